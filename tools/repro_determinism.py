"""Run-to-run bit reproducibility of the batch-256 forward (BASELINE config 2), compared ON THE DEVICE so that
thousands of iterations fit in seconds.  For every differing output it prints which layers / images / rows /
columns differ and by how much -- enough to tell a lost head-average contribution (whole 16-column groups of a
query tile, ~p/12) from a reordered sum (last bit).

    ITERS=2000 SMALL=0 python tools/repro_determinism.py            # the library under test
    VITB200_LIB=ab_old/libvitb200_old.so ITERS=2000 python tools/repro_determinism.py   # a library built from another commit

Round 2 finding: the round-1 library loses the last head's averaging MMAs now and then (the accumulator was read
behind o_full instead of p_free); see profiles/README.md."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import interactive_vit_b200.engine as E
from interactive_vit_b200 import vit_plugin as P

model = os.environ.get("MODEL", "vit_b_16")
B = int(os.environ.get("BATCH", "256"))
cfg = E.CONFIGS[model]
eng = E.VitEngine(cfg, 0, B)
eng.load_state_dict(P.build_torchvision_vit(cfg, seed=0).state_dict())
x = torch.rand(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1234)).cuda()
x1 = [x[i:i + 1].contiguous() for i in (0, B // 2, B - 1)]
flags = E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT
small = int(os.environ.get("SMALL", "1"))
iters = int(os.environ.get("ITERS", "200"))
L, H, N = cfg.num_layers, cfg.num_heads, cfg.tokens
pitch = (N + 15) // 16 * 16
stream = torch.cuda.Stream()


def run(imgs):
    eng.forward_device(imgs, flags, stream.cuda_stream)
    stream.synchronize()


def outputs():
    return {"logits": eng.device_output(0, (B, cfg.num_classes)).clone(),
            "cls_maps": eng.device_output(E.EMIT_CLS, (L, B, H, N)).clone(),
            "avg_maps": eng.device_output(E.EMIT_AVG, (L, B, N, pitch))[..., :N].clone(),
            "rollout": eng.device_output(E.EMIT_ROLLOUT, (B, N - 1)).clone()}


run(x)
ref = outputs()
bad = 0
for it in range(iters):
    if small:
        for xi in x1:
            run(xi)
    run(x)
    got = outputs()
    for k in ("logits", "cls_maps", "avg_maps", "rollout"):
        if torch.equal(got[k], ref[k]):
            continue
        bad += 1
        d = (got[k] - ref[k]).abs()
        msg = f"iter {it}: {k} differs in {int((d > 0).sum())} elements, max |diff| {d.max().item():.3e}"
        if k == "avg_maps":
            idx = (d > 0).nonzero()
            lay = sorted(set(idx[:, 0].tolist()))
            img = sorted(set(idx[:, 1].tolist()))
            rows = sorted(set(idx[:, 2].tolist()))
            cols = sorted(set(idx[:, 3].tolist()))
            rel = (d / ref[k].abs().clamp_min(1e-30))[d > 0]
            msg += (f"; layers {lay[:8]} images {img[:8]} ({len(img)}) rows {rows[0]}..{rows[-1]} ({len(rows)}) "
                    f"cols {cols[0]}..{cols[-1]} ({len(cols)}); relative diff median {rel.median().item():.3e} "
                    f"(a lost head contributes ~1/{H} = {1 / H:.3f})")
        print(msg, flush=True)
print(f"lib={os.path.basename(E.LIB_PATH)} model={model} batch={B} iters={iters} SMALL={small} "
      f"SPLIT={os.environ.get('VITB200_ATTN_SPLIT', '1')}: {bad} differing outputs")
eng.close()
