"""Run-to-run bit reproducibility of the batch-256 forward, with single-image forwards in between (what
tests/test_gpu_forward.py::test_bench_size_properties does).  Prints which outputs / layers / images differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import interactive_vit_b200.engine as E
from interactive_vit_b200 import vit_plugin as P

cfg = E.CONFIGS["vit_b_16"]
eng = E.VitEngine(cfg, 0, 256)
eng.load_state_dict(P.build_torchvision_vit(cfg, seed=0).state_dict())
x = torch.rand(256, 3, 224, 224, generator=torch.Generator().manual_seed(1234))
flags = E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT
small = int(os.environ.get("SMALL", "1"))
big = eng.forward_host(x, flags)
bad = 0
for it in range(int(os.environ.get("ITERS", "12"))):
    if small:
        for i in (0, 127, 255, 255):
            eng.forward_host(x[i:i + 1].contiguous(), flags)
    again = eng.forward_host(x, flags)
    for k in ("logits", "cls_maps", "avg_maps", "rollout"):
        if not torch.equal(again[k], big[k]):
            bad += 1
            d = (again[k] - big[k]).abs()
            bdim = 1 if k in ("cls_maps", "avg_maps") else 0
            imgs = d.movedim(bdim, 0).flatten(1).max(1).values.nonzero().flatten().tolist()
            lay = d.flatten(1).max(1).values.nonzero().flatten().tolist() if bdim == 1 else None
            print(f"iter {it}: {k} differs, max {d.max().item():.3e}, images {imgs[:12]}{'...' if len(imgs) > 12 else ''} ({len(imgs)}), layers {lay}")
print("SPLIT=%s SMALL=%d: %d differing outputs" % (os.environ.get("VITB200_ATTN_SPLIT", "1"), small, bad))
