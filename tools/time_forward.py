"""Time the batch forward of the library in VITB200_LIB (default: the in-tree one): ms per step over BLOCKS blocks of
STEPS forwards (CUDA events on the launch stream), and the per-kernel breakdown of profiled forwards.  One line of JSON.
Used by tools/ab.sh to compare two builds on ONE box in alternation (box-to-box spread is +-5 %)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import interactive_vit_b200.engine as E
from interactive_vit_b200 import vit_plugin as P

model = os.environ.get("MODEL", "vit_b_16")
B = int(os.environ.get("BATCH", "256"))
steps, blocks = int(os.environ.get("STEPS", "100")), int(os.environ.get("BLOCKS", "3"))
warm = int(os.environ.get("WARM", "100"))   # reach the power-capped steady state before timing (short runs drift by +-4 %)
flags = int(os.environ.get("FLAGS", str(E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT)))
cfg = E.CONFIGS[model]
eng = E.VitEngine(cfg, 0, B)
eng.load_state_dict(P.build_torchvision_vit(cfg, seed=0).state_dict())
x = torch.rand(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1234)).cuda()
st = torch.cuda.Stream()
for _ in range(warm):
    eng.forward_device(x, flags, st.cuda_stream)
st.synchronize()
ms = []
for _ in range(blocks):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(steps):
        eng.forward_device(x, flags, st.cuda_stream)
    b.record(st)
    st.synchronize()
    ms.append(a.elapsed_time(b) / steps)
runs = [eng.profile_forward(x, flags) for _ in range(3)]
kern = {k: round(sorted(r[k][1] for r in runs)[1], 3) for k in runs[0]}
print(json.dumps({"lib": os.path.basename(E.LIB_PATH), "model": model, "batch": B, "ms_per_step": [round(v, 3) for v in ms],
                  "img_per_s": round(B / min(ms) * 1e3), "kernels_ms": kern}))
eng.close()
