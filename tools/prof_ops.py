"""Profiling driver: a handful of launches of each hot kernel at the bench shapes (ViT-B/16, batch 256) so that
`ncu -k regex:<name>` can capture them.   python tools/prof_ops.py [attention|gemm|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from interactive_vit_b200 import engine as E

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B, N, H = 256, 197, 12
M = B * N
if what in ("attention", "all"):
    qkv = torch.randn(M, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3):
        E.op_attention(qkv, B, N, H, True, True, False)
    torch.cuda.synchronize()
if what in ("gemm", "all"):
    for (Nn, K, gelu, f32, resid) in ((2304, 768, False, False, False), (768, 768, False, True, True),
                                      (3072, 768, True, False, False), (768, 3072, False, True, True)):
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(Nn, K, device="cuda") * 0.05).bfloat16()
        bs = torch.randn(Nn, device="cuda")
        rs = torch.randn(M, Nn, device="cuda") if resid else None
        for _ in range(2):
            E.op_gemm(a, w, bs, rs, gelu, f32)
        torch.cuda.synchronize()
print("ok")
