"""Phase timing of the two-heads-in-flight attention kernel (attention_pp.cuh) from in-kernel clock64() stamps
(tracing build, CTA 0 only):
    nvcc <flags of __graft_entry__> -DVITB200_ATTN_TRACE -o gpurun_out/libvitb200_trace.so engine.cu
    VITB200_LIB=gpurun_out/libvitb200_trace.so python tools/attn_trace_pp.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from interactive_vit_b200 import engine as E

lib = E.load_library()
B, N, H = 256, 197, 12
qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
names = ["loop top", "s_full passed", "pass 1a (max) done", "pass 1b (fp16 d) done, S released", "pass 2 (exp) done",
         "exchange barrier passed", "factor computed", "p_free passed", "P stored", "fence + p_full arrive"]
for maps in ((True, True), (False, False)):
    for _ in range(3):
        E.op_attention(qkv, B, N, H, maps[0], maps[1], False)
    buf = (C.c_longlong * (64 * 32))()
    lib.vitb200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
    assert lib.vitb200_debug_attn_trace(buf, 64 * 32) == 0
    t = [[buf[h * 32 + k] for k in range(32)] for h in range(H)]
    t0 = t[0][0]
    print(f"maps={maps}: absolute timeline of CTA 0 (cycles since head 0's loop top); softmax stamps from quarter 0 / column half 0 of the head's group")
    print("  head | " + " ".join(f"{k:>6d}" for k in range(10)) + " |   QK iss  PV iss  PV end  ep beg  ep end")
    for h in range(H):
        print(f"  {h:4d} | " + " ".join(f"{t[h][k] - t0:6d}" for k in range(10)) + " | " +
              " ".join(f"{t[h][k] - t0:7d}" for k in range(16, 21)))
    hs = range(4, H - 2)
    print("  phase durations, mean over heads 4..%d:" % (H - 3))
    for k in range(1, 10):
        print(f"    {names[k]:36s} {sum(t[h][k] - t[h][k - 1] for h in hs) / len(hs):7.0f}")
    print("    group period (2 heads)               %7.0f" % (sum(t[h + 2][0] - t[h][0] for h in hs) / len(hs)))
    print("    head period                          %7.0f" % ((t[H - 1][9] - t[2][9]) / (H - 3)))
    print("    S ready (s_full) after QK issue      %7.0f" % (sum(t[h][1] - t[h][16] for h in hs) / len(hs)))
    print("    QK(h+1) issue after S(h) released    %7.0f" % (sum(t[h + 1][16] - t[h][3] for h in hs) / len(hs)))
    print("    PV issue after P stored+arrive       %7.0f" % (sum(t[h][17] - t[h][9] for h in hs) / len(hs)))
    print("    whole CTA 0: %d cycles" % (max(max(r[:10] + r[16:21]) for r in t) - t0))
