"""Phase timing of the attention kernel from in-kernel clock64() stamps (tracing build, CTA 0 only).
    nvcc ... -DVITB200_ATTN_TRACE -o gpurun_out/libvitb200_trace.so engine.cu ; VITB200_LIB=... python tools/attn_trace.py"""
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
for maps in ((True, True), (False, False)):
    for _ in range(3):
        E.op_attention(qkv, B, N, H, maps[0], maps[1], False)
    buf = (C.c_longlong * (64 * 32))()
    lib.vitb200_debug_attn_trace.argtypes = [C.c_void_p, C.c_int]
    assert lib.vitb200_debug_attn_trace(buf, 64 * 32) == 0
    t = [[buf[h * 32 + k] for k in range(32)] for h in range(H)]
    names = {1: "s_full passed", 2: "S ld done", 4: "mask+max done", 5: "exp done", 6: "bar passed", 9: "inv computed",
             10: "p_free passed", 11: "O staged (h-1)", 12: "P stored", 7: "fence+arrive", 13: "O store issued", 14: "cls_free passed",
             8: "cls staged / end"}
    print(f"maps={maps}: softmax warp 4 lane 0, cycles since loop top of the head (mean over heads 2..{H - 2})")
    for k in (1, 2, 4, 5, 6, 9, 10, 11, 12, 7, 13, 14, 8):
        d = [t[h][k] - t[h][0] for h in range(2, H - 1)]
        # warp 9 (quarter 1, column group 1): its stamps 4..15 live in slots 20..31, same time base (warp 4's loop top)
        d9 = [t[h][16 + k] - t[h][0] for h in range(2, H - 1)] if 4 <= k < 16 else None
        print(f"   {names[k]:16s} {sum(d) / len(d):8.0f}" + (f"   warp 9: {sum(d9) / len(d9):8.0f}" if d9 else ""))
    per_head = [t[h + 1][0] - t[h][0] for h in range(2, H - 2)]
    print(f"   head period    {sum(per_head) / len(per_head):8.0f}")
    print("  MMA thread: p_full wait %.0f, o_free wait %.0f, issue PV(+avg) %.0f, period %.0f" % (
        sum(t[h][17] - t[h][16] for h in range(2, H - 1)) / (H - 3), sum(t[h][18] - t[h][17] for h in range(2, H - 1)) / (H - 3),
        sum(t[h][19] - t[h][18] for h in range(2, H - 1)) / (H - 3), sum(t[h + 1][16] - t[h][16] for h in range(2, H - 2)) / (H - 4)))
    print("  softmax P stored -> MMA saw p_full: %.0f" % (sum(t[h][17] - t[h][7] for h in range(2, H - 1)) / (H - 3)))
    print("  whole CTA 0 (first stamp of head 0 -> last stamp): %.0f cycles" % (max(t[H - 1][8], t[H - 1][19]) - t[0][0]))
