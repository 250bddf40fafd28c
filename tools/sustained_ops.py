"""Energy view of the step (round 2): the forward is POWER-bound in steady state (sw_power_cap, ~965 W), so what limits
throughput is energy per image, not idle gaps.  This tool runs each hot kernel of the ViT-B/16 batch-256 layer back to
back for SECONDS each (so the power governor settles), and prints its sustained rate, SM clock and power -- next to
cuBLAS (torch.matmul, a MEASUREMENT REFERENCE only, never on the product path) on the same shapes.

    python tools/sustained_ops.py        # one JSON line per op
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import interactive_vit_b200.engine as E

SECONDS = float(os.environ.get("SECONDS", "1.5"))
B, N, H, d, mlp = 256, 197, 12, 768, 3072
M = B * N
dev = "cuda"
torch.manual_seed(0)


class Sampler:
    def __init__(self):
        import pynvml

        pynvml.nvmlInit()
        self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        self.clk, self.pw, self.stop = [], [], threading.Event()

    def __enter__(self):
        def run():
            while not self.stop.is_set():
                self.clk.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.pw.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                self.stop.wait(0.05)

        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join()

    def med(self, xs):
        xs = sorted(xs[len(xs) // 3:])     # the settled part
        return xs[len(xs) // 2] if xs else None


def sustained(name, fn, flop=None, bytes_=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    n = max(10, int(SECONDS * 1e3 / max(a.elapsed_time(b), 1e-3)))
    with Sampler() as s:
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
    us = a.elapsed_time(b) / n * 1e3
    out = {"op": name, "us": round(us, 2), "sm_mhz": s.med(s.clk), "power_w": round(s.med(s.pw) or 0, 1)}
    if flop:
        out["tflops"] = round(flop / us / 1e6, 1)
        out["tflop_per_joule"] = round(flop / us / 1e6 / max(out["power_w"], 1), 3)
    if bytes_:
        out["gbs"] = round(bytes_ / us / 1e3, 1)
    print(json.dumps(out), flush=True)
    time.sleep(1.0)


def bf(*shape, scale=0.05):
    return (torch.randn(*shape, device=dev) * scale).bfloat16()


x_b, w_qkv, w_o, w_fc1, w_fc2 = bf(M, d, scale=1.0), bf(3 * d, d), bf(d, d), bf(mlp, d), bf(d, mlp)
h_b = bf(M, mlp, scale=1.0)
bias3, bias1, biasd = torch.randn(3 * d, device=dev), torch.randn(mlp, device=dev), torch.randn(d, device=dev)
resid = torch.randn(M, d, device=dev)
shapes = {"qkv": (3 * d, d), "out_proj": (d, d), "fc1": (mlp, d), "fc2": (d, mlp)}

# cuBLAS reference on the same shapes (bf16 in, bf16 out, no epilogue)
for name, (n_, k_) in shapes.items():
    a_ = h_b if k_ == mlp else x_b
    w_ = {"qkv": w_qkv, "out_proj": w_o, "fc1": w_fc1, "fc2": w_fc2}[name]
    out = torch.empty(M, n_, device=dev, dtype=torch.bfloat16)
    sustained(f"cublas_{name}", lambda: torch.matmul(a_, w_.t(), out=out), flop=2.0 * M * n_ * k_)
big = bf(8192, 8192)
outb = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
sustained("cublas_8192^3", lambda: torch.matmul(big, big.t(), out=outb), flop=2.0 * 8192 ** 3)

# ours, same epilogues as in the forward
sustained("ours_qkv(+bias)", lambda: E.op_gemm(x_b, w_qkv, bias3), flop=2.0 * M * 3 * d * d)
sustained("ours_fc1(+bias+gelu)", lambda: E.op_gemm(x_b, w_fc1, bias1, gelu=True), flop=2.0 * M * mlp * d)
sustained("ours_fc2(+bias+resid f32)", lambda: E.op_gemm(h_b, w_fc2, biasd, resid, out_f32=True), flop=2.0 * M * d * mlp)
sustained("ours_out_proj(+bias+resid f32)", lambda: E.op_gemm(x_b, w_o, biasd, resid, out_f32=True), flop=2.0 * M * d * d)
sustained("ours_out_proj(+resid+xb+stats)", lambda: E.op_gemm_residual_stats(x_b, w_o, biasd, resid), flop=2.0 * M * d * d)
qkv = bf(M, 3 * d, scale=1.0)
qkv[:, 2 * d:] = qkv[:, 2 * d:].float().half().view(torch.bfloat16)   # the V third in fp16, as the qkv GEMM writes it
lib = E.load_library()
ctx = torch.empty(M, d, device=dev, dtype=torch.bfloat16)
avg = torch.empty(B, N, 208, device=dev)
cls = torch.empty(B, H, N, device=dev)


def attn(a, c):
    E.check(lib.vitb200_op_attention_ex(qkv.data_ptr(), ctx.data_ptr(), a.data_ptr() if a is not None else None,
                                        c.data_ptr() if c is not None else None, None, B, N, H, 64, 208, 1, None))


sustained("ours_attention(avg+cls)", lambda: attn(avg, cls), flop=4.0 * B * H * N * N * 64,
          bytes_=M * 3 * d * 2 + M * d * 2 + B * N * 208 * 4)
sustained("ours_attention(ctx only)", lambda: attn(None, None), flop=4.0 * B * H * N * N * 64)
maps = torch.rand(12, B, N, 208, device=dev)
sustained("ours_rollout", lambda: E.op_rollout(maps), bytes_=maps.numel() * 4)
