"""Developer harness: run every CUDA kernel (and the whole forward) once against a torch fp32 reference on the
GPU box and print error figures.  Each case runs in its own subprocess with a timeout, so a trap / illegal
access / hang in one kernel does not hide the results of the others.  Not part of the product or the tests.

    gpurun -- python tools/gpu_check.py [case ...]
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _err(name, got, ref):
    import torch
    got, ref = got.float(), ref.float()
    diff = (got - ref).abs()
    rel = diff.max().item() / max(ref.abs().max().item(), 1e-30)
    bad = (~torch.isfinite(got)).sum().item()
    print(f"    {name}: max_abs={diff.max().item():.4e} rel_to_max={rel:.4e} mean_abs={diff.mean().item():.4e} "
          f"ref_max={ref.abs().max().item():.3e} nonfinite={bad}", flush=True)
    return rel


@case
def patchify():
    import torch
    from interactive_vit_b200 import engine as E
    x = torch.rand(3, 3, 224, 224, device="cuda")
    got = E.op_patchify(x, 16)
    ref = torch.nn.functional.unfold(x, kernel_size=16, stride=16).transpose(1, 2).reshape(3 * 196, 768)
    _err("patchify", got, ref.bfloat16())


@case
def layernorm():
    import torch
    from interactive_vit_b200 import engine as E
    for d in (128, 384, 768, 1024):
        x = torch.randn(1000, d, device="cuda") * 2 + 0.5
        g = torch.randn(d, device="cuda")
        b = torch.randn(d, device="cuda")
        got = E.op_layernorm(x, g, b)
        ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-6)
        _err(f"layernorm d={d}", got, ref)


def _gemm_case(M, N, K, bias, resid, gelu, out_f32):
    import torch
    from interactive_vit_b200 import engine as E
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bs = torch.randn(N, device="cuda") if bias else None
    rs = torch.randn(M, N, device="cuda") if resid else None
    got = E.op_gemm(a, w, bs, rs, gelu, out_f32)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + bs
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + rs
    return _err(f"gemm M={M} N={N} K={K} bias={bias} resid={resid} gelu={gelu} f32={out_f32}", got, ref)


@case
def gemm_small():
    _gemm_case(128, 128, 64, False, False, False, True)
    _gemm_case(128, 256, 64, False, False, False, True)
    _gemm_case(128, 128, 128, False, False, False, True)
    _gemm_case(256, 256, 256, True, False, False, True)


@case
def gemm_shapes():
    _gemm_case(197 * 4, 768, 768, True, True, False, True)
    _gemm_case(197 * 4, 2304, 768, True, False, False, False)
    _gemm_case(197 * 4, 3072, 768, True, False, True, False)
    _gemm_case(197 * 4, 768, 3072, True, True, False, True)
    _gemm_case(7, 1000, 768, True, False, False, True)
    _gemm_case(197 * 3, 1152, 384, True, False, False, False)
    _gemm_case(197 * 300, 768, 768, True, True, False, True)


@case
def gemm_cluster():
    """4-CTA multicast clusters forced onto small shapes: share-W units, the share-A last row (even and odd n_tiles),
    M tails, every epilogue."""
    os.environ["VITB200_GEMM_CLUSTER_MIN_TILES"] = "0"
    os.environ["VITB200_GEMM_PAIR"] = "4"
    _gemm_case(512, 256, 64, False, False, False, True)
    _gemm_case(512, 256, 768, True, False, False, True)
    _gemm_case(768, 512, 256, True, False, False, True)      # odd m_tiles: one share-W row + one share-A unit
    _gemm_case(768, 768, 256, True, False, False, True)      # share-A row with odd n_tiles (second pair idle)
    _gemm_case(256, 768, 128, True, False, False, True)      # share-A only
    _gemm_case(700, 2304, 768, True, False, True, False)     # M tail, GELU, bf16 out
    _gemm_case(197 * 9, 768, 3072, True, True, False, True)  # residual epilogue
    _gemm_case(197 * 64, 2304, 768, True, False, False, False)
    _gemm_case(197 * 256, 768, 768, True, True, False, True)


@case
def gemm_split():
    """fp32x3 GEMM (split-bf16 operands) against an fp64 reference."""
    import torch
    from interactive_vit_b200 import engine as E
    for (M, N, K, gelu, f32, resid) in ((788, 768, 768, False, True, True), (788, 2304, 768, False, False, False),
                                        (788, 3072, 768, True, False, False), (300, 768, 3072, False, True, True)):
        torch.manual_seed(M + N)
        a = torch.randn(M, K, device="cuda") * 0.5
        w = torch.randn(N, K, device="cuda") * 0.05
        b = torch.randn(N, device="cuda")
        r = torch.randn(M, N, device="cuda") if resid else None
        got = E.op_gemm_split(a, w, b, r, gelu, f32)
        ref = a.double() @ w.double().t() + b.double()
        if gelu:
            ref = torch.nn.functional.gelu(ref)
        if resid:
            ref = ref + r.double()
        _err(f"gemm_split M={M} N={N} K={K} gelu={gelu} f32={f32} resid={resid}", got.double(), ref)


@case
def gemm_ln_perf():
    """Plain bias epilogue vs the folded-LayerNorm consumer epilogue, and plain residual vs the producer epilogue."""
    import torch
    from interactive_vit_b200 import engine as E
    M, d = 256 * 197, 768

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n

    a = (torch.randn(M, d, device="cuda") * 0.5).bfloat16()
    x = torch.randn(M, d, device="cuda")
    stats = torch.stack([x.reshape(M, d // 32, 32).sum(-1), (x * x).reshape(M, d // 32, 32).sum(-1)], -1).contiguous()
    for (N, gelu) in ((2304, False), (3072, True)):
        w = (torch.randn(N, d, device="cuda") * 0.05).bfloat16()
        b = torch.randn(N, device="cuda")
        cs = w.float().sum(-1).contiguous()
        t_plain = timeit(lambda: E.op_gemm(a, w, b, None, gelu, False))
        t_ln = timeit(lambda: E.op_gemm_ln(a, stats, w, cs, b, gelu))
        print(f"    N={N} gelu={gelu}: plain {t_plain * 1e3:.1f} us   folded-LN {t_ln * 1e3:.1f} us", flush=True)
    w = (torch.randn(d, d, device="cuda") * 0.05).bfloat16()
    b = torch.randn(d, device="cuda")
    t_plain = timeit(lambda: E.op_gemm(a, w, b, x, False, True))
    t_st = timeit(lambda: E.op_gemm_residual_stats(a, w, b, x))
    print(f"    out_proj: plain residual {t_plain * 1e3:.1f} us   + bf16 copy + stats {t_st * 1e3:.1f} us (incl. torch allocs)", flush=True)


@case
def gemm_perf_cluster():
    os.environ["VITB200_GEMM_PAIR"] = "4"
    gemm_perf()


@case
def gemm_perf():
    import torch
    from interactive_vit_b200 import engine as E
    M = 256 * 197
    for (M, N, K, gelu, f32, resid) in ((M, 2304, 768, False, False, False), (M, 768, 768, False, True, True),
                                        (M, 3072, 768, True, False, False), (M, 768, 3072, False, True, True),
                                        (74 * 256, 256, 16384, False, False, False), (74 * 256 * 4, 512, 4096, False, False, False)):
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        bs = torch.randn(N, device="cuda")
        rs = torch.randn(M, N, device="cuda") if resid else None
        for _ in range(3):
            E.op_gemm(a, w, bs, rs, gelu, f32)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            E.op_gemm(a, w, bs, rs, gelu, f32)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"    gemm M={M} N={N} K={K} gelu={gelu}: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
        t0.record()
        for _ in range(10):
            c = a @ w.t()
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"    cublas same shape: {ms:.3f} ms  {2 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)


def _attn_ref(qkv, B, N, H, D=64):
    import torch
    d = H * D
    q, k, v = qkv.float().reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
    s = (q * D ** -0.5) @ k.transpose(-1, -2)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, d)
    return o, p


@case
def attention():
    import torch
    from interactive_vit_b200 import engine as E
    for (B, N, H, scale) in ((2, 197, 12, 1.0), (3, 197, 6, 3.0), (2, 64, 2, 1.0), (1, 17, 2, 2.0)):
        torch.manual_seed(B * N)
        qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * scale).bfloat16()
        ctx, avg, cls, hm = E.op_attention(qkv, B, N, H, True, True, True)
        torch.cuda.synchronize()
        o, p = _attn_ref(qkv, B, N, H)
        print(f"  attention B={B} N={N} H={H} scale={scale}")
        _err("ctx", ctx, o)
        _err("avg", avg, p.mean(1))
        _err("cls", cls, p[:, :, 0, :])
        _err("heads", hm, p)


@case
def attention_long():
    """Key-blocked path: 577 tokens (384 px / patch 16), head dims 64 and 80, odd sizes."""
    import torch
    from interactive_vit_b200 import engine as E
    for (B, N, H, D, scale) in ((2, 577, 12, 64, 1.0), (2, 577, 16, 80, 1.0), (1, 257, 4, 80, 2.0), (3, 197, 6, 80, 1.0),
                                (1, 300, 2, 96, 1.0), (1, 129, 2, 128, 1.0), (2, 209, 3, 64, 3.0)):
        torch.manual_seed(B * N)
        qkv = (torch.randn(B * N, 3 * H * D, device="cuda") * scale).bfloat16()
        ctx, avg, cls, hm = E.op_attention(qkv, B, N, H, True, True, True, head_dim=D)
        torch.cuda.synchronize()
        o, p = _attn_ref(qkv, B, N, H, D)
        print(f"  attention_long B={B} N={N} H={H} D={D} scale={scale}")
        _err("ctx", ctx, o)
        _err("avg", avg, p.mean(1))
        _err("cls", cls, p[:, :, 0, :])
        _err("heads", hm, p)
        ctx2, _, _, _ = E.op_attention(qkv, B, N, H, False, False, False, head_dim=D)
        print("    ctx identical without maps:", bool(torch.equal(ctx, ctx2)), flush=True)


@case
def attention_perf():
    import torch
    from interactive_vit_b200 import engine as E
    B, N, H = 256, 197, 12
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    for flags in ((True, True, False), (True, False, False), (False, True, False), (False, False, False)):
        for _ in range(3):
            E.op_attention(qkv, B, N, H, *flags)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            E.op_attention(qkv, B, N, H, *flags)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 10
        print(f"    attention B={B} maps={flags}: {ms:.3f} ms  {4 * B * H * N * N * 64 / ms / 1e9:.1f} TFLOP/s (incl. torch allocs)", flush=True)


@case
def rollout():
    import torch
    from interactive_vit_b200 import engine as E
    from oracle.vit_oracle import rollout_from_avg
    L, B, N, pitch = 5, 3, 197, 208
    p = torch.softmax(torch.randn(L, B, N, N) * 2, dim=-1)
    padded = torch.zeros(L, B, N, pitch)
    padded[..., :N] = p
    got = E.op_rollout(padded.cuda())
    ref = rollout_from_avg(list(p))
    _err("rollout", got.cpu(), ref)


def _forward_case(name, batch, init):
    import torch
    from interactive_vit_b200 import engine as E
    from oracle import vit_oracle as O
    ocfg = O.ORACLE_CONFIGS[name]
    model = O.build_vit(ocfg, seed=0, init=init)
    x = O.synthetic_images(batch, ocfg.image_size)
    ref = O.forward_with_maps(model, x)
    cfg = E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                      ocfg.num_classes)
    eng = E.VitEngine(cfg, 0, batch)
    eng.load_state_dict(model.state_dict())
    got = eng.forward_host(x, E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT | E.EMIT_HEADS | E.EMIT_HIDDEN)
    print(f"  forward {name} batch={batch} init={init}")
    for k in ("logits", "avg_maps", "cls_maps", "rollout", "heads", "hidden"):
        _err(k, got[k], ref[k])
    _err("hidden[0]", got["hidden"][0], ref["hidden"][0])
    print("    top1 equal:", bool((got["logits"].argmax(-1) == ref["logits"].argmax(-1)).all()), flush=True)
    # stage path
    eng.stage_embed(x)
    _err("stage embed", eng.get_tokens(batch), ref["embed"])


@case
def forward_precise():
    """fp32x3 precision mode (split-bf16 operands) against the fp32 CPU oracle: north_star asks <= 1e-3."""
    import torch
    from interactive_vit_b200 import engine as E
    from oracle import vit_oracle as O
    for (name, batch, init) in (("vit_small_test", 3, "stress"), ("vit_577_test", 2, "stress"), ("vit_b_16", 2, "default"),
                                ("vit_b_16", 1, "stress")):
        ocfg = O.ORACLE_CONFIGS[name]
        model = O.build_vit(ocfg, seed=0, init=init)
        x = O.synthetic_images(batch, ocfg.image_size)
        ref = O.forward_with_maps(model, x)
        cfg = E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                          ocfg.num_classes)
        eng = E.VitEngine(cfg, 0, batch, precision="fp32x3")
        eng.load_state_dict(model.state_dict())
        got = eng.forward_host(x, E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT | E.EMIT_HEADS | E.EMIT_HIDDEN)
        print(f"  forward_precise {name} batch={batch} init={init}")
        for k in ("logits", "avg_maps", "cls_maps", "rollout", "heads", "hidden"):
            _err(k, got[k], ref[k])
        print("    top1 equal:", bool((got["logits"].argmax(-1) == ref["logits"].argmax(-1)).all()), flush=True)
        eng.close()


@case
def forward_small():
    _forward_case("vit_small_test", 3, "stress")


@case
def forward_vitb():
    _forward_case("vit_b_16", 2, "default")
    _forward_case("vit_b_16", 2, "stress")


@case
def forward_perf():
    import torch
    from interactive_vit_b200 import engine as E
    from oracle import vit_oracle as O
    ocfg = O.ORACLE_CONFIGS["vit_b_16"]
    model = O.build_vit(ocfg, seed=0)
    cfg = E.CONFIGS["vit_b_16"]
    B = 256
    eng = E.VitEngine(cfg, 0, B)
    eng.load_state_dict(model.state_dict())
    x = torch.rand(B, 3, 224, 224, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for flags in (0, E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT):
        for _ in range(3):
            eng.forward_device(x, flags, st)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            eng.forward_device(x, flags, st)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        ips = B / ms * 1e3
        print(f"    forward vit_b_16 B={B} flags={flags}: {ms:.3f} ms  {ips:.0f} img/s  "
              f"{ips * cfg.gflop_per_image() / 1e3:.1f} TFLOP/s", flush=True)


@case
def forward_profile():
    """Per-kernel event times inside one real forward (vit_b_16, B=256, all maps): where the step goes."""
    import torch
    from interactive_vit_b200 import engine as E
    from oracle import vit_oracle as O
    name = os.environ.get("VITB200_PROFILE_MODEL", "vit_b_16")
    B = int(os.environ.get("VITB200_PROFILE_BATCH", "256"))
    ocfg = O.ORACLE_CONFIGS[name]
    model = O.build_vit(ocfg, seed=0)
    cfg = E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                      ocfg.num_classes)
    eng = E.VitEngine(cfg, 0, B)
    eng.load_state_dict(model.state_dict())
    x = torch.rand(B, 3, ocfg.image_size, ocfg.image_size, device="cuda")
    flags = E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT
    for _ in range(3):
        eng.forward_device(x, flags)
    eng.synchronize()
    runs = [eng.profile_forward(x, flags) for _ in range(5)]
    keys = list(runs[0])
    print(f"    {name} B={B}: median of 5 profiled forwards (event-to-event ms, gaps included)")
    for k in keys:
        ms = sorted(r[k][1] for r in runs)[2]
        print(f"    {k:20s} x{runs[0][k][0]:3d}  {ms:8.3f} ms  {1e3 * ms / max(runs[0][k][0], 1):8.1f} us each", flush=True)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream().cuda_stream
    eng.forward_device(x, flags, st)
    torch.cuda.synchronize()
    t0.record()
    for _ in range(5):
        eng.forward_device(x, flags, st)
    t1.record()
    torch.cuda.synchronize()
    print(f"    unprofiled forward: {t0.elapsed_time(t1) / 5:.3f} ms", flush=True)


@case
def forward_graph():
    """Eager launches vs one CUDA-graph replay of the same forward: what the launch gaps cost."""
    import torch
    from interactive_vit_b200 import engine as E
    from oracle import vit_oracle as O
    ocfg = O.ORACLE_CONFIGS["vit_b_16"]
    model = O.build_vit(ocfg, seed=0)
    cfg = E.CONFIGS["vit_b_16"]
    B = 256
    eng = E.VitEngine(cfg, 0, B)
    eng.load_state_dict(model.state_dict())
    x = torch.rand(B, 3, 224, 224, device="cuda")
    flags = E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            eng.forward_device(x, flags, s.cuda_stream)
        s.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(s)
        for _ in range(10):
            eng.forward_device(x, flags, s.cuda_stream)
        t1.record(s)
        s.synchronize()
        print(f"    eager: {t0.elapsed_time(t1) / 10:.3f} ms", flush=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            eng.forward_device(x, flags, s.cuda_stream)
        for _ in range(3):
            g.replay()
        s.synchronize()
        t0.record(s)
        for _ in range(10):
            g.replay()
        t1.record(s)
        s.synchronize()
        print(f"    graph: {t0.elapsed_time(t1) / 10:.3f} ms", flush=True)


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--run":
        CASES[sys.argv[2]]()
        return
    names = sys.argv[1:] or list(CASES)
    for n in names:
        t = time.time()
        print(f"=== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", n], timeout=300, cwd=ROOT)
            print(f"=== {n}: exit {r.returncode} in {time.time() - t:.1f}s", flush=True)
        except subprocess.TimeoutExpired:
            print(f"=== {n}: TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
