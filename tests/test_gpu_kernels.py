"""Per-kernel parity on the B200: every CUDA kernel is called through the C ABI (libvitb200.so) and compared
with a plain torch fp32 evaluation of the same op on the same seeded inputs.

Tolerances: the GEMM / attention operands are bf16 by design, so outputs that are *stored* as bf16 carry one
bf16 rounding (2^-9 relative); fp32 outputs of the GEMM differ from the fp32 reference only by summation order.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_EPS = 2.0 ** -8     # one bf16 rounding of an O(ref_max) value, with margin
F32_EPS = 2e-5


def _rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-30)).item()


@pytest.fixture(scope="module")
def E(built_library):
    import interactive_vit_b200.engine as E

    assert torch.cuda.is_available()
    return E


def test_patchify_is_exact(E):
    x = torch.rand(3, 3, 224, 224, device="cuda")
    got = E.op_patchify(x, 16)
    ref = torch.nn.functional.unfold(x, kernel_size=16, stride=16).transpose(1, 2).reshape(3 * 196, 768).bfloat16()
    assert torch.equal(got, ref)
    x = torch.rand(2, 3, 64, 64, device="cuda")
    ref = torch.nn.functional.unfold(x, kernel_size=8, stride=8).transpose(1, 2).reshape(2 * 64, 192).bfloat16()
    assert torch.equal(E.op_patchify(x, 8), ref)


@pytest.mark.parametrize("d", [128, 256, 384, 768, 1024, 1280])
def test_layernorm(E, d):
    torch.manual_seed(d)
    x = torch.randn(1001, d, device="cuda") * 3 + 0.7
    g, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    got = E.op_layernorm(x, g, b, 1e-6)
    ref = torch.nn.functional.layer_norm(x, (d,), g, b, 1e-6)
    assert _rel(got, ref) < BF16_EPS


GEMM_CASES = [
    # M, N, K, bias, resid, gelu, out_f32
    (128, 128, 64, False, False, False, True),
    (128, 256, 64, True, False, False, True),
    (1, 128, 64, True, False, False, True),          # single row, M tail
    (129, 256, 128, True, True, False, True),        # one row into the second tile
    (788, 768, 768, True, True, False, True),        # out_proj + residual
    (788, 2304, 768, True, False, False, False),     # qkv
    (788, 3072, 768, True, False, True, False),      # fc1 + GELU
    (788, 768, 3072, True, True, False, True),       # fc2 + residual
    (7, 1000, 768, True, False, False, True),        # classifier head, N tail (1000 = 7*128 + 104)
    (591, 1152, 384, True, False, False, False),     # ViT-S qkv (BN=128 path)
    (591, 1536, 384, True, False, True, False),      # ViT-S fc1
    (300, 4096, 1024, True, False, True, False),     # ViT-L fc1
]


@pytest.mark.parametrize("M,N,K,bias,resid,gelu,out_f32", GEMM_CASES)
def test_gemm(E, M, N, K, bias, resid, gelu, out_f32):
    torch.manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bs = torch.randn(N, device="cuda") if bias else None
    rs = torch.randn(M, N, device="cuda") if resid else None
    got = E.op_gemm(a, w, bs, rs, gelu, out_f32)
    ref = a.float() @ w.float().t()
    if bias:
        ref = ref + bs
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + rs
    assert got.shape == (M, N) and torch.isfinite(got.float()).all()
    assert _rel(got, ref) < (F32_EPS if out_f32 else BF16_EPS)


@pytest.mark.parametrize("M,N,K,resid,gelu,out_f32", [(788, 768, 768, True, False, True), (788, 2304, 768, False, False, False),
                                                       (788, 3072, 768, False, True, False), (300, 768, 3072, True, False, True),
                                                       (7, 1000, 768, False, False, True), (591, 1152, 384, False, False, False)])
def test_gemm_split_bf16_operands(E, M, N, K, resid, gelu, out_f32):
    """fp32x3 precision mode: fp32 operands as hi + lo bf16, product = hi*hi + lo*hi + hi*lo in three K passes, fp32
    accumulation in TMEM.  Against an fp64 matmul: 5e-5 of the output range (a plain bf16 GEMM sits at 2e-3); the
    non-fp32 output is itself a hi + lo pair."""
    torch.manual_seed(M + N)
    a = torch.randn(M, K, device="cuda") * 0.5
    w = torch.randn(N, K, device="cuda") * 0.05
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if resid else None
    got = E.op_gemm_split(a, w, b, r, gelu, out_f32)
    ref = a.double() @ w.double().t() + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if resid:
        ref = ref + r.double()
    assert got.shape == (M, N) and torch.isfinite(got).all()
    assert ((got.double() - ref).abs().max() / ref.abs().max()).item() < 5e-5


def test_split_bf16_is_a_17_bit_representation(E):
    x = torch.randn(1 << 16, device="cuda") * torch.logspace(-3, 3, 1 << 16, device="cuda")
    hi, lo = E.op_split_bf16(x)
    assert torch.equal(hi, x.bfloat16())
    assert ((hi.float() + lo.float() - x).abs() / x.abs().clamp_min(1e-30)).max().item() < 2.0 ** -16


def test_gemm_full_size_linearity(E):
    """Size-independent property at the bench shape (M = 256 * 197): GEMM(a1 + a2) == GEMM(a1) + GEMM(a2) when
    the sum a1 + a2 is exact in bf16, and every row depends only on its own input row."""
    M, N, K = 256 * 197, 768, 768
    torch.manual_seed(0)
    a1 = (torch.randint(-8, 9, (M, K), device="cuda").float() / 8).bfloat16()
    a2 = (torch.randint(-8, 9, (M, K), device="cuda").float() / 8).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    y1, y2 = E.op_gemm(a1, w, out_f32=True), E.op_gemm(a2, w, out_f32=True)
    y12 = E.op_gemm((a1.float() + a2.float()).bfloat16(), w, out_f32=True)
    assert _rel(y12, y1 + y2) < F32_EPS
    rows = torch.tensor([0, 127, 128, 50431], device="cuda")
    assert torch.equal(E.op_gemm(a1[rows].contiguous(), w, out_f32=True), y1[rows])   # bit-exact row independence


@pytest.mark.parametrize("M,d,n_out,gelu", [(788, 768, 2304, False), (788, 768, 3072, True), (300, 384, 1152, False),
                                             (257, 1280, 5120, True), (50432, 768, 2304, False)])
def test_layernorm_folded_into_gemms(E, M, d, n_out, gelu):
    """The residual GEMM's producer epilogue (bf16 copy + per-chunk partial sums) and the consumer epilogue
    (rstd * (acc - mean * colsum) + bias') together reproduce Linear(LayerNorm(x)) [-> GELU] of the reference
    (EncoderBlock, vision_transformer.py:110-119) for a residual stream with a non-zero mean and uneven scales."""
    torch.manual_seed(M + d)
    a = (torch.randn(M, d, device="cuda") * 0.5).bfloat16()
    w0 = (torch.randn(d, d, device="cuda") * 0.05).bfloat16()
    b0 = torch.randn(d, device="cuda")
    resid = torch.randn(M, d, device="cuda") * torch.linspace(0.5, 3.0, d, device="cuda") + 0.7
    x, xb, stats = E.op_gemm_residual_stats(a, w0, b0, resid)
    x_ref = a.float() @ w0.float().t() + b0 + resid
    assert _rel(x, x_ref) < F32_EPS
    assert torch.equal(xb, x.bfloat16())                                     # the copy is the rounded fp32 result
    sw = E.ln_slot_width(d)       # one slot per epilogue-warp column group (128 columns, 64 for widths not divisible by 256)
    assert stats.shape == (M, d // sw, 2)
    chunks = x.reshape(M, d // sw, sw)
    assert _rel(stats[..., 0], chunks.sum(-1)) < 1e-5 and _rel(stats[..., 1], (chunks * chunks).sum(-1)) < 1e-5
    # consumer
    gamma = 1.0 + 0.1 * torch.randn(d, device="cuda")
    beta = 0.05 * torch.randn(d, device="cuda")
    w1 = torch.randn(n_out, d, device="cuda") * 0.05
    b1 = torch.randn(n_out, device="cuda") * 0.02
    wq, colsum, bias_f = E.op_fold_ln(w1, gamma, beta, b1)
    assert torch.equal(wq, (w1 * gamma).bfloat16())
    assert _rel(colsum, wq.float().sum(-1)) < 1e-5 and _rel(bias_f, b1 + w1 @ beta) < 1e-5
    got = E.op_gemm_ln(xb, stats, wq, colsum, bias_f, gelu)
    ref = torch.nn.functional.layer_norm(x_ref, (d,), gamma, beta, 1e-6) @ w1.t() + b1
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert _rel(got, ref) < 2 * BF16_EPS      # bf16 operands (x and gamma * W) and a bf16 result
    # deterministic: fixed statistics slots, no atomics
    x2, xb2, stats2 = E.op_gemm_residual_stats(a, w0, b0, resid)
    assert torch.equal(stats, stats2) and torch.equal(E.op_gemm_ln(xb2, stats2, wq, colsum, bias_f, gelu), got)


@pytest.mark.parametrize("B,H,W,resize,crop", [(2, 300, 400, 256, 224), (1, 512, 512, 256, 224), (1, 224, 224, 256, 224),
                                               (3, 100, 80, 256, 224), (1, 683, 1024, 256, 224), (1, 480, 640, 384, 384),
                                               (1, 341, 256, 256, 224)])
def test_preprocess_matches_torchvision_preset(E, B, H, W, resize, crop):
    """`<model>:transform`: antialiased bilinear resize + centre crop + normalise == torchvision's ImageClassification
    preset on the CPU (what the reference's VggModel runs for its transform node, static/models/vgg16.py:40-42)."""
    from oracle import vit_oracle as O

    torch.manual_seed(H * W)
    x = torch.rand(B, 3, H, W)
    ref = O.preprocess(x, crop, resize)
    got = E.op_preprocess(x.cuda(), resize, crop).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 2e-5


def _attn_ref(qkv, B, N, H, D=64):
    q, k, v = qkv.float().reshape(B, N, 3, H, D).permute(2, 0, 3, 1, 4)
    p = torch.softmax((q * D ** -0.5) @ k.transpose(-1, -2), dim=-1)
    o = (p @ v).permute(0, 2, 1, 3).reshape(B * N, H * D)
    return o, p


@pytest.mark.parametrize("B,N,H,scale", [(2, 197, 12, 1.0), (3, 197, 6, 3.0), (2, 64, 2, 1.0), (1, 17, 2, 2.0),
                                         (1, 128, 1, 1.0), (2, 129, 3, 1.0), (5, 197, 16, 0.5),
                                         # more work items than SMs with a short last round: its items are split over
                                         # two CTAs by heads and the head average is combined with a TMA reduce-add
                                         (100, 197, 4, 1.0), (90, 197, 3, 1.0), (180, 100, 2, 1.0)])
def test_attention(E, B, N, H, scale):
    torch.manual_seed(B * 1000 + N)
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * scale).bfloat16()
    ctx, avg, cls, hm = E.op_attention(qkv, B, N, H, True, True, True)
    o, p = _attn_ref(qkv, B, N, H)
    assert _rel(ctx, o) < 2 * BF16_EPS           # P is rounded to bf16 for the P.V product, the output again
    assert _rel(hm, p) < 1e-5                    # probabilities are emitted in fp32
    # the head average is summed on the tensor pipe (fp32 accumulate) from the SAME bf16 probabilities that feed P.V
    assert _rel(avg, p.mean(1)) < BF16_EPS
    assert _rel(cls, p[:, :, 0, :]) < 1e-5
    assert (hm.sum(-1) - 1).abs().max() < 1e-5   # rows of a softmax
    # outputs selected independently give the same context (same kernel: per-head maps select the one-head-in-flight
    # kernel, so keep the two-heads-in-flight kernel of the 197-token shape out of this comparison)
    os.environ["VITB200_ATTN_PP"] = "0"
    try:
        ctx2, _, _, _ = E.op_attention(qkv, B, N, H, False, False, False)
    finally:
        del os.environ["VITB200_ATTN_PP"]
    assert torch.equal(ctx, ctx2)


@pytest.mark.parametrize("B,N,H,scale", [(2, 197, 12, 1.0), (3, 197, 6, 3.0), (5, 197, 16, 0.5), (1, 197, 1, 1.0),
                                         (1, 193, 2, 1.0), (2, 200, 5, 2.0), (2, 197, 3, 8.0),
                                         # short last round: items split over two CTAs by heads (1 + 2 heads for H = 3)
                                         (100, 197, 4, 1.0), (90, 197, 3, 1.0), (256, 197, 12, 1.0)])
def test_attention_two_heads_in_flight(E, B, N, H, scale):
    """attention_pp.cuh (193..200 tokens, no per-head maps): two groups of softmax warps on alternate heads, scores
    carried as fp16 differences to the thread-local row maximum.  Against the fp32 reference and against the
    one-head-in-flight kernel on the same input."""
    torch.manual_seed(B * 1000 + N + H)
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * scale).bfloat16()
    ctx, avg, cls, _ = E.op_attention(qkv, B, N, H, True, True, False)
    o, p = _attn_ref(qkv, B, N, H)
    assert _rel(ctx, o) < 2 * BF16_EPS
    assert _rel(avg, p.mean(1)) < BF16_EPS
    # class-token rows: fp16 exponentials (relative rounding 4.9e-4, plus the fp16 rounding of the difference to the
    # row maximum) times the fp32 normalising factor
    assert _rel(cls, p[:, :, 0, :]) < 2e-3
    assert (avg.sum(-1) - 1).abs().max() < 2e-3   # fp16 probabilities: rows of the head average sum to 1 within 1e-3
    os.environ["VITB200_ATTN_PP"] = "0"
    try:
        ctx1, avg1, cls1, _ = E.op_attention(qkv, B, N, H, True, True, False)
    finally:
        del os.environ["VITB200_ATTN_PP"]
    assert _rel(ctx, ctx1) < 2 * BF16_EPS and _rel(avg, avg1) < 2e-3 and _rel(cls, cls1) < 2e-3
    # outputs selected independently give the same context, and a second run the same bits
    ctx2, _, _, _ = E.op_attention(qkv, B, N, H, False, False, False)
    assert torch.equal(ctx, ctx2)
    ctx3, avg3, cls3, _ = E.op_attention(qkv, B, N, H, True, True, False)
    assert torch.equal(ctx, ctx3) and torch.equal(avg, avg3) and torch.equal(cls, cls3)


@pytest.mark.parametrize("B,N,H,heads", [(1, 197, 12, False), (1, 197, 12, True), (2, 197, 6, False), (5, 197, 16, False),
                                         (8, 197, 12, False), (3, 100, 5, False), (24, 197, 12, False), (1, 197, 3, False)])
def test_attention_head_split_of_small_launches(E, B, N, H, heads):
    """Launches with at most a third of the SMs' worth of (image, query tile) items split every item over up to H CTAs
    by heads (engine.cu launch_attention); the parts of the head average meet in index order (avg_parts_sum_kernel).
    Context, class-token rows and per-head maps must not depend on the split at all, the head average only through
    the order of its fp32 sum; every variant is bit-reproducible."""
    torch.manual_seed(B * 100 + H)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    outs = {}
    for mode in ("0", "2", "1"):
        os.environ["VITB200_ATTN_SPLIT"] = mode
        try:
            outs[mode] = E.op_attention(qkv, B, N, H, True, True, heads)
            again = E.op_attention(qkv, B, N, H, True, True, heads)
        finally:
            del os.environ["VITB200_ATTN_SPLIT"]
        for a, b in zip(outs[mode], again):
            assert (a is None and b is None) or torch.equal(a, b)
    ctx0, avg0, cls0, hm0 = outs["0"]
    _, p = _attn_ref(qkv, B, N, H)
    for mode in ("2", "1"):
        ctx, avg, cls, hm = outs[mode]
        assert torch.equal(ctx, ctx0) and torch.equal(cls, cls0)
        assert (hm is None and hm0 is None) or torch.equal(hm, hm0)
        assert (avg - avg0).abs().max() < 1e-6
        assert _rel(avg, p.mean(1)) < BF16_EPS


@pytest.mark.parametrize("B,N,H,D,scale", [(2, 577, 12, 64, 1.0), (2, 577, 16, 80, 1.0), (1, 257, 4, 80, 2.0),
                                           (3, 197, 6, 80, 1.0), (1, 300, 2, 96, 1.0), (1, 129, 2, 128, 1.0),
                                           (2, 209, 3, 64, 3.0), (1, 768, 1, 64, 1.0)])
def test_attention_long(E, B, N, H, D, scale):
    """Key-blocked two-kernel path: more than 208 tokens and / or head dims other than 64 (ViT-H: 80).  Every map is
    formed from fp32 probabilities here (the head average accumulates in registers)."""
    torch.manual_seed(B * 1000 + N + D)
    qkv = (torch.randn(B * N, 3 * H * D, device="cuda") * scale).bfloat16()
    ctx, avg, cls, hm = E.op_attention(qkv, B, N, H, True, True, True, head_dim=D)
    o, p = _attn_ref(qkv, B, N, H, D)
    assert _rel(ctx, o) < 2 * BF16_EPS
    assert _rel(hm, p) < 1e-5
    assert _rel(avg, p.mean(1)) < 1e-5
    assert _rel(cls, p[:, :, 0, :]) < 1e-5
    assert (hm.sum(-1) - 1).abs().max() < 1e-5
    ctx2, _, _, _ = E.op_attention(qkv, B, N, H, False, False, False, head_dim=D)
    assert torch.equal(ctx, ctx2)


@pytest.mark.parametrize("B,N,H,D", [(2, 577, 3, 64), (1, 577, 2, 80), (1, 700, 2, 64), (2, 257, 2, 128)])
def test_attention_long_online_rescale(E, B, N, H, D):
    """The context kernel of the key-blocked path makes ONE pass over the keys: every thread (query row, 64-key half of a
    128-key block) keeps its own reference maximum and raises it -- rescaling its row of its O accumulator in TMEM --
    only when a later block exceeds it by more than 2^8 (attention_long.cuh).  Keys whose scores grow block by block
    force that path in every block; keys that shrink never take it; both against the fp32 reference and against the
    two-pass kernel (VITB200_ATTN_LONG_ONLINE=0)."""
    torch.manual_seed(N + D)
    d = H * D
    for growth in (1.35, 1 / 1.35, 1.0):
        qkv = torch.randn(B, N, 3, H, D, device="cuda")
        ramp = growth ** (torch.arange(N, device="cuda") // 64).float()           # per 64-key half block
        qkv[:, :, 1] *= ramp[None, :, None, None] * 2.0
        qkv[:, :, 0] *= 2.0
        qkv = qkv.reshape(B * N, 3 * d).bfloat16()
        ctx, avg, cls, hm = E.op_attention(qkv, B, N, H, True, True, True, head_dim=D)
        o, p = _attn_ref(qkv, B, N, H, D)
        assert torch.isfinite(ctx.float()).all()
        assert _rel(ctx, o) < 2 * BF16_EPS
        # (scores reach several hundred here: the fp32 rounding of s * c alone is ~1e-5 relative in exp2; the two-pass
        # kernel measures the same on these inputs)
        assert _rel(hm, p) < 2e-5 and _rel(avg, p.mean(1)) < 2e-5 and _rel(cls, p[:, :, 0, :]) < 2e-5
        assert (hm.sum(-1) - 1).abs().max() < 2e-5
        os.environ["VITB200_ATTN_LONG_ONLINE"] = "0"
        try:
            ctx2, avg2, _, _ = E.op_attention(qkv, B, N, H, True, False, False, head_dim=D)
        finally:
            del os.environ["VITB200_ATTN_LONG_ONLINE"]
        assert _rel(ctx, ctx2) < 2 * BF16_EPS and _rel(avg, avg2) < 2e-5
        assert _rel(avg2, p.mean(1)) < 2e-5
        ctx3, _, _, _ = E.op_attention(qkv, B, N, H, False, False, False, head_dim=D)
        assert torch.equal(ctx, ctx3)            # bit-reproducible, independent of the outputs selected


def test_attention_masks_padded_keys(E):
    """Keys beyond N come from the next image (or TMA zero fill) and must get probability 0: changing image 1 must
    not change image 0's outputs."""
    B, N, H = 2, 197, 4
    torch.manual_seed(1)
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    ctx_a, avg_a, _, _ = E.op_attention(qkv, B, N, H)
    qkv2 = qkv.clone()
    qkv2[N:] = (torch.randn(N, 3 * H * 64, device="cuda") * 50).bfloat16()
    ctx_b, avg_b, _, _ = E.op_attention(qkv2, B, N, H)
    assert torch.equal(ctx_a[:N], ctx_b[:N]) and torch.equal(avg_a[0], avg_b[0])


# Batch sizes on either side of the cluster switch (engine.cu launch_rollout: clusters of 8 / 4 / 2 CTAs per image up to
# 74 images on 148 SMs, one CTA per image beyond), both token counts, pad columns poisoned.
@pytest.mark.parametrize("L,B,N,pitch", [(5, 3, 197, 208), (12, 1, 197, 208), (4, 1, 577, 592), (3, 20, 197, 208),
                                         (3, 40, 50, 64), (2, 80, 197, 208), (1, 2, 197, 208), (6, 2, 17, 32)])
def test_rollout(E, L, B, N, pitch):
    from oracle.vit_oracle import rollout_from_avg

    torch.manual_seed(2)
    p = torch.softmax(torch.randn(L, B, N, N) * 2, dim=-1)
    padded = torch.full((L, B, N, pitch), float("nan"))
    padded[..., :N] = p
    dev = padded.cuda()
    got = E.op_rollout(dev).cpu()
    ref = rollout_from_avg(list(p))
    assert torch.isfinite(got).all()
    assert _rel(got, ref) < 1e-5
    assert (got.sum(-1) - ref.sum(-1)).abs().max() < 1e-5
    assert torch.equal(E.op_rollout(dev).cpu(), got)      # fixed summation order: bit-reproducible
