"""CPU oracle for the ViT forward + attention-map path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product (interactive-vit_b200/) never does.

What it restates
----------------
The reference (0Marble/interactive-vit) ships no ViT: its hot path is ``Model.compute`` calling
``sub(x)`` under ``torch.no_grad()`` on CPU fp32 (main/context.py:79-88), reached from
``Context.compute`` (main/context.py:143-147).  The arithmetic therefore lives in an un-vendored third-party
dependency, **torchvision** (reference pin: torchvision 0.24.1 / torch 2.9.1, requirements.txt:78,89,90;
this image: torchvision 0.26.0 / torch 2.11.0).  The oracle *calls torchvision itself* on CPU fp32 — no
re-implementation of the math — and only re-states the glue torchvision does not expose:

* ``EncoderBlock.forward`` (torchvision/models/vision_transformer.py:110-119) with ``need_weights=True,
  average_attn_weights=False`` so that ``nn.MultiheadAttention`` returns the per-head probabilities
  ``softmax((q/sqrt(D)) k^T)`` (torch/nn/functional.py:6630-6659).  Outputs are bit-identical to the default
  ``need_weights=False`` forward on CPU (checked in tests/test_oracle.py).
* ``VisionTransformer.forward`` (vision_transformer.py:289-306) split at the block-granular node boundaries
  the plugin exposes: embed / layer.i / head.
* attention rollout (Abnar & Zuidema 2020), which neither the reference nor torchvision has (north_star
  feature):  A_l = rownorm(0.5 * mean_h P_l + 0.5 I),  R = A_L ... A_1,  map = R[0, 1:].

Parity pinning
--------------
The reference has no tests, golden vectors or fixtures for any path (main/tests.py:1-3 is empty), so the
oracle is pinned against outputs of the reference itself run in the build container: tests/golden/*.pt are
produced by oracle/make_golden.py, which drives the *unmodified* reference modules main/graph.py,
main/context.py and main/message.py (imported from /root/reference behind a 6-line django.conf stub) with this
oracle's model as the hosted ``Model`` plugin, and tests/test_oracle.py re-checks them on every run.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
from torchvision.models.vision_transformer import VisionTransformer


@dataclass(frozen=True)
class OracleConfig:
    image_size: int = 224
    patch_size: int = 16
    num_layers: int = 12
    num_heads: int = 12
    hidden_dim: int = 768
    mlp_dim: int = 3072
    num_classes: int = 1000

    @property
    def tokens(self) -> int:
        return (self.image_size // self.patch_size) ** 2 + 1


ORACLE_CONFIGS: Dict[str, OracleConfig] = {
    "vit_s_16": OracleConfig(224, 16, 12, 6, 384, 1536),
    "vit_b_16": OracleConfig(224, 16, 12, 12, 768, 3072),
    "vit_l_16": OracleConfig(224, 16, 24, 16, 1024, 4096),
    # 384 px / patch 16 -> 577 tokens (BASELINE.json config 5; SURVEY.md section 8 on the "H/14 384px (577 tokens)" wording):
    # torchvision ships this geometry as ViT-B/16 SWAG_E2E (vision_transformer.py:374-392); "vit_h_16_384" is ViT-H's
    # width / depth / heads (1280 / 32 / 16, head dim 80) on the same geometry
    "vit_b_16_384": OracleConfig(384, 16, 12, 12, 768, 3072),
    "vit_h_16_384": OracleConfig(384, 16, 32, 16, 1280, 5120),
    # small shapes for fast CPU tests (same code path, same 197-token geometry or smaller)
    "vit_tiny_test": OracleConfig(64, 16, 2, 2, 128, 256, 16),
    "vit_small_test": OracleConfig(224, 16, 3, 4, 256, 512, 40),
    # ViT-H's layer shape (head dim 80, 577 tokens) at 2 layers: the key-blocked attention path end to end in seconds
    "vit_h_test": OracleConfig(384, 16, 2, 16, 1280, 5120, 40),
    "vit_577_test": OracleConfig(384, 16, 2, 4, 256, 512, 40),
}


def build_vit(cfg: OracleConfig, seed: int = 0, init: str = "default") -> VisionTransformer:
    """Random-init torchvision ViT (no weights can be downloaded here).

    init="default": torchvision's own initialisation under ``torch.manual_seed(seed)``, then the classifier
        head re-drawn from N(0, 0.02) because torchvision zero-inits it (vision_transformer.py:264-266), which
        would make every logit 0 and top-1 degenerate.
    init="stress": additionally perturbs everything torchvision initialises to a constant (biases, LayerNorm
        affine, class token) and sharpens the attention logits, so that a kernel that drops a bias, a gamma or
        the softmax scale cannot pass.
    """
    torch.manual_seed(seed)
    m = VisionTransformer(
        image_size=cfg.image_size,
        patch_size=cfg.patch_size,
        num_layers=cfg.num_layers,
        num_heads=cfg.num_heads,
        hidden_dim=cfg.hidden_dim,
        mlp_dim=cfg.mlp_dim,
        num_classes=cfg.num_classes,
    )
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        m.heads.head.weight.normal_(0.0, 0.02, generator=g)
        m.heads.head.bias.normal_(0.0, 0.02, generator=g)
        if init == "stress":
            d = cfg.hidden_dim
            m.class_token.normal_(0.0, 0.02, generator=g)
            m.conv_proj.bias.normal_(0.0, 0.02, generator=g)
            for blk in m.encoder.layers:
                for ln in (blk.ln_1, blk.ln_2):
                    ln.weight.copy_(1.0 + 0.1 * torch.randn(d, generator=g))
                    ln.bias.normal_(0.0, 0.05, generator=g)
                att = blk.self_attention
                att.in_proj_weight[: 2 * d].mul_(2.0)  # peakier softmax (4x larger attention logits)
                att.in_proj_bias.normal_(0.0, 0.02, generator=g)
                att.out_proj.bias.normal_(0.0, 0.02, generator=g)
                blk.mlp[0].bias.normal_(0.0, 0.02, generator=g)
                blk.mlp[3].bias.normal_(0.0, 0.02, generator=g)
            m.encoder.ln.weight.copy_(1.0 + 0.1 * torch.randn(d, generator=g))
            m.encoder.ln.bias.normal_(0.0, 0.05, generator=g)
        elif init != "default":
            raise ValueError(init)
    return m.eval()


def synthetic_images(batch: int, image_size: int, seed: int = 1234) -> torch.Tensor:
    """Uniform [0,1) fp32 images (SURVEY.md §8d synthetic inputs)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 3, image_size, image_size, generator=g)


def default_resize(image_size: int, params=None) -> int:
    """Resize target of the `transform` node: the node's `resize` param, else torchvision's preset for the geometry
    (ViT_B_16_Weights.IMAGENET1K_V1: resize 256 / crop 224, vision_transformer.py:354; SWAG 384: resize = crop)."""
    if params is not None and str(params.get("resize", "")).strip():
        return int(params["resize"])
    return 256 if image_size == 224 else image_size


def preprocess(x: torch.Tensor, crop: int, resize: int) -> torch.Tensor:
    """torchvision's ImageClassification preset (transforms/_presets.py:58-65) on [3,H,W] or [B,3,H,W] in [0,1]."""
    from torchvision.transforms._presets import ImageClassification

    return ImageClassification(crop_size=crop, resize_size=resize)(x)


# ---- node-granular stages (each follows the torchvision lines cited) -----------------------------------
@torch.no_grad()
def embed(model: VisionTransformer, images: torch.Tensor) -> torch.Tensor:
    """vision_transformer.py:291-296 (_process_input + class token) and the pos-embedding add of Encoder.forward
    (vision_transformer.py:155-156; dropout is identity in eval).  [B,3,S,S] -> [B,N,d]."""
    x = model._process_input(images)
    cls = model.class_token.expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1)
    return x + model.encoder.pos_embedding


@torch.no_grad()
def encoder_layer(model: VisionTransformer, i: int, x: torch.Tensor):
    """EncoderBlock.forward (vision_transformer.py:110-119) with the attention weights kept.
    Returns (tokens [B,N,d], per-head probabilities [B,H,N,N])."""
    blk = model.encoder.layers[i]
    h = blk.ln_1(x)
    a, p = blk.self_attention(h, h, h, need_weights=True, average_attn_weights=False)
    a = blk.dropout(a)
    a = a + x
    y = blk.mlp(blk.ln_2(a))
    return a + y, p


@torch.no_grad()
def encoder_attn_half(model: VisionTransformer, i: int, x: torch.Tensor):
    """First half of EncoderBlock.forward (vision_transformer.py:112-116): x + dropout(MHA(ln_1 x)), weights kept.
    Returns (tokens [B,N,d], per-head probabilities [B,H,N,N])."""
    blk = model.encoder.layers[i]
    h = blk.ln_1(x)
    a, p = blk.self_attention(h, h, h, need_weights=True, average_attn_weights=False)
    return blk.dropout(a) + x, p


@torch.no_grad()
def encoder_mlp_half(model: VisionTransformer, i: int, x: torch.Tensor) -> torch.Tensor:
    """Second half (vision_transformer.py:118-119): x + mlp(ln_2 x)."""
    blk = model.encoder.layers[i]
    return x + blk.mlp(blk.ln_2(x))


@torch.no_grad()
def head(model: VisionTransformer, x: torch.Tensor) -> torch.Tensor:
    """Final LayerNorm (vision_transformer.py:157), class token (302), classifier (304).  [B,N,d] -> [B,classes]."""
    return model.heads(model.encoder.ln(x)[:, 0])


def rollout_from_avg(avg_maps: List[torch.Tensor]) -> torch.Tensor:
    """Attention rollout, class-token row.  avg_maps: L tensors [B,N,N] (layer 0 first) -> [B,N-1]."""
    R: Optional[torch.Tensor] = None
    for a in avg_maps:
        N = a.shape[-1]
        ah = 0.5 * a + 0.5 * torch.eye(N, dtype=a.dtype)
        ah = ah / ah.sum(dim=-1, keepdim=True)
        R = ah if R is None else ah @ R
    assert R is not None
    return R[:, 0, 1:]


@torch.no_grad()
def forward_with_maps(model: VisionTransformer, images: torch.Tensor) -> Dict[str, torch.Tensor]:
    """Whole path.  Returns logits [B,C], hidden [L,B,N,d], heads [L,B,H,N,N], avg_maps [L,B,N,N],
    cls_maps [L,B,H,N], rollout [B,N-1], embed [B,N,d]."""
    x = embed(model, images)
    out = {"embed": x}
    hidden, heads = [], []
    for i in range(len(model.encoder.layers)):
        x, p = encoder_layer(model, i, x)
        hidden.append(x)
        heads.append(p)
    out["hidden"] = torch.stack(hidden)
    out["heads"] = torch.stack(heads)
    out["avg_maps"] = out["heads"].mean(dim=2)
    out["cls_maps"] = out["heads"][:, :, :, 0, :].contiguous()
    out["rollout"] = rollout_from_avg(list(out["avg_maps"]))
    out["logits"] = head(model, x)
    return out
