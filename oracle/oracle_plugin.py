"""The oracle as a plugin: torchvision's ViT hosted behind the reference's ``Model`` API on CPU fp32.

TEST INFRASTRUCTURE ONLY (see oracle/vit_oracle.py).  This is "the reference's PyTorch/CPU ViT path" that
BASELINE.json's north_star names: the reference ships no ViT, so — exactly like ``VggModel`` hosts VGG16
(static/models/vgg16.py:10-62) — a ``Model`` subclass hosts torchvision's VisionTransformer with block-granular
nodes.  Node names, ``io()`` and output shapes are the contract the product plugin
(interactive-vit_b200/vit_plugin.py) is checked against; the arithmetic is torchvision's own.

``make_oracle_model_class(Model, Pinout)`` binds to either the reference's unmodified main.context.Model
(oracle/refhost.py, only available in the build container) or the product's mirror (tests on the GPU box).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import vit_oracle as O


def make_oracle_model_class(ModelBase, PinoutCls):
    class VitOracleModel(ModelBase):
        def __init__(self, name: str, cfg: O.OracleConfig, module: Optional[torch.nn.Module] = None):
            module = module if module is not None else O.build_vit(cfg)
            super().__init__(module, name)
            self.cfg = cfg
            self.node_names = ([self.prefix() + "embed"] + [self.prefix() + f"layer.{i}" for i in range(cfg.num_layers)]
                               + [self.prefix() + "head", self.prefix() + "rollout", self.prefix() + "transform"])

        def list_node_names(self) -> List[str]:
            return self.node_names

        def fine_node_names(self) -> List[str]:
            return [self.prefix() + f"layer.{i}.{half}" for i in range(self.cfg.num_layers) for half in ("attn", "mlp")]

        def io(self, node_name: str, params=None) -> Dict:
            sub = node_name.removeprefix(self.prefix())
            if sub.startswith("layer.") and sub.endswith(".mlp"):
                return {"ins": ["o"], "outs": ["o"]}
            if sub.startswith("layer."):
                return {"ins": ["o"], "outs": ["o", "attn", "cls"]}
            if sub == "rollout":
                return {"ins": [f"a{i}" for i in range(self.cfg.num_layers)], "outs": ["o"]}
            return {"ins": ["o"], "outs": ["o"]}

        def contents(self, node_name: str) -> str:
            return f"<p>{node_name}</p> <p>torchvision CPU oracle</p>"

        def compute(self, node_name: str, pinin, params=None):
            sub = node_name.removeprefix(self.prefix())
            c = self.cfg
            g = c.image_size // c.patch_size
            out = PinoutCls()
            with torch.no_grad():
                if sub == "embed":
                    x = pinin.get("o")
                    assert x is not None
                    batched = x.dim() == 4
                    t = O.embed(self.model, x if batched else x[None])
                    out.set("o", t if batched else t[0])
                elif sub.startswith("layer.") and sub.endswith(".mlp"):
                    i = int(sub[len("layer."):-len(".mlp")])
                    x = pinin.get("o")
                    assert x is not None
                    batched = x.dim() == 3
                    t = O.encoder_mlp_half(self.model, i, x if batched else x[None])
                    out.set("o", t if batched else t[0])
                elif sub.startswith("layer."):
                    half = sub.endswith(".attn")
                    i = int(sub[len("layer."):-len(".attn")] if half else sub[len("layer."):])
                    x = pinin.get("o")
                    assert x is not None
                    batched = x.dim() == 3
                    fn = O.encoder_attn_half if half else O.encoder_layer
                    t, p = fn(self.model, i, x if batched else x[None])
                    avg = p.mean(dim=1)
                    cls = p[:, :, 0, 1:].reshape(p.shape[0], c.num_heads, g, g).contiguous()
                    out.set("o", t if batched else t[0])
                    out.set("attn", avg if batched else avg[0])
                    out.set("cls", cls if batched else cls[0])
                elif sub == "head":
                    x = pinin.get("o")
                    assert x is not None
                    batched = x.dim() == 3
                    y = O.head(self.model, x if batched else x[None])
                    out.set("o", y if batched else y[0])
                elif sub == "rollout":
                    maps = [pinin.get(f"a{i}") for i in range(c.num_layers)]
                    assert all(m is not None for m in maps)
                    batched = maps[0].dim() == 3
                    r = O.rollout_from_avg([m if batched else m[None] for m in maps])
                    r = r.reshape(r.shape[0], g, g)
                    out.set("o", r if batched else r[0])
                elif sub == "transform":
                    # what VggModel does for its `transform` pseudo-node (static/models/vgg16.py:40-42): the weights'
                    # torchvision preset on the CPU
                    x = pinin.get("o")
                    assert x is not None
                    out.set("o", O.preprocess(x, c.image_size, O.default_resize(c.image_size, params)))
                else:
                    raise KeyError(node_name)
            return out

    return VitOracleModel


def vit_graph_request(name: str, num_layers: int, image: torch.Tensor):
    """(nodes, edges, tensors) of the request a browser would POST for the full ViT graph: image -> embed ->
    layers -> head, attn_i -> rollout (the wire JSON of main/message.py:61-73)."""
    nodes = [{"endpoint": f"{name}:embed", "params": {}}]
    nodes += [{"endpoint": f"{name}:layer.{i}", "params": {}} for i in range(num_layers)]
    nodes += [{"endpoint": f"{name}:head", "params": {}}, {"endpoint": f"{name}:rollout", "params": {}}]
    head_idx, rollout_idx = 1 + num_layers, 2 + num_layers
    edges = [{"tensor": 0, "out_port": {"node": 0, "channel": "o"}}]
    for i in range(1, head_idx + 1):
        edges.append({"in_port": {"node": i - 1, "channel": "o"}, "out_port": {"node": i, "channel": "o"}})
    for i in range(num_layers):
        edges.append({"in_port": {"node": 1 + i, "channel": "attn"}, "out_port": {"node": rollout_idx, "channel": f"a{i}"}})
    return nodes, edges, [image]
