"""The ViT ``Model`` plugin: block-granular graph nodes whose ``compute`` runs on the B200 engine.

It plays the role ``static/models/vgg16.py:10-62`` plays for VGG16 in the reference — a ``Model`` subclass that
overrides ``list_node_names / compute / io / contents / generate_graph_json`` — because the generic leaf-module
wrapper cannot express a ViT (it drops nn.MultiheadAttention, the class token, the position
embedding and both residual adds, and Graph.connect cannot fan a channel out server-side).

Nodes (``<name>`` is the model name, e.g. ``vit_b_16``):

    <name>:embed      ins [o]            outs [o]              [3,S,S] or [B,3,S,S] -> tokens [N,d] / [B,N,d]
    <name>:layer.<i>  ins [o]            outs [o, attn, cls]   tokens -> tokens, head-averaged map [N,N],
                                          (+ heads if params["heads"] == "1")   per-head CLS maps [H,g,g], [H,N,N]
    <name>:layer.<i>.attn  ins [o]       outs [o, attn, cls]   first half of the block: x + out_proj(MHA(LN1 x)) and its maps
    <name>:layer.<i>.mlp   ins [o]       outs [o]              second half: x + MLP(LN2 x)   (finer-grained graphs, SURVEY 8f-4;
                                                               registered, but not part of the default graph file)
    <name>:head       ins [o]            outs [o]              tokens -> logits [classes]
    <name>:rollout    ins [a0..a{L-1}]   outs [o]              head-averaged maps -> rollout map [g,g]

Every output is a CPU fp32 tensor (the wire format, main/message.py:111-121).  Between consecutive nodes of
one request the token stream stays on the GPU: the plugin remembers which tensor object it handed out last and
skips the upload when that same object comes back as the next node's input.

The class is produced by ``make_vit_model_class(Model, Pinout)`` so that the same code subclasses either this
package's mirror of the plugin API or the reference's own ``main.context.Model``.
"""
from __future__ import annotations

import math
import os
import threading
from typing import Dict, List, Optional

import torch

from . import engine as E


def build_torchvision_vit(cfg: E.VitConfig, seed: int = 0) -> torch.nn.Module:
    """Random-init torchvision ViT (weights cannot be downloaded offline).  The classifier is re-drawn from
    N(0, 0.02): torchvision zero-inits it (vision_transformer.py:264-266), which would make every logit 0."""
    from torchvision.models.vision_transformer import VisionTransformer

    torch.manual_seed(seed)
    m = VisionTransformer(image_size=cfg.image_size, patch_size=cfg.patch_size, num_layers=cfg.num_layers,
                          num_heads=cfg.num_heads, hidden_dim=cfg.hidden_dim, mlp_dim=cfg.mlp_dim,
                          num_classes=cfg.num_classes)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        m.heads.head.weight.normal_(0.0, 0.02, generator=g)
        m.heads.head.bias.normal_(0.0, 0.02, generator=g)
    return m.eval()


def imagenet_categories(num_classes: int) -> List[str]:
    if num_classes == 1000:
        try:
            from torchvision.models import ViT_B_16_Weights

            return list(ViT_B_16_Weights.IMAGENET1K_V1.meta["categories"])
        except Exception:
            pass
    return [f"class {i}" for i in range(num_classes)]


def vit_graph_request(name: str, num_layers: int, image: torch.Tensor):
    """(nodes, edges, tensors) of the request a browser POSTs for the full ViT graph (the wire JSON of
    main/message.py:61-73): image -> embed -> layer.0.. -> head, layer.i `attn` -> rollout."""
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


def make_vit_model_class(ModelBase, PinoutCls):
    class VitB200Model(ModelBase):
        """One ViT replica on one GPU behind the reference's ``Model`` plugin interface."""

        def __init__(self, name: str, cfg: E.VitConfig, module: Optional[torch.nn.Module] = None, device: int = 0,
                     max_batch: int = 1, engine: Optional[E.VitEngine] = None):
            module = module if module is not None else build_torchvision_vit(cfg)
            super().__init__(module, name)  # .eval(), self.name (main/context.py:39-42)
            self.cfg = cfg
            self.engine = engine if engine is not None else E.VitEngine(cfg, device, max_batch)
            self.engine.load_state_dict(module.state_dict())
            # Deferred host outputs (default on; VITB200_DEFERRED=0 restores a wait per call): node outputs are pinned CPU
            # fp32 tensors that wait for the engine's stream on first access (engine.PendingTensor), so a request costs
            # one wait -- when Response.encode reads the first tensor -- instead of four per node.
            deferred = os.environ.get("VITB200_DEFERRED", "1") != "0"
            self.engine.set_deferred(deferred)
            if deferred and hasattr(self.engine, "prewarm_host_outputs"):
                self.engine.prewarm_host_outputs(1, int(os.environ.get("VITB200_PREWARM_REQUESTS", "16")))
            self._lock = threading.Lock()
            self._tokens_out: Optional[torch.Tensor] = None   # last token tensor handed out (device-resident copy valid)
            self._tokens_batch = 0
            self._maps_out: Dict[int, tuple] = {}             # layer -> (avg-map tensor handed out, batch): resident on the device
            self.node_names = ([self.prefix() + "embed"] + [self.prefix() + f"layer.{i}" for i in range(cfg.num_layers)]
                               + [self.prefix() + "head", self.prefix() + "rollout", self.prefix() + "transform"])
            self._images_out: Optional[torch.Tensor] = None   # last preprocessed image tensor handed out (still on the device)
            self._generation = self.engine.workspace_generation() if hasattr(self.engine, "workspace_generation") else 0

        # ---- catalogue ---------------------------------------------------------------------------
        def list_node_names(self) -> List[str]:
            return self.node_names

        def _kind(self, node_name: str) -> str:
            sub = node_name.removeprefix(self.prefix())
            if sub in ("embed", "head", "rollout", "transform"):
                return sub
            if sub.startswith("layer."):
                idx, _, half = sub[len("layer."):].partition(".")
                if idx.isdigit() and 0 <= int(idx) < self.cfg.num_layers and half in ("", "attn", "mlp"):
                    return {"": "layer", "attn": "attn_block", "mlp": "mlp_block"}[half]
            raise KeyError(node_name)

        def _layer_index(self, node_name: str) -> int:
            return int(node_name.removeprefix(self.prefix())[len("layer."):].partition(".")[0])

        def fine_node_names(self) -> List[str]:
            """Half-block nodes (`layer.<i>.attn`, `layer.<i>.mlp`): registered next to the block-granular ones."""
            return [self.prefix() + f"layer.{i}.{half}" for i in range(self.cfg.num_layers) for half in ("attn", "mlp")]

        def io(self, node_name: str, params: Optional[Dict[str, str]] = None) -> Dict:
            kind = self._kind(node_name)
            if kind == "mlp_block":
                return {"ins": ["o"], "outs": ["o"]}
            if kind in ("layer", "attn_block"):
                outs = ["o", "attn", "cls"]
                if params is not None and str(params.get("heads", "0")) == "1":
                    outs.append("heads")
                return {"ins": ["o"], "outs": outs}
            if kind == "rollout":
                return {"ins": [f"a{i}" for i in range(self.cfg.num_layers)], "outs": ["o"]}
            return {"ins": ["o"], "outs": ["o"]}

        def contents(self, node_name: str) -> str:
            kind = self._kind(node_name)
            c = self.cfg
            what = {
                "embed": f"patch {c.patch_size}x{c.patch_size} embedding + class token + position ({c.tokens} tokens x {c.hidden_dim})",
                "layer": f"EncoderBlock: {c.num_heads}-head attention + MLP {c.mlp_dim} (outs: o, attn, cls)",
                "attn_block": f"x + {c.num_heads}-head attention(LayerNorm x) (outs: o, attn, cls)",
                "mlp_block": f"x + MLP {c.mlp_dim}(LayerNorm x)",
                "head": f"LayerNorm + Linear -> {c.num_classes} logits",
                "rollout": f"attention rollout over {c.num_layers} layers -> {c.image_size // c.patch_size}x{c.image_size // c.patch_size}",
                "transform": f"resize (antialiased bilinear) + centre crop {c.image_size} + ImageNet normalisation",
            }[kind]
            return f"<p>{node_name}</p> <p>{what}</p> <p>B200 engine</p>"

        def generate_graph_json(self) -> Dict:
            """Same file format as Model.generate_graph_json (main/context.py:55-73) and VggModel's category tail
            (static/models/vgg16.py:16-29): chain embed -> layers -> head -> category, plus attn_i -> rollout."""
            names = self.list_node_names()
            L = self.cfg.num_layers
            w = int(math.sqrt(len(names) + 1))
            nodes, edges = [], []
            for i, endpoint in enumerate(names):
                nodes.append({"instance": {"kind": "net_node", "endpoint": endpoint, "params": {}},
                              "pos": {"x": (i % w) * 200, "y": int(i / w) * 200}})
            head_idx, rollout_idx = 1 + L, 2 + L
            for i in range(1, head_idx + 1):  # embed -> layer.0 -> ... -> head
                edges.append({"in_port": {"node": i - 1, "channel": "o"}, "out_port": {"node": i, "channel": "o"}})
            for i in range(L):
                edges.append({"in_port": {"node": 1 + i, "channel": "attn"},
                              "out_port": {"node": rollout_idx, "channel": f"a{i}"}})
            # preprocessing in front of the chain, as VggModel's `transform` pseudo-node (static/models/vgg16.py:31-35)
            edges.append({"in_port": {"node": names.index(self.prefix() + "transform"), "channel": "o"},
                          "out_port": {"node": 0, "channel": "o"}})
            cat_idx = len(nodes)
            nodes.append({"instance": {"kind": "category", "cats": imagenet_categories(self.cfg.num_classes)},
                          "pos": {"x": (cat_idx % w) * 200, "y": int(cat_idx / w) * 200}})
            edges.append({"in_port": {"node": head_idx, "channel": "o"}, "out_port": {"node": cat_idx, "channel": "o"}})
            return {"nodes": nodes, "edges": edges}

        def generate_fine_graph_json(self) -> Dict:
            """The half-block variant of the graph file: transform -> embed -> layer.0.attn ->
            layer.0.mlp -> ... -> head -> category, layer.i.attn `attn` -> rollout.  Same file format; every output
            channel still has one server-side consumer, so it loads in the unpatched reference too."""
            L = self.cfg.num_layers
            names = ([self.prefix() + "transform", self.prefix() + "embed"] + self.fine_node_names()
                     + [self.prefix() + "head", self.prefix() + "rollout"])
            w = int(math.sqrt(len(names) + 1))
            nodes = [{"instance": {"kind": "net_node", "endpoint": n, "params": {}},
                      "pos": {"x": (i % w) * 200, "y": int(i / w) * 200}} for i, n in enumerate(names)]
            head_idx, rollout_idx = 2 + 2 * L, 3 + 2 * L
            edges = [{"in_port": {"node": i - 1, "channel": "o"}, "out_port": {"node": i, "channel": "o"}}
                     for i in range(1, head_idx + 1)]
            edges += [{"in_port": {"node": 2 + 2 * i, "channel": "attn"}, "out_port": {"node": rollout_idx, "channel": f"a{i}"}}
                      for i in range(L)]
            cat_idx = len(nodes)
            nodes.append({"instance": {"kind": "category", "cats": imagenet_categories(self.cfg.num_classes)},
                          "pos": {"x": (cat_idx % w) * 200, "y": int(cat_idx / w) * 200}})
            edges.append({"in_port": {"node": head_idx, "channel": "o"}, "out_port": {"node": cat_idx, "channel": "o"}})
            return {"nodes": nodes, "edges": edges}

        def _graphs_dir(self) -> str:
            """static/graphs under the host's base directory: Django's settings.BASE_DIR when the class is bound to the
            reference's main.context.Model (main/context.py:99), this package's base dir otherwise."""
            if ModelBase.__module__.startswith("main."):
                from django.conf import settings

                return os.path.join(str(settings.BASE_DIR), "static/graphs")
            from .context import get_base_dir

            return os.path.join(get_base_dir(), "static/graphs")

        # ---- compute -----------------------------------------------------------------------------
        @staticmethod
        def _need(pinin, ch: str) -> torch.Tensor:
            t = pinin.get(ch)
            if t is None:
                raise Exception(f"missing input: {ch}")
            if not isinstance(t, torch.Tensor):
                raise Exception(f"input {ch} is not a tensor")
            return t

        @staticmethod
        def _host(t: torch.Tensor) -> torch.Tensor:
            return t.detach().to(device="cpu", dtype=torch.float32).contiguous()

        def _reserve(self, batch: int, flags: int = 0) -> None:
            """First thing a node call does once it knows its batch size: grow the engine's workspace NOW (nothing later
            in this call can then re-allocate it) and, if ANY call since the last one re-allocated it -- a request with
            a larger batch interleaved with this one -- forget what we believed to be resident on the device: growth
            frees and re-allocates the token stream and the map buffers without copying (and the layer stride of the
            maps changes with the capacity).  The tensors handed out earlier are CPU copies; they are uploaded again."""
            eng = self.engine
            if hasattr(eng, "reserve"):
                eng.reserve(batch, flags)
                gen = eng.workspace_generation()
                if gen != self._generation:
                    self._generation = gen
                    self._tokens_out, self._tokens_batch = None, 0
                    self._maps_out.clear()
                    self._images_out = None

        @staticmethod
        def _shape(x) -> tuple:
            """Shape of a node input.  Engine outputs carry it as a plain tuple (`_wire`, engine._issue): a metadata call
            on a deferred output is a `__torch_function__` trip (~1 us each, several per node call)."""
            w = getattr(x, "_wire", None)
            return w[2] if w is not None else tuple(x.shape)

        def _bind_tokens(self, x: torch.Tensor, flags: int = 0) -> int:
            """Make the engine's token stream equal to `x` ([N,d] or [B,N,d]); returns the batch size."""
            c = self.cfg
            shp = self._shape(x)
            if len(shp) not in (2, 3) or shp[-2:] != (c.tokens, c.hidden_dim):
                raise Exception(f"expected tokens of shape [{c.tokens}, {c.hidden_dim}] (optionally batched), got {list(shp)}")
            batch = 1 if len(shp) == 2 else shp[0]
            self._reserve(batch, flags)
            if x is self._tokens_out and batch == self._tokens_batch:
                return batch  # still resident from the previous node of this request
            self.engine.set_tokens(self._host(x).reshape(batch, c.tokens, c.hidden_dim))
            # the engine now holds x, not what it handed out last (a fanned-out graph may come back to an older tensor)
            self._tokens_out, self._tokens_batch = x, batch
            return batch

        def _emit_tokens(self, batch: int, batched: bool) -> torch.Tensor:
            c = self.cfg
            t = self.engine.get_tokens(batch, (batch, c.tokens, c.hidden_dim) if batched else (c.tokens, c.hidden_dim))
            self._tokens_out, self._tokens_batch = t, batch
            return t

        def compute(self, node_name: str, pinin, params: Optional[Dict[str, str]] = None):
            kind = self._kind(node_name)
            c = self.cfg
            g = c.image_size // c.patch_size
            out = PinoutCls()
            with self._lock:
                if kind in ("transform", "embed") and hasattr(self.engine, "begin_request"):
                    self.engine.begin_request()     # a request's first node: its outputs get a pinned slab of their own
                if kind == "transform":
                    x = self._need(pinin, "o")
                    if x.dim() not in (3, 4) or x.shape[-3] != 3:
                        raise Exception(f"expected an image of shape [3, H, W] (optionally batched), got {list(x.shape)}")
                    batched = x.dim() == 4
                    imgs = self._host(x).reshape(-1, 3, x.shape[-2], x.shape[-1])
                    self._reserve(imgs.shape[0])
                    resize = int(params["resize"]) if params and str(params.get("resize", "")).strip() else (256 if c.image_size == 224 else c.image_size)
                    y = self.engine.stage_transform(imgs, resize)
                    y = y if batched else y[0]
                    self._images_out = y
                    out.set("o", y)
                elif kind == "embed":
                    x = self._need(pinin, "o")
                    if x.dim() not in (3, 4) or tuple(x.shape[-3:]) != (3, c.image_size, c.image_size):
                        raise Exception(f"expected an image of shape [3, {c.image_size}, {c.image_size}] (optionally batched), got {list(x.shape)}")
                    batched = x.dim() == 4
                    nimg = x.shape[0] if batched else 1
                    self._reserve(nimg)
                    if x is self._images_out:
                        self.engine.stage_embed_resident(nimg)   # preprocessed by the transform node: still on the device
                    else:
                        self.engine.stage_embed(self._host(x).reshape(-1, 3, c.image_size, c.image_size))
                    self._images_out = None
                    out.set("o", self._emit_tokens(nimg, batched))
                elif kind == "mlp_block":
                    i = self._layer_index(node_name)
                    x = self._need(pinin, "o")
                    batched = len(self._shape(x)) == 3
                    batch = self._bind_tokens(x)
                    self.engine.stage_mlp_block(i, batch)
                    out.set("o", self._emit_tokens(batch, batched))
                elif kind in ("layer", "attn_block"):
                    i = self._layer_index(node_name)
                    x = self._need(pinin, "o")
                    batched = len(self._shape(x)) == 3
                    want_heads = params is not None and str(params.get("heads", "0")) == "1"
                    flags = E.EMIT_AVG | E.EMIT_CLS | (E.EMIT_HEADS if want_heads else 0)
                    batch = self._bind_tokens(x, flags)
                    lead = (batch,) if batched else ()
                    if hasattr(self.engine, "stage_layer_fetch"):
                        # the node and its three output copies in one call into the library
                        tok, amap, cls = self.engine.stage_layer_fetch(
                            i, batch, flags, kind != "layer", lead + (c.tokens, c.hidden_dim), lead + (c.tokens, c.tokens),
                            lead + (c.num_heads, g, g))
                        self._tokens_out, self._tokens_batch = tok, batch
                    else:
                        if kind == "layer":
                            self.engine.stage_layer(i, batch, flags)
                        else:
                            self.engine.stage_attn_block(i, batch, flags)
                        tok = self._emit_tokens(batch, batched)
                        amap = self.engine.get_avg_map(i, batch, lead + (c.tokens, c.tokens))
                        cls = self.engine.get_cls_grid(i, batch, lead + (c.num_heads, g, g))
                    out.set("o", tok)
                    self._maps_out[i] = (amap, batch)
                    out.set("attn", amap)
                    out.set("cls", cls)
                    if want_heads:
                        out.set("heads", self.engine.get_head_map(i, batch, lead + (c.num_heads, c.tokens, c.tokens)))
                elif kind == "head":
                    x = self._need(pinin, "o")
                    batched = len(self._shape(x)) == 3
                    batch = self._bind_tokens(x)
                    out.set("o", self.engine.stage_head(batch, (batch, c.num_classes) if batched else (c.num_classes,)))
                else:  # rollout
                    maps = [self._need(pinin, f"a{i}") for i in range(c.num_layers)]
                    shapes = [self._shape(m) for m in maps]
                    batched = len(shapes[0]) == 3
                    batch = shapes[0][0] if batched else 1
                    for i, ms in enumerate(shapes):
                        if (ms[-2:] != (c.tokens, c.tokens) or (len(ms) == 3) != batched
                                or (batched and ms[0] != batch)):
                            raise Exception(f"a{i}: expected a [{c.tokens}, {c.tokens}] map"
                                            f"{f' for each of {batch} images' if batched else ''}, got {list(ms)}")
                    self._reserve(batch, E.EMIT_AVG | E.EMIT_ROLLOUT)
                    for i, m in enumerate(maps):
                        res = self._maps_out.get(i)
                        if res is None or res[0] is not m or res[1] != batch:   # not (or no longer) resident
                            self.engine.set_avg_map(i, self._host(m).reshape(batch, c.tokens, c.tokens))
                            # the engine now holds m for layer i, not the map it handed out last: the request that map
                            # belongs to (interleaved on another thread) must upload it again for ITS rollout
                            self._maps_out[i] = (m, batch)
                    out.set("o", self.engine.stage_rollout(batch, (batch, g, g) if batched else (g, g)))
            return out

        # ---- registration: ModelNode forwards (params are dropped by the reference's ModelNode,
        #      main/context.py:119-129, so the per-node params variant is wired through our own node class)
        def register(self, ctx) -> None:
            super().register(ctx)  # graph json + one ModelNode per name
            for node_name in self.list_node_names() + self.fine_node_names():
                ctx.register(_ParamNode(self, node_name))
            try:  # the half-block graph next to the default one; same policy as the reference: log, keep going
                path = os.path.join(self._graphs_dir(), self.name + "_fine.json")
                if not os.path.exists(path):
                    import json

                    with open(path, "w") as f:
                        f.write(json.dumps(self.generate_fine_graph_json()))
            except Exception as e:
                import logging

                logging.getLogger(__name__).error("could not generate the fine-grained graph: %s", e)

    class _ParamNode:
        """Same surface as ModelNode (get_name / compute / contents / io / register) but passes ``params`` on."""

        def __init__(self, parent, name: str):
            self.parent = parent
            self.name = name

        def get_name(self) -> str:
            return self.name

        def compute(self, params, inputs):
            return self.parent.compute(self.name, inputs, params)

        def contents(self, params) -> str:
            return self.parent.contents(self.name)

        def io(self, params) -> Dict:
            return self.parent.io(self.name, params)

        def register(self, ctx) -> None:
            ctx.register(self)

    return VitB200Model


def _default_class():
    from .context import Model
    from .graph import Pinout

    return make_vit_model_class(Model, Pinout)


_cls_cache = None


def VitB200Model(*args, **kwargs):
    """Plugin class bound to this package's mirror of the plugin API."""
    global _cls_cache
    if _cls_cache is None:
        _cls_cache = _default_class()
    return _cls_cache(*args, **kwargs)


def instances() -> list:
    """Plugin entry point (what scan_nodes calls, main/context.py:154-176).  Models come from the environment:
    VITB200_MODELS="vit_b_16[,vit_s_16,...]", VITB200_DEVICE, VITB200_MAX_BATCH, VITB200_WEIGHTS_<NAME>=path.pt."""
    names = [n for n in os.environ.get("VITB200_MODELS", "vit_b_16").split(",") if n]
    device = int(os.environ.get("VITB200_DEVICE", "0"))
    max_batch = int(os.environ.get("VITB200_MAX_BATCH", "1"))
    out = []
    for n in names:
        cfg = E.CONFIGS[n]
        module = build_torchvision_vit(cfg)
        wpath = os.environ.get(f"VITB200_WEIGHTS_{n.upper()}")
        if wpath:
            module.load_state_dict(torch.load(wpath, map_location="cpu"))
        out.append(VitB200Model(n, cfg, module, device, max_batch))
    return out
