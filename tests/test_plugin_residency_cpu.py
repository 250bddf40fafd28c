"""The plugin's residency bookkeeping on CPU (no GPU, no library): `vit_plugin.VitB200Model` driven over a MODEL of the
engine -- one resident token stream, one resident head-averaged map per layer, every stage computed by the oracle's stage
functions (test infrastructure) -- so that what the plugin believes to be resident can be checked against what "the device"
really holds.  The requests interleave the way the reference's thread-per-request server lets them (one plugin lock per node
call, `ref:main/context.py:149-152`): every request must come out as if it had been served alone.

This is the CPU twin of tests/test_gpu_forward.py::test_concurrent_requests_from_two_threads, which found that a rollout node
uploading request A's layer map left request B's record claiming that layer resident."""
import itertools

import torch

from interactive_vit_b200 import vit_plugin as P
from interactive_vit_b200.graph import Pinout
from oracle import vit_oracle as O


class ModelEngine:
    """What the plugin may call on an engine (vit_plugin.py), with the device state held as CPU tensors."""

    def __init__(self, ocfg, module):
        self.ocfg, self.module = ocfg, module
        self.x = None                 # resident token stream [B,N,d]
        self.maps = {}                # layer -> resident head-averaged map [B,N,N]
        self.cls = {}
        self.uploads = {"tokens": 0, "maps": 0}
        self.gen = 0
        self.cap = 0

    # construction-time surface
    def load_state_dict(self, sd):
        pass

    def set_deferred(self, on):
        pass

    def workspace_generation(self):
        return self.gen

    def reserve(self, batch, flags=0):
        if batch > self.cap:          # growth re-allocates: resident state is gone
            self.cap, self.gen = batch, self.gen + 1
            self.x, self.maps, self.cls = None, {}, {}

    def begin_request(self):
        pass

    # stages
    def stage_transform(self, images, resize):
        self.images = O.preprocess(images, self.ocfg.image_size, resize)      # stays "on the device" for the embed node
        return self.images.clone()

    def stage_embed_resident(self, batch):
        self.x = O.embed(self.module, self.images[:batch])

    def stage_embed(self, images):
        self.uploads["images"] = self.uploads.get("images", 0) + 1
        self.images = images.clone()
        self.x = O.embed(self.module, images)

    def set_tokens(self, tokens):
        self.uploads["tokens"] += 1
        self.x = tokens.clone()

    def get_tokens(self, batch, shape=None):
        return self.x[:batch].clone().reshape(shape) if shape is not None else self.x[:batch].clone()

    def _after_attention(self, layer, p):
        self.maps[layer] = p.mean(1)
        self.cls[layer] = p[:, :, 0, 1:]

    def stage_layer(self, layer, batch, flags):
        self.x, p = O.encoder_layer(self.module, layer, self.x)
        self._after_attention(layer, p)

    def stage_attn_block(self, layer, batch, flags=0):
        self.x, p = O.encoder_attn_half(self.module, layer, self.x)
        self._after_attention(layer, p)

    def stage_mlp_block(self, layer, batch):
        self.x = O.encoder_mlp_half(self.module, layer, self.x)

    def get_avg_map(self, layer, batch, shape=None):
        return self.maps[layer][:batch].clone().reshape(shape)

    def get_cls_grid(self, layer, batch, shape=None):
        return self.cls[layer][:batch].clone().reshape(shape)

    def set_avg_map(self, layer, amap):
        self.uploads["maps"] += 1
        self.maps[layer] = amap.clone()

    def stage_head(self, batch, shape=None):
        return O.head(self.module, self.x[:batch]).reshape(shape)

    def stage_rollout(self, batch, shape=None):
        L = self.ocfg.num_layers
        return O.rollout_from_avg([self.maps[i][:batch] for i in range(L)]).reshape(shape)


def _plugin(name="vit_tiny_test"):
    import interactive_vit_b200.engine as E

    ocfg = O.ORACLE_CONFIGS[name]
    module = O.build_vit(ocfg, seed=0, init="stress")
    cfg = E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                      ocfg.num_classes)
    eng = ModelEngine(ocfg, module)
    return P.VitB200Model(name, cfg, module, 0, 1, engine=eng), eng, ocfg, module


def _steps(name, L, image, half_blocks=False, transform=False):
    """The node calls of one request, in scheduler order, as (node, input channels -> source step / image)."""
    steps = [("transform", {"o": ("image", image)}), ("embed", {"o": ("step", 0, "o")})] if transform else [("embed", {"o": ("image", image)})]
    prev = len(steps) - 1
    map_src = []
    for i in range(L):
        if half_blocks:
            steps.append((f"layer.{i}.attn", {"o": ("step", prev, "o")}))
            map_src.append(len(steps) - 1)
            steps.append((f"layer.{i}.mlp", {"o": ("step", len(steps) - 1, "o")}))
        else:
            steps.append((f"layer.{i}", {"o": ("step", prev, "o")}))
            map_src.append(len(steps) - 1)
        prev = len(steps) - 1
    steps.append(("head", {"o": ("step", prev, "o")}))
    steps.append(("rollout", {f"a{i}": ("step", s, "attn") for i, s in enumerate(map_src)}))
    return steps


class _Request:
    def __init__(self, plug, name, steps):
        self.plug, self.name, self.steps, self.outs, self.pos = plug, name, steps, [], 0

    def done(self):
        return self.pos == len(self.steps)

    def step(self):
        node, ins = self.steps[self.pos]
        pin = Pinout()
        for ch, src in ins.items():
            pin.set(ch, src[1] if src[0] == "image" else self.outs[src[1]].get(src[2]))
        self.outs.append(self.plug.compute(f"{self.name}:{node}", pin))
        self.pos += 1

    def result(self):
        return [{ch: t.clone() for ch, t in out.items()} for out in self.outs]


def _serve_alone(plug, name, steps):
    r = _Request(plug, name, steps)
    while not r.done():
        r.step()
    return r.result()


def _same(a, b):
    return len(a) == len(b) and all(sorted(x) == sorted(y) and all(torch.equal(x[c], y[c]) for c in x) for x, y in zip(a, b))


def test_interleaved_requests_equal_requests_served_alone():
    plug, eng, ocfg, module = _plugin()
    name, L = "vit_tiny_test", ocfg.num_layers
    imgs = O.synthetic_images(3, ocfg.image_size, seed=3)
    plans = [_steps(name, L, imgs[0]), _steps(name, L, imgs[1]), _steps(name, L, imgs[2], half_blocks=True)]
    alone = [_serve_alone(plug, name, p) for p in plans]
    # the oracle's own forward: the model engine and the plugin plumbing reproduce it
    ref = O.forward_with_maps(module, imgs[:1])
    assert torch.allclose(alone[0][1 + L]["o"], ref["logits"][0], atol=1e-5)
    assert torch.allclose(alone[0][2 + L]["o"].reshape(-1), ref["rollout"][0], atol=1e-6)
    # every way of interleaving two requests step by step that a pair of threads can produce is too many; take
    # round-robin with every phase shift, plus "B overtakes A just before A's rollout / head" patterns
    n = [len(p) for p in plans]
    schedules = []
    for a, b in itertools.permutations(range(3), 2):
        for shift in range(0, n[a], 2):
            order = [a] * shift
            ia, ib = shift, 0
            while ia < n[a] or ib < n[b]:
                if ib < n[b]:
                    order.append(b), (ib := ib + 1)
                if ia < n[a]:
                    order.append(a), (ia := ia + 1)
            schedules.append(((a, b), order))
        schedules.append(((a, b), [a] * (n[a] - 1) + [b] * n[b] + [a]))          # B runs whole before A's last node
        schedules.append(((a, b), [a] * (n[a] - 2) + [b] * (n[b] - 1) + [a, a, b]))
    for (a, b), order in schedules:
        reqs = {a: _Request(plug, name, plans[a]), b: _Request(plug, name, plans[b])}
        for who in order:
            reqs[who].step()
        assert reqs[a].done() and reqs[b].done()
        assert _same(reqs[a].result(), alone[a]), ("request", a, "disturbed by", b, order)
        assert _same(reqs[b].result(), alone[b]), ("request", b, "disturbed by", a, order)


def test_undisturbed_request_uploads_nothing():
    """The residency shortcuts are what keep a request on the device: served alone, no token stream and no map is uploaded."""
    plug, eng, ocfg, _ = _plugin()
    name, L = "vit_tiny_test", ocfg.num_layers
    img = O.synthetic_images(1, ocfg.image_size, seed=4)[0]
    _serve_alone(plug, name, _steps(name, L, img))
    assert eng.uploads == {"tokens": 0, "maps": 0, "images": 1}
    _serve_alone(plug, name, _steps(name, L, img, half_blocks=True))
    assert eng.uploads == {"tokens": 0, "maps": 0, "images": 2}
    raw = torch.rand(3, 80, 72)
    _serve_alone(plug, name, _steps(name, L, raw, transform=True))      # preprocessed on the device: embed takes it from there
    assert eng.uploads == {"tokens": 0, "maps": 0, "images": 2}


def test_growing_batch_forgets_resident_state():
    """A batched request between A's layers and A's rollout re-allocates the workspace (ModelEngine.reserve drops its
    state like the library does): A's rollout and head upload what they need again and still match."""
    plug, eng, ocfg, _ = _plugin()
    name, L = "vit_tiny_test", ocfg.num_layers
    imgs = O.synthetic_images(3, ocfg.image_size, seed=5)
    plan_a = _steps(name, L, imgs[0])
    want = _serve_alone(plug, name, plan_a)
    a = _Request(plug, name, plan_a)
    for _ in range(1 + L):
        a.step()
    b = _Request(plug, name, _steps(name, L, imgs[1:3]))
    b.step(), b.step()
    while not a.done():
        a.step()
    assert _same(a.result(), want)
    assert eng.uploads["maps"] == L and eng.uploads["tokens"] >= 1


def test_transform_nodes_of_interleaved_requests():
    """`<model>:transform` leaves the preprocessed image on the device for the embed node; another request's transform or
    embed in between must make the embed node upload its own image again."""
    plug, eng, ocfg, _ = _plugin()
    name, L = "vit_tiny_test", ocfg.num_layers
    g = torch.Generator().manual_seed(11)
    raw = [torch.rand(3, 90, 70, generator=g), torch.rand(3, 66, 100, generator=g)]
    plans = [_steps(name, L, raw[0], transform=True), _steps(name, L, raw[1], transform=True),
             _steps(name, L, O.synthetic_images(1, ocfg.image_size, seed=6)[0])]
    alone = [_serve_alone(plug, name, p) for p in plans]
    for a, b in itertools.permutations(range(3), 2):
        for lead in (1, 2):           # B's nodes start after A's transform / after A's embed
            reqs = {a: _Request(plug, name, plans[a]), b: _Request(plug, name, plans[b])}
            order = [a] * lead + [b] * 2 + [a] * (len(plans[a]) - lead) + [b] * (len(plans[b]) - 2)
            for who in order:
                reqs[who].step()
            assert _same(reqs[a].result(), alone[a]) and _same(reqs[b].result(), alone[b]), (a, b, lead)
