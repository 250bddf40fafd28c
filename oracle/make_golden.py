"""Generate tests/golden/* by running the UNMODIFIED reference boundary (oracle/refhost.py) in the build
container.  TEST INFRASTRUCTURE ONLY.   Usage:  python -m oracle.make_golden

The reference has no golden vectors of its own (main/tests.py:1-3 is empty), so these files are the pins:

wire_tiny.{request,response}.bin   one full POST /compute round trip: browser-format request bytes ->
        reference Request.decode -> reference Context.compute over reference Model-hosted oracle nodes ->
        reference Response.encode bytes (main/views.py:32-39).  The product's codec + scheduler + oracle plugin
        must reproduce the response byte for byte on CPU; the CUDA path must match its tensors within tolerance.
graph_kats.json   Graph.order() visit orders, Model.generate_graph_json() and node-name enumeration of the
        reference for toy inputs, plus the cos node round trip (main/nodes/cos.py) and the fan-out failure mode.
vit_small_b2.pt / vit_b16_b1.pt   forward + attention-map tensors from torchvision CPU fp32 (seeded init and
        input; see oracle/vit_oracle.py) for the CUDA parity tests and for oracle self-consistency.
"""
from __future__ import annotations

import json
import os

import torch

from . import oracle_plugin, refhost, vit_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _client_encode(nodes, edges, tensors) -> bytes:
    """Browser-side Request.encode (main/static/main/nodes/net_node.js:56-175), restated with struct."""
    import struct

    js = json.dumps({"nodes": nodes, "edges": edges}).encode()
    pad = (-(16 + len(js))) % 4
    blocks = b""
    for t in tensors:
        data = t.contiguous().numpy().tobytes()
        dims = list(t.shape)
        blocks += struct.pack(f"<II{len(dims)}I", 8 + 4 * len(dims) + len(data), len(dims), *dims) + data
    body = js + b"\0" * pad + blocks
    return struct.pack("<IIII", 16 + len(body), 0x69BABE69, len(tensors), len(js)) + body


def wire_round_trip():
    graph, context, message, _ = refhost.load()
    cfg = O.ORACLE_CONFIGS["vit_tiny_test"]
    Cls = oracle_plugin.make_oracle_model_class(context.Model, graph.Pinout)
    model = Cls("vit_tiny_test", cfg, O.build_vit(cfg, seed=0, init="stress"))
    ctx = context.Context()
    model.register(ctx)
    img = O.synthetic_images(1, cfg.image_size, seed=1234)[0]
    nodes, edges, tensors = oracle_plugin.vit_graph_request("vit_tiny_test", cfg.num_layers, img)
    req_bytes = _client_encode(nodes, edges, tensors)
    req = message.Request()
    req.decode(req_bytes)
    ctx.compute(req.graph)
    resp_bytes = message.Response(req.graph).encode()
    with open(os.path.join(GOLD, "wire_tiny.request.bin"), "wb") as f:
        f.write(req_bytes)
    with open(os.path.join(GOLD, "wire_tiny.response.bin"), "wb") as f:
        f.write(resp_bytes)
    print("wire_tiny: request", len(req_bytes), "B, response", len(resp_bytes), "B")


def graph_kats():
    graph, context, message, _ = refhost.load()
    kats = {}

    def build(n, edges, inputs):
        g = graph.Graph()
        ns = [g.add_node(f"n{i}", {}) for i in range(n)]
        for (a, ach, b, bch) in edges:
            g.connect(ns[a], ach, ns[b], bch)
        for (b, bch) in inputs:
            g.add_input(torch.zeros(1), ns[b], bch)
        return g

    cases = {
        "chain4": (4, [(0, "o", 1, "o"), (1, "o", 2, "o"), (2, "o", 3, "o")], [(0, "o")]),
        "reversed_chain": (4, [(3, "o", 2, "o"), (2, "o", 1, "o"), (1, "o", 0, "o")], [(3, "o")]),
        "diamond_two_channels": (4, [(0, "a", 1, "o"), (0, "b", 2, "o"), (1, "o", 3, "x"), (2, "o", 3, "y")], [(0, "o")]),
        "vit_like": (6, [(0, "o", 1, "o"), (1, "o", 2, "o"), (2, "o", 3, "o"), (3, "o", 4, "o"),
                         (1, "attn", 5, "a0"), (2, "attn", 5, "a1"), (3, "attn", 5, "a2")], [(0, "o")]),
        "isolated": (3, [], []),
    }
    kats["order"] = {}
    for name, (n, edges, inputs) in cases.items():
        g = build(n, edges, inputs)
        kats["order"][name] = {"n": n, "edges": edges, "inputs": inputs, "order": [x.index for x in g.order()]}

    # Model wrapper over a toy module: node names + generated graph json
    toy = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.ReLU(), torch.nn.Sequential(torch.nn.Linear(4, 2), torch.nn.Tanh()))
    m = context.Model(toy, "toy")
    kats["model"] = {"node_names": m.list_node_names(), "graph_json": m.generate_graph_json(), "io": m.io("toy:0"),
                     "contents": m.contents("toy:0")}
    torch.manual_seed(3)
    with torch.no_grad():
        for p in toy.parameters():
            p.copy_(torch.randn(p.shape))
    pin = graph.Pinout()
    x = torch.tensor([0.5, -1.0, 2.0, 0.25])
    pin.set("o", x)
    kats["model"]["state"] = {k: v.tolist() for k, v in toy.state_dict().items()}
    kats["model"]["x"] = x.tolist()
    kats["model"]["y_toy0"] = m.compute("toy:0", pin).get("o").tolist()

    # cos node through the reference context singleton (registered by scan_nodes at import)
    cos = context.context().get_node("cos")
    pin = graph.Pinout()
    pin.set("o", torch.tensor([0.0, 1.0, 2.0]))
    kats["cos"] = {"io": cos.io({}), "contents": cos.contents({"A": "2", "b": "0.5"}),
                   "y": cos.compute({"A": "2", "b": "0.5"}, pin).get("o").tolist()}
    try:
        cos.compute({}, graph.Pinout())
        kats["cos"]["missing_input_error"] = None
    except Exception as e:
        kats["cos"]["missing_input_error"] = str(e)

    # fan-out on one output channel: the second connect overwrites the producer-side edge (graph.py:64-70)
    g = graph.Graph()
    a, b, c = g.add_node("a", {}), g.add_node("b", {}), g.add_node("c", {})
    g.connect(a, "o", b, "o")
    g.connect(a, "o", c, "o")
    p = graph.Pinout()
    p.set("o", torch.ones(1))
    a.set_pinout(p)
    fan = {"b_has_tensor": b.inputs["o"].tensor is not None, "c_has_tensor": c.inputs["o"].tensor is not None}
    try:
        b.get_pinin()
        fan["b_get_pinin"] = "ok"
    except AssertionError:
        fan["b_get_pinin"] = "AssertionError"
    kats["fanout"] = fan

    # unknown endpoint -> KeyError from get_node (main/context.py:140-141)
    try:
        context.Context().get_node("nope")
    except KeyError as e:
        kats["unknown_node_error"] = type(e).__name__
    with open(os.path.join(GOLD, "graph_kats.json"), "w") as f:
        json.dump(kats, f, indent=1, sort_keys=True)
    print("graph_kats: ok")


def vit_goldens():
    for name, batch, init, fname in (("vit_small_test", 2, "stress", "vit_small_b2.pt"), ("vit_b_16", 1, "default", "vit_b16_b1.pt")):
        cfg = O.ORACLE_CONFIGS[name]
        model = O.build_vit(cfg, seed=0, init=init)
        x = O.synthetic_images(batch, cfg.image_size, seed=1234)
        r = O.forward_with_maps(model, x)
        assert torch.allclose(r["logits"], model(x), rtol=1e-4, atol=1e-5)
        gold = {
            "config": name, "batch": batch, "init": init, "seed": 0, "image_seed": 1234,
            "logits": r["logits"], "rollout": r["rollout"], "cls_maps": r["cls_maps"],
            "avg_rows": r["avg_maps"][:, :, ::16, :].contiguous(),          # every 16th query row of every layer
            "hidden_cls": r["hidden"][:, :, 0, :].contiguous(),            # class-token row after every layer
            "embed_rows": r["embed"][:, :4, :].contiguous(),
            "weight_probe": model.state_dict()["encoder.layers.encoder_layer_0.mlp.0.weight"][:2, :8].clone(),
            "image_probe": x[0, 0, 0, :8].clone(),
        }
        if name == "vit_small_test":
            gold["avg_maps"] = r["avg_maps"]
            gold["hidden"] = r["hidden"].to(torch.float16)
        gold = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in gold.items()}  # drop view storage
        torch.save(gold, os.path.join(GOLD, fname))
        print(fname, os.path.getsize(os.path.join(GOLD, fname)) // 1024, "KiB")


def main():
    os.makedirs(GOLD, exist_ok=True)
    wire_round_trip()
    graph_kats()
    vit_goldens()


if __name__ == "__main__":
    main()
