"""Host-side mirror of the reference boundary (graph / context / message) against the known answers recorded
from the unmodified reference (tests/golden/graph_kats.json, wire_tiny.*.bin).  CPU only."""
import json
import os
import struct

import pytest
import torch

from interactive_vit_b200 import context as C
from interactive_vit_b200 import graph as G
from interactive_vit_b200 import message as M
from oracle import oracle_plugin, refhost, vit_oracle as O


@pytest.fixture(scope="module")
def kats(golden_dir):
    with open(os.path.join(golden_dir, "graph_kats.json")) as f:
        return json.load(f)


def _build(n, edges, inputs):
    g = G.Graph()
    ns = [g.add_node(f"n{i}", {}) for i in range(n)]
    for (a, ach, b, bch) in edges:
        g.connect(ns[a], ach, ns[b], bch)
    for (b, bch) in inputs:
        g.add_input(torch.zeros(1), ns[b], bch)
    return g


def test_order_matches_reference_visit_order(kats):
    for name, case in kats["order"].items():
        g = _build(case["n"], case["edges"], case["inputs"])
        assert [x.index for x in g.order()] == case["order"], name


def test_order_raises_on_cycle_instead_of_spinning():
    g = G.Graph()
    a, b = g.add_node("a", {}), g.add_node("b", {})
    g.connect(a, "o", b, "o")
    g.connect(b, "o", a, "o")
    with pytest.raises(ValueError):
        g.order()


def test_pinout_and_dangling_outputs():
    g = G.Graph()
    a, b = g.add_node("a", {}), g.add_node("b", {})
    g.connect(a, "o", b, "o")
    p = G.Pinout()
    t1, t2 = torch.ones(2), torch.zeros(3)
    p.set("o", t1)
    p.set("extra", t2)
    a.set_pinout(p)
    assert b.get_pinin().get("o") is t1                      # by reference, no copy
    assert a.outputs["extra"].output is None                 # dangling edge still carries the tensor
    assert a.get_pinout().get("extra") is t2
    assert G.Pinout().get("missing") is None
    with pytest.raises(AssertionError):
        g.add_node("c", {}).get_pinin() if False else G.Node._collect({"o": G.Edge(None, None)})


def test_fanout_failure_mode_is_preserved(kats):
    g = G.Graph()
    a, b, c = g.add_node("a", {}), g.add_node("b", {}), g.add_node("c", {})
    g.connect(a, "o", b, "o")
    g.connect(a, "o", c, "o")
    p = G.Pinout()
    p.set("o", torch.ones(1))
    a.set_pinout(p)
    assert (b.inputs["o"].tensor is not None) == kats["fanout"]["b_has_tensor"]
    assert (c.inputs["o"].tensor is not None) == kats["fanout"]["c_has_tensor"]
    if kats["fanout"]["b_get_pinin"] == "AssertionError":
        with pytest.raises(AssertionError):
            b.get_pinin()


def test_fan_out_extension_feeds_every_consumer():
    """with Graph(fan_out=True) one output channel reaches all its consumers (the default
    keeps the reference's overwrite, test above), the response still lists the channel once, and the visit order
    is the same worklist."""
    g = G.Graph(fan_out=True)
    a, b, c, d = (g.add_node(n, {}) for n in "abcd")
    g.connect(a, "o", b, "o")
    g.connect(a, "o", c, "o")
    g.connect(a, "o", d, "x")
    assert a.outputs["o"].output.node is b and [t.output.node for t in a.outputs["o"].taps] == [c, d]
    pin = G.Pinout()
    t = torch.arange(3.0)
    pin.set("o", t)
    a.set_pinout(pin)
    assert b.get_pinin().get("o") is t and c.get_pinin().get("o") is t and d.get_pinin().get("x") is t
    assert list(a.get_pinout().pinout) == ["o"]
    order = [n.name for n in g.order()]
    assert order[0] == "a" and sorted(order[1:]) == ["b", "c", "d"]

    # through the wire: a "logit lens" request (the head applied to the embedding AND to the last layer) decodes into
    # a graph whose embed output has two consumers
    from interactive_vit_b200 import message as M

    nodes = [{"endpoint": "m:embed", "params": {}}, {"endpoint": "m:layer.0", "params": {}},
             {"endpoint": "m:head", "params": {}}, {"endpoint": "m:head", "params": {}}]
    edges = [{"tensor": 0, "out_port": {"node": 0, "channel": "o"}},
             {"in_port": {"node": 0, "channel": "o"}, "out_port": {"node": 1, "channel": "o"}},
             {"in_port": {"node": 1, "channel": "o"}, "out_port": {"node": 2, "channel": "o"}},
             {"in_port": {"node": 0, "channel": "o"}, "out_port": {"node": 3, "channel": "o"}}]
    blob = M.encode_request(nodes, edges, [torch.zeros(3, 8, 8)])
    req = M.Request(fan_out=True)
    req.decode(blob)
    n = req.graph.nodes
    assert n[0].outputs["o"].output.node is n[1] and n[0].outputs["o"].taps[0].output.node is n[3]
    ref_like = M.Request()
    ref_like.decode(blob)
    assert ref_like.graph.nodes[0].outputs["o"].output.node is ref_like.graph.nodes[3]   # reference: last consumer wins


def test_model_wrapper_matches_reference(kats, tmp_path):
    toy = torch.nn.Sequential(torch.nn.Linear(4, 4), torch.nn.ReLU(), torch.nn.Sequential(torch.nn.Linear(4, 2), torch.nn.Tanh()))
    m = C.Model(toy, "toy")
    k = kats["model"]
    assert m.list_node_names() == k["node_names"]
    assert m.generate_graph_json() == k["graph_json"]
    assert m.io("toy:0") == k["io"]
    assert m.contents("toy:0") == k["contents"]
    toy.load_state_dict({n: torch.tensor(v) for n, v in k["state"].items()})
    pin = G.Pinout()
    pin.set("o", torch.tensor(k["x"]))
    assert torch.allclose(m.compute("toy:0", pin).get("o"), torch.tensor(k["y_toy0"]), rtol=1e-6, atol=1e-7)
    assert not toy.training
    # register(): writes static/graphs/<name>.json once, registers one ModelNode per node name
    os.makedirs(tmp_path / "static" / "graphs")
    C.set_base_dir(str(tmp_path))
    try:
        ctx = C.Context()
        m.register(ctx)
        assert json.load(open(tmp_path / "static" / "graphs" / "toy.json")) == k["graph_json"]
        assert sorted(ctx.nodes) == sorted(k["node_names"])
        node = ctx.get_node("toy:0")
        assert isinstance(node, C.ModelNode) and node.io({}) == k["io"]
        with pytest.raises(KeyError):
            ctx.get_node("nope")
    finally:
        C.set_base_dir(None)


def test_nodekind_defaults_and_plugin_discovery(kats, tmp_path):
    base = C.NodeKind("x")
    assert base.contents({"A": "2"}) == "x?A=2"
    with pytest.raises(Exception, match="TODO: implement Node.io"):
        base.io({})
    with pytest.raises(Exception, match="TODO: implement Node.compute"):
        base.compute({}, G.Pinout())
    # a cos-like plugin written against the mirror behaves like the reference's main/nodes/cos.py
    os.makedirs(tmp_path / "main" / "nodes")
    (tmp_path / "main" / "nodes" / "cosine.py").write_text(
        "import torch\nfrom interactive_vit_b200.context import NodeKind\nfrom interactive_vit_b200.graph import Pinout\n"
        "class Cos(NodeKind):\n"
        "    def __init__(self): super().__init__('cos')\n"
        "    def io(self, params): return {'ins': ['o'], 'outs': ['o']}\n"
        "    def compute(self, params, inputs):\n"
        "        x = inputs.get('o')\n"
        "        if x is None: raise Exception('missing input: o')\n"
        "        r = Pinout(); r.set('o', torch.cos(float(params.get('A', 1.0)) * x + float(params.get('b', 0.0)))); return r\n"
        "def instances(): return [Cos()]\n")
    (tmp_path / "main" / "nodes" / "broken.py").write_text("raise RuntimeError('boom')\n")
    C.set_base_dir(str(tmp_path))
    try:
        ctx = C.Context()
        ok = C.scan_nodes(["main/nodes"], ctx)
        assert len(ok) == 1 and "cos" in ctx.nodes          # broken plugin is logged and skipped, not fatal
        g = G.Graph()
        n = g.add_node("cos", {"A": "2", "b": "0.5"})
        g.add_input(torch.tensor([0.0, 1.0, 2.0]), n, "o")
        ctx.compute(g)
        assert torch.allclose(n.get_pinout().get("o"), torch.tensor(kats["cos"]["y"]))
        g2 = G.Graph()
        g2.add_node("cos", {})
        with pytest.raises(Exception, match=kats["cos"]["missing_input_error"]):
            ctx.compute(g2)
    finally:
        C.set_base_dir(None)


def test_wire_header_known_answers(golden_dir):
    req = open(os.path.join(golden_dir, "wire_tiny.request.bin"), "rb").read()
    size, magic, blocks, jsize = struct.unpack_from("<IIII", req, 0)
    assert size == len(req) and magic == 0x69BABE69 and blocks == 1
    resp = open(os.path.join(golden_dir, "wire_tiny.response.bin"), "rb").read()
    size, magic, blocks, jsize = struct.unpack_from("<IIII", resp, 0)
    assert size == len(resp) and magic == 0xDEADBEEF
    assert M.align_next(17, 4) == 20 and M.align_next(16, 4) == 16


def test_wire_round_trip_reproduces_reference_bytes(golden_dir):
    """Browser request bytes -> mirror Request.decode -> mirror Context.compute over the oracle plugin hosted by
    the mirror Model -> mirror Response.encode == the bytes the unmodified reference produced."""
    req_bytes = open(os.path.join(golden_dir, "wire_tiny.request.bin"), "rb").read()
    want = open(os.path.join(golden_dir, "wire_tiny.response.bin"), "rb").read()
    cfg = O.ORACLE_CONFIGS["vit_tiny_test"]
    Cls = oracle_plugin.make_oracle_model_class(C.Model, G.Pinout)
    model = Cls("vit_tiny_test", cfg, O.build_vit(cfg, seed=0, init="stress"))
    ctx = C.Context()
    for name in model.list_node_names():
        C.ModelNode(model, name).register(ctx)
    req = M.Request()
    req.decode(req_bytes)
    assert [n.name for n in req.graph.nodes][:2] == ["vit_tiny_test:embed", "vit_tiny_test:layer.0"]
    ctx.compute(req.graph)
    got = M.Response(req.graph).encode()
    assert len(got) == len(want)
    js = struct.unpack_from("<I", want, 12)[0]
    hdr = M.align_next(16 + js, 4)
    assert got[:hdr] == want[:hdr]                          # header + json index + padding: byte-identical
    a, b = M.decode_response(got), M.decode_response(want)
    assert {k: sorted(v) for k, v in a.items()} == {k: sorted(v) for k, v in b.items()}
    for node in b:
        for ch in b[node]:
            assert a[node][ch].shape == b[node][ch].shape
            assert torch.allclose(a[node][ch], b[node][ch], rtol=1e-4, atol=1e-6), (node, ch)
    # the client-side encoder of the mirror reproduces the browser's request bytes exactly
    img = O.synthetic_images(1, cfg.image_size, seed=1234)[0]
    nodes, edges, tensors = oracle_plugin.vit_graph_request("vit_tiny_test", cfg.num_layers, img)
    assert M.encode_request(nodes, edges, tensors) == req_bytes


def test_codec_edge_cases():
    # 0-d, empty and non-contiguous tensors; json length not a multiple of 4
    g = G.Graph()
    n = g.add_node("x", {})
    p = G.Pinout()
    p.set("s", torch.tensor(3.5))
    p.set("e", torch.zeros(0, 4))
    p.set("t", torch.arange(6, dtype=torch.float32).reshape(2, 3).t())
    p.set("d", torch.arange(4, dtype=torch.float64))      # converted instead of mis-encoded
    n.set_pinout(p)
    out = M.decode_response(M.Response(g).encode())[0]
    assert out["s"].shape == () and out["s"].item() == 3.5
    assert out["e"].shape == (0, 4)
    assert torch.equal(out["t"], torch.arange(6, dtype=torch.float32).reshape(2, 3).t())
    assert torch.equal(out["d"], torch.arange(4, dtype=torch.float32))
    for extra in ("", "a", "ab", "abc"):
        b = M.encode_request([{"endpoint": "cos" + extra, "params": {}}], [{"tensor": 0, "out_port": {"node": 0, "channel": "o"}}],
                             [torch.ones(2, 2)])
        r = M.Request()
        r.decode(b)
        assert torch.equal(r.graph.nodes[0].get_pinin().get("o"), torch.ones(2, 2))
    bad = bytearray(b)
    bad[4] ^= 0xFF
    with pytest.raises(AssertionError):
        M.Request().decode(bytes(bad))


@pytest.mark.skipif(not refhost.available(), reason="reference tree only exists in the build container")
def test_mirror_and_reference_codecs_interoperate():
    graph, context, message, _ = refhost.load()
    t = torch.randn(3, 5)
    b = M.encode_request([{"endpoint": "cos", "params": {"A": "2"}}], [{"tensor": 0, "out_port": {"node": 0, "channel": "o"}}], [t])
    ref_req = message.Request()
    ref_req.decode(b)
    context.context().compute(ref_req.graph)
    ref_bytes = message.Response(ref_req.graph).encode()
    mine = M.Request()
    mine.decode(b)
    ctx = C.Context()

    class Cos(C.NodeKind):
        def compute(self, params, inputs):
            r = G.Pinout()
            r.set("o", torch.cos(float(params["A"]) * inputs.get("o")))
            return r

    ctx.register(Cos("cos"))
    ctx.compute(mine.graph)
    assert M.Response(mine.graph).encode() == ref_bytes


def test_vit_plugin_catalogue_without_gpu():
    """io()/contents()/graph-json of the B200 plugin class do not need the device: build the class around a stub
    engine and compare the catalogue with the oracle plugin (the contract the UI sees)."""
    from interactive_vit_b200 import engine as E, vit_plugin as P

    class StubEngine:
        def load_state_dict(self, sd):
            self.keys = sorted(sd)

        def set_deferred(self, on):
            self.deferred = on

    ocfg = O.ORACLE_CONFIGS["vit_tiny_test"]
    cfg = E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                      ocfg.num_classes)
    plug = P.VitB200Model("vit_tiny_test", cfg, O.build_vit(ocfg), engine=StubEngine())
    oracle = oracle_plugin.make_oracle_model_class(C.Model, G.Pinout)("vit_tiny_test", ocfg, O.build_vit(ocfg))
    assert plug.list_node_names() == oracle.list_node_names()
    for n in plug.list_node_names():
        assert plug.io(n) == oracle.io(n)
        assert n in plug.contents(n)
        assert "/" not in n                                  # node names are URL path segments (main/urls.py:12-13)
    assert plug.fine_node_names() == oracle.fine_node_names() and len(plug.fine_node_names()) == 2 * ocfg.num_layers
    for n in plug.fine_node_names():                         # half-block nodes: registered, not in the default graph
        assert plug.io(n) == oracle.io(n) and n in plug.contents(n) and "/" not in n
    assert plug.io("vit_tiny_test:layer.1.mlp") == {"ins": ["o"], "outs": ["o"]}
    with pytest.raises(KeyError):
        plug.io("vit_tiny_test:layer.1.qkv")
    assert plug.io("vit_tiny_test:layer.0", {"heads": "1"})["outs"] == ["o", "attn", "cls", "heads"]
    with pytest.raises(KeyError):
        plug.io("vit_tiny_test:layer.7")
    assert "encoder.layers.encoder_layer_1.mlp.3.bias" in plug.engine.keys and len(plug.engine.keys) == 4 + 12 * 2 + 4
    g = plug.generate_graph_json()
    L = ocfg.num_layers
    assert [n["instance"].get("endpoint") for n in g["nodes"][:-1]] == plug.list_node_names()
    assert g["nodes"][-1]["instance"]["kind"] == "category" and len(g["nodes"][-1]["instance"]["cats"]) == ocfg.num_classes
    chain = [(e["in_port"]["node"], e["in_port"]["channel"], e["out_port"]["node"], e["out_port"]["channel"]) for e in g["edges"]]
    assert (0, "o", 1, "o") in chain and (L, "o", L + 1, "o") in chain and (L + 1, "o", L + 4, "o") in chain
    assert (L + 3, "o", 0, "o") in chain                  # transform -> embed (VggModel's transform pseudo-node)
    assert all((1 + i, "attn", L + 2, f"a{i}") in chain for i in range(L))
    # every output channel has at most one server-side consumer (Graph.connect keeps one edge per channel)
    outs = [(a, ch) for (a, ch, _, _) in chain]
    assert len(outs) == len(set(outs))
    # register(): graph json written once, nodes registered under their names, params reach compute()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "static", "graphs"))
        C.set_base_dir(d)
        try:
            ctx = C.Context()
            plug.register(ctx)
            assert sorted(ctx.nodes) == sorted(plug.list_node_names() + plug.fine_node_names())
            assert ctx.get_node("vit_tiny_test:layer.1.attn").io({})["outs"] == ["o", "attn", "cls"]
            assert json.load(open(os.path.join(d, "static", "graphs", "vit_tiny_test.json"))) == g
            assert ctx.get_node("vit_tiny_test:layer.1").io({"heads": "1"})["outs"][-1] == "heads"
            # the half-block graph file: a chain with one consumer per channel, every endpoint registered
            fine = json.load(open(os.path.join(d, "static", "graphs", "vit_tiny_test_fine.json")))
            assert fine == plug.generate_fine_graph_json()
            ends = [n["instance"].get("endpoint") for n in fine["nodes"][:-1]]
            assert ends[:4] == ["vit_tiny_test:transform", "vit_tiny_test:embed", "vit_tiny_test:layer.0.attn", "vit_tiny_test:layer.0.mlp"]
            assert all(e in ctx.nodes for e in ends)
            fo = [(e["in_port"]["node"], e["in_port"]["channel"]) for e in fine["edges"]]
            assert len(fo) == len(set(fo)) and (2 + 2 * L, "o") in fo
        finally:
            C.set_base_dir(None)


def test_pending_tensor_waits_on_first_data_access_only():
    """engine.PendingTensor (deferred node outputs) without a GPU: it is a torch.Tensor for the reference's graph
    code; metadata never waits; numpy() / data_ptr() / any torch function / the reference codec idiom wait exactly
    once through the engine's drain; results of operations are plain tensors; nested arguments are found."""
    import numpy as np

    import interactive_vit_b200.engine as E

    class FakeEngine:
        def __init__(self):
            self.drains = []

        def _drain(self, seq):
            self.drains.append(seq)

    def pending(eng, seq, *shape):
        t = torch.arange(float(np.prod(shape))).reshape(*shape).as_subclass(E.PendingTensor)
        t._seq, t._engine = seq, eng
        return t

    eng = FakeEngine()
    t = pending(eng, 7, 2, 3)
    assert isinstance(t, torch.Tensor)
    assert tuple(t.shape) == (2, 3) and t.dim() == 2 and t.ndim == 2 and t.numel() == 6 and len(t) == 2
    assert t.dtype == torch.float32 and t.device.type == "cpu" and t.is_contiguous() and t.size(1) == 3
    assert eng.drains == [], "metadata must not wait"
    raw = t.numpy().tobytes()                       # main/message.py:115
    assert eng.drains == [7] and len(raw) == 24
    assert type(t + 1) is torch.Tensor and type(t[0]) is torch.Tensor and type(t.detach()) is torch.Tensor
    assert eng.drains == [7], "a tensor waits once"
    for touch in (lambda x: x.data_ptr(), lambda x: np.asarray(x), lambda x: x.sum().item(), lambda x: x.reshape(3, 2),
                  lambda x: x.to(torch.float64), lambda x: torch.equal(x, x), lambda x: x.tolist(), lambda x: x.clone()):
        e2 = FakeEngine()
        touch(pending(e2, 1, 2, 3))
        assert e2.drains == [1], touch
    e3 = FakeEngine()
    a, b = pending(e3, 1, 2, 3), pending(e3, 2, 2, 3)
    out = torch.cat([a, b])                         # tensors inside a list argument
    assert sorted(e3.drains) == [1, 2] and type(out) is torch.Tensor and out.shape == (4, 3)
    e4 = FakeEngine()
    torch.add(torch.zeros(2, 3), other=pending(e4, 5, 2, 3))   # keyword argument
    assert e4.drains == [5]
    # through this package's encoder and the byte-compatible decoder
    from interactive_vit_b200 import message as M

    e5 = FakeEngine()
    r = M.Response.__new__(M.Response)
    r.outputs = {}
    r.set_output(0, "o", pending(e5, 9, 2, 3))
    back = M.decode_response(r.encode())
    assert e5.drains == [9] and torch.equal(back[0]["o"], torch.arange(6.0).reshape(2, 3))


def test_pending_tensor_through_the_unmodified_reference_graph_and_encoder():
    """The reference's OWN code (main/graph.py Node/Pinout/Edge plumbing, main/message.py Response.encode, imported
    unmodified from /root/reference) handles PendingTensors: they pass through set_pinout / get_pinin by reference
    without waiting, and Response.encode's t.numpy().tobytes() (message.py:115) waits once and ships the right bytes."""
    from oracle import refhost

    if not refhost.available():
        pytest.skip("/root/reference is only present in the build container")
    import interactive_vit_b200.engine as E
    from interactive_vit_b200 import message as M

    rgraph, _, rmessage, _ = refhost.load()

    class FakeEngine:
        def __init__(self):
            self.drains = []

        def _drain(self, seq):
            self.drains.append(seq)

    eng = FakeEngine()
    vals = torch.arange(12.0).reshape(3, 4)
    t = vals.clone().as_subclass(E.PendingTensor)
    t._seq, t._engine = 3, eng
    g = rgraph.Graph()
    a = g.add_node("producer", {})
    b = g.add_node("consumer", {})
    g.connect(a, "o", b, "o")
    out = rgraph.Pinout()
    out.set("o", t)
    a.set_pinout(out)                       # graph.py:22-29
    got = b.get_pinin().get("o")            # graph.py:15-20
    assert got is t and eng.drains == [], "moving the tensor between nodes must neither copy nor wait"
    wire = rmessage.Response(g).encode()    # message.py:75-121, unmodified
    assert eng.drains == [3]
    back = M.decode_response(wire)
    assert torch.equal(back[a.index]["o"], vals)


def test_response_encoded_in_place_from_a_wire_ready_slab():
    """message.Response.encode's zero-copy path (engine outputs of one request sit in ONE slab with block-header gaps,
    engine.VitEngine._host_out): the response is a view of the slab, decodes to the same tensors as the copying encoder's
    bytes, lists the blocks in slab order whatever the node order, and falls back to the copying encoder when a tensor
    does not belong to the slab.  Simulated here with an ordinary (unpinned) slab; the GPU suite drives the real one."""
    import interactive_vit_b200.engine as E
    from interactive_vit_b200 import message as M

    slab = torch.zeros(1 << 20, dtype=torch.uint8)
    off = E.VitEngine.PREFIX_GAP
    made = []
    for shape in [(7, 5), (3, 4, 2), (11,), (5, 5)]:
        n = 1
        for d in shape:
            n *= d
        hdr = 8 + 4 * len(shape)
        o = off + hdr
        v = slab[o:o + n * 4].view(torch.float32).view(*shape)
        v.copy_(torch.arange(n, dtype=torch.float32).reshape(shape) + len(made))
        t = v.as_subclass(E.PendingTensor)
        t._wire, t._seq, t._engine = (slab, o, tuple(shape)), len(made) + 1, None
        made.append(t)
        off = o + n * 4

    def response(order):
        r = M.Response.__new__(M.Response)
        r.outputs = {}
        for node, ch, idx in order:
            r.outputs.setdefault(node, {})[ch] = made[idx]
        return r

    # node order == slab order, and node order != slab order (the scheduler ran node 2 before node 1)
    for order in ([(0, "o", 0), (1, "o", 1), (1, "attn", 2), (2, "o", 3)], [(0, "o", 0), (2, "o", 1), (2, "attn", 2), (1, "o", 3)]):
        r = response(order)
        enc = r.encode()
        assert isinstance(enc, memoryview) and enc.obj is not None
        plain = M.Response.__new__(M.Response)
        plain.outputs = {n: {ch: t.clone().as_subclass(torch.Tensor) for ch, t in chans.items()} for n, chans in r.outputs.items()}
        joined = plain.encode()
        assert isinstance(joined, bytes) and len(joined) == len(enc)
        a, b = M.decode_response(enc), M.decode_response(joined)
        assert {n: sorted(v) for n, v in a.items()} == {n: sorted(v) for n, v in b.items()}
        assert all(torch.equal(a[n][ch], b[n][ch]) for n in b for ch in b[n])
        assert bytes(r.encode()) == bytes(enc)      # idempotent
    # a tensor from elsewhere, or a missing block of the slab: the copying encoder takes over
    r = response([(0, "o", 0), (1, "o", 1)])
    r.outputs[2] = {"o": torch.ones(3)}
    assert isinstance(r.encode(), bytes)
    assert isinstance(response([(0, "o", 0), (1, "o", 2)]).encode(), bytes)      # block 1 skipped: not contiguous
    assert M.decode_response(response([(0, "o", 0), (1, "o", 2)]).encode())[1]["o"].shape == (11,)
