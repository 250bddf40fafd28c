"""Host logic of the ViT plugin (vit_plugin.py) on the CPU, with the oracle standing in for the engine
(tests/_oracle_engine.py): the class bound to the UNMODIFIED reference `main.context.Model`,
residency shortcuts under interleaved requests of different batch sizes (workspace re-allocation), batched requests
through the wire codec.  The same scenarios run against the real engine in tests/test_gpu_forward.py."""
import json
import os

import pytest
import torch

from interactive_vit_b200 import context as C
from interactive_vit_b200 import engine as E
from interactive_vit_b200 import graph as G
from interactive_vit_b200 import message as M
from interactive_vit_b200 import vit_plugin as P
from oracle import oracle_plugin, refhost, vit_oracle as O

from _oracle_engine import OracleEngine

NAME = "vit_tiny_test"


def _cfg(ocfg):
    return E.VitConfig(ocfg.image_size, ocfg.patch_size, ocfg.num_layers, ocfg.num_heads, ocfg.hidden_dim, ocfg.mlp_dim,
                       ocfg.num_classes)


def _plugin(cls=None, max_batch=1):
    ocfg = O.ORACLE_CONFIGS[NAME]
    module = O.build_vit(ocfg, seed=0, init="stress")
    cfg = _cfg(ocfg)
    eng = OracleEngine(cfg, module, max_batch)
    plug = (cls or P.VitB200Model)(NAME, cfg, module, engine=eng)
    return ocfg, module, plug, eng


def test_plugin_class_bound_to_the_unmodified_reference_model(golden_dir):
    """`make_vit_model_class(main.context.Model, main.graph.Pinout)` -- the plugin subclassing the
    reference's OWN `Model` (imported unmodified from /root/reference), registered in the reference's OWN `Context`,
    driven by the reference's OWN `Request.decode -> Context.compute -> Response.encode`.  With the oracle as the
    engine the response bytes equal the golden response the oracle plugin produced through the same reference code."""
    if not refhost.available():
        pytest.skip("/root/reference is only present in the build container")
    rgraph, rcontext, rmessage, base = refhost.load()
    Cls = P.make_vit_model_class(rcontext.Model, rgraph.Pinout)
    assert issubclass(Cls, rcontext.Model)
    ocfg, module, plug, eng = _plugin(Cls)
    ctx = rcontext.Context()
    for fname in (NAME + ".json", NAME + "_fine.json"):  # (another test's oracle plugin may have written them: the
        path = os.path.join(base, "static/graphs", fname)   # reference only generates a graph file that is absent)
        if os.path.exists(path):
            os.remove(path)
    plug.register(ctx)                                   # main/context.py:98-112 + the plugin's params-aware nodes
    L = ocfg.num_layers
    want_nodes = {f"{NAME}:{s}" for s in ["embed", "head", "rollout", "transform"] + [f"layer.{i}" for i in range(L)]
                  + [f"layer.{i}.{h}" for i in range(L) for h in ("attn", "mlp")]}
    assert want_nodes <= set(ctx.nodes.keys())
    for fname in (NAME + ".json", NAME + "_fine.json"):  # both graph files land in the reference's static/graphs
        with open(os.path.join(base, "static/graphs", fname)) as f:
            gj = json.load(f)
        assert gj["nodes"][-1]["instance"]["kind"] == "category" and len(gj["edges"]) > L
    assert ctx.get_node(f"{NAME}:layer.0").io({"heads": "1"})["outs"] == ["o", "attn", "cls", "heads"]
    assert ctx.get_node(f"{NAME}:rollout").io({}) == {"ins": [f"a{i}" for i in range(L)], "outs": ["o"]}
    assert NAME in ctx.get_node(f"{NAME}:embed").contents({})
    # the golden request through the reference's own codec and scheduler
    req = rmessage.Request()
    req.decode(open(os.path.join(golden_dir, "wire_tiny.request.bin"), "rb").read())
    ctx.compute(req.graph)
    got = rmessage.Response(req.graph).encode()
    want = open(os.path.join(golden_dir, "wire_tiny.response.bin"), "rb").read()
    assert bytes(got) == want
    assert eng.uploads == {"tokens": 0, "maps": 0, "images": 1}, "tokens and maps stay resident between the nodes of a request"


def _run_request(ctx, image):
    ocfg = O.ORACLE_CONFIGS[NAME]
    nodes, edges, tensors = P.vit_graph_request(NAME, ocfg.num_layers, image)
    req = M.Request()
    req.decode(M.encode_request(nodes, edges, tensors))
    ctx.compute(req.graph)
    return M.decode_response(M.Response(req.graph).encode())


def _context_for(plug):
    ctx = C.Context()
    for n in plug.list_node_names() + plug.fine_node_names():
        C.ModelNode(plug, n).register(ctx)
    return ctx


def test_batched_wire_request_matches_per_image_requests():
    """a [B,3,S,S] tensor through Request.decode -> Context.compute -> Response.encode gives, per image,
    what B single-image requests give (every node output carries the leading batch dimension)."""
    ocfg, module, plug, eng = _plugin()
    ctx = _context_for(plug)
    x = O.synthetic_images(3, ocfg.image_size)
    batched = _run_request(ctx, x)
    L = ocfg.num_layers
    g = ocfg.image_size // ocfg.patch_size
    assert batched[0]["o"].shape == (3, ocfg.tokens, ocfg.hidden_dim)
    assert batched[1]["attn"].shape == (3, ocfg.tokens, ocfg.tokens) and batched[1]["cls"].shape == (3, ocfg.num_heads, g, g)
    assert batched[1 + L]["o"].shape == (3, ocfg.num_classes) and batched[2 + L]["o"].shape == (3, g, g)
    for b in range(3):
        one = _run_request(ctx, x[b])
        for node in one:
            for ch in one[node]:
                torch.testing.assert_close(batched[node][ch][b], one[node][ch], rtol=1e-5, atol=1e-6)


def test_interleaved_requests_of_different_batch_sizes_never_read_reallocated_state():
    """request A (batch 1) has run its layers; request B's embed arrives with a larger batch and makes the
    engine re-allocate its workspace.  A's later nodes (head, rollout, the next layer) must notice -- the workspace
    generation moved -- and upload their inputs again instead of trusting the residency shortcuts."""
    ocfg, module, plug, eng = _plugin()
    L = ocfg.num_layers
    xa = O.synthetic_images(1, ocfg.image_size, seed=1)[0]
    xb = O.synthetic_images(3, ocfg.image_size, seed=2)
    ref_a = O.forward_with_maps(module, xa[None])

    def call(node, **chans):
        p = G.Pinout()
        for k, v in chans.items():
            p.set(k, v)
        return plug.compute(f"{NAME}:{node}", p)

    tok = call("embed", o=xa).get("o")
    maps = []
    for i in range(L):
        out = call(f"layer.{i}", o=tok)
        tok = out.get("o")
        maps.append(out.get("attn"))
    gen0 = eng.workspace_generation()
    tok_b = call("embed", o=xb).get("o")            # grows the workspace: everything of request A is gone on the device
    assert eng.workspace_generation() != gen0 and tok_b.shape[0] == 3
    up0 = dict(eng.uploads)
    roll = call("rollout", **{f"a{i}": m for i, m in enumerate(maps)}).get("o")
    assert eng.uploads["maps"] == up0["maps"] + L, "every map of request A has to be uploaded again"
    logits = call("head", o=tok).get("o")
    assert eng.uploads["tokens"] == up0["tokens"] + 1
    g = ocfg.image_size // ocfg.patch_size
    torch.testing.assert_close(roll, ref_a["rollout"][0].reshape(g, g), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(logits, ref_a["logits"][0], rtol=1e-5, atol=1e-6)
    # request B continues unharmed (its tokens are what the engine holds now? no: A's head re-uploaded A's) and the
    # identity shortcut must not fire for a tensor whose batch differs from what is resident
    out_b = call("layer.0", o=tok_b)
    ref_b = O.forward_with_maps(module, xb)
    torch.testing.assert_close(out_b.get("o"), ref_b["hidden"][0], rtol=1e-5, atol=1e-5)
    # maps of different batch sizes in one rollout call are rejected
    with pytest.raises(Exception, match="expected a"):
        call("rollout", a0=maps[0], **{f"a{i}": out_b.get("attn") for i in range(1, L)})
