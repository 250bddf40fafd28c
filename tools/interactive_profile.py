"""Where a single-image request through the plugin path spends its time (decode / per-node compute / encode)."""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import interactive_vit_b200.engine as E
from interactive_vit_b200 import context as C, message as M, vit_plugin as P

cfg = E.CONFIGS["vit_b_16"]
plug = P.VitB200Model("vit_b_16", cfg, P.build_torchvision_vit(cfg, seed=0), 0, 1)
with tempfile.TemporaryDirectory() as d:
    os.makedirs(os.path.join(d, "static", "graphs"))
    C.set_base_dir(d)
    ctx = C.Context()
    plug.register(ctx)
    C.set_base_dir(None)
img = torch.rand(3, 224, 224)
body = M.encode_request(*P.vit_graph_request("vit_b_16", cfg.num_layers, img))
acc = {"decode": 0.0, "compute": 0.0, "encode": 0.0}
per_node = {}
orig = {}
for name, node in ctx.nodes.items():
    def wrap(fn, name=name):
        def f(params, inputs):
            t = time.perf_counter()
            r = fn(params, inputs)
            k = name.split(":")[1].split(".")[0]
            per_node[k] = per_node.get(k, 0.0) + time.perf_counter() - t
            return r
        return f
    node.compute = wrap(node.compute)
N = 30
WARM = 12     # past the first cyclic-GC passes: finished requests' pinned blocks are back in torch's cache
totals = []
for it in range(N + WARM):
    if it == WARM:
        acc = {k: 0.0 for k in acc}; per_node = {}; totals = []
    t0 = time.perf_counter(); req = M.Request(); req.decode(body)
    t1 = time.perf_counter(); ctx.compute(req.graph)
    t2 = time.perf_counter(); out = M.Response(req.graph).encode()
    t3 = time.perf_counter()
    acc["decode"] += t1 - t0; acc["compute"] += t2 - t1; acc["encode"] += t3 - t2
    totals.append((t3 - t0) * 1e3)
print({k: round(v / N * 1e3, 3) for k, v in acc.items()}, "ms per request;", len(out), "response bytes")
print({k: round(v / N * 1e3, 3) for k, v in per_node.items()}, "ms per request by node kind")
totals.sort()
print("request total ms: median %.3f  min %.3f  max %.3f  (deferred=%s)" % (totals[len(totals) // 2], totals[0], totals[-1],
                                                                              os.environ.get("VITB200_DEFERRED", "1")))
