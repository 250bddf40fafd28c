#!/usr/bin/env python
"""Benchmark of the hot path: images/sec of a ViT-B/16 224 px forward with attention-map extraction.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model vit_b_16] [--batch 256]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one forward of `batch` synthetic images PER GPU (weak scaling:
images are independent units, sharded with no data-path collective; the only exchange is the gather of logits /
CLS maps / rollout to rank 0, which is inside the timed step).

value      whole-job img/s with the images already resident in HBM, CUDA-event timed on the launch stream.
e2e        the same metric through the public host API (VitEngine.submit_host / wait): pinned-host images in, H2D +
           forward + D2H of logits, per-head CLS maps and rollout inside the timed region.
roofline   the dominant kernel (the tcgen05 GEMM instance with the largest share of the step) timed INSIDE real forwards
           with CUDA events around every launch: achieved TFLOP/s = 2*M*N*K / duration; `frac` is against the measured
           BURST bf16 peak in MEASURED_PEAKS.json (the profiled forwards are a ~60 ms region), the sustained-peak
           fraction is reported beside it.
sustained  a second timed block of >= 3 s (--long-seconds) with its own clock / power samples and per-rank step times,
           so that "power-bound" and "gap-bound" can be told apart from the line alone.
cpu_baseline / --impl reference
           the reference's own path on the host cores: wire request -> Request.decode -> Context.compute over the
           torchvision-CPU plugin (one unbatched fp32 image per request, main/context.py:79-88,143-147) ->
           Response.encode.  /root/reference is not on the GPU box: the scheduler/codec are this repo's mirror of
           it and the arithmetic is torchvision itself (kind = "port").
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# forward, copy-in, copy-out, reader and NCCL streams must not share a hardware queue (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

METRIC = "images/sec ViT-B/16 224px fwd+attn maps"


def _peaks():
    p = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        p.update({k: m[k] for k in ("bf16_tflops", "bf16_tflops_sustained", "hbm_gbs") if k in m})
        p["source"] = "measured"
    except Exception:
        pass
    return p


class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU while the timed region runs."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        s, w = sorted(self.samples), sorted(self.power)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_min_mhz": s[0] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s),
                "power_w": {"median": round(w[len(w) // 2], 1), "max": round(w[-1], 1)} if w else None}


# ------------------------------------------------------------------------------------------ reference arm
class ReferencePath:
    """The reference's path on the host cores: one unbatched fp32 image per request through
    wire decode -> Context.compute (torchvision CPU plugin behind the Model API) -> wire encode."""

    def __init__(self, model_name: str):
        import torch
        from interactive_vit_b200 import context as C, graph as G, message as M
        from oracle import oracle_plugin, vit_oracle as O

        torch.set_num_threads(os.cpu_count() or 1)
        self.cores = torch.get_num_threads()
        self.M, self.plugin_mod, self.name = M, oracle_plugin, model_name
        self.ocfg = O.ORACLE_CONFIGS[model_name]
        Cls = oracle_plugin.make_oracle_model_class(C.Model, G.Pinout)
        plug = Cls(model_name, self.ocfg, O.build_vit(self.ocfg, seed=0))
        self.ctx = C.Context()
        for n in plug.list_node_names():
            C.ModelNode(plug, n).register(self.ctx)
        self.imgs = O.synthetic_images(4, self.ocfg.image_size)
        self.one(0)  # warm-up (thread pool, allocator)

    def one(self, i: int) -> bytes:
        nodes, edges, tensors = self.plugin_mod.vit_graph_request(self.name, self.ocfg.num_layers, self.imgs[i % 4])
        req = self.M.Request()
        req.decode(self.M.encode_request(nodes, edges, tensors))
        self.ctx.compute(req.graph)
        return self.M.Response(req.graph).encode()

    def batched(self, batch: int = 32, reps: int = 2):
        """SURVEY.md section 8d's second CPU figure: the same torchvision model called on a whole batch (forward + maps +
        rollout, no wire), which the reference cannot do through its UI but a fair CPU comparison should show."""
        import torch
        from oracle import vit_oracle as O

        model = O.build_vit(self.ocfg, seed=0)
        x = O.synthetic_images(batch, self.ocfg.image_size)
        O.forward_with_maps(model, x[:2])
        t0 = time.perf_counter()
        for _ in range(reps):
            O.forward_with_maps(model, x)
        dt = time.perf_counter() - t0
        return {"value": batch * reps / dt, "unit": "img/s", "batch": batch,
                "what": "torchvision CPU fp32 forward + maps + rollout on one batch, no wire codec"}

    def run(self, seconds: float, max_images: int):
        """(images done, seconds): stops at max_images or once `seconds` have elapsed."""
        t0 = time.perf_counter()
        done = 0
        while done < max_images and (time.perf_counter() - t0 < seconds or done == 0):
            self.one(done)
            done += 1
        return done, time.perf_counter() - t0


def _workload(model: str, batch: int) -> str:
    import interactive_vit_b200.engine as E

    cfg = E.CONFIGS[model]
    return (f"{model} {cfg.image_size}px batch {batch} per GPU, forward + head-averaged maps + per-head CLS maps + rollout "
            f"for all {cfg.num_layers} layers; random-init weights")


def run_reference(args, rank: int):
    if rank != 0:
        return
    ref = ReferencePath(args.model)
    per_step = 4
    for _ in range(args.warmup):
        ref.run(1e9, 1)
    steps = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        steps.append(ref.run(1e9, per_step))
        if time.perf_counter() - t_all > 150:
            break
    imgs = sum(d for d, _ in steps)
    secs = sum(t for _, t in steps)
    value = imgs / secs
    sample = f"{imgs} single-image requests ({len(steps)} steps x {per_step}) through decode -> compute -> encode, fp32"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": args.gpus, "steps": len(steps),
        "warmup": args.warmup, "ms_per_step": secs / len(steps) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the product arm's line (its `config.workload`); each reference step is a bounded sample of
        # it, run the only way the reference's path runs: one unbatched image per request on the host cores
        "config": {"workload": _workload(args.model, args.batch), "global_batch": args.batch * max(1, args.gpus),
                   "parallelism": "cpu", "reference_sample": f"{per_step} unbatched single-image requests per step through the "
                   "reference path (wire decode -> scheduler -> torchvision CPU fp32 plugin with attention maps -> wire encode)"},
        "cpu_baseline": {"value": value, "unit": "img/s", "cores": ref.cores, "kind": "port", "sample": sample,
                         "batched": ref.batched()},
        "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def plugin_request_latency(model_name, cfg, eng, P, n=20):
    """ms per single-image request through decode -> scheduler -> B200 plugin nodes -> encode (median of n)."""
    import tempfile

    import torch
    from interactive_vit_b200 import context as C, message as M

    plug = P.VitB200Model(model_name, cfg, P.build_torchvision_vit(cfg, seed=0), engine=eng)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "static", "graphs"))
        C.set_base_dir(d)
        try:
            ctx = C.Context()
            plug.register(ctx)
        finally:
            C.set_base_dir(None)
    img = torch.rand(3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1234))
    nodes, edges, tensors = P.vit_graph_request(model_name, cfg.num_layers, img)
    body = M.encode_request(nodes, edges, tensors)

    def one():
        req = M.Request()
        req.decode(body)
        ctx.compute(req.graph)
        return M.Response(req.graph).encode()

    for _ in range(3):
        one()
    times = []
    for _ in range(n):
        t0 = time.perf_counter()
        resp = one()
        times.append((time.perf_counter() - t0) * 1e3)
    times.sort()
    return {"ms_per_request": times[len(times) // 2], "img_per_s": 1e3 / times[len(times) // 2], "requests": n,
            "response_bytes": len(resp),
            "path": "wire request -> Request.decode -> Context.compute (embed, layer.0.., head, rollout nodes on the GPU) -> "
                    "Response.encode; every node output returned as CPU fp32 like the reference's (pinned, deferred: one wait for "
                    "the device per request, single-copy encode)"}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank: int, local_rank: int, world: int):
    import math

    import torch
    import torch.distributed as dist
    import interactive_vit_b200.engine as E
    from interactive_vit_b200 import dist as D, vit_plugin as P

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        # NCCL announces its version on stdout when the communicator comes up; stdout carries exactly one JSON line,
        # so the file descriptor points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    cfg = E.CONFIGS[args.model]
    B = args.batch
    device = torch.device("cuda", local_rank)
    eng = E.VitEngine(cfg, local_rank, B)
    eng.load_state_dict(P.build_torchvision_vit(cfg, seed=0).state_dict())
    flags = E.EMIT_AVG | E.EMIT_CLS | E.EMIT_ROLLOUT

    def rank_images(r):
        return torch.rand(B, 3, cfg.image_size, cfg.image_size, generator=torch.Generator().manual_seed(1234 + r))

    host_images = rank_images(rank).pin_memory()
    images = host_images.cuda(non_blocking=True)
    stream = torch.cuda.Stream()
    L, H, N = cfg.num_layers, cfg.num_heads, cfg.tokens
    total = B * world

    # N > 1, the result exchange (logits + per-head CLS maps + rollout of every rank to rank 0), inside the timed step:
    #   push (default): NO collective and no barrier -- the producing kernels store into rank 0's receive set over NVLink
    #     peer memory (dist.PeerPush, vitb200_bind_outputs); per-rank completion flags order rank 0's reader behind them
    #   nccl: one packed gather per step on a side stream, overlapping the next forward (dist.PackedGather)
    spec = {"logits": ((cfg.num_classes,), 0), "cls_maps": ((L, H, N), 1), "rollout": ((N - 1,), 0)}
    top1 = [None]

    def consume(s, views):
        # rank 0 USES the gathered set on its reader stream: top-1 of every image, a checksum of the maps
        top1[0] = (views["logits"].argmax(-1), views["cls_maps"].sum(), views["rollout"].sum())

    gatherer, pusher, gather_how = None, None, "none"
    if world > 1:
        if args.gather in ("auto", "push"):
            try:
                pusher = D.PeerPush(eng, total, device, stream=stream.cuda_stream, consumer=consume)
            except Exception as ex:      # raised on every rank alike (PeerPush agrees before it raises)
                print(f"[bench] rank {rank}: peer-memory push unavailable ({type(ex).__name__}: {str(ex)[:200]})",
                      file=sys.stderr, flush=True)
                if args.gather == "push":
                    raise
                pusher = None
        if pusher is not None:
            gather_how = ("logits + CLS maps + rollout to rank 0 with NO collective and no barrier: the producing kernels store "
                          "into rank 0's receive set over NVLink peer memory (CUDA IPC mapping), per-rank completion flags, "
                          "three sets in rotation; rank 0's reader stream consumes every set (top-1 + checksums)")
        else:
            gatherer = D.PackedGather(spec, total, device)
            gather_how = ("logits + CLS maps + rollout to rank 0: one packed NCCL gather per step on a side stream, "
                          "overlapping the next forward")

    def step():
        if pusher is not None:
            pusher.begin()
            eng.forward_device(images, flags, stream.cuda_stream)
            pusher.end()
            return
        eng.forward_device(images, flags, stream.cuda_stream)
        if gatherer is not None:
            gatherer.submit({"logits": eng.device_output(0, (B, cfg.num_classes)),
                             "cls_maps": eng.device_output(E.EMIT_CLS, (L, B, H, N)),
                             "rollout": eng.device_output(E.EMIT_ROLLOUT, (B, N - 1))})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_block(nsteps):
        """nsteps steps between a barrier + synchronize on both sides; CUDA events on the launch stream.  Returns
        (ms per step incl. the exchange, ms per step of this rank's OWN forwards, clock / power summary)."""
        barrier()
        ev0, ev_own, ev1 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        with ClockSampler(local_rank) as clocks:
            ev0.record(stream)
            for _ in range(nsteps):
                step()
            ev_own.record(stream)            # this rank's forwards (and its completion flags) are done here ...
            if gatherer is not None:
                gatherer.finish()
            if pusher is not None:
                pusher.finish()              # ... rank 0 also waits for its reader: every rank's results consumed
            ev1.record(stream)
            barrier()
        return ev0.elapsed_time(ev1) / nsteps, ev0.elapsed_time(ev_own) / nsteps, clocks.summary()

    def all_ranks(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
        launches0, replays0 = eng.launch_count(), eng.graph_replays()
        ms_local, own_local, clk = timed_block(args.steps)
        launches = eng.launch_count() - launches0
        replays = eng.graph_replays() - replays0
        per_rank_ms, per_rank_own = all_ranks(ms_local), all_ranks(own_local)
        ms = max(per_rank_ms)

        # ---- per-kernel times inside real forwards (CUDA events in front of every launch on the launch stream; the
        # figures are event-to-event, so each includes the gap to the next launch), and the dominant kernel's roofline.
        roof, kernels = None, None
        if rank == 0:
            peaks = _peaks()
            runs = [eng.profile_forward(images, flags) for _ in range(5)]
            kernels = {}
            for k in runs[0]:
                n = runs[0][k][0]
                msk = sorted(r[k][1] for r in runs)[len(runs) // 2]
                kernels[k] = {"launches": n, "ms": round(msk, 4)} if k != "total" else {"ms": round(msk, 4)}
            M_ = B * N
            gemms = {"gemm_qkv": (3 * cfg.hidden_dim, cfg.hidden_dim), "gemm_out_proj": (cfg.hidden_dim, cfg.hidden_dim),
                     "gemm_fc1_gelu": (cfg.mlp_dim, cfg.hidden_dim), "gemm_fc2": (cfg.hidden_dim, cfg.mlp_dim)}
            for k, (n_, k_) in gemms.items():
                kernels[k]["tflops"] = round(2.0 * M_ * n_ * k_ * kernels[k]["launches"] / (kernels[k]["ms"] * 1e-3) / 1e12, 1)
            dom = max(gemms, key=lambda k: kernels[k]["ms"])
            n_, k_ = gemms[dom]
            kms = kernels[dom]["ms"] / kernels[dom]["launches"]
            achieved = 2.0 * M_ * n_ * k_ / (kms * 1e-3) / 1e12
            traffic = None
            try:
                with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                    traffic = json.load(f).get(args.model, {}).get(dom)
            except Exception:
                pass
            # the profiled forwards are a ~60 ms region: the BURST peak is the denominator; the
            # sustained-peak fraction is beside it
            roof = {"bound": "tensor", "kernel": f"gemm_bf16_kernel ({dom}) M={M_} N={n_} K={k_}", "achieved": achieved,
                    "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                    "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"], "traffic": traffic,
                    "peak_source": peaks["source"] + " burst (cuBLAS bf16, best of 10)", "kernel_ms": kms,
                    "timing": "CUDA events on the launch stream around every launch of 5 profiled forwards, median"}

        # ---- a second, LONG block (>= --long-seconds): sustained clocks / power, per-rank step times
        sustained = None
        if args.long_seconds > 0:
            nlong = max(args.steps, int(math.ceil(args.long_seconds * 1e3 / ms)))
            lms, lown, lclk = timed_block(nlong)
            l_ms, l_own = all_ranks(lms), all_ranks(lown)
            sustained = {"steps": nlong, "seconds": round(max(l_ms) * nlong / 1e3, 2), "ms_per_step": max(l_ms),
                         "value": total / max(l_ms) * 1e3, "unit": "img/s",
                         "per_rank_ms_per_step": [round(v, 4) for v in l_ms],
                         "per_rank_own_forward_ms_per_step": [round(v, 4) for v in l_own], "clocks_rank0": lclk}

        push_check = None
        if pusher is not None:
            # EVERY rank's slice of what was pushed to rank 0, against rank 0's own forward on that rank's images
            # (seed 1234 + r; the forward is batch-invariant per image, so the bits must agree) -- outside the timed region
            if rank == 0:
                got = {k: pusher.result(k).clone() for k in ("logits", "cls_maps", "rollout")}
            pusher.close()
            if rank == 0:
                push_check = []
                for r in range(world):
                    xr = images if r == 0 else rank_images(r).cuda()
                    eng.forward_device(xr, flags, stream.cuda_stream)
                    torch.cuda.synchronize()
                    sl = slice(r * B, (r + 1) * B)
                    ok = bool(torch.equal(got["logits"][sl], eng.device_output(0, (B, cfg.num_classes))) and
                              torch.equal(got["cls_maps"][:, sl], eng.device_output(E.EMIT_CLS, (L, B, H, N))) and
                              torch.equal(got["rollout"][sl], eng.device_output(E.EMIT_ROLLOUT, (B, N - 1))))
                    push_check.append(ok)
                if not all(push_check):
                    raise RuntimeError(f"peer-memory push: slices {[r for r, ok in enumerate(push_check) if not ok]} of rank "
                                       "0's receive set do not match rank 0's own forward on those ranks' images")
                gather_how += f"; all {world} ranks' slices verified bit-identical after the run"
            dist.barrier()

        # ---- e2e: public host API (VitEngine.submit_host / wait), pinned host buffers.  Every step copies its images
        # host -> device and its results device -> host inside the timed region; two requests are in flight, so the
        # copies of neighbouring steps overlap the forward (separate copy streams).  N > 1: every rank feeds its GPU from
        # its own pinned host images; the results are bound to rank 0's receive set (the same push as above, on the
        # engine's own stream) and rank 0's reader copies the WHOLE gathered set to rank 0's pinned host memory.
        e2e_flags = E.EMIT_CLS | E.EMIT_ROLLOUT
        e2e_push, e2e_gatherer, e2e_how = None, None, "VitEngine.submit_host / wait: 2 requests in flight, copies overlap the forward"
        if world > 1:
            host_sets = {}

            def to_host(s, views):
                for k, v in views.items():
                    host_sets[s][k].copy_(v, non_blocking=True)

            if pusher is not None:      # the transport that worked for the device-resident arm
                e2e_push = D.PeerPush(eng, total, device, stream=None, consumer=to_host)
                if rank == 0:       # pinned destinations of the gathered results, one per receive set
                    shapes = {"logits": (total, cfg.num_classes), "cls_maps": (L, total, H, N), "rollout": (total, N - 1)}
                    for s_ in range(e2e_push.sets):
                        host_sets[s_] = {k: torch.empty(shp).pin_memory() for k, shp in shapes.items()}
                e2e_how += ("; results bound to rank 0's receive set (no collective), rank 0 copies the gathered set "
                            "device -> host from its reader stream")
            else:
                e2e_gatherer = D.PackedGather(spec, total, device)
                e2e_how += "; staged results gathered to rank 0 with one packed NCCL gather per step"
        outs = [{} if e2e_push is not None else
                {"logits": torch.empty(B, cfg.num_classes).pin_memory(), "cls_maps": torch.empty(L, B, H, N).pin_memory(),
                 "rollout": torch.empty(B, N - 1).pin_memory()} for _ in range(2)]

        def e2e_finish(ticket):
            eng.wait(ticket)
            if e2e_gatherer is not None:
                e2e_gatherer.submit({"logits": eng.staged_output(ticket, 0, (B, cfg.num_classes)),
                                     "cls_maps": eng.staged_output(ticket, E.EMIT_CLS, (L, B, H, N)),
                                     "rollout": eng.staged_output(ticket, E.EMIT_ROLLOUT, (B, N - 1))})

        def e2e_run(n):
            pending = []
            for i in range(n):
                if e2e_push is not None:
                    e2e_push.begin()
                pending.append(eng.submit_host(host_images, e2e_flags, outs[i & 1]))
                if e2e_push is not None:
                    e2e_push.end()
                if len(pending) == 2:
                    e2e_finish(pending.pop(0))
            while pending:
                e2e_finish(pending.pop(0))
            if e2e_gatherer is not None:
                e2e_gatherer.finish()
            if e2e_push is not None:
                e2e_push.finish()
            torch.cuda.synchronize()

        e2e_run(max(2, args.warmup // 2))
        barrier()
        t0 = time.perf_counter()
        e2e_run(args.steps)
        barrier()
        e2e_ms = (time.perf_counter() - t0) / args.steps * 1e3
        if e2e_push is not None:
            e2e_push.close()

        # ---- BASELINE config 1 beside it: ONE unbatched image through the reference-facing path (wire request ->
        # Request.decode -> Context.compute over the B200 plugin's nodes -> Response.encode), i.e. what the reference's
        # UI would wait for; same request bytes as the CPU arm below
        interactive = None
        if rank == 0 and not args.no_cpu_baseline:
            try:
                interactive = plugin_request_latency(args.model, cfg, eng, P)
                st1 = torch.cuda.Stream()
                x1 = images[:1].contiguous()
                for _ in range(5):
                    eng.forward_device(x1, flags, st1.cuda_stream)
                st1.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st1)
                for _ in range(50):
                    eng.forward_device(x1, flags, st1.cuda_stream)
                b.record(st1)
                st1.synchronize()
                interactive["forward_ms_batch1"] = round(a.elapsed_time(b) / 50, 4)
            except Exception as ex:  # reported, never fatal for the headline
                interactive = {"error": str(ex)[:200]}

    e2e_ms = max(all_ranks(e2e_ms))

    if rank == 0:
        peaks = _peaks()
        value = total / ms * 1e3
        flop = cfg.gflop_per_image()
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = ReferencePath(args.model)
            done, dt = ref.run(12.0, 400)
            cpu = {"value": done / dt, "unit": "img/s", "cores": ref.cores, "kind": "port",
                   "sample": f"{done} single-image requests in {dt:.1f}s through decode -> compute -> encode (torchvision CPU fp32)",
                   "batched": ref.batched()}
        res_bytes = (cfg.num_classes + L * H * N + (N - 1)) * 4
        line = {
            "metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": _workload(args.model, B), "global_batch": total,
                       "l2": f"inputs larger than L2 ({B * 3 * cfg.image_size ** 2 * 4 >> 20} MiB images, "
                             f"{B * N * cfg.hidden_dim * 4 >> 20} MiB token stream per step)",
                       "parallelism": f"dp{world}", "gather": gather_how},
            "clocks": clk,
            "per_rank": {"ms_per_step": [round(v, 4) for v in per_rank_ms],
                         "own_forward_ms_per_step": [round(v, 4) for v in per_rank_own],
                         "what": "CUDA events per rank: whole step incl. the exchange (rank 0: until its reader has consumed "
                                 "every set) / the rank's own forwards only"},
            "sustained": sustained,
            "e2e": {"value": total / e2e_ms * 1e3, "unit": "img/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": world * B * 3 * cfg.image_size ** 2 * 4,
                    "d2h_bytes_per_step": total * res_bytes,
                    "outputs": "logits, per-head CLS maps (all layers), rollout" + (" -- of all ranks, on rank 0's host" if world > 1 else ""),
                    "api": e2e_how},
            "gpu_launches": launches,
            "graph_replays": replays,
            "roofline": roof,
            "kernels": kernels,
            "step_tensor": {"achieved": value * flop / 1e3, "unit": "TFLOP/s", "gflop_per_image": flop,
                            "frac_of_burst_peak": value * flop / 1e3 / peaks["bf16_tflops"] / world,
                            "frac_of_sustained_peak": value * flop / 1e3 / peaks["bf16_tflops_sustained"] / world},
            "cpu_baseline": cpu,
            "interactive": interactive,
        }
        if sustained is not None:
            sustained["frac_of_burst_peak"] = sustained["value"] * flop / 1e3 / peaks["bf16_tflops"] / world
            sustained["frac_of_sustained_peak"] = sustained["value"] * flop / 1e3 / peaks["bf16_tflops_sustained"] / world
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vit_b_16")
    ap.add_argument("--batch", type=int, default=256, help="images per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--long-seconds", type=float, default=3.0,
                    help="length of the second timed block (sustained clocks / power, per-rank times); 0 skips it")
    ap.add_argument("--gather", default="auto", choices=["auto", "push", "nccl"],
                    help="N > 1 result exchange of the device-resident arm: peer-memory push (no collective) or packed NCCL gather")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    args.warmup = max(args.warmup, 3)
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
