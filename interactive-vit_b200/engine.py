"""ctypes binding of libvitb200.so (C ABI: include/vitb200.h) — the only way Python reaches the CUDA kernels.

The reference has no FFI (it calls torch modules directly inside ``Model.compute``, main/context.py:79-88);
this module is the stub a maintainer would add.  There is deliberately no fallback: if the shared library is
missing or no sm_100 device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Dict, Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VITB200_LIB", os.path.join(_HERE, "libvitb200.so"))  # override: tracing builds (tools/)

EMIT_AVG = 1
EMIT_CLS = 2
EMIT_ROLLOUT = 4
EMIT_HEADS = 8
EMIT_HIDDEN = 16


class EngineError(RuntimeError):
    """Raised for every non-zero status of the C ABI; ``str(e)`` is vitb200_last_error()."""


class _Config(C.Structure):
    _fields_ = [
        ("image_size", C.c_int),
        ("patch_size", C.c_int),
        ("num_layers", C.c_int),
        ("num_heads", C.c_int),
        ("hidden_dim", C.c_int),
        ("mlp_dim", C.c_int),
        ("num_classes", C.c_int),
        ("max_batch", C.c_int),
        ("device", C.c_int),
        ("precision", C.c_int),
    ]


class _HostOutputs(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p),
        ("avg_maps", C.c_void_p),
        ("cls_maps", C.c_void_p),
        ("rollout", C.c_void_p),
        ("heads", C.c_void_p),
        ("hidden", C.c_void_p),
    ]


_lib: Optional[C.CDLL] = None

# name -> (restype, argtypes); kept in one table so the symbol test can walk it against include/vitb200.h
_P, _I, _U32, _F, _L = C.c_void_p, C.c_int, C.c_uint32, C.c_float, C.c_long
SIGNATURES = {
    "vitb200_last_error": (C.c_char_p, []),
    "vitb200_version": (_I, []),
    "vitb200_create": (_I, [C.POINTER(_Config), C.POINTER(_P)]),
    "vitb200_destroy": (None, [_P]),
    "vitb200_load_weight": (_I, [_P, C.c_char_p, _P, C.c_size_t]),
    "vitb200_weights_ready": (_I, [_P]),
    "vitb200_forward_host": (_I, [_P, _P, _I, _U32, C.POINTER(_HostOutputs)]),
    "vitb200_forward_device": (_I, [_P, _P, _I, _U32, _P]),
    "vitb200_engine_stream": (_I, [_P, C.POINTER(_P)]),
    "vitb200_workspace_generation": (C.c_uint64, [_P]),
    "vitb200_reserve": (_I, [_P, _I, _U32]),
    "vitb200_set_graphs": (_I, [_P, _I]),
    "vitb200_graph_replays": (C.c_uint64, [_P]),
    "vitb200_bind_outputs": (_I, [_P, _P, _P, _L, _P]),
    "vitb200_peer_alloc": (_I, [_I, C.c_size_t, C.POINTER(_P), _P]),
    "vitb200_peer_open": (_I, [_I, _P, C.POINTER(_P)]),
    "vitb200_peer_close": (_I, [_I, _P]),
    "vitb200_peer_free": (_I, [_I, _P]),
    "vitb200_flag_signal": (_I, [_P, _U32, _P]),
    "vitb200_flag_wait": (_I, [_P, _U32, _P]),
    "vitb200_submit_host": (_I, [_P, _P, _I, _U32, C.POINTER(_HostOutputs), C.POINTER(C.c_uint64)]),
    "vitb200_wait": (_I, [_P, C.c_uint64]),
    "vitb200_staged_output": (_I, [_P, C.c_uint64, _U32, C.POINTER(_P)]),
    "vitb200_profile_forward": (_I, [_P, _P, _I, _U32, C.c_char_p, C.c_size_t]),
    "vitb200_device_output": (_I, [_P, _U32, C.POINTER(_P), C.POINTER(_I)]),
    "vitb200_synchronize": (_I, [_P]),
    "vitb200_set_deferred": (_I, [_P, _I]),
    "vitb200_stage_embed": (_I, [_P, _P, _I]),
    "vitb200_stage_transform": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "vitb200_stage_embed_resident": (_I, [_P, _I]),
    "vitb200_stage_layer": (_I, [_P, _I, _I, _U32]),
    "vitb200_stage_layer_fetch": (_I, [_P, _I, _I, _U32, _I, _P, _P, _P]),
    "vitb200_stage_attn_block": (_I, [_P, _I, _I, _U32]),
    "vitb200_stage_mlp_block": (_I, [_P, _I, _I]),
    "vitb200_stage_head": (_I, [_P, _I, _P]),
    "vitb200_stage_rollout": (_I, [_P, _I, _P]),
    "vitb200_set_tokens": (_I, [_P, _P, _I]),
    "vitb200_get_tokens": (_I, [_P, _P, _I]),
    "vitb200_set_avg_map": (_I, [_P, _I, _P, _I]),
    "vitb200_get_avg_map": (_I, [_P, _I, _P, _I]),
    "vitb200_get_cls_map": (_I, [_P, _I, _P, _I]),
    "vitb200_get_head_map": (_I, [_P, _I, _P, _I]),
    "vitb200_get_cls_grid": (_I, [_P, _I, _P, _I]),
    "vitb200_launch_count": (C.c_uint64, [_P]),
    "vitb200_op_gemm": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vitb200_op_gemm_ex": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _F, _P]),
    "vitb200_op_fold_ln": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "vitb200_op_split_bf16": (_I, [_P, _P, _P, C.c_size_t, _P]),
    "vitb200_op_gemm_split": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P]),
    "vitb200_op_layernorm": (_I, [_P, _P, _P, _P, _I, _I, _F, _P]),
    "vitb200_op_attention": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "vitb200_op_attention_ex": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "vitb200_op_preprocess": (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    "vitb200_op_patchify": (_I, [_P, _P, _I, _I, _I, _P]),
    "vitb200_op_rollout": (_I, [_P, _L, _I, _I, _I, _I, _P, _P]),
}


def load_library() -> C.CDLL:
    """dlopen libvitb200.so (built in-tree by __graft_entry__.build()) and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)"
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load_library().vitb200_last_error()
        raise EngineError(f"vitb200 error {status}: {msg.decode() if msg else '?'}")


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass(frozen=True)
class VitConfig:
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

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    def gflop_per_image(self) -> float:
        """Tensor-core FLOPs of one forward (SURVEY.md §8d formula; GEMMs + attention matmuls only)."""
        n, N, d, m, L = self.grid ** 2, self.tokens, self.hidden_dim, self.mlp_dim, self.num_layers
        f = 2 * n * (3 * self.patch_size ** 2) * d
        f += L * (2 * N * d * 3 * d + 4 * N * N * d + 2 * N * d * d + 4 * N * d * m)
        f += 2 * d * self.num_classes
        return f / 1e9


CONFIGS: Dict[str, VitConfig] = {
    "vit_s_16": VitConfig(224, 16, 12, 6, 384, 1536),
    "vit_b_16": VitConfig(224, 16, 12, 12, 768, 3072),
    "vit_l_16": VitConfig(224, 16, 24, 16, 1024, 4096),
    "vit_b_16_384": VitConfig(384, 16, 12, 12, 768, 3072),    # 577 tokens
    "vit_h_16_384": VitConfig(384, 16, 32, 16, 1280, 5120),   # ViT-H width/depth, head dim 80, 577 tokens
}


def _pending_args(args):
    for a in args:
        if isinstance(a, PendingTensor):
            yield a
        elif isinstance(a, (list, tuple)):
            yield from _pending_args(a)


class PendingTensor(torch.Tensor):
    """CPU fp32 tensor over pinned host memory whose contents are still being written by a device-to-host copy queued
    on the engine's stream (deferred mode, `VitEngine.set_deferred`).  It IS a `torch.Tensor` for the reference's
    graph code (`Edge.tensor`, main/graph.py:53; shipped by `t.numpy().tobytes()`, main/message.py:111-121): the first
    operation that can see the data -- `numpy()`, `data_ptr()`, any torch function -- waits for the engine's stream
    once (one wait per request instead of four per node); metadata (`shape`, `dim()`, `dtype`, `device` ...) does not.
    Results of operations are plain tensors."""

    _NO_WAIT = None

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if cls._NO_WAIT is None:
            T = torch.Tensor
            cls._NO_WAIT = {T.dim, T.size, T.numel, T.stride, T.is_contiguous, T.element_size, T.shape.__get__,
                            T.ndim.__get__, T.dtype.__get__, T.device.__get__, T.is_cuda.__get__, T.layout.__get__,
                            T.requires_grad.__get__, T.is_pinned, T.__len__}
        if func not in cls._NO_WAIT:
            for a in _pending_args(args):
                a.wait()
            for a in _pending_args(tuple(kwargs.values())):
                a.wait()
        # the idiom of PyTorch's "Extending torch" notes for a subclass that returns plain tensors (torch >= 2.0; the
        # reference pins 2.9.1); older builds only have the broader DisableTorchFunction
        guard = getattr(torch._C, "DisableTorchFunctionSubclass", None) or torch._C.DisableTorchFunction
        with guard():
            return func(*args, **kwargs)

    def wait(self) -> None:
        eng = getattr(self, "_engine", None)
        if eng is not None:
            eng._drain(self._seq)
            self._engine = None


class VitEngine:
    """One engine = one model replica on one GPU.  Thread-safe (the C side serialises calls)."""

    def __init__(self, cfg: VitConfig, device: int = 0, max_batch: int = 1, precision: str = "bf16"):
        """precision: "bf16" (bf16 operands, fp32 accumulation) or "fp32x3" (split-bf16 operands: hi*hi + lo*hi +
        hi*lo, <= 1e-3 of the fp32 reference, ~3x the tensor work; attention of head dims other than 64 in fp32 on the
        CUDA cores)."""
        self.lib = load_library()
        self.cfg = cfg
        self.device = device
        self.precision = precision
        c = _Config(cfg.image_size, cfg.patch_size, cfg.num_layers, cfg.num_heads, cfg.hidden_dim, cfg.mlp_dim,
                    cfg.num_classes, max_batch, device, {"bf16": 0, "fp32x3": 1}[precision])
        h = C.c_void_p()
        check(self.lib.vitb200_create(C.byref(c), C.byref(h)))
        self._h = h
        # deferred mode (set_deferred): host outputs handed out as PendingTensor; `_issued` counts them, `_drained` is the
        # highest count known to have landed, `_keep` holds their pinned storage allocated until then
        self._deferred = False
        self._issued = 0
        self._drained = 0
        self._keep = []
        self._slab: Optional[torch.Tensor] = None    # current request's pinned slab (uint8) and its fill level
        self._slab_f32: Optional[torch.Tensor] = None
        self._slab_off = 0
        self._book = threading.Lock()    # guards the three fields above (requests may be encoded on other threads)
        self._wire_pending: Dict[int, tuple] = {}    # data_ptr -> (slab, offset) between _host_out and _issue

    def close(self) -> None:
        if getattr(self, "_h", None):
            if getattr(self, "_deferred", False):
                try:       # outstanding PendingTensors become plain, valid CPU tensors before the engine goes away
                    self._drain(self._issued)
                except Exception:
                    pass
            self.lib.vitb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights ---------------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        """torchvision VisionTransformer state-dict (fp32) -> device (bf16 GEMM operands, fp32 the rest)."""
        for name, t in sd.items():
            t = t.detach().to(torch.float32).contiguous().cpu()
            check(self.lib.vitb200_load_weight(self._h, name.encode(), t.data_ptr(), t.numel()))
        check(self.lib.vitb200_weights_ready(self._h))

    # ---- whole forward ---------------------------------------------------------------------------
    def forward_host(self, images: torch.Tensor, flags: int = 0, out: Optional[Dict[str, torch.Tensor]] = None
                     ) -> Dict[str, torch.Tensor]:
        """images: CPU fp32 [B,3,S,S] (pinned for async copies).  Returns CPU fp32 tensors."""
        assert images.device.type == "cpu" and images.dtype == torch.float32 and images.is_contiguous()
        cfg = self.cfg
        B, N, L, H, d = images.shape[0], cfg.tokens, cfg.num_layers, cfg.num_heads, cfg.hidden_dim
        shapes = {"logits": (B, cfg.num_classes)}
        if flags & EMIT_AVG:
            shapes["avg_maps"] = (L, B, N, N)
        if flags & EMIT_CLS:
            shapes["cls_maps"] = (L, B, H, N)
        if flags & EMIT_ROLLOUT:
            shapes["rollout"] = (B, N - 1)
        if flags & EMIT_HEADS:
            shapes["heads"] = (L, B, H, N, N)
        if flags & EMIT_HIDDEN:
            shapes["hidden"] = (L, B, N, d)
        res = {}
        for k, shp in shapes.items():
            if out is not None and k in out:
                assert tuple(out[k].shape) == shp and out[k].dtype == torch.float32 and out[k].is_contiguous()
                res[k] = out[k]
            else:
                res[k] = torch.empty(shp, dtype=torch.float32)
        ho = _HostOutputs(*[_ptr(res.get(k)) for k in ("logits", "avg_maps", "cls_maps", "rollout", "heads", "hidden")])
        check(self.lib.vitb200_forward_host(self._h, images.data_ptr(), B, flags, C.byref(ho)))
        return res

    def submit_host(self, images: torch.Tensor, flags: int, out: Dict[str, torch.Tensor]) -> int:
        """Pipelined forward_host: enqueue H2D + forward + D2H into the caller's (pinned) `out` tensors and return a
        ticket; up to two requests are in flight.  `out` keys: logits, cls_maps, rollout, avg_maps."""
        assert images.device.type == "cpu" and images.dtype == torch.float32 and images.is_contiguous()
        cfg = self.cfg
        B, N, L, H = images.shape[0], cfg.tokens, cfg.num_layers, cfg.num_heads
        shapes = {"logits": (B, cfg.num_classes), "avg_maps": (L, B, N, N), "cls_maps": (L, B, H, N), "rollout": (B, N - 1)}
        for k, t in out.items():
            assert tuple(t.shape) == shapes[k] and t.dtype == torch.float32 and t.is_contiguous() and t.device.type == "cpu", k
        ho = _HostOutputs(_ptr(out.get("logits")), _ptr(out.get("avg_maps")), _ptr(out.get("cls_maps")), _ptr(out.get("rollout")),
                          None, None)
        ticket = C.c_uint64()
        check(self.lib.vitb200_submit_host(self._h, images.data_ptr(), B, flags, C.byref(ho), C.byref(ticket)))
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[ticket.value] = (images, out)   # keep the host buffers alive until wait()
        return ticket.value

    def wait(self, ticket: int) -> None:
        check(self.lib.vitb200_wait(self._h, ticket))
        getattr(self, "_inflight", {}).pop(ticket, None)

    def staged_output(self, ticket: int, which: int, shape) -> torch.Tensor:
        """Device view of one staged output of an in-flight / just-completed ticket (dense layout)."""
        p = C.c_void_p()
        check(self.lib.vitb200_staged_output(self._h, ticket, which, C.byref(p)))
        return _device_view(p.value, shape, self.device)

    def forward_device(self, images: torch.Tensor, flags: int = 0, stream: Optional[int] = None) -> None:
        """images: CUDA fp32 [B,3,S,S] on this engine's device; enqueues on `stream`, a raw cudaStream_t taken literally
        (0 = the legacy default stream, which is also what torch reports for its default stream).  None = the engine's
        own stream (`engine_stream()`), the one `synchronize()` waits for."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        if stream is None:
            stream = self.engine_stream()
        check(self.lib.vitb200_forward_device(self._h, images.data_ptr(), images.shape[0], flags, stream or None))

    def engine_stream(self) -> int:
        """Raw cudaStream_t of the engine's own stream."""
        p = C.c_void_p()
        check(self.lib.vitb200_engine_stream(self._h, C.byref(p)))
        return int(p.value or 0)

    def workspace_generation(self) -> int:
        """Changes whenever the engine re-allocated its activation buffers (device-resident state is gone)."""
        return int(self.lib.vitb200_workspace_generation(self._h))

    def reserve(self, batch: int, flags: int = 0) -> None:
        """Grow the workspace for `batch` images and the outputs in `flags` now, so that no later call can."""
        check(self.lib.vitb200_reserve(self._h, batch, flags))

    def set_graphs(self, on: bool) -> None:
        check(self.lib.vitb200_set_graphs(self._h, 1 if on else 0))

    def graph_replays(self) -> int:
        return int(self.lib.vitb200_graph_replays(self._h))

    def bind_outputs(self, logits: Optional[int] = None, cls_maps: Optional[int] = None, cls_layer_stride: int = 0,
                     rollout: Optional[int] = None) -> None:
        """Raw device addresses (possibly peer-mapped: another GPU's memory over NVLink) that forward_device's producing
        kernels store logits / CLS maps / rollout into; all None restores the engine's own buffers.  See
        include/vitb200.h (vitb200_bind_outputs) for the layouts and dist.PeerPush for the multi-GPU use."""
        check(self.lib.vitb200_bind_outputs(self._h, logits, cls_maps, cls_layer_stride, rollout))

    def profile_forward(self, images: torch.Tensor, flags: int = 0) -> Dict[str, tuple]:
        """One forward with a CUDA event in front of every launch: {kernel: (launches, total_ms)} plus "total"."""
        assert images.is_cuda and images.dtype == torch.float32 and images.is_contiguous()
        buf = C.create_string_buffer(8192)
        check(self.lib.vitb200_profile_forward(self._h, images.data_ptr(), images.shape[0], flags, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split(",")
            out[name] = (int(n), float(ms))
        return out

    def device_output(self, which: int, shape, batch_capacity: Optional[int] = None) -> torch.Tensor:
        """Zero-copy torch view of an engine-owned device buffer (valid until the next growing call)."""
        p, pitch = C.c_void_p(), C.c_int()
        check(self.lib.vitb200_device_output(self._h, which, C.byref(p), C.byref(pitch)))
        return _device_view(p.value, shape, self.device)

    def synchronize(self) -> None:
        check(self.lib.vitb200_synchronize(self._h))

    # ---- deferred host outputs -----------------------------------------------------------------------
    def set_deferred(self, on: bool) -> None:
        """Node-granular calls without a host input stop synchronising; their results come back as PendingTensor
        (pinned, valid on first access).  See vitb200_set_deferred in include/vitb200.h."""
        if not on and self._deferred:
            self._drain(self._issued)
        check(self.lib.vitb200_set_deferred(self._h, 1 if on else 0))
        self._deferred = bool(on)

    def _request_bytes(self, batch: int) -> int:
        """Pinned bytes the outputs of one default-graph request need (embed + L layers with maps + head + rollout)."""
        c = self.cfg
        floats = ((c.num_layers + 1) * c.tokens * c.hidden_dim + c.num_layers * (c.tokens * c.tokens + c.num_heads * (c.tokens - 1))
                  + c.num_classes + c.tokens - 1)
        return batch * floats * 4 + 256 * (3 * c.num_layers + 4) + self.PREFIX_GAP

    def prewarm_host_outputs(self, batch: int = 1, requests: int = 16) -> None:
        """Fill torch's pinned-memory cache with `requests` request slabs.  The reference's request graphs are reference
        cycles (Node <-> Edge, main/graph.py:6-53), so the tensors of a finished request are only released by the cyclic
        collector some requests later; without this every request until the first collections pays a cudaHostAlloc
        (milliseconds)."""
        slab = self._request_bytes(batch)
        rounded = 1 << (slab - 1).bit_length()          # torch's pinned cache rounds block sizes up to a power of two
        requests = max(2, min(requests, (512 << 20) // rounded))   # at most 512 MiB of pinned memory for the warm-up
        held = [torch.empty(slab, dtype=torch.uint8, pin_memory=True) for _ in range(requests)]
        del held

    def begin_request(self) -> None:
        # (the slab fields are guarded by _book: responses may be encoded on other threads)
        """Deferred mode: the next host output starts a fresh pinned slab (called by the plugin at a request's first
        node).  Outputs are bump-allocated views of ONE pinned allocation per request -- a region is never handed out
        twice, so earlier responses stay valid, and the slab returns to torch's cache when its last view dies."""
        with self._book:
            self._slab = None

    # Wire-ready layout of the request slab (deferred mode): the slab is laid out as the RESPONSE the host mirror will
    # encode (message.py: 16-byte header, JSON index, then per tensor an 8 + 4 * ndim byte block header followed by the
    # fp32 payload), so that `Response.encode` writes the small headers into the gaps and returns a view of the slab --
    # the 9.8 MB of a ViT-B/16 request are then copied once (device -> pinned host), not twice.  `_host_out` leaves the
    # gap of the block header in front of every payload and `PREFIX_GAP` bytes in front of the first one.
    PREFIX_GAP = 16384

    def _host_out(self, *shape, final=None) -> torch.Tensor:
        """A host buffer for an output of `shape`; `final` is the shape of the view that will be handed out (it fixes
        the size of the wire block header reserved in front of the payload)."""
        if not self._deferred:
            return torch.empty(*shape, dtype=torch.float32)
        n = 1
        for d_ in shape:
            n *= d_
        fs = tuple(final) if final is not None else tuple(shape)
        hdr = 8 + 4 * len(fs)
        strides, acc = [], 1
        for d_ in reversed(fs):
            strides.append(acc)
            acc *= d_
        with self._book:
            if self._slab is None or self._slab_off + hdr + n * 4 > self._slab.numel():
                size = (max(self._request_bytes(shape[0]), self.PREFIX_GAP + hdr + n * 4) + 3) // 4 * 4
                self._slab = torch.empty(size, dtype=torch.uint8, pin_memory=True)
                self._slab_f32 = self._slab.view(torch.float32)
                self._slab_off = self.PREFIX_GAP
            off = self._slab_off + hdr
            # ONE torch call per output (a slice + two views cost ~8 us of host time each, 40 outputs per request): the
            # view is created in its final shape straight away
            v = self._slab_f32.as_strided(fs, strides[::-1], off >> 2)
            self._slab_off = off + n * 4
            self._wire_pending[v.data_ptr()] = (self._slab, off, fs)   # fs: the handed-out shape, cached for the encoder
        return v

    def _checked_out(self, status: int, out: torch.Tensor) -> None:
        """`check` for calls that may already have enqueued a copy into `out`: on failure wait for the stream before
        the exception lets go of the pinned buffer (it would return to torch's cache with the copy still in flight)."""
        if status != 0 and self._deferred:
            try:
                self.synchronize()
            except Exception:
                pass
        check(status)

    def _issue(self, out: torch.Tensor, shape=None) -> torch.Tensor:
        """Called right AFTER the copy into `out` was enqueued: a count read before a synchronize therefore only
        covers copies that the synchronize waits for.  `shape`: the view to hand out (taken here, on the plain
        tensor, because a view of a PendingTensor would have to wait)."""
        if shape is not None and tuple(out.shape) != tuple(shape):
            out = out.view(shape)
        if not self._deferred:
            return out
        if len(self._keep) >= 1024:      # nobody looked at the results: do not pile up pinned buffers
            self._drain(self._issued)
        t = out.as_subclass(PendingTensor)
        with self._book:
            self._issued += 1
            t._seq, t._engine = self._issued, self
            self._keep.append((t._seq, out))
            # (slab, payload offset, shape): message.Response.encode and the plugin read these instead of calling
            # tensor methods -- every metadata call on a PendingTensor goes through __torch_function__ (~1 us; ~350 of
            # them per single-image request before)
            t._wire = self._wire_pending.pop(out.data_ptr(), None)
        return t

    def _drain(self, seq: int) -> None:
        with self._book:
            if seq <= self._drained or not self._h:      # a closed engine drained everything in close()
                return
            mark = self._issued      # every copy counted here was enqueued before the wait below starts
        self.synchronize()
        with self._book:
            if mark > self._drained:
                self._drained = mark
            self._keep = [(s_, t) for s_, t in self._keep if s_ > self._drained]

    def launch_count(self) -> int:
        return int(self.lib.vitb200_launch_count(self._h))

    # ---- node-granular stages ----------------------------------------------------------------------
    def stage_embed(self, images: torch.Tensor) -> None:
        assert images.device.type == "cpu" and images.dtype == torch.float32 and images.is_contiguous()
        check(self.lib.vitb200_stage_embed(self._h, images.data_ptr(), images.shape[0]))

    def stage_transform(self, images: torch.Tensor, resize: int) -> torch.Tensor:
        """`<model>:transform`: images CPU fp32 [B,3,H,W] in [0,1] -> preprocessed CPU fp32 [B,3,S,S]; the result also
        stays on the device for stage_embed_resident."""
        assert images.device.type == "cpu" and images.dtype == torch.float32 and images.is_contiguous() and images.dim() == 4
        B, _, H, W = images.shape
        out = torch.empty(B, 3, self.cfg.image_size, self.cfg.image_size)
        check(self.lib.vitb200_stage_transform(self._h, images.data_ptr(), B, H, W, resize, out.data_ptr()))
        return out

    def stage_embed_resident(self, batch: int) -> None:
        check(self.lib.vitb200_stage_embed_resident(self._h, batch))

    def stage_layer(self, layer: int, batch: int, flags: int) -> None:
        check(self.lib.vitb200_stage_layer(self._h, layer, batch, flags))

    def stage_layer_fetch(self, layer: int, batch: int, flags: int, half: bool, tok_shape, map_shape, cls_shape):
        """A whole layer node -- `stage_layer` (or `stage_attn_block` with `half`) and its three host outputs (token
        stream, head-averaged map, class-token grid; shapes as in the getters) -- in ONE call into the library."""
        c = self.cfg
        tok = self._host_out(batch, c.tokens, c.hidden_dim, final=tok_shape)
        amap = self._host_out(batch, c.tokens, c.tokens, final=map_shape)
        cls = self._host_out(batch, c.num_heads, c.tokens - 1, final=cls_shape)
        status = self.lib.vitb200_stage_layer_fetch(self._h, layer, batch, flags, 1 if half else 0, tok.data_ptr(),
                                                    amap.data_ptr(), cls.data_ptr())
        self._checked_out(status, tok)
        return self._issue(tok, tok_shape), self._issue(amap, map_shape), self._issue(cls, cls_shape)

    def stage_attn_block(self, layer: int, batch: int, flags: int = EMIT_AVG | EMIT_CLS) -> None:
        """First half of an EncoderBlock on the resident token stream: x <- x + out_proj(MHA(LN1 x)), maps per `flags`."""
        check(self.lib.vitb200_stage_attn_block(self._h, layer, batch, flags))

    def stage_mlp_block(self, layer: int, batch: int) -> None:
        """Second half: x <- x + MLP(LN2 x)."""
        check(self.lib.vitb200_stage_mlp_block(self._h, layer, batch))

    def stage_head(self, batch: int, shape=None) -> torch.Tensor:
        out = self._host_out(batch, self.cfg.num_classes, final=shape)
        self._checked_out(self.lib.vitb200_stage_head(self._h, batch, out.data_ptr()), out)
        return self._issue(out, shape)

    def stage_rollout(self, batch: int, shape=None) -> torch.Tensor:
        out = self._host_out(batch, self.cfg.tokens - 1, final=shape)
        self._checked_out(self.lib.vitb200_stage_rollout(self._h, batch, out.data_ptr()), out)
        return self._issue(out, shape)

    def set_tokens(self, tokens: torch.Tensor) -> None:
        assert tokens.device.type == "cpu" and tokens.dtype == torch.float32 and tokens.is_contiguous()
        check(self.lib.vitb200_set_tokens(self._h, tokens.data_ptr(), tokens.shape[0]))

    def get_tokens(self, batch: int, shape=None) -> torch.Tensor:
        out = self._host_out(batch, self.cfg.tokens, self.cfg.hidden_dim, final=shape)
        self._checked_out(self.lib.vitb200_get_tokens(self._h, out.data_ptr(), batch), out)
        return self._issue(out, shape)

    def set_avg_map(self, layer: int, amap: torch.Tensor) -> None:
        assert amap.device.type == "cpu" and amap.dtype == torch.float32 and amap.is_contiguous()
        check(self.lib.vitb200_set_avg_map(self._h, layer, amap.data_ptr(), amap.shape[0]))

    def get_avg_map(self, layer: int, batch: int, shape=None) -> torch.Tensor:
        N = self.cfg.tokens
        out = self._host_out(batch, N, N, final=shape)
        self._checked_out(self.lib.vitb200_get_avg_map(self._h, layer, out.data_ptr(), batch), out)
        return self._issue(out, shape)

    def get_cls_map(self, layer: int, batch: int, shape=None) -> torch.Tensor:
        out = self._host_out(batch, self.cfg.num_heads, self.cfg.tokens, final=shape)
        self._checked_out(self.lib.vitb200_get_cls_map(self._h, layer, out.data_ptr(), batch), out)
        return self._issue(out, shape)

    def get_cls_grid(self, layer: int, batch: int, shape=None) -> torch.Tensor:
        """[B, H, N-1]: the class token's attention to the patch tokens per head (class column dropped by the copy)."""
        out = self._host_out(batch, self.cfg.num_heads, self.cfg.tokens - 1, final=shape)
        self._checked_out(self.lib.vitb200_get_cls_grid(self._h, layer, out.data_ptr(), batch), out)
        return self._issue(out, shape)

    def get_head_map(self, layer: int, batch: int, shape=None) -> torch.Tensor:
        N = self.cfg.tokens
        out = self._host_out(batch, self.cfg.num_heads, N, N, final=shape)
        self._checked_out(self.lib.vitb200_get_head_map(self._h, layer, out.data_ptr(), batch), out)
        return self._issue(out, shape)


# ---- peer memory / completion flags (multi-GPU result exchange, dist.PeerPush) --------------------------------
def peer_alloc(device: int, nbytes: int):
    """(device pointer, 64-byte IPC handle) of a zeroed allocation on `device` that other processes can map."""
    lib = load_library()
    p, h = C.c_void_p(), C.create_string_buffer(64)
    check(lib.vitb200_peer_alloc(device, nbytes, C.byref(p), h))
    return int(p.value), bytes(h.raw)


def peer_open(device: int, handle: bytes) -> int:
    lib = load_library()
    p = C.c_void_p()
    check(lib.vitb200_peer_open(device, C.create_string_buffer(handle, 64), C.byref(p)))
    return int(p.value)


def peer_close(device: int, ptr: int) -> None:
    check(load_library().vitb200_peer_close(device, ptr))


def peer_free(device: int, ptr: int) -> None:
    check(load_library().vitb200_peer_free(device, ptr))


def flag_signal(flag_ptr: int, value: int, stream: int) -> None:
    """Enqueue on `stream` (raw cudaStream_t): *flag = value (system-scope release), behind everything already queued."""
    check(load_library().vitb200_flag_signal(flag_ptr, value, stream or None))


def flag_wait(flag_ptr: int, value: int, stream: int) -> None:
    """Enqueue on `stream`: block the stream until *flag >= value (device-side spin with back-off)."""
    check(load_library().vitb200_flag_wait(flag_ptr, value, stream or None))


def _device_view(ptr: int, shape, device: int, typestr: str = "<f4") -> torch.Tensor:
    """Wrap a raw device pointer as a torch tensor through __cuda_array_interface__ (no copy)."""

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {
        "shape": tuple(int(s) for s in shape),
        "typestr": typestr,
        "data": (int(ptr), False),
        "version": 2,
    }
    return torch.as_tensor(h, device=f"cuda:{device}")


# ---- single-kernel wrappers used by the parity tests (device tensors in, device tensors out) ---------
def op_gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, resid: Optional[torch.Tensor] = None,
            gelu: bool = False, out_f32: bool = False) -> torch.Tensor:
    lib = load_library()
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    check(lib.vitb200_op_gemm(a.data_ptr(), w.data_ptr(), _ptr(bias), _ptr(resid), out.data_ptr(), M, N, K, int(gelu),
                              int(out_f32), None))
    return out


def op_layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    lib = load_library()
    rows, d = x.shape
    y = torch.empty(rows, d, device=x.device, dtype=torch.bfloat16)
    check(lib.vitb200_op_layernorm(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), rows, d, eps, None))
    return y


def ln_slot_width(d: int) -> int:
    """Columns per LayerNorm partial-sum slot of a row of width d (csrc/rowwise.cuh ln_slot_width)."""
    return 128 if d % 256 == 0 else 64


def op_gemm_residual_stats(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, resid: torch.Tensor):
    """Residual GEMM with the LayerNorm-folding producer epilogue: (x fp32 [M,N], bf16 copy, partial sums
    [M, N / ln_slot_width(N), 2])."""
    lib = load_library()
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device)
    xb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    stats = torch.zeros(M, N // ln_slot_width(N), 2, device=a.device)
    check(lib.vitb200_op_gemm_ex(a.data_ptr(), w.data_ptr(), bias.data_ptr(), resid.data_ptr(), out.data_ptr(), M, N, K, 0, 1,
                                 xb.data_ptr(), stats.data_ptr(), None, None, 0.0, None))
    return out, xb, stats


def op_fold_ln(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: torch.Tensor):
    """(W' bf16, colsum, bias') of a Linear that consumes LayerNorm(x; gamma, beta)."""
    lib = load_library()
    N, K = w.shape
    wq = torch.empty(N, K, device=w.device, dtype=torch.bfloat16)
    colsum = torch.empty(N, device=w.device)
    bias_out = torch.empty(N, device=w.device)
    check(lib.vitb200_op_fold_ln(w.data_ptr(), gamma.data_ptr(), beta.data_ptr(), bias.data_ptr(), wq.data_ptr(),
                                 colsum.data_ptr(), bias_out.data_ptr(), N, K, None))
    return wq, colsum, bias_out


def op_gemm_ln(xb: torch.Tensor, stats: torch.Tensor, wq: torch.Tensor, colsum: torch.Tensor, bias: torch.Tensor,
               gelu: bool = False, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm(x) W^T + b [-> GELU] from the bf16 copy of x, its partial sums and folded weights; bf16 result."""
    lib = load_library()
    M, K = xb.shape
    N = wq.shape[0]
    out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    check(lib.vitb200_op_gemm_ex(xb.data_ptr(), wq.data_ptr(), bias.data_ptr(), None, out.data_ptr(), M, N, K, int(gelu), 0,
                                 None, None, stats.data_ptr(), colsum.data_ptr(), eps, None))
    return out


def op_split_bf16(x: torch.Tensor):
    """(hi, lo) bf16 tensors with hi + lo ~= x to ~2^-17 relative."""
    lib = load_library()
    x = x.contiguous()
    hi = torch.empty_like(x, dtype=torch.bfloat16)
    lo = torch.empty_like(x, dtype=torch.bfloat16)
    check(lib.vitb200_op_split_bf16(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), None))
    return hi, lo


def op_gemm_split(a: torch.Tensor, w: torch.Tensor, bias=None, resid=None, gelu=False, out_f32=True):
    """epilogue(a @ w.T) for fp32 a [M,K], w [N,K] through split-bf16 operands (the fp32x3 precision mode)."""
    lib = load_library()
    M, K = a.shape
    N = w.shape[0]
    ah, al = op_split_bf16(a)
    wh, wl = op_split_bf16(w)
    out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    out_lo = None if out_f32 else torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    check(lib.vitb200_op_gemm_split(ah.data_ptr(), al.data_ptr(), wh.data_ptr(), wl.data_ptr(), _ptr(bias), _ptr(resid),
                                    out.data_ptr(), _ptr(out_lo), M, N, K, int(gelu), int(out_f32), None))
    return out if out_f32 else out.float() + out_lo.float()


def op_attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, want_avg=True, want_cls=True, want_heads=False,
                 head_dim: int = 64):
    lib = load_library()
    d = heads * head_dim
    pitch = (tokens + 15) // 16 * 16
    ctx = torch.zeros(batch * tokens, d, device=qkv.device, dtype=torch.bfloat16)
    avg = torch.zeros(batch, tokens, pitch, device=qkv.device) if want_avg else None
    cls = torch.zeros(batch, heads, tokens, device=qkv.device) if want_cls else None
    hm = torch.zeros(batch, heads, tokens, pitch, device=qkv.device) if want_heads else None
    check(lib.vitb200_op_attention_ex(qkv.data_ptr(), ctx.data_ptr(), _ptr(avg), _ptr(cls), _ptr(hm), batch, tokens, heads,
                                      head_dim, pitch, 0, None))
    return ctx, (avg[..., :tokens] if avg is not None else None), cls, (hm[..., :tokens] if hm is not None else None)


def op_preprocess(images: torch.Tensor, resize: int, crop: int) -> torch.Tensor:
    lib = load_library()
    B, _, H, W = images.shape
    out = torch.empty(B, 3, crop, crop, device=images.device)
    check(lib.vitb200_op_preprocess(images.data_ptr(), out.data_ptr(), B, H, W, resize, crop, None))
    return out


def op_patchify(images: torch.Tensor, patch: int) -> torch.Tensor:
    lib = load_library()
    B, _, S, _ = images.shape
    n = (S // patch) ** 2
    out = torch.empty(B * n, 3 * patch * patch, device=images.device, dtype=torch.bfloat16)
    check(lib.vitb200_op_patchify(images.data_ptr(), out.data_ptr(), B, S, patch, None))
    return out


def op_rollout(maps: torch.Tensor) -> torch.Tensor:
    """maps: CUDA fp32 [L, B, N, pitch] -> [B, N-1] (pitch may exceed N)."""
    lib = load_library()
    L, B, N, pitch = maps.shape
    out = torch.empty(B, N - 1, device=maps.device)
    check(lib.vitb200_op_rollout(maps.data_ptr(), B * N * pitch, L, B, N, pitch, out.data_ptr(), None))
    return out
