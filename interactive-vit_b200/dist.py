"""Data-parallel sharding of image batches over the GPUs of one box (one process per GPU).

The path shards over independent units — images: weights are replicated, each rank runs the
whole forward on a contiguous slice of the batch, and the only exchange step is a gather of the (small) results
to rank 0.  The reference has nothing here (single process, CPU); `torch.distributed` is plumbing: NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, start+count) slice of `total` images owned by `rank`; the first `total % world` ranks
    get one extra image, so counts differ by at most one and empty shards only occur when total < world."""
    if total < 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad shard request total={total} world={world} rank={rank}")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def gather_to_rank0(local: torch.Tensor, total: int, batch_dim: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-rank result slices (batch on `batch_dim`) into one tensor on rank 0, in image order.

    Shards may be uneven (see shard_range): every rank pads its slice to the largest shard so that a single
    fixed-size gather is issued, and rank 0 trims the padding.  Returns None on the other ranks."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [shard_range(total, world, r)[1] for r in range(world)]
    cap = max(counts) if counts else 0
    x = local.movedim(batch_dim, 0).contiguous()
    if x.shape[0] != counts[rank]:
        raise ValueError(f"rank {rank} holds {x.shape[0]} images, expected {counts[rank]}")
    if x.shape[0] < cap:
        pad = torch.zeros((cap - x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], dim=0)
    if world == 1:
        out = x[: counts[0]]
        return out.movedim(0, batch_dim)
    bufs: Optional[List[torch.Tensor]] = None
    if rank == 0:
        bufs = [torch.empty_like(x) for _ in range(world)]
    dist.gather(x, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    return out.movedim(0, batch_dim)


def gather_results(local: Dict[str, torch.Tensor], total: int, batch_dims: Dict[str, int], group=None
                   ) -> Optional[Dict[str, torch.Tensor]]:
    """Gather a dict of per-rank results (logits [b,C], avg_maps [L,b,N,N], cls_maps [L,b,H,N], rollout [b,n]...)."""
    out = {}
    for k in sorted(local):
        g = gather_to_rank0(local[k], total, batch_dims.get(k, 0), group)
        if g is not None:
            out[k] = g
    return out if dist.get_rank(group) == 0 else None


RESULT_BATCH_DIMS = {"logits": 0, "rollout": 0, "avg_maps": 1, "cls_maps": 1, "heads": 1, "hidden": 1}


class PackedGather:
    """One fixed-size gather per step, overlapped with the next step's forward.

    Every rank packs its results image-major into ONE staging row block ``[cap, F]`` (F = floats per image over all
    results), a single ``dist.gather`` moves it into rank 0's preallocated ``[world, cap, F]`` buffer, and rank 0 reads
    the results as views of that buffer — no per-result collectives, no allocation and no concatenation inside the
    step.  Two staging / receive sets alternate, and on CUDA the collective is issued from a side stream behind an
    event, so the gather of step i runs under the forward of step i + 1; ``submit`` only makes the compute stream wait
    for the gather that used the same set two steps earlier.

    spec: ``{name: (per-image shape, batch_dim of the local tensor)}``, e.g. ``{"logits": ((1000,), 0),
    "cls_maps": ((12, 12, 197), 1)}`` for a local ``[L, b, H, N]`` tensor.  Results come back image-major:
    ``result(name)`` is ``[total, *per-image shape]`` (for cls_maps: ``[total, L, H, N]``).
    """

    def __init__(self, spec: Dict[str, Tuple[Tuple[int, ...], int]], total: int, device, group=None,
                 dtype: torch.dtype = torch.float32):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.total = total
        self.counts = [shard_range(total, self.world, r)[1] for r in range(self.world)]
        self.cap = max(self.counts) if self.counts else 0
        self.spec, self.offsets, off = dict(spec), {}, 0
        for name in sorted(spec):
            n = 1
            for s in spec[name][0]:
                n *= s
            self.offsets[name] = (off, n)
            off += n
        self.width = off
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.staging = [torch.zeros(self.cap, self.width, dtype=dtype, device=self.device) for _ in range(2)]
        self.recv = ([torch.empty(self.world, self.cap, self.width, dtype=dtype, device=self.device) for _ in range(2)]
                     if self.rank == 0 else [None, None])
        self.work = [None, None]
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self.step = 0
        self.last = -1

    def submit(self, local: Dict[str, torch.Tensor]) -> int:
        """Pack `local` (on the current stream) and start its gather; returns the set index to pass to `result`."""
        s = self.step & 1
        self.step += 1
        if self.work[s] is not None:
            self.work[s].wait()          # CUDA: the current stream waits for the collective that last used this set
            self.work[s] = None
        stage = self.staging[s]
        mine = self.counts[self.rank]
        for name, (off, n) in self.offsets.items():
            shape, bdim = self.spec[name]
            x = local[name].movedim(bdim, 0)
            if x.shape[0] != mine or tuple(x.shape[1:]) != tuple(shape):
                raise ValueError(f"{name}: rank {self.rank} holds {tuple(local[name].shape)}, expected {mine} images of {shape}")
            stage[:mine, off:off + n].view((mine,) + tuple(shape)).copy_(x)
        if self.world == 1:
            self.last = s
            return s
        bufs = list(self.recv[s].unbind(0)) if self.rank == 0 else None
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self.side):
                self.side.wait_event(ready)
                self.work[s] = dist.gather(stage, bufs, dst=0, group=self.group, async_op=True)
        else:
            dist.gather(stage, bufs, dst=0, group=self.group)
        self.last = s
        return s

    def finish(self) -> None:
        """Make the current stream (CUDA) wait for every gather still in flight."""
        for s in (0, 1):
            if self.work[s] is not None:
                self.work[s].wait()
                self.work[s] = None

    def result(self, name: str, s: Optional[int] = None) -> Optional[torch.Tensor]:
        """Rank 0: `[total, *per-image shape]` of set `s` (default: the last submitted), after `finish()` (or after
        the set's own wait); image order is rank order because shards are contiguous.  Other ranks: None."""
        if self.rank != 0:
            return None
        s = self.last if s is None else s
        off, n = self.offsets[name]
        shape = tuple(self.spec[name][0])
        if self.world == 1:
            return self.staging[s][: self.total, off:off + n].reshape((self.total,) + shape)
        rows = self.recv[s][:, :, off:off + n]
        if all(c == self.cap for c in self.counts):
            return rows.reshape((self.total,) + shape)
        return torch.cat([rows[r, : self.counts[r]] for r in range(self.world)], dim=0).reshape((self.total,) + shape)


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class PushLayout:
    """Where every rank's results live inside ONE receive set on rank 0 (offsets in floats).

    A set holds the gathered tensors in their final, image-ordered layouts, so rank 0 reads them as plain views:
    ``logits [total, classes]``, ``cls_maps [L, total, H, N]``, ``rollout [total, N - 1]``.  Rank r owns the images
    ``shard_range(total, world, r)`` of each; for the CLS maps that is a strided region (one slab per layer), which is
    why the engine takes a layer stride (``vitb200_bind_outputs``).  Regions start on 16-byte boundaries."""

    def __init__(self, total: int, world: int, classes: int, layers: int, heads: int, tokens: int):
        self.total, self.world = total, world
        self.classes, self.layers, self.heads, self.tokens = classes, layers, heads, tokens
        self.starts = [shard_range(total, world, r)[0] for r in range(world)]
        self.counts = [shard_range(total, world, r)[1] for r in range(world)]
        self.logits_off = 0
        self.cls_off = _round_up(total * classes, 4)
        self.cls_layer_stride = total * heads * tokens
        self.rollout_off = self.cls_off + _round_up(layers * self.cls_layer_stride, 4)
        self.set_floats = _round_up(self.rollout_off + total * (tokens - 1), 4)

    def rank_offsets(self, rank: int) -> Dict[str, int]:
        """Float offsets (inside a set) of the first image of `rank` in each result."""
        s = self.starts[rank]
        return {"logits": self.logits_off + s * self.classes,
                "cls_maps": self.cls_off + s * self.heads * self.tokens,
                "rollout": self.rollout_off + s * (self.tokens - 1)}

    def views(self, set_buf: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Rank 0: the three gathered results as views of one set (a flat float tensor of `set_floats`)."""
        t, L, H, N = self.total, self.layers, self.heads, self.tokens
        return {"logits": set_buf[self.logits_off:self.logits_off + t * self.classes].view(t, self.classes),
                "cls_maps": set_buf[self.cls_off:self.cls_off + L * self.cls_layer_stride].view(L, t, H, N),
                "rollout": set_buf[self.rollout_off:self.rollout_off + t * (N - 1)].view(t, N - 1)}


class PeerPush:
    """The result exchange of the data-parallel forward without a collective and without a barrier: every rank's
    PRODUCING kernels (head GEMM epilogue, the attention kernel's CLS-row writer, the rollout kernel) store their
    results straight into rank 0's receive set over NVLink / NVSwitch peer memory (`VitEngine.bind_outputs`), already
    in the final image-ordered layout.  There is no pack kernel, no staging copy, no NCCL call and nothing to
    concatenate; ordering is carried by per-rank completion flags next to the sets (csrc/peer.cuh):

        done[r][s]   rank r, behind its forward into set s          (monotone step counter, system-scope release store)
        free[s]      rank 0's reader, behind its reads of set s

    Per rank and step i (set s = i % sets), all stream-ordered on the stream the forward runs on: ``begin()`` waits for
    ``free[s] >= i - sets + 1`` (only from the second rotation on) and binds the set, the forward writes, ``end()``
    signals ``done[rank][s] = i + 1``.  Rank 0 additionally runs a READER stream: wait for every rank's ``done[.][s]``,
    run the consumer (``consumer(s, views)``: torch work on the current = reader stream, e.g. the device-to-host copy
    of the gathered results), signal ``free[s]``.  No rank waits for another rank's forward of the same step: with
    three sets a rank may be two forwards ahead of rank 0's reader, and the reader is never on a forward's stream.
    On rank 0 the two streams of its OWN GPU are ordered with CUDA events, not flags (kernels that spin on a flag
    written by another launch on the same GPU are not guaranteed to make progress).

    The receive sets and flags are ONE `cudaMalloc` allocation of rank 0 whose CUDA IPC handle travels to the other
    ranks through ``torch.distributed`` (public API; plumbing) and is mapped there with ``cudaIpcOpenMemHandle``
    (``engine.peer_alloc / peer_open``).  Raises on every rank if any rank cannot map it -- callers that want a
    fallback transport use `PackedGather` (NCCL)."""

    FLAG_BYTES = 128   # one flag per 128-byte line

    def __init__(self, eng, total: int, device, group=None, sets: int = 3, stream: Optional[int] = None,
                 consumer=None):
        from . import engine as E

        self.E, self.eng = E, eng
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        cfg = eng.cfg
        self.layout = PushLayout(total, self.world, cfg.num_classes, cfg.num_layers, cfg.num_heads, cfg.tokens)
        self.sets = sets
        self.device = torch.device(device)
        self.dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        # the stream the forwards run on: the flags are enqueued on it, in front of and behind each forward
        self.stream = int(stream) if stream is not None else eng.engine_stream()
        self.consumer = consumer
        self.flag_off = _round_up(4 * sets * self.layout.set_floats, self.FLAG_BYTES)
        nbytes = self.flag_off + self.FLAG_BYTES * (self.world * sets + sets)
        self.root_ptr = 0
        self.owner = False
        # (1) rank 0 allocates, (2) its IPC handle is broadcast, (3) everyone maps it, (4) agree (MIN over ranks)
        err: Optional[Exception] = None
        payload = [None]
        if self.rank == 0:
            try:
                self.root_ptr, handle = E.peer_alloc(self.dev_index, nbytes)
                self.owner = True
                payload = [handle]
            except Exception as ex:
                err = ex
        dist.broadcast_object_list(payload, src=dist.get_global_rank(self.group, 0), group=self.group)
        if self.rank != 0:
            try:
                if payload[0] is None:
                    raise RuntimeError("rank 0 could not allocate the receive sets")
                self.root_ptr = E.peer_open(self.dev_index, payload[0])   # rank 0's memory as mapped into THIS process
            except Exception as ex:
                err = ex
        ok = torch.tensor([0.0 if err is not None else 1.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if ok.item() == 0:
            self._release()
            raise RuntimeError("PeerPush: the receive sets could not be mapped on at least one rank"
                               f"{'' if err is None else f' (here: {type(err).__name__}: {err})'}")
        self.buf = (E._device_view(self.root_ptr, (sets * self.layout.set_floats,), self.dev_index)
                    if self.rank == 0 else None)
        self.reader = torch.cuda.Stream(device=self.device) if self.rank == 0 else None
        self.closed = [None] * sets     # rank 0: event behind the reader's pass over set s
        self.step = 0
        self.last = -1

    # ---- flag addresses (in rank 0's allocation, as mapped here)
    def _done(self, rank: int, s: int) -> int:
        return self.root_ptr + self.flag_off + self.FLAG_BYTES * (rank * self.sets + s)

    def _free(self, s: int) -> int:
        return self.root_ptr + self.flag_off + self.FLAG_BYTES * (self.world * self.sets + s)

    def begin(self) -> int:
        """Bind this step's receive set as the engine's output destination; returns the set index."""
        i, s = self.step, self.step % self.sets
        if i >= self.sets:
            if self.rank == 0:     # own GPU: an event, not a flag
                _raw_stream_wait_event(self.stream, self.closed[s], self.device)
            else:
                self.E.flag_wait(self._free(s), i - self.sets + 1, self.stream)
        off = self.layout.rank_offsets(self.rank)
        base = self.root_ptr + 4 * s * self.layout.set_floats
        self.eng.bind_outputs(base + 4 * off["logits"], base + 4 * off["cls_maps"], self.layout.cls_layer_stride,
                              base + 4 * off["rollout"])
        return s

    def end(self) -> int:
        """Behind the forward: publish that this rank's part of the set is written; rank 0 also queues its reader."""
        i, s = self.step, self.step % self.sets
        self.step += 1
        self.last = s
        if self.rank != 0:
            self.E.flag_signal(self._done(self.rank, s), i + 1, self.stream)
            return s
        written = torch.cuda.Event()
        _raw_stream_record_event(self.stream, written, self.device)
        with torch.cuda.stream(self.reader):
            self.reader.wait_event(written)
            for r in range(1, self.world):
                self.E.flag_wait(self._done(r, s), i + 1, self.reader.cuda_stream)
            if self.consumer is not None:
                n = self.layout.set_floats
                self.consumer(s, self.layout.views(self.buf[s * n:(s + 1) * n]))
            self.E.flag_signal(self._free(s), i + 1, self.reader.cuda_stream)
            done = torch.cuda.Event()
            done.record(self.reader)
        self.closed[s] = done
        return s

    def wait(self, s: Optional[int] = None) -> None:
        """Rank 0: make the current stream wait until set `s` (default: the last one) is complete and consumed."""
        s = self.last if s is None else s
        if self.rank == 0 and self.closed[s] is not None:
            torch.cuda.current_stream(self.device).wait_event(self.closed[s])

    def finish(self) -> None:
        for s in range(self.sets):
            self.wait(s)

    def result(self, name: str, s: Optional[int] = None) -> Optional[torch.Tensor]:
        """Rank 0: view of the gathered result in set `s` (valid after `wait(s)`); other ranks: None."""
        if self.rank != 0:
            return None
        s = self.last if s is None else s
        n = self.layout.set_floats
        return self.layout.views(self.buf[s * n:(s + 1) * n])[name]

    def _release(self) -> None:
        if self.root_ptr:
            try:
                if self.owner:
                    self.E.peer_free(self.dev_index, self.root_ptr)
                else:
                    self.E.peer_close(self.dev_index, self.root_ptr)
            except Exception:
                pass
        self.root_ptr = 0
        self.buf = None

    def close(self) -> None:
        """Collective: every rank unbinds and unmaps; rank 0 frees the sets once nobody can write them any more."""
        self.finish()
        self.eng.bind_outputs()
        torch.cuda.synchronize(self.device)
        if self.rank != 0:
            self._release()
        dist.barrier(group=self.group)
        if self.rank == 0:
            self._release()


def _raw_stream(handle: int, device) -> "torch.cuda.Stream":
    """torch view of a raw cudaStream_t (the engine's own stream, or a torch stream's handle)."""
    return torch.cuda.ExternalStream(handle, device=device) if handle else torch.cuda.default_stream(device)


def _raw_stream_wait_event(handle: int, event, device) -> None:
    if event is not None:
        _raw_stream(handle, device).wait_event(event)


def _raw_stream_record_event(handle: int, event, device) -> None:
    event.record(_raw_stream(handle, device))
