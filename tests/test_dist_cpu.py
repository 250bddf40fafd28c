"""Data-parallel sharding + gather to rank 0 on CPU (gloo, world_size 2 and 3, uneven shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from interactive_vit_b200.dist import RESULT_BATCH_DIMS, PackedGather, gather_results, shard_range


def test_shard_range_partitions_every_image_exactly_once():
    for total in (0, 1, 2, 7, 8, 255, 256, 4096):
        for world in (1, 2, 3, 4, 8):
            covered = []
            for r in range(world):
                s, c = shard_range(total, world, r)
                covered += list(range(s, s + c))
            assert covered == list(range(total))
            counts = [shard_range(total, world, r)[1] for r in range(world)]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, count = shard_range(total, world, rank)
        L, N, H, C = 2, 5, 3, 8
        idx = torch.arange(start, start + count, dtype=torch.float32)
        local = {
            "logits": idx[:, None] + torch.arange(C)[None, :] * 0.001,
            "rollout": idx[:, None].expand(count, N - 1).clone(),
            "avg_maps": idx[None, :, None, None].expand(L, count, N, N).clone() + torch.arange(L)[:, None, None, None] * 1000,
            "cls_maps": idx[None, :, None, None].expand(L, count, H, N).clone(),
        }
        out = gather_results(local, total, RESULT_BATCH_DIMS)
        # the pipelined packed gather (what bench.py uses): two steps in flight, image-major results, same values
        pg = PackedGather({"logits": ((C,), 0), "cls_maps": ((L, H, N), 1), "rollout": ((N - 1,), 0)}, total, "cpu")
        packed_ok = True
        for step in range(3):
            shifted = {k: v + step for k, v in local.items()}
            s_idx = pg.submit(shifted)
            pg.finish()
            if rank == 0:
                want = torch.arange(total, dtype=torch.float32) + step
                packed_ok = packed_ok and pg.result("logits", s_idx).shape == (total, C)
                packed_ok = packed_ok and torch.equal(pg.result("logits")[:, 0], want)
                packed_ok = packed_ok and pg.result("cls_maps").shape == (total, L, H, N)
                packed_ok = packed_ok and torch.equal(pg.result("cls_maps")[:, 1, 2, 4], want)
                packed_ok = packed_ok and torch.equal(pg.result("rollout")[:, 0], want)
            else:
                assert pg.result("logits") is None
        if rank == 0:
            assert packed_ok
            ok = out["logits"].shape == (total, C) and torch.equal(out["logits"][:, 0], torch.arange(total, dtype=torch.float32))
            ok = ok and out["avg_maps"].shape == (L, total, N, N)
            ok = ok and torch.equal(out["avg_maps"][1, :, 0, 0], torch.arange(total, dtype=torch.float32) + 1000)
            ok = ok and torch.equal(out["cls_maps"][0, :, 2, 4], torch.arange(total, dtype=torch.float32))
            ok = ok and torch.equal(out["rollout"][:, 0], torch.arange(total, dtype=torch.float32))
            q.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 8), (2, 7), (3, 4)])
def test_gather_results_gloo(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_push_layout_regions_are_disjoint_aligned_and_in_image_order():
    """dist.PushLayout: every rank's slice of every result lies inside its region of the receive set, slices of
    different ranks do not overlap, regions start on 16-byte boundaries, and the views rank 0 reads are the final
    image-ordered tensors (rank order = image order because shards are contiguous)."""
    from interactive_vit_b200.dist import PushLayout

    for total, world in ((8, 2), (7, 3), (512, 8), (3, 4)):
        C, L, H, N = 1000, 3, 6, 197
        lay = PushLayout(total, world, C, L, H, N)
        assert lay.cls_off % 4 == 0 and lay.rollout_off % 4 == 0 and lay.set_floats % 4 == 0
        owner = torch.full((lay.set_floats,), -1, dtype=torch.int32)
        for r in range(world):
            off, cnt = lay.rank_offsets(r), lay.counts[r]
            spans = [(off["logits"], cnt * C), (off["rollout"], cnt * (N - 1))]
            spans += [(off["cls_maps"] + l * lay.cls_layer_stride, cnt * H * N) for l in range(L)]
            for a, n in spans:
                assert a >= 0 and a + n <= lay.set_floats
                assert (owner[a:a + n] == -1).all(), "two ranks (or two results) share floats of the set"
                owner[a:a + n] = r
        v = lay.views(owner)
        assert v["logits"].shape == (total, C) and v["cls_maps"].shape == (L, total, H, N) and v["rollout"].shape == (total, N - 1)
        for r in range(world):
            s, c = lay.starts[r], lay.counts[r]
            assert (v["logits"][s:s + c] == r).all() and (v["cls_maps"][:, s:s + c] == r).all() and (v["rollout"][s:s + c] == r).all()
