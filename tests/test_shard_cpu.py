"""CPU, world_size 2, gloo: the multi-GPU host logic (checkpoint broadcast, batch sharding, ordered gather)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    from reface_b200 import shard, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lay, total = synth.flat_layout()
    small = 4096
    flat = torch.arange(small, dtype=torch.float32) if rank == 0 else torch.zeros(small)
    shard.broadcast_checkpoint(flat, 0)
    ok_bcast = bool(torch.equal(flat, torch.arange(small, dtype=torch.float32)))
    batch = {"x": torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3), "y": torch.arange(7)}
    mine = shard.shard_batch(batch, rank, world)
    out = shard.gather_batch(mine["x"] * 2, 7)
    ok_gather = bool(torch.equal(out, batch["x"] * 2))
    even = shard.gather_batch(torch.full((2, 2), float(rank)), 4)
    ok_even = bool(torch.equal(even, torch.tensor([[0., 0.], [0., 0.], [1., 1.], [1., 1.]])))
    q.put((rank, ok_bcast, ok_gather, ok_even, mine["x"].shape[0], total))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
    assert [r[4] for r in res] == [4, 3]                 # 7 items -> 4 + 3, contiguous
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert res[0][5] > 1.3e9                             # flat checkpoint covers all 1.3 B parameters


def test_shard_range_partitions():
    from reface_b200.shard import shard_range
    for total in (1, 7, 8, 64, 240):
        for world in (1, 2, 4, 8):
            rs = [shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def test_video_segments_cover_every_frame_once():
    """Video mode: contiguous chunks round-robined over the ranks; together the ranks cover every frame exactly once
    and the frame index (the reference's segment_id) restores the order."""
    from reface_b200.shard import video_segments
    for n, chunk, world in ((240, 30, 8), (240, 30, 1), (7, 3, 2), (31, 30, 4), (5, 30, 8)):
        seen = []
        for r in range(world):
            for lo, hi in video_segments(n, chunk, r, world):
                assert 0 < hi - lo <= chunk
                seen += list(range(lo, hi))
        assert sorted(seen) == list(range(n))
    assert video_segments(240, 30, 1, 8) == [(30, 60)]
