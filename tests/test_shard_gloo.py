"""The N > 1 host logic on CPU: world_size-2 gloo run of the row-band sharding + gather
(pwnfps_b200/shard.py).  The per-rank "trace" is the CPU oracle (test infrastructure);
what is under test is band ownership, packing, the gather and the stitch."""
import os
import socket

import numpy as np
import pytest

from pwnfps_b200 import shard


def test_band_ownership_partitions_rows():
    for h in (1, 31, 32, 33, 240, 1080, 4320):
        for n in (1, 2, 3, 4, 8):
            rows = [shard.owned_rows(h, n, r) for r in range(n)]
            allr = np.sort(np.concatenate(rows))
            assert (allr == np.arange(h)).all()
            assert shard.max_owned_rows(h, n) == max(len(r) for r in rows)
            for r in range(n):
                assert all((y // 32) % n == r for y in rows[r])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import orcbind
        from pwnfps_b200 import api, levels
        lv = levels.default_level()
        o = orcbind.Oracle()
        o.level_set(lv.cells, lv.portals_i32())
        cam = api.mat4_roty(api.cam_spawn(lv), 0.8)
        w, h = 200, 150
        sb = np.zeros((h, w), np.uint32)
        zb = np.zeros((h, w), np.float32)
        for b in shard.owned_bands(h, world, rank):
            o.trace_frame(cam, 0.5, w, h, y0=b * 32, y1=min(h, b * 32 + 32), out=(sb, zb), nthreads=1)
        mine = shard.owned_rows(h, world, rank)
        other = np.setdiff1d(np.arange(h), mine)
        assert (sb[other] == 0).all()
        full = shard.gather_frame(torch.from_numpy(sb.view(np.int32)), h, rank, world, dist)
        zfull = shard.gather_frame(torch.from_numpy(zb), h, rank, world, dist)
        if rank == 0:
            want_s, want_z, _ = o.trace_frame(cam, 0.5, w, h, nthreads=1)
            ok = (full.numpy().view(np.uint32) == want_s).all() and (zfull.numpy().view(np.uint32) == want_z.view(np.uint32)).all()
            q.put(bool(ok))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_band_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_rebalance_is_a_partition_and_evens_the_cost():
    """shard.rebalance: boundaries stay a partition of the bands with at least one band per shard, and iterating it on a
    known cost profile brings the blocks' costs together"""
    rng = np.random.default_rng(3)
    for nb, world in ((135, 8), (135, 2), (34, 4), (9, 8), (8, 8), (68, 3)):
        cost = 1.0 + 0.4 * np.linspace(0, 1, nb) + 0.05 * rng.random(nb)          # rows below the horizon cost more
        bounds = [s * nb // world for s in range(world + 1)]
        for _ in range(4):
            times = [cost[bounds[s]:bounds[s + 1]].sum() for s in range(world)]
            bounds = shard.rebalance(bounds, times, nb)
            assert bounds[0] == 0 and bounds[-1] == nb and len(bounds) == world + 1
            assert all(bounds[s + 1] > bounds[s] for s in range(world)), bounds
        times = np.array([cost[bounds[s]:bounds[s + 1]].sum() for s in range(world)])
        even = np.array([cost[s * nb // world:(s + 1) * nb // world].sum() for s in range(world)])
        if nb >= 4 * world:
            assert times.max() / times.mean() <= min(even.max() / even.mean(), 1.0 + 2.0 * cost.max() / times.mean()), (nb, world, times)


class _FakeTracer:
    """stands in for api.Tracer in balance_blocks: trace_cam 'takes' the summed cost of the rank's block"""

    def __init__(self, cost, world, rank):
        self.cost, self.world, self.rank = cost, world, rank
        self.bounds = None
        self.last = 0.0

    def shard_layout(self, blocks=False, halo_bands=0):
        self.blocks = blocks

    def shard_config(self, n, r):
        assert (n, r) == (self.world, self.rank)

    def shard_blocks(self, bounds):
        self.bounds = list(bounds)

    def trace_cam(self, cam, w, h, sec, flags):
        self.last = float(self.cost[self.bounds[self.rank]:self.bounds[self.rank + 1]].sum())

    def last_kernel_ms(self):
        return self.last, 0.0


def _balance_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nb = 135
        cost = 1.0 + 0.35 * np.linspace(0, 1, nb)
        tr = _FakeTracer(cost, world, rank)
        poses = np.zeros(2, [("cam", "<f4", (16,)), ("sec_current", "<f4")])
        bounds = shard.balance_blocks(tr, rank, world, dist, poses, 7680, 4320)
        q.put((rank, bounds, tr.bounds))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_block_balancing():
    """balance_blocks over gloo: both ranks time their own block, all-reduce the times and arrive at the same boundaries,
    which split the (known) cost evenly"""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_balance_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = dict()
    for _ in range(2):
        rank, bounds, set_on_tracer = q.get(timeout=5)
        got[rank] = bounds
        assert bounds == set_on_tracer
    assert got[0] == got[1] and got[0][0] == 0 and got[0][2] == 135
    cost = 1.0 + 0.35 * np.linspace(0, 1, 135)
    a, b = cost[:got[0][1]].sum(), cost[got[0][1]:].sum()
    assert abs(a - b) / (a + b) < 0.02, got


def test_block_shape_arithmetic():
    """pwn_block_shape (host only): for every frame height, shard count, halo width and boundary set the blocks partition the
    bands, a shard never sends more halo bands than it has, what it expects from its neighbours is exactly what they send,
    and the rows it will hold contain its own block"""
    from pwnfps_b200 import api
    rng = np.random.default_rng(11)
    for dimy in (1, 31, 32, 33, 360, 1080, 2160, 4320, 4321):
        nbands = (dimy + 31) // 32
        for n in range(1, 9):
            for halo in (0, 1, 2, 3, 9, 200):
                for explicit in (False, True):
                    bounds = None
                    if explicit:
                        cuts = np.sort(rng.integers(0, nbands + 1, n - 1)).tolist()
                        bounds = [0] + cuts + [nbands]
                    shapes = [api.block_shape(dimy, n, r, halo, bounds) for r in range(n)]
                    assert shapes[0][0] == 0 and shapes[-1][0] + shapes[-1][1] == nbands
                    for r, (b0, nb, ht, hb, vlo, vhi) in enumerate(shapes):
                        assert nb >= 0 and 0 <= ht and 0 <= hb and ht + hb <= nb
                        if r + 1 < n:
                            assert b0 + nb == shapes[r + 1][0]
                        if r == 0:
                            assert ht == 0 and vlo == b0
                        else:
                            assert vlo == b0 - shapes[r - 1][3]          # what the shard above sends down
                        if r == n - 1:
                            assert hb == 0 and vhi == b0 + nb
                        else:
                            assert vhi == b0 + nb + shapes[r + 1][2]     # what the shard below sends up
                        assert 0 <= vlo <= b0 and b0 + nb <= vhi <= nbands
                        if bounds is not None:
                            assert (b0, nb) == (bounds[r], bounds[r + 1] - bounds[r])
    # boundaries given for another frame height are ignored: the even split
    assert api.block_shape(4320, 2, 1, 0, [0, 10, 99]) == api.block_shape(4320, 2, 1, 0, None)
    with pytest.raises(api.PwnError):
        api.block_shape(4320, 9, 0)
