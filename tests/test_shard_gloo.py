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
