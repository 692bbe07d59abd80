#!/usr/bin/env python3
"""tools/blk_time.py -- under torchrun: the sharded 8K maze frame WITH the blur pass, interleaved vs block layout;
device-timed per frame (events on the library's stream, max over ranks), frames compared with the one-GPU frames."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels, shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
tr = api.Tracer(local)
stream = torch.cuda.ExternalStream(tr.stream_ptr(), device=torch.device("cuda", local))
w, h = 7680, 4320
lv = levels.maze(dim=128, seed=1)
poses = levels.poses_random(lv, 64, seed=9)[:int(sys.argv[1]) if len(sys.argv) > 1 else 32]
tr.level_upload(lv)
tr.level_prepare_render()
tr.frame_config(w, h)


def allmax(v):
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


want = tr.trace_poses(poses[:4], w, h, api.PWN_F_BLUR, want_sums=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tr.trace_poses(poses, w, h, api.PWN_F_BLUR)
a.record(stream)
tr.trace_poses(poses, w, h, api.PWN_F_BLUR)
b.record(stream)
b.synchronize()
one = a.elapsed_time(b) / len(poses)
tr.shard_config(world, rank)
for layout in (sys.argv[2].split(",") if len(sys.argv) > 2 else ("interleaved", "blocks", "blocks balanced")):
    tr.shard_layout(blocks=(layout != "interleaved"))
    if layout == "blocks balanced":
        print("boundaries", shard.balance_blocks(tr, rank, world, dist, poses[::4], w, h)) if rank == 0 else shard.balance_blocks(tr, rank, world, dist, poses[::4], w, h)
    dist.barrier()
    bases = shard.open_all_peers(tr, rank, world, dist)
    ok = True
    for k in range(4):
        tr.trace_shard_blur(poses[k]["cam"], w, h, float(poses[k]["sec_current"]))
        tr.sync()
        dist.barrier()
        if rank == 0:
            s0, _, _ = tr.read_frame(want_z=False)
            ok = ok and int(api.checksum_ref(s0, w, h)) == int(want[k])
        dist.barrier()
    best = 1e9
    k1 = 0.0
    for rep in range(3):
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        a.record(stream)
        th = time.perf_counter()
        for p in poses:
            tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]))
        th = (time.perf_counter() - th) / len(poses) * 1e3
        tr.join()
        b.record(stream)
        b.synchronize()
        best = min(best, allmax(a.elapsed_time(b)) / len(poses))
        if rep == 2 and rank == 0:
            print("   host submission %.4f ms/frame" % th)
        k1 = allmax(tr.last_kernel_ms()[0])
    # per rank: K1 and the rest (barrier A .. barrier B) of single frames with a host sync in between
    k1s, rests = [], []
    for p in poses[:8]:
        dist.barrier()
        tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]))
        a1, b1 = tr.last_kernel_ms()
        k1s.append(a1)
        rests.append(b1)
    allk = [None] * world
    dist.all_gather_object(allk, (float(np.mean(k1s)), float(np.mean(rests))))
    if rank == 0:
        print("   per rank (K1 ms, rest ms):", [(round(a1, 3), round(b1, 3)) for a1, b1 in allk])
    to = tr.peers_timed_out()
    dist.barrier()
    shard.close_all_peers(tr, bases)
    if rank == 0:
        print("%-12s N=%d: %.4f ms/frame (one GPU %.4f: x%.2f), last frame K1 max over ranks %.4f ms, frames equal one GPU: %s, timed out: %s"
              % (layout, world, best, one, one / best, k1, ok, to))
tr.shard_blocks(None)
tr.shard_layout(blocks=False)
tr.shard_config(1, 0)
tr.close()
dist.barrier()
dist.destroy_process_group()
