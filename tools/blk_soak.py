#!/usr/bin/env python3
"""tools/blk_soak.py NFRAMES JOIN_EVERY -- under torchrun: many sharded frames of the block layout back to back without a
host sync; reports time-outs and where the error word was first seen (synced every CHECK frames if given)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels, shard
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nframes, join_every = int(sys.argv[1]), int(sys.argv[2])
check = int(sys.argv[3]) if len(sys.argv) > 3 else 0
tr = api.Tracer(local)
w, h = 7680, 4320
lv = levels.maze(dim=128, seed=1)
poses = levels.poses_random(lv, 64, seed=9)
tr.level_upload(lv); tr.level_prepare_render(); tr.frame_config(w, h)
tr.shard_config(world, rank); tr.shard_layout(blocks=True)
if "balance" in sys.argv:
    shard.balance_blocks(tr, rank, world, dist, poses[::8], w, h)
dist.barrier()
bases = shard.open_all_peers(tr, rank, world, dist)
t0 = time.time()
for i in range(nframes):
    p = poses[i % 64]
    tr.trace_shard_blur(p["cam"], w, h, float(p["sec_current"]))
    if join_every and (i + 1) % join_every == 0:
        tr.join()
    if check and (i + 1) % check == 0:
        f = tr.debug_flags(40)
        if f[16]:
            print("rank %d: error word set by frame %d (%.2f s)" % (rank, i + 1, time.time() - t0), flush=True)
            break
if "dump" in sys.argv:
    time.sleep(1.0)
    f = tr.debug_flags(64 + 140)
    g = tr.debug_flags(64 + 140, alt=True)
    def rle(v):
        out, prev, cnt = [], None, 0
        for x in v:
            if x == prev: cnt += 1
            else:
                if prev is not None: out.append("%gx%d" % (prev, cnt))
                prev, cnt = x, 1
        out.append("%gx%d" % (prev, cnt))
        return " ".join(out)
    print("rank %d after 1 s: A %s B %s err %d defer %d | group0 %s | group1 %s" % (rank, f[0:2], f[8:10], f[16], f[32], rle((f[64:64+135] / 245760).round(3).tolist()), rle((g[64:64+135] / 245760).round(3).tolist())), flush=True)
try:
    tr.sync()
    print("rank %d: %d frames ok in %.2f s" % (rank, nframes, time.time() - t0), flush=True)
except Exception as e:
    print("rank %d: %s (%.2f s)" % (rank, str(e)[:90], time.time() - t0), flush=True)
dist.barrier()
tr.frame_peers(None)
tr.close()
dist.destroy_process_group()
