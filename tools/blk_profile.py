#!/usr/bin/env python3
"""tools/blk_profile.py -- one GPU: K1 time of each contiguous block of an 8K maze frame (block layout, no blur),
averaged over poses: the load profile a block partition has to balance."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels
tr = api.Tracer(0)
w, h = 7680, 4320
lv = levels.maze(dim=128, seed=1)
poses = levels.poses_random(lv, 64, seed=9)[:16]
tr.level_upload(lv); tr.level_prepare_render(); tr.frame_config(w, h)
for n in (1, 2, 8):
    tr.shard_layout(blocks=True)
    prof = np.zeros((len(poses), n))
    for r in range(n):
        tr.shard_config(n, r)
        for k, p in enumerate(poses):
            tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
            prof[k, r] = tr.last_kernel_ms()[0]
    m = prof.mean(axis=0)
    print("N=%d mean K1 ms per block:" % n, np.round(m, 4), "sum %.4f max %.4f -> balance %.3f; per-pose max/mean: %s" % (m.sum(), m.max(), m.sum() / n / m.max(), np.round((prof.max(axis=1) / prof.mean(axis=1)), 2)))
tr.shard_layout(blocks=False)
for n in (2, 8):
    prof = np.zeros((len(poses), n))
    for r in range(n):
        tr.shard_config(n, r)
        for k, p in enumerate(poses):
            tr.trace_cam(p["cam"], w, h, float(p["sec_current"]), 0)
            prof[k, r] = tr.last_kernel_ms()[0]
    m = prof.mean(axis=0)
    print("interleaved N=%d mean K1 ms per shard:" % n, np.round(m, 4), "sum %.4f max %.4f -> balance %.3f" % (m.sum(), m.max(), m.sum() / n / m.max()))
