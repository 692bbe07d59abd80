#!/usr/bin/env python3
"""tools/parts_time.py -- cost of the per-frame sphere binning + upload (pwn_parts_upload =
level_prepare_render, level.h:64-81) at 14, 1000 and OBJ_MAX = 10000 spheres on the 10k-sector maze,
beside K1 on the same frame."""
import sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pwnfps_b200 import api, levels
lv = levels.scatter_spheres(levels.maze(dim=128, seed=1), 10000, seed=3)
t = api.Tracer(0)
t.level_upload(lv)
t.frame_config(1920, 1080)
for n in (14, 1000, 10000):
    lv2 = levels.maze(dim=128, seed=1); lv2.spheres = lv.spheres[:n].copy()
    t.level_prepare_render(lv2)
    t0 = time.perf_counter()
    for _ in range(50):
        t.level_prepare_render(lv2)
    t.sync()
    dt = (time.perf_counter() - t0) / 50
    poses = levels.poses_random(lv2, 4, seed=1)
    t.trace_cam(poses["cam"][0], 1920, 1080, 0.0, 0)
    k1, _ = t.last_kernel_ms()
    print("spheres %5d: parts_upload %.3f ms per frame, K1 1080p %.3f ms" % (n, dt * 1e3, k1))
