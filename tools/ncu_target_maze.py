#!/usr/bin/env python3
"""tools/ncu_target_maze.py -- short single-GPU workload for ncu: the 10k-sector maze at 3840x2160,
a few of the seeded poses of BASELINE configs[3] (single-frame path, no blur)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

lv = levels.maze(dim=128, seed=1)
t = api.Tracer(0)
t.level_upload(lv)
t.level_prepare_render()
t.frame_config(3840, 2160)
for p in levels.poses_random(lv, 4, seed=1):
    t.trace_cam(p["cam"], 3840, 2160, float(p["sec_current"]), 0)
t.sync()
t.close()
