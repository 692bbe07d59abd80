#!/usr/bin/env python3
"""tools/ncu_target_spheres.py -- short single-GPU workload for ncu: the default level with the 14 game
spheres, a few 1080p frames around the spawn (single-frame path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

lv = levels.game_spheres(levels.default_level(), 0.3)
t = api.Tracer(0)
t.level_upload(lv)
t.level_prepare_render(lv)
t.frame_config(1920, 1080)
poses = api.flythrough(lv, 160)[::40]
for p in poses:
    t.trace_cam(p["cam"], 1920, 1080, float(p["sec_current"]), 0)
t.sync()
t.close()
