#!/usr/bin/env python3
"""tools/ncu_target.py -- short single-GPU workload for ncu: a few frames of the bench workload."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
size = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
lv = levels.default_level()
t = api.Tracer(0)
t.level_upload(lv)
t.level_prepare_render()
t.frame_config(*size)
poses = api.flythrough(lv, 600)
for i in range(0, 600, 600 // n):
    t.trace_cam(poses["cam"][i], size[0], size[1], float(poses["sec_current"][i]), api.PWN_F_BLUR)
# the pose pipeline: trace of frame i + blur of frame i-1 in one launch (k_trace<..., BLURP>)
t.trace_poses(poses[:: 600 // n], size[0], size[1], api.PWN_F_BLUR)
# K3: the frame above to a surface twice as wide and high
t.screen_upscale(size[0], size[1], 2, read=False)
t.sync()
t.close()
