#!/usr/bin/env python3
"""tools/sanitize_target.py -- small workload for compute-sanitizer: every kernel variant once."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

t = api.Tracer(0)
for lv in (levels.default_level(), levels.open_border(), levels.maze()):
    for spheres in (False, True):
        if spheres:
            lv = levels.game_spheres(lv, 0.3)
        t.level_upload(lv)
        t.level_prepare_render()
        w, h = 203, 131
        t.frame_config(w, h, want_ids=True)
        poses = api.flythrough(lv, 8)
        for flags in (0, api.PWN_F_BLUR, api.PWN_F_IDS):
            t.trace_cam(poses["cam"][3], w, h, 0.3, flags)
        out = np.zeros((8, h, w), np.uint32)
        t.trace_poses(poses, w, h, api.PWN_F_BLUR, out=out, want_sums=True)
        t.trace_poses(poses, w, h, 0, want_sums=True)
        for n in (2, 3):
            for r in range(n):
                t.shard_config(n, r)
                t.trace_cam(poses["cam"][1], w, h, 0.0, 0)
        t.shard_config(1, 0)
        t.read_frame(want_ids=True)
t.debug_libm(np.linspace(-200, 200, 1000).astype(np.float32))
t.close()
print("sanitize target done")
