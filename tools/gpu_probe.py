#!/usr/bin/env python3
"""tools/gpu_probe.py -- first look at a GPU box: host CPU, math tables, kernel times."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

print(subprocess.run("lscpu | grep -E 'Model name|^CPU\\(s\\)|Thread|Socket'; nproc; nvidia-smi -L", shell=True, capture_output=True, text=True).stdout)
rb, sb, rt, st = api.host_math_probe()
print("host math bits", rb, sb)
g = np.load(os.path.join(ROOT, "tests/golden/host_tables.npz"))
print("tables equal to golden host:", rb == int(g["rcp_bits"]) and (rt == g["rcp_tab"]).all(), sb == int(g["rsqrt_bits"]) and (st == g["rsqrt_tab"]).all())

t = api.Tracer(0)
for name, lv in (("default", levels.default_level()), ("hall", levels.mirror_hall()), ("maze", levels.maze())):
    t.level_upload(lv)
    t.level_prepare_render()
    poses = api.flythrough(lv, 240)
    for (w, h) in [(1920, 1080), (3840, 2160), (7680, 4320)]:
        t.frame_config(w, h)
        for flags in (0, api.PWN_F_BLUR):
            ms = []
            for i in (0, 60, 120, 180):
                t.time_frames(poses["cam"][i], w, h, float(poses["sec_current"][i]), flags, 2)
                ms.append(t.time_frames(poses["cam"][i], w, h, float(poses["sec_current"][i]), flags, 5) / 5)
            ms = np.array(ms)
            print("%-8s %dx%d blur=%d: %.3f ms/frame (min %.3f max %.3f) -> %.0f Mrays/s" % (
                name, w, h, flags, ms.mean(), ms.min(), ms.max(), w * h / ms.mean() / 1e3))
