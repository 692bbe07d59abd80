#!/usr/bin/env python3
"""tools/ncu_target_r2.py -- short single-GPU workload for ncu, one K1 launch per workload family:
default level 1080p, mirror hall 1080p, 10k-sector maze 8K (the headline), maze + 3000 spheres 4K,
the game's 14 spheres 1080p; then two frames of the 8K pose pipeline (k_trace<..., BLURP> + k_blur)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402

t = api.Tracer(0)
dl = levels.default_level()
hall = levels.mirror_hall()
mz = levels.maze(dim=128, seed=1)
mzs = levels.scatter_spheres(levels.maze(dim=128, seed=1), 3000, seed=2)
g = levels.game_spheres(levels.default_level(), 0.3)
fly = api.flythrough(dl, 600)
batch = levels.poses_random(mz, 64, seed=9)
work = [(dl, (1920, 1080), fly["cam"][300], float(fly["sec_current"][300])),
        (hall, (1920, 1080), api.mat4_roty(api.cam_spawn(hall), 1.5707964 + 0.4), 0.25),
        (mz, (7680, 4320), batch["cam"][5], float(batch["sec_current"][5])),
        (mzs, (3840, 2160), batch["cam"][5], float(batch["sec_current"][5])),
        (g, (1920, 1080), api.flythrough(g, 160)["cam"][40], 0.3)]
for lv, (w, h), cam, sec in work:
    t.level_upload(lv)
    t.level_prepare_render(lv)
    t.frame_config(w, h)
    t.trace_cam(cam, w, h, sec, 0)
    t.sync()
t.level_upload(mz)
t.level_prepare_render(mz)
t.frame_config(7680, 4320)
t.trace_poses(batch[5:7], 7680, 4320, api.PWN_F_BLUR)
t.sync()
t.close()
