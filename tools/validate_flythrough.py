#!/usr/bin/env python3
"""tools/validate_flythrough.py -- one-off full-size parity run of BASELINE configs[1]:
all 600 poses of the scripted fly-through at 1920x1080, blur on, device (pose pipeline, as the
bench runs it) against the CPU oracle, frame by frame, by checksum; plus the 16 seeded poses of
the 10k-sector maze at 3840x2160 (configs[3]).  Prints one summary line per config."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pwnfps_b200 import api, levels  # noqa: E402
from oracle import orcbind  # noqa: E402
from conftest import oracle_set_level  # noqa: E402


def run(name, lv, poses, w, h, tr, o):
    tr.level_upload(lv)
    tr.level_prepare_render()
    tr.frame_config(w, h)
    oracle_set_level(o, lv)
    t0 = time.time()
    sums = tr.trace_poses(poses, w, h, api.PWN_F_BLUR, want_sums=True)
    t_dev = time.time() - t0
    bad = []
    t0 = time.time()
    for i, pz in enumerate(poses):
        so, zo, _ = o.trace_frame(pz["cam"], float(pz["sec_current"]), w, h)
        sb = o.blur(so, zo, w, h)
        if int(api.checksum_ref(sb, w, h)) != int(sums[i]):
            bad.append(i)
    t_cpu = time.time() - t0
    print("%s: %d frames %dx%d blur on: %d frames differ from the oracle %s (device %.2f s, oracle %.1f s on %d threads)"
          % (name, len(poses), w, h, len(bad), bad[:8], t_dev, t_cpu, o.lib.orc_max_threads()), flush=True)
    return len(bad)


def main():
    tr = api.Tracer(0)
    o = orcbind.Oracle()
    lv = levels.default_level()
    nbad = run("configs[1] default level fly-through", lv, api.flythrough(lv, 600), 1920, 1080, tr, o)
    mz = levels.maze(dim=128, seed=1)
    nbad += run("configs[3] 10k-sector maze", mz, levels.poses_random(mz, 16, seed=1), 3840, 2160, tr, o)
    hall = levels.mirror_hall()
    cam0 = api.cam_spawn(hall)
    px = np.zeros(8, api.POSE_DTYPE)
    for k in range(8):
        px["cam"][k] = api.mat4_roty(cam0, 1.5707964 + 0.4 * k)
        px["sec_current"][k] = 0.25 * k
    nbad += run("configs[2] mirror hall", hall, px, 1920, 1080, tr, o)
    tr.close()
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
