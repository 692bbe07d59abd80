#!/usr/bin/env python3
"""tools/ab_time.py -- A/B timing of an experimental libpwntrace build (PWN_LIBPWNTRACE=path).
Device-timed pose batches: default level 1080p, 14 game spheres 1080p, 60 scattered spheres on the
maze at 4K, maze 4K.  Prints ms per frame and a checksum of checksums per workload."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pwnfps_b200 import api, levels  # noqa: E402


def main():
    tr = api.Tracer(0)
    stream = torch.cuda.ExternalStream(tr.stream_ptr(), device=torch.device("cuda", 0))
    out = []
    lvd = levels.default_level()
    mz = levels.maze(dim=128, seed=1)
    work = [("default 1080p", lvd, (1920, 1080), api.flythrough(lvd, 600)[::5].copy()),
            ("game spheres 1080p", levels.game_spheres(levels.default_level(), 0.3), (1920, 1080), api.flythrough(lvd, 160)[::4].copy()),
            ("maze+60 spheres 4K", levels.scatter_spheres(levels.maze(dim=128, seed=1), 60, seed=3), (3840, 2160), levels.poses_random(mz, 16, seed=1)),
            ("maze 4K", mz, (3840, 2160), levels.poses_random(mz, 16, seed=1))]
    for name, lv, (w, h), poses in work:
        tr.level_upload(lv)
        tr.level_prepare_render(lv)
        tr.frame_config(w, h)
        sums = tr.trace_poses(poses, w, h, api.PWN_F_BLUR, want_sums=True)
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            tr.trace_poses(poses, w, h, api.PWN_F_BLUR)
            b.record(stream)
            b.synchronize()
            best = min(best, a.elapsed_time(b) / len(poses))
        out.append("%-22s %.4f ms/frame  %8.1f Mrays/s  sum %016x" % (name, best, w * h / best / 1e3, int(np.bitwise_xor.reduce(sums))))
    # K1 and K2 as single launches (what bench.py's roofline times): per-launch events over 150 fly-through poses
    tr.level_upload(lvd)
    tr.level_prepare_render(lvd)
    tr.frame_config(1920, 1080)
    fl = api.flythrough(lvd, 600)[::4]
    for rep in range(2):
        k1 = k2 = 0.0
        for i in range(len(fl)):
            tr.trace_cam(fl["cam"][i], 1920, 1080, float(fl["sec_current"][i]), api.PWN_F_BLUR)
            a, b = tr.last_kernel_ms()
            k1 += a
            k2 += b
    out.append("single launches 1080p  K1 %.4f ms  K2 %.4f ms" % (k1 / len(fl), k2 / len(fl)))
    print(os.environ.get("PWN_LIBPWNTRACE", "in-tree"))
    print("\n".join(out))


if __name__ == "__main__":
    main()
