#!/usr/bin/env python3
"""tests/tools/validate_flythrough.py -- one-off full-size parity run of BASELINE configs[1]:
all 600 poses of the scripted fly-through at 1920x1080, blur on, device (pose pipeline, as the
bench runs it) against the CPU oracle, frame by frame, by checksum; plus the 16 seeded poses of
the 10k-sector maze at 3840x2160 (configs[3]), 8 poses of its 64-pose batch at 7680x4320 (configs[4]) and the
mirror hall (configs[2]).  Prints one summary line per config."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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


def vs_reference(tr, lv, poses, w, h):
    """north_star acceptance against the REFERENCE ITSELF (oracle/_ref, shipped -ffast-math flags):
    zbuf bit-exact; RGB within 1 LSB per channel on >= 99.9% of pixels."""
    import tempfile
    from oracle import refbind
    from conftest import rgb_diff
    if not refbind.available("libref.so"):
        print("oracle/_ref not built: reference comparison skipped")
        return 0
    refs = {False: refbind.Ref("libref_noblur.so"), True: refbind.Ref("libref.so")}
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "level.txt")
        lv.save_txt(p)
        devnull, saved = os.open(os.devnull, os.O_WRONLY), os.dup(1)
        os.dup2(devnull, 1)
        try:
            for r in refs.values():
                r.level_load(p)
        finally:
            os.dup2(saved, 1)
    tr.level_upload(lv)
    tr.level_prepare_render()
    tr.frame_config(w, h)
    worst = {False: [1.0, 0, 0], True: [1.0, 0, 0]}
    zbad = 0
    for pz in poses:
        for blur in (False, True):
            tr.frame_clear()
            tr.trace_cam(pz["cam"], w, h, float(pz["sec_current"]), api.PWN_F_BLUR if blur else 0)
            s, z, _ = tr.read_frame()
            rs, rz, _, _ = refs[blur].trace_frame(pz["cam"], float(pz["sec_current"]), w, h)
            dd = rgb_diff(s, rs)
            worst[blur][0] = min(worst[blur][0], float((dd <= 1).mean()))
            worst[blur][1] += int((dd > 0).sum())
            worst[blur][2] = max(worst[blur][2], int(dd.max()))
            zbad += int((z.view(np.uint32) != rz.view(np.uint32)).sum())
    n = len(poses) * w * h
    for blur in (False, True):
        print("vs reference (shipped flags), %d frames %dx%d, blur %s: worst frame has %.6f%% of pixels within 1 LSB; "
              "%d of %d pixels differ at all (%.2e), max channel delta %d"
              % (len(poses), w, h, "on" if blur else "off", 100 * worst[blur][0], worst[blur][1], n, worst[blur][1] / n, worst[blur][2]), flush=True)
    print("vs reference: zbuf words that differ: %d" % zbad, flush=True)
    return 0 if (zbad == 0 and worst[False][0] >= 0.999 and worst[True][0] >= 0.999) else 1


def main():
    tr = api.Tracer(0)
    o = orcbind.Oracle()
    lv = levels.default_level()
    nbad = run("configs[1] default level fly-through", lv, api.flythrough(lv, 600), 1920, 1080, tr, o)
    mz = levels.maze(dim=128, seed=1)
    nbad += run("configs[3] 10k-sector maze", mz, levels.poses_random(mz, 16, seed=1), 3840, 2160, tr, o)
    # configs[4]: the same maze at 7680x4320, every 8th pose of the 64-pose batch the multi-GPU bench shards (seed 9)
    nbad += run("configs[4] 10k-sector maze 8K", mz, levels.poses_random(mz, 64, seed=9)[::8], 7680, 4320, tr, o)
    hall = levels.mirror_hall()
    cam0 = api.cam_spawn(hall)
    px = np.zeros(8, api.POSE_DTYPE)
    for k in range(8):
        px["cam"][k] = api.mat4_roty(cam0, 1.5707964 + 0.4 * k)
        px["sec_current"][k] = 0.25 * k
    nbad += run("configs[2] mirror hall", hall, px, 1920, 1080, tr, o)
    lv = levels.default_level()
    nbad += vs_reference(tr, lv, api.flythrough(lv, 600)[::30], 1920, 1080)
    tr.close()
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
