#!/usr/bin/env python3
"""tests/tools/make_golden.py -- generates the committed fixtures from the reference itself.

Runs HERE (where /root/reference exists), through oracle/_ref (the reference's own
hot path compiled headless, oracle/make_ref.py).  Writes:

  pwnfps_b200/data/level_default.npz   the default level as the reference's OWN loader
                                       parses level.txt (cells, portal table, spawn)
  tests/golden/host_tables.npz         rcpps/rsqrtps tables of the generating CPU, so a
                                       device can reproduce that CPU anywhere
  tests/golden/kat_util.npz            known answers of util.h: randfs stream, col_ftoint,
                                       v_normalise, mat4_roty
  tests/golden/upscale.npz             screen_upscale of a seeded frame for several (rscale, pitch)
  tests/golden/frames_default.npz      reference frames (sbuf pre-blur + blurred, zbuf, ids)
                                       on the default level: spawn pose 320x240
                                       and three more poses at 160x120
The reference ships no golden vectors; these are its outputs on this host.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refbind  # noqa: E402
from pwnfps_b200 import api  # noqa: E402

REF = "/root/reference"


def main():
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    ref = refbind.Ref("libref_ids.so")
    refb = refbind.Ref("libref.so")
    for r in (ref, refb):
        r.level_load(os.path.join(REF, "level.txt"))
    cells, pm = ref.level_get()
    sx, sz = ref.spawn()
    np.savez_compressed(os.path.join(ROOT, "pwnfps_b200", "data", "level_default.npz"),
                        cells=cells.astype(np.int16), portals=pm.astype(np.int32), spawn=np.array([sx, sz], np.int32))

    rb, sb, rt, st = api.host_math_probe()
    fma_build, exact = api.host_libm_probe()
    assert exact
    np.savez_compressed(os.path.join(gold, "host_tables.npz"), rcp_bits=rb, rsqrt_bits=sb, rcp_tab=rt, rsqrt_tab=st,
                        libm_fma=fma_build)

    # util.h known answers
    seeds = [1, 12345, 0x7FFFFFFF, 415135, 0xDEADBEEF]
    rs = {}
    for s in seeds:
        v, stt = ref.randfs_seq(s, 64)
        rs["randfs_%d" % s] = v
        rs["randst_%d" % s] = stt
    rng = np.random.RandomState(7)
    cols = np.concatenate([rng.uniform(-0.5, 1.5, (200, 4)), rng.uniform(-40, 40, (40, 4)),
                           np.array([[np.nan, np.inf, -np.inf, 1e20], [0.5/255, 1.5/255, 2.5/255, 254.5/255],
                                     [1.0, 0.0, -0.0, 30.0]])]).astype(np.float32)
    packed = np.array([ref.col_ftoint(c) for c in cols], np.uint32)
    vecs = rng.uniform(-3, 3, (100, 4)).astype(np.float32)
    vecs[:, 3] = 0
    norm = np.stack([ref.v_normalise(v) for v in vecs])
    eye = np.eye(4, dtype=np.float32).reshape(16)
    angs = np.linspace(-7, 7, 57).astype(np.float32)
    roty = np.stack([ref.mat4_roty(eye, a) for a in angs])
    np.savez_compressed(os.path.join(gold, "kat_util.npz"), seeds=np.array(seeds, np.uint64), cols=cols, packed=packed,
                        vecs=vecs, norm=norm, angs=angs, roty=roty, **rs)

    # screen_upscale (screen.h:126-149): exact and padded pitches
    frame = np.random.default_rng(11).integers(0, 2**32, (9, 20), dtype=np.uint32)
    ups = {"frame": frame, "cases": np.array([[1, 20], [2, 40], [3, 60], [4, 80], [1, 28], [2, 52], [3, 64]], np.int32)}
    for rs_, pitch in ups["cases"]:
        ups["out_%d_%d" % (rs_, pitch)] = refb.screen_upscale(frame, int(rs_), int(pitch))
    np.savez_compressed(os.path.join(gold, "upscale.npz"), **ups)

    # frames
    out = {}
    cam0 = refbind.spawn_cam(sx, sz)
    poses = [("c1", cam0, 0.0, 320, 240)]
    for i, (yaw, sec) in enumerate([(0.7, 0.25), (2.9, 1.5), (4.4, 3.125)]):
        poses.append(("p%d" % i, ref.mat4_roty(cam0, yaw), sec, 160, 120))
    for name, cam, sec, w, h in poses:
        s, z, ids, _ = ref.trace_frame(cam, sec, w, h, want_ids=True)
        sb_, zb_, _, _ = refb.trace_frame(cam, sec, w, h)
        assert (zb_ == z).all()
        out[name + "_cam"] = np.asarray(cam, np.float32)
        out[name + "_sec"] = np.float32(sec)
        out[name + "_size"] = np.array([w, h], np.int32)
        out[name + "_sbuf"] = s
        out[name + "_sbuf_blur"] = sb_
        out[name + "_zbuf"] = z
        out[name + "_ids"] = ids
        print(name, w, h, hashlib.sha256(s.tobytes()).hexdigest()[:16], hashlib.sha256(sb_.tobytes()).hexdigest()[:16],
              hashlib.sha256(z.tobytes()).hexdigest()[:16])
    np.savez_compressed(os.path.join(gold, "frames_default.npz"), names=np.array([p[0] for p in poses]), **out)
    for f in sorted(os.listdir(gold)):
        print(f, os.path.getsize(os.path.join(gold, f)))


if __name__ == "__main__":
    main()
