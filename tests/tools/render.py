#!/usr/bin/env python3
"""tests/tools/render.py -- render poses of a level on the GPU and dump them for visual diffing.

  python tests/tools/render.py out_prefix [--level default|hall|maze|PATH] [--size 640x360] [--frames 4]
                         [--spheres] [--check]

Writes out_prefix_NNN.ppm (host/pwn_image.c) and .png (zlib, no dependencies); with --check the
same pose is traced by the CPU oracle (tests' checker) and out_prefix_NNN_diff.png shows
|device - oracle| x 64 per colour byte (black = identical)."""
import argparse
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pwnfps_b200 import api, levels  # noqa: E402


def save_png(path, sbuf):
    h, w = sbuf.shape
    rgb = np.empty((h, w, 3), np.uint8)
    rgb[..., 0] = (sbuf >> 16) & 255
    rgb[..., 1] = (sbuf >> 8) & 255
    rgb[..., 2] = sbuf & 255
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgb.reshape(h, w * 3)], axis=1).tobytes()

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("prefix")
    ap.add_argument("--level", default="default")
    ap.add_argument("--size", default="640x360")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--spheres", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--no-blur", action="store_true")
    a = ap.parse_args()
    w, h = (int(v) for v in a.size.split("x"))
    lv = {"default": levels.default_level, "hall": levels.mirror_hall,
          "maze": lambda: levels.maze(dim=128, seed=1)}.get(a.level, lambda: api.level_load(a.level))()
    game = api.Game(lv) if a.spheres else None
    tr = api.Tracer(0)
    tr.level_upload(lv)
    tr.frame_config(w, h)
    poses = api.flythrough(lv, 60 * a.frames)[::60]
    orc = None
    if a.check:
        from oracle import orcbind
        orc = orcbind.Oracle()
        orc.level_set(lv.cells, lv.portals_i32())
    flags = 0 if a.no_blur else api.PWN_F_BLUR
    for i, p in enumerate(poses):
        if game is not None:
            for k in range(60):
                game.tick(float(p["sec_current"]) + k / 60.0, 1.0 / 60.0)
            lv.spheres = game.spheres()
        tr.level_prepare_render(lv)
        s = tr.trace_screen_centred(lv, 0, 0, w, h, p["cam"], float(p["sec_current"]), blur=bool(flags))
        name = "%s_%03d" % (a.prefix, i)
        api.save_ppm(name + ".ppm", s)
        save_png(name + ".png", s)
        if orc is not None:
            if len(lv.spheres):
                sp = np.zeros((len(lv.spheres), 10), np.float32)
                sp[:, 0], sp[:, 1], sp[:, 2:6], sp[:, 6:10] = lv.spheres["r"], lv.spheres["refl"], lv.spheres["pos"], lv.spheres["col"]
                orc.spheres_set(sp)
            so, _, _ = orc.trace_frame(p["cam"], float(p["sec_current"]), w, h, blur=bool(flags))
            d = api.frame_diff(s, so)
            save_png(name + "_diff.png", d)
            print(name, "differing pixels vs oracle:", int(np.count_nonzero(d)))
        else:
            print(name)


if __name__ == "__main__":
    main()
