"""The drop-in proof (-m gpu): the REFERENCE HOST -- its own defs.h / util.h / level.h loader and
the headless stand-in for main.c (oracle/ref_harness.c) -- built with pwn_screen.h in place of
trace.h + screen.h and linked against libpwntrace.so (oracle/_ref/libref_shim*.so), must render
what the unmodified reference renders."""
import numpy as np
import pytest

from conftest import rgb_diff
from pwnfps_b200 import api, levels

pytestmark = pytest.mark.gpu


def test_reference_host_through_the_shim(refs, default_level, tmp_path):
    from oracle import refbind
    if not refbind.available("libref_shim.so"):
        pytest.skip("libref_shim.so not built")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    pairs = [(refs("libref_shim_noblur.so"), refs("libref_noblur.so")), (refs("libref_shim.so"), refs("libref.so"))]
    lv = levels.game_spheres(levels.default_level(), sec=0.3)
    cam0 = api.cam_spawn(default_level)
    for shim, ref in pairs:
        for r in (shim, ref):
            r.level_load(p)          # the reference's own loader in both builds
            r.objs_clear()
        for with_spheres in (False, True):
            if with_spheres:
                for r in (shim, ref):
                    for sp in lv.spheres:
                        r.obj_sphere(sp["r"], sp["refl"], sp["pos"][0], sp["pos"][1], sp["pos"][2],
                                     sp["col"][0], sp["col"][1], sp["col"][2])
            for yaw, sec, (w, h) in [(0.0, 0.3, (320, 240)), (1.9, 0.3, (640, 360))]:
                cam = api.mat4_roty(cam0, yaw)
                s, z, _, _ = shim.trace_frame(cam, sec, w, h)
                rs, rz, _, _ = ref.trace_frame(cam, sec, w, h)
                assert (z.view(np.uint32) == rz.view(np.uint32)).all()
                d = rgb_diff(s, rs)
                # 1 LSB on >= 99.9% pre-blur; a blurred pixel averages 4 gathers of such pixels
                assert (d <= 1).mean() >= (0.995 if shim.has_blur else 0.999)
        for r in (shim, ref):
            r.objs_clear()
