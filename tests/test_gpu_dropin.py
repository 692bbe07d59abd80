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


def test_pwnbench_c_host(tracer, default_level, tmp_path):
    """The pure-C headless host (host/pwnbench.c: level.txt loader + camera script + C ABI) renders
    the frames the Python client renders."""
    import os
    import re
    import subprocess
    from pwnfps_b200 import build
    exe = build.BENCH_BIN
    if not os.path.isfile(exe):
        pytest.skip("pwnbench not built")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    w, h, n = 320, 200, 120
    out = subprocess.run([exe, p, str(n), str(w), str(h), "1", "0"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    sums = {int(m.group(1)): int(m.group(2), 16) for m in re.finditer(r"frame\s+(\d+) fnv1a64 ([0-9a-f]{16})", out.stdout)}
    assert 0 in sums and 100 in sums and "Mrays/s" in out.stdout
    tracer.level_upload(default_level)
    tracer.level_prepare_render()
    tracer.frame_config(w, h)
    poses = api.flythrough(default_level, n)

    def fnv(a):
        hsh = 0xcbf29ce484222325
        for v in a.reshape(-1).tolist():
            hsh = ((hsh ^ v) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
        return hsh

    for f in (0, 100):
        tracer.trace_cam(poses["cam"][f], w, h, float(poses["sec_current"][f]), api.PWN_F_BLUR)
        s, _, _ = tracer.read_frame(want_z=False)
        assert fnv(s) == sums[f], f
