"""Pins the CPU restatement (oracle/pwn_oracle.c) against the reference's own code compiled
headless (oracle/_ref) and against the committed fixtures generated from it.

Bars:  ids and zbuf bit-exact vs the shipped-flags build, sbuf within 1 LSB on < 1e-5 of pixels
(gcc's fast-math code generation of the colour arithmetic, DESIGN.md "Oracle"); the restatement
built without the one mirrored fast-math folding (randfs) is bit-exact, sbuf included, vs the same
reference sources built without -ffast-math."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, oracle_set_level, pack_edge_cols, rgb_diff
from pwnfps_b200 import api, levels

IDF = ("cell", "ldir", "kind", "nstep", "nportal")


def same_tables_as_golden():
    g = np.load(os.path.join(GOLDEN, "host_tables.npz"))
    rb, sb, rt, st = api.host_math_probe()
    return rb == int(g["rcp_bits"]) and sb == int(g["rsqrt_bits"]) and (rt == g["rcp_tab"]).all() and (st == g["rsqrt_tab"]).all()


def test_util_kats(oracle):
    """RNG / pack known answers dumped from the reference build (exact integer parts)."""
    k = np.load(os.path.join(GOLDEN, "kat_util.npz"))
    for s in k["seeds"]:
        v, st = oracle.randfs_seq(int(s), 64)
        assert (st == k["randst_%d" % s]).all()
        assert (v.view(np.uint32) == k["randfs_%d" % s].view(np.uint32)).all()
    got = np.array([oracle.col_ftoint(c) for c in k["cols"]], np.uint32)
    assert (got == k["packed"]).all()
    for a, want in zip(k["angs"], k["roty"]):
        got = api.mat4_roty(np.eye(4, dtype=np.float32).reshape(16), a)
        assert np.allclose(got, want, atol=1e-6)


def test_upscale_golden(oracle):
    """screen_upscale (screen.h:126-149) fixtures from the reference, incl. the padded-pitch drift"""
    g = np.load(os.path.join(GOLDEN, "upscale.npz"))
    for rs, pitch in g["cases"]:
        got = oracle.upscale(g["frame"], int(rs), int(pitch))
        assert np.array_equal(got, g["out_%d_%d" % (rs, pitch)]), (rs, pitch)


def test_upscale_vs_reference(oracle, refs):
    ref = refs("libref.so")
    rng = np.random.default_rng(3)
    for w, h, rs, pitch in [(64, 32, 1, None), (64, 32, 2, None), (33, 7, 3, None), (64, 16, 4, None),
                            (20, 6, 2, 48), (20, 6, 5, 128), (1, 1, 7, 9)]:
        f = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
        assert np.array_equal(oracle.upscale(f, rs, pitch), ref.screen_upscale(f, rs, pitch)), (w, h, rs, pitch)
    # the drop-in shim's host-side screen_upscale (pwnfps_b200/host/pwn_screen.h) writes the same words
    from oracle import refbind
    if refbind.available("libref_shim.so"):
        shim = refs("libref_shim.so")
        for w, h, rs, pitch in [(64, 32, 2, None), (33, 7, 3, None), (20, 6, 2, 48), (20, 6, 5, 128)]:
            f = rng.integers(0, 2**32, (h, w), dtype=np.uint32)
            assert np.array_equal(shim.screen_upscale(f, rs, pitch), ref.screen_upscale(f, rs, pitch)), (w, h, rs, pitch)


def test_col_ftoint_edges_vs_reference(oracle, refs):
    """the pack on ties, saturation, NaN, inf and out-of-int-range values: restatement = reference"""
    ref = refs("libref_noblur.so")
    cols = pack_edge_cols()
    with np.errstate(all="ignore"):
        got = np.array([oracle.col_ftoint(c) for c in cols], np.uint32)
        want = np.array([ref.col_ftoint(c) for c in cols], np.uint32)
    assert (got == want).all(), cols[np.flatnonzero(got != want)[:3]]
    assert len(np.unique(want)) > 500


def test_randfs_exhaustive_vs_reference(oracle, refs):
    """every residue r in 0..3758 gives the reference's float"""
    ref = refs("libref_noblur.so")
    # seeds that produce each residue: walk the LCG long enough to see them all
    seen = {}
    v_o, st_o = oracle.randfs_seq(99, 40000)
    v_r, st_r = ref.randfs_seq(99, 40000)
    assert (st_o == st_r).all()
    assert (v_o.view(np.uint32) == v_r.view(np.uint32)).all()
    for s in st_o:
        seen[int(s) % 3759] = 1
    assert len(seen) == 3759


@pytest.mark.skipif(not same_tables_as_golden(), reason="this CPU's rcpps/rsqrtps differ from the fixture's")
def test_golden_frames(oracle, default_level):
    g = np.load(os.path.join(GOLDEN, "frames_default.npz"))
    oracle_set_level(oracle, default_level)
    for name in g["names"]:
        w, h = g[name + "_size"]
        s, z, ids = oracle.trace_frame(g[name + "_cam"], float(g[name + "_sec"]), int(w), int(h), want_ids=True)
        assert (z.view(np.uint32) == g[name + "_zbuf"].view(np.uint32)).all()
        for f in IDF:
            assert (ids[f] == g[name + "_ids"][f]).all(), f
        d = rgb_diff(s, g[name + "_sbuf"])
        assert d.max() <= 1 and (d > 0).mean() < 1e-4
        if d.max() == 0:
            b = oracle.blur(s, z, int(w), int(h))
            assert (b == g[name + "_sbuf_blur"]).all()


def _compare(oracle, ref, ref_strict, cam, sec, w, h, dim_ref=64):
    """ref: an instrumented build with the shipped flags; ref_strict: the same sources without -ffast-math
    (instrumented or not).  Against ref: zbuf and the primary segment's ids bit-exact; bounce-segment ids
    and sbuf may differ on a few grazing pixels per million (gcc's fast-math code generation moves a
    jittered bounce ray by an ulp) -- counted and bounded.  Against ref_strict the strict build of the
    restatement (oracle.twin) is exact: sbuf, zbuf and, when ref_strict is instrumented, every id."""
    s, z, ids, _ = ref.trace_frame(cam, sec, w, h, want_ids=True)
    so, zo, io = oracle.trace_frame(cam, sec, w, h, want_ids=True)
    assert (z.view(np.uint32) == zo.view(np.uint32)).all()
    for f in IDF:
        assert (ids[..., 0][f] == io[..., 0][f]).all(), f
    graze = np.zeros((h, w), bool)
    for seg in (1, 2):
        for f in IDF:
            graze |= ids[..., seg][f] != io[..., seg][f]
    d = rgb_diff(s, so)
    budget = max(2, int(1e-5 * w * h))
    assert graze.sum() <= budget, int(graze.sum())
    assert (d[~graze] <= 1).all() and ((d > 0) & ~graze).sum() <= budget
    if ref_strict is not None:
        want_ids = ref_strict.has_ids
        ss, zs, ids_s, _ = ref_strict.trace_frame(cam, sec, w, h, want_ids=want_ids)
        st, zt, it = oracle.twin.trace_frame(cam, sec, w, h, want_ids=want_ids)
        assert (ss == st).all() and (zs.view(np.uint32) == zt.view(np.uint32)).all()
        if want_ids:
            for f in IDF:
                assert (ids_s[f] == it[f]).all(), f
        assert (zs.view(np.uint32) == zo.view(np.uint32)).all() and (rgb_diff(ss, so) > 0).sum() <= budget
    return s, z, so, zo


def _strict(refs):
    """the reference sources without -ffast-math: the instrumented build when it exists (ids of all segments)"""
    from oracle import refbind
    return refs("libref_ids_strict.so") if refbind.available("libref_ids_strict.so") else refs("libref_strict.so")


def _wide_strict(refs):
    from oracle import refbind
    return refs("libref_wide_ids_strict.so") if refbind.available("libref_wide_ids_strict.so") else None


def _compare_ids_only(oracle, ref, cam, sec, w, h):
    s, z, ids, _ = ref.trace_frame(cam, sec, w, h, want_ids=True)
    so, zo, io = oracle.trace_frame(cam, sec, w, h, want_ids=True)
    assert (z.view(np.uint32) == zo.view(np.uint32)).all()
    for f in IDF:
        assert (ids[f] == io[f]).all(), f
    return so, zo, io


def test_default_level_vs_reference(oracle, refs, default_level, tmp_path):
    ref, strict, refb = refs("libref_ids.so"), _strict(refs), refs("libref.so")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    for r in (ref, strict, refb):
        r.level_load(p)
    oracle_set_level(oracle, default_level)
    cam0 = api.cam_spawn(default_level)
    for yaw, sec, (w, h) in [(0.0, 0.0, (320, 240)), (0.9, 0.7, (333, 77)), (2.2, 1.3, (640, 360)),
                             (3.9, 5.0, (160, 300)), (5.1, 2.0, (642, 33))]:
        cam = api.mat4_roty(cam0, yaw)
        s, z, so, zo = _compare(oracle, ref, strict, cam, sec, w, h)
        # blur pass: same input -> bit-identical output (integer pavgb + LCG)
        if w % 4:
            continue   # the reference's blur needs rwidth % 4 == 0 (aligned store, screen.h:88,117)
        sb, _, _, _ = refb.trace_frame(cam, sec, w, h)
        if (s == so).all():
            assert (oracle.blur(so, zo, w, h) == sb).all()


def test_compile_time_switches_vs_reference(oracle, refs, default_level, tmp_path):
    """REFLECT 0 / 1 and REFLECT_BLUR off (defs.h:6-8), each against the reference built that way
    (strict flags: bit-exact)"""
    from oracle import refbind
    if not refbind.available("libref_reflect1.so"):
        pytest.skip("oracle/_ref switch variants not built")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    oracle_set_level(oracle, default_level)
    cam0 = api.cam_spawn(default_level)
    try:
        for variant, reflect, rblur in (("libref_reflect0.so", 0, True), ("libref_reflect1.so", 1, True),
                                        ("libref_norblur.so", 2, False)):
            ref = refs(variant)
            ref.level_load(p)
            oracle.config(reflect, rblur)
            for yaw, sec, (w, h) in [(0.0, 0.0, (320, 240)), (2.2, 1.3, (333, 77))]:
                cam = api.mat4_roty(cam0, yaw)
                sr, zr, _, _ = ref.trace_frame(cam, sec, w, h)
                so, zo, _ = oracle.twin.trace_frame(cam, sec, w, h)      # strict flags on both sides
                assert (zr.view(np.uint32) == zo.view(np.uint32)).all(), variant
                assert (sr == so).all(), (variant, int((sr != so).sum()))
    finally:
        oracle.config(2, True)


def test_lowercase_portal_quirk_vs_reference(oracle, refs, default_level, tmp_path):
    """level.txt:5 has a lower-case 'm': the cell joins pair M and then BECOMES 'N' (level.h:144-161), so
    M at (4,3) leads into a cell that answers as the unpaired N.  Poses that look straight at both."""
    ref, strict = refs("libref_ids.so"), _strict(refs)
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    for r in (ref, strict):
        r.level_load(p)
    cells, pm = ref.level_get()
    assert (cells == default_level.cells).all() and cells[4, 4] == ord("N")
    assert tuple(pm[12][:4]) == (4, 3, 4, 4) and pm[13][2] == -1
    oracle_set_level(oracle, default_level)
    saw = 0
    for (x, z, yaw) in [(4.5, 2.5, 0.0), (4.3, 2.2, 0.3), (4.5, 5.5, 3.1415927), (5.5, 2.5, 5.9)]:
        cam = api.mat4_roty(api.cam_spawn(default_level), yaw)
        cam[12], cam[14] = x, z
        s, z_, so, zo = _compare(oracle, ref, strict, cam, 0.5, 320, 200)
        _, _, ids = oracle.trace_frame(cam, 0.5, 320, 200, want_ids=True)
        saw += int((ids[..., 0]["nportal"] > 0).sum())
    assert saw > 1000          # primary rays did cross M


def test_axis_aligned_rays_vs_reference(oracle, refs, default_level, tmp_path):
    """power-of-two frames from the unrotated spawn camera: the column x = w/2 - 1 has ray.x == 0 and
    the row y = h/2 has ray.y == 0 exactly, so the EPSILON clamp of trace.h:220-222 and the
    infinite/huge reciprocals behind it are on the path (they are almost never taken otherwise)"""
    ref, strict = refs("libref_ids.so"), _strict(refs)
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    for r in (ref, strict):
        r.level_load(p)
    oracle_set_level(oracle, default_level)
    cam = api.cam_spawn(default_level)
    for (w, h) in [(64, 32), (256, 128), (32, 64)]:
        rb, rx, ry = api.ray_setup(cam, w, h)
        assert rb[0] + np.float32(w // 2) * rx[0] == 0 and rb[1] + np.float32(h // 2) * ry[1] == 0
        _compare(oracle, ref, strict, cam, 0.25, w, h)
    # the same looking straight along -x (a quarter turn of the camera rows, exact)
    cam2 = cam.copy()
    cam2[0:3] = (0.0, 0.0, 1.0)
    cam2[8:11] = (-1.0, 0.0, 0.0)
    _compare(oracle, ref, strict, cam2, 0.25, 64, 32)
    # the camera exactly on cell boundaries and on the floor / ceiling planes: zero distances to the first planes
    for pos in ((9.0, 0.5, 4.0), (9.0, 0.0, 5.0), (10.0, 1.0, 4.5), (9.0, 0.5, 5.0)):
        for yaw in (0.0, 0.8):
            c3 = api.mat4_roty(cam, yaw)
            c3[12:15] = pos
            _compare(oracle, ref, strict, c3, 0.25, 96, 64)


def test_flythrough_poses_vs_reference(oracle, refs, default_level, tmp_path):
    ref = refs("libref_ids.so")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    ref.level_load(p)
    oracle_set_level(oracle, default_level)
    poses = api.flythrough(default_level, 600)
    for i in range(0, 600, 75):
        _compare(oracle, ref, None, poses["cam"][i], float(poses["sec_current"][i]), 240, 136)


def test_mirror_hall_vs_reference(oracle, refs, tmp_path):
    lv = levels.mirror_hall()
    ref, strict = refs("libref_ids.so"), _strict(refs)
    p = str(tmp_path / "hall.txt")
    lv.save_txt(p)
    ref.level_load(p)
    strict.level_load(p)
    oracle_set_level(oracle, lv)
    cam0 = api.cam_spawn(lv)
    deep = 0
    for yaw in (0.0, 1.57, 0.4, 3.3, 4.6):
        cam = api.mat4_roty(cam0, yaw)
        _compare(oracle, ref, strict, cam, 0.5, 320, 180)
        # a wide, flat frame: pixel pitch 2/dimx in tan(angle) keeps the centre rows in the chain
        _, _, ids = _compare_ids_only(oracle, ref, cam, 0.5, 1920, 8)
        deep = max(deep, int(ids["nportal"].max()))
    assert deep >= 12   # the chain really is walked deep
    # a pose inside the endless hall: walks that exhaust maxsteps (trace.h:677) must agree too
    last = lv.portals[16]
    cam = np.eye(4, dtype=np.float32)
    cam[3, :3] = (last["x1"] - 2.5, 0.5, last["z1"] + 0.5)
    cam = api.mat4_roty(cam.reshape(16), 1.5707964)
    _, _, so, zo = _compare(oracle, ref, strict, cam, 0.0, 256, 144)


def test_spheres_vs_reference(oracle, refs, default_level, tmp_path):
    ref, strict = refs("libref_ids.so"), _strict(refs)
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    lv = levels.game_spheres(levels.default_level(), sec=0.3)
    for r in (ref, strict):
        r.level_load(p)
        r.objs_clear()
        for sp in lv.spheres:
            r.obj_sphere(sp["r"], sp["refl"], sp["pos"][0], sp["pos"][1], sp["pos"][2], sp["col"][0], sp["col"][1], sp["col"][2])
    oracle_set_level(oracle, lv)
    cam0 = api.cam_spawn(lv)
    hits = 0
    for yaw in (0.0, 0.35, -0.4):
        cam = api.mat4_roty(cam0, yaw)
        _compare(oracle, ref, strict, cam, 0.3, 320, 240)
        _, _, ids = oracle.trace_frame(cam, 0.3, 320, 240, want_ids=True)
        hits += int((ids["kind"] == 1).sum())
    assert hits > 1000
    for r in (ref, strict):
        r.objs_clear()


def test_wide_reference_equals_unmodified_on_default_level(refs, default_level, tmp_path):
    """The widened build (128x128 int cells, 4096 portals) is accepted only if it reproduces
    the unmodified build bit for bit."""
    ref, wide = refs("libref_ids.so"), refs("libref_wide_ids.so")
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    ref.level_load(p)
    cells = np.full((wide.dim, wide.dim), ord("."), np.int32)
    cells[:64, :64] = default_level.cells
    wide.level_set(cells, default_level.portals_i32(), default_level.sx, default_level.sz)
    cam0 = api.cam_spawn(default_level)
    for yaw in (0.0, 2.0, 4.0):
        cam = api.mat4_roty(cam0, yaw)
        a = ref.trace_frame(cam, 0.25, 320, 240, want_ids=True)
        b = wide.trace_frame(cam, 0.25, 320, 240, want_ids=True)
        assert (a[0] == b[0]).all() and (a[1].view(np.uint32) == b[1].view(np.uint32)).all()
        for f in ("ldir", "kind", "nstep", "nportal"):
            assert (a[2][f] == b[2][f]).all()
        # cell ids use each build's own row stride
        ca, cb = a[2]["cell"], b[2]["cell"]
        assert ((ca // 64 == cb // 128) & (ca % 64 == cb % 128)).all()


def test_maze_vs_wide_reference(oracle, refs):
    lv = levels.maze(dim=128, seed=1)
    wide, wide_strict = refs("libref_wide_ids.so"), _wide_strict(refs)
    for r in (wide, wide_strict):
        if r is not None:
            r.level_set(lv.cells, lv.portals_i32(), lv.sx, lv.sz)
    oracle_set_level(oracle, lv)
    poses = levels.poses_random(lv, 16, seed=1)          # the 16 poses of BASELINE configs[3]
    # pose 10 at a finer grid: a quarter-turn portal seen through a thin column of pixels (see tests/test_gpu_parity.py)
    _compare(oracle, wide, wide_strict, poses["cam"][10], float(poses["sec_current"][10]), 960, 540)
    crossed = 0
    for pz in poses:
        s, z, so, zo = _compare(oracle, wide, wide_strict, pz["cam"], float(pz["sec_current"]), 256, 144)
        _, _, ids = oracle.trace_frame(pz["cam"], float(pz["sec_current"]), 64, 36, want_ids=True)
        crossed += int(ids["nportal"].sum())
    assert crossed > 0


def test_open_border_level_vs_reference(oracle, refs, tmp_path):
    """rays that leave the grid: get_cell folds the lookup to index 0 on that axis (util.h:151-158)"""
    lv = levels.open_border()
    ref, strict = refs("libref_ids.so"), _strict(refs)
    p = str(tmp_path / "open.txt")
    lv.save_txt(p)
    ref.level_load(p)
    strict.level_load(p)
    oracle_set_level(oracle, lv)
    outside = 0
    for (x, z, yaw) in [(4.5, 6.5, 0.0), (4.5, 6.5, 1.6), (4.5, 6.5, 3.1), (4.5, 6.5, 4.7), (60.5, 62.5, 0.3),
                        (60.5, 57.5, 2.0), (31.5, 2.5, 3.14), (31.5, 2.5, 0.5)]:
        cam = np.eye(4, dtype=np.float32)
        cam[3, :3] = (x, 0.5, z)
        cam = api.mat4_roty(cam.reshape(16), yaw)
        _compare(oracle, ref, strict, cam, 0.25, 320, 180)
        _, _, ids = oracle.trace_frame(cam, 0.25, 320, 180, want_ids=True)
        c = ids["cell"]
        cx, cz = c % 64, c // 64
        outside += int(((c < 0) | (cx >= 64)).sum())   # cz*64+cx with cx or cz outside the grid
    assert outside > 1000


def test_many_spheres_vs_wide_reference(oracle, refs):
    """per-cell sphere lists at scale (level_prepare_render, level.h:64-81) on the 10k-sector maze"""
    lv = levels.scatter_spheres(levels.maze(dim=128, seed=1), 1500, seed=2)
    wide = refs("libref_wide_ids.so")
    wide.level_set(lv.cells, lv.portals_i32(), lv.sx, lv.sz)
    wide.objs_clear()
    for sp in lv.spheres:
        wide.obj_sphere(sp["r"], sp["refl"], sp["pos"][0], sp["pos"][1], sp["pos"][2], sp["col"][0], sp["col"][1], sp["col"][2])
    oracle_set_level(oracle, lv)
    hits = 0
    try:
        for pz in levels.poses_random(lv, 4, seed=3):
            _compare(oracle, wide, None, pz["cam"], float(pz["sec_current"]), 256, 144)
            _, _, ids = oracle.trace_frame(pz["cam"], float(pz["sec_current"]), 256, 144, want_ids=True)
            hits += int((ids["kind"] == 1).sum())
    finally:
        wide.objs_clear()
    assert hits > 500


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_random_levels_vs_reference(oracle, refs, tmp_path, seed):
    """seeded random level.txt files (tests/fuzz_levels.py): the loader restatement reads what the
    reference's level_load reads, and every frame matches (ids + zbuf bit-exact vs the shipped
    flags, sbuf bit-exact vs the strict build)"""
    import fuzz_levels
    ref, strict = refs("libref_ids.so"), _strict(refs)
    txt = fuzz_levels.random_level_text(seed, open_edges=(seed % 4 == 3))
    p = str(tmp_path / "level.txt")
    open(p, "wb").write(txt)
    for r in (ref, strict):
        r.level_load(p)
    lv = api.level_parse(txt)
    cells, pm = ref.level_get()
    assert (cells == lv.cells).all()
    assert (pm == lv.portals_i32()).all()
    assert ref.spawn() == (lv.sx, lv.sz)
    for r in (ref, strict):
        r.objs_clear()
    if seed % 3 == 1:
        fuzz_levels.random_spheres(lv, seed, 10)
        for r in (ref, strict):
            for sp in lv.spheres:
                r.obj_sphere(sp["r"], sp["refl"], sp["pos"][0], sp["pos"][1], sp["pos"][2], sp["col"][0], sp["col"][1], sp["col"][2])
    oracle_set_level(oracle, lv)
    try:
        for pose in fuzz_levels.random_poses(lv, seed, 3):
            _compare(oracle, ref, strict, pose["cam"], float(pose["sec_current"]), 160, 96)
    finally:
        for r in (ref, strict):
            r.objs_clear()


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_wide_levels_vs_wide_reference(oracle, refs, seed):
    """random 96x96 levels with 120 portal pairs (ids far beyond the 26 letters) against the widened
    build of the reference: ids + zbuf bit-exact, sbuf within 1 LSB"""
    import fuzz_levels
    lv = fuzz_levels.random_wide_level(seed)
    wide, wide_strict = refs("libref_wide_ids.so"), _wide_strict(refs)
    for r in (wide, wide_strict):
        if r is not None:
            r.level_set(lv.cells, lv.portals_i32(), lv.sx, lv.sz)
    oracle_set_level(oracle, lv)
    crossed = 0
    for pose in fuzz_levels.random_poses(lv, seed, 3, dim=96):
        _compare(oracle, wide, wide_strict, pose["cam"], float(pose["sec_current"]), 160, 96)
        _, _, ids = oracle.trace_frame(pose["cam"], float(pose["sec_current"]), 160, 96, want_ids=True)
        crossed += int(ids["nportal"].sum())
    assert crossed > 0


def test_loader_on_random_bytes_vs_reference(refs, tmp_path):
    """the level.txt loader restatement against the reference's level_load on 300 random files: random
    draws from the level alphabet with CR/LF noise, over-long and short rows, more than 64 rows, and
    arbitrary bytes (>= 128 included: `char data[64][64]` is signed, so are the cells here)"""
    ref = refs("libref_noblur.so")
    rng = np.random.default_rng(0)
    alphabet = np.frombuffer(b";;;;;;##$&\"><,^...  \n\n\r\nABCDEFabcdXYZyz*@!~\t0", np.uint8)
    p = str(tmp_path / "level.txt")
    for it in range(300):
        n = int(rng.integers(0, 6000))
        if it % 5 == 0:
            txt = rng.integers(1, 256, n, dtype=np.uint8).tobytes()
        else:
            txt = alphabet[rng.integers(0, len(alphabet), n)].tobytes()
        open(p, "wb").write(txt)
        ref.level_load(p)
        cells, pm = ref.level_get()
        lv = api.level_parse(txt)
        mine = lv.portals_i32()
        assert (cells == lv.cells).all(), it
        assert (pm[:, :4] == mine[:, :4]).all(), it
        # rot12 / c1 / c2 come from the four neighbours of each end (level.h:194-221); for an end on the
        # rim of the grid the reference reads outside `data` (util.h:140-149, unchecked): undefined, skipped
        inner = np.ones(len(pm), bool)
        for k in (0, 1, 2, 3):
            inner &= (pm[:, k] == -1) | ((pm[:, k] >= 1) & (pm[:, k] <= 62))
        assert (pm[inner] == mine[inner]).all(), it
        assert ref.spawn() == (lv.sx, lv.sz), it


def test_ftz_daz_environment_changes_nothing(refs, default_level, tmp_path):
    """Which FP environment is parity pinned to?  The shipped executable runs with MXCSR FTZ|DAZ on (-ffast-math links
    crtfastmath.o into an EXECUTABLE); the shared objects under oracle/_ref, loaded into Python, run with denormals on,
    and the device is built -ftz=false to match them.  The two environments give the same frames: the reference's own
    code (shipped flags, blur on) traced both ways -- default level, the fog-heavy fly-through poses, the mirror hall
    incl. its exhausted walks, twelve fuzz levels with spheres, and a level that is all fog -- differs in 0 pixels of
    sbuf and 0 words of zbuf.  (The smallest magnitudes on the path are the EPSILON clamp, 1e-13, and exp(-0.6*fog):
    the latter is denormal only for an accumulated fog length in (145.6, 172.2) units, i.e. a walk of that length
    through fog cells alone; the all-fog corridor with portals at both ends below is the level built to come close,
    and the pixels that see farther than 100 units in it are counted in the printed line.)"""
    import fuzz_levels
    ref = refs("libref.so")
    cases = []
    p = str(tmp_path / "level.txt")
    default_level.save_txt(p)
    poses = api.flythrough(default_level, 600)
    cases.append((p, [(poses["cam"][i], float(poses["sec_current"][i])) for i in range(0, 600, 60)], None))
    hall = levels.mirror_hall()
    ph = str(tmp_path / "hall.txt")
    hall.save_txt(ph)
    last = hall.portals[16]
    cam = np.eye(4, dtype=np.float32)
    cam[3, :3] = (last["x1"] - 2.5, 0.5, last["z1"] + 0.5)
    cases.append((ph, [(api.mat4_roty(api.cam_spawn(hall), 1.5707964), 0.5), (api.mat4_roty(cam.reshape(16), 1.5707964), 0.0)], None))
    for seed in range(12):
        txt = fuzz_levels.random_level_text(seed, open_edges=(seed % 4 == 3))
        pf = str(tmp_path / ("fuzz%d.txt" % seed))
        open(pf, "wb").write(txt)
        lv = api.level_parse(txt)
        fuzz_levels.random_spheres(lv, seed, 10)
        cases.append((pf, [(q["cam"], float(q["sec_current"])) for q in fuzz_levels.random_poses(lv, seed, 2)], lv.spheres))
    # every open cell foggy ('$'), a long corridor with mirrors of portals at both ends: fog accumulates over hundreds of cells
    rows = [b"." * 40] + [b"." + b"$" * 38 + b"." for _ in range(3)] + [b"." * 40]
    rows[2] = b"A" + b"$" * 38 + b"A"
    pfog = str(tmp_path / "fog.txt")
    open(pfog, "wb").write(b"\n".join(rows) + b"\n")
    camf = np.eye(4, dtype=np.float32)
    camf[3, :3] = (20.5, 0.5, 2.5)
    cases.append((pfog, [(api.mat4_roty(camf.reshape(16), 1.5707964), 0.0), (api.mat4_roty(camf.reshape(16), 1.4), 0.3)], None))
    total = differing = 0
    denormal_fog = 0
    try:
        for path, views, spheres in cases:
            ref.level_load(path)
            ref.objs_clear()
            if spheres is not None:
                for sp in spheres:
                    ref.obj_sphere(sp["r"], sp["refl"], sp["pos"][0], sp["pos"][1], sp["pos"][2], sp["col"][0], sp["col"][1], sp["col"][2])
            for cam, sec in views:
                out = []
                for env in (0, 1):
                    assert bool(ref.set_fpenv(env)) == bool(env)
                    s, z, _, _ = ref.trace_frame(cam, sec, 320, 180)
                    out.append((s, z))
                total += out[0][0].size
                differing += int((out[0][0] != out[1][0]).sum()) + int((out[0][1].view(np.uint32) != out[1][1].view(np.uint32)).sum())
                if path == pfog:
                    denormal_fog += int((out[0][1] > 100.0).sum())
    finally:
        ref.set_fpenv(0)
        ref.objs_clear()
    print("FTZ|DAZ on vs off: %d of %d pixels differ (%d pixels of the all-fog level see farther than 100 units)" % (differing, total, denormal_fog))
    assert differing == 0, (differing, total)
