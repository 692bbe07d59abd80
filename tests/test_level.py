"""Host logic: the C level loader (pwnfps_b200/host/pwn_level.c) against the reference's own
loader (level.h:107-228 via oracle/_ref) and the synthetic level generators."""
import os

import numpy as np
import pytest

from conftest import REFERENCE
from pwnfps_b200 import api, levels


def test_default_level_kat(default_level):
    """portal table printed by level_load for level.txt, spawn 9,4."""
    lv = default_level
    assert (lv.w, lv.h, lv.sx, lv.sz) == (64, 64, 9, 4)
    got = " ".join("%s%d" % (chr(65 + i), p["rot12"]) for i, p in enumerate(lv.portals) if p["x2"] != -1)
    assert got == "A0 B0 C0 D0 E1 F0 G0 H3 I1 J2 K3 L0 M0 O3 P0 Q0 R0 S0 T2 X0 Y1 Z2"


@pytest.mark.skipif(not os.path.isfile(os.path.join(REFERENCE, "level.txt")), reason="reference absent")
def test_loader_matches_reference_loader(refs, default_level):
    ref = refs("libref_ids.so")
    ref.level_load(os.path.join(REFERENCE, "level.txt"))
    cells, pm = ref.level_get()
    lv = api.level_load(os.path.join(REFERENCE, "level.txt"))
    assert (lv.cells == cells).all()
    assert (lv.portals_i32() == pm).all()
    assert (lv.sx, lv.sz) == ref.spawn()
    # and the committed copy of the parsed level is that same level
    assert (default_level.cells == cells).all() and (default_level.portals_i32() == pm).all()


def test_loader_quirks():
    # lowercase double registration (level.h:144-161): 'm' closes M and opens N
    txt = ".....\r\n.M;m.\r\n.....\r\n..*..\r\n"
    lv = api.level_parse(txt)
    assert chr(lv.cells[1, 1]) == "M" and chr(lv.cells[1, 3]) == "N"
    m, n = lv.portals[12], lv.portals[13]
    assert (m["x1"], m["z1"], m["x2"], m["z2"]) == (1, 1, 3, 1)
    assert (n["x1"], n["z1"], n["x2"]) == (3, 1, -1)
    assert (lv.sx, lv.sz) == (2, 3) and chr(lv.cells[3, 2]) == ";"
    # a third occurrence of a letter is ignored by the table but stays in the grid
    lv = api.level_parse(";A;\n;A;\n;A;\n")
    a = lv.portals[0]
    assert (a["x1"], a["z1"], a["x2"], a["z2"]) == (1, 0, 1, 1) and chr(lv.cells[2, 1]) == "A"


def test_mirror_hall_roundtrips_through_level_txt(tmp_path, refs):
    lv = levels.mirror_hall()
    paired = lv.portals["x2"] != -1
    assert paired.sum() == 17
    assert set(lv.portals["rot12"][paired].tolist()) == {0, 1, 2, 3}
    p = str(tmp_path / "hall.txt")
    lv.save_txt(p)
    back = api.level_load(p)
    assert (back.cells == lv.cells).all() and (back.portals_i32() == lv.portals_i32()).all()
    ref = refs("libref_ids.so")
    ref.level_load(p)
    cells, pm = ref.level_get()
    assert (cells == lv.cells).all() and (pm == lv.portals_i32()).all()


def test_maze_is_deterministic_and_sized():
    a, b = levels.maze(seed=1), levels.maze(seed=1)
    assert (a.cells == b.cells).all() and (a.portals_i32() == b.portals_i32()).all()
    walk = np.isin(a.cells, [ord(c) for c in ';$"#&><,^']).sum()
    assert 9500 <= walk <= 10500
    paired = a.portals["x2"] != -1
    assert paired.all() and len(a.portals) >= 100
    assert (levels.maze(seed=2).cells != a.cells).any()
    # portal cells sit off the border with exactly one walkable neighbour
    for pm in a.portals:
        for (x, z) in ((pm["x1"], pm["z1"]), (pm["x2"], pm["z2"])):
            assert 1 <= x < a.w - 1 and 1 <= z < a.h - 1
            nb = [a.cells[z, x + 1], a.cells[z + 1, x], a.cells[z, x - 1], a.cells[z - 1, x]]
            assert sum(1 for c in nb if chr(c) in ';$"#&><,^') == 1 if all(c < 256 for c in nb) else True


def test_flythrough_is_deterministic(default_level):
    a = api.flythrough(default_level, 300)
    b = api.flythrough(default_level, 300)
    assert a.tobytes() == b.tobytes()
    # it moves, turns and crosses at least one portal (a jump of more than 2 cells in one tick)
    pos = a["cam"][:, 12:15]
    step = np.abs(np.diff(pos[:, [0, 2]], axis=0)).max(axis=1)
    assert (step > 2.0).any() and np.ptp(a["cam"][:, 8]) > 0.5
    assert abs(float(a["sec_current"][60]) - 1.0) < 1e-4


def test_binary_level_roundtrip(tmp_path):
    lv = levels.maze(dim=128, seed=3)
    p = str(tmp_path / "maze.pwnlv")
    lv.save_bin(p)
    back = api.level_load_bin(p)
    assert (back.cells == lv.cells).all() and (back.portals_i32() == lv.portals_i32()).all()
    assert (back.sx, back.sz) == (lv.sx, lv.sz)
    with open(p, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(api.PwnError):
        api.level_load_bin(p)
    with pytest.raises(api.PwnError):
        lv.save_txt(str(tmp_path / "maze.txt"))    # does not fit 64x64 / 26 letters


def test_player_portal_traversal_keeps_the_view(oracle, default_level):
    """The restated player physics (main.c:276-379) teleports and turns the camera through a portal
    with the same quarter-turn maths as the rays: the frame after a crossing must look like the frame
    before it (the portal is a window), unlike a frame from elsewhere on the path."""
    from conftest import oracle_set_level
    oracle_set_level(oracle, default_level)
    poses = api.flythrough(default_level, 600)
    pos = poses["cam"][:, [12, 14]]
    jumps = np.nonzero(np.abs(np.diff(pos, axis=0)).max(axis=1) > 2.0)[0]
    assert len(jumps) >= 2
    w, h = 160, 90

    def frame(i):
        s, _, _ = oracle.trace_frame(poses["cam"][i], 0.0, w, h, nthreads=4)   # same sec: only the pose differs
        return s.view(np.uint8).reshape(h, w, 4)[..., :3].astype(np.int32)

    for j in jumps[:3]:
        before, after = frame(j), frame(j + 1)
        far = frame((j + 150) % 600)
        d_cross = np.abs(before - after).mean()
        d_far = np.abs(before - far).mean()
        assert d_cross < 12.0 and d_cross < 0.5 * d_far, (j, d_cross, d_far)


def test_game_script_spheres(default_level):
    """host/pwn_game.c: the 14 spheres of game.lua start where the script creates them, keep their
    radii, never enter a '.' cell (the script's only wall) and turn there."""
    g = api.Game(default_level)
    s0 = g.spheres()
    assert len(s0) == 14
    assert np.allclose(s0["pos"][0], (9.5, 0.3, 5.5, 1.0)) and np.isclose(s0["r"][0], 0.3) and np.isclose(s0["refl"][0], 0.6)
    assert np.allclose(s0["pos"][2][:3], (9.8, 0.3, 5.5))
    seen_cells = set()
    sec = 0.0
    for i in range(1200):
        sp = g.tick(sec, 1.0 / 60.0)
        sec += 1.0 / 60.0
        cx, cz = int(np.floor(sp["pos"][0][0])), int(np.floor(sp["pos"][0][2]))
        assert chr(default_level.cells[cz, cx]) != ".", (i, cx, cz)      # the script only turns at '.' (game.lua:70)
        seen_cells.add((cx, cz))
        assert np.isclose(np.hypot(sp["pos"][2][0] - sp["pos"][0][0], sp["pos"][2][2] - sp["pos"][0][2]), 0.3, atol=1e-5)
    assert len(seen_cells) >= 5       # the cluster walks
    assert set(np.round(g.spheres()["col"][1][:3].astype(np.float64), 3).tolist()) <= {0.3, 1.3}   # sphere 2 blinks (game.lua:36-40)


def test_ppm_dump_and_diff(tmp_path):
    """host/pwn_image.c: the frame dump used for visual diffing"""
    rng = np.random.default_rng(5)
    a = rng.integers(0, 2**32, (7, 12), dtype=np.uint32)
    p = str(tmp_path / "f.ppm")
    api.save_ppm(p, a, w=10, h=6)
    raw = open(p, "rb").read()
    head = b"P6\n10 6\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + 10 * 6 * 3
    px = np.frombuffer(raw[len(head):], np.uint8).reshape(6, 10, 3)
    assert (px[..., 0] == ((a[:6, :10] >> 16) & 255)).all()
    assert (px[..., 1] == ((a[:6, :10] >> 8) & 255)).all()
    assert (px[..., 2] == (a[:6, :10] & 255)).all()
    b = a.copy()
    b[2, 3] ^= 0x00010000          # one LSB of red
    b[4, 5] ^= 0xFF000000          # alpha only: ignored
    d = api.frame_diff(a, b, gain=64)
    assert d[2, 3] == 64 << 16 and np.count_nonzero(d) == 1


def test_extended_text_format(tmp_path):
    """the "!pwnlv" text format: level.txt rules on any grid, '@' cells for pairs >= 26;
    a level round-trips through it, and the default level written in it loads as level.txt does"""
    lv = levels.maze(dim=128, seed=1)
    assert len(lv.portals) > 26
    p = str(tmp_path / "maze.pwnlv.txt")
    lv.save_txtx(p)
    assert open(p, "rb").read(6) == b"!pwnlv"
    back = api.level_load(p)
    assert back.cells.shape == lv.cells.shape and (back.cells == lv.cells).all()
    assert (back.sx, back.sz) == (lv.sx, lv.sz)
    assert len(back.portals) == len(lv.portals)
    assert (back.portals_i32() == lv.portals_i32()).all()
    d = levels.default_level()
    p2 = str(tmp_path / "default.pwnlv.txt")
    d.save_txtx(p2)
    p3 = str(tmp_path / "default.txt")
    d.save_txt(p3)
    assert b"m" in open(p3, "rb").read()        # the lower-case mark of level.txt:5 is restored from the table
    for back in (api.level_load(p2), api.level_load(p3)):
        assert (back.cells[:64, :64] == d.cells).all() and (back.portals_i32()[:, :5] == d.portals_i32()[:, :5]).all()
        assert (back.sx, back.sz) == (d.sx, d.sz)
    # hand-written: short rows, blank line, CRLF, lower-case double registration, one '@' pair + a stray third
    txt = b"!pwnlv 70 5 30\r\n" + b"." * 70 + b"\r\n\r\n" + b".*;;a;;@;;@;;@.\n" + b".;;;A;;;;;;;;;.\n" + b"...\n" + b"@ 27 7 1\n@ 27 10 1\n@ 27 13 1\n"
    lv2 = api.level_parse(txt)
    assert lv2.cells.shape == (5, 70) and (lv2.sx, lv2.sz) == (1, 1)
    assert lv2.cells[1, 4] == ord("B") and lv2.cells[2, 4] == ord("A")      # 'a' registers in A, becomes B
    pm = lv2.portals_i32()
    assert tuple(pm[0][:4]) == (4, 1, 4, 2) and tuple(pm[1][:3]) == (4, 1, -1)
    assert lv2.cells[1, 7] == 0x100 + 27 and tuple(pm[27][:4]) == (7, 1, 10, 1)   # third '@' cell ignored by the table
    assert lv2.cells[1, 13] == 0x100 + 27
    with pytest.raises(api.PwnError):
        api.level_parse(b"!pwnlv 10 10 30\n..@..\n@ 5 2 0\n")      # id < 26 must be written as a letter


def _tick_pin(lv_text_or_level, refs, variant, nticks, seed, tdiffs, key_every=7):
    """host/pwn_camera.c's tick against the reference's own block (main.c:188-379, extracted at build time into
    oracle/_ref/libref_tick*.so): same keys, same tdiff, cam and gravity compared bit for bit after every tick.
    Returns the number of ticks that differed (after a difference both sides continue from the reference's state)."""
    import tempfile
    ref = refs(variant)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "level.txt")
        if isinstance(lv_text_or_level, (bytes, str)):
            txt = lv_text_or_level if isinstance(lv_text_or_level, bytes) else lv_text_or_level.encode("latin-1")
            open(p, "wb").write(txt)
            lv = api.level_parse(txt)
        else:
            lv = lv_text_or_level
            lv.save_txt(p)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(1)
        os.dup2(devnull, 1)          # level_load prints its portal table
        try:
            ref.level_load(p)
        finally:
            os.dup2(saved, 1)
            os.close(devnull)
            os.close(saved)
    rng = np.random.default_rng(seed)
    pl = api.Player(lv)
    cam, grav = pl.cam.copy(), pl.gravity.copy()
    keys = np.zeros(6, np.int32)
    bad = crossed = 0
    for t in range(nticks):
        if t % key_every == 0:
            keys = (rng.random(6) < 0.45).astype(np.int32)
        td = tdiffs[t % len(tdiffs)]
        before = cam[[12, 14]].copy()
        pl.tick(keys, td)
        ref.player_tick(cam, grav, np.array([keys[0], keys[1], 0, 0, keys[2], keys[3], keys[4], keys[5]], np.int32), td)
        crossed += int(np.abs(cam[[12, 14]] - before).max() > 1.5)
        if not (np.array_equal(cam.view(np.uint32), pl.cam.view(np.uint32)) and np.array_equal(grav.view(np.uint32), pl.gravity.view(np.uint32))):
            bad += 1
            pl.state[:16] = cam
            pl.state[16:] = grav
    return bad, crossed


def test_player_tick_pinned_to_the_reference_block(refs, default_level):
    """the product's player tick == main.c:188-379 compiled from the reference (strict flags) bit for
    bit -- the 600-tick bench script, 20 000 random-key ticks on the default level (wall slides, raised floors,
    all its portals), and 1000 random key sequences on the fuzz levels (every cell class next to every other,
    third-occurrence and unpaired portal letters, open grid edges).  Against the shipped -ffast-math build of the
    same block the tick may differ where gcc's fast-math code generation does (counted, as for the trace)."""
    from oracle import refbind
    import fuzz_levels
    if not refbind.available("libref_tick_strict.so"):
        pytest.skip("libref_tick_strict.so not built")
    # the bench script itself: pwn_flythrough replayed through the reference's block
    ref = refs("libref_tick_strict.so")
    poses = api.flythrough(default_level, 600)
    bad, crossed = _tick_pin(default_level, refs, "libref_tick_strict.so", 0, 0, [1.0 / 60.0])
    cam = api.cam_spawn(default_level)
    grav = np.zeros(4, np.float32)
    for f in range(600):
        assert np.array_equal(cam.view(np.uint32), poses["cam"][f].view(np.uint32)), f
        t = f % 240
        fwd, tl, tr, ml = (1, 0, 0, 0) if t < 90 else (1, 1, 0, 0) if t < 150 else (0, 0, 1, 0) if t < 200 else (1, 0, 0, 1)
        ref.player_tick(cam, grav, np.array([tl, tr, 0, 0, fwd, 0, ml, 0], np.int32), np.float32(1.0) / np.float32(60.0))
    bad, crossed = _tick_pin(default_level, refs, "libref_tick_strict.so", 20000, 1, [1.0 / 60.0, 0.013, 0.031, 0.1])
    assert bad == 0 and crossed > 20, (bad, crossed)
    total_cross = 0
    for seed in range(100):
        txt = fuzz_levels.random_level_text(seed, open_edges=(seed % 4 == 3))
        for k in range(10):
            bad, crossed = _tick_pin(txt, refs, "libref_tick_strict.so", 120, 1000 * seed + k, [1.0 / 60.0, 0.05, 0.2], key_every=5)
            assert bad == 0, (seed, k, bad)
            total_cross += crossed
    assert total_cross > 50
    if refbind.available("libref_tick.so"):
        bad, _ = _tick_pin(default_level, refs, "libref_tick.so", 20000, 1, [1.0 / 60.0, 0.013, 0.031, 0.1])
        assert bad <= 40, bad          # shipped flags: fast-math code generation, a few ticks in 20 000 (6 here)


# game.lua:1-31 as data: the table the script builds.  When /root/reference is there the test parses it from the
# script itself and checks this copy against it.
_GAME_LUA_OPOS = [
    (0.0, 0.0, 0.0, 0.3, 0.8, 0.8, 0.8, 0.6), (0.0, 0.3, 0.0, 0.1, 0.4, 0.4, 0.4, 0.2),
    (0.3, 0.0, 0.0, 0.1, 0.7, 0.7, 1.0, 0.4), (0.0, 0.0, 0.3, 0.1, 0.7, 1.0, 0.7, 0.4),
    (-0.3, 0.0, 0.0, 0.1, 1.0, 0.7, 0.7, 0.4), (0.0, 0.0, -0.3, 0.1, 0.5, 1.0, 1.0, 0.4),
    (0.3, 0.0, 0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (0.1, 0.0, 0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
    (-0.3, 0.0, 0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
    (0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (-0.1, 0.0, 0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
    (-0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (-0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2)]


class _LuaGame:
    """A second, independent evaluation of game.lua in Python floats (= Lua 5.1 numbers = C doubles; math.fmod,
    math.sin, math.cos, math.floor are the same libm calls Lua's math library makes): the globals and on_tick of
    game.lua:22-87 statement by statement.  obj_set narrows to float when it stores into the part (script.h:22-32)."""

    def __init__(self, level_get, opos):
        import math
        self.m = math
        self.level_get = level_get
        self.opos = [list(r) for r in opos]
        self.obx, self.oby, self.obz = 9.5, 0.3, 5.5
        self.obvx, self.obvz = 1, 0
        self.obj = [None] * len(self.opos)
        for i, o in enumerate(self.opos):
            self.obj_set(i, o[3], o[7], self.obx + o[0], self.oby + o[1], self.obz + o[2], o[4], o[5], o[6])

    def obj_set(self, i, r, refl, x, y, z, cb, cg, cr):
        self.obj[i] = np.array([r, refl, x, y, z, 1.0, cb, cg, cr], np.float32)

    def on_tick(self, sec_current, sec_delta):
        m, opos = self.m, self.opos
        if m.fmod(sec_current, 0.5) < 0.15:
            opos[1][4], opos[1][5], opos[1][6] = 0.3, 0.3, 1.3
        else:
            opos[1][4], opos[1][5], opos[1][6] = 0.3, 0.3, 0.3
        for i, o in enumerate(opos):
            rx, ry, rz = o[0], o[1], o[2]
            rs = m.sin(sec_current * m.pi * 2 / 2)
            rc = m.cos(sec_current * m.pi * 2 / 2)
            rx, rz = rc * rx + rs * rz, rc * rz - rs * rx
            self.obj_set(i, o[3], o[7], self.obx + rx, self.oby + ry, self.obz + rz, o[4], o[5], o[6])
        c1 = self.level_get(m.floor(self.obx), m.floor(self.obz))
        spd = 2.0
        nobx = self.obx + self.obvx * sec_delta * spd
        nobz = self.obz + self.obvz * sec_delta * spd
        c2 = self.level_get(m.floor(nobx + self.obvx * 0.5), m.floor(nobz + self.obvz * 0.5))
        turns = 0
        if c1 != c2:
            if c2 == "." or ((c1 == "#" or c1 == "&") and c2 == '"'):
                self.obvx, self.obvz = self.obvz, -self.obvx
                nobx = self.obx + self.obvx * sec_delta * spd
                nobz = self.obz + self.obvz * sec_delta * spd
                c2 = self.level_get(m.floor(nobx + self.obvx * 0.5), m.floor(nobz + self.obvz * 0.5))
                turns = 1
                if c2 == "." or ((c1 == "#" or c1 == "&") and c2 == '"'):
                    self.obvx, self.obvz = -self.obvx, -self.obvz
                    nobx = self.obx + self.obvx * sec_delta * spd
                    nobz = self.obz + self.obvz * sec_delta * spd
                    c2 = self.level_get(m.floor(nobx + self.obvx * 0.5), m.floor(nobz + self.obvz * 0.5))
                    turns = 2
        self.obx, self.obz = nobx, nobz
        return turns


def _parse_game_lua(path):
    """the opos table and the start values straight from the reference's script"""
    import re
    src = open(path).read()
    body = src[src.index("opos = {"):src.index("obx, oby, obz")]
    rows = re.findall(r"\{([^{}]+)\}", body)
    return [tuple(float(v) for v in r.split(",") if v.strip()) for r in rows]


def test_game_script_pinned_to_a_second_evaluation(default_level):
    """host/pwn_game.c against an independent evaluation of game.lua:22-87 (there is no Lua interpreter
    in this image): 10 000 ticks at 60 Hz and 3 000 at irregular steps, every sphere (r, refl, pos, col as floats)
    bit for bit after every tick, the turn rule of game.lua:69-83 taken both once and twice; then the same on the
    fuzz levels, where raised floors beside 2-high rooms make the second clause of the rule fire."""
    import fuzz_levels
    opos = _GAME_LUA_OPOS
    lua = os.path.join(REFERENCE, "game.lua")
    if os.path.isfile(lua):
        assert _parse_game_lua(lua) == [tuple(r) for r in opos]

    def run(lv, nticks, deltas):
        def level_get(x, z):          # lf_level_get -> get_cell (script.h:53-63, util.h:151-158)
            x, z = int(x), int(z)
            if x < 0 or x >= 64:
                x = 0
            if z < 0 or z >= 64:
                z = 0
            c = int(lv.cells[z, x]) if (z < lv.h and x < lv.w) else ord(".")
            return chr(c) if c < 256 else "?"
        py = _LuaGame(level_get, opos)
        g = api.Game(lv)
        s0 = g.spheres()
        for i in range(14):
            assert np.array_equal(np.concatenate([[s0["r"][i], s0["refl"][i]], s0["pos"][i], s0["col"][i][:3]]).astype(np.float32).view(np.uint32),
                                  py.obj[i].view(np.uint32)), i
        sec = 0.0
        turns = [0, 0, 0]
        for t in range(nticks):
            dt = deltas[t % len(deltas)]
            sec += dt
            turns[py.on_tick(sec, dt)] += 1
            sp = g.tick(sec, dt)
            got = np.concatenate([sp["r"][:, None], sp["refl"][:, None], sp["pos"], sp["col"][:, :3]], axis=1).astype(np.float32)
            want = np.stack(py.obj)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (t, np.argwhere(got != want)[:3])
        return turns

    turns = run(default_level, 10000, [1.0 / 60.0])
    assert turns[1] > 10, turns
    t2 = run(default_level, 3000, [0.013, 0.05, 0.2, 1.0 / 30.0, 0.4])
    both = [turns[k] + t2[k] for k in range(3)]
    for seed in range(40):
        lv = api.level_parse(fuzz_levels.random_level_text(seed + 200, dim=22))
        # (the script starts at (9.5, 5.5): whatever cell that is in the fuzz level)
        t3 = run(lv, 300, [1.0 / 60.0, 0.1, 0.3])
        both = [both[k] + t3[k] for k in range(3)]
    assert both[1] > 20 and both[2] > 0, both
