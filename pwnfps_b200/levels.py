"""pwnfps_b200.levels -- levels for the headless harness: the default level (as parsed by
the reference's own loader, tests/tools/make_golden.py) and the deterministic synthetic levels of
BASELINE.json configs 3-5.  Generators use their own LCG so that a level
is a pure function of its seed on every platform."""
import os

import numpy as np

from . import api

HERE = os.path.dirname(os.path.abspath(__file__))

_PORTAL_DTYPE = api.PORTAL_DTYPE


def default_level():
    """level.txt of the reference, 30x28 cells in the 64x64 grid, 22 portal pairs."""
    d = np.load(os.path.join(HERE, "data", "level_default.npz"))
    pm = np.zeros(len(d["portals"]), _PORTAL_DTYPE)
    pm.view(np.int32).reshape(-1, 7)[:] = d["portals"]
    return api.Level(d["cells"].astype(np.int32), pm, int(d["spawn"][0]), int(d["spawn"][1]))


class _Lcg:
    """Generator-private LCG (not the trace's): x = x*6364136223846793005 + 1442695040888963407 mod 2^64."""

    def __init__(self, seed):
        self.s = (seed * 2 + 1) & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return self.s >> 33

    def below(self, n):
        return self.next() % n

    def shuffle(self, a):
        for i in range(len(a) - 1, 0, -1):
            j = self.below(i + 1)
            a[i], a[j] = a[j], a[i]


def _grid(w, h):
    return np.full((h, w), ord("."), np.int32)


def mirror_hall(nrooms=17):
    """C3: `nrooms` 5x5 rooms chained by nrooms-1 portal pairs A.. plus one endless hall.

    Room i is entered through the middle of one wall and left through the middle of the
    OPPOSITE wall, so a ray down the axis threads the whole chain (16 crossings deep for 17
    rooms); the entry wall cycles through all four facings so every rot12 value occurs.
    Room types cycle ';' '#' '$' '&'; 2-high rooms carry a raised '"' strip.  The last room
    also holds a pair facing itself across the room: an endless hall that runs the walk into
    maxsteps (trace.h:247,677).  Fits the reference's 64x64 / 26-letter limits."""
    assert 2 <= nrooms <= 24
    g = _grid(64, 64)
    kinds = [";", "#", "$", "&"]
    sides = ["xn", "zn", "xp", "zp"]
    opposite = {"xn": "xp", "xp": "xn", "zn": "zp", "zp": "zn"}
    origins = [(2 + (i % 8) * 7, 2 + (i // 8) * 8) for i in range(nrooms)]   # interior corner
    entry = [sides[(i * (i + 1) // 2) % 4] for i in range(nrooms)]   # steps of 1,2,3,0 quarter turns
    for i, (ox, oz) in enumerate(origins):
        k = kinds[i % 4]
        g[oz:oz + 5, ox:ox + 5] = ord(k)
        if k in "#&":
            # raised strip (a wall from the lower floor) along the edge parallel to the chain axis
            if entry[i] in ("xn", "xp"):
                g[oz, ox:ox + 5] = ord('"')
            else:
                g[oz:oz + 5, ox] = ord('"')

    def slot(room, side, off=2):
        ox, oz = origins[room]
        return {"xn": (ox - 1, oz + off), "xp": (ox + 5, oz + off),
                "zn": (ox + off, oz - 1), "zp": (ox + off, oz + 5)}[side]

    for i in range(nrooms - 1):
        code = api.portal_code(i)
        for (x, z) in (slot(i, opposite[entry[i]]), slot(i + 1, entry[i + 1])):
            assert g[z, x] == ord("."), (i, x, z)
            g[z, x] = code
    # endless hall in the last room, off the chain axis
    code = api.portal_code(nrooms - 1)
    last = nrooms - 1
    a, b = ("zn", "zp") if entry[last] in ("xn", "xp") else ("xn", "xp")
    for (x, z) in (slot(last, a, 3), slot(last, b, 3)):
        assert g[z, x] == ord(".")
        g[z, x] = code
    sx, sz = origins[0][0] + 2, origins[0][1] + 2
    return api.Level.from_grid(g, 26, sx, sz)


def open_border():
    """A 64x64 level whose rooms are open to the edge of the grid: rays leave the map and the
    lookup folds to index 0 on that axis (get_cell, util.h:151-158), for good."""
    g = _grid(64, 64)
    g[0:8, 0:8] = ord(";")
    g[2:5, 3:6] = ord("#")
    g[5, 0:3] = ord("$")
    g[56:64, 56:64] = ord(";")
    g[58:61, 60:64] = ord("&")
    g[0:8, 30:34] = ord(";")          # a corridor open to the top edge only
    g[3, 8] = api.portal_code(0)      # links the two corner rooms
    g[59, 55] = api.portal_code(0)
    return api.Level.from_grid(g, 26, 4, 6)


def maze(dim=128, seed=1, target_cells=10000, portal_pairs=1000):
    """C4/C5: a `dim` x `dim` level carved by a seeded randomised DFS to about `target_cells`
    walkable cells of mixed types, dead ends capped with portal cells paired by a seeded
    shuffle.  Each portal cell has exactly one walkable neighbour and is off the border
    (find_free_dir_2d reads neighbours unchecked, util.h:140-149)."""
    rnd = _Lcg(seed)
    g = _grid(dim, dim)
    n = (dim - 3) // 2          # maze nodes per axis at odd coordinates 1,3,5..
    seen = np.zeros((n, n), bool)
    floor = ord(";")
    stack = [(rnd.below(n), rnd.below(n))]
    seen[stack[0][1], stack[0][0]] = True
    g[1 + 2 * stack[0][1], 1 + 2 * stack[0][0]] = floor
    while stack:
        x, z = stack[-1]
        nb = [(x + dx, z + dz, dx, dz) for dx, dz in ((1, 0), (-1, 0), (0, 1), (0, -1))
              if 0 <= x + dx < n and 0 <= z + dz < n and not seen[z + dz, x + dx]]
        if not nb:
            stack.pop()
            continue
        nx, nz, dx, dz = nb[rnd.below(len(nb))]
        seen[nz, nx] = True
        g[1 + 2 * z + dz, 1 + 2 * x + dx] = floor
        g[1 + 2 * nz, 1 + 2 * nx] = floor
        stack.append((nx, nz))
    # widen: knock out walls between corridors until the walkable count reaches the target
    walk = int((g == floor).sum())
    tries = 0
    while walk < target_cells and tries < 20 * dim * dim:
        tries += 1
        x, z = 2 + rnd.below(dim - 4), 2 + rnd.below(dim - 4)
        if g[z, x] != ord("."):
            continue
        if (g[z, x - 1] == floor and g[z, x + 1] == floor) or (g[z - 1, x] == floor and g[z + 1, x] == floor):
            g[z, x] = floor
            walk += 1
    # mixed cell types in 4x4 blocks: ';' 70%, '#' 15%, '$'/'&' 10%, ramps 5%
    for bz in range(0, dim, 4):
        for bx in range(0, dim, 4):
            r = rnd.below(100)
            if r < 70:
                continue
            k = "#" if r < 85 else ("$" if r < 90 else ("&" if r < 95 else "><,^"[rnd.below(4)]))
            blk = g[bz:bz + 4, bx:bx + 4]
            blk[blk == floor] = ord(k)
    walkable = set(ord(c) for c in ';$"#&><,^')
    # dead ends: solid cells with exactly one walkable 4-neighbour, none of whose neighbours is a portal yet
    cand = []
    for z in range(2, dim - 2):
        for x in range(2, dim - 2):
            if g[z, x] != ord("."):
                continue
            nbw = sum(1 for dx, dz in ((1, 0), (-1, 0), (0, 1), (0, -1)) if int(g[z + dz, x + dx]) in walkable)
            if nbw == 1:
                cand.append((x, z))
    rnd.shuffle(cand)
    chosen = []
    taken = np.zeros((dim, dim), bool)
    for (x, z) in cand:
        if len(chosen) >= 2 * portal_pairs:
            break
        if taken[z - 1:z + 2, x - 1:x + 2].any():
            continue
        # ramps in front of a portal confuse nothing in the trace, but keep the facing cell flat
        front = [int(g[z + dz, x + dx]) for dx, dz in ((1, 0), (-1, 0), (0, 1), (0, -1)) if int(g[z + dz, x + dx]) in walkable][0]
        if chr(front) in "><,^":
            continue
        taken[z, x] = True
        chosen.append((x, z))
    if len(chosen) % 2:
        chosen.pop()
    npairs = len(chosen) // 2
    # row-major registration order decides which end is "1": sort each pair that way
    for p in range(npairs):
        for (x, z) in (chosen[2 * p], chosen[2 * p + 1]):
            g[z, x] = api.portal_code(p)
    # spawn: first walkable 1-high cell in scan order from the centre
    sx = sz = 1
    for z in list(range(dim // 2, dim - 1)) + list(range(1, dim // 2)):
        xs = np.nonzero(g[z] == floor)[0]
        if len(xs):
            sx, sz = int(xs[len(xs) // 2]), z
            break
    return api.Level.from_grid(g, max(npairs, 1), sx, sz)


def poses_random(lv, n, seed=1):
    """n seeded poses at walkable 1-/2-high cells with a seeded yaw: POSE_DTYPE[n]."""
    rnd = _Lcg(seed + 77)
    ok = np.argwhere(np.isin(lv.cells, [ord(c) for c in ";$#&"]))
    poses = np.zeros(n, api.POSE_DTYPE)
    for i in range(n):
        z, x = ok[rnd.below(len(ok))]
        cam = np.eye(4, dtype=np.float32)
        cam[3, :3] = (x + 0.5, 0.5, z + 0.5)
        yaw = (rnd.below(6283) / 1000.0)
        poses["cam"][i] = api.mat4_roty(cam.reshape(16), yaw)
        poses["sec_current"][i] = np.float32(i) / np.float32(16.0)
    return poses


def game_spheres(lv, sec=0.0):
    """The 14 spheres game.lua creates (game.lua:2-30) at their initial offsets, rotated as
    on_tick does (game.lua:48-58) for time `sec`, about (9.5, 0.3, 5.5)."""
    opos = [
        (0.0, 0.0, 0.0, 0.3, 0.8, 0.8, 0.8, 0.6), (0.0, 0.3, 0.0, 0.1, 0.4, 0.4, 0.4, 0.2),
        (0.3, 0.0, 0.0, 0.1, 0.7, 0.7, 1.0, 0.4), (0.0, 0.0, 0.3, 0.1, 0.7, 1.0, 0.7, 0.4),
        (-0.3, 0.0, 0.0, 0.1, 1.0, 0.7, 0.7, 0.4), (0.0, 0.0, -0.3, 0.1, 0.5, 1.0, 1.0, 0.4),
        (0.3, 0.0, 0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (0.1, 0.0, 0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
        (-0.3, 0.0, 0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
        (0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (-0.1, 0.0, 0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
        (-0.3, 0.0, -0.1, 0.03, 0.4, 0.4, 0.4, 0.2), (-0.1, 0.0, -0.3, 0.03, 0.4, 0.4, 0.4, 0.2),
    ]
    obx, oby, obz = 9.5, 0.3, 5.5
    import math
    rs, rc = math.sin(sec * math.pi), math.cos(sec * math.pi)
    lv.spheres = np.zeros(0, api.SPHERE_DTYPE)
    for (x, y, z, r, cb, cg, cr, refl) in opos:
        rx, rz = rc * x + rs * z, rc * z - rs * x
        lv.sphere_add(r, refl, obx + rx, oby + y, obz + rz, cb, cg, cr)
    return lv


def scatter_spheres(lv, n, seed=1):
    """n seeded spheres over the walkable cells of `lv` (radius 0.03-0.3, mixed reflectance, some
    overlapping cell borders): exercises the per-cell lists at scale (OBJ_MAX is 10000, defs.h:4)."""
    rnd = _Lcg(seed + 1234)
    ok = np.argwhere(np.isin(lv.cells, [ord(c) for c in ";$#&"]))
    lv.spheres = np.zeros(0, api.SPHERE_DTYPE)
    sp = np.zeros(n, api.SPHERE_DTYPE)
    for i in range(n):
        z, x = ok[rnd.below(len(ok))]
        r = 0.03 + rnd.below(28) / 100.0
        sp["r"][i] = r
        sp["refl"][i] = (0.0, 0.2, 0.4, 0.6)[rnd.below(4)]
        sp["pos"][i] = (x + rnd.below(1000) / 1000.0, 0.05 + rnd.below(900) / 1000.0, z + rnd.below(1000) / 1000.0, 1.0)
        sp["col"][i] = (rnd.below(100) / 80.0, rnd.below(100) / 80.0, rnd.below(100) / 80.0, 0.0)
    # keep every bounding box inside the grid: the reference indexes its cell lists unchecked (level.h:6-17)
    keep = ((sp["pos"][:, 0] - sp["r"] >= 1) & (sp["pos"][:, 0] + sp["r"] < lv.w - 1)
            & (sp["pos"][:, 2] - sp["r"] >= 1) & (sp["pos"][:, 2] + sp["r"] < lv.h - 1))
    lv.spheres = sp[keep]
    return lv
