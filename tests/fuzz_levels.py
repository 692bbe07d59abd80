"""tests/fuzz_levels.py -- seeded random level.txt files and poses for the parity fuzz tests.

Everything the loader and the trace define is fair game: every cell class next to every other,
portal letters anywhere in the interior (facing nothing, facing each other, third occurrences,
unpaired, lower-case double registration), rooms open to the edge of the grid, cameras inside
solid cells.  The generator is a pure function of its seed (own LCG, no numpy RNG)."""
import numpy as np

from pwnfps_b200 import api


class Lcg:
    def __init__(self, seed):
        self.s = (seed * 2654435761 + 12345) & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return self.s >> 33

    def below(self, n):
        return self.next() % n

    def unit(self):
        return self.below(1 << 24) / float(1 << 24)


# (char, weight): rooms dominate so that rays travel; every class is present
CLASSES = [(";", 40), ("#", 16), ("$", 6), ("&", 6), ('"', 6), (">", 2), ("<", 2), (",", 2), ("^", 2), (".", 18)]


def random_level_text(seed, dim=22, open_edges=False):
    """a dim x dim block in the top-left corner of the 64x64 grid, as level.txt bytes"""
    r = Lcg(seed)
    tot = sum(w for _, w in CLASSES)
    g = [["." for _ in range(dim)] for _ in range(dim)]
    # blobs of one class (rooms) plus salt
    for _ in range(dim):
        x0, z0 = r.below(dim), r.below(dim)
        w, h = 1 + r.below(6), 1 + r.below(6)
        k = r.below(tot)
        for c, wgt in CLASSES:
            if k < wgt:
                break
            k -= wgt
        for z in range(z0, min(dim, z0 + h)):
            for x in range(x0, min(dim, x0 + w)):
                g[z][x] = c
    for _ in range(dim * 2):
        k = r.below(tot)
        for c, wgt in CLASSES:
            if k < wgt:
                break
            k -= wgt
        g[r.below(dim)][r.below(dim)] = c
    if not open_edges:
        for i in range(dim):
            g[0][i] = g[dim - 1][i] = g[i][0] = g[i][dim - 1] = "."
    # portal letters in the interior: pairs, some third occurrences, some unpaired, some lower case
    nlet = 3 + r.below(6)
    for k in range(nlet):
        letter = chr(ord("A") + r.below(12))
        copies = (2, 2, 2, 2, 1, 3)[r.below(6)]
        for _ in range(copies):
            x, z = 1 + r.below(dim - 2), 1 + r.below(dim - 2)
            ch = letter
            if r.below(8) == 0 and letter != "Z":
                ch = letter.lower()
            g[z][x] = ch
    # spawn on a walkable cell if there is one
    walk = [(x, z) for z in range(dim) for x in range(dim) if g[z][x] == ";"]
    if walk:
        x, z = walk[r.below(len(walk))]
        g[z][x] = "*"
    return ("\n".join("".join(row) for row in g) + "\n").encode("latin-1")


def random_poses(lv, seed, n, dim=22):
    r = Lcg(seed + 999)
    walk = np.argwhere(np.isin(lv.cells[:dim, :dim], [ord(c) for c in ';$#&"><,^']))
    poses = np.zeros(n, api.POSE_DTYPE)
    for i in range(n):
        if len(walk) and r.below(8) != 0:
            z, x = walk[r.below(len(walk))]
        else:
            z, x = r.below(dim), r.below(dim)          # anywhere, solid cells included
        cam = np.eye(4, dtype=np.float32)
        cam[3, :3] = (x + 0.05 + 0.9 * r.unit(), 0.1 + 0.8 * r.unit(), z + 0.05 + 0.9 * r.unit())
        poses["cam"][i] = api.mat4_roty(cam.reshape(16), 6.2831853 * r.unit())
        poses["sec_current"][i] = np.float32(10.0 * r.unit())
    return poses


def random_spheres(lv, seed, n, dim=22):
    """n spheres inside the block (bounding boxes stay inside the grid: level_part_add_bbox, level.h:1-39,
    writes unchecked), around walkable cells, some overlapping cell borders and each other"""
    r = Lcg(seed + 4242)
    walk = np.argwhere(np.isin(lv.cells[:dim, :dim], [ord(c) for c in ';$#&"']))
    lv.spheres = lv.spheres[:0]
    if len(walk) == 0:
        return lv
    for _ in range(n):
        z, x = walk[r.below(len(walk))]
        rad = 0.05 + 0.4 * r.unit()
        px = min(max(x + r.unit(), 1.0), dim - 1.0)
        pz = min(max(z + r.unit(), 1.0), dim - 1.0)
        lv.sphere_add(np.float32(rad), np.float32((0.0, 0.3, 0.6, 0.9)[r.below(4)]), np.float32(px),
                      np.float32(0.1 + 1.3 * r.unit()), np.float32(pz),
                      np.float32(r.unit()), np.float32(r.unit()), np.float32(r.unit()))
    return lv


def random_wide_level(seed, dim=128, block=96, pairs=120):
    """a block x block random level in a dim x dim grid with `pairs` portal pairs, ids far beyond the 26
    letters (api.portal_code), paired and derived by the host library the way the loader does"""
    r = Lcg(seed + 31337)
    tot = sum(w for _, w in CLASSES)
    g = np.full((dim, dim), ord("."), np.int32)

    def pick():
        k = r.below(tot)
        for c, wgt in CLASSES:
            if k < wgt:
                return ord(c)
            k -= wgt
    for _ in range(block * 3):
        x0, z0 = 1 + r.below(block - 2), 1 + r.below(block - 2)
        w, h = 1 + r.below(9), 1 + r.below(9)
        g[z0:min(block - 1, z0 + h), x0:min(block - 1, x0 + w)] = pick()
    for _ in range(block * 4):
        g[1 + r.below(block - 2), 1 + r.below(block - 2)] = pick()
    for p in range(pairs):
        copies = (2, 2, 2, 2, 1, 3)[r.below(6)]
        for _ in range(copies):
            g[1 + r.below(block - 2), 1 + r.below(block - 2)] = api.portal_code(p)
    walk = np.argwhere(g == ord(";"))
    z, x = walk[r.below(len(walk))] if len(walk) else (1, 1)
    return api.Level.from_grid(g, pairs, int(x), int(z))
