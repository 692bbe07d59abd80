#!/usr/bin/env python3
"""oracle/make_ref.py -- TEST INFRASTRUCTURE.  Build recipe for oracle/_ref/.

Compiles the reference's own hot path (defs.h util.h trace.h screen.h level.h,
#included from /root/reference where they lie) into shared objects under
oracle/_ref/ using the reference Makefile's flags minus SDL/Lua
(Makefile:3-8: -g -O3 -fopenmp -ffast-math -funroll-loops).

Variants (all from ref_harness.c):
  libref.so            as shipped (POSTPROC_BLUR on)
  libref_noblur.so     #undef POSTPROC_BLUR (pre-blur frame)
  libref_ids.so        build-time INSTRUMENTED copy of trace.h: ORC_* hooks at the
                       sites SURVEY 8(c) lists; emits per-pixel cell/ldir/steps/
                       portal-crossings for the three segments; no blur
  libref_wide.so       build-time WIDENED copy (LV_DIM x LV_DIM int cells, PM_MAX
  libref_wide_ids.so   portals) for the levels that exceed 64x64 / 26 letters
  libref_strict.so     same sources, -O2 and no -ffast-math (diagnostic)
  libref_tick.so       main.c:188-379 -- one game tick of player motion (yaw, move, wall push-back,
  libref_tick_strict.so  gravity, raised floors, portal traversal) -- extracted at build time into the
                       body of ref_player_tick(); shipped flags and strict flags
  libref_shim.so       the reference host (defs.h util.h level.h) with the PRODUCT's drop-in
  libref_shim_noblur.so  pwn_screen.h in place of trace.h + screen.h, linked to libpwntrace.so

The patched copies live only in a temporary directory during the build and are
deleted afterwards: no reference source text enters this repository.  Every
patch asserts its anchor text, so a changed reference fails loudly.

Does nothing (exit 0) when /root/reference is absent and the prebuilt
libraries exist -- the GPU box case.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PWN_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
HARNESS = os.path.join(HERE, "ref_harness.c")
HEADERS = ["defs.h", "util.h", "trace.h", "screen.h", "level.h"]

# Makefile:3-5 minus -Wall/-Wextra noise and the sdl/lua pkg-config parts
REF_CFLAGS = ["-g", "-O3", "-fopenmp", "-ffast-math", "-funroll-loops", "-w"]
STRICT_CFLAGS = ["-g", "-O2", "-fopenmp", "-ffp-contract=off", "-w"]


def _read(name):
    with open(os.path.join(REF, name), "r", newline="") as f:
        return f.read().split("\n")


def _write(d, name, lines):
    with open(os.path.join(d, name), "w", newline="") as f:
        f.write("\n".join(lines))


def _expect(lines, lineno, text):
    got = lines[lineno - 1].strip()
    assert got == text, "reference changed at line %d: %r != %r" % (lineno, got, text)


def instrument_trace(lines):
    """ORC_* hooks (defined in ref_harness.c) into a copy of trace.h."""
    lines = list(lines)
    _expect(lines, 204, "pos.m = ifrom->m;")
    _expect(lines, 250, "while(maxsteps-- > 0)")
    _expect(lines, 251, "{")
    _expect(lines, 576, "ldir = (ldir - rot) & 3;")
    _expect(lines, 678, "return ray.m;")
    nwall = nsph = 0
    for i in range(186, 679):  # body of trace_ray only (trace.h:186-679)
        s = lines[i]
        if "return trace_hit_wall(" in s:
            lines[i] = s.replace("return trace_hit_wall(", "ORC_WALL(); return trace_hit_wall(")
            nwall += 1
        elif "return trace_hit_bounce(" in s:
            lines[i] = s.replace("return trace_hit_bounce(", "ORC_SPH(); return trace_hit_bounce(")
            nsph += 1
    assert (nwall, nsph) == (8, 7), (nwall, nsph)
    lines[677] = lines[677].replace("return ray.m;", "ORC_EXHAUST(); return ray.m;")
    lines[575] += " ORC_PORTAL()"
    lines[250] += " ORC_STEP()"
    lines[203] += " ORC_ENTER()"
    return lines


def _sub(lines, lineno, old, new, count=1):
    s = lines[lineno - 1]
    assert s.count(old) == count, "reference changed at line %d: %r lacks %r x%d" % (lineno, s, old, count)
    lines[lineno - 1] = s.replace(old, new)


def widen(files):
    """64x64 char grid / 26 letters -> LV_DIM x LV_DIM int cells / PM_MAX portals.

    Portal p is the cell code 'A'+p for p < 26 (unchanged) and 0x100+p above.
    Sites: SURVEY 0 ("hard limits")."""
    d, u, t, lv = (list(files[k]) for k in ("defs.h", "util.h", "trace.h", "level.h"))
    # defs.h
    _sub(d, 92, "char c1, c2;", "int c1, c2;")
    _sub(d, 103, "pmap[26]", "pmap[PM_MAX]")
    _sub(d, 105, "char data[64][64]", "int data[LV_DIM][LV_DIM]")
    for ln in (107, 108, 109):
        _sub(d, ln, "[64][64]", "[LV_DIM][LV_DIM]")
    d.insert(0, "#define IS_PORTAL(c) (((c) >= 'A' && (c) <= 'Z') || (c) >= 0x100)")
    d.insert(1, "#define PIDX(c) ((c) >= 0x100 ? (c) - 0x100 : (c) - 'A')")
    # util.h (line numbers local to util.h; SURVEY numbers them after defs.h)
    _sub(u, 112, "char c, char oldcell", "int c, int oldcell")
    _sub(u, 122, "if(c >= 'A' && c <= 'Z')", "if(IS_PORTAL(c))")
    _sub(u, 123, "pmap[c - 'A']", "pmap[PIDX(c)]")
    _sub(u, 128, "char c", "int c")
    _sub(u, 151, "static char get_cell", "static int get_cell")
    _sub(u, 154, "cx >= 64", "cx >= LV_DIM")
    _sub(u, 155, "cz >= 64", "cz >= LV_DIM")
    # trace.h
    _sub(t, 244, "char cell", "int cell")
    _sub(t, 252, "< 64", "< LV_DIM", 2)
    _sub(t, 299, "char this_cell", "int this_cell")
    _sub(t, 404, "char xcell", "int xcell")
    _sub(t, 405, "if(xcell >= 'A' && xcell <= 'Z')", "if(IS_PORTAL(xcell))")
    _sub(t, 407, "pmap[xcell - 'A']", "pmap[PIDX(xcell)]")
    _sub(t, 508, "if(cell >= 'A' && cell <= 'Z')", "if(IS_PORTAL(cell))")
    _sub(t, 510, "pmap[cell - 'A']", "pmap[PIDX(cell)]")
    # level.h
    _sub(lv, 69, "z < 64", "z < LV_DIM")
    _sub(lv, 70, "x < 64", "x < LV_DIM")
    _sub(lv, 91, "memset(lv->data, '.', 64*64);",
         "{ int z_, x_; for(z_ = 0; z_ < LV_DIM; z_++) for(x_ = 0; x_ < LV_DIM; x_++) lv->data[z_][x_] = '.'; }")
    _sub(lv, 94, "i < 26", "i < PM_MAX")
    _sub(lv, 118, "z < 64", "z < LV_DIM")
    _sub(lv, 120, "x < 64", "x < LV_DIM")
    _sub(lv, 186, "z < 64", "z < LV_DIM")
    _sub(lv, 187, "x < 64", "x < LV_DIM")
    _sub(lv, 194, "i < 26", "i < PM_MAX")
    out = dict(files)
    out.update({"defs.h": d, "util.h": u, "trace.h": t, "level.h": lv})
    return out


def extract_tick():
    """main.c:188-379: the part of mainloop()'s for(;;) body between the event switch and the end of the loop.
    It reads and writes only `cam`, `gravity`, the eight key flags, `tdiff` and `lv` (get_cell,
    celltype_is_solid, lv->pmap) -- no SDL, no Lua -- so it compiles as the body of a function that declares
    those names (ref_harness.c, -DREF_TICK)."""
    with open(os.path.join(REF, "main.c"), "r", newline="") as f:
        m = f.read().split("\n")
    _expect(m, 186, "}")
    _expect(m, 188, "mat4_roty(&cam, tdiff*3.0f*(k_turnleft - k_turnright));")
    _expect(m, 196, "int cx1 = (int)cam.v.w.v.x;")
    _expect(m, 267, "cam.v.w.m = _mm_add_ps(cam.v.w.m, gravity.m);")
    _expect(m, 296, 'printf("TRAVERSE %c\\n", c2);')
    _expect(m, 377, "cam.v.z.v.z = rvzz;")
    _expect(m, 379, "}")
    _expect(m, 380, "}")
    return m[187:379]


def cc(out_name, inc_dirs, defines, cflags, libs=()):
    cmd = ["gcc"] + cflags + ["-shared", "-fPIC"]
    for i in inc_dirs:
        cmd += ["-I", i]
    cmd += ["-D" + x for x in defines]
    cmd += ["-o", os.path.join(OUT, out_name), HARNESS, "-lm"] + list(libs)
    subprocess.run(cmd, check=True)


VARIANTS = ["libref.so", "libref_noblur.so", "libref_ids.so", "libref_wide.so",
            "libref_wide_ids.so", "libref_strict.so", "libref_tick.so"]


def main():
    os.makedirs(OUT, exist_ok=True)
    if not os.path.isfile(os.path.join(REF, "trace.h")):
        missing = [v for v in VARIANTS if not os.path.isfile(os.path.join(OUT, v))]
        if missing:
            print("make_ref: reference absent and prebuilt libraries missing: %s" % missing)
            return 1
        print("make_ref: reference absent, using prebuilt oracle/_ref")
        return 0

    files = {h: _read(h) for h in HEADERS}
    tmp = tempfile.mkdtemp(prefix="pwnref_")
    try:
        cc("libref.so", [REF], [], REF_CFLAGS)
        cc("libref_noblur.so", [REF], ["REF_NO_BLUR"], REF_CFLAGS)
        cc("libref_strict.so", [REF], ["REF_NO_BLUR"], STRICT_CFLAGS)
        # the compile-time switches of defs.h:6-8 (strict flags, so the restatement must match to the bit)
        cc("libref_reflect0.so", [REF], ["REF_NO_BLUR", "REF_REFLECT=0"], STRICT_CFLAGS)
        cc("libref_reflect1.so", [REF], ["REF_NO_BLUR", "REF_REFLECT=1"], STRICT_CFLAGS)
        cc("libref_norblur.so", [REF], ["REF_NO_BLUR", "REF_NO_REFLECT_BLUR"], STRICT_CFLAGS)

        # the reference HOST (defs.h util.h level.h + this harness as main.c) with the product's
        # drop-in shim in place of trace.h/screen.h, linked against libpwntrace.so
        root = os.path.dirname(HERE)
        pkg = os.path.join(root, "pwnfps_b200")
        if os.path.isfile(os.path.join(pkg, "libpwntrace.so")):
            shim_inc = [os.path.join(pkg, "host"), os.path.join(root, "include"), REF]
            shim_lib = ["-L" + pkg, "-lpwntrace", "-Wl,-rpath,$ORIGIN/../../pwnfps_b200"]
            cc("libref_shim.so", shim_inc, ["REF_SHIM"], REF_CFLAGS, shim_lib)
            cc("libref_shim_noblur.so", shim_inc, ["REF_SHIM", "REF_NO_BLUR"], REF_CFLAGS, shim_lib)

        # one game tick of the player (main.c:188-379) as a function
        d_tick = os.path.join(tmp, "tick")
        os.makedirs(d_tick)
        _write(d_tick, "tick_body.inc", extract_tick())
        cc("libref_tick.so", [d_tick, REF], ["REF_NO_BLUR", "REF_TICK"], REF_CFLAGS)
        cc("libref_tick_strict.so", [d_tick, REF], ["REF_NO_BLUR", "REF_TICK"], STRICT_CFLAGS)

        d_ids = os.path.join(tmp, "ids")
        os.makedirs(d_ids)
        _write(d_ids, "trace.h", instrument_trace(files["trace.h"]))
        cc("libref_ids.so", [d_ids, REF], ["REF_NO_BLUR", "REF_IDS"], REF_CFLAGS)
        cc("libref_ids_strict.so", [d_ids, REF], ["REF_NO_BLUR", "REF_IDS"], STRICT_CFLAGS)

        wide = widen(files)
        d_w = os.path.join(tmp, "wide")
        os.makedirs(d_w)
        for h in HEADERS:
            _write(d_w, h, wide[h])
        cc("libref_wide.so", [d_w], ["REF_NO_BLUR", "REF_WIDE"], REF_CFLAGS)
        cc("libref_wide_blur.so", [d_w], ["REF_WIDE"], REF_CFLAGS)

        d_wi = os.path.join(tmp, "wide_ids")
        os.makedirs(d_wi)
        for h in HEADERS:
            _write(d_wi, h, wide[h])
        _write(d_wi, "trace.h", instrument_trace(wide["trace.h"]))
        cc("libref_wide_ids.so", [d_wi], ["REF_NO_BLUR", "REF_WIDE", "REF_IDS"], REF_CFLAGS)
        # the same instrumented wide sources without -ffast-math: the strict build of the restatement must
        # match it to the bit, ids of all three segments included
        cc("libref_wide_ids_strict.so", [d_wi], ["REF_NO_BLUR", "REF_WIDE", "REF_IDS"], STRICT_CFLAGS)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print("make_ref: built %s" % sorted(os.listdir(OUT)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
