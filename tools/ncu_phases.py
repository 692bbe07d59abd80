#!/usr/bin/env python3
"""tools/ncu_phases.py REPORT.ncu-rep SKIP [SKIP...] -- executed warp-instructions of captured K1 launches by PHASE of the
kernel (ray generation, segment setup, room step, ..., unwind, pack): the source lines of the ncu source page grouped by
the marker comments of pwnfps_b200/csrc/trace_kernel.cuh as it is NOW (capture and source must be the same build)."""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "pwnfps_b200", "csrc", "trace_kernel.cuh")
MARKS = [("prologue + ray generation", "template <bool SPH, bool IDS, bool LFMA, bool BLURP, bool STD>"),
         ("segment setup", "// ================= trace.h:203-248: segment setup"),
         ("walk: loop head", "// ================= trace.h:250-675: the walk"),
         ("walk: sphere lists", "if(SPH)   // trace.h:252-296"),
         ("walk: room step", "if(cell & 8u)"),
         ("walk: ramp", "else if(cell & 4u)"),
         ("walk: portal", "else if((cell & 15u) == CL_PORTAL)"),
         ("walk: solid + loop tail", "// solid (trace.h:651-664)"),
         ("ids / zbuf store / exhausted", "if(IDS && P.ids != nullptr)"),
         ("shade (trace_hit_wall)", "// ================= shading"),
         ("bounce: mirror, floor normal, reflect", "// ================= trace_hit_bounce"),
         ("bounce: jitter (RNG)", "if(P.reflect_blur)"),
         ("bounce: park blend state", "// (indexed directly:"),
         ("unwind blends (+ fog)", "// the colour the deepest segment returns"),
         ("opcount (ids builds)", "if(IDS && P.opcount != nullptr)"),
         ("pack + frame stores + band counters", "// ---- col_ftoint (util.h:48-59)"),
         ("blur of the previous frame (BLURP)", "// ---- pose batches")]
HELPERS = [("rcpps emulation", "__device__ __forceinline__ float rcp_emul"), ("rsqrtps emulation", "__device__ __forceinline__ float rsqrt_emul"),
           ("randfs", "__device__ __forceinline__ float randfs"), ("col_ftoint", "__device__ __noinline__ uint32_t pack_huge"),
           ("k_pack_eval", "__global__ void k_pack_eval"), ("packed helpers / bit_of / xmov", "__device__ __forceinline__ float xmov"),
           ("sphere_normal / eps_clamp", "struct Aux"), ("band_arrive", "__device__ __forceinline__ void band_arrive")]


def ranges():
    lines = open(SRC).read().split("\n")
    def find(txt):
        for i, l in enumerate(lines):
            if txt in l:
                return i + 1
        raise SystemExit("marker not found: " + txt)
    pts = sorted([(find(t), n) for n, t in MARKS + HELPERS])
    out = []
    for k, (ln, name) in enumerate(pts):
        end = pts[k + 1][0] - 1 if k + 1 < len(pts) else len(lines)
        out.append((ln, end, name))
    return out


def per_line(rep, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(skip), "--launch-count", "1"],
                         check=True, capture_output=True, text=True).stdout
    cur = head = None
    per = collections.Counter()
    kernels = 0
    name = "?"
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "Kernel Name":
            kernels += 1
            if kernels > 1:
                break
            name = r[1]
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            head, ie = r, r.index("Instructions Executed")
            continue
        if head is None or r[0] in ("", "-", "Function Name") or r[2] not in ("-", ""):
            continue
        try:
            per[(cur, int(r[0]))] += int(r[ie])
        except ValueError:
            pass
    return name, per


def main():
    rep = sys.argv[1]
    rg = ranges()
    for skip in sys.argv[2:]:
        name, per = per_line(rep, int(skip))
        tot = sum(per.values())
        agg = collections.Counter()
        for (f, l), n in per.items():
            if f == "trace_kernel.cuh":
                lab = "other"
                for a, b, nm in rg:
                    if a <= l <= b:
                        lab = nm
                        break
                agg[lab] += n
            else:
                agg["(" + str(f) + ")"] += n
        print("capture %s: %s -- %d warp-instructions" % (skip, name, tot))
        for k, v in agg.most_common():
            print("  %-44s %13d  %6.2f%%" % (k, v, 100.0 * v / tot))
        print()


if __name__ == "__main__":
    main()
