#!/usr/bin/env python3
"""tools/sass_loop.py LIB [mangled-substring] -- static look at K1's walk loop in a build: the loop is
the backward branch with the longest span; prints its length, the fast room path length (up to the first
backward-target exit) and the opcode histogram of the whole loop."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "ILb0ELb0ELb1ELb0E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
for f in funcs:
    if "k_trace" in f.split("\n")[0] and want in f.split("\n")[0]:
        ins = re.findall(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)
        ins = [(int(a, 16), t.strip()) for a, t in ins]
        best = None
        for a, t in ins:
            m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                span = a - int(m.group(1), 16)
                # the walk loop: the SHORTEST backward branch longer than 150 instructions (the segment loop
                # around it is longer, the sphere-list and libm loops inside it are shorter)
                if span > 150*16 and (best is None or span < best[0]):
                    best = (span, int(m.group(1), 16), a)
        span, head, tail = best
        body = [(a, t) for a, t in ins if head <= a <= tail]
        ops = collections.Counter()
        for a, t in body:
            o = t.split()[1] if t.startswith("@") else t.split()[0]
            o = "IMAD.MOV" if o.startswith("IMAD.MOV") else o.split(".")[0]
            ops[o] += 1
        # fast room path: from the head to the first forward branch that lands on the loop latch region
        print("%s: function %d instr, loop 0x%x..0x%x = %d instr" % (want, len(ins), head, tail, len(body)))
        first_fwd = None
        for a, t in body:
            m = re.search(r"BRA\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) > a and int(m.group(1), 16) >= tail - 0x60 and "@" in t:
                first_fwd = (a, int(m.group(1), 16))
                break
        if first_fwd:
            n = (first_fwd[0] - head) // 16 + 1 + (tail - first_fwd[1]) // 16 + 1
            print("  fast room step: head..0x%x then 0x%x..tail = %d instr" % (first_fwd[0], first_fwd[1], n))
            seg = [t for a, t in body if a <= first_fwd[0] or a >= first_fwd[1]]
            print("  MOVs on it: %d, BSSY/BSYNC/BREAK: %d" % (sum("IMAD.MOV" in t or t.startswith("MOV") for t in seg),
                                                           sum(any(k in t for k in ("BSSY", "BSYNC", "BREAK")) for t in seg)))
        print("  loop opcode histogram:", dict(ops.most_common(12)))
