#!/usr/bin/env python3
"""tools/ncu_lines.py REPORT.ncu-rep [N] -- executed warp-instructions per CUDA source line of the first
kernel in an ncu report captured with --import-source on (SourceCounters section)."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         check=True, capture_output=True, text=True).stdout
    cur = head = None
    per, text, total = collections.Counter(), {}, 0
    for r in csv.reader(io.StringIO(out)):
        if not r:
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
            n = int(r[ie])
        except ValueError:
            continue
        per[(cur, int(r[0]))] += n
        text[(cur, int(r[0]))] = r[1].strip()
        total += n
    print("total warp-instructions", total)
    for (f, l), n in per.most_common(top):
        print("%-18s %4d %10d %5.2f%%  %s" % (f, l, n, 100.0 * n / total, text[(f, l)][:88]))


if __name__ == "__main__":
    main()
