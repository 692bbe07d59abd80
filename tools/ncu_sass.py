#!/usr/bin/env python3
"""tools/ncu_sass.py REPORT.ncu-rep SKIP [--listing FILE] -- the SASS of captured kernel number SKIP with the
executed warp-instruction count of every instruction (ncu source page, needs --import-source on or at least the
SourceCounters section): a per-opcode histogram of EXECUTED instructions with each opcode's pipe, and optionally
the whole annotated listing (address, executed count, share, average active threads, stall samples, instruction)."""
import collections
import csv
import io
import subprocess
import sys

# which issue pipe an opcode occupies (B300_MICROARCH.md "Pipe rates": FFMA/FMUL/FADD/IMAD on the FMA pipe;
# IADD3/LOP3/SHF/PRMT/FMNMX/ISETP/FSETP/SEL/FSEL/MOV on the half-rate ALU pipe)
PIPE = {"FADD": "fma", "FMUL": "fma", "FFMA": "fma", "IMAD": "fma", "IMAD.MOV": "fma", "FADD2": "fma", "FMUL2": "fma", "FFMA2": "fma",
        "HFMA2": "fma", "I2FP": "alu", "F2IP": "alu", "FSEL": "alu", "SEL": "alu", "FSETP": "alu", "ISETP": "alu", "LOP3": "alu", "IADD3": "alu", "VIADD": "alu",
        "SHF": "alu", "LEA": "alu", "PRMT": "alu", "FMNMX": "alu", "IMNMX": "alu", "VIMNMX": "alu", "MOV": "alu", "IABS": "alu", "PLOP3": "alu", "P2R": "alu", "R2P": "alu",
        "FCHK": "alu", "FMNMX3": "alu", "VIMNMX3": "alu", "LOP": "alu", "SGXT": "alu", "BMSK": "alu", "FLO": "alu", "POPC": "alu", "I2F": "xu", "F2I": "xu", "F2F": "xu", "MUFU": "xu",
        "DFMA": "fp64", "DMUL": "fp64", "DADD": "fp64", "DSETP": "fp64", "F2F.F64": "fp64", "LDG": "lsu", "STG": "lsu", "LDS": "lsu", "STS": "lsu", "LDC": "lsu/const",
        "LDCU": "uniform", "ATOMG": "lsu", "RED": "lsu", "BRA": "branch", "BSSY": "branch", "BSYNC": "branch", "BREAK": "branch", "EXIT": "branch", "CALL": "branch", "RET": "branch",
        "WARPSYNC": "branch", "BAR": "branch", "NOP": "-", "S2R": "-", "S2UR": "uniform", "R2UR": "uniform", "UMOV": "uniform", "ULEA": "uniform", "UIADD3": "uniform",
        "ULOP3": "uniform", "USHF": "uniform", "UIMAD": "uniform", "UISETP": "uniform", "USEL": "uniform", "UPRMT": "uniform", "UFLO": "uniform", "VOTE": "alu", "VOTEU": "uniform",
        "SHFL": "lsu", "REDUX": "alu", "CS2R": "-", "ACQBULK": "-", "DEPBAR": "-", "ERRBAR": "-", "MEMBAR": "-", "CCTL": "lsu", "ISCADD": "alu", "LDGDEPBAR": "lsu"}


def opcode(text):
    t = text.strip()
    if t.startswith("@"):
        t = t.split(None, 1)[1]
    op = t.split()[0].rstrip(";")
    if op.startswith("IMAD.MOV"):
        return "IMAD.MOV"
    if op.startswith("F2F.F64") or op.startswith("F2F.F32.F64"):
        return "F2F.F64"
    return op.split(".")[0]


def main():
    rep, skip = sys.argv[1], int(sys.argv[2])
    listing = sys.argv[sys.argv.index("--listing") + 1] if "--listing" in sys.argv else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(skip), "--launch-count", "1"],
                         check=True, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    head = rows[1]
    col = {n: i for i, n in enumerate(head)}
    ie, it, isamp = col["Instructions Executed"], col["Avg. Threads Executed"], col["# Samples"]
    per, static = collections.Counter(), collections.Counter()
    total = 0
    lines = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":      # (the filter can match more than one capture: the first one only)
            break
        if len(r) <= ie or not r[0].startswith("0x"):
            continue
        n = int(r[ie])
        op = opcode(r[1])
        per[op] += n
        static[op] += 1
        total += n
        lines.append((r[0], n, r[it], r[isamp], r[1].strip()))
    print("kernel %s" % name)
    print("static instructions %d, executed warp-instructions %d" % (len(lines), total))
    pipes = collections.Counter()
    for op, n in per.items():
        pipes[PIPE.get(op, "?")] += n
    print("by pipe: " + ", ".join("%s %.1f%%" % (p, 100.0 * n / total) for p, n in pipes.most_common()))
    print("%-10s %-8s %8s %14s %7s" % ("opcode", "pipe", "static", "executed", "share"))
    for op, n in per.most_common():
        if n == 0:
            continue
        print("%-10s %-8s %8d %14d %6.2f%%" % (op, PIPE.get(op, "?"), static[op], n, 100.0 * n / total))
    if listing:
        base = int(lines[0][0], 16)
        with open(listing, "w") as f:
            f.write("# %s\n# offset  executed  share  avg.threads  stall-samples  instruction\n" % name)
            for a, n, t, sm, txt in lines:
                f.write("%05x %12d %6.3f%% %5s %6s  %s\n" % (int(a, 16) - base, n, 100.0 * n / total, t, sm, txt))


if __name__ == "__main__":
    main()
