#!/usr/bin/env python3
"""Dynamic SASS opcode mix of one kernel of an .ncu-rep (executed warp instructions per opcode, pipe guess).
usage: ncu_opmix.py report.ncu-rep kernel-regex [units]   (units: divide counts by this, e.g. windows processed)"""
import csv
import collections
import subprocess
import sys

FMA_PIPE = ("IMAD", "FFMA", "FMUL", "FADD", "IDP", "IMUL")


def main(path, kernel, units=1.0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel, "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    mix = collections.Counter()
    for r in rows:
        if r and r[0] == "Address":
            if hdr is not None:
                break  # first matching launch only
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < 6:
            continue
        src = r[hdr["Source"]].strip()
        toks = src.split()
        if toks and toks[0].startswith("@"):
            toks = toks[1:]
        if not toks:
            continue
        op = toks[0].rstrip(";")
        base = op.split(".")[0]
        key = base if base not in ("IMAD",) else (op if op.startswith(("IMAD.MOV", "IMAD.SHL", "IMAD.IADD")) else "IMAD")
        mix[key] += int(r[hdr["Instructions Executed"]])
    tot = sum(mix.values())
    fma = sum(v for k, v in mix.items() if k.split(".")[0] in FMA_PIPE)
    print("total %.1f per unit; FMA-pipe %.1f (%.0f%%)" % (tot / units, fma / units, 100.0 * fma / max(tot, 1)))
    for k, v in mix.most_common(40):
        print("  %-12s %10.1f" % (k, v / units))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
