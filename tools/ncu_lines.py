#!/usr/bin/env python3
"""Per source line: executed warp instructions and stall samples of one kernel of an .ncu-rep (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep kernel-regex|launch-index [top_n]"""
import csv
import subprocess
import sys


def main(path, kernel, top=40):
    sel = ["--kernel-id", ":::" + kernel] if kernel.isdigit() else ["--kernel-name", "regex:" + kernel]
    if kernel.isdigit():  # the N-th profiled launch of the report (0-based): ncu has no direct flag, so filter by its launch-skip
        sel = ["--launch-skip", kernel, "--launch-count", "1"]
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + sel + ["--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = ""
    lines = []
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < 10 or not r[0]:
            continue
        try:
            inst = int(r[hdr["Instructions Executed"]])
            samples = int(r[hdr["# Samples"]])
        except ValueError:
            continue
        lines.append((inst, samples, fname, r[0], r[1].strip()))
    tot_i = sum(l[0] for l in lines) or 1
    tot_s = sum(l[1] for l in lines) or 1
    print("total warp instructions %d, samples %d" % (tot_i, tot_s))
    for inst, samples, f, ln, src in sorted(lines, key=lambda l: -l[1])[:top]:
        print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100.0 * inst / tot_i, 100.0 * samples / tot_s, f, ln, src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40)
