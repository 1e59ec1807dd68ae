#!/usr/bin/env python3
"""Summarise an ncu report per CUDA source line: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top]
(needs the kernels built with -lineinfo and captured with --import-source on).  Prints, per source line, the share of
warp-stall samples and of executed warp instructions -- the view used to pick optimisation targets (profiles/)."""
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    cmd = ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"]
    if len(sys.argv) > 2 and sys.argv[2]:
        cmd += ["-k", "regex:" + sys.argv[2]]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    cur_file = ""
    lines = []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            lines.append((cur_file, r))
    if not hdr:
        print(out[:2000])
        return
    i_s = hdr.index("# Samples") if "# Samples" in hdr else 4
    i_x = hdr.index("Instructions Executed")
    i_t = hdr.index("Thread Instructions Executed")
    tot_s = sum(int(r[i_s] or 0) for _, r in lines) or 1
    tot_x = sum(int(r[i_x] or 0) for _, r in lines) or 1
    print("total samples %d, warp instructions %d" % (tot_s, tot_x))
    for f, r in sorted(lines, key=lambda fr: -int(fr[1][i_x] or 0))[:top]:
        x, t, s = int(r[i_x] or 0), int(r[i_t] or 0), int(r[i_s] or 0)
        print("%5.1f%% inst %5.1f%% stall  lanes %4.1f  %s:%s  %s" % (100.0 * x / tot_x, 100.0 * s / tot_s, t / max(1, x), f, r[0], r[1].strip()[:110]))


if __name__ == "__main__":
    main()
