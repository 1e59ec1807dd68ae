#!/usr/bin/env python3
"""Opcode histogram (weighted by executed warp instructions) and hottest SASS regions of one kernel in an ncu report:
python tools/ncu_sass.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys
from collections import Counter

def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    i_x, i_t = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data = []
    for r in rows:
        if len(r) > i_t and r[0].startswith("0x"):
            data.append((r[1].strip(), int(r[i_x] or 0), int(r[i_t] or 0)))
    tot = sum(d[1] for d in data) or 1
    print("total warp instructions", tot, "SASS lines", len(data))
    c = Counter()
    for sx, x, t in data:
        parts = sx.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        c[op.split(".")[0]] += x
    for op, v in c.most_common(top):
        print("%-10s %5.1f%%" % (op, 100.0 * v / tot))

if __name__ == "__main__":
    main()
