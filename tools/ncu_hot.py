#!/usr/bin/env python3
"""SASS instructions of one kernel grouped into runs of equal execution count (basic blocks / loop bodies), hottest first:
python tools/ncu_hot.py report.ncu-rep kernel-regex [top]"""
import csv, io, subprocess, sys

def main():
    rep, rx = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "-k", "regex:" + rx], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(r for r in rows if r and r[0] == "Address")
    i_x, i_t = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
    data, seen = [], set()
    for r in rows:
        if len(r) > i_t and r[0].startswith("0x") and r[0] not in seen:
            seen.add(r[0])
            data.append((r[1].strip(), int(r[i_x] or 0), int(r[i_t] or 0)))
    tot = sum(d[1] for d in data) or 1
    blocks = []  # (start, end, count-sum, exec)
    s = 0
    for k in range(1, len(data) + 1):
        if k == len(data) or abs(data[k][1] - data[s][1]) > 0.02 * max(data[s][1], 1):
            blocks.append((s, k, sum(d[1] for d in data[s:k]), data[s][1], sum(d[2] for d in data[s:k])))
            s = k
    print("total warp instructions %d, %d SASS lines, %d blocks" % (tot, len(data), len(blocks)))
    for s, e, c, x, t in sorted(blocks, key=lambda b: -b[2])[:top]:
        print("%5.1f%%  lines %5d-%5d (%4d instr) x %9d execs  lanes %4.1f   first: %s" % (100.0 * c / tot, s, e - 1, e - s, x, t / max(c, 1), data[s][0][:60]))

if __name__ == "__main__":
    main()
