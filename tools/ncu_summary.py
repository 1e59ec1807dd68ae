#!/usr/bin/env python3
"""Write the judged summaries from an ncu report: python tools/ncu_summary.py report.ncu-rep out_prefix [kernel-key]
 -> <out_prefix>_summary.txt (key raw metrics per launch) and, with a kernel key, profiles/traffic.json (dram bytes/launch)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "lts__t_bytes.sum"]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["source: %s  (ncu --set full --clock-control none --import-source on)" % os.path.basename(rep)]
    traffic = None
    for r in rows[2:]:
        lines.append("--- %s  (launch id %s)" % (r[hdr.index("Kernel Name")], r[0]))
        for k in KEYS:
            if k in hdr:
                lines.append("%-66s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        if traffic is None and "dram__bytes_read.sum" in hdr:
            i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            traffic = to_bytes(r[i], units[i]) + to_bytes(r[j], units[j])
    open(prefix + "_summary.txt", "w").write("\n".join(lines) + "\n")
    if len(sys.argv) > 3 and traffic is not None:
        p = os.path.join(os.path.dirname(os.path.abspath(prefix)), "traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[sys.argv[3]] = traffic
        d[sys.argv[3] + "_source"] = os.path.basename(prefix) + "_summary.txt"
        json.dump(d, open(p, "w"), indent=1)
    print("\n".join(lines[:24]))


if __name__ == "__main__":
    main()
