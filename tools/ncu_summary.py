#!/usr/bin/env python3
"""Write the judged summaries from an ncu report: python tools/ncu_summary.py report.ncu-rep out_prefix [--traffic] [--share N]
 -> <out_prefix>_summary.txt (key raw metrics per launch) and, with --traffic, profiles/traffic.json: per kernel (short name,
 last launch wins) dram bytes per launch, duration, executed warp instructions and issue-slot utilisation -- the numbers
 bench.py quotes under "roofline" (it never runs under a profiler itself).  --share N: the capture was made with GPV_DEBUG_OWN=N,0
 (rank 0's share of an N-rank gathering call, on one GPU): stored under "share"/"N".  traffic.json is stamped with a hash of the
 kernel sources (csrc_sha); bench.py prints the ncu numbers only while the sources still hash to it."""
import csv
import io
import json
import os
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "lts__t_bytes.sum",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__thread_inst_executed.sum"]
FP32 = KEYS[-4:-1]


def csrc_sha():
    """sha256 over the device sources: what the executed-instruction counts of a capture belong to"""
    import hashlib
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpview_b200", "csrc")
    h = hashlib.sha256()
    for f in ("gpv_kernels.cuh", "gpv_math.h", "gpv_abi.cu"):
        h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()[:16]


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    f = float(v.replace(",", ""))
    return f * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)


def short(name):
    return re.sub(r"^(void )?(gpv::)?", "", name).split("(")[0]


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = ["source: %s  (ncu --set full --clock-control none --import-source on)" % os.path.basename(rep)]
    per_kernel = {}
    col = {k: hdr.index(k) for k in KEYS if k in hdr}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append("--- %s  (launch id %s)" % (name, r[0]))
        for k in KEYS:
            if k in col:
                lines.append("%-66s %s %s" % (k, r[col[k]], units[col[k]]))
        if "dram__bytes_read.sum" in col:
            i, j, t = col["dram__bytes_read.sum"], col["dram__bytes_write.sum"], col["gpu__time_duration.sum"]
            per_kernel[short(name)] = {
                "dram_bytes": to_bytes(r[i], units[i]) + to_bytes(r[j], units[j]),
                "time_us": to_us(r[t], units[t]),
                "warp_inst_executed": float(r[col["smsp__inst_executed.sum"]].replace(",", "")),
                "issue_active_pct": float(r[col["smsp__issue_active.avg.pct_of_peak_sustained_active"]]),
                "lanes_per_inst": float(r[col["smsp__thread_inst_executed_per_inst_executed.ratio"]]),
                "pipe_fma_pct": float(r[col["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]]),
                "pipe_alu_pct": float(r[col["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]]),
                "source": os.path.basename(prefix) + "_summary.txt"}
            if all(k in col for k in FP32):  # executed FP32 thread-level instructions (FADD, FMUL, FFMA), predicated-on lanes only
                fa, fm, ff = [float(r[col[k]].replace(",", "")) for k in FP32]
                per_kernel[short(name)].update(fp32_fadd=fa, fp32_fmul=fm, fp32_ffma=ff)
            if "smsp__thread_inst_executed.sum" in col:
                per_kernel[short(name)]["thread_inst_executed"] = float(r[col["smsp__thread_inst_executed.sum"]].replace(",", ""))
    open(prefix + "_summary.txt", "w").write("\n".join(lines) + "\n")
    if "--traffic" in sys.argv[3:]:
        p = os.path.join(os.path.dirname(os.path.abspath(prefix)), "traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        sha = csrc_sha()
        if d.get("csrc_sha") != sha:  # captures of other sources do not mix
            d = {"csrc_sha": sha}
        if "--share" in sys.argv[3:]:
            d.setdefault("share", {}).setdefault(sys.argv[sys.argv.index("--share") + 1], {}).update(per_kernel)
        else:
            d.setdefault("kernels", {}).update(per_kernel)
        json.dump(d, open(p, "w"), indent=1)
    print("\n".join(lines[:24]))


if __name__ == "__main__":
    main()
