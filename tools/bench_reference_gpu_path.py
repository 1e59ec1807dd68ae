#!/usr/bin/env python3
"""The reference's own GPU path next to the product, like for like (SURVEY.md 8c "Tier C": the kernels to beat).

  python tools/bench_reference_gpu_path.py [--mesh cessna|torus|block|sphere|cad] [--l1 64] [--l2 4] [--reps 3]

Three ways through the same model on cuda:0, each in its own process:
  reference   the reference's UNMODIFIED host code (Object::ClassifyTessellationCUDA incl. its two-pass buffer re-run,
              Object::ClassifyInOutTessellationLevel2CUDA) on the reference's OWN kernels recompiled for sm_100a, strict IEEE
              (oracle/_ref/libgpvref_refcuda.so) -- what a GPView user runs today, minus the GL solid fill, which is seeded
              from the oracle because it cannot run headless
  compat      the same host code on the product's three operator symbols (oracle/_ref/libgpvref_b200.so): the drop-in at the
              reference's own boundary
  native      gpv_voxelize_host: host triangles in, host streams out (what bench.py reports as e2e)
Times are host wall clock around the whole call (the reference's path is dominated by its host loops and D2H copies, which
is the point).  Level-1 / Level-2 states of the three must agree; the reference's kernels rely on zero-filled cudaMalloc
memory, so a mismatch of `reference` alone is reported, not fatal.  TEST / MEASUREMENT TOOL: it loads
oracle/ and is not part of the product.  Needs a GPU and the prebuilt oracle/_ref."""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def mesh_file(name, tmp):
    from util import mesh_path
    return mesh_path(name, tmp)


def worker(kind, path, l1, l2, reps):
    from oracle import oraclebind as O
    out = {"kind": kind}
    if kind == "native":
        import gpview_b200 as gpv
        from gpview_b200 import binding as B
        L = B.lib()
        mesh = gpv.load_mesh(path)
        ctx = gpv.Context(0)
        res = ctx.voxelize(mesh, gpv.Params(l1, l2, gpv.GPV_NORMALS))
        cells, nb, n23 = res.cells, res.nb, res.n23
        hb = {k: L.gpv_alloc_host(n + 64) for k, n in (("l1", cells), ("pre", cells * 4), ("bi", nb * 4), ("l2", nb * n23), ("n1", cells * 3), ("n2", nb * n23 * 3))}
        hs = B.CHostStreams(hb["l1"], hb["pre"], hb["bi"], hb["l2"], hb["n1"], hb["n2"], nb * n23 + 64, nb + 16)
        prm = gpv.Params(l1, l2, gpv.GPV_NORMALS)
        times = []
        for _ in range(reps + 1):
            t0 = time.perf_counter()
            r = ctx.voxelize_host(mesh, prm, hs)
            times.append(time.perf_counter() - t0)
        l1s = np.ctypeslib.as_array(C.cast(hb["l1"], C.POINTER(C.c_uint8)), shape=(cells,)) // 127
        l2s = np.ctypeslib.as_array(C.cast(hb["l2"], C.POINTER(C.c_uint8)), shape=(nb * n23,)) // 127
        out.update(seconds=times[1:], first_call_seconds=times[0], counts=r.counts, l1=sha(l1s.astype(np.uint8)), l2=sha(l2s.astype(np.uint8)),
                   note="normals included (the reference always computes them)")
    else:
        from oracle import refbind
        # "emulated": the reference's kernel source executed on the host (oracle/ref_kernels_host.cpp) -- no GPU, checks this tool's plumbing
        refbind.LIB_PATH = os.path.join(ROOT, "oracle", "_ref", {"reference": "libgpvref_refcuda.so", "compat": "libgpvref_b200.so", "emulated": "libgpvref_emu.so"}[kind])
        refbind._lib = None
        L = refbind.lib()
        L.ref_cuda_path.argtypes = [C.c_void_p]
        L.ref_cuda_path.restype = C.c_int
        fill = O.OracleMesh(path).voxelize(l1, l2, O.FILL_CERTIFIED | O.NO_L2 | O.NO_NORMALS, os.cpu_count() or 8).l1_fill_only.astype(np.float32)
        times, used = [], 0
        for _ in range(reps + 1):
            ro = refbind.RefObject(path)
            ro.setup(l1, l2)
            C.memmove(L.ref_level1InOut(ro.h), fill.ctypes.data, fill.nbytes)  # the GL fill's stand-in
            t0 = time.perf_counter()
            used = L.ref_cuda_path(ro.h)
            times.append(time.perf_counter() - t0)
            l1s, l2s, cnt = ro.level1_inout().astype(np.uint8), ro.level2_inout().astype(np.uint8), ro.count()
            ro.close()
        out.update(seconds=times[1:], first_call_seconds=times[0], counts=cnt, l1=sha(l1s), l2=sha(l2s), tri_buffer=used)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="cessna")
    ap.add_argument("--l1", type=int, default=64)
    ap.add_argument("--l2", type=int, default=4)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--emulated-only", action="store_true", help="no GPU: run the reference's host code on its host-executed kernel source only (plumbing check)")
    ap.add_argument("--worker", default=None)
    ap.add_argument("--path", default=None)
    a = ap.parse_args()
    if a.worker:
        return worker(a.worker, a.path, a.l1, a.l2, a.reps)
    path = mesh_file(a.mesh, tempfile.mkdtemp(prefix="gpvref"))
    rows = {}
    for kind in (("emulated",) if a.emulated_only else ("reference", "compat", "native")):
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", kind, "--path", path, "--l1", str(a.l1), "--l2", str(a.l2), "--reps", str(a.reps)],
                           capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
        rows[kind] = json.loads(line[-1]) if line else {"kind": kind, "error": (r.stderr or r.stdout)[-400:]}
    best = {k: min(v["seconds"]) for k, v in rows.items() if "seconds" in v}
    summary = {"config": "%s Level1 %d + Level2 %d^3, normals on" % (a.mesh, a.l1, a.l2), "best_seconds": best, "runs": rows}
    if "native" in best:
        summary["speedup_over"] = {k: best[k] / best["native"] for k in best if k != "native"}
    if "native" in rows and all("l1" in v for v in rows.values()):
        summary["states_agree"] = {k: (rows[k]["l1"] == rows["native"]["l1"] and rows[k]["l2"] == rows["native"]["l2"] and rows[k]["counts"] == rows["native"]["counts"])
                                   for k in rows if k != "native"}
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
