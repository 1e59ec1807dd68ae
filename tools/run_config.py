#!/usr/bin/env python3
"""Run one benchmark configuration on cuda:0 and (optionally) check every stream against the CPU oracle.

  python tools/run_config.py --mesh sphere|torus|cad|cessna|block --l1 512 --l2 8 [--check] [--reps 5] [--normals] [--slabs R]

Synthetic meshes come from gpview_b200.meshgen (seeded, float32); they are handed over as triangle arrays
(gpv_mesh_from_triangles: bbox over the vertices + the reference's padding), which is what the OBJ loader yields for a mesh
whose vertices are all referenced.  --check is test infrastructure (loads oracle/): hashes of the oracle's streams must
equal the GPU's.  --slabs R additionally runs R z-slabs on the same device and checks that they concatenate to the whole.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_tris(name, scale=1):
    from gpview_b200 import meshgen as M
    if name == "sphere":
        V, F = M.uv_sphere(1000 // scale, 502 // scale if scale > 1 else 502)
    elif name == "torus":
        V, F = M.torus(1000 // scale, 500 // scale)
    elif name == "cad":
        V, F = M.cad_body(2500 // scale, 2001 // scale if scale > 1 else 2001)
    elif name == "block":
        V, F = M.drilled_block(n_seg=220, n_grid=40)
    elif name == "cessna":
        z = np.load(os.path.join(ROOT, "tests", "golden", "cessna_mesh.npz"))
        V, F = z["V"], z["F"]
    else:
        raise SystemExit("unknown mesh")
    return M.triangles(V, F)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="sphere")
    ap.add_argument("--scale", type=int, default=1, help="divide the mesh resolution (quick runs)")
    ap.add_argument("--l1", type=int, default=512)
    ap.add_argument("--l2", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--normals", action="store_true")
    ap.add_argument("--slabs", type=int, default=0)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    a = ap.parse_args()
    import gpview_b200 as gpv
    t0 = time.time()
    tris = make_tris(a.mesh, a.scale)
    mesh = gpv.mesh_from_triangles(tris)
    print("mesh %s: %d triangles (%.1fs to build)" % (a.mesh, mesh.ntri, time.time() - t0), flush=True)
    ctx = gpv.Context(0)
    d = ctx.upload(mesh)
    flags = gpv.GPV_PROFILE | (gpv.GPV_NORMALS if a.normals else 0)
    res = None
    times = []
    for i in range(a.reps):
        t0 = time.perf_counter()
        res = ctx.voxelize_device(d, mesh, gpv.Params(a.l1, a.l2, flags))
        times.append((time.perf_counter() - t0) * 1e3)
    ph = {k: round(v, 3) for k, v in res.phase_ms.items() if v > 0}
    out = {"mesh": a.mesh, "ntri": mesh.ntri, "l1": a.l1, "l2": a.l2, "grid": [int(x) for x in res.num_div], "cells": res.cells, "n_boundary": res.nb,
           "l2_voxels": res.nb * res.n23, "counts": res.counts, "stats": res.stats, "wall_ms": [round(t, 3) for t in times], "device_ms": round(sum(ph.values()), 3),
           "phase_ms": ph}
    tests = res.stats["l1_box_tests"] + res.stats["l2_box_tests"]
    out["g_tribox_tests_per_s"] = tests / (min(times) * 1e-3) / 1e9
    print(json.dumps(out), flush=True)
    if a.slabs > 1:
        whole = {"l1": res.level1_inout(), "pre": res.prefix(), "bi": res.boundary_index(), "l2": res.level2_inout()}
        nz = int(res.num_div[2])
        parts = {k: [] for k in whole}
        base = 0
        for r in range(a.slabs):
            s = ctx.voxelize_device(d, mesh, gpv.Params(a.l1, a.l2, 0, nz * r // a.slabs, nz * (r + 1) // a.slabs))
            parts["l1"].append(s.level1_inout()); parts["pre"].append(s.prefix() + base); parts["bi"].append(s.boundary_index()); parts["l2"].append(s.level2_inout())
            base += s.nb
        ok = all(np.array_equal(np.concatenate(parts[k]), whole[k]) for k in whole)
        print("slabs x%d concatenate to the whole grid: %s" % (a.slabs, ok), flush=True)
        if not ok:
            raise SystemExit(1)
    if a.check:
        from oracle import oraclebind as O
        t0 = time.time()
        om = O.OracleMesh(tris=tris)
        assert np.array_equal(om.bmin, mesh.bbox_min) and np.array_equal(om.bmax, mesh.bbox_max)
        ores = om.voxelize(a.l1, a.l2, O.FILL_CERTIFIED | (0 if a.normals else O.NO_NORMALS), a.threads)
        print("oracle: %.1fs on %d threads, counts %s" % (time.time() - t0, a.threads, ores.counts), flush=True)
        res = ctx.voxelize_device(d, mesh, gpv.Params(a.l1, a.l2, flags | gpv.GPV_KEEP_LISTS))  # canonical list order only on request
        checks = {"counts": res.counts == ores.counts, "l1": sha(res.level1_inout()) == sha(ores.l1_state * 127), "prefix": sha(res.prefix()) == sha(ores.prefix),
                  "boundary_index": sha(res.boundary_index()) == sha(ores.boundary_index), "l2": sha(res.level2_inout()) == sha(ores.l2_state * 127),
                  "cell_lists": sha(res.cell_tris()) == sha(ores.cell_tris), "l1_tests": res.stats["l1_box_tests"] == ores.stats["l1BoxTests"],
                  "ill": res.stats["fill_ill_conditioned"] == ores.stats["fillIllConditioned"], "crossings": res.stats["fill_crossings"] == ores.stats["fillCrossings"]}
        if a.normals:
            checks["n1"] = sha(res.level1_normal()) == sha(ores.l1_normal)
            checks["n2"] = sha(res.level2_normal()) == sha(ores.l2_normal)
        print("parity vs oracle:", checks, flush=True)
        if not all(checks.values()):
            raise SystemExit(2)
    ctx.free_device(d)
    ctx.close()


if __name__ == "__main__":
    main()
