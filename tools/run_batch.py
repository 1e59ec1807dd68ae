#!/usr/bin/env python3
"""Config 5 (BASELINE.json): batched dataset generation -- N drilled-block .off meshes (~5k triangles) at Level-1 64 + Level-2 4^3.

  python tools/run_batch.py [--models 2000] [--distinct 100] [--threads 16] [--gpus 1] [--save] [--check 5] [--parse-only]

Writes `distinct` seeded meshes (gpview_b200.meshgen.drilled_block, seed = SEED_BASE + i) once, then runs gpv_voxelize_batch over
`models` paths cycling through them (every path is parsed again: parsing is part of the pipeline).  --check K compares the
files of the first K models with the oracle's writer (test infrastructure).  Prints one JSON line with models/s."""
import argparse
import filecmp
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def write_meshes(d, distinct):
    from gpview_b200 import meshgen as M
    files = []
    for i in range(distinct):
        V, F = M.drilled_block(seed=M.SEED_BASE + i, n_seg=160, n_grid=36)
        p = os.path.join(d, "block%05d.off" % i)
        M.write_off(p, V, F)
        files.append(p)
    return files


def run_procs(a):
    """--procs P: P copies of this tool, each on its share of the models; wall clock from the moment all are started (their
    warm-up batches included, so take --models large) to the last one's exit."""
    import subprocess
    d = tempfile.mkdtemp(prefix="gpvbatch")
    write_meshes(d, a.distinct)
    base = [sys.executable, os.path.abspath(__file__), "--models", str(a.models), "--distinct", str(a.distinct), "--threads", str(a.threads), "--gpus", str(a.gpus),
            "--l1", str(a.l1), "--l2", str(a.l2), "--mesh-dir", d] + (["--normals"] if a.normals else []) + (["--save"] if a.save else [])
    t0 = time.time()
    ps = [subprocess.Popen(base + ["--share", "%d/%d" % (k, a.procs)], stdout=subprocess.PIPE, text=True) for k in range(a.procs)]
    outs = [p.communicate()[0] for p in ps]
    dt = time.time() - t0
    rows = [json.loads(o.strip().splitlines()[-1]) for o in outs if o.strip()]
    print(json.dumps({"config": "drilled-block .off meshes, %d processes x %d threads on %d GPU(s)" % (a.procs, a.threads, a.gpus), "models": a.models,
                      "wall_seconds_incl_startup": dt, "models_per_s_sum_of_processes": sum(r["models_per_s"] for r in rows),
                      "per_process": [{k: r[k] for k in ("models", "models_per_s", "seconds", "failed")} for r in rows]}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", type=int, default=2000)
    ap.add_argument("--distinct", type=int, default=100)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--l1", type=int, default=64)
    ap.add_argument("--l2", type=int, default=4)
    ap.add_argument("--save", action="store_true")
    ap.add_argument("--normals", action="store_true")
    ap.add_argument("--six-files", action="store_true", help="without --normals: still write the two normal files (127-filled), as gpv_save does")
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--parse-only", action="store_true", help="host side alone (no GPU needed): models/s the loaders can feed on --threads host threads")
    ap.add_argument("--procs", type=int, default=1, help="experiment: P worker processes (own CUDA contexts) of --threads threads each, every process runs "
                    "gpv_voxelize_batch on its share of the models; one process saturates its own launch path at ~7 k models/s per GPU")
    ap.add_argument("--mesh-dir", default=None, help=argparse.SUPPRESS)   # worker processes of --procs: meshes already written here
    ap.add_argument("--share", default=None, help=argparse.SUPPRESS)      # "k/P": this worker's share of the model list
    a = ap.parse_args()
    if a.procs > 1 and not a.share:
        return run_procs(a)
    import gpview_b200 as gpv
    from gpview_b200 import binding as B, meshgen as M
    d = a.mesh_dir or tempfile.mkdtemp(prefix="gpvbatch")
    t0 = time.time()
    files = write_meshes(d, a.distinct) if not a.mesh_dir else [os.path.join(d, "block%05d.off" % i) for i in range(a.distinct)]
    gen_s = time.time() - t0
    paths = [files[i % a.distinct] for i in range(a.models)]
    if a.share:
        k, P = [int(x) for x in a.share.split("/")]
        paths = paths[k::P]
    if a.parse_only:
        import threading
        nxt, lock, tris = [0], threading.Lock(), [0]

        def worker():
            n = 0
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= len(paths):
                    break
                n += gpv.load_mesh(paths[i]).ntri   # ctypes releases the GIL inside gpv_load_mesh
            with lock:
                tris[0] += n
        t0 = time.time()
        ts = [threading.Thread(target=worker) for _ in range(a.threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        dt = time.time() - t0
        print(json.dumps({"config": "parse only: drilled-block .off meshes (~5k triangles)", "models": a.models, "threads": a.threads, "models_per_s": a.models / dt,
                          "ms_per_model_per_thread": 1e3 * dt * a.threads / a.models, "triangles": tris[0], "mesh_generation_s": gen_s}))
        return
    out = os.path.join(d, "out") if (a.save or a.check) else None
    if out:
        os.makedirs(out)
    flags = gpv.GPV_NORMALS if a.normals else (0 if a.six_files else gpv.GPV_SAVE_COMPUTED_ONLY)
    B.voxelize_batch(paths[:min(len(paths), 4 * a.threads)], gpv.Params(a.l1, a.l2, flags), list(range(a.gpus)), a.threads, None)  # warm-up: contexts, pools
    st = B.voxelize_batch(paths, gpv.Params(a.l1, a.l2, flags), list(range(a.gpus)), a.threads, out, 0, False)
    line = {"config": "drilled-block .off meshes (~5k triangles), Level1 %d + Level2 %d^3" % (a.l1, a.l2), "models": len(paths), "distinct_meshes": a.distinct,
            "threads": a.threads, "gpus": a.gpus, "saved": bool(out), "normals": a.normals, "models_per_s": st["models_done"] / st["seconds"],
            "seconds": st["seconds"], "per_model_ms_summed_over_threads": {k: 1e3 * st[k + "_seconds"] / max(1, st["models_done"]) for k in ("parse", "gpu", "save")},
            "mesh_generation_s": gen_s, "failed": st["models_failed"]}
    if a.check:
        from oracle import oraclebind as O
        ok = True
        for i in range(min(a.check, a.models)):
            ref = os.path.join(d, "ref%d" % i)
            os.makedirs(ref)
            O.OracleMesh(paths[i]).voxelize(a.l1, a.l2, O.FILL_CERTIFIED | (0 if a.normals else O.NO_NORMALS), 2).save(i, ref)
            for n in os.listdir(ref):
                if not a.normals and "Normal" in n:
                    continue
                ok = ok and filecmp.cmp(os.path.join(ref, n), os.path.join(out, n), shallow=False)
        line["files_match_oracle"] = ok
        # restart: nothing is recomputed when the outputs exist
        st2 = B.voxelize_batch(paths, gpv.Params(a.l1, a.l2, flags), list(range(a.gpus)), a.threads, out, 0, True)
        line["restart_skipped"] = st2["models_skipped"]
    print(json.dumps(line))
    if a.check and not line["files_match_oracle"]:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
