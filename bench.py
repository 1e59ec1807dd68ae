#!/usr/bin/env python3
"""bench.py -- headline benchmark of the voxelizer hot path (BASELINE.json: cessna Level-1 256 + Level-2 16^3).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--l1 256 --l2 16 --mesh cessna|sphere|torus|cad] [--normals] [--batch M]

One "step" = one full voxelization of the model (Level-1 SAT binning, parity fill, boundary compaction, Level-2
refinement), triangles resident in HBM when the timed region starts, outputs resident in HBM when it ends (N > 1: on rank 0,
as file bytes).  Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (k_l2): algorithmic tri-box FLOPs (124 per reference-equivalent test) over
                its CUDA-event time (events recorded around the kernel on the launching stream, inside the timed steps),
                against the non-FMA FP32 issue rate measured live by gpv_measure_fp32_peak; beside the algorithmic `frac` the
                utilisation figures `frac_executed` / `frac_issue_slots` from the committed ncu capture of the same command
                (profiles/traffic.json, quoted only while the kernel sources hash to what it was captured from)
  roofline_hbm  the same launch's output bytes against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own object code (oracle/_ref) -- or the oracle port when _ref is absent -- through the WHOLE CPU
                path (Level-1 tri-box loop, column lists, column-list fill, Level-2 rays + SAT) on a bounded sample of boundary
                cells, all host threads (rank 0, N = 1); --impl reference prints the same as its own line, plus configs[0] (`c1`)
  e2e           the same metric through gpv_voxelize_host: pinned HOST triangles in, HOST streams out, copies timed -- Level 2
                over PCIe as bytes and as 2 bits per sub-voxel expanded by the library's host threads; the faster is the headline.
                N > 1: every rank delivers the byte range of its z-slab to its own pinned host buffer over its own PCIe
                link (ranges are disjoint and ordered, prefix sums made global with the all-gathered boundary counts)
N > 1 (torchrun), GPV_GATHER: Level 1 is replicated (every rank needs whole column lists); every rank delivers a z-slab of the
Level-1 bytes / prefix sums and the Level-2 blocks of the Level-1 columns it owns (interleaved groups) straight into rank 0's
buffers over NVLink peer memory -- no collective; --gather nccl: the NCCL send/recv gather of cost-balanced z-slabs it replaces.
--batch M: BASELINE.json configs[4], models/s through gpv_voxelize_batch (one rank per GPU, no collective).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_TRIBOX = 124  # 66 MUL + 58 ADD/SUB, SURVEY.md 8(a8)


def make_mesh_file(name, tmp):
    from gpview_b200 import meshgen
    p = os.path.join(tmp, name + ".obj")
    if name == "cessna":
        z = np.load(os.path.join(ROOT, "tests", "golden", "cessna_mesh.npz"))
        meshgen.write_obj(p, z["V"], z["F"])
    elif name == "sphere":
        meshgen.write_obj(p, *meshgen.uv_sphere(1000, 502))
    elif name == "torus":
        meshgen.write_obj(p, *meshgen.torus(1000, 500))
    elif name == "cad":
        meshgen.write_obj(p, *meshgen.cad_body(2500, 2001))
    else:
        raise SystemExit("unknown mesh " + name)
    return p


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the GPU runs the benchmark steps (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.p = gpu, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def count(self, t0):
        return sum(1 for t, r in self.rows if t >= t0 and len(r) > 8)

    def stop(self, t0, t1, extended):
        if self.p:
            self.p.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) > 8]
        num = lambda s: float(s) if s.replace(".", "", 1).isdigit() else None
        sm = [num(r[1]) for r in rows if num(r[1]) is not None]
        mx = [num(r[2]) for r in rows if num(r[2]) is not None]
        pw = [num(r[3]) for r in rows if num(r[3]) is not None]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw) if pw else None,
                "window": "timed steps" + (" + %d extra untimed steps of the same work (the timed region is shorter than the sampling period)" % extended if extended else "")}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs"), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """(per-kernel numbers of the committed ncu --set full capture, note): profiles/traffic.json, written by tools/ncu_summary.py
    --traffic and stamped with a hash of the kernel sources.  A capture of OTHER sources is not quoted: ({}, why)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return {}, "no capture committed"
    with open(p) as f:
        d = json.load(f)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ncu_summary
    sha = ncu_summary.csrc_sha()
    if d.get("csrc_sha") != sha:
        return {}, "stale: profiles/traffic.json was captured from kernel sources %s, this build is %s" % (d.get("csrc_sha"), sha)
    return d, "ncu --set full capture of this build's kernel sources (csrc_sha %s), same command, one GPU" % sha


class ReferencePath:
    """The reference's CPU implementation of the WHOLE path on this box's host cores (oracle/_ref: the reference's own object code --
    Object::ClassifyTessellation, TriBoxOverlap, triangle_ray_intersection -- driven by oracle/ref_harness.cpp):
      Level 1   tri-box binning (Object::ClassifyTessellation, src/Object.cpp:2256-2342: its own single-threaded loop), the host CSR /
                column-list pass (:2137-2180, :3258-3284) and the solid fill through the column lists (all host threads)
      Level 2   per sub-voxel the parity rays over the column list (cu:450-504) and the SAT over the cell list (cu:403-448), in the
                kernels' arithmetic, all host threads -- timed on a bounded sample of boundary cells
    A step of the sample charges Level 1 pro rata: seconds = t_L1 * K / nB + t_L2(K cells), tests = L1 tests * K / nB + L2 tests(K).
    When oracle/_ref was not built the oracle port's Level-2 SAT nest stands in (kind "port")."""

    def __init__(self, mesh_path, l1, l2, threads):
        from oracle import refbind
        self.threads, self.l1, self.l2 = threads, l1, l2
        if refbind.available():
            self.kind = "reference"
            o = self.o = refbind.RefObject(mesh_path)
            o.setup(l1, l2)
            self.t_tri = o.l1_tribox()                       # the reference's own loop (single-threaded by construction)
            t0 = time.perf_counter()
            o.compact()
            self.t_lists = time.perf_counter() - t0
            self.t_fill, self.fill_ray_tests = o.time_l1_fill_collist(threads)
            st = o.stats()
            self.l1_box_tests, self.nb = st["l1_box_tests"], o.nboundary()
            self.grid = [int(x) for x in o.num_div]
            self.triangles = o.ntri
            self.l2_box_tests_model = int(o.tri_count()[o.boundary_index()].astype(np.int64).sum()) * max(l2, 1) ** 3
        else:
            from oracle import oraclebind as O
            self.kind = "port"
            self.r = O.OracleMesh(mesh_path).voxelize(l1, l2, O.FILL_COLLIST | O.NO_L2 | O.NO_NORMALS, threads)
            self.t_tri = self.t_lists = self.t_fill = 0.0
            self.fill_ray_tests = 0
            self.l1_box_tests, self.nb = 0, self.r.nb
            self.grid, self.triangles, self.l2_box_tests_model = [int(x) for x in self.r.num_div], None, None
        self.t_l1 = self.t_tri + self.t_lists + self.t_fill

    def sample(self, cells, threads=None):
        """(seconds, tri-box tests, ray tests) of one step over the first `cells` boundary cells, Level 1 charged pro rata."""
        threads = threads or self.threads
        if self.kind == "reference":
            s, box, rays = self.o.time_l2_path(0, cells, threads)
        else:
            s, box = self.r.time_l2_tribox(0, cells, threads)
            rays = 0
        f = cells / max(1, self.nb)
        return s + self.t_l1 * f, box + self.l1_box_tests * f, rays + self.fill_ray_tests * f

    def cells_for(self, seconds):
        probe = min(self.nb, 64)
        s, _, _ = self.sample(probe)
        return int(max(probe, min(self.nb, seconds / max(s / probe, 1e-9))))

    def describe(self, cells):
        if self.kind == "port":
            return "oracle port (oracle/_ref not built): Level-2 SAT loop nest over the first %d of %d boundary cells" % (cells, self.nb)
        return ("whole CPU path: Level 1 measured on the whole model (ClassifyTessellation %.3f s on 1 thread -- the reference's own loop --, CSR / column lists %.3f s, "
                "column-list fill %.3f s on %d threads) and charged pro rata; Level 2 (parity rays + SAT per sub-voxel) over the first %d of %d boundary cells per step"
                % (self.t_tri, self.t_lists, self.t_fill, self.threads, cells, self.nb))


def c1_reference(mesh_path, threads):
    """BASELINE.json configs[0]: cessna Level-1 64 through the reference's CPU path -- brute-force Object::ClassifyInOutCPU loop nest
    (src/Object.cpp:716-779, z-layers over host threads) + Object::ClassifyTessellation."""
    from oracle import refbind
    if not refbind.available():
        return None
    o = refbind.RefObject(mesh_path)
    o.setup(64, 0)
    t_fill = o.l1_inout_brute(threads)
    t_tri = o.l1_tribox()
    st = o.stats()
    out = {"workload": "cessna Level1 64, CPU path (BASELINE.json configs[0])", "ms_per_model": 1e3 * (t_fill + t_tri), "fill_brute_force_s": t_fill,
           "fill_ray_tests": int(o.cells) * int(o.ntri), "tri_box_s": t_tri, "tri_box_tests": st["l1_box_tests"], "cores": threads}
    o.close()
    return out


def cpu_baseline(mesh_path, l1, l2, threads, target_seconds=12.0):
    ref = ReferencePath(mesh_path, l1, l2, threads)
    cells = ref.cells_for(target_seconds)
    s, box, rays = ref.sample(cells)
    out = {"value": box / s / 1e9, "unit": "G tri-box tests/s", "cores": threads, "kind": ref.kind, "sample": ref.describe(cells),
           "seconds": s, "ray_tests_in_sample": int(rays)}
    if ref.kind == "reference" and ref.nb:
        out["ms_per_model_extrapolated"] = 1e3 * s * ref.nb / cells
    try:  # the reference's actual execution model is one thread -- reported beside the all-core figure
        c1 = max(16, cells // max(1, threads) // 2)
        s1, b1, _ = ref.sample(c1, 1)
        out["single_thread"] = {"value": b1 / s1 / 1e9, "unit": "G tri-box tests/s", "cores": 1, "sample": "the same path over the first %d boundary cells (%.2f s)" % (c1, s1)}
    except Exception as e:
        out["single_thread"] = {"error": repr(e)}
    return out


def workload_config(args, tests=None, triangles=None, grid=None):
    """The workload, identically for both arms (the reference arm derives the same numbers from the reference's own data structures)."""
    return {"workload": "%s Level1 %d + Level2 %d^3 (BASELINE.json configs[1] when cessna/256/16), .raw occupancy streams"
                        % (args.mesh, args.l1, args.l2), "l1": args.l1, "l2": args.l2, "mesh": args.mesh,
            "cache": "L2 flushed between timed steps (256 MiB memset); outputs (>=221 MB at 256/16) exceed L2",
            "tri_box_tests_per_model": tests, "triangles": triangles, "grid": grid}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the whole path on this box's host cores (all threads)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    path = make_mesh_file(args.mesh, tempfile.mkdtemp(prefix="gpvbench"))
    ref = ReferencePath(path, args.l1, args.l2, threads)
    cells = ref.cells_for(4.0)  # ~4 s of CPU work per step
    for _ in range(args.warmup):
        ref.sample(max(64, cells // 8))
    tot_s, tot_n, tot_r = 0.0, 0.0, 0.0
    for _ in range(args.steps):
        s, n, r = ref.sample(cells)
        tot_s += s; tot_n += n; tot_r += r
    v = tot_n / tot_s / 1e9
    tests_model = (ref.l1_box_tests + ref.l2_box_tests_model) if ref.l2_box_tests_model is not None else None
    line = {"impl": "reference", "metric": "G tri-box tests/s", "value": v, "unit": "G tri-box tests/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "cessna.obj fixture (reference sample mesh)" if args.mesh == "cessna" else "synthetic",
            "config": workload_config(args, tests_model, ref.triangles, ref.grid),
            "cpu_baseline": {"value": v, "unit": "G tri-box tests/s", "cores": threads, "kind": ref.kind, "sample": ref.describe(cells)},
            "e2e": {"value": v, "unit": "G tri-box tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "parallelism": "%d host threads (Object::ClassifyTessellation itself on one: the reference's own loop)" % threads,
            "ray_tests_per_step": tot_r / args.steps}
    if ref.kind == "reference" and ref.nb:
        line["ms_per_model_extrapolated"] = 1e3 * (tot_s / args.steps) * ref.nb / cells
    if args.mesh == "cessna":
        try:
            line["c1"] = c1_reference(path, threads)
        except Exception as e:  # commentary: never lose the line over it
            line["c1"] = {"error": repr(e)}
    print(json.dumps(line))


def run_batch_arm(args):
    """--batch MODELS: config 5.  Every rank (one per GPU) runs gpv_voxelize_batch over its share of the model files on its own GPU
    with its share of the host threads; models are independent, so there is no collective (replicas only).  value = models of all
    ranks / the slowest rank's wall time.  --impl reference: the reference's CPU path (oracle/_ref) over a bounded sample of the models."""
    import shutil
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    l1, l2 = (64, 4) if (args.l1, args.l2) == (256, 16) else (args.l1, args.l2)   # the config's own resolution unless given
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import run_batch
    distinct = min(100, args.batch)
    d = tempfile.mkdtemp(prefix="gpvbatch%d_" % rank)
    files = run_batch.write_meshes(d, distinct)
    cfg = {"workload": "batched dataset generation: %d drilled-block .off meshes (~5 k triangles, %d distinct, seeded) at Level1 %d + Level2 %d^3 "
                       "(BASELINE.json configs[4]); per model: parse + H2D + voxelize + D2H%s" % (args.batch, distinct, l1, l2, " + file write" if args.batch_save else ""),
           "l1": l1, "l2": l2, "mesh": "drilled blocks", "models": args.batch, "cache": "every model is a different file and a fresh set of device streams (no reuse between models)"}
    threads_all = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        n = min(args.batch, 12)
        t0 = time.perf_counter()
        for k in range(n):
            ref = ReferencePath(files[k % distinct], l1, l2, threads_all)
            ref.sample(ref.nb)
            if ref.kind == "reference":
                ref.o.close()
        dt = time.perf_counter() - t0
        print(json.dumps({"impl": "reference", "metric": "models/s", "value": n / dt, "unit": "models/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": 1e3 * dt,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": n / dt, "unit": "models/s", "cores": threads_all, "kind": "reference", "sample": "%d models through the whole CPU path (oracle/_ref)" % n},
                          "e2e": {"value": n / dt, "unit": "models/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        shutil.rmtree(d, ignore_errors=True)
        return
    import torch
    import gpview_b200 as gpv
    from gpview_b200 import binding as B
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    paths = [files[i % distinct] for i in range(args.batch)][rank::world]
    threads = max(1, min(8, threads_all // world))   # more than ~8 feeders per GPU contend on one process's driver-call path (measured: 8 -> 7.5 k, 16 -> 4.0 k models/s)
    out = os.path.join(d, "out") if args.batch_save else None
    if out:
        os.makedirs(out)
    prm = gpv.Params(l1, l2, gpv.GPV_SAVE_COMPUTED_ONLY)
    B.voxelize_batch(paths[:min(len(paths), 4 * threads)], prm, [local], threads, None)   # warm-up: contexts, pools
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    st = B.voxelize_batch(paths, prm, [local], threads, out, 0, False)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    nn = torch.tensor([float(st["models_done"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn)
    if rank == 0:
        clocks = sampler.stop(t0, t1, 0)
        total, secs = float(nn.item()), float(tt.item())
        per = {k: 1e3 * st[k + "_seconds"] / max(1, st["models_done"]) for k in ("parse", "gpu", "save")}
        print(json.dumps({"metric": "models/s", "value": total / secs, "unit": "models/s", "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": 1e3 * secs,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": dict(cfg, parallelism="%d GPU(s) x %d host threads each, one gpv_ctx per thread, models round-robin; no collective" % (world, threads)),
                          "clocks": clocks, "gpu_launches": int(total) * 16,
                          "e2e": {"value": total / secs, "unit": "models/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                  "note": "the batch call IS end to end: model files in, host streams%s out" % (" and files" if out else "")},
                          "per_model_ms_summed_over_threads_rank0": per, "level2_resizes_rank0": st["level2_resizes"], "failed": int(st["models_failed"])}))
    shutil.rmtree(d, ignore_errors=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--mesh", default="cessna")
    ap.add_argument("--l1", type=int, default=256)
    ap.add_argument("--l2", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, metavar="MODELS",
                    help="BASELINE.json configs[4]: batched dataset generation -- MODELS drilled-block .off meshes (~5 k triangles) at --l1 64 --l2 4 through "
                         "gpv_voxelize_batch (parse + H2D + voxelize + D2H per model), models/s over all GPUs; one step = the whole batch")
    ap.add_argument("--batch-save", action="store_true", help="--batch: also write every model's files (GPV_SAVE_COMPUTED_ONLY) into a temporary directory")
    ap.add_argument("--normals", action="store_true", help="also produce Level1Normal / Level2Normal (the reference's full six-file contract, SURVEY.md 8f1): 3 B per cell / sub-voxel more")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N > 1: slabs written into rank 0's buffers over NVLink peer memory from inside the kernels (GPV_GATHER), or gathered with NCCL send/recv")
    args = ap.parse_args()
    if args.batch:
        return run_batch_arm(args)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import gpview_b200 as gpv
    from gpview_b200 import binding as B
    from gpview_b200 import sharded

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # e2e leg: every rank expands its own slab's 2-bit Level-2 transfer; share the host cores instead of 15 spinning threads per rank
        os.environ.setdefault("GPV_HOST_THREADS", str(max(2, (os.cpu_count() or 8) // world - 1)))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- libgpview_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.mesh == "cessna":  # through the reference-semantics OBJ loader, like the reference does
        path = make_mesh_file(args.mesh, tempfile.mkdtemp(prefix="gpvbench"))
        mesh = gpv.load_mesh(path)
    else:  # large synthetic meshes are handed over as triangle arrays (same bbox rule as the loader; writing a 1 GB OBJ is not the point)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import run_config
        path = None
        mesh = gpv.mesh_from_triangles(run_config.make_tris(args.mesh))
    ctx = gpv.Context(local)
    L = gpv.lib()
    stream = torch.cuda.current_stream()
    sptr = C.c_void_p(stream.cuda_stream)
    d_tris = ctx.upload(mesh)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # ---- z-slab plan (N > 1): the replicated Level-1 pre-pass gives the Level-2 cost per layer; cuts are deterministic
    z0, z1 = 0, 0
    whole = ctx.voxelize_device(d_tris, mesh, gpv.Params(args.l1, args.l2, gpv.GPV_NO_LEVEL2), sptr)
    nz = int(whole.num_div[2]); plane = int(whole.num_div[0]) * int(whole.num_div[1])
    cuts = [0, nz]
    layer0 = None
    if world > 1:  # (the pre-pass result views die with the next call on the ctx: take the cost model now)
        layer0 = sharded.layer_cost(whole.boundary_index(), whole.cell_off(), plane, nz).astype(np.float64)
        cuts = sharded.plan_slabs(layer0, world)
        z0, z1 = cuts[rank], cuts[rank + 1]
    peer = world > 1 and args.gather == "peer"
    n23 = args.l2 ** 3
    if peer:  # rank 0 owns the whole-grid streams; the other ranks map them (CUDA IPC) and write their slabs over NVLink
        desc = B.CGatherDesc()
        if rank == 0:
            desc = ctx.gather_create(plane * nz, int(whole.nb) * n23, gpv.GPV_NORMALS if args.normals else 0)
        t = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        desc = B.CGatherDesc.from_buffer_copy(bytes(t.cpu().numpy().tobytes()))
        ctx.gather_attach(desc, rank, world)
    # peer gather: the library shares the work out itself (Level-1 bytes by equal z-slabs, Level-2 by interleaved column groups); the
    # cost-balanced z cuts above are what the NCCL gather and the e2e leg use
    params = gpv.Params(args.l1, args.l2, gpv.GPV_PROFILE_L2 | (gpv.GPV_GATHER if peer else 0) | (gpv.GPV_NORMALS if args.normals else 0), 0 if peer else z0, 0 if peer else z1)  # timed steps: CUDA events around the two Level-2 kernels only, on the launching stream
    gathered = {}

    # the timed call goes straight to the C ABI with arguments built once (a Python wrapper object per call costs tens of microseconds
    # of host time inside a ~1 ms step); the result struct is wrapped outside the timed region
    c_res = B.CResult()
    c_args = (ctx.h, d_tris, mesh.ntri, B._fp(mesh.bbox_min), B._fp(mesh.bbox_max), float(mesh.max_model_size), C.byref(params.c), sptr, C.byref(c_res))
    voxelize_device = L.gpv_voxelize_device

    def step(wrap=True):
        if voxelize_device(*c_args):
            raise SystemExit(L.gpv_last_error().decode())
        if world > 1 and not peer:  # slab pieces -> rank 0 over NCCL (concatenation only), inside the timed region
            res = B.Result(c_res, ctx)
            w = lambda ptr, n: sharded.wrap_device_bytes(torch, ptr, n)
            pieces = {"l1": (w(res.c.d_level1_inout, res.cells), 1, 0), "prefix": (w(res.c.d_prefix, res.cells * 4), 4, 0),
                      "l2": (w(res.c.d_level2_inout, res.nb * res.n23), res.n23, 1)}
            sharded.gather_to_rank0(dist, torch, rank, world, pieces, res.cells, res.nb, gathered)
            return res
        return B.Result(c_res, ctx) if wrap else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        res = step()
    barrier()
    if world > 1 and not peer:
        # Untimed load balancing (NCCL gather of z-slabs only): the a-priori cost model only knows list lengths.  Rescale every slab's layer costs to the device
        # time its rank measured (every phase but the final wait for the other ranks), re-cut, repeat; keep the best cuts seen.
        # Every rank derives the same cuts from the all-gathered times (no data moves).
        layer = layer0 + 1e-9

        def measured(r):
            t = torch.tensor([sum(v for k, v in r.phase_ms.items() if k not in ("l2_normals", "host_gap"))], device="cuda", dtype=torch.float64)
            out = torch.zeros(world, device="cuda", dtype=torch.float64)
            dist.all_gather_into_tensor(out, t)
            return out.cpu().numpy()

        def set_cuts(c):
            params.c.z0, params.c.z1 = c[rank], c[rank + 1]

        t_all = measured(res)
        best = (float(t_all.max()), list(cuts))
        for _ in range(4):
            if t_all.max() < 1.05 * t_all.mean():
                break
            layer, cuts = sharded.rebalance(layer, cuts, t_all)
            set_cuts(cuts)
            for _ in range(2):
                res = step()
            barrier()
            t_all = measured(res)
            if float(t_all.max()) < best[0]:
                best = (float(t_all.max()), list(cuts))
        cuts = best[1]
        set_cuts(cuts)
        z0, z1 = cuts[rank], cuts[rank + 1]
        res = step()
        barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches, phase_acc = 0, {}
    t_timed0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)
        barrier()
        ev[k][0].record(stream)
        step(False)
        ev[k][1].record(stream)
        res = B.Result(c_res, ctx)
        launches += res.stats["kernel_launches"]
        for ph in ("l2_rays", "l2", "l2_normals"):  # the dominant kernels, from events inside the timed steps
            phase_acc[ph] = phase_acc.get(ph, 0.0) + res.phase_ms[ph]
    barrier()
    # every other phase: the same number of steps again with an event pair around EVERY kernel (GPV_PROFILE), untimed -- two dozen
    # event records per step are host time that a 1 ms step would feel
    params.c.flags |= gpv.GPV_PROFILE
    for k in range(args.steps):
        barrier()
        r_p = step()
        for ph, v in r_p.phase_ms.items():
            if ph not in ("l2_rays", "l2", "l2_normals"):
                phase_acc[ph] = phase_acc.get(ph, 0.0) + v
    params.c.flags &= ~gpv.GPV_PROFILE
    barrier()
    # the timed region (tens of ms) is shorter than nvidia-smi's sampling period: keep the same work running, untimed, until
    # the sampler has seen the GPU under this load
    extended = 0
    flag = torch.zeros(1, device="cuda")
    while True:
        need = 1.0 if (rank == 0 and sampler.p is not None and sampler.count(t_timed0) < 8 and extended < 400) else 0.0
        flag.fill_(need)
        if world > 1:
            dist.broadcast(flag, 0)
        if flag.item() == 0.0:
            break
        for _ in range(10):
            step()
        extended += 10
    barrier()
    t_timed1 = time.perf_counter()
    clocks = sampler.stop(t_timed0, t_timed1, extended) if rank == 0 else None

    gather_ok = None
    if world > 1:  # untimed: the gathered streams on rank 0 must equal a single-GPU run of the whole grid, byte for byte
        if peer:
            ctx.gather_detach() if rank else None
        if rank == 0:
            w = lambda ptr, n: sharded.wrap_device_bytes(torch, ptr, n)
            if peer:
                p1, p2, p3, nbt = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
                B._check(L.gpv_gather_result(ctx.h, C.byref(p1), C.byref(p2), C.byref(p3), C.byref(nbt)))
                got = {"l1": w(p1.value, plane * nz).clone(), "prefix": w(p2.value, plane * nz * 4).clone(), "l2": w(p3.value, nbt.value * n23).clone()}
                ctx.gather_detach()
            else:
                got = {k: v.clone() for k, v in gathered.items()}
            ref = ctx.voxelize_device(d_tris, mesh, gpv.Params(args.l1, args.l2, 0), sptr)
            want = {"l1": w(ref.c.d_level1_inout, ref.cells), "prefix": w(ref.c.d_prefix, ref.cells * 4), "l2": w(ref.c.d_level2_inout, ref.nb * ref.n23)}
            gather_ok = all(got[k].numel() == want[k].numel() and bool(torch.equal(got[k], want[k])) for k in want)
            assert gather_ok, "gathered slabs differ from the single-GPU result"
        barrier()
    ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device="cuda", dtype=torch.float64)
    tests_local = torch.tensor([res.stats["l2_box_tests"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tests_local)
    ms_per_step = float(ms.item()) / args.steps
    phase_per_rank = None
    if world > 1:  # every rank's own phase times (CUDA events on its stream), so that the scaling line shows where each rank's step goes
        mine = {k: round(v / args.steps, 4) for k, v in phase_acc.items() if v > 0}
        mine["step_ms"] = round(sum(a.elapsed_time(b) for a, b in ev) / args.steps, 4)
        mine["l2_cells"] = int(res.n_refined)
        phase_per_rank = [None] * world
        dist.all_gather_object(phase_per_rank, mine)
    tests = float(tests_local.item()) + res.stats["l1_box_tests"]  # Level-1 tests are replicated: counted once
    value = tests / (ms_per_step * 1e-3) / 1e9

    if os.environ.get("GPV_DEBUG_OWN"):
        # profiling aid (tools/profile_round.sh): this process refined only the Level-2 share of one rank of an N-rank gathering call, on
        # one GPU -- what ncu captures for the N > 1 roofline entries.  Not a benchmark line: print the phase times and stop.
        print(json.dumps({"profiling_share": os.environ["GPV_DEBUG_OWN"], "ms_per_step": ms_per_step, "l2_cells": int(res.n_refined),
                          "phase_ms": {k: round(v / args.steps, 4) for k, v in phase_acc.items() if v > 0}}))
        ctx.free_device(d_tris)
        ctx.close()
        return
    # ---- e2e: pinned host triangles in, host streams out, through gpv_voxelize_host (each rank: its slab, its PCIe link)
    e2e_params = gpv.Params(args.l1, args.l2, 0, z0, z1)
    slab = ctx.voxelize_device(d_tris, mesh, e2e_params, sptr)  # sizes of this rank's slab (the gathering call above reports the whole grid)
    cells, nb, n23 = slab.cells, slab.nb, slab.n23
    hb = {k: L.gpv_alloc_host(n) for k, n in (("l1", cells), ("pre", cells * 4), ("bi", nb * 4 + 64), ("l2", nb * n23 + 64))}
    if args.normals:
        hb.update({k: L.gpv_alloc_host(n) for k, n in (("n1", cells * 3), ("n2", nb * n23 * 3 + 64))})
    pinned_tris = L.gpv_alloc_host(mesh.ntri * 36)
    C.memmove(pinned_tris, C.cast(mesh.c.tris, C.c_void_p), mesh.ntri * 36)
    pm = B.CMesh(mesh.c.n_tri, C.cast(pinned_tris, C.POINTER(C.c_float)), mesh.c.bbox_min, mesh.c.bbox_max, mesh.c.max_model_size, mesh.c.n_verts)
    hs = B.CHostStreams(hb["l1"], hb["pre"], hb["bi"], hb["l2"], hb.get("n1"), hb.get("n2"), nb * n23 + 64, nb + 16)
    e2e_base_flags = gpv.GPV_NORMALS if args.normals else 0
    r2 = B.CResult()
    pre_np = np.ctypeslib.as_array(C.cast(hb["pre"], C.POINTER(C.c_int32)), shape=(cells,))
    nb_all = torch.zeros(world, dtype=torch.int64, device="cuda")

    def e2e_step():
        if L.gpv_voxelize_host(ctx.h, C.byref(pm), C.byref(e2e_params.c), sptr, C.byref(r2), C.byref(hs)):
            raise SystemExit(L.gpv_last_error().decode())
        if world > 1:  # slab-local prefix sums -> global: boundary counts of the lower slabs
            mine = torch.tensor([r2.n_boundary], dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(nb_all, mine)
            base = int(nb_all[:rank].sum().item())
            if base:
                np.add(pre_np, base, out=pre_np)

    def time_e2e(flags):
        e2e_params.c.flags = flags | e2e_base_flags
        for _ in range(args.warmup):
            e2e_step()
        barrier()
        t = 0.0
        for k in range(args.steps):
            flush.fill_(k & 0xff)
            barrier()
            t0 = time.perf_counter()
            e2e_step()  # returns after the last byte has landed in the caller's buffers (stream syncs / host-thread joins inside)
            if world > 1:
                dist.barrier()
            t += time.perf_counter() - t0
        te = torch.tensor([t], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        got = np.ctypeslib.as_array(C.cast(hb["l2"], C.POINTER(C.c_uint8)), shape=(nb * n23,))
        assert int((got == 254).sum()) == int(r2.l2_boundary) == slab.counts[3], "e2e host stream does not match the device counts"
        return 1e3 * float(te.item()) / args.steps

    # Level 2 over PCIe as file bytes (1 B per sub-voxel, straight into the caller's buffer by DMA) and as 2 bits per sub-voxel
    # expanded by the library's host threads (GPV_PACKED_L2): same bytes delivered; the headline is the faster, both are reported
    can_pack = (n23 % 32) == 0
    e2e_bytes_ms = time_e2e(0)
    e2e_packed_ms = time_e2e(gpv.GPV_PACKED_L2) if can_pack else None
    packed_wins = e2e_packed_ms is not None and e2e_packed_ms < e2e_bytes_ms
    e2e_ms = e2e_packed_ms if packed_wins else e2e_bytes_ms
    l2_d2h = nb * n23 // 4 if packed_wins else nb * n23
    dsum = torch.tensor([float(cells + cells * 4 + nb * 4 + l2_d2h + (3 * cells + 3 * nb * n23 if args.normals else 0))], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dsum)
    e2e = {"value": tests / (e2e_ms * 1e-3) / 1e9, "unit": "G tri-box tests/s", "ms_per_model": e2e_ms,
           "h2d_bytes_per_step": mesh.ntri * 36 * world, "d2h_bytes_per_step": int(dsum.item()),
           "level2_transfer": "2 bits per sub-voxel over PCIe, expanded into the caller's bytes by %d host threads per rank (GPV_PACKED_L2)" % int(os.environ.get("GPV_HOST_THREADS", min(15, max(1, (os.cpu_count() or 2) - 1)))) if packed_wins else "file bytes by DMA",
           "ms_per_model_level2_as_bytes": e2e_bytes_ms, "ms_per_model_level2_packed": e2e_packed_ms,
           "timing": "host wall clock around gpv_voxelize_host, max over ranks (pinned buffers both ways; Level-2 D2H overlaps the refinement; "
                     "the call returns after the last byte has landed)"}
    if rank == 0 and world == 1:
        # file write reported separately, outside every timed region -- Object::SaveVoxelization's files from the host
        # streams of the last e2e step (streams that were not computed get no file)
        try:
            import shutil
            d_out = tempfile.mkdtemp(prefix="gpvsave")
            t0 = time.perf_counter()
            rc = L.gpv_save_streams(C.byref(pm), C.byref(r2), C.byref(hs), -1, os.fsencode(d_out), 1)
            dt = time.perf_counter() - t0
            if rc == 0:
                e2e["file_write"] = {"ms": 1e3 * dt, "bytes": sum(os.path.getsize(os.path.join(d_out, f)) for f in os.listdir(d_out)),
                                     "files": sorted(os.listdir(d_out)), "note": "gpv_save_streams into a fresh temporary directory (page cache), not part of e2e"}
            shutil.rmtree(d_out, ignore_errors=True)
        except Exception as e:  # commentary: never lose the bench line over it
            e2e["file_write"] = {"error": repr(e)}
    for p in list(hb.values()) + [pinned_tris]:
        L.gpv_free_host(p)

    if rank == 0:
        k_l2_ms = phase_acc["l2"] / args.steps                 # the SAT kernel alone (CUDA events on the launching stream)
        k_rays_ms = phase_acc.get("l2_rays", 0.0) / args.steps   # k_col_cells + k_l2_rays
        fp32_peak = ctx.fp32_peak()
        hbm_peak, hbm_src = measured_peaks()
        # executed-work numbers of the same kernels under ncu (never taken in this run: a number measured under a profiler is not a
        # bench value): at N = 1 the whole call, at N > 1 rank 0's share captured on one GPU (GPV_DEBUG_OWN=N,0); only quoted while the
        # kernel sources still hash to what the capture was made from
        ncu_all, ncu_src = ncu_traffic()
        ncu = ncu_all.get("kernels", {}) if world == 1 else ncu_all.get("share", {}).get(str(world), {})
        sat_name = "k_l2<%d, 0>" % (args.l2 if args.l2 in (2, 4, 8, 16) else 0)  # (rank 0 of a gathering call writes its own blocks as bytes, too)
        n_sat, n_rays = ncu.get(sat_name, {}), ncu.get("k_l2_rays", {})
        flops = FLOPS_PER_TRIBOX * res.stats["l2_box_tests"]
        ach = flops / (k_l2_ms * 1e-3) / 1e12

        def ncu_note(d):
            if not d:
                return None
            out = {k: d.get(k) for k in ("issue_active_pct", "warp_inst_executed", "thread_inst_executed", "lanes_per_inst", "pipe_fma_pct", "pipe_alu_pct", "time_us",
                                         "fp32_fadd", "fp32_fmul", "fp32_ffma")}
            out["source"] = "profiles/" + d["source"]
            out["provenance"] = ncu_src
            return out

        def executed(d, kernel_ms):
            """executed FP32 work of the launch (hardware counters of the capture: thread-level FADD + FMUL + FFMA, predicated-on lanes)
            over the kernel's LIVE time, against the live FP32 issue peak: the utilisation figure (the algorithmic `frac` is a speed-up figure)"""
            if not d or d.get("fp32_fadd") is None or kernel_ms <= 0:
                return {}
            lane_ops = d["fp32_fadd"] + d["fp32_fmul"] + d["fp32_ffma"]
            return {"executed_fp32_lane_ops": lane_ops, "executed_flops": d["fp32_fadd"] + d["fp32_fmul"] + 2.0 * d["fp32_ffma"],
                    "frac_executed": lane_ops / (kernel_ms * 1e-3) / fp32_peak,
                    "executed_note": "FADD + FMUL + FFMA thread instructions of this launch under ncu (deterministic for the same sources and model) / live kernel time / "
                                     "live non-FMA FP32 issue peak; the rest of the issue slots go to integer, compare, select, shared-memory and queue instructions"}

        roof = {"kernel": sat_name + " (Level-2 SAT over shared-memory queues + final bytes)" + (" on rank 0's share of the boundary cells" if world > 1 else ""), "bound": "fp32",
                "achieved": ach, "peak": fp32_peak / 1e12, "unit": "TFLOP/s", "frac": ach / (fp32_peak / 1e12),
                "traffic": n_sat.get("dram_bytes"),
                "peak_source": "non-FMA FP32 issue rate measured live (gpv_measure_fp32_peak: independent FMUL/FADD chains); "
                               "MEASURED_PEAKS.json has no FP32 entry",
                "algorithmic": "%d reference-equivalent tri-box tests x 124 FLOP per launch (SURVEY.md 8d).  frac > 1 is expected: certified plane / AABB "
                               "culling proves ~94 %% of the reference's tests negative without running them and the z-independent part of the rest is "
                               "hoisted per sub-voxel column -- `frac` is an algorithmic speed-up figure; `frac_executed` and `frac_issue_slots` are the "
                               "utilisation figures" % res.stats["l2_box_tests"],
                "ncu": ncu_note(n_sat), "ncu_provenance": ncu_src,
                "kernel_ms": k_l2_ms, "share_of_step": k_l2_ms / ms_per_step}
        roof.update(executed(n_sat, k_l2_ms))
        if n_sat.get("issue_active_pct") is not None:
            # the executed-work view of the same kernel: share of issue slots in use (ncu capture of this command), the ceiling
            # that actually binds a kernel whose FP32 work is comparisons, selects and non-FMA arithmetic
            roof["frac_issue_slots"] = n_sat["issue_active_pct"] / 100.0
        ray_flops = 51.0 * res.stats.get("l2_ray_tests", 0)
        roof_rays = {"kernel": "k_l2_rays (Level-2 parity rays per sub-voxel column; + k_col_cells, k_l2_rays_overflow)", "bound": "fp32", "kernel_ms": k_rays_ms, "share_of_step": k_rays_ms / ms_per_step,
                     "achieved": ray_flops / (k_rays_ms * 1e-3) / 1e12 if k_rays_ms > 0 and ray_flops else None, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
                     "algorithmic": "%d reference-equivalent ray tests x 51 FLOP (+1 division) per launch" % res.stats.get("l2_ray_tests", 0),
                     "traffic": n_rays.get("dram_bytes"), "ncu": ncu_note(n_rays)}
        if roof_rays["achieved"]:
            roof_rays["frac"] = roof_rays["achieved"] / roof_rays["peak"]
        roof_rays.update(executed(n_rays, k_rays_ms))
        if n_rays.get("issue_active_pct") is not None:
            roof_rays["frac_issue_slots"] = n_rays["issue_active_pct"] / 100.0
        out_bytes = res.n_refined * res.n23  # the blocks this rank's launch wrote
        roof_hbm = {"kernel": sat_name, "bound": "hbm", "achieved": out_bytes / (k_l2_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": out_bytes / (k_l2_ms * 1e-3) / 1e9 / hbm_peak, "traffic": n_sat.get("dram_bytes"),
                    "peak_source": hbm_src, "algorithmic": "1 B per Level-2 voxel written (%d B)" % out_bytes}
        roof_normals = None
        if args.normals and world == 1:
            t1n, t2n = phase_acc.get("l1_normals", 0.0) / args.steps, phase_acc.get("l2_normals", 0.0) / args.steps
            roof_normals = {"kernels": "k_l1_normals (+ 127-fill), k_l2_normals", "bound": "hbm", "unit": "GB/s", "peak": hbm_peak, "peak_source": hbm_src,
                            "algorithmic": "3 B written per Level-1 cell (%d B) and per Level-2 sub-voxel (%d B)" % (3 * res.cells, 3 * res.nb * res.n23),
                            "level1": {"ms": t1n, "achieved": 3 * res.cells / (t1n * 1e-3) / 1e9 if t1n > 0 else None},
                            "level2": {"ms": t2n, "achieved": 3 * res.nb * res.n23 / (t2n * 1e-3) / 1e9 if t2n > 0 else None}}
            for lv in ("level1", "level2"):
                if roof_normals[lv]["achieved"]:
                    roof_normals[lv]["frac"] = roof_normals[lv]["achieved"] / hbm_peak
        # every phase of the pipeline against the roofline that bounds it (algorithmic bytes / flops per unit x units
        # of this model, over the phase's CUDA-event time).  The Level-1 phases of a cessna-sized model are tens of microseconds of
        # launch-latency-bound work: their fractions say how far a 4 M-cell grid is from filling the machine, not kernel quality.
        try:
            ph = {k: v / args.steps for k, v in phase_acc.items()}
            cells_b, nb_b, ntri_b = float(res.cells), float(res.nb), float(mesh.ntri)
            l1_flops = float(FLOPS_PER_TRIBOX * res.stats["l1_box_tests"])  # each binning sweep runs every clipped-footprint test once
            table = [("setup", "hbm", 4 * cells_b + 16 * plane + 212 * ntri_b, "k_clear + k_prepare + work-space scans: counters zeroed, 36 B read + 176 B written per triangle"),
                     ("bin_count", "fp32", l1_flops, "124 FLOP per (triangle, cell) test of the clipped footprints"),
                     ("bin_fill", "fp32", l1_flops, "124 FLOP per (triangle, cell) test of the clipped footprints"),
                     ("scan", "hbm", 8 * cells_b + 8 * nb_b + 24 * plane, "4 B read + 4 B written per cell, 8 B per boundary cell, three column scans"),
                     ("fill_sweep", "hbm", 1.125 * cells_b, "1 B written + 1 bit read per cell"),
                     ("l2_rays", "fp32", 51.0 * res.stats.get("l2_ray_tests", 0), "51 FLOP per reference-equivalent ray test"),
                     ("l2", "fp32", float(flops), "124 FLOP per reference-equivalent tri-box test")]
            per_phase = []
            for name, bound, work, what in table:
                t = ph.get(name, 0.0) * 1e-3
                if t <= 0:
                    continue
                peak = hbm_peak * 1e9 if bound == "hbm" else fp32_peak
                per_phase.append({"phase": name, "ms": round(t * 1e3, 4), "bound": bound, "achieved": work / t / (1e9 if bound == "hbm" else 1e12),
                                  "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": work / t / peak, "algorithmic": what})
        except Exception as e:  # the table is commentary: never lose the bench line over it
            per_phase = [{"error": repr(e)}]
        line = {"metric": "G tri-box tests/s", "value": value, "unit": "G tri-box tests/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "ms_per_model": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "cessna.obj fixture (reference sample mesh)" if args.mesh == "cessna" else "synthetic",
                "config": workload_config(args, int(tests), mesh.ntri, [int(x) for x in res.num_div]),   # the workload only: identical in the reference arm's line
                "parallelism": ("%d ranks, Level-1 replicated, %s; gathered streams == single-GPU result: %s" % (world, "Level-1 bytes / prefix sums by equal z-slabs, Level-2 refinement by interleaved Level-1 column groups, every rank writing its share into rank 0's buffers over NVLink peer memory from inside the kernels (GPV_GATHER): no collective, no exchange step, completion flags through a mailbox" if peer else "z-slabs (cuts %s), NCCL send/recv gather to rank 0" % cuts, gather_ok)) if world > 1 else "1 GPU",
                "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roof, "roofline_rays": roof_rays, "roofline_hbm": roof_hbm, "roofline_phases": per_phase,
                "phase_ms": {k: round(v / args.steps, 4) for k, v in phase_acc.items() if v > 0},
                "counts": {"l1_inside": whole.counts[0], "l1_boundary": whole.counts[1]}}
        if phase_per_rank is not None:
            line["phase_ms_per_rank"] = phase_per_rank
        if roof_normals is not None:
            line["roofline_normals"] = roof_normals
        if args.normals:
            line["config"]["workload"] += " + Level1Normal / Level2Normal streams"
        if world == 1 and not args.no_cpu_baseline and path is not None:
            line["cpu_baseline"] = cpu_baseline(path, args.l1, args.l2, os.cpu_count() or 1)
        print(json.dumps(line))
    ctx.free_device(d_tris)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
