#!/usr/bin/env python3
"""Generate tests/golden/fullsize.json: stream hashes of BASELINE.json's full-size configurations (3, 4, 5) from the CPU oracle.

TEST INFRASTRUCTURE ONLY.  The small fixtures (oracle/gen_golden.py) come from the unmodified reference and pin the oracle;
the reference's own loops cannot finish these sizes (brute-force fill: 134 M cells x 1 M triangles), so here the pinned
oracle (oracle/gpv_oracle.c, certified column-list fill == brute force, tests/test_oracle_*.py) is the source.  The GPU box
has no reference: tests/test_gpu_scale.py compares the device streams' sha256 with this file.  The meshes are made by numpy
(sin / cos in float64, rounded to float32), whose last bit may depend on the host CPU's SIMD path: every case records the
sha256 of its triangle array, and when the array built on the box differs the test calls expected() there instead (the
same oracle, live) -- so the comparison never silently weakens.

  python oracle/gen_fullsize.py [case ...]        (no args = every case; ~20 s for the 1 M-triangle cases, minutes and
                                                   ~20 GB of RAM for the 10 M-triangle one, 8 host threads)
Meshes are the seeded generators of gpview_b200.meshgen handed over as triangle arrays (bbox over the vertices + the
reference's padding); config 5's drilled blocks go through an ASCII .off file and the loader, as in a dataset run."""
import hashlib
import json
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from gpview_b200 import meshgen as M  # noqa: E402
from oracle import oraclebind as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "fullsize.json")
BLOCK_SEEDS = 4


def fullsize_tris(name):
    """The full-size synthetic meshes of SURVEY.md 8d (the same functions tests/test_gpu_scale.py calls)."""
    if name == "sphere":
        V, F = M.uv_sphere(1000, 502)          # 1,000,000 triangles
    elif name == "torus":
        V, F = M.torus(1000, 500)              # 1,000,000 triangles
    elif name == "cad":
        V, F = M.cad_body(2500, 2001)          # 10,000,000 triangles
    else:
        raise KeyError(name)
    return M.triangles(V, F)


CASES = {"sphere_512_8": ("sphere", 512, 8), "torus_512_8": ("torus", 512, 8), "cad_1024_2": ("cad", 1024, 2)}
CASES.update({"block%d_64_4" % i: ("block%d" % i, 64, 4) for i in range(BLOCK_SEEDS)})


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def case_mesh(case, workdir):
    """(triangle array or None, .off path or None) of a case.  Blocks go through a file: parsing is part of config 5."""
    name = CASES[case][0]
    if name.startswith("block"):
        V, F = M.drilled_block(seed=M.SEED_BASE + int(name[5:]), n_seg=160, n_grid=36)
        p = os.path.join(workdir, name + ".off")
        M.write_off(p, V, F)
        return None, p
    return fullsize_tris(name), None


def expected(case, tris=None, path=None):
    """What the oracle says about a case (the dictionary stored per case in fullsize.json)."""
    name, l1, l2 = CASES[case]
    om = O.OracleMesh(path) if path else O.OracleMesh(tris=tris)
    flags = O.FILL_CERTIFIED | (0 if path else O.NO_NORMALS)
    r = om.voxelize(l1, l2, flags, os.cpu_count() or 8)
    out = {"mesh": name, "l1": l1, "l2": l2, "triangles": int(om.ntri), "triangles_sha256": sha(om.tris),
           "num_div": [int(x) for x in r.num_div], "counts": [int(x) for x in r.counts], "n_boundary": int(r.nb),
           "sha256": {"level1_inout": sha((r.l1_state * 127).astype(np.uint8)), "prefix": sha(r.prefix),
                      "boundary_index": sha(r.boundary_index), "level2_inout": sha((r.l2_state * 127).astype(np.uint8))}}
    if not (flags & O.NO_NORMALS):
        out["sha256"]["level1_normal"] = sha(r.l1_normal)
        out["sha256"]["level2_normal"] = sha(r.l2_normal)
    return out


def run(case):
    t0 = time.time()
    with tempfile.TemporaryDirectory() as d:
        tris, path = case_mesh(case, d)
        out = expected(case, tris, path)
    print("%-14s %8d triangles  grid %s  %d boundary cells  %.1f s" % (case, out["triangles"], out["num_div"], out["n_boundary"], time.time() - t0), flush=True)
    return out


def main():
    want = sys.argv[1:] or list(CASES)
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    for case in want:
        data[case] = run(case)
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
