#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference (oracle/_ref/libgpvref.so, built by oracle/Makefile).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference); the GPU box only sees the committed
fixtures.  Usage:  python oracle/gen_golden.py [case ...]      (no args = every case)

Per case the fixture holds
  <case>.json : grid, counts, reference-equivalent test counts, FNV-1a-64 + sha256 of every output stream
  <case>.npz  : (small cases) the streams themselves, zlib-compressed: l1 state, prefix sum, boundary index,
                L2 state (kernel form), normals as the reference's uchar encoding, CSR cell lists / column lists
What "reference output" means for each stream is spelled out in oracle/ref_harness.cpp's header.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import refbind as R  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("GPVIEW_REF", "/root/reference")

# name -> (mesh path, L1, L2, brute-force fill mode, store arrays?)
CASES = {
    "cessna_8_4": (os.path.join(REF, "files/cessna.obj"), 8, 4, "member", True),
    "cessna_64_4": (os.path.join(REF, "files/cessna.obj"), 64, 4, "member", True),
    "cessna_128_8": (os.path.join(REF, "files/cessna.obj"), 128, 8, "mt", False),
    "cessna_256_16": (os.path.join(REF, "files/cessna.obj"), 256, 16, "mt", False),
}


def mesh_cases():
    """Synthetic meshes written by gpview_b200.meshgen (committed generator, seeded) and voxelized by the reference."""
    d = os.path.join(GOLD, "meshes")
    out = {}
    if os.path.isdir(d):
        for fn in sorted(os.listdir(d)):
            if fn.endswith((".off", ".obj")):
                stem = fn.rsplit(".", 1)[0]
                out["%s_32_4" % stem] = (os.path.join(d, fn), 32, 4, "member", True)
    return out


def h(b):
    b = np.ascontiguousarray(b)
    return {"fnv1a64": R.fnv1a64(b.tobytes()) if b.nbytes < (64 << 20) else None,
            "sha256": hashlib.sha256(b.tobytes()).hexdigest(), "bytes": int(b.nbytes)}


def to_u8_state(f):
    return (f * np.float32(127.0)).astype(np.uint8)


def to_u8_normal(n):
    return (n * np.float32(256.0 / 3.0) + np.float32(127.0)).astype(np.uint8)


def run_case(name, path, l1, l2, mode, store):
    t0 = time.time()
    o = R.RefObject(path, obj_id=-1)
    o.setup(l1, l2)
    info = {"case": name, "mesh": os.path.basename(path), "l1": l1, "l2": l2, "ntri": o.ntri,
            "bbox_min": [float(x) for x in o.bmin], "bbox_max": [float(x) for x in o.bmax],
            "bbox_min_hex": [np.float32(x).tobytes().hex() for x in o.bmin],
            "bbox_max_hex": [np.float32(x).tobytes().hex() for x in o.bmax],
            "max_model_size": float(o.max_model_size),
            "num_div": [int(x) for x in o.num_div], "grid_size": [float(x) for x in o.grid_size],
            "grid_size_hex": [np.float32(x).tobytes().hex() for x in o.grid_size],
            "grid_size2": [float(x) for x in o.grid_size2],
            "fill": "Object::ClassifyInOutCPU" if mode == "member" else "ClassifyInOutCPU loop nest, 8 host threads"}
    info["t_l1_inout_s"] = o.l1_inout_brute(0 if mode == "member" else 8)
    fill_only = o.level1_inout().copy()
    info["t_l1_tribox_s"] = o.l1_tribox()
    o.compact()
    info["collist_fill_mismatches"] = int(o.collist_mismatch())
    info["t_l2_kernelform_s"] = o.l2_kernelform(8)
    k2 = o.level2_inout_kernel()
    kn = o.level2_normal_kernel().reshape(-1, 4)[:, :3]
    if store:  # cross-check against the reference's CPU Level-2 twins (f64 centres) on the small cases
        a, b = o.l2_cpu()
        info["t_l2_cpu_inout_s"], info["t_l2_cpu_tribox_s"] = a, b
        info["l2_cpu_twin_vs_kernelform_diffs"] = int((o.level2_inout() != k2).sum())
    o.adopt_kernelform()
    cnt = o.count()
    info["l1_inside"], info["l1_boundary"], info["l2_inside"], info["l2_boundary"] = cnt
    info.update(o.stats())
    l1s = to_u8_state(o.level1_inout())
    l1n = to_u8_normal(o.level1_normal())
    pre = o.prefix()
    bidx = o.boundary_index()
    l2s = to_u8_state(k2)
    l2n = to_u8_normal(kn.reshape(-1))
    info["streams"] = {"Level1InOut": h(l1s), "Level1FillOnly": h(to_u8_state(fill_only)), "Level1Normal": h(l1n),
                       "Level1BoundaryPrefixSum": h(pre), "BoundaryIndex": h(bidx), "Level2InOut": h(l2s), "Level2Normal": h(l2n)}
    tc, xc = o.tri_count(), o.xy_count()
    info["tri_flat_len"], info["xy_flat_len"] = int(tc.sum()), int(xc.sum())
    info["max_xy_count"] = int(xc.max())
    n23 = max(l2, 1) ** 3
    per_in = (l2s.reshape(-1, n23) == 127).sum(1).astype(np.int32)
    per_bd = (l2s.reshape(-1, n23) == 254).sum(1).astype(np.int32)
    os.makedirs(GOLD, exist_ok=True)
    arrays = {"l1_state_bits": np.packbits(np.stack([(l1s == 127), (l1s == 254)]).astype(np.uint8)),
              "boundary_index": bidx, "l2_inside_per_cell": per_in, "l2_boundary_per_cell": per_bd, "tri_count_boundary": tc[bidx]}
    if store:
        # canonical (ascending) CSR in linear-index order + column lists
        tfi, tf = o.tri_flat_index(), o.tri_flat()
        cells = [np.sort(tf[tfi[c]:tfi[c] + tc[c]]) for c in bidx]
        xfi, xf = o.xy_flat_index(), o.xy_flat()
        cols = [np.sort(xf[xfi[c]:xfi[c] + xc[c]]) for c in range(len(xc))]
        arrays.update({"l2_state": l2s, "l1_normal_u8": l1n, "l2_normal_u8": l2n, "xy_count": xc,
                       "cell_lists": np.concatenate(cells) if cells else np.zeros(0, np.int32),
                       "col_lists": np.concatenate(cols) if cols else np.zeros(0, np.int32)})
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **arrays)
    info["t_total_s"] = time.time() - t0
    with open(os.path.join(GOLD, name + ".json"), "w") as f:
        json.dump(info, f, indent=1)
    o.close()
    print(name, "done in %.1fs" % info["t_total_s"], "L1 in/bd", cnt[:2], "L2 in/bd", cnt[2:], flush=True)


def main():
    cases = dict(CASES)
    cases.update(mesh_cases())
    want = sys.argv[1:] or list(cases)
    for name in want:
        run_case(name, *cases[name])


if __name__ == "__main__":
    main()
