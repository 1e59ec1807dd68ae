"""The product's intersection arithmetic (gpview_b200/csrc/gpv_math.h, compiled for the host by tests/cpu_probe) must
agree bit-for-bit with the oracle -- and, in the build container, with the reference's own object code -- on random and
adversarial inputs: touching, degenerate, axis-aligned, coplanar, sliver and far-away triangles."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from util import HAVE_REF, ROOT

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def probe():
    so = os.path.join(HERE, "cpu_probe", "math_probe.so")
    src = os.path.join(HERE, "cpu_probe", "math_probe.cpp")
    hdr = os.path.join(ROOT, "gpview_b200", "csrc", "gpv_math.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    fp, bp = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    for n in ("probe_sat_full", "probe_sat_row"):
        getattr(L, n).argtypes = [C.c_int64, fp, fp, fp, bp]
    L.probe_ray.argtypes = [C.c_int64, fp, fp, bp]
    L.probe_column.argtypes = [C.c_int64, fp, fp, bp]
    L.probe_candidates.argtypes = [C.c_int64, fp, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.probe_cell_of.argtypes = [C.c_float, C.c_float, C.c_float, C.c_int]
    L.probe_encode_normal.argtypes = [C.c_float]
    return L


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _bp(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def sat_cases(rng, n):
    """boxes + triangles with many near-boundary situations"""
    c = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    h = rng.uniform(0.05, 0.6, (n, 3)).astype(np.float32)
    t = (c[:, None, :] + rng.normal(0, 0.7, (n, 3, 3)) * rng.uniform(0.1, 2.0, (n, 1, 1))).astype(np.float32)
    k = n // 8
    # vertices exactly on box faces / corners
    t[:k, 0, :] = c[:k] + h[:k] * rng.choice([-1, 1], (k, 3))
    # axis-aligned triangles lying in a face plane of the box
    t[k:2 * k, :, 0] = (c[k:2 * k, 0] + h[k:2 * k, 0])[:, None]
    # degenerate: two equal vertices, or all collinear
    t[2 * k:3 * k, 1] = t[2 * k:3 * k, 0]
    t[3 * k:4 * k, 2] = (t[3 * k:4 * k, 0] + (t[3 * k:4 * k, 1] - t[3 * k:4 * k, 0]) * np.float32(2.0)).astype(np.float32)
    # big triangles swallowing the box, tiny triangles inside the box
    t[4 * k:5 * k] = (c[4 * k:5 * k, None, :] + rng.normal(0, 20, (k, 3, 3))).astype(np.float32)
    t[5 * k:6 * k] = (c[5 * k:6 * k, None, :] + rng.normal(0, 0.01, (k, 3, 3))).astype(np.float32)
    return c, h, t.reshape(n, 9)


def test_sat_matches_oracle(probe, oracle):
    rng = np.random.default_rng(1)
    n = 400000
    c, h, t = sat_cases(rng, n)
    want = oracle.tribox_batch(c, h, t)
    full = np.zeros(n, np.uint8); row = np.zeros(n, np.uint8)
    probe.probe_sat_full(n, _fp(c), _fp(h), _fp(t), _bp(full))
    probe.probe_sat_row(n, _fp(c), _fp(h), _fp(t), _bp(row))
    assert 0.05 < want.mean() < 0.95
    assert np.array_equal(full, want)
    assert np.array_equal(row, want)


def ray_cases(rng, n):
    o = rng.uniform(-2, 2, (n, 3)).astype(np.float32)
    t = rng.uniform(-2.5, 2.5, (n, 3, 3)).astype(np.float32)
    k = n // 8
    # origin exactly under a vertex / an edge midpoint (inclusive-edge semantics, App. A.5)
    o[:k, :2] = t[:k, 0, :2]
    o[k:2 * k, :2] = ((t[k:2 * k, 0, :2] + t[k:2 * k, 1, :2]) * np.float32(0.5)).astype(np.float32)
    # vertical and near-vertical triangles (det ~ 0), huge coordinates (noise-level determinants)
    t[2 * k:3 * k, 2, :2] = t[2 * k:3 * k, 0, :2]
    t[3 * k:4 * k] = (t[3 * k:4 * k] * np.float32(1000)).astype(np.float32)
    o[3 * k:4 * k] = (o[3 * k:4 * k] * np.float32(1000)).astype(np.float32)
    # slivers: third vertex almost on the line through the first two
    a = rng.uniform(0, 1, (k, 1)).astype(np.float32)
    t[4 * k:5 * k, 2] = (t[4 * k:5 * k, 0] + (t[4 * k:5 * k, 1] - t[4 * k:5 * k, 0]) * a + rng.normal(0, 1e-6, (k, 3))).astype(np.float32)
    # origin on the triangle's plane (t ~ 0)
    o[5 * k:6 * k] = ((t[5 * k:6 * k, 0] + t[5 * k:6 * k, 1] + t[5 * k:6 * k, 2]) / np.float32(3)).astype(np.float32)
    return o, t.reshape(n, 9)


def test_ray_split_matches_general_form(probe, oracle):
    rng = np.random.default_rng(2)
    n = 400000
    o, t = ray_cases(rng, n)
    want = oracle.triray_batch(o, t)
    got = np.zeros(n, np.uint8)
    probe.probe_ray(n, _fp(o), _fp(t), _bp(got))
    assert 0.01 < want.mean() < 0.6
    assert np.array_equal(got, want)
    assert np.array_equal(oracle.triray_batch(o, t, specialised=True), want)


@pytest.mark.skipif(not HAVE_REF, reason="reference object code only exists in the build container")
def test_oracle_predicates_match_reference_object_code(oracle):
    from oracle import refbind
    rng = np.random.default_rng(3)
    n = 200000
    c, h, t = sat_cases(rng, n)
    assert np.array_equal(oracle.tribox_batch(c, h, t), refbind.tribox_batch(c, h, t))
    o, t2 = ray_cases(rng, n)
    assert np.array_equal(oracle.triray_batch(o, t2), refbind.triray_batch(o, t2))


def test_certified_candidates_cover_brute_force(probe):
    """For every triangle, every grid column whose +Z ray passes the column part of Moller-Trumbore must lie inside the
    candidate rectangle (or the triangle is flagged ill-conditioned = test every column)."""
    rng = np.random.default_rng(4)
    nx = ny = 48
    gs = np.float32(0.125)
    mn = np.float32(-3.0)
    cx = (mn + (np.arange(nx) + 0.5) * gs).astype(np.float32)
    n = 3000
    t = rng.uniform(-3, 3, (n, 3, 3)).astype(np.float32)
    k = n // 6
    a = rng.uniform(-1, 2, (k, 1)).astype(np.float32)
    t[:k, 2] = (t[:k, 0] + (t[:k, 1] - t[:k, 0]) * a + rng.normal(0, 3e-6, (k, 3))).astype(np.float32)          # slivers
    t[k:2 * k] = (t[k:2 * k] * np.float32(300)).astype(np.float32)                                              # large coordinates
    t[2 * k:3 * k, 2, :2] = (t[2 * k:3 * k, 0, :2] + rng.normal(0, 1e-5, (k, 2))).astype(np.float32)            # near-vertical
    t[3 * k:4 * k] = (t[3 * k:4 * k, :1] + rng.normal(0, 0.05, (k, 3, 3))).astype(np.float32)                   # small
    t = t.reshape(n, 9)
    cand = np.zeros((n, 5), np.int32)
    probe.probe_candidates(n, _fp(t), mn, mn, gs, gs, nx, ny, cand.ctypes.data_as(C.POINTER(C.c_int32)))
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    oxy = np.stack([cx[ii.reshape(-1)], cx[jj.reshape(-1)]], -1).astype(np.float32)
    kinds = np.bincount(cand[:, 0], minlength=3)
    assert kinds[1] > n // 2 and kinds[2] > 0
    for q in range(n):
        tt = np.repeat(t[q:q + 1], nx * ny, 0)
        hit = np.zeros(nx * ny, np.uint8)
        probe.probe_column(nx * ny, _fp(oxy), _fp(np.ascontiguousarray(tt)), _bp(hit))
        hit = hit.reshape(ny, nx).astype(bool)
        kind, i0, i1, j0, j1 = cand[q]
        if kind == 2:
            continue
        if kind == 0:
            assert not hit.any(), q
            continue
        outside = np.ones((ny, nx), bool)
        outside[j0:j1 + 1, i0:i1 + 1] = False
        assert not (hit & outside).any(), (q, cand[q], np.argwhere(hit & outside)[:4])


def test_cell_of_and_normal_encoding(probe):
    assert probe.probe_cell_of(C.c_float(1.0), C.c_float(0.0), C.c_float(1.0), 8) == 7      # max-edge fix-up (cu:336)
    assert probe.probe_cell_of(C.c_float(0.0), C.c_float(0.0), C.c_float(1.0), 8) == 0
    assert probe.probe_cell_of(C.c_float(0.5), C.c_float(0.0), C.c_float(1.0), 8) == 4
    for x in np.linspace(-1, 1, 101).astype(np.float32):
        want = int(np.uint8(np.float32(np.float32(x * np.float32(256.0 / 3.0)) + np.float32(127.0))))
        assert probe.probe_encode_normal(C.c_float(float(x))) == want


def test_certified_plane_culling_never_drops_a_passing_subvoxel(probe, oracle):
    """The Level-2 kernel skips sub-voxels outside gpv::plane_row_interval.  Sweep cells x triangles over nine orders of
    magnitude of coordinates / cell sizes / triangle sizes, slivers and axis-parallel planes included: no sub-voxel whose
    plane predicate passes in the reference's f32 arithmetic may fall outside the interval."""
    probe.probe_plane_cull.argtypes = [C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    rng = np.random.default_rng(7)
    total = np.zeros(4, np.int64)
    for n2 in (2, 4, 8, 16):
        n = 60000 if n2 < 16 else 25000
        gs = (10.0 ** rng.uniform(-3, 1, (n, 1)) * rng.uniform(0.7, 1.3, (n, 3))).astype(np.float32)
        off = 10.0 ** rng.uniform(-2, 4, (n, 1)) * rng.choice([-1, 1], (n, 3)) * rng.uniform(0, 1, (n, 3))
        mid = (off * (rng.uniform(0, 1, (n, 1)) < 0.7)).astype(np.float32)
        size = 10.0 ** rng.uniform(-2, 2, (n, 1, 1)) * gs[:, None, :]
        ctr = mid[:, None, :] + rng.uniform(-0.7, 0.7, (n, 1, 3)) * gs[:, None, :]
        tri = (ctr + rng.normal(0, 1, (n, 3, 3)) * size).astype(np.float32)
        k = n // 6
        tri[:k, :, 0] = tri[:k, :1, 0] + (rng.normal(0, 1e-4, (k, 3)) * gs[:k, :1]).astype(np.float32)         # plane ~ perpendicular to x
        tri[k:2 * k, :, 2] = tri[k:2 * k, :1, 2]                                                              # Nx = Ny = 0 exactly
        a = rng.uniform(0, 1, (k, 1))
        tri[2 * k:3 * k, 2] = (tri[2 * k:3 * k, 0] + (tri[2 * k:3 * k, 1] - tri[2 * k:3 * k, 0]) * a).astype(np.float32)  # slivers
        tri[3 * k:4 * k, :, 1] = tri[3 * k:4 * k, :1, 1] + (rng.normal(0, 1e-6, (k, 3)) * gs[3 * k:4 * k, 1:2]).astype(np.float32)  # Nx ~ 0
        tri = np.ascontiguousarray(tri.reshape(n, 9))
        # keep only triangles that overlap their Level-1 cell (the kernel only ever sees those)
        keep = oracle.tribox_batch(mid, (gs / np.float32(2.0)).astype(np.float32), tri).astype(bool)
        mid_k, gs_k, tri_k = np.ascontiguousarray(mid[keep]), np.ascontiguousarray(gs[keep]), np.ascontiguousarray(tri[keep])
        out = np.zeros(4, np.int64)
        probe.probe_plane_cull(len(mid_k), _fp(mid_k), _fp(gs_k), n2, _fp(tri_k), out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert out[0] == 0, (n2, out)
        print('n2=%d: kept %d, plane-pass %d, SAT-pass %d of %d sub-voxels' % (n2, out[1], out[2], out[3], len(mid_k) * n2 ** 3))
        total += out
    assert total[2] > 1000000 and total[3] > 100000        # the sweep exercised plenty of passing sub-voxels
    assert total[1] < 4 * total[2]                         # and the interval is not vacuous: it keeps < 4x what passes
    print("plane culling: kept %d sub-voxels, %d pass the plane predicate, %d the full SAT" % (total[1], total[2], total[3]))


def test_certified_z_runs_agree_with_per_cell_evaluation(probe):
    """gpv::ray_z_run decides a whole run of cells along a column from one evaluation; whenever it does, the per-cell
    ray_cell results must all agree (no hit at all / every cell hit).  Origins are placed inside the triangle's projection
    so that the column test passes; near-vertical, huge-coordinate and plane-grazing cases included."""
    probe.probe_z_run.argtypes = [C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    rng = np.random.default_rng(11)
    tot = np.zeros(3, np.int64)
    for nz in (2, 4, 16, 32):
        n = 400000
        scale = 10.0 ** rng.uniform(-2, 3, (n, 1, 1))
        tri = (rng.normal(0, 1, (n, 3, 3)) * scale + rng.normal(0, 1, (n, 1, 3)) * scale * rng.choice([0, 1, 30], (n, 1, 1))).astype(np.float32)
        k = n // 5
        tri[:k, 2, :2] = (tri[:k, 0, :2] + (tri[:k, 1, :2] - tri[:k, 0, :2]) * rng.uniform(0, 1, (k, 1)) +
                          rng.normal(0, 1e-4, (k, 2)) * scale[:k, 0]).astype(np.float32)          # near-vertical: tiny 2-D determinant
        w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
        o = (tri * w[:, :, None]).sum(1)                                                           # a point of the triangle
        oxy = np.ascontiguousarray(o[:, :2], np.float32)
        dz = (10.0 ** rng.uniform(-3, 0, n) * scale[:, 0, 0]).astype(np.float32)
        z0 = (o[:, 2] - dz * rng.uniform(-1.5 * nz, 2.5 * nz, n) * (rng.uniform(0, 1, n) < 0.8)).astype(np.float32)  # runs below / across / above / grazing
        zrun = np.ascontiguousarray(np.stack([z0, dz], -1), np.float32)
        out = np.zeros(3, np.int64)
        probe.probe_z_run(n, _fp(oxy), _fp(zrun), nz, _fp(np.ascontiguousarray(tri.reshape(n, 9))), out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert out[0] == 0, (nz, out)
        tot += out
    assert tot[2] > 400000 and tot[1] > 0.3 * tot[2], tot
    print("z-runs: %d column-passing pairs, %d decided from one evaluation" % (tot[2], tot[1]))


def test_cell_mask_equals_per_sub_voxel_evaluation(probe):
    """gpv::ray_cell_mask (Level-2 parity bits of one cell from a certified prefix + a certified stop) must equal n2
    independent ray_cell evaluations bit for bit: crossings inside, below, above and grazing the cell; near-vertical and
    huge-coordinate triangles; cells from 1/1000 of the triangle to 30x its size."""
    probe.probe_cell_mask.argtypes = [C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    rng = np.random.default_rng(23)
    tot = np.zeros(3, np.int64)
    for n2 in (1, 2, 4, 5, 8, 16, 32):
        n = 300000
        scale = 10.0 ** rng.uniform(-2, 3, (n, 1, 1))
        tri = (rng.normal(0, 1, (n, 3, 3)) * scale + rng.normal(0, 1, (n, 1, 3)) * scale * rng.choice([0, 1, 30, 3000], (n, 1, 1))).astype(np.float32)
        k = n // 5
        tri[:k, 2, :2] = (tri[:k, 0, :2] + (tri[:k, 1, :2] - tri[:k, 0, :2]) * rng.uniform(0, 1, (k, 1)) +
                          rng.normal(0, 1e-4, (k, 2)) * scale[:k, 0]).astype(np.float32)          # near-vertical: tiny 2-D determinant
        tri[k:2 * k, :, 2] = tri[k:2 * k, :1, 2] + rng.normal(0, 1e-5, (k, 3)).astype(np.float32) * scale[k:2 * k, 0]  # nearly horizontal
        w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
        o = (tri * w[:, :, None]).sum(1)
        oxy = np.ascontiguousarray(o[:, :2], np.float32)
        gs = (10.0 ** rng.uniform(-3, 1.5, n) * scale[:, 0, 0]).astype(np.float32)
        # cell centre: crossing inside the cell (60 %), exactly on a sub-voxel centre (10 %), cell far below / above (30 %)
        u = rng.uniform(0, 1, n)
        off = np.where(u < 0.6, rng.uniform(-0.6, 0.6, n), np.where(u < 0.7, (rng.integers(0, n2, n) + 0.5) / n2 - 0.5, rng.uniform(-40, 40, n)))
        mid = (o[:, 2] + gs * off).astype(np.float32)
        cellz = np.ascontiguousarray(np.stack([mid, gs], -1), np.float32)
        out = np.zeros(3, np.int64)
        probe.probe_cell_mask(n, _fp(oxy), _fp(cellz), n2, _fp(np.ascontiguousarray(tri.reshape(n, 9))), out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert out[0] == 0, (n2, out)
        tot += out
    assert tot[1] > 400000 and tot[2] > 100000, tot   # plenty of passing pairs, plenty of cells with the crossing inside
    print("cell masks: %d column-passing pairs, %d with the crossing inside the cell" % (tot[1], tot[2]))


def test_column_thresholds_classify_cells_like_per_sub_voxel_evaluation(probe):
    """gpv::ray_col_thresholds gives, per (triangle, sub-voxel column), one height below which every cell of the grid column is
    all hit and one from which on no cell is hit (two evaluations of t, certified slope and rounding bounds).  Every cell the
    thresholds classify must agree with n2 independent ray_cell evaluations; columns of 4 to 1000 cells, cells from 1/300 of
    the triangle to 30x its size, columns far from the origin, near-vertical / near-horizontal triangles."""
    probe.probe_col_thresholds.argtypes = [C.c_int64, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]
    rng = np.random.default_rng(29)
    tot = np.zeros(4, np.int64)
    for n2 in (1, 2, 4, 8, 16, 32):
        n = 40000
        scale = 10.0 ** rng.uniform(-2, 3, (n, 1, 1))
        tri = (rng.normal(0, 1, (n, 3, 3)) * scale + rng.normal(0, 1, (n, 1, 3)) * scale * rng.choice([0, 1, 30, 3000], (n, 1, 1))).astype(np.float32)
        k = n // 5
        tri[:k, 2, :2] = (tri[:k, 0, :2] + (tri[:k, 1, :2] - tri[:k, 0, :2]) * rng.uniform(0, 1, (k, 1)) +
                          rng.normal(0, 1e-4, (k, 2)) * scale[:k, 0]).astype(np.float32)          # near-vertical
        tri[k:2 * k, :, 2] = tri[k:2 * k, :1, 2] + rng.normal(0, 1e-5, (k, 3)).astype(np.float32) * scale[k:2 * k, 0]  # nearly horizontal
        w = rng.dirichlet([1, 1, 1], n).astype(np.float32)
        o = (tri * w[:, :, None]).sum(1)
        oxy = np.ascontiguousarray(o[:, :2], np.float32)
        gs = (10.0 ** rng.uniform(-2.5, 1.5, n) * scale[:, 0, 0]).astype(np.float32)
        ncell = rng.choice([4, 16, 64, 256, 1000], n)
        # the crossing somewhere inside the column (80 %), or the whole column below / above it
        frac = np.where(rng.uniform(0, 1, n) < 0.8, rng.uniform(0, 1, n), rng.choice([-0.5, 1.5], n))
        zmin = (o[:, 2] - gs * ncell * frac).astype(np.float32)
        colz = np.ascontiguousarray(np.stack([zmin, gs, ncell.astype(np.float32)], -1), np.float32)
        out = np.zeros(4, np.int64)
        probe.probe_col_thresholds(n, _fp(oxy), _fp(colz), n2, _fp(np.ascontiguousarray(tri.reshape(n, 9))), out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert out[0] == 0, (n2, out)
        tot += out
    assert tot[1] > 50000 and tot[2] > 0.8 * tot[3], tot   # the thresholds decide the bulk of the cells (88 % in this hostile mix)
    print("column thresholds: %d pairs, %d of %d cells classified without a per-cell evaluation" % (tot[1], tot[2], tot[3]))


@pytest.mark.parametrize("origin,gs", [(0.0, 0.125), (1000.0, 0.125), (-250000.0, 1.0), (3.0, 1e-3), (0.0, 40.0)])
def test_certified_candidates_tight_and_covering_on_shifted_grids(probe, origin, gs):
    """Same property as above on grids far from the origin / with tiny or huge cells, with triangles from 1/20 of a cell
    to 30 cells across; also checks that the rectangle is tight (not more than ~2 columns of slack per side on average)."""
    rng = np.random.default_rng(int(abs(origin)) + int(gs * 1000))
    nx = ny = 40
    gs = np.float32(gs)
    mn = np.float32(origin)
    cx = (np.float64(mn) + (np.arange(nx) + 0.5) * np.float64(gs)).astype(np.float32)
    n = 2500
    size = (10.0 ** rng.uniform(-1.3, 1.5, (n, 1, 1))) * float(gs)
    ctr = float(mn) + rng.uniform(0, nx, (n, 1, 3)) * float(gs)
    t = (ctr + rng.normal(0, 1, (n, 3, 3)) * size).astype(np.float32).reshape(n, 9)
    cand = np.zeros((n, 5), np.int32)
    probe.probe_candidates(n, _fp(t), mn, mn, gs, gs, nx, ny, cand.ctypes.data_as(C.POINTER(C.c_int32)))
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    oxy = np.ascontiguousarray(np.stack([cx[ii.reshape(-1)], cx[jj.reshape(-1)]], -1), np.float32)
    area_c = area_h = 0
    for q in range(n):
        tt = np.ascontiguousarray(np.repeat(t[q:q + 1], nx * ny, 0))
        hit = np.zeros(nx * ny, np.uint8)
        probe.probe_column(nx * ny, _fp(oxy), _fp(tt), _bp(hit))
        hit = hit.reshape(ny, nx).astype(bool)
        kind, i0, i1, j0, j1 = cand[q]
        if kind == 2:
            continue
        if kind == 0:
            assert not hit.any(), q
            continue
        outside = np.ones((ny, nx), bool)
        outside[j0:j1 + 1, i0:i1 + 1] = False
        assert not (hit & outside).any(), (q, cand[q], np.argwhere(hit & outside)[:4])
        if hit.any():
            ys, xs = np.nonzero(hit)
            area_h += (xs.max() - xs.min() + 1) * (ys.max() - ys.min() + 1)
            area_c += (i1 - i0 + 1) * (j1 - j0 + 1)
    assert area_c <= 2.5 * area_h + 4 * n, (area_c, area_h)
