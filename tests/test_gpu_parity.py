"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs and against the committed
fixtures the unmodified reference produced (tests/golden, oracle/gen_golden.py).  Bit-exact for every stream."""
import numpy as np
import pytest

from util import GOLDEN_CASES, golden, l1_bits, mesh_path, sha

pytestmark = pytest.mark.gpu


def run_product(product, ctx, path, l1, l2, flags=None):
    mesh = product.load_mesh(path)
    if flags is None:
        flags = product.GPV_NORMALS | product.GPV_KEEP_LISTS
    res = ctx.voxelize(mesh, product.Params(l1, l2, flags))
    return mesh, res


@pytest.mark.parametrize("name,l1,l2", GOLDEN_CASES)
def test_streams_match_oracle_and_golden(product, oracle, ctx, tmp_path_factory, name, l1, l2):
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh, res = run_product(product, ctx, path, l1, l2)
    om = oracle.OracleMesh(path)
    assert np.array_equal(mesh.tris, om.tris)
    assert np.array_equal(mesh.bbox_min, om.bmin) and np.array_equal(mesh.bbox_max, om.bmax)
    ores = om.voxelize(l1, l2, oracle.FILL_CERTIFIED, 8)
    info, z = golden("%s_%d_%d" % (name, l1, l2))

    assert list(res.num_div) == list(ores.num_div) == info["num_div"]
    assert res.counts == ores.counts == [info["l1_inside"], info["l1_boundary"], info["l2_inside"], info["l2_boundary"]]
    assert res.stats["l1_box_tests"] == ores.stats["l1BoxTests"] == info["l1_box_tests"]
    assert res.stats["l1_box_hits"] == ores.stats["l1BoxHits"] == info["l1_box_hits"]
    assert res.stats["l2_box_tests"] == ores.stats["l2BoxTests"] == info["l2_box_tests"]
    assert res.stats["l2_ray_tests"] == ores.stats["l2RayTests"] == info["l2_ray_tests"]
    assert res.stats["fill_ill_conditioned"] == ores.stats["fillIllConditioned"]
    assert res.stats["fill_crossings"] == ores.stats["fillCrossings"]

    l1s = res.level1_inout()
    assert np.array_equal(l1s, ores.l1_state * 127)
    assert np.array_equal(l1_bits(l1s), z["l1_state_bits"])
    assert sha(l1s) == info["streams"]["Level1InOut"]["sha256"]
    pre = res.prefix()
    assert np.array_equal(pre, ores.prefix)
    assert sha(pre) == info["streams"]["Level1BoundaryPrefixSum"]["sha256"]
    assert np.array_equal(res.boundary_index(), ores.boundary_index)
    assert np.array_equal(res.boundary_index(), z["boundary_index"])
    l2s = res.level2_inout()
    n23 = l2 ** 3
    bad = np.nonzero((l2s != ores.l2_state * 127).reshape(-1, n23).any(1))[0]
    assert bad.size == 0, "Level-2 blocks differ for boundary ranks %s" % bad[:10]
    assert sha(l2s) == info["streams"]["Level2InOut"]["sha256"]
    # canonical CSR lists
    assert np.array_equal(res.cell_off().astype(np.int64), np.concatenate([[0], np.cumsum(ores.cell_count[ores.boundary_index])]))
    assert np.array_equal(res.cell_tris(), ores.cell_tris)
    cols = res.col_lists()
    assert np.array_equal(np.array([len(c) for c in cols]), ores.col_count)
    assert np.array_equal(np.concatenate(cols) if cols else np.zeros(0, np.int32), ores.col_tris)
    # normals: bit-exact too (same ascending accumulation order, same f32 ops)
    assert np.array_equal(res.level1_normal(), ores.l1_normal)
    assert sha(res.level1_normal()) == info["streams"]["Level1Normal"]["sha256"]
    assert np.array_equal(res.level2_normal(), ores.l2_normal)
    assert sha(res.level2_normal()) == info["streams"]["Level2Normal"]["sha256"]
    # the plain call (no normals, no lists for the caller) skips the list sorts: occupancy must not depend on the list order
    plain = ctx.voxelize(mesh, product.Params(l1, l2, 0))
    assert plain.counts == res.counts and plain.stats["l1_box_tests"] == info["l1_box_tests"] and plain.stats["l2_ray_tests"] == info["l2_ray_tests"]
    assert sha(plain.level1_inout()) == info["streams"]["Level1InOut"]["sha256"]
    assert sha(plain.prefix()) == info["streams"]["Level1BoundaryPrefixSum"]["sha256"]
    assert sha(plain.level2_inout()) == info["streams"]["Level2InOut"]["sha256"]


def test_no_level2_and_no_normals(product, oracle, ctx, tmp_path_factory):
    path = mesh_path("torus", tmp_path_factory.getbasetemp())
    mesh, res = run_product(product, ctx, path, 32, 4, flags=product.GPV_NO_LEVEL2)
    ores = oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED | oracle.NO_L2 | oracle.NO_NORMALS, 4)
    assert np.array_equal(res.level1_inout(), ores.l1_state * 127)
    assert res.c.d_level2_inout is None and res.c.d_level1_normal is None


def test_zslabs_concatenate_to_whole(product, oracle, ctx, tmp_path_factory):
    """a z-slab owns a contiguous byte range of every stream; R slabs concatenate to the 1-GPU result."""
    path = mesh_path("cessna", tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    whole = ctx.voxelize(mesh, product.Params(64, 4, product.GPV_NORMALS))
    w = dict(l1=whole.level1_inout(), pre=whole.prefix(), bi=whole.boundary_index(), l2=whole.level2_inout(), n1=whole.level1_normal(),
             n2=whole.level2_normal())
    nz = int(whole.num_div[2])
    for R in (2, 3, 8):
        parts = dict(l1=[], pre=[], bi=[], l2=[], n1=[], n2=[])
        base = 0
        for r in range(R):
            z0, z1 = nz * r // R, nz * (r + 1) // R
            s = ctx.voxelize(mesh, product.Params(64, 4, product.GPV_NORMALS, z0, z1))
            parts["l1"].append(s.level1_inout()); parts["pre"].append(s.prefix() + base); parts["bi"].append(s.boundary_index())
            parts["l2"].append(s.level2_inout()); parts["n1"].append(s.level1_normal()); parts["n2"].append(s.level2_normal())
            base += s.nb
        for k in parts:
            assert np.array_equal(np.concatenate(parts[k]), w[k]), (R, k)


@pytest.mark.parametrize("packed", [False, True])
@pytest.mark.parametrize("name,l1,l2", [("cessna", 64, 4), ("cessna", 256, 16), ("torus", 32, 4), ("cessna", 128, 8)])
def test_host_call_delivers_the_same_streams(product, oracle, ctx, tmp_path_factory, name, l1, l2, packed):
    """gpv_voxelize_host (H2D + pipeline + chunked, overlapped D2H) must hand the host exactly what the device holds -- with
    Level 2 crossing PCIe as file bytes, or (GPV_PACKED_L2) as 2 bits per sub-voxel expanded by the host-thread pool."""
    from gpview_b200 import binding as B
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    info, _ = golden("%s_%d_%d" % (name, l1, l2))
    cells = int(np.prod(info["num_div"])); nb = info["l1_boundary"]; n23 = l2 ** 3
    l1s = np.full(cells, 7, np.uint8); pre = np.full(cells, -1, np.int32); bi = np.full(nb, -1, np.int32)
    l2s = np.full(nb * n23, 7, np.uint8); n1 = np.full(cells * 3, 7, np.uint8); n2 = np.full(nb * n23 * 3, 7, np.uint8)
    hs = B.CHostStreams(l1s.ctypes.data, pre.ctypes.data, bi.ctypes.data, l2s.ctypes.data, n1.ctypes.data, n2.ctypes.data, l2s.nbytes, nb)
    flags = product.GPV_NORMALS | (product.GPV_PACKED_L2 if packed else 0)
    res = ctx.voxelize_host(mesh, product.Params(l1, l2, flags), hs)
    assert res.counts == [info["l1_inside"], info["l1_boundary"], info["l2_inside"], info["l2_boundary"]]
    s = info["streams"]
    assert sha(l1s) == s["Level1InOut"]["sha256"] and sha(pre) == s["Level1BoundaryPrefixSum"]["sha256"] and sha(bi) == s["BoundaryIndex"]["sha256"]
    assert sha(l2s) == s["Level2InOut"]["sha256"]
    assert sha(n1) == s["Level1Normal"]["sha256"] and sha(n2) == s["Level2Normal"]["sha256"]
    # too-small host buffers are refused, not overrun
    hs2 = B.CHostStreams(l1s.ctypes.data, pre.ctypes.data, bi.ctypes.data, l2s.ctypes.data, None, None, l2s.nbytes - 1, nb)
    with pytest.raises(product.GpvError):
        ctx.voxelize_host(mesh, product.Params(l1, l2, product.GPV_PACKED_L2 if packed else 0), hs2)
    if packed:  # the call after a refused one, a second packed call (pool reuse), and a destination that is not 64-byte aligned
        raw = np.full(nb * n23 + 128, 7, np.uint8)
        off = (-raw.ctypes.data) % 64 + 8
        hs3 = B.CHostStreams(None, None, None, raw.ctypes.data + off, None, None, nb * n23, nb)
        for _ in range(2):
            ctx.voxelize_host(mesh, product.Params(l1, l2, product.GPV_PACKED_L2), hs3)
            assert sha(raw[off:off + nb * n23]) == s["Level2InOut"]["sha256"]
        assert set(raw[:off]) == {7} and set(raw[off + nb * n23:]) == {7}


def test_save_from_gpu_matches_reference_files(product, oracle, ctx, tmp_path_factory, tmp_path):
    """mesh file -> gpv_voxelize_host -> gpv_save: the six ObjN* files byte-identical to the oracle's writer (== the
    reference's Object::SaveVoxelization, tests/test_oracle_ref.py)."""
    import filecmp, os
    from gpview_b200 import binding as B
    path = mesh_path("block", tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    ores = oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED, 4)
    cells, nb, n23 = ores.cells, ores.nb, ores.n23
    bufs = [np.zeros(cells, np.uint8), np.zeros(cells, np.int32), np.zeros(nb, np.int32), np.zeros(nb * n23, np.uint8), np.zeros(cells * 3, np.uint8),
            np.zeros(nb * n23 * 3, np.uint8)]
    hs = B.CHostStreams(*[b.ctypes.data for b in bufs], bufs[3].nbytes, nb)
    res = ctx.voxelize_host(mesh, product.Params(32, 4, product.GPV_NORMALS), hs)
    d1, d2 = tmp_path / "gpu", tmp_path / "ora"
    d1.mkdir(); d2.mkdir()
    B.save(mesh, res, hs, 3, str(d1))
    ores.save(3, str(d2))
    names = sorted(os.listdir(d1))
    assert len(names) == 6 and names == sorted(os.listdir(d2))
    for n in names:
        assert filecmp.cmp(d1 / n, d2 / n, shallow=False), n


def test_cli_writes_the_reference_file_set(product, oracle, tmp_path_factory, tmp_path):
    """tools/gpview_voxelize (the headless stand-in for GPView's `t` key, C++ Object facade over the C ABI) on two meshes: file
    names follow GPView's objID numbering (-1, 0, ...) and every file equals the oracle's writer byte for byte."""
    import filecmp, os, subprocess
    from util import ROOT
    exe = os.path.join(ROOT, "tools", "gpview_voxelize")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])
    paths = [mesh_path("torus", tmp_path_factory.getbasetemp()), mesh_path("sphere", tmp_path_factory.getbasetemp())]
    out = tmp_path / "cli"
    out.mkdir()
    log = subprocess.run([exe, "--l1", "32", "--l2", "4", "--out", str(out)] + paths, capture_output=True, text=True)
    assert log.returncode == 0, log.stderr
    assert "Boundary Voxels Level2" in log.stdout
    for k, p in enumerate(paths):
        ref = tmp_path / ("ora%d" % k)
        ref.mkdir()
        oracle.OracleMesh(p).voxelize(32, 4, oracle.FILL_CERTIFIED, 4).save(k - 1, str(ref))
        names = sorted(os.listdir(ref))
        assert len(names) == 6
        for n in names:
            assert filecmp.cmp(ref / n, out / n, shallow=False), n
    bad = subprocess.run([exe, str(tmp_path / "nope.obj")], capture_output=True, text=True)
    assert bad.returncode == 1 and "Unable to open file" in bad.stderr


@pytest.mark.parametrize("trial", range(9))
def test_hostile_triangle_soups_match_brute_force_oracle(product, oracle, ctx, trial):
    """Random triangle soups (not manifolds) with slivers, near-vertical walls, tiny and huge triangles, large coordinates:
    exercises the ill-conditioned (test-every-column) path of the certified fill, long per-cell lists, footprints covering the
    whole grid and every culling shortcut -- against the oracle's literal brute force (Object::ClassifyInOutCPU semantics)."""
    rng = np.random.default_rng(100 + trial)
    n = 400
    scale = [1.0, 1.0, 50.0, 500.0, 1.0, 2000.0, 1.0, 30.0, 1.0][trial]
    t = rng.uniform(-1, 1, (n, 3, 3)) * scale
    k = n // 5
    t[:k, 2] = t[:k, 0] + (t[:k, 1] - t[:k, 0]) * rng.uniform(-0.5, 1.5, (k, 1)) + rng.normal(0, 1e-6 * scale, (k, 3))      # slivers
    t[k:2 * k, 2, :2] = t[k:2 * k, 0, :2] + rng.normal(0, 1e-7 * scale, (k, 2))                                            # vertical walls
    t[2 * k:3 * k] = t[2 * k:3 * k, :1] + rng.normal(0, 0.02 * scale, (k, 3, 3))                                            # small triangles
    t[3 * k:3 * k + 5] *= 3.0                                                                                              # a few giants
    tris = t.reshape(n, 9).astype(np.float32)
    l1, l2 = [(24, 2), (20, 4), (16, 8), (12, 16), (28, 3), (24, 5), (8, 32), (20, 1), (12, 12)][trial]
    mesh = product.mesh_from_triangles(tris)
    om = oracle.OracleMesh(tris=tris)
    assert np.array_equal(mesh.bbox_min, om.bmin) and np.array_equal(mesh.bbox_max, om.bmax)
    res = ctx.voxelize(mesh, product.Params(l1, l2, product.GPV_NORMALS))
    want = om.voxelize(l1, l2, oracle.FILL_BRUTE | oracle.L2_NAIVE, 8)
    assert res.stats["fill_ill_conditioned"] == om.voxelize(l1, l2, oracle.FILL_CERTIFIED | oracle.NO_L2 | oracle.NO_NORMALS, 4).stats["fillIllConditioned"]
    assert res.counts == want.counts
    assert np.array_equal(res.level1_inout(), want.l1_state * 127)
    assert np.array_equal(res.prefix(), want.prefix)
    assert np.array_equal(res.level2_inout(), want.l2_state * 127)
    assert np.array_equal(res.cell_tris(), want.cell_tris)
    assert np.array_equal(res.level1_normal(), want.l1_normal)
    assert np.array_equal(res.level2_normal(), want.l2_normal)
    if l2 % 4 == 0:  # the 2-bit packed host transfer on every n2 it supports (ballot paths 8 / 16, the two-layer path 4, bit by bit 12 / 32)
        from gpview_b200 import binding as B
        got = np.full(want.nb * want.n23 + 64, 9, np.uint8)
        hs = B.CHostStreams(None, None, None, got.ctypes.data, None, None, want.nb * want.n23, want.nb)
        r2 = ctx.voxelize_host(mesh, product.Params(l1, l2, product.GPV_PACKED_L2), hs)
        assert r2.counts == want.counts
        assert np.array_equal(got[:want.nb * want.n23], want.l2_state * 127) and set(got[want.nb * want.n23:]) == {9}


@pytest.mark.parametrize("name,l1,l2,R", [("cessna", 64, 4, 3), ("torus", 32, 4, 2), ("cessna", 128, 8, 4), ("sphere", 24, 16, 5), ("cad", 40, 3, 2), ("block", 40, 2, 7)])
def test_peer_memory_gather_equals_whole(product, ctx, tmp_path_factory, name, l1, l2, R):
    """GPV_GATHER: R ranks -- here R contexts of one process on one device, each on its own thread and stream --
    write their shares straight into rank 0's whole-grid streams: Level-1 bytes and prefix sums by z-slab, Level-2 blocks by
    Level-1 column (every rank runs Level 1 over the whole grid, so boundary ranks are global and nothing is exchanged).  Rank 0's
    buffers must equal the single-call result byte for byte, its counts the whole grid's; run twice to exercise the epochs, once
    with explicit slabs and once with the default split.  (tests/test_gpu_multiproc.py is the one-process-per-rank version.)"""
    import threading
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    whole = ctx.voxelize(mesh, product.Params(l1, l2, 0))
    w_l1, w_pre, w_l2 = whole.level1_inout(), whole.prefix(), whole.level2_inout()
    nz, cells, n23 = int(whole.num_div[2]), whole.cells, whole.n23
    ranks = [product.Context(0) for _ in range(R)]
    try:
        ranks[0].gather_create(cells, whole.nb * n23)
        for r in range(R):
            ranks[r].gather_attach_local(ranks[0], r, R)
        cuts = [nz * r // R for r in range(R + 1)]
        cuts[1] = min(cuts[1] + 1, cuts[2]) if R > 2 else cuts[1]   # uneven explicit slabs
        # Same-process ranks only: CUDA serialises streams around every cudaMalloc / cudaFree, so nothing may allocate while a rank's
        # polling kernel is in flight.  Upload first and let one plain whole-grid call grow every context's pools.  (One process per
        # GPU -- the deployment, bench.py -- has no such constraint.)
        dev = [ranks[r].upload(mesh) for r in range(R)]
        for r in range(R):
            ranks[r].voxelize_device(dev[r], mesh, product.Params(l1, l2, 0))
        for rep in range(2):
            errs, shares = [], [0] * R
            res0 = [None]

            def work(r):
                try:
                    z0, z1 = (cuts[r], cuts[r + 1]) if rep == 0 else (0, 0)
                    res = ranks[r].voxelize_device(dev[r], mesh, product.Params(l1, l2, product.GPV_GATHER, z0, z1), ranks[r].stream())
                    shares[r] = res.n_refined
                    if r == 0:
                        res0[0] = res
                except Exception as e:  # noqa: BLE001
                    errs.append((r, repr(e)))
            th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not errs, errs
            g_l1, g_pre, g_l2, nb = ranks[0].gather_result(cells, n23)
            assert nb == whole.nb
            assert np.array_equal(g_l1, w_l1), (rep, "l1")
            assert np.array_equal(g_pre, w_pre), (rep, "prefix")
            assert np.array_equal(g_l2, w_l2), (rep, "l2")
            assert res0[0].counts == whole.counts, (rep, res0[0].counts, whole.counts)   # rank 0 reports the whole grid's counts
            assert sum(shares) == whole.nb, shares                                      # every boundary cell refined by exactly one rank
        for r in range(R):
            ranks[r].free_device(dev[r])
        # without an attached gather the flag is refused, not ignored
        with pytest.raises(product.GpvError):
            ctx.voxelize(mesh, product.Params(l1, l2, product.GPV_GATHER))
        # a gather buffer that is too small for the grid's boundary cells is refused by the call, nothing is overrun
        ranks[0].gather_detach()
        ranks[0].gather_create(cells, max(0, whole.nb * n23 - 1))
        ranks[0].gather_attach_local(ranks[0], 0, 1)
        with pytest.raises(product.GpvError):
            ranks[0].voxelize(mesh, product.Params(l1, l2, product.GPV_GATHER))
    finally:
        for c in ranks:
            c.close()


@pytest.mark.parametrize("name,l1,l2,R", [("cessna", 64, 4, 3), ("torus", 32, 8, 2), ("cad", 36, 3, 4), ("sphere", 20, 16, 2)])
def test_peer_memory_gather_with_normals(product, ctx, tmp_path_factory, name, l1, l2, R):
    """GPV_GATHER | GPV_NORMALS (the six-file contract across several GPUs): every rank also delivers the Level-1
    normals of its z-slab and the Level-2 normals of the blocks it refined; rank 0's six streams == the single call's, bit for bit
    (n2 = 4, 8, 16: 2-bit blocks kept locally for the normals and copied over; n2 = 3: byte blocks)."""
    import threading
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    whole = ctx.voxelize(mesh, product.Params(l1, l2, product.GPV_NORMALS))
    want = dict(l1=whole.level1_inout(), pre=whole.prefix(), l2=whole.level2_inout(), n1=whole.level1_normal(), n2=whole.level2_normal())
    cells, n23 = whole.cells, whole.n23
    ranks = [product.Context(0) for _ in range(R)]
    try:
        ranks[0].gather_create(cells, whole.nb * n23, product.GPV_NORMALS)
        for r in range(R):
            ranks[r].gather_attach_local(ranks[0], r, R)
        dev = [ranks[r].upload(mesh) for r in range(R)]
        for r in range(R):
            ranks[r].voxelize_device(dev[r], mesh, product.Params(l1, l2, product.GPV_NORMALS))   # grows the pools (see the test above)
        for rep in range(2):
            errs = []

            def work(r):
                try:
                    ranks[r].voxelize_device(dev[r], mesh, product.Params(l1, l2, product.GPV_GATHER | product.GPV_NORMALS), ranks[r].stream())
                except Exception as e:  # noqa: BLE001
                    errs.append((r, repr(e)))
            th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            assert not errs, errs
            g_l1, g_pre, g_l2, nb = ranks[0].gather_result(cells, n23)
            g_n1, g_n2 = ranks[0].gather_normals(cells, nb, n23)
            assert nb == whole.nb
            for k, got in (("l1", g_l1), ("pre", g_pre), ("l2", g_l2), ("n1", g_n1), ("n2", g_n2)):
                assert np.array_equal(got, want[k]), (rep, k)
        for r in range(R):
            ranks[r].free_device(dev[r])
        # buffers created without normal streams refuse the combination
        ranks[0].gather_detach()
        ranks[0].gather_create(cells, whole.nb * n23)
        ranks[0].gather_attach_local(ranks[0], 0, 1)
        with pytest.raises(product.GpvError):
            ranks[0].voxelize(mesh, product.Params(l1, l2, product.GPV_GATHER | product.GPV_NORMALS))
    finally:
        for c in ranks:
            c.close()


def test_large_grid_paths_match_oracle(product, oracle, ctx, tmp_path_factory):
    """A grid beyond 16 M cells takes the 4-sub-tile boundary scan (k_scan<CELLS, 4>) and, at n2 = 2, the flat multi-cell pair
    space of k_l2 with 64 cells per CTA: every stream against the oracle (certified fill; ~25 M cells, n2 = 2)."""
    path = mesh_path("torus", tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    res = ctx.voxelize(mesh, product.Params(448, 2, product.GPV_KEEP_LISTS))
    assert res.cells > (16 << 20)
    ores = oracle.OracleMesh(path).voxelize(448, 2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 8)
    assert res.counts == ores.counts
    assert np.array_equal(res.level1_inout(), ores.l1_state * 127)
    assert np.array_equal(res.prefix(), ores.prefix)
    assert np.array_equal(res.boundary_index(), ores.boundary_index)
    assert np.array_equal(res.level2_inout(), ores.l2_state * 127)
    assert np.array_equal(res.cell_tris(), ores.cell_tris)
