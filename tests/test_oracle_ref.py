"""Pins the oracle against the UNMODIFIED reference run live (oracle/_ref, build container only): loaders, grid sizing,
Object::ClassifyInOutCPU, Object::ClassifyTessellation, the Level-2 kernel arithmetic, and Object::SaveVoxelization."""
import filecmp
import os

import numpy as np
import pytest

from util import HAVE_REF, REF, mesh_path

pytestmark = [pytest.mark.ref, pytest.mark.skipif(not HAVE_REF, reason="needs /root/reference and oracle/_ref (build container)")]


@pytest.mark.parametrize("name,l1,l2", [("cessna", 8, 4), ("cessna", 32, 4), ("torus", 16, 2), ("block", 24, 4), ("cad", 20, 8)])
def test_oracle_equals_reference_live(oracle, tmp_path_factory, tmp_path, name, l1, l2):
    from oracle import refbind
    path = os.path.join(REF, "files", "cessna.obj") if name == "cessna" else mesh_path(name, tmp_path_factory.getbasetemp())
    ro = refbind.RefObject(path, obj_id=7)
    om = oracle.OracleMesh(path)
    assert np.array_equal(ro.tris, om.tris)
    assert np.array_equal(ro.bmin, om.bmin) and np.array_equal(ro.bmax, om.bmax) and ro.max_model_size == om.max_model_size
    ro.setup(l1, l2)
    ro.l1_inout_brute(0)                 # Object::ClassifyInOutCPU itself
    ro.l1_tribox()                       # Object::ClassifyTessellation
    ro.compact()
    ro.l2_kernelform(4)
    cpu2 = None
    if l2 <= 4:
        ro.l2_cpu()                      # the reference's CPU Level-2 twins (f64 centres) -- cross-check only
        cpu2 = ro.level2_inout()
    ro.adopt_kernelform()
    cnt = ro.count()
    r = om.voxelize(l1, l2, oracle.FILL_BRUTE | oracle.L2_NAIVE, 4)
    assert list(r.num_div) == list(ro.num_div)
    assert np.array_equal(r.grid_size, ro.grid_size) and np.array_equal(r.grid_size2, ro.grid_size2)
    assert np.array_equal(r.l1_state, ro.level1_inout().astype(np.uint8))
    assert np.array_equal(r.prefix, ro.prefix())
    assert np.array_equal(r.boundary_index, ro.boundary_index())
    assert np.array_equal(r.l2_state, ro.level2_inout_kernel().astype(np.uint8))
    assert r.counts == cnt
    if cpu2 is not None:
        # The reference's CPU twins compute sub-voxel centres in f64 (src/Object.cpp:2389-2391, :1158-1160), the shipped CUDA
        # kernels in f32 (cu:423-425): the two are NOT bit-identical in general (e.g. 6 of 4,608 voxels on
        # torus 16/2).  The oracle follows the kernels; the twins only have to agree almost everywhere.
        assert (cpu2 != ro.level2_inout_kernel()).mean() < 0.01
    # the six files, byte for byte, written by Object::SaveVoxelization vs the oracle's writer
    d1, d2 = tmp_path / "ref", tmp_path / "ora"
    d1.mkdir(); d2.mkdir()
    ro.save(str(d1))
    r.save(7, str(d2))
    names = sorted(os.listdir(d1))
    assert names == sorted(os.listdir(d2)) and len(names) == 6
    for n in names:
        assert filecmp.cmp(d1 / n, d2 / n, shallow=False), n
    ro.close()


def test_threaded_brute_force_equals_member_function():
    """oracle/gen_golden.py uses the 8-thread loop nest for the 128/256 fixtures: it must equal Object::ClassifyInOutCPU."""
    from oracle import refbind
    path = os.path.join(REF, "files", "cessna.obj")
    a = refbind.RefObject(path); a.setup(24, 2); a.l1_inout_brute(0)
    b = refbind.RefObject(path); b.setup(24, 2); b.l1_inout_brute(4)
    assert np.array_equal(a.level1_inout(), b.level1_inout())
    a.close(); b.close()


def test_oracle_equals_reference_on_random_meshes(oracle, tmp_path):
    """The pin beyond the fixture meshes: random triangle soups, scattered small triangles, quantised (degenerate-rich) soups,
    slivers / near-vertical walls and a perturbed closed body, at random small resolutions -- reference live (brute-force fill,
    its own tri-box classification, kernel-form Level 2) == oracle brute force == the oracle's certified fast paths, which
    is what the GPU tests compare with.  (A 15-minute soak of this generator, 28,017 meshes, found no difference.)"""
    from oracle import refbind
    from gpview_b200 import meshgen
    rng = np.random.default_rng(424242)
    p = str(tmp_path / "s.obj")
    for it in range(150):
        nt, kind = int(rng.integers(4, 200)), it % 5
        if kind == 0:
            V = rng.uniform(-1, 1, (nt * 3, 3))
        elif kind == 1:
            V = (rng.uniform(-1, 1, (nt, 1, 3)) + rng.normal(0, 0.05, (nt, 3, 3))).reshape(-1, 3)
        elif kind == 2:
            V = np.round(rng.uniform(-1, 1, (nt * 3, 3)) * 4) / 4
        elif kind == 3:
            a, b, t = rng.uniform(-1, 1, (nt, 3)), rng.uniform(-1, 1, (nt, 3)), rng.uniform(0, 1, (nt, 1))
            V = np.stack([a, b, a + (b - a) * t + rng.normal(0, 1e-5, (nt, 3))], 1).reshape(-1, 3)
        else:
            Vs, Fs = meshgen.uv_sphere(12, 8)
            V = Vs[Fs].reshape(-1, 3) + rng.normal(0, 0.02, (len(Fs) * 3, 3))
        V = V.astype(np.float32)
        meshgen.write_obj(p, V, np.arange(len(V)).reshape(-1, 3))
        l1, l2 = int(rng.choice([4, 8, 12, 20])), int(rng.choice([1, 2, 3, 4]))
        ro, om = refbind.RefObject(p, obj_id=7), oracle.OracleMesh(p)
        assert np.array_equal(ro.tris, om.tris) and np.array_equal(ro.bmin, om.bmin) and np.array_equal(ro.bmax, om.bmax), it
        ro.setup(l1, l2); ro.l1_inout_brute(0); ro.l1_tribox(); ro.compact(); ro.l2_kernelform(2); ro.adopt_kernelform()
        cnt = ro.count()
        r = om.voxelize(l1, l2, oracle.FILL_BRUTE | oracle.L2_NAIVE, 2)
        rc = om.voxelize(l1, l2, oracle.FILL_CERTIFIED, 2)
        where = (it, kind, l1, l2)
        assert list(r.num_div) == list(ro.num_div), where
        assert np.array_equal(r.l1_state, ro.level1_inout().astype(np.uint8)), where
        assert np.array_equal(r.prefix, ro.prefix()) and np.array_equal(r.boundary_index, ro.boundary_index()), where
        assert np.array_equal(r.l2_state, ro.level2_inout_kernel().astype(np.uint8)) and r.counts == cnt, where
        assert np.array_equal(rc.l1_state, r.l1_state) and np.array_equal(rc.l2_state, r.l2_state) and rc.counts == r.counts, where
        if it % 3 == 0:  # the normals as Object::SaveVoxelization encodes them (5,810 meshes in a soak of their own)
            d = tmp_path / ("n%d" % it)
            d.mkdir()
            ro.save(str(d))
            assert np.array_equal(np.fromfile(d / "Obj7Level1Normal.raw", np.uint8), rc.l1_normal), where
            assert np.array_equal(np.fromfile(d / "Obj7Level2Normal.raw", np.uint8), rc.l2_normal), where
        ro.close()


def test_product_obj_reader_equals_the_live_reference_on_quirky_files(tmp_path):
    """The product's strict OBJ reader against Object::ReadObject itself on 1,500 random files from the subset the reference
    survives (it assert()s on `f` / `vn` / `vt` lines with other field counts, throws on stof range errors and writes out of
    bounds on long `v` lines): tab- or space-separated lines, short `v` lines (stale coordinates), numeric prefixes ("1.5abc",
    "4\\r"), "a/b", "a//c", "a/b/c", leading zeros and plus signs, comments and other keywords, a missing final newline --
    triangles and padded bounding box bit for bit.  (261,151 files in a soak.)"""
    import gpview_b200 as gpv
    from oracle import refbind
    rng = np.random.default_rng(31337)
    p = str(tmp_path / "r.obj")
    nums = ["0", "1", "-1.5", "2.25e0", "+3", ".5", "7.", "1e-3", "0.333333343", "1.5abc", "4\r", "-0", "00012.5", "3.4e38"]
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    for it in range(1500):
        lines, nv = ["v 0 0 0", "v 1 0 0", "v 0 1 0"], 3
        for _ in range(int(rng.integers(1, 14))):
            r, d1 = rng.random(), (" " if rng.random() < 0.8 else "\t")
            if r < 0.45:
                k = 3 if rng.random() < 0.8 else int(rng.integers(0, 3))
                lines.append(d1.join(["v"] + [nums[rng.integers(len(nums))] for _ in range(k)]))
                nv += 1
            elif r < 0.85:
                def ix():
                    s, v = str(int(rng.integers(1, nv + 1))), rng.random()
                    return s if v < 0.5 else s + "/1" if v < 0.65 else s + "//2" if v < 0.8 else s + "/1/2" if v < 0.9 else ("0" + s if v < 0.95 else "+" + s)
                lines.append(d1.join(["f", ix(), ix(), ix()]) + ("\r" if rng.random() < 0.15 else ""))
            else:
                lines.append(["vn 0 0 1", "vt 0.5 0.5", "# comment", "g grp", "", "usemtl m", "s off", "o obj", "vn\t0\t1\t0"][rng.integers(9)])
        txt = "\n".join(lines) + ("\n" if rng.random() < 0.85 else "")
        with open(p, "w", newline="") as f:
            f.write(txt)
        ro = refbind.RefObject(p)
        pm = gpv.load_mesh(p)
        want = ro.tris.reshape(-1, 9)
        assert pm.ntri == len(want) and np.array_equal(bits(pm.tris), bits(want)), (it, txt)
        assert np.array_equal(bits(pm.bbox_min), bits(ro.bmin)) and np.array_equal(bits(pm.bbox_max), bits(ro.bmax)), (it, txt)
        ro.close()


def test_product_off_reader_equals_the_live_reference(tmp_path):
    """The product's strict OFF reader against Object::ReadOFFObject itself on 1,500 random well-formed files: any whitespace
    between tokens (the reference reads with operator>>), number forms like "+3", ".5", "7.", "1E2", "00012.5", face counts other
    than 3 (read and ignored: exactly three indices follow), leading zeros in indices, trailing tokens, no final newline --
    triangles and padded bounding box (over the referenced vertices) bit for bit.  (104,504 files in a soak.)"""
    import gpview_b200 as gpv
    from oracle import refbind
    rng = np.random.default_rng(4242)
    p = str(tmp_path / "r.off")
    nums = ["0", "1", "-1.5", "2.25e0", "+3", ".5", "7.", "1e-3", "0.333333343", "-0", "00012.5", "3.4e38", "-19.6246052", "1E2", "5e+0"]
    ws = [" ", "\n", "\t", "  ", "\r\n", " \n "]
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    for it in range(1500):
        nv, nf = int(rng.integers(3, 12)), int(rng.integers(1, 10))
        toks = ["OFF", str(nv), str(nf), str(int(rng.integers(0, 50)))]
        toks += [nums[rng.integers(len(nums))] for _ in range(nv * 3)]
        for _ in range(nf):
            toks.append(str(int(rng.choice([3, 3, 3, 4, 0, 7]))))
            toks += [str(int(rng.integers(0, nv))) if rng.random() < 0.9 else "0" + str(int(rng.integers(0, nv))) for _ in range(3)]
        if rng.random() < 0.2:
            toks += ["trailing", "1", "2"]
        txt = "".join(t + ws[rng.integers(len(ws))] for t in toks)
        if rng.random() < 0.1:
            txt = txt.rstrip()
        with open(p, "w", newline="") as f:
            f.write(txt)
        ro = refbind.RefObject(p)
        pm = gpv.load_mesh(p)
        want = ro.tris.reshape(-1, 9)
        assert pm.ntri == len(want) and np.array_equal(bits(pm.tris), bits(want)), (it, txt)
        assert np.array_equal(bits(pm.bbox_min), bits(ro.bmin)) and np.array_equal(bits(pm.bbox_max), bits(ro.bmax)), (it, txt)
        ro.close()


def test_grid_sizing_equals_the_live_reference_on_random_models(tmp_path):
    """gpv_make_grid (+ the loader's padded bounding box and maxModelSize) against the reference's own arithmetic
    (src/Object.cpp:572-583, 3094-3134) on 600 random models: coordinates from 1e-3 to 1e4, offsets up to 1e4, cubes (ties
    between the axes) and flat boxes, Level-1 counts around the GetNextDiv4 steps, Level-2 1..32 -- numDiv, gridSize, gridSize2
    bit for bit.  (165,280 models in a soak.)"""
    import gpview_b200 as gpv
    from gpview_b200 import meshgen
    from oracle import refbind
    rng = np.random.default_rng(99)
    p = str(tmp_path / "g.obj")
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    refused = 0
    for it in range(600):
        scale, off = 10.0 ** rng.uniform(-3, 4), rng.uniform(-1, 1, 3) * 10.0 ** rng.uniform(-2, 4)
        ext = np.ones(3) if rng.random() < 0.2 else 10.0 ** rng.uniform(-1.5, 0, 3)
        V = (rng.uniform(-1, 1, (12, 3)) * ext * scale + off).astype(np.float32)
        meshgen.write_obj(p, V, np.arange(12).reshape(-1, 3))
        l1, l2 = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 16, 31, 33, 63, 64, 65, 100, 127, 129])), int(rng.integers(1, 33))
        ro, pm = refbind.RefObject(p), gpv.load_mesh(p)
        ro.setup(l1, l2)
        try:
            g = gpv.grid_for(pm.bbox_min, pm.bbox_max, pm.max_model_size, l1, l2)
        except gpv.GpvError:      # a box that collapses in f32 (tiny model far from the origin): refused, the reference divides by zero
            refused += 1
            ro.close()
            continue
        where = (it, l1, l2, list(V[0]))
        assert pm.max_model_size == ro.max_model_size and list(g.num_div) == [int(x) for x in ro.num_div], where
        assert np.array_equal(bits(list(g.grid_size)), bits(ro.grid_size)) and np.array_equal(bits(list(g.grid_size2)), bits(ro.grid_size2)), where
        ro.close()
    assert refused < 100, refused


def test_writer_equals_save_voxelization_on_random_models(tmp_path):
    """gpv_save fed with the reference's own arrays against Object::SaveVoxelization on 200 random models whose sizes span eleven
    orders of magnitude (the %g text of ObjNVoxelConfig.txt: fixed, scientific, negative), object ids -1..499: the config file,
    Level1InOut, Level1BoundaryPrefixSum and Level2InOut byte for byte.  (14,779 models in a soak.)"""
    import shutil
    import gpview_b200 as gpv
    from gpview_b200 import binding as B, meshgen
    from oracle import refbind
    rng = np.random.default_rng(2024)
    p = str(tmp_path / "m.obj")
    Vs, Fs = meshgen.uv_sphere(8, 6)
    done = 0
    for it in range(200):
        scale = 10.0 ** rng.uniform(-5, 6)
        V = (Vs * rng.uniform(0.3, 1, 3) * scale + rng.uniform(-1, 1, 3) * scale * 10.0 ** rng.uniform(-1, 2)).astype(np.float32)
        meshgen.write_obj(p, V, Fs)
        l1, l2, oid = int(rng.choice([4, 8, 12])), int(rng.choice([1, 2, 3])), int(rng.integers(-1, 500))
        ro, pm = refbind.RefObject(p, obj_id=oid), gpv.load_mesh(p)
        ro.setup(l1, l2); ro.l1_inout_brute(0); ro.l1_tribox(); ro.compact(); ro.l2_kernelform(1); ro.adopt_kernelform()
        cnt = ro.count()
        d1, d2 = tmp_path / "ref", tmp_path / "gpv"
        for d in (d1, d2):
            shutil.rmtree(d, ignore_errors=True)
            d.mkdir()
        ro.save(str(d1))
        try:
            g = gpv.grid_for(pm.bbox_min, pm.bbox_max, pm.max_model_size, l1, l2)
        except gpv.GpvError:
            ro.close()
            continue
        res = B.CResult()
        res.grid = g
        l1s = (ro.level1_inout() * np.float32(127)).astype(np.uint8)
        pre = ro.prefix().copy()
        l2s = (ro.level2_inout() * np.float32(127)).astype(np.uint8)
        res.cells, res.n_boundary, res.n23 = len(l1s), ro.nboundary(), l2 ** 3
        res.l1_inside, res.l1_boundary, res.l2_inside, res.l2_boundary = cnt
        hs = B.CHostStreams(l1s.ctypes.data, pre.ctypes.data, None, l2s.ctypes.data, None, None, l2s.nbytes, 0)

        class R:
            c = res
        B.save(pm, R, hs, oid, str(d2))
        names = sorted(os.listdir(d1))
        assert names == sorted(os.listdir(d2)) and len(names) == 6, (it, names)
        for n in names:
            if "Normal" not in n:
                assert filecmp.cmp(d1 / n, d2 / n, shallow=False), (it, n, scale, open(d1 / n, "rb").read()[:300], open(d2 / n, "rb").read()[:300])
        ro.close()
        done += 1
    assert done > 150


EMU = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libgpvref_emu.so")


def _emu_lib():
    import ctypes as C
    from oracle import refbind
    refbind.LIB_PATH, refbind._lib = EMU, None
    L = refbind.lib()
    L.ref_cuda_path.argtypes = [C.c_void_p]
    L.ref_cuda_path.restype = C.c_int
    return L


def _restore_ref_lib():
    from oracle import refbind
    refbind.LIB_PATH, refbind._lib = os.path.join(os.path.dirname(EMU), "libgpvref.so"), None


def _reference_gpu_path_vs_oracle(oracle, L, path, l1, l2, files_dir=None):
    """Object::ClassifyTessellationCUDA (with its two-pass buffer re-run) + Object::ClassifyInOutTessellationLevel2CUDA on the
    host-executed kernels; the GL solid fill, which cannot run headless, is seeded from the oracle.  Returns the buffer size used."""
    import ctypes as C
    from oracle import refbind
    want = oracle.OracleMesh(path).voxelize(l1, l2, oracle.FILL_CERTIFIED, 4)
    ro = refbind.RefObject(path, obj_id=7)
    ro.setup(l1, l2)
    fill = want.l1_fill_only.astype(np.float32)
    C.memmove(L.ref_level1InOut(ro.h), fill.ctypes.data, fill.nbytes)
    used = L.ref_cuda_path(ro.h)
    where = (path, l1, l2, used)
    assert np.array_equal(ro.level1_inout().astype(np.uint8), want.l1_state), where
    assert np.array_equal(ro.boundary_index(), want.boundary_index), where
    assert np.array_equal(ro.level2_inout().astype(np.uint8), want.l2_state), where
    assert ro.count() == want.counts, where
    # Level-2 normals after the reference's host averaging (src/Object.cpp:2613-2632) in its file encoding: the emulated threads run
    # in ascending order, which is the oracle's canonical accumulation order, so even these agree bit for bit
    n = ro.level2_normal().reshape(-1, 4)[:, :3].reshape(-1)
    assert np.array_equal((n * np.float32(256.0 / 3.0) + np.float32(127.0)).astype(np.uint8), want.l2_normal), where
    if files_dir is not None:  # ... and Object::SaveVoxelization behind that path writes the oracle's six files, byte for byte
        d1, d2 = os.path.join(files_dir, "ref"), os.path.join(files_dir, "ora")
        for d in (d1, d2):
            os.makedirs(d)
        ro.save(d1)
        want.save(7, d2)
        names = sorted(os.listdir(d1))
        assert names == sorted(os.listdir(d2)) and len(names) == 6, where
        for f in names:
            assert filecmp.cmp(os.path.join(d1, f), os.path.join(d2, f), shallow=False), (where, f)
    ro.close()
    return used


@pytest.mark.skipif(not os.path.exists(EMU), reason="oracle/_ref/libgpvref_emu.so not built")
@pytest.mark.parametrize("name,l1,l2", [("cessna", 32, 4), ("torus", 24, 8), ("cessna", 64, 4), ("block", 20, 3), ("cad", 16, 5)])
def test_reference_gpu_path_on_host_executed_kernels_equals_the_oracle(oracle, tmp_path_factory, tmp_path, name, l1, l2):
    """The last link of the pin: the reference's GPU path SOURCE FOR SOURCE -- its unmodified host code driving its unmodified
    kernel source (cuda/CUDAClassifyTessellation.cu compiled as C++ and run thread by thread, oracle/ref_kernels_host.cpp) --
    gives the oracle's Level-1 states, boundary list, Level-2 states, counts and normals, and Object::SaveVoxelization behind it the
    oracle's six files byte for byte.  So the "kernel form" the
    oracle restates IS what the reference's kernels compute (under strict IEEE; g++ -ffp-contract=off == nvcc -fmad=false)."""
    path = os.path.join(REF, "files", "cessna.obj") if name == "cessna" else mesh_path(name, tmp_path_factory.getbasetemp())
    try:
        used = _reference_gpu_path_vs_oracle(oracle, _emu_lib(), path, l1, l2, str(tmp_path))
    finally:
        _restore_ref_lib()
    if name == "cessna":
        assert used == (946 if l1 == 32 else 329)   # more triangles per cell than the default buffer of 50: the re-run (src/Object.cpp:3204-3214)


@pytest.mark.skipif(not os.path.exists(EMU), reason="oracle/_ref/libgpvref_emu.so not built")
def test_reference_gpu_path_on_host_executed_kernels_random_meshes(oracle, tmp_path):
    from gpview_b200 import meshgen
    rng = np.random.default_rng(8080)
    p = str(tmp_path / "s.obj")
    try:
        L = _emu_lib()
        for it in range(60):
            nt, kind = int(rng.integers(4, 150)), it % 4
            if kind == 0:
                V = rng.uniform(-1, 1, (nt * 3, 3))
            elif kind == 1:
                V = (rng.uniform(-1, 1, (nt, 1, 3)) + rng.normal(0, 0.08, (nt, 3, 3))).reshape(-1, 3)
            elif kind == 2:
                V = np.round(rng.uniform(-1, 1, (nt * 3, 3)) * 4) / 4
            else:
                Vs, Fs = meshgen.uv_sphere(10, 7)
                V = Vs[Fs].reshape(-1, 3) + rng.normal(0, 0.02, (len(Fs) * 3, 3))
            V = V.astype(np.float32)
            meshgen.write_obj(p, V, np.arange(len(V)).reshape(-1, 3))
            _reference_gpu_path_vs_oracle(oracle, L, p, int(rng.choice([4, 8, 12, 16])), int(rng.choice([1, 2, 3, 4])))
    finally:
        _restore_ref_lib()


@pytest.mark.skipif(not os.path.exists(EMU), reason="oracle/_ref/libgpvref_emu.so not built")
def test_reference_gpu_path_tool_plumbing(oracle, tmp_path_factory):
    """tools/bench_reference_gpu_path.py (reference's kernels vs compat vs native, for the GPU box) with its reference worker on the
    host-executed kernels: the worker runs through to its JSON line and reports the oracle's counts."""
    import json
    import subprocess
    import sys
    from util import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_reference_gpu_path.py"), "--emulated-only", "--mesh", "torus", "--l1", "16", "--l2", "2", "--reps", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    run = out["runs"]["emulated"]
    want = oracle.OracleMesh(mesh_path("torus", tmp_path_factory.getbasetemp())).voxelize(16, 2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 2)
    assert run["counts"] == want.counts and run["tri_buffer"] == 50 and len(run["seconds"]) == 1


@pytest.mark.parametrize("name,l1", [("sphere", 16), ("sphere", 32), ("block", 24), ("torus", 32), ("cessna", 64)])
def test_collision_boxes_and_hierarchy_equal_the_reference(oracle, tmp_path_factory, name, l1):
    """the oracle's restatement of Object::CollisionInitCUDA's host loop (src/Object.cpp:3530-3552) and of
    Object::BuildHierarchy / CombineBBox (:2750-2867) against the reference's own code on its own BBoxData array -- every box's
    midPoint, halfSize, solid flag and child indices bit for bit, on the grids the reference's loop is defined on (all dimensions powers
    of two); on the others (GetNextDiv4 grids in general) it indexes out of bounds and the oracle refuses."""
    import ctypes as C
    from oracle import refbind
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    r = oracle.OracleMesh(path).voxelize(l1, 0, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 4)
    ro = refbind.RefObject(path)
    ro.setup(l1, 0)
    fill = r.l1_fill_only.astype(np.float32)
    C.memmove(refbind.lib().ref_level1InOut(ro.h), fill.ctypes.data, fill.nbytes)      # the GL fill's stand-in, as everywhere in this file
    ro.l1_tribox()                                                                      # Object::ClassifyTessellation: marks the boundary cells, builds bBox[]
    assert np.array_equal(ro.level1_inout().astype(np.uint8), r.l1_state)
    inv, mid, ext = r.collision_boxes()
    rinv, rmid, rext = ro.collision_boxes()
    assert len(inv) == r.counts[0] + r.counts[1] and np.array_equal(inv, rinv) and np.array_equal(mid, rmid) and np.array_equal(ext, rext)
    h = r.build_hierarchy()
    pow2 = all(int(n) & (int(n) - 1) == 0 for n in r.num_div)
    assert (h is not None) == pow2
    if h is not None:
        lv, hm, hh, hs, hc = h
        rlv, rm, rh, rs, rc = ro.build_hierarchy(r.l1_fill_only)
        assert lv == rlv and np.array_equal(hm, rm) and np.array_equal(hh, rh) and np.array_equal(hs, rs) and np.array_equal(hc, rc)
        assert hs[-1] == (1 if r.l1_fill_only.any() else 0)                             # the root is solid iff any cell's fill is
    ro.close()
