"""Host side of the native tier on the CPU: loaders with Object::ReadObject / ReadOFFObject semantics, grid sizing of
Object::PerformVoxelization, and the Object::SaveVoxelization file contract -- product (gpv_*) vs oracle, and vs the
reference itself in the build container.  No compute entry point is called (no GPU here)."""
import ctypes as C
import filecmp
import os

import numpy as np
import pytest

from util import GOLDEN_CASES, HAVE_REF, REF, golden, mesh_path


def _same_mesh(a, b):
    return (np.array_equal(a.tris, b.tris) and np.array_equal(np.asarray(a.bbox_min if hasattr(a, "bbox_min") else a.bmin), b.bmin)
            and np.array_equal(np.asarray(a.bbox_max if hasattr(a, "bbox_max") else a.bmax), b.bmax))


@pytest.mark.parametrize("name", ["cessna", "sphere", "torus", "block", "cad"])
def test_loader_matches_oracle(product, oracle, tmp_path_factory, name):
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    pm, om = product.load_mesh(path), oracle.OracleMesh(path)
    assert pm.ntri == om.ntri and _same_mesh(pm, om)
    assert pm.max_model_size == om.max_model_size


OBJ_QUIRKS = {
    # the reference splits on single ' ' OR single '\\t' (whichever yields more fields), reads "a/b/c" faces, keeps stale
    # coordinates on short `v` lines, takes the bbox over ALL `v` lines and drops an unterminated last line (App. B9)
    "tabs": "v\t0\t0\t0\nv\t1\t0\t0\nv\t0\t1\t0\nv\t0\t0\t1\nf\t1\t2\t3\nf\t1\t2\t4\n",
    "slashes": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\nf 1/1/1 2/1/1 3/1/1\nf 1//1 2//1 4//1\nf 2/1 3/1 4/1\n",
    "unreferenced_vertex_in_bbox": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 9 9 9\nf 1 2 3\n",
    "no_final_newline": "v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 2 4",
    "crlf": "v 0 0 0\r\nv 1 0 0\r\nv 0 1 0\r\nf 1 2 3\r\n",
    "comments_and_groups": "# c\no obj\ng grp\nv 0 0 0\nv 1.5e0 0 0\nv 0 2.25 0\nusemtl m\ns off\nf 1 2 3\n",
    "short_vertex_line": "v 1 2 3\nv 4 5\nv 0 0 0\nf 1 2 3\n",
}


@pytest.mark.parametrize("case", sorted(OBJ_QUIRKS))
def test_obj_quirks(product, oracle, tmp_path, case):
    p = tmp_path / (case + ".obj")
    p.write_bytes(OBJ_QUIRKS[case].encode())
    pm, om = product.load_mesh(str(p)), oracle.OracleMesh(str(p))
    assert pm.ntri == om.ntri and _same_mesh(pm, om)
    if case == "no_final_newline":
        assert pm.ntri == 1
    if case == "unreferenced_vertex_in_bbox":
        assert pm.bbox_max[0] > 9
    if case == "short_vertex_line":
        assert list(pm.tris[0][3:6]) == [4, 5, 3]
    if HAVE_REF:
        from oracle import refbind
        ro = refbind.RefObject(str(p))
        assert np.array_equal(ro.tris, pm.tris) and np.array_equal(ro.bmin, pm.bbox_min) and np.array_equal(ro.bmax, pm.bbox_max)
        ro.close()


def test_obj_errors_are_reported_not_aborted(product, tmp_path):
    for name, txt in {"double_space": "v 0  0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "bad_index": "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n",
                      "empty": "\n"}.items():
        p = tmp_path / (name + ".obj")
        p.write_text(txt)
        with pytest.raises(product.GpvError):
            product.load_mesh(str(p))
    with pytest.raises(product.GpvError):
        product.load_mesh(str(tmp_path / "missing.obj"))
    with pytest.raises(product.GpvError):
        product.load_mesh(str(tmp_path / "mesh.stl"))


def test_off_reads_exactly_three_indices_and_bbox_over_referenced(product, oracle, tmp_path):
    p = tmp_path / "q.off"
    p.write_text("OFF\n5 2 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n50 50 50\n3 0 1 2\n3 0 1 3\n")
    pm, om = product.load_mesh(str(p)), oracle.OracleMesh(str(p))
    assert pm.ntri == 2 and _same_mesh(pm, om)
    assert pm.bbox_max[0] < 2          # vertex 4 is not referenced: OFF bbox ignores it (src/Object.cpp:257-266)
    p2 = tmp_path / "oneline.off"
    p2.write_text("OFF 4 2 0 0 0 0 1 0 0 0 1 0 0 0 1 3 0 1 2 3 0 1 3")
    assert product.load_mesh(str(p2)).ntri == 2


def _load_with_threads(product, path, t):
    os.environ["GPV_LOAD_THREADS"] = str(t)
    try:
        return product.load_mesh(path)
    finally:
        os.environ.pop("GPV_LOAD_THREADS", None)


def test_chunk_parallel_loading_is_independent_of_the_chunk_count(product, oracle, tmp_path_factory, tmp_path):
    """the loaders cut a file into chunks parsed on host threads.  Every float, the triangle order, the bbox and
    the error that is reported must not depend on the number of chunks: fixtures, every OBJ quirk, a 60 k-triangle mesh in both
    formats (against the oracle's sequential reader), and files whose first error sits in the middle."""
    from gpview_b200 import meshgen
    files = [mesh_path(n, tmp_path_factory.getbasetemp()) for n in ("cessna", "sphere", "torus", "block", "cad")]
    for case, txt in OBJ_QUIRKS.items():
        p = tmp_path / (case + ".obj")
        p.write_bytes(txt.encode())
        files.append(str(p))
    V, F = meshgen.uv_sphere(200, 152)
    meshgen.write_obj(str(tmp_path / "big.obj"), V, F)
    meshgen.write_off(str(tmp_path / "big.off"), V, F)
    files += [str(tmp_path / "big.obj"), str(tmp_path / "big.off")]
    for path in files:
        om = oracle.OracleMesh(path)
        for t in (1, 2, 3, 7):
            pm = _load_with_threads(product, path, t)
            assert pm.ntri == om.ntri and _same_mesh(pm, om) and pm.max_model_size == om.max_model_size, (path, t)
    # errors: the earliest one in file order, whatever the cut
    lines = ["v %d 0 0" % i for i in range(3000)] + ["f %d %d %d" % (i + 1, i + 2, i + 3) for i in range(2900)]
    broken = {
        "bad_index_mid": lines[:4000] + ["f 1 2 99999"] + lines[4000:] + ["f 1 2 0"],
        "bad_float_mid": lines[:1500] + ["v 1 x 3"] + lines[1500:4000] + ["f 1 2 99999"] + lines[4000:],
        "forward_reference": ["v 0 0 0", "v 1 0 0", "f 1 2 3", "v 0 1 0"] + lines,
    }
    for name, ls in broken.items():
        p = tmp_path / (name + ".obj")
        p.write_text("\n".join(ls) + "\n")
        msgs = []
        for t in (1, 2, 5, 7):
            with pytest.raises(product.GpvError) as ei:
                _load_with_threads(product, str(p), t)
            msgs.append(str(ei.value))
        assert len(set(msgs)) == 1, (name, msgs)
    off_bad = tmp_path / "bad.off"
    off_bad.write_text("OFF\n4 3 0\n0 0 0\n1 0 0\n0 1 0\n0 0 1\n3 0 1 2\n3 0 1 9\n3 0 x 2\n")
    msgs = []
    for t in (1, 2, 3):
        with pytest.raises(product.GpvError) as ei:
            _load_with_threads(product, str(off_bad), t)
        msgs.append(str(ei.value))
    assert len(set(msgs)) == 1 and "out of range" in msgs[0], msgs


@pytest.mark.parametrize("name,l1,l2", GOLDEN_CASES)
def test_grid_sizing_matches_reference_fixture(product, tmp_path_factory, name, l1, l2):
    info, _ = golden("%s_%d_%d" % (name, l1, l2))
    m = product.load_mesh(mesh_path(name, tmp_path_factory.getbasetemp()))
    g = product.grid_for(m.bbox_min, m.bbox_max, m.max_model_size, l1, l2)
    assert list(g.num_div) == info["num_div"]
    assert [np.float32(x).tobytes().hex() for x in g.grid_size] == info["grid_size_hex"]
    assert [float(np.float32(x)) for x in g.grid_size2] == pytest.approx(info["grid_size2"], rel=0, abs=0)


def test_grid_sizing_fuzz_vs_oracle(product, oracle):
    rng = np.random.default_rng(9)
    for _ in range(300):
        lo = rng.uniform(-100, 100, 3).astype(np.float32)
        hi = (lo + 10.0 ** rng.uniform(-3, 3, 3)).astype(np.float32)
        ms = np.float32(np.max(hi - lo))
        l1, l2 = int(rng.integers(1, 1500)), int(rng.integers(1, 33))
        a, b = product.grid_for(lo, hi, ms, l1, l2), oracle.make_grid(lo, hi, ms, l1, l2)
        assert list(a.num_div) == list(b.numDiv) and list(a.grid_size) == list(b.gridSize) and list(a.grid_size2) == list(b.gridSize2)
        assert list(a.ext1) == list(b.ext1) and list(a.ext2) == list(b.ext2)


def test_save_writes_the_reference_file_contract(product, oracle, tmp_path_factory, tmp_path):
    """gpv_save from host streams == the oracle's writer == (build container) Object::SaveVoxelization, byte for byte."""
    from gpview_b200 import binding as B
    path = mesh_path("torus", tmp_path_factory.getbasetemp())
    pm = product.load_mesh(path)
    r = oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED, 4)
    res = B.CResult()
    res.grid = product.grid_for(pm.bbox_min, pm.bbox_max, pm.max_model_size, 32, 4)
    res.cells, res.n_boundary, res.n23 = r.cells, r.nb, r.n23
    res.l1_inside, res.l1_boundary, res.l2_inside, res.l2_boundary = r.counts
    l1 = (r.l1_state * 127).astype(np.uint8); l2 = (r.l2_state * 127).astype(np.uint8)
    keep = [l1, r.prefix, l2, r.l1_normal, r.l2_normal]
    hs = B.CHostStreams(l1.ctypes.data, r.prefix.ctypes.data, None, l2.ctypes.data, r.l1_normal.ctypes.data, r.l2_normal.ctypes.data, l2.nbytes, 0)
    d1, d2 = tmp_path / "gpv", tmp_path / "ora"
    d1.mkdir(); d2.mkdir()

    class R:  # minimal stand-in for binding.Result
        c = res
    B.save(pm, R, hs, -1, str(d1))
    r.save(-1, str(d2))
    names = sorted(os.listdir(d1))
    assert names == ["Obj-1Level1BoundaryPrefixSum.raw", "Obj-1Level1InOut.raw", "Obj-1Level1Normal.raw", "Obj-1Level2InOut.raw",
                     "Obj-1Level2Normal.raw", "Obj-1VoxelConfig.txt"]
    for n in names:
        assert filecmp.cmp(d1 / n, d2 / n, shallow=False), n
    # streams that were not computed: gpv_save writes their neutral value (the reference's contract is six files), gpv_save_streams
    # with omit_absent leaves them out (dataset runs: the filler is 3/4 of the bytes); gpv_load_voxels reads either set
    hs2 = B.CHostStreams(l1.ctypes.data, r.prefix.ctypes.data, None, l2.ctypes.data, None, None, l2.nbytes, 0)
    d3, d4 = tmp_path / "filled", tmp_path / "lean"
    d3.mkdir(); d4.mkdir()
    B.save(pm, R, hs2, 3, str(d3))
    B.save(pm, R, hs2, 3, str(d4), omit_absent=True)
    assert sorted(os.listdir(d3)) == [n.replace("Obj-1", "Obj3") for n in names]
    assert sorted(os.listdir(d4)) == [n.replace("Obj-1", "Obj3") for n in names if "Normal" not in n]
    assert set(open(d3 / "Obj3Level2Normal.raw", "rb").read()) == {127} and os.path.getsize(d3 / "Obj3Level2Normal.raw") == l2.nbytes * 3
    for n in os.listdir(d4):
        assert filecmp.cmp(d3 / n, d4 / n, shallow=False), n
    va, vb = B.load_voxels(str(d3), 3), B.load_voxels(str(d4), 3)
    assert vb["level1_normal"] is None and vb["level2_normal"] is None and va["level2_normal"] is not None
    assert np.array_equal(va["level2_inout"], vb["level2_inout"]) and np.array_equal(vb["level2_inout"], l2) and np.array_equal(vb["prefix_sum"], r.prefix)
    assert keep


def test_voxel_files_round_trip(product, oracle, tmp_path_factory, tmp_path):
    """gpv_load_voxels reads back what SaveVoxelization-format writers produce (sizes from ObjNVoxelConfig.txt), with and
    without Level 2 / normals, and refuses files whose sizes contradict the config."""
    from gpview_b200 import binding as B
    path = mesh_path("sphere", tmp_path_factory.getbasetemp())
    r = oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED, 4)
    r.save(5, str(tmp_path))
    v = B.load_voxels(str(tmp_path), 5)
    assert v["name"] == "Obj5" and v["num_div"] == list(r.num_div) and v["num_div2"] == [4, 4, 4] and v["counts"] == r.counts
    assert np.array_equal(v["level1_inout"], r.l1_state * 127) and np.array_equal(v["prefix_sum"], r.prefix)
    assert np.array_equal(v["level2_inout"], r.l2_state * 127)
    assert np.array_equal(v["level1_normal"], r.l1_normal) and np.array_equal(v["level2_normal"], r.l2_normal)
    assert v["grid_size"] == pytest.approx([float(x) for x in r.grid_size], rel=1e-5)     # the config prints 6 significant digits
    os.remove(tmp_path / "Obj5Level2Normal.raw")
    assert B.load_voxels(str(tmp_path), 5)["level2_normal"] is None                        # optional stream
    with open(tmp_path / "Obj5Level2InOut.raw", "ab") as f:
        f.write(b"\\0")
    with pytest.raises(product.GpvError):
        B.load_voxels(str(tmp_path), 5)                                                    # size contradicts the config
    r1 = oracle.OracleMesh(path).voxelize(32, 0, oracle.FILL_CERTIFIED, 4)
    d2 = tmp_path / "l1only"
    d2.mkdir()
    r1.save(0, str(d2))
    v1 = B.load_voxels(str(d2), 0)
    assert v1["num_div2"] is None and v1["level2_inout"] is None and np.array_equal(v1["level1_inout"], r1.l1_state * 127)
    with pytest.raises(product.GpvError):
        B.load_voxels(str(d2), 9)


def test_interrupted_save_leaves_no_config_and_check_voxels_sees_truncation(product, oracle, tmp_path_factory, tmp_path):
    """ObjNVoxelConfig.txt is the completeness marker (batch restart, Dataset listing): gpv_save_streams writes it LAST, by
    rename, and removes a stale one first -- a save that fails on a stream leaves no config; gpv_check_voxels accepts only sets
    whose streams have exactly the sizes the config implies (a truncated stream of a killed run is not 'complete')."""
    from gpview_b200 import binding as B
    from gpview_b200 import dataset
    path = mesh_path("block", tmp_path_factory.getbasetemp())
    pm = product.load_mesh(path)
    r = oracle.OracleMesh(path).voxelize(16, 2, oracle.FILL_CERTIFIED, 2)
    res = B.CResult()
    res.grid = product.grid_for(pm.bbox_min, pm.bbox_max, pm.max_model_size, 16, 2)
    res.cells, res.n_boundary, res.n23 = r.cells, r.nb, r.n23
    res.l1_inside, res.l1_boundary, res.l2_inside, res.l2_boundary = r.counts
    l1 = (r.l1_state * 127).astype(np.uint8); l2 = (r.l2_state * 127).astype(np.uint8)
    hs = B.CHostStreams(l1.ctypes.data, r.prefix.ctypes.data, None, l2.ctypes.data, None, None, l2.nbytes, 0)

    class R:
        c = res
    d = tmp_path / "set"
    d.mkdir()
    assert not B.check_voxels(str(d), 7)
    B.save(pm, R, hs, 7, str(d), omit_absent=True)
    assert B.check_voxels(str(d), 7) and dataset.list_object_ids(str(d)) == [7]
    assert not [n for n in os.listdir(d) if n.endswith(".tmp")]
    # a later save of the same id that cannot open a stream (a directory sits where Level2InOut.raw belongs): error, and the
    # config of the earlier, complete run is gone too -- it must not vouch for streams that were being rewritten
    os.remove(d / "Obj7Level2InOut.raw")
    os.mkdir(d / "Obj7Level2InOut.raw")
    with pytest.raises(product.GpvError):
        B.save(pm, R, hs, 7, str(d), omit_absent=True)
    assert not os.path.exists(d / "Obj7VoxelConfig.txt") and not B.check_voxels(str(d), 7) and dataset.list_object_ids(str(d)) == []
    os.rmdir(d / "Obj7Level2InOut.raw")
    B.save(pm, R, hs, 7, str(d), omit_absent=True)
    assert B.check_voxels(str(d), 7)
    # truncated stream (what a killed run or a full disk leaves when an OLD writer put the config first)
    with open(d / "Obj7Level2InOut.raw", "r+b") as f:
        f.truncate(l2.nbytes - 1)
    assert not B.check_voxels(str(d), 7) and "size" in product.lib().gpv_last_error().decode()
    assert dataset.list_object_ids(str(d)) == []
    with open(d / "Obj7Level2InOut.raw", "ab") as f:
        f.write(b"\0")
    os.remove(d / "Obj7Level1BoundaryPrefixSum.raw")
    assert not B.check_voxels(str(d), 7) and "missing" in product.lib().gpv_last_error().decode()


# ------------------------------------------------------------------------------------------------ number fields of the loaders
@pytest.fixture(scope="module")
def parse_probe():
    """gpview_b200/csrc/gpv_parse.h compiled for the test (tests/cpu_probe/parse_probe.cpp)."""
    import subprocess
    from util import ROOT
    here = os.path.join(ROOT, "tests", "cpu_probe")
    so, src, hdr = os.path.join(here, "parse_probe.so"), os.path.join(here, "parse_probe.cpp"), os.path.join(ROOT, "gpview_b200", "csrc", "gpv_parse.h")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-o", so, src])
    L = C.CDLL(so)
    for n in ("probe_parse_float_fuzz", "probe_parse_long_fuzz"):
        getattr(L, n).restype = C.c_int64
        getattr(L, n).argtypes = [C.c_uint64, C.c_int64, C.c_int, C.c_char_p, C.POINTER(C.c_int64)]
    L.probe_parse_float.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_float)]
    L.probe_parse_long.argtypes = [C.c_char_p, C.c_int64, C.POINTER(C.c_long)]
    L.probe_eight_digits.argtypes = [C.c_char_p]
    L.probe_eight_digits.restype = C.c_uint32
    return L


FIELD_KINDS = ["%.9g of random floats", "fixed-point coordinates", "float midpoints (exact, cut, nudged)", "digit soup", "garbage and tails", "integers"]


@pytest.mark.parametrize("kind", range(6))
def test_number_fields_equal_strtof_and_strtol(parse_probe, kind):
    """The loaders' scanners must return what std::stof / std::stoi (strtof / strtol) return on the same field -- value bits,
    whether a conversion happened, and where it ended -- on 1.5 M seeded fields of each kind, both as coordinates and as
    indices.  Without the float-midpoint guard of gpv_parse.h the midpoint kind alone disagrees ~4000 times per million."""
    first, fast = C.create_string_buffer(64), C.c_int64()
    for seed in (1, 20240607, 77):
        bad = parse_probe.probe_parse_float_fuzz(seed, 500_000, kind, first, C.byref(fast))
        assert bad == 0, (FIELD_KINDS[kind], "float", first.value)
        if kind < 4:
            assert fast.value > 100_000, (FIELD_KINDS[kind], "the short path is never taken", fast.value)
        bad = parse_probe.probe_parse_long_fuzz(seed, 500_000, kind, first, C.byref(fast))
        assert bad == 0, (FIELD_KINDS[kind], "long", first.value)


def test_number_fields_known_cases(parse_probe):
    """Hand-picked fields: (text, converts?, value) as strtof / strtol define them."""
    import struct
    f32 = lambda x: struct.unpack("f", struct.pack("f", x))[0]
    cases = [(b"1", 1.0), (b"-0", -0.0), (b"+.5", 0.5), (b"1.", 1.0), (b"1.e2", 100.0), (b"1e", 1.0), (b"1e+", 1.0), (b"1.5abc", 1.5), (b"1.25\r", 1.25),
             (b"0x10", 16.0), (b"0x1p3", 8.0), (b"1e39", float("inf")), (b"1e-46", 0.0), (b"1.17549435e-38", f32(1.17549435e-38)), (b"1e-40", f32(1e-40)),
             (b"16777217", 16777216.0), (b"16777219", 16777220.0), (b"0.1", f32(0.1)), (b"-19.624605178833008", f32(-19.624605178833008)),
             (b"3.4028234663852886e38", f32(3.4028234663852886e38)), (b"inf", float("inf")), (b"-Infinity", float("-inf")), (b" 2", 2.0),
             (b"00000000000000000000000001.5", 1.5), (b"0.000000000000000000000000000015e29", f32(1.5)), (b"123456789012345678901234567890", f32(1.2345678901234568e29))]
    v = C.c_float()
    for txt, want in cases:
        assert parse_probe.probe_parse_float(txt, len(txt), C.byref(v)) == 1, txt
        assert struct.pack("f", v.value) == struct.pack("f", want), (txt, v.value, want)
    for txt in (b"", b".", b"-", b"e5", b"x", b"+-1", b"- 1", b"/3"):
        assert parse_probe.probe_parse_float(txt, len(txt), C.byref(v)) == 0, txt
    n = C.c_long()
    for txt, want in [(b"7", 7), (b"-12", -12), (b"+3", 3), (b"007", 7), (b"12/5", 12), (b" 4", 4), (b"99999999999999999999", 2**63 - 1), (b"-99999999999999999999", -2**63), (b"1e5", 1)]:
        assert parse_probe.probe_parse_long(txt, len(txt), C.byref(n)) == 1 and n.value == want, (txt, n.value)
    for txt in (b"", b"-", b"x1", b"/1", b"+ 1"):
        assert parse_probe.probe_parse_long(txt, len(txt), C.byref(n)) == 0, txt
    rng = np.random.default_rng(5)
    for x in list(rng.integers(0, 10**8, 2000)) + [0, 99999999, 10**7, 12345678]:
        assert parse_probe.probe_eight_digits(b"%08d" % x) == x
    for txt in (b"1234567 ", b"/2345678", b"1234:678", b"12345678"[:7] + b"\xb9"):
        assert parse_probe.probe_eight_digits(txt) == 0xFFFFFFFF, txt


def test_loader_keeps_its_results_across_reuse_of_the_thread_scratch(product, oracle, tmp_path_factory, tmp_path):
    """The loaders keep their scratch memory and the last freed triangle block per host thread: loading large, small, failing
    and large files again in one thread must give every time what a fresh load gives (== the oracle's sequential reader)."""
    from gpview_b200 import meshgen
    V, F = meshgen.uv_sphere(60, 40)
    meshgen.write_obj(str(tmp_path / "s.obj"), V, F)
    meshgen.write_off(str(tmp_path / "s.off"), V, F)
    (tmp_path / "tiny.off").write_text("OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    (tmp_path / "bad.off").write_text("OFF\n3 2 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n3 0 1\n")
    (tmp_path / "lying.off").write_text("OFF\n3 2000000000 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    (tmp_path / "lying2.off").write_text("OFF\n2000000000 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    seq = ["s.obj", "tiny.off", "s.off", "bad.off", "s.obj", "lying.off", "lying2.off", "tiny.off", "s.off", "cessna", "tiny.off", "s.obj"]
    held = []
    for i, name in enumerate(seq):
        path = mesh_path("cessna", tmp_path_factory.getbasetemp()) if name == "cessna" else str(tmp_path / name)
        if name in ("bad.off", "lying.off", "lying2.off"):
            with pytest.raises(product.GpvError):
                product.load_mesh(path)
            continue
        pm, om = product.load_mesh(path), oracle.OracleMesh(path)
        assert pm.ntri == om.ntri and _same_mesh(pm, om), (i, name)
        held.append((pm, om))  # meshes stay alive (and are freed in a different order): blocks must not be shared
        if i % 3 == 2:
            held.pop(0)
    for pm, om in held:
        assert np.array_equal(pm.tris, om.tris)


def test_obj_line_fuzz_vs_oracle(product, oracle, tmp_path):
    """Random OBJ text -- spaces, tabs, both, doubled and trailing delimiters, CR, slashed and signed indices, short and long `v`
    lines, other keywords starting with v / f -- through the product's in-place field walk and the oracle's literal
    split-twice reader: same mesh, or both refuse.  (The reference itself cannot referee here: it assert()s on `f` / `vn` / `vt`
    lines with an unexpected field count and writes out of bounds on long `v` lines; test_obj_quirks covers it on the forms it survives.)"""
    rng = np.random.default_rng(20240607)
    nums = ["0", "1", "-1.5", "2.25e0", "+3", ".5", "7.", "1e-3", "0x10", "4\r", "1.5abc", "", "x", "-", "1e", "3/4"]
    idxs = ["1", "2", "3", "4", "1/2", "2//3", "3/1/1", "04", "+2", "4\r", "", "/1", "0", "9", "-1", "2.7", "1e0", "x"]
    heads = ["v", "v", "v", "f", "f", "vn", "vt", "#", "g", "fo", "", "vv", "V", "F"]
    agree_ok = agree_fail = 0
    for it in range(400):
        lines = ["v 0 0 0", "v 1 0 0", "v 0 1 0", "v 0 0 1"]
        for _ in range(int(rng.integers(1, 12))):
            h = heads[rng.integers(len(heads))]
            pool = idxs if h == "f" else nums
            risky = rng.random() < 0.35                       # most lines well-formed, so that many files load
            k = int(rng.integers(0, 6)) if risky else 3
            toks = [pool[rng.integers(len(pool) if risky else 4)] for _ in range(k)]
            style = rng.integers(5)
            d = [" ", "\t", " ", "\t", " "][style]
            line = d.join([h] + toks)
            if style == 2 and risky:                          # mixed: some delimiters swapped for the other kind
                line = "".join(("\t" if (c == " " and rng.random() < 0.4) else c) for c in line)
            if style == 3 and risky:
                line = "".join((" " if (c == "\t" and rng.random() < 0.5) else c) for c in line)
            if risky and rng.random() < 0.2:
                line += d
            lines.append(line)
        p = tmp_path / ("fuzz%d.obj" % it)
        p.write_bytes(("\n".join(lines) + ("\n" if rng.random() < 0.9 else "")).encode())
        try:
            om = oracle.OracleMesh(str(p))
        except RuntimeError:
            om = None
        try:
            pm = product.load_mesh(str(p))
        except product.GpvError:
            pm = None
        assert (om is None) == (pm is None), (it, lines)
        if om is not None:
            assert pm.ntri == om.ntri and _same_mesh(pm, om), (it, lines)
            agree_ok += 1
        else:
            agree_fail += 1
    assert agree_ok > 40 and agree_fail > 40, (agree_ok, agree_fail)
    print("obj fuzz: %d files load identically, %d refused by both" % (agree_ok, agree_fail))


# ------------------------------------------------------------------------------------------------ tolerant readers (extension)
def test_tolerant_readers_equal_the_strict_ones_on_clean_files(product, oracle, tmp_path_factory):
    """GPV_LOAD_TOLERANT is a superset: on files the reference's readers take as meant (all fixtures) it yields the same mesh."""
    for name in ("cessna", "sphere", "torus", "block", "cad"):
        path = mesh_path(name, tmp_path_factory.getbasetemp())
        a, b = product.load_mesh(path), product.load_mesh(path, tolerant=True)
        assert a.ntri == b.ntri and np.array_equal(a.tris, b.tris) and np.array_equal(a.bbox_min, b.bbox_min) and np.array_equal(a.bbox_max, b.bbox_max)
        for t in (2, 5):
            os.environ["GPV_LOAD_THREADS"] = str(t)
            try:
                c = product.load_mesh(path, tolerant=True)
            finally:
                os.environ.pop("GPV_LOAD_THREADS", None)
            assert np.array_equal(c.tris, a.tris)


def test_tolerant_obj_reads_what_the_strict_reader_refuses(product, tmp_path):
    txt = ("# a quad, a pentagon, relative indices, free-form blanks, CRLF, a w coordinate, no final newline\r\n"
           "v  0 0 0\r\nv\t1  0 0 1.0\r\nv 1 1 0\r\n  v 0 1 0\r\nvn 0 0 1\r\nv 0.5 1.5 0 # apex\r\n"
           "f 1 2 3 4\r\nf  1/1/1 2/1/1  3/1/1 5//1 4\r\nf -5 -4 -1\r\nf 1 2 3")
    p = tmp_path / "wild.obj"
    p.write_bytes(txt.encode())
    with pytest.raises(product.GpvError):
        product.load_mesh(str(p))
    m = product.load_mesh(str(p), tolerant=True)
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 1.5, 0]], np.float32)
    want = [(0, 1, 2), (0, 2, 3),            # quad
            (0, 1, 2), (0, 2, 4), (0, 4, 3),  # pentagon 1 2 3 5 4
            (0, 1, 4),                        # -5 -4 -1 with five vertices defined
            (0, 1, 2)]                        # the unterminated last line
    assert m.ntri == len(want)
    assert np.array_equal(m.tris, np.array([V[list(t)].reshape(9) for t in want], np.float32))
    assert m.bbox_max[1] > 1.5 and m.bbox_min[0] < 0
    for bad in ("v 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 0\n",
                "v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 7\n", "v 0 0 0\nv 1 0 0\nv 0 1 0\nf -4 1 2\n", "v 0 0 0\nv 1 0 0\nf 1 2 3\nv 0 1 0\n"):
        q = tmp_path / "bad.obj"
        q.write_text(bad)
        with pytest.raises(product.GpvError):
            product.load_mesh(str(q), tolerant=True)


def test_tolerant_off_reads_polygons_comments_and_colours(product, oracle, tmp_path):
    txt = ("OFF # header with a comment\n# counts on their own line\n5 3 0\n\n0 0 0\n1 0 0 255 0 0\n1 1 0\n0 1 0\n0.5 1.5 0\n"
           "4 0 1 2 3  0.5 0.5 0.5\n3 3 2 4\n5 0 1 2 4 3\n trailing junk is ignored\n")
    p = tmp_path / "wild.off"
    p.write_text(txt)
    m = product.load_mesh(str(p), tolerant=True)
    V = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0.5, 1.5, 0]], np.float32)
    want = [(0, 1, 2), (0, 2, 3), (3, 2, 4), (0, 1, 2), (0, 2, 4), (0, 4, 3)]
    assert m.ntri == len(want) and np.array_equal(m.tris, np.array([V[list(t)].reshape(9) for t in want], np.float32))
    # the strict reader (reference semantics: three indices whatever the count says) misreads the same file
    try:
        s = product.load_mesh(str(p))
        assert not (s.ntri == m.ntri and np.array_equal(s.tris, m.tris))
    except product.GpvError:
        pass
    one = tmp_path / "oneline.off"
    one.write_text("OFF 3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n")
    assert product.load_mesh(str(one), tolerant=True).ntri == 1
    for bad in ("COFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n", "OFF\n3 1 0\n0 0 0\n1 0 0\n0 1\n3 0 1 2\n", "OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 3\n",
                "OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 0\n4 0 1 2\n", "OFF\n3 2 0\n0 0 0\n1 0 0\n0 1 0\n3 0 1 2\n", "OFF\n99999999 1 0\n0 0 0\n", ""):
        q = tmp_path / "bad.off"
        q.write_text(bad)
        with pytest.raises(product.GpvError):
            product.load_mesh(str(q), tolerant=True)


# ------------------------------------------------------------------------------------------------ consumer side: dense grids, Dataset
def test_dense_expansion_and_dataset(product, oracle, tmp_path_factory, tmp_path):
    """gpv_expand_dense / gpview_b200.dataset: the two-level file set as one grid at the effective resolution, against a direct
    numpy construction; the torch Dataset over a directory of sets."""
    from gpview_b200 import binding as B, dataset as D
    path = mesh_path("torus", tmp_path_factory.getbasetemp())
    for obj_id, (l1, l2) in enumerate([(16, 4), (12, 3), (8, 1)]):
        r = oracle.OracleMesh(path).voxelize(l1, l2, oracle.FILL_CERTIFIED, 4)
        r.save(obj_id, str(tmp_path))
        nx, ny, nz = [int(x) for x in r.num_div]
        n2 = l2
        want = np.repeat(np.repeat(np.repeat((r.l1_state * 127).reshape(nz, ny, nx), n2, 0), n2, 1), n2, 2)
        blocks = (r.l2_state * 127).reshape(r.nb, n2, n2, n2)
        for b, cell in enumerate(r.boundary_index):
            z, y, x = cell // (nx * ny), (cell // nx) % ny, cell % nx
            want[z * n2:(z + 1) * n2, y * n2:(y + 1) * n2, x * n2:(x + 1) * n2] = blocks[b]
        got = B.expand_dense(r.l1_state * 127, r.prefix, r.l2_state * 127, r.num_div, n2)
        assert got.shape == want.shape and np.array_equal(got, want)
        assert int((got == 254).sum()) == r.counts[3] and int((got == 127).sum()) == r.counts[0] * n2 ** 3 + r.counts[2]
        g2, meta = D.load_grid(str(tmp_path), obj_id, "dense")
        assert np.array_equal(g2, want // 127) and meta["num_div"] == [nx, ny, nz]
        g1, _ = D.load_grid(str(tmp_path), obj_id, "level1", occupancy=True)
        assert np.array_equal(g1, (r.l1_state > 0).reshape(nz, ny, nx))
    assert D.list_object_ids(str(tmp_path)) == [0, 1, 2]
    ds = D.VoxelFolder(str(tmp_path))
    x, meta = ds[1]
    assert len(ds) == 3 and tuple(x.shape) == (1, meta["num_div"][2] * 3, meta["num_div"][1] * 3, meta["num_div"][0] * 3) and int(x.max()) == 2
    # streams that do not belong together are refused, not read out of bounds
    r = oracle.OracleMesh(path).voxelize(16, 4, oracle.FILL_CERTIFIED, 4)
    with pytest.raises(product.GpvError):
        B.expand_dense(r.l1_state * 127, r.prefix, (r.l2_state * 127)[: 10 * 64], r.num_div, 4)


def test_off_token_fuzz_vs_oracle(product, oracle, tmp_path):
    """Random OFF text -- every kind of whitespace between tokens, odd but valid number forms (nan, 1e-40, 00012.5, 1.5abc), face
    counts other than 3, signed / fractional / out-of-range indices, truncations and trailing tokens -- through the product's
    fused token scanner and the oracle's operator>>-style reader: bit-identical mesh (NaNs included), or both refuse.
    (A four-minute soak of this generator, 634,767 files, found no difference.)"""
    rng = np.random.default_rng(77)
    nums = ["0", "1", "-1.5", "2.25e0", "+3", ".5", "7.", "1e-3", "0.333333343", "-19.6246052", "1e10", "3.4e38", "1e-40", "00012.5"]
    bad = ["x", "1.5abc", "", "--1", "e5", "0x10", "nan", "inf", "1e999"]
    ws = [" ", "\n", "\t", "  ", "\r\n", " \n "]
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
    same = refused = 0
    p = str(tmp_path / "f.off")
    for it in range(3000):
        nv, nf = int(rng.integers(3, 12)), int(rng.integers(1, 10))
        toks = ["OFF", str(nv), str(nf), "0"]
        risky = rng.random() < 0.3
        for _ in range(nv * 3):
            toks.append(nums[rng.integers(len(nums))] if not (risky and rng.random() < 0.03) else bad[rng.integers(len(bad))])
        for _ in range(nf):
            toks.append(str(3 if not risky or rng.random() < 0.8 else int(rng.integers(0, 6))))
            for _ in range(3):
                toks.append(str(int(rng.integers(0, nv))) if not (risky and rng.random() < 0.05) else ["-1", str(nv), "99", "x", "1.5", "+1", "01"][rng.integers(7)])
        if risky and rng.random() < 0.3:
            toks = toks[:int(rng.integers(1, len(toks)))]
        if risky and rng.random() < 0.2:
            toks += ["extra", "tokens", "1", "2"]
        txt = "".join(t + ws[rng.integers(len(ws))] for t in toks)
        if rng.random() < 0.1:
            txt = txt.rstrip()
        with open(p, "w", newline="") as f:
            f.write(txt)
        try:
            om = oracle.OracleMesh(p)
        except RuntimeError:
            om = None
        try:
            pm = product.load_mesh(p)
        except product.GpvError:
            pm = None
        assert (om is None) == (pm is None), (it, txt)
        if om is None:
            refused += 1
            continue
        same += 1
        assert pm.ntri == om.ntri and np.array_equal(bits(pm.tris), bits(om.tris)), (it, txt)
        assert np.array_equal(bits(pm.bbox_min), bits(om.bmin)) and np.array_equal(bits(pm.bbox_max), bits(om.bmax)), (it, txt)
    assert same > 1500 and refused > 300, (same, refused)


# an independent statement of the tolerant OBJ semantics (include/gpview_b200.h, gpv_load_mesh_ex) for the model-based test below
import re  # noqa: E402


def _f32(s):
    # strtof-prefix semantics via the product's strict number scanner is what we test elsewhere; here tokens are clean numbers
    return np.float32(float(s))
def _tolerant_obj_model(txt):
    """the documented tolerant OBJ semantics, independently: returns list of triangles (9 floats) or None on error"""
    if not txt.endswith("\n"): txt+="\n"
    V=[]; T=[]
    for line in txt.split("\n")[:-1]:
        toks=[t for t in re.split(r"[ \t\r]+", line) if t!=""]
        if not toks or toks[0] not in ("v","f"): continue
        body=[]
        for t in toks[1:]:
            if t.startswith("#"): break
            body.append(t)
        if toks[0]=="v":
            if len(body)<3: return None
            try: V.append([_f32(body[0]),_f32(body[1]),_f32(body[2])])
            except ValueError: return None
        else:
            idx=[]
            for t in body:
                h=t.split("/")[0]
                if not re.fullmatch(r"[+-]?\d+", h): return None
                i=int(h)
                if i==0: return None
                idx.append(i)
            if len(idx)<3: return None
            nv=len(V); res=[]
            for i in idx:
                j=i-1 if i>0 else nv+i
                if j<0 or j>=nv: return None
                res.append(j)
            for k in range(1,len(res)-1): T.append(V[res[0]]+V[res[k]]+V[res[k+1]])
    if not V: return None
    return np.array(T,np.float32).reshape(-1,9)


def test_tolerant_obj_reader_against_an_independent_model(product, tmp_path):
    """2,500 random OBJ texts (runs of blanks, CR, comments, short and long `v` lines, polygons of 3-6 vertices, positive /
    negative / slashed / zero / out-of-range indices, other keywords, missing final newline): the tolerant reader and a
    30-line Python statement of its documented semantics agree bit for bit, or both refuse.  (83,723 files in a soak.)"""
    rng = np.random.default_rng(5)
    p = str(tmp_path / "t.obj")
    nums = ["0", "1", "-1.5", "2.25", "+3", ".5", "7.", "1e-3", "0.333333343"]
    seps = [" ", "  ", "\t", " \t "]
    loaded = refused = 0
    for it in range(2500):
        lines, nv = [], 0
        for _ in range(int(rng.integers(3, 14))):
            r = rng.random()
            sep = lambda: seps[rng.integers(len(seps))]
            if r < 0.5 or nv < 3:
                k = 3 if rng.random() < 0.9 else int(rng.integers(0, 6))
                lines.append((" " if rng.random() < 0.1 else "") + "v" + sep() + sep().join(nums[rng.integers(len(nums))] for _ in range(k)) +
                             (" # c" if rng.random() < 0.1 else "") + ("\r" if rng.random() < 0.2 else ""))
                nv += 1
            elif r < 0.9:
                k = int(rng.integers(3, 7)) if rng.random() < 0.9 else int(rng.integers(0, 3))

                def ix():
                    u = rng.random()
                    i = int(rng.integers(1, nv + 1)) if u < 0.8 else (-int(rng.integers(1, nv + 1)) if u < 0.92 else int(rng.choice([0, nv + 1, -nv - 1, 99])))
                    s, v = str(i), rng.random()
                    return s if v < 0.6 else s + "/1" if v < 0.75 else s + "//2" if v < 0.9 else s + "/1/2"
                lines.append("f" + sep() + sep().join(ix() for _ in range(k)) + ("\r" if rng.random() < 0.2 else ""))
            else:
                lines.append(["vn 0 0 1", "# comment", "g grp", "", "vt 0 0", "usemtl m"][rng.integers(6)])
        txt = "\n".join(lines) + ("\n" if rng.random() < 0.8 else "")
        with open(p, "w", newline="") as f:
            f.write(txt)
        want = _tolerant_obj_model(txt)
        try:
            got = np.array(product.load_mesh(p, tolerant=True).tris)
        except product.GpvError:
            got = None
        assert (want is None) == (got is None), (it, txt)
        if want is None:
            refused += 1
            continue
        loaded += 1
        assert want.shape == got.shape and np.array_equal(want.view(np.uint32), got.view(np.uint32)), (it, txt)
    assert loaded > 500 and refused > 500, (loaded, refused)


# ... and of the tolerant OFF semantics
_INT = re.compile(r"[+-]?\d+")
def _tolerant_off_model(txt):
    rows=[]
    for line in txt.split("\n"):
        line=line.split("#")[0]
        t=line.split()
        if t: rows.append(t)
    if not rows or rows[0][0]!="OFF": return None
    i=0
    if len(rows[0])>=4: cnt=rows[0][1:]
    else:
        if len(rows)<2: return None
        cnt=rows[1]; i=1
    i+=1
    if len(cnt)<2 or not _INT.fullmatch(cnt[0]) or not _INT.fullmatch(cnt[1]): return None
    nv=int(cnt[0]); nf=int(cnt[1])
    if nv<=0 or nf<=0: return None
    V=[]
    for _ in range(nv):
        if i>=len(rows) or len(rows[i])<3: return None
        try: V.append([np.float32(float(x)) for x in rows[i][:3]])
        except ValueError: return None
        i+=1
    T=[]
    for _ in range(nf):
        if i>=len(rows): return None
        r=rows[i]; i+=1
        if not _INT.fullmatch(r[0]): return None
        n=int(r[0])
        if n<3 or len(r)<n+1: return None
        idx=[]
        for x in r[1:n+1]:
            if not _INT.fullmatch(x): return None
            q=int(x)
            if q<0 or q>=nv: return None
            idx.append(q)
        for k in range(1,n-1): T.append(V[idx[0]]+V[idx[k]]+V[idx[k+1]])
    return np.array(T,np.float32).reshape(-1,9)


def test_tolerant_off_reader_against_an_independent_model(product, tmp_path):
    """2,500 random OFF texts (counts on the header line or the next, comments, blank lines, colour fields behind vertices and
    faces, polygons of 3-6 vertices, short records, bad indices, truncations, trailing lines): the tolerant reader and a Python
    statement of its documented semantics agree bit for bit, or both refuse.  (78,930 files in a soak.)"""
    rng = np.random.default_rng(9)
    p = str(tmp_path / "t.off")
    nums = ["0", "1", "-1.5", "2.25", "3", ".5", "7.", "1e-3", "0.333333343"]
    loaded = refused = 0
    for it in range(2500):
        nv, nf, risky = int(rng.integers(3, 9)), int(rng.integers(1, 6)), rng.random() < 0.35
        L = []
        head = "OFF" if rng.random() < 0.95 or not risky else "COFF"
        if rng.random() < 0.3:
            L.append("%s %d %d 0" % (head, nv, nf))
        else:
            L.append(head + (" # hdr" if rng.random() < 0.2 else ""))
            if rng.random() < 0.2:
                L.append("# comment line")
            L.append("%d %d 0" % (nv, nf))
        for _ in range(nv):
            k = 3 if not risky or rng.random() < 0.9 else int(rng.integers(0, 3))
            L.append(" ".join(nums[rng.integers(len(nums))] for _ in range(k)) + (" 255 0 0" if rng.random() < 0.2 else "") + ("\r" if rng.random() < 0.1 else ""))
            if rng.random() < 0.1:
                L.append("")
        for _ in range(nf):
            n = int(rng.integers(3, 7)) if not risky or rng.random() < 0.85 else int(rng.integers(0, 3))
            ids = [str(int(rng.integers(0, nv))) if not risky or rng.random() < 0.93 else str(int(rng.choice([-1, nv, 99])))
                   for _ in range(n if not risky or rng.random() < 0.9 else max(0, n - 1))]
            L.append(" ".join([str(n)] + ids) + ("  0.5 0.5 0.5" if rng.random() < 0.2 else ""))
        if risky and rng.random() < 0.3:
            L = L[:int(rng.integers(1, len(L)))]
        if rng.random() < 0.2:
            L.append("trailing junk")
        txt = "\n".join(L) + ("\n" if rng.random() < 0.8 else "")
        with open(p, "w", newline="") as f:
            f.write(txt)
        want = _tolerant_off_model(txt)
        try:
            got = np.array(product.load_mesh(p, tolerant=True).tris)
        except product.GpvError:
            got = None
        assert (want is None) == (got is None), (it, txt)
        if want is None:
            refused += 1
            continue
        loaded += 1
        assert want.shape == got.shape and np.array_equal(want.view(np.uint32), got.view(np.uint32)), (it, txt)
    assert loaded > 800 and refused > 300, (loaded, refused)


@pytest.mark.parametrize("isa", ["", "avx2", "scalar"])
def test_expand_packed_l2_matches_numpy_on_every_isa_path(product, isa):
    """gpv_expand_packed_l2 (the host half of GPV_PACKED_L2 / the consumer of the 2-bit Level-2 stream): (inside, boundary) mask
    pairs -> bytes 0 / 127 / 254, on the AVX-512, AVX2 and scalar paths (GPV_EXPAND_ISA picks a narrower one), aligned and
    unaligned destinations, odd word counts."""
    import subprocess, sys, textwrap
    from util import ROOT
    code = textwrap.dedent('''
        import ctypes as C, numpy as np, sys
        sys.path.insert(0, %r)
        import gpview_b200 as gpv
        L = gpv.lib()
        rng = np.random.default_rng(7)
        for n, shift in [(1, 0), (2, 0), (3, 32), (64, 0), (1001, 0), (1001, 32), (4096, 1), (5000, 17), (70001, 0)]:
            inside = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
            bd = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
            inside &= ~bd                                  # a sub-voxel is never both
            packed = np.stack([inside, bd], 1).copy()
            raw = np.zeros(n * 32 + 128, np.uint8)
            base = (-raw.ctypes.data) %% 64 + shift            # 64-byte aligned + shift
            out = raw[base:base + n * 32]
            assert L.gpv_expand_packed_l2(packed.ctypes.data, n, out.ctypes.data) == 0
            bits = lambda w: ((w[:, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8).reshape(-1)
            want = bits(inside) * 127 + bits(bd) * 254
            assert np.array_equal(out, want), (n, shift)
            assert raw[:base].sum() == 0 and raw[base + n * 32:].sum() == 0
        print("ok")
    ''') % ROOT
    env = dict(os.environ)
    if isa:
        env["GPV_EXPAND_ISA"] = isa
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
