"""BASELINE.json's full-size configurations on the GPU (configs[2], [3], [4]; the file is named to run after the other GPU
files): 1 M-triangle sphere and torus at Level-1 512 + Level-2 8^3, the 10 M-triangle CAD body at 1024 + 2^3 (2048^3 effective), drilled-block .off meshes at 64 + 4^3.

The CPU oracle needs 20 s to minutes for these, so the device streams are compared with sha256 values committed in
tests/golden/fullsize.json (written by oracle/gen_fullsize.py from the pinned oracle in the build container), and checked
through size-independent properties (tests/properties.py): stream structure, invariance under a permutation of the
triangle list, and the volume bracket inside <= V <= inside + boundary at both levels on a copy in generic position."""
import hashlib
import json
import os

import numpy as np
import pytest

from properties import check_stream_structure, check_volume_bracket, epsilon_blind_fraction, rotated
from util import GOLD

pytestmark = pytest.mark.gpu

with open(os.path.join(GOLD, "fullsize.json")) as _f:
    FULL = json.load(_f)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _gold(case, tris=None, path=None):
    """The committed oracle hashes of a case -- or, when numpy built a different mesh on this host than in the build
    container (last-bit differences of sin / cos between SIMD paths), the same oracle run here on the mesh at hand."""
    from oracle.gen_fullsize import expected
    from oracle import oraclebind as O
    gold = FULL[case]
    have = _sha(O.OracleMesh(path).tris if path else tris)
    if have != gold["triangles_sha256"]:
        gold = expected(case, tris, path)
    return gold


def _streams(res, normals=False):
    """D2H of every stream of the last call on the context + their hashes (the views die with the next call)."""
    s = {"level1_inout": res.level1_inout(), "prefix": res.prefix(), "boundary_index": res.boundary_index(), "level2_inout": res.level2_inout()}
    if normals:
        s["level1_normal"], s["level2_normal"] = res.level1_normal(), res.level2_normal()
    return s, {k: _sha(v) for k, v in s.items()}


@pytest.mark.parametrize("case", ["sphere_512_8", "torus_512_8", "cad_1024_2"])
def test_fullsize_config_matches_oracle_hashes_and_properties(product, ctx, case):
    from oracle.gen_fullsize import fullsize_tris
    tris = fullsize_tris(FULL[case]["mesh"])
    gold = _gold(case, tris=tris)
    assert len(tris) == gold["triangles"]
    mesh = product.mesh_from_triangles(tris)
    prm = product.Params(gold["l1"], gold["l2"])
    res = ctx.voxelize(mesh, prm)
    assert list(res.num_div) == gold["num_div"] and res.counts == gold["counts"] and res.nb == gold["n_boundary"]
    s, h = _streams(res)
    assert h == gold["sha256"], case
    check_stream_structure(s["level1_inout"], s["prefix"], s["boundary_index"], s["level2_inout"], res.counts, res.n23)
    del s
    # the result is a function of the triangle SET: any order of the list gives the same bytes
    perm = np.random.default_rng(20240607).permutation(len(tris))
    res2 = ctx.voxelize(product.mesh_from_triangles(tris[perm]), product.Params(gold["l1"], gold["l2"], product.GPV_KEEP_LISTS))  # and with the list sorts
    assert res2.counts == gold["counts"]
    assert _streams(res2)[1] == gold["sha256"], case + " (permuted triangle list)"
    # volume bracket at both levels, on the body in generic position (counts only: no stream leaves the device) -- where the
    # reference's absolute |det| < 1e-6 ray rejection leaves the fill meaningful: not on the 10 M-triangle body (properties.py)
    rt = rotated(tris)
    if epsilon_blind_fraction(rt) < 2e-2:  # sphere 3e-3, torus 1e-3, CAD body 1e-1
        res3 = ctx.voxelize(product.mesh_from_triangles(rt), prm)
        check_volume_bracket(rt, res3.grid_size, res3.grid_size2, res3.n23, res3.counts)
    else:
        assert case == "cad_1024_2", "only the 10 M-triangle body is expected to sit below the reference's ray epsilon"


@pytest.mark.parametrize("i", range(4))
def test_dataset_block_matches_oracle_hashes(product, ctx, tmp_path, i):
    """config 5: a drilled-block .off written with %.9g, read by the product's loader, 64 + 4^3 with normals."""
    from gpview_b200 import meshgen
    from oracle.gen_fullsize import case_mesh
    _, p = case_mesh("block%d_64_4" % i, str(tmp_path))
    gold = _gold("block%d_64_4" % i, path=p)
    mesh = product.load_mesh(p)
    assert mesh.ntri == gold["triangles"]
    res = ctx.voxelize(mesh, product.Params(64, 4, product.GPV_NORMALS))
    assert list(res.num_div) == gold["num_div"] and res.counts == gold["counts"]
    s, h = _streams(res, normals=True)
    assert h == gold["sha256"]
    check_stream_structure(s["level1_inout"], s["prefix"], s["boundary_index"], s["level2_inout"], res.counts, res.n23)
    rt = rotated(mesh.tris)
    res3 = ctx.voxelize(product.mesh_from_triangles(rt), product.Params(64, 4))
    check_volume_bracket(rt, res3.grid_size, res3.grid_size2, res3.n23, res3.counts)


def test_plain_c_example_writes_the_reference_file_set(product, oracle, tmp_path_factory, tmp_path):  # last: written after this round's GPU budget was spent
    """tools/example_native.c (INTEGRATION.md section 2 as a C99 program: Level-1-only call to size the host buffers, then the full
    call with normals, then gpv_save): its six files equal the oracle's writer byte for byte."""
    import filecmp
    import subprocess
    from util import ROOT, mesh_path
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools"), "example_native"])
    path = mesh_path("torus", tmp_path_factory.getbasetemp())
    out, ref = tmp_path / "c", tmp_path / "ora"
    out.mkdir(); ref.mkdir()
    log = subprocess.run([os.path.join(ROOT, "tools", "example_native"), path, "32", "4", str(out)], capture_output=True, text=True)
    assert log.returncode == 0, log.stderr
    oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED, 4).save(-1, str(ref))
    names = sorted(os.listdir(ref))
    assert names == sorted(os.listdir(out)) and len(names) == 6
    for n in names:
        assert filecmp.cmp(ref / n, out / n, shallow=False), n
