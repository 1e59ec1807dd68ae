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
    assert res.stats["l2_box_tests"] == ores.stats["l2BoxTests"]
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
