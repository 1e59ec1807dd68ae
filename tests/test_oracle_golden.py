"""The oracle (oracle/gpv_oracle.c) against the committed fixtures the UNMODIFIED reference produced (oracle/gen_golden.py):
every stream bit-for-bit -- occupancy, prefix sums, boundary indices, CSR lists, and the normals."""
import numpy as np
import pytest

from util import GOLDEN_CASES, golden, l1_bits, mesh_path, sha


@pytest.mark.parametrize("name,l1,l2", GOLDEN_CASES)
def test_oracle_matches_reference_fixture(oracle, tmp_path_factory, name, l1, l2):
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    m = oracle.OracleMesh(path)
    info, z = golden("%s_%d_%d" % (name, l1, l2))
    assert m.ntri == info["ntri"]
    assert [np.float32(x).tobytes().hex() for x in m.bmin] == info["bbox_min_hex"]
    assert [np.float32(x).tobytes().hex() for x in m.bmax] == info["bbox_max_hex"]
    r = m.voxelize(l1, l2, oracle.FILL_CERTIFIED, 8)
    assert list(r.num_div) == info["num_div"]
    assert [np.float32(x).tobytes().hex() for x in r.grid_size] == info["grid_size_hex"]
    assert r.counts == [info["l1_inside"], info["l1_boundary"], info["l2_inside"], info["l2_boundary"]]
    assert r.stats["l1BoxTests"] == info["l1_box_tests"] and r.stats["l1BoxHits"] == info["l1_box_hits"]
    assert r.stats["maxPerCell"] == info["max_per_cell"]
    s = info["streams"]
    assert sha(r.l1_state * 127) == s["Level1InOut"]["sha256"]
    assert sha(r.l1_fill_only * 127) == s["Level1FillOnly"]["sha256"]       # == Object::ClassifyInOutCPU brute force
    assert sha(r.prefix) == s["Level1BoundaryPrefixSum"]["sha256"]
    assert sha(r.boundary_index) == s["BoundaryIndex"]["sha256"]
    assert sha(r.l2_state * 127) == s["Level2InOut"]["sha256"]
    assert sha(r.l1_normal) == s["Level1Normal"]["sha256"]
    assert sha(r.l2_normal) == s["Level2Normal"]["sha256"]
    assert np.array_equal(l1_bits(r.l1_state * 127), z["l1_state_bits"])
    assert np.array_equal(r.cell_count[r.boundary_index], z["tri_count_boundary"])
    n23 = l2 ** 3
    assert np.array_equal((r.l2_state.reshape(-1, n23) == 1).sum(1), z["l2_inside_per_cell"])
    assert np.array_equal((r.l2_state.reshape(-1, n23) == 2).sum(1), z["l2_boundary_per_cell"])
    if "cell_lists" in z:
        assert np.array_equal(r.cell_tris, z["cell_lists"]) and np.array_equal(r.col_tris, z["col_lists"])
        assert np.array_equal(r.l2_state * 127, z["l2_state"])


@pytest.mark.parametrize("name,l1,l2", [c for c in GOLDEN_CASES if c[1] <= 64])
def test_fill_variants_agree(oracle, tmp_path_factory, name, l1, l2):
    """brute force (Object::ClassifyInOutCPU literally) == certified column culling == SAT column lists; naive L2 rays ==
    factorised L2 rays."""
    m = oracle.OracleMesh(mesh_path(name, tmp_path_factory.getbasetemp()))
    a = m.voxelize(l1, l2, oracle.FILL_BRUTE | oracle.L2_NAIVE, 8)
    b = m.voxelize(l1, l2, oracle.FILL_CERTIFIED, 8)
    c = m.voxelize(l1, l2, oracle.FILL_COLLIST, 8)
    for x in (b, c):
        assert np.array_equal(a.l1_fill_only, x.l1_fill_only)
        assert np.array_equal(a.l2_state, x.l2_state)
        assert np.array_equal(a.l2_normal, x.l2_normal)


def test_certified_fill_equals_brute_force_on_hostile_soups(oracle):
    """Triangle soups with slivers, near-vertical walls and large coordinates: the certified fill must reproduce the brute
    force parity exactly (noise-level hits of degenerate triangles included)."""
    rng = np.random.default_rng(5)
    for trial in range(6):
        n = 300
        scale = [1.0, 1.0, 50.0, 500.0, 1.0, 2000.0][trial]
        t = rng.uniform(-1, 1, (n, 3, 3)) * scale
        k = n // 4
        t[:k, 2] = t[:k, 0] + (t[:k, 1] - t[:k, 0]) * rng.uniform(-0.5, 1.5, (k, 1)) + rng.normal(0, 1e-6 * scale, (k, 3))   # slivers
        t[k:2 * k, 2, :2] = t[k:2 * k, 0, :2] + rng.normal(0, 1e-7 * scale, (k, 2))                                         # vertical walls
        t[2 * k:3 * k] = t[2 * k:3 * k, :1] + rng.normal(0, 0.02 * scale, (k, 3, 3))                                         # small
        m = oracle.OracleMesh(tris=t.reshape(n, 9).astype(np.float32))
        a = m.voxelize(24, 2, oracle.FILL_BRUTE | oracle.NO_NORMALS, 8)
        b = m.voxelize(24, 2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 8)
        assert np.array_equal(a.l1_fill_only, b.l1_fill_only), trial
        assert np.array_equal(a.l2_state, b.l2_state), trial
