"""The size-independent property checks (tests/properties.py) and the full-size hash file (tests/golden/fullsize.json) on the
CPU: the checkers must accept what the oracle produces, reject corrupted streams, and the committed hashes must be what the
oracle yields today.  (The GPU box applies the same checkers to the device streams at full size: tests/test_gpu_scale.py.)"""
import json
import os

import numpy as np
import pytest

from properties import check_stream_structure, check_volume_bracket, epsilon_blind_fraction, mesh_volume, rotated
from util import GOLD, mesh_path


@pytest.mark.parametrize("name,l1,l2", [("sphere", 32, 4), ("torus", 32, 4), ("block", 32, 4), ("cad", 32, 4), ("torus", 48, 2), ("cad", 100, 2)])
def test_checkers_accept_the_oracle(oracle, tmp_path_factory, name, l1, l2):
    t0 = oracle.OracleMesh(mesh_path(name, tmp_path_factory.getbasetemp())).tris
    for tris, generic in ((t0, False), (rotated(t0), True)):
        om = oracle.OracleMesh(tris=tris)
        r = om.voxelize(l1, l2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 4)
        nb = check_stream_structure(r.l1_state * 127, r.prefix, r.boundary_index, r.l2_state * 127, r.counts, r.n23)
        assert nb == r.nb
        if generic:  # axis-aligned bodies of revolution put rays exactly on mesh edges (see properties.rotated)
            out = check_volume_bracket(tris, r.grid_size, r.grid_size2, r.n23, r.counts)
            assert out["l1"][0] < out["l2"][0] <= out["l2"][1] < out["l1"][1]
    assert mesh_volume(rotated(t0)) == pytest.approx(mesh_volume(t0), rel=1e-5)


def test_checkers_reject_corrupted_streams(oracle, tmp_path_factory):
    om = oracle.OracleMesh(mesh_path("torus", tmp_path_factory.getbasetemp()))
    r = om.voxelize(32, 4, oracle.FILL_CERTIFIED | oracle.NO_NORMALS, 4)
    l1, l2 = (r.l1_state * 127).astype(np.uint8), (r.l2_state * 127).astype(np.uint8)
    good = dict(l1=l1, prefix=r.prefix, bidx=r.boundary_index, l2=l2, counts=r.counts, n23=r.n23, chunk=997)  # many chunk seams
    check_stream_structure(**good)

    def broken(**kw):
        with pytest.raises(AssertionError):
            check_stream_structure(**dict(good, **kw))
    b = int(r.boundary_index[len(r.boundary_index) // 2])
    x = l1.copy(); x[b] = 127; broken(l1=x)                                # a boundary cell lost
    x = l1.copy(); x[0] = 1; broken(l1=x)                                  # a state outside the encoding
    x = r.prefix.copy(); x[b + 1:] += 1; broken(prefix=x)                  # prefix off by one behind a cell
    x = r.boundary_index.copy(); x[[3, 4]] = x[[4, 3]]; broken(bidx=x)     # list not ascending
    x = l2.copy(); x[np.flatnonzero(x == 254)[0]] = 0; broken(l2=x)        # a Level-2 voxel lost
    broken(l2=l2[:-1])                                                     # a short Level-2 stream
    broken(counts=[r.counts[0] + 1] + r.counts[1:])
    with pytest.raises(AssertionError):                                    # half the solid missing
        check_volume_bracket(om.tris, r.grid_size, r.grid_size2, r.n23, [r.counts[0] // 2, r.counts[1], r.counts[2], r.counts[3]])


def test_fullsize_hash_file_is_what_the_oracle_says(oracle, tmp_path):
    """The small cases of tests/golden/fullsize.json recomputed here (the large ones take a minute: oracle/gen_fullsize.py)."""
    from oracle.gen_fullsize import CASES, case_mesh, expected
    with open(os.path.join(GOLD, "fullsize.json")) as f:
        full = json.load(f)
    assert sorted(full) == sorted(CASES)
    for case in ("block0_64_4", "block3_64_4"):
        _, p = case_mesh(case, str(tmp_path))
        now = expected(case, path=p)
        if now["triangles_sha256"] != full[case]["triangles_sha256"]:
            pytest.skip("numpy builds a different mesh on this host (last-bit sin/cos): the GPU test falls back to the live oracle")
        assert now == full[case]


def test_epsilon_blind_fraction():
    """Two triangles of projected area 1 and 4e-7 (|det| = 2 and 8e-7): the rays cannot see the second."""
    big = [0, 0, 0, 1, 0, 0, 0, 2, 1]
    small = [0, 0, 0, 1e-3, 0, 5, 0, 8e-4, 0]
    assert epsilon_blind_fraction(np.array([big], np.float32)) == 0.0
    assert epsilon_blind_fraction(np.array([small], np.float32)) == 1.0
    assert epsilon_blind_fraction(np.array([big, small], np.float32)) == pytest.approx(8e-7 / (2 + 8e-7), rel=1e-3)
