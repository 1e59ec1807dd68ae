"""GPU: batched dataset generation (BASELINE.json configs[4], gpv_voxelize_batch) through the C ABI -- the replacement of the
reference's one-GLUT-session-per-model loop (main's argv loop, src/GPView.cpp:1642-1659, then key `t`).  Every file of every
model against the oracle's writer (== Object::SaveVoxelization, tests/test_oracle_ref.py), the restart rule, and the Level-2
host-buffer growth path."""
import filecmp
import os

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu

L1, L2 = 64, 4


@pytest.fixture(scope="module")
def block_files(tmp_path_factory):
    """36 drilled-block .off meshes (seeded, four tessellation densities: ~0.4 k - 3.4 k triangles) plus the four fixture meshes
    (.obj and .off), smallest boundary count first so that a one-thread run has to grow its Level-2 buffer more than once."""
    from gpview_b200 import meshgen as M
    d = tmp_path_factory.mktemp("batch_meshes")
    files = [os.path.join(ROOT, "tests", "golden", "meshes", n) for n in ("torus.off", "cad.obj", "sphere.obj", "block.off")]
    for i in range(36):
        ns, ng = [(24, 8), (48, 14), (96, 20), (160, 36)][i % 4]
        V, F = M.drilled_block(seed=M.SEED_BASE + 1000 + i, n_seg=ns, n_grid=ng)
        p = str(d / ("block%03d.off" % i))
        M.write_off(p, V, F)
        files.append(p)
    return files


def oracle_files(oracle, path, obj_id, out, normals=True):
    os.makedirs(out, exist_ok=True)
    oracle.OracleMesh(path).voxelize(L1, L2, oracle.FILL_CERTIFIED | (0 if normals else oracle.NO_NORMALS), 4).save(obj_id, out)


def compare_sets(ref, got, ids, skip_normals=False):
    for i in ids:
        names = sorted(n for n in os.listdir(ref) if n.startswith("Obj%d" % i) and n[len("Obj%d" % i)].isalpha())
        assert len(names) == 6, names
        for n in names:
            if skip_normals and "Normal" in n:
                assert not os.path.exists(os.path.join(got, n)), n
                continue
            assert filecmp.cmp(os.path.join(ref, n), os.path.join(got, n), shallow=False), n


def test_batch_files_equal_the_oracle_writer_and_restart_skips(product, oracle, block_files, tmp_path):
    from gpview_b200 import binding as B
    ndev = product.lib().gpv_device_count()
    assert ndev >= 1
    out, ref = str(tmp_path / "out"), str(tmp_path / "ref")
    os.makedirs(out)
    first = 100
    st = B.voxelize_batch(block_files, product.Params(L1, L2, product.GPV_NORMALS), list(range(ndev)), 4, out, first, False)
    assert st["models_done"] == len(block_files) and st["models_failed"] == 0 and st["models_skipped"] == 0
    assert st["level2_resizes"] >= 1   # every worker's first model outgrows the initial 1,024-cell Level-2 buffer
    for k, p in enumerate(block_files):
        oracle_files(oracle, p, first + k, ref)
    compare_sets(ref, out, range(first, first + len(block_files)))
    assert len(os.listdir(out)) == 6 * len(block_files)
    # restart: complete sets are not recomputed ...
    st2 = B.voxelize_batch(block_files, product.Params(L1, L2, product.GPV_NORMALS), list(range(ndev)), 4, out, first, True)
    assert st2["models_skipped"] == len(block_files) and st2["models_done"] == 0
    # ... a set with a missing or truncated stream is (its config alone does not vouch for it), and only that one
    os.remove(os.path.join(out, "Obj%dLevel2InOut.raw" % (first + 3)))
    with open(os.path.join(out, "Obj%dLevel1InOut.raw" % (first + 7)), "r+b") as f:
        f.truncate(100)
    os.remove(os.path.join(out, "Obj%dVoxelConfig.txt" % (first + 11)))
    before = {n: os.path.getmtime(os.path.join(out, n)) for n in os.listdir(out)}
    st3 = B.voxelize_batch(block_files, product.Params(L1, L2, product.GPV_NORMALS), list(range(ndev)), 3, out, first, True)
    assert st3["models_done"] == 3 and st3["models_skipped"] == len(block_files) - 3
    compare_sets(ref, out, range(first, first + len(block_files)))
    untouched = [n for n in before if not any(n.startswith("Obj%d" % (first + k)) and n[len("Obj%d" % (first + k))].isalpha() for k in (3, 7, 11))]
    assert all(os.path.getmtime(os.path.join(out, n)) == before[n] for n in untouched)


def test_batch_one_thread_grows_the_level2_buffer_and_lean_sets(product, oracle, block_files, tmp_path):
    """One worker, models ordered by boundary count: the Level-2 host buffer is outgrown repeatedly (gpv_batch.cpp: sized from a
    Level-1-only run, then the model is voxelized again) -- results must not depend on it.  GPV_SAVE_COMPUTED_ONLY: no normal
    files, the four computed files equal the oracle's."""
    from gpview_b200 import binding as B
    files = block_files[:4] + block_files[4:16]
    nb = []
    for p in files:
        nb.append(oracle.OracleMesh(p).voxelize(L1, L2, oracle.FILL_CERTIFIED | oracle.NO_L2 | oracle.NO_NORMALS, 4).nb)
    order = list(np.argsort(nb))
    files = [files[i] for i in order]
    out, ref = str(tmp_path / "out"), str(tmp_path / "ref")
    os.makedirs(out)
    st = B.voxelize_batch(files, product.Params(L1, L2, product.GPV_SAVE_COMPUTED_ONLY), [0], 1, out, -1, False)
    assert st["models_done"] == len(files) and st["models_failed"] == 0
    assert st["level2_resizes"] >= 2, st
    for k, p in enumerate(files):
        oracle_files(oracle, p, k - 1, ref)
    compare_sets(ref, out, range(-1, len(files) - 1), skip_normals=True)
    assert all(B.check_voxels(out, k - 1) for k in range(len(files)))


def test_batch_reports_a_bad_model_and_finishes_the_rest(product, block_files, tmp_path):
    from gpview_b200 import binding as B
    bad = str(tmp_path / "broken.off")
    with open(bad, "w") as f:
        f.write("OFF\n3 1 0\n0 0 0\n1 0 0\n0 1 x\n3 0 1 2\n")
    files = block_files[:3] + [bad, str(tmp_path / "missing.obj")] + block_files[3:6]
    out = str(tmp_path / "out")
    os.makedirs(out)
    with pytest.raises(product.GpvError) as e:
        B.voxelize_batch(files, product.Params(L1, L2, product.GPV_SAVE_COMPUTED_ONLY), [0], 2, out, 0, False)
    assert "broken.off" in str(e.value) or "missing.obj" in str(e.value)
    done = sorted(int(n[3:-len("VoxelConfig.txt")]) for n in os.listdir(out) if n.endswith("VoxelConfig.txt"))
    assert done == [0, 1, 2, 5, 6, 7]
