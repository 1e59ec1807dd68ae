"""GPU: voxel hierarchy / collision structures over the Level-1 grid -- gpv_collision_boxes (Object::CollisionInitCUDA,
src/Object.cpp:3530-3572) and gpv_build_hierarchy (Object::BuildHierarchy + CombineBBox, :2750-2867) through the C ABI, bit for bit
against the oracle's restatement (which tests/test_oracle_ref.py pins to the reference's own code)."""
import numpy as np
import pytest

from util import mesh_path

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,l1,l2", [("cessna", 64, 4), ("torus", 32, 0), ("sphere", 32, 4), ("cad", 48, 2), ("cessna", 256, 0)])
def test_collision_boxes_equal_the_oracle(product, oracle, ctx, tmp_path_factory, name, l1, l2):
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    res = ctx.voxelize(mesh, product.Params(l1, max(l2, 1), 0 if l2 else product.GPV_NO_LEVEL2))
    inv, mid, ext = ctx.collision_boxes()
    want = oracle.OracleMesh(path).voxelize(l1, l2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS | oracle.NO_L2, 4)
    winv, wmid, wext = want.collision_boxes()
    assert len(inv) == res.counts[0] + res.counts[1] == len(winv)
    assert np.array_equal(inv, winv) and np.array_equal(mid, wmid) and np.array_equal(ext, wext)
    # a z-slab call: the slab's occupied cells, global indices
    nz = int(res.num_div[2])
    z0, z1 = nz // 4, nz // 4 + max(1, nz // 3)
    ctx.voxelize(mesh, product.Params(l1, max(l2, 1), product.GPV_NO_LEVEL2, z0, z1))
    sinv, smid, sext = ctx.collision_boxes()
    plane = int(res.num_div[0]) * int(res.num_div[1])
    sel = (winv >= z0 * plane) & (winv < z1 * plane)
    assert np.array_equal(sinv, winv[sel]) and np.array_equal(smid, wmid[sel]) and np.array_equal(sext, wext[sel])


@pytest.mark.parametrize("name,l1", [("sphere", 16), ("sphere", 32), ("sphere", 64), ("sphere", 128)])
def test_hierarchy_equals_the_oracle(product, oracle, ctx, tmp_path_factory, name, l1):
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    mesh = product.load_mesh(path)
    res = ctx.voxelize(mesh, product.Params(l1, 2, product.GPV_COLLISION))
    assert all(int(n) & (int(n) - 1) == 0 for n in res.num_div)
    lv, mid, half, solid, child = ctx.build_hierarchy()
    want = oracle.OracleMesh(path).voxelize(l1, 2, oracle.FILL_CERTIFIED | oracle.NO_NORMALS | oracle.NO_L2, 4)
    wlv, wmid, whalf, wsolid, wchild = want.build_hierarchy()
    assert lv == wlv and len(solid) == res.cells - 1
    assert np.array_equal(solid, wsolid) and np.array_equal(child, wchild)
    assert np.array_equal(mid, wmid) and np.array_equal(half, whalf)
    # the occupancy streams of a GPV_COLLISION call are the plain call's
    plain = ctx.voxelize(mesh, product.Params(l1, 2, 0))
    assert plain.counts == res.counts


def test_hierarchy_refuses_what_the_reference_cannot_do(product, ctx, tmp_path_factory):
    mesh = product.load_mesh(mesh_path("cessna", tmp_path_factory.getbasetemp()))
    ctx.voxelize(mesh, product.Params(64, 4, product.GPV_COLLISION))          # 60 x 16 x 64: not powers of two
    with pytest.raises(product.GpvError) as e:
        ctx.build_hierarchy()
    assert "power of two" in str(e.value)
    sphere = product.load_mesh(mesh_path("sphere", tmp_path_factory.getbasetemp()))
    ctx.voxelize(sphere, product.Params(32, 4, 0))                            # no GPV_COLLISION: the parity of the boundary cells was not kept
    with pytest.raises(product.GpvError):
        ctx.build_hierarchy()
    ctx.voxelize(sphere, product.Params(32, 4, product.GPV_COLLISION, 4, 12))  # a slab
    with pytest.raises(product.GpvError):
        ctx.build_hierarchy()


def test_cli_facade_reports_the_oracle_hierarchy(product, oracle, tmp_path_factory, tmp_path):
    """The Object-shaped C++ facade (include/gpview_b200.hpp: Object::CollisionInitCUDA / BuildHierarchy, same member names as the
    reference) through the headless CLI: counts, levels and the root box the oracle computes; files still the oracle writer's."""
    import filecmp, os, re, subprocess
    from util import ROOT
    exe = os.path.join(ROOT, "tools", "gpview_voxelize")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])
    path = mesh_path("sphere", tmp_path_factory.getbasetemp())
    out = tmp_path / "cli"
    out.mkdir()
    log = subprocess.run([exe, "--l1", "32", "--l2", "4", "--collision", "--hierarchy", "--out", str(out), path], capture_output=True, text=True)
    assert log.returncode == 0, log.stderr
    want = oracle.OracleMesh(path).voxelize(32, 4, oracle.FILL_CERTIFIED, 4)
    lv, mid, half, solid, child = want.build_hierarchy()
    field = lambda name: re.search(name + r"\s*: (.*)", log.stdout).group(1)
    assert int(field("Collision Boxes")) == want.counts[0] + want.counts[1]
    assert int(field("Hierarchy Levels")) == lv and int(field("Hierarchy Boxes")) == want.cells - 1 and int(field("Solid Boxes")) == int(solid.sum())
    root = [np.float32(x) for x in field("Root Box").replace("+-", " ").split()]
    assert np.array_equal(np.array(root[:3], np.float32), mid[-1]) and np.array_equal(np.array(root[3:], np.float32), half[-1])
    ref = tmp_path / "ora"
    ref.mkdir()
    want.save(-1, str(ref))
    for n in sorted(os.listdir(ref)):
        assert filecmp.cmp(ref / n, out / n, shallow=False), n      # (the facade moves Level 2 as 2 bits per sub-voxel by default)
    bad = subprocess.run([exe, "--l1", "64", "--hierarchy", mesh_path("cessna", tmp_path_factory.getbasetemp())], capture_output=True, text=True, cwd=str(tmp_path))
    assert bad.returncode == 1 and "power of two" in bad.stderr
