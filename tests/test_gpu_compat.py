"""Drop-in boundary (SURVEY.md 8b tier 1): the reference's UNMODIFIED Object.cpp host code -- Object::ClassifyTessellationCUDA
and Object::ClassifyInOutTessellationLevel2CUDA -- linked against libgpview_b200.so (oracle/_ref/libgpvref_b200.so), i.e. the
three extern "C" operators resolved by the product; and the product's compat kernels against the reference's own CUDA
kernels compiled strict-IEEE (oracle/_ref/libgpvref_cuda.so).  Needs the prebuilt oracle/_ref (built in the build container)."""
import ctypes as C
import os

import numpy as np
import pytest

from util import ROOT, mesh_path

pytestmark = pytest.mark.gpu
REFB200 = os.path.join(ROOT, "oracle", "_ref", "libgpvref_b200.so")


@pytest.mark.skipif(not os.path.exists(REFB200), reason="oracle/_ref/libgpvref_b200.so not built")
@pytest.mark.parametrize("name,l1,l2", [("cessna", 32, 4), ("torus", 24, 8), ("cessna", 64, 4)])
def test_reference_host_code_runs_on_the_product_operators(product, oracle, tmp_path_factory, name, l1, l2):
    from oracle import refbind
    refbind.LIB_PATH = REFB200
    refbind._lib = None
    L = refbind.lib()
    L.ref_cuda_path.argtypes = [C.c_void_p]; L.ref_cuda_path.restype = C.c_int
    path = mesh_path(name, tmp_path_factory.getbasetemp())
    om = oracle.OracleMesh(path)
    want = om.voxelize(l1, l2, oracle.FILL_CERTIFIED, 4)
    ro = refbind.RefObject(path)
    ro.setup(l1, l2)
    # Level-1 fill: the reference would use its GL path here; seed the buffer with the oracle's parity fill
    fill = want.l1_fill_only.astype(np.float32)
    C.memmove(L.ref_level1InOut(ro.h), fill.ctypes.data, fill.nbytes)
    used = L.ref_cuda_path(ro.h)
    assert used >= 50
    if name == "cessna" and l1 == 64:
        assert used == 329           # max triangles per cell > the default buffer of 50: the reference's two-pass re-run (src/Object.cpp:3204-3214)
    assert np.array_equal(ro.level1_inout().astype(np.uint8), want.l1_state)
    assert np.array_equal(ro.boundary_index(), want.boundary_index)
    got2 = ro.level2_inout()
    assert np.array_equal(got2.astype(np.uint8), want.l2_state)
    # Level-2 normals after the reference's host averaging (src/Object.cpp:2613-2632), in its uchar encoding.
    # The reference's per-cell list order is the atomic slot order, ours canonical: sums may differ in the last bit.
    n = ro.level2_normal().reshape(-1, 4)[:, :3].reshape(-1)
    enc = (n * np.float32(256.0 / 3.0) + np.float32(127.0)).astype(np.uint8)
    assert np.abs(enc.astype(int) - want.l2_normal.astype(int)).max() <= 1
    assert ro.count() == want.counts
    ro.close()
    refbind.LIB_PATH = os.path.join(ROOT, "oracle", "_ref", "libgpvref.so")
    refbind._lib = None
