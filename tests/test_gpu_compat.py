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


@pytest.mark.parametrize("n,w", [(1, 1), (7, 7), (255, 5), (256, 16), (65537, 257), (1 << 22, 2048), (3 * 1000 * 1000 + 1, 1)])
def test_thrust_device_find_max_equals_numpy(product, n, w):
    """THRUSTDeviceFindMax (cuda/THRUSTUtilities.cu:44-61: thrust::max_element over w*h device floats + one-element read-back), the
    fourth device symbol the reference's host code links against -- hand-written reduction, no Thrust: against np.max on random
    data, all-negative data (the identity must not be 0), a maximum in the last element, and +inf."""
    L = product.lib()
    L.THRUSTDeviceFindMax.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.THRUSTDeviceFindMax.restype = C.c_float
    h = n // w
    n = w * h
    rng = np.random.default_rng(n)
    cases = [rng.standard_normal(n).astype(np.float32) * 1e3, -np.abs(rng.standard_normal(n)).astype(np.float32) - 1.0,
             np.zeros(n, np.float32), np.full(n, -3.0e38, np.float32)]
    cases[2][-1] = 5.5
    if n > 3:
        big = rng.standard_normal(n).astype(np.float32)
        big[n // 3] = np.inf
        cases.append(big)
    d = L.gpv_alloc_device(n * 4)
    assert d
    try:
        for a in cases:
            assert L.gpv_memcpy_h2d(d, a.ctypes.data, a.nbytes, None) == 0 and L.gpv_stream_sync(None) == 0
            got = L.THRUSTDeviceFindMax(d, w, h)
            assert np.float32(got) == a.max(), (n, got, a.max())
        assert L.THRUSTDeviceFindMax(d, 0, 5) == 0.0   # empty range: the reference dereferences end(); we return 0
    finally:
        L.gpv_free_device(d)
