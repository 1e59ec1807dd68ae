"""ctypes binding of oracle/_ref/libgpvref.so (the unmodified reference compiled by oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/gen_golden.py and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libgpvref.so")


def available():
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, ci, cl, cd = C.c_void_p, C.c_int, C.c_long, C.c_double
        fp, ip, lp, bp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_ubyte)
        sig = {
            "ref_open": (vp, [C.c_char_p, ci, ci]), "ref_ntri": (ci, [vp]), "ref_tris": (fp, [vp]),
            "ref_bbox": (None, [vp, fp, fp, fp]), "ref_setup": (None, [vp, ci, ci]), "ref_grid": (None, [vp, ip, fp, fp]),
            "ref_l1_inout_brute": (cd, [vp]), "ref_l1_inout_brute_mt": (cd, [vp, ci]), "ref_l1_tribox": (cd, [vp]),
            "ref_compact": (None, [vp]), "ref_l1_inout_collist_mismatch": (cl, [vp]),
            "ref_l2_cpu": (None, [vp, C.POINTER(cd), C.POINTER(cd)]), "ref_l2_kernelform": (cd, [vp, ci]),
            "ref_l2_adopt_kernelform": (None, [vp]), "ref_count": (None, [vp, lp]), "ref_save": (None, [vp, C.c_char_p]),
            "ref_level1InOut": (fp, [vp]), "ref_level1Normal": (fp, [vp]), "ref_prefix": (ip, [vp]),
            "ref_nboundary": (cl, [vp]), "ref_boundaryIndex": (ip, [vp]), "ref_level2InOut": (fp, [vp]),
            "ref_level2Normal": (fp, [vp]), "ref_level2InOutKernel": (fp, [vp]), "ref_level2NormalKernel": (fp, [vp]),
            "ref_triCount": (ip, [vp]), "ref_triFlatIndex": (ip, [vp]), "ref_triFlat": (ip, [vp]), "ref_triFlatLen": (cl, [vp]),
            "ref_xyCount": (ip, [vp]), "ref_xyFlatIndex": (ip, [vp]), "ref_xyFlat": (ip, [vp]), "ref_xyFlatLen": (cl, [vp]),
            "ref_stats": (None, [vp, lp]), "ref_tribox_batch": (None, [cl, fp, fp, fp, bp]),
            "ref_triray_batch": (None, [cl, fp, fp, bp]),
            "ref_time_l2_tribox": (cd, [vp, cl, cl, ci, lp]), "ref_close": (None, [vp]),
            "ref_set_solid": (None, [vp, C.c_void_p]), "ref_build_hierarchy": (ci, [vp, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
            "ref_collision_boxes": (cl, [vp, C.c_void_p, C.c_void_p, C.c_void_p]),
            "ref_time_l1_fill_collist": (cd, [vp, ci, lp]), "ref_time_l2_path": (cd, [vp, cl, cl, ci, lp]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class RefObject:
    """The reference's Object, loaded from an .obj/.off file by the reference's own loader."""

    def __init__(self, path, obj_id=-1):
        L = lib()
        is_off = path.lower().endswith("off")
        self.h = L.ref_open(os.fsencode(path), int(is_off), obj_id)
        self.ntri = L.ref_ntri(self.h)
        self.tris = _arr(L.ref_tris(self.h), self.ntri * 9, np.float32).reshape(-1, 9)
        bmin, bmax, ms = np.zeros(3, np.float32), np.zeros(3, np.float32), C.c_float()
        L.ref_bbox(self.h, _fp(bmin), _fp(bmax), C.byref(ms))
        self.bmin, self.bmax, self.max_model_size = bmin, bmax, np.float32(ms.value)

    def setup(self, l1, l2):
        L = lib()
        L.ref_setup(self.h, l1, l2)
        nd, gs, gs2 = np.zeros(3, np.int32), np.zeros(3, np.float32), np.zeros(3, np.float32)
        L.ref_grid(self.h, nd.ctypes.data_as(C.POINTER(C.c_int)), _fp(gs), _fp(gs2))
        self.num_div, self.grid_size, self.grid_size2, self.n2 = nd, gs, gs2, max(l2, 1)
        self.cells = int(nd[0]) * int(nd[1]) * int(nd[2])

    def l1_inout_brute(self, threads=0):
        L = lib()
        return L.ref_l1_inout_brute(self.h) if threads <= 0 else L.ref_l1_inout_brute_mt(self.h, threads)

    def l1_tribox(self):
        return lib().ref_l1_tribox(self.h)

    def compact(self):
        lib().ref_compact(self.h)

    def collist_mismatch(self):
        return lib().ref_l1_inout_collist_mismatch(self.h)

    def l2_cpu(self):
        a, b = C.c_double(), C.c_double()
        lib().ref_l2_cpu(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def l2_kernelform(self, threads=8):
        return lib().ref_l2_kernelform(self.h, threads)

    def adopt_kernelform(self):
        lib().ref_l2_adopt_kernelform(self.h)

    def count(self):
        out = (C.c_long * 4)()
        lib().ref_count(self.h, out)
        return [int(x) for x in out]

    def save(self, d):
        lib().ref_save(self.h, os.fsencode(d))

    def stats(self):
        out = (C.c_long * 6)()
        lib().ref_stats(self.h, out)
        k = ["l1_box_tests", "l1_box_hits", "max_per_cell", "l2_box_tests", "l2_ray_tests", "l1_col_ray_tests"]
        return dict(zip(k, [int(x) for x in out]))

    def time_l2_tribox(self, b0, b1, threads):
        n = C.c_long()
        s = lib().ref_time_l2_tribox(self.h, b0, b1, threads, C.byref(n))
        return s, n.value

    def time_l1_fill_collist(self, threads):
        """(seconds, ray tests) of the Level-1 fill through the column lists on `threads` host threads (needs compact())."""
        n = C.c_long()
        s = lib().ref_time_l1_fill_collist(self.h, threads, C.byref(n))
        return s, n.value

    def time_l2_path(self, b0, b1, threads):
        """(seconds, tri-box tests, ray tests) of Level 2 -- parity rays over the column list, then SAT over the cell list, per
        sub-voxel -- over boundary cells [b0,b1) on `threads` host threads."""
        out = (C.c_long * 2)()
        s = lib().ref_time_l2_path(self.h, b0, b1, threads, out)
        return s, int(out[0]), int(out[1])

    def build_hierarchy(self, fill_only):
        """The reference's own Object::BuildHierarchy over its bBox[] (needs l1_tribox()); `fill_only` = the parity fill (bBox[].solid)."""
        f = np.ascontiguousarray(fill_only, np.uint8)
        lib().ref_set_solid(self.h, f.ctypes.data)
        n = self.cells - 1
        mid, half, solid, child = np.zeros(n * 3, np.float32), np.zeros(n * 3, np.float32), np.zeros(n, np.uint8), np.zeros(n * 2, np.int32)
        lv = lib().ref_build_hierarchy(self.h, mid.ctypes.data, half.ctypes.data, solid.ctypes.data, child.ctypes.data)
        return lv, mid.reshape(-1, 3), half.reshape(-1, 3), solid, child.reshape(-1, 2)

    def collision_boxes(self):
        inv, mid, ext = np.zeros(self.cells, np.int32), np.zeros(self.cells * 3, np.float32), np.zeros(self.cells * 3, np.float32)
        n = lib().ref_collision_boxes(self.h, inv.ctypes.data, mid.ctypes.data, ext.ctypes.data)
        return inv[:n].copy(), mid[:n * 3].reshape(-1, 3).copy(), ext[:n * 3].reshape(-1, 3).copy()

    # arrays
    def level1_inout(self): return _arr(lib().ref_level1InOut(self.h), self.cells, np.float32)
    def level1_normal(self): return _arr(lib().ref_level1Normal(self.h), self.cells * 3, np.float32)
    def prefix(self): return _arr(lib().ref_prefix(self.h), self.cells, np.int32)
    def nboundary(self): return int(lib().ref_nboundary(self.h))
    def boundary_index(self): return _arr(lib().ref_boundaryIndex(self.h), self.nboundary(), np.int32)
    def level2_inout(self): return _arr(lib().ref_level2InOut(self.h), self.nboundary() * self.n2 ** 3, np.float32)
    def level2_normal(self): return _arr(lib().ref_level2Normal(self.h), self.nboundary() * self.n2 ** 3 * 4, np.float32)
    def level2_inout_kernel(self): return _arr(lib().ref_level2InOutKernel(self.h), self.nboundary() * self.n2 ** 3, np.float32)
    def level2_normal_kernel(self): return _arr(lib().ref_level2NormalKernel(self.h), self.nboundary() * self.n2 ** 3 * 4, np.float32)
    def tri_count(self): return _arr(lib().ref_triCount(self.h), self.cells, np.int32)
    def tri_flat_index(self): return _arr(lib().ref_triFlatIndex(self.h), self.cells, np.int32)
    def tri_flat(self): return _arr(lib().ref_triFlat(self.h), lib().ref_triFlatLen(self.h), np.int32)
    def xy_count(self): return _arr(lib().ref_xyCount(self.h), int(self.num_div[0]) * int(self.num_div[1]), np.int32)
    def xy_flat_index(self): return _arr(lib().ref_xyFlatIndex(self.h), int(self.num_div[0]) * int(self.num_div[1]), np.int32)
    def xy_flat(self): return _arr(lib().ref_xyFlat(self.h), lib().ref_xyFlatLen(self.h), np.int32)

    def close(self):
        if self.h:
            lib().ref_close(self.h)
            self.h = None


def tribox_batch(c, h, tri9):
    c = np.ascontiguousarray(c, np.float32); h = np.ascontiguousarray(h, np.float32); tri9 = np.ascontiguousarray(tri9, np.float32)
    n = len(c); out = np.zeros(n, np.uint8)
    lib().ref_tribox_batch(n, _fp(c), _fp(h), _fp(tri9), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out


def triray_batch(org, tri9):
    org = np.ascontiguousarray(org, np.float32); tri9 = np.ascontiguousarray(tri9, np.float32)
    n = len(org); out = np.zeros(n, np.uint8)
    lib().ref_triray_batch(n, _fp(org), _fp(tri9), out.ctypes.data_as(C.POINTER(C.c_ubyte)))
    return out


def fnv1a64(b):
    """FNV-1a 64 of a bytes-like (the hash SURVEY.md 8(c) quotes)."""
    h = 1469598103934665603
    for x in memoryview(b).cast("B").tobytes():
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
