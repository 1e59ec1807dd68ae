"""ctypes binding of oracle/libgpv_oracle.so (plain-C restatement of the reference voxelizer path).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpv_oracle.so")

FILL_BRUTE, FILL_CERTIFIED, FILL_COLLIST, L2_NAIVE, NO_NORMALS, NO_L2 = 0, 1, 2, 16, 32, 64


class Mesh(C.Structure):
    _fields_ = [("nTri", C.c_int64), ("tris", C.POINTER(C.c_float)), ("bmin", C.c_float * 3), ("bmax", C.c_float * 3),
                ("maxModelSize", C.c_float), ("nVerts", C.c_int64)]


class Grid(C.Structure):
    _fields_ = [("numDiv", C.c_int * 3), ("gridSize", C.c_float * 3), ("gridSize2", C.c_float * 3), ("ext1", C.c_float * 3),
                ("ext2", C.c_float * 3), ("n2", C.c_int)]


class Result(C.Structure):
    _fields_ = [("g", Grid), ("cells", C.c_int64), ("nBoundary", C.c_int64), ("n23", C.c_int64),
                ("l1State", C.POINTER(C.c_uint8)), ("l1FillOnly", C.POINTER(C.c_uint8)), ("prefix", C.POINTER(C.c_int32)),
                ("boundaryIndex", C.POINTER(C.c_int32)), ("cellCount", C.POINTER(C.c_int32)), ("cellOffset", C.POINTER(C.c_int64)),
                ("cellTris", C.POINTER(C.c_int32)), ("colCount", C.POINTER(C.c_int32)), ("colOffset", C.POINTER(C.c_int64)),
                ("colTris", C.POINTER(C.c_int32)), ("l1Normal", C.POINTER(C.c_uint8)), ("l2State", C.POINTER(C.c_uint8)),
                ("l2Normal", C.POINTER(C.c_uint8))] + [(k, C.c_int64) for k in (
                    "l1Inside", "l1Boundary", "l2Inside", "l2Boundary", "l1BoxTests", "l1BoxHits", "maxPerCell", "l2BoxTests",
                    "l2RayTests", "l1ColRayTests", "fillIllConditioned", "fillCrossings")]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        fp, bp = C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.gpvo_load_obj.argtypes = [C.c_char_p, C.POINTER(Mesh)]
        L.gpvo_load_off.argtypes = [C.c_char_p, C.POINTER(Mesh)]
        L.gpvo_mesh_from_tris.argtypes = [fp, C.c_int64, C.POINTER(Mesh)]; L.gpvo_mesh_from_tris.restype = None
        L.gpvo_free_mesh.argtypes = [C.POINTER(Mesh)]; L.gpvo_free_mesh.restype = None
        L.gpvo_make_grid.argtypes = [fp, fp, C.c_float, C.c_int, C.c_int, C.POINTER(Grid)]; L.gpvo_make_grid.restype = None
        L.gpvo_axis_table.argtypes = [fp, C.POINTER(Grid), C.c_int, fp]; L.gpvo_axis_table.restype = None
        L.gpvo_tribox_batch.argtypes = [C.c_int64, fp, fp, fp, bp]; L.gpvo_tribox_batch.restype = None
        L.gpvo_triray_batch.argtypes = [C.c_int64, fp, fp, bp]; L.gpvo_triray_batch.restype = None
        L.gpvo_triray_z_batch.argtypes = [C.c_int64, fp, fp, bp]; L.gpvo_triray_z_batch.restype = None
        L.gpvo_voxelize.argtypes = [C.POINTER(Mesh), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Result)]
        L.gpvo_free_result.argtypes = [C.POINTER(Result)]; L.gpvo_free_result.restype = None
        L.gpvo_save.argtypes = [C.POINTER(Mesh), C.POINTER(Result), C.c_int, C.c_char_p]
        L.gpvo_time_l2_tribox.argtypes = [C.POINTER(Mesh), C.POINTER(Result), C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_int64)]
        L.gpvo_time_l2_tribox.restype = C.c_double
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _np(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


class OracleMesh:
    def __init__(self, path=None, tris=None):
        self.m = Mesh()
        L = lib()
        if path is not None:
            fn = L.gpvo_load_off if path.lower().endswith("off") else L.gpvo_load_obj
            rc = fn(os.fsencode(path), C.byref(self.m))
            if rc:
                raise RuntimeError("oracle loader failed (%d) on %s" % (rc, path))
        else:
            t = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
            L.gpvo_mesh_from_tris(_fp(t), len(t), C.byref(self.m))
        self.ntri = int(self.m.nTri)
        self.tris = _np(self.m.tris, self.ntri * 9, np.float32).reshape(-1, 9)
        self.bmin = np.array(list(self.m.bmin), np.float32)
        self.bmax = np.array(list(self.m.bmax), np.float32)
        self.max_model_size = np.float32(self.m.maxModelSize)

    def set_bbox(self, bmin, bmax, max_model_size):
        for a in range(3):
            self.m.bmin[a] = float(bmin[a]); self.m.bmax[a] = float(bmax[a])
        self.m.maxModelSize = float(max_model_size)
        self.bmin = np.array(list(self.m.bmin), np.float32); self.bmax = np.array(list(self.m.bmax), np.float32)
        self.max_model_size = np.float32(self.m.maxModelSize)

    def voxelize(self, l1, l2, flags=FILL_CERTIFIED, threads=8):
        return OracleResult(self, l1, l2, flags, threads)

    def __del__(self):
        try:
            lib().gpvo_free_mesh(C.byref(self.m))
        except Exception:
            pass


class OracleResult:
    def __init__(self, mesh, l1, l2, flags, threads):
        self.mesh = mesh
        self.r = Result()
        rc = lib().gpvo_voxelize(C.byref(mesh.m), l1, l2, flags, threads, C.byref(self.r))
        if rc:
            raise RuntimeError("gpvo_voxelize failed %d" % rc)
        r = self.r
        self.num_div = np.array(list(r.g.numDiv), np.int32)
        self.grid_size = np.array(list(r.g.gridSize), np.float32)
        self.grid_size2 = np.array(list(r.g.gridSize2), np.float32)
        self.ext1 = np.array(list(r.g.ext1), np.float32)
        self.ext2 = np.array(list(r.g.ext2), np.float32)
        self.n2, self.cells, self.nb, self.n23 = int(r.g.n2), int(r.cells), int(r.nBoundary), int(r.n23)
        ncol = int(self.num_div[0]) * int(self.num_div[1])
        self.l1_state = _np(r.l1State, self.cells, np.uint8)
        self.l1_fill_only = _np(r.l1FillOnly, self.cells, np.uint8)
        self.prefix = _np(r.prefix, self.cells, np.int32)
        self.boundary_index = _np(r.boundaryIndex, self.nb, np.int32)
        self.cell_count = _np(r.cellCount, self.cells, np.int32)
        self.cell_offset = _np(r.cellOffset, self.cells + 1, np.int64)
        self.cell_tris = _np(r.cellTris, self.cell_offset[-1], np.int32)
        self.col_count = _np(r.colCount, ncol, np.int32)
        self.col_offset = _np(r.colOffset, ncol + 1, np.int64)
        self.col_tris = _np(r.colTris, self.col_offset[-1], np.int32)
        self.l1_normal = _np(r.l1Normal, self.cells * 3, np.uint8)
        self.l2_state = _np(r.l2State, self.nb * self.n23, np.uint8)
        self.l2_normal = _np(r.l2Normal, self.nb * self.n23 * 3, np.uint8)
        self.counts = [int(r.l1Inside), int(r.l1Boundary), int(r.l2Inside), int(r.l2Boundary)]
        self.stats = {k: int(getattr(r, k)) for k in ("l1BoxTests", "l1BoxHits", "maxPerCell", "l2BoxTests", "l2RayTests",
                                                     "l1ColRayTests", "fillIllConditioned", "fillCrossings")}

    def save(self, obj_id, d):
        rc = lib().gpvo_save(C.byref(self.mesh.m), C.byref(self.r), obj_id, os.fsencode(d))
        if rc:
            raise RuntimeError("gpvo_save failed")

    def time_l2_tribox(self, b0, b1, threads):
        n = C.c_int64()
        s = lib().gpvo_time_l2_tribox(C.byref(self.mesh.m), C.byref(self.r), b0, b1, threads, C.byref(n))
        return s, int(n.value)

    def collision_boxes(self):
        """Object::CollisionInitCUDA's arrays (src/Object.cpp:3530-3572): (inv_index, centre[n,3], extent[n,3]) of the occupied cells."""
        L = lib()
        L.gpvo_collision_boxes.argtypes = [C.POINTER(Mesh), C.POINTER(Result), C.c_void_p, C.c_void_p, C.c_void_p]
        L.gpvo_collision_boxes.restype = C.c_int64
        inv, mid, ext = np.zeros(self.cells, np.int32), np.zeros(self.cells * 3, np.float32), np.zeros(self.cells * 3, np.float32)
        n = L.gpvo_collision_boxes(C.byref(self.mesh.m), C.byref(self.r), inv.ctypes.data, mid.ctypes.data, ext.ctypes.data)
        return inv[:n].copy(), mid[:n * 3].reshape(-1, 3).copy(), ext[:n * 3].reshape(-1, 3).copy()

    def build_hierarchy(self):
        """Object::BuildHierarchy (src/Object.cpp:2790-2867): (levels, mid[n,3], half[n,3], solid[n], child[n,2]) with n = cells - 1;
        None for a grid whose dimensions are not all powers of two (the reference's loop is not defined there)."""
        L = lib()
        L.gpvo_build_hierarchy.argtypes = [C.POINTER(Mesh), C.POINTER(Result), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        n = self.cells - 1
        mid, half, solid, child = np.zeros(max(n, 1) * 3, np.float32), np.zeros(max(n, 1) * 3, np.float32), np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1) * 2, np.int32)
        lv = L.gpvo_build_hierarchy(C.byref(self.mesh.m), C.byref(self.r), mid.ctypes.data, half.ctypes.data, solid.ctypes.data, child.ctypes.data)
        if lv < 0:
            return None
        return lv, mid.reshape(-1, 3), half.reshape(-1, 3), solid, child.reshape(-1, 2)

    def __del__(self):
        try:
            lib().gpvo_free_result(C.byref(self.r))
        except Exception:
            pass


def make_grid(bmin, bmax, max_model_size, l1, l2):
    g = Grid()
    bmin = np.ascontiguousarray(bmin, np.float32); bmax = np.ascontiguousarray(bmax, np.float32)
    lib().gpvo_make_grid(_fp(bmin), _fp(bmax), float(max_model_size), l1, l2, C.byref(g))
    return g


def axis_table(bmin, g, axis):
    bmin = np.ascontiguousarray(bmin, np.float32)
    out = np.zeros(g.numDiv[axis], np.float32)
    lib().gpvo_axis_table(_fp(bmin), C.byref(g), axis, _fp(out))
    return out


def tribox_batch(c, h, tri9):
    c = np.ascontiguousarray(c, np.float32); h = np.ascontiguousarray(h, np.float32); tri9 = np.ascontiguousarray(tri9, np.float32)
    out = np.zeros(len(c), np.uint8)
    lib().gpvo_tribox_batch(len(c), _fp(c), _fp(h), _fp(tri9), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out


def triray_batch(org, tri9, specialised=False):
    org = np.ascontiguousarray(org, np.float32); tri9 = np.ascontiguousarray(tri9, np.float32)
    out = np.zeros(len(org), np.uint8)
    fn = lib().gpvo_triray_z_batch if specialised else lib().gpvo_triray_batch
    fn(len(org), _fp(org), _fp(tri9), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out
