"""ctypes binding of libgpview_b200.so (include/gpview_b200.h).  Harness only -- see gpview_b200/__init__.py."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPVIEW_B200_LIB", os.path.join(HERE, "libgpview_b200.so"))  # override: A/B builds of the same ABI

GPV_NORMALS, GPV_NO_LEVEL2, GPV_KEEP_LISTS, GPV_PROFILE, GPV_GATHER, GPV_SAVE_COMPUTED_ONLY, GPV_BATCH_TOLERANT_LOAD, GPV_PROFILE_L2, GPV_PACKED_L2, GPV_COLLISION = 1, 2, 4, 8, 16, 32, 64, 128, 256, 512


class GpvError(RuntimeError):
    pass


class CMesh(C.Structure):
    _fields_ = [("n_tri", C.c_int64), ("tris", C.POINTER(C.c_float)), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3),
                ("max_model_size", C.c_float), ("n_verts", C.c_int64)]


class CGrid(C.Structure):
    _fields_ = [("num_div", C.c_int * 3), ("grid_size", C.c_float * 3), ("grid_size2", C.c_float * 3), ("ext1", C.c_float * 3),
                ("ext2", C.c_float * 3), ("n2", C.c_int)]


class CParams(C.Structure):
    _fields_ = [("voxel_count", C.c_int), ("voxel_count2", C.c_int), ("flags", C.c_int), ("z0", C.c_int), ("z1", C.c_int)]


class CResult(C.Structure):
    _fields_ = [("grid", CGrid), ("z0", C.c_int), ("z1", C.c_int), ("cells", C.c_int64), ("n_boundary", C.c_int64), ("n23", C.c_int64),
                ("d_level1_inout", C.c_void_p), ("d_prefix", C.c_void_p), ("d_boundary_index", C.c_void_p), ("d_level2_inout", C.c_void_p),
                ("d_level1_normal", C.c_void_p), ("d_level2_normal", C.c_void_p), ("d_cell_off", C.c_void_p), ("d_cell_tris", C.c_void_p),
                ("d_col_off", C.c_void_p), ("d_col_count", C.c_void_p), ("d_col_tris", C.c_void_p)] + [(k, C.c_int64) for k in (
                    "l1_inside", "l1_boundary", "l2_inside", "l2_boundary", "l1_box_tests", "l1_box_hits", "l2_box_tests", "l2_ray_tests",
                    "tri_total", "fill_crossings", "fill_ill_conditioned", "kernel_launches")] + [("phase_ms", C.c_float * 16), ("n_refined", C.c_int64)]


PHASES = ["setup", "bin_count", "cross_count", "scan", "host_gap", "bin_fill", "cross_fill", "sort", "fill_sweep", "l1_normals", "l2_rays", "l2", "l2_normals"]


class CHostStreams(C.Structure):
    _fields_ = [("level1_inout", C.c_void_p), ("prefix", C.c_void_p), ("boundary_index", C.c_void_p), ("level2_inout", C.c_void_p),
                ("level1_normal", C.c_void_p), ("level2_normal", C.c_void_p), ("level2_capacity", C.c_int64), ("boundary_capacity", C.c_int64)]


class CVoxelFile(C.Structure):
    _fields_ = [("name", C.c_char * 64), ("bbox_min", C.c_float * 3), ("bbox_max", C.c_float * 3), ("num_div", C.c_int * 3), ("grid_size", C.c_float * 3),
                ("l1_inside", C.c_int64), ("l1_boundary", C.c_int64), ("has_level2", C.c_int), ("num_div2", C.c_int * 3), ("grid_size2", C.c_float * 3),
                ("l2_inside", C.c_int64), ("l2_boundary", C.c_int64), ("cells", C.c_int64), ("n_boundary", C.c_int64), ("n23", C.c_int64),
                ("level1_inout", C.POINTER(C.c_uint8)), ("level1_normal", C.POINTER(C.c_uint8)), ("prefix_sum", C.POINTER(C.c_int32)),
                ("level2_inout", C.POINTER(C.c_uint8)), ("level2_normal", C.POINTER(C.c_uint8))]


class CGatherDesc(C.Structure):
    _fields_ = [("l1", C.c_ubyte * 64), ("prefix", C.c_ubyte * 64), ("l2", C.c_ubyte * 64), ("mailbox", C.c_ubyte * 64), ("cells_total", C.c_int64),
                ("l2_capacity", C.c_int64), ("owner_device", C.c_int32), ("reserved", C.c_int32)]


class CCollision(C.Structure):
    _fields_ = [("count", C.c_int64), ("index_base", C.c_int64), ("d_inv_index", C.c_void_p), ("d_center", C.c_void_p), ("d_extent", C.c_void_p)]


class CHierarchy(C.Structure):
    _fields_ = [("num_levels", C.c_int), ("n_boxes", C.c_int64), ("d_mid", C.c_void_p), ("d_half", C.c_void_p), ("d_solid", C.c_void_p), ("d_child", C.c_void_p)]


class CBatchStats(C.Structure):
    _fields_ = [("models_done", C.c_int64), ("models_failed", C.c_int64), ("models_skipped", C.c_int64), ("seconds", C.c_double),
                ("parse_seconds", C.c_double), ("gpu_seconds", C.c_double), ("save_seconds", C.c_double), ("level2_resizes", C.c_int64)]


# every symbol include/gpview_b200.h declares (tests/test_abi_symbols.py checks the header against this and the .so)
NATIVE_SYMBOLS = ["gpv_last_error", "gpv_device_count", "gpv_create", "gpv_destroy", "gpv_stream", "gpv_load_obj", "gpv_load_off", "gpv_load_mesh", "gpv_load_mesh_ex",
                  "gpv_mesh_from_triangles", "gpv_free_mesh", "gpv_make_grid", "gpv_alloc_host", "gpv_free_host", "gpv_alloc_device",
                  "gpv_free_device", "gpv_memcpy_h2d", "gpv_memcpy_d2h", "gpv_stream_sync", "gpv_voxelize_device", "gpv_voxelize_host",
                  "gpv_save", "gpv_save_streams", "gpv_load_voxels", "gpv_check_voxels", "gpv_free_voxels", "gpv_expand_dense", "gpv_expand_packed_l2", "gpv_collision_boxes", "gpv_build_hierarchy", "gpv_voxelize_batch", "gpv_batch_release", "gpv_measure_fp32_peak", "gpv_measure_copy_peak",
                  "gpv_gather_create", "gpv_gather_create_ex", "gpv_gather_normals", "gpv_gather_attach", "gpv_gather_attach_local", "gpv_gather_detach", "gpv_gather_set_timeout", "gpv_gather_result"]
COMPAT_SYMBOLS = ["CUDAClassifyTessellation", "CUDAClassifyTessellationLevel2", "CUDAClassifyInOutLevel2", "THRUSTDeviceFindMax"]

_lib = None


def build(verbose=False):
    """Compile libgpview_b200.so in-tree (nvcc, sm_100a, -fmad=false).  Cross-compiles without a GPU."""
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc")] + ([] if verbose else ["-s"]))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GpvError("libgpview_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "there is no fallback implementation")
        L = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        vp = C.c_void_p
        L.gpv_last_error.restype = C.c_char_p
        L.gpv_device_count.restype = C.c_int
        L.gpv_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.gpv_destroy.argtypes = [vp]; L.gpv_destroy.restype = None
        L.gpv_stream.argtypes = [vp]; L.gpv_stream.restype = vp
        for n in ("gpv_load_obj", "gpv_load_off", "gpv_load_mesh"):
            getattr(L, n).argtypes = [C.c_char_p, C.POINTER(CMesh)]
        L.gpv_mesh_from_triangles.argtypes = [fp, C.c_int64, C.POINTER(CMesh)]
        L.gpv_free_mesh.argtypes = [C.POINTER(CMesh)]; L.gpv_free_mesh.restype = None
        L.gpv_make_grid.argtypes = [fp, fp, C.c_float, C.c_int, C.c_int, C.POINTER(CGrid)]
        L.gpv_alloc_host.argtypes = [C.c_int64]; L.gpv_alloc_host.restype = vp
        L.gpv_free_host.argtypes = [vp]; L.gpv_free_host.restype = None
        L.gpv_alloc_device.argtypes = [C.c_int64]; L.gpv_alloc_device.restype = vp
        L.gpv_free_device.argtypes = [vp]; L.gpv_free_device.restype = None
        L.gpv_memcpy_h2d.argtypes = [vp, vp, C.c_int64, vp]
        L.gpv_memcpy_d2h.argtypes = [vp, vp, C.c_int64, vp]
        L.gpv_stream_sync.argtypes = [vp]
        L.gpv_voxelize_device.argtypes = [vp, vp, C.c_int64, fp, fp, C.c_float, C.POINTER(CParams), vp, C.POINTER(CResult)]
        L.gpv_voxelize_host.argtypes = [vp, C.POINTER(CMesh), C.POINTER(CParams), vp, C.POINTER(CResult), C.POINTER(CHostStreams)]
        L.gpv_save.argtypes = [C.POINTER(CMesh), C.POINTER(CResult), C.POINTER(CHostStreams), C.c_int, C.c_char_p]
        L.gpv_save_streams.argtypes = [C.POINTER(CMesh), C.POINTER(CResult), C.POINTER(CHostStreams), C.c_int, C.c_char_p, C.c_int]
        L.gpv_voxelize_batch.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.POINTER(CParams), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                         C.POINTER(CBatchStats)]
        L.gpv_expand_packed_l2.argtypes = [vp, C.c_int64, vp]
        L.gpv_collision_boxes.argtypes = [vp, vp, C.POINTER(CCollision)]
        L.gpv_build_hierarchy.argtypes = [vp, vp, C.POINTER(CHierarchy)]
        L.gpv_batch_release.argtypes = []; L.gpv_batch_release.restype = None
        L.gpv_measure_fp32_peak.argtypes = [vp, vp, C.POINTER(C.c_double)]
        L.gpv_measure_copy_peak.argtypes = [vp, vp, C.POINTER(C.c_double)]
        L.gpv_gather_create.argtypes = [vp, C.c_int64, C.c_int64, C.POINTER(CGatherDesc)]
        L.gpv_gather_create_ex.argtypes = [vp, C.c_int64, C.c_int64, C.c_int, C.POINTER(CGatherDesc)]
        L.gpv_gather_normals.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        L.gpv_gather_attach.argtypes = [vp, C.POINTER(CGatherDesc), C.c_int, C.c_int]
        L.gpv_gather_attach_local.argtypes = [vp, vp, C.c_int, C.c_int]
        L.gpv_gather_detach.argtypes = [vp]; L.gpv_gather_detach.restype = None
        L.gpv_gather_set_timeout.argtypes = [vp, C.c_double]
        L.gpv_gather_result.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _check(rc):
    if rc:
        raise GpvError(lib().gpv_last_error().decode(errors="replace"))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Mesh:
    """Host mesh in the reference's flat layout (Object::CreateFlatTriangleData) with its padded bbox."""

    def __init__(self, cmesh):
        self.c = cmesh
        self.ntri = int(cmesh.n_tri)
        self.bbox_min = np.array(list(cmesh.bbox_min), np.float32)
        self.bbox_max = np.array(list(cmesh.bbox_max), np.float32)
        self.max_model_size = np.float32(cmesh.max_model_size)

    @property
    def tris(self):
        return np.ctypeslib.as_array(self.c.tris, shape=(self.ntri * 9,)).reshape(-1, 9)

    def set_bbox(self, bmin, bmax, max_model_size):
        for a in range(3):
            self.c.bbox_min[a] = float(bmin[a]); self.c.bbox_max[a] = float(bmax[a])
        self.c.max_model_size = float(max_model_size)
        self.bbox_min = np.array(list(self.c.bbox_min), np.float32); self.bbox_max = np.array(list(self.c.bbox_max), np.float32)
        self.max_model_size = np.float32(self.c.max_model_size)

    def __del__(self):
        try:
            lib().gpv_free_mesh(C.byref(self.c))
        except Exception:
            pass


def load_mesh(path, tolerant=False):
    """gpv_load_mesh (the reference's reader semantics), or gpv_load_mesh_ex(GPV_LOAD_TOLERANT): polygons, free-form blanks, ..."""
    m = CMesh()
    if tolerant:
        L = lib()
        L.gpv_load_mesh_ex.argtypes = [C.c_char_p, C.c_uint, C.POINTER(CMesh)]
        _check(L.gpv_load_mesh_ex(os.fsencode(path), 1, C.byref(m)))
    else:
        _check(lib().gpv_load_mesh(os.fsencode(path), C.byref(m)))
    return Mesh(m)


def mesh_from_triangles(tris):
    t = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    m = CMesh()
    _check(lib().gpv_mesh_from_triangles(_fp(t), len(t), C.byref(m)))
    return Mesh(m)


def grid_for(bmin, bmax, max_model_size, l1, l2):
    g = CGrid()
    bmin = np.ascontiguousarray(bmin, np.float32); bmax = np.ascontiguousarray(bmax, np.float32)
    _check(lib().gpv_make_grid(_fp(bmin), _fp(bmax), float(max_model_size), l1, l2, C.byref(g)))
    return g


class Params:
    def __init__(self, l1=8, l2=4, flags=0, z0=0, z1=0):
        self.c = CParams(l1, l2, flags, z0, z1)


class Result:
    """Device-resident result of one voxelization (views are valid until the next call on the same Context)."""

    def __init__(self, cres, ctx):
        self.c, self.ctx = cres, ctx
        g = cres.grid
        self.num_div = np.array(list(g.num_div), np.int32)
        self.grid_size = np.array(list(g.grid_size), np.float32)
        self.grid_size2 = np.array(list(g.grid_size2), np.float32)
        self.n2, self.cells, self.nb, self.n23 = int(g.n2), int(cres.cells), int(cres.n_boundary), int(cres.n23)
        self.z0, self.z1 = int(cres.z0), int(cres.z1)
        self.n_refined = int(cres.n_refined)
        self.counts = [int(cres.l1_inside), int(cres.l1_boundary), int(cres.l2_inside), int(cres.l2_boundary)]
        self.stats = {k: int(getattr(cres, k)) for k in ("l1_box_tests", "l1_box_hits", "l2_box_tests", "l2_ray_tests", "tri_total", "fill_crossings",
                                                         "fill_ill_conditioned", "kernel_launches")}
        self.phase_ms = dict(zip(PHASES, [float(x) for x in cres.phase_ms]))

    def _d2h(self, ptr, n, dtype):
        out = np.empty(int(n), dtype)
        if n and ptr:
            _check(lib().gpv_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, None))
            _check(lib().gpv_stream_sync(None))
        return out

    def level1_inout(self): return self._d2h(self.c.d_level1_inout, self.cells, np.uint8)
    def prefix(self): return self._d2h(self.c.d_prefix, self.cells, np.int32)
    def boundary_index(self): return self._d2h(self.c.d_boundary_index, self.nb, np.int32)
    def level2_inout(self): return self._d2h(self.c.d_level2_inout, self.nb * self.n23, np.uint8)
    def level1_normal(self): return self._d2h(self.c.d_level1_normal, self.cells * 3, np.uint8)
    def level2_normal(self): return self._d2h(self.c.d_level2_normal, self.nb * self.n23 * 3, np.uint8)
    def cell_off(self): return self._d2h(self.c.d_cell_off, self.nb + 1, np.uint32)
    def cell_tris(self): return self._d2h(self.c.d_cell_tris, int(self.c.tri_total), np.int32)
    def col_off(self): return self._d2h(self.c.d_col_off, int(self.num_div[0]) * int(self.num_div[1]) + 1, np.uint32)
    def col_count(self): return self._d2h(self.c.d_col_count, int(self.num_div[0]) * int(self.num_div[1]), np.int32)

    def col_lists(self):
        off, cnt = self.col_off(), self.col_count()
        flat = self._d2h(self.c.d_col_tris, int(off[-1]), np.int32)
        return [flat[o:o + c] for o, c in zip(off[:-1], cnt)]


class Context:
    """One per host thread and device (gpv_ctx)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().gpv_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if self.h:
            lib().gpv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, mesh):
        """Copy the mesh's flat triangles to the device; returns a device pointer owned by the caller (free_device)."""
        L = lib()
        d = L.gpv_alloc_device(mesh.ntri * 36)
        if not d:
            raise GpvError("device allocation failed")
        _check(L.gpv_memcpy_h2d(d, C.cast(mesh.c.tris, C.c_void_p), mesh.ntri * 36, None))
        _check(L.gpv_stream_sync(None))
        return d

    def free_device(self, d):
        lib().gpv_free_device(d)

    def voxelize_device(self, d_tris, mesh, params, stream=None):
        r = CResult()
        _check(lib().gpv_voxelize_device(self.h, d_tris, mesh.ntri, _fp(mesh.bbox_min), _fp(mesh.bbox_max), float(mesh.max_model_size),
                                         C.byref(params.c), stream, C.byref(r)))
        return Result(r, self)

    def voxelize(self, mesh, params):
        """Convenience: upload, voxelize, free the upload (results stay on the device)."""
        d = self.upload(mesh)
        try:
            return self.voxelize_device(d, mesh, params)
        finally:
            self.free_device(d)

    def voxelize_host(self, mesh, params, host, stream=None):
        r = CResult()
        _check(lib().gpv_voxelize_host(self.h, C.byref(mesh.c), C.byref(params.c), stream, C.byref(r), C.byref(host)))
        return Result(r, self)

    def stream(self):
        """The ctx's own non-blocking cudaStream_t (as a c_void_p) for callers that run several contexts side by side."""
        return C.c_void_p(lib().gpv_stream(self.h))

    # ---- multi-GPU gather over NVLink peer memory (gpv_gather_*)
    def gather_create(self, cells_total, l2_capacity, flags=0):
        d = CGatherDesc()
        _check(lib().gpv_gather_create_ex(self.h, int(cells_total), int(l2_capacity), int(flags), C.byref(d)))
        return d

    def gather_normals(self, cells_total, nb, n23):
        """Gathering rank: (level1_normal, level2_normal) copied to the host."""
        p1, p2 = C.c_void_p(), C.c_void_p()
        _check(lib().gpv_gather_normals(self.h, C.byref(p1), C.byref(p2)))
        out = []
        for ptr, n in ((p1, cells_total * 3), (p2, nb * n23 * 3)):
            a = np.empty(int(n), np.uint8)
            if n:
                _check(lib().gpv_memcpy_d2h(a.ctypes.data, ptr, a.nbytes, None))
                _check(lib().gpv_stream_sync(None))
            out.append(a)
        return out[0], out[1]

    def gather_attach(self, desc, rank, world):
        _check(lib().gpv_gather_attach(self.h, C.byref(desc), rank, world))

    def gather_attach_local(self, owner, rank, world):
        _check(lib().gpv_gather_attach_local(self.h, owner.h, rank, world))

    def gather_detach(self):
        lib().gpv_gather_detach(self.h)

    def gather_result(self, cells_total, n23):
        """Gathering rank: (level1_inout, prefix, level2_inout, n_boundary_total) copied to the host."""
        l1, pre, l2, nb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
        _check(lib().gpv_gather_result(self.h, C.byref(l1), C.byref(pre), C.byref(l2), C.byref(nb)))
        out = []
        for ptr, n, dt in ((l1, cells_total, np.uint8), (pre, cells_total, np.int32), (l2, nb.value * n23, np.uint8)):
            a = np.empty(int(n), dt)
            if n:
                _check(lib().gpv_memcpy_d2h(a.ctypes.data, ptr, a.nbytes, None))
                _check(lib().gpv_stream_sync(None))
            out.append(a)
        return out[0], out[1], out[2], int(nb.value)

    def _d2h(self, ptr, n, dtype):
        out = np.empty(int(n), dtype)
        if n and ptr:
            _check(lib().gpv_memcpy_d2h(out.ctypes.data, ptr, out.nbytes, None))
            _check(lib().gpv_stream_sync(None))
        return out

    def collision_boxes(self):
        """gpv_collision_boxes on the last call's streams: (inv_index, centre[n,3], extent[n,3]) copied to the host."""
        c = CCollision()
        _check(lib().gpv_collision_boxes(self.h, None, C.byref(c)))
        n = int(c.count)
        return (self._d2h(c.d_inv_index, n, np.int32) + int(c.index_base), self._d2h(c.d_center, n * 3, np.float32).reshape(-1, 3),
                self._d2h(c.d_extent, n * 3, np.float32).reshape(-1, 3))

    def build_hierarchy(self):
        """gpv_build_hierarchy on the last call's streams (made with GPV_COLLISION): (levels, mid[n,3], half[n,3], solid[n], child[n,2])."""
        h = CHierarchy()
        _check(lib().gpv_build_hierarchy(self.h, None, C.byref(h)))
        n = int(h.n_boxes)
        return (int(h.num_levels), self._d2h(h.d_mid, n * 3, np.float32).reshape(-1, 3), self._d2h(h.d_half, n * 3, np.float32).reshape(-1, 3),
                self._d2h(h.d_solid, n, np.uint8), self._d2h(h.d_child, n * 2, np.int32).reshape(-1, 2))

    def fp32_peak(self):
        v = C.c_double()
        _check(lib().gpv_measure_fp32_peak(self.h, None, C.byref(v)))
        return v.value

    def copy_peak(self):
        v = C.c_double()
        _check(lib().gpv_measure_copy_peak(self.h, None, C.byref(v)))
        return v.value


def check_voxels(directory, obj_id):
    """gpv_check_voxels: True when the six-file set is complete (config parses, every stream has the size it implies)."""
    L = lib()
    L.gpv_check_voxels.argtypes = [C.c_char_p, C.c_int]
    return L.gpv_check_voxels(os.fsencode(directory), int(obj_id)) == 0


def load_voxels(directory, obj_id):
    """gpv_load_voxels -> dict of numpy copies of the streams plus the VoxelConfig fields."""
    v = CVoxelFile()
    L = lib()
    L.gpv_load_voxels.argtypes = [C.c_char_p, C.c_int, C.POINTER(CVoxelFile)]
    L.gpv_free_voxels.argtypes = [C.POINTER(CVoxelFile)]; L.gpv_free_voxels.restype = None
    _check(L.gpv_load_voxels(os.fsencode(directory), obj_id, C.byref(v)))
    try:
        cp = lambda p, n, dt: (np.ctypeslib.as_array(p, shape=(int(n),)).astype(dt, copy=True) if p and n else None)
        l2n = int(v.n_boundary * v.n23)
        return {"name": v.name.decode(), "num_div": list(v.num_div), "num_div2": list(v.num_div2) if v.has_level2 else None,
                "bbox_min": list(v.bbox_min), "bbox_max": list(v.bbox_max), "grid_size": list(v.grid_size), "grid_size2": list(v.grid_size2),
                "counts": [int(v.l1_inside), int(v.l1_boundary), int(v.l2_inside), int(v.l2_boundary)],
                "level1_inout": cp(v.level1_inout, v.cells, np.uint8), "level1_normal": cp(v.level1_normal, v.cells * 3, np.uint8),
                "prefix_sum": cp(v.prefix_sum, v.cells, np.int32), "level2_inout": cp(v.level2_inout, l2n, np.uint8),
                "level2_normal": cp(v.level2_normal, l2n * 3, np.uint8)}
    finally:
        L.gpv_free_voxels(C.byref(v))


def voxelize_batch(paths, params, devices=(0,), threads=4, out_dir=None, first_obj_id=0, skip_existing=False):
    """gpv_voxelize_batch: returns the stats dict; raises GpvError if a model failed."""
    arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    dev = (C.c_int * len(devices))(*devices)
    st = CBatchStats()
    rc = lib().gpv_voxelize_batch(arr, len(paths), C.byref(params.c), dev, len(devices), threads, os.fsencode(out_dir) if out_dir else None, first_obj_id,
                                  int(skip_existing), C.byref(st))
    stats = {k: getattr(st, k) for k, _ in CBatchStats._fields_}
    _check(rc)
    return stats


def save(mesh, result, host, obj_id, directory, omit_absent=False):
    """gpv_save / gpv_save_streams: the six files of Object::SaveVoxelization (omit_absent: no file for a stream that was not computed)."""
    if omit_absent:
        _check(lib().gpv_save_streams(C.byref(mesh.c), C.byref(result.c), C.byref(host), obj_id, os.fsencode(directory), 1))
    else:
        _check(lib().gpv_save(C.byref(mesh.c), C.byref(result.c), C.byref(host), obj_id, os.fsencode(directory)))


def expand_dense(level1_inout, prefix, level2_inout, num_div, n2):
    """gpv_expand_dense: (nz*n2, ny*n2, nx*n2) uint8 array in the file encoding (0 outside / 127 inside / 254 boundary)."""
    l1 = np.ascontiguousarray(level1_inout, np.uint8)
    pre = np.ascontiguousarray(prefix, np.int32)
    l2 = np.ascontiguousarray(level2_inout, np.uint8)
    nd = (C.c_int * 3)(*[int(x) for x in num_div])
    n2 = int(n2)
    out = np.empty((int(num_div[2]) * n2, int(num_div[1]) * n2, int(num_div[0]) * n2), np.uint8)
    L = lib()
    L.gpv_expand_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int64, C.c_void_p, C.c_int64]
    _check(L.gpv_expand_dense(l1.ctypes.data, pre.ctypes.data, l2.ctypes.data, nd, n2, l2.size // max(1, n2 ** 3), out.ctypes.data, out.nbytes))
    return out
