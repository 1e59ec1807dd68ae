/* include/gpview_b200.h -- C ABI of libgpview_b200.so, the B200-native replacement of GPView's hybrid two-level
 * voxelizer hot path.  Plain pointers and sizes only; no CUDA, torch or C++ types in any signature.
 *
 * Two tiers:
 *   1. COMPATIBILITY TIER -- the reference's own operator boundary, same names, argument meaning and error behaviour:
 *        CUDAClassifyTessellation / CUDAClassifyTessellationLevel2 / CUDAClassifyInOutLevel2
 *        (declared includes/CUDAUtilities.h:87-89, defined cuda/CUDAClassifyTessellation.cu:507,518,531)
 *      plus THRUSTDeviceFindMax (includes/CUDAUtilities.h:70; cuda/THRUSTUtilities.cu:44), the only other device symbol the
 *      reference's Object.cpp/CUDAUtilities.cpp link against.  The reference's unmodified Object.cpp links against this
 *      library.
 *   2. NATIVE TIER -- gpv_*: owns the whole path (Object::PerformVoxelization, src/Object.cpp:3077-3430, minus GL): loaders
 *      with Object::ReadObject / ReadOFFObject semantics, grid sizing, the six-kernel device pipeline, and the
 *      Object::SaveVoxelization file contract.
 *
 * Error behaviour: native functions return 0 on success, non-zero on failure, message via gpv_last_error() (per thread).
 * There is NO CPU fallback: every compute entry point fails if no sm_100 device is usable.
 * Threading: one gpv_ctx per host thread; a ctx is bound to one CUDA device.
 */
#ifndef GPVIEW_B200_H
#define GPVIEW_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ compatibility tier (device pointers, caller-owned) */
/* layout-identical to CUDA's float3 / int3 (vector_types.h), which is what the reference passes by value */
typedef struct { float x, y, z; } gpv_float3;
typedef struct { int x, y, z; } gpv_int3;

/* replaces cuda/CUDAClassifyTessellation.cu:507-515.  inOut[cell]=2.0f on a hit, count[cell]++ (always), triIndex[cell*buf+slot]=t
 * only while slot < triBufferLen (the reference writes out of bounds, App. B4).  Returns 1 like the reference. */
int CUDAClassifyTessellation(float* trianglesCUDAData, int numTriangles, float* inOutCUDAData, int* voxelTriCountCUDAData,
                             int* voxelTriIndexCUDAData, gpv_float3 objBoxMin, gpv_float3 objBoxMax, gpv_float3 boxExtents,
                             gpv_int3 numDiv, int triBufferLen);
/* replaces cuda/CUDAClassifyTessellation.cu:518-527 */
int CUDAClassifyTessellationLevel2(float* trianglesCUDAData, float* level2InOutCUDAData, float* level2NormalCUDAData,
                                   float* level1MidPointCUDAData, int* level2IndexCUDAData, int* voxelTriCountCUDAData,
                                   int* level1TriFlatIndexCUDAData, int* level1TriFlatCUDAData, int numBoundary,
                                   gpv_int3 numDiv2, gpv_float3 boxExtentsLevel1, gpv_float3 boxExtentsLevel2);
/* replaces cuda/CUDAClassifyTessellation.cu:531-540 */
int CUDAClassifyInOutLevel2(float* trianglesCUDAData, float* level2InOutCUDAData, float* level1MidPointCUDAData,
                            int* level2IndexCUDAData, int* level1XYTriCountCUDAData, int* level1XYTriFlatIndexCUDAData,
                            int* level1XYTriFlatCUDAData, int numBoundary, gpv_int3 numDiv, gpv_int3 numDiv2,
                            gpv_float3 boxExtentsLevel1, gpv_float3 boxExtentsLevel2);
/* replaces cuda/THRUSTUtilities.cu:44-61 (hand-written max reduction, no Thrust) */
float THRUSTDeviceFindMax(float* dataCUDAPointer, int w, int h);

/* ------------------------------------------------------------------ native tier */
typedef struct gpv_ctx gpv_ctx;

typedef struct {
	int64_t n_tri;
	float* tris;               /* n_tri*9 floats: v0xyz v1xyz v2xyz, file face order (Object::CreateFlatTriangleData, src/Object.cpp:3496) */
	float bbox_min[3], bbox_max[3]; /* padded bounding box (src/Object.cpp:572-583) */
	float max_model_size;      /* largest padded extent (src/Object.cpp:583) */
	int64_t n_verts;
} gpv_mesh;

typedef struct {
	int num_div[3];            /* Level-1 resolution (src/Object.cpp:3098-3106) */
	float grid_size[3];        /* Level-1 cell size  (:3107-3109) */
	float grid_size2[3];       /* Level-2 cell size  (:3128-3130) */
	float ext1[3], ext2[3];    /* half extents (:2551-2552) */
	int n2;                    /* Level-2 resolution per boundary cell (GLParameters::voxelCount2) */
} gpv_grid;

/* flags */
#define GPV_NORMALS      1     /* also produce Level1Normal / Level2Normal streams */
#define GPV_NO_LEVEL2    2     /* GLParameters::level2Voxels == false */
#define GPV_KEEP_LISTS   4     /* CSR cell lists / column lists in canonical (ascending) order for the caller (gpv_result list pointers);
                                  without it (and without GPV_NORMALS) the lists stay in the order the binning left them: occupancy does not depend on it */
#define GPV_PROFILE      8     /* record a CUDA event pair around every kernel of the pipeline -> gpv_result.phase_ms */
#define GPV_PROFILE_L2 128     /* the same for the two Level-2 kernels only (three events: negligible host time, for timed runs) */
#define GPV_PACKED_L2  256     /* gpv_voxelize_host: Level2InOut crosses PCIe as 2 bits per sub-voxel and is expanded into the caller's bytes
                                  by a pool of host threads while the next chunk is on the bus (same bytes, a quarter of the transfer).
                                  Ignored when n2^3 is not a multiple of 32.  GPV_HOST_THREADS sets the pool size (default: all cores but one). */
#define GPV_COLLISION  512     /* also keep the parity fill of every cell (incl. boundary cells) on the device: needed by gpv_build_hierarchy */
#define GPV_GATHER      16     /* multi-GPU: write this rank's share of the streams straight into the gathering rank's buffers (gpv_gather_*) */
#define GPV_BATCH_TOLERANT_LOAD 64 /* gpv_voxelize_batch: read the meshes with GPV_LOAD_TOLERANT (gpv_load_mesh_ex) */
#define GPV_SAVE_COMPUTED_ONLY 32 /* gpv_voxelize_batch: write only the streams that were computed -- no 127-filled normal files when
                                  GPV_NORMALS is off (74 % of a 64 + 4^3 model's bytes, and file writing is what bounds a dataset run) */

typedef struct {
	int voxel_count;           /* GLParameters::voxelCount  (Level-1 cells along the longest axis; reference default 8) */
	int voxel_count2;          /* GLParameters::voxelCount2 (Level-2 cells per boundary cell axis; reference default 4) */
	int flags;
	int z0, z1;                /* z-slab [z0,z1) owned by this call; z1 <= 0 means the whole grid */
} gpv_params;

/* Result of one voxelization.  All d_* pointers are DEVICE memory owned by the ctx, valid until the next call on the ctx.
 * Streams are in the reference's file layout; for a slab they cover only the slab's cells / boundary cells
 * (contiguous ranges of the whole-grid streams; prefix sums and Level-2 blocks are slab-local, boundary_index is global). */
typedef struct {
	gpv_grid grid;
	int z0, z1;
	int64_t cells;             /* (z1-z0)*ny*nx */
	int64_t n_boundary;        /* boundary cells in the slab */
	int64_t n23;               /* n2^3 */
	uint8_t* d_level1_inout;   /* cells bytes: 0 outside / 127 inside / 254 boundary  (ObjNLevel1InOut.raw) */
	int32_t* d_prefix;         /* cells+1 int32: exclusive boundary prefix sum (ObjNLevel1BoundaryPrefixSum.raw), [cells] = n_boundary */
	int32_t* d_boundary_index; /* n_boundary int32: global linear index of each boundary cell, ascending */
	uint8_t* d_level2_inout;   /* n_boundary*n23 bytes (ObjNLevel2InOut.raw) */
	uint8_t* d_level1_normal;  /* cells*3 bytes or NULL */
	uint8_t* d_level2_normal;  /* n_boundary*n23*3 bytes or NULL */
	/* CSR lists (GPV_KEEP_LISTS): ascending triangle ids */
	uint32_t* d_cell_off;      /* n_boundary+1, indexed by slab-local boundary rank */
	int32_t* d_cell_tris;
	uint32_t* d_col_off;       /* nx*ny+1 */
	int32_t* d_col_count;      /* nx*ny */
	int32_t* d_col_tris;
	/* counts (src/Object.cpp:3353-3378) */
	int64_t l1_inside, l1_boundary, l2_inside, l2_boundary;
	/* reference-equivalent work */
	int64_t l1_box_tests, l1_box_hits, l2_box_tests, l2_ray_tests, tri_total, fill_crossings, fill_ill_conditioned;
	int64_t kernel_launches;   /* kernels launched by this call */
	/* GPV_PROFILE: device time of each phase in ms (CUDA events on the caller's stream), indexed by GPV_PHASE_* */
	float phase_ms[16];
	int64_t n_refined;         /* boundary cells whose Level-2 blocks THIS call produced: n_boundary, or with GPV_GATHER this rank's share */
} gpv_result;

enum { GPV_PHASE_SETUP = 0, GPV_PHASE_BIN_COUNT, GPV_PHASE_CROSS_COUNT, GPV_PHASE_SCAN, GPV_PHASE_HOST_GAP, GPV_PHASE_BIN_FILL,
       GPV_PHASE_CROSS_FILL, GPV_PHASE_SORT, GPV_PHASE_FILL_SWEEP, GPV_PHASE_L1_NORMALS, GPV_PHASE_L2_RAYS, GPV_PHASE_L2, GPV_PHASE_L2_NORMALS, GPV_PHASE_COUNT };

/* host copies of the streams (caller-allocated; any pointer may be NULL to skip that stream) */
typedef struct {
	uint8_t* level1_inout; int32_t* prefix; int32_t* boundary_index; uint8_t* level2_inout;
	uint8_t* level1_normal; uint8_t* level2_normal;
	int64_t level2_capacity;   /* bytes available behind level2_inout (level2_normal must hold 3x) */
	int64_t boundary_capacity; /* entries available behind boundary_index */
} gpv_host_streams;

const char* gpv_last_error(void);
int gpv_device_count(void);

/* context: device buffers are pooled (grow-only) inside the ctx */
int gpv_create(int device, gpv_ctx** out);
void gpv_destroy(gpv_ctx* ctx);
/* a non-blocking cudaStream_t owned by the ctx: pass it as `stream` when several contexts run side by side (the NULL stream is
 * CUDA's legacy default stream and serialises them) */
void* gpv_stream(gpv_ctx* ctx);

/* loaders with the reference's semantics (Object::ReadObject src/Object.cpp:395-584, Object::ReadOFFObject :171-317);
 * gpv_load_mesh dispatches on the last three characters like main() (src/GPView.cpp:1642-1659) */
int gpv_load_obj(const char* path, gpv_mesh* out);
/* Where the reference's readers die, these report: a field std::stof / std::stoi would throw invalid_argument on, a face index out
 * of range, a missing file -> non-zero + gpv_last_error() (the reference abort()s, crashes or reads garbage).  One leniency: a
 * coordinate std::stof rejects as out of range ("1e-40", "1e39": strtof sets ERANGE) is taken with strtof's value. */
int gpv_load_off(const char* path, gpv_mesh* out);
int gpv_load_mesh(const char* path, gpv_mesh* out);
/* EXTENSION (not reference behaviour): flags = GPV_LOAD_TOLERANT reads what files in the wild contain and the
 * reference's readers refuse or misread -- any run of blanks between fields, CRLF, an unterminated last line, comments, a fourth
 * vertex coordinate or colour fields, negative (relative) OBJ indices, and polygons with more than three vertices (fan
 * triangulation a0, a_k, a_k+1) in both formats.  Lenient about layout, not about digits: a number field must be a number as a
 * whole ("1.5abc", or "0.5" where an index belongs, is an error; the strict readers keep the reference's strtof / strtol prefix rule).  A file with three-vertex faces, single delimiters, full `v` lines and a final
 * newline gives the same mesh either way.  flags = 0 is gpv_load_mesh. */
#define GPV_LOAD_TOLERANT 1u
int gpv_load_mesh_ex(const char* path, unsigned flags, gpv_mesh* out);
int gpv_mesh_from_triangles(const float* tris, int64_t n_tri, gpv_mesh* out); /* bbox over the given vertices + padding */
void gpv_free_mesh(gpv_mesh* m); /* the only way to release gpv_mesh.tris (the block has a header in front of the floats and may be
                                  * parked for the calling thread's next load) */

/* grid sizing of Object::PerformVoxelization (src/Object.cpp:3094-3134) */
int gpv_make_grid(const float bbox_min[3], const float bbox_max[3], float max_model_size, int voxel_count, int voxel_count2, gpv_grid* out);

/* pinned host memory for the callers' staging buffers */
void* gpv_alloc_host(int64_t bytes);
void gpv_free_host(void* p);
void* gpv_alloc_device(int64_t bytes);
void gpv_free_device(void* p);
int gpv_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream);
int gpv_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream);
int gpv_stream_sync(void* stream);

/* the hot path.  d_tris: DEVICE pointer to n_tri*9 floats.  stream: a cudaStream_t (NULL = default stream).
 * Asynchronous except for one internal size read-back; counts in `out` are final on return (the call ends with a stream sync). */
int gpv_voxelize_device(gpv_ctx* ctx, const float* d_tris, int64_t n_tri, const float bbox_min[3], const float bbox_max[3],
                        float max_model_size, const gpv_params* params, void* stream, gpv_result* out);
/* same, from HOST triangles into HOST streams: H2D + pipeline + D2H (the reference-facing call: what
 * Object::PerformVoxelization does between CreateFlatTriangleData and SaveVoxelization) */
int gpv_voxelize_host(gpv_ctx* ctx, const gpv_mesh* mesh, const gpv_params* params, void* stream, gpv_result* out, gpv_host_streams* host);

/* ---- multi-GPU gather over NVLink peer memory (one process per GPU, one ctx per process).
 * The gathering rank (rank 0) allocates the whole-grid streams once and exports them as CUDA IPC handles; every other rank maps
 * them (peer access over NVLink / NVSwitch).  A call with GPV_GATHER then delivers its share of the streams directly at their
 * final offsets in rank 0's memory -- there is no separate collective and NO exchange step on the data path: every rank runs
 * Level 1 over the whole grid (the parity rays need whole column lists, cu:461-463), so boundary ranks and prefix sums are
 * global on every rank.  Shared out are
 *   - the Level1InOut bytes and prefix sums by z-slab: [z0,z1) of gpv_params, or an equal share of the layers when z1 <= 0; the
 *     two contiguous ranges leave by the copy engines beside the Level-2 kernels;
 *   - the Level-2 refinement by Level-1 column: groups of max(1, 256/n2^2) consecutive columns are dealt to the ranks round-robin
 *     (skewed by the grid row), so that every column list is walked by exactly one rank and neighbouring columns (similar cost)
 *     land on different ranks.  A peer's blocks are stored from inside k_l2 as 2 bits per sub-voxel (n2 a multiple of 4; file bytes
 *     otherwise) and expanded into Level2InOut bytes on rank 0, peer by peer as they finish.
 * Completion flags and the ranks' shares of the counts travel through a mailbox in rank 0's memory (one-thread kernels on the
 * callers' streams); rank 0's call returns once every rank has signalled, with the whole grid's counts in its gpv_result (the
 * other ranks report their own share).  All ranks must issue their GPV_GATHER calls in the same order; a rank's call does not
 * touch rank 0's buffers before rank 0 has entered the same call.  gpv_gather_desc is plain bytes: ship it to the other ranks
 * with any transport (torch.distributed broadcast, a file, a pipe).  The host-stream call is not supported together with
 * GPV_GATHER; GPV_NORMALS is, on buffers created by gpv_gather_create_ex(GPV_NORMALS): every rank sends the Level-1 normals of its z-slab and
 * the Level-2 normals of the blocks it refined.  Every poll gives up after a timeout (5 s; gpv_gather_set_timeout) with an error return instead of
 * hanging the GPU.  If several gathering contexts share ONE process and device (tests): raise CUDA_DEVICE_MAX_CONNECTIONS so
 * that their streams do not share a hardware work queue, and do not allocate device memory while a gathering call is in flight
 * (CUDA serialises streams around cudaMalloc/cudaFree) -- run one plain call first to grow the pools. */
typedef struct {
	unsigned char l1[64], prefix[64], l2[64], mailbox[64];   /* cudaIpcMemHandle_t of the four allocations */
	int64_t cells_total;                                     /* nx*ny*nz of the grid the buffers were sized for */
	int64_t l2_capacity;                                     /* bytes behind l2 */
	int32_t owner_device, reserved;                          /* reserved: flags the buffers were created with (GPV_NORMALS) */
} gpv_gather_desc;
int gpv_gather_create(gpv_ctx* ctx, int64_t cells_total, int64_t l2_capacity, gpv_gather_desc* out);      /* gathering rank only */
/* flags = GPV_NORMALS: room for the two normal streams as well (3 B per cell / sub-voxel), so that GPV_GATHER | GPV_NORMALS calls can be made */
int gpv_gather_create_ex(gpv_ctx* ctx, int64_t cells_total, int64_t l2_capacity, int flags, gpv_gather_desc* out);
int gpv_gather_attach(gpv_ctx* ctx, const gpv_gather_desc* desc, int rank, int world);                    /* every rank, the gathering one included */
int gpv_gather_attach_local(gpv_ctx* ctx, gpv_ctx* owner, int rank, int world);                           /* same-process ranks (tests, several GPUs in one process) */
void gpv_gather_detach(gpv_ctx* ctx);                                /* the creating ctx keeps its buffers and may be attached again (a new session) */
int gpv_gather_set_timeout(gpv_ctx* ctx, double seconds);            /* how long a poll waits for a peer before the call fails (default 5 s) */
/* gathering rank, after its own GPV_GATHER call returned (it returns once every rank has signalled completion): device views of
 * the whole-grid streams and the total boundary count */
int gpv_gather_result(gpv_ctx* ctx, uint8_t** d_level1_inout, int32_t** d_prefix, uint8_t** d_level2_inout, int64_t* n_boundary_total);
int gpv_gather_normals(gpv_ctx* ctx, uint8_t** d_level1_normal, uint8_t** d_level2_normal);   /* gathering rank, buffers of gpv_gather_create_ex(GPV_NORMALS) */

/* Object::SaveVoxelization (src/Object.cpp:2934-3075): the six ObjN*.{txt,raw} files into `dir` from host streams */
int gpv_save(const gpv_mesh* mesh, const gpv_result* res, const gpv_host_streams* host, int obj_id, const char* dir);
/* the same; omit_absent != 0: a stream whose host pointer is NULL gets no file (gpv_save writes its neutral value instead: 127
 * for normals, 0 otherwise, so that all six files of the reference's contract always exist).  gpv_load_voxels accepts both. */
int gpv_save_streams(const gpv_mesh* mesh, const gpv_result* res, const gpv_host_streams* host, int obj_id, const char* dir, int omit_absent);

/* Reader of the six-file set written by gpv_save / Object::SaveVoxelization: sizes come from ObjNVoxelConfig.txt (the reference can
 * only read back one hard-coded 48x64x64 grid, Object::ReadRAWObject src/Object.cpp:319-392).  Arrays are malloc'ed; the two
 * normal streams are NULL when their files are absent.  Free with gpv_free_voxels. */
typedef struct {
	char name[64];             /* "Obj<N>" */
	float bbox_min[3], bbox_max[3];
	int num_div[3]; float grid_size[3];
	int64_t l1_inside, l1_boundary;
	int has_level2; int num_div2[3]; float grid_size2[3];
	int64_t l2_inside, l2_boundary;
	int64_t cells, n_boundary, n23;
	uint8_t* level1_inout; uint8_t* level1_normal; int32_t* prefix_sum; uint8_t* level2_inout; uint8_t* level2_normal;
} gpv_voxel_file;
int gpv_load_voxels(const char* dir, int obj_id, gpv_voxel_file* out);
/* 0 if the set ObjN* in `dir` is complete: ObjNVoxelConfig.txt parses and every stream it implies exists with exactly the size
 * it implies (normal streams may be absent, not truncated).  Reads nothing but the config.  gpv_save_streams writes the streams
 * first and renames the config into place last, so an interrupted run never leaves a set that passes this test. */
int gpv_check_voxels(const char* dir, int obj_id);
void gpv_free_voxels(gpv_voxel_file* v);

/* The two-level result as one dense grid at the effective resolution (num_div[a] * n2 per axis, z-major, file encoding 0/127/254):
 * the input of a 3-D CNN (the reference's stated consumer, README.md:22-27).  Inside / outside Level-1 cells are filled with
 * their state, boundary cells with their Level-2 block.  Host memory in, host memory out (out_bytes >= the dense size). */
int gpv_expand_dense(const uint8_t* level1_inout, const int32_t* prefix, const uint8_t* level2_inout, const int num_div[3], int n2,
                     int64_t n_boundary, uint8_t* out, int64_t out_bytes);

/* ---- voxel hierarchy / collision structures over the Level-1 grid, from the streams the LAST call on the ctx left on
 * the device (valid until the next call).  The reference builds both on the host, from its per-cell BBoxData array.
 * gpv_collision_boxes = Object::CollisionInitCUDA (src/Object.cpp:3530-3572): the occupied cells (inside or boundary), ascending:
 *   d_inv_index[k] = linear cell index - index_base (VoxelData::invIndex), d_center / d_extent = 3 floats per box (boxCenterCUDAData /
 *   boxExtentCUDAData: midPoint, halfSize of the cell).
 * gpv_build_hierarchy = Object::BuildHierarchy (:2790-2867) + CombineBBox (:2750-2788): the binary AABB hierarchy (cells - 1 boxes: x-neighbour
 *   pairs first, then x / y / z halved in rotation), per box midPoint, halfSize, solid, the two child indices (children of the first level
 *   are cell indices -- the reference leaves those uninitialised --, of every further level indices into this array).  `solid` of a cell is
 *   its parity fill BEFORE the SAT pass (also for boundary cells, :3165-3193): the call must have been made with GPV_COLLISION, on the whole
 *   grid.  Every grid dimension must be a power of two: the reference's loop indexes out of bounds otherwise (an error here). */
typedef struct { int64_t count, index_base; int32_t* d_inv_index; float* d_center; float* d_extent; } gpv_collision;
typedef struct { int num_levels; int64_t n_boxes; float* d_mid; float* d_half; uint8_t* d_solid; int32_t* d_child; } gpv_hierarchy;
int gpv_collision_boxes(gpv_ctx* ctx, void* stream, gpv_collision* out);
int gpv_build_hierarchy(gpv_ctx* ctx, void* stream, gpv_hierarchy* out);

/* 2-bit packed Level-2 words -> file bytes (host memory in, host memory out).  One word pair (uint32 inside mask, uint32 boundary
 * mask) per 32 consecutive sub-voxels of Level2InOut.raw, bit k = sub-voxel 32*w + k; out gets 32*n_words bytes 0 / 127 / 254.
 * This is the format Level 2 crosses NVLink (GPV_GATHER) and PCIe (GPV_PACKED_L2) in; exported for consumers that keep it packed. */
int gpv_expand_packed_l2(const void* packed, int64_t n_words, uint8_t* out);

/* Batched dataset generation (BASELINE.json config 5): `threads` host threads, each with its own ctx on devices[w % n_devices],
 * pull paths from a shared queue: load -> gpv_voxelize_host -> gpv_save(out_dir, obj id = first_obj_id + index).  out_dir NULL:
 * nothing is written.  skip_existing: a model whose set is complete (gpv_check_voxels) is skipped (restartable).  The *_seconds are
 * summed over threads.  Returns non-zero if any model failed (message of the first failure in gpv_last_error()). */
typedef struct {
	int64_t models_done, models_failed, models_skipped;
	double seconds, parse_seconds, gpu_seconds, save_seconds;
	int64_t level2_resizes;    /* models whose Level-2 host buffer had to grow: sized from a Level-1-only run, then voxelized again */
} gpv_batch_stats;
int gpv_voxelize_batch(const char* const* paths, int64_t n_paths, const gpv_params* params, const int* devices, int n_devices, int threads,
                       const char* out_dir, int first_obj_id, int skip_existing, gpv_batch_stats* stats);
/* The workers' contexts (streams, pooled device buffers) are kept for the next gpv_voxelize_batch call of the process; this frees them. */
void gpv_batch_release(void);

/* micro-benchmarks used for the roofline denominators (bench.py): achieved non-FMA FP32 lane-ops/s and copy GB/s */
int gpv_measure_fp32_peak(gpv_ctx* ctx, void* stream, double* ops_per_s);
int gpv_measure_copy_peak(gpv_ctx* ctx, void* stream, double* gb_per_s);

#ifdef __cplusplus
}
#endif
#endif
