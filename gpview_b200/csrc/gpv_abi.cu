// gpview_b200/csrc/gpv_abi.cu -- C ABI (include/gpview_b200.h): context, device pipeline orchestration, compat tier.
//
// Pipeline of gpv_voxelize_device (what Object::PerformVoxelization, src/Object.cpp:3077-3430, does between
// CreateFlatTriangleData and SaveVoxelization, re-designed for one B200):
//
//   k_clear                 the small counters of the call (cellCount and the (triangle, column) bitmaps clean up after themselves)
//   k_prepare               48 B triangle / ray / plane records, footprints, work-item counts, centre tables
//   k_scan_offs3            balanced work spaces of the two triangle-parallel sweeps (two scans, one launch)
//   k_bin<count>            K1 count sweep: cellCount, colCount, l1Hits               | side stream: k_cross<count>  K2a crossCount
//   k_scan<CELLS>           K3: prefix, boundaryIndex, bTriOff, bmask, per-column boundary-cell counts (of the columns this rank owns)
//   k_scan_offs3            column-list, crossing-list and column-cell offsets (three scans, one launch)
//   k_ray_units             work list of the Level-2 parity rays: (column, <= 16 of its boundary cells), expensive units first
//   ---- one 128-byte read-back (sizes of the variable-length buffers) ----
//   k_bin<fill>             cell lists, column lists                                  | side stream: k_cross<fill>, k_fill_sweep K2b (final
//   k_sort_segments / k_sort_long   (GPV_NORMALS / GPV_KEEP_LISTS) canonical order    |   Level-1 bytes + inside count), Level-1 streams to the
//   k_l1_normals            (GPV_NORMALS)                                             |   host / the gathering rank, k_clear_bits
//   k_col_cells, k_l2_rays, k_l2_rays_overflow   K4a: boundary cells by column, Level-2 parity bits per sub-voxel column
//   k_l2<n2, out>           K4: Level-2 SAT + counts + the blocks as file bytes / staged bytes / 2 bits per sub-voxel (chunked with a host sink)
//   k_scatter_blocks, k_l2_normals   (GPV_NORMALS)
//   k_gather_begin / done / wait_rank + k_gather_expand / wait   (GPV_GATHER) mailbox flags, rank 0 expands the peers' 2-bit blocks
// A small model (<= 1 M cells) runs all of it on the caller's stream alone (no fork / join events: the driver calls bound a dataset run).
// gpv_collision_boxes / gpv_build_hierarchy (gpv_collision.cuh) work on the streams the last call left on the device.
#include "../../include/gpview_b200.h"
#include "gpv_internal.h"
#include "gpv_kernels.cuh"
#include "gpv_collision.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace gpv {

static thread_local std::string g_err;
int fail(const std::string& msg) { g_err = msg; return 1; }

#define GPV_CUDA(call)                                                                                              \
	do {                                                                                                            \
		cudaError_t e_ = (call);                                                                                    \
		if (e_ != cudaSuccess) return gpv::fail(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
	} while (0)

struct DevBuf {
	void* p = nullptr;
	size_t cap = 0;
	int ensure(size_t bytes)
	{
		if (bytes <= cap) return 0;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = bytes + bytes / 8 + 256; // grow-only pool with slack so that repeated models do not reallocate
		cudaError_t e = cudaMalloc(&p, want);
		if (e != cudaSuccess) return gpv::fail(std::string("cudaMalloc(") + std::to_string(want) + "): " + cudaGetErrorString(e));
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
	template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

} // namespace gpv

struct gpv_ctx {
	int device = 0;
	int smCount = 0;
	gpv::DevBuf tri48, ray48, tabX, tabY, tabZ, cellCount, colCount, crossCount, prefix, bmask, boundaryIndex, bTriOff, cellTris,
	    colOff, colTris, crossOff, crossTri, l1State, l2State, l1Normal, l2Normal, desc, totals, scratch, crossFp, binCnt, binOff, crossCnt, crossWorkOff, plane16, aabbxy16, longList, colCursor, binBits, colCellCnt, colCellOff, colCellList, l2Par, cellMid, rayOver, rayOverflow, l2Packed, peerCells, solidWords, occCount, occOff, occInv, occCenter, occExtent, hierMid, hierHalf, hierSolid, hierChild;
	gpv::Totals* hTotals = nullptr; // pinned
	cudaEvent_t ev[GPV_PHASE_COUNT + 1] = {}, evEnd[GPV_PHASE_COUNT + 1] = {};
	bool haveEvents = false;
	cudaStream_t copyStream = nullptr;  // D2H of finished streams overlaps the rest of the pipeline (gpv_voxelize_host)
	cudaStream_t ownStream = nullptr;   // gpv_stream(): a non-blocking stream for callers that run several contexts side by side
	cudaStream_t sideStream = nullptr;  // the parity-fill branch of the Level-1 pipeline runs beside the binning / sorting branch
	cudaEvent_t evFork[2] = {}, evJoin[2] = {};
	cudaEvent_t evChunk[17] = {};
	cudaEvent_t evCopied[16] = {};      // GPV_PACKED_L2: a chunk of packed Level-2 words has landed in hPacked
	void* hPacked = nullptr; size_t hPackedCap = 0; // pinned staging of the packed Level-2 stream (gpv_voxelize_host)
	bool sortAttrSet = false;
	// Counters that clean up after themselves: a call that completes leaves cellCount all zero again (the fill sweep of the binning
	// takes its slots by atomic decrement) and wipes the part of the (triangle, column) bitmaps it used on the side stream, beside
	// Level 2 -- so the next call clears neither on its critical path.  What is known to be zero:
	void* cleanCellCount = nullptr; size_t cleanCellBytes = 0; // [cleanCellCount, +cleanCellBytes)
	void* cleanBits = nullptr;                                  // the whole binBits pool at this address
	cudaEvent_t evBinDone = nullptr;
	// what the last completed call left on the device, for gpv_collision_boxes / gpv_build_hierarchy
	struct { bool valid = false, whole = false, solid = false; gpv::GridP g{}; long long cells = 0; } last;
	int debugOwnWorld = 0, debugOwnRank = 0; // GPV_DEBUG_OWN (profiling one rank's share of a gathering call on a single GPU)
	// GPV_GATHER (gpv_gather_*): the gathering rank's whole-grid streams and mailbox, local or mapped over NVLink
	struct {
		bool on = false, owner = false, ipc = false;
		int rank = 0, world = 1;
		unsigned epoch = 0;
		uint8_t* l1 = nullptr; int32_t* prefix = nullptr; uint8_t* l2 = nullptr; gpv::GatherMail* mail = nullptr;
		uint8_t* l2p = nullptr; // 2-bit packed Level-2 blocks of the peers (behind the byte stream in the same allocation)
		uint8_t* l1n = nullptr; uint8_t* l2n = nullptr; // normal streams (behind the Level-1 bytes / the packed blocks), null without GPV_NORMALS
		int flags = 0;
		int64_t cellsTotal = 0, l2Cap = 0, nbTotal = 0;
		unsigned long long timeoutNs = gpv::kGatherTimeoutNsDefault;
	} gather;
	gpv::DevBuf gatherL1, gatherPrefix, gatherL2, gatherMail;
};

using namespace gpv;

extern "C" const char* gpv_last_error(void) { return g_err.c_str(); }

extern "C" int gpv_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

// CUDA loads kernels lazily, and loading one synchronises the context.  With GPV_GATHER a rank's one-thread polling kernel may be
// running while another context of the same process launches a kernel for the first time -- the load would wait for the poll,
// the poll for that rank.  So every kernel of the pipeline is loaded when the first context of a device is created.
static int preload_kernels()
{
	cudaFuncAttributes a;
#define GPV_LOAD(...) GPV_CUDA(cudaFuncGetAttributes(&a, __VA_ARGS__))
	GPV_LOAD(k_clear); GPV_LOAD(k_clear_bits); GPV_LOAD(k_prepare); GPV_LOAD(k_scan_offs3); GPV_LOAD((k_scan<MODE_CELLS, 1>)); GPV_LOAD((k_scan<MODE_CELLS, 4>));
	GPV_LOAD(k_bin<false>); GPV_LOAD(k_bin<true>); GPV_LOAD(k_cross<false>); GPV_LOAD(k_cross<true>); GPV_LOAD(k_fill_sweep);
	GPV_LOAD(k_sort_segments); GPV_LOAD(k_sort_long);
	GPV_LOAD(k_col_cells); GPV_LOAD(k_l2_rays); GPV_LOAD(k_l2_rays_overflow); GPV_LOAD(k_ray_units); GPV_LOAD(k_l1_normals); GPV_LOAD(k_l2_normals); GPV_LOAD(k_scatter_blocks); GPV_LOAD(k_occupied_count); GPV_LOAD(k_occupied_write); GPV_LOAD(k_hier_leaves); GPV_LOAD(k_hier_level);
	GPV_LOAD(k_l2<16, 0>); GPV_LOAD(k_l2<8, 0>); GPV_LOAD(k_l2<4, 0>); GPV_LOAD(k_l2<2, 0>); GPV_LOAD(k_l2<0, 0>);
	GPV_LOAD(k_l2<16, 1>); GPV_LOAD(k_l2<8, 1>); GPV_LOAD(k_l2<4, 1>); GPV_LOAD(k_l2<2, 1>); GPV_LOAD(k_l2<0, 1>);
	GPV_LOAD(k_l2<16, 2>); GPV_LOAD(k_l2<8, 2>); GPV_LOAD(k_l2<4, 2>); GPV_LOAD(k_l2<0, 2>); GPV_LOAD(k_gather_expand); GPV_LOAD(k_gather_wait_rank);
	GPV_LOAD(k_gather_begin); GPV_LOAD(k_gather_done); GPV_LOAD(k_gather_wait);
#undef GPV_LOAD
	return 0;
}

extern "C" int gpv_create(int device, gpv_ctx** out)
{
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) return fail("no CUDA device: libgpview_b200 has no CPU fallback");
	if (device < 0 || device >= n) return fail("gpv_create: bad device index");
	GPV_CUDA(cudaSetDevice(device));
	cudaDeviceProp prop;
	GPV_CUDA(cudaGetDeviceProperties(&prop, device));
	if (prop.major != 10) return fail(std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
	                                  "; libgpview_b200 carries sm_100a code only");
	{
		static std::mutex mu;
		static bool loaded[64] = {};
		std::lock_guard<std::mutex> lock(mu);
		if (device < 64 && !loaded[device]) { if (preload_kernels()) return 1; loaded[device] = true; }
	}
	gpv_ctx* c = new gpv_ctx();
	c->device = device;
	c->smCount = prop.multiProcessorCount;
	if (const char* e = getenv("GPV_DEBUG_OWN")) { if (sscanf(e, "%d,%d", &c->debugOwnWorld, &c->debugOwnRank) != 2 || c->debugOwnRank < 0 || c->debugOwnRank >= c->debugOwnWorld) c->debugOwnWorld = 0; }
	static_assert(sizeof(Totals) <= 128, "gpv_ctx::totals: Totals in the first 128 bytes, the sort's long-list counters behind");
	auto init = [c]() -> int {
		GPV_CUDA(cudaHostAlloc((void**)&c->hTotals, 256, cudaHostAllocDefault)); // Totals + the counters behind it
		if (c->totals.ensure(256)) return 1;
		GPV_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
		GPV_CUDA(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
		GPV_CUDA(cudaStreamCreateWithFlags(&c->sideStream, cudaStreamNonBlocking));
		for (int k = 0; k < 2; k++) { GPV_CUDA(cudaEventCreateWithFlags(&c->evFork[k], cudaEventDisableTiming)); GPV_CUDA(cudaEventCreateWithFlags(&c->evJoin[k], cudaEventDisableTiming)); }
		for (cudaEvent_t& e : c->evChunk) GPV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		for (cudaEvent_t& e : c->evCopied) GPV_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
		GPV_CUDA(cudaEventCreateWithFlags(&c->evBinDone, cudaEventDisableTiming));
		return 0;
	};
	if (init()) { // a half-built context owns nothing the caller could release: gpv_destroy() frees whatever was created
		const std::string why = g_err;
		gpv_destroy(c);
		return fail(why);
	}
	*out = c;
	return 0;
}

extern "C" void gpv_destroy(gpv_ctx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	DevBuf* all[] = { &c->tri48, &c->ray48, &c->tabX, &c->tabY, &c->tabZ, &c->cellCount, &c->colCount, &c->crossCount, &c->prefix, &c->bmask,
		              &c->boundaryIndex, &c->bTriOff, &c->cellTris, &c->colOff, &c->colTris, &c->crossOff, &c->crossTri, &c->l1State, &c->l2State,
		              &c->l1Normal, &c->l2Normal, &c->desc, &c->totals, &c->scratch, &c->crossFp, &c->binCnt, &c->binOff, &c->crossCnt, &c->crossWorkOff, &c->plane16, &c->aabbxy16, &c->longList, &c->colCursor, &c->binBits, &c->colCellCnt, &c->colCellOff, &c->colCellList, &c->l2Par, &c->cellMid, &c->rayOver, &c->rayOverflow, &c->l2Packed, &c->peerCells, &c->solidWords, &c->occCount, &c->occOff, &c->occInv, &c->occCenter, &c->occExtent, &c->hierMid, &c->hierHalf, &c->hierSolid, &c->hierChild, &c->gatherL1, &c->gatherPrefix, &c->gatherL2, &c->gatherMail };
	gpv_gather_detach(c);
	for (DevBuf* b : all) b->release();
	if (c->hTotals) cudaFreeHost(c->hTotals);
	if (c->haveEvents) { for (cudaEvent_t e : c->ev) cudaEventDestroy(e); for (cudaEvent_t e : c->evEnd) cudaEventDestroy(e); }
	for (int k = 0; k < 2; k++) { if (c->evFork[k]) cudaEventDestroy(c->evFork[k]); if (c->evJoin[k]) cudaEventDestroy(c->evJoin[k]); }
	if (c->evBinDone) cudaEventDestroy(c->evBinDone);
	if (c->sideStream) cudaStreamDestroy(c->sideStream);
	for (cudaEvent_t e : c->evChunk) if (e) cudaEventDestroy(e);
	for (cudaEvent_t e : c->evCopied) if (e) cudaEventDestroy(e);
	if (c->hPacked) cudaFreeHost(c->hPacked);
	if (c->copyStream) cudaStreamDestroy(c->copyStream);
	if (c->ownStream) cudaStreamDestroy(c->ownStream);
	delete c;
}

extern "C" void* gpv_stream(gpv_ctx* c) { return c ? (void*)c->ownStream : nullptr; }

extern "C" void* gpv_alloc_host(int64_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, (size_t)(bytes > 0 ? bytes : 1), cudaHostAllocDefault) != cudaSuccess) { fail("cudaHostAlloc failed"); return nullptr; }
	return p;
}
extern "C" void gpv_free_host(void* p) { if (p) cudaFreeHost(p); }
extern "C" void* gpv_alloc_device(int64_t bytes)
{
	void* p = nullptr;
	if (cudaMalloc(&p, (size_t)(bytes > 0 ? bytes : 1)) != cudaSuccess) { fail("cudaMalloc failed"); return nullptr; }
	return p;
}
extern "C" void gpv_free_device(void* p) { if (p) cudaFree(p); }
extern "C" int gpv_memcpy_h2d(void* dst, const void* src, int64_t bytes, void* stream)
{
	GPV_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
	return 0;
}
extern "C" int gpv_memcpy_d2h(void* dst, const void* src, int64_t bytes, void* stream)
{
	GPV_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	return 0;
}
extern "C" int gpv_stream_sync(void* stream)
{
	GPV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
	return 0;
}

// offset scans (k_scan_offs3): up to three per launch; the look-back descriptors live in regions of c->desc cleared by k_clear
struct ScanReq { const int* in; long long n; unsigned* off; unsigned* totalOut; unsigned long long* totalOut64; size_t descOff; };
static size_t desc_bytes(long long n) { return (size_t)(((n + kScanTile - 1) / kScanTile + 1) * 8 + 16 + 15) & ~(size_t)15; }
static void launch_scans(gpv_ctx* c, cudaStream_t st, const ScanReq* r, int count, int64_t& launches)
{
	ScanIO3 io3{};
	long long maxTiles = 1;
	for (int k = 0; k < count; k++) {
		ScanIO& io = io3.s[k];
		io.in = r[k].in; io.n = r[k].n;
		io.tileCounter = reinterpret_cast<unsigned*>(c->desc.as<char>() + r[k].descOff);
		io.desc = reinterpret_cast<unsigned long long*>(c->desc.as<char>() + r[k].descOff) + 1;
		io.off = r[k].off; io.totalOut = r[k].totalOut; io.totalOut64 = r[k].totalOut64;
		maxTiles = std::max(maxTiles, (r[k].n + kScanTile - 1) / kScanTile);
	}
	k_scan_offs3<<<dim3((unsigned)maxTiles, (unsigned)count), kScanThreads, 0, st>>>(io3);
	launches++;
}

constexpr unsigned kRayOverflowCap = 1u << 20; // sub-columns handed to k_l2_rays_overflow per call (16 MB of entries; cessna-256 needs ~300)
constexpr int kMaxChunks = 16; // Level-2 chunks whose D2H copies overlap the next chunk's refinement (host sink only)

static int voxelize_body(gpv_ctx* c, const float* d_tris, int64_t n_tri, const float bmin[3], const float bmax[3], float max_model_size,
                         const gpv_params* prm, void* stream, gpv_result* out, const gpv_host_streams* sink);

static int voxelize_impl(gpv_ctx* c, const float* d_tris, int64_t n_tri, const float bmin[3], const float bmax[3], float max_model_size,
                         const gpv_params* prm, void* stream, gpv_result* out, const gpv_host_streams* sink)
{
	const int rc = voxelize_body(c, d_tris, n_tri, bmin, bmax, max_model_size, prm, stream, out, sink);
	if (rc && c) {
		// An error return must not leave work in flight on the context's own streams: it would still be reading pooled buffers, or
		// writing the caller's host memory, when the next call (or the caller) reuses them.
		const std::string why = g_err;
		cudaSetDevice(c->device);
		cudaStreamSynchronize((cudaStream_t)stream); cudaStreamSynchronize(c->sideStream); cudaStreamSynchronize(c->copyStream);
		cudaGetLastError();
		g_err = why;
	}
	return rc;
}

static int voxelize_body(gpv_ctx* c, const float* d_tris, int64_t n_tri, const float bmin[3], const float bmax[3], float max_model_size,
                         const gpv_params* prm, void* stream, gpv_result* out, const gpv_host_streams* sink)
{
	if (!c) return fail("gpv_voxelize_device: null ctx");
	if (!d_tris || !bmin || !bmax || !prm || !out) return fail("gpv_voxelize_device: null argument");
	if (n_tri <= 0 || n_tri > 0x7fffffff) return fail("gpv_voxelize_device: triangle count out of range");
	memset(out, 0, sizeof *out);
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	const bool wantL2 = !(prm->flags & GPV_NO_LEVEL2) && prm->voxel_count2 > 0;
	const bool wantN = (prm->flags & GPV_NORMALS) != 0;
	const bool gather = (prm->flags & GPV_GATHER) != 0;
	const bool wantSolid = (prm->flags & GPV_COLLISION) != 0;
	c->last.valid = false;
	if (gather && !c->gather.on) return fail("GPV_GATHER without gpv_gather_attach");
	if (gather && sink) return fail("GPV_GATHER does not support the host-stream call");
	if (gather && wantN && !c->gather.l1n) return fail("GPV_GATHER with GPV_NORMALS: the gather buffers were created without normal streams (gpv_gather_create_ex)");
	if (gpv_make_grid(bmin, bmax, max_model_size, prm->voxel_count, wantL2 ? prm->voxel_count2 : 1, &out->grid)) return 1;
	const gpv_grid& gg = out->grid;
	if (gg.n2 > 32) return fail("voxel_count2 > 32 is not supported (Level-2 z parity is kept in one 32-bit word)");
	GridP g{};
	g.nx = gg.num_div[0]; g.ny = gg.num_div[1]; g.nz = gg.num_div[2];
	g.z0 = 0; g.z1 = g.nz;
	if (prm->z1 > 0) { g.z0 = prm->z0; g.z1 = prm->z1; }
	if (g.z0 < 0 || g.z1 > g.nz || g.z0 >= g.z1) return fail("gpv_voxelize_device: bad z-slab");
	// GPV_GATHER: every rank runs Level 1 over the WHOLE grid (global boundary ranks, whole column lists); [oz0,oz1) is only the slab of
	// Level-1 bytes / prefix sums this rank delivers (default: an equal share of the layers), the Level-2 work is shared out by column (Own)
	int oz0 = g.z0, oz1 = g.z1;
	Own own{ 1, 0, 1, 1 };
	if (!gather && c->debugOwnWorld > 1) { // profiling aid (GPV_DEBUG_OWN=world,rank): a plain call refines only the Level-2 share rank `rank` of `world` would
		own.world = c->debugOwnWorld; own.rank = c->debugOwnRank; own.group = std::max(1, 256 / (gg.n2 * gg.n2)); own.nx = g.nx;
	}
	if (gather) {
		if (prm->z1 <= 0) { oz0 = (int)((long long)g.nz * c->gather.rank / c->gather.world); oz1 = (int)((long long)g.nz * (c->gather.rank + 1) / c->gather.world); }
		g.z0 = 0; g.z1 = g.nz;
		own.world = c->gather.world; own.rank = c->gather.rank; own.group = std::max(1, 256 / (gg.n2 * gg.n2)); own.nx = g.nx; // = the columns one CTA of k_l2_rays walks
	}
	g.minx = bmin[0]; g.miny = bmin[1]; g.minz = bmin[2]; g.maxx = bmax[0]; g.maxy = bmax[1]; g.maxz = bmax[2];
	g.gsx = gg.grid_size[0]; g.gsy = gg.grid_size[1]; g.gsz = gg.grid_size[2];
	g.h1x = gg.ext1[0]; g.h1y = gg.ext1[1]; g.h1z = gg.ext1[2];
	g.h2x = gg.ext2[0]; g.h2y = gg.ext2[1]; g.h2z = gg.ext2[2];
	g.n2 = gg.n2;
	const long long ncol = (long long)g.nx * g.ny;
	const long long cells = ncol * (g.z1 - g.z0);
	if (ncol * g.nz > 0x7fffffffLL) return fail("grid exceeds 2^31 cells: boundary_index is int32 (file contract); shard finer");
	if (g.nx > 32767 || g.ny > 32767 || g.nz > 32767) return fail("grid axis exceeds 32767 cells (footprints are packed in 16-bit fields)");
	if (gather && ncol * g.nz != c->gather.cellsTotal) return fail("GPV_GATHER: the gather buffers were created for a different grid");
	if (gather && oz1 > oz0 && ((long long)oz0 * ncol) % 8 != 0) return fail("GPV_GATHER: slab offset is not a multiple of 8 cells"); // cannot happen: nx, ny are multiples of 4
	const unsigned epoch = gather ? ++c->gather.epoch : 0; // every rank counts its GPV_GATHER calls: same order on all ranks
	const int nTri = (int)n_tri;
	int64_t launches = 0;
	const bool profLight = (prm->flags & GPV_PROFILE_L2) != 0 && !(prm->flags & GPV_PROFILE); // events around the two Level-2 kernels only
	const bool prof = (prm->flags & (GPV_PROFILE | GPV_PROFILE_L2)) != 0;
	if (prof && !c->haveEvents) {
		for (cudaEvent_t& e : c->ev) GPV_CUDA(cudaEventCreate(&e));
		for (cudaEvent_t& e : c->evEnd) GPV_CUDA(cudaEventCreate(&e));
		c->haveEvents = true;
	}
	// phase boundaries: ev[k] is recorded when phase k starts, ev[GPV_PHASE_COUNT] when the last one ends
	// (phases of the side branch carry their own end event: they overlap the main branch)
	bool marked[GPV_PHASE_COUNT + 1] = {}, sidePhase[GPV_PHASE_COUNT + 1] = {};
	// A small model (a dataset block: 10^5 cells, a few thousand triangles) gains nothing from the side / copy streams -- its kernels are
	// launch-latency-sized -- but pays for every fork and join: a dozen driver calls of ~45 per model, and with sixteen host threads feeding
	// one GPU (gpv_voxelize_batch) the driver calls are what bounds the throughput.  Such a call runs on the caller's stream alone.
	const bool serial = !gather && !prof && cells <= (1ll << 20) && nTri <= 200000;
	cudaStream_t side = serial ? st : c->sideStream, copy = serial ? st : c->copyStream;
	auto ev_record = [&](cudaEvent_t e, cudaStream_t s) { return serial ? cudaSuccess : cudaEventRecord(e, s); };
	auto ev_wait = [&](cudaStream_t s, cudaEvent_t e) { return serial ? cudaSuccess : cudaStreamWaitEvent(s, e, 0); };
	auto mark = [&](int phase) { if (prof && (!profLight || phase >= GPV_PHASE_L2_RAYS)) { cudaEventRecord(c->ev[phase], st); marked[phase] = true; } };
	auto mark_side = [&](int phase, bool begin) { if (prof && !profLight) { cudaEventRecord(begin ? c->ev[phase] : c->evEnd[phase], side); marked[phase] = sidePhase[phase] = true; } };

	// ---- fixed-size buffers
	if (c->tri48.ensure((size_t)nTri * 48) || c->ray48.ensure((size_t)nTri * 48) || c->tabX.ensure((size_t)g.nx * 4) || c->tabY.ensure((size_t)g.ny * 4) ||
	    c->tabZ.ensure((size_t)g.nz * 4) || c->cellCount.ensure((size_t)cells * 4 + 32) || c->colCount.ensure((size_t)ncol * 4 + 32) ||
	    c->crossCount.ensure((size_t)ncol * 4 + 32) || c->prefix.ensure((size_t)(cells + 1) * 4 + 32) || c->bmask.ensure((size_t)cells / 8 + 64) ||
	    c->boundaryIndex.ensure((size_t)cells * 4 + 32) || c->bTriOff.ensure((size_t)(cells + 1) * 4 + 32) || c->colOff.ensure((size_t)(ncol + 1) * 4 + 32) ||
	    c->crossOff.ensure((size_t)(ncol + 1) * 4 + 32) || c->l1State.ensure((size_t)cells + 32) || c->crossFp.ensure((size_t)nTri * 16) || c->plane16.ensure((size_t)nTri * 16) || c->aabbxy16.ensure((size_t)nTri * 16) ||
	    c->binCnt.ensure((size_t)nTri * 4 + 32) || c->binOff.ensure((size_t)(nTri + 1) * 4 + 32) || c->crossCnt.ensure((size_t)nTri * 4 + 32) ||
	    c->crossWorkOff.ensure((size_t)(nTri + 1) * 4 + 32) || c->colCellCnt.ensure((size_t)ncol * 4 + 32) || c->colCursor.ensure((size_t)ncol * 4 + 32) || c->colCellOff.ensure((size_t)(ncol + 1) * 4 + 32))
		return 1;
	const long long rayCap = ncol + cells / kRayChunk + 64; // units of k_l2_rays: at most one per column that has boundary cells + one per kRayChunk cells
	if (wantL2 && (c->rayOver.ensure((size_t)rayCap * 8) || c->rayOverflow.ensure((size_t)kRayOverflowCap * 16))) return 1;
	// (triangle, column) bitmaps of the two binning sweeps: one bit per work item at most.  The work-space size is known on the
	// device only; a first guess here, the exact size after the read-back (then the call starts over, once per growth).
	if (!c->binBits.cap && c->binBits.ensure((size_t)std::max<long long>(1 << 20, 64ll * nTri) / 8 * 2 + 64)) return 1;
	const unsigned long long bitsCap = ((unsigned long long)(c->binBits.cap - 64) / 2 / 4) * 32; // bits per sweep (whole words)
	if (c->cleanBits != c->binBits.p) { // a new pool, or a call that did not get to wipe what it used: zero all of it once
		GPV_CUDA(cudaMemsetAsync(c->binBits.p, 0, c->binBits.cap, st));
	}
	c->cleanBits = nullptr; // dirty from here on; restored when the call completes
	Totals* dT = c->totals.as<Totals>();
	mark(GPV_PHASE_SETUP);
	// look-back descriptor regions of the six scans of this call
	size_t dOff[7];
	{
		const long long ns[6] = { nTri, nTri, cells, ncol, ncol, ncol };
		dOff[0] = 0;
		for (int k = 0; k < 6; k++) dOff[k + 1] = dOff[k] + desc_bytes(ns[k]);
		if (c->desc.ensure(dOff[6] + 32)) return 1;
	}
	{ // one launch zeroes every counter of the call (instead of a dozen memsets)
		ClearList cl{};
		int k = 0;
		auto add = [&](void* p, size_t bytes) { cl.p[k] = reinterpret_cast<uint4*>(p); cl.n16[k] = (bytes + 15) / 16; k++; };
		add(dT, 256);                                   // Totals + the two long-list counters of the sort (gpv_ctx::totals is 256 B)
		if (!(c->cleanCellCount == c->cellCount.p && c->cleanCellBytes >= (size_t)cells * 4)) add(c->cellCount.p, (size_t)cells * 4);
		c->cleanCellCount = nullptr; c->cleanCellBytes = 0; // dirty from here on; restored when the call completes
		add(c->colCount.p, (size_t)ncol * 4);
		add(c->crossCount.p, (size_t)ncol * 4);
		add(c->colCellCnt.p, (size_t)ncol * 4);
		add(c->colCursor.p, (size_t)ncol * 4);
		add(c->desc.p, dOff[6]);
		k_clear<<<dim3((unsigned)std::min<long long>(c->smCount * 8, (cells * 4 / 16 + 255) / 256 + 1), (unsigned)k), 256, 0, st>>>(cl);
		launches++;
	}

	float *cx = c->tabX.as<float>(), *cy = c->tabY.as<float>(), *cz = c->tabZ.as<float>();
	float4 *tri48 = c->tri48.as<float4>(), *ray48 = c->ray48.as<float4>();
	{
		int m = g.nx > g.ny ? g.nx : g.ny; m = m > g.nz ? m : g.nz;
		m = m > nTri ? m : nTri; // k_prepare also fills the per-axis centre tables
		k_prepare<<<(m + 255) / 256, 256, 0, st>>>(d_tris, nTri, g, tri48, ray48, c->plane16.as<float4>(), c->aabbxy16.as<float4>(), c->crossFp.as<int4>(), c->binCnt.as<int>(), c->crossCnt.as<int>(), dT, cx, cy, cz);
		launches++;
	}
	// balanced work spaces: exclusive scans of the per-triangle item counts; totals stay on the device (persistent grids read them)
	{
		const ScanReq r[2] = { { c->binCnt.as<int>(), nTri, c->binOff.as<unsigned>(), nullptr, &dT->binWork, dOff[0] },
			                   { c->crossCnt.as<int>(), nTri, c->crossWorkOff.as<unsigned>(), nullptr, &dT->crossWork, dOff[1] } };
		launch_scans(c, st, r, 2, launches);
	}
	GPV_CUDA(ev_record(c->evFork[0], st)); // the work spaces are scanned: the crossing count may start (side stream, launched below)
	BinOut bo{};
	bo.cellCount = c->cellCount.as<int>(); bo.colCount = c->colCount.as<int>(); bo.totals = dT;
	bo.bits = c->binBits.as<unsigned>(); bo.bitsCap = bitsCap;
	// (the main stream's kernels are enqueued first: it carries the critical path, the side stream's work hides beside it)
	mark(GPV_PHASE_BIN_COUNT);
	k_bin<false><<<kWorkGrid, kWorkThreads, 0, st>>>(tri48, nTri, c->binOff.as<unsigned>(), g, cx, cy, cz, bo);
	launches++;
	mark(GPV_PHASE_SCAN);
	{ // K3 boundary compaction
		const int V = cells > (1ll << 20) ? 4 : 1; // sub-tiles per tile: beyond a million cells the per-tile latency is amortised over 32 KB of input
		long long tiles = (cells + (long long)V * kScanTile - 1) / ((long long)V * kScanTile);
		ScanIO io{};
		io.in = c->cellCount.as<int>(); io.n = cells;
		io.tileCounter = reinterpret_cast<unsigned*>(c->desc.as<char>() + dOff[2]); io.desc = reinterpret_cast<unsigned long long*>(c->desc.as<char>() + dOff[2]) + 1;
		io.prefix = c->prefix.as<int>(); io.boundaryIndex = c->boundaryIndex.as<int>(); io.bTriOff = c->bTriOff.as<unsigned>();
		io.bmask = c->bmask.as<unsigned char>(); io.globalBase = (long long)g.z0 * ncol; io.totals = dT;
		io.colCells = c->colCellCnt.as<int>(); io.plane = ncol; io.own = own;
		if (V == 4) k_scan<MODE_CELLS, 4><<<(unsigned)tiles, kScanThreads, 0, st>>>(io);
		else k_scan<MODE_CELLS, 1><<<(unsigned)tiles, kScanThreads, 0, st>>>(io);
		launches++;
	}
	// the two count sweeps are independent: the crossing count runs on the side stream beside the SAT count and the cell scan
	GPV_CUDA(ev_wait(side, c->evFork[0]));
	mark_side(GPV_PHASE_CROSS_COUNT, true);
	k_cross<false><<<kWorkGrid, kWorkThreads, 0, side>>>(ray48, c->crossFp.as<int4>(), nTri, c->crossWorkOff.as<unsigned>(), g, cx, cy, c->crossCount.as<int>(), nullptr, nullptr, dT);
	launches++;
	mark_side(GPV_PHASE_CROSS_COUNT, false);
	GPV_CUDA(ev_record(c->evJoin[0], side));
	GPV_CUDA(ev_wait(st, c->evJoin[0]));
	{
		const ScanReq r[3] = { { c->colCount.as<int>(), ncol, c->colOff.as<unsigned>(), &dT->colTotalOver, nullptr, dOff[3] },
			                   { c->crossCount.as<int>(), ncol, c->crossOff.as<unsigned>(), &dT->crossTotal, nullptr, dOff[4] },
			                   { c->colCellCnt.as<int>(), ncol, c->colCellOff.as<unsigned>(), &dT->nLocalCells, nullptr, dOff[5] } };
		launch_scans(c, st, r, wantL2 ? 3 : 2, launches);
		if (wantL2) { // the work list of the Level-2 parity rays: (column, chunk of its boundary cells), expensive units first
			k_ray_units<<<(unsigned)((ncol + 255) / 256), 256, 0, st>>>(c->colCellOff.as<unsigned>(), c->colCount.as<int>(), ncol, c->rayOver.as<int2>(), rayCap, dT);
			launches++;
		}
	}

	// ---- the one size read-back
	mark(GPV_PHASE_HOST_GAP);
	GPV_CUDA(cudaMemcpyAsync(c->hTotals, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
	GPV_CUDA(cudaStreamSynchronize(st));
	GPV_CUDA(cudaGetLastError());
	const Totals T1 = *c->hTotals;
	if (T1.l1Hits > 0x7ffffff0ull) return fail("more than 2^31 (cell, triangle) pairs in one slab: shard finer");
	if (T1.binWork > 0xfffffff0ull || T1.crossWork > 0xfffffff0ull) return fail("more than 2^32 (triangle, cell) work items: work offsets are 32-bit");
	if (T1.binWork > bitsCap) { // the binning sweeps left without doing anything: grow the bitmaps and start the call over (first call of a larger model only)
		if (c->binBits.ensure((size_t)((T1.binWork + 31) / 32 * 4) * 2 + 64)) return 1; // (a new pool: zeroed by the repeated call)
		if (gather) c->gather.epoch--; // nothing of this call has reached the mailbox yet: the repeated call takes the same epoch
		return voxelize_impl(c, d_tris, n_tri, bmin, bmax, max_model_size, prm, stream, out, sink);
	}
	const long long nB = T1.nBoundary;
	const long long n23 = (long long)g.n2 * g.n2 * g.n2;
	if (c->cellTris.ensure((size_t)T1.triTotal * 4 + 32) || c->colTris.ensure((size_t)T1.colTotalOver * 4 + 32) || c->crossTri.ensure((size_t)T1.crossTotal * 4 + 32))
		return 1;
	if (wantL2 && (c->l2State.ensure(gather ? 32 : (size_t)(nB * n23) + 32) || c->colCellList.ensure((size_t)nB * 8 + 32) || c->cellMid.ensure((size_t)nB * 16 + 32) || c->l2Par.ensure((size_t)nB * g.n2 * g.n2 * 4 + 32))) return 1;
	if (wantN && (c->l1Normal.ensure((size_t)cells * 3 + 32) || (wantL2 && c->l2Normal.ensure((size_t)(nB * n23) * 3 + 32)))) return 1;

	bo.prefix = c->prefix.as<int>(); bo.bTriOff = c->bTriOff.as<unsigned>(); bo.cellTris = c->cellTris.as<int>();
	bo.colOff = c->colOff.as<unsigned>(); bo.colTris = c->colTris.as<int>(); bo.colCursor = c->colCursor.as<int>();
	bo.bits = c->binBits.as<unsigned>() + bitsCap / 32; // the fill sweep's own bitmap
	if (c->longList.ensure((size_t)(nB + ncol) * 4 + 64)) return 1;
	if (wantSolid && !gather && c->solidWords.ensure((size_t)((g.z1 - g.z0 + 31) / 32) * ncol * 4 + 64)) return 1; // (every allocation of the call happens before the exchange and the fork: a growing pool's cudaFree synchronises the device)
	if (gather) {
		// every rank sees the same global boundary count: a gather buffer that is too small is refused by all of them, before anything is written
		if (wantL2 && nB * n23 > c->gather.l2Cap) return fail("GPV_GATHER: Level-2 gather buffer too small for the boundary cells of the grid");
	}
	if (sink) {
		if ((sink->level2_inout || (sink->level2_normal && wantN)) && wantL2 && (int64_t)(nB * n23) > sink->level2_capacity) return fail("gpv_voxelize_host: level2 host buffer too small");
		if (sink->boundary_index && nB > sink->boundary_capacity) return fail("gpv_voxelize_host: boundary_index host buffer too small");
	}
	// main stream first (it carries the critical path): the fill sweep of the binning
	mark(GPV_PHASE_BIN_FILL);
	k_bin<true><<<kWorkGrid, kWorkThreads, 0, st>>>(tri48, nTri, c->binOff.as<unsigned>(), g, cx, cy, cz, bo);
	GPV_CUDA(ev_record(c->evBinDone, st)); // both bitmaps are dead now: the side stream wipes what this call used (below)
	launches++;
	// side stream: the parity-fill branch (crossing lists -> fill sweep -> final Level-1 bytes), independent of the binning / sorting /
	// Level-2 branch on the main stream until the end of the call.  (No fork event: the host has just synchronised the main stream.)
	if (gather) { // peers wait here until rank 0 has entered this call: nothing is stored into its buffers before
		k_gather_begin<<<1, 1, 0, side>>>(c->gather.mail, c->gather.rank, epoch, c->gather.timeoutNs, dT);
		GPV_CUDA(ev_record(c->evFork[1], side)); // the main stream's Level-2 stores wait for it too (below)
		launches++;
	}
	mark_side(GPV_PHASE_CROSS_FILL, true);
	k_cross<true><<<kWorkGrid, kWorkThreads, 0, side>>>(ray48, c->crossFp.as<int4>(), nTri, c->crossWorkOff.as<unsigned>(), g, cx, cy, c->crossCount.as<int>(), c->crossOff.as<unsigned>(),
	                                                     c->crossTri.as<int>(), dT);
	mark_side(GPV_PHASE_CROSS_FILL, false);
	mark_side(GPV_PHASE_FILL_SWEEP, true);
	{
		GridP gs = g; // the slab of Level-1 bytes this call delivers ([oz0,oz1) with GPV_GATHER, the call's slab otherwise)
		gs.z0 = oz0; gs.z1 = oz1;
		const size_t slabOff = gather ? (size_t)oz0 * ncol : 0; // offset of the slab inside the whole-grid arrays of a gathering call
		dim3 grid((g.nx + 31) / 32, g.ny, (oz1 - oz0 + 127) / 128), block(32, 4);
		k_fill_sweep<<<grid, block, 0, side>>>(ray48, gs, cx, cy, cz, c->crossOff.as<unsigned>(), c->crossTri.as<int>(), c->bmask.as<unsigned char>() + slabOff / 8,
		                                       c->l1State.as<unsigned char>() + slabOff, dT, (wantSolid && !gather) ? c->solidWords.as<unsigned>() : nullptr);
		launches += 2;
		if (gather) {
			// This slab of the Level-1 bytes and of the (global) prefix sums -> their final place on the gathering rank: two contiguous
			// ranges, moved by the copy engines over NVLink beside the Level-2 kernels.  (Stored from inside the fill sweep -- 32-byte
			// writes over NVLink -- the slab took 3x as long and the SMs' store queues held up the parity-ray kernel running beside it.)
			const size_t n = (size_t)(oz1 - oz0) * ncol;
			GPV_CUDA(cudaMemcpyAsync(c->gather.l1 + slabOff, c->l1State.as<unsigned char>() + slabOff, n, cudaMemcpyDefault, side));
			GPV_CUDA(cudaMemcpyAsync(c->gather.prefix + slabOff, c->prefix.as<int>() + slabOff, n * 4, cudaMemcpyDefault, side));
		}
	}
	mark_side(GPV_PHASE_FILL_SWEEP, false);
	if (sink) { // host call: the Level-1 streams are final here and leave on this stream, beside the Level-2 chunks on the copy stream
		if (sink->level1_inout) GPV_CUDA(cudaMemcpyAsync(sink->level1_inout, c->l1State.p, (size_t)cells, cudaMemcpyDeviceToHost, side));
		if (sink->prefix) GPV_CUDA(cudaMemcpyAsync(sink->prefix, c->prefix.p, (size_t)cells * 4, cudaMemcpyDeviceToHost, side));
		if (sink->boundary_index && nB) GPV_CUDA(cudaMemcpyAsync(sink->boundary_index, c->boundaryIndex.p, (size_t)nB * 4, cudaMemcpyDeviceToHost, side));
	}
	GPV_CUDA(ev_wait(side, c->evBinDone));
	k_clear_bits<<<c->smCount * 4, 256, 0, side>>>(c->binBits.as<unsigned>(), bitsCap, dT); // the next call finds the bitmaps zeroed
	launches++;
	GPV_CUDA(ev_record(c->evJoin[1], side));
	mark(GPV_PHASE_SORT);
	if (wantN || (prm->flags & GPV_KEEP_LISTS)) {
		// canonical (ascending) order of the cell and column lists: needed only where the order shows -- the f32 sums of the normals and
		// lists handed to the caller.  Occupancy does not depend on it (the SAT ORs its hits, the parity rays XOR theirs), and the
		// column lists are duplicate-free by construction.  Warp per list; lists longer than kSortSmem go through a work list to
		// k_sort_long (CTA per list).
		unsigned* longCnt = reinterpret_cast<unsigned*>(c->totals.as<char>() + 128); // [0] cell lists, [1] column lists (zeroed by k_clear)
		int* longCells = c->longList.as<int>() + 16;
		int* longCols = longCells + nB;
		if (!c->sortAttrSet) { // function attributes are per device: once per context, not once per process
			GPV_CUDA(cudaFuncSetAttribute(k_sort_long, cudaFuncAttributeMaxDynamicSharedMemorySize, kSortLongSmem * 4));
			c->sortAttrSet = true;
		}
		const SortSeg sc = { c->bTriOff.as<unsigned>(), (int)nB, c->cellTris.as<int>(), nullptr, longCells, longCnt };
		const SortSeg sk = { c->colOff.as<unsigned>(), (int)ncol, c->colTris.as<int>(), c->colCount.as<int>(), longCols, longCnt + 1 };
		const long long segs = std::max<long long>(nB, ncol);
		k_sort_segments<<<dim3((unsigned)((segs * 32 + 255) / 256), 2), 256, 0, st>>>(sc, sk);
		k_sort_long<<<dim3(c->smCount, 2), kSortLongThreads, kSortLongSmem * 4, st>>>(sc, sk);
		launches += 2;
	}
	mark(GPV_PHASE_L1_NORMALS);
	if (wantN) {
		GPV_CUDA(cudaMemsetAsync(c->l1Normal.p, 127, (size_t)cells * 3, st));
		if (nB > 0) {
			k_l1_normals<<<(unsigned)((nB + 127) / 128), 128, 0, st>>>(tri48, c->boundaryIndex.as<int>(), c->bTriOff.as<unsigned>(), c->cellTris.as<int>(), (int)nB,
			                                                         (long long)g.z0 * ncol, c->l1Normal.as<unsigned char>());
			launches++;
		}
		if (gather) { // this rank's z-slab of the Level-1 normals -> rank 0 (copy engine; peers store into rank 0 only once it has entered the call)
			GPV_CUDA(ev_wait(st, c->evFork[1]));
			GPV_CUDA(cudaMemcpyAsync(c->gather.l1n + (size_t)oz0 * ncol * 3, c->l1Normal.as<unsigned char>() + (size_t)oz0 * ncol * 3, (size_t)(oz1 - oz0) * ncol * 3, cudaMemcpyDefault, st));
		}
	}
	if (sink && sink->level1_normal && wantN) { // (the other Level-1 streams left from the side stream, right behind the fill sweep)
		GPV_CUDA(ev_record(c->evChunk[kMaxChunks], st));
		GPV_CUDA(ev_wait(copy, c->evChunk[kMaxChunks]));
		GPV_CUDA(cudaMemcpyAsync(sink->level1_normal, c->l1Normal.p, (size_t)cells * 3, cudaMemcpyDeviceToHost, copy));
	}
	L2IO lio{};
	long long packedChunks = 0, packedPer = 0, packedCells = 0; // GPV_PACKED_L2: chunks of 2-bit words on their way to hPacked
	mark(GPV_PHASE_L2_RAYS);
	if (wantL2 && nB > 0) {
		lio.tri48 = tri48; lio.ray48 = ray48; lio.plane16 = c->plane16.as<float4>(); lio.aabbxy16 = c->aabbxy16.as<float4>(); lio.boundaryIndex = c->boundaryIndex.as<int>(); lio.bTriOff = c->bTriOff.as<unsigned>();
		lio.cellTris = c->cellTris.as<int>(); lio.colOff = c->colOff.as<unsigned>(); lio.colCount = c->colCount.as<int>(); lio.colTris = c->colTris.as<int>();
		lio.cx = cx; lio.cy = cy; lio.cz = cz; lio.nBoundary = (int)nB; lio.totals = dT;
		lio.l2State = gather ? c->gather.l2 : c->l2State.as<unsigned char>(); lio.own = own;
		lio.colCellOff = c->colCellOff.as<unsigned>(); lio.colCellList = c->colCellList.as<int2>(); lio.l2Par = c->l2Par.as<unsigned>(); lio.cellMid = c->cellMid.as<float4>();
		const L2K K = l2_constants(g.n2);
		const int G = K.G;
		const bool canPack = (n23 % 32) == 0; // the blocks can travel as 2 bits per sub-voxel
		const size_t smem = (size_t)K.total;
		// how the blocks leave k_l2 (L2_OUT_*): file bytes into local HBM; or 2 bits per sub-voxel -- to the gathering rank over NVLink
		// (peers of a GPV_GATHER call; rank 0 expands them after the wait) or into the staging buffer of the host call (GPV_PACKED_L2;
		// host threads expand); or, when n2^3 is not a multiple of 32, staged bytes for the peers of a gathering call
		const bool packHost = sink && sink->level2_inout && (prm->flags & GPV_PACKED_L2) && canPack;
		const bool packPeer = gather && c->gather.rank != 0 && canPack;
		// (a peer that also computes normals keeps its blocks: they are written locally and copied to rank 0 by k_scatter_blocks)
		const bool peer = gather && c->gather.rank != 0, peerLocal = peer && wantN;
		const int outMode = (packHost || packPeer) ? L2_OUT_PACKED : (peer && !peerLocal) ? L2_OUT_STAGED : L2_OUT_BYTES;
		typedef void (*L2Fn)(GridP, L2IO, L2K);
		static const L2Fn table[3][5] = { { k_l2<16, 0>, k_l2<8, 0>, k_l2<4, 0>, k_l2<2, 0>, k_l2<0, 0> },
			                              { k_l2<16, 1>, k_l2<8, 1>, k_l2<4, 1>, k_l2<2, 1>, k_l2<0, 1> },
			                              { k_l2<16, 2>, k_l2<8, 2>, k_l2<4, 2>, k_l2<0, 2>, k_l2<0, 2> } }; // (n2 = 2 cannot be packed: canPack is false)
		const L2Fn l2fn = table[outMode][g.n2 == 16 ? 0 : g.n2 == 8 ? 1 : g.n2 == 4 ? 2 : g.n2 == 2 ? 3 : 4];
		if (peerLocal && (packPeer ? c->l2Packed.ensure((size_t)(nB * n23) / 4 + 64) : c->l2State.ensure((size_t)(nB * n23) + 64))) return 1;
		if (peerLocal && !packPeer) lio.l2State = c->l2State.as<unsigned char>();
		if (packHost) {
			if (c->l2Packed.ensure((size_t)(nB * n23) / 4 + 64)) return 1;
			if (c->hPackedCap < (size_t)(nB * n23) / 4) {
				if (c->hPacked) cudaFreeHost(c->hPacked);
				c->hPacked = nullptr; c->hPackedCap = 0;
				const size_t want = (size_t)(nB * n23) / 4 + (size_t)(nB * n23) / 32 + 4096;
				GPV_CUDA(cudaHostAlloc(&c->hPacked, want, cudaHostAllocDefault));
				c->hPackedCap = want;
			}
		}
		lio.l2Packed = (packHost || (packPeer && peerLocal)) ? c->l2Packed.as<unsigned char>() : packPeer ? c->gather.l2p : nullptr;
		if (smem > 48 * 1024) return fail("k_l2: shared-memory layout exceeds 48 KB"); // cannot happen for n2 <= 32 (41 KB)
		// K4a: boundary cells grouped by Level-1 column, then the parity bits of every sub-voxel column, one walk of the column list per column
		// gathering rank: the cells of every peer, for the peer-by-peer expansion of their 2-bit blocks (counters in the totals block: zeroed by k_clear)
		const bool peerLists = gather && c->gather.rank == 0 && c->gather.world > 1 && canPack;
		unsigned* peerCount = reinterpret_cast<unsigned*>(c->totals.as<char>() + 144);
		if (peerLists && c->peerCells.ensure((size_t)c->gather.world * nB * 4 + 32)) return 1;
		const PeerLists pl{ peerLists ? c->peerCells.as<int>() : nullptr, peerCount, nB };
		k_col_cells<<<(unsigned)((nB + 255) / 256), 256, 0, st>>>(lio.boundaryIndex, (int)nB, (int)ncol, g.nx, cx, cy, cz, lio.colCellOff, c->colCellCnt.as<int>(), c->colCellList.as<int2>(),
		                                                        c->cellMid.as<float4>(), lio.colCount, lio.bTriOff, own, pl, dT);
		const long long nUnits = (long long)T1.nRayHeavy + T1.nRayLight;
		RayWork rw{ c->rayOver.as<int2>(), rayCap, dT, { c->rayOverflow.as<int4>(), reinterpret_cast<unsigned*>(c->totals.as<char>() + 128) + 2, kRayOverflowCap } }; // (counter zeroed by k_clear)
		if (nUnits > 0) {
			k_l2_rays<<<(unsigned)((nUnits + G - 1) / G), 256, 0, st>>>(g, lio, rw);
			k_l2_rays_overflow<<<c->smCount, 256, 0, st>>>(g, lio, rw.ov); // the few sub-columns with more crossings than register slots
			launches += 2;
		}
		launches++;
		mark(GPV_PHASE_L2);
		// With a host sink the boundary cells are refined in chunks and every finished chunk's bytes start their way to the
		// host on the copy stream while the next chunk computes (e2e is PCIe-bound: 1 B per Level-2 voxel).
		long long chunks = 1;
		if (sink && sink->level2_inout) chunks = std::min<long long>(kMaxChunks, std::max<long long>(1, (nB * n23 + (24ll << 20) - 1) / (24ll << 20)));
		if (packHost) chunks = std::min<long long>(kMaxChunks, std::max<long long>(1, (nB * n23 + (12ll << 20) - 1) / (12ll << 20))); // 3 MB on the bus, 12 MB for the host threads per chunk
		// the cells this call refines: all boundary ranks, or (GPV_GATHER) this rank's share as listed, column by column, in colCellList
		if (gather) GPV_CUDA(ev_wait(st, c->evFork[1])); // rank 0 has entered the call: its Level-2 buffer may be written
		const bool byList = gather || own.world > 1;
		const long long nRefine = byList ? (long long)T1.nLocalCells : nB;
		lio.cellList = byList ? c->colCellList.as<int2>() : nullptr;
		const long long per = ((nRefine + chunks - 1) / chunks + G - 1) / G * G; // whole CTAs per chunk
		long long nChunks = 0;
		for (long long k = 0; k < chunks; k++) {
			const long long bb = k * per, be = std::min(nRefine, bb + per);
			if (bb >= be) break;
			nChunks = k + 1;
			lio.bBegin = (int)bb; lio.nBoundary = (int)be;
			l2fn<<<(unsigned)((be - bb + G - 1) / G), kL2Threads, smem, st>>>(g, lio, K);
			launches++;
			if (sink && sink->level2_inout) {
				GPV_CUDA(ev_record(c->evChunk[k], st));
				GPV_CUDA(ev_wait(copy, c->evChunk[k]));
				if (packHost) {
					GPV_CUDA(cudaMemcpyAsync((char*)c->hPacked + bb * n23 / 4, c->l2Packed.as<uint8_t>() + bb * n23 / 4, (size_t)((be - bb) * n23 / 4), cudaMemcpyDeviceToHost, copy));
					GPV_CUDA(cudaEventRecord(c->evCopied[k], copy));
				} else
					GPV_CUDA(cudaMemcpyAsync(sink->level2_inout + bb * n23, c->l2State.as<uint8_t>() + bb * n23, (size_t)((be - bb) * n23), cudaMemcpyDeviceToHost, copy));
			}
		}
		if (packHost) { packedChunks = nChunks; packedPer = per; packedCells = nRefine; }
		if (peerLocal && nRefine > 0) { // the blocks went to local memory (the normals read them back): their copies to rank 0
			const int per = (int)(packPeer ? n23 / 4 : n23);
			k_scatter_blocks<<<(unsigned)std::min<long long>(c->smCount * 8, (nRefine * (per % 8 ? per : per / 8) + 255) / 256), 256, 0, st>>>(packPeer ? c->l2Packed.as<unsigned char>() : c->l2State.as<unsigned char>(),
			                                                                                                    packPeer ? c->gather.l2p : c->gather.l2, c->colCellList.as<int2>(), nRefine, per);
			launches++;
		}
		mark(GPV_PHASE_L2_NORMALS);
		if (wantN && nRefine > 0) {
			lio.bBegin = 0; lio.nBoundary = (int)nRefine; // (slots again: the cells this call refined)
			const bool fromPacked = packHost || (peerLocal && packPeer);
			k_l2_normals<<<(unsigned)((nRefine * n23 + kNormalVoxels - 1) / kNormalVoxels), 256, 0, st>>>(g, lio, fromPacked ? nullptr : lio.l2State, fromPacked ? c->l2Packed.as<uint2>() : nullptr,
			                                                                     gather ? c->gather.l2n : c->l2Normal.as<unsigned char>());
			launches++;
			if (sink && sink->level2_normal) GPV_CUDA(cudaMemcpyAsync(sink->level2_normal, c->l2Normal.p, (size_t)(nB * n23) * 3, cudaMemcpyDeviceToHost, st));
		}
		lio.cellList = nullptr;
		lio.bBegin = 0; lio.nBoundary = (int)nB;
	}
	GPV_CUDA(ev_wait(st, c->evJoin[1])); // the parity-fill branch joins
	if (gather) { // completion flag behind this rank's last store; the gathering rank returns when every rank has signalled
		k_gather_done<<<1, 1, 0, st>>>(c->gather.mail, c->gather.rank, epoch, dT);
		launches++;
		if (c->gather.rank == 0) {
			if (wantL2 && nB > 0 && c->gather.world > 1 && (n23 % 32) == 0) { // the peers sent 2 bits per sub-voxel: the file bytes of their blocks, peer by peer as they finish
				const unsigned* peerCount = reinterpret_cast<const unsigned*>(c->totals.as<char>() + 144);
				const long long perPeer = (nB / c->gather.world + 1) * (n23 / 32); // grid: about one thread per word of a peer's share
				for (int q = 1; q < c->gather.world; q++) {
					k_gather_wait_rank<<<1, 1, 0, st>>>(c->gather.mail, q, epoch, c->gather.timeoutNs, dT);
					k_gather_expand<<<(unsigned)std::min<long long>(c->smCount * 8, (perPeer + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint2*>(c->gather.l2p), c->gather.l2, (int)(n23 / 32),
					                                                                                                        c->peerCells.as<int>() + (size_t)q * nB, peerCount + q);
					launches += 2;
				}
			}
			k_gather_wait<<<1, 1, 0, st>>>(c->gather.mail, c->gather.world, epoch, c->gather.timeoutNs, dT);
			launches++;
		}
	}
	mark(GPV_PHASE_COUNT);
	GPV_CUDA(cudaMemcpyAsync(c->hTotals, dT, 144, cudaMemcpyDeviceToHost, st)); // Totals + the sort's and the ray overflow's counters
	if (packedChunks) {
		// Every launch and copy of the call is queued: the calling thread now follows the packed Level-2 chunks as they land and hands
		// each to the host threads, which write the caller's byte stream while the next chunk is on the bus.
		ExpandPool* pool = expand_pool_get();
		expand_pool_begin(pool);
		for (long long k = 0; k < packedChunks; k++) {
			const long long bb = k * packedPer, be = std::min(packedCells, bb + packedPer);
			const cudaError_t e = cudaEventSynchronize(c->evCopied[k]);
			if (e != cudaSuccess) { expand_pool_end(pool); return fail(std::string("cudaEventSynchronize: ") + cudaGetErrorString(e)); }
			expand_pool_submit(pool, (const char*)c->hPacked + bb * n23 / 4, sink->level2_inout + bb * n23, (size_t)((be - bb) * n23 / 32));
		}
		expand_pool_end(pool);
	}
	GPV_CUDA(cudaStreamSynchronize(st));
	if (sink && !serial) GPV_CUDA(cudaStreamSynchronize(c->copyStream));
	GPV_CUDA(cudaGetLastError());
	const Totals T2 = *c->hTotals;
	if (reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(c->hTotals) + 128)[2] > kRayOverflowCap)
		return fail("more than 2^20 Level-2 sub-columns with over 16 ray crossings: not supported");
	c->cleanCellCount = c->cellCount.p; c->cleanCellBytes = (size_t)cells * 4; c->cleanBits = c->binBits.p; // the call ran to its end: counters are back to zero
	if (gather && T2.gatherError) return fail("GPV_GATHER: timed out waiting for a peer rank");
	if (gather) c->gather.nbTotal = nB; // the boundary ranks are global on every rank

	c->last.valid = !gather; c->last.whole = g.z0 == 0 && g.z1 == g.nz; c->last.solid = wantSolid && !gather; c->last.g = g; c->last.cells = cells;
	out->z0 = oz0; out->z1 = oz1;
	out->cells = cells; out->n_boundary = nB; out->n23 = n23; out->n_refined = wantL2 ? (gather ? (int64_t)T2.nLocalCells : nB) : 0;
	out->d_level1_inout = c->l1State.as<uint8_t>() + (gather ? (size_t)oz0 * ncol : 0); // this call's slab of the bytes (gather: a copy went to rank 0)
	out->d_prefix = c->prefix.as<int32_t>();                                                          // the call's own sums (gather: the whole grid's)
	out->d_boundary_index = c->boundaryIndex.as<int32_t>();
	out->d_level2_inout = !wantL2 ? nullptr : gather ? c->gather.l2 : c->l2State.as<uint8_t>();
	out->d_level1_normal = wantN ? c->l1Normal.as<uint8_t>() : nullptr;
	out->d_level2_normal = (wantN && wantL2) ? (gather ? c->gather.l2n : c->l2Normal.as<uint8_t>()) : nullptr;
	out->d_cell_off = c->bTriOff.as<uint32_t>(); out->d_cell_tris = c->cellTris.as<int32_t>();
	out->d_col_off = c->colOff.as<uint32_t>(); out->d_col_count = c->colCount.as<int32_t>(); out->d_col_tris = c->colTris.as<int32_t>();
	out->l1_inside = (int64_t)T2.l1Inside; out->l1_boundary = nB;
	out->l2_inside = (int64_t)T2.l2Inside; out->l2_boundary = (int64_t)T2.l2Boundary;
	out->l1_box_tests = (int64_t)T2.binWork; out->l1_box_hits = (int64_t)T2.l1Hits; // sum of clipped footprints = the reference's loop nest (cu:374-378)
	out->tri_total = T2.triTotal;
	out->l2_box_tests = wantL2 ? (int64_t)T2.l2LocalTris * n23 : 0; // reference-equivalent: n2^3 x sum of cell list lengths (cu:428) over the cells this call refined
	out->l2_ray_tests = wantL2 ? (int64_t)T2.l2ColPairs * n23 : 0; // reference-equivalent: n2^3 x sum of column list lengths (cu:461-463)
	out->fill_crossings = (int64_t)T2.crossPairs; out->fill_ill_conditioned = (int64_t)T2.nIll;
	out->kernel_launches = launches;
	if (prof) {
		for (int k = 0; k < GPV_PHASE_COUNT; k++) {
			if (!marked[k]) continue;
			float ms = 0.f;
			if (sidePhase[k]) { // side-branch phases overlap the main branch: their own start / end events
				if (cudaEventElapsedTime(&ms, c->ev[k], c->evEnd[k]) == cudaSuccess) out->phase_ms[k] = ms;
				continue;
			}
			int nxt = k + 1;
			while (nxt < GPV_PHASE_COUNT && (!marked[nxt] || sidePhase[nxt])) nxt++;
			if (cudaEventElapsedTime(&ms, c->ev[k], c->ev[nxt]) == cudaSuccess) out->phase_ms[k] = ms;
		}
	}
	return 0;
}

extern "C" int gpv_voxelize_device(gpv_ctx* c, const float* d_tris, int64_t n_tri, const float bmin[3], const float bmax[3], float max_model_size,
                                   const gpv_params* prm, void* stream, gpv_result* out)
{
	return voxelize_impl(c, d_tris, n_tri, bmin, bmax, max_model_size, prm, stream, out, nullptr);
}

extern "C" int gpv_voxelize_host(gpv_ctx* c, const gpv_mesh* mesh, const gpv_params* prm, void* stream, gpv_result* out, gpv_host_streams* h)
{
	if (!c || !mesh || !mesh->tris) return fail("gpv_voxelize_host: null argument");
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	if (c->scratch.ensure((size_t)mesh->n_tri * 36)) return 1;
	GPV_CUDA(cudaMemcpyAsync(c->scratch.p, mesh->tris, (size_t)mesh->n_tri * 36, cudaMemcpyHostToDevice, st));
	// every stream goes to the host from inside the pipeline (Level-1 streams during Level-2, Level-2 chunk by chunk)
	if (voxelize_impl(c, c->scratch.as<float>(), mesh->n_tri, mesh->bbox_min, mesh->bbox_max, mesh->max_model_size, prm, stream, out, h)) return 1;
	GPV_CUDA(cudaStreamSynchronize(st));
	return 0;
}

// ------------------------------------------------------------------------------------------------ gather over peer memory
extern "C" void gpv_gather_detach(gpv_ctx* c)
{
	if (!c || !c->gather.on) return;
	cudaSetDevice(c->device);
	if (c->gather.ipc) {
		cudaIpcCloseMemHandle(c->gather.l1); cudaIpcCloseMemHandle(c->gather.prefix); cudaIpcCloseMemHandle(c->gather.l2); cudaIpcCloseMemHandle(c->gather.mail);
	}
	// the ctx that created the buffers keeps owning them (and their sizes): it can be attached again, as rank 0 of a new session
	const bool owner = c->gather.owner;
	const int gflags = c->gather.flags;
	const int64_t cellsTotal = c->gather.cellsTotal, l2Cap = c->gather.l2Cap;
	const unsigned long long timeoutNs = c->gather.timeoutNs;
	c->gather = {};
	c->gather.timeoutNs = timeoutNs;
	if (owner) { c->gather.owner = true; c->gather.cellsTotal = cellsTotal; c->gather.l2Cap = l2Cap; c->gather.flags = gflags; }
}

extern "C" int gpv_gather_set_timeout(gpv_ctx* c, double seconds)
{
	if (!c || !(seconds > 0)) return fail("gpv_gather_set_timeout: bad argument");
	c->gather.timeoutNs = (unsigned long long)(seconds * 1e9);
	return 0;
}

static size_t gather_packed_offset(int64_t l2Cap) { return ((size_t)l2Cap + 255) & ~(size_t)255; } // the peers' 2-bit blocks live behind the byte stream
static size_t gather_l2n_offset(int64_t l2Cap) { return gather_packed_offset(l2Cap) + (((size_t)l2Cap / 4 + 255) & ~(size_t)255); } // ... and the Level-2 normals behind those
static size_t gather_l1n_offset(int64_t cells) { return ((size_t)cells + 255) & ~(size_t)255; }                                   // Level-1 normals behind the Level-1 bytes

extern "C" int gpv_gather_create(gpv_ctx* c, int64_t cells_total, int64_t l2_capacity, gpv_gather_desc* out)
{
	return gpv_gather_create_ex(c, cells_total, l2_capacity, 0, out);
}

extern "C" int gpv_gather_create_ex(gpv_ctx* c, int64_t cells_total, int64_t l2_capacity, int flags, gpv_gather_desc* out)
{
	if (!c || !out || cells_total <= 0 || l2_capacity < 0) return fail("gpv_gather_create: bad argument");
	GPV_CUDA(cudaSetDevice(c->device));
	const bool n = (flags & GPV_NORMALS) != 0;
	if (c->gatherL1.ensure(gather_l1n_offset(cells_total) + (n ? (size_t)cells_total * 3 : 0) + 64) || c->gatherPrefix.ensure((size_t)cells_total * 4 + 64) ||
	    c->gatherL2.ensure(gather_l2n_offset(l2_capacity) + (n ? (size_t)l2_capacity * 3 : 0) + 64) ||
	    c->gatherMail.ensure(sizeof(GatherMail)))
		return 1;
	GPV_CUDA(cudaMemset(c->gatherMail.p, 0, sizeof(GatherMail)));
	memset(out, 0, sizeof *out);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "gpv_gather_desc carries 64-byte IPC handles");
	cudaIpcMemHandle_t h;
	GPV_CUDA(cudaIpcGetMemHandle(&h, c->gatherL1.p)); memcpy(out->l1, &h, 64);
	GPV_CUDA(cudaIpcGetMemHandle(&h, c->gatherPrefix.p)); memcpy(out->prefix, &h, 64);
	GPV_CUDA(cudaIpcGetMemHandle(&h, c->gatherL2.p)); memcpy(out->l2, &h, 64);
	GPV_CUDA(cudaIpcGetMemHandle(&h, c->gatherMail.p)); memcpy(out->mailbox, &h, 64);
	out->cells_total = cells_total; out->l2_capacity = l2_capacity; out->owner_device = c->device; out->reserved = flags & GPV_NORMALS;
	const unsigned long long timeoutNs = c->gather.timeoutNs;
	c->gather = {};
	c->gather.timeoutNs = timeoutNs;
	c->gather.owner = true; c->gather.cellsTotal = cells_total; c->gather.l2Cap = l2_capacity; c->gather.flags = flags & GPV_NORMALS;
	return 0;
}

static int gather_bind(gpv_ctx* c, void* l1, void* prefix, void* l2, void* mail, int64_t cellsTotal, int64_t l2Cap, int rank, int world, bool ipc, int flags)
{
	if (rank < 0 || world < 1 || rank >= world || world > 16) return fail("gpv_gather_attach: bad rank / world (at most 16 ranks)");
	c->gather.on = true; c->gather.ipc = ipc; c->gather.rank = rank; c->gather.world = world; c->gather.epoch = 0;
	c->gather.l1 = (uint8_t*)l1; c->gather.prefix = (int32_t*)prefix; c->gather.l2 = (uint8_t*)l2; c->gather.mail = (GatherMail*)mail;
	c->gather.l2p = (uint8_t*)l2 + gather_packed_offset(l2Cap);
	c->gather.flags = flags;
	c->gather.l1n = (flags & GPV_NORMALS) ? (uint8_t*)l1 + gather_l1n_offset(cellsTotal) : nullptr;
	c->gather.l2n = (flags & GPV_NORMALS) ? (uint8_t*)l2 + gather_l2n_offset(l2Cap) : nullptr;
	c->gather.cellsTotal = cellsTotal; c->gather.l2Cap = l2Cap;
	return 0;
}

extern "C" int gpv_gather_attach(gpv_ctx* c, const gpv_gather_desc* d, int rank, int world)
{
	if (!c || !d) return fail("gpv_gather_attach: null argument");
	GPV_CUDA(cudaSetDevice(c->device));
	if (c->gather.owner) { // the gathering rank uses its own allocations
		if (rank != 0) return fail("gpv_gather_attach: the ctx that created the gather buffers is rank 0");
		// a new session starts from epoch 0 on every rank: flags left by an earlier session on the same buffers must not satisfy its polls.
		// (Attaching is collective: no peer is inside a call.)
		GPV_CUDA(cudaMemset(c->gatherMail.p, 0, sizeof(GatherMail)));
		return gather_bind(c, c->gatherL1.p, c->gatherPrefix.p, c->gatherL2.p, c->gatherMail.p, d->cells_total, d->l2_capacity, rank, world, false, d->reserved);
	}
	void* p[4] = {};
	const unsigned char* hs[4] = { d->l1, d->prefix, d->l2, d->mailbox };
	for (int k = 0; k < 4; k++) {
		cudaIpcMemHandle_t h;
		memcpy(&h, hs[k], 64);
		GPV_CUDA(cudaIpcOpenMemHandle(&p[k], h, cudaIpcMemLazyEnablePeerAccess));
	}
	return gather_bind(c, p[0], p[1], p[2], p[3], d->cells_total, d->l2_capacity, rank, world, true, d->reserved);
}

extern "C" int gpv_gather_attach_local(gpv_ctx* c, gpv_ctx* owner, int rank, int world)
{
	if (!c || !owner || !owner->gather.owner) return fail("gpv_gather_attach_local: the owner ctx has no gather buffers");
	GPV_CUDA(cudaSetDevice(c->device));
	if (c->device != owner->device) {
		cudaError_t e = cudaDeviceEnablePeerAccess(owner->device, 0);
		if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
		cudaGetLastError();
	}
	const bool own = c == owner;
	if (own && rank != 0) return fail("gpv_gather_attach_local: the owner is rank 0");
	const int64_t cellsTotal = owner->gather.cellsTotal, l2Cap = owner->gather.l2Cap;
	const int gflags = owner->gather.flags;
	if (own) { cudaSetDevice(owner->device); GPV_CUDA(cudaMemset(owner->gatherMail.p, 0, sizeof(GatherMail))); GPV_CUDA(cudaSetDevice(c->device)); } // new session (see gpv_gather_attach)
	if (!own) { const unsigned long long t = c->gather.timeoutNs; c->gather = {}; c->gather.timeoutNs = t; }
	return gather_bind(c, owner->gatherL1.p, owner->gatherPrefix.p, owner->gatherL2.p, owner->gatherMail.p, cellsTotal, l2Cap, rank, world, false, gflags);
}

extern "C" int gpv_gather_normals(gpv_ctx* c, uint8_t** l1n, uint8_t** l2n)
{
	if (!c || !c->gather.on || c->gather.rank != 0) return fail("gpv_gather_normals: not the gathering rank");
	if (!c->gather.l1n) return fail("gpv_gather_normals: the gather buffers were created without normal streams");
	if (l1n) *l1n = c->gather.l1n;
	if (l2n) *l2n = c->gather.l2n;
	return 0;
}

extern "C" int gpv_gather_result(gpv_ctx* c, uint8_t** l1, int32_t** prefix, uint8_t** l2, int64_t* nb)
{
	if (!c || !c->gather.on || c->gather.rank != 0) return fail("gpv_gather_result: not the gathering rank");
	if (l1) *l1 = c->gather.l1;
	if (prefix) *prefix = c->gather.prefix;
	if (l2) *l2 = c->gather.l2;
	if (nb) *nb = c->gather.nbTotal;
	return 0;
}

// ------------------------------------------------------------------------------------------------ voxel hierarchy / collision structures
extern "C" int gpv_collision_boxes(gpv_ctx* c, void* stream, gpv_collision* out)
{
	if (!c || !out) return fail("gpv_collision_boxes: null argument");
	memset(out, 0, sizeof *out);
	if (!c->last.valid) return fail("gpv_collision_boxes: no completed (non-gathering) voxelization on this ctx");
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	const GridP& g = c->last.g;
	const long long cells = c->last.cells;
	const long long nBlocks = (cells + kOccBlock - 1) / kOccBlock;
	const size_t dsz = desc_bytes(nBlocks);
	if (c->occCount.ensure((size_t)nBlocks * 4 + 32) || c->occOff.ensure((size_t)(nBlocks + 1) * 4 + 32) || c->desc.ensure(dsz + 32)) return 1;
	GPV_CUDA(cudaMemsetAsync(c->desc.p, 0, dsz, st));
	GPV_CUDA(cudaMemsetAsync(c->totals.as<char>() + 208, 0, 8, st));
	k_occupied_count<<<(unsigned)nBlocks, 256, 0, st>>>(c->l1State.as<unsigned char>(), cells, c->occCount.as<int>());
	int64_t launches = 0;
	const ScanReq r[1] = { { c->occCount.as<int>(), nBlocks, c->occOff.as<unsigned>(), reinterpret_cast<unsigned*>(c->totals.as<char>() + 208), nullptr, 0 } };
	launch_scans(c, st, r, 1, launches);
	unsigned n = 0;
	GPV_CUDA(cudaMemcpyAsync(&n, c->totals.as<char>() + 208, 4, cudaMemcpyDeviceToHost, st));
	GPV_CUDA(cudaStreamSynchronize(st));
	if (c->occInv.ensure((size_t)n * 4 + 32) || c->occCenter.ensure((size_t)n * 12 + 32) || c->occExtent.ensure((size_t)n * 12 + 32)) return 1;
	if (n) k_occupied_write<<<(unsigned)nBlocks, 256, 0, st>>>(c->l1State.as<unsigned char>(), cells, c->occOff.as<unsigned>(), g.nx, g.ny, c->tabX.as<float>(), c->tabY.as<float>(),
	                                                          c->tabZ.as<float>() + g.z0, g.h1x, g.h1y, g.h1z, c->occInv.as<int>(), c->occCenter.as<float>(), c->occExtent.as<float>());
	GPV_CUDA(cudaStreamSynchronize(st));
	GPV_CUDA(cudaGetLastError());
	out->count = n; out->d_inv_index = c->occInv.as<int32_t>(); out->d_center = c->occCenter.as<float>(); out->d_extent = c->occExtent.as<float>();
	out->index_base = (int64_t)g.z0 * g.nx * g.ny;
	return 0;
}

extern "C" int gpv_build_hierarchy(gpv_ctx* c, void* stream, gpv_hierarchy* out)
{
	if (!c || !out) return fail("gpv_build_hierarchy: null argument");
	memset(out, 0, sizeof *out);
	if (!c->last.valid || !c->last.whole || !c->last.solid) return fail("gpv_build_hierarchy: needs a completed whole-grid voxelization made with GPV_COLLISION on this ctx");
	const GridP& g = c->last.g;
	auto pow2 = [](int x) { return x > 0 && (x & (x - 1)) == 0; };
	if (!pow2(g.nx) || !pow2(g.ny) || !pow2(g.nz) || (long long)g.nx * g.ny * g.nz < 2)
		return fail("gpv_build_hierarchy: every grid dimension must be a power of two (Object::BuildHierarchy, src/Object.cpp:2790-2867, indexes out of bounds on other grids)");
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	const int total = g.nx * g.ny * g.nz;
	int numLevels = 0;
	{ float y = (float)total; while (y > 1) { y /= 2; numLevels++; } } // GetExponent2, src/Utilities.cpp:349
	if (c->hierMid.ensure((size_t)(total - 1) * 12 + 32) || c->hierHalf.ensure((size_t)(total - 1) * 12 + 32) || c->hierSolid.ensure((size_t)total + 32) || c->hierChild.ensure((size_t)(total - 1) * 8 + 32)) return 1;
	HierOut h{ c->hierMid.as<float>(), c->hierHalf.as<float>(), c->hierSolid.as<unsigned char>(), c->hierChild.as<int>() };
	int numLevelBoxes = total / 2;
	k_hier_leaves<<<(numLevelBoxes + 255) / 256, 256, 0, st>>>(total, g.nx, g.ny, c->tabX.as<float>(), c->tabY.as<float>(), c->tabZ.as<float>(), g.h1x, g.h1y, g.h1z, c->solidWords.as<unsigned>(), h);
	int dX = g.nx / 2, dY = g.ny, dZ = g.nz, prevLevelIndex = 0, levelIndex = numLevelBoxes;
	numLevelBoxes /= 2;
	for (int level = 2; level < numLevels + 1; level++) { // the reference's level schedule, :2822-2866
		int iSkip = (level % 3 == 1 && dX > 1) ? 2 : 1, jSkip = (level % 3 == 2 && dY > 1) ? 2 : 1, kSkip = (level % 3 == 0 && dZ > 1) ? 2 : 1;
		if (iSkip == 1 && jSkip == 1 && kSkip == 1) { if (dX > 1) iSkip = 2; else if (dY > 1) jSkip = 2; else if (dZ > 1) kSkip = 2; }
		k_hier_level<<<(numLevelBoxes + 255) / 256, 256, 0, st>>>(dX, dY, dZ, iSkip, jSkip, kSkip, prevLevelIndex, levelIndex, h);
		if (iSkip == 2) dX /= 2;
		if (jSkip == 2) dY /= 2;
		if (kSkip == 2) dZ /= 2;
		prevLevelIndex += numLevelBoxes * 2;
		levelIndex += numLevelBoxes;
		numLevelBoxes /= 2;
	}
	GPV_CUDA(cudaStreamSynchronize(st));
	GPV_CUDA(cudaGetLastError());
	out->num_levels = numLevels; out->n_boxes = total - 1;
	out->d_mid = h.mid; out->d_half = h.half; out->d_solid = h.solid; out->d_child = h.child;
	return 0;
}

// ------------------------------------------------------------------------------------------------ roofline probes
// non-FMA FP32 issue rate: independent FMUL/FADD chains, 8 accumulators per thread
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters)
{
	float a0 = threadIdx.x * 1e-3f + 1.f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f, a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
	const float m = 1.0000001f, d = 1e-7f;
	for (int i = 0; i < iters; i++) {
		a0 = a0 * m; a1 = a1 + d; a2 = a2 * m; a3 = a3 + d; a4 = a4 * m; a5 = a5 + d; a6 = a6 * m; a7 = a7 + d;
		a0 = a0 + d; a1 = a1 * m; a2 = a2 + d; a3 = a3 * m; a4 = a4 + d; a5 = a5 * m; a6 = a6 + d; a7 = a7 * m;
	}
	if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678f) out[0] = a0;
}

extern "C" int gpv_expand_packed_l2(const void* packed, int64_t n_words, uint8_t* out)
{
	if (!packed || !out || n_words < 0) return fail("gpv_expand_packed_l2: bad argument");
	expand_packed(packed, out, (size_t)n_words);
	return 0;
}

extern "C" int gpv_measure_fp32_peak(gpv_ctx* c, void* stream, double* ops)
{
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	if (c->scratch.ensure(1024)) return 1;
	cudaEvent_t e0, e1;
	GPV_CUDA(cudaEventCreate(&e0)); GPV_CUDA(cudaEventCreate(&e1));
	const int iters = 4096, blocks = c->smCount * 16;
	k_fp32_peak<<<blocks, 256, 0, st>>>(c->scratch.as<float>(), 64);
	double best = 0;
	for (int rep = 0; rep < 5; rep++) {
		GPV_CUDA(cudaEventRecord(e0, st));
		k_fp32_peak<<<blocks, 256, 0, st>>>(c->scratch.as<float>(), iters);
		GPV_CUDA(cudaEventRecord(e1, st));
		GPV_CUDA(cudaEventSynchronize(e1));
		float ms = 0;
		GPV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		double v = (double)blocks * 256 * iters * 16 / (ms * 1e-3);
		if (v > best) best = v;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*ops = best;
	return 0;
}

extern "C" int gpv_measure_copy_peak(gpv_ctx* c, void* stream, double* gbs)
{
	GPV_CUDA(cudaSetDevice(c->device));
	cudaStream_t st = (cudaStream_t)stream;
	const size_t bytes = (size_t)1 << 30;
	if (c->scratch.ensure(2 * bytes)) return 1;
	cudaEvent_t e0, e1;
	GPV_CUDA(cudaEventCreate(&e0)); GPV_CUDA(cudaEventCreate(&e1));
	char* p = c->scratch.as<char>();
	double best = 0;
	for (int rep = 0; rep < 6; rep++) {
		GPV_CUDA(cudaEventRecord(e0, st));
		GPV_CUDA(cudaMemcpyAsync(p + bytes, p, bytes, cudaMemcpyDeviceToDevice, st));
		GPV_CUDA(cudaEventRecord(e1, st));
		GPV_CUDA(cudaEventSynchronize(e1));
		float ms = 0;
		GPV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
		double v = 2.0 * bytes / (ms * 1e-3) / 1e9;
		if (rep && v > best) best = v;
	}
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	*gbs = best;
	return 0;
}

// ------------------------------------------------------------------------------------------------ compatibility tier
// The reference's three operator entry points (cu:507-540), same signatures and launch semantics (asynchronous on the
// default stream, always return 1, errors surface at the caller's CUDACheckErrors), strict-IEEE arithmetic.
namespace gpv {

__global__ void __launch_bounds__(128) k_compat_l1(const float* __restrict__ tris, int nTri, float* inOut, int* count, int* triIndex,
                                                   gpv_float3 mn, gpv_float3 mx, gpv_float3 ext, gpv_int3 nd, int bufLen)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= nTri) return;
	const float* v = tris + (size_t)t * 9;
	float a0 = v[0], a1 = v[1], a2 = v[2], b0 = v[3], b1 = v[4], b2 = v[5], c0 = v[6], c1 = v[7], c2 = v[8];
	int x0 = cell_of(a0, mn.x, mx.x, nd.x), x1 = cell_of(b0, mn.x, mx.x, nd.x), x2 = cell_of(c0, mn.x, mx.x, nd.x);
	int y0 = cell_of(a1, mn.y, mx.y, nd.y), y1 = cell_of(b1, mn.y, mx.y, nd.y), y2 = cell_of(c1, mn.y, mx.y, nd.y);
	int z0 = cell_of(a2, mn.z, mx.z, nd.z), z1 = cell_of(b2, mn.z, mx.z, nd.z), z2 = cell_of(c2, mn.z, mx.z, nd.z);
	int lox = max(0, min(x0, min(x1, x2))), hix = max(x0, max(x1, x2));
	int loy = max(0, min(y0, min(y1, y2))), hiy = max(y0, max(y1, y2));
	int loz = max(0, min(z0, min(z1, z2))), hiz = max(z0, max(z1, z2));
	for (int p = lox; p <= hix && p < nd.x; p++) for (int q = loy; q <= hiy && q < nd.y; q++) for (int r = loz; r <= hiz && r < nd.z; r++) {
		float mx_ = (float)((p + 0.5) * (double)ext.x * 2 + (double)mn.x); // cu:382-384
		float my_ = (float)((q + 0.5) * (double)ext.y * 2 + (double)mn.y);
		float mz_ = (float)((r + 0.5) * (double)ext.z * 2 + (double)mn.z);
		if (tri_box_overlap(mx_, my_, mz_, ext.x, ext.y, ext.z, a0, a1, a2, b0, b1, b2, c0, c1, c2)) {
			int index = r * nd.y * nd.x + q * nd.x + p;
			inOut[index] = 2;
			int slot = atomicAdd(count + index, 1);
			if (slot < bufLen) triIndex[(size_t)index * bufLen + slot] = t; // the reference writes unconditionally (App. B4)
		}
	}
}

__device__ __forceinline__ void compat_l2_centre(int loc, gpv_int3 n2, const float* mid, int b, gpv_float3 e1, gpv_float3 e2, float& cxv, float& cyv, float& czv)
{
	int r = loc / (n2.x * n2.y), pq = loc - r * n2.x * n2.y, q = pq / n2.x, p = pq % n2.x;
	cxv = (float)(2 * p + 1) * e2.x + mid[b * 3 + 0] - e1.x; // cu:423-425
	cyv = (float)(2 * q + 1) * e2.y + mid[b * 3 + 1] - e1.y;
	czv = (float)(2 * r + 1) * e2.z + mid[b * 3 + 2] - e1.z;
}

__global__ void __launch_bounds__(256) k_compat_l2_sat(const float* __restrict__ tris, float* l2InOut, float* l2Normal, const float* __restrict__ mid,
                                                       const int* __restrict__ l2Index, const int* __restrict__ triCount, const int* __restrict__ triFlatIndex,
                                                       const int* __restrict__ triFlat, int nB, gpv_int3 n2, gpv_float3 e1, gpv_float3 e2)
{
	const long long n23 = (long long)n2.x * n2.y * n2.z;
	long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nB * n23) return;
	int b = (int)(v / n23), loc = (int)(v - (long long)b * n23);
	int cell = l2Index[b], nT = triCount[cell], first = triFlatIndex[cell];
	float cxv, cyv, czv;
	compat_l2_centre(loc, n2, mid, b, e1, e2, cxv, cyv, czv);
	// The accumulators start from ZERO whatever the buffer holds: the reference's caller zeroes only a quarter of it
	// (cudaMemset with element counts, src/Object.cpp:2595-2596); zero-initialised is the intended semantics.
	float ax = 0.f, ay = 0.f, az = 0.f, an = 0.f;
	bool hit = false;
	for (int k = 0; k < nT; k++) {
		const float* d = tris + (size_t)triFlat[first + k] * 9;
		if (tri_box_overlap(cxv, cyv, czv, e2.x, e2.y, e2.z, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8])) {
			hit = true;
			float ux = d[3] - d[0], uy = d[4] - d[1], uz = d[5] - d[2], wx = d[6] - d[0], wy = d[7] - d[1], wz = d[8] - d[2];
			ax += uy * wz - uz * wy; ay += uz * wx - ux * wz; az += ux * wy - uy * wx; an += 1; // cu:40-46, 311-318
		}
	}
	if (hit) l2InOut[v] = 2;
	l2Normal[v * 4] = ax; l2Normal[v * 4 + 1] = ay; l2Normal[v * 4 + 2] = az; l2Normal[v * 4 + 3] = an;
}

__global__ void __launch_bounds__(256) k_compat_l2_ray(const float* __restrict__ tris, float* l2InOut, const float* __restrict__ mid, const int* __restrict__ l2Index,
                                                       const int* __restrict__ xyCount, const int* __restrict__ xyFlatIndex, const int* __restrict__ xyFlat, int nB,
                                                       gpv_int3 nd, gpv_int3 n2, gpv_float3 e1, gpv_float3 e2)
{
	const long long n23 = (long long)n2.x * n2.y * n2.z;
	long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= nB * n23) return;
	int b = (int)(v / n23), loc = (int)(v - (long long)b * n23);
	int xy = l2Index[b] % (nd.x * nd.y), nT = xyCount[xy], first = xyFlatIndex[xy];
	float cxv, cyv, czv;
	compat_l2_centre(loc, n2, mid, b, e1, e2, cxv, cyv, czv);
	int n = 0;
	for (int k = 0; k < nT; k++) {
		const float* d = tris + (size_t)xyFlat[first + k] * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8]);
		n += s.ok && ray_column(s, cxv, cyv, rc) && ray_cell(s, rc, czv);
	}
	l2InOut[v] = (n % 2 == 1) ? 1.f : 0.f; // cu:497-501; every voxel is written (0 = the intended zero-initialisation, App. B2)
}

__global__ void __launch_bounds__(256) k_max_reduce(const float* __restrict__ in, long long n, float* out)
{
	__shared__ float s[8];
	float m = -3.402823466e+38f;
	for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) m = fmaxf(m, in[i]);
	for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
	__syncthreads();
	if (threadIdx.x < 8) {
		m = s[threadIdx.x];
		for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
		if (threadIdx.x == 0) out[blockIdx.x] = m;
	}
}

} // namespace gpv

extern "C" int CUDAClassifyTessellation(float* tris, int nTri, float* inOut, int* count, int* triIndex, gpv_float3 mn, gpv_float3 mx, gpv_float3 ext,
                                        gpv_int3 nd, int bufLen)
{
	// the reference's caller zeroes only cells BYTES of the int counter array (src/Object.cpp:2110, App. B1): zero all of it
	cudaMemsetAsync(count, 0, (size_t)nd.x * nd.y * nd.z * sizeof(int));
	if (nTri > 0) k_compat_l1<<<(nTri + 127) / 128, 128>>>(tris, nTri, inOut, count, triIndex, mn, mx, ext, nd, bufLen);
	return 1;
}

extern "C" int CUDAClassifyTessellationLevel2(float* tris, float* l2InOut, float* l2Normal, float* mid, int* l2Index, int* triCount, int* triFlatIndex,
                                              int* triFlat, int nB, gpv_int3 n2, gpv_float3 e1, gpv_float3 e2)
{
	long long total = (long long)nB * n2.x * n2.y * n2.z;
	if (total > 0) k_compat_l2_sat<<<(unsigned)((total + 255) / 256), 256>>>(tris, l2InOut, l2Normal, mid, l2Index, triCount, triFlatIndex, triFlat, nB, n2, e1, e2);
	return 1;
}

extern "C" int CUDAClassifyInOutLevel2(float* tris, float* l2InOut, float* mid, int* l2Index, int* xyCount, int* xyFlatIndex, int* xyFlat, int nB,
                                       gpv_int3 nd, gpv_int3 n2, gpv_float3 e1, gpv_float3 e2)
{
	long long total = (long long)nB * n2.x * n2.y * n2.z;
	if (total > 0) k_compat_l2_ray<<<(unsigned)((total + 255) / 256), 256>>>(tris, l2InOut, mid, l2Index, xyCount, xyFlatIndex, xyFlat, nB, nd, n2, e1, e2);
	return 1;
}

namespace gpv { __device__ float g_maxPart[256 + 1]; } // per-device scratch of THRUSTDeviceFindMax: no allocation per call

extern "C" float THRUSTDeviceFindMax(float* data, int w, int h)
{
	// cuda/THRUSTUtilities.cu:44-61: thrust::max_element over w*h floats on the default stream, then a one-element D2H.
	long long n = (long long)w * h;
	if (n <= 0 || !data) return 0.f;
	float* part = nullptr;
	if (cudaGetSymbolAddress((void**)&part, gpv::g_maxPart) != cudaSuccess) return 0.f;
	const int blocks = 256;
	k_max_reduce<<<blocks, 256>>>(data, n, part);
	k_max_reduce<<<1, 256>>>(part, blocks, part + blocks);
	float r = 0.f;
	if (cudaMemcpy(&r, part + blocks, sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return 0.f;
	return r;
}
