// gpview_b200/csrc/gpv_kernels.cuh -- hand-written sm_100a kernels of the two-level voxelizer hot path.
//
// Compiled ONLY with  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (see DESIGN.md "Numerics").
// Kernel inventory (SURVEY.md 2.3 "new kernels"):
//   k_clear         every counter of the call in one launch
//   k_prepare       36 B flat triangles -> 48 B float4x3 records (TMA-able, footprint in the w lanes) + 48 B +Z ray records;
//                   also the per-axis Level-1 cell-centre tables (kills all FP64 / int->float work in the inner loops)
//   k_bin<FILL>     K1  triangle -> Level-1 cell SAT binning (count / fill sweeps), TMA-staged triangle tiles
//   k_cross<FILL>   K2a certified (column, triangle) crossing detection for the parity fill
//   k_fill_sweep    K2b +Z parity sweep per Level-1 column, coalesced along x, final Level-1 state bytes
//   k_scan<MODE,V>  K3  single-pass decoupled-look-back scan: boundary prefix sum / index compaction / per-column cell counts;
//   k_scan_offs3        up to three CSR offset scans per launch
//   k_sort_segments, k_sort_long   canonical (ascending) order of every cell / column list; de-duplicates column lists
//   k_ray_units, k_col_cells, k_l2_rays, k_l2_rays_overflow   K4a Level-2 parity rays per sub-voxel column of each Level-1 column
//                   (certified cell masks), over a work list of (column, <= 16 boundary cells) units
//   k_l2<N2,OUT>    K4  Level-2 SAT hoisted along z per sub-voxel column, three stages over shared-memory queues; blocks leave as file
//                   bytes, staged bytes per cell, or 2 bits per sub-voxel (NVLink / PCIe)
//   k_l1_normals, k_l2_normals   K5 normals in the reference's uchar encoding
//   k_gather_*, k_scatter_blocks   multi-GPU: every rank's share written into the gathering rank's buffers over NVLink peer memory
//                   (mailbox flags; the peers' 2-bit Level-2 blocks expanded on rank 0)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gpv_math.h"

namespace gpv {

struct GridP {
	int nx, ny, nz;       // Level-1 resolution
	int z0, z1;           // z-slab owned by this launch [z0,z1)
	float minx, miny, minz, maxx, maxy, maxz; // padded bbox
	float gsx, gsy, gsz;  // Level-1 cell size
	float h1x, h1y, h1z;  // Level-1 half extents
	float h2x, h2y, h2z;  // Level-2 half extents
	int n2;               // Level-2 resolution per boundary cell
};

// device-side totals block (one per context), read back once per model
struct Totals {
	unsigned long long binWork, crossWork; // totals of the two balanced work spaces (exclusive scans of binCnt / crossCnt)
	unsigned long long l1Hits, crossPairs, nIll, l1Inside, l2Inside, l2Boundary;
	unsigned long long l2ColPairs; // sum over boundary cells of their column-list length (reference-equivalent Level-2 ray tests / n2^3)
	unsigned long long l2LocalTris; // sum of the cell-list lengths of the boundary cells THIS rank refines (all of them without GPV_GATHER)
	unsigned long long gatherError; // 2 timed out waiting for a peer
	unsigned int nBoundary, triTotal, colTotalOver, crossTotal;
	unsigned int nLocalCells;       // boundary cells this rank refines (GPV_GATHER: those in the Level-1 columns it owns)
	unsigned int nRayHeavy, nRayLight; // units of k_l2_rays (k_ray_units): heavy ones are listed from the front, light ones from the back
};

// GPV_GATHER: which rank refines a Level-1 column.  Columns are dealt out in groups of `group` consecutive columns (the columns one
// CTA of k_l2_rays used to walk), group k of grid row j to rank (k + j) % world: neighbouring columns cost about the same, so the
// interleaving balances the Level-2 work without a cost model, every column list is walked by exactly one rank, and the skew by the
// row keeps a wall of the model that runs along x OR along y from landing on one rank (nx is a multiple of 4: without the skew rank r
// would own the same x positions in every row).  world <= 1: everything is owned.
struct Own {
	int world, rank, group, nx;
	__host__ __device__ __forceinline__ int owner(unsigned col) const { return (int)((col / (unsigned)group + col / (unsigned)nx) % (unsigned)world); }
	__host__ __device__ __forceinline__ bool operator()(unsigned col) const { return world <= 1 || owner(col) == rank; }
};

constexpr int kWorkThreads = 128;                        // threads per CTA of the balanced triangle-work kernels
constexpr int kTileTris = 128;                           // triangle records staged per TMA tile (6 KB + 2 KB)
constexpr int kWorkGrid = 148 * 8;                       // persistent grid: 8 CTAs per SM, work blocks strided over it

// ------------------------------------------------------------------------------------------------ TMA staging
// 1-D bulk async copies (TMA, SASS UBLKCP) from global to shared memory, completion on an mbarrier (SYNCS).
__device__ __forceinline__ void mbar_init(uint64_t* bar)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"((uint32_t)__cvta_generic_to_shared(smemDst)), "l"(gmemSrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	uint32_t done = 0, addr = (uint32_t)__cvta_generic_to_shared(bar);
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(addr), "r"(parity) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------ k_prepare
// flat float[9] per triangle (src/Object.cpp:3496-3527 layout) ->
//   tri48  v0|v1|v2 as float4; the three w lanes carry the clipped Level-1 footprint (cu:333-378) as packed 16-bit fields
//          w0 = lox | loy<<16, w1 = loz | dx<<16, w2 = dy | dz<<16
//   ray48  v1xyz e1xyz e2xyz det inv ok  (gpv::RayTri)
//   plane16 plane record of the certified Level-2 plane culling, normalised by Nz (gpv::PlaneRec on permuted axes)
//   aabbxy16 xmin xmax ymin ymax of the triangle: the x / y AABB predicates of the SAT per sub-voxel column, exactly
//           (fl(t - c) is monotone in t, so min_i fl(t_i - c) = fl(min_i t_i - c))
//   crossFp i0 | j0<<16, di | dj<<16, kind, -   certified candidate columns of the +Z parity fill (gpv::fill_candidates)
//   binCnt / crossCnt  number of (triangle, cell) / (triangle, column) work items
// All records are 16-byte aligned so that contiguous triangle ranges can be moved by TMA bulk copies.
// Centre tables: centre[p] = fl32((p + 0.5) * ext * 2 + min) evaluated in double exactly like cu:382-384 (== ray origin
// src/Object.cpp:743-745 == mid point :2567-2569, see oracle/gpv_oracle.c gpvo_axis_table).
__global__ void __launch_bounds__(256) k_prepare(const float* __restrict__ flat, long long nTri, GridP g, float4* __restrict__ tri48,
                                                 float4* __restrict__ ray48, float4* __restrict__ plane16, float4* __restrict__ aabbxy16, int4* __restrict__ crossFp,
                                                 int* __restrict__ binCnt, int* __restrict__ crossCnt, Totals* totals, float* cx, float* cy, float* cz)
{
	__shared__ float s[256 * 9];
	long long base = (long long)blockIdx.x * 256;
	{ // the per-axis centre tables of k_tables, fused (the grid covers max(nx, ny, nz) threads as well)
		const long long i = base + threadIdx.x;
		if (i < g.nx) cx[i] = (float)((i + 0.5) * (double)g.h1x * 2 + (double)g.minx);
		if (i < g.ny) cy[i] = (float)((i + 0.5) * (double)g.h1y * 2 + (double)g.miny);
		if (i < g.nz) cz[i] = (float)((i + 0.5) * (double)g.h1z * 2 + (double)g.minz);
	}
	int n = (int)max((long long)0, min((long long)256, nTri - base));
	for (int i = threadIdx.x; i < n * 9; i += 256) s[i] = flat[base * 9 + i]; // coalesced
	__syncthreads();
	int t = threadIdx.x;
	unsigned long long ill = 0;
	if (t < n) {
		const float* v = s + t * 9;
		// vertex -> cell, footprint min/max (cu:333-371); loops are clipped to < numDiv (cu:374-378) and to >= 0
		int x0 = cell_of(v[0], g.minx, g.maxx, g.nx), x1 = cell_of(v[3], g.minx, g.maxx, g.nx), x2 = cell_of(v[6], g.minx, g.maxx, g.nx);
		int y0 = cell_of(v[1], g.miny, g.maxy, g.ny), y1 = cell_of(v[4], g.miny, g.maxy, g.ny), y2 = cell_of(v[7], g.miny, g.maxy, g.ny);
		int w0 = cell_of(v[2], g.minz, g.maxz, g.nz), w1 = cell_of(v[5], g.minz, g.maxz, g.nz), w2 = cell_of(v[8], g.minz, g.maxz, g.nz);
		int lox = max(0, min(x0, min(x1, x2))), hix = min(g.nx - 1, max(x0, max(x1, x2)));
		int loy = max(0, min(y0, min(y1, y2))), hiy = min(g.ny - 1, max(y0, max(y1, y2)));
		int loz = max(0, min(w0, min(w1, w2))), hiz = min(g.nz - 1, max(w0, max(w1, w2)));
		int dx = max(0, hix - lox + 1), dy = max(0, hiy - loy + 1), dz = max(0, hiz - loz + 1);
		if (dx == 0 || dy == 0 || dz == 0) { dx = dy = dz = 0; lox = loy = loz = 0; }
		tri48[(base + t) * 3 + 0] = make_float4(v[0], v[1], v[2], __int_as_float(lox | (loy << 16)));
		tri48[(base + t) * 3 + 1] = make_float4(v[3], v[4], v[5], __int_as_float(loz | (dx << 16)));
		tri48[(base + t) * 3 + 2] = make_float4(v[6], v[7], v[8], __int_as_float(dy | (dz << 16)));
		binCnt[base + t] = dx * dy * dz;
		RayTri r;
		ray_tri_setup(r, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
		ray48[(base + t) * 3 + 0] = make_float4(r.v1x, r.v1y, r.v1z, r.e1x);
		ray48[(base + t) * 3 + 1] = make_float4(r.e1y, r.e1z, r.e2x, r.e2y);
		ray48[(base + t) * 3 + 2] = make_float4(r.e2z, r.det, r.inv, r.ok ? (r.well ? 2.f : 1.f) : 0.f);
		// normalised by Nz, for intervals along a sub-voxel column: the same function on cyclically permuted coordinates (x,y,z) <- (z,x,y)
		PlaneRec pl = plane_rec_setup(v[2], v[0], v[1], v[5], v[3], v[4], v[8], v[6], v[7], g.gsz, g.gsx, g.gsy, g.h2z, g.h2x, g.h2y);
		plane16[base + t] = make_float4(pl.sx, pl.ny, pl.nz, pl.R);
		aabbxy16[base + t] = make_float4(fminf(v[0], fminf(v[3], v[6])), fmaxf(v[0], fmaxf(v[3], v[6])), fminf(v[1], fminf(v[4], v[7])), fmaxf(v[1], fmaxf(v[4], v[7])));
		int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
		int kind = fill_candidates(r, g.minx, g.miny, g.gsx, g.gsy, g.nx, g.ny, i0, i1, j0, j1);
		if (kind == 2) { i0 = 0; j0 = 0; i1 = g.nx - 1; j1 = g.ny - 1; ill = 1; }
		int di = kind ? i1 - i0 + 1 : 0, dj = kind ? j1 - j0 + 1 : 0;
		crossFp[base + t] = make_int4(i0 | (j0 << 16), di | (dj << 16), kind, 0);
		crossCnt[base + t] = di * dj;
	}
	ill = __reduce_add_sync(0xffffffffu, (unsigned)ill);
	if ((threadIdx.x & 31) == 0 && ill) atomicAdd(&totals->nIll, ill);
}

__device__ __forceinline__ void load_ray(RayTri& r, const float4* __restrict__ ray48, int t)
{
	float4 a = __ldg(ray48 + (size_t)t * 3), b = __ldg(ray48 + (size_t)t * 3 + 1), c = __ldg(ray48 + (size_t)t * 3 + 2);
	r.v1x = a.x; r.v1y = a.y; r.v1z = a.z; r.e1x = a.w; r.e1y = b.x; r.e1z = b.y; r.e2x = b.z; r.e2y = b.w;
	r.e2z = c.x; r.det = c.y; r.inv = c.z; r.ok = c.w != 0.f; r.well = c.w == 2.f;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// ------------------------------------------------------------------------------------------------ balanced triangle work
// The (triangle, cell) items of all triangles form one flat index space (off = exclusive scan of the per-triangle item
// counts), cut into equal contiguous ranges, one per CTA of a persistent grid (one binary search per CTA).  The triangles
// a range touches are a CONTIGUOUS range too, staged tile by tile (<= 128 records) into shared memory by TMA; each
// thread finds the owner of its item by binary search in the tile's offsets (shared memory) and decodes the cell from the
// item's rank inside the footprint.  This removes the footprint skew (cessna-256: mean 64, max 6,762 cells per triangle;
// an ill-conditioned triangle of the parity fill has nx*ny candidate columns) that a triangle-per-thread map suffers from.
struct WorkSmem {
	float4 rec[kTileTris * 3];
	int4 fp[kTileTris];
	unsigned off[kTileTris + 1];
	int t0;
	uint64_t bar;
};

template <bool HAS_FP, class F>
__device__ __forceinline__ void for_each_work_item(WorkSmem& sm, const unsigned* __restrict__ off, int nTri, const unsigned long long* totalPtr,
                                                   const float4* __restrict__ rec48, const int4* __restrict__ fp, F&& f)
{
	const int tid = threadIdx.x;
	if (tid == 0) mbar_init(&sm.bar);
	__syncthreads();
	const unsigned long long total = *totalPtr;
	uint32_t parity = 0;
	// this CTA's contiguous item range [R0,R1): the items are dealt out in units of one item per thread (128), so that a small model
	// (cessna-256: 478 k items) still spreads over the whole persistent grid instead of 1,024 items on each of a third of the CTAs
	const unsigned long long nBlocks = (total + kWorkThreads - 1) / kWorkThreads, per = (nBlocks + gridDim.x - 1) / gridDim.x;
	const unsigned long long R0 = min(total, blockIdx.x * per * kWorkThreads), R1 = min(total, (blockIdx.x + 1ull) * per * kWorkThreads);
	if (R0 >= R1) return;
	if (tid == 0) { // largest t with off[t] <= R0   (off[0] = 0 <= R0 < total = off[nTri]); ONE search per CTA
		int lo = 0, hi = nTri;
		while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= R0) lo = mid; else hi = mid; }
		sm.t0 = lo;
	}
	__syncthreads();
	int t = sm.t0;
	for (;;) {
		const int nt = min(kTileTris, nTri - t);
		__syncthreads(); // the previous round's readers of sm.off / sm.t0 are done
		for (int i = tid; i <= nt; i += kWorkThreads) sm.off[i] = off[t + i];
		if (tid == 0) sm.t0 = 0x7fffffff;
		__syncthreads();
		const unsigned o0 = sm.off[0];
		if (sm.off[nt] == o0) {
			// no item in these 128 triangles (e.g. a run of triangles that can never be hit): fast-forward.  Every thread
			// probes the end offset of one of the next 128 tiles; the first tile that ends beyond o0 holds the next item.
			if (o0 >= R1 || t + nt >= nTri) break;
			const long long probe = (long long)t + (long long)(tid + 1) * kTileTris;
			if (off[min((long long)nTri, probe)] > o0) atomicMin(&sm.t0, tid);
			__syncthreads();
			const int first = sm.t0;
			if (first == 0x7fffffff) { if ((long long)t + (long long)kWorkThreads * kTileTris >= nTri) break; t += kWorkThreads * kTileTris; }
			else t += first * kTileTris;
			continue;
		}
		if (tid == 0) {
			mbar_expect(&sm.bar, (uint32_t)nt * (HAS_FP ? 64u : 48u));
			tma_bulk_g2s(sm.rec, rec48 + (size_t)t * 3, (uint32_t)nt * 48u, &sm.bar);
			if (HAS_FP) tma_bulk_g2s(sm.fp, fp + t, (uint32_t)nt * 16u, &sm.bar);
		}
		mbar_wait(&sm.bar, parity);
		parity ^= 1;
		__syncthreads();
		const unsigned long long cb = max(R0, (unsigned long long)sm.off[0]), ce = min(R1, (unsigned long long)sm.off[nt]);
		const bool last = sm.off[nt] >= R1 || t + nt >= nTri;
		for (unsigned long long c = cb + tid; c < ce; c += kWorkThreads) {
			int lo = 0, hi = nt; // largest k in [0,nt) with off[k] <= c
			while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (sm.off[mid] <= c) lo = mid; else hi = mid; }
			f(t + lo, sm.rec + lo * 3, sm.fp + lo, (unsigned)(c - sm.off[lo]), sm.off[lo]);
		}
		__syncthreads();
		if (last) break;
		t += nt;
	}
}

// ------------------------------------------------------------------------------------------------ k_bin
// K1.  Replaces CUDAClassifyTessellationKernel (cu:320-401) + the host CSR flatten and per-column de-duplication
// (src/Object.cpp:2137-2180).  One thread = one (triangle, Level-1 cell) SAT test of the clipped footprint (cu:374-395).
// A triangle hits several cells of a column but enters the COLUMN's list once: every (triangle, column of its footprint) pair
// owns one bit of a bitmap laid out over the same flat work space (bit = the triangle's work offset + the column's index in
// its footprint), and the thread whose atomicOr finds the bit clear is the one that adds.  The reference de-duplicates on the
// host with a bool[numTriangles] per column; here the lists are duplicate-free by construction and need no sort to be USED
// (only to be compared -- GPV_KEEP_LISTS -- and for the order-dependent f32 sums of the normals).  Two sweeps over the same
// balanced work space, each with its own bitmap:
//   FILL=false  cellCount[cell]++ (slab cells only), colCount[col]++
//   FILL=true   cellTris[bTriOff[prefix[cell]] + slot] (slot by atomic decrement), colTris[colOff[col] + slot] (slot by a cursor)
struct BinOut {
	int* cellCount;            // slab-local linear index
	int* colCount;             // nx*ny: length of every column list (kept)
	int* colCursor;            // FILL: nx*ny, zeroed
	unsigned* bits;            // this sweep's (triangle, column) bitmap, zeroed over the work space (k_clear_bits)
	unsigned long long bitsCap; // bits available; a larger work space makes the sweep leave (the host grows the pool and starts over)
	const int* prefix;         // FILL: slab-local boundary rank of each cell
	const unsigned* bTriOff;   // FILL
	int* cellTris;             // FILL
	const unsigned* colOff;    // FILL
	int* colTris;              // FILL
	Totals* totals;
};

// zeroes both sweeps' bitmaps over the work space of this model (its size is on the device only)
__global__ void __launch_bounds__(256) k_clear_bits(unsigned* bits, unsigned long long capBits, const Totals* totals)
{
	const unsigned long long n = totals->binWork;
	if (n > capBits) return;
	const unsigned long long words = (n + 31) / 32, capWords = capBits / 32;
	for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < words; i += (unsigned long long)gridDim.x * 256) { bits[i] = 0u; bits[capWords + i] = 0u; }
}

template <bool FILL>
__global__ void __launch_bounds__(kWorkThreads) k_bin(const float4* __restrict__ tri48, int nTri, const unsigned* __restrict__ binOff, GridP g,
                                                      const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz, BinOut o)
{
	__shared__ __align__(128) WorkSmem sm;
	if (o.totals->binWork > o.bitsCap || o.totals->binWork > 0xfffffff0ull) return; // bitmap too small (the host grows it and starts over) / 32-bit work offsets would wrap (the host reports it)
	unsigned long long hits = 0;
	for_each_work_item<false>(sm, binOff, nTri, &o.totals->binWork, tri48, nullptr, [&](int t, const float4* rec, const int4*, unsigned local, unsigned base) {
		const float4 a = rec[0], b = rec[1], c = rec[2];
		const unsigned w0 = __float_as_uint(a.w), w1 = __float_as_uint(b.w), w2 = __float_as_uint(c.w);
		const unsigned dx = w1 >> 16, dy = w2 & 0xffffu;
		const unsigned rest = local / dx, rz = rest / dy;
		const int p = (int)((w0 & 0xffffu) + (local - rest * dx)), q = (int)((w0 >> 16) + (rest - rz * dy)), r = (int)((w1 & 0xffffu) + rz);
		if (!tri_box_overlap(cx[p], cy[q], cz[r], g.h1x, g.h1y, g.h1z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z)) return;
		hits++;
		const int col = q * g.nx + p;
		if (r >= g.z0 && r < g.z1) {
			const size_t li = (size_t)(r - g.z0) * g.ny * g.nx + col;
			if (!FILL) atomicAdd(o.cellCount + li, 1);
			else o.cellTris[o.bTriOff[o.prefix[li]] + atomicSub(o.cellCount + li, 1) - 1] = t;
		}
		const unsigned long long bit = (unsigned long long)base + (local - rz * dx * dy); // < base + dx*dy <= the triangle's end of the work space
		const unsigned mask = 1u << (bit & 31);
		if (atomicOr(o.bits + (bit >> 5), mask) & mask) return; // another cell of this column has entered the triangle already
		if (!FILL) atomicAdd(o.colCount + col, 1);
		else o.colTris[o.colOff[col] + atomicAdd(o.colCursor + col, 1)] = t;
	});
	if (!FILL) {
		hits = warp_sum(hits);
		if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&o.totals->l1Hits, hits);
	}
}

// ------------------------------------------------------------------------------------------------ k_cross
// K2a.  For every triangle, the Level-1 columns whose +Z ray passes the det/u/v part of Moller-Trumbore (constant along
// the column, App. A.6).  Candidate columns come from gpv::fill_candidates (certified superset of what the reference's
// brute force Object::ClassifyInOutCPU, src/Object.cpp:716-779, can hit); ill-conditioned triangles test every column.
// One thread = one (triangle, candidate column) test over the balanced work space.
template <bool FILL>
__global__ void __launch_bounds__(kWorkThreads) k_cross(const float4* __restrict__ ray48, const int4* __restrict__ crossFp, int nTri,
                                                        const unsigned* __restrict__ workOff, GridP g, const float* __restrict__ cx,
                                                        const float* __restrict__ cy, int* crossCount, const unsigned* __restrict__ crossOff,
                                                        int* crossTri, Totals* totals)
{
	__shared__ __align__(128) WorkSmem sm;
	if (totals->crossWork > 0xfffffff0ull) return; // work offsets are 32-bit: the host reports the error after the read-back (gpv_abi.cu), nothing may run on wrapped offsets
	unsigned long long found = 0;
	for_each_work_item<true>(sm, workOff, nTri, &totals->crossWork, ray48, crossFp, [&](int t, const float4* rec, const int4* fp, unsigned local, unsigned) {
		const float4 a = rec[0], b = rec[1], c = rec[2];
		RayTri s;
		s.v1x = a.x; s.v1y = a.y; s.v1z = a.z; s.e1x = a.w; s.e1y = b.x; s.e1z = b.y; s.e2x = b.z; s.e2y = b.w;
		s.e2z = c.x; s.det = c.y; s.inv = c.z; s.ok = true; s.well = false;
		const unsigned f0 = (unsigned)fp->x, f1 = (unsigned)fp->y;
		const unsigned di = f1 & 0xffffu, jj = local / di;
		const int i = (int)((f0 & 0xffffu) + (local - jj * di)), j = (int)((f0 >> 16) + jj);
		RayCol rc;
		if (!ray_column(s, cx[i], cy[j], rc)) return;
		found++;
		const int col = j * g.nx + i;
		if (!FILL) atomicAdd(crossCount + col, 1);
		else crossTri[crossOff[col] + atomicSub(crossCount + col, 1) - 1] = t;
	});
	if (!FILL) {
		found = warp_sum(found);
		if ((threadIdx.x & 31) == 0 && found) atomicAdd(&totals->crossPairs, found);
	}
}

// ------------------------------------------------------------------------------------------------ k_fill_sweep
// K2b.  One thread = one Level-1 column x one chunk of 32 z-layers; a warp = 32 x-adjacent columns, so every store
// instruction writes 32 consecutive state bytes.  For each crossing triangle of the column only `t` is evaluated per
// cell.  Writes the FINAL Level-1 file byte: 254 boundary (bit set in bmask), 127 inside, 0 outside
// (fill first, SAT overwrites: src/Object.cpp:3158 then :3202; uchar(state*127) :3031).
__global__ void __launch_bounds__(128) k_fill_sweep(const float4* __restrict__ ray48, GridP g, const float* __restrict__ cx, const float* __restrict__ cy,
                                                    const float* __restrict__ cz, const unsigned* __restrict__ crossOff, const int* __restrict__ crossTri,
                                                    const unsigned char* __restrict__ bmask, unsigned char* __restrict__ l1State, Totals* totals,
                                                    unsigned* __restrict__ solidWords)
{
	const int i = blockIdx.x * 32 + threadIdx.x;
	const int j = blockIdx.y;
	const int kbase = g.z0 + (blockIdx.z * 4 + threadIdx.y) * 32;
	unsigned inside = 0;
	if (i < g.nx && kbase < g.z1) {
		const int col = j * g.nx + i;
		const int kn = min(32, g.z1 - kbase);
		unsigned par = 0;
		const float ox = cx[i], oy = cy[j];
		for (unsigned q = crossOff[col]; q < crossOff[col + 1]; q++) {
			RayTri s;
			load_ray(s, ray48, crossTri[q]);
			RayCol rc;
			if (!ray_column(s, ox, oy, rc)) continue; // cannot happen (listed because it passed); keeps c0..c2 defined
			unsigned m = 0;
			const int run = ray_z_run(s, rc, s.well, cz[kbase], cz[kbase + kn - 1]); // whole chunk below / above the crossing?
			if (run == 0) continue;
			if (run == 1) m = kn == 32 ? 0xffffffffu : ((1u << kn) - 1);
			else for (int kk = 0; kk < kn; kk++) m |= (unsigned)ray_cell(s, rc, cz[kbase + kk]) << kk;
			par ^= m;
		}
		const size_t plane = (size_t)g.ny * g.nx;
		// GPV_COLLISION: the parity of every cell INCLUDING the boundary ones (BBoxData::solid is the fill before the SAT pass overwrites it,
		// src/Object.cpp:3165-3193): one word per (column, 32 z-layers), bit kk = layer kbase + kk
		if (solidWords) solidWords[(size_t)((kbase - g.z0) >> 5) * plane + col] = par;
		for (int kk = 0; kk < kn; kk++) {
			size_t li = (size_t)(kbase - g.z0 + kk) * plane + col;
			bool bd = (bmask[li >> 3] >> (li & 7)) & 1;
			unsigned char st = bd ? 254 : (((par >> kk) & 1) ? 127 : 0);
			l1State[li] = st;
			inside += st == 127;
		}
	}
	unsigned long long s = warp_sum((unsigned long long)inside);
	if (threadIdx.x == 0 && s) atomicAdd(&totals->l1Inside, s);
}

// ------------------------------------------------------------------------------------------------ k_scan
// K3.  Single-pass chained scan with decoupled look-back (one 64-bit descriptor per 2048-item tile: 2 status bits + 62
// value bits; tile ids handed out by an atomic counter so every predecessor of a running tile is itself running or done).
//   MODE_CELLS  in = cellCount[n]  ->  prefix[n] (exclusive count of boundary cells, src/Object.cpp:3270-3279),
//               boundaryIndex[b] (global linear index), bTriOff[b] (exclusive sum of list lengths), bmask (1 bit/cell),
//               totals.nBoundary / triTotal.  The two running sums travel packed as (boundary << 31 | tris).
//   MODE_OFFS   in = count[n]      ->  off[n+1] exclusive offsets, total in *totalOut
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;
constexpr unsigned long long kDescAgg = 1ull << 62, kDescIncl = 2ull << 62, kDescValue = (1ull << 62) - 1;
enum { MODE_CELLS = 0, MODE_OFFS = 1 };

struct ScanIO {
	const int* in; long long n;
	unsigned long long* desc; unsigned* tileCounter;
	// MODE_CELLS
	int* prefix; int* boundaryIndex; unsigned* bTriOff; unsigned char* bmask; long long globalBase; Totals* totals;
	int* colCells; long long plane; // MODE_CELLS also counts the boundary cells of every Level-1 column (plane = nx*ny) this rank owns
	Own own;
	// MODE_OFFS
	unsigned* off; unsigned* totalOut; unsigned long long* totalOut64; // either total pointer may be null
};

__device__ __forceinline__ unsigned long long ld_desc(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_desc(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// V = sub-tiles of kScanTile items per tile: one tile id, one look-back and one descriptor per V * 2048 items, so the fixed
// latency of a tile (atomic, look-back, barriers) is paid once per 8 KB * V of input -- V = 4 for grids beyond 1 M cells.
template <int MODE, int V>
__device__ __forceinline__ void scan_body(const ScanIO& io)
{
	static_assert(V * (kScanThreads / 32) <= 32, "the warp totals of a tile are scanned by one warp");
	__shared__ unsigned sTile;
	__shared__ unsigned long long sWarp[V * (kScanThreads / 32)];
	__shared__ unsigned long long sExcl;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int kWarps = kScanThreads / 32;
	if (tid == 0) sTile = atomicAdd(io.tileCounter, 1u);
	__syncthreads();
	const unsigned tile = sTile;
	const long long tileFirst = (long long)tile * (V * kScanTile) + (long long)tid * kScanItems;

	// the tile's items are read twice: once for the sums (before the look-back) and once, from L2, for the outputs -- holding
	// 8 * V values per thread across the look-back would cost the registers that keep enough tiles in flight per SM
	auto load8 = [&](long long first, int* x) {
		if (first + kScanItems <= io.n) {
			const int4* p = reinterpret_cast<const int4*>(io.in + first);
			const int4 a = __ldcg(p), b = __ldcg(p + 1);
			x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
		} else {
#pragma unroll
			for (int k = 0; k < kScanItems; k++) x[k] = (first + k < io.n) ? io.in[first + k] : 0;
		}
	};
	auto item_of = [](int x) { return (MODE == MODE_CELLS) ? (((unsigned long long)(x > 0) << 31) | (unsigned)x) : (unsigned long long)(unsigned)x; };
	unsigned long long threadSum[V], incl[V];
#pragma unroll
	for (int s = 0; s < V; s++) {
		int v[kScanItems];
		load8(tileFirst + (long long)s * kScanTile, v);
		unsigned long long t = 0;
#pragma unroll
		for (int k = 0; k < kScanItems; k++) t += item_of(v[k]);
		threadSum[s] = t;
		unsigned long long in = t;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned long long y = __shfl_up_sync(0xffffffffu, in, o);
			if (lane >= o) in += y;
		}
		incl[s] = in;
		if (lane == 31) sWarp[s * kWarps + warp] = in;
	}
	__syncthreads();
	if (warp == 0) {
		// exclusive prefix over the V * 8 warp totals in item order (sub-tile major), then the tile's look-back
		unsigned long long w = lane < V * kWarps ? sWarp[lane] : 0, wi = w;
#pragma unroll
		for (int o = 1; o < V * kWarps; o <<= 1) {
			unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
			if (lane >= o) wi += y;
		}
		const unsigned long long blockAgg = __shfl_sync(0xffffffffu, wi, V * kWarps - 1);
		if (lane < V * kWarps) sWarp[lane] = wi - w;
		unsigned long long excl = 0;
		if (tile == 0) {
			if (lane == 0) st_desc(io.desc, kDescIncl | blockAgg);
		} else {
			if (lane == 0) st_desc(io.desc + tile, kDescAgg | blockAgg);
			long long look = (long long)tile - 1;
			for (;;) {
				long long idx = look - lane;
				unsigned long long d = kDescIncl; // tiles before the first count as "inclusive 0"
				if (idx >= 0) { do { d = ld_desc(io.desc + idx); } while ((d >> 62) == 0); }
				unsigned inclMask = __ballot_sync(0xffffffffu, (d >> 62) == 2);
				int firstIncl = inclMask ? __ffs(inclMask) - 1 : 32;
				unsigned long long part = (lane <= firstIncl) ? (d & kDescValue) : 0;
				for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
				excl += part;
				if (inclMask) break;
				look -= 32;
			}
			if (lane == 0) st_desc(io.desc + tile, kDescIncl | (excl + blockAgg));
		}
		if (lane == 0) sExcl = excl;
	}
	__syncthreads();

#pragma unroll
	for (int s = 0; s < V; s++) {
		const long long first = tileFirst + (long long)s * kScanTile;
		unsigned long long run = sExcl + sWarp[s * kWarps + warp] + (incl[s] - threadSum[s]);
		int v[kScanItems];
		load8(first, v);
		if (MODE == MODE_CELLS) {
			int pre[kScanItems];
			unsigned flags = 0;
			const unsigned col0 = (unsigned)(io.globalBase + first) % (unsigned)io.plane; // once per thread, not per boundary cell (linear indices are < 2^31)
#pragma unroll
			for (int k = 0; k < kScanItems; k++) {
				unsigned b = (unsigned)(run >> 31), ts = (unsigned)(run & 0x7fffffffu);
				pre[k] = (int)b;
				if (v[k] > 0) {
					flags |= 1u << k;
					io.boundaryIndex[b] = (int)(io.globalBase + first + k);
					unsigned col = col0 + (unsigned)k; // < 2 * plane + 8: the column of cell first + k
					while (col >= (unsigned)io.plane) col -= (unsigned)io.plane;
					if (io.own(col)) atomicAdd(io.colCells + col, 1);
					io.bTriOff[b] = ts;
				}
				run += item_of(v[k]);
			}
			if (first + kScanItems <= io.n) {
				int4* p = reinterpret_cast<int4*>(io.prefix + first);
				p[0] = make_int4(pre[0], pre[1], pre[2], pre[3]);
				p[1] = make_int4(pre[4], pre[5], pre[6], pre[7]);
				io.bmask[first >> 3] = (unsigned char)flags;
			} else {
#pragma unroll
				for (int k = 0; k < kScanItems; k++) if (first + k < io.n) io.prefix[first + k] = pre[k];
				if (first < io.n) io.bmask[first >> 3] = (unsigned char)flags;
			}
			if (first <= io.n - 1 && io.n - 1 < first + kScanItems) { // the thread that owns the last item publishes the totals
				unsigned nb = (unsigned)(run >> 31), tt = (unsigned)(run & 0x7fffffffu);
				io.prefix[io.n] = (int)nb;
				io.bTriOff[nb] = tt;
				io.totals->nBoundary = nb;
				io.totals->triTotal = tt;
			}
		} else {
#pragma unroll
			for (int k = 0; k < kScanItems; k++) {
				if (first + k < io.n) io.off[first + k] = (unsigned)run;
				run += item_of(v[k]);
			}
			if (first <= io.n - 1 && io.n - 1 < first + kScanItems) {
				io.off[io.n] = (unsigned)run;
				if (io.totalOut) *io.totalOut = (unsigned)run;
				if (io.totalOut64) *io.totalOut64 = run;
			}
		}
	}
}

template <int MODE, int V>
__global__ void __launch_bounds__(kScanThreads) k_scan(ScanIO io) { scan_body<MODE, V>(io); }

// up to three independent offset scans in one launch (blockIdx.y selects the scan; blocks beyond a scan's tile count leave)
struct ScanIO3 { ScanIO s[3]; };
__global__ void __launch_bounds__(kScanThreads) k_scan_offs3(ScanIO3 io3)
{
	const ScanIO& io = io3.s[blockIdx.y];
	if ((long long)blockIdx.x * kScanTile >= io.n) return;
	scan_body<MODE_OFFS, 1>(io);
}

// one launch instead of a dozen cudaMemsetAsync calls: zero up to 12 buffers (16-byte granules; every buffer carries >= 32 B of slack)
struct ClearList { uint4* p[12]; unsigned long long n16[12]; };
__global__ void __launch_bounds__(256) k_clear(ClearList cl)
{
	uint4* p = cl.p[blockIdx.y];
	const unsigned long long n = cl.n16[blockIdx.y];
	const uint4 z = make_uint4(0u, 0u, 0u, 0u);
	for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * 256) p[i] = z;
}

// ------------------------------------------------------------------------------------------------ k_sort_segments
// One warp per segment [off[s], off[s]+len).  Canonical order = ascending triangle id (what the reference's CPU path
// yields, src/Object.cpp:2705-2748; the GPU path's atomic slots are nondeterministic).  len <= 32: rank sort in
// registers; longer: in-place bitonic network in the all-ascending "flip" form, valid for any length.
// UNIQUE (column lists): drops duplicates (a triangle hits several cells of one column) and stores the unique count.
constexpr int kSortSmem = 128; // longest list the warp-per-list kernel takes (rank sort in shared memory); longer ones go to k_sort_long

template <bool UNIQUE>
__device__ __forceinline__ void sort_segments_body(const unsigned* __restrict__ off, int nSeg, int* data, int* uniqueCount, int* longList, unsigned* longCount)
{
	const int lane = threadIdx.x & 31;
	const int seg = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
	if (seg >= nSeg) return;
	const unsigned beg = off[seg];
	const int len = (int)(off[seg + 1] - beg);
	int* a = data + beg;
	if (len <= 1) { if (UNIQUE && lane == 0) uniqueCount[seg] = len; return; }
	if (len <= 32) {
		int x = lane < len ? a[lane] : 0x7fffffff;
		int rank = 0;
		for (int j = 0; j < len; j++) { // lanes >= len hold INT_MAX and rank behind every entry: only the list's own entries are compared
			int y = __shfl_sync(0xffffffffu, x, j);
			rank += (y < x) || (y == x && j < lane);
		}
		__syncwarp();
		if (!UNIQUE) { if (lane < len) a[rank] = x; return; }
		// sorted value of position `lane` = the x whose rank == lane: exchange through shuffles
		int sorted = 0x7fffffff;
		for (int j = 0; j < len; j++) {
			int y = __shfl_sync(0xffffffffu, x, j), ry = __shfl_sync(0xffffffffu, rank, j);
			if (ry == lane) sorted = y;
		}
		int prev = __shfl_up_sync(0xffffffffu, sorted, 1);
		bool keep = lane < len && (lane == 0 || sorted != prev);
		unsigned km = __ballot_sync(0xffffffffu, keep);
		if (keep) a[__popc(km & ((1u << lane) - 1))] = sorted;
		if (lane == 0) uniqueCount[seg] = __popc(km);
		return;
	}
	// lists beyond kSortSmem entries (a column under a pole of a finely tessellated body collects thousands of triangles)
	// are handed to k_sort_long, one CTA of 1024 threads per list
	if (len > kSortSmem) {
		if (lane == 0) longList[atomicAdd(longCount, 1u)] = seg;
		return;
	}
	__shared__ int sSort[8][kSortSmem];
	int* dst = a; // where the sorted (and de-duplicated) list ends up
	{
		int* sb = sSort[threadIdx.x >> 5];
		for (int i = lane; i < len; i += 32) sb[i] = a[i];
		__syncwarp();
		a = sb;
	}
	// Rank (enumeration) sort: every element counts the elements that must precede it and is stored at that position.  All
	// lanes read the same a[j] at the same time (shared-memory broadcast) and there are no barriers between compare rounds:
	// for lists of up to 128 entries (<= 512 rounds per lane) it beats a bitonic network on the single warp that owns the list.
	if (!UNIQUE) {
		for (int i = lane; i < len; i += 32) {
			const int x = a[i];
			int r = 0;
			for (int j = 0; j < len; j++) { const int y = a[j]; r += (y < x) || (y == x && j < i); }
			dst[r] = x;
		}
	} else {
		// first occurrences only: keep[i] = no equal element before i; position = number of kept elements smaller than a[i]
		__shared__ unsigned char sKeep[8][kSortSmem];
		unsigned char* keep = sKeep[threadIdx.x >> 5];
		for (int i = lane; i < len; i += 32) {
			const int x = a[i];
			bool dup = false;
			for (int j = 0; j < i; j++) dup |= a[j] == x;
			keep[i] = !dup;
		}
		__syncwarp();
		int kept = 0;
		for (int i = lane; i < len; i += 32) {
			if (!keep[i]) continue;
			const int x = a[i];
			int r = 0;
			for (int j = 0; j < len; j++) r += keep[j] && a[j] < x;
			dst[r] = x;
			kept++;
		}
		kept = __reduce_add_sync(0xffffffffu, kept);
		if (lane == 0) uniqueCount[seg] = kept;
	}
}

// cell lists (blockIdx.y = 0) and column lists (blockIdx.y = 1, de-duplicated) in one launch
struct SortSeg { const unsigned* off; int nSeg; int* data; int* uniqueCount; int* longList; unsigned* longCount; };
__global__ void __launch_bounds__(256) k_sort_segments(SortSeg cells, SortSeg cols)
{
	if (blockIdx.y == 0) sort_segments_body<false>(cells.off, cells.nSeg, cells.data, nullptr, cells.longList, cells.longCount);
	else sort_segments_body<true>(cols.off, cols.nSeg, cols.data, cols.uniqueCount, cols.longList, cols.longCount);
}

// Long lists: one CTA of 1024 threads per list, same all-ascending bitonic network with CTA barriers; up to kSortLongSmem
// entries in (opt-in, 64 KB) dynamic shared memory, in place in global memory beyond that.
constexpr int kSortLongThreads = 1024, kSortLongSmem = 16384;

template <bool UNIQUE>
__device__ __forceinline__ void sort_long_body(const unsigned* __restrict__ off, const int* __restrict__ longList,
                                               const unsigned* __restrict__ longCount, int* data, int* uniqueCount, int* sLong)
{
	__shared__ int sW;
	const int tid = threadIdx.x;
	for (unsigned w = blockIdx.x; w < *longCount; w += gridDim.x) {
		const int seg = longList[w];
		const unsigned beg = off[seg];
		const int len = (int)(off[seg + 1] - beg);
		int* dst = data + beg;
		int* a = dst;
		if (len <= kSortLongSmem) {
			for (int i = tid; i < len; i += kSortLongThreads) sLong[i] = dst[i];
			a = sLong;
		}
		__syncthreads();
		for (int k = 2; (k >> 1) < len; k <<= 1) {
			for (int i = tid; i < len; i += kSortLongThreads) {
				int l = i ^ (k - 1);
				if (l > i && l < len) { int x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
			}
			__syncthreads();
			for (int j = k >> 2; j > 0; j >>= 1) {
				for (int i = tid; i < len; i += kSortLongThreads) {
					int l = i ^ j;
					if (l > i && l < len) { int x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
				}
				__syncthreads();
			}
		}
		if (UNIQUE) { // CTA-wide compaction in chunks of 1024: reads of a chunk finish (barrier) before its writes
			if (tid == 0) sW = 0;
			__syncthreads();
			for (int b0 = 0; b0 < len; b0 += kSortLongThreads) {
				const int i = b0 + tid;
				const int x = i < len ? a[i] : 0;
				const bool keep = i < len && (i == 0 || x != a[i - 1]);
				const unsigned km = __ballot_sync(0xffffffffu, keep);
				__shared__ int sWarpCnt[32];
				if ((tid & 31) == 0) sWarpCnt[tid >> 5] = __popc(km);
				__syncthreads();
				int before = 0;
				for (int q = 0; q < (tid >> 5); q++) before += sWarpCnt[q];
				int total = 0;
				for (int q = 0; q < 32; q++) total += sWarpCnt[q];
				const int base = sW;
				__syncthreads();
				if (keep) dst[base + before + __popc(km & ((1u << (tid & 31)) - 1))] = x;
				if (tid == 0) sW = base + total;
				__syncthreads();
			}
			if (tid == 0) uniqueCount[seg] = sW;
		} else if (a != dst) {
			for (int i = tid; i < len; i += kSortLongThreads) dst[i] = a[i];
		}
		__syncthreads();
	}
}

__global__ void __launch_bounds__(kSortLongThreads) k_sort_long(SortSeg cells, SortSeg cols)
{
	extern __shared__ int sLongDyn[];
	if (blockIdx.y == 0) sort_long_body<false>(cells.off, cells.longList, cells.longCount, cells.data, nullptr, sLongDyn);
	else sort_long_body<true>(cols.off, cols.longList, cols.longCount, cols.data, cols.uniqueCount, sLongDyn);
}

// ------------------------------------------------------------------------------------------------ k_l2
// K4.  Replaces CUDAClassifyInOutLevel2Kernel (cu:450-504) + CUDAClassifyTessellationLevel2Kernel (cu:403-448).
// A CTA of 256 threads refines G = max(1, 256/n2^2) boundary cells; a thread owns (cell, q, r): first as an xy-column
// (p=q', q=r') of the parity-ray phase, then as a row of n2 sub-voxels along x in the SAT phase; each row leaves as ONE
// vector store of final file bytes (n2 = 16: 128-bit; rows of a CTA are contiguous in Level2InOut.raw).
// Sub-voxel centre (cu:423-425 / 472-474): ((2p+1)*ext2 + mid) - ext1, all f32.
struct L2IO {
	const float4* tri48; const float4* ray48; const float4* plane16; const float4* aabbxy16;
	const int* boundaryIndex; const unsigned* bTriOff; const int* cellTris;
	const unsigned* colOff; const int* colCount; const int* colTris;
	const unsigned* colCellOff; const int2* colCellList; // boundary cells (slab-local rank, centre height) of every Level-1 column, CSR
	unsigned* l2Par;                                    // [boundary rank][n2*n2] parity bits along z of every sub-voxel column (k_l2_rays)
	const float4* cellMid;                              // [boundary rank] centre of the Level-1 cell (k_col_cells)
	const float* cx; const float* cy; const float* cz;
	unsigned char* l2State; // nBoundary * n2^3 file bytes (local, or the gathering rank's buffer over NVLink)
	unsigned char* l2Packed; // L2_OUT_PACKED: nBoundary * n2^3 / 4 bytes, 2 bits per sub-voxel (local, or the gathering rank's buffer)
	const int2* cellList;   // GPV_GATHER: k_l2 refines the cells cellList[bBegin .. nBoundary) (.x = boundary rank; this rank's share, grouped
	                        // by column = colCellList); null: the boundary ranks [bBegin, nBoundary) themselves
	int bBegin;             // this launch refines slots [bBegin, nBoundary)
	int nBoundary;
	Own own;                // k_col_cells / k_l2_rays: the Level-1 columns this rank refines
	Totals* totals;
};

// exact a / d for 0 <= a < 2^21 with inv = 1.f / (float)d: (a + 0.5) / d is at least 0.5/d away from an integer, the f32 error
// of the product is below (a/d) * 2^-22
__device__ __forceinline__ int fast_div(int a, float inv) { return __float2int_rz(((float)a + 0.5f) * inv); }

// Boundary cells grouped by Level-1 column: slot by atomic decrement of the per-column count (left at 0; order inside a
// column is irrelevant, every cell is refined independently).  Entry = (slab-local boundary rank, centre height of the cell).
// Also the centre of every boundary cell by rank, so that k_l2 does not decode linear indices.
// GPV_GATHER, gathering rank only: the boundary ranks of the cells every PEER refines (list[q * cap + i], i < count[q]), so that the
// peers' 2-bit blocks can be expanded peer by peer as they finish without scanning the whole grid once per peer.
struct PeerLists { int* list; unsigned* count; long long cap; };

__global__ void k_col_cells(const int* __restrict__ boundaryIndex, int nBoundary, int plane, int nx, const float* __restrict__ cx, const float* __restrict__ cy,
                            const float* __restrict__ cz, const unsigned* __restrict__ colCellOff, int* colCellCnt, int2* colCellList, float4* cellMid,
                            const int* __restrict__ colCount, const unsigned* __restrict__ bTriOff, Own own, PeerLists peers, Totals* totals)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long pairs = 0, tris = 0;
	int l1 = 0, kz = 0, col = 0;
	bool mine = false;
	if (b < nBoundary) {
		l1 = boundaryIndex[b]; kz = l1 / plane; col = l1 - kz * plane;
		mine = own((unsigned)col);
		if (mine) { pairs = (unsigned long long)colCount[col]; tris = bTriOff[b + 1] - bTriOff[b]; }
		else if (peers.list) { const int q = own.owner((unsigned)col); peers.list[(long long)q * peers.cap + atomicAdd(peers.count + q, 1u)] = b; }
	}
	pairs = warp_sum(pairs); tris = warp_sum(tris);
	if ((threadIdx.x & 31) == 0 && (pairs | tris)) { atomicAdd(&totals->l2ColPairs, pairs); atomicAdd(&totals->l2LocalTris, tris); }
	if (!mine) return;
	const int jy = col / nx, ix = col - jy * nx;
	const float mz = cz[kz];
	cellMid[b] = make_float4(cx[ix], cy[jy], mz, 0.f);
	colCellList[colCellOff[col] + atomicSub(colCellCnt + col, 1) - 1] = make_int2(b, __float_as_int(mz));
}

// K4a.  Level-2 parity rays, one thread per sub-voxel COLUMN of a Level-1 column (replaces CUDAClassifyInOutLevel2Kernel,
// cu:450-504).  The det/u/v part of the +Z ray test depends only on the xy position of the sub-column, and all boundary cells
// of one Level-1 column share their n2 x n2 sub-columns and their candidate triangles (the column list, cu:461-463).  So the
// column list is walked ONCE per Level-1 column (cessna-256: ~9 boundary cells per column with boundary cells, 24 triangles per
// list), the few triangles that pass (typically 2-4) are kept in registers, and for each boundary cell of the column each of
// them contributes its parity bits through gpv::ray_cell_mask (certified: only the sub-voxels next to the crossing are
// evaluated).  Output: one word per (boundary cell, sub-column), bit r = parity of sub-voxel r; k_l2 only loads it.
constexpr int kRaySlots = 16; // crossings kept in registers (16-bit list positions); further ones are applied in a second walk of the list (rare)
constexpr int kRayCells = 8;  // boundary cells of the column refined per register chunk
constexpr int kL2Stage = 512; // ray records staged per chunk (24 KB); a list that fits stays staged for rays_apply

struct RaySub { // one sub-voxel column: origin, list, cells, output slot, height range of the grid column
	float ox, oy, zMin, zMax, inv101, inv099; unsigned off; int cnt; unsigned cb, ce; int item; int col;
};
// sub-columns with more crossings than kRaySlots (or list positions beyond 16 bits) are finished by k_l2_rays_overflow, a warp each
struct RayOverflow { int4* list; unsigned* count; unsigned cap; }; // (column, first cell, end cell, item)

__device__ __forceinline__ void rays_apply(const GridP& g, const L2IO& io, const RayOverflow& ov, const RaySub& u, const uint4 pk0, const uint4 pk1, const unsigned n, const float4* sRec)
{
	const int rows = g.n2 * g.n2;
	const unsigned nk = min(n, (unsigned)kRaySlots);
	const unsigned fullRun = g.n2 >= 32 ? 0xffffffffu : ((1u << g.n2) - 1u);
	for (unsigned cc = u.cb; cc < u.ce; cc += kRayCells) {
		unsigned par[kRayCells];
		float midz[kRayCells];
		int bb[kRayCells];
#pragma unroll
		for (int c = 0; c < kRayCells; c++) {
			par[c] = 0u; bb[c] = -1; midz[c] = 0.f;
			if (cc + c < u.ce) { const int2 e = __ldg(io.colCellList + cc + c); bb[c] = e.x; midz[c] = __int_as_float(e.y); }
		}
		for (unsigned j = 0; j < nk; j++) {
			const unsigned w = j >> 1;
			const unsigned v = w == 0 ? pk0.x : w == 1 ? pk0.y : w == 2 ? pk0.z : w == 3 ? pk0.w : w == 4 ? pk1.x : w == 5 ? pk1.y : w == 6 ? pk1.z : pk1.w;
			RayTri s;
			const unsigned pos = (v >> ((j & 1) * 16)) & 0xffffu;
			if (sRec) { // the whole column list is still staged in shared memory (G == 1, list <= kL2Stage): no dependent global loads per crossing
				const float4 a = sRec[pos * 3], b = sRec[pos * 3 + 1], c4 = sRec[pos * 3 + 2];
				s.v1x = a.x; s.v1y = a.y; s.v1z = a.z; s.e1x = a.w; s.e1y = b.x; s.e1z = b.y; s.e2x = b.z; s.e2y = b.w;
				s.e2z = c4.x; s.det = c4.y; s.inv = c4.z; s.ok = c4.w != 0.f; s.well = c4.w == 2.f;
			} else load_ray(s, io.ray48, io.colTris[u.off + pos]);
			RayCol rc;
			if (!ray_column(s, u.ox, u.oy, rc)) continue; // cannot happen (listed because it passed); keeps rc defined
			const RayColZ k1 = ray_col_bound(s, rc, u.zMin, u.zMax, g.gsz, u.inv101, u.inv099);
			// two heights per crossing classify every cell of the chunk with two comparisons (gpv::ray_col_thresholds: all hit below
			// zAll, no hit from zNone on) ...
			const RayThr thr = ray_col_thresholds(s, rc, k1, u.zMin, u.zMax);
			unsigned undecided = 0;
#pragma unroll
			for (int c = 0; c < kRayCells; c++) {
				if (bb[c] < 0) continue;
				const float zLoC = l2_centre(0, g.h2z, midz[c], g.h1z), zHiC = l2_centre(g.n2 - 1, g.h2z, midz[c], g.h1z);
				if (zHiC < thr.zAll) par[c] ^= fullRun;
				else if (!(zLoC >= thr.zNone)) undecided |= 1u << c;
			}
			// ... the cells in between (normally the one that holds the crossing) get their masks from gpv::ray_cell_mask, one per lane
			// and round: the lanes of a warp run this part together although their crossings lie in different cells
			while (undecided) {
				const int c = __ffs(undecided) - 1;
				undecided &= undecided - 1;
				float mz = midz[0];
#pragma unroll
				for (int q = 1; q < kRayCells; q++) if (q == c) mz = midz[q];
				const unsigned m = ray_cell_mask(s, rc, k1, mz, g.h1z, g.h2z, g.n2);
#pragma unroll
				for (int q = 0; q < kRayCells; q++) if (q == c) par[q] ^= m;
			}
		}
#pragma unroll
		for (int c = 0; c < kRayCells; c++) if (bb[c] >= 0) io.l2Par[(size_t)bb[c] * rows + u.item] = par[c];
	}
	// more crossings than slots (0.02 % of cessna-256's sub-columns have more than 15), or list positions beyond 16 bits: the remaining
	// crossings are folded in by k_l2_rays_overflow, a warp per sub-column (walking the list again here, one dependent load after the
	// other, made this thread the longest-running one of the whole kernel)
	if (n > (unsigned)kRaySlots || u.cnt > 65536) {
		const unsigned at = atomicAdd(ov.count, 1u);
		if (at < ov.cap) ov.list[at] = make_int4(u.col, (int)u.cb, (int)u.ce, u.item); // (beyond the capacity the host sees the count and fails the call)
	}
}

__device__ __forceinline__ void rays_note(uint4& pk0, uint4& pk1, unsigned& n, unsigned pos)
{
	const unsigned v = (pos & 0xffffu) << ((n & 1u) * 16);
	const unsigned w = n >> 1;
	if (w == 0) pk0.x |= v; else if (w == 1) pk0.y |= v; else if (w == 2) pk0.z |= v; else if (w == 3) pk0.w |= v;
	else if (w == 4) pk1.x |= v; else if (w == 5) pk1.y |= v; else if (w == 6) pk1.z |= v; else if (w == 7) pk1.w |= v;
	n++;
}

// A Level-1 column with many boundary cells (a column that runs along a wall of the model: cessna-256 has columns with more than a
// hundred) would keep ONE CTA busy long after the rest of the grid has drained -- the kernel's time was that column's time, on one
// GPU as on eight.  So the unit of work is a column AND at most kRayChunk of its boundary cells.  k_ray_units lists the units of
// all columns that have boundary cells (so empty columns cost nothing), the expensive ones -- long column list or a full chunk --
// from the front of the array and the cheap ones from its back: CTAs are dealt out in index order, so the long units start first
// and the short ones fill the tail.  A further chunk of a column repeats the walk of its list.
constexpr int kRayChunk = 16;
struct RayWork { const int2* units; long long cap; const Totals* totals; RayOverflow ov; }; // unit = (column, first entry of colCellList); cap = array length

__global__ void __launch_bounds__(256) k_ray_units(const unsigned* __restrict__ colCellOff, const int* __restrict__ colCount, long long ncol, int2* units, long long cap, Totals* totals)
{
	const long long col = (long long)blockIdx.x * 256 + threadIdx.x;
	const int lane = threadIdx.x & 31;
	unsigned cb = 0, cnt = 0;
	int len = 0;
	if (col < ncol) { cb = colCellOff[col]; cnt = colCellOff[col + 1] - cb; len = colCount[col]; }
	const unsigned n = (cnt + kRayChunk - 1) / kRayChunk;
	const bool heavy = len > 64 || cnt >= kRayChunk;
	unsigned nh = heavy ? n : 0u, nl = heavy ? 0u : n, ih = nh, il = nl;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned a = __shfl_up_sync(0xffffffffu, ih, o), b = __shfl_up_sync(0xffffffffu, il, o);
		if (lane >= o) { ih += a; il += b; }
	}
	unsigned bh = 0, bl = 0;
	if (lane == 31) { if (ih) bh = atomicAdd(&totals->nRayHeavy, ih); if (il) bl = atomicAdd(&totals->nRayLight, il); }
	bh = __shfl_sync(0xffffffffu, bh, 31) + ih - nh; bl = __shfl_sync(0xffffffffu, bl, 31) + il - nl;
	for (unsigned k = 0; k < n; k++) {
		const int2 u = make_int2((int)col, (int)(cb + k * kRayChunk));
		if (heavy) units[bh + k] = u; else units[cap - 1 - (long long)(bl + k)] = u;
	}
}

// G == 1 (n2 >= 16): the CTA's 256 threads are 256 sub-voxel columns of ONE Level-1 column; cells [cb, ce) of colCellList
__device__ __forceinline__ void rays_unit_cta(const GridP& g, const L2IO& io, const RayOverflow& ov, float4* sStage, int col, unsigned cb, unsigned ce)
{
	const int n2 = g.n2, rows = n2 * n2, tid = threadIdx.x;
	const float invN2 = 1.f / (float)n2;
	RaySub u;
	u.cb = cb; u.ce = ce; u.col = col;
	const int jy = col / g.nx, ix = col - jy * g.nx;
	u.off = io.colOff[col]; u.cnt = io.colCount[col];
	u.zMin = io.cz[0] - g.gsz; u.zMax = io.cz[g.nz - 1] + g.gsz; u.inv101 = 1.f / (2.02f * g.h2z); u.inv099 = 1.f / (1.98f * g.h2z);
	const float mx = io.cx[ix], my = io.cy[jy];
	for (int item0 = 0; item0 < rows; item0 += 256) { // one round unless n2 = 32
		u.item = item0 + tid;
		const int q = fast_div(u.item, invN2), p = u.item - q * n2;
		u.ox = l2_centre(p, g.h2x, mx, g.h1x); u.oy = l2_centre(q, g.h2y, my, g.h1y); // cu:472-473
		uint4 pk0 = make_uint4(0u, 0u, 0u, 0u), pk1 = pk0;
		unsigned n = 0;
		for (int k0 = 0; k0 < u.cnt; k0 += kL2Stage) {
			const int nrec = min(kL2Stage, u.cnt - k0);
			__syncthreads();
			for (int i = tid; i < nrec * 3; i += 256) {
				const int rec = i / 3;
				sStage[i] = __ldg(io.ray48 + (size_t)io.colTris[u.off + k0 + rec] * 3 + (i - rec * 3));
			}
			__syncthreads();
			if (u.item < rows) {
				for (int k = 0; k < nrec; k++) {
					const float4 a = sStage[k * 3], b = sStage[k * 3 + 1], c4 = sStage[k * 3 + 2];
					if (c4.w == 0.f) continue;
					RayTri s;
					s.v1x = a.x; s.v1y = a.y; s.v1z = a.z; s.e1x = a.w; s.e1y = b.x; s.e1z = b.y; s.e2x = b.z; s.e2y = b.w;
					s.e2z = c4.x; s.det = c4.y; s.inv = c4.z; s.ok = true; s.well = false;
					RayCol rc;
					if (ray_column(s, u.ox, u.oy, rc)) rays_note(pk0, pk1, n, (unsigned)(k0 + k));
				}
			}
		}
		if (u.item < rows) rays_apply(g, io, ov, u, pk0, pk1, n, u.cnt <= kL2Stage ? sStage : nullptr);
	}
}

// G > 1 (n2 < 16): a thread is one sub-voxel column of its own Level-1 column; no CTA-wide step
__device__ __forceinline__ void rays_unit_thread(const GridP& g, const L2IO& io, const RayOverflow& ov, long long col, unsigned cb, unsigned ce, int item)
{
	const int n2 = g.n2;
	const float invN2 = 1.f / (float)n2;
	RaySub u;
	u.cb = cb; u.ce = ce; u.item = item; u.col = (int)col;
	const int jy = (int)(col / g.nx), ix = (int)(col - (long long)jy * g.nx);
	const int q = fast_div(u.item, invN2), p = u.item - q * n2;
	u.ox = l2_centre(p, g.h2x, io.cx[ix], g.h1x); u.oy = l2_centre(q, g.h2y, io.cy[jy], g.h1y);
	u.off = io.colOff[col]; u.cnt = io.colCount[col];
	u.zMin = io.cz[0] - g.gsz; u.zMax = io.cz[g.nz - 1] + g.gsz; u.inv101 = 1.f / (2.02f * g.h2z); u.inv099 = 1.f / (1.98f * g.h2z);
	uint4 pk0 = make_uint4(0u, 0u, 0u, 0u), pk1 = pk0;
	unsigned n = 0;
	// the walk is a chain of dependent loads (list entry -> ray record): four entries at a time keep four chains in flight
	for (int k0 = 0; k0 < u.cnt; k0 += 4) {
		int t[4];
		float4 ra[4], rb[4], rc4[4];
#pragma unroll
		for (int j = 0; j < 4; j++) t[j] = k0 + j < u.cnt ? io.colTris[u.off + k0 + j] : -1;
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (t[j] >= 0) { ra[j] = __ldg(io.ray48 + (size_t)t[j] * 3); rb[j] = __ldg(io.ray48 + (size_t)t[j] * 3 + 1); rc4[j] = __ldg(io.ray48 + (size_t)t[j] * 3 + 2); }
			else rc4[j] = make_float4(0.f, 0.f, 0.f, 0.f); // class 0: never hits
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (rc4[j].w == 0.f) continue;
			RayTri s;
			s.v1x = ra[j].x; s.v1y = ra[j].y; s.v1z = ra[j].z; s.e1x = ra[j].w; s.e1y = rb[j].x; s.e1z = rb[j].y; s.e2x = rb[j].z; s.e2y = rb[j].w;
			s.e2z = rc4[j].x; s.det = rc4[j].y; s.inv = rc4[j].z; s.ok = true; s.well = false;
			RayCol rc;
			if (ray_column(s, u.ox, u.oy, rc)) rays_note(pk0, pk1, n, (unsigned)(k0 + j));
		}
	}
	rays_apply(g, io, ov, u, pk0, pk1, n, nullptr);
}

// one CTA per G units (G = columns per CTA: 1 for n2 >= 16, else 256 / n2^2)
__global__ void __launch_bounds__(256, 3) k_l2_rays(GridP g, L2IO io, RayWork w)
{
	__shared__ float4 sStage[kL2Stage * 3]; // G == 1: ray records of the column list, kL2Stage at a time
	const int rows = g.n2 * g.n2;
	const int G = max(1, 256 / rows);
	const int tid = threadIdx.x;
	const long long nHeavy = w.totals->nRayHeavy, nUnits = nHeavy + w.totals->nRayLight;
	auto unit = [&](long long k) { return k < nHeavy ? w.units[k] : w.units[w.cap - 1 - (k - nHeavy)]; };
	if (G == 1) {
		if (blockIdx.x >= nUnits) return;
		const int2 e = unit(blockIdx.x);
		rays_unit_cta(g, io, w.ov, sStage, e.x, (unsigned)e.y, min(io.colCellOff[e.x + 1], (unsigned)e.y + kRayChunk));
	} else {
		const int gi = fast_div(tid, 1.f / (float)rows);
		const long long k = (long long)blockIdx.x * G + gi;
		if (gi >= G || k >= nUnits) return;
		const int2 e = unit(k);
		rays_unit_thread(g, io, w.ov, e.x, (unsigned)e.y, min(io.colCellOff[e.x + 1], (unsigned)e.y + kRayChunk), tid - gi * rows);
	}
}

// The sub-columns k_l2_rays could not finish (RayOverflow): a warp per sub-column walks the column list 32 entries at a time; the
// crossings are numbered in list order by ballot, those beyond the kRaySlots k_l2_rays has applied already (or all of them when
// the list is too long for 16-bit positions: the words are zeroed first) are XOR-ed into the sub-column's words of every cell of
// the unit.
__global__ void __launch_bounds__(256) k_l2_rays_overflow(GridP g, L2IO io, RayOverflow ov)
{
	const int lane = threadIdx.x & 31, rows = g.n2 * g.n2;
	const unsigned nEntries = min(*ov.count, ov.cap);
	const float zMin = io.cz[0] - g.gsz, zMax = io.cz[g.nz - 1] + g.gsz, inv101 = 1.f / (2.02f * g.h2z), inv099 = 1.f / (1.98f * g.h2z);
	for (unsigned w = (blockIdx.x * 256 + threadIdx.x) >> 5; w < nEntries; w += (gridDim.x * 256) >> 5) {
		const int4 e = ov.list[w];
		const int col = e.x, item = e.w;
		const unsigned cb = (unsigned)e.y, ce = (unsigned)e.z;
		const int jy = col / g.nx, ix = col - jy * g.nx;
		const int q = fast_div(item, 1.f / (float)g.n2), p = item - q * g.n2;
		const float ox = l2_centre(p, g.h2x, io.cx[ix], g.h1x), oy = l2_centre(q, g.h2y, io.cy[jy], g.h1y);
		const unsigned off = io.colOff[col];
		const int cnt = io.colCount[col];
		const bool all = cnt > 65536;
		if (all) {
			for (unsigned cc = cb + lane; cc < ce; cc += 32) io.l2Par[(size_t)io.colCellList[cc].x * rows + item] = 0u;
			__syncwarp();
		}
		unsigned seen = 0;
		for (int k0 = 0; k0 < cnt; k0 += 32) {
			const int k = k0 + lane;
			bool hit = false;
			RayTri s;
			RayCol rc;
			if (k < cnt) {
				load_ray(s, io.ray48, io.colTris[off + k]);
				hit = s.ok && ray_column(s, ox, oy, rc);
			}
			const unsigned m = __ballot_sync(0xffffffffu, hit);
			const unsigned ordinal = seen + __popc(m & ((1u << lane) - 1u));
			seen += __popc(m);
			if (hit && (all || ordinal >= (unsigned)kRaySlots)) {
				const RayColZ k1 = ray_col_bound(s, rc, zMin, zMax, g.gsz, inv101, inv099);
				for (unsigned cc = cb; cc < ce; cc++) {
					const int2 c2 = io.colCellList[cc];
					const unsigned msk = ray_cell_mask(s, rc, k1, __int_as_float(c2.y), g.h1z, g.h2z, g.n2);
					if (msk) atomicXor(io.l2Par + (size_t)c2.x * rows + item, msk);
				}
			}
		}
	}
}

#ifndef GPV_L2_MINBLOCKS
#define GPV_L2_MINBLOCKS 6 // 40 registers: 6 CTAs per SM (a 42-register build drops to 5 and measured 7 % slower on cessna 256/16)
#endif
constexpr int kL2Threads = 256;
constexpr int kL2Batch = 8;   // triangles per round of the (row, triangle) queue

// Launch constants of k_l2, computed once on the host (no integer divisions / layout arithmetic per CTA): geometry of the
// item space and the shared-memory layout (byte offsets).
struct L2K {
	int n2, rows, G, nItems;       // G cells per CTA, nItems = G*rows (cell, sub-voxel column) items
	float invRows, invN2, inv3N2;
	int sat, info, q1, q2, qn, total;
};
inline L2K l2_constants(int n2)
{
	L2K K{};
	K.n2 = n2; K.rows = n2 * n2; K.G = K.rows >= kL2Threads ? 1 : kL2Threads / K.rows; K.nItems = K.G * K.rows;
	K.invRows = 1.f / (float)K.rows; K.invN2 = 1.f / (float)n2; K.inv3N2 = 1.f / (float)(3 * n2);
	int o = K.G * 3 * n2 * 4;                // [G][3][n2] sub-voxel centres
	K.sat = o; o += K.nItems * 4;            // [nItems] SAT hit bits along z per sub-voxel column
	K.info = o; o += (K.G * 4 + 4) * 4;      // [G][2] triOff, triCnt (0 for cells past the end); [G+1] prefix of the cells' pair counts; [G] boundary rank of each cell (-1 past the end)
	o = (o + 15) & ~15;
	K.q1 = o; o += kL2Threads * kL2Batch * 8;    // (column, triangle) queue: item | rlo<<16 | rhi<<24, triangle
	K.q2 = o; o += kL2Threads * n2 * 2;          // sub-voxel queue of one slice of kL2Threads (column, triangle) entries: entry<<5 | r
	o = (o + 15) & ~15;
	K.qn = o; o += 16;                       // queue fills: [0..1] (column, triangle) queue, ping-pong; [2..3] sub-voxel queue, ping-pong
	K.total = o;
	return K;
}

// K4.  Level-2 SAT + final bytes (replaces CUDAClassifyTessellationLevel2Kernel, cu:403-448, and the 2-overwrites-1 merge with
// the parity kernel's result).  A CTA of 256 threads refines G = max(1, 256/n2^2) boundary cells; an item is one sub-voxel
// COLUMN (cell, p, q) -- the same unit k_l2_rays works on, so the SAT bits and the parity bits of a column share one word
// layout (bit r = sub-voxel r) and no transposition is needed.  Everything of the SAT that does not involve z is hoisted
// per (column, triangle) (gpv::SatCol).
// N2 > 0: n2 fixed at compile time (2, 4, 8, 16: index arithmetic by shifts, unrolled byte loop); N2 = 0: any n2 <= 32.
// OUT: how the cells' blocks leave the CTA
//   L2_OUT_BYTES   file bytes straight to local HBM (a warp's byte stores cover 32 consecutive sub-voxels)
//   L2_OUT_STAGED  file bytes staged in shared memory, every cell's n2^3 bytes leave as one run of 128-bit stores (peer memory over NVLink)
//   L2_OUT_PACKED  2 bits per sub-voxel: one uint2 (inside mask, boundary mask) per 32 consecutive sub-voxels of Level2InOut.raw, staged
//                  in shared memory, n2^3 / 4 bytes per cell at (boundary rank) * n2^3 / 4 of io.l2Packed -- a quarter of the bytes over
//                  NVLink (gather) or PCIe (host call); k_l2_expand / the host threads of gpv_voxelize_host turn it into the file bytes.
//                  Needs n2^3 % 32 == 0 (n2 a multiple of 4).
enum { L2_OUT_BYTES = 0, L2_OUT_STAGED = 1, L2_OUT_PACKED = 2 };
template <int N2, int OUT>
__global__ void __launch_bounds__(kL2Threads, GPV_L2_MINBLOCKS) k_l2(GridP g, L2IO io, L2K K)
{
	extern __shared__ __align__(16) unsigned char smemRaw[];
	const int n2 = N2 ? N2 : K.n2, rows = N2 ? N2 * N2 : K.rows, G = N2 ? (N2 * N2 >= kL2Threads ? 1 : kL2Threads / (N2 * N2)) : K.G, nItems = G * rows;
	float* sC = reinterpret_cast<float*>(smemRaw);
	unsigned* sSat = reinterpret_cast<unsigned*>(smemRaw + K.sat);
	int* sInfo = reinterpret_cast<int*>(smemRaw + K.info);
	uint2* sQ1 = reinterpret_cast<uint2*>(smemRaw + K.q1);
	unsigned short* sQ2 = reinterpret_cast<unsigned short*>(smemRaw + K.q2);
	int* sQn = reinterpret_cast<int*>(smemRaw + K.qn);
	const int tid = threadIdx.x, lane = tid & 31;
	const long long b0 = io.bBegin + (long long)blockIdx.x * G; // first slot of this CTA; slot -> boundary rank through io.cellList (GPV_GATHER) or directly
	int* sB = sInfo + 3 * G + 2;
	auto div_rows = [&](int a) { return N2 ? a / (N2 ? N2 * N2 : 1) : a / K.rows; }; // generic n2: exact (stage A divides pair indices up to rows * list length, beyond fast_div's 2^21)
	auto div_n2 = [&](int a) { return N2 ? a / (N2 ? N2 : 1) : fast_div(a, K.invN2); };
	auto div_3n2 = [&](int a) { return N2 ? a / (N2 ? 3 * N2 : 1) : fast_div(a, K.inv3N2); };
	auto rank_of = [&](long long slot) -> long long { return slot >= io.nBoundary ? -1ll : (io.cellList ? (long long)__ldg(&io.cellList[slot].x) : slot); };

	if (tid < 4) sQn[tid] = 0;
	for (int k = tid; k < G * 3 * n2; k += kL2Threads) {
		const int gi = div_3n2(k), rem = k - gi * 3 * n2, ax = div_n2(rem), p = rem - ax * n2;
		const long long b = rank_of(b0 + gi);
		float val = 0.f;
		if (b >= 0) {
			const float4 m = __ldg(io.cellMid + b);
			const float mid = ax == 0 ? m.x : (ax == 1 ? m.y : m.z);
			const float e2 = ax == 0 ? g.h2x : (ax == 1 ? g.h2y : g.h2z), e1 = ax == 0 ? g.h1x : (ax == 1 ? g.h1y : g.h1z);
			val = l2_centre(p, e2, mid, e1);
		}
		sC[k] = val;
	}
	for (int gi = tid; gi < G; gi += kL2Threads) {
		const long long b = rank_of(b0 + gi);
		int off = 0, cnt = 0;
		if (b >= 0) { off = (int)io.bTriOff[b]; cnt = (int)(io.bTriOff[b + 1] - io.bTriOff[b]); }
		sInfo[gi * 2] = off; sInfo[gi * 2 + 1] = cnt; sB[gi] = (int)b;
	}
	if (N2 != 16) __syncthreads();
	if (N2 != 16 && tid < 32) { // exclusive prefix of the cells' pair counts (rows * triangles), one warp (n2 = 16: one cell, not needed)
		int* pre = sInfo + 2 * G + 1;
		int carry = 0;
		for (int g0 = 0; g0 < G; g0 += 32) {
			const int gi = g0 + tid;
			const int v = gi < G ? rows * sInfo[gi * 2 + 1] : 0;
			int incl = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += y; }
			if (gi < G) pre[gi] = carry + incl - v;
			carry += __shfl_sync(0xffffffffu, incl, 31);
		}
		if (tid == 0) pre[G] = carry;
	}
	for (int item = tid; item < nItems; item += kL2Threads) sSat[item] = 0;
	__syncthreads();

	// ---- SAT, three stages over shared-memory queues so that every stage runs on full warps.
	//   A  every (column, triangle) pair of the cell gets the certified plane interval along z (gpv::plane_row_interval, ~12
	//      instructions); pairs that can still hit go to the (column, triangle) queue (ballot/popc per warp, one atomicAdd per warp).
	//   B1 one queue entry per thread: the z-independent predicates of the SAT (gpv::sat_col_setup: x/y AABB, the three Z-axis
	//      tests) and the certified z-AABB clip; survivors are expanded into one sub-voxel queue entry per r of their interval
	//      (warp scan of the interval lengths, one atomicAdd per warp).
	//   B2 one sub-voxel per thread: the column state is re-created (gpv::sat_col_values, no predicates) and the remaining
	//      predicates are evaluated (gpv::sat_col_test); hits are OR-ed into the column's bit mask in shared memory.
	// Triangles are taken kL2Batch at a time and queue entries kL2Threads at a time, so neither queue can overflow; the queue
	// fills are ping-pong counters, reset one round ahead, which keeps it to two barriers per round.
	{
		const float inv2h = 1.f / (2.f * g.h2z);
		int round1 = 0, round2 = 0;
		// Stage A enumerates the (column, triangle) pairs of the CTA, kL2Threads * kL2Batch per round.
		//   n2 = 16 (one cell, one column per thread): thread = column, triangles 8 at a time; the column state stays in registers.
		//   otherwise: ONE flat pair space over all cells of the CTA, pair -> (cell, triangle, column) through the prefix sums of
		//   the cells' pair counts.  A CTA holds up to 256 cells with very different list lengths (a cell at a pole of a finely
		//   tessellated body carries hundreds of triangles, its neighbours ten): the flat space keeps every thread busy whatever
		//   the skew, where a thread-per-column loop runs as long as the longest list of the CTA.
		constexpr bool kFlat = N2 != 16;
		const int* sPre = sInfo + 2 * G + 1; // [G+1] exclusive prefix of rows * triCnt
		const int pairTotal = kFlat ? sPre[G] : rows * sInfo[1];
		float cx2 = 0.f, cy2 = 0.f, cz0 = 0.f, slack = 0.f;
		if (!kFlat) {
			const int q = div_n2(tid), p = tid - q * n2;
			cx2 = sC[p]; cy2 = sC[n2 + q]; cz0 = sC[2 * n2];
			slack = 9.5367431640625e-07f * (fabsf(cz0) + 2.f * g.gsz); // 16u(|mid_z| + gs_z) >= |(c_r - c_0) - 2*h2z*r|
		}
		{
			for (int pairBase = 0; pairBase < pairTotal; pairBase += kL2Threads * kL2Batch) {
				int* q1n = sQn + (round1 & 1);
				if (tid == 0) sQn[(round1 + 1) & 1] = 0; // the other counter: every read of it lies behind a barrier, its next use after the next one
				// stage A
				for (int j = 0; j < kL2Batch; j++) {
					const int pr = pairBase + j * kL2Threads + tid;
					if (pairBase + j * kL2Threads >= pairTotal) break; // uniform
					bool alive = false;
					int rlo = 0, rhi = -1, t = 0, item = tid;
					if (pr < pairTotal) {
						int k;
						if (kFlat) {
							int lo = 0, hi = G; // largest gi with sPre[gi] <= pr
							while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sPre[mid] <= pr) lo = mid; else hi = mid; }
							const int local = pr - sPre[lo];
							k = div_rows(local);
							const int pq = local - k * rows, q = div_n2(pq), p = pq - q * n2;
							const float* c = sC + lo * 3 * n2;
							cx2 = c[p]; cy2 = c[n2 + q]; cz0 = c[2 * n2];
							slack = 9.5367431640625e-07f * (fabsf(cz0) + 2.f * g.gsz);
							item = lo * rows + pq;
							t = io.cellTris[sInfo[lo * 2] + k];
						} else {
							k = pr >> 8; // rows == kL2Threads == 256: the pair's column is this thread's
							t = io.cellTris[sInfo[0] + k];
						}
						bool inBox = true;
						if (kFlat) { // the exact x / y AABB predicates of this column (cu:284-292) from the triangle's extent: small triangles
							         // (1-2 cells across) miss most columns of a cell; n2 = 16 models have triangles larger than their cells
							const float4 bb = __ldg(io.aabbxy16 + t);
							inBox = !(bb.x - cx2 > g.h2x || bb.y - cx2 < -g.h2x || bb.z - cy2 > g.h2y || bb.w - cy2 < -g.h2y);
						}
						if (inBox) {
							const float4 A = __ldg(io.tri48 + (size_t)t * 3), pl = __ldg(io.plane16 + t);
							PlaneRec P; P.sx = pl.x; P.ny = pl.y; P.nz = pl.z; P.R = pl.w;
							alive = plane_row_interval(P, A.z - cz0, A.x - cx2, A.y - cy2, inv2h, slack, n2, rlo, rhi);
						}
					}
					const unsigned m = __ballot_sync(0xffffffffu, alive);
					if (m) {
						int base = 0;
						if (lane == 0) base = atomicAdd(q1n, __popc(m));
						base = __shfl_sync(0xffffffffu, base, 0);
						if (alive) sQ1[base + __popc(m & ((1u << lane) - 1))] = make_uint2((unsigned)item | ((unsigned)rlo << 16) | ((unsigned)rhi << 24), (unsigned)t);
					}
				}
				__syncthreads();
				const int n1 = *q1n;
				round1++;
				for (int s0 = 0; s0 < n1; s0 += kL2Threads) {
					int* q2n = sQn + 2 + (round2 & 1);
					// stage B1
					int len = 0, rlo = 0;
					if (s0 + tid < n1) {
						const uint2 e = sQ1[s0 + tid];
						const int it = (int)(e.x & 0xffffu);
						const int gi = div_rows(it), pq = it - gi * rows, q = div_n2(pq), p = pq - q * n2;
						const float* c = sC + gi * 3 * n2;
						const float4 A = __ldg(io.tri48 + (size_t)e.y * 3), B = __ldg(io.tri48 + (size_t)e.y * 3 + 1), C = __ldg(io.tri48 + (size_t)e.y * 3 + 2);
						SatCol s;
						if (sat_col_setup(s, c[p], c[n2 + q], g.h2x, g.h2y, A.x, A.y, B.x, B.y, C.x, C.y)) {
							int rhi = (int)(e.x >> 24);
							rlo = (int)((e.x >> 16) & 0xffu);
							const float sl = 9.5367431640625e-07f * (fabsf(c[2 * n2]) + 2.f * g.gsz);
							axis_clip(fminf(A.z, fminf(B.z, C.z)), fmaxf(A.z, fmaxf(B.z, C.z)), c[2 * n2], g.h2z, g.gsz, inv2h, sl, n2, rlo, rhi);
							len = max(0, rhi - rlo + 1);
						}
					}
					int incl = len;
#pragma unroll
					for (int o = 1; o < 32; o <<= 1) {
						const int y = __shfl_up_sync(0xffffffffu, incl, o);
						if (lane >= o) incl += y;
					}
					const int warpTot = __shfl_sync(0xffffffffu, incl, 31);
					if (warpTot) {
						int base = 0;
						if (lane == 31) base = atomicAdd(q2n, warpTot);
						base = __shfl_sync(0xffffffffu, base, 31) + incl - len;
						for (int j = 0; j < len; j++) sQ2[base + j] = (unsigned short)((tid << 5) | (rlo + j));
					}
					__syncthreads();
					const int n2q = *q2n;
					if (tid == 0) sQn[2 + ((round2 + 1) & 1)] = 0;
					round2++;
					// stage B2
					for (int v = tid; v < n2q; v += kL2Threads) {
						const unsigned x = sQ2[v];
						const uint2 e = sQ1[s0 + (int)(x >> 5)];
						const int r = (int)(x & 31u), it = (int)(e.x & 0xffffu);
						const int gi = div_rows(it), pq = it - gi * rows, q = div_n2(pq), p = pq - q * n2;
						const float* c = sC + gi * 3 * n2;
						const float4 A = __ldg(io.tri48 + (size_t)e.y * 3), B = __ldg(io.tri48 + (size_t)e.y * 3 + 1), C = __ldg(io.tri48 + (size_t)e.y * 3 + 2);
						SatCol s;
						sat_col_values(s, c[p], c[n2 + q], g.h2x, g.h2y, A.x, A.y, B.x, B.y, C.x, C.y);
						if (sat_col_test(s, c[2 * n2 + r], g.h2x, g.h2y, g.h2z, A.z, B.z, C.z)) atomicOr(sSat + it, 1u << r);
					}
					__syncthreads();
				}
			}
		}
	}
	__syncthreads();

	// ---- the file bytes: SAT hit 254, else inside 127, else 0 (2 overwrites 1: src/Object.cpp:2603-2606)
	// four sub-voxels at once: (nibble * 0x204081) & 0x01010101 spreads bit k of the nibble to byte k
	auto four = [](unsigned par, unsigned sat, int r) { return (((par >> r) & 15u) * 0x204081u & 0x01010101u) * 127u + (((sat >> r) & 15u) * 0x204081u & 0x01010101u) * 254u; };
	unsigned nIn = 0, nBd = 0;
	if (OUT == L2_OUT_PACKED) {
		uint2* sPk = reinterpret_cast<uint2*>(smemRaw + K.q1); // [G][n2^3 / 32] (inside, boundary) masks in file order
		const int n23 = rows * n2, wpc = n23 >> 5;           // words per cell
		constexpr bool kBallot = N2 == 16 || N2 == 8;          // rows % 32 == 0: a warp holds 32 consecutive columns of one cell
		if (!kBallot) {
			for (int k = tid; k < G * wpc; k += kL2Threads) sPk[k] = make_uint2(0u, 0u);
			__syncthreads();
		}
		for (int item0 = 0; item0 < nItems; item0 += kL2Threads) { // (uniform trip count: ballots inside)
			const int item = item0 + tid;
			const int gi = item < nItems ? div_rows(item) : 0, pq = item - gi * rows;
			const long long b = item < nItems ? sB[gi] : -1;
			unsigned sat = 0u, par = 0u;
			if (b >= 0) { sat = sSat[item]; par = io.l2Par[(size_t)b * rows + pq] & ~sat; }
			nIn += __popc(par); nBd += __popc(sat);
			if (kBallot) {
#pragma unroll
				for (int r = 0; r < (N2 ? N2 : 1); r++) {
					const unsigned mi = __ballot_sync(0xffffffffu, (par >> r) & 1u), mb = __ballot_sync(0xffffffffu, (sat >> r) & 1u);
					if (lane == (r & 31)) sPk[gi * wpc + ((r * rows + pq) >> 5)] = make_uint2(mi, mb); // lane r's own (gi, pq) lies in the same word as lane 0's
				}
			} else if (N2 == 4) { // rows = 16: lanes 0-15 are one cell, 16-31 the next; a word is two z-layers of 16 columns
#pragma unroll
				for (int k = 0; k < 2; k++) {
					const unsigned i0 = __ballot_sync(0xffffffffu, (par >> (2 * k)) & 1u), i1 = __ballot_sync(0xffffffffu, (par >> (2 * k + 1)) & 1u);
					const unsigned b0m = __ballot_sync(0xffffffffu, (sat >> (2 * k)) & 1u), b1m = __ballot_sync(0xffffffffu, (sat >> (2 * k + 1)) & 1u);
					if (lane == 0) sPk[gi * wpc + k] = make_uint2((i0 & 0xffffu) | (i1 << 16), (b0m & 0xffffu) | (b1m << 16));
					if (lane == 16) sPk[gi * wpc + k] = make_uint2((i0 >> 16) | (i1 & 0xffff0000u), (b0m >> 16) | (b1m & 0xffff0000u));
				}
			} else if (b >= 0) { // any other n2 with n2^3 % 32 == 0: bit by bit
				unsigned* w = reinterpret_cast<unsigned*>(sPk);
				for (int r = 0; r < n2; r++) {
					const int v = gi * n23 + r * rows + pq;
					if ((par >> r) & 1u) atomicOr(w + (v >> 5) * 2, 1u << (v & 31));
					if ((sat >> r) & 1u) atomicOr(w + (v >> 5) * 2 + 1, 1u << (v & 31));
				}
			}
		}
		__syncthreads();
		const int nValid = (int)max(0ll, min((long long)G, (long long)io.nBoundary - b0));
		const int perCell = n23 >> 2, total = nValid * perCell; // bytes; perCell is a multiple of 16
		const unsigned char* src = reinterpret_cast<const unsigned char*>(sPk);
		for (int i = tid * 16; i < total; i += kL2Threads * 16) {
			const int gi = i / perCell, off = i - gi * perCell;
			*reinterpret_cast<uint4*>(io.l2Packed + (size_t)sB[gi] * perCell + off) = *reinterpret_cast<const uint4*>(src + i);
		}
	} else if (OUT == L2_OUT_STAGED) {
		// GPV_GATHER: the bytes go over NVLink into the gathering rank's buffer.  Every thread expands its column into the CTA's block
		// of Level2InOut.raw staged in shared memory (the queue area, idle now; the cells of a CTA are consecutive boundary ranks, so
		// the block is contiguous in the file); the block then leaves as 128-bit coalesced stores, the granularity NVLink likes.
		unsigned char* sOut = smemRaw + K.q1;
		const int n23 = rows * n2;
		for (int item = tid; item < nItems; item += kL2Threads) {
			const int gi = div_rows(item), pq = item - gi * rows;
			const long long b = sB[gi];
			if (b < 0) continue;
			const unsigned sat = sSat[item];
			const unsigned par = io.l2Par[(size_t)b * rows + pq] & ~sat;
			nIn += __popc(par); nBd += __popc(sat);
			unsigned char* o = sOut + gi * n23 + pq;
			int r = 0;
			for (; r + 4 <= n2; r += 4) {
				const unsigned w = four(par, sat, r);
				o[r * rows] = (unsigned char)w; o[(r + 1) * rows] = (unsigned char)(w >> 8);
				o[(r + 2) * rows] = (unsigned char)(w >> 16); o[(r + 3) * rows] = (unsigned char)(w >> 24);
			}
			for (; r < n2; r++) o[r * rows] = ((sat >> r) & 1) ? 254 : (((par >> r) & 1) ? 127 : 0);
		}
		__syncthreads();
		// every cell's block is n2^3 contiguous bytes of Level2InOut.raw at (boundary rank) * n2^3 -- the ranks are global (every rank
		// scans the whole grid), so no offset has to be exchanged; the cells of a CTA are not neighbours in the file (they are
		// grouped by column), each leaves as its own run of 128-bit stores
		const int nValid = (int)max(0ll, min((long long)G, (long long)io.nBoundary - b0));
		const int total = nValid * n23;
		if ((n23 & 15) == 0) {
			for (int i = tid * 16; i < total; i += kL2Threads * 16) {
				const int gi = i / n23, off = i - gi * n23;
				*reinterpret_cast<uint4*>(io.l2State + (size_t)sB[gi] * n23 + off) = *reinterpret_cast<const uint4*>(sOut + i);
			}
		} else if ((n23 & 7) == 0) {
			for (int i = tid * 8; i < total; i += kL2Threads * 8) {
				const int gi = i / n23, off = i - gi * n23;
				*reinterpret_cast<uint2*>(io.l2State + (size_t)sB[gi] * n23 + off) = *reinterpret_cast<const uint2*>(sOut + i);
			}
		} else {
			for (int i = tid; i < total; i += kL2Threads) {
				const int gi = i / n23, off = i - gi * n23;
				io.l2State[(size_t)sB[gi] * n23 + off] = sOut[i];
			}
		}
	} else {
		// local HBM: a warp's byte stores cover 32 consecutive sub-voxels of the file (whole sectors); measured 10 % faster for the
		// kernel than the staged form (no barrier, the warps retire independently)
		for (int item = tid; item < nItems; item += kL2Threads) {
			const int gi = div_rows(item), pq = item - gi * rows;
			const long long b = sB[gi];
			if (b < 0) continue;
			const unsigned sat = sSat[item];
			const unsigned par = io.l2Par[(size_t)b * rows + pq] & ~sat;
			nIn += __popc(par); nBd += __popc(sat);
			unsigned char* o = io.l2State + ((size_t)b * rows * n2 + pq);
			int r = 0;
			for (; r + 4 <= n2; r += 4) {
				const unsigned w = four(par, sat, r);
				o[(size_t)r * rows] = (unsigned char)w; o[(size_t)(r + 1) * rows] = (unsigned char)(w >> 8);
				o[(size_t)(r + 2) * rows] = (unsigned char)(w >> 16); o[(size_t)(r + 3) * rows] = (unsigned char)(w >> 24);
			}
			for (; r < n2; r++) o[(size_t)r * rows] = ((sat >> r) & 1) ? 254 : (((par >> r) & 1) ? 127 : 0);
		}
	}
	nIn = __reduce_add_sync(0xffffffffu, nIn); nBd = __reduce_add_sync(0xffffffffu, nBd);
	if (lane == 0 && (nIn | nBd)) { atomicAdd(&io.totals->l2Inside, (unsigned long long)nIn); atomicAdd(&io.totals->l2Boundary, (unsigned long long)nBd); }
}

// Level-2 blocks this rank refined (cellList[0 .. nCells)), bytesPerCell each, from a local whole-grid array to the
// same offsets of the gathering rank's array: used when the blocks are needed locally as well (GPV_NORMALS reads the states back).
__global__ void __launch_bounds__(256) k_scatter_blocks(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, const int2* __restrict__ cellList,
                                                         long long nCells, int bytesPerCell)
{
	const int unit = (bytesPerCell & 7) ? 1 : 8, per = bytesPerCell / unit; // 8 bytes at a time when the blocks allow it (any even n2)
	for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nCells * per; i += (long long)gridDim.x * 256) {
		const long long cell = i / per;
		const size_t off = (size_t)cellList[cell].x * bytesPerCell + (size_t)(i - cell * per) * unit;
		if (unit == 8) *reinterpret_cast<uint2*>(dst + off) = *reinterpret_cast<const uint2*>(src + off);
		else dst[off] = src[off];
	}
}

// ------------------------------------------------------------------------------------------------ gather over peer memory
// Multi-GPU: every rank writes its share of the streams straight into the gathering rank's buffers over NVLink peer
// memory, from inside the kernels that produce them.  Level 1 is computed by every rank over the whole grid (the parity rays need
// whole column lists, cu:461-463), so every rank knows the GLOBAL boundary ranks and prefix sums: no offset has to be exchanged.
// What is shared out is the output: the Level-1 bytes and prefix sums by z-slab (contiguous byte ranges), the Level-2 refinement by
// Level-1 column (struct Own) -- each block lands at (global boundary rank) * n2^3.  The mailbox in the gathering rank's memory
// carries three kinds of flags, all tagged with the call's epoch (every rank counts its GPV_GATHER calls):
//   begin   rank 0 has entered call `epoch`: its previous result has been consumed, peers may overwrite the buffers
//   done[r] rank r's last store of call `epoch` is visible (system-scope fence before the flag)
//   stat    rank r's share of the counts (Level-1 inside cells of its slab, Level-2 inside / boundary voxels of its columns), summed by rank 0
struct GatherMail { unsigned long long begin; unsigned long long done[16]; unsigned long long stat[2][16][4]; };
constexpr unsigned long long kGatherTimeoutNsDefault = 5000000000ull; // a missing peer must not hang the GPU (gpv_gather_set_timeout)

__device__ __forceinline__ unsigned long long ld_sys(const unsigned long long* p)
{
	unsigned long long v;
	asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_sys(unsigned long long* p, unsigned long long v)
{
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

// <<<1, 1>>> before this rank's first store into the gathering rank's buffers.  Rank 0 announces the call; the others wait for the
// announcement, so that a fast rank cannot overwrite a result rank 0's caller is still reading (rank 0 enters its next call only
// after the previous one returned).
__global__ void k_gather_begin(GatherMail* mail, int rank, unsigned epoch, unsigned long long timeoutNs, Totals* totals)
{
	if (rank == 0) { st_sys(&mail->begin, (unsigned long long)epoch); return; }
	const unsigned long long t0 = global_ns();
	while (ld_sys(&mail->begin) < (unsigned long long)epoch)
		if (global_ns() - t0 > timeoutNs) { totals->gatherError = 2; return; }
}

// <<<1, 1>>> behind the last kernel of the call: everything this rank wrote to the gathering rank is ordered before the flag
__global__ void k_gather_done(GatherMail* mail, int rank, unsigned epoch, const Totals* totals)
{
	unsigned long long* st = mail->stat[epoch & 1][rank];
	st[0] = totals->l1Inside; st[1] = totals->l2Inside; st[2] = totals->l2Boundary; st[3] = totals->gatherError;
	__threadfence_system();
	st_sys(&mail->done[rank], (unsigned long long)epoch);
}

// Gathering rank: the peers sent their Level-2 blocks as 2 bits per sub-voxel (L2_OUT_PACKED); these two turn them into the file
// bytes, peer by peer: a one-thread kernel waits for rank q's completion flag, then k_gather_expand expands the blocks of the
// cells rank q refined (listed per peer by k_col_cells; one thread per 32 sub-voxels: a uint2 in, two 128-bit stores out) -- the
// blocks of the ranks that finish early are expanded while the later ones are still computing.  (Filtering the whole grid by owner
// once per peer instead cost 47 us per peer: 0.33 ms of waiting and scanning on rank 0 at 8 ranks.)  (A single resident grid whose threads spin on the flags was tried
// and dropped: with several ranks on ONE device, as in the tests, the spinning grid starves the ranks it waits for.)
__global__ void k_gather_wait_rank(GatherMail* mail, int q, unsigned epoch, unsigned long long timeoutNs, Totals* totals)
{
	const unsigned long long t0 = global_ns();
	while (ld_sys(&mail->done[q]) < (unsigned long long)epoch)
		if (global_ns() - t0 > timeoutNs) { totals->gatherError = 2; return; }
}

__global__ void __launch_bounds__(256) k_gather_expand(const uint2* __restrict__ packed, unsigned char* __restrict__ bytes, int wordsPerCell, const int* __restrict__ cells,
                                                        const unsigned* __restrict__ nCells)
{
	const long long nWords = (long long)*nCells * wordsPerCell; // the blocks rank q refined (k_col_cells listed them)
	for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < nWords; i += (long long)gridDim.x * 256) {
		const long long c = i / wordsPerCell;
		const long long w = (long long)cells[c] * wordsPerCell + (i - c * wordsPerCell);
		const uint2 m = __ldcg(packed + w); // (written by a peer over NVLink: not through the read-only path)
		unsigned o[8];
#pragma unroll
		for (int k = 0; k < 8; k++)
			o[k] = (((m.x >> (4 * k)) & 15u) * 0x204081u & 0x01010101u) * 127u + (((m.y >> (4 * k)) & 15u) * 0x204081u & 0x01010101u) * 254u;
		uint4* dst = reinterpret_cast<uint4*>(bytes + w * 32);
		dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
		dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
	}
}

// <<<1, 1>>> on the gathering rank: returns when every rank has signalled completion of this epoch; the counts become the whole grid's
__global__ void k_gather_wait(GatherMail* mail, int world, unsigned epoch, unsigned long long timeoutNs, Totals* totals)
{
	const unsigned long long t0 = global_ns();
	unsigned long long in1 = 0, in2 = 0, bd2 = 0;
	for (int q = 0; q < world; q++) {
		while (ld_sys(&mail->done[q]) < (unsigned long long)epoch) {
			if (global_ns() - t0 > timeoutNs) { totals->gatherError = 2; return; }
		}
		const unsigned long long* st = mail->stat[epoch & 1][q];
		in1 += st[0]; in2 += st[1]; bd2 += st[2];
		if (st[3]) totals->gatherError = st[3];
	}
	totals->l1Inside = in1; totals->l2Inside = in2; totals->l2Boundary = bd2;
}

// ------------------------------------------------------------------------------------------------ normals
// K5a. Level-1 normals (src/Object.cpp:3219-3253): normalise(mean of unit face normals), list in ascending order.
// One thread per boundary cell; the other cells keep the memset value 127 (= uchar(0*85.33+127)).
__global__ void k_l1_normals(const float4* __restrict__ tri48, const int* __restrict__ boundaryIndex, const unsigned* __restrict__ bTriOff,
                             const int* __restrict__ cellTris, int nBoundary, long long globalBase, unsigned char* l1Normal)
{
	int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nBoundary) return;
	float sx = 0, sy = 0, sz = 0;
	unsigned beg = bTriOff[b], end = bTriOff[b + 1];
	for (unsigned k = beg; k < end; k++) {
		int t = cellTris[k];
		float4 A = __ldg(tri48 + (size_t)t * 3), B = __ldg(tri48 + (size_t)t * 3 + 1), C = __ldg(tri48 + (size_t)t * 3 + 2);
		float ax = B.x - A.x, ay = B.y - A.y, az = B.z - A.z, bx = C.x - A.x, by = C.y - A.y, bz = C.z - A.z;
		float nx = ay * bz - by * az, ny = az * bx - bz * ax, nz = ax * by - bx * ay; // VectorCrossProduct, FloatVector.h:293
		normalize3(nx, ny, nz);
		sx += nx; sy += ny; sz += nz;
	}
	float nt = (float)(end - beg);
	float ax = sx / nt, ay = sy / nt, az = sz / nt;
	normalize3(ax, ay, az);
	size_t li = (size_t)((long long)boundaryIndex[b] - globalBase);
	l1Normal[li * 3] = encode_normal(ax); l1Normal[li * 3 + 1] = encode_normal(ay); l1Normal[li * 3 + 2] = encode_normal(az);
}

// K5b. Level-2 normals: only boundary sub-voxels (SAT bit set) visit the cell list, accumulating the UN-normalised cross(e01,e02) of
// every hit in ascending order (cu:311-318: normalize() result discarded; cu:40-46), then the host averaging of
// src/Object.cpp:2613-2632; the others get 127,127,127.  A CTA takes 16,384 consecutive sub-voxels of the cells this call refined
// (slots, like k_l2): every thread writes the neutral value for its sub-voxels and appends the boundary ones to a dense list in
// shared memory; the list is then worked off by full warps (6 % of cessna's sub-voxels are boundary: one thread per sub-voxel
// walking the list ran at a few lanes per warp and took 2.8 ms, five times the SAT kernel that found the hits).
// A triangle is put through the full 13-predicate SAT only where the certified plane interval of its sub-voxel column (the one
// k_l2 culls with, gpv::plane_row_interval) contains the sub-voxel: the predicate fails outside it, so the set of hits -- and with
// the ascending list order the f32 sums -- are those of the reference's loop over the whole list.
// `state`: file bytes (254 = boundary), or null with `packed` = the 2-bit words of L2_OUT_PACKED (boundary mask in .y).
constexpr int kNormalVoxels = 16384; // (four cessna cells: ~1,000 boundary sub-voxels in the dense list, four full rounds of the CTA)
__global__ void __launch_bounds__(256) k_l2_normals(GridP g, L2IO io, const unsigned char* __restrict__ state, const uint2* __restrict__ packed,
                                                     unsigned char* __restrict__ l2Normal)
{
	__shared__ unsigned short sList[kNormalVoxels];
	__shared__ int sCount;
	const int n2 = g.n2, n23 = n2 * n2 * n2, tid = threadIdx.x;
	const long long total = (long long)(io.nBoundary - io.bBegin) * n23, V0 = (long long)blockIdx.x * kNormalVoxels;
	const long long slot0 = V0 / n23;
	const int rem0 = (int)(V0 - slot0 * n23);
	if (tid == 0) sCount = 0;
	__syncthreads();
	auto locate = [&](int i, long long& b, int& loc) { // sub-voxel i of this CTA -> (boundary rank, index inside the block)
		const int s = (rem0 + i) / n23;
		loc = rem0 + i - s * n23;
		const long long slot = io.bBegin + slot0 + s;
		b = io.cellList ? (long long)__ldg(&io.cellList[slot].x) : slot;
	};
	for (int i = tid; i < kNormalVoxels && V0 + i < total; i += 256) {
		long long b; int loc;
		locate(i, b, loc);
		const long long gv = b * n23 + loc; // index in Level2InOut.raw
		const bool boundary = state ? state[gv] == 254 : ((__ldcg(&packed[gv >> 5].y) >> (gv & 31)) & 1u) != 0;
		l2Normal[gv * 3] = 127; l2Normal[gv * 3 + 1] = 127; l2Normal[gv * 3 + 2] = 127;
		if (boundary) sList[atomicAdd(&sCount, 1)] = (unsigned short)i;
	}
	__syncthreads(); // (also orders this CTA's neutral bytes before the boundary sub-voxels' stores below)
	const float inv2h = 1.f / (2.f * g.h2z);
	for (int k = tid; k < sCount; k += 256) {
		long long b; int loc;
		locate((int)sList[k], b, loc);
		const int r = loc / (n2 * n2), pq = loc - r * n2 * n2, q = pq / n2, p = pq - q * n2;
		const float4 mid = __ldg(io.cellMid + b);
		const float cxv = l2_centre(p, g.h2x, mid.x, g.h1x), cyv = l2_centre(q, g.h2y, mid.y, g.h1y), czv = l2_centre(r, g.h2z, mid.z, g.h1z);
		const float cz0 = l2_centre(0, g.h2z, mid.z, g.h1z);
		const float slack = 9.5367431640625e-07f * (fabsf(cz0) + 2.f * g.gsz);
		float sx = 0, sy = 0, sz = 0, cnt = 0;
		const unsigned kEnd = io.bTriOff[b + 1];
		for (unsigned kk = io.bTriOff[b]; kk < kEnd; kk++) {
			const int t = io.cellTris[kk];
			const float4 A = __ldg(io.tri48 + (size_t)t * 3), pl = __ldg(io.plane16 + t);
			PlaneRec P; P.sx = pl.x; P.ny = pl.y; P.nz = pl.z; P.R = pl.w;
			int rlo, rhi;
			if (!plane_row_interval(P, A.z - cz0, A.x - cxv, A.y - cyv, inv2h, slack, n2, rlo, rhi) || r < rlo || r > rhi) continue; // certified: the plane predicate fails here
			const float4 B = __ldg(io.tri48 + (size_t)t * 3 + 1), C = __ldg(io.tri48 + (size_t)t * 3 + 2);
			if (tri_box_overlap(cxv, cyv, czv, g.h2x, g.h2y, g.h2z, A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z)) {
				float ax = B.x - A.x, ay = B.y - A.y, az = B.z - A.z, bx = C.x - A.x, by = C.y - A.y, bz = C.z - A.z;
				sx += ay * bz - az * by; sy += az * bx - ax * bz; sz += ax * by - ay * bx; // cutil_math.h:409 cross()
				cnt += 1;
			}
		}
		if (cnt > 0) {
			float ax = sx / cnt, ay = sy / cnt, az = sz / cnt;
			normalize3(ax, ay, az);
			const long long gv = b * n23 + loc;
			l2Normal[gv * 3] = encode_normal(ax); l2Normal[gv * 3 + 1] = encode_normal(ay); l2Normal[gv * 3 + 2] = encode_normal(az);
		}
	}
}

} // namespace gpv
