// gpview_b200/csrc/gpv_kernels.cuh -- hand-written sm_100a kernels of the two-level voxelizer hot path.
//
// Compiled ONLY with  nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (see DESIGN.md "Numerics").
// Kernel inventory (SURVEY.md 2.3 "new kernels"):
//   k_tables        per-axis Level-1 cell-centre tables (kills all FP64 / int->float work in the inner loops)
//   k_repack        36 B flat triangles -> 48 B float4x3 records (TMA-able) + 48 B +Z ray records
//   k_bin<FILL>     K1  triangle -> Level-1 cell SAT binning (count / fill sweeps), TMA-staged triangle tiles
//   k_cross<FILL>   K2a certified (column, triangle) crossing detection for the parity fill
//   k_fill_sweep    K2b +Z parity sweep per Level-1 column, coalesced along x, final Level-1 state bytes
//   k_scan<MODE>    K3  single-pass decoupled-look-back scan: boundary prefix sum / index compaction / CSR offsets
//   k_sort_segments canonical (ascending) order of every cell / column list; de-duplicates column lists
//   k_l2            K4  Level-2 refinement: parity rays then hoisted SAT per sub-voxel row, 128-bit row stores
//   k_l1_normals, k_l2_normals   K5 normals in the reference's uchar encoding
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
	unsigned long long l1Tests, l1Hits, colPairsOver, crossPairs, nIll, l1Inside, l2Inside, l2Boundary, l2BoxTests, l2RayTests;
	unsigned int nBoundary, triTotal, colTotalOver, crossTotal;
};

constexpr int kBinThreads = 128;  // triangles per tile
constexpr int kBigFootprint = 32; // cells; larger footprints are spread over the warp

// ------------------------------------------------------------------------------------------------ TMA tile load
// One elected thread issues a 1-D bulk async copy (TMA, SASS UBLKCP) of `bytes` (multiple of 16) from global to shared
// memory; completion is signalled on an mbarrier that every thread of the CTA then waits on.
__device__ __forceinline__ void tma_load_tile(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar, int tid)
{
	uint32_t barAddr = (uint32_t)__cvta_generic_to_shared(bar);
	uint32_t dstAddr = (uint32_t)__cvta_generic_to_shared(smemDst);
	if (tid == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barAddr));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		             ::"r"(dstAddr), "l"(gmemSrc), "r"(bytes), "r"(barAddr) : "memory");
	}
	uint32_t done = 0;
	while (!done) {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(barAddr) : "memory");
	}
}

// ------------------------------------------------------------------------------------------------ k_tables
// centre[p] = fl32((p + 0.5) * ext * 2 + min) evaluated in double exactly like cu:382-384 (== ray origin
// src/Object.cpp:743-745 == mid point :2567-2569, see oracle/gpv_oracle.c gpvo_axis_table).
__global__ void k_tables(GridP g, float* cx, float* cy, float* cz)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < g.nx) cx[i] = (float)((i + 0.5) * (double)g.h1x * 2 + (double)g.minx);
	if (i < g.ny) cy[i] = (float)((i + 0.5) * (double)g.h1y * 2 + (double)g.miny);
	if (i < g.nz) cz[i] = (float)((i + 0.5) * (double)g.h1z * 2 + (double)g.minz);
}

// ------------------------------------------------------------------------------------------------ k_repack
// flat float[9] per triangle (src/Object.cpp:3496-3527 layout) -> tri48 (v0|v1|v2 as float4, w = 0) and ray48
// (v1xyz e1xyz e2xyz det inv ok), both 16-byte aligned records so that tiles can be moved by TMA bulk copies.
__global__ void __launch_bounds__(256) k_repack(const float* __restrict__ flat, long long nTri, float4* __restrict__ tri48, float4* __restrict__ ray48)
{
	__shared__ float s[256 * 9];
	long long base = (long long)blockIdx.x * 256;
	int n = (int)min((long long)256, nTri - base);
	for (int i = threadIdx.x; i < n * 9; i += 256) s[i] = flat[base * 9 + i]; // coalesced
	__syncthreads();
	int t = threadIdx.x;
	if (t >= n) return;
	const float* v = s + t * 9;
	tri48[(base + t) * 3 + 0] = make_float4(v[0], v[1], v[2], 0.f);
	tri48[(base + t) * 3 + 1] = make_float4(v[3], v[4], v[5], 0.f);
	tri48[(base + t) * 3 + 2] = make_float4(v[6], v[7], v[8], 0.f);
	RayTri r;
	ray_tri_setup(r, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8]);
	ray48[(base + t) * 3 + 0] = make_float4(r.v1x, r.v1y, r.v1z, r.e1x);
	ray48[(base + t) * 3 + 1] = make_float4(r.e1y, r.e1z, r.e2x, r.e2y);
	ray48[(base + t) * 3 + 2] = make_float4(r.e2z, r.det, r.inv, r.ok ? 1.f : 0.f);
}

__device__ __forceinline__ void load_ray(RayTri& r, const float4* __restrict__ ray48, int t)
{
	float4 a = __ldg(ray48 + (size_t)t * 3), b = __ldg(ray48 + (size_t)t * 3 + 1), c = __ldg(ray48 + (size_t)t * 3 + 2);
	r.v1x = a.x; r.v1y = a.y; r.v1z = a.z; r.e1x = a.w; r.e1y = b.x; r.e1z = b.y; r.e2x = b.z; r.e2y = b.w;
	r.e2z = c.x; r.det = c.y; r.inv = c.z; r.ok = c.w != 0.f;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
	for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
	return v;
}

// ------------------------------------------------------------------------------------------------ k_bin
// K1.  Replaces CUDAClassifyTessellationKernel (cu:320-401) + the host CSR flatten (src/Object.cpp:2137-2180).
// One thread owns one triangle of a 128-triangle tile (staged by TMA); footprints larger than kBigFootprint cells are
// spread over the 32 lanes of the warp (cessna-256: mean 64, max 6,762 cells per triangle).  Two sweeps:
//   FILL=false  cellCount[cell]++ (slab cells only), colCount[col]++ (every hit: column lists are de-duplicated later)
//   FILL=true   cellTris[bTriOff[prefix[cell]] + slot], colTris[colOff[col] + slot]   (slots by atomic decrement)
struct BinOut {
	int* cellCount;            // slab-local linear index
	int* colCount;             // nx*ny
	const int* prefix;         // FILL: slab-local boundary rank of each cell
	const unsigned* bTriOff;   // FILL
	int* cellTris;             // FILL
	const unsigned* colOff;    // FILL
	int* colTris;              // FILL
	Totals* totals;
};

template <bool FILL>
__device__ __forceinline__ void bin_emit(const GridP& g, const BinOut& o, int p, int q, int r, int t)
{
	int col = q * g.nx + p;
	if (!FILL) {
		atomicAdd(o.colCount + col, 1);
		if (r >= g.z0 && r < g.z1) atomicAdd(o.cellCount + ((size_t)(r - g.z0) * g.ny * g.nx + col), 1);
	} else {
		int slot = atomicSub(o.colCount + col, 1) - 1;
		o.colTris[o.colOff[col] + slot] = t;
		if (r >= g.z0 && r < g.z1) {
			size_t li = (size_t)(r - g.z0) * g.ny * g.nx + col;
			int s2 = atomicSub(o.cellCount + li, 1) - 1;
			o.cellTris[o.bTriOff[o.prefix[li]] + s2] = t;
		}
	}
}

template <bool FILL>
__global__ void __launch_bounds__(kBinThreads) k_bin(const float4* __restrict__ tri48, int nTri, GridP g,
                                                     const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz, BinOut o)
{
	__shared__ __align__(128) float4 sTri[kBinThreads * 3];
	__shared__ __align__(8) uint64_t bar;
	const int tid = threadIdx.x, lane = tid & 31;
	const int base = blockIdx.x * kBinThreads;
	const int n = min(kBinThreads, nTri - base);
	tma_load_tile(sTri, tri48 + (size_t)base * 3, (uint32_t)n * 48u, &bar, tid);

	float4 a = make_float4(0, 0, 0, 0), b = a, c = a;
	int lox = 0, loy = 0, loz = 0, dx = 0, dy = 0, dz = 0;
	const bool valid = tid < n;
	if (valid) {
		a = sTri[tid * 3]; b = sTri[tid * 3 + 1]; c = sTri[tid * 3 + 2];
		// vertex -> cell, footprint min/max (cu:333-371); loops are clipped to < numDiv (cu:374-378) and to >= 0
		int x0 = cell_of(a.x, g.minx, g.maxx, g.nx), x1 = cell_of(b.x, g.minx, g.maxx, g.nx), x2 = cell_of(c.x, g.minx, g.maxx, g.nx);
		int y0 = cell_of(a.y, g.miny, g.maxy, g.ny), y1 = cell_of(b.y, g.miny, g.maxy, g.ny), y2 = cell_of(c.y, g.miny, g.maxy, g.ny);
		int w0 = cell_of(a.z, g.minz, g.maxz, g.nz), w1 = cell_of(b.z, g.minz, g.maxz, g.nz), w2 = cell_of(c.z, g.minz, g.maxz, g.nz);
		lox = max(0, min(x0, min(x1, x2))); int hix = min(g.nx - 1, max(x0, max(x1, x2)));
		loy = max(0, min(y0, min(y1, y2))); int hiy = min(g.ny - 1, max(y0, max(y1, y2)));
		loz = max(0, min(w0, min(w1, w2))); int hiz = min(g.nz - 1, max(w0, max(w1, w2)));
		dx = max(0, hix - lox + 1); dy = max(0, hiy - loy + 1); dz = max(0, hiz - loz + 1);
	}
	const long long ncell = (long long)dx * dy * dz;
	unsigned long long tests = 0, hits = 0;

	// big footprints: all 32 lanes walk the cells of one triangle, p fastest (adjacent atomics)
	unsigned big = __ballot_sync(0xffffffffu, ncell > kBigFootprint);
	while (big) {
		int src = __ffs(big) - 1;
		big &= big - 1;
		float t0x = __shfl_sync(0xffffffffu, a.x, src), t0y = __shfl_sync(0xffffffffu, a.y, src), t0z = __shfl_sync(0xffffffffu, a.z, src);
		float t1x = __shfl_sync(0xffffffffu, b.x, src), t1y = __shfl_sync(0xffffffffu, b.y, src), t1z = __shfl_sync(0xffffffffu, b.z, src);
		float t2x = __shfl_sync(0xffffffffu, c.x, src), t2y = __shfl_sync(0xffffffffu, c.y, src), t2z = __shfl_sync(0xffffffffu, c.z, src);
		int slx = __shfl_sync(0xffffffffu, lox, src), sly = __shfl_sync(0xffffffffu, loy, src), slz = __shfl_sync(0xffffffffu, loz, src);
		int sdx = __shfl_sync(0xffffffffu, dx, src), sdy = __shfl_sync(0xffffffffu, dy, src);
		long long tot = __shfl_sync(0xffffffffu, ncell, src);
		int t = base + (tid - lane) + src;
		for (long long i = lane; i < tot; i += 32) {
			int p = slx + (int)(i % sdx);
			long long rest = i / sdx;
			int q = sly + (int)(rest % sdy), r = slz + (int)(rest / sdy);
			tests++;
			if (tri_box_overlap(cx[p], cy[q], cz[r], g.h1x, g.h1y, g.h1z, t0x, t0y, t0z, t1x, t1y, t1z, t2x, t2y, t2z)) {
				hits++;
				bin_emit<FILL>(g, o, p, q, r, t);
			}
		}
	}
	// small footprints: the owning lane walks its own cells
	if (valid && ncell <= kBigFootprint) {
		int t = base + tid;
		for (int r = loz; r < loz + dz; r++) for (int q = loy; q < loy + dy; q++) for (int p = lox; p < lox + dx; p++) {
			tests++;
			if (tri_box_overlap(cx[p], cy[q], cz[r], g.h1x, g.h1y, g.h1z, a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z)) {
				hits++;
				bin_emit<FILL>(g, o, p, q, r, t);
			}
		}
	}
	if (!FILL) {
		tests = warp_sum(tests); hits = warp_sum(hits);
		if (lane == 0) { atomicAdd(&o.totals->l1Tests, tests); atomicAdd(&o.totals->l1Hits, hits); }
	}
}

// ------------------------------------------------------------------------------------------------ k_cross
// K2a.  For every triangle, the Level-1 columns whose +Z ray passes the det/u/v part of Moller-Trumbore (constant along
// the column, App. A.6).  Candidate columns come from gpv::fill_candidates (certified superset of what the reference's
// brute force Object::ClassifyInOutCPU, src/Object.cpp:716-779, can hit); ill-conditioned triangles test every column.
template <bool FILL>
__global__ void __launch_bounds__(kBinThreads) k_cross(const float4* __restrict__ ray48, int nTri, GridP g,
                                                       const float* __restrict__ cx, const float* __restrict__ cy,
                                                       int* crossCount, const unsigned* __restrict__ crossOff, int* crossTri, Totals* totals)
{
	__shared__ __align__(128) float4 sRay[kBinThreads * 3];
	__shared__ __align__(8) uint64_t bar;
	const int tid = threadIdx.x, lane = tid & 31;
	const int base = blockIdx.x * kBinThreads;
	const int n = min(kBinThreads, nTri - base);
	tma_load_tile(sRay, ray48 + (size_t)base * 3, (uint32_t)n * 48u, &bar, tid);

	RayTri s;
	int i0 = 0, j0 = 0, di = 0, dj = 0, kind = 0;
	if (tid < n) {
		float4 a = sRay[tid * 3], b = sRay[tid * 3 + 1], c = sRay[tid * 3 + 2];
		s.v1x = a.x; s.v1y = a.y; s.v1z = a.z; s.e1x = a.w; s.e1y = b.x; s.e1z = b.y; s.e2x = b.z; s.e2y = b.w;
		s.e2z = c.x; s.det = c.y; s.inv = c.z; s.ok = c.w != 0.f;
		int i1, j1;
		kind = fill_candidates(s, g.minx, g.miny, g.gsx, g.gsy, g.nx, g.ny, i0, i1, j0, j1);
		if (kind == 2) { i0 = 0; j0 = 0; i1 = g.nx - 1; j1 = g.ny - 1; }
		if (kind) { di = i1 - i0 + 1; dj = j1 - j0 + 1; }
	} else { s.ok = false; s.v1x = s.v1y = s.v1z = s.e1x = s.e1y = s.e1z = s.e2x = s.e2y = s.e2z = s.det = s.inv = 0.f; }
	const long long ncol = (long long)di * dj;
	unsigned long long found = 0;

	unsigned big = __ballot_sync(0xffffffffu, ncol > kBigFootprint);
	while (big) {
		int src = __ffs(big) - 1;
		big &= big - 1;
		RayTri w;
		w.v1x = __shfl_sync(0xffffffffu, s.v1x, src); w.v1y = __shfl_sync(0xffffffffu, s.v1y, src);
		w.e1x = __shfl_sync(0xffffffffu, s.e1x, src); w.e1y = __shfl_sync(0xffffffffu, s.e1y, src); w.e1z = __shfl_sync(0xffffffffu, s.e1z, src);
		w.e2x = __shfl_sync(0xffffffffu, s.e2x, src); w.e2y = __shfl_sync(0xffffffffu, s.e2y, src); w.e2z = __shfl_sync(0xffffffffu, s.e2z, src);
		w.inv = __shfl_sync(0xffffffffu, s.inv, src);
		int si0 = __shfl_sync(0xffffffffu, i0, src), sj0 = __shfl_sync(0xffffffffu, j0, src), sdi = __shfl_sync(0xffffffffu, di, src);
		long long tot = __shfl_sync(0xffffffffu, ncol, src);
		int t = base + (tid - lane) + src;
		for (long long k = lane; k < tot; k += 32) {
			int i = si0 + (int)(k % sdi), j = sj0 + (int)(k / sdi);
			RayCol rc;
			if (ray_column(w, cx[i], cy[j], rc)) {
				found++;
				int col = j * g.nx + i;
				if (!FILL) atomicAdd(crossCount + col, 1);
				else crossTri[crossOff[col] + atomicSub(crossCount + col, 1) - 1] = t;
			}
		}
	}
	if (kind && ncol <= kBigFootprint) {
		int t = base + tid;
		for (int j = j0; j < j0 + dj; j++) for (int i = i0; i < i0 + di; i++) {
			RayCol rc;
			if (ray_column(s, cx[i], cy[j], rc)) {
				found++;
				int col = j * g.nx + i;
				if (!FILL) atomicAdd(crossCount + col, 1);
				else crossTri[crossOff[col] + atomicSub(crossCount + col, 1) - 1] = t;
			}
		}
	}
	if (!FILL) {
		found = warp_sum(found);
		unsigned long long ill = warp_sum((unsigned long long)(kind == 2));
		if (lane == 0) { atomicAdd(&totals->crossPairs, found); if (ill) atomicAdd(&totals->nIll, ill); }
	}
}

// ------------------------------------------------------------------------------------------------ k_fill_sweep
// K2b.  One thread = one Level-1 column x one chunk of 32 z-layers; a warp = 32 x-adjacent columns, so every store
// instruction writes 32 consecutive state bytes.  For each crossing triangle of the column only `t` is evaluated per
// cell.  Writes the FINAL Level-1 file byte: 254 boundary (bit set in bmask), 127 inside, 0 outside
// (fill first, SAT overwrites: src/Object.cpp:3158 then :3202; uchar(state*127) :3031).
__global__ void __launch_bounds__(128) k_fill_sweep(const float4* __restrict__ ray48, GridP g, const float* __restrict__ cx, const float* __restrict__ cy,
                                                    const float* __restrict__ cz, const unsigned* __restrict__ crossOff, const int* __restrict__ crossTri,
                                                    const unsigned char* __restrict__ bmask, unsigned char* __restrict__ l1State, Totals* totals)
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
			for (int kk = 0; kk < kn; kk++) m |= (unsigned)ray_cell(s, rc, cz[kbase + kk]) << kk;
			par ^= m;
		}
		const size_t plane = (size_t)g.ny * g.nx;
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
	// MODE_OFFS
	unsigned* off; unsigned* totalOut;
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

template <int MODE>
__global__ void __launch_bounds__(kScanThreads) k_scan(ScanIO io)
{
	__shared__ unsigned sTile;
	__shared__ unsigned long long sWarp[kScanThreads / 32];
	__shared__ unsigned long long sExcl;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) sTile = atomicAdd(io.tileCounter, 1u);
	__syncthreads();
	const unsigned tile = sTile;
	const long long first = (long long)tile * kScanTile + (long long)tid * kScanItems;

	int v[kScanItems];
	if (first + kScanItems <= io.n) {
		const int4* p = reinterpret_cast<const int4*>(io.in + first);
		int4 a = p[0], b = p[1];
		v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
	} else {
#pragma unroll
		for (int k = 0; k < kScanItems; k++) v[k] = (first + k < io.n) ? io.in[first + k] : 0;
	}
	unsigned long long item[kScanItems], threadSum = 0;
#pragma unroll
	for (int k = 0; k < kScanItems; k++) {
		item[k] = (MODE == MODE_CELLS) ? (((unsigned long long)(v[k] > 0) << 31) | (unsigned)v[k]) : (unsigned long long)(unsigned)v[k];
		threadSum += item[k];
	}
	unsigned long long incl = threadSum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		unsigned long long y = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += y;
	}
	if (lane == 31) sWarp[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		unsigned long long w = lane < kScanThreads / 32 ? sWarp[lane] : 0, wi = w;
#pragma unroll
		for (int o = 1; o < 8; o <<= 1) {
			unsigned long long y = __shfl_up_sync(0xffffffffu, wi, o);
			if (lane >= o) wi += y;
		}
		const unsigned long long blockAgg = __shfl_sync(0xffffffffu, wi, kScanThreads / 32 - 1);
		if (lane < kScanThreads / 32) sWarp[lane] = wi - w;
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
	unsigned long long run = sExcl + sWarp[warp] + (incl - threadSum);

	if (MODE == MODE_CELLS) {
		int pre[kScanItems];
		unsigned flags = 0;
#pragma unroll
		for (int k = 0; k < kScanItems; k++) {
			unsigned b = (unsigned)(run >> 31), ts = (unsigned)(run & 0x7fffffffu);
			pre[k] = (int)b;
			if (v[k] > 0) {
				flags |= 1u << k;
				io.boundaryIndex[b] = (int)(io.globalBase + first + k);
				io.bTriOff[b] = ts;
			}
			run += item[k];
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
			run += item[k];
		}
		if (first <= io.n - 1 && io.n - 1 < first + kScanItems) { io.off[io.n] = (unsigned)run; *io.totalOut = (unsigned)run; }
	}
}

// ------------------------------------------------------------------------------------------------ k_sort_segments
// One warp per segment [off[s], off[s]+len).  Canonical order = ascending triangle id (what the reference's CPU path
// yields, src/Object.cpp:2705-2748; the GPU path's atomic slots are nondeterministic).  len <= 32: rank sort in
// registers; longer: in-place bitonic network in the all-ascending "flip" form, valid for any length.
// UNIQUE (column lists): drops duplicates (a triangle hits several cells of one column) and stores the unique count.
template <bool UNIQUE>
__global__ void __launch_bounds__(256) k_sort_segments(const unsigned* __restrict__ off, int nSeg, int* data, int* uniqueCount)
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
		for (int j = 0; j < 32; j++) {
			int y = __shfl_sync(0xffffffffu, x, j);
			rank += (y < x) || (y == x && j < lane);
		}
		__syncwarp();
		if (!UNIQUE) { if (lane < len) a[rank] = x; return; }
		// sorted value of position `lane` = the x whose rank == lane: exchange through shuffles
		int sorted = 0x7fffffff;
		for (int j = 0; j < 32; j++) {
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
	for (int k = 2; (k >> 1) < len; k <<= 1) {
		for (int i = lane; i < len; i += 32) { // flip: mirror inside blocks of k
			int l = i ^ (k - 1);
			if (l > i && l < len) { int x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
		}
		__syncwarp();
		for (int j = k >> 2; j > 0; j >>= 1) { // disperse
			for (int i = lane; i < len; i += 32) {
				int l = i ^ j;
				if (l > i && l < len) { int x = a[i], y = a[l]; if (x > y) { a[i] = y; a[l] = x; } }
			}
			__syncwarp();
		}
	}
	if (UNIQUE) { // in-place compaction, 32 elements at a time, reads finish before writes of the same round
		int w = 0, carry = 0;
		for (int b0 = 0; b0 < len; b0 += 32) {
			int i = b0 + lane;
			int x = i < len ? a[i] : 0;
			int prev = __shfl_up_sync(0xffffffffu, x, 1);
			if (lane == 0) prev = carry;
			bool keep = i < len && (i == 0 || x != prev);
			unsigned km = __ballot_sync(0xffffffffu, keep);
			carry = __shfl_sync(0xffffffffu, x, 31);
			__syncwarp();
			if (keep) a[w + __popc(km & ((1u << lane) - 1))] = x;
			w += __popc(km);
			__syncwarp();
		}
		if (lane == 0) uniqueCount[seg] = w;
	}
}

// ------------------------------------------------------------------------------------------------ k_l2
// K4.  Replaces CUDAClassifyInOutLevel2Kernel (cu:450-504) + CUDAClassifyTessellationLevel2Kernel (cu:403-448).
// A CTA of 256 threads refines G = max(1, 256/n2^2) boundary cells; a thread owns (cell, q, r): first as an xy-column
// (p=q', q=r') of the parity-ray phase, then as a row of n2 sub-voxels along x in the SAT phase, so the y/z part of every
// SAT test is hoisted out of the x loop (gpv::SatRow) and each row leaves as ONE vector store of final file bytes
// (n2 = 16: 128-bit; rows of a CTA are contiguous in Level2InOut.raw).
// Sub-voxel centre (cu:423-425 / 472-474): ((2p+1)*ext2 + mid) - ext1, all f32.
struct L2IO {
	const float4* tri48; const float4* ray48;
	const int* boundaryIndex; const unsigned* bTriOff; const int* cellTris;
	const unsigned* colOff; const int* colCount; const int* colTris;
	const float* cx; const float* cy; const float* cz;
	unsigned char* l2State; // nBoundary * n2^3 file bytes
	int nBoundary;
	Totals* totals;
};

constexpr int kL2Threads = 256;

__global__ void __launch_bounds__(kL2Threads) k_l2(GridP g, L2IO io)
{
	extern __shared__ unsigned char smemRaw[];
	const int n2 = g.n2, rows = n2 * n2;
	const int G = max(1, kL2Threads / rows);
	float* sC = reinterpret_cast<float*>(smemRaw);                   // [G][3][n2] sub-voxel centres
	unsigned* sPar = reinterpret_cast<unsigned*>(sC + G * 3 * n2);    // [G][rows] parity bits along z per xy-column
	int* sInfo = reinterpret_cast<int*>(sPar + G * rows);             // [G][4] triOff, triCnt, colOff, colCnt
	const int tid = threadIdx.x;
	const long long b0 = (long long)blockIdx.x * G;

	for (int k = tid; k < G * 3 * n2; k += kL2Threads) {
		int gi = k / (3 * n2), rem = k - gi * 3 * n2, ax = rem / n2, p = rem - ax * n2;
		long long b = b0 + gi;
		float val = 0.f;
		if (b < io.nBoundary) {
			int l1 = io.boundaryIndex[b];
			int kz = l1 / (g.nx * g.ny), ij = l1 - kz * g.nx * g.ny, jy = ij / g.nx, ix = ij - jy * g.nx;
			float mid = ax == 0 ? io.cx[ix] : (ax == 1 ? io.cy[jy] : io.cz[kz]);
			float e2 = ax == 0 ? g.h2x : (ax == 1 ? g.h2y : g.h2z), e1 = ax == 0 ? g.h1x : (ax == 1 ? g.h1y : g.h1z);
			val = (float)(2 * p + 1) * e2 + mid - e1;
		}
		sC[k] = val;
	}
	for (int gi = tid; gi < G; gi += kL2Threads) {
		long long b = b0 + gi;
		int* inf = sInfo + gi * 4;
		if (b < io.nBoundary) {
			int l1 = io.boundaryIndex[b];
			int col = l1 % (g.nx * g.ny);
			inf[0] = (int)io.bTriOff[b]; inf[1] = (int)(io.bTriOff[b + 1] - io.bTriOff[b]);
			inf[2] = (int)io.colOff[col]; inf[3] = io.colCount[col];
		} else { inf[0] = inf[1] = inf[2] = inf[3] = 0; }
	}
	__syncthreads();

	// ---- phase 1: parity rays.  item = (cell gi, xy-column pq); n2 <= 32 so the z parity fits one word
	for (int item = tid; item < G * rows; item += kL2Threads) {
		int gi = item / rows, pq = item - gi * rows, q = pq / n2, p = pq - q * n2;
		const float* c = sC + gi * 3 * n2;
		const float ox = c[p], oy = c[n2 + q];
		const int* inf = sInfo + gi * 4;
		unsigned par = 0;
		for (int k = 0; k < inf[3]; k++) {
			RayTri s;
			load_ray(s, io.ray48, io.colTris[inf[2] + k]);
			RayCol rc;
			if (!s.ok || !ray_column(s, ox, oy, rc)) continue;
			for (int r = 0; r < n2; r++) par ^= (unsigned)ray_cell(s, rc, c[2 * n2 + r]) << r;
		}
		sPar[item] = par;
	}
	__syncthreads();

	// ---- phase 2: SAT per row of n2 sub-voxels, then the row's file bytes
	unsigned long long nIn = 0, nBd = 0;
	for (int item = tid; item < G * rows; item += kL2Threads) {
		int gi = item / rows, row = item - gi * rows, r = row / n2, q = row - r * n2;
		long long b = b0 + gi;
		if (b >= io.nBoundary) continue;
		const float* c = sC + gi * 3 * n2;
		const float cy2 = c[n2 + q], cz2 = c[2 * n2 + r];
		const int* inf = sInfo + gi * 4;
		unsigned sat = 0;
		const unsigned full = n2 == 32 ? 0xffffffffu : ((1u << n2) - 1);
		for (int k = 0; k < inf[1] && sat != full; k++) {
			int t = io.cellTris[inf[0] + k];
			float4 A = __ldg(io.tri48 + (size_t)t * 3), B = __ldg(io.tri48 + (size_t)t * 3 + 1), C = __ldg(io.tri48 + (size_t)t * 3 + 2);
			SatRow s;
			if (!sat_row_setup(s, cy2, cz2, g.h2y, g.h2z, A.y, A.z, B.y, B.z, C.y, C.z)) continue;
			for (int p = 0; p < n2; p++) {
				if ((sat >> p) & 1) continue;
				if (sat_row_test(s, c[p], g.h2x, g.h2y, g.h2z, A.x, B.x, C.x)) sat |= 1u << p;
			}
		}
		unsigned par = 0;
		const unsigned* pr = sPar + gi * rows + q * n2;
		for (int p = 0; p < n2; p++) par |= ((pr[p] >> r) & 1u) << p;
		par &= ~sat;
		nIn += __popc(par); nBd += __popc(sat);
		unsigned char* out = io.l2State + ((size_t)b * rows + row) * n2;
		if (n2 == 16) {
			unsigned w[4];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				unsigned x = 0;
#pragma unroll
				for (int e = 0; e < 4; e++) {
					int p = k * 4 + e;
					x |= (((sat >> p) & 1) ? 254u : (((par >> p) & 1) ? 127u : 0u)) << (8 * e);
				}
				w[k] = x;
			}
			*reinterpret_cast<uint4*>(out) = make_uint4(w[0], w[1], w[2], w[3]);
		} else if (n2 == 8) {
			unsigned w[2];
#pragma unroll
			for (int k = 0; k < 2; k++) {
				unsigned x = 0;
#pragma unroll
				for (int e = 0; e < 4; e++) {
					int p = k * 4 + e;
					x |= (((sat >> p) & 1) ? 254u : (((par >> p) & 1) ? 127u : 0u)) << (8 * e);
				}
				w[k] = x;
			}
			*reinterpret_cast<uint2*>(out) = make_uint2(w[0], w[1]);
		} else if (n2 == 4) {
			unsigned x = 0;
#pragma unroll
			for (int e = 0; e < 4; e++) x |= (((sat >> e) & 1) ? 254u : (((par >> e) & 1) ? 127u : 0u)) << (8 * e);
			*reinterpret_cast<unsigned*>(out) = x;
		} else {
			for (int p = 0; p < n2; p++) out[p] = ((sat >> p) & 1) ? 254 : (((par >> p) & 1) ? 127 : 0);
		}
	}
	nIn = warp_sum(nIn); nBd = warp_sum(nBd);
	if ((tid & 31) == 0 && (nIn | nBd)) { atomicAdd(&io.totals->l2Inside, nIn); atomicAdd(&io.totals->l2Boundary, nBd); }
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

// K5b. Level-2 normals: only boundary sub-voxels (byte 254) re-run the SAT over the cell list, accumulating the
// UN-normalised cross(e01,e02) of every hit in ascending order (cu:311-318: normalize() result discarded; cu:40-46),
// then the host averaging of src/Object.cpp:2613-2632.  One thread per sub-voxel; others write 127,127,127.
__global__ void __launch_bounds__(256) k_l2_normals(GridP g, L2IO io, unsigned char* __restrict__ l2Normal)
{
	const int n2 = g.n2;
	const long long n23 = (long long)n2 * n2 * n2;
	long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= (long long)io.nBoundary * n23) return;
	unsigned char n0 = 127, n1 = 127, n2b = 127;
	if (io.l2State[v] == 254) {
		long long b = v / n23;
		int loc = (int)(v - b * n23), r = loc / (n2 * n2), pq = loc - r * n2 * n2, q = pq / n2, p = pq - q * n2;
		int l1 = io.boundaryIndex[b];
		int kz = l1 / (g.nx * g.ny), ij = l1 - kz * g.nx * g.ny, jy = ij / g.nx, ix = ij - jy * g.nx;
		float cxv = (float)(2 * p + 1) * g.h2x + io.cx[ix] - g.h1x;
		float cyv = (float)(2 * q + 1) * g.h2y + io.cy[jy] - g.h1y;
		float czv = (float)(2 * r + 1) * g.h2z + io.cz[kz] - g.h1z;
		float sx = 0, sy = 0, sz = 0, cnt = 0;
		for (unsigned k = io.bTriOff[b]; k < io.bTriOff[b + 1]; k++) {
			int t = io.cellTris[k];
			float4 A = __ldg(io.tri48 + (size_t)t * 3), B = __ldg(io.tri48 + (size_t)t * 3 + 1), C = __ldg(io.tri48 + (size_t)t * 3 + 2);
			if (tri_box_overlap(cxv, cyv, czv, g.h2x, g.h2y, g.h2z, A.x, A.y, A.z, B.x, B.y, B.z, C.x, C.y, C.z)) {
				float ax = B.x - A.x, ay = B.y - A.y, az = B.z - A.z, bx = C.x - A.x, by = C.y - A.y, bz = C.z - A.z;
				sx += ay * bz - az * by; sy += az * bx - ax * bz; sz += ax * by - ay * bx; // cutil_math.h:409 cross()
				cnt += 1;
			}
		}
		if (cnt > 0) {
			float ax = sx / cnt, ay = sy / cnt, az = sz / cnt;
			normalize3(ax, ay, az);
			n0 = encode_normal(ax); n1 = encode_normal(ay); n2b = encode_normal(az);
		}
	}
	l2Normal[v * 3] = n0; l2Normal[v * 3 + 1] = n1; l2Normal[v * 3 + 2] = n2b;
}

} // namespace gpv
