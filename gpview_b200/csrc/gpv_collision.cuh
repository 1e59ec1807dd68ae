// gpview_b200/csrc/gpv_collision.cuh -- voxel hierarchy / collision structures over the Level-1 grid, on the device:
//   k_occupied_count / k_occupied_write   Object::CollisionInitCUDA (src/Object.cpp:3530-3572): the occupied cells (state >= 1) as an
//                                         ascending inverse index + box centre / extent arrays (the reference builds them on the host
//                                         and uploads them)
//   k_hier_leaves, k_hier_level           Object::BuildHierarchy (:2790-2867) with CombineBBox (:2750-2788): binary AABB hierarchy --
//                                         pairs of x-neighbours first, then halving x / y / z in rotation, one launch per level
// Arithmetic as in the reference (f32 min / max / +, - and an exact halving); oracle: oracle/gpv_oracle_collision.c, pinned against
// the reference's own BuildHierarchy by tests/test_oracle_ref.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gpv {

constexpr int kOccBlock = 1024; // cells per counting block

// occupied cells per block of kOccBlock cells (state byte != 0: inside 127 or boundary 254)
__global__ void __launch_bounds__(256) k_occupied_count(const unsigned char* __restrict__ state, long long cells, int* __restrict__ blockCount)
{
	const long long base = (long long)blockIdx.x * kOccBlock;
	int n = 0;
	for (int i = threadIdx.x; i < kOccBlock; i += 256) n += (base + i < cells) && state[base + i] != 0;
	n = __reduce_add_sync(0xffffffffu, n);
	__shared__ int s[8];
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = n;
	__syncthreads();
	if (threadIdx.x == 0) blockCount[blockIdx.x] = s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
}

// ascending inverse index + centre / extent of every occupied cell; blockOff = exclusive scan of blockCount
__global__ void __launch_bounds__(256) k_occupied_write(const unsigned char* __restrict__ state, long long cells, const unsigned* __restrict__ blockOff, int nx, int ny,
                                                         const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz, float h1x, float h1y, float h1z,
                                                         int* __restrict__ invIndex, float* __restrict__ center, float* __restrict__ extent)
{
	__shared__ int sWarp[8];
	__shared__ int sBase;
	const long long base = (long long)blockIdx.x * kOccBlock;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) sBase = (int)blockOff[blockIdx.x];
	__syncthreads();
	for (int i0 = 0; i0 < kOccBlock; i0 += 256) { // four rounds of 256 consecutive cells: ranks stay ascending
		const long long idx = base + i0 + threadIdx.x;
		const bool occ = idx < cells && state[idx] != 0;
		const unsigned m = __ballot_sync(0xffffffffu, occ);
		if (lane == 0) sWarp[warp] = __popc(m);
		__syncthreads();
		int before = 0, total = 0;
		for (int w = 0; w < 8; w++) { before += w < warp ? sWarp[w] : 0; total += sWarp[w]; }
		if (occ) {
			const long long k = (long long)sBase + before + __popc(m & ((1u << lane) - 1u));
			const int plane = nx * ny, kz = (int)(idx / plane), rem = (int)(idx - (long long)kz * plane), jy = rem / nx, ix = rem - jy * nx;
			invIndex[k] = (int)idx;
			center[k * 3] = cx[ix]; center[k * 3 + 1] = cy[jy]; center[k * 3 + 2] = cz[kz];
			extent[k * 3] = h1x; extent[k * 3 + 1] = h1y; extent[k * 3 + 2] = h1z;
		}
		__syncthreads();
		if (threadIdx.x == 0) sBase += total;
		__syncthreads();
	}
}

struct HierBox { float mx, my, mz, hx, hy, hz; int solid; int index; };

// CombineBBox (src/Object.cpp:2750-2788)
__device__ __forceinline__ void hier_combine(const HierBox& a, const HierBox& b, HierBox& o)
{
	o.mx = o.my = o.mz = o.hx = o.hy = o.hz = 0.f;
	if (a.solid == 0 && b.solid == 0) { o.solid = 0; return; }
	o.solid = 1;
	if (a.solid == 0) { o.mx = b.mx; o.my = b.my; o.mz = b.mz; o.hx = b.hx; o.hy = b.hy; o.hz = b.hz; return; }
	if (b.solid == 0) { o.mx = a.mx; o.my = a.my; o.mz = a.mz; o.hx = a.hx; o.hy = a.hy; o.hz = a.hz; return; }
	const float xlo = fminf(a.mx - a.hx, b.mx - b.hx), xhi = fmaxf(a.mx + a.hx, b.mx + b.hx);
	const float ylo = fminf(a.my - a.hy, b.my - b.hy), yhi = fmaxf(a.my + a.hy, b.my + b.hy);
	const float zlo = fminf(a.mz - a.hz, b.mz - b.hz), zhi = fmaxf(a.mz + a.hz, b.mz + b.hz);
	o.mx = (xhi + xlo) / 2.0f; o.my = (yhi + ylo) / 2.0f; o.mz = (zhi + zlo) / 2.0f;
	o.hx = (xhi - xlo) / 2.0f; o.hy = (yhi - ylo) / 2.0f; o.hz = (zhi - zlo) / 2.0f;
}

struct HierOut { float* mid; float* half; unsigned char* solid; int* child; };
__device__ __forceinline__ void hier_store(const HierOut& h, int at, const HierBox& o, int c1, int c2)
{
	h.mid[(size_t)at * 3] = o.mx; h.mid[(size_t)at * 3 + 1] = o.my; h.mid[(size_t)at * 3 + 2] = o.mz;
	h.half[(size_t)at * 3] = o.hx; h.half[(size_t)at * 3 + 1] = o.hy; h.half[(size_t)at * 3 + 2] = o.hz;
	h.solid[at] = (unsigned char)o.solid; h.child[(size_t)at * 2] = c1; h.child[(size_t)at * 2 + 1] = c2;
}
__device__ __forceinline__ HierBox hier_load(const HierOut& h, int at)
{
	HierBox b;
	b.mx = h.mid[(size_t)at * 3]; b.my = h.mid[(size_t)at * 3 + 1]; b.mz = h.mid[(size_t)at * 3 + 2];
	b.hx = h.half[(size_t)at * 3]; b.hy = h.half[(size_t)at * 3 + 1]; b.hz = h.half[(size_t)at * 3 + 2];
	b.solid = h.solid[at]; b.index = at;
	return b;
}

// level 1 (:2806-2814): box i = leaves 2i, 2i+1 (x-neighbours; nx is even).  A leaf's solid flag is the parity fill of the cell -- also
// for boundary cells (bBox[].solid is set before the SAT pass, :3165-3193) --, bit (kz & 31) of solidWords[(kz >> 5) * plane + col].
__global__ void __launch_bounds__(256) k_hier_leaves(int total, int nx, int ny, const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ cz,
                                                      float h1x, float h1y, float h1z, const unsigned* __restrict__ solidWords, HierOut h)
{
	const int i = blockIdx.x * 256 + threadIdx.x;
	if (i >= total / 2) return;
	const int plane = nx * ny;
	HierBox leaf[2];
	for (int s = 0; s < 2; s++) {
		const int idx = 2 * i + s, kz = idx / plane, col = idx - kz * plane, jy = col / nx, ix = col - jy * nx;
		leaf[s].mx = cx[ix]; leaf[s].my = cy[jy]; leaf[s].mz = cz[kz];
		leaf[s].hx = h1x; leaf[s].hy = h1y; leaf[s].hz = h1z;
		leaf[s].solid = (int)((solidWords[(size_t)(kz >> 5) * plane + col] >> (kz & 31)) & 1u);
		leaf[s].index = idx;
	}
	HierBox o;
	hier_combine(leaf[0], leaf[1], o);
	hier_store(h, i, o, leaf[0].index, leaf[1].index);
}

// one further level (:2822-2866): the boxes of the previous level form a dX x dY x dZ grid at prevLevelIndex; the axis whose skip is 2 is halved
__global__ void __launch_bounds__(256) k_hier_level(int dX, int dY, int dZ, int iSkip, int jSkip, int kSkip, int prevLevelIndex, int levelIndex, HierOut h)
{
	const int oX = dX / iSkip, oY = dY / jSkip, oZ = dZ / kSkip;
	const int index2 = blockIdx.x * 256 + threadIdx.x;
	if (index2 >= oX * oY * oZ) return;
	const int k2 = index2 / (oY * oX), r = index2 - k2 * oY * oX, j2 = r / oX, i2 = r - j2 * oX;
	const int i = i2 * iSkip, j = j2 * jSkip, k = k2 * kSkip;
	const int index1 = k * dY * dX + j * dX + i;
	const int skip = (kSkip - 1) * dY * dX + (jSkip - 1) * dX + (iSkip - 1);
	const HierBox a = hier_load(h, prevLevelIndex + index1), b = hier_load(h, prevLevelIndex + index1 + skip);
	HierBox o;
	hier_combine(a, b, o);
	hier_store(h, levelIndex + index2, o, a.index, b.index);
}

} // namespace gpv
