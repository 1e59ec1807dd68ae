// oracle/ref_kernels_host.cpp -- TEST INFRASTRUCTURE ONLY: the reference's CUDA kernels EXECUTED ON THE HOST.
//
// cuda/CUDAClassifyTessellation.cu holds the three kernels of the reference's operator boundary.  They are plain per-thread
// code (no shared memory, no barriers; one atomicAdd), so the unmodified kernel source compiles as C++ once the CUDA built-ins
// exist: oracle/Makefile pipes this file to the compiler with the reference file's text up to its <<< >>> launch wrappers
// spliced in at the marker below (read from $(GPVIEW_REF) where it lies; no copy is written anywhere).  This file supplies
// threadIdx / blockIdx / blockDim / atomicAdd in front of that text and, behind it, re-creates the three extern "C" operators
// as loops over the thread indices the wrappers would launch (cu:507-540).  A dozen CUDA runtime calls are answered with host
// memory, so that the reference's UNMODIFIED host code (Object::ClassifyTessellationCUDA,
// Object::ClassifyInOutTessellationLevel2CUDA) runs against them without a GPU:
// oracle/_ref/libgpvref_emu.so = the reference's GPU path, source for source, under g++ -O2 -ffp-contract=off (the strict-IEEE
// twin of nvcc -fmad=false: + - * / and sqrt round identically).  tests/test_oracle_ref.py compares it with the oracle.
//
// Threads run in ascending index order, so atomic slots come out ascending -- one of the orders the GPU may produce.
#include <cuda_runtime.h> // float3 / int3 / uint3 / dim3 / make_float3 for the host compiler
#include <cstdlib>
#include <cstring>

static uint3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
static inline int atomicAdd(int* p, int v) { const int old = *p; *p += v; return old; }
static inline float atomicAdd(float* p, float v) { const float old = *p; *p += v; return old; }

// @@REFERENCE_KERNEL_SOURCE@@  oracle/Makefile splices the reference's kernel source in here, on the fly (see the top of this file)

static inline void at(unsigned x, unsigned y)
{
	blockDim = dim3(1, 1, 1); gridDim = dim3(1, 1, 1);
	threadIdx.x = threadIdx.y = threadIdx.z = 0;
	blockIdx.x = x; blockIdx.y = y; blockIdx.z = 0;
}

// cu:507-515
extern "C" int CUDAClassifyTessellation(float* tris, int numTriangles, float* inOut, int* triCount, int* triIndex, float3 objBoxMin, float3 objBoxMax,
                                        float3 boxExtents, int3 numDiv, int triBufferLen)
{
	for (int i = 0; i < numTriangles; i++) { at((unsigned)i, 0); CUDAClassifyTessellationKernel(tris, inOut, triCount, triIndex, numTriangles, objBoxMin, objBoxMax, boxExtents, numDiv, triBufferLen); }
	return 1;
}
// cu:518-528
extern "C" int CUDAClassifyTessellationLevel2(float* tris, float* l2InOut, float* l2Normal, float* l1Mid, int* l2Index, int* triCount, int* triFlatIndex, int* triFlat,
                                              int numBoundary, int3 numDiv2, float3 ext1, float3 ext2)
{
	const int n = numDiv2.x * numDiv2.y * numDiv2.z;
	for (int b = 0; b < numBoundary; b++) for (int k = 0; k < n; k++) {
		at((unsigned)b, (unsigned)k);
		CUDAClassifyTessellationLevel2Kernel(tris, l2InOut, l2Normal, l1Mid, l2Index, triCount, triFlatIndex, triFlat, numBoundary, numDiv2, ext1, ext2);
	}
	return 1;
}
// cu:531-540
extern "C" int CUDAClassifyInOutLevel2(float* tris, float* l2InOut, float* l1Mid, int* l2Index, int* xyCount, int* xyFlatIndex, int* xyFlat, int numBoundary, int3 numDiv,
                                       int3 numDiv2, float3 ext1, float3 ext2)
{
	const int n = numDiv2.x * numDiv2.y * numDiv2.z;
	for (int b = 0; b < numBoundary; b++) for (int k = 0; k < n; k++) {
		at((unsigned)b, (unsigned)k);
		CUDAClassifyInOutLevel2Kernel(tris, l2InOut, l1Mid, l2Index, xyCount, xyFlatIndex, xyFlat, numBoundary, numDiv, numDiv2, ext1, ext2);
	}
	return 1;
}
extern "C" float THRUSTDeviceFindMax(float*, int, int) { abort(); } // cuda/THRUSTUtilities.cu: not on the voxelizer path

// ---- the CUDA runtime calls of the reference's host code, answered with host memory.  "Device" memory is zero-filled like a
// fresh cudaMalloc in practice (the reference's own cudaMemset calls cover a quarter of its buffers) and
// carries 1 MB of slack: the first Level-1 pass writes past the slot of any cell with more triangles than the buffer (App. B4).
extern "C" {
cudaError_t cudaMalloc(void** p, size_t n) { *p = calloc(n + (1u << 20), 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
cudaError_t cudaThreadSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) { memset(p, 0, sizeof *p); strcpy(p->name, "host emulation"); p->major = 10; return cudaSuccess; }
}
