// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Drives the UNMODIFIED reference sources (compiled where they lie under $GPVIEW_REF, default
// /root/reference, against the generated no-op GL shim) through the reference's own CPU voxelizer path
// and exposes the results over a small C ABI for ctypes (tests/, oracle/gen_golden.py, bench.py's
// cpu_baseline leg).  Output: oracle/_ref/libgpvref.so (git-ignored).
//
// What is the reference's code and what is restated here:
//   * reference, called as is: Object::ReadObject / ReadOFFObject / CreateFlatTriangleData,
//     Object::ClassifyInOutCPU (src/Object.cpp:716), Object::ClassifyTessellation (:2256),
//     Object::ClassifyInOutLevel2CPU (:1111), Object::ClassifyTessellationLevel2 (:2344),
//     Object::SaveVoxelization (:2934), TriBoxOverlap (src/TriBoxIntersection.cpp:135),
//     triangle_ray_intersection (src/TriRayIntersection.cpp:78).
//   * restated (Object::PerformVoxelization cannot run headless: its L1 fill is GL, :3158): the grid set-up
//     (:3094-3154), the bBox[] initialisation (:3165-3193), L1 normals (:3219-3253), boundary prefix sum
//     (:3258-3284), voxel counting (:3353-3378); the host CSR/column-list construction of
//     ClassifyTessellationCUDA (:2137-2180); and the arithmetic of the two Level-2 CUDA kernels
//     (cuda/CUDAClassifyTessellation.cu:403-504) evaluated on the CPU under strict IEEE with the
//     reference's own TriBoxOverlap / triangle_ray_intersection object code ("kernel form").
#include "Object.h"
#include <thread>
#include <atomic>
#include <unistd.h>

// The reference's Object.o references these; the CPU path never calls them.  With -DGPV_LINK_PRODUCT the harness is linked
// against libgpview_b200.so instead (the drop-in boundary test, tests/test_gpu_compat.py) and the stubs disappear.
#ifndef GPV_LINK_PRODUCT
extern "C" int CUDAClassifyTessellation(float*, int, float*, int*, int*, float3, float3, float3, int3, int) { abort(); }
extern "C" int CUDAClassifyTessellationLevel2(float*, float*, float*, float*, int*, int*, int*, int*, int, int3, float3, float3) { abort(); }
extern "C" int CUDAClassifyInOutLevel2(float*, float*, float*, int*, int*, int*, int*, int, int3, int3, float3, float3) { abort(); }
extern "C" float THRUSTDeviceFindMax(float*, int, int) { abort(); }
#endif

int TriBoxOverlap(float boxcenter[3], float boxhalfsize[3], float triverts[3][3]);
int triangle_ray_intersection(const float V1[3], const float V2[3], const float V3[3], const float O[3], const float D[3], float* out);

namespace {
struct Ref {
	Object* o = nullptr;
	GLParameters* gp = nullptr;
	bool setup = false, boxes = false, compacted = false;
	// column lists (ClassifyTessellationCUDA host code, canonical ascending order inside a cell)
	std::vector<int> triFlat, triFlatIndex, triCount, xyFlat, xyFlatIndex, xyCount;
	std::vector<float> l2k;       // kernel-form Level-2 state (float 0/1/2)
	std::vector<float> l2kNormal; // kernel-form Level-2 normals, 4 floats per voxel (averaged like :2613-2632)
	long l1BoxTests = 0, l1BoxHits = 0, l2BoxTests = 0, l2RayTests = 0, l1ColRayTests = 0;
	int maxPerCell = 0;
};
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}

extern "C" {

void* ref_open(const char* path, int isOff, int objID)
{
	Ref* r = new Ref();
	r->gp = new GLParameters();
	r->gp->saveVoxels = false;
	r->o = new Object();
	Face* f = new Face();
	f->trimmed = false;
	f->surfID = 0;
	r->o->faces.push_back(f);
	r->o->objID = objID;
	if (isOff) r->o->ReadOFFObject((char*)path);
	else r->o->ReadObject((char*)path);
	r->o->CreateFlatTriangleData();
	return r;
}

int ref_ntri(void* h) { return ((Ref*)h)->o->totalNumTriangles; }
const float* ref_tris(void* h) { return ((Ref*)h)->o->flatCPUTriangleData; }
void ref_bbox(void* h, float* bmin, float* bmax, float* maxModelSize)
{
	Object* o = ((Ref*)h)->o;
	for (int a = 0; a < 3; a++) { bmin[a] = o->bBoxMin[a]; bmax[a] = o->bBoxMax[a]; }
	*maxModelSize = o->maxModelSize;
}

// Grid set-up, restated from Object::PerformVoxelization (src/Object.cpp:3084-3154).
void ref_setup(void* h, int voxelCount, int voxelCount2)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	GLParameters* gp = r->gp;
	gp->voxelCount = voxelCount;
	gp->voxelCount2 = voxelCount2;
	gp->level2Voxels = voxelCount2 > 0;
	if (o->voxelInit) delete o->voxelData;
	o->voxelData = new VoxelData();
	o->voxelInit = true;
	VoxelData* vd = o->voxelData;
	vd->level1VoxelStripDLID = 0;
	vd->storeBoxData = true;
	float nominalGridSize = o->maxModelSize / (1.0*(gp->voxelCount));
	vd->numDivX = (int((o->bBoxMax[0] - o->bBoxMin[0]) / nominalGridSize));
	vd->numDivY = (int((o->bBoxMax[1] - o->bBoxMin[1]) / nominalGridSize));
	vd->numDivZ = (int((o->bBoxMax[2] - o->bBoxMin[2]) / nominalGridSize));
	if (vd->numDivX == 0) vd->numDivX++;
	if (vd->numDivY == 0) vd->numDivY++;
	if (vd->numDivZ == 0) vd->numDivZ++;
	vd->numDivX = GetNextDiv4(vd->numDivX);
	vd->numDivY = GetNextDiv4(vd->numDivY);
	vd->numDivZ = GetNextDiv4(vd->numDivZ);
	float gx = (o->bBoxMax[0] - o->bBoxMin[0]) / (vd->numDivX*1.0);
	float gy = (o->bBoxMax[1] - o->bBoxMin[1]) / (vd->numDivY*1.0);
	float gz = (o->bBoxMax[2] - o->bBoxMin[2]) / (vd->numDivZ*1.0);
	vd->gridSizeX = gx; vd->gridSizeY = gy; vd->gridSizeZ = gz;
	int n2 = voxelCount2 > 0 ? voxelCount2 : 1;
	vd->numDivX2 = vd->numDivY2 = vd->numDivZ2 = n2;
	vd->gridSizeX2 = gx / (n2*1.0);
	vd->gridSizeY2 = gy / (n2*1.0);
	vd->gridSizeZ2 = gz / (n2*1.0);
	size_t N = (size_t)vd->numDivX * vd->numDivY * vd->numDivZ;
	vd->level1InOut = new inOutDType[N];
	memset(vd->level1InOut, 0, N * sizeof(inOutDType));
	vd->level1Normal = new float[N * 3];
	memset(vd->level1Normal, 0, N * 3 * sizeof(float));
	vd->boundaryPrefixSum = new int[N];
	vd->numLevel1InsideVoxels = vd->numLevel1BoundaryVoxels = 0;
	vd->numLevel2InsideVoxels = vd->numLevel2BoundaryVoxels = 0;
	r->setup = true; r->boxes = false; r->compacted = false;
}

void ref_grid(void* h, int* numDiv, float* gridSize, float* gridSize2)
{
	VoxelData* vd = ((Ref*)h)->o->voxelData;
	numDiv[0] = vd->numDivX; numDiv[1] = vd->numDivY; numDiv[2] = vd->numDivZ;
	gridSize[0] = vd->gridSizeX; gridSize[1] = vd->gridSizeY; gridSize[2] = vd->gridSizeZ;
	gridSize2[0] = vd->gridSizeX2; gridSize2[1] = vd->gridSizeY2; gridSize2[2] = vd->gridSizeZ2;
}

// L1 solid fill, the reference's own brute force: Object::ClassifyInOutCPU (src/Object.cpp:716-779).
double ref_l1_inout_brute(void* h)
{
	Ref* r = (Ref*)h;
	double t0 = now();
	r->o->ClassifyInOutCPU(r->gp);
	return now() - t0;
}

// Same loop nest (src/Object.cpp:733-771) around the reference's triangle_ray_intersection object code,
// z-layers split over host threads -- only so that the 256^3 golden finishes in minutes.  Proven equal to
// ref_l1_inout_brute on the 64^3 case by tests/test_oracle_ref.py.
double ref_l1_inout_brute_mt(void* h, int nThreads)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	float gx = vd->gridSizeX, gy = vd->gridSizeY, gz = vd->gridSizeZ;
	int nTri = o->totalNumTriangles;
	const float* T = o->flatCPUTriangleData;
	double t0 = now();
	std::atomic<int> next(0);
	auto work = [&]() {
		float rayDir[3] = { 0, 0, 1 };
		for (;;) {
			int k = next.fetch_add(1);
			if (k >= nz) break;
			for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
				float c[3];
				c[0] = o->bBoxMin[0] + (i + 0.5)*gx;
				c[1] = o->bBoxMin[1] + (j + 0.5)*gy;
				c[2] = o->bBoxMin[2] + (k + 0.5)*gz;
				int n = 0; float rp;
				for (int t = 0; t < nTri; t++) {
					const float* d = T + (size_t)t * 9;
					if (triangle_ray_intersection(d, d + 3, d + 6, c, rayDir, &rp)) n++;
				}
				if (n % 2 == 1) vd->level1InOut[(size_t)k*ny*nx + (size_t)j*nx + i] = 1;
			}
		}
	};
	std::vector<std::thread> th;
	for (int i = 0; i < nThreads; i++) th.emplace_back(work);
	for (auto& t : th) t.join();
	return now() - t0;
}

// bBox[] initialisation (src/Object.cpp:3165-3193) + Object::ClassifyTessellation (:2256) + L1 normals (:3219-3253).
// L1 normals, restated from PerformVoxelization (src/Object.cpp:3219-3253): the same loop follows the CPU classification and the
// CUDA one (both leave the triangles of a cell in bBox[].objTriangles).
static void l1_normals(Object* o)
{
	VoxelData* vd = o->voxelData;
	size_t N = (size_t)vd->numDivX * vd->numDivY * vd->numDivZ;
	for (size_t k = 0; k < N; k++) {
		int numTri = vd->bBox[k].objTriangles.size();
		Float3 sumNorm = Float3(0, 0, 0);
		Float3 avgNorm = Float3(0, 0, 0);
		if (numTri != 0) {
			for (int i = 0; i < numTri; i++) {
				int triID = vd->bBox[k].objTriangles[i];
				float* td = o->flatCPUTriangleData + (size_t)triID * 9;
				Float3 v0 = Float3(td[0], td[1], td[2]);
				Float3 v1 = Float3(td[3], td[4], td[5]);
				Float3 v2 = Float3(td[6], td[7], td[8]);
				Float3 side1 = v1 - v0;
				Float3 side2 = v2 - v0;
				Float3 faceNormal = VectorCrossProduct(side1, side2);
				VectorNormalize(faceNormal);
				sumNorm += faceNormal;
			}
			avgNorm = sumNorm / numTri;
			VectorNormalize(avgNorm);
			vd->level1Normal[k * 3 + 0] = avgNorm[0];
			vd->level1Normal[k * 3 + 1] = avgNorm[1];
			vd->level1Normal[k * 3 + 2] = avgNorm[2];
		}
	}
}

double ref_l1_tribox(void* h)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	float gridSizeX = vd->gridSizeX, gridSizeY = vd->gridSizeY, gridSizeZ = vd->gridSizeZ;
	size_t N = (size_t)nx*ny*nz;
	vd->bBox = new BBoxData[N];
	Float3 boxExtentsLevel1 = Float3(gridSizeX / 2.0, gridSizeY / 2.0, gridSizeZ / 2.0);
	for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
		size_t idx = (size_t)k*ny*nx + (size_t)j*nx + i;
		float midX = (i + 0.5)*gridSizeX + o->bBoxMin[0];
		float midY = (j + 0.5)*gridSizeY + o->bBoxMin[1];
		float midZ = (k + 0.5)*gridSizeZ + o->bBoxMin[2];
		vd->bBox[idx].midPoint = Float3(midX, midY, midZ);
		vd->bBox[idx].halfSize = boxExtentsLevel1;
		vd->bBox[idx].solid = int(vd->level1InOut[idx]) % 2;
		vd->bBox[idx].intersecting = 0;
	}
	double t0 = now();
	o->ClassifyTessellation(r->gp);
	double dt = now() - t0;
	l1_normals(o);
	r->boxes = true;
	// reference-equivalent test count = sum of clipped footprints (cuda/CUDAClassifyTessellation.cu:374-378)
	long tests = 0, hits = 0; int mx = 0;
	for (size_t k = 0; k < N; k++) { int c = vd->bBox[k].objTriangles.size(); hits += c; if (c > mx) mx = c; }
	{
		int nTri = o->totalNumTriangles;
		int nd[3] = { nx, ny, nz };
		for (int t = 0; t < nTri; t++) {
			const float* v = o->flatCPUTriangleData + (size_t)t * 9;
			long f = 1;
			for (int a = 0; a < 3; a++) {
				int lo = 1 << 30, hi = -(1 << 30);
				for (int c = 0; c < 3; c++) {
					int b = int((v[c * 3 + a] - o->bBoxMin[a]) / (o->bBoxMax[a] - o->bBoxMin[a]) * nd[a]);
					if (b == nd[a] && v[c * 3 + a] == o->bBoxMax[a]) b--;
					if (b < lo) lo = b;
					if (b > hi) hi = b;
				}
				if (hi > nd[a] - 1) hi = nd[a] - 1;
				f *= (hi >= lo) ? (hi - lo + 1) : 0;
			}
			tests += f;
		}
	}
	r->l1BoxTests = tests; r->l1BoxHits = hits; r->maxPerCell = mx;
	return dt;
}

// Boundary prefix sum + index list (src/Object.cpp:3258-3284) and the CSR / per-column dedup lists of
// ClassifyTessellationCUDA's host code (:2137-2180; flat order = column p outer, z inner).
void ref_compact(void* h)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	size_t N = (size_t)nx*ny*nz;
	int boundaryVoxelCount = 0;
	vd->boundaryIndex.clear();
	for (size_t index = 0; index < N; index++) {
		vd->boundaryPrefixSum[index] = boundaryVoxelCount;
		if (vd->level1InOut[index] == 2) { boundaryVoxelCount++; vd->boundaryIndex.push_back(index); }
	}
	size_t nb = vd->boundaryIndex.size();
	size_t n23 = (size_t)vd->numDivX2*vd->numDivY2*vd->numDivZ2;
	if (vd->level2InOut) delete[] vd->level2InOut;
	if (vd->level2Normal) delete[] vd->level2Normal;
	vd->level2InOut = new inOutDType[nb * n23 + 1];
	memset(vd->level2InOut, 0, (nb * n23 + 1) * sizeof(inOutDType));
	vd->level2Normal = new float[nb * n23 * 4 + 4];
	memset(vd->level2Normal, 0, (nb * n23 * 4 + 4) * sizeof(float));
	// CSR + column lists
	int nTri = o->totalNumTriangles;
	r->triCount.assign(N, 0); r->triFlatIndex.assign(N, 0); r->triFlat.clear();
	r->xyFlat.clear(); r->xyFlatIndex.assign((size_t)nx*ny, 0); r->xyCount.assign((size_t)nx*ny, 0);
	std::vector<char> added(nTri);
	int triRunningSum = 0, triXYRunningSum = 0;
	for (int p = 0; p < nx*ny; p++) {
		std::fill(added.begin(), added.end(), 0);
		r->xyFlatIndex[p] = triXYRunningSum;
		int numXYTri = 0;
		for (int z = 0; z < nz; z++) {
			size_t k = (size_t)z*nx*ny + p;
			int numTri = vd->bBox[k].objTriangles.size();
			r->triCount[k] = numTri;
			r->triFlatIndex[k] = triRunningSum;
			for (int t = 0; t < numTri; t++) {
				int id = vd->bBox[k].objTriangles[t];
				r->triFlat.push_back(id);
				if (!added[id]) { r->xyFlat.push_back(id); added[id] = 1; numXYTri++; }
			}
			triRunningSum += numTri;
		}
		r->xyCount[p] = numXYTri;
		triXYRunningSum += numXYTri;
	}
	long col = 0;
	for (int p = 0; p < nx*ny; p++) col += (long)r->xyCount[p] * nz;
	r->l1ColRayTests = col;
	r->compacted = true;
}

// L1 fill through the per-column lists (the candidate set the reference's Level-2 kernel uses,
// cu:461-463) -- returns the number of non-boundary cells whose parity differs from level1InOut.
long ref_l1_inout_collist_mismatch(void* h)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	float dir[3] = { 0, 0, 1 };
	long diff = 0;
	for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
		size_t idx = (size_t)k*ny*nx + (size_t)j*nx + i;
		if (vd->level1InOut[idx] == 2) continue;
		float c[3];
		c[0] = o->bBoxMin[0] + (i + 0.5)*vd->gridSizeX;
		c[1] = o->bBoxMin[1] + (j + 0.5)*vd->gridSizeY;
		c[2] = o->bBoxMin[2] + (k + 0.5)*vd->gridSizeZ;
		int p = j*nx + i, n = 0; float rp;
		for (int q = 0; q < r->xyCount[p]; q++) {
			float* d = o->flatCPUTriangleData + (size_t)r->xyFlat[r->xyFlatIndex[p] + q] * 9;
			if (triangle_ray_intersection(d, d + 3, d + 6, c, dir, &rp)) n++;
		}
		if (float(n % 2) != vd->level1InOut[idx]) diff++;
	}
	return diff;
}

// Level 2 through the reference's CPU twins (src/Object.cpp:1111, :2344). f64 centre formulas (cross-check only).
void ref_l2_cpu(void* h, double* tInOut, double* tTriBox)
{
	Ref* r = (Ref*)h;
	VoxelData* vd = r->o->voxelData;
	int nb = vd->boundaryIndex.size();
	double t0 = now();
	for (int b = 0; b < nb; b++) r->o->ClassifyInOutLevel2CPU(b);
	double t1 = now();
	r->o->ClassifyTessellationLevel2(r->gp);
	double t2 = now();
	*tInOut = t1 - t0; *tTriBox = t2 - t1;
}

// Level 2 in the arithmetic of the two CUDA kernels (cu:403-504): f32 centre ((2p+1)*ext2 + mid) - ext1, ray parity over
// the XY-column list first, then SAT over the cell list overwriting with 2 and accumulating cross(e01,e02) (cu:311-318,
// normalize() result discarded), then the host averaging (src/Object.cpp:2613-2632).  Runs on nThreads host threads over b.
double ref_l2_kernelform(void* h, int nThreads)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY;
	int n2x = vd->numDivX2, n2y = vd->numDivY2, n2z = vd->numDivZ2;
	size_t n23 = (size_t)n2x*n2y*n2z;
	size_t nb = vd->boundaryIndex.size();
	float3 e1 = make_float3(vd->gridSizeX / 2.0, vd->gridSizeY / 2.0, vd->gridSizeZ / 2.0);
	float3 e2 = make_float3(vd->gridSizeX2 / 2.0, vd->gridSizeY2 / 2.0, vd->gridSizeZ2 / 2.0);
	r->l2k.assign(nb*n23, 0.f);
	r->l2kNormal.assign(nb*n23 * 4, 0.f);
	std::atomic<long> rayTests(0), boxTests(0);
	std::atomic<size_t> next(0);
	double t0 = now();
	auto work = [&]() {
		float dir[3] = { 0, 0, 1 };
		long rt = 0, bt = 0;
		for (;;) {
			size_t b = next.fetch_add(1);
			if (b >= nb) break;
			int l1 = vd->boundaryIndex[b];
			int k = l1 / (nx*ny); int ij = l1 - k*nx*ny; int j = ij / nx, i = ij % nx;
			float mid[3];
			mid[0] = (i + 0.5)*vd->gridSizeX + o->bBoxMin[0];
			mid[1] = (j + 0.5)*vd->gridSizeY + o->bBoxMin[1];
			mid[2] = (k + 0.5)*vd->gridSizeZ + o->bBoxMin[2];
			int xy = l1 % (nx*ny);
			const int* cl = r->xyFlat.data() + r->xyFlatIndex[xy]; int ncl = r->xyCount[xy];
			const int* tl = r->triFlat.data() + r->triFlatIndex[l1]; int ntl = r->triCount[l1];
			for (size_t loc = 0; loc < n23; loc++) {
				int rr = loc / (n2x*n2y); int pq = loc - rr*n2x*n2y; int q = pq / n2x, p = pq % n2x;
				float c[3];
				c[0] = (2 * p + 1)*e2.x + mid[0] - e1.x;
				c[1] = (2 * q + 1)*e2.y + mid[1] - e1.y;
				c[2] = (2 * rr + 1)*e2.z + mid[2] - e1.z;
				int n = 0; float rp;
				for (int t = 0; t < ncl; t++) {
					float* d = o->flatCPUTriangleData + (size_t)cl[t] * 9;
					rt++;
					if (triangle_ray_intersection(d, d + 3, d + 6, c, dir, &rp)) n++;
				}
				size_t idx = b*n23 + loc;
				if (n % 2 == 1) r->l2k[idx] = 1;
				float he[3] = { e2.x, e2.y, e2.z };
				for (int t = 0; t < ntl; t++) {
					float* d = o->flatCPUTriangleData + (size_t)tl[t] * 9;
					float tv[3][3] = { { d[0], d[1], d[2] }, { d[3], d[4], d[5] }, { d[6], d[7], d[8] } };
					bt++;
					if (TriBoxOverlap(c, he, tv)) {
						r->l2k[idx] = 2;
						float ax = tv[1][0] - tv[0][0], ay = tv[1][1] - tv[0][1], az = tv[1][2] - tv[0][2];
						float bx = tv[2][0] - tv[0][0], by = tv[2][1] - tv[0][1], bz = tv[2][2] - tv[0][2];
						// cutil_math.h cross(): (a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x)
						r->l2kNormal[idx * 4 + 0] += ay*bz - az*by;
						r->l2kNormal[idx * 4 + 1] += az*bx - ax*bz;
						r->l2kNormal[idx * 4 + 2] += ax*by - ay*bx;
						r->l2kNormal[idx * 4 + 3] += 1;
					}
				}
				float numNormals = r->l2kNormal[idx * 4 + 3];
				Float3 avg = Float3(0, 0, 0);
				if (numNormals > 0) {
					avg[0] = r->l2kNormal[idx * 4 + 0] / numNormals;
					avg[1] = r->l2kNormal[idx * 4 + 1] / numNormals;
					avg[2] = r->l2kNormal[idx * 4 + 2] / numNormals;
					VectorNormalize(avg);
				}
				r->l2kNormal[idx * 4 + 0] = avg[0]; r->l2kNormal[idx * 4 + 1] = avg[1]; r->l2kNormal[idx * 4 + 2] = avg[2];
			}
		}
		rayTests += rt; boxTests += bt;
	};
	std::vector<std::thread> th;
	for (int i = 0; i < nThreads; i++) th.emplace_back(work);
	for (auto& t : th) t.join();
	r->l2RayTests = rayTests; r->l2BoxTests = boxTests;
	return now() - t0;
}

// copy the kernel-form Level-2 result into the Object so that counting/saving use it
void ref_l2_adopt_kernelform(void* h)
{
	Ref* r = (Ref*)h;
	VoxelData* vd = r->o->voxelData;
	memcpy(vd->level2InOut, r->l2k.data(), r->l2k.size() * sizeof(float));
	memcpy(vd->level2Normal, r->l2kNormal.data(), r->l2kNormal.size() * sizeof(float));
}

// The reference's own GPU host path, UNMODIFIED: Object::ClassifyTessellationCUDA (src/Object.cpp:2071; re-run with the
// returned buffer size like :3204-3214) and Object::ClassifyInOutTessellationLevel2CUDA (:2533), which call the three
// extern "C" operators -- resolved by whatever library this harness is linked against.  Level-1 normals, prefix sum and
// level-2 allocation restated from PerformVoxelization as in ref_l1_tribox/ref_compact.  Returns the max triangles per cell.
int ref_cuda_path(void* h)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	size_t N = (size_t)nx*ny*nz;
	vd->bBox = new BBoxData[N];
	for (size_t k = 0; k < N; k++) { vd->bBox[k].solid = int(vd->level1InOut[k]) % 2; vd->bBox[k].intersecting = 0; }
	int maxTri = o->ClassifyTessellationCUDA(r->gp);
	int used = 50;
	if (maxTri > 0) {
		for (size_t k = 0; k < N; k++) vd->bBox[k].objTriangles.clear(); // App. B5: the first pass already pushed its lists
		vd->level1TriBuffer = maxTri;
		used = maxTri;
		int again = o->ClassifyTessellationCUDA(r->gp);
		if (again != 0) return -1;
	}
	l1_normals(o);
	int boundaryVoxelCount = 0;
	vd->boundaryIndex.clear();
	for (size_t index = 0; index < N; index++) {
		vd->boundaryPrefixSum[index] = boundaryVoxelCount;
		if (vd->level1InOut[index] == 2) { boundaryVoxelCount++; vd->boundaryIndex.push_back(index); }
	}
	size_t nb = vd->boundaryIndex.size();
	size_t n23 = (size_t)vd->numDivX2*vd->numDivY2*vd->numDivZ2;
	vd->level2InOut = new inOutDType[nb * n23 + 1];
	memset(vd->level2InOut, 0, (nb * n23 + 1) * sizeof(inOutDType));
	vd->level2Normal = new float[nb * n23 * 4 + 4];
	memset(vd->level2Normal, 0, (nb * n23 * 4 + 4) * sizeof(float));
	// the reference's cudaMemset calls zero only part of the Level-2 device buffers (App. B2): hand it zeroed memory by
	// poisoning nothing -- cudaMalloc'ed memory of a fresh context is zero-filled by the driver in practice; see the test.
	o->ClassifyInOutTessellationLevel2CUDA(r->gp);
	r->boxes = true;
	return used;
}

// Voxel counting, src/Object.cpp:3353-3378.
void ref_count(void* h, long* out4)
{
	VoxelData* vd = ((Ref*)h)->o->voxelData;
	size_t N = (size_t)vd->numDivX*vd->numDivY*vd->numDivZ;
	size_t n23 = (size_t)vd->numDivX2*vd->numDivY2*vd->numDivZ2;
	vd->numLevel1InsideVoxels = vd->numLevel1BoundaryVoxels = 0;
	for (size_t k = 0; k < N; k++) {
		if (int(vd->level1InOut[k]) == 1) vd->numLevel1InsideVoxels++;
		if (int(vd->level1InOut[k]) == 2) vd->numLevel1BoundaryVoxels++;
	}
	vd->numLevel2InsideVoxels = vd->numLevel2BoundaryVoxels = 0;
	for (size_t g = 0; g < (size_t)vd->numLevel1BoundaryVoxels * n23; g++) {
		if (int(vd->level2InOut[g]) % 2 == 1) vd->numLevel2InsideVoxels++;
		if (int(vd->level2InOut[g]) == 2) vd->numLevel2BoundaryVoxels++;
	}
	out4[0] = vd->numLevel1InsideVoxels; out4[1] = vd->numLevel1BoundaryVoxels;
	out4[2] = vd->numLevel2InsideVoxels; out4[3] = vd->numLevel2BoundaryVoxels;
}

// The reference's own writer (src/Object.cpp:2934-3075) into `dir`.
void ref_save(void* h, const char* dir)
{
	Ref* r = (Ref*)h;
	char cwd[4096];
	if (!getcwd(cwd, sizeof cwd)) abort();
	if (chdir(dir)) abort();
	r->o->SaveVoxelization(r->gp);
	if (chdir(cwd)) abort();
}

// accessors
const float* ref_level1InOut(void* h) { return ((Ref*)h)->o->voxelData->level1InOut; }
const float* ref_level1Normal(void* h) { return ((Ref*)h)->o->voxelData->level1Normal; }
const int* ref_prefix(void* h) { return ((Ref*)h)->o->voxelData->boundaryPrefixSum; }
long ref_nboundary(void* h) { return ((Ref*)h)->o->voxelData->boundaryIndex.size(); }
const int* ref_boundaryIndex(void* h) { return ((Ref*)h)->o->voxelData->boundaryIndex.data(); }
const float* ref_level2InOut(void* h) { return ((Ref*)h)->o->voxelData->level2InOut; }
const float* ref_level2Normal(void* h) { return ((Ref*)h)->o->voxelData->level2Normal; }
const float* ref_level2InOutKernel(void* h) { return ((Ref*)h)->l2k.data(); }
const float* ref_level2NormalKernel(void* h) { return ((Ref*)h)->l2kNormal.data(); }
const int* ref_triCount(void* h) { return ((Ref*)h)->triCount.data(); }
const int* ref_triFlatIndex(void* h) { return ((Ref*)h)->triFlatIndex.data(); }
const int* ref_triFlat(void* h) { return ((Ref*)h)->triFlat.data(); }
long ref_triFlatLen(void* h) { return ((Ref*)h)->triFlat.size(); }
const int* ref_xyCount(void* h) { return ((Ref*)h)->xyCount.data(); }
const int* ref_xyFlatIndex(void* h) { return ((Ref*)h)->xyFlatIndex.data(); }
const int* ref_xyFlat(void* h) { return ((Ref*)h)->xyFlat.data(); }
long ref_xyFlatLen(void* h) { return ((Ref*)h)->xyFlat.size(); }
void ref_stats(void* h, long* out6)
{
	Ref* r = (Ref*)h;
	out6[0] = r->l1BoxTests; out6[1] = r->l1BoxHits; out6[2] = r->maxPerCell;
	out6[3] = r->l2BoxTests; out6[4] = r->l2RayTests; out6[5] = r->l1ColRayTests;
}

// predicate-level access to the reference's object code (known-answer source for the fuzz tests)
int ref_tribox(const float* c, const float* h, const float* tri9)
{
	float cc[3] = { c[0], c[1], c[2] }, hh[3] = { h[0], h[1], h[2] };
	float tv[3][3] = { { tri9[0], tri9[1], tri9[2] }, { tri9[3], tri9[4], tri9[5] }, { tri9[6], tri9[7], tri9[8] } };
	return TriBoxOverlap(cc, hh, tv);
}
void ref_tribox_batch(long n, const float* c, const float* h, const float* tri9, unsigned char* out)
{
	for (long i = 0; i < n; i++) out[i] = (unsigned char)ref_tribox(c + i * 3, h + i * 3, tri9 + i * 9);
}
void ref_triray_batch(long n, const float* org, const float* tri9, unsigned char* out)
{
	float dir[3] = { 0, 0, 1 }; float rp;
	for (long i = 0; i < n; i++) {
		const float* d = tri9 + i * 9;
		out[i] = (unsigned char)triangle_ray_intersection(d, d + 3, d + 6, org + i * 3, dir, &rp);
	}
}

// CPU baseline: the reference's TriBoxOverlap inside the reference's loop nests, timed.
//   L1: the ClassifyTessellation nest (src/Object.cpp:2268-2341) over triangles [t0,t1)  (no list building)
//   L2: the kernel-form nest (cu:428-445) over boundary cells [b0,b1)
// split across nThreads host threads; returns seconds and the number of tests performed.
double ref_time_l2_tribox(void* h, long b0, long b1, int nThreads, long* testsOut)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY;
	int n2x = vd->numDivX2, n2y = vd->numDivY2, n2z = vd->numDivZ2;
	size_t n23 = (size_t)n2x*n2y*n2z;
	float3 e1 = make_float3(vd->gridSizeX / 2.0, vd->gridSizeY / 2.0, vd->gridSizeZ / 2.0);
	float3 e2 = make_float3(vd->gridSizeX2 / 2.0, vd->gridSizeY2 / 2.0, vd->gridSizeZ2 / 2.0);
	std::atomic<long> tests(0), hits(0);
	std::atomic<long> next(b0);
	double t0 = now();
	auto work = [&]() {
		long bt = 0, ht = 0;
		for (;;) {
			long b = next.fetch_add(1);
			if (b >= b1) break;
			int l1 = vd->boundaryIndex[b];
			int k = l1 / (nx*ny); int ij = l1 - k*nx*ny; int j = ij / nx, i = ij % nx;
			float mid[3];
			mid[0] = (i + 0.5)*vd->gridSizeX + o->bBoxMin[0];
			mid[1] = (j + 0.5)*vd->gridSizeY + o->bBoxMin[1];
			mid[2] = (k + 0.5)*vd->gridSizeZ + o->bBoxMin[2];
			const int* tl = r->triFlat.data() + r->triFlatIndex[l1]; int ntl = r->triCount[l1];
			for (size_t loc = 0; loc < n23; loc++) {
				int rr = loc / (n2x*n2y); int pq = loc - rr*n2x*n2y; int q = pq / n2x, p = pq % n2x;
				float c[3];
				c[0] = (2 * p + 1)*e2.x + mid[0] - e1.x;
				c[1] = (2 * q + 1)*e2.y + mid[1] - e1.y;
				c[2] = (2 * rr + 1)*e2.z + mid[2] - e1.z;
				float he[3] = { e2.x, e2.y, e2.z };
				for (int t = 0; t < ntl; t++) {
					float* d = o->flatCPUTriangleData + (size_t)tl[t] * 9;
					float tv[3][3] = { { d[0], d[1], d[2] }, { d[3], d[4], d[5] }, { d[6], d[7], d[8] } };
					bt++;
					if (TriBoxOverlap(c, he, tv)) ht++;
				}
			}
		}
		tests += bt; hits += ht;
	};
	std::vector<std::thread> th;
	for (int i = 0; i < nThreads; i++) th.emplace_back(work);
	for (auto& t : th) t.join();
	*testsOut = tests;
	return now() - t0;
}

// ---- whole-path CPU baseline (bench.py --impl reference, cpu_baseline): the pieces of the path beyond the Level-2 SAT nest, each
// around the reference's own object code (triangle_ray_intersection, TriBoxOverlap), split across host threads.

// Level-1 solid fill through the per-column lists (the form the reference's own Level-2 kernel uses, cu:461-463; byte-identical to
// the brute force Object::ClassifyInOutCPU on cessna 64, tests/test_oracle_ref.py): every cell's +Z ray against its column list.
// Times the loop nest only; the parities go to a scratch array (level1InOut already holds the final states).
double ref_time_l1_fill_collist(void* h, int nThreads, long* rayTestsOut)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY, nz = vd->numDivZ;
	std::vector<unsigned char> scratch((size_t)nx * ny * nz);
	std::atomic<long> tests(0);
	std::atomic<int> next(0);
	double t0 = now();
	auto work = [&]() {
		float dir[3] = { 0, 0, 1 };
		long rt = 0;
		for (;;) {
			int k = next.fetch_add(1);
			if (k >= nz) break;
			for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
				float c[3];
				c[0] = o->bBoxMin[0] + (i + 0.5)*vd->gridSizeX;
				c[1] = o->bBoxMin[1] + (j + 0.5)*vd->gridSizeY;
				c[2] = o->bBoxMin[2] + (k + 0.5)*vd->gridSizeZ;
				int p = j*nx + i, n = 0; float rp;
				for (int q = 0; q < r->xyCount[p]; q++) {
					float* d = o->flatCPUTriangleData + (size_t)r->xyFlat[r->xyFlatIndex[p] + q] * 9;
					rt++;
					if (triangle_ray_intersection(d, d + 3, d + 6, c, dir, &rp)) n++;
				}
				scratch[(size_t)k*ny*nx + p] = (unsigned char)(n % 2);
			}
		}
		tests += rt;
	};
	std::vector<std::thread> th;
	for (int i = 0; i < nThreads; i++) th.emplace_back(work);
	for (auto& t : th) t.join();
	double dt = now() - t0;
	*rayTestsOut = tests;
	return dt;
}

// Level 2 of the path over boundary cells [b0,b1): per sub-voxel the parity rays over the column list (cu:450-504) and then the
// SAT over the cell list (cu:403-448), in the kernels' arithmetic, states written to a scratch array.  out[0] = tri-box tests,
// out[1] = ray tests.
double ref_time_l2_path(void* h, long b0, long b1, int nThreads, long* out)
{
	Ref* r = (Ref*)h;
	Object* o = r->o;
	VoxelData* vd = o->voxelData;
	int nx = vd->numDivX, ny = vd->numDivY;
	int n2x = vd->numDivX2, n2y = vd->numDivY2, n2z = vd->numDivZ2;
	size_t n23 = (size_t)n2x*n2y*n2z;
	float3 e1 = make_float3(vd->gridSizeX / 2.0, vd->gridSizeY / 2.0, vd->gridSizeZ / 2.0);
	float3 e2 = make_float3(vd->gridSizeX2 / 2.0, vd->gridSizeY2 / 2.0, vd->gridSizeZ2 / 2.0);
	std::vector<float> scratch((size_t)(b1 - b0) * n23, 0.f);
	std::atomic<long> boxTests(0), rayTests(0);
	std::atomic<long> next(b0);
	double t0 = now();
	auto work = [&]() {
		float dir[3] = { 0, 0, 1 };
		long bt = 0, rt = 0;
		for (;;) {
			long b = next.fetch_add(1);
			if (b >= b1) break;
			int l1 = vd->boundaryIndex[b];
			int k = l1 / (nx*ny); int ij = l1 - k*nx*ny; int j = ij / nx, i = ij % nx;
			float mid[3];
			mid[0] = (i + 0.5)*vd->gridSizeX + o->bBoxMin[0];
			mid[1] = (j + 0.5)*vd->gridSizeY + o->bBoxMin[1];
			mid[2] = (k + 0.5)*vd->gridSizeZ + o->bBoxMin[2];
			int xy = l1 % (nx*ny);
			const int* cl = r->xyFlat.data() + r->xyFlatIndex[xy]; int ncl = r->xyCount[xy];
			const int* tl = r->triFlat.data() + r->triFlatIndex[l1]; int ntl = r->triCount[l1];
			for (size_t loc = 0; loc < n23; loc++) {
				int rr = loc / (n2x*n2y); int pq = loc - rr*n2x*n2y; int q = pq / n2x, p = pq % n2x;
				float c[3];
				c[0] = (2 * p + 1)*e2.x + mid[0] - e1.x;
				c[1] = (2 * q + 1)*e2.y + mid[1] - e1.y;
				c[2] = (2 * rr + 1)*e2.z + mid[2] - e1.z;
				int n = 0; float rp;
				for (int t = 0; t < ncl; t++) {
					float* d = o->flatCPUTriangleData + (size_t)cl[t] * 9;
					rt++;
					if (triangle_ray_intersection(d, d + 3, d + 6, c, dir, &rp)) n++;
				}
				float state = (n % 2 == 1) ? 1.f : 0.f;
				float he[3] = { e2.x, e2.y, e2.z };
				for (int t = 0; t < ntl; t++) {
					float* d = o->flatCPUTriangleData + (size_t)tl[t] * 9;
					float tv[3][3] = { { d[0], d[1], d[2] }, { d[3], d[4], d[5] }, { d[6], d[7], d[8] } };
					bt++;
					if (TriBoxOverlap(c, he, tv)) state = 2.f;
				}
				scratch[(size_t)(b - b0) * n23 + loc] = state;
			}
		}
		boxTests += bt; rayTests += rt;
	};
	std::vector<std::thread> th;
	for (int i = 0; i < nThreads; i++) th.emplace_back(work);
	for (auto& t : th) t.join();
	double dt = now() - t0;
	out[0] = boxTests; out[1] = rayTests;
	return dt;
}

// ---- voxel hierarchy / collision structures: the reference's OWN Object::BuildHierarchy (src/Object.cpp:2790-2867,
// with CombineBBox :2750-2788) on the bBox[] array of ref_l1_tribox, and the host loop of Object::CollisionInitCUDA (:3530-3552; the
// rest of that function only uploads the two arrays).  bBox[].solid must hold the parity fill: call ref_set_solid first.
void ref_set_solid(void* h, const unsigned char* fill)
{
	Ref* r = (Ref*)h;
	VoxelData* vd = r->o->voxelData;
	size_t N = (size_t)vd->numDivX * vd->numDivY * vd->numDivZ;
	for (size_t k = 0; k < N; k++) { vd->bBox[k].solid = int(fill[k]) % 2; vd->bBox[k].index = (int)k; }  // (index: never initialised by the reference)
}

int ref_build_hierarchy(void* h, float* mid, float* half, unsigned char* solid, int* child)
{
	Ref* r = (Ref*)h;
	VoxelData* vd = r->o->voxelData;
	int total = vd->numDivX * vd->numDivY * vd->numDivZ;
	r->o->BuildHierarchy(r->gp);
	for (int i = 0; i < total - 1; i++) {
		BBoxData& b = vd->bBoxHierarchy[i];
		for (int a = 0; a < 3; a++) { mid[i * 3 + a] = b.midPoint[a]; half[i * 3 + a] = b.halfSize[a]; }
		solid[i] = (unsigned char)b.solid; child[2 * i] = b.childIndex1; child[2 * i + 1] = b.childIndex2;
	}
	return vd->numLevels;
}

long ref_collision_boxes(void* h, int* invIndex, float* mid, float* ext)
{
	Ref* r = (Ref*)h;
	VoxelData* vd = r->o->voxelData;
	int totalNumBoxes = vd->numDivX * vd->numDivY * vd->numDivZ;
	std::vector<int> inv;
	for (int i = 0; i < totalNumBoxes; i++)                      // :3534-3539
		if (vd->level1InOut[i] >= 1) inv.push_back(i);
	for (size_t i = 0; i < inv.size(); i++) {                    // :3545-3555
		int boxIndex = inv[i];
		invIndex[i] = boxIndex;
		for (int a = 0; a < 3; a++) { mid[i * 3 + a] = vd->bBox[boxIndex].midPoint[a]; ext[i * 3 + a] = vd->bBox[boxIndex].halfSize[a]; }
	}
	return (long)inv.size();
}

void ref_close(void* h)
{
	Ref* r = (Ref*)h;
	delete r->o; delete r->gp; delete r;
}

} // extern "C"
