/* oracle/gpv_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's hybrid two-level voxelizer (idealab-isu/GPView), the checker the CUDA
 * path is compared against.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (gpview_b200/) never does.
 *
 * Parity status: PINNED.  The reference ships no golden vectors, so the oracle is pinned against the
 * outputs of the reference's own code run in the build container (oracle/_ref, tests/test_oracle_ref.py) and against the
 * fixtures those runs produced (tests/golden/, tests/test_oracle_golden.py).
 *
 * Every function cites the reference lines it restates (paths relative to the reference root).
 * Build: gcc -std=c11 -O2 -ffp-contract=off (binary32 ops in source order, no FMA; double exactly where C promotes).
 */
#ifndef GPV_ORACLE_H
#define GPV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	int64_t nTri;
	float* tris;          /* nTri*9: v0xyz v1xyz v2xyz in file face order (src/Object.cpp:3496-3527) */
	float bmin[3], bmax[3]; /* padded bbox (src/Object.cpp:572-583 / :285-296) */
	float maxModelSize;
	int64_t nVerts;
} gpvo_mesh;

typedef struct {
	int numDiv[3];
	float gridSize[3], gridSize2[3], ext1[3], ext2[3];
	int n2;
} gpvo_grid;

typedef struct {
	gpvo_grid g;
	int64_t cells, nBoundary, n23;
	uint8_t* l1State;        /* cells, 0 outside / 1 inside / 2 boundary */
	uint8_t* l1FillOnly;     /* cells, parity fill before the SAT overwrite */
	int32_t* prefix;         /* cells, exclusive boundary prefix sum (src/Object.cpp:3270-3279) */
	int32_t* boundaryIndex;  /* nBoundary */
	int32_t* cellCount;      /* cells */
	int64_t* cellOffset;     /* cells+1, CSR in LINEAR index order, lists ascending */
	int32_t* cellTris;
	int32_t* colCount;       /* nx*ny */
	int64_t* colOffset;      /* nx*ny+1, lists ascending */
	int32_t* colTris;
	uint8_t* l1Normal;       /* cells*3, reference uchar encoding */
	uint8_t* l2State;        /* nBoundary*n23 */
	uint8_t* l2Normal;       /* nBoundary*n23*3 */
	int64_t l1Inside, l1Boundary, l2Inside, l2Boundary;
	int64_t l1BoxTests, l1BoxHits, maxPerCell, l2BoxTests, l2RayTests, l1ColRayTests;
	int64_t fillIllConditioned; /* triangles the certified fill had to test against every column */
	int64_t fillCrossings;
} gpvo_result;

/* flags for gpvo_voxelize */
#define GPVO_FILL_BRUTE      0  /* Object::ClassifyInOutCPU literally: every cell x every triangle */
#define GPVO_FILL_CERTIFIED  1  /* same result, per-triangle certified column culling + per-column t sweep */
#define GPVO_FILL_COLLIST    2  /* parity over the SAT column lists only (the candidate set of cu:461-463) */
#define GPVO_L2_NAIVE        16 /* Level-2 rays as written in the kernel (every voxel x every column triangle) */
#define GPVO_NO_NORMALS      32
#define GPVO_NO_L2           64

int gpvo_load_obj(const char* path, gpvo_mesh* m);
int gpvo_load_off(const char* path, gpvo_mesh* m);
void gpvo_mesh_from_tris(const float* tris, int64_t nTri, gpvo_mesh* m); /* bbox over the given vertices, padded */
void gpvo_free_mesh(gpvo_mesh* m);
void gpvo_make_grid(const float bmin[3], const float bmax[3], float maxModelSize, int voxelCount, int voxelCount2, gpvo_grid* g);
void gpvo_axis_table(const float bmin[3], const gpvo_grid* g, int axis, float* out); /* numDiv[axis] cell centres */

int gpvo_tribox(const float c[3], const float h[3], const float tri[9]);
int gpvo_triray(const float V1[3], const float V2[3], const float V3[3], const float O[3], const float D[3]);
void gpvo_tribox_batch(int64_t n, const float* c, const float* h, const float* tri9, uint8_t* out);
void gpvo_triray_batch(int64_t n, const float* org, const float* tri9, uint8_t* out);
void gpvo_triray_z_batch(int64_t n, const float* org, const float* tri9, uint8_t* out); /* axis-specialised form */

int gpvo_voxelize(const gpvo_mesh* m, int voxelCount, int voxelCount2, int flags, int nThreads, gpvo_result* r);
void gpvo_free_result(gpvo_result* r);
int gpvo_save(const gpvo_mesh* m, const gpvo_result* r, int objID, const char* dir); /* six files, src/Object.cpp:2934-3075 */

/* voxel hierarchy / collision structures over the Level-1 grid (gpv_oracle_collision.c) */
int64_t gpvo_collision_boxes(const gpvo_mesh* m, const gpvo_result* r, int32_t* invIndex, float* mid, float* ext);         /* Object::CollisionInitCUDA, src/Object.cpp:3530-3572 */
int gpvo_build_hierarchy(const gpvo_mesh* m, const gpvo_result* r, float* mid, float* half, uint8_t* solid, int32_t* child); /* Object::BuildHierarchy, :2790-2867 */

/* timed loop nests for bench.py's cpu_baseline ("port"): returns seconds, writes the number of tests done */
double gpvo_time_l2_tribox(const gpvo_mesh* m, const gpvo_result* r, int64_t b0, int64_t b1, int nThreads, int64_t* tests);

#ifdef __cplusplus
}
#endif
#endif
