// include/gpview_b200.hpp -- header-only C++ facade over the C ABI, shaped like the reference's Object / GLParameters so that
// GPView-style callers port by recompiling (SURVEY.md 8b tier 3).  Same member names and argument meaning:
//   Object::ReadObject / ReadOFFObject (src/Object.cpp:395, :171), CreateFlatTriangleData (:3496), PerformVoxelization (:3077),
//   SaveVoxelization (:2934), BuildHierarchy (:2790), CollisionInitCUDA (:3530); GLParameters::voxelCount / voxelCount2 / level2Voxels /
//   saveVoxels (src/GLParameters.cpp:28-69).
// Error behaviour: the reference abort()s on file errors and only prints CUDA errors; the facade throws gpview::Error.
#pragma once
#include "gpview_b200.h"
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace gpview {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) { if (rc) throw Error(gpv_last_error()); }

struct GLParameters {            // the voxelizer-relevant subset of the reference's GLParameters, same defaults
	int voxelCount = 8;          // src/GLParameters.cpp:66
	int voxelCount2 = 4;         // src/GLParameters.cpp:67
	bool level2Voxels = true;
	bool saveVoxels = true;
	bool normals = true;         // the reference always computes normals
	bool packedTransfer = true;  // Level2InOut crosses PCIe as 2 bits per sub-voxel, expanded by host threads (GPV_PACKED_L2): same bytes, a quarter of the transfer
	bool collision = false;      // keep what BuildHierarchy needs (the parity fill of the boundary cells, GPV_COLLISION)
	int device = 0;
};

struct BBoxData {                // one box of the hierarchy: the reference's BBoxData without the per-cell triangle vectors (includes/Utilities.h:57-78)
	float midPoint[3], halfSize[3];
	int solid, childIndex1, childIndex2, index;
};

struct VoxelData {               // host copies of the streams, file encoding
	int numDivX = 0, numDivY = 0, numDivZ = 0, numDivX2 = 0, numDivY2 = 0, numDivZ2 = 0;
	float gridSizeX = 0, gridSizeY = 0, gridSizeZ = 0, gridSizeX2 = 0, gridSizeY2 = 0, gridSizeZ2 = 0;
	int64_t numLevel1InsideVoxels = 0, numLevel1BoundaryVoxels = 0, numLevel2InsideVoxels = 0, numLevel2BoundaryVoxels = 0;
	std::vector<uint8_t> level1InOut, level1Normal, level2InOut, level2Normal;
	std::vector<int32_t> boundaryPrefixSum, boundaryIndex;
	// Object::CollisionInitCUDA: occupied cells (ascending) and their boxes; the two arrays stay on the device like the reference's
	std::vector<int> invIndex;
	const float* boxCenterCUDAData = nullptr; const float* boxExtentCUDAData = nullptr;
	bool collisionInit = false;
	// Object::BuildHierarchy
	int numLevels = 0;
	std::vector<BBoxData> bBoxHierarchy;
	gpv_result result{};
};

class Object {
public:
	int objID = -1;              // the first OBJ on GPView's command line gets -1 (src/GPView.cpp:181, App. B8)
	int totalNumTriangles = 0;
	float bBoxMin[3] = { 0, 0, 0 }, bBoxMax[3] = { 0, 0, 0 }, maxModelSize = 0;
	const float* flatCPUTriangleData = nullptr;
	VoxelData voxelData;
	bool voxelInit = false;

	Object() { mesh_.tris = nullptr; }
	~Object() { gpv_free_mesh(&mesh_); if (ctx_) gpv_destroy(ctx_); }
	Object(const Object&) = delete;
	Object& operator=(const Object&) = delete;

	void ReadObject(const char* fname) { gpv_free_mesh(&mesh_); check(gpv_load_obj(fname, &mesh_)); adopt(); }
	void ReadOFFObject(const char* fname) { gpv_free_mesh(&mesh_); check(gpv_load_off(fname, &mesh_)); adopt(); }
	void ReadMesh(const char* fname, bool tolerant = false) { gpv_free_mesh(&mesh_); check(gpv_load_mesh_ex(fname, tolerant ? GPV_LOAD_TOLERANT : 0u, &mesh_)); adopt(); } // tolerant: extension, see gpv_load_mesh_ex
	void CreateFlatTriangleData() { flatCPUTriangleData = mesh_.tris; totalNumTriangles = (int)mesh_.n_tri; } // already flat

	// bufferSize is accepted for source compatibility and ignored: lists are CSR and cannot overflow (src/Object.cpp:3088)
	void PerformVoxelization(const GLParameters* glParam, int /*bufferSize*/ = -1)
	{
		if (!ctx_) check(gpv_create(glParam->device, &ctx_));
		gpv_params p{ glParam->voxelCount, glParam->voxelCount2, (glParam->normals ? GPV_NORMALS : 0) | (glParam->level2Voxels ? 0 : GPV_NO_LEVEL2) |
			                                                  (glParam->packedTransfer ? GPV_PACKED_L2 : 0) | (glParam->collision ? GPV_COLLISION : 0), 0, 0 };
		gpv_grid g;
		check(gpv_make_grid(mesh_.bbox_min, mesh_.bbox_max, mesh_.max_model_size, p.voxel_count, glParam->level2Voxels ? p.voxel_count2 : 1, &g));
		VoxelData& v = voxelData;
		const size_t cells = (size_t)g.num_div[0] * g.num_div[1] * g.num_div[2];
		v.level1InOut.resize(cells); v.boundaryPrefixSum.resize(cells);
		if (glParam->normals) v.level1Normal.resize(cells * 3);
		// Level-2 sizes are only known after the Level-1 pass: run Level 1 alone first, then the full call into exact buffers
		gpv_params p1 = p; p1.flags |= GPV_NO_LEVEL2; p1.flags &= ~(GPV_NORMALS | GPV_COLLISION);
		gpv_host_streams none{};
		check(gpv_voxelize_host(ctx_, &mesh_, &p1, nullptr, &v.result, &none));
		const size_t nb = (size_t)v.result.n_boundary, n23 = (size_t)g.n2 * g.n2 * g.n2;
		v.boundaryIndex.resize(nb);
		if (glParam->level2Voxels) { v.level2InOut.resize(nb * n23); if (glParam->normals) v.level2Normal.resize(nb * n23 * 3); }
		streams_ = gpv_host_streams{ v.level1InOut.data(), v.boundaryPrefixSum.data(), v.boundaryIndex.data(), glParam->level2Voxels ? v.level2InOut.data() : nullptr,
			                         glParam->normals ? v.level1Normal.data() : nullptr, (glParam->normals && glParam->level2Voxels) ? v.level2Normal.data() : nullptr,
			                         (int64_t)(nb * n23), (int64_t)nb };
		check(gpv_voxelize_host(ctx_, &mesh_, &p, nullptr, &v.result, &streams_));
		v.numDivX = g.num_div[0]; v.numDivY = g.num_div[1]; v.numDivZ = g.num_div[2];
		v.numDivX2 = v.numDivY2 = v.numDivZ2 = g.n2;
		v.gridSizeX = g.grid_size[0]; v.gridSizeY = g.grid_size[1]; v.gridSizeZ = g.grid_size[2];
		v.gridSizeX2 = g.grid_size2[0]; v.gridSizeY2 = g.grid_size2[1]; v.gridSizeZ2 = g.grid_size2[2];
		v.numLevel1InsideVoxels = v.result.l1_inside; v.numLevel1BoundaryVoxels = v.result.l1_boundary;
		v.numLevel2InsideVoxels = v.result.l2_inside; v.numLevel2BoundaryVoxels = v.result.l2_boundary;
		voxelInit = true;
		if (glParam->saveVoxels) SaveVoxelization(glParam);
	}

	// Object::CollisionInitCUDA (src/Object.cpp:3530-3572): the occupied cells of the last voxelization as an inverse index (host) and box
	// centre / extent arrays (device, owned by the library, valid until the next voxelization of this object)
	void CollisionInitCUDA(const GLParameters*)
	{
		if (!voxelInit) throw Error("CollisionInitCUDA before PerformVoxelization");
		gpv_collision c;
		check(gpv_collision_boxes(ctx_, nullptr, &c));
		voxelData.invIndex.resize((size_t)c.count);
		if (c.count) { check(gpv_memcpy_d2h(voxelData.invIndex.data(), c.d_inv_index, c.count * 4, nullptr)); check(gpv_stream_sync(nullptr)); }
		voxelData.boxCenterCUDAData = c.d_center; voxelData.boxExtentCUDAData = c.d_extent;
		voxelData.collisionInit = true;
	}

	// Object::BuildHierarchy (src/Object.cpp:2790-2867): needs GLParameters::collision = true at PerformVoxelization and a grid whose
	// dimensions are powers of two (the reference's loop is not defined on others: Error)
	void BuildHierarchy(const GLParameters*)
	{
		if (!voxelInit) throw Error("BuildHierarchy before PerformVoxelization");
		gpv_hierarchy h;
		check(gpv_build_hierarchy(ctx_, nullptr, &h));
		const size_t n = (size_t)h.n_boxes;
		std::vector<float> mid(n * 3), half(n * 3);
		std::vector<uint8_t> solid(n);
		std::vector<int32_t> child(n * 2);
		check(gpv_memcpy_d2h(mid.data(), h.d_mid, (int64_t)n * 12, nullptr)); check(gpv_memcpy_d2h(half.data(), h.d_half, (int64_t)n * 12, nullptr));
		check(gpv_memcpy_d2h(solid.data(), h.d_solid, (int64_t)n, nullptr)); check(gpv_memcpy_d2h(child.data(), h.d_child, (int64_t)n * 8, nullptr));
		check(gpv_stream_sync(nullptr));
		voxelData.numLevels = h.num_levels;
		voxelData.bBoxHierarchy.resize(n);
		for (size_t i = 0; i < n; i++) {
			BBoxData& b = voxelData.bBoxHierarchy[i];
			for (int a = 0; a < 3; a++) { b.midPoint[a] = mid[i * 3 + a]; b.halfSize[a] = half[i * 3 + a]; }
			b.solid = solid[i]; b.childIndex1 = child[2 * i]; b.childIndex2 = child[2 * i + 1]; b.index = (int)i;
		}
	}

	void SaveVoxelization(const GLParameters*, const char* dir = ".")
	{
		if (!voxelInit) throw Error("SaveVoxelization before PerformVoxelization");
		check(gpv_save(&mesh_, &voxelData.result, &streams_, objID, dir));
	}

private:
	void adopt()
	{
		for (int a = 0; a < 3; a++) { bBoxMin[a] = mesh_.bbox_min[a]; bBoxMax[a] = mesh_.bbox_max[a]; }
		maxModelSize = mesh_.max_model_size;
		totalNumTriangles = (int)mesh_.n_tri;
		flatCPUTriangleData = mesh_.tris;
	}
	gpv_mesh mesh_{};
	gpv_ctx* ctx_ = nullptr;
	gpv_host_streams streams_{};
};

} // namespace gpview
