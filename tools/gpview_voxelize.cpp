// tools/gpview_voxelize.cpp -- headless replacement of "run GPView, press t" (src/GPView.cpp:1598-1695, :1374-1381) for the
// voxelizer path: load every mesh named on the command line (type by the last three characters, like main()), voxelize, write
// the six ObjN*.{txt,raw} files.  Build: make -C tools.  No GLEW / freeglut.
//
//   gpview_voxelize [--l1 N] [--l2 M] [--no-level2] [--no-normals] [--tolerant] [--obj-id K] [--out DIR] [--device D] [--collision] [--hierarchy] mesh.obj|mesh.off ...
//   --tolerant: polygons, free-form blanks, comments, relative indices (GPV_LOAD_TOLERANT, an extension to the reference's readers)
//   --collision: also Object::CollisionInitCUDA (occupied boxes); --hierarchy: also Object::BuildHierarchy (power-of-two grids only)
#include "../include/gpview_b200.hpp"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

int main(int argc, char** argv)
{
	gpview::GLParameters gp;
	gp.saveVoxels = false;
	const char* out = ".";
	bool tolerant = false, collision = false, hierarchy = false;
	int objID = -1; // GPView numbers the first OBJ -1, the second 0, ... (dlID - 3, src/GPView.cpp:181)
	std::vector<const char*> files;
	for (int i = 1; i < argc; i++) {
		if (!strcmp(argv[i], "--l1") && i + 1 < argc) gp.voxelCount = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--l2") && i + 1 < argc) gp.voxelCount2 = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--no-level2")) gp.level2Voxels = false;
		else if (!strcmp(argv[i], "--no-normals")) gp.normals = false;
		else if (!strcmp(argv[i], "--tolerant")) tolerant = true;
		else if (!strcmp(argv[i], "--collision")) collision = true;
		else if (!strcmp(argv[i], "--hierarchy")) { hierarchy = true; gp.collision = true; }
		else if (!strcmp(argv[i], "--obj-id") && i + 1 < argc) objID = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
		else if (!strcmp(argv[i], "--device") && i + 1 < argc) gp.device = atoi(argv[++i]);
		else files.push_back(argv[i]);
	}
	if (files.empty()) { fprintf(stderr, "usage: %s [--l1 N] [--l2 M] [--no-level2] [--no-normals] [--tolerant] [--obj-id K] [--out DIR] mesh.obj|mesh.off ...\n", argv[0]); return 2; }
	try {
		for (const char* f : files) {
			gpview::Object o;
			o.objID = objID++;
			auto t0 = std::chrono::steady_clock::now();
			o.ReadMesh(f, tolerant);
			o.CreateFlatTriangleData();
			auto t1 = std::chrono::steady_clock::now();
			o.PerformVoxelization(&gp);
			auto t2 = std::chrono::steady_clock::now();
			o.SaveVoxelization(&gp, out);
			auto t3 = std::chrono::steady_clock::now();
			const gpview::VoxelData& v = o.voxelData;
			auto s = [](auto a, auto b) { return std::chrono::duration<double>(b - a).count(); };
			// the same report lines as the reference prints (src/Object.cpp:3122-3146, :3415-3423)
			printf("Level 1 Resolution     : %d x %d x %d\n", v.numDivX, v.numDivY, v.numDivZ);
			if (gp.level2Voxels) printf("Level 2 Resolution     : %d x %d x %d\n", v.numDivX2, v.numDivY2, v.numDivZ2);
			printf("Voxels Level 1         : %lld\nInside Voxels          : %lld\nBoundary Voxels        : %lld\n", (long long)v.result.cells,
			       (long long)v.numLevel1InsideVoxels, (long long)v.numLevel1BoundaryVoxels);
			if (gp.level2Voxels)
				printf("Voxels Level2          : %lld\nInside Voxels Level2   : %lld\nBoundary Voxels Level2 : %lld\n", (long long)(v.result.n_boundary * v.result.n23),
				       (long long)v.numLevel2InsideVoxels, (long long)v.numLevel2BoundaryVoxels);
			if (collision) {
				o.CollisionInitCUDA(&gp);
				printf("Collision Boxes        : %zu\n", v.invIndex.size());
			}
			if (hierarchy) {
				o.BuildHierarchy(&gp);
				size_t solid = 0;
				for (const gpview::BBoxData& b : v.bBoxHierarchy) solid += b.solid != 0;
				const gpview::BBoxData& root = v.bBoxHierarchy.back();
				printf("Hierarchy Levels       : %d\nHierarchy Boxes        : %zu\nSolid Boxes            : %zu\nRoot Box               : %.9g %.9g %.9g +- %.9g %.9g %.9g\n", v.numLevels,
				       v.bBoxHierarchy.size(), solid, root.midPoint[0], root.midPoint[1], root.midPoint[2], root.halfSize[0], root.halfSize[1], root.halfSize[2]);
			}
			printf("Load Time              : %g\nVoxelize Time          : %g\nSave Time              : %g\n\n", s(t0, t1), s(t1, t2), s(t2, t3));
		}
	} catch (const gpview::Error& e) {
		fprintf(stderr, "gpview_voxelize: %s\n", e.what());
		return 1;
	}
	return 0;
}
