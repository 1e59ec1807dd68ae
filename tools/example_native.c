/* tools/example_native.c -- the native tier from plain C: load a mesh, voxelize it with host buffers,
 * write the six files.  Build:  cc -std=c99 -Iinclude tools/example_native.c -Lgpview_b200 -lgpview_b200 -Wl,-rpath,$PWD/gpview_b200
 * Usage:  example_native mesh.obj|mesh.off [level1=64] [level2=4] [out_dir=.]                                                  */
#include "gpview_b200.h"
#include <stdio.h>
#include <stdlib.h>

static int die(const char* what)
{
	fprintf(stderr, "example_native: %s: %s\n", what, gpv_last_error());
	return 1;
}

int main(int argc, char** argv)
{
	if (argc < 2) { fprintf(stderr, "usage: %s mesh.obj|mesh.off [level1] [level2] [out_dir]\n", argv[0]); return 2; }
	const int l1 = argc > 2 ? atoi(argv[2]) : 64, l2 = argc > 3 ? atoi(argv[3]) : 4;
	const char* out = argc > 4 ? argv[4] : ".";

	gpv_mesh mesh;
	if (gpv_load_mesh(argv[1], &mesh)) return die("load");                 /* Object::ReadObject / ReadOFFObject semantics */
	gpv_grid grid;
	if (gpv_make_grid(mesh.bbox_min, mesh.bbox_max, mesh.max_model_size, l1, l2, &grid)) return die("grid");
	printf("%lld triangles, Level-1 grid %d x %d x %d\n", (long long)mesh.n_tri, grid.num_div[0], grid.num_div[1], grid.num_div[2]);

	gpv_ctx* ctx;
	if (gpv_create(0, &ctx)) return die("create");                         /* fails without an sm_100 GPU: there is no CPU fallback */

	/* host buffers (pinned): the Level-2 sizes are known only after Level 1 -- size them from a Level-1-only call */
	gpv_params p1 = { l1, l2, GPV_NO_LEVEL2, 0, 0 };
	gpv_result r;
	gpv_host_streams none = { 0 };
	if (gpv_voxelize_host(ctx, &mesh, &p1, gpv_stream(ctx), &r, &none)) return die("level 1");
	const int64_t cells = r.cells, nb = r.n_boundary, n23 = (int64_t)l2 * l2 * l2;
	gpv_host_streams h = { 0 };
	h.level1_inout = gpv_alloc_host(cells);
	h.prefix = gpv_alloc_host(cells * 4);
	h.boundary_index = gpv_alloc_host(nb * 4 + 4);
	h.level2_inout = gpv_alloc_host(nb * n23 + 1);
	h.level1_normal = gpv_alloc_host(cells * 3);
	h.level2_normal = gpv_alloc_host(nb * n23 * 3 + 1);
	h.level2_capacity = nb * n23;
	h.boundary_capacity = nb;
	if (!h.level1_inout || !h.prefix || !h.boundary_index || !h.level2_inout || !h.level1_normal || !h.level2_normal) return die("pinned memory");

	gpv_params p = { l1, l2, GPV_NORMALS, 0, 0 };
	if (gpv_voxelize_host(ctx, &mesh, &p, gpv_stream(ctx), &r, &h)) return die("voxelize");
	printf("Level 1: %lld inside, %lld boundary;  Level 2: %lld inside, %lld boundary;  %lld kernel launches\n", (long long)r.l1_inside,
	       (long long)r.l1_boundary, (long long)r.l2_inside, (long long)r.l2_boundary, (long long)r.kernel_launches);
	if (gpv_save(&mesh, &r, &h, /* obj id */ -1, out)) return die("save");  /* Obj-1VoxelConfig.txt, Obj-1Level1InOut.raw, ... */

	gpv_free_host(h.level1_inout); gpv_free_host(h.prefix); gpv_free_host(h.boundary_index);
	gpv_free_host(h.level2_inout); gpv_free_host(h.level1_normal); gpv_free_host(h.level2_normal);
	gpv_destroy(ctx);
	gpv_free_mesh(&mesh);
	return 0;
}
