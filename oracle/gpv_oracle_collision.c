/* oracle/gpv_oracle_collision.c -- TEST INFRASTRUCTURE ONLY (see gpv_oracle.h): plain-C restatement of the reference's voxel
 * hierarchy / collision structures over the Level-1 grid:
 *   Object::CollisionInitCUDA   src/Object.cpp:3530-3572   occupied cells (state >= 1) -> inverse index + box centre / extent arrays
 *   CombineBBox                 src/Object.cpp:2750-2788   union of two boxes, `solid` flags
 *   Object::BuildHierarchy      src/Object.cpp:2790-2867   binary AABB hierarchy: pairs along x, then halving x / y / z in rotation
 * Leaf boxes are the bBox[] of PerformVoxelization (:3165-3193): midPoint = fl32((i + 0.5) * gridSize + bBoxMin) (double arithmetic),
 * halfSize = fl32(gridSize / 2.0), solid = int(level1InOut) % 2 taken BEFORE the SAT pass (the parity fill, also for boundary cells).
 * The reference never initialises the leaves' `index`, so the child indices of its first level are indeterminate; here a leaf's index
 * is its linear cell index.  BuildHierarchy is only well defined when every grid dimension is a power of two (it indexes out of
 * bounds otherwise -- GetNextDiv4 grids generally are not): gpvo_build_hierarchy returns -1 for other grids.
 * Pinned against the live reference by tests/test_oracle_ref.py. */
#include "gpv_oracle.h"
#include <stdlib.h>
#include <string.h>

int64_t gpvo_collision_boxes(const gpvo_mesh* m, const gpvo_result* r, int32_t* invIndex, float* mid, float* ext)
{
	const gpvo_grid* g = &r->g;
	const int nx = g->numDiv[0], ny = g->numDiv[1], nz = g->numDiv[2];
	float* cx = (float*)malloc(sizeof(float) * (size_t)(nx + ny + nz));
	float *cy = cx + nx, *cz = cy + ny;
	gpvo_axis_table(m->bmin, g, 0, cx); gpvo_axis_table(m->bmin, g, 1, cy); gpvo_axis_table(m->bmin, g, 2, cz);
	int64_t n = 0;
	for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
		const int64_t idx = ((int64_t)k * ny + j) * nx + i;
		if (r->l1State[idx] >= 1) {                                    /* :3537 level1InOut[i] >= 1 */
			invIndex[n] = (int32_t)idx;
			mid[n * 3] = cx[i]; mid[n * 3 + 1] = cy[j]; mid[n * 3 + 2] = cz[k];   /* bBox[].midPoint, :3165-3193 */
			ext[n * 3] = g->ext1[0]; ext[n * 3 + 1] = g->ext1[1]; ext[n * 3 + 2] = g->ext1[2];
			n++;
		}
	}
	free(cx);
	return n;
}

typedef struct { float mid[3], half[3]; int solid, c1, c2, index; } Box;

static void combine(const Box* b1, const Box* b2, Box* out)   /* CombineBBox, :2750-2788 */
{
	float midVal[3] = { 0, 0, 0 }, halfVal[3] = { 0, 0, 0 };
	if (b1->solid == 0 && b2->solid == 0) out->solid = 0;
	else if (b1->solid == 0) { out->solid = 1; memcpy(midVal, b2->mid, sizeof midVal); memcpy(halfVal, b2->half, sizeof halfVal); }
	else if (b2->solid == 0) { out->solid = 1; memcpy(midVal, b1->mid, sizeof midVal); memcpy(halfVal, b1->half, sizeof halfVal); }
	else {
		out->solid = 1;
		for (int a = 0; a < 3; a++) {
			const float lo1 = b1->mid[a] - b1->half[a], hi1 = b1->mid[a] + b1->half[a];
			const float lo2 = b2->mid[a] - b2->half[a], hi2 = b2->mid[a] + b2->half[a];
			const float mx = hi1 > hi2 ? hi1 : hi2, mn = lo1 < lo2 ? lo1 : lo2;   /* std::max / std::min */
			midVal[a] = (mx + mn) / 2.0f;                                          /* Float3 / float */
			halfVal[a] = (mx - mn) / 2.0f;
		}
	}
	memcpy(out->mid, midVal, sizeof midVal); memcpy(out->half, halfVal, sizeof halfVal);
	out->c1 = b1->index; out->c2 = b2->index;
}

static int is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

/* outputs sized cells-1: mid / half 3 floats each, solid 1 byte, child 2 ints; returns the number of levels, -1 for a grid the
 * reference's loop is not defined on */
int gpvo_build_hierarchy(const gpvo_mesh* m, const gpvo_result* r, float* mid, float* half, uint8_t* solid, int32_t* child)
{
	const gpvo_grid* g = &r->g;
	int nx = g->numDiv[0], ny = g->numDiv[1], nz = g->numDiv[2];
	if (!is_pow2(nx) || !is_pow2(ny) || !is_pow2(nz) || (int64_t)nx * ny * nz < 2) return -1;
	const int total = nx * ny * nz;
	int numLevels = 0;
	{ float y = (float)total; while (y > 1) { y /= 2; numLevels++; } }     /* GetExponent2, src/Utilities.cpp:349 */
	float* cx = (float*)malloc(sizeof(float) * (size_t)(nx + ny + nz));
	float *cy = cx + nx, *cz = cy + ny;
	gpvo_axis_table(m->bmin, g, 0, cx); gpvo_axis_table(m->bmin, g, 1, cy); gpvo_axis_table(m->bmin, g, 2, cz);
	Box* leaf = (Box*)calloc((size_t)total, sizeof(Box));
	Box* h = (Box*)calloc((size_t)total - 1, sizeof(Box));
	for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
		Box* b = leaf + ((size_t)k * ny + j) * nx + i;
		b->mid[0] = cx[i]; b->mid[1] = cy[j]; b->mid[2] = cz[k];
		b->half[0] = g->ext1[0]; b->half[1] = g->ext1[1]; b->half[2] = g->ext1[2];
		b->solid = r->l1FillOnly[((size_t)k * ny + j) * nx + i] % 2;
		b->index = (int)(((size_t)k * ny + j) * nx + i);
	}
	int numLevelBoxes = total / 2;
	for (int i = 0; i < numLevelBoxes; i++) { combine(leaf + 2 * i, leaf + 2 * i + 1, h + i); h[i].index = i; }   /* :2806-2814 */
	int numDivX = nx / 2, numDivY = ny, numDivZ = nz, prevLevelIndex = 0, levelIndex = numLevelBoxes;
	numLevelBoxes /= 2;
	for (int level = 2; level < numLevels + 1; level++) {                                                            /* :2822-2866 */
		int iSkip = (level % 3 == 1 && numDivX > 1) ? 2 : 1;
		int jSkip = (level % 3 == 2 && numDivY > 1) ? 2 : 1;
		int kSkip = (level % 3 == 0 && numDivZ > 1) ? 2 : 1;
		if (iSkip == 1 && jSkip == 1 && kSkip == 1) {
			if (numDivX > 1) iSkip = 2; else if (numDivY > 1) jSkip = 2; else if (numDivZ > 1) kSkip = 2;
		}
		for (int k = 0; k < numDivZ; k += kSkip) for (int j = 0; j < numDivY; j += jSkip) for (int i = 0; i < numDivX; i += iSkip) {
			const int index1 = k * numDivY * numDivX + j * numDivX + i;
			const int index2 = (k / kSkip) * (numDivY / jSkip) * (numDivX / iSkip) + (j / jSkip) * (numDivX / iSkip) + (i / iSkip);
			const int skip = (kSkip - 1) * numDivY * numDivX + (jSkip - 1) * numDivX + (iSkip - 1);
			Box* out = h + levelIndex + index2;
			combine(h + prevLevelIndex + index1, h + prevLevelIndex + index1 + skip, out);
			out->index = levelIndex + index2;
		}
		if (iSkip == 2) numDivX /= 2;
		if (jSkip == 2) numDivY /= 2;
		if (kSkip == 2) numDivZ /= 2;
		prevLevelIndex += numLevelBoxes * 2;
		levelIndex += numLevelBoxes;
		numLevelBoxes /= 2;
	}
	for (int i = 0; i < total - 1; i++) {
		memcpy(mid + (size_t)i * 3, h[i].mid, 12); memcpy(half + (size_t)i * 3, h[i].half, 12);
		solid[i] = (uint8_t)h[i].solid; child[2 * i] = h[i].c1; child[2 * i + 1] = h[i].c2;
	}
	free(cx); free(leaf); free(h);
	return numLevels;
}
