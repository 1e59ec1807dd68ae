// tests/cpu_probe/math_probe.cpp -- TEST-ONLY: compiles the product's device arithmetic header for the host
// (g++ -O2 -ffp-contract=off) so that the CPU suite can compare it bit-for-bit with the oracle.  Never shipped.
#include "../../gpview_b200/csrc/gpv_math.h"
#include <cstdint>
using namespace gpv;

extern "C" {
void probe_sat_full(int64_t n, const float* c, const float* h, const float* t, uint8_t* out)
{
	for (int64_t i = 0; i < n; i++) {
		const float *C = c + i * 3, *H = h + i * 3, *T = t + i * 9;
		out[i] = tri_box_overlap(C[0], C[1], C[2], H[0], H[1], H[2], T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
	}
}
void probe_sat_row(int64_t n, const float* c, const float* h, const float* t, uint8_t* out)
{
	// the hoisted Level-2 form (z-independent part per column, then the per-sub-voxel part), both ways k_l2 uses it:
	// setup + test, and values (no predicates) + test for pairs that passed the setup
	for (int64_t i = 0; i < n; i++) {
		const float *C = c + i * 3, *H = h + i * 3, *T = t + i * 9;
		SatCol s, v;
		const bool pre = sat_col_setup(s, C[0], C[1], H[0], H[1], T[0], T[1], T[3], T[4], T[6], T[7]);
		sat_col_values(v, C[0], C[1], H[0], H[1], T[0], T[1], T[3], T[4], T[6], T[7]);
		const bool a = pre && sat_col_test(s, C[2], H[0], H[1], H[2], T[2], T[5], T[8]);
		const bool b = pre && sat_col_test(v, C[2], H[0], H[1], H[2], T[2], T[5], T[8]);
		out[i] = a == b ? a : 2;
	}
}
void probe_ray(int64_t n, const float* o, const float* t, uint8_t* out)
{
	for (int64_t i = 0; i < n; i++) {
		const float *O = o + i * 3, *T = t + i * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		out[i] = s.ok && ray_column(s, O[0], O[1], rc) && ray_cell(s, rc, O[2]);
	}
}
// certified candidates: kind and ranges for each triangle
void probe_candidates(int64_t n, const float* t, float minx, float miny, float gsx, float gsy, int nx, int ny, int32_t* out5)
{
	for (int64_t i = 0; i < n; i++) {
		const float* T = t + i * 9;
		RayTri s;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
		int k = fill_candidates(s, minx, miny, gsx, gsy, nx, ny, i0, i1, j0, j1);
		out5[i * 5] = k; out5[i * 5 + 1] = i0; out5[i * 5 + 2] = i1; out5[i * 5 + 3] = j0; out5[i * 5 + 4] = j1;
	}
}
// does the +Z ray through (ox,oy) pass the column part for triangle t?
void probe_column(int64_t n, const float* oxy, const float* t, uint8_t* out)
{
	for (int64_t i = 0; i < n; i++) {
		const float* T = t + i * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		out[i] = s.ok && ray_column(s, oxy[i * 2], oxy[i * 2 + 1], rc);
	}
}
int probe_cell_of(float v, float mn, float mx, int n) { return cell_of(v, mn, mx, n); }
int probe_encode_normal(float x) { return encode_normal(x); }
}

// ---- certified plane culling: for one Level-1 cell (centre mid, size gs, n2 sub-voxels per axis) and one triangle, walk
// every row (q,r): returns the number of sub-voxels whose PLANE predicate (reference arithmetic) passes outside the
// certified interval (must be 0), and counts how many sub-voxels the interval keeps / how many pass the plane / full SAT.
static bool plane_pred_ref(float cx, float cy, float cz, float hx, float hy, float hz, const float* T)
{
	float v0x = T[0] - cx, v0y = T[1] - cy, v0z = T[2] - cz, v1x = T[3] - cx, v1y = T[4] - cy, v1z = T[5] - cz, v2x = T[6] - cx, v2y = T[7] - cy, v2z = T[8] - cz;
	float e0x = v1x - v0x, e0y = v1y - v0y, e0z = v1z - v0z, e1x = v2x - v1x, e1y = v2y - v1y, e1z = v2z - v1z;
	float nx = e0y * e1z - e0z * e1y, ny = e0z * e1x - e0x * e1z, nz = e0x * e1y - e0y * e1x;
	float mnx = (nx > 0.0f) ? (-hx - v0x) : (hx - v0x), mxx = (nx > 0.0f) ? (hx - v0x) : (-hx - v0x);
	float mny = (ny > 0.0f) ? (-hy - v0y) : (hy - v0y), mxy = (ny > 0.0f) ? (hy - v0y) : (-hy - v0y);
	float mnz = (nz > 0.0f) ? (-hz - v0z) : (hz - v0z), mxz = (nz > 0.0f) ? (hz - v0z) : (-hz - v0z);
	if (nx * mnx + ny * mny + nz * mnz > 0.0f) return false;
	return nx * mxx + ny * mxy + nz * mxz >= 0.0f;
}
extern "C" void probe_plane_cull(int64_t n, const float* mid3, const float* gs3, int n2, const float* tri9, int64_t* out4)
{
	// k_l2's culling of one sub-voxel column (p,q): plane interval along z (plane record on cyclically permuted axes), then the
	// certified z-AABB clip.  violations = sub-voxels outside the plane interval whose plane predicate passes, or outside the
	// clipped interval whose full SAT passes.
	int64_t violations = 0, kept = 0, planePass = 0, satPass = 0;
	for (int64_t i = 0; i < n; i++) {
		const float *mid = mid3 + i * 3, *gs = gs3 + i * 3, *T = tri9 + i * 9;
		float h1[3], h2[3], c[3][32];
		for (int a = 0; a < 3; a++) {
			h1[a] = gs[a] / 2.0; float g2 = gs[a] / (n2 * 1.0); h2[a] = g2 / 2.0;
			for (int p = 0; p < n2; p++) c[a][p] = (float)(2 * p + 1) * h2[a] + mid[a] - h1[a];
		}
		PlaneRec pl = plane_rec_setup(T[2], T[0], T[1], T[5], T[3], T[4], T[8], T[6], T[7], gs[2], gs[0], gs[1], h2[2], h2[0], h2[1]);
		float inv2h = 1.f / (2.f * h2[2]);
		float slack = 9.5367431640625e-07f * (fabsf(c[2][0]) + 2.f * gs[2]); // as in k_l2
		const float zmin = fminf(T[2], fminf(T[5], T[8])), zmax = fmaxf(T[2], fmaxf(T[5], T[8]));
		for (int q = 0; q < n2; q++) for (int p = 0; p < n2; p++) {
			int rlo, rhi;
			bool any = plane_row_interval(pl, T[2] - c[2][0], T[0] - c[0][p], T[1] - c[1][q], inv2h, slack, n2, rlo, rhi);
			if (!any) { rlo = 0; rhi = -1; }
			int clo = rlo, chi = rhi;
			axis_clip(zmin, zmax, c[2][0], h2[2], gs[2], inv2h, slack, n2, clo, chi);
			for (int r = 0; r < n2; r++) {
				bool pp = plane_pred_ref(c[0][p], c[1][q], c[2][r], h2[0], h2[1], h2[2], T);
				bool in = r >= rlo && r <= rhi, inc = r >= clo && r <= chi;
				kept += inc; planePass += pp;
				if (pp && !in) violations++;
				bool sat = tri_box_overlap(c[0][p], c[1][q], c[2][r], h2[0], h2[1], h2[2], T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
				if (sat && !inc) violations++;
				if (sat) satPass++;
			}
		}
	}
	out4[0] = violations; out4[1] = kept; out4[2] = planePass; out4[3] = satPass;
}

// ---- certified z-runs: for each (triangle, origin xy, run of nz centres z0 + k*dz) compare gpv::ray_z_run with the per-cell truth.
// out3: [0] violations, [1] runs decided without per-cell work, [2] runs that passed the column test
extern "C" void probe_z_run(int64_t n, const float* oxy, const float* zrun /* z0, dz per item */, int nz, const float* tri9, int64_t* out3)
{
	int64_t bad = 0, decided = 0, passed = 0;
	for (int64_t i = 0; i < n; i++) {
		const float* T = tri9 + i * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		if (!s.ok || !ray_column(s, oxy[i * 2], oxy[i * 2 + 1], rc)) continue;
		passed++;
		float zs[64];
		for (int k = 0; k < nz; k++) zs[k] = zrun[i * 2] + (float)k * zrun[i * 2 + 1];
		int cls = ray_z_run(s, rc, s.well, zs[0], zs[nz - 1]);
		int hits = 0;
		for (int k = 0; k < nz; k++) hits += ray_cell(s, rc, zs[k]);
		if (cls == 0 && hits != 0) bad++;
		if (cls == 1 && hits != nz) bad++;
		if (cls != 2) decided++;
	}
	out3[0] = bad; out3[1] = decided; out3[2] = passed;
}

// ---- ray_cell_mask: the bracketed evaluation of one cell's sub-voxel column must equal n2 independent ray_cell calls.
// cellz: midz, gsz per item (h1 = gs/2, h2 = gs/n2/2 like gpv_make_grid).  out3: [0] mismatches, [1] pairs that passed the
// column test, [2] exact evaluations skipped
extern "C" void probe_cell_mask(int64_t n, const float* oxy, const float* cellz, int n2, const float* tri9, int64_t* out3)
{
	int64_t bad = 0, passed = 0, mixed = 0;
	for (int64_t i = 0; i < n; i++) {
		const float* T = tri9 + i * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		if (!s.ok || !ray_column(s, oxy[i * 2], oxy[i * 2 + 1], rc)) continue;
		passed++;
		const float mid = cellz[i * 2], gs = cellz[i * 2 + 1];
		const float h1 = gs / 2.0; const float g2 = gs / (n2 * 1.0); const float h2 = g2 / 2.0;
		unsigned want = 0;
		for (int r = 0; r < n2; r++) want |= (unsigned)ray_cell(s, rc, l2_centre(r, h2, mid, h1)) << r;
		// the column bound may be taken over any height range that contains the cell: tight, and as wide as a 1000-cell column
		const RayColZ za = ray_col_bound(s, rc, mid - gs, mid + gs, gs, 1.f / (2.02f * h2), 1.f / (1.98f * h2)), zb = ray_col_bound(s, rc, mid - 700.f * gs, mid + 300.f * gs, gs, 1.f / (2.02f * h2), 1.f / (1.98f * h2));
		RayColZ zn = za; zn.k1 = -1.f;
		if (ray_cell_mask(s, rc, za, mid, h1, h2, n2) != want || ray_cell_mask(s, rc, zb, mid, h1, h2, n2) != want || ray_cell_mask(s, rc, zn, mid, h1, h2, n2) != want) bad++;
		const unsigned full = n2 >= 32 ? 0xffffffffu : ((1u << n2) - 1u);
		if (want != 0 && want != full) mixed++;
	}
	out3[0] = bad; out3[1] = passed; out3[2] = mixed;
}

// ---- ray_col_thresholds: for cells at many heights of a grid column [zMin, zMax] the threshold classification (all hit below
// zAll, no hit from zNone on) must agree with n2 independent ray_cell evaluations.  colz: zMin, gs (cell size), ncell per item.
// out4: [0] mismatches, [1] passing pairs, [2] cells classified by a threshold, [3] cells checked
extern "C" void probe_col_thresholds(int64_t n, const float* oxy, const float* colz, int n2, const float* tri9, int64_t* out4)
{
	int64_t bad = 0, passed = 0, byThr = 0, cells = 0;
	for (int64_t i = 0; i < n; i++) {
		const float* T = tri9 + i * 9;
		RayTri s; RayCol rc;
		ray_tri_setup(s, T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
		if (!s.ok || !ray_column(s, oxy[i * 2], oxy[i * 2 + 1], rc)) continue;
		passed++;
		const float zmin0 = colz[i * 3], gs = colz[i * 3 + 1];
		const int ncell = (int)colz[i * 3 + 2];
		const float h1 = gs / 2.0; const float g2 = gs / (n2 * 1.0); const float h2 = g2 / 2.0;
		// centre table like k_prepare: fl32((k + 0.5) * h1 * 2 + min) in double
		const float zMin = (float)((0 + 0.5) * (double)h1 * 2 + (double)zmin0) - gs, zMax = (float)((ncell - 1 + 0.5) * (double)h1 * 2 + (double)zmin0) + gs;
		const RayColZ z = ray_col_bound(s, rc, zMin, zMax, gs, 1.f / (2.02f * h2), 1.f / (1.98f * h2));
		const RayThr thr = ray_col_thresholds(s, rc, z, zMin, zMax);
		const unsigned full = n2 >= 32 ? 0xffffffffu : ((1u << n2) - 1u);
		for (int k = 0; k < ncell; k++) {
			const float mid = (float)((k + 0.5) * (double)h1 * 2 + (double)zmin0);
			unsigned want = 0;
			for (int r = 0; r < n2; r++) want |= (unsigned)ray_cell(s, rc, l2_centre(r, h2, mid, h1)) << r;
			const float zLoC = l2_centre(0, h2, mid, h1), zHiC = l2_centre(n2 - 1, h2, mid, h1);
			cells++;
			if (zHiC < thr.zAll) { byThr++; if (want != full) bad++; }
			else if (zLoC >= thr.zNone) { byThr++; if (want != 0u) bad++; }
			else if (ray_cell_mask(s, rc, z, mid, h1, h2, n2) != want) bad++;
		}
	}
	out4[0] = bad; out4[1] = passed; out4[2] = byThr; out4[3] = cells;
}
