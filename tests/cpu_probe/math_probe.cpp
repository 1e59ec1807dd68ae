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
	for (int64_t i = 0; i < n; i++) {
		const float *C = c + i * 3, *H = h + i * 3, *T = t + i * 9;
		SatRow s;
		out[i] = sat_row_setup(s, C[1], C[2], H[1], H[2], T[1], T[2], T[4], T[5], T[7], T[8]) && sat_row_test(s, C[0], H[0], H[1], H[2], T[0], T[3], T[6]);
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
