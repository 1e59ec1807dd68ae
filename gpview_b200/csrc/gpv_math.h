// gpview_b200/csrc/gpv_math.h -- intersection arithmetic of the voxelizer path, shared by every kernel.
//
// Bit-exactness contract (DESIGN.md "Numerics"): every binary32 operation below is a separate IEEE-754 round-to-nearest
// operation in the reference's operand order.  The translation unit that includes this header MUST be compiled with
// FMA contraction disabled (nvcc -fmad=false; g++ -ffp-contract=off for the CPU-side unit probes in tests/).
// The functions are __host__ __device__ only so that tests/ can exercise the very same source on the CPU against the
// oracle; the product never calls them on the host.
//
// Reference arithmetic restated here:
//   SAT  tri/box : TriBoxOverlapCUDA  cuda/CUDAClassifyTessellation.cu:234-309 (== src/TriBoxIntersection.cpp:135-210)
//   ray  tri/+Z  : TriRayIntersectCUDA cuda/CUDAClassifyTessellation.cu:102-155 (== src/TriRayIntersection.cpp:78-131)
//   cell index   : cuda/CUDAClassifyTessellation.cu:333-361
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GPV_HD __host__ __device__ __forceinline__
#else
#define GPV_HD inline
#endif

namespace gpv {

// fl32(1e-6): the reference compares floats against the DOUBLE literal 0.000001 (cu:76).  For a float x:
//   (double)x < 1e-6  <=>  x <= kEps      (double)x > -1e-6  <=>  x >= -kEps      (double)x > 1e-6  <=>  x > kEps
constexpr float kEps = 9.99999997475242707878e-07f;
constexpr float kFltMax = 3.402823466e+38f;

// ---- separating-axis helper: `if (min>rad || max<-rad) return 0;` of every AXISTEST macro (cu:182-231)
GPV_HD bool axis_separates(float pa, float pb, float rad)
{
	float mn = fminf(pa, pb), mx = fmaxf(pa, pb); // pa,pb are never NaN for finite input; min/max == the macro's if/else
	return mn > rad || mx < -rad;
}

// ---- full 13-predicate SAT in the reference's order (used by Level-1 binning, where nothing can be hoisted)
GPV_HD bool tri_box_overlap(float cx, float cy, float cz, float hx, float hy, float hz,
                            float t0x, float t0y, float t0z, float t1x, float t1y, float t1z, float t2x, float t2y, float t2z)
{
	// translate so the box centre is the origin (cu:251-253); edges are formed AFTER the translation (cu:256-258)
	float v0x = t0x - cx, v0y = t0y - cy, v0z = t0z - cz;
	float v1x = t1x - cx, v1y = t1y - cy, v1z = t1z - cz;
	float v2x = t2x - cx, v2y = t2y - cy, v2z = t2z - cz;
	// cheap predicates first (legal: the result is an AND of side-effect-free predicates)
	{
		float mn = fminf(v0x, fminf(v1x, v2x)), mx = fmaxf(v0x, fmaxf(v1x, v2x));
		if (mn > hx || mx < -hx) return false;
		mn = fminf(v0y, fminf(v1y, v2y)); mx = fmaxf(v0y, fmaxf(v1y, v2y));
		if (mn > hy || mx < -hy) return false;
		mn = fminf(v0z, fminf(v1z, v2z)); mx = fmaxf(v0z, fmaxf(v1z, v2z));
		if (mn > hz || mx < -hz) return false;
	}
	float e0x = v1x - v0x, e0y = v1y - v0y, e0z = v1z - v0z;
	float e1x = v2x - v1x, e1y = v2y - v1y, e1z = v2z - v1z;
	// plane of the triangle against the box (cu:304-306, planeBoxOverlapCUDA cu:157-179)
	{
		float nx = e0y * e1z - e0z * e1y, ny = e0z * e1x - e0x * e1z, nz = e0x * e1y - e0y * e1x;
		float mnx = (nx > 0.0f) ? (-hx - v0x) : (hx - v0x), mxx = (nx > 0.0f) ? (hx - v0x) : (-hx - v0x);
		float mny = (ny > 0.0f) ? (-hy - v0y) : (hy - v0y), mxy = (ny > 0.0f) ? (hy - v0y) : (-hy - v0y);
		float mnz = (nz > 0.0f) ? (-hz - v0z) : (hz - v0z), mxz = (nz > 0.0f) ? (hz - v0z) : (-hz - v0z);
		if (nx * mnx + ny * mny + nz * mnz > 0.0f) return false;
		if (!(nx * mxx + ny * mxy + nz * mxz >= 0.0f)) return false;
	}
	float e2x = v0x - v2x, e2y = v0y - v2y, e2z = v0z - v2z;
	float fex, fey, fez;
	// edge 0: X01 Y02 Z12 (cu:262-267)
	fex = fabsf(e0x); fey = fabsf(e0y); fez = fabsf(e0z);
	if (axis_separates(e0z * v0y - e0y * v0z, e0z * v2y - e0y * v2z, fez * hy + fey * hz)) return false;
	if (axis_separates(-e0z * v0x + e0x * v0z, -e0z * v2x + e0x * v2z, fez * hx + fex * hz)) return false;
	if (axis_separates(e0y * v1x - e0x * v1y, e0y * v2x - e0x * v2y, fey * hx + fex * hy)) return false;
	// edge 1: X01 Y02 Z0 (cu:269-274)
	fex = fabsf(e1x); fey = fabsf(e1y); fez = fabsf(e1z);
	if (axis_separates(e1z * v0y - e1y * v0z, e1z * v2y - e1y * v2z, fez * hy + fey * hz)) return false;
	if (axis_separates(-e1z * v0x + e1x * v0z, -e1z * v2x + e1x * v2z, fez * hx + fex * hz)) return false;
	if (axis_separates(e1y * v0x - e1x * v0y, e1y * v1x - e1x * v1y, fey * hx + fex * hy)) return false;
	// edge 2: X2 Y1 Z12 (cu:276-281)
	fex = fabsf(e2x); fey = fabsf(e2y); fez = fabsf(e2z);
	if (axis_separates(e2z * v0y - e2y * v0z, e2z * v1y - e2y * v1z, fez * hy + fey * hz)) return false;
	if (axis_separates(-e2z * v0x + e2x * v0z, -e2z * v1x + e2x * v1z, fez * hx + fex * hz)) return false;
	if (axis_separates(e2y * v1x - e2x * v1y, e2y * v2x - e2x * v2y, fey * hx + fex * hy)) return false;
	return true;
}

// ---- Level-2 SAT, hoisted along z.  A column of sub-voxels shares (cx,cy); everything that touches only x/y components is
// computed once per (column, triangle): the translated x/y coordinates, the x/y edge components, the three Z-axis tests
// (Z12 of edge 0, Z0 of edge 1, Z12 of edge 2, cu:266/273/280), the x/y AABB tests and the z component of the normal.
// Exact: those sub-expressions do not involve cz at all.  (z is the axis of the parity rays too, so the SAT bits and the
// parity bits of a sub-voxel column end up in the same word layout.)
struct SatCol {
	float v0x, v0y, v1x, v1y, v2x, v2y;
	float e0x, e0y, e1x, e1y, e2x, e2y;
	float nz;                 // e0x*e1y - e0y*e1x
	float px0, px1, py0, py1; // plane-box candidates: (-hx - v0x, hx - v0x), (-hy - v0y, hy - v0y)
};

// the column quantities without any predicate
GPV_HD void sat_col_values(SatCol& s, float cx, float cy, float hx, float hy,
                           float t0x, float t0y, float t1x, float t1y, float t2x, float t2y)
{
	s.v0x = t0x - cx; s.v0y = t0y - cy; s.v1x = t1x - cx; s.v1y = t1y - cy; s.v2x = t2x - cx; s.v2y = t2y - cy;
	s.e0x = s.v1x - s.v0x; s.e0y = s.v1y - s.v0y;
	s.e1x = s.v2x - s.v1x; s.e1y = s.v2y - s.v1y;
	s.e2x = s.v0x - s.v2x; s.e2y = s.v0y - s.v2y;
	s.nz = s.e0x * s.e1y - s.e0y * s.e1x;
	s.px0 = -hx - s.v0x; s.px1 = hx - s.v0x; s.py0 = -hy - s.v0y; s.py1 = hy - s.v0y;
}

// returns false when a z-independent predicate already separates the triangle from every box of the column
GPV_HD bool sat_col_setup(SatCol& s, float cx, float cy, float hx, float hy,
                          float t0x, float t0y, float t1x, float t1y, float t2x, float t2y)
{
	sat_col_values(s, cx, cy, hx, hy, t0x, t0y, t1x, t1y, t2x, t2y);
	{
		float mn = fminf(s.v0x, fminf(s.v1x, s.v2x)), mx = fmaxf(s.v0x, fmaxf(s.v1x, s.v2x));
		if (mn > hx || mx < -hx) return false;
		mn = fminf(s.v0y, fminf(s.v1y, s.v2y)); mx = fmaxf(s.v0y, fmaxf(s.v1y, s.v2y));
		if (mn > hy || mx < -hy) return false;
	}
	// Z12(e0), Z0(e1), Z12(e2): a = e.y, b = e.x, rad = |e.y|*hx + |e.x|*hy
	if (axis_separates(s.e0y * s.v1x - s.e0x * s.v1y, s.e0y * s.v2x - s.e0x * s.v2y, fabsf(s.e0y) * hx + fabsf(s.e0x) * hy)) return false;
	if (axis_separates(s.e1y * s.v0x - s.e1x * s.v0y, s.e1y * s.v1x - s.e1x * s.v1y, fabsf(s.e1y) * hx + fabsf(s.e1x) * hy)) return false;
	if (axis_separates(s.e2y * s.v1x - s.e2x * s.v1y, s.e2y * s.v2x - s.e2x * s.v2y, fabsf(s.e2y) * hx + fabsf(s.e2x) * hy)) return false;
	return true;
}

// Certified clip of an index interval [lo,hi] along one axis to the sub-voxels whose AABB predicate on that axis (cu:284-298)
// can pass: fl(tmin - c_k) <= h and fl(tmax - c_k) >= -h need c_k within [tmin - h, tmax + h] up to rounding.  c0 = centre of
// the first sub-voxel of the line, |c_k - c0 - 2*h*k| <= slack (see plane_row_interval); the 128u margin covers that, the
// rounding of the predicate's own subtraction and of the bounds below.  Never drops a sub-voxel that passes; the exact test still runs.
GPV_HD void axis_clip(float tmin, float tmax, float c0, float h, float gs, float inv2h, float slack, int n2, int& lo, int& hi)
{
	const float s2 = slack + 7.62939453125e-06f * (fabsf(tmin) + fabsf(tmax) + fabsf(c0) + gs); // 2^-17 = 128u
	const float a = ceilf((tmin - c0 - h - s2) * inv2h), b = floorf((tmax - c0 + h + s2) * inv2h);
	if (a == a && a > (float)lo) lo = a > (float)n2 ? n2 : (int)a;
	if (b == b && b < (float)hi) hi = b < -1.f ? -1 : (int)b;
}

// the remaining predicates for one box of the column (box centre z = cz)
GPV_HD bool sat_col_test(const SatCol& s, float cz, float hx, float hy, float hz, float t0z, float t1z, float t2z)
{
	float v0z = t0z - cz, v1z = t1z - cz, v2z = t2z - cz;
	{
		float mn = fminf(v0z, fminf(v1z, v2z)), mx = fmaxf(v0z, fmaxf(v1z, v2z));
		if (mn > hz || mx < -hz) return false;
	}
	float e0z = v1z - v0z, e1z = v2z - v1z;
	{
		float nx = s.e0y * e1z - e0z * s.e1y, ny = e0z * s.e1x - s.e0x * e1z;
		float mnx = (nx > 0.0f) ? s.px0 : s.px1, mxx = (nx > 0.0f) ? s.px1 : s.px0;
		float mny = (ny > 0.0f) ? s.py0 : s.py1, mxy = (ny > 0.0f) ? s.py1 : s.py0;
		float mnz = (s.nz > 0.0f) ? (-hz - v0z) : (hz - v0z), mxz = (s.nz > 0.0f) ? (hz - v0z) : (-hz - v0z);
		if (nx * mnx + ny * mny + s.nz * mnz > 0.0f) return false;
		if (!(nx * mxx + ny * mxy + s.nz * mxz >= 0.0f)) return false;
	}
	float e2z = v0z - v2z;
	float fez;
	// edge 0: X01, Y02
	fez = fabsf(e0z);
	if (axis_separates(e0z * s.v0y - s.e0y * v0z, e0z * s.v2y - s.e0y * v2z, fez * hy + fabsf(s.e0y) * hz)) return false;
	if (axis_separates(-e0z * s.v0x + s.e0x * v0z, -e0z * s.v2x + s.e0x * v2z, fez * hx + fabsf(s.e0x) * hz)) return false;
	// edge 1: X01, Y02
	fez = fabsf(e1z);
	if (axis_separates(e1z * s.v0y - s.e1y * v0z, e1z * s.v2y - s.e1y * v2z, fez * hy + fabsf(s.e1y) * hz)) return false;
	if (axis_separates(-e1z * s.v0x + s.e1x * v0z, -e1z * s.v2x + s.e1x * v2z, fez * hx + fabsf(s.e1x) * hz)) return false;
	// edge 2: X2, Y1
	fez = fabsf(e2z);
	if (axis_separates(e2z * s.v0y - s.e2y * v0z, e2z * s.v1y - s.e2y * v1z, fez * hy + fabsf(s.e2y) * hz)) return false;
	if (axis_separates(-e2z * s.v0x + s.e2x * v0z, -e2z * s.v1x + s.e2x * v1z, fez * hx + fabsf(s.e2x) * hz)) return false;
	return true;
}

// ---- certified plane culling for the Level-2 SAT (DESIGN.md "Certified plane culling").
// The plane predicate of the SAT (planeBoxOverlapCUDA, cu:157-179) passes only if |N.(t0 - c)| <= r with N = e0 x e1 and
// r = sum |N_q| h_q -- in exact arithmetic.  In the reference's f32 arithmetic the two dot products it compares with 0 differ
// from -r -/+ N.(t0-c) by at most 488 u M^3 (u = 2^-24, M = bound on every translated coordinate of the triangle seen from
// any sub-voxel centre of a Level-1 cell the triangle overlaps); our own f32 evaluation of the left side adds < 500 u M^3.
// With E = 2048 u M^3 = 2^-13 M^3:   |N.(t0 - c)| > r + E   ==>   the reference's plane predicate FAILS.
// (The product calls these two functions with the coordinates permuted cyclically, (x,y,z) <- (z,x,y): a cyclic relabelling maps
// e0 x e1 onto itself, so "x" below is the z axis of the grid and the interval runs along a sub-voxel COLUMN.)
// Along a row of sub-voxels (fixed cy, cz) N.(t0 - c_p) = D0 - Nx * s_p with s_p = c_p - c_0, so the sub-voxels that can
// pass form an index interval.  The record is normalised by Nx so that the interval needs no division per row:
//   x = 1: (1, Ny/Nx, Nz/Nx, (r+E)/|Nx|)      x = 0: Nx too small to normalise, (0, Ny, Nz, r+E): whole row or nothing
//   w = +inf: no culling (degenerate magnitudes)
struct PlaneRec { float sx, ny, nz, R; };

GPV_HD PlaneRec plane_rec_setup(float t0x, float t0y, float t0z, float t1x, float t1y, float t1z, float t2x, float t2y, float t2z,
                                float gsx, float gsy, float gsz, float hx, float hy, float hz)
{
	PlaneRec p;
	float e0x = t1x - t0x, e0y = t1y - t0y, e0z = t1z - t0z, e1x = t2x - t1x, e1y = t2y - t1y, e1z = t2z - t1z;
	float Nx = e0y * e1z - e0z * e1y, Ny = e0z * e1x - e0x * e1z, Nz = e0x * e1y - e0y * e1x;
	float ex = fmaxf(t0x, fmaxf(t1x, t2x)) - fminf(t0x, fminf(t1x, t2x));
	float ey = fmaxf(t0y, fmaxf(t1y, t2y)) - fminf(t0y, fminf(t1y, t2y));
	float ez = fmaxf(t0z, fmaxf(t1z, t2z)) - fminf(t0z, fminf(t1z, t2z));
	float amax = fmaxf(fmaxf(fabsf(t0x), fmaxf(fabsf(t1x), fabsf(t2x))), fmaxf(fmaxf(fabsf(t0y), fmaxf(fabsf(t1y), fabsf(t2y))),
	                                                                            fmaxf(fabsf(t0z), fmaxf(fabsf(t1z), fabsf(t2z)))));
	float gmax = fmaxf(gsx, fmaxf(gsy, gsz));
	float M = 1.01f * fmaxf(ex + gsx, fmaxf(ey + gsy, ez + gsz)) + 1e-6f * (amax + gmax);
	float E = M * M * M * 1.220703125e-4f; // 2^-13
	float R = fabsf(Nx) * hx + fabsf(Ny) * hy + fabsf(Nz) * hz + E;
	p.sx = 0.f; p.ny = Ny; p.nz = Nz; p.R = R;
	if (!(M > 1e-9f && M < 1e9f) || !(R <= kFltMax)) { p.R = INFINITY; return p; }
	float inv = 1.f / Nx;
	float ny = Ny * inv, nz = Nz * inv, Rn = R * fabsf(inv) * 1.000001f;
	if (fabsf(Nx) > 1e-30f && fabsf(inv) <= kFltMax && fabsf(ny) <= kFltMax && fabsf(nz) <= kFltMax && Rn <= kFltMax) { p.sx = 1.f; p.ny = ny; p.nz = nz; p.R = Rn; }
	return p;
}

// Index interval [plo,phi] of the sub-voxels of one row that can pass the plane predicate; false = none can.
//   v0x0 = t0x - c_x[0], v0y = t0y - cy, v0z = t0z - cz (the translated first vertex at p = 0);
//   inv2h = 1/(2*h2x); slack >= 6u(|mid_x| + gs_x) bounds |(c_p - c_0) - 2*h2x*p| (f32 rounding of the centre formula).
GPV_HD bool plane_row_interval(const PlaneRec& pl, float v0x0, float v0y, float v0z, float inv2h, float slack, int n2, int& plo, int& phi)
{
	plo = 0; phi = n2 - 1;
	if (!(pl.R <= kFltMax)) return true;
	float D = pl.sx * v0x0 + pl.ny * v0y + pl.nz * v0z;
	if (pl.sx == 0.f) return !(fabsf(D) > pl.R);
	// index k can pass only if lo <= k <= hi as real numbers; the f32 error of the two bounds (<= 3u(|D|+R), i.e. below 1e-5 of an
	// index for any cell the triangle overlaps) is inside the E budget folded into R and again inside the 2^-7 widening
	const float xl = (D - pl.R - slack) * inv2h, xh = (D + pl.R + slack) * inv2h;
	float lo = ceilf(xl - (0.0078125f + 1e-6f * fabsf(xl))), hi = floorf(xh + (0.0078125f + 1e-6f * fabsf(xh)));
	if (!(lo == lo) || !(hi == hi)) return true;
	if (lo > 0.f) plo = lo > (float)n2 ? n2 : (int)lo;
	if (hi < (float)(n2 - 1)) phi = hi < -1.f ? -1 : (int)hi;
	return plo <= phi;
}

// ---- Moller-Trumbore for the fixed direction D = (0,0,1), split by what each part depends on.
// Exact for finite intermediates: the dropped terms are products with D's zero components, i.e. additions of +-0.
struct RayTri { // per triangle
	float v1x, v1y, v1z, e1x, e1y, e1z, e2x, e2y, e2z, det, inv;
	bool ok;   // false: |det| inside the epsilon band, or not finite -- the ray test can never return 1
	bool well; // det is well conditioned: whole z-runs can be decided by gpv::ray_z_run
};
GPV_HD void ray_tri_setup(RayTri& s, float t0x, float t0y, float t0z, float t1x, float t1y, float t1z, float t2x, float t2y, float t2z)
{
	s.v1x = t0x; s.v1y = t0y; s.v1z = t0z;
	s.e1x = t1x - t0x; s.e1y = t1y - t0y; s.e1z = t1z - t0z;
	s.e2x = t2x - t0x; s.e2y = t2y - t0y; s.e2z = t2z - t0z;
	s.det = s.e1x * (-s.e2y) + s.e1y * s.e2x; // P = D x e2 = (-e2y, e2x, 0)
	float ad = fabsf(s.det);
	s.ok = ad > kEps && ad <= kFltMax;
	{
		float S = fabsf(s.e1x * s.e2y) + fabsf(s.e1y * s.e2x);
		s.well = s.ok && ad >= 9.765625e-4f * S && S <= kFltMax; // 2^-10
	}
	s.inv = 1.f / s.det;
}
struct RayCol { float c0, c1, c2; }; // per (triangle, xy origin)
GPV_HD bool ray_column(const RayTri& s, float ox, float oy, RayCol& c)
{
	float Tx = ox - s.v1x, Ty = oy - s.v1y;
	float u = (Tx * (-s.e2y) + Ty * s.e2x) * s.inv;
	if (u < 0.f || u > 1.f) return false;
	float Q2 = Tx * s.e1y - Ty * s.e1x;
	float v = Q2 * s.inv;
	if (v < 0.f || u + v > 1.f) return false;
	c.c0 = Ty * s.e1z; c.c1 = Tx * s.e1z; c.c2 = s.e2z * Q2;
	return true;
}
GPV_HD float ray_cell_t(const RayTri& s, const RayCol& c, float oz)
{
	float Tz = oz - s.v1z;
	float Q0 = c.c0 - Tz * s.e1y;
	float Q1 = Tz * s.e1x - c.c1;
	return (s.e2x * Q0 + s.e2y * Q1 + c.c2) * s.inv;
}
GPV_HD bool ray_cell(const RayTri& s, const RayCol& c, float oz) { return ray_cell_t(s, c, oz) > kEps; }

// ---- certified classification of a whole run of cells along one column (DESIGN.md "Certified z-runs").
// For a (triangle, column) pair that passed ray_column, t as a function of the origin height is  t = alpha - Tz * (D/det') in
// exact arithmetic on the f32 inputs, D the exact 2-D determinant.  When det' is well conditioned (|det'| >= 2^-10 * S,
// S = |e1x*e2y| + |e1y*e2x|, flagged per triangle) the slope lies in [0.998, 1.002], so t falls monotonically with height.
// ray_cell's f32 result differs from that exact t by at most bnd (16u times the magnitudes of its intermediates; the
// actual rounding count is <= 5u per term).  Evaluating ray_cell ONCE at the lowest centre therefore decides the whole run:
//   t0 + 2 bnd <= eps            -> no cell of the run is hit
//   t0 - 2 bnd - 1.01 span > eps -> every cell of the run is hit
// returns 0 / 1 for those, 2 when the cells must be evaluated one by one.
GPV_HD int ray_z_run(const RayTri& s, const RayCol& c, bool wellConditioned, float ozLo, float ozHi)
{
	if (!wellConditioned) return 2;
	const float TzLo = ozLo - s.v1z, TzHi = ozHi - s.v1z;
	const float Tm = fmaxf(fabsf(TzLo), fabsf(TzHi));
	const float Q0 = c.c0 - TzLo * s.e1y, Q1 = TzLo * s.e1x - c.c1;
	const float t0 = (s.e2x * Q0 + s.e2y * Q1 + c.c2) * s.inv; // == what ray_cell computes at ozLo
	const float span = (ozHi - ozLo) * 1.01f;
	const float bnd = 9.5367431640625e-07f * (fabsf(s.inv) * (fabsf(s.e2x) * (fabsf(c.c0) + Tm * fabsf(s.e1y)) + fabsf(s.e2y) * (fabsf(c.c1) + Tm * fabsf(s.e1x)) + fabsf(c.c2)) +
	                                          fabsf(t0) + span) + 1e-30f;
	if (!(bnd <= kFltMax) || !(t0 == t0)) return 2;
	if (t0 + 2.f * bnd <= kEps) return 0;
	if (t0 - 2.f * bnd - span > kEps) return 1;
	return 2;
}

// ---- Level-2 sub-voxel centre along one axis (cu:423-425 / 472-474): ((2p+1)*ext2 + mid) - ext1, all f32
GPV_HD float l2_centre(int p, float h2, float mid, float h1) { return (float)(2 * p + 1) * h2 + mid - h1; }

// ---- parity bits of one (triangle, sub-voxel column) pair over the n2 <= 32 sub-voxels of ONE Level-1 cell: bit r is set iff
// ray_cell(s, c, z_r), z_r = l2_centre(r, h2z, midz, h1z).  Bit-identical to n2 calls of ray_cell, but for a well-conditioned
// triangle only the sub-voxels next to the crossing are evaluated (certificates below).  The part of ray_z_run's error bound
// that does not depend on the cell is computed once per (triangle, sub-column) by ray_col_bound for ALL heights of the grid
// column [zMin, zMax] (a larger |Tz| or span only makes the bound more conservative).
struct RayColZ { float k1, span, inv101, inv099; }; // k1 < 0: no certificate; inv101 = 1/(1.01*2*h2z), inv099 = 1/(0.99*2*h2z): index estimates per unit of t
GPV_HD RayColZ ray_col_bound(const RayTri& s, const RayCol& c, float zMin, float zMax, float gsz, float inv101, float inv099)
{
	RayColZ z;
	// upper bound of 1.01 (z_hi - z_lo) of every cell of the column: the centres span (2 n2 - 2) h2z < gsz, plus their own rounding
	z.span = 1.01f * (gsz + 9.5367431640625e-07f * (fabsf(zMin) + fabsf(zMax)));
	z.inv101 = inv101; z.inv099 = inv099;
	z.k1 = -1.f;
	if (!s.well) return z;
	const float Tm = fmaxf(fabsf(zMin - s.v1z), fabsf(zMax - s.v1z));
	const float k1 = 9.5367431640625e-07f * (fabsf(s.inv) * (fabsf(s.e2x) * (fabsf(c.c0) + Tm * fabsf(s.e1y)) + fabsf(s.e2y) * (fabsf(c.c1) + Tm * fabsf(s.e1x)) + fabsf(c.c2)) + z.span) + 1e-30f;
	if (k1 <= kFltMax) z.k1 = k1; // NaN / inf -> no certificate
	return z;
}

// For a well-conditioned triangle t falls with height with slope in [0.998, 1.002] (ray_z_run), and every f32 t of the cell is
// within bnd_r <= bnd + 16u |t_r| <= 3 bnd of the exact one.  So with d_r = z_r - z_0 >= 0:
//   t0 - 2 bnd - 1.01 d_r > eps   ->  sub-voxel r (and every lower one) is hit        (ray_z_run's "all hit" claim for the run z_0..z_r)
//   t0 + 4 bnd - 0.99 d_r <= eps  ->  sub-voxel r (and every higher one) is not hit   (t_r <= t0 + bnd_0 + bnd_r - 0.998 d_r)
// The two index estimates only choose where to start; both claims are checked against the actual centres, and the sub-voxels
// between them (none or one, as the gap is 2 % of the height plus 6 bnd) are evaluated.
// ---- per (triangle, sub-voxel column): two heights that classify most cells of the grid column without evaluating anything per
// cell.  With t falling at slope in [0.998, 1.002] and every f32 t of the column within bndc = k1 + 16u(|t| + 2.2 H) of the
// exact one (H = height of the grid column; ray_z_run's bound with the column-wide |Tz| and span):
//   claim 1 (ray_z_run "all hit", run [zMin, z]):   t(zMin) - 2 bndc - 1.01 (z - zMin) > eps  ->  every centre in [zMin, z] is hit
//   a second evaluation at zr = that height (clamped to the column) tightens both sides, the slope uncertainty now acting on
//   the short distance from zr to the crossing instead of on the whole column:
//   claim 2 (same, run [zr, z]):                     t(zr) - 2 bndc - 1.01 (z - zr) > eps      ->  every centre in [zr, z] is hit
//   claim 3 (t_z <= t_zr + 2 bndc - 0.998 (z - zr)): t(zr) + 4 bndc - 0.99 (z - zr) <= eps     ->  no centre at or above z is hit (z >= zr)
// A cell whose highest centre lies below zAll is all hit, one whose lowest centre lies at or above zNone is not hit at all;
// only the cells in between (the one that holds the crossing, sometimes a neighbour) go through ray_cell_class / _exact.
// The margins absorb the f32 rounding of the threshold arithmetic itself.  No certificate (k1 < 0): zAll = -inf, zNone = +inf.
struct RayThr { float zAll, zNone; };
GPV_HD RayThr ray_col_thresholds(const RayTri& s, const RayCol& c, const RayColZ& z, float zMin, float zMax)
{
	RayThr r;
	r.zAll = -INFINITY; r.zNone = INFINITY;
	if (!(z.k1 >= 0.f)) return r;
	const float H = zMax - zMin;
	const float t1 = ray_cell_t(s, c, zMin);
	const float b1 = z.k1 + 9.5367431640625e-07f * (fabsf(t1) + 2.2f * H);
	const float mg = 1.9073486328125e-06f * (fabsf(zMin) + fabsf(zMax) + fabsf(t1)) + 1e-30f; // 32u of the magnitudes involved
	float zr = zMin + (t1 - 2.f * b1 - kEps) * 0.9900990128517151f - mg; // 1 / 1.01
	if (!(zr > zMin)) zr = zMin; // (also NaN)
	if (zr > zMax) zr = zMax;
	const float t2 = ray_cell_t(s, c, zr);
	const float b2 = z.k1 + 9.5367431640625e-07f * (fabsf(t2) + 2.2f * H);
	const float za = zr + (t2 - 2.f * b2 - kEps) * 0.9900990128517151f - mg;
	const float zn = zr + (t2 + 4.f * b2 - kEps) * 1.0101009607315063f + mg;  // 1 / 0.99
	if (!(b1 <= kFltMax) || !(b2 <= kFltMax) || !(t1 == t1) || !(t2 == t2)) return r;
	r.zAll = za;                 // NaN compares false everywhere: no cell is classified by it
	r.zNone = zn > zr ? zn : zr; // claim 3 speaks about heights at or above zr only
	return r;
}

// Fast classification of one cell: 0 = no sub-voxel is hit, 1 = every sub-voxel is hit, 2 = undecided (the crossing lies in or
// next to the cell, or there is no certificate): call ray_cell_mask_exact.  Split from the exact part so that a warp can run the
// cheap classification over many cells first and then the exact part with all undecided lanes together.
GPV_HD int ray_cell_class(const RayTri& s, const RayCol& c, const RayColZ& z, float midz, float h1z, float h2z)
{
	if (!(z.k1 >= 0.f)) return 2;
	const float t0 = ray_cell_t(s, c, l2_centre(0, h2z, midz, h1z));
	const float bnd = z.k1 + 9.5367431640625e-07f * fabsf(t0); // >= ray_z_run's bound for this cell (column-wide |Tz| and span)
	if (t0 + 2.f * bnd <= kEps) return 0;              // every comparison is false for NaN / inf: undecided
	if (t0 - 2.f * bnd - z.span > kEps) return 1;
	return 2;
}

// The undecided cells: certified hit prefix, certified miss suffix, exact evaluation in between (or of every sub-voxel when
// there is no certificate).
GPV_HD unsigned ray_cell_mask_exact(const RayTri& s, const RayCol& c, const RayColZ& z, float midz, float h1z, float h2z, int n2)
{
	const float zLo = l2_centre(0, h2z, midz, h1z);
	int first = 0, last = n2; // sub-voxels [first, last) are evaluated; below first: hit, from last on: not hit
	if (z.k1 >= 0.f) {
		const float t0 = ray_cell_t(s, c, zLo);
		const float bnd = z.k1 + 9.5367431640625e-07f * fabsf(t0);
		if (bnd <= kFltMax && t0 == t0) {
			const float hi = t0 - 2.f * bnd;
			const float A = (hi - kEps) * z.inv101;
			int a = A >= 0.f ? (A < (float)n2 ? (int)A : n2 - 1) : -1;
			while (a >= 0 && !(hi - 1.01f * (l2_centre(a, h2z, midz, h1z) - zLo) > kEps)) a--;
			first = a + 1;
			const float lo = t0 + 4.f * bnd;
			const float B = ceilf((lo - kEps) * z.inv099);
			int m = B < (float)n2 ? (B > (float)first ? (int)B : first) : n2;
			while (m < n2 && !(lo - 0.99f * (l2_centre(m, h2z, midz, h1z) - zLo) <= kEps)) m++;
			last = m;
		}
	}
	unsigned mask = first >= 32 ? 0xffffffffu : ((1u << first) - 1u);
	for (int r = first; r < last; r++) mask |= (unsigned)(ray_cell_t(s, c, l2_centre(r, h2z, midz, h1z)) > kEps) << r;
	return mask;
}

GPV_HD unsigned ray_cell_mask(const RayTri& s, const RayCol& c, const RayColZ& z, float midz, float h1z, float h2z, int n2)
{
	const int cls = ray_cell_class(s, c, z, midz, h1z, h2z);
	if (cls == 0) return 0u;
	if (cls == 1) return n2 >= 32 ? 0xffffffffu : ((1u << n2) - 1u);
	return ray_cell_mask_exact(s, c, z, midz, h1z, h2z, n2);
}

// ---- certified candidate columns for the +Z parity fill of one triangle (DESIGN.md "Certified fill").
// kind 0: no column can be hit; 1: candidates [i0,i1]x[j0,j1]; 2: ill-conditioned -> every column must be tested.
GPV_HD int fill_candidates(const RayTri& s, float minx, float miny, float gsx, float gsy, int nx, int ny, int& i0, int& i1, int& j0, int& j1)
{
	if (!s.ok) return 0;
	float ax = fabsf(s.e1x), ay = fabsf(s.e1y), bx = fabsf(s.e2x), by = fabsf(s.e2y);
	float B = fmaxf(fmaxf(ax * by + ay * bx, 2.f * (bx * by)), 2.f * (ax * ay));
	if (!(fabsf(s.det) >= 1.6e-5f * B)) return 2;
	// With the signs: u' >= 0 and v' >= 0 force U, V >= -2.05/127 = -0.0162 and u'+v' <= 1 forces U+V <= 1.04, so
	// T = U*a + V*b lies in the projected triangle {0, a, b} grown by 6 % of (|a|+|b|) per axis.  Column centres are
	// x_i = minx + (i+0.5)*gs up to f32 rounding: i >= (lo-minx)/gs - 0.5; the slack covers the rounding of lo/hi themselves
	// (u*|v1|), of the centre table (u*|coord|) and of this index arithmetic.
	const float mx = 0.06f * (ax + bx) + 1e-30f, my = 0.06f * (ay + by) + 1e-30f;
	const float lox = s.v1x + fminf(0.f, fminf(s.e1x, s.e2x)) - mx, hix = s.v1x + fmaxf(0.f, fmaxf(s.e1x, s.e2x)) + mx;
	const float loy = s.v1y + fminf(0.f, fminf(s.e1y, s.e2y)) - my, hiy = s.v1y + fmaxf(0.f, fmaxf(s.e1y, s.e2y)) + my;
	const float sx = 0.02f + 4.8e-7f * (fabsf(lox) + fabsf(hix) + fabsf(minx)) / gsx, sy = 0.02f + 4.8e-7f * (fabsf(loy) + fabsf(hiy) + fabsf(miny)) / gsy;
	float lo = ceilf((lox - minx) / gsx - 0.5f - sx), hi = floorf((hix - minx) / gsx - 0.5f + sx);
	if (!(lo == lo) || !(hi == hi)) return 2;
	i0 = lo < 0.f ? 0 : (lo > 2e9f ? 2000000000 : (int)lo);
	i1 = hi > (float)(nx - 1) ? nx - 1 : (hi < -1.f ? -1 : (int)hi);
	lo = ceilf((loy - miny) / gsy - 0.5f - sy); hi = floorf((hiy - miny) / gsy - 0.5f + sy);
	if (!(lo == lo) || !(hi == hi)) return 2;
	j0 = lo < 0.f ? 0 : (lo > 2e9f ? 2000000000 : (int)lo);
	j1 = hi > (float)(ny - 1) ? ny - 1 : (hi < -1.f ? -1 : (int)hi);
	return (i0 <= i1 && j0 <= j1) ? 1 : 0;
}

// ---- vertex -> Level-1 cell index (cu:333-361): int(((v-min)/(max-min))*n), with the max-edge fix-up
GPV_HD int cell_of(float v, float mn, float mx, int n)
{
	int b = (int)((v - mn) / (mx - mn) * (float)n);
	if (b == n && v == mx) b--;
	return b;
}

// ---- reference uchar encodings (src/Object.cpp:2940-2942, :3031-3034)
GPV_HD unsigned char encode_normal(float n) { return (unsigned char)(n * 85.33333587646484375f + 127.0f); } // float(256.0/3.0)

// VectorNormalize (includes/FloatVector.h:333-342)
GPV_HD void normalize3(float& x, float& y, float& z)
{
	float mag = sqrtf(x * x + y * y + z * z);
	if (mag != 0) { x /= mag; y /= mag; z /= mag; }
}

} // namespace gpv
