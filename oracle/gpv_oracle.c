/* oracle/gpv_oracle.c -- TEST INFRASTRUCTURE ONLY (see gpv_oracle.h).
 *
 * CPU restatement of the reference voxelizer path.  Strict IEEE-754: compile with -O2 -ffp-contract=off; every f32
 * expression below is written in the reference's operand order, and `double` appears exactly where the reference's C++
 * promotes (a double literal or variable in the expression).
 */
#define _GNU_SOURCE
#include "gpv_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------------------------------ predicates */

/* one separating-axis check: projections pa,pb of the two relevant vertices against radius rad
 * (the `if (min>rad || max<-rad) return 0;` tail of every AXISTEST_* macro, src/TriBoxIntersection.cpp:84-133) */
static inline int sep(float pa, float pb, float rad)
{
	float mn, mx;
	if (pa < pb) { mn = pa; mx = pb; } else { mn = pb; mx = pa; }
	return mn > rad || mx < -rad;
}

/* TriBoxOverlap, src/TriBoxIntersection.cpp:135-210 == TriBoxOverlapCUDA, cuda/CUDAClassifyTessellation.cu:234-309.
 * AND of 13 side-effect-free predicates, evaluated here in the reference's order. */
int gpvo_tribox(const float c[3], const float h[3], const float tri[9])
{
	float v0[3], v1[3], v2[3], e0[3], e1[3], e2[3];
	for (int a = 0; a < 3; a++) { v0[a] = tri[a] - c[a]; v1[a] = tri[3 + a] - c[a]; v2[a] = tri[6 + a] - c[a]; }
	for (int a = 0; a < 3; a++) { e0[a] = v1[a] - v0[a]; e1[a] = v2[a] - v1[a]; e2[a] = v0[a] - v2[a]; }
	float fex, fey, fez;

	fex = fabsf(e0[0]); fey = fabsf(e0[1]); fez = fabsf(e0[2]);
	/* X01(e0z,e0y) */ if (sep(e0[2] * v0[1] - e0[1] * v0[2], e0[2] * v2[1] - e0[1] * v2[2], fez * h[1] + fey * h[2])) return 0;
	/* Y02(e0z,e0x) */ if (sep(-e0[2] * v0[0] + e0[0] * v0[2], -e0[2] * v2[0] + e0[0] * v2[2], fez * h[0] + fex * h[2])) return 0;
	/* Z12(e0y,e0x) */ if (sep(e0[1] * v1[0] - e0[0] * v1[1], e0[1] * v2[0] - e0[0] * v2[1], fey * h[0] + fex * h[1])) return 0;

	fex = fabsf(e1[0]); fey = fabsf(e1[1]); fez = fabsf(e1[2]);
	/* X01(e1z,e1y) */ if (sep(e1[2] * v0[1] - e1[1] * v0[2], e1[2] * v2[1] - e1[1] * v2[2], fez * h[1] + fey * h[2])) return 0;
	/* Y02(e1z,e1x) */ if (sep(-e1[2] * v0[0] + e1[0] * v0[2], -e1[2] * v2[0] + e1[0] * v2[2], fez * h[0] + fex * h[2])) return 0;
	/* Z0 (e1y,e1x) */ if (sep(e1[1] * v0[0] - e1[0] * v0[1], e1[1] * v1[0] - e1[0] * v1[1], fey * h[0] + fex * h[1])) return 0;

	fex = fabsf(e2[0]); fey = fabsf(e2[1]); fez = fabsf(e2[2]);
	/* X2 (e2z,e2y) */ if (sep(e2[2] * v0[1] - e2[1] * v0[2], e2[2] * v1[1] - e2[1] * v1[2], fez * h[1] + fey * h[2])) return 0;
	/* Y1 (e2z,e2x) */ if (sep(-e2[2] * v0[0] + e2[0] * v0[2], -e2[2] * v1[0] + e2[0] * v1[2], fez * h[0] + fex * h[2])) return 0;
	/* Z12(e2y,e2x) */ if (sep(e2[1] * v1[0] - e2[0] * v1[1], e2[1] * v2[0] - e2[0] * v2[1], fey * h[0] + fex * h[1])) return 0;

	for (int a = 0; a < 3; a++) { /* FINDMINMAX + per-axis AABB test, :187-199 */
		float mn = v0[a], mx = v0[a];
		if (v1[a] < mn) mn = v1[a];
		if (v1[a] > mx) mx = v1[a];
		if (v2[a] < mn) mn = v2[a];
		if (v2[a] > mx) mx = v2[a];
		if (mn > h[a] || mx < -h[a]) return 0;
	}
	/* plane of the triangle vs box, CROSS(normal,e0,e1) then planeBoxOverlap (:64-80, :205-207) */
	float n[3] = { e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2], e0[0] * e1[1] - e0[1] * e1[0] };
	float vmin[3], vmax[3];
	for (int q = 0; q < 3; q++) {
		float v = v0[q];
		if (n[q] > 0.0f) { vmin[q] = -h[q] - v; vmax[q] = h[q] - v; }
		else { vmin[q] = h[q] - v; vmax[q] = -h[q] - v; }
	}
	if (n[0] * vmin[0] + n[1] * vmin[1] + n[2] * vmin[2] > 0.0f) return 0;
	if (n[0] * vmax[0] + n[1] * vmax[1] + n[2] * vmax[2] >= 0.0f) return 1;
	return 0;
}

#define EPS_D 0.000001 /* the reference's EPSILON is a double literal: comparisons promote (src/TriRayIntersection.cpp:38) */

/* triangle_ray_intersection, src/TriRayIntersection.cpp:78-131 == TriRayIntersectCUDA, cu:102-155 (general form) */
int gpvo_triray(const float V1[3], const float V2[3], const float V3[3], const float O[3], const float D[3])
{
	float e1[3], e2[3], P[3], Q[3], T[3];
	for (int a = 0; a < 3; a++) { e1[a] = V2[a] - V1[a]; e2[a] = V3[a] - V1[a]; }
	P[0] = D[1] * e2[2] - D[2] * e2[1];
	P[1] = D[2] * e2[0] - D[0] * e2[2];
	P[2] = D[0] * e2[1] - D[1] * e2[0];
	float det = e1[0] * P[0] + e1[1] * P[1] + e1[2] * P[2];
	if (det > -EPS_D && det < EPS_D) return 0;
	float inv = 1.f / det;
	for (int a = 0; a < 3; a++) T[a] = O[a] - V1[a];
	float u = (T[0] * P[0] + T[1] * P[1] + T[2] * P[2]) * inv;
	if (u < 0.f || u > 1.f) return 0;
	Q[0] = T[1] * e1[2] - T[2] * e1[1];
	Q[1] = T[2] * e1[0] - T[0] * e1[2];
	Q[2] = T[0] * e1[1] - T[1] * e1[0];
	float v = (D[0] * Q[0] + D[1] * Q[1] + D[2] * Q[2]) * inv;
	if (v < 0.f || u + v > 1.f) return 0;
	float t = (e2[0] * Q[0] + e2[1] * Q[1] + e2[2] * Q[2]) * inv;
	if (t > EPS_D) return 1;
	return 0;
}

/* Axis-specialised Moller-Trumbore for D = (0,0,1), split into the part that is constant along a
 * z-column and the per-cell part.  f32 constant: fl32(1e-6) = 9.99999997e-7 < 1e-6, so for a float x
 *   (double)x <  1e-6  <=>  x <= EPS_F      and      (double)x > 1e-6  <=>  x > EPS_F.
 * Exact for finite intermediates (adding a +-0 term never changes a non-zero f32; comparisons ignore the sign of zero). */
#define EPS_F 9.99999997475242707878e-07f

typedef struct { float e1x, e1y, e1z, e2x, e2y, e2z, v1x, v1y, v1z, det, inv; int ok; } tri_z; /* per triangle */

static inline void triz_setup(const float* tri, tri_z* s)
{
	s->v1x = tri[0]; s->v1y = tri[1]; s->v1z = tri[2];
	s->e1x = tri[3] - tri[0]; s->e1y = tri[4] - tri[1]; s->e1z = tri[5] - tri[2];
	s->e2x = tri[6] - tri[0]; s->e2y = tri[7] - tri[1]; s->e2z = tri[8] - tri[2];
	/* P = (-e2y, e2x, 0);  det = (e1x*P0 + e1y*P1) + e1z*0 */
	s->det = s->e1x * (-s->e2y) + s->e1y * s->e2x;
	s->ok = fabsf(s->det) > EPS_F && fabsf(s->det) <= 3.402823466e+38f; /* NaN/inf det can never reach `t > EPS` */
	s->inv = 1.f / s->det;
}

/* column part: u, v tests for ray origin (ox,oy,*).  Returns 1 if the ray passes them; outputs what the t test needs. */
typedef struct { float c0, c1, c2; } col_z;
static inline int triz_column(const tri_z* s, float ox, float oy, col_z* c)
{
	float Tx = ox - s->v1x, Ty = oy - s->v1y;
	float u = (Tx * (-s->e2y) + Ty * s->e2x) * s->inv;
	if (u < 0.f || u > 1.f) return 0;
	float Q2 = Tx * s->e1y - Ty * s->e1x;
	float v = Q2 * s->inv;
	if (v < 0.f || u + v > 1.f) return 0;
	c->c0 = Ty * s->e1z; /* first product of Q0 = Ty*e1z - Tz*e1y */
	c->c1 = Tx * s->e1z; /* second product of Q1 = Tz*e1x - Tx*e1z */
	c->c2 = s->e2z * Q2; /* last term of DOT(e2,Q) */
	return 1;
}
static inline int triz_cell(const tri_z* s, const col_z* c, float oz)
{
	float Tz = oz - s->v1z;
	float Q0 = c->c0 - Tz * s->e1y;
	float Q1 = Tz * s->e1x - c->c1;
	float t = (s->e2x * Q0 + s->e2y * Q1 + c->c2) * s->inv;
	return t > EPS_F;
}

void gpvo_tribox_batch(int64_t n, const float* c, const float* h, const float* tri9, uint8_t* out)
{
	for (int64_t i = 0; i < n; i++) out[i] = (uint8_t)gpvo_tribox(c + i * 3, h + i * 3, tri9 + i * 9);
}
void gpvo_triray_batch(int64_t n, const float* org, const float* tri9, uint8_t* out)
{
	const float D[3] = { 0, 0, 1 };
	for (int64_t i = 0; i < n; i++) { const float* d = tri9 + i * 9; out[i] = (uint8_t)gpvo_triray(d, d + 3, d + 6, org + i * 3, D); }
}
void gpvo_triray_z_batch(int64_t n, const float* org, const float* tri9, uint8_t* out)
{
	for (int64_t i = 0; i < n; i++) {
		tri_z s; col_z c;
		triz_setup(tri9 + i * 9, &s);
		out[i] = (uint8_t)(s.ok && triz_column(&s, org[i * 3], org[i * 3 + 1], &c) && triz_cell(&s, &c, org[i * 3 + 2]));
	}
}

/* ------------------------------------------------------------------------------------------------ loaders */

/* bbox padding + maxModelSize, src/Object.cpp:572-583 (OBJ) == :285-296 (OFF); VectorMagnitude includes/FloatVector.h:323 */
static void pad_bbox(float mn[3], float mx[3], gpvo_mesh* m)
{
	float d0 = mx[0] - mn[0], d1 = mx[1] - mn[1], d2 = mx[2] - mn[2];
	float mag = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
	double modelSize = mag;
	float offset = 0.001 * modelSize;
	for (int a = 0; a < 3; a++) { m->bmax[a] = offset + mx[a]; m->bmin[a] = mn[a] - offset; }
	float s0 = m->bmax[0] - m->bmin[0], s1 = m->bmax[1] - m->bmin[1], s2 = m->bmax[2] - m->bmin[2];
	float s12 = s1 > s2 ? s1 : s2;
	m->maxModelSize = s0 > s12 ? s0 : s12; /* ___max(a, ___max(b, c)), includes/Includes.h:116 */
}

/* split(), src/Utilities.cpp:989-1004: single-character delimiter, empty tokens kept, trailing delimiter adds an empty token */
static int split1(const char* s, size_t len, char delim, const char** tok, size_t* tlen, int maxTok)
{
	int n = 0;
	size_t i = 0;
	while (i < len) {
		size_t j = i;
		while (j < len && s[j] != delim) j++;
		if (n < maxTok) { tok[n] = s + i; tlen[n] = j - i; }
		n++;
		if (j < len) { i = j + 1; if (i == len) { if (n < maxTok) { tok[n] = s + i; tlen[n] = 0; } n++; } }
		else i = j;
	}
	return n;
}

static int tok_float(const char* t, size_t len, float* out)
{
	char buf[128];
	if (len == 0 || len >= sizeof buf) return -1;
	memcpy(buf, t, len); buf[len] = 0;
	char* end;
	*out = strtof(buf, &end); /* std::stof -> strtof */
	return end == buf ? -1 : 0;
}
static int tok_int(const char* t, size_t len, long* out)
{
	char buf[64];
	if (len == 0 || len >= sizeof buf) return -1;
	memcpy(buf, t, len); buf[len] = 0;
	char* end;
	*out = strtol(buf, &end, 10); /* std::stoi */
	return end == buf ? -1 : 0;
}

/* Object::ReadObject, src/Object.cpp:395-584 (only what feeds the voxelizer: vertices, triangles, bbox). */
int gpvo_load_obj(const char* path, gpvo_mesh* m)
{
	memset(m, 0, sizeof *m);
	FILE* f = fopen(path, "rb");
	if (!f) return -1;
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	char* buf = (char*)malloc(sz + 1);
	if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
	fclose(f);
	int64_t capV = 1024, capT = 1024, nV = 0, nT = 0;
	float* V = (float*)malloc(capV * 3 * sizeof(float));
	float* T = (float*)malloc(capT * 9 * sizeof(float));
	float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
	long pos = 0;
	int rc = 0;
	while (pos < sz) {
		long e = pos;
		while (e < sz && buf[e] != '\n') e++;
		if (e >= sz) break; /* getline hit EOF: `if (!in.good()) break;` drops an unterminated last line (:425) */
		const char* line = buf + pos; size_t len = e - pos;
		pos = e + 1;
		const char *t1[64], *t2[64]; size_t l1[64], l2[64];
		int n1 = split1(line, len, ' ', t1, l1, 64), n2 = split1(line, len, '\t', t2, l2, 64);
		const char** tk = n1 > n2 ? t1 : t2; size_t* tl = n1 > n2 ? l1 : l2; int n = n1 > n2 ? n1 : n2; /* :419-423 */
		if (n > 64) n = 64;
		if (n == 0) continue;
		if (tl[0] == 1 && tk[0][0] == 'v') {
			float pt[3] = { 0, 0, 0 };
			if (nV > 0) { pt[0] = V[(nV - 1) * 3]; pt[1] = V[(nV - 1) * 3 + 1]; pt[2] = V[(nV - 1) * 3 + 2]; } /* `pt` persists across lines */
			for (int i = 1; i < n && i <= 3; i++) if (tok_float(tk[i], tl[i], &pt[i - 1])) { rc = -2; goto done; }
			if (nV == capV) { capV *= 2; V = (float*)realloc(V, capV * 3 * sizeof(float)); }
			V[nV * 3] = pt[0]; V[nV * 3 + 1] = pt[1]; V[nV * 3 + 2] = pt[2];
			if (nV == 0) for (int a = 0; a < 3; a++) mn[a] = mx[a] = pt[a];
			else for (int a = 0; a < 3; a++) { mn[a] = mn[a] < pt[a] ? mn[a] : pt[a]; mx[a] = mx[a] > pt[a] ? mx[a] : pt[a]; } /* MinFloat3/MaxFloat3 */
			nV++;
		} else if (tl[0] == 1 && tk[0][0] == 'f') {
			long idx[3] = { 0, 0, 0 };
			for (int i = 1; i < n && i <= 3; i++) {
				size_t k = 0;
				while (k < tl[i] && tk[i][k] != '/') k++; /* "a", "a/b", "a/b/c", "a//c": vertex index is the first field (:480-505) */
				if (tok_int(tk[i], k ? k : tl[i], &idx[i - 1])) { rc = -2; goto done; }
			}
			for (int c = 0; c < 3; c++) { idx[c] -= 1; if (idx[c] < 0 || idx[c] >= nV) { rc = -3; goto done; } }
			if (nT == capT) { capT *= 2; T = (float*)realloc(T, capT * 9 * sizeof(float)); }
			for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) T[nT * 9 + c * 3 + a] = V[idx[c] * 3 + a];
			nT++;
		}
	}
done:
	free(buf);
	if (rc == 0 && nV == 0) rc = -4;
	if (rc) { free(V); free(T); return rc; }
	m->tris = T; m->nTri = nT; m->nVerts = nV;
	pad_bbox(mn, mx, m);
	free(V);
	return 0;
}

/* Object::ReadOFFObject, src/Object.cpp:171-317: `>>` token stream; exactly three indices are read per face whatever its
 * count field says (:219-222); bbox over the vertices the triangles reference (:257-266). */
static const char* next_tok(const char* p, const char* end, const char** tok, size_t* len)
{
	while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
	if (p >= end) return NULL;
	*tok = p;
	while (p < end && !(*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
	*len = p - *tok;
	return p;
}
int gpvo_load_off(const char* path, gpvo_mesh* m)
{
	memset(m, 0, sizeof *m);
	FILE* f = fopen(path, "rb");
	if (!f) return -1;
	fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
	char* buf = (char*)malloc(sz + 1);
	if (fread(buf, 1, sz, f) != (size_t)sz) { fclose(f); free(buf); return -1; }
	fclose(f);
	const char *p = buf, *end = buf + sz, *tk; size_t tl;
	long hdr[3];
	int rc = 0;
	float *V = NULL, *T = NULL;
	if (!(p = next_tok(p, end, &tk, &tl))) { rc = -2; goto done; } /* header word */
	for (int i = 0; i < 3; i++) { if (!(p = next_tok(p, end, &tk, &tl)) || tok_int(tk, tl, &hdr[i])) { rc = -2; goto done; } }
	long nV = hdr[0], nF = hdr[1];
	if (nV <= 0 || nF <= 0) { rc = -4; goto done; }
	V = (float*)malloc((size_t)nV * 3 * sizeof(float));
	T = (float*)malloc((size_t)nF * 9 * sizeof(float));
	for (long i = 0; i < nV * 3; i++) { if (!(p = next_tok(p, end, &tk, &tl)) || tok_float(tk, tl, &V[i])) { rc = -2; goto done; } }
	for (long i = 0; i < nF; i++) {
		long q[4];
		for (int c = 0; c < 4; c++) { if (!(p = next_tok(p, end, &tk, &tl)) || tok_int(tk, tl, &q[c])) { rc = -2; goto done; } }
		for (int c = 0; c < 3; c++) {
			if (q[c + 1] < 0 || q[c + 1] >= nV) { rc = -3; goto done; }
			for (int a = 0; a < 3; a++) T[i * 9 + c * 3 + a] = V[q[c + 1] * 3 + a];
		}
	}
	{
		float mn[3], mx[3];
		for (int a = 0; a < 3; a++) mn[a] = mx[a] = T[a];
		for (long i = 0; i < nF * 3; i++) for (int a = 0; a < 3; a++) {
			float x = T[i * 3 + a];
			mn[a] = mn[a] < x ? mn[a] : x; mx[a] = mx[a] > x ? mx[a] : x;
		}
		m->tris = T; m->nTri = nF; m->nVerts = nV; T = NULL;
		pad_bbox(mn, mx, m);
	}
done:
	free(buf); free(V); free(T);
	return rc;
}

void gpvo_mesh_from_tris(const float* tris, int64_t nTri, gpvo_mesh* m)
{
	memset(m, 0, sizeof *m);
	m->tris = (float*)malloc((size_t)nTri * 9 * sizeof(float));
	memcpy(m->tris, tris, (size_t)nTri * 9 * sizeof(float));
	m->nTri = nTri;
	float mn[3], mx[3];
	for (int a = 0; a < 3; a++) mn[a] = mx[a] = tris[a];
	for (int64_t i = 0; i < nTri * 3; i++) for (int a = 0; a < 3; a++) {
		float x = tris[i * 3 + a];
		mn[a] = mn[a] < x ? mn[a] : x; mx[a] = mx[a] > x ? mx[a] : x;
	}
	pad_bbox(mn, mx, m);
}

void gpvo_free_mesh(gpvo_mesh* m) { free(m->tris); m->tris = NULL; }

/* ------------------------------------------------------------------------------------------------ grid */

static int next_div4(int a) { return (a % 4 == 0) ? a : a + (4 - a % 4); } /* includes/Utilities.h:313 */

/* Grid sizing, Object::PerformVoxelization src/Object.cpp:3094-3134; half extents :2087, :2551-2552 */
void gpvo_make_grid(const float bmin[3], const float bmax[3], float maxModelSize, int voxelCount, int voxelCount2, gpvo_grid* g)
{
	float nominalGridSize = maxModelSize / (1.0 * voxelCount);
	int n2 = voxelCount2 > 0 ? voxelCount2 : 1;
	g->n2 = n2;
	for (int a = 0; a < 3; a++) {
		int n = (int)((bmax[a] - bmin[a]) / nominalGridSize);
		if (n == 0) n++;
		n = next_div4(n);
		g->numDiv[a] = n;
		g->gridSize[a] = (bmax[a] - bmin[a]) / (n * 1.0);
		g->gridSize2[a] = g->gridSize[a] / (n2 * 1.0);
		g->ext1[a] = g->gridSize[a] / 2.0;
		g->ext2[a] = g->gridSize2[a] / 2.0;
	}
}

/* L1 cell centre along one axis.  Three reference expressions, all double -> float:
 *   SAT centre   (p + 0.5)*boxExtents*2 + bBoxMin   cu:382-384 == src/Object.cpp:2329-2331
 *   ray origin   bBoxMin + (i + 0.5)*gridSize       src/Object.cpp:743-745
 *   mid point    (i + 0.5)*gridSize + bBoxMin       src/Object.cpp:2567-2569, :3178-3180
 * They agree whenever 2*ext == gridSize exactly (always, short of subnormal grid sizes); the oracle aborts otherwise. */
void gpvo_axis_table(const float bmin[3], const gpvo_grid* g, int a, float* out)
{
	for (int p = 0; p < g->numDiv[a]; p++) {
		float sat = (p + 0.5) * g->ext1[a] * 2 + bmin[a];
		float ray = bmin[a] + (p + 0.5) * g->gridSize[a];
		float mid = (p + 0.5) * g->gridSize[a] + bmin[a];
		if (sat != ray || ray != mid) { fprintf(stderr, "gpv_oracle: centre formulas disagree (axis %d cell %d)\n", a, p); abort(); }
		out[p] = sat;
	}
}

/* vertex -> cell index, cu:333-361 == src/Object.cpp:2280-2308 */
static inline int cell_of(float v, float mn, float mx, int n)
{
	int b = (int)((v - mn) / (mx - mn) * n);
	if (b == n && v == mx) b--;
	return b;
}

/* ------------------------------------------------------------------------------------------------ voxelize */

typedef struct { int64_t key; } pair_t; /* (bucket << 32) | tri */
static int cmp_i64(const void* a, const void* b) { int64_t x = *(const int64_t*)a, y = *(const int64_t*)b; return (x > y) - (x < y); }

typedef struct { int64_t* v; int64_t n, cap; } vec64;
static void push64(vec64* v, int64_t x)
{
	if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 1 << 16; v->v = (int64_t*)realloc(v->v, v->cap * sizeof(int64_t)); }
	v->v[v->n++] = x;
}

typedef struct job job;
struct job {
	const gpvo_mesh* m; gpvo_result* r; const float *cx, *cy, *cz;
	int flags, nThreads, tid;
	volatile int64_t* next;
	/* crossing CSR for the certified / collist fills */
	const int64_t* crossOff; const int32_t* crossTri;
	int64_t rayTests, boxTests;
	int64_t b0, b1; /* boundary range for timed loops */
};

static int64_t fetch_add(volatile int64_t* p, int64_t d) { return __sync_fetch_and_add(p, d); }

static void run_threads(void* (*fn)(void*), job* proto, int nThreads, int64_t* sumRay, int64_t* sumBox)
{
	if (nThreads < 1) nThreads = 1;
	if (nThreads > 256) nThreads = 256;
	pthread_t th[256]; job jobs[256];
	volatile int64_t next = proto->b0;
	for (int i = 0; i < nThreads; i++) { jobs[i] = *proto; jobs[i].tid = i; jobs[i].nThreads = nThreads; jobs[i].next = &next; jobs[i].rayTests = jobs[i].boxTests = 0; }
	for (int i = 1; i < nThreads; i++) pthread_create(&th[i], NULL, fn, &jobs[i]);
	fn(&jobs[0]);
	for (int i = 1; i < nThreads; i++) pthread_join(th[i], NULL);
	for (int i = 0; i < nThreads; i++) { if (sumRay) *sumRay += jobs[i].rayTests; if (sumBox) *sumBox += jobs[i].boxTests; }
}

/* Object::ClassifyInOutCPU, src/Object.cpp:733-771: every cell against every triangle, general Moller-Trumbore. */
static void* fill_brute_worker(void* arg)
{
	job* j = (job*)arg;
	const gpvo_grid* g = &j->r->g;
	int nx = g->numDiv[0], ny = g->numDiv[1], nz = g->numDiv[2];
	const float D[3] = { 0, 0, 1 };
	for (;;) {
		int64_t k = fetch_add(j->next, 1);
		if (k >= nz) break;
		for (int jj = 0; jj < ny; jj++) for (int i = 0; i < nx; i++) {
			float O[3] = { j->cx[i], j->cy[jj], j->cz[k] };
			int n = 0;
			for (int64_t t = 0; t < j->m->nTri; t++) {
				const float* d = j->m->tris + t * 9;
				if (gpvo_triray(d, d + 3, d + 6, O, D)) n++;
			}
			if (n % 2 == 1) j->r->l1FillOnly[k * ny * nx + (int64_t)jj * nx + i] = 1;
		}
	}
	return NULL;
}

/* per-column sweep over a crossing CSR: u/v once per (column, triangle), t per cell (App. A.6) */
static void* fill_cross_worker(void* arg)
{
	job* j = (job*)arg;
	const gpvo_grid* g = &j->r->g;
	int nx = g->numDiv[0], ny = g->numDiv[1], nz = g->numDiv[2];
	for (;;) {
		int64_t col = fetch_add(j->next, 1);
		if (col >= (int64_t)nx * ny) break;
		int i = col % nx, jj = col / nx;
		for (int64_t q = j->crossOff[col]; q < j->crossOff[col + 1]; q++) {
			tri_z s; col_z c;
			triz_setup(j->m->tris + (int64_t)j->crossTri[q] * 9, &s);
			if (!s.ok || !triz_column(&s, j->cx[i], j->cy[jj], &c)) continue;
			for (int k = 0; k < nz; k++) if (triz_cell(&s, &c, j->cz[k])) j->r->l1FillOnly[(int64_t)k * ny * nx + col] ^= 1;
		}
	}
	return NULL;
}

/* Certified candidate columns of one triangle for the +Z parity fill.
 * Returns 0: no column can be hit; 1: candidates are [i0,i1]x[j0,j1]; 2: ill-conditioned, test every column.
 * Proof sketch (DESIGN.md "certified fill"): with a=e1.xy, b=e2.xy, T=fl(O-V1).xy and real U,V with T=U*a+V*b, the f32
 * numerators obey |nu'-U*D|,|nv'-V*D| <= g*B*(|U|+|V|), |det'-D| <= g*B, g=2^-23(1+2^-24),
 * B=max(|ax*by|+|ay*bx|, 2|bx*by|, 2|ax*ay|).  Passing 0<=u'<=1, v'>=0, u'+v'<=1 needs |nu'|,|nv'| <= |det'|(1+4*2^-24).
 * If |det'| >= 128*g*B this forces |U|,|V| <= 1.03, i.e. |O-V1| <= 1.04*(|a|+|b|) per axis. */
static int fill_candidates(const tri_z* s, const float bmin[3], const gpvo_grid* g, int* i0, int* i1, int* j0, int* j1)
{
	if (!s->ok) return 0;
	float ax = fabsf(s->e1x), ay = fabsf(s->e1y), bx = fabsf(s->e2x), by = fabsf(s->e2y);
	float B = fmaxf(fmaxf(ax * by + ay * bx, 2.f * (bx * by)), 2.f * (ax * ay));
	if (!(fabsf(s->det) >= 1.6e-5f * B)) return 2; /* 1.6e-5 > 128*2^-23*(1+slop); NaN/inf B lands here too */
	float rx = 1.04f * (ax + bx) + 1e-30f, ry = 1.04f * (ay + by) + 1e-30f;
	float lo, hi;
	lo = floorf((s->v1x - rx - bmin[0]) / g->gridSize[0]) - 1.f; hi = floorf((s->v1x + rx - bmin[0]) / g->gridSize[0]) + 1.f;
	if (!(lo == lo) || !(hi == hi)) return 2;
	*i0 = lo < 0.f ? 0 : (lo > 2e9f ? 2000000000 : (int)lo); *i1 = hi > (float)(g->numDiv[0] - 1) ? g->numDiv[0] - 1 : (int)hi;
	lo = floorf((s->v1y - ry - bmin[1]) / g->gridSize[1]) - 1.f; hi = floorf((s->v1y + ry - bmin[1]) / g->gridSize[1]) + 1.f;
	if (!(lo == lo) || !(hi == hi)) return 2;
	*j0 = lo < 0.f ? 0 : (lo > 2e9f ? 2000000000 : (int)lo); *j1 = hi > (float)(g->numDiv[1] - 1) ? g->numDiv[1] - 1 : (int)hi;
	return (*i0 <= *i1 && *j0 <= *j1) ? 1 : 0;
}

/* CSR from (bucket<<32 | tri) keys: sorts, optionally uniques; returns offsets (nb+1) and values */
static void build_csr(vec64* keys, int64_t nb, int unique, int64_t** offOut, int32_t** valOut)
{
	qsort(keys->v, keys->n, sizeof(int64_t), cmp_i64);
	if (unique) {
		int64_t w = 0;
		for (int64_t i = 0; i < keys->n; i++) if (i == 0 || keys->v[i] != keys->v[i - 1]) keys->v[w++] = keys->v[i];
		keys->n = w;
	}
	int64_t* off = (int64_t*)calloc(nb + 1, sizeof(int64_t));
	int32_t* val = (int32_t*)malloc((keys->n ? keys->n : 1) * sizeof(int32_t));
	for (int64_t i = 0; i < keys->n; i++) { off[(keys->v[i] >> 32) + 1]++; val[i] = (int32_t)(keys->v[i] & 0xffffffff); }
	for (int64_t b = 0; b < nb; b++) off[b + 1] += off[b];
	*offOut = off; *valOut = val;
}

/* reference uchar encodings, src/Object.cpp:2940-2942, :3031-3034 */
static inline uint8_t enc_normal(float n) { float normScale = 256.0 / 3.0; float normBias = 127.0; return (uint8_t)(n * normScale + normBias); }

/* VectorNormalize, includes/FloatVector.h:333-342 */
static inline void normalize3(float* a)
{
	float mag = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
	if (mag != 0) { a[0] /= mag; a[1] /= mag; a[2] /= mag; }
}

/* Level 2 for boundary cells [next..): kernels cu:403-504 in f32 + host averaging src/Object.cpp:2613-2632 */
static void* l2_worker(void* arg)
{
	job* j = (job*)arg;
	gpvo_result* r = j->r;
	const gpvo_grid* g = &r->g;
	int nx = g->numDiv[0], ny = g->numDiv[1], n2 = g->n2;
	int64_t n23 = r->n23;
	const float D[3] = { 0, 0, 1 };
	float* c2[3];
	for (int a = 0; a < 3; a++) c2[a] = (float*)malloc(n2 * sizeof(float));
	float* acc = (float*)malloc(n23 * 4 * sizeof(float));
	for (;;) {
		int64_t b = fetch_add(j->next, 1);
		if (b >= j->b1) break;
		int64_t l1 = r->boundaryIndex[b];
		int k = l1 / ((int64_t)nx * ny); int64_t ij = l1 - (int64_t)k * nx * ny; int jj = ij / nx, i = ij % nx;
		float mid[3] = { j->cx[i], j->cy[jj], j->cz[k] };
		/* L2 centre, cu:423-425 / :472-474:  (2p+1)*ext2 + mid - ext1  in f32, left to right */
		for (int a = 0; a < 3; a++) for (int p = 0; p < n2; p++) c2[a][p] = (2 * p + 1) * g->ext2[a] + mid[a] - g->ext1[a];
		uint8_t* st = r->l2State + b * n23;
		int64_t col = l1 % ((int64_t)nx * ny);
		const int32_t* cl = r->colTris + r->colOffset[col]; int64_t ncl = r->colOffset[col + 1] - r->colOffset[col];
		const int32_t* tl = r->cellTris + r->cellOffset[l1]; int64_t ntl = r->cellOffset[l1 + 1] - r->cellOffset[l1];
		/* CUDAClassifyInOutLevel2Kernel cu:450-504: parity over the XY-column list */
		if (j->flags & GPVO_L2_NAIVE) {
			for (int64_t loc = 0; loc < n23; loc++) {
				int rr = loc / (n2 * n2), pq = loc - rr * n2 * n2, q = pq / n2, p = pq % n2;
				float O[3] = { c2[0][p], c2[1][q], c2[2][rr] };
				int n = 0;
				for (int64_t t = 0; t < ncl; t++) { const float* d = j->m->tris + (int64_t)cl[t] * 9; if (gpvo_triray(d, d + 3, d + 6, O, D)) n++; }
				if (n % 2 == 1) st[loc] = 1;
			}
		} else {
			for (int64_t t = 0; t < ncl; t++) {
				tri_z s; triz_setup(j->m->tris + (int64_t)cl[t] * 9, &s);
				if (!s.ok) continue;
				for (int q = 0; q < n2; q++) for (int p = 0; p < n2; p++) {
					col_z c;
					if (!triz_column(&s, c2[0][p], c2[1][q], &c)) continue;
					for (int rr = 0; rr < n2; rr++) if (triz_cell(&s, &c, c2[2][rr])) st[rr * n2 * n2 + q * n2 + p] ^= 1;
				}
			}
		}
		j->rayTests += ncl * n23;
		/* CUDAClassifyTessellationLevel2Kernel cu:403-448: SAT over the cell list, 2 overwrites, normals accumulate */
		int wantN = !(j->flags & GPVO_NO_NORMALS);
		if (wantN) memset(acc, 0, n23 * 4 * sizeof(float));
		for (int64_t loc = 0; loc < n23; loc++) {
			int rr = loc / (n2 * n2), pq = loc - rr * n2 * n2, q = pq / n2, p = pq % n2;
			float c[3] = { c2[0][p], c2[1][q], c2[2][rr] };
			for (int64_t t = 0; t < ntl; t++) {
				const float* d = j->m->tris + (int64_t)tl[t] * 9;
				if (gpvo_tribox(c, g->ext2, d)) {
					st[loc] = 2;
					if (wantN) { /* CalculateNormal cu:311-318: cross(e01,e02), normalize() result discarded */
						float ax = d[3] - d[0], ay = d[4] - d[1], az = d[5] - d[2];
						float bx = d[6] - d[0], by = d[7] - d[1], bz = d[8] - d[2];
						acc[loc * 4 + 0] += ay * bz - az * by;
						acc[loc * 4 + 1] += az * bx - ax * bz;
						acc[loc * 4 + 2] += ax * by - ay * bx;
						acc[loc * 4 + 3] += 1;
					}
				}
			}
			if (wantN) {
				float avg[3] = { 0, 0, 0 };
				float numNormals = acc[loc * 4 + 3];
				if (numNormals > 0) {
					avg[0] = acc[loc * 4 + 0] / numNormals; avg[1] = acc[loc * 4 + 1] / numNormals; avg[2] = acc[loc * 4 + 2] / numNormals;
					normalize3(avg);
				}
				uint8_t* o = r->l2Normal + (b * n23 + loc) * 3;
				o[0] = enc_normal(avg[0]); o[1] = enc_normal(avg[1]); o[2] = enc_normal(avg[2]);
			}
		}
		j->boxTests += ntl * n23;
	}
	for (int a = 0; a < 3; a++) free(c2[a]);
	free(acc);
	return NULL;
}

int gpvo_voxelize(const gpvo_mesh* m, int voxelCount, int voxelCount2, int flags, int nThreads, gpvo_result* r)
{
	memset(r, 0, sizeof *r);
	gpvo_grid* g = &r->g;
	gpvo_make_grid(m->bmin, m->bmax, m->maxModelSize, voxelCount, voxelCount2, g);
	if (voxelCount2 <= 0) flags |= GPVO_NO_L2;
	int nx = g->numDiv[0], ny = g->numDiv[1], nz = g->numDiv[2];
	int64_t cells = (int64_t)nx * ny * nz, ncol = (int64_t)nx * ny;
	r->cells = cells;
	r->n23 = (int64_t)g->n2 * g->n2 * g->n2;
	float* cx = (float*)malloc(nx * sizeof(float)); float* cy = (float*)malloc(ny * sizeof(float)); float* cz = (float*)malloc(nz * sizeof(float));
	gpvo_axis_table(m->bmin, g, 0, cx); gpvo_axis_table(m->bmin, g, 1, cy); gpvo_axis_table(m->bmin, g, 2, cz);
	r->l1State = (uint8_t*)calloc(cells, 1);
	r->l1FillOnly = (uint8_t*)calloc(cells, 1);
	r->prefix = (int32_t*)malloc(cells * sizeof(int32_t));
	r->cellCount = (int32_t*)calloc(cells, sizeof(int32_t));

	/* ---- L1 triangle classification: Object::ClassifyTessellation src/Object.cpp:2256-2342 == kernel cu:320-401 ---- */
	vec64 cellKeys = { 0 }, colKeys = { 0 };
	for (int64_t t = 0; t < m->nTri; t++) {
		const float* v = m->tris + t * 9;
		int lo[3], hi[3];
		for (int a = 0; a < 3; a++) {
			int b0 = cell_of(v[a], m->bmin[a], m->bmax[a], g->numDiv[a]);
			int b1 = cell_of(v[3 + a], m->bmin[a], m->bmax[a], g->numDiv[a]);
			int b2 = cell_of(v[6 + a], m->bmin[a], m->bmax[a], g->numDiv[a]);
			lo[a] = b0 < (b1 < b2 ? b1 : b2) ? b0 : (b1 < b2 ? b1 : b2);
			hi[a] = b0 > (b1 > b2 ? b1 : b2) ? b0 : (b1 > b2 ? b1 : b2);
		}
		for (int p = lo[0]; p <= hi[0] && p < nx; p++) for (int q = lo[1]; q <= hi[1] && q < ny; q++) {
			int any = 0;
			for (int rr = lo[2]; rr <= hi[2] && rr < nz; rr++) {
				r->l1BoxTests++;
				if (p < 0 || q < 0 || rr < 0) continue; /* the reference would index out of bounds; cannot occur for a bbox built from the mesh */
				float c[3] = { cx[p], cy[q], cz[rr] };
				if (gpvo_tribox(c, g->ext1, v)) {
					int64_t idx = (int64_t)rr * ny * nx + (int64_t)q * nx + p;
					push64(&cellKeys, (idx << 32) | t);
					r->cellCount[idx]++;
					any = 1;
				}
			}
			if (any) push64(&colKeys, (((int64_t)q * nx + p) << 32) | t); /* per-column dedup list, src/Object.cpp:2146-2180 */
		}
	}
	r->l1BoxHits = cellKeys.n;
	build_csr(&cellKeys, cells, 0, &r->cellOffset, &r->cellTris);
	build_csr(&colKeys, ncol, 1, &r->colOffset, &r->colTris);
	r->colCount = (int32_t*)malloc(ncol * sizeof(int32_t));
	for (int64_t c = 0; c < ncol; c++) { r->colCount[c] = (int32_t)(r->colOffset[c + 1] - r->colOffset[c]); r->l1ColRayTests += (int64_t)r->colCount[c] * nz; }
	for (int64_t c = 0; c < cells; c++) if (r->cellCount[c] > r->maxPerCell) r->maxPerCell = r->cellCount[c];
	free(cellKeys.v); free(colKeys.v);

	/* ---- L1 solid fill: Object::ClassifyInOutCPU src/Object.cpp:716-779 ---- */
	job proto; memset(&proto, 0, sizeof proto);
	proto.m = m; proto.r = r; proto.cx = cx; proto.cy = cy; proto.cz = cz; proto.flags = flags; proto.b0 = 0;
	int fill = flags & 3;
	if (fill == GPVO_FILL_BRUTE) {
		run_threads(fill_brute_worker, &proto, nThreads, NULL, NULL);
	} else {
		int64_t* off; int32_t* val;
		if (fill == GPVO_FILL_COLLIST) { off = r->colOffset; val = r->colTris; }
		else {
			vec64 keys = { 0 };
			for (int64_t t = 0; t < m->nTri; t++) {
				tri_z s; triz_setup(m->tris + t * 9, &s);
				int i0, i1, j0, j1;
				int kind = fill_candidates(&s, m->bmin, g, &i0, &i1, &j0, &j1);
				if (kind == 0) continue;
				if (kind == 2) { i0 = 0; i1 = nx - 1; j0 = 0; j1 = ny - 1; r->fillIllConditioned++; }
				for (int jj = j0; jj <= j1; jj++) for (int i = i0; i <= i1; i++) {
					col_z c;
					if (triz_column(&s, cx[i], cy[jj], &c)) push64(&keys, (((int64_t)jj * nx + i) << 32) | t);
				}
			}
			r->fillCrossings = keys.n;
			build_csr(&keys, ncol, 0, &off, &val);
			free(keys.v);
		}
		proto.crossOff = off; proto.crossTri = val;
		run_threads(fill_cross_worker, &proto, nThreads, NULL, NULL);
		if (fill != GPVO_FILL_COLLIST) { free(off); free(val); }
	}

	/* ---- state = fill, then SAT overwrites with 2 (src/Object.cpp:3158 then :3202); prefix sum :3270-3279 ---- */
	int32_t nb = 0;
	for (int64_t c = 0; c < cells; c++) {
		r->l1State[c] = r->cellCount[c] > 0 ? 2 : r->l1FillOnly[c];
		r->prefix[c] = nb;
		if (r->l1State[c] == 2) nb++;
	}
	r->nBoundary = nb;
	r->boundaryIndex = (int32_t*)malloc((nb ? nb : 1) * sizeof(int32_t));
	for (int64_t c = 0, b = 0; c < cells; c++) if (r->l1State[c] == 2) r->boundaryIndex[b++] = (int32_t)c;

	/* ---- L1 normals, src/Object.cpp:3219-3253 (canonical list order = ascending triangle id) ---- */
	if (!(flags & GPVO_NO_NORMALS)) {
		r->l1Normal = (uint8_t*)malloc((size_t)cells * 3);
		for (int64_t c = 0; c < cells; c++) {
			float avg[3] = { 0, 0, 0 };
			int numTri = r->cellCount[c];
			if (numTri != 0) {
				float sum[3] = { 0, 0, 0 };
				for (int64_t q = r->cellOffset[c]; q < r->cellOffset[c + 1]; q++) {
					const float* d = m->tris + (int64_t)r->cellTris[q] * 9;
					float s1[3] = { d[3] - d[0], d[4] - d[1], d[5] - d[2] }, s2[3] = { d[6] - d[0], d[7] - d[1], d[8] - d[2] };
					/* VectorCrossProduct, includes/FloatVector.h:293 */
					float fn[3] = { s1[1] * s2[2] - s2[1] * s1[2], s1[2] * s2[0] - s2[2] * s1[0], s1[0] * s2[1] - s2[0] * s1[1] };
					normalize3(fn);
					sum[0] += fn[0]; sum[1] += fn[1]; sum[2] += fn[2];
				}
				float nt = (float)numTri;
				avg[0] = sum[0] / nt; avg[1] = sum[1] / nt; avg[2] = sum[2] / nt;
				normalize3(avg);
			}
			r->l1Normal[c * 3] = enc_normal(avg[0]); r->l1Normal[c * 3 + 1] = enc_normal(avg[1]); r->l1Normal[c * 3 + 2] = enc_normal(avg[2]);
		}
	}

	/* ---- Level 2 ---- */
	if (!(flags & GPVO_NO_L2)) {
		r->l2State = (uint8_t*)calloc((size_t)nb * r->n23 + 1, 1);
		if (!(flags & GPVO_NO_NORMALS)) r->l2Normal = (uint8_t*)malloc((size_t)nb * r->n23 * 3 + 1);
		proto.b0 = 0; proto.b1 = nb;
		run_threads(l2_worker, &proto, nThreads, &r->l2RayTests, &r->l2BoxTests);
	}

	/* ---- counts, src/Object.cpp:3353-3378 ---- */
	for (int64_t c = 0; c < cells; c++) { if (r->l1State[c] == 1) r->l1Inside++; if (r->l1State[c] == 2) r->l1Boundary++; }
	if (r->l2State) for (int64_t i = 0; i < (int64_t)nb * r->n23; i++) { if (r->l2State[i] % 2 == 1) r->l2Inside++; if (r->l2State[i] == 2) r->l2Boundary++; }
	free(cx); free(cy); free(cz);
	return 0;
}

void gpvo_free_result(gpvo_result* r)
{
	free(r->l1State); free(r->l1FillOnly); free(r->prefix); free(r->boundaryIndex); free(r->cellCount); free(r->cellOffset);
	free(r->cellTris); free(r->colCount); free(r->colOffset); free(r->colTris); free(r->l1Normal); free(r->l2State); free(r->l2Normal);
	memset(r, 0, sizeof *r);
}

/* Object::SaveVoxelization, src/Object.cpp:2934-3075.  ostream << float == "%g" (precision 6). */
int gpvo_save(const gpvo_mesh* m, const gpvo_result* r, int objID, const char* dir)
{
	char path[4096];
	const gpvo_grid* g = &r->g;
	int l2 = r->l2State != NULL;
	snprintf(path, sizeof path, "%s/Obj%dVoxelConfig.txt", dir, objID);
	FILE* f = fopen(path, "w");
	if (!f) return -1;
	fprintf(f, "Obj%d\n", objID);
	fprintf(f, "%g\t%g\t%g\n", m->bmin[0], m->bmin[1], m->bmin[2]);
	fprintf(f, "%g\t%g\t%g\n", m->bmax[0], m->bmax[1], m->bmax[2]);
	fprintf(f, "%d\t%d\t%d\n", g->numDiv[0], g->numDiv[1], g->numDiv[2]);
	fprintf(f, "%g\t%g\t%g\n", g->gridSize[0], g->gridSize[1], g->gridSize[2]);
	fprintf(f, "%ld\n%ld\n", (long)r->l1Inside, (long)r->l1Boundary);
	if (l2) {
		fprintf(f, "%d\t%d\t%d\n", g->n2, g->n2, g->n2);
		fprintf(f, "%g\t%g\t%g\n", g->gridSize2[0], g->gridSize2[1], g->gridSize2[2]);
		fprintf(f, "%ld\n%ld\n", (long)r->l2Inside, (long)r->l2Boundary);
	}
	fclose(f);
	struct { const char* name; const void* p; size_t n; int scale; } out[5] = {
		{ "Level1InOut.raw", r->l1State, (size_t)r->cells, 127 }, { "Level1Normal.raw", r->l1Normal, (size_t)r->cells * 3, 1 },
		{ "Level1BoundaryPrefixSum.raw", r->prefix, (size_t)r->cells * 4, 1 },
		{ "Level2InOut.raw", r->l2State, (size_t)r->nBoundary * r->n23, 127 }, { "Level2Normal.raw", r->l2Normal, (size_t)r->nBoundary * r->n23 * 3, 1 } };
	for (int i = 0; i < (l2 ? 5 : 2); i++) {
		snprintf(path, sizeof path, "%s/Obj%d%s", dir, objID, out[i].name);
		f = fopen(path, "wb");
		if (!f) return -1;
		if (out[i].scale == 127) { /* uchar(state * 127.0f), src/Object.cpp:3031 */
			uint8_t* tmp = (uint8_t*)malloc(out[i].n ? out[i].n : 1);
			for (size_t k = 0; k < out[i].n; k++) tmp[k] = (uint8_t)(((const uint8_t*)out[i].p)[k] * 127.0f);
			fwrite(tmp, 1, out[i].n, f); free(tmp);
		} else if (out[i].p) fwrite(out[i].p, 1, out[i].n, f);
		fclose(f);
	}
	return 0;
}

/* timed kernel-form Level-2 SAT nest (cu:428-445) over boundary cells [b0,b1) -- bench.py cpu_baseline, kind "port" */
static void* time_l2_worker(void* arg)
{
	job* j = (job*)arg;
	gpvo_result* r = j->r;
	const gpvo_grid* g = &r->g;
	int nx = g->numDiv[0], ny = g->numDiv[1], n2 = g->n2;
	int64_t n23 = r->n23, hits = 0;
	for (;;) {
		int64_t b = fetch_add(j->next, 1);
		if (b >= j->b1) break;
		int64_t l1 = r->boundaryIndex[b];
		int k = l1 / ((int64_t)nx * ny); int64_t ij = l1 - (int64_t)k * nx * ny; int jj = ij / nx, i = ij % nx;
		float mid[3] = { j->cx[i], j->cy[jj], j->cz[k] };
		const int32_t* tl = r->cellTris + r->cellOffset[l1]; int64_t ntl = r->cellOffset[l1 + 1] - r->cellOffset[l1];
		for (int64_t loc = 0; loc < n23; loc++) {
			int rr = loc / (n2 * n2), pq = loc - rr * n2 * n2, q = pq / n2, p = pq % n2;
			float c[3];
			c[0] = (2 * p + 1) * g->ext2[0] + mid[0] - g->ext1[0];
			c[1] = (2 * q + 1) * g->ext2[1] + mid[1] - g->ext1[1];
			c[2] = (2 * rr + 1) * g->ext2[2] + mid[2] - g->ext1[2];
			for (int64_t t = 0; t < ntl; t++) hits += gpvo_tribox(c, g->ext2, j->m->tris + (int64_t)tl[t] * 9);
		}
		j->boxTests += ntl * n23;
	}
	j->rayTests = hits;
	return NULL;
}
double gpvo_time_l2_tribox(const gpvo_mesh* m, const gpvo_result* r, int64_t b0, int64_t b1, int nThreads, int64_t* tests)
{
	const gpvo_grid* g = &r->g;
	float* cx = (float*)malloc(g->numDiv[0] * sizeof(float)); float* cy = (float*)malloc(g->numDiv[1] * sizeof(float)); float* cz = (float*)malloc(g->numDiv[2] * sizeof(float));
	gpvo_axis_table(m->bmin, g, 0, cx); gpvo_axis_table(m->bmin, g, 1, cy); gpvo_axis_table(m->bmin, g, 2, cz);
	job proto; memset(&proto, 0, sizeof proto);
	proto.m = m; proto.r = (gpvo_result*)r; proto.cx = cx; proto.cy = cy; proto.cz = cz; proto.b0 = b0; proto.b1 = b1;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	int64_t hits = 0, n = 0;
	run_threads(time_l2_worker, &proto, nThreads, &hits, &n);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	*tests = n;
	free(cx); free(cy); free(cz);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
