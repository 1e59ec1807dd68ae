// gpview_b200/csrc/gpv_host.cpp -- host side of the native tier: mesh loaders, grid sizing, output writer.
//
// Mirrors the reference's host semantics for the voxelizer path (compile with -O2 -ffp-contract=off):
//   gpv_load_obj   Object::ReadObject       src/Object.cpp:395-584
//   gpv_load_off   Object::ReadOFFObject    src/Object.cpp:171-317
//   flat layout    Object::CreateFlatTriangleData src/Object.cpp:3496-3527
//   gpv_make_grid  Object::PerformVoxelization    src/Object.cpp:3094-3134
//   gpv_save       Object::SaveVoxelization       src/Object.cpp:2934-3075
// Only what feeds the voxelizer is kept (positions, faces, bbox); normals/texcoords/adjacency of the viewer are not.
#include "../../include/gpview_b200.h"
#include "gpv_internal.h"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

bool read_file(const char* path, std::vector<char>& buf)
{
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	buf.resize(sz > 0 ? (size_t)sz : 0);
	bool ok = sz <= 0 || fread(buf.data(), 1, (size_t)sz, f) == (size_t)sz;
	fclose(f);
	return ok;
}

// split() of src/Utilities.cpp:989-1004 for a one-character delimiter: empty fields are kept, a trailing delimiter adds an
// empty field, an empty string has no fields.
struct Field { const char* p; size_t n; };
void split_fields(const char* s, size_t len, char delim, std::vector<Field>& out)
{
	out.clear();
	size_t i = 0;
	while (i < len) {
		size_t j = i;
		while (j < len && s[j] != delim) j++;
		out.push_back({ s + i, j - i });
		if (j == len) break;
		i = j + 1;
		if (i == len) out.push_back({ s + i, 0 });
	}
}

bool field_to_float(const Field& f, float& v) // std::stof == strtof on the field
{
	char tmp[128];
	if (f.n == 0 || f.n >= sizeof tmp) return false;
	memcpy(tmp, f.p, f.n);
	tmp[f.n] = 0;
	char* end;
	v = strtof(tmp, &end);
	return end != tmp;
}
bool field_to_long(const Field& f, long& v) // std::stoi
{
	char tmp[64];
	if (f.n == 0 || f.n >= sizeof tmp) return false;
	memcpy(tmp, f.p, f.n);
	tmp[f.n] = 0;
	char* end;
	v = strtol(tmp, &end, 10);
	return end != tmp;
}

// bbox padding by 0.001*|diagonal| and maxModelSize (src/Object.cpp:572-583; VectorMagnitude includes/FloatVector.h:323)
void finish_bbox(const float mn[3], const float mx[3], gpv_mesh* m)
{
	float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
	float diag = sqrtf(dx * dx + dy * dy + dz * dz);
	double modelSize = diag;
	float offset = 0.001 * modelSize;
	for (int a = 0; a < 3; a++) {
		m->bbox_max[a] = offset + mx[a];
		m->bbox_min[a] = mn[a] - offset;
	}
	float ex = m->bbox_max[0] - m->bbox_min[0], ey = m->bbox_max[1] - m->bbox_min[1], ez = m->bbox_max[2] - m->bbox_min[2];
	float eyz = ey > ez ? ey : ez;
	m->max_model_size = ex > eyz ? ex : eyz;
}

int export_mesh(std::vector<float>& tris, int64_t nVerts, const float mn[3], const float mx[3], gpv_mesh* out)
{
	out->n_tri = (int64_t)(tris.size() / 9);
	out->n_verts = nVerts;
	out->tris = (float*)malloc(tris.size() * sizeof(float) + 16);
	if (!out->tris) return gpv::fail("out of host memory");
	memcpy(out->tris, tris.data(), tris.size() * sizeof(float));
	finish_bbox(mn, mx, out);
	return 0;
}

} // namespace

extern "C" int gpv_load_obj(const char* path, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	std::vector<char> buf;
	if (!read_file(path, buf)) return gpv::fail(std::string("Unable to open file \"") + path + "\""); // the reference abort()s (:409-413)
	std::vector<float> verts, tris;
	std::vector<Field> bySpace, byTab;
	float mn[3] = { 0, 0, 0 }, mx[3] = { 0, 0, 0 };
	float pt[3] = { 0, 0, 0 }; // declared outside the loop in the reference: short `v` lines keep stale coordinates
	size_t pos = 0, lineNo = 0;
	while (pos < buf.size()) {
		size_t e = pos;
		while (e < buf.size() && buf[e] != '\n') e++;
		if (e == buf.size()) break; // getline reached EOF: `if (!in.good()) break;` drops an unterminated last line (:425)
		const char* line = buf.data() + pos;
		size_t len = e - pos;
		pos = e + 1;
		lineNo++;
		split_fields(line, len, ' ', bySpace);
		split_fields(line, len, '\t', byTab);
		const std::vector<Field>& w = bySpace.size() > byTab.size() ? bySpace : byTab; // :419-423
		if (w.empty()) continue;
		if (w[0].n == 1 && w[0].p[0] == 'v') {
			for (size_t i = 1; i < w.size() && i <= 3; i++)
				if (!field_to_float(w[i], pt[i - 1])) return gpv::fail("OBJ line " + std::to_string(lineNo) + ": bad vertex coordinate (std::stof would throw)");
			if (verts.empty()) for (int a = 0; a < 3; a++) mn[a] = mx[a] = pt[a];
			else for (int a = 0; a < 3; a++) { mn[a] = mn[a] < pt[a] ? mn[a] : pt[a]; mx[a] = mx[a] > pt[a] ? mx[a] : pt[a]; }
			verts.insert(verts.end(), pt, pt + 3);
		} else if (w[0].n == 1 && w[0].p[0] == 'f') {
			long idx[3] = { 0, 0, 0 };
			for (size_t i = 1; i < w.size() && i <= 3; i++) { // "a", "a/b", "a/b/c", "a//c": the vertex index is the first '/' field (:480-505)
				Field f = w[i];
				size_t k = 0;
				while (k < f.n && f.p[k] != '/') k++;
				if (k > 0) f.n = k;
				if (!field_to_long(f, idx[i - 1])) return gpv::fail("OBJ line " + std::to_string(lineNo) + ": bad face index (std::stoi would throw)");
			}
			const long nv = (long)(verts.size() / 3);
			for (int c = 0; c < 3; c++) {
				idx[c] -= 1;
				if (idx[c] < 0 || idx[c] >= nv) return gpv::fail("OBJ line " + std::to_string(lineNo) + ": face index out of range");
			}
			for (int c = 0; c < 3; c++) tris.insert(tris.end(), verts.begin() + idx[c] * 3, verts.begin() + idx[c] * 3 + 3);
		}
	}
	if (verts.empty()) return gpv::fail(std::string("no vertices in ") + path);
	return export_mesh(tris, (int64_t)(verts.size() / 3), mn, mx, out);
}

extern "C" int gpv_load_off(const char* path, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	std::vector<char> buf;
	if (!read_file(path, buf)) return gpv::fail(std::string("Unable to open file \"") + path + "\""); // the reference abort()s (:187-191)
	const char *p = buf.data(), *end = buf.data() + buf.size();
	auto next = [&](Field& f) -> bool { // operator>> tokenisation: skip whitespace, take the non-space run
		while (p < end && isspace((unsigned char)*p)) p++;
		if (p >= end) return false;
		f.p = p;
		while (p < end && !isspace((unsigned char)*p)) p++;
		f.n = (size_t)(p - f.p);
		return true;
	};
	Field f;
	long nV = 0, nF = 0, nE = 0;
	if (!next(f)) return gpv::fail("OFF: empty file");                                         // in >> header
	if (!next(f) || !field_to_long(f, nV) || !next(f) || !field_to_long(f, nF) || !next(f) || !field_to_long(f, nE))
		return gpv::fail("OFF: bad counts line");                                                // in >> v_len >> f_len >> n_len
	if (nV <= 0 || nF <= 0) return gpv::fail("OFF: no vertices or faces");
	std::vector<float> verts((size_t)nV * 3), tris((size_t)nF * 9);
	for (long i = 0; i < nV * 3; i++) if (!next(f) || !field_to_float(f, verts[i])) return gpv::fail("OFF: bad vertex record");
	for (long i = 0; i < nF; i++) {
		long q[4]; // f_count then EXACTLY three indices whatever f_count says (:219-222)
		for (int c = 0; c < 4; c++) if (!next(f) || !field_to_long(f, q[c])) return gpv::fail("OFF: bad face record");
		for (int c = 0; c < 3; c++) {
			if (q[c + 1] < 0 || q[c + 1] >= nV) return gpv::fail("OFF: face index out of range");
			memcpy(&tris[(size_t)i * 9 + c * 3], &verts[(size_t)q[c + 1] * 3], 3 * sizeof(float));
		}
	}
	float mn[3], mx[3]; // bbox over the vertices the triangles reference (:257-266)
	for (int a = 0; a < 3; a++) mn[a] = mx[a] = tris[a];
	for (size_t i = 0; i < tris.size(); i += 3) for (int a = 0; a < 3; a++) {
		float x = tris[i + a];
		mn[a] = mn[a] < x ? mn[a] : x;
		mx[a] = mx[a] > x ? mx[a] : x;
	}
	return export_mesh(tris, nV, mn, mx, out);
}

// main()'s dispatch on the last three characters (src/GPView.cpp:1642-1659)
extern "C" int gpv_load_mesh(const char* path, gpv_mesh* out)
{
	size_t n = strlen(path);
	if (n >= 3) {
		const char* ext = path + n - 3;
		if (!strcmp(ext, "obj") || !strcmp(ext, "OBJ")) return gpv_load_obj(path, out);
		if (!strcmp(ext, "off") || !strcmp(ext, "OFF")) return gpv_load_off(path, out);
	}
	return gpv::fail(std::string("unknown mesh extension: ") + path);
}

extern "C" int gpv_mesh_from_triangles(const float* tris, int64_t n_tri, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	if (n_tri <= 0) return gpv::fail("gpv_mesh_from_triangles: no triangles");
	std::vector<float> t(tris, tris + n_tri * 9);
	float mn[3], mx[3];
	for (int a = 0; a < 3; a++) mn[a] = mx[a] = t[a];
	for (size_t i = 0; i < t.size(); i += 3) for (int a = 0; a < 3; a++) {
		float x = t[i + a];
		mn[a] = mn[a] < x ? mn[a] : x;
		mx[a] = mx[a] > x ? mx[a] : x;
	}
	return export_mesh(t, n_tri * 3, mn, mx, out);
}

extern "C" void gpv_free_mesh(gpv_mesh* m)
{
	if (m && m->tris) { free(m->tris); m->tris = nullptr; m->n_tri = 0; }
}

static int next_div4(int a) { return (a % 4 == 0) ? a : a + (4 - a % 4); } // includes/Utilities.h:313

extern "C" int gpv_make_grid(const float bmin[3], const float bmax[3], float max_model_size, int voxel_count, int voxel_count2, gpv_grid* g)
{
	if (voxel_count <= 0) return gpv::fail("voxel_count must be positive");
	float nominalGridSize = max_model_size / (1.0 * voxel_count);        // :3094
	int n2 = voxel_count2 > 0 ? voxel_count2 : 1;
	g->n2 = n2;
	for (int a = 0; a < 3; a++) {
		int n = int((bmax[a] - bmin[a]) / nominalGridSize);                 // :3098-3100
		if (n == 0) n++;                                                   // :3101-3103
		n = next_div4(n);                                                  // :3104-3106
		g->num_div[a] = n;
		g->grid_size[a] = (bmax[a] - bmin[a]) / (n * 1.0);                  // :3107-3109
		g->grid_size2[a] = g->grid_size[a] / (n2 * 1.0);                    // :3128-3130
		g->ext1[a] = g->grid_size[a] / 2.0;                                 // :2551
		g->ext2[a] = g->grid_size2[a] / 2.0;                                // :2552
		if (!(g->grid_size[a] > 0.f) || n <= 0) return gpv::fail("degenerate bounding box");
	}
	return 0;
}

// ostream << float prints like "%g" (precision 6); file names "Obj" + to_string(objID) + suffix (:2958-2974)
extern "C" int gpv_save(const gpv_mesh* mesh, const gpv_result* res, const gpv_host_streams* h, int obj_id, const char* dir)
{
	const gpv_grid& g = res->grid;
	const bool l2 = h->level2_inout != nullptr;
	std::string prefix = std::string(dir) + "/Obj" + std::to_string(obj_id);
	FILE* f = fopen((prefix + "VoxelConfig.txt").c_str(), "w");
	if (!f) return gpv::fail("Unable to open output file for writing"); // the reference abort()s (:2988-2992)
	fprintf(f, "Obj%d\n", obj_id);
	fprintf(f, "%g\t%g\t%g\n", mesh->bbox_min[0], mesh->bbox_min[1], mesh->bbox_min[2]);
	fprintf(f, "%g\t%g\t%g\n", mesh->bbox_max[0], mesh->bbox_max[1], mesh->bbox_max[2]);
	fprintf(f, "%d\t%d\t%d\n", g.num_div[0], g.num_div[1], g.num_div[2]);
	fprintf(f, "%g\t%g\t%g\n", g.grid_size[0], g.grid_size[1], g.grid_size[2]);
	fprintf(f, "%lld\n%lld\n", (long long)res->l1_inside, (long long)res->l1_boundary);
	if (l2) {
		fprintf(f, "%d\t%d\t%d\n", g.n2, g.n2, g.n2);
		fprintf(f, "%g\t%g\t%g\n", g.grid_size2[0], g.grid_size2[1], g.grid_size2[2]);
		fprintf(f, "%lld\n%lld\n", (long long)res->l2_inside, (long long)res->l2_boundary);
	}
	fclose(f);
	auto dump = [&](const char* name, const void* p, size_t bytes, uint8_t fill) -> bool {
		FILE* o = fopen((prefix + name).c_str(), "wb");
		if (!o) return false;
		if (p) fwrite(p, 1, bytes, o);
		else { std::vector<uint8_t> z(bytes, fill); fwrite(z.data(), 1, bytes, o); } // stream not requested: neutral value
		fclose(o);
		return true;
	};
	const size_t cells = (size_t)res->cells, l2n = (size_t)res->n_boundary * (size_t)res->n23;
	bool ok = dump("Level1InOut.raw", h->level1_inout, cells, 0) && dump("Level1Normal.raw", h->level1_normal, cells * 3, 127);
	if (l2) ok = ok && dump("Level1BoundaryPrefixSum.raw", h->prefix, cells * 4, 0) && dump("Level2InOut.raw", h->level2_inout, l2n, 0) &&
	             dump("Level2Normal.raw", h->level2_normal, l2n * 3, 127);
	return ok ? 0 : gpv::fail("Unable to open output file for writing");
}

// ---- reader of the six-file set.  The reference can only read back one hard-coded 48x64x64 uchar grid
// (Object::ReadRAWObject, src/Object.cpp:319-392); this generalises it by parsing ObjNVoxelConfig.txt (written by
// SaveVoxelization, :3009-3022) for the sizes, so that outputs can be re-loaded, diffed and fed to a 3-D CNN loader.
namespace {
bool read_exact(const std::string& path, void* dst, size_t bytes)
{
	FILE* f = fopen(path.c_str(), "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	bool ok = (size_t)sz == bytes && (bytes == 0 || fread(dst, 1, bytes, f) == bytes);
	fclose(f);
	return ok;
}
}

extern "C" void gpv_free_voxels(gpv_voxel_file* v)
{
	if (!v) return;
	free(v->level1_inout); free(v->level1_normal); free(v->prefix_sum); free(v->level2_inout); free(v->level2_normal);
	v->level1_inout = v->level1_normal = v->level2_inout = v->level2_normal = nullptr;
	v->prefix_sum = nullptr;
}

extern "C" int gpv_load_voxels(const char* dir, int obj_id, gpv_voxel_file* v)
{
	memset(v, 0, sizeof *v);
	const std::string prefix = std::string(dir) + "/Obj" + std::to_string(obj_id);
	FILE* f = fopen((prefix + "VoxelConfig.txt").c_str(), "r");
	if (!f) return gpv::fail("Unable to open " + prefix + "VoxelConfig.txt");
	long long a = 0, b = 0;
	int n = fscanf(f, "%63s %f %f %f %f %f %f %d %d %d %f %f %f %lld %lld", v->name, &v->bbox_min[0], &v->bbox_min[1], &v->bbox_min[2], &v->bbox_max[0],
	               &v->bbox_max[1], &v->bbox_max[2], &v->num_div[0], &v->num_div[1], &v->num_div[2], &v->grid_size[0], &v->grid_size[1], &v->grid_size[2], &a, &b);
	if (n != 15) { fclose(f); return gpv::fail(prefix + "VoxelConfig.txt: malformed Level-1 header"); }
	v->l1_inside = a; v->l1_boundary = b;
	n = fscanf(f, "%d %d %d %f %f %f %lld %lld", &v->num_div2[0], &v->num_div2[1], &v->num_div2[2], &v->grid_size2[0], &v->grid_size2[1], &v->grid_size2[2], &a, &b);
	fclose(f);
	v->has_level2 = n == 8;
	if (v->has_level2) { v->l2_inside = a; v->l2_boundary = b; }
	else if (n > 0) return gpv::fail(prefix + "VoxelConfig.txt: malformed Level-2 block");
	if (v->num_div[0] <= 0 || v->num_div[1] <= 0 || v->num_div[2] <= 0) return gpv::fail(prefix + "VoxelConfig.txt: bad resolution");
	v->cells = (int64_t)v->num_div[0] * v->num_div[1] * v->num_div[2];
	v->n_boundary = v->l1_boundary;
	v->n23 = v->has_level2 ? (int64_t)v->num_div2[0] * v->num_div2[1] * v->num_div2[2] : 0;
	auto grab = [&](const char* suffix, size_t bytes, void** dst, bool required) -> int {
		*dst = malloc(bytes ? bytes : 1);
		if (!*dst) return gpv::fail("out of host memory");
		if (read_exact(prefix + suffix, *dst, bytes)) return 0;
		free(*dst); *dst = nullptr;
		return required ? gpv::fail(prefix + suffix + ": missing or not the size ObjNVoxelConfig.txt implies") : 0;
	};
	int rc = grab("Level1InOut.raw", (size_t)v->cells, (void**)&v->level1_inout, true) || grab("Level1Normal.raw", (size_t)v->cells * 3, (void**)&v->level1_normal, false);
	if (!rc && v->has_level2)
		rc = grab("Level1BoundaryPrefixSum.raw", (size_t)v->cells * 4, (void**)&v->prefix_sum, true) ||
		     grab("Level2InOut.raw", (size_t)(v->n_boundary * v->n23), (void**)&v->level2_inout, true) ||
		     grab("Level2Normal.raw", (size_t)(v->n_boundary * v->n23) * 3, (void**)&v->level2_normal, false);
	if (rc) { gpv_free_voxels(v); return 1; }
	return 0;
}
