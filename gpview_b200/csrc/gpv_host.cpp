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
#include "gpv_parse.h"
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <string>
#include <thread>
#include <vector>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

// Grow-only, uninitialised scratch memory kept per host thread.  A dataset run loads thousands of ~170 KB files per thread; a
// fresh std::vector per file is zero-filled and, above malloc's 128 KB mmap threshold, mapped and unmapped every time: ~140
// page faults per load, as expensive as the parse itself.  Blocks above 64 MB are given back when the load returns.
template <class T>
struct Scratch {
	T* p = nullptr;
	size_t cap = 0;
	T* get(size_t n)
	{
		if (n > cap) {
			free(p);
			cap = n + n / 2 + 64;
			p = (T*)malloc(cap * sizeof(T));
			if (!p) cap = 0;
		}
		return p;
	}
	void trim() { if (cap * sizeof(T) > ((size_t)64 << 20)) { free(p); p = nullptr; cap = 0; } }
	~Scratch() { free(p); }
};

// the whole file into `buf`; false = cannot open / read / allocate
bool read_file(const char* path, Scratch<char>& buf, size_t& size)
{
	size = 0;
	FILE* f = fopen(path, "rb");
	if (!f) return false;
	fseek(f, 0, SEEK_END);
	long sz = ftell(f);
	fseek(f, 0, SEEK_SET);
	bool ok = sz <= 0;
	if (sz > 0 && buf.get((size_t)sz + 1)) { // one byte of slack: the tolerant OBJ reader terminates an unterminated last line
		size = (size_t)sz;
		ok = fread(buf.p, 1, size, f) == size;
	}
	fclose(f);
	return ok;
}

struct Field { const char* p; size_t n; };

// std::stof / std::stoi on a field: same values as strtof / strtol, by the short exact path of gpv_parse.h when there is one
inline bool field_to_float(const Field& f, float& v) { return gpv::parse_float(f.p, f.n, v); }
inline bool field_to_long(const Field& f, long& v) { return gpv::parse_long(f.p, f.n, v); }

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

// The triangle block a gpv_mesh owns: 16 bytes of header (the block's capacity) in front of the floats.  free_tris() parks the
// thread's last block (up to 16 MB) for the thread's next load instead of handing it back to malloc: same page-fault
// argument as Scratch.  Only gpv_free_mesh() releases a mesh's triangles.
thread_local char* g_triParked = nullptr;   // plain thread-locals (no destructor): still valid when a static object releases a
thread_local bool g_triParkingClosed = false; // mesh after the thread's destructors have run -- then nothing is parked any more
struct TriParkingGuard {
	~TriParkingGuard()
	{
		free(g_triParked);
		g_triParked = nullptr;
		g_triParkingClosed = true;
	}
};
thread_local TriParkingGuard g_triGuard;
const size_t kTriHeader = 16;

float* alloc_tris(size_t nTri)
{
	const size_t need = kTriHeader + nTri * 9 * sizeof(float) + 16;
	if (g_triParked) {
		size_t cap;
		memcpy(&cap, g_triParked, sizeof cap);
		if (cap >= need && cap <= 4 * need) {
			char* b = g_triParked;
			g_triParked = nullptr;
			return (float*)(b + kTriHeader);
		}
	}
	const size_t cap = need + need / 8;
	char* b = (char*)malloc(cap);
	if (!b) return nullptr;
	memcpy(b, &cap, sizeof cap);
	return (float*)(b + kTriHeader);
}
void free_tris(float* tris)
{
	char* b = (char*)tris - kTriHeader;
	size_t cap;
	memcpy(&cap, b, sizeof cap);
	if (cap <= ((size_t)16 << 20) && !g_triParkingClosed) {
		(void)&g_triGuard;          // instantiates the guard: its destructor gives the parked block back when the thread ends
		std::swap(b, g_triParked);  // keep the newer block, release the one that was parked
	}
	free(b);
}

// bbox of n floats taken as xyz triples (n > 0)
void bbox_of(const float* xyz, size_t n, float mn[3], float mx[3])
{
	for (int a = 0; a < 3; a++) mn[a] = mx[a] = xyz[a];
	for (size_t i = 0; i < n; i += 3) for (int a = 0; a < 3; a++) {
		const float x = xyz[i + a];
		mn[a] = mn[a] < x ? mn[a] : x;
		mx[a] = mx[a] > x ? mx[a] : x;
	}
}

// `tris` comes from alloc_tris(nTri) and is owned by the mesh from here on
int export_mesh(float* tris, int64_t nTri, int64_t nVerts, const float mn[3], const float mx[3], gpv_mesh* out)
{
	out->n_tri = nTri;
	out->n_verts = nVerts;
	out->tris = tris;
	finish_bbox(mn, mx, out);
	return 0;
}

} // namespace

namespace {
// ------------------------------------------------------------------------------------------------ chunk-parallel parsing
// Both loaders cut the file into chunks at token boundaries (OBJ: lines), parse the chunks on host threads and stitch the
// results in file order, so that a single 10 M-triangle file (0.7 GB of text) loads in seconds instead of half a minute
// (SURVEY.md 8f3).  One chunk on one thread IS the sequential loader: the same code runs either way, and the result --
// every float, the order of the triangles, the error that is reported -- does not depend on the number of chunks
// (tests/test_host_cpu.py cuts the fixtures into 1..7 chunks).  GPV_LOAD_THREADS overrides the thread count (default: one per
// 4 MB of text, at most the hardware concurrency).
int load_threads(size_t bytes)
{
	long t = 0;
	if (const char* e = getenv("GPV_LOAD_THREADS")) t = strtol(e, nullptr, 10);
	if (t <= 0) {
		const size_t hw = std::max(1u, std::thread::hardware_concurrency());
		t = (long)std::min<size_t>(hw, std::max<size_t>(1, bytes >> 22));
	}
	return (int)std::min<long>(t, 256);
}

template <class F>
void run_chunks(int n, F&& f)
{
	if (n <= 1) { f(0); return; }
	std::vector<std::thread> pool;
	for (int k = 1; k < n; k++) pool.emplace_back([&f, k] { f(k); });
	f(0);
	for (auto& t : pool) t.join();
}

struct ObjChunk {
	size_t begin = 0, end = 0;            // [begin, end): whole lines, each ending with '\n'
	size_t nLines = 0;
	std::vector<float> v;                 // 3 per `v` line (missing coordinates are filled in when the chunks are stitched)
	std::vector<unsigned char> vParsed;   // coordinates present on the line (0..3)
	std::vector<long> f;                  // 3 raw indices per `f` line
	std::vector<size_t> fLine, fVertsBefore; // chunk-local line number; `v` lines of this chunk before the face
	size_t errLine = (size_t)-1;          // chunk-local line number of the first parse error
	int errKind = 0;                      // 1 vertex coordinate, 2 face index
	void reset()
	{
		begin = end = nLines = 0;
		v.clear(); vParsed.clear(); f.clear(); fLine.clear(); fVertsBefore.clear();
		errLine = (size_t)-1;
		errKind = 0;
	}
};

// per host thread: what a load needs besides the triangle block it returns
struct LoadScratch {
	Scratch<char> file;
	Scratch<float> verts;
	Scratch<long> faceTok;
	ObjChunk obj;
};
thread_local LoadScratch g_loadScratch;
struct ScratchTrim { // a single huge file must not pin its text for the life of the thread
	LoadScratch& s;
	~ScratchTrim()
	{
		s.file.trim(); s.verts.trim(); s.faceTok.trim();
		if (s.obj.v.capacity() * sizeof(float) > ((size_t)64 << 20)) s.obj = ObjChunk();
	}
};
struct TriGuard { // error returns between alloc_tris() and export_mesh()
	float* p;
	~TriGuard() { if (p) free_tris(p); }
};

void parse_obj_chunk(const char* data, ObjChunk& c)
{
	size_t pos = c.begin;
	while (pos < c.end) {
		const char* line = data + pos;
		const char* le = (const char*)memchr(line, '\n', c.end - pos); // the chunk ends with '\n'
		const size_t len = (size_t)(le - line);
		pos += len + 1;
		const size_t lineNo = c.nLines++;
		if (len == 0 || (line[0] != 'v' && line[0] != 'f')) continue; // only lines whose first field is exactly "v" or "f" matter
		// The reference splits the line on " " and on "\t" and keeps the split with more fields, the tab split on a tie (:419-423).
		// split() (src/Utilities.cpp:989-1004) keeps empty fields and adds one behind a trailing delimiter, so a non-empty line has
		// one field more than it has delimiters: the delimiter counts decide, and the fields are walked where they stand.
		char delim = ' ';
		if (memchr(line, '\t', len)) {
			size_t nSpace = 0, nTab = 0;
			for (const char* q = line; q < le; q++) { nSpace += *q == ' '; nTab += *q == '\t'; }
			delim = nSpace > nTab ? ' ' : '\t';
		}
		if (len > 1 && line[1] != delim) continue; // "vn", "vt", "vp", ... or a first field longer than one character
		const bool isV = line[0] == 'v';
		float pt[3] = { 0, 0, 0 };
		long idx[3] = { 0, 0, 0 };
		unsigned char got = 0;
		const char* cur = line + 1; // at the delimiter that ends the previous field, or at the end of the line
		for (int i = 0; i < 3 && cur < le; i++) {
			const char* fs = cur + 1;
			const char* fe = fs;
			while (fe < le && *fe != delim) fe++;
			if (isV) {
				if (!gpv::parse_float(fs, (size_t)(fe - fs), pt[i])) { c.errLine = lineNo; c.errKind = 1; return; } // std::stof would throw: the load ends here
			} else { // "a", "a/b", "a/b/c", "a//c": the vertex index is the first '/' field (:480-505)
				const char* sl = fs;
				while (sl < fe && *sl != '/') sl++;
				if (!gpv::parse_long(fs, (size_t)((sl > fs ? sl : fe) - fs), idx[i])) { c.errLine = lineNo; c.errKind = 2; return; }
			}
			got++;
			cur = fe;
		}
		if (isV) {
			c.v.insert(c.v.end(), pt, pt + 3);
			c.vParsed.push_back(got);
		} else {
			c.f.insert(c.f.end(), idx, idx + 3);
			c.fLine.push_back(lineNo);
			c.fVertsBefore.push_back(c.vParsed.size());
		}
	}
}

// Number fields of the tolerant readers: the WHOLE field must be the number ("1.5abc", "0.5" as an index, "3/" are errors).  The
// strict readers keep the reference's strtof / strtol prefix semantics; the tolerant ones are lenient about layout, not about digits.
bool whole_long(const char* p, size_t n, long& v)
{
	const char* q = gpv::scan_long(p, p + n, v);
	return q == p + n && n > 0;
}
bool whole_float(const char* p, size_t n, float& v)
{
	if (const char* q = gpv::scan_float(p, p + n, v)) return q == p + n;
	char tmp[128]; // no short exact path (long mantissa, subnormal, inf / nan, ...): the library, which must also consume the whole field
	if (n == 0 || n >= sizeof tmp) return false;
	memcpy(tmp, p, n);
	tmp[n] = 0;
	if (gpv::is_space((unsigned char)tmp[0])) return false;
	char* end;
	v = strtof(tmp, &end);
	return end == tmp + n;
}

// Tolerant twin (GPV_LOAD_TOLERANT, an extension -- SURVEY.md 8f3): what OBJ files in the wild need and the reference's reader
// refuses or misreads.  Fields are runs of non-blank characters (any mix of spaces, tabs, CR); `v` takes its first three numbers
// (a fourth, w or a colour, is ignored; fewer is an error, not a stale coordinate); `f` takes any number >= 3 of vertices
// "a", "a/b", "a/b/c", "a//c" and is fan-triangulated (a0, a_k, a_k+1); negative indices count back from the vertices defined
// so far.  Every triangle becomes one entry of the chunk's face list, so the stitching below is shared with the strict reader.
void parse_obj_chunk_tolerant(const char* data, ObjChunk& c)
{
	auto blank = [](char ch) { return ch == ' ' || ch == '\t' || ch == '\r'; };
	std::vector<long> poly;
	size_t pos = c.begin;
	while (pos < c.end) {
		const char* p = data + pos;
		const char* le = (const char*)memchr(p, '\n', c.end - pos); // the chunk ends with '\n'
		pos += (size_t)(le - p) + 1;
		const size_t lineNo = c.nLines++;
		while (p < le && blank(*p)) p++;
		if (le - p < 2 || (*p != 'v' && *p != 'f') || !blank(p[1])) continue;
		const bool isV = *p == 'v';
		p++;
		float pt[3] = { 0, 0, 0 };
		int got = 0;
		poly.clear();
		bool bad = false;
		for (;;) {
			while (p < le && blank(*p)) p++;
			if (p >= le || *p == '#') break;
			const char* fs = p;
			while (p < le && !blank(*p)) p++;
			if (isV) {
				if (got < 3 && !whole_float(fs, (size_t)(p - fs), pt[got])) { bad = true; break; }
				got++;
			} else {
				const char* sl = fs;
				while (sl < p && *sl != '/') sl++;
				long idx = 0;
				if (!whole_long(fs, (size_t)(sl - fs), idx) || idx == 0) { bad = true; break; }
				poly.push_back(idx);
			}
		}
		if (bad || (isV && got < 3) || (!isV && poly.size() < 3)) { c.errLine = lineNo; c.errKind = isV ? 1 : 2; return; }
		if (isV) {
			c.v.insert(c.v.end(), pt, pt + 3);
			c.vParsed.push_back(3);
		} else {
			for (size_t k = 1; k + 1 < poly.size(); k++) {
				const long tri[3] = { poly[0], poly[k], poly[k + 1] };
				c.f.insert(c.f.end(), tri, tri + 3);
				c.fLine.push_back(lineNo);
				c.fVertsBefore.push_back(c.vParsed.size());
			}
		}
	}
}

// [0, usable) cut into n pieces at line starts
std::vector<size_t> line_cuts(const char* buf, size_t usable, int n)
{
	std::vector<size_t> cut(1, 0);
	for (int k = 1; k < n; k++) {
		size_t p = std::max(cut.back(), usable * (size_t)k / (size_t)n);
		while (p < usable && p > 0 && buf[p - 1] != '\n') p++;
		cut.push_back(std::min(p, usable));
	}
	cut.push_back(usable);
	return cut;
}

} // namespace

static int load_obj_impl(const char* path, gpv_mesh* out, bool tolerant)
{
	memset(out, 0, sizeof *out);
	LoadScratch& S = g_loadScratch;
	ScratchTrim trimOnReturn{ S };
	size_t fileSize = 0;
	if (!read_file(path, S.file, fileSize)) return gpv::fail(std::string("Unable to open file \"") + path + "\""); // the reference abort()s (:409-413)
	const char* buf = S.file.p;
	size_t usable = fileSize; // getline at EOF: `if (!in.good()) break;` drops an unterminated last line (:425)
	if (tolerant && fileSize > 0 && buf[fileSize - 1] != '\n') { S.file.p[fileSize] = '\n'; usable = fileSize + 1; } // ... which the tolerant reader keeps
	while (usable > 0 && buf[usable - 1] != '\n') usable--;
	const int nChunks = load_threads(usable);
	const std::vector<size_t> cut = line_cuts(buf, usable, nChunks);
	std::vector<ObjChunk> many(nChunks > 1 ? (size_t)nChunks : 0);
	ObjChunk* ch = nChunks > 1 ? many.data() : &S.obj; // one chunk: the thread's own, its vectors keep their capacity
	for (int k = 0; k < nChunks; k++) { ch[k].reset(); ch[k].begin = cut[k]; ch[k].end = cut[k + 1]; }
	run_chunks(nChunks, [&](int k) { if (tolerant) parse_obj_chunk_tolerant(buf, ch[k]); else parse_obj_chunk(buf, ch[k]); });

	// ---- stitch in file order.  The sequential reader stops at its first error: everything behind the first chunk with a parse
	// error is ignored, and the earliest error -- parse error or face index out of range -- is the one reported.
	size_t nUsed = (size_t)nChunks, lineBase = 0;
	std::vector<size_t> vertBase((size_t)nChunks + 1, 0), faceBase((size_t)nChunks + 1, 0), firstLine((size_t)nChunks + 1, 0);
	for (size_t k = 0; k < (size_t)nChunks; k++) {
		firstLine[k] = lineBase;
		lineBase += ch[k].nLines;
		vertBase[k + 1] = vertBase[k] + ch[k].vParsed.size();
		faceBase[k + 1] = faceBase[k] + ch[k].fLine.size();
		if (ch[k].errKind) { nUsed = k + 1; break; }
	}
	const size_t nVerts = vertBase[nUsed], nFaces = faceBase[nUsed];
	float* verts = nChunks > 1 ? S.verts.get(nVerts * 3 + 3) : ch[0].v.data(); // one chunk: its vertex array is the vertex array
	if (nChunks > 1 && !verts) return gpv::fail("out of host memory");
	bool shortLine = false;
	for (size_t k = 0; k < nUsed; k++) {
		if (nChunks > 1 && !ch[k].v.empty()) memcpy(&verts[vertBase[k] * 3], ch[k].v.data(), ch[k].v.size() * sizeof(float));
		for (unsigned char g : ch[k].vParsed) shortLine |= g < 3;
	}
	if (shortLine) { // `pt` lives outside the reference's loop: a short `v` line keeps the previous line's trailing coordinates (zero at first)
		float pt[3] = { 0, 0, 0 };
		size_t i = 0;
		for (size_t k = 0; k < nUsed; k++) for (unsigned char g : ch[k].vParsed) {
			for (int a = 0; a < 3; a++) { if (a < g) pt[a] = verts[i * 3 + a]; else verts[i * 3 + a] = pt[a]; }
			i++;
		}
	}
	// faces -> triangles; a face may only name vertices that were defined before its line
	float* tris = alloc_tris(nFaces);
	if (!tris) return gpv::fail("out of host memory");
	TriGuard guard{ tris };
	std::vector<size_t> badFace(nUsed, (size_t)-1);
	run_chunks((int)nUsed, [&](int k) {
		const ObjChunk& c = ch[k];
		for (size_t j = 0; j < c.fLine.size(); j++) {
			const long nv = (long)(vertBase[k] + c.fVertsBefore[j]);
			float* t = &tris[(faceBase[k] + j) * 9];
			for (int q = 0; q < 3; q++) {
				const long raw = c.f[j * 3 + q];
				const long idx = (tolerant && raw < 0) ? nv + raw : raw - 1; // tolerant: -1 is the vertex defined last
				if (idx < 0 || idx >= nv) { badFace[k] = j; return; }
				memcpy(t + q * 3, &verts[(size_t)idx * 3], 3 * sizeof(float));
			}
		}
	});
	for (size_t k = 0; k < nUsed; k++) { // the earliest error in file order
		const size_t faceErr = badFace[k] == (size_t)-1 ? (size_t)-1 : ch[k].fLine[badFace[k]];
		const size_t parseErr = ch[k].errKind ? ch[k].errLine : (size_t)-1;
		if (faceErr == (size_t)-1 && parseErr == (size_t)-1) continue;
		const bool isFace = faceErr < parseErr;
		const std::string where = "OBJ line " + std::to_string(firstLine[k] + std::min(faceErr, parseErr) + 1);
		if (isFace) return gpv::fail(where + ": face index out of range");
		if (tolerant) return gpv::fail(where + (ch[k].errKind == 1 ? ": a vertex needs three numbers" : ": a face needs at least three non-zero integer vertex indices"));
		return gpv::fail(where + (ch[k].errKind == 1 ? ": bad vertex coordinate (std::stof would throw)" : ": bad face index (std::stoi would throw)"));
	}
	if (nVerts == 0) return gpv::fail(std::string("no vertices in ") + path);
	float mn[3], mx[3]; // bbox over ALL `v` lines (:441-450)
	bbox_of(verts, nVerts * 3, mn, mx);
	guard.p = nullptr;
	return export_mesh(tris, (int64_t)nFaces, (int64_t)nVerts, mn, mx, out);
}

extern "C" int gpv_load_obj(const char* path, gpv_mesh* out) { return load_obj_impl(path, out, false); }

namespace {
// operator>> tokens of [begin, end): whitespace-separated runs (chunks are cut at whitespace, so a run never straddles a cut).
// f(k, field) is called with the chunk-local token number; returns the number of tokens.
template <class F>
size_t off_tokens(const char* data, size_t begin, size_t end, F&& f)
{
	const char *p = data + begin, *e = data + end;
	size_t k = 0;
	for (;;) {
		while (p < e && gpv::is_space((unsigned char)*p)) p++;
		if (p >= e) return k;
		Field fld{ p, 0 };
		while (p < e && !gpv::is_space((unsigned char)*p)) p++;
		fld.n = (size_t)(p - fld.p);
		if (!f(k, fld)) return k;
		k++;
	}
}
} // namespace

namespace {
// Pass 2 over one chunk: token number i = base + local is a header token (i < vTok), a coordinate (i < fTok) or a face token
// (i < needTok).  Numbers are converted where they stand (gpv::scan_*): when the conversion ends at whitespace the token is done;
// anything else -- trailing characters, a form without a short exact path -- takes the whole token through the library,
// exactly like the field-based route.  Returns the chunk-local number of the token it stopped at (`bad` = its global number
// when it stopped because of a failed conversion).
size_t off_parse_chunk(const char* data, size_t begin, size_t end, size_t base, size_t vTok, size_t fTok, size_t needTok, float* verts, long* faceTok, size_t& bad)
{
	const char *p = data + begin, *e = data + end;
	size_t local = 0;
	for (;;) {
		while (p < e && gpv::is_space((unsigned char)*p)) p++;
		if (p >= e) return local;
		const size_t i = base + local;
		if (i >= needTok) return local;
		if (i >= vTok) {
			const char* q = i < fTok ? gpv::scan_float(p, e, verts[i - vTok]) : gpv::scan_long(p, e, faceTok[i - fTok]);
			if (q && (q == e || gpv::is_space((unsigned char)*q))) { p = q; local++; continue; }
		}
		const char* s = p;
		while (p < e && !gpv::is_space((unsigned char)*p)) p++;
		if (i >= vTok) {
			const bool ok = i < fTok ? gpv::parse_float(s, (size_t)(p - s), verts[i - vTok]) : gpv::parse_long(s, (size_t)(p - s), faceTok[i - fTok]);
			if (!ok) { bad = i; return local; }
		}
		local++;
	}
}
} // namespace

extern "C" int gpv_load_off(const char* path, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	LoadScratch& S = g_loadScratch;
	ScratchTrim trimOnReturn{ S };
	size_t fileSize = 0;
	if (!read_file(path, S.file, fileSize)) return gpv::fail(std::string("Unable to open file \"") + path + "\""); // the reference abort()s (:187-191)
	const char* buf = S.file.p;
	// pass 1: count the tokens of every chunk (chunks are cut at whitespace) -> the global number of each chunk's first token
	const int nChunks = load_threads(fileSize);
	std::vector<size_t> cut(1, 0);
	for (int k = 1; k < nChunks; k++) {
		size_t p = std::max(cut.back(), fileSize * (size_t)k / (size_t)nChunks);
		while (p < fileSize && !gpv::is_space((unsigned char)buf[p])) p++;
		cut.push_back(p);
	}
	cut.push_back(fileSize);
	std::vector<size_t> base((size_t)nChunks + 1, 0);
	if (nChunks > 1) { // one chunk (every file below 8 MB): its tokens are counted by pass 2 itself
		run_chunks(nChunks, [&](int k) { base[k + 1] = off_tokens(buf, cut[k], cut[k + 1], [](size_t, const Field&) { return true; }); });
		for (int k = 0; k < nChunks; k++) base[k + 1] += base[k];
	}
	// header and counts: tokens 0..3 (in >> header; in >> v_len >> f_len >> n_len)
	Field head[4] = {};
	size_t nHead = 0;
	off_tokens(buf, 0, fileSize, [&](size_t k, const Field& f) { head[k] = f; nHead = k + 1; return k < 3; });
	size_t nTok = nChunks > 1 ? base[nChunks] : nHead;
	long nV = 0, nF = 0, nE = 0;
	if (nTok < 1) return gpv::fail("OFF: empty file");
	if (nTok < 4 || !field_to_long(head[1], nV) || !field_to_long(head[2], nF) || !field_to_long(head[3], nE)) return gpv::fail("OFF: bad counts line");
	if (nV <= 0 || nF <= 0) return gpv::fail("OFF: no vertices or faces");
	// pass 2: token 4 + i (i < 3 nV) is a coordinate; then every face takes FOUR tokens: f_count and exactly three indices whatever
	// f_count says (:219-222).  The sequential reader fails at the first bad or missing token: so does this one.
	const size_t vTok = 4, fTok = 4 + (size_t)nV * 3, needTok = fTok + (size_t)nF * 4;
	// a vertex record takes >= 6 bytes of text and a face record >= 8: counts the file cannot hold end in a missing token below,
	// and must not size the arrays
	if ((size_t)nV > fileSize / 6 + 1) return gpv::fail("OFF: bad vertex record");
	const size_t nFcap = std::min<size_t>((size_t)nF, fileSize / 8 + 2);
	float* verts = S.verts.get((size_t)nV * 3);
	long* faceTok = S.faceTok.get(nFcap * 4);
	float* tris = alloc_tris(nFcap);
	TriGuard guard{ tris };
	if (!verts || !faceTok || !tris) return gpv::fail("out of host memory");
	std::vector<size_t> bad((size_t)nChunks, (size_t)-1); // first bad token (global number) of each chunk
	std::vector<size_t> seen((size_t)nChunks, 0); // tokens of the chunk walked before pass 2 stopped
	run_chunks(nChunks, [&](int k) { seen[k] = off_parse_chunk(buf, cut[k], cut[k + 1], base[k], vTok, fTok, needTok, verts, faceTok, bad[k]); });
	if (nChunks == 1) nTok = bad[0] == (size_t)-1 ? seen[0] : needTok; // ran out of tokens (< needTok) or saw enough; behind a bad token the count no longer matters
	size_t firstBad = nTok < needTok ? nTok : (size_t)-1; // a missing token fails like a bad one
	for (int k = 0; k < nChunks; k++) firstBad = std::min(firstBad, bad[k]);
	if (firstBad != (size_t)-1 && firstBad < fTok) return gpv::fail("OFF: bad vertex record");
	// faces (vertices are complete here); the earliest failure -- bad token or index out of range -- is reported
	std::vector<size_t> badIdx((size_t)nChunks, (size_t)-1);
	const size_t okFaces = firstBad == (size_t)-1 ? (size_t)nF : (firstBad - fTok) / 4;
	run_chunks(nChunks, [&](int k) {
		const size_t j0 = okFaces * (size_t)k / (size_t)nChunks, j1 = okFaces * (size_t)(k + 1) / (size_t)nChunks;
		for (size_t j = j0; j < j1; j++) for (int c = 0; c < 3; c++) {
			const long q = faceTok[j * 4 + 1 + c];
			if (q < 0 || q >= nV) { badIdx[k] = j; return; }
			memcpy(&tris[j * 9 + c * 3], &verts[(size_t)q * 3], 3 * sizeof(float));
		}
	});
	size_t firstBadFace = (size_t)-1;
	for (int k = 0; k < nChunks; k++) firstBadFace = std::min(firstBadFace, badIdx[k]);
	if (firstBadFace != (size_t)-1) return gpv::fail("OFF: face index out of range");
	if (firstBad != (size_t)-1) return gpv::fail("OFF: bad face record");
	float mn[3], mx[3]; // bbox over the vertices the triangles reference (:257-266)
	bbox_of(tris, (size_t)nF * 9, mn, mx);
	guard.p = nullptr;
	return export_mesh(tris, (int64_t)nF, nV, mn, mx, out);
}

namespace {
// Tolerant OFF reader (GPV_LOAD_TOLERANT, an extension): line-aware where the reference's `>>` chain is not.  '#' starts a
// comment; the counts may follow "OFF" on the same line or come on the next one; a vertex line gives its first three numbers
// (trailing colour fields are ignored); a face line gives `n i_1 ... i_n` with any n >= 3, fan-triangulated, and whatever follows
// the n indices (colours) is ignored -- the reference reads exactly three indices whatever n says (src/Object.cpp:219-222), so
// one quad shifts every later record.  Sequential: these are small files by nature.
int load_off_tolerant(const char* path, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	LoadScratch& S = g_loadScratch;
	ScratchTrim trimOnReturn{ S };
	size_t fileSize = 0;
	if (!read_file(path, S.file, fileSize)) return gpv::fail(std::string("Unable to open file \"") + path + "\"");
	const char *p = S.file.p, *end = p + fileSize;
	std::vector<Field> tok;
	size_t lineNo = 0;
	auto next_line = [&]() -> bool { // tokens of the next line that has any (comments stripped)
		while (p < end) {
			const char* le = (const char*)memchr(p, '\n', (size_t)(end - p));
			if (!le) le = end;
			const char* stop = (const char*)memchr(p, '#', (size_t)(le - p));
			if (!stop) stop = le;
			tok.clear();
			for (const char* q = p; q < stop;) {
				while (q < stop && gpv::is_space((unsigned char)*q)) q++;
				if (q >= stop) break;
				const char* b = q;
				while (q < stop && !gpv::is_space((unsigned char)*q)) q++;
				tok.push_back({ b, (size_t)(q - b) });
			}
			p = le < end ? le + 1 : end;
			lineNo++;
			if (!tok.empty()) return true;
		}
		return false;
	};
	auto where = [&]() { return "OFF line " + std::to_string(lineNo); };
	if (!next_line()) return gpv::fail("OFF: empty file");
	if (tok[0].n != 3 || memcmp(tok[0].p, "OFF", 3) != 0) return gpv::fail(where() + ": header is not \"OFF\" (COFF / NOFF / 4OFF variants are not supported)");
	size_t c0 = 1;
	if (tok.size() < 4) { if (!next_line()) return gpv::fail("OFF: bad counts line"); c0 = 0; }
	long nV = 0, nF = 0;
	if (tok.size() < c0 + 2 || !whole_long(tok[c0].p, tok[c0].n, nV) || !whole_long(tok[c0 + 1].p, tok[c0 + 1].n, nF)) return gpv::fail(where() + ": bad counts line");
	if (nV <= 0 || nF <= 0) return gpv::fail("OFF: no vertices or faces");
	if ((size_t)nV > fileSize / 6 + 1 || (size_t)nF > fileSize / 8 + 1) return gpv::fail("OFF: the counts line promises more records than the file can hold");
	float* verts = S.verts.get((size_t)nV * 3);
	if (!verts) return gpv::fail("out of host memory");
	for (long i = 0; i < nV; i++) {
		if (!next_line()) return gpv::fail("OFF: file ends inside the vertex list");
		if (tok.size() < 3 || !whole_float(tok[0].p, tok[0].n, verts[i * 3]) || !whole_float(tok[1].p, tok[1].n, verts[i * 3 + 1]) || !whole_float(tok[2].p, tok[2].n, verts[i * 3 + 2]))
			return gpv::fail(where() + ": a vertex needs three numbers");
	}
	std::vector<long> idx; // 3 vertex indices per triangle
	idx.reserve((size_t)nF * 3);
	for (long f = 0; f < nF; f++) {
		if (!next_line()) return gpv::fail("OFF: file ends inside the face list");
		long n = 0;
		if (!whole_long(tok[0].p, tok[0].n, n) || n < 3 || tok.size() < (size_t)n + 1) return gpv::fail(where() + ": a face needs its vertex count (>= 3) and that many indices");
		long a = 0, b = 0, c = 0;
		for (long k = 0; k < n; k++) {
			long q = 0;
			if (!whole_long(tok[(size_t)k + 1].p, tok[(size_t)k + 1].n, q) || q < 0 || q >= nV) return gpv::fail(where() + ": face index missing, not an integer or out of range");
			if (k == 0) a = q;
			else { b = c; c = q; if (k >= 2) { idx.push_back(a); idx.push_back(b); idx.push_back(c); } }
		}
	}
	const size_t nTri = idx.size() / 3;
	float* tris = alloc_tris(nTri);
	if (!tris) return gpv::fail("out of host memory");
	for (size_t i = 0; i < idx.size(); i++) memcpy(&tris[i * 3], &verts[(size_t)idx[i] * 3], 3 * sizeof(float));
	float mn[3], mx[3]; // bbox over the vertices the triangles reference, like the strict reader (:257-266)
	bbox_of(tris, nTri * 9, mn, mx);
	return export_mesh(tris, (int64_t)nTri, nV, mn, mx, out);
}
} // namespace

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

extern "C" int gpv_load_mesh_ex(const char* path, unsigned flags, gpv_mesh* out)
{
	if (!(flags & GPV_LOAD_TOLERANT)) return gpv_load_mesh(path, out);
	size_t n = strlen(path);
	if (n >= 3) {
		const char* ext = path + n - 3;
		if (!strcmp(ext, "obj") || !strcmp(ext, "OBJ")) return load_obj_impl(path, out, true);
		if (!strcmp(ext, "off") || !strcmp(ext, "OFF")) return load_off_tolerant(path, out);
	}
	return gpv::fail(std::string("unknown mesh extension: ") + path);
}

extern "C" int gpv_mesh_from_triangles(const float* tris, int64_t n_tri, gpv_mesh* out)
{
	memset(out, 0, sizeof *out);
	if (n_tri <= 0) return gpv::fail("gpv_mesh_from_triangles: no triangles");
	float* t = alloc_tris((size_t)n_tri);
	if (!t) return gpv::fail("out of host memory");
	memcpy(t, tris, (size_t)n_tri * 9 * sizeof(float));
	float mn[3], mx[3];
	bbox_of(t, (size_t)n_tri * 9, mn, mx);
	return export_mesh(t, n_tri, n_tri * 3, mn, mx, out);
}

extern "C" void gpv_free_mesh(gpv_mesh* m)
{
	if (m && m->tris) { free_tris(m->tris); m->tris = nullptr; m->n_tri = 0; }
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
	return gpv_save_streams(mesh, res, h, obj_id, dir, 0);
}

extern "C" int gpv_save_streams(const gpv_mesh* mesh, const gpv_result* res, const gpv_host_streams* h, int obj_id, const char* dir, int omit_absent)
{
	if (!mesh || !res || !h || !dir) return gpv::fail("gpv_save: null argument");
	const gpv_grid& g = res->grid;
	const bool l2 = h->level2_inout != nullptr;
	std::string prefix = std::string(dir) + "/Obj" + std::to_string(obj_id);
	// ObjNVoxelConfig.txt is what marks a set as complete (gpv_voxelize_batch's skip_existing, gpv_load_voxels, the Dataset): it is
	// written LAST, to a temporary name, and renamed into place only after every stream has been written and closed.  A run that is
	// killed or hits a full disk leaves no config behind, so a restart recomputes the model instead of skipping a truncated set.
	const std::string cfgPath = prefix + "VoxelConfig.txt", cfgTmp = cfgPath + ".tmp";
	unlink(cfgPath.c_str()); // a stale config of an earlier run must not vouch for the streams that are about to be rewritten
	// The streams: small sets (a dataset model: a few MB) are written by the calling thread; large ones (cessna 256 / 16 with
	// normals: 0.9 GB) are cut into 8 MB pieces written with pwrite() by a few threads, so that the five files fill side by side.
	// (Buffered writes to ONE file are serialised by the kernel's inode lock: the largest stream -- Level2Normal, 3/4 of the bytes
	// -- still goes at one thread's ~2.4 GB/s; measured 0.39 -> 0.36 s for the 915 MB set in the build container.)  A stream that
	// was not computed is written as its neutral value (or left out: omit_absent).  Any short write or failed close fails the
	// call: a full disk must not pass for a saved model.
	struct Stream { const char* name; const void* p; size_t bytes; uint8_t fill; int fd; };
	const size_t cells = (size_t)res->cells, l2n = (size_t)res->n_boundary * (size_t)res->n23;
	Stream st[5] = { { "Level1InOut.raw", h->level1_inout, cells, 0, -1 }, { "Level1Normal.raw", h->level1_normal, cells * 3, 127, -1 },
		             { "Level1BoundaryPrefixSum.raw", h->prefix, cells * 4, 0, -1 }, { "Level2InOut.raw", h->level2_inout, l2n, 0, -1 },
		             { "Level2Normal.raw", h->level2_normal, l2n * 3, 127, -1 } };
	const int nStreams = l2 ? 5 : 2;
	const size_t kPiece = (size_t)8 << 20;
	struct Piece { int s; size_t off, len; };
	std::vector<Piece> pieces;
	std::string failed;
	size_t total = 0;
	for (int k = 0; k < nStreams && failed.empty(); k++) {
		if (!st[k].p && omit_absent) continue; // not computed, not written
		st[k].fd = open((prefix + st[k].name).c_str(), O_WRONLY | O_CREAT | O_TRUNC | O_CLOEXEC, 0666);
		if (st[k].fd < 0) { failed = "Unable to open output file for writing: " + prefix + st[k].name; break; } // the reference abort()s (:2988-2992)
		for (size_t off = 0; off < st[k].bytes; off += kPiece) pieces.push_back({ k, off, std::min(kPiece, st[k].bytes - off) });
		total += st[k].bytes;
	}
	std::atomic<size_t> next(0);
	std::atomic<int> badStream(-1);
	auto work = [&]() {
		std::vector<uint8_t> filler;
		for (;;) {
			const size_t i = next.fetch_add(1);
			if (i >= pieces.size() || badStream.load() >= 0) return;
			const Piece& pc = pieces[i];
			const Stream& sm = st[pc.s];
			const uint8_t* src = (const uint8_t*)sm.p + pc.off;
			if (!sm.p) {
				if (filler.size() < pc.len || filler[0] != sm.fill) filler.assign(std::max(filler.size(), pc.len), sm.fill);
				src = filler.data();
			}
			for (size_t done = 0; done < pc.len;) {
				const ssize_t w = pwrite(sm.fd, src + done, pc.len - done, (off_t)(pc.off + done));
				if (w < 0 && errno == EINTR) continue;
				if (w <= 0) { int none = -1; badStream.compare_exchange_strong(none, pc.s); return; }
				done += (size_t)w;
			}
		}
	};
	if (failed.empty()) {
		const int nThreads = total < ((size_t)32 << 20) ? 1 : (int)std::min<size_t>({ (size_t)8, (size_t)std::max(1u, std::thread::hardware_concurrency()), pieces.size() });
		run_chunks(nThreads, [&](int) { work(); });
		if (badStream.load() >= 0) failed = "write error on " + prefix + st[badStream.load()].name;
	}
	for (int k = 0; k < nStreams; k++)
		if (st[k].fd >= 0 && close(st[k].fd) != 0 && failed.empty()) failed = "write error on " + prefix + st[k].name;
	if (!failed.empty()) return gpv::fail(failed);
	FILE* f = fopen(cfgTmp.c_str(), "w");
	if (!f) return gpv::fail("Unable to open output file for writing: " + cfgPath); // the reference abort()s (:2988-2992)
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
	const bool cfgBad = ferror(f) != 0;
	if (fclose(f) != 0 || cfgBad) { unlink(cfgTmp.c_str()); return gpv::fail("write error on " + cfgPath); }
	if (rename(cfgTmp.c_str(), cfgPath.c_str()) != 0) { unlink(cfgTmp.c_str()); return gpv::fail("cannot rename " + cfgTmp + " into place"); }
	return 0;
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

// The two-level result as ONE dense grid at the effective resolution (nx*n2, ny*n2, nz*n2), z-major like every other stream:
// what a 3-D CNN consumes (README.md:22-27 of the reference).  An outside / inside Level-1 cell becomes n2^3 voxels of its
// state, a boundary cell its Level-2 block (Level2InOut index = prefix[cell]*n2^3 + r*n2^2 + q*n2 + p, cu:466-469 / :499).
// States stay in the file encoding (0 / 127 / 254).  Rows of n2 voxels are copied at a time; z-layers are dealt to host threads.
extern "C" int gpv_expand_dense(const uint8_t* level1_inout, const int32_t* prefix, const uint8_t* level2_inout, const int num_div[3], int n2,
                                int64_t n_boundary, uint8_t* out, int64_t out_bytes)
{
	if (!level1_inout || !prefix || !level2_inout || !num_div || !out || n2 < 1) return gpv::fail("gpv_expand_dense: bad argument");
	const int64_t nx = num_div[0], ny = num_div[1], nz = num_div[2], n23 = (int64_t)n2 * n2 * n2;
	const int64_t X = nx * n2, Y = ny * n2, Z = nz * n2;
	if (nx < 1 || ny < 1 || nz < 1 || out_bytes < X * Y * Z) return gpv::fail("gpv_expand_dense: output buffer smaller than (nx*n2)*(ny*n2)*(nz*n2) bytes");
	std::atomic<int64_t> next(0);
	std::atomic<bool> bad(false);
	const int threads = (int)std::min<int64_t>({ (int64_t)16, (int64_t)std::max(1u, std::thread::hardware_concurrency()), nz, std::max<int64_t>(1, X * Y * Z >> 22) });
	run_chunks(threads, [&](int) {
		for (;;) {
			const int64_t z = next.fetch_add(1);
			if (z >= nz) return;
			for (int64_t y = 0; y < ny; y++) {
				const uint8_t* row = level1_inout + (z * ny + y) * nx;
				for (int64_t x = 0; x < nx;) {
					if (row[x] != 254) { // a run of cells of one plain state: one memset per dense row
						int64_t x1 = x + 1;
						while (x1 < nx && row[x1] == row[x]) x1++;
						for (int r = 0; r < n2; r++) for (int q = 0; q < n2; q++)
							memset(out + ((z * n2 + r) * Y + (y * n2 + q)) * X + x * n2, row[x], (size_t)((x1 - x) * n2));
						x = x1;
						continue;
					}
					const int64_t b = prefix[(z * ny + y) * nx + x];
					if (b < 0 || b >= n_boundary) { bad = true; return; } // streams that do not belong together
					const uint8_t* block = level2_inout + b * n23;
					uint8_t* dst0 = out + ((z * n2) * Y + y * n2) * X + x * n2;
					switch (n2) { // fixed-size copies for the usual resolutions: the compiler turns them into single moves
					case 2: for (int r = 0; r < 2; r++) for (int q = 0; q < 2; q++) memcpy(dst0 + (r * Y + q) * X, block + (r * 2 + q) * 2, 2); break;
					case 4: for (int r = 0; r < 4; r++) for (int q = 0; q < 4; q++) memcpy(dst0 + (r * Y + q) * X, block + (r * 4 + q) * 4, 4); break;
					case 8: for (int r = 0; r < 8; r++) for (int q = 0; q < 8; q++) memcpy(dst0 + (r * Y + q) * X, block + (r * 8 + q) * 8, 8); break;
					case 16: for (int r = 0; r < 16; r++) for (int q = 0; q < 16; q++) memcpy(dst0 + (r * Y + q) * X, block + (r * 16 + q) * 16, 16); break;
					default: for (int r = 0; r < n2; r++) for (int q = 0; q < n2; q++) memcpy(dst0 + (r * Y + q) * X, block + ((int64_t)r * n2 + q) * n2, (size_t)n2);
					}
					x++;
				}
			}
		}
	});
	return bad ? gpv::fail("gpv_expand_dense: a boundary cell's prefix sum points outside the Level-2 stream") : 0;
}

extern "C" void gpv_free_voxels(gpv_voxel_file* v)
{
	if (!v) return;
	free(v->level1_inout); free(v->level1_normal); free(v->prefix_sum); free(v->level2_inout); free(v->level2_normal);
	v->level1_inout = v->level1_normal = v->level2_inout = v->level2_normal = nullptr;
	v->prefix_sum = nullptr;
}

// ObjNVoxelConfig.txt -> the scalar fields of gpv_voxel_file (no stream is touched)
static int read_voxel_config(const std::string& prefix, gpv_voxel_file* v)
{
	memset(v, 0, sizeof *v);
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
	return 0;
}

// Is the set ObjN* in `dir` complete?  The config parses and every stream it implies is there with exactly the size it implies
// (the normal streams may be absent -- GPV_SAVE_COMPUTED_ONLY -- but not truncated).  Nothing is read but the config: this is the
// restart test of gpv_voxelize_batch (skip_existing) and of the Dataset's directory listing.
extern "C" int gpv_check_voxels(const char* dir, int obj_id)
{
	if (!dir) return gpv::fail("gpv_check_voxels: null directory");
	gpv_voxel_file v;
	const std::string prefix = std::string(dir) + "/Obj" + std::to_string(obj_id);
	if (read_voxel_config(prefix, &v)) return 1;
	auto sized = [&](const char* suffix, int64_t bytes, bool required) -> int {
		struct stat sb;
		if (stat((prefix + suffix).c_str(), &sb) != 0) return required ? gpv::fail(prefix + suffix + ": missing") : 0;
		return (int64_t)sb.st_size == bytes ? 0 : gpv::fail(prefix + suffix + ": not the size ObjNVoxelConfig.txt implies");
	};
	if (sized("Level1InOut.raw", v.cells, true) || sized("Level1Normal.raw", v.cells * 3, false)) return 1;
	if (v.has_level2 && (sized("Level1BoundaryPrefixSum.raw", v.cells * 4, true) || sized("Level2InOut.raw", v.n_boundary * v.n23, true) ||
	                     sized("Level2Normal.raw", v.n_boundary * v.n23 * 3, false)))
		return 1;
	return 0;
}

extern "C" int gpv_load_voxels(const char* dir, int obj_id, gpv_voxel_file* v)
{
	const std::string prefix = std::string(dir) + "/Obj" + std::to_string(obj_id);
	if (read_voxel_config(prefix, v)) return 1;
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
