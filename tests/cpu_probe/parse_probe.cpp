// tests/cpu_probe/parse_probe.cpp -- TEST-ONLY: the loaders' number scanners (gpview_b200/csrc/gpv_parse.h) against the C
// library they stand in for.  Fields are generated here (xorshift, seeded) because the interesting ones -- decimal strings
// next to the midpoint of two adjacent floats, 17-20 digit mantissas, subnormals, overflow, hexadecimal, garbage tails --
// are cheap to make in C and the comparison runs at tens of millions of fields per minute.  Never shipped.
#include "../../gpview_b200/csrc/gpv_parse.h"
#include <cmath>
#include <cstdio>
#include <cstring>

namespace {
struct Rng {
	uint64_t s;
	uint64_t next() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }
	uint32_t below(uint32_t n) { return (uint32_t)((next() >> 11) % n); }
};

float random_float(Rng& r)
{
	uint32_t b = (uint32_t)r.next();
	if ((b & 0x7F800000u) == 0x7F800000u) b &= ~0x00800000u; // no inf / nan
	float f;
	memcpy(&f, &b, 4);
	return f;
}

// one field of the given kind into out (NUL-terminated), returns its length
int make_field(Rng& r, int kind, char* out)
{
	switch (kind) {
	case 0: return sprintf(out, "%.9g", (double)random_float(r)); // what the mesh writers print
	case 1: { // moderate coordinates with 1..12 decimals, optional sign, as modelling tools write them
		const int dec = (int)r.below(12) + 1;
		const double x = ((double)(r.next() >> 11) / 9007199254740992.0 - 0.5) * pow(10.0, (double)r.below(8) - 2);
		return sprintf(out, "%.*f", dec, x);
	}
	case 2: { // the midpoint of two adjacent floats, printed exactly or cut / nudged in the last digits
		float f = random_float(r);
		if (r.below(4)) { uint32_t b; memcpy(&b, &f, 4); b = (b & 0x807FFFFFu) | ((uint32_t)(100 + r.below(60)) << 23); memcpy(&f, &b, 4); } // near 1
		const double mid = ((double)f + (double)nextafterf(f, f < 0 ? -INFINITY : INFINITY)) / 2;
		const int digits = 8 + (int)r.below(18);
		int n = sprintf(out, "%.*e", digits, mid);
		if (r.below(2)) { // nudge one mantissa digit
			int k = 2 + (int)r.below((uint32_t)digits);
			if (out[k] >= '0' && out[k] <= '9') out[k] = (char)('0' + r.below(10));
		}
		return n;
	}
	case 3: { // digit soup: optional sign, up to 24 integer and 24 fraction digits, optional exponent
		int n = 0;
		if (r.below(3) == 0) out[n++] = r.below(2) ? '-' : '+';
		int ni = (int)r.below(r.below(4) ? 4 : 25), nf = (int)r.below(r.below(4) ? 12 : 25);
		for (int i = 0; i < ni; i++) out[n++] = (char)('0' + (r.below(5) ? r.below(10) : 0));
		if (r.below(8)) { out[n++] = '.'; for (int i = 0; i < nf; i++) out[n++] = (char)('0' + (r.below(5) ? r.below(10) : 0)); }
		if (r.below(3) == 0) {
			out[n++] = r.below(2) ? 'e' : 'E';
			if (r.below(2)) out[n++] = r.below(2) ? '-' : '+';
			int ne = (int)r.below(4);
			for (int i = 0; i < ne; i++) out[n++] = (char)('0' + r.below(r.below(3) ? 4 : 10));
		}
		out[n] = 0;
		return n;
	}
	case 4: { // garbage from the alphabet of numbers, and tails behind valid numbers
		static const char A[] = "0123456789012345678901234567890123456789+-..eExXinfatyINFNA \t\r/pP";
		int n = (int)r.below(14) + 1;
		for (int i = 0; i < n; i++) out[i] = A[r.below(sizeof A - 1)];
		out[n] = 0;
		return n;
	}
	default: { // integers: indices, signs, leading zeros, overflow
		int n = 0;
		if (r.below(4) == 0) out[n++] = r.below(2) ? '-' : '+';
		int nd = (int)r.below(r.below(6) ? 8 : 24);
		for (int i = 0; i < nd; i++) out[n++] = (char)('0' + r.below(10));
		if (r.below(6) == 0) { static const char T[] = "/ .e-x\t"; out[n++] = T[r.below(sizeof T - 1)]; out[n++] = (char)('0' + r.below(10)); }
		out[n] = 0;
		return n;
	}
	}
}

bool same_float(float a, float b)
{
	if (a != a || b != b) return (a != a) == (b != b);
	uint32_t x, y;
	memcpy(&x, &a, 4);
	memcpy(&y, &b, 4);
	return x == y;
}
} // namespace

extern "C" {
// n fields of `kind`; returns the number of disagreements with strtof (value bits, success flag, end of the consumed prefix
// when the scanner took its short path) and copies the first offending field into `first` (>= 64 bytes).  `fast` counts
// the fields that took the short path.
int64_t probe_parse_float_fuzz(uint64_t seed, int64_t n, int kind, char* first, int64_t* fast)
{
	Rng r{ seed * 0x9E3779B97F4A7C15ull + 0x1234567ull };
	char s[128];
	int64_t bad = 0;
	*fast = 0;
	for (int64_t i = 0; i < n; i++) {
		const int len = make_field(r, kind, s);
		char* end;
		const float want = strtof(s, &end);
		const bool wantOk = end != s;
		float got = 12345.0f, got2 = 54321.0f;
		const bool gotOk = gpv::parse_float(s, (size_t)len, got);
		const char* q = gpv::scan_float(s, s + len, got2);
		bool ok = gotOk == wantOk && (!wantOk || same_float(got, want));
		if (q) { (*fast)++; ok = ok && q == end && same_float(got2, want); }
		if (!ok && bad++ == 0) { strncpy(first, s, 63); first[63] = 0; }
	}
	return bad;
}
int64_t probe_parse_long_fuzz(uint64_t seed, int64_t n, int kind, char* first, int64_t* fast)
{
	Rng r{ seed * 0x9E3779B97F4A7C15ull + 0x7654321ull };
	char s[128];
	int64_t bad = 0;
	*fast = 0;
	for (int64_t i = 0; i < n; i++) {
		const int len = make_field(r, kind, s);
		char* end;
		const long want = strtol(s, &end, 10);
		const bool wantOk = end != s;
		long got = 777, got2 = 888;
		const bool gotOk = gpv::parse_long(s, (size_t)len, got);
		const char* q = gpv::scan_long(s, s + len, got2);
		bool ok = gotOk == wantOk && (!wantOk || got == want);
		if (q) { (*fast)++; ok = ok && q == end && got2 == want; }
		if (!ok && bad++ == 0) { strncpy(first, s, 63); first[63] = 0; }
	}
	return bad;
}
// one field, for the hand-written cases: returns 1/0 (converted or not), value through *v
int probe_parse_float(const char* s, int64_t n, float* v) { return gpv::parse_float(s, (size_t)n, *v); }
int probe_parse_long(const char* s, int64_t n, long* v) { return gpv::parse_long(s, (size_t)n, *v); }
uint32_t probe_eight_digits(const char* s) { uint64_t x; memcpy(&x, s, 8); return gpv::eight_digits(x) ? gpv::eight_digits_value(x) : 0xFFFFFFFFu; }
}
