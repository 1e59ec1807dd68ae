// gpview_b200/csrc/gpv_parse.h -- number fields of the mesh loaders (host only).
//
// The reference converts every coordinate with std::stof (OBJ, src/Object.cpp:433) or istream >> float (OFF, :202), i.e.
// strtof, and every index with std::stoi / >> int, i.e. strtol.  Dataset generation (BASELINE.json config 5) is bound by
// exactly these two calls: ~27 k fields per 5 k-triangle model at ~100 ns each.  parse_float()/parse_long() return the SAME
// value as strtof()/strtol() on the field -- they take a short exact path when one exists and call the C library otherwise:
//
//   float: sign, <= 19 decimal digits w, decimal exponent e with |e| <= 22 and w <= 2^53.  Then w and 10^|e| are
//   exact doubles and ONE double multiplication or division gives d = RN_double(x) (Clinger's fast path).  Rounding is
//   monotone, so RN_float(d) == RN_float(x) unless d sits exactly on the midpoint of two adjacent floats (low 29 mantissa
//   bits == 1000...0), where d may hide on which side x was: those, and anything outside the normal float range, go to
//   strtof.  Hexadecimal floats, inf/nan, leading blanks, longer mantissas or exponents go to strtof as well.
//
// tests/test_host_cpu.py compares both functions with strtof/strtol on millions of random and adversarial fields
// (through tests/cpu_probe).  Compile with -ffp-contract=off (no effect here, one operation) and the default rounding mode.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>

namespace gpv {

// isspace() of the "C" locale without the locale lookup
inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// strtof on the field [p, p+n) (copied and terminated: fields are not NUL-terminated in the file buffer)
inline bool parse_float_libc(const char* p, size_t n, float& v)
{
	char tmp[128];
	if (n == 0 || n >= sizeof tmp) return false;
	memcpy(tmp, p, n);
	tmp[n] = 0;
	char* end;
	v = strtof(tmp, &end);
	return end != tmp;
}
inline bool parse_long_libc(const char* p, size_t n, long& v)
{
	char tmp[64];
	if (n == 0 || n >= sizeof tmp) return false;
	memcpy(tmp, p, n);
	tmp[n] = 0;
	char* end;
	v = strtol(tmp, &end, 10);
	return end != tmp;
}

// ---- eight ASCII digits at a time (little-endian loads; x86-64 and aarch64 hosts)
inline bool eight_digits(uint64_t x) // every byte in '0'..'9'
{
	return (((x + 0x4646464646464646ull) | (x - 0x3030303030303030ull)) & 0x8080808080808080ull) == 0;
}
inline uint32_t eight_digits_value(uint64_t x) // first byte in memory = most significant digit
{
	x -= 0x3030303030303030ull;
	x = x * 10 + (x >> 8);                                            // pairs: bytes 0, 2, 4, 6 hold two-digit values
	const uint64_t lo = (x & 0x000000FF000000FFull) * (100ull + (1000000ull << 32));
	const uint64_t hi = ((x >> 16) & 0x000000FF000000FFull) * (1ull + (10000ull << 32));
	return (uint32_t)((lo + hi) >> 32);
}

// Prefix scanners.  On success they return the end of the consumed prefix -- what strtof / strtol would report in endptr --
// and the library's value; nullptr means "no short exact path here": the caller converts the field with the library.
inline const char* scan_float(const char* s, const char* e, float& v)
{
	static const double P10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
	bool neg = false;
	if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
	uint64_t w = 0;
	int nd = 0, exp10 = 0; // digits taken into w (leading zeros included, at most 19: w < 10^19 < 2^64); decimal exponent
	while (s < e && (unsigned)(*s - '0') < 10u) {
		if (nd == 19) return nullptr;
		w = w * 10 + (uint64_t)(*s - '0');
		nd++;
		s++;
	}
	if (s < e && (*s == 'x' || *s == 'X')) return nullptr; // hexadecimal float
	if (s < e && *s == '.') {
		s++;
		const int before = nd;
		while (e - s >= 8 && nd <= 11) {
			uint64_t x;
			memcpy(&x, s, 8);
			if (!eight_digits(x)) break;
			w = w * 100000000ull + eight_digits_value(x);
			nd += 8;
			s += 8;
		}
		while (s < e && (unsigned)(*s - '0') < 10u) {
			if (nd == 19) return nullptr;
			w = w * 10 + (uint64_t)(*s - '0');
			nd++;
			s++;
		}
		exp10 = before - nd;
	}
	if (nd == 0) return nullptr; // blanks, inf, nan, garbage: the library decides
	if (s < e && (*s == 'e' || *s == 'E')) { // consumed only when at least one exponent digit follows
		const char* t = s + 1;
		bool eneg = false;
		if (t < e && (*t == '-' || *t == '+')) { eneg = *t == '-'; t++; }
		if (t < e && (unsigned)(*t - '0') < 10u) {
			int x = 0;
			while (t < e && (unsigned)(*t - '0') < 10u) {
				if (x > 9999) return nullptr;
				x = x * 10 + (*t - '0');
				t++;
			}
			exp10 += eneg ? -x : x;
			s = t;
		}
	}
	if (w == 0) { v = neg ? -0.0f : 0.0f; return s; }
	if (w > (1ull << 53) || exp10 < -22 || exp10 > 22) return nullptr;
	double d = (double)w;
	d = exp10 < 0 ? d / P10[-exp10] : d * P10[exp10];
	if (!(d >= 1.17549435082228750797e-38 && d <= 3.4028234e38)) return nullptr; // subnormal or near overflow
	uint64_t bits;
	memcpy(&bits, &d, sizeof bits);
	if ((bits & 0x1FFFFFFFull) == 0x10000000ull) return nullptr; // a float midpoint: the double may hide the side
	const float f = (float)d;
	v = neg ? -f : f;
	return s;
}

inline const char* scan_long(const char* s, const char* e, long& v)
{
	bool neg = false;
	if (s < e && (*s == '-' || *s == '+')) { neg = *s == '-'; s++; }
	const char* d0 = s;
	uint64_t w = 0;
	while (s < e && (unsigned)(*s - '0') < 10u) {
		if (s - d0 == 18) return nullptr; // may overflow: strtol saturates
		w = w * 10 + (uint64_t)(*s - '0');
		s++;
	}
	if (s == d0) return nullptr; // leading blanks or garbage
	v = neg ? -(long)w : (long)w;
	return s;
}

// std::stof / std::stoi on the field [p, p+n).  false = no conversion could be performed (std::stof would throw); the longest
// valid prefix counts, like strtof.
inline bool parse_float(const char* p, size_t n, float& v) { return scan_float(p, p + n, v) ? true : parse_float_libc(p, n, v); }
inline bool parse_long(const char* p, size_t n, long& v) { return scan_long(p, p + n, v) ? true : parse_long_libc(p, n, v); }

} // namespace gpv
