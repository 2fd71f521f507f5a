// Compatibility shim for `al2o3_cmath/scalar.h`: scalar min/max/abs helpers and float->half.
// Semantics chosen here are part of the parity contract:
//   Math_Min*/Max*  : plain ternaries (a<b?a:b / a>b?a:b), i.e. NaN/equal pick the second operand
//   Math_Float2Half : IEEE-754 binary32 -> binary16, round-to-nearest-even (== CUDA __float2half_rn)
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#ifdef __cplusplus
extern "C" {
#endif

static inline float Math_MinF(float a, float b) { return a < b ? a : b; }
static inline float Math_MaxF(float a, float b) { return a > b ? a : b; }
static inline float Math_AbsF(float a) { return fabsf(a); }
static inline double Math_MinD(double a, double b) { return a < b ? a : b; }
static inline double Math_MaxD(double a, double b) { return a > b ? a : b; }
static inline double Math_AbsD(double a) { return fabs(a); }
static inline uint32_t Math_MinU32(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t Math_MaxU32(uint32_t a, uint32_t b) { return a > b ? a : b; }

static inline uint16_t Math_Float2Half(float f) {
	uint32_t x;
	memcpy(&x, &f, 4);
	uint32_t const sign = (x >> 16) & 0x8000u;
	uint32_t const absx = x & 0x7FFFFFFFu;
	if (absx >= 0x7F800000u) { // inf / nan
		return (uint16_t) (sign | 0x7C00u | ((absx > 0x7F800000u) ? 0x200u : 0u));
	}
	if (absx >= 0x477FF000u) { // rounds to >= 65520 -> inf
		return (uint16_t) (sign | 0x7C00u);
	}
	if (absx < 0x33000001u) { // <= 2^-25 -> +-0 (ties to even gives 0 at exactly 2^-25)
		return (uint16_t) sign;
	}
	int32_t const e = (int32_t) (absx >> 23) - 127;
	uint32_t m = (absx & 0x7FFFFFu) | 0x800000u;
	uint32_t shift, hexp;
	if (e < -14) { // subnormal half
		shift = (uint32_t) (13 + (-14 - e));
		hexp = 0;
	} else {
		shift = 13;
		hexp = (uint32_t) (e + 15);
	}
	uint32_t const halfway = 1u << (shift - 1);
	uint32_t const rem = m & ((1u << shift) - 1u);
	uint32_t q = m >> shift;
	if (rem > halfway || (rem == halfway && (q & 1u))) q++;
	// for normals q includes the implicit bit (0x400); adding (hexp-1)<<10 folds mantissa carry into exponent
	uint32_t const h = (hexp == 0) ? q : (((hexp - 1) << 10) + q);
	return (uint16_t) (sign | h);
}

static inline float Math_Half2Float(uint16_t h) {
	uint32_t const sign = ((uint32_t) h & 0x8000u) << 16;
	uint32_t const e = (h >> 10) & 0x1Fu;
	uint32_t m = h & 0x3FFu;
	uint32_t out;
	if (e == 0) {
		if (m == 0) out = sign;
		else {
			int s = 0;
			while (!(m & 0x400u)) { m <<= 1; s++; }
			m &= 0x3FFu;
			out = sign | ((uint32_t) (127 - 15 - s + 1) << 23) | (m << 13);
		}
	} else if (e == 31) {
		out = sign | 0x7F800000u | (m << 13);
	} else {
		out = sign | ((e + 127 - 15) << 23) | (m << 13);
	}
	float f;
	memcpy(&f, &out, 4);
	return f;
}

#ifdef __cplusplus
}
#endif
