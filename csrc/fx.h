// fx.h — the arithmetic of the lean kernel's fixed-point double sums (FX, agg_direct.cuh), kept free of device intrinsics so that the
// same functions run in the kernel and in the host-side check (tests/cpp/fx_check.cpp, tests/test_fx_limbs.py).
//
// A slot's sum is a 96-bit two's-complement integer {ext, mid, hi} (32-bit limbs, ext lowest) in units of 2^-(F+32); values are
// scaled by 2^F (exact) and rounded to the integer grid of the range they fall in:
//   FX_MAIN  2^(M-15) <= |x * 2^F| < 2^M     : round(x * 2^F) added to {mid, hi}
//   FX_FINE  2^(M-47) <= |x * 2^F| < 2^(M-15): round(x * 2^(F+32)) added to {ext, mid, hi}
//   FX_ZERO  x == 0                           : nothing
//   FX_EXACT everything else (beyond the range, denormal products, Inf, NaN): the caller adds x as a double elsewhere
// fx_lo = 1023 + M - 47 is the biased exponent of the fine range's floor.
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#if defined(__CUDACC__)
#define BK_HD __host__ __device__ __forceinline__
#else
#define BK_HD inline
#endif

namespace bk {

constexpr int FX_MAIN_BINADES = 15;   // width of the two-limb range below 2^M
constexpr int FX_FINE_BINADES = 32;   // width of the three-limb range below that
constexpr int FX_MARGIN = 2;          // binades of head room above the largest sampled value
enum : int { FX_ZERO = 0, FX_MAIN = 1, FX_FINE = 2, FX_EXACT = 3 };

BK_HD uint32_t fx_hi_word(double y) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(y);
#else
    uint64_t b; memcpy(&b, &y, 8); return (uint32_t)(b >> 32);
#endif
}
BK_HD long long fx_round(double y) {   // round to nearest, ties to even; |y| < 2^63
#if defined(__CUDA_ARCH__)
    return __double2ll_rn(y);
#else
    return llrint(y);
#endif
}
// M: magnitude bits a single value may have so that `rows` additions into one slot cannot overflow 63 bits
// The caller ORs FX_MIN_ROWS into `rows`, so M never exceeds FX_MAX_M = 45: a fine value is rounded after a further scaling by 2^32, and
// |x * 2^(F+32)| < 2^(M - 15 + 32) must stay below 2^63.  (The cap is applied to the row count, not as a min() on the result: with the
// min() ptxas allocated the loop's registers differently — 24 instead of 12 bytes of spills — and the kernel measured 6 % slower.)
constexpr uint64_t FX_MIN_ROWS = (uint64_t)1 << 16;
constexpr int FX_MAX_M = 45;
BK_HD int fx_magnitude_bits(uint64_t rows) {
    int h = 0;
    while (h < 62 && (rows >> h) != 0) h++;   // rows < 2^h
    return 62 - h;
}
// F: the scale's exponent from the largest biased exponent sampled (values below 2^(emax - 1022)); clamped so that 2^F is a normal double
BK_HD int fx_scale_exp(int M, uint32_t emax) {
    const int F = M - ((int)emax - 1022 + FX_MARGIN);
    return F > 1000 ? 1000 : (F < -1000 ? -1000 : F);
}
BK_HD double fx_pow2(int e) {   // 2^e for -1022 <= e <= 1023
    const uint64_t b = (uint64_t)(1023 + e) << 52;
    double d; memcpy(&d, &b, 8); return d;
}
BK_HD uint32_t fx_floor_exp(int M) { return (uint32_t)(1023 + M - FX_MAIN_BINADES - FX_FINE_BINADES); }
// classify x and split it into its lowest limb `lo` and the sign-extended rest `up`
BK_HD int fx_split(double x, double scale, uint32_t fx_lo, uint32_t& lo, uint64_t& up) {
    const double y = x * scale;
    const uint32_t d = ((fx_hi_word(y) >> 20) & 0x7FFu) - fx_lo;   // binades above the fine range's floor (wraps for smaller exponents)
    if (d - (uint32_t)FX_FINE_BINADES < (uint32_t)FX_MAIN_BINADES) {
        const long long f = fx_round(y);
        lo = (uint32_t)f; up = (uint64_t)(f >> 32);
        return FX_MAIN;
    }
    if (d < (uint32_t)FX_FINE_BINADES) {
        const long long f = fx_round(y * 4294967296.0);
        lo = (uint32_t)f; up = (uint64_t)(f >> 32);
        return FX_FINE;
    }
    return x == 0.0 ? FX_ZERO : FX_EXACT;
}
// the double a slot's limbs stand for: top = {mid, hi} as one signed 64-bit integer
BK_HD double fx_combine(long long top, uint32_t ext, int F) {
    return scalbn((double)top * 4294967296.0 + (double)ext, -(F + 32));
}

}  // namespace bk
