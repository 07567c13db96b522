// cosf / sinf of the descriptor rotation (reference src/ORBextractor.cc:160: `float a = (float)cos(angle), b = (float)sin(angle);` with a
// float argument and `using namespace std`, i.e. libm's cosf / sinf).  The algorithm lives in a dependency that is not part of the reference
// tree: glibc's libm (>= 2.28; the image has 2.39), sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h — the ARM optimized-routines
// single-precision sine / cosine: the argument widened to double, n = round(x * 2/pi) by a scaled truncation (`reduce_fast`), x - n * pi/2,
// a degree-7 sine or degree-8 cosine polynomial in double (`sinf_poly`; the second coefficient set yields -cos), ONE rounding to float.
// Restated here for 0 <= x < 120 (the rotation angle is at most 2 pi).  Pinned exhaustively: tools/cpp/sincos_exhaustive.cu compares this
// function with the host's cosf / sinf on every float angle in [0, 360] degrees — 1 135 869 953 arguments, no difference, with the FMA
// contractions of glibc's -mfma ifunc variant (used here) and without them (profiles/r2final_sincos_exhaustive.txt).
#pragma once
#include <cstdint>

namespace orbtrig {

__device__ __forceinline__ float poly(double x, double x2, bool neg_cos, int n)
{
    if ((n & 1) == 0) {                                   // sine: x + x^3 (s0 + x^2 s1 + x^4 s2), evaluated as sinf_poly does
        const double s0 = -0x1.555545995a603p-3, s1 = 0x1.1107605230bc4p-7, s2 = -0x1.994eb3774cf24p-13;
        const double x3 = __dmul_rn(x, x2), t1 = fma(x2, s2, s1), x7 = __dmul_rn(x3, x2), s = fma(x3, s0, x);
        return (float)fma(x7, t1, s);
    }
    const double sg = neg_cos ? -1.0 : 1.0;               // __sincosf_table[1] holds the negated cosine coefficients
    const double c0 = sg, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5, c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
    const double x4 = __dmul_rn(x2, x2), q2 = fma(x2, c4, c3), q1 = fma(x2, c1, c0), x6 = __dmul_rn(x4, x2), c = fma(x4, c2, q1);
    return (float)fma(x6, q2, c);
}

// cs = cosf(y), sn = sinf(y) for 0 <= y < 120
__device__ __forceinline__ void sincosf_glibc(float y, float& sn, float& cs)
{
    const uint32_t top = (__float_as_uint(y) >> 20) & 0x7ff;                  // abstop12
    double x = (double)y;
    if (top < 0x3f4) {                                    // abstop12 (y) < abstop12 (pi/4)
        if (top < 0x398) { cs = 1.0f; sn = y; return; }   // |y| < 2^-12
        const double x2 = __dmul_rn(x, x);
        sn = poly(x, x2, false, 0); cs = poly(x, x2, false, 1);
        return;
    }
    const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);                    // x * 2/pi * 2^24
    const int n = (__double2int_rz(r) + 0x800000) >> 24;
    x = fma(-(double)n, 0x1.921FB54442D18p0, x);
    const bool neg = (n & 2) != 0;
    const double xs = ((n + 1) & 2) ? -x : x, x2 = __dmul_rn(x, x);           // sign[n & 3] = { 1, -1, -1, 1 }
    sn = poly(xs, x2, neg, n); cs = poly(xs, x2, neg, n ^ 1);
}

} // namespace orbtrig
