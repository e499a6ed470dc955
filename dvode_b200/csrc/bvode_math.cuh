// bvode_math.cuh -- scalar helpers shared by every kernel of the batched VODE path.
//
// Two builds of the same source exist (see build.py):
//   fast   : default nvcc FP64 code generation (FMA contraction on), CUDA pow().
//   strict : -fmad=false and the portable pow below, so that every floating-point
//            operation is individually rounded exactly like the CPU oracle built
//            with -ffp-contract=off and pow_mode = 1.  The strict build exists to
//            prove the logic bit for bit; the fast build is the product.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef BVODE_STRICT
#define BVODE_STRICT 0
#endif

#if BVODE_STRICT
#define BVODE_NS bvode_strict
#else
#define BVODE_NS bvode_fast
#endif

#define BV_DEV __device__ __forceinline__
#include "bvode_real.h"

namespace BVODE_NS {

// REALIFY-OFF
#if BVODE_REAL_IS_FLOAT
constexpr real kUround = 1.1920929e-07f;  // epsilon(1.0_real32), M:36 with wp = real32 (K:29-31)
#else
constexpr real kUround = 2.220446049250313e-16;  // epsilon(1.0_real64), M:36
#endif
// REALIFY-ON

BV_DEV real dmax2(real a, real b) { return a > b ? a : b; }
BV_DEV real dmin2(real a, real b) { return a < b ? a : b; }
BV_DEV real dsign(real a, real b) { return copysign(fabs(a), b); }

// x**n for integer n the way gfortran lowers it (libgcc __powidf2), M:1956, M:2287.
BV_DEV real powi(real x, int m) {
    unsigned int n = m < 0 ? (unsigned int)(-m) : (unsigned int)m;
    real y = (n & 1u) ? x : BV_R(1.0);
    while (n >>= 1) {
        x = x * x;
        if (n & 1u) y *= x;
    }
    return m < 0 ? BV_R(1.0) / y : y;
}

// x**m for 2 <= m <= MAXP.  Strict build: powi (the reference's operation order).  Fast build,
// small MAXP (BDF): straight-line products and a select -- no divergent loop in the step loop.
template <int MAXP>
BV_DEV real powi_small(real x, int m) {
#if !BVODE_STRICT
    if constexpr (MAXP <= 7) {
        const real x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
        real r = x2;
        r = (m == 3) ? x3 : r;
        r = (m == 4) ? x4 : r;
        r = (m == 5) ? x4 * x : r;
        r = (m == 6) ? x3 * x3 : r;
        r = (m == 7) ? x4 * x3 : r;
        return r;
    }
#endif
    return powi(x, m);
}

// REALIFY-OFF (double inside whatever `real` is)
// Portable x**y, x > 0: same operation sequence as dvo_ppow in oracle/dvode_oracle.c.
// Only meaningful bit for bit when compiled with -fmad=false.
static __device__ __noinline__ double ppow(double x, double y) {
    if (x == 0.0) return 0.0;
    if (x == 1.0 || y == 0.0) return 1.0;
    uint64_t u = (uint64_t)__double_as_longlong(x);
    int e = (int)((u >> 52) & 0x7ff);
    if (e == 0) {
        x = x * 18014398509481984.0;
        u = (uint64_t)__double_as_longlong(x);
        e = (int)((u >> 52) & 0x7ff) - 54;
    }
    e -= 1023;
    u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    double m = __longlong_as_double((long long)u);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double p = 1.0 / 25.0;
    p = p * s2 + 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0;
    p = p * s2 + 1.0 / 19.0;
    p = p * s2 + 1.0 / 17.0;
    p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0;
    p = p * s2 + 1.0 / 11.0;
    p = p * s2 + 1.0 / 9.0;
    p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0;
    p = p * s2 + 1.0 / 3.0;
    double lm = 2.0 * s + (2.0 * s) * (s2 * p);
    const double ln2_hi = 6.93147180369123816490e-01;
    const double ln2_lo = 1.90821492927058770002e-10;
    double ed = (double)e;
    double lx = (ed * ln2_hi + lm) + ed * ln2_lo;
    double z = y * lx;
    if (z > 709.0) return __longlong_as_double(0x7ff0000000000000LL);
    if (z < -745.0) return 0.0;
    double kd = z * 1.44269504088896338700e+00;
    kd = (kd >= 0.0) ? (double)(long long)(kd + 0.5) : (double)(long long)(kd - 0.5);
    double r = (z - kd * ln2_hi) - kd * ln2_lo;
    r = r * 0.0625;
    double q = 1.0 / 3628800.0;
    q = q * r + 1.0 / 362880.0;
    q = q * r + 1.0 / 40320.0;
    q = q * r + 1.0 / 5040.0;
    q = q * r + 1.0 / 720.0;
    q = q * r + 1.0 / 120.0;
    q = q * r + 1.0 / 24.0;
    q = q * r + 1.0 / 6.0;
    q = q * r + 0.5;
    q = q * r + 1.0;
    q = q * r;
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    q = 2.0 * q + q * q;
    double er = 1.0 + q;
    int k = (int)kd;
    int k1 = k / 2, k2 = k - k1;
    double t1 = __longlong_as_double((long long)((uint64_t)(k1 + 1023) << 52));
    double t2 = __longlong_as_double((long long)((uint64_t)(k2 + 1023) << 52));
    return (er * t1) * t2;
}

// REALIFY-ON
// x**y with real y: the reference calls libm pow through gfortran (M:2201, 2275, 2282, 2292).
// (REAL32: the portable pow works in double and its result is rounded to real -- on both sides.)
BV_DEV real rpow(real x, real y) {
#if BVODE_STRICT
    return (real)ppow((real)x, (real)y);
#else
    return pow(x, y);
#endif
}

// a/b and 1/b on the step loop.  Strict build: the IEEE operations.  Fast build: the
// hardware reciprocal seed (MUFU.RCP64H, relative error 2^-23) refined by one cubic step
// (error^3 -> below 2^-53) and, for a/b, one residual correction -- faithfully rounded
// (<= 1 ulp, equal to the IEEE quotient in all but a few per cent of cases), 6-7 FP64
// instructions, no range test and no out-of-line slow path (CUDA's `/` is ~15 instructions
// plus a call per site; the kernel has ~70 sites and the loop body must fit the instruction
// cache).  Valid for normal, finite divisors with |b| in [2^-1000, 2^1000]: step sizes, error
// weights and the order coefficients of a running integration are all far inside that.
#if BVODE_STRICT
BV_DEV real qrcp(real b) { return BV_R(1.0) / b; }
BV_DEV real qdiv(real a, real b) { return a / b; }
#else
BV_DEV real qrcp(real b) {
    real r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    const real e = fma(-b, r, BV_R(1.0));
    const real t = fma(e, e, e);
    return fma(r, t, r);
}
BV_DEV real qdiv(real a, real b) {
    const real r = qrcp(b);
    const real q = a * r;
    const real rem = fma(-b, q, a);
    return fma(rem, r, q);
}
#endif

// 1/(double)k for the small integers of the order logic (k = 0..15): the same bits as the
// reference's `one/real(k)` divide, from a constant-bank table (one indexed LDC; lanes at
// different orders cost one replay per distinct k) -- no divide, no divergent switch.
__constant__ real kRecipTab[16] = {BV_R(0.0),        BV_R(1.0),        BV_R(1.0) / BV_R(2.0),  BV_R(1.0) / BV_R(3.0),  BV_R(1.0) / BV_R(4.0),  BV_R(1.0) / BV_R(5.0),
                                     BV_R(1.0) / BV_R(6.0),  BV_R(1.0) / BV_R(7.0),  BV_R(1.0) / BV_R(8.0),  BV_R(1.0) / BV_R(9.0),  BV_R(1.0) / BV_R(10.0), BV_R(1.0) / BV_R(11.0),
                                     BV_R(1.0) / BV_R(12.0), BV_R(1.0) / BV_R(13.0), BV_R(1.0) / BV_R(14.0), BV_R(1.0) / BV_R(15.0)};
BV_DEV real recip_int(int k) { return kRecipTab[k & 15]; }

#if !BVODE_STRICT
// Fast build: the step-size ratio eta = 1/(x**(1/k) + addon) (M:2201, 2275, 2282, 2292) is
// evaluated as 1/(exp2(log2(x)/k) + addon) with the double-precision log2 / exp2 below
// (rounding-level difference from pow).
typedef real lg_t;
// log2 / exp2 in double precision, ~2 ulp (checked against libm over 2e6 random arguments:
// max relative error 4.3e-16 / 2.2e-16), about a third of the instructions of CUDA's
// correctly-rounded-ish log2() / exp2(): atanh series on the mantissa reduced to
// [sqrt(1/2), sqrt(2)], degree-13 Taylor polynomial of 2^f on [-1/2, 1/2].
BV_DEV lg_t log2_fast(real x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    int e = (int)(u >> 52) - 1023;
    u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;
    real m = __longlong_as_double((long long)u);
    if (m > BV_R(1.4142135623730951)) { m *= BV_R(0.5); e += 1; }
    const real s = qdiv(m - BV_R(1.0), m + BV_R(1.0)), z = s * s;
    real p = BV_R(1.0) / BV_R(21.0);
    p = fma(p, z, BV_R(1.0) / BV_R(19.0));
    p = fma(p, z, BV_R(1.0) / BV_R(17.0));
    p = fma(p, z, BV_R(1.0) / BV_R(15.0));
    p = fma(p, z, BV_R(1.0) / BV_R(13.0));
    p = fma(p, z, BV_R(1.0) / BV_R(11.0));
    p = fma(p, z, BV_R(1.0) / BV_R(9.0));
    p = fma(p, z, BV_R(1.0) / BV_R(7.0));
    p = fma(p, z, BV_R(1.0) / BV_R(5.0));
    p = fma(p, z, BV_R(1.0) / BV_R(3.0));
    const real t = s * BV_R(2.8853900817779268);  // 2/ln 2
    return fma(t * z, p, t) + (real)e;
}
// log2 of a positive normal double in single precision: exponent exactly, mantissa through the
// hardware lg2 (absolute error < 3e-7 + 6e-8 |log2 x|).  Only used to ORDER step-ratio
// candidates that are far apart; near_f32 says when two such values are too close to call.
BV_DEV float log2_f32(real x) {
    const unsigned hi = (unsigned)__double2hiint(x), lo = (unsigned)__double2loint(x);
    const int e = (int)(hi >> 20) - 1023;
    const float m = __uint_as_float(0x3f800000u | ((hi & 0xfffffu) << 3) | (lo >> 29));
    return (float)e + __log2f(m);
}
BV_DEV bool near_f32(float a, float b) { return fabsf(a - b) <= 1.0e-5f * (1.0f + fabsf(a) + fabsf(b)); }
BV_DEV real exp2_fast(lg_t l) {
    const real n = rint(l), f = l - n;
    real p = BV_R(1.369148885390412888e-12);
    p = fma(p, f, BV_R(2.567843599348820514e-11));
    p = fma(p, f, BV_R(4.445538271870811498e-10));
    p = fma(p, f, BV_R(7.054911620801123329e-9));
    p = fma(p, f, BV_R(1.017808600923969973e-7));
    p = fma(p, f, BV_R(1.321548679014430949e-6));
    p = fma(p, f, BV_R(1.525273380405984028e-5));
    p = fma(p, f, BV_R(1.540353039338160995e-4));
    p = fma(p, f, BV_R(1.333355814642844342e-3));
    p = fma(p, f, BV_R(9.618129107628477232e-3));
    p = fma(p, f, BV_R(5.550410866482157995e-2));
    p = fma(p, f, BV_R(2.402265069591007123e-1));
    p = fma(p, f, BV_R(6.931471805599453094e-1));
    p = fma(p, f, BV_R(1.0));
    const int ni = (int)n;
    if (ni < -1020) return BV_R(0.0);
    if (ni > 1020) return __longlong_as_double(0x7ff0000000000000LL);
    return p * __longlong_as_double((long long)(ni + 1023) << 52);
}
#endif

}  // namespace BVODE_NS
