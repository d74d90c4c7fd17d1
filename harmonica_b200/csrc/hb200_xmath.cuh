// hb200_xmath.cuh -- float64 division / sqrt / log / atan2 sequences for the
// merged prism path, written for the B200 FP64 pipe.
//
// Why not CUDA's libm here: ncu on the libm build of prism_kernel<g_z> shows the
// kernel is ISSUE-bound, not FP64-bound (1276 thread-instructions per pair of
// which only 478 go to the FP64 pipe): libm's log/atan2/div/sqrt spend most of
// their instructions on special-case branches, integer exponent handling and
// on materialising polynomial coefficients with MOVs. The arguments on the
// merged path are tame (positive, finite, far from the subnormal range), so
// the sequences below drop the special cases, take their coefficients from
// the constant bank (operands of DFMA, no MOV), and use table-driven argument
// reduction so that the polynomials are short:
//   fast_div   MUFU.RCP64H seed + one cubic Newton step         4 FP64 instr
//   fast_sqrt  MUFU.RSQ64H seed + 2 Goldschmidt steps + fix-up   9 FP64 instr (correctly rounded)
//   fast_sqrt_1ulp  MUFU.RSQ64H seed + one cubic step            5 FP64 instr
//   fast_log   128-bucket reciprocal table, degree-7 log1p     ~11 FP64 instr
//   fast_atan2 17-entry atan table, one division, degree-11     ~20 FP64 instr
// (CUDA 12.9 libm: div 8, sqrt 8, log ~30, atan2 ~45 FP64 instr + ~2x as many
// integer/branch/move instructions.)
// Accuracy (tests/test_pair_math_host.py, tests/test_gpu_parity.py): <= 2 ulp
// for div/log/atan2 results, sqrt correctly rounded in all sampled cases.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "hb200_math.cuh"

#if defined(__CUDACC__)
#define HB_TABLE static __device__ const
#define HB_COEF static __constant__ const
#else
#define HB_TABLE static const
#define HB_COEF static const
#endif
#include "hb200_tables.h"

namespace hb {

HB_HD double flip_sign_if(double x, bool neg)
{
    return make_double(hi_word(x) ^ (neg ? (int)0x80000000 : 0), lo_word(x));
}

// ~20-bit seeds (MUFU.RCP64H / MUFU.RSQ64H look at the upper 32 bits only)
HB_HD double rcp_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return make_double(hi_word(1.0 / make_double(hi_word(x), 0)), 0);  // emulates the truncation
#endif
}
HB_HD double rsqrt_seed(double x)
{
#if defined(__CUDA_ARCH__)
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return make_double(hi_word(1.0 / sqrt(make_double(hi_word(x), 0))), 0);
#endif
}

// 1/x for finite normal x != 0: seed error e ~ 2^-20, y = y0 (1 + e + e^2), error e^3
HB_HD double fast_rcp(double x)
{
    const double y0 = rcp_seed(x);
    const double e = fma(-x, y0, 1.0);
    const double t = fma(e, e, e);
    return fma(y0, t, y0);
}
HB_HD double fast_div(double a, double b) { return a * fast_rcp(b); }

// sqrt(x) for finite normal x > 0 (Goldschmidt; last step makes it correctly
// rounded except for rare half-way cases)
HB_HD double fast_sqrt(double x)
{
    const double y0 = rsqrt_seed(x);
    double g = x * y0;
    double h = 0.5 * y0;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}

// sqrt(x) for finite normal x > 0 to within 1 ulp: g = g0 (1 + e/2 + 3 e^2/8), g0 = x y0,
// e = 1 - g0 y0 (5 FP64 instructions). Enough for the merged path, where r no longer decides
// a branch (needs_exact_path() has already diverted every pair on which r == |x| could hold).
HB_HD double fast_sqrt_1ulp(double x)
{
    const double y0 = rsqrt_seed(x);
    const double g0 = x * y0;
    const double e = fma(-g0, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    p = p * e;
    return fma(g0, p, g0);
}

// 1/sqrt(x) for finite normal x > 0: seed error e ~ 2^-20, y = y0 (1 + e/2 + 3e^2/8), error O(e^3)
HB_HD double fast_rsqrt(double x)
{
    const double y0 = rsqrt_seed(x);
    const double t = x * y0;
    const double e = fma(-t, y0, 1.0);
    double p = fma(e, 0.375, 0.5);
    p = p * e;
    return fma(y0, p, y0);
}

HB_HD double point_rsqrt(double d2) { return fast_rsqrt(d2); }

// log1p Taylor coefficients z^2 .. z^11
HB_COEF double kLogC[10] = {-0.5, 1.0 / 3.0, -0.25, 0.2, -1.0 / 6.0, 1.0 / 7.0,
                            -0.125, 1.0 / 9.0, -0.1, 1.0 / 11.0};
// atan Taylor coefficients s^1 .. s^5 (s = t^2)
HB_COEF double kAtanC[5] = {-1.0 / 3.0, 0.2, -1.0 / 7.0, 1.0 / 9.0, -1.0 / 11.0};
HB_COEF double kLn2 = 0.693147180559945309417232121458;

// log(a) for finite normal a > 0.
// a = 2^k m, m in [sqrt(1/2), sqrt(2)); c ~ 1/m from a 128-bucket table with
// c == 1 for the bucket around m == 1; z = m c - 1 (exact in one fma), |z| < 2^-8;
// log a = k ln2 - log c + log1p(z).
HB_HD double fast_log(double a)
{
    const int ix = hi_word(a);
    const int tmp = ix - HB_LOG_OFF;
    const int k = tmp >> 20;
    const int idx = (tmp >> (20 - HB_LOG_BITS)) & ((1 << HB_LOG_BITS) - 1);
    const double m = make_double(ix - (k << 20), lo_word(a));
    const double c = hb_log_tab[idx][0];
    const double lc = hb_log_tab[idx][1];
    const double z = fma(m, c, -1.0);
    const double z2 = z * z;
    double p = fma(kLogC[5], z, kLogC[4]);
    p = fma(p, z, kLogC[3]);
    p = fma(p, z, kLogC[2]);
    p = fma(p, z, kLogC[1]);
    p = fma(p, z, kLogC[0]);
    const double l1p = fma(z2, p, z);
    return fma((double)k, kLn2, lc) + l1p;
}

// log1p(z) for |z| < 2^-8 (the polynomial of fast_log without the table reduction)
HB_HD double log1p_small(double z)
{
    const double z2 = z * z;
    double p = fma(kLogC[5], z, kLogC[4]);
    p = fma(p, z, kLogC[3]);
    p = fma(p, z, kLogC[2]);
    p = fma(p, z, kLogC[1]);
    p = fma(p, z, kLogC[0]);
    return fma(z2, p, z);
}

// log1p(z) = z + z^2 (c0 + c1 z + ...) with the Taylor degree chosen from the class of |z|
// (m = upper word of |z|): |z| < 2^-11: z^5, < 2^-8: z^7, < 2^-5: z^11 (truncation below 2^-55
// |z| in every class). The classes differ only in the LEADING Horner steps (log1p_head_*), the
// last three steps (log1p_tail) are shared, so a warp whose lanes fall into two classes
// repeats little. m >= kLog1pMax is not allowed.
constexpr int kLog1pTiny = 0x3f400000;   // 2^-11
constexpr int kLog1pSmall = 0x3f700000;  // 2^-8
constexpr int kLog1pMax = 0x3fa00000;    // 2^-5
HB_HD double log1p_head_tiny(double z) { return fma(kLogC[3], z, kLogC[2]); }
HB_HD double log1p_head_small(double z)
{
    double p = fma(kLogC[5], z, kLogC[4]);
    p = fma(p, z, kLogC[3]);
    return fma(p, z, kLogC[2]);
}
HB_HD double log1p_head_mid(double z)
{
    double p = fma(kLogC[9], z, kLogC[8]);
    p = fma(p, z, kLogC[7]);
    p = fma(p, z, kLogC[6]);
    p = fma(p, z, kLogC[5]);
    p = fma(p, z, kLogC[4]);
    p = fma(p, z, kLogC[3]);
    return fma(p, z, kLogC[2]);
}
HB_HD double log1p_tail(double z, double p)
{
    p = fma(p, z, kLogC[1]);
    p = fma(p, z, kLogC[0]);
    return fma(z * z, p, z);
}
HB_HD double log1p_nested(double z, int m)
{
    double p;
    if (m < kLog1pTiny) p = log1p_head_tiny(z);
    else if (m < kLog1pSmall) p = log1p_head_small(z);
    else p = log1p_head_mid(z);
    return log1p_tail(z, p);
}

// log(1 + z) for |z| < 1/4 without a table: log(1 + z) = 2 atanh(w), w = z / (2 + z), |w| < 1/7;
// with t = 2 w:  t + t^3 (1/12 + t^2 / 80 + ...), coefficients 1 / ((2k + 1) 4^k), k = 1 .. 9
// (the next term is below 2^-59 of the result). One division and a polynomial in t^2: no
// exponent extraction, no table loads, no integer-to-double conversion -- the ratios of the
// potential's vertex pairs (first differences along an axis) mostly fall in this class.
constexpr int kLogAtanhMax = 0x3fd00000;  // 2^-2
HB_COEF double kAtanhC[9] = {1.0 / 12.0, 1.0 / 80.0, 1.0 / 448.0, 1.0 / 2304.0, 1.0 / 11264.0,
                             1.0 / 53248.0, 1.0 / 245760.0, 1.0 / 1114112.0, 1.0 / 4980736.0};
HB_HD double log1p_atanh(double z)
{
    const double t = (z + z) * fast_rcp(2.0 + z);
    const double u = t * t;
    double p = fma(kAtanhC[8], u, kAtanhC[7]);
    p = fma(p, u, kAtanhC[6]);
    p = fma(p, u, kAtanhC[5]);
    p = fma(p, u, kAtanhC[4]);
    p = fma(p, u, kAtanhC[3]);
    p = fma(p, u, kAtanhC[2]);
    p = fma(p, u, kAtanhC[1]);
    p = fma(p, u, kAtanhC[0]);
    return fma(t * u, p, t);
}

// sufficient for x > 0 and |y| < x / 32 (resp. x / 512): compared on the upper words only
// (positive doubles order like integers; subtracting 5 (9) from the exponent field divides by 32
// (512)). Pairs that fail take the general sequence, which is valid everywhere.
HB_HD bool small_angle(double y, double x)
{
    return hi_word(x) - (5 << 20) > (hi_word(y) & 0x7fffffff);
}
HB_HD bool tiny_angle(double y, double x)
{
    return hi_word(x) - (9 << 20) > (hi_word(y) & 0x7fffffff);
}

// atan(y / x) for x > 0, |y| < x / 32: same polynomial as fast_atan2 after its reduction;
// tiny (|y| < x / 512): the three leading Horner steps are dropped (truncation s^3 / 7 < 2^-56)
HB_HD double atan_poly_small(double s)
{
    double p = fma(kAtanC[4], s, kAtanC[3]);
    p = fma(p, s, kAtanC[2]);
    p = fma(p, s, kAtanC[1]);
    return fma(p, s, kAtanC[0]);
}
HB_HD double atan_poly_tiny(double s) { return fma(kAtanC[1], s, kAtanC[0]); }
HB_HD double atan_small(double y, double x, bool tiny = false)
{
    const double a = y * fast_rcp(x);
    const double s = a * a;
    const double p = tiny ? atan_poly_tiny(s) : atan_poly_small(s);
    return fma(a * s, p, a);
}

// atan2(y, x) in (-pi, pi] for finite x, y not both zero.
// q = min/max in [0, 1]; c = round(16 q)/16 from a 20-bit estimate;
// atan q = atan c + atan((min - c max)/(max + c min)), |argument| <= 1/32.
HB_HD double fast_atan2(double y, double x)
{
    const int hx = hi_word(x), hy = hi_word(y);
    const double ax = make_double(hx & 0x7fffffff, lo_word(x));
    const double ay = make_double(hy & 0x7fffffff, lo_word(y));
    const bool swap = ay > ax;
    const double mx = swap ? ay : ax;
    const double mn = swap ? ax : ay;
    const double q0 = mn * rcp_seed(mx);
    const double magic = 6755399441055744.0;  // 1.5 * 2^52
    const double t = fma(q0, 16.0, magic);
    const int idx = lo_word(t) & 31;
    const double c = (t - magic) * 0.0625;
    const double num = fma(-c, mx, mn);
    const double den = fma(c, mn, mx);
    const double a = num * fast_rcp(den);
    const double s = a * a;
    double p = fma(kAtanC[4], s, kAtanC[3]);
    p = fma(p, s, kAtanC[2]);
    p = fma(p, s, kAtanC[1]);
    p = fma(p, s, kAtanC[0]);
    const double as = a * s;
    double r = fma(as, p, a) + hb_atan_tab[idx];
    // quadrant: swap -> pi/2 - r; x < 0 -> pi - (.)
    const bool xneg = hx < 0;
    const double base = swap ? (kPi / 2) : (xneg ? kPi : 0.0);
    r = base + flip_sign_if(r, swap != xneg);
    return flip_sign_if(r, hy < 0);
}

}  // namespace hb
