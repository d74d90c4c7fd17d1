// hb200_trig.cuh -- the library's own sin / cos / acos for BOUNDED angles (tesseroid walks).
//
// CUDA's sin / cos / acos spend most of their instructions on argument ranges and special values
// that cannot occur here: the arguments are longitudes / latitudes in radians (|x| < 4 pi) and
// cosines of angular distances (|x| <= 1, or a last-place excess that must give NaN like libm's).
//   fast_sincos: k = rint(x * 2/pi), r = x - k * pi/2 in three parts (exact products for |k| < 2^20),
//                then the classic minimax kernels on |r| <= pi/4 (fdlibm's __kernel_sin / __kernel_cos
//                coefficients, public domain), quadrant by k & 3.
//   fast_acos:   fdlibm's e_acos.c scheme (rational R(z) = p(z) / q(z); for |x| > 1/2 through
//                s = sqrt((1 - |x|) / 2) with a two-part square root), one division, one sqrt.
// All steps are written with explicit fma() so that the host build (tests) and the device build
// execute the same arithmetic. Accuracy (tests/test_tesseroid_host.py): < 1 ulp against glibc.
#pragma once
#include "hb200_math.cuh"

namespace hb {

HB_HD void fast_sincos(double x, double& s, double& c)
{
    const double kTwoOverPi = 6.36619772367581382433e-01;
    const double kPio2_1 = 1.57079632673412561417e+00;   // first 33 bits of pi/2
    const double kPio2_2 = 6.07710050630396597660e-11;   // next 33 bits
    const double kPio2_2t = 2.02226624879595063154e-21;  // pi/2 - (kPio2_1 + kPio2_2)
    const double kq = rint(x * kTwoOverPi);
    double r = fma(-kq, kPio2_1, x);
    r = fma(-kq, kPio2_2, r);
    r = fma(-kq, kPio2_2t, r);
    const double z = r * r;
    // __kernel_sin: r + r^3 (S1 + z (S2 + ... ))
    double ps = fma(1.58969099521155010221e-10, z, -2.50507602534068634195e-08);
    ps = fma(ps, z, 2.75573137070700676789e-06);
    ps = fma(ps, z, -1.98412698298579493134e-04);
    ps = fma(ps, z, 8.33333333332248946124e-03);
    ps = fma(ps, z, -1.66666666666666324348e-01);
    const double sr = fma(r * z, ps, r);
    // __kernel_cos: 1 - z/2 + z^2 (C1 + z (C2 + ... ))
    double pc = fma(-1.13596475577881948265e-11, z, 2.08757232129817482790e-09);
    pc = fma(pc, z, -2.75573143513906633035e-07);
    pc = fma(pc, z, 2.48015872894767294178e-05);
    pc = fma(pc, z, -1.38888888888741095749e-03);
    pc = fma(pc, z, 4.16666666666666019037e-02);
    const double hz = 0.5 * z;
    const double w = 1.0 - hz;
    const double cr = w + (((1.0 - w) - hz) + z * z * pc);
    const int q = (int)kq & 3;
    const double s0 = (q & 1) ? cr : sr;
    const double c0 = (q & 1) ? sr : cr;
    s = (q & 2) ? -s0 : s0;
    c = ((q + 1) & 2) ? -c0 : c0;
}

HB_HD double fast_cos(double x)
{
    double s, c;
    fast_sincos(x, s, c);
    return c;
}

// acos(x) for |x| <= 1; NaN beyond (like libm), which the split test then treats as "no split"
HB_HD double fast_acos(double x)
{
    const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
    const double pi = 3.14159265358979311600e+00;
    const double ax = fabs(x);
    if (!(ax <= 1.0)) return (x - x) / (x - x);  // NaN
    if (ax == 1.0) return x > 0.0 ? 0.0 : pi + 2.0 * pio2_lo;
    const bool small = ax < 0.5;
    const double z = small ? x * x : (1.0 - ax) * 0.5;
    double p = fma(3.47933107596021167570e-05, z, 7.91534994289814532176e-04);
    p = fma(p, z, -4.00555345006794114027e-02);
    p = fma(p, z, 2.01212532134862925881e-01);
    p = fma(p, z, -3.25565818622400915405e-01);
    p = fma(p, z, 1.66666666666666657415e-01);
    p = p * z;
    double q = fma(7.70381505559019352791e-02, z, -6.88283971605453293030e-01);
    q = fma(q, z, 2.02094576023350569471e+00);
    q = fma(q, z, -2.40339491173441421878e+00);
    q = fma(q, z, 1.0);
    const double r = p / q;
    if (small) return pio2_hi - (x - (pio2_lo - x * r));
    const double s = sqrt(z);
    if (x < 0.0) {  // x < -0.5
        const double w = fma(r, s, -pio2_lo);
        return pi - 2.0 * (s + w);
    }
    // x > 0.5: s = df + c with df the head of sqrt(z)
#if defined(__CUDA_ARCH__)
    const double df = __hiloint2double(__double2hiint(s), 0);
#else
    double df = s;
    {
        uint64_t bits;
        memcpy(&bits, &df, 8);
        bits &= 0xffffffff00000000ull;
        memcpy(&df, &bits, 8);
    }
#endif
    const double cc = fma(-df, df, z) / (s + df);
    const double w = fma(r, s, cc);
    return 2.0 * (df + w);
}

}  // namespace hb
