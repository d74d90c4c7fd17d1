// hb200_fast.cuh -- merged-transcendental evaluation of the 8-vertex prism sum.
//
// Valid for the pairs classify_pair<FS>() lets through: every pair whose observer is not
// on (the extension of) an edge or a vertex of the prism -- ONE axis may hold a zero shift
// (observer in the plane of a face), the reference's face rule is then applied after the
// merged evaluation (PAIR_FAST_CHECK). Every other pair takes prism_pair_direct, which
// carries the reference's singular-point rules verbatim.
//
// The reference (choclo kernels behind gravity.py:526-537 / magnetic.py:319)
// evaluates per vertex 1-3 safe_log and 1-3 safe_atan2 and forms an
// alternating sum. Here vertex terms that share their prefactor are merged
// BEFORE the transcendental:
//   * logs:  sum_v s_v log(num_v/den_v) = log(prod num^s / prod den^s)
//            with (num, den) chosen by the safe_log rule (T = r + |x|)
//              x >= 0:  (T, 1)        [log(x + r)]
//              x <  0:  (y^2 + z^2, T) [log((y^2 + z^2) / (r - x))]
//            (the r == 0 and on-axis r == -x rules only fire on pairs that
//            needs_exact_path() routes to the direct path)
//   * atans: atan(y0/x0) - atan(y1/x1) = atan2(y0 x1 - x0 y1, x0 x1 + y0 y1)
//            for the two vertices that differ in one factor of y; x0 and x1
//            then have the same sign, so the identity is exact, and it
//            simplifies to  atan2(c a (b0 r1 - b1 r0), a^2 r0 r1 + b0 b1 c^2).
// Merged forms are algebraically identical to the vertex sum and round better
// (the far-field cancellation happens inside the ratio, not between logs).
// g_z: 16 log + 8 atan + 24 div per pair  ->  4 log + 2..4 atan2 + 6 div.
#pragma once
#include "hb200_math.cuh"
#include "hb200_xmath.cuh"

namespace hb {

// XM = false: CUDA libm; XM = true: the sequences of hb200_xmath.cuh
template <bool XM> HB_HD double x_sqrt(double x) { return XM ? fast_sqrt_1ulp(x) : sqrt(x); }

// Far-field shortcuts (log1p / small-angle atan without the table reductions) are chosen PER
// LANE from the lane's own arguments: the value of an (observer, prism) pair never depends on
// which other observers share its warp, so results are bit-reproducible under any batching,
// chunking or sharding of the observers. (A warp whose lanes all agree issues one side only.)

// log(top / bot) with y = 1 / bot, z = top y - 1 and m = upper word of |z| given: the far field
// (|z| < 2^-5) needs no table reduction
HB_HD double log_from_z(double z, double top, double y, int m)
{
    if (m < kLog1pMax) return log1p_nested(z, m);
    if (m < kLogAtanhMax) return log1p_atanh(z);
    return fast_log(top * y);
}

// log(top / bot)
template <bool XM> HB_HD double x_log_ratio(double top, double bot)
{
    if (!XM) return log(top / bot);
    const double y = fast_rcp(bot);
    const double z = fma(top, y, -1.0);
    return log_from_z(z, top, y, hi_word(z) & 0x7fffffff);
}

// atan2(y, x): far from the prism x > 0 and |y| < x / 32, no quadrant or table reduction
template <bool XM> HB_HD double x_atan2(double y, double x)
{
    if (!XM) return atan2(y, x);
    if (small_angle(y, x)) return atan_small(y, x, tiny_angle(y, x));
    return fast_atan2(y, x);
}

// sign bit of x as a mask for the upper word of a double
HB_HD unsigned sign_mask(double x) { return (unsigned)hi_word(x) & 0x80000000u; }
HB_HD bool is_neg(double x) { return hi_word(x) < 0; }
HB_HD double flip_by(double x, unsigned mask)
{
    return make_double((int)((unsigned)hi_word(x) ^ mask), lo_word(x));
}

struct FastCtx {
    double se[2], sn[2], su[2];
    double se2[2], sn2[2], su2[2];
    double en2[2][2];
    double r[2][2][2];
};

// Vertex distances carried from one record of a prism layer to the next (records are emitted
// easting-outer / northing-inner, layer.py:591-594): where the next prism has the same west /
// east bounds and bottom and its south bound IS the previous north bound (pack_layer_kernel marks
// such records), its two bottom-south vertices are the previous bottom-north ones: the same
// inputs, hence bit-identical distances, and two of the eight square roots are not redone.
// (Carrying the whole context -- shifts, squares, all shared distances -- was measured too: the
// loop-carried registers cost more moves than the 22 FP64 instructions they save, 70.8 against
// 73.4 G pair/s; these two doubles give 75.0 G.)
struct LayerCarry {
    double r[2];  // r[i][north][bottom] of the previous record, i = east / west
    bool valid;   // the previous record of this lane went through the merged path
};

template <int FS, bool XM>
HB_HD void make_fast_ctx(FastCtx& c, const PairGeom& g, const LayerCarry* carry = nullptr, bool reuse = false)
{
#pragma unroll
    for (int i = 0; i < 2; i++) {
        c.se[i] = g.se[i]; c.sn[i] = g.sn[i]; c.su[i] = g.su[i];
        c.se2[i] = g.se2[i]; c.sn2[i] = g.sn2[i]; c.su2[i] = g.su2[i];
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) c.en2[a][b] = add_rn(g.se2[a], g.sn2[b]);
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++) {
                if (carry && j == 1 && k == 1) continue;  // below
                c.r[i][j][k] = x_sqrt<XM>(add_rn(c.en2[i][j], g.su2[k]));
            }
    if (carry) {
        if (reuse) {
            c.r[0][1][1] = carry->r[0];
            c.r[1][1][1] = carry->r[1];
        } else {
            c.r[0][1][1] = x_sqrt<XM>(add_rn(c.en2[0][1], g.su2[1]));
            c.r[1][1][1] = x_sqrt<XM>(add_rn(c.en2[1][1], g.su2[1]));
        }
    }
}

// safe_log type X (0: x = e, 1: x = n, 2: x = u) at vertex ijk:
//   safe_log(x, y, z, r) = log(T) for x >= 0 and log(Y / T) for x < 0, T = r + |x|, Y = y^2 + z^2
// (the on-axis and r == 0 branches cannot occur on this path: classify_pair() sends such pairs
// to the direct path). Y does not depend on the X index, so in every alternating sum in which
// both X indices carry the SAME sign of x the Y factors cancel identically and
//   sum over a sign-symmetric vertex set of s L  =  sigma * log(prod T^s),  sigma = -1 for x < 0:
// no selects and no Y at all (SAME = true below). Only when the observer lies strictly inside the
// prism's extent on an axis (x_0 and x_1 of opposite sign: rare; ONE test per pair decides for
// all of its logs) do the Y factors stay: the general arrangement with selects (SAME = false),
//   x >= 0: (num, den) = (T, 1);   x < 0: (num, den) = (Y, T).
template <int X> HB_HD double log_x(const FastCtx& c, int xi)
{
    return (X == 0) ? c.se[xi] : (X == 1) ? c.sn[xi] : c.su[xi];
}
template <int X> HB_HD double log_T(const FastCtx& c, int i, int j, int k)
{
    const double x = (X == 0) ? c.se[i] : (X == 1) ? c.sn[j] : c.su[k];
    return c.r[i][j][k] + fabs(x);
}
template <int X> HB_HD double log_Y(const FastCtx& c, int i, int j, int k)
{
    return (X == 0) ? add_rn(c.sn2[j], c.su2[k]) : (X == 1) ? add_rn(c.se2[i], c.su2[k]) : c.en2[i][j];
}
// true when x_0 and x_1 of axis X have different sign bits
template <int X> HB_HD bool log_mixed(const FastCtx& c)
{
    return (hi_word(log_x<X>(c, 0)) ^ hi_word(log_x<X>(c, 1))) < 0;
}

// map (fixed axis F with index f, the other two indices a, b in axis order) -> ijk
template <int F> HB_HD void ijk_of(int f, int a, int b, int& i, int& j, int& k)
{
    if (F == 0) { i = f; j = a; k = b; }
    else if (F == 1) { i = a; j = f; k = b; }
    else { i = a; j = b; k = f; }
}

// The 4 vertices with index f fixed on axis F (F != X), signs (-1)^(a+b):
//   sum = +- log(top / bot), the sign flipped where the bit of `flip` is set
// [xi][o] = [X index][remaining index]
template <int X, int F, bool SAME>
HB_HD void log_group4_tb(const FastCtx& c, int f, double& top, double& bot, unsigned& flip)
{
    constexpr bool x_first = (X == (F == 0 ? 1 : 0));
    double T[2][2];
#pragma unroll
    for (int xi = 0; xi < 2; xi++)
#pragma unroll
        for (int o = 0; o < 2; o++) {
            int i, j, k;
            ijk_of<F>(f, x_first ? xi : o, x_first ? o : xi, i, j, k);
            T[xi][o] = log_T<X>(c, i, j, k);
        }
    if (SAME) {
        top = T[0][0] * T[1][1];
        bot = T[0][1] * T[1][0];
        flip = sign_mask(log_x<X>(c, 0));
    } else {
        double P[2], Q[2];
#pragma unroll
        for (int xi = 0; xi < 2; xi++) {
            double Y[2];
#pragma unroll
            for (int o = 0; o < 2; o++) {
                int i, j, k;
                ijk_of<F>(f, x_first ? xi : o, x_first ? o : xi, i, j, k);
                Y[o] = log_Y<X>(c, i, j, k);
            }
            const bool neg = is_neg(log_x<X>(c, xi));
            P[xi] = neg ? Y[0] * T[xi][1] : T[xi][0];
            Q[xi] = neg ? T[xi][0] * Y[1] : T[xi][1];
        }
        top = P[0] * Q[1];
        bot = Q[0] * P[1];
        flip = 0u;
    }
}

// sum over all 8 vertices of s_ijk L^X = +- log(top / bot): per X index xi the four vertices
// split by parity of the other two indices into side A (00, 11) and side B (01, 10)
template <int X, bool SAME>
HB_HD void log_sum8_tb(const FastCtx& c, double& top, double& bot, unsigned& flip)
{
    double TA[2], TB[2], YA = 1.0, YB = 1.0;
#pragma unroll
    for (int xi = 0; xi < 2; xi++) {
        double T[2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                int i, j, k;
                ijk_of<X>(xi, a, b, i, j, k);
                T[a][b] = log_T<X>(c, i, j, k);
            }
        TA[xi] = T[0][0] * T[1][1];
        TB[xi] = T[0][1] * T[1][0];
    }
    if (SAME) {
        top = TA[0] * TB[1];
        bot = TB[0] * TA[1];
        flip = sign_mask(log_x<X>(c, 0));
    } else {
        double Y[2][2];
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                int i, j, k;
                ijk_of<X>(0, a, b, i, j, k);
                Y[a][b] = log_Y<X>(c, i, j, k);  // independent of the X index
            }
        YA = Y[0][0] * Y[1][1];
        YB = Y[0][1] * Y[1][0];
        double P[2], Q[2];
#pragma unroll
        for (int xi = 0; xi < 2; xi++) {
            const bool neg = is_neg(log_x<X>(c, xi));
            P[xi] = neg ? YA * TB[xi] : TA[xi];
            Q[xi] = neg ? TA[xi] * YB : TB[xi];
        }
        top = P[0] * Q[1];
        bot = Q[0] * P[1];
        flip = 0u;
    }
}

// L^X(x_0) - L^X(x_1) along X's own axis; (a, b) = the other two indices in axis order:
// same signs: -+ log(T0 / T1); general: (num, den) per vertex as above (Y is shared)
template <int X, bool SAME>
HB_HD void log_pair_tb(const FastCtx& c, int a, int b, double& top, double& bot, unsigned& flip)
{
    int i0, j0, k0, i1, j1, k1;
    ijk_of<X>(0, a, b, i0, j0, k0);
    ijk_of<X>(1, a, b, i1, j1, k1);
    const double T0 = log_T<X>(c, i0, j0, k0), T1 = log_T<X>(c, i1, j1, k1);
    if (SAME) {
        top = T0;
        bot = T1;
        flip = sign_mask(log_x<X>(c, 0));
    } else {
        const double Y = log_Y<X>(c, i0, j0, k0);
        const bool neg0 = is_neg(log_x<X>(c, 0)), neg1 = is_neg(log_x<X>(c, 1));
        const double n0 = neg0 ? Y : T0, d0 = neg0 ? T0 : 1.0;
        const double n1 = neg1 ? Y : T1, d1 = neg1 ? T1 : 1.0;
        top = n0 * d1;
        bot = n1 * d0;
        flip = 0u;
    }
}

HB_HD bool any_mixed(const FastCtx& c) { return log_mixed<0>(c) | log_mixed<1>(c) | log_mixed<2>(c); }

// NQ merged log ratios with ONE class decision per lane: out[q] = +-log(top[q] / bot[q])
template <bool XM, int NQ>
HB_HD void x_log_ratio4(const double (&top)[NQ], const double (&bot)[NQ], const unsigned (&flip)[NQ],
                        double (&out)[NQ])
{
    if (!XM) {
#pragma unroll
        for (int q = 0; q < NQ; q++) out[q] = flip_by(log(top[q] / bot[q]), flip[q]);
        return;
    }
    double y[NQ], z[NQ];
    int worst = 0;
#pragma unroll
    for (int q = 0; q < NQ; q++) {
        y[q] = fast_rcp(bot[q]);
        z[q] = fma(top[q], y[q], -1.0);
        const int m = hi_word(z[q]) & 0x7fffffff;
        worst = m > worst ? m : worst;
    }
    if (worst < kLog1pMax) {
        // ONE class for the NQ logs of the lane; the classes differ in the leading Horner steps only
        double p[NQ];
        if (worst < kLog1pTiny) {  // the far field: almost all pairs
#pragma unroll
            for (int q = 0; q < NQ; q++) p[q] = log1p_head_tiny(z[q]);
        } else if (worst < kLog1pSmall) {
#pragma unroll
            for (int q = 0; q < NQ; q++) p[q] = log1p_head_small(z[q]);
        } else {
#pragma unroll
            for (int q = 0; q < NQ; q++) p[q] = log1p_head_mid(z[q]);
        }
#pragma unroll
        for (int q = 0; q < NQ; q++) out[q] = flip_by(log1p_tail(z[q], p[q]), flip[q]);
    } else if (worst < kLogAtanhMax) {  // ratios within 1/4 of 1: no table either
#pragma unroll
        for (int q = 0; q < NQ; q++) out[q] = flip_by(log1p_atanh(z[q]), flip[q]);
    } else {
#pragma unroll
        for (int q = 0; q < NQ; q++) out[q] = flip_by(fast_log(top[q] * y[q]), flip[q]);
    }
}

// sufficient for re > |im| (both finite), i.e. for the angle of (re, im) to lie inside
// (-pi/4, pi/4): compared on the upper words only (integer pipe, one instruction). Pairs that
// fail the test take the general sequence, which is valid everywhere.
HB_HD bool angle_below_quarter_pi(double im, double re)
{
    return hi_word(re) > (hi_word(im) & 0x7fffffff);
}

template <int X, bool XM> HB_HD double log_sum8(const FastCtx& c)
{
    double top, bot;
    unsigned flip;
    if (!log_mixed<X>(c)) log_sum8_tb<X, true>(c, top, bot, flip);
    else log_sum8_tb<X, false>(c, top, bot, flip);
    return flip_by(x_log_ratio<XM>(top, bot), flip);
}

// the 12 four-vertex log groups of the three acceleration components, slots as used by
// prism_pair_fast (FS_ACC3)
template <bool SAME>
HB_HD void acc3_logs_t(const FastCtx& c, double (&top)[12], double (&bot)[12], unsigned (&flip)[12])
{
#pragma unroll
    for (int f = 0; f < 2; f++) {
        log_group4_tb<2, 1, SAME>(c, f, top[0 + f], bot[0 + f], flip[0 + f]);     // E: n * L^u over (i, k)
        log_group4_tb<1, 2, SAME>(c, f, top[2 + f], bot[2 + f], flip[2 + f]);     // E: u * L^n over (i, j)
        log_group4_tb<0, 2, SAME>(c, f, top[4 + f], bot[4 + f], flip[4 + f]);     // N: u * L^e over (i, j)
        log_group4_tb<2, 0, SAME>(c, f, top[6 + f], bot[6 + f], flip[6 + f]);     // N: e * L^u over (j, k)
        log_group4_tb<1, 0, SAME>(c, f, top[8 + f], bot[8 + f], flip[8 + f]);     // U: e * L^n over (j, k)
        log_group4_tb<0, 1, SAME>(c, f, top[10 + f], bot[10 + f], flip[10 + f]);  // U: n * L^e over (i, k)
    }
}
template <bool XM>
HB_HD void acc3_logs(const FastCtx& c, double (&top)[12], double (&bot)[12], unsigned (&flip)[12])
{
    acc3_logs_t<true>(c, top, bot, flip);
    if (any_mixed(c)) {  // redo the groups of the axis the observer is inside of
#pragma unroll
        for (int f = 0; f < 2; f++) {
            if (log_mixed<2>(c)) {
                log_group4_tb<2, 1, false>(c, f, top[0 + f], bot[0 + f], flip[0 + f]);
                log_group4_tb<2, 0, false>(c, f, top[6 + f], bot[6 + f], flip[6 + f]);
            }
            if (log_mixed<1>(c)) {
                log_group4_tb<1, 2, false>(c, f, top[2 + f], bot[2 + f], flip[2 + f]);
                log_group4_tb<1, 0, false>(c, f, top[8 + f], bot[8 + f], flip[8 + f]);
            }
            if (log_mixed<0>(c)) {
                log_group4_tb<0, 2, false>(c, f, top[4 + f], bot[4 + f], flip[4 + f]);
                log_group4_tb<0, 1, false>(c, f, top[10 + f], bot[10 + f], flip[10 + f]);
            }
        }
    }
}

// the 12 vertex-pair logs of the potential: [0..3] L^u over [i][j], [4..7] L^e over [j][k],
// [8..11] L^n over [i][k]
template <bool SAME>
HB_HD void pot_logs_t(const FastCtx& c, double (&top)[12], double (&bot)[12], unsigned (&flip)[12])
{
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int q = 2 * a + b;
            log_pair_tb<2, SAME>(c, a, b, top[q], bot[q], flip[q]);
            log_pair_tb<0, SAME>(c, a, b, top[4 + q], bot[4 + q], flip[4 + q]);
            log_pair_tb<1, SAME>(c, a, b, top[8 + q], bot[8 + q], flip[8 + q]);
        }
}
HB_HD void pot_logs(const FastCtx& c, double (&top)[12], double (&bot)[12], unsigned (&flip)[12])
{
    pot_logs_t<true>(c, top, bot, flip);
    if (any_mixed(c)) {  // redo the pairs of the axis the observer is inside of
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int b = 0; b < 2; b++) {
                const int q = 2 * a + b;
                if (log_mixed<2>(c)) log_pair_tb<2, false>(c, a, b, top[q], bot[q], flip[q]);
                if (log_mixed<0>(c)) log_pair_tb<0, false>(c, a, b, top[4 + q], bot[4 + q], flip[4 + q]);
                if (log_mixed<1>(c)) log_pair_tb<1, false>(c, a, b, top[8 + q], bot[8 + q], flip[8 + q]);
            }
    }
}

// (im, re) with atan2(im, re) = A^X(b_0) - A^X(b_1) at fixed index f on axis X and index m on the
// remaining axis, A^X = atan(b c / (a r)), a = shift on axis X:
//   im = c a (b0 r1 - b1 r0),  re = a^2 r0 r1 + b0 b1 c^2
template <int X> HB_HD void atan_pair_terms(const FastCtx& c, int f, int m, double& im, double& re)
{
    const double a = (X == 0) ? c.se[f] : (X == 1) ? c.sn[f] : c.su[f];
    const double a2 = (X == 0) ? c.se2[f] : (X == 1) ? c.sn2[f] : c.su2[f];
    // paired variable b (first remaining axis), remaining variable cc (second remaining axis)
    const double b0 = (X == 0) ? c.sn[0] : c.se[0];
    const double b1 = (X == 0) ? c.sn[1] : c.se[1];
    const double cc = (X == 2) ? c.sn[m] : c.su[m];
    const double cc2 = (X == 2) ? c.sn2[m] : c.su2[m];
    int i, j, k;
    ijk_of<X>(f, 0, m, i, j, k);
    const double r0 = c.r[i][j][k];
    ijk_of<X>(f, 1, m, i, j, k);
    const double r1 = c.r[i][j][k];
    im = (cc * a) * (b0 * r1 - b1 * r0);
    re = a2 * (r0 * r1) + (b0 * b1) * cc2;
}

// S[f] = sum over the 4 vertices with index f on axis X of (-1)^(b+c) A^X = D_0 - D_1.
// Far from the prism both D are small and their difference is one more complex product:
// one atan2 instead of two (valid while |D_0|, |D_1| < pi/4, checked per pair).
template <int X, bool XM> HB_HD double atan_sum4(const FastCtx& c, int f)
{
    double im0, re0, im1, re1;
    atan_pair_terms<X>(c, f, 0, im0, re0);
    atan_pair_terms<X>(c, f, 1, im1, re1);
    if (XM && angle_below_quarter_pi(im0, re0) && angle_below_quarter_pi(im1, re1))
        return x_atan2<XM>(im0 * re1 - re0 * im1, re0 * re1 + im0 * im1);
    return x_atan2<XM>(im0, re0) - x_atan2<XM>(im1, re1);
}

// sum over all 8 vertices of s_ijk A^X = (D_00 - D_01) - (D_10 - D_11): one atan2 of the product
// z_00 conj(z_01) conj(z_10) z_11 while every |D| < pi/4 (then |sum| < pi), else pairwise.
template <int X, bool XM> HB_HD double atan_sum8(const FastCtx& c)
{
    if (XM) {
        double im00, re00, im01, re01, im10, re10, im11, re11;
        atan_pair_terms<X>(c, 0, 0, im00, re00);
        atan_pair_terms<X>(c, 0, 1, im01, re01);
        atan_pair_terms<X>(c, 1, 0, im10, re10);
        atan_pair_terms<X>(c, 1, 1, im11, re11);
        if (angle_below_quarter_pi(im00, re00) && angle_below_quarter_pi(im01, re01)
            && angle_below_quarter_pi(im10, re10) && angle_below_quarter_pi(im11, re11)) {
            const double are = re00 * re01 + im00 * im01, aim = im00 * re01 - re00 * im01;
            const double bre = re11 * re10 + im11 * im10, bim = im11 * re10 - re11 * im10;
            return x_atan2<XM>(aim * bre + are * bim, are * bre - aim * bim);
        }
    }
    return atan_sum4<X, XM>(c, 0) - atan_sum4<X, XM>(c, 1);
}

// atan(y[t] / x[t]), t = 0, 1, both small angles: one (per-lane) tiny / small decision for the
// two, the classes share everything but the leading Horner steps
HB_HD void atan_small_two(const double (&y)[2], const double (&x)[2], double& r0, double& r1)
{
    const double a0 = y[0] * fast_rcp(x[0]), a1 = y[1] * fast_rcp(x[1]);
    const double s0 = a0 * a0, s1 = a1 * a1;
    double p0, p1;
    if (tiny_angle(y[0], x[0]) && tiny_angle(y[1], x[1])) {
        p0 = atan_poly_tiny(s0);
        p1 = atan_poly_tiny(s1);
    } else {
        p0 = atan_poly_small(s0);
        p1 = atan_poly_small(s1);
    }
    r0 = fma(a0 * s0, p0, a0);
    r1 = fma(a1 * s1, p1, a1);
}

// The 8-vertex atan sums of two diagonal kernels (types XA, XB) with one branch for the merge
// test and one (per-lane) decision for the far-field sequence.
template <int XA, int XB, bool XM> HB_HD void atan_sum8_two(const FastCtx& c, double& sa, double& sb)
{
    if (!XM) {
        sa = atan_sum8<XA, XM>(c);
        sb = atan_sum8<XB, XM>(c);
        return;
    }
    double im[2][4], re[2][4];
    bool ok = true;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        atan_pair_terms<XA>(c, q >> 1, q & 1, im[0][q], re[0][q]);
        atan_pair_terms<XB>(c, q >> 1, q & 1, im[1][q], re[1][q]);
        ok = ok && angle_below_quarter_pi(im[0][q], re[0][q]) && angle_below_quarter_pi(im[1][q], re[1][q]);
    }
    if (ok) {
        double y[2], x[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            // z_00 conj(z_01) conj(z_10) z_11 (q = 2 f + m)
            const double are = re[t][0] * re[t][1] + im[t][0] * im[t][1];
            const double aim = im[t][0] * re[t][1] - re[t][0] * im[t][1];
            const double bre = re[t][3] * re[t][2] + im[t][3] * im[t][2];
            const double bim = im[t][3] * re[t][2] - re[t][3] * im[t][2];
            y[t] = aim * bre + are * bim;
            x[t] = are * bre - aim * bim;
        }
        if (small_angle(y[0], x[0]) && small_angle(y[1], x[1])) {
            atan_small_two(y, x, sa, sb);
        } else {
            sa = fast_atan2(y[0], x[0]);
            sb = fast_atan2(y[1], x[1]);
        }
    } else {
        sa = atan_sum8<XA, XM>(c);
        sb = atan_sum8<XB, XM>(c);
    }
}

// S[0], S[1] of atan_sum4 (both indices of axis X) with one branch for the merge test and one
// (per-lane) decision for the far-field sequence.
template <int X, bool XM> HB_HD void atan_sum4_both(const FastCtx& c, double (&S)[2])
{
    if (!XM) {
        S[0] = atan_sum4<X, XM>(c, 0);
        S[1] = atan_sum4<X, XM>(c, 1);
        return;
    }
    double im[2][2], re[2][2];
    bool ok = true;
#pragma unroll
    for (int f = 0; f < 2; f++)
#pragma unroll
        for (int m = 0; m < 2; m++) {
            atan_pair_terms<X>(c, f, m, im[f][m], re[f][m]);
            ok = ok && angle_below_quarter_pi(im[f][m], re[f][m]);
        }
    if (ok) {
        double y[2], x[2];
#pragma unroll
        for (int f = 0; f < 2; f++) {
            y[f] = im[f][0] * re[f][1] - re[f][0] * im[f][1];
            x[f] = re[f][0] * re[f][1] + im[f][0] * im[f][1];
        }
        if (small_angle(y[0], x[0]) && small_angle(y[1], x[1])) {
            atan_small_two(y, x, S[0], S[1]);
        } else {
            S[0] = fast_atan2(y[0], x[0]);
            S[1] = fast_atan2(y[1], x[1]);
        }
    } else {
#pragma unroll
        for (int f = 0; f < 2; f++) S[f] = fast_atan2(im[f][0], re[f][0]) - fast_atan2(im[f][1], re[f][1]);
    }
}

// One acceleration component: LA, LB = the two safe_log types with the axes FA, FB their
// prefactors run over (pa, pb), AX = the atan type with prefactor px.
template <int LA, int FA, int LB, int FB, int AX, bool XM>
HB_HD double accel_component(const FastCtx& c, const double* pa, const double* pb, const double* px)
{
    double top[4], bot[4], L[4], S[2];
    unsigned flip[4];
    log_group4_tb<LA, FA, true>(c, 0, top[0], bot[0], flip[0]);
    log_group4_tb<LA, FA, true>(c, 1, top[1], bot[1], flip[1]);
    log_group4_tb<LB, FB, true>(c, 0, top[2], bot[2], flip[2]);
    log_group4_tb<LB, FB, true>(c, 1, top[3], bot[3], flip[3]);
    if (log_mixed<LA>(c) | log_mixed<LB>(c)) {
        // observer inside the prism's extent on one of the two axes: redo that axis' groups
        if (log_mixed<LA>(c)) {
            log_group4_tb<LA, FA, false>(c, 0, top[0], bot[0], flip[0]);
            log_group4_tb<LA, FA, false>(c, 1, top[1], bot[1], flip[1]);
        }
        if (log_mixed<LB>(c)) {
            log_group4_tb<LB, FB, false>(c, 0, top[2], bot[2], flip[2]);
            log_group4_tb<LB, FB, false>(c, 1, top[3], bot[3], flip[3]);
        }
    }
    x_log_ratio4<XM, 4>(top, bot, flip, L);
    atan_sum4_both<AX, XM>(c, S);
    return pa[0] * L[0] - pa[1] * L[1] + pb[0] * L[2] - pb[1] * L[3] - (px[0] * S[0] - px[1] * S[1]);
}

// build switches of two restructurings (kept to measure them against the one-batch forms)
#ifndef HB_ACC3_BY_COMPONENT
#define HB_ACC3_BY_COMPONENT 0  // measured: 28.8 (by component) against 31.0 G pair/s (twelve logs at once)
#endif
#ifndef HB_POT_BY_AXIS
#define HB_POT_BY_AXIS 1
#endif
constexpr bool kAcc3ByComponent = HB_ACC3_BY_COMPONENT != 0;
constexpr bool kPotByAxis = HB_POT_BY_AXIS != 0;

// +0.0 for x == -0.0 (a shift bound - observer is -0 when the bound is -0 and the coordinate +0):
// the sign-bit tests of the merged path must see a zero shift as non-negative
HB_HD double plus_zero(double x) { return x == 0.0 ? 0.0 : x; }

template <int FS, bool XM>
HB_HD void prism_pair_fast(const PairGeom& g0, const double* prm, double* acc, int cls = PAIR_FAST,
                           unsigned mag_rules = 0u, unsigned* flags = nullptr, LayerCarry* carry = nullptr,
                           bool reuse = false)
{
    typedef Traits<FS> T;
    PairGeom g = g0;
    if (cls == PAIR_FAST_CHECK) {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            g.se[i] = plus_zero(g.se[i]); g.sn[i] = plus_zero(g.sn[i]); g.su[i] = plus_zero(g.su[i]);
        }
    }
    FastCtx c;
    make_fast_ctx<FS, XM>(c, g, carry, reuse);
    if (carry) {
        carry->r[0] = c.r[0][0][1];
        carry->r[1] = c.r[1][0][1];
    }
    const double* e = c.se;
    const double* n = c.sn;
    const double* u = c.su;
    if (FS == F_U) {
        acc[0] += prm[0] * -accel_component<1, 0, 0, 1, 2, XM>(c, e, n, u);
    } else if (FS == F_E) {
        acc[0] += prm[0] * -accel_component<2, 1, 1, 2, 0, XM>(c, n, u, e);
    } else if (FS == F_N) {
        acc[0] += prm[0] * -accel_component<0, 2, 2, 0, 1, XM>(c, u, e, n);
    } else if (FS == FS_ACC3 && XM && kAcc3ByComponent) {
        // the three single-component evaluations on ONE shared context (shifts, squares, the
        // eight vertex distances): four logs and one class decision at a time keep the kernel
        // inside its registers (twelve at once: 128 registers and spills)
        const double ve = accel_component<2, 1, 1, 2, 0, XM>(c, n, u, e);
        const double vn = accel_component<0, 2, 2, 0, 1, XM>(c, u, e, n);
        const double vu = accel_component<1, 0, 0, 1, 2, XM>(c, e, n, u);
        acc[0] += prm[0] * -ve;
        acc[1] += prm[0] * -vn;
        acc[2] += prm[0] * -vu;
    } else if (FS == F_POT && XM && kPotByAxis) {
        // the potential's twelve vertex-pair logs in three batches of four (one per log type):
        // a class decision per batch and a third of the live registers (all twelve at once
        // spilled 100 bytes per thread)
        double v = 0.0;
        {
            double top[4], bot[4], L[4];
            unsigned flip[4];
#pragma unroll
            for (int q = 0; q < 4; q++) log_pair_tb<2, true>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            if (log_mixed<2>(c)) {
#pragma unroll
                for (int q = 0; q < 4; q++) log_pair_tb<2, false>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            }
            x_log_ratio4<XM, 4>(top, bot, flip, L);
            v += e[0] * (n[0] * L[0] - n[1] * L[1]) - e[1] * (n[0] * L[2] - n[1] * L[3]);
        }
        {
            double top[4], bot[4], L[4];
            unsigned flip[4];
#pragma unroll
            for (int q = 0; q < 4; q++) log_pair_tb<0, true>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            if (log_mixed<0>(c)) {
#pragma unroll
                for (int q = 0; q < 4; q++) log_pair_tb<0, false>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            }
            x_log_ratio4<XM, 4>(top, bot, flip, L);
            v += n[0] * (u[0] * L[0] - u[1] * L[1]) - n[1] * (u[0] * L[2] - u[1] * L[3]);
        }
        {
            double top[4], bot[4], L[4];
            unsigned flip[4];
#pragma unroll
            for (int q = 0; q < 4; q++) log_pair_tb<1, true>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            if (log_mixed<1>(c)) {
#pragma unroll
                for (int q = 0; q < 4; q++) log_pair_tb<1, false>(c, q >> 1, q & 1, top[q], bot[q], flip[q]);
            }
            x_log_ratio4<XM, 4>(top, bot, flip, L);
            v += e[0] * (u[0] * L[0] - u[1] * L[1]) - e[1] * (u[0] * L[2] - u[1] * L[3]);
        }
        double S[2];
        atan_sum4_both<0, XM>(c, S);
        v -= 0.5 * (c.se2[0] * S[0] - c.se2[1] * S[1]);
        atan_sum4_both<1, XM>(c, S);
        v -= 0.5 * (c.sn2[0] * S[0] - c.sn2[1] * S[1]);
        atan_sum4_both<2, XM>(c, S);
        v -= 0.5 * (c.su2[0] * S[0] - c.su2[1] * S[1]);
        acc[0] += prm[0] * v;
    } else if (FS == FS_ACC3 && XM) {
        // the three single-component formulas on one shared context: second differences
        // (4-vertex groups), so the far-field log1p shortcut applies; one vote for all 12 logs
        double top[12], bot[12], L[12], Se[2], Sn[2], Su[2];
        unsigned flip[12];
        acc3_logs<XM>(c, top, bot, flip);
        x_log_ratio4<XM, 12>(top, bot, flip, L);
        atan_sum4_both<0, XM>(c, Se);
        atan_sum4_both<1, XM>(c, Sn);
        atan_sum4_both<2, XM>(c, Su);
        const double ve = n[0] * L[0] - n[1] * L[1] + u[0] * L[2] - u[1] * L[3]
                        - (e[0] * Se[0] - e[1] * Se[1]);
        const double vn = u[0] * L[4] - u[1] * L[5] + e[0] * L[6] - e[1] * L[7]
                        - (n[0] * Sn[0] - n[1] * Sn[1]);
        const double vu = e[0] * L[8] - e[1] * L[9] + n[0] * L[10] - n[1] * L[11]
                        - (u[0] * Su[0] - u[1] * Su[1]);
        acc[0] += prm[0] * -ve;
        acc[1] += prm[0] * -vn;
        acc[2] += prm[0] * -vu;
    } else if (FS == F_POT || FS == FS_ACC3) {
        double Pu[2][2], Pe[2][2], Pn[2][2], SA[3][2];
        {
            double top[12], bot[12], L[12];
            unsigned flip[12];
            pot_logs(c, top, bot, flip);
            x_log_ratio4<XM, 12>(top, bot, flip, L);
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    Pu[a][b] = L[2 * a + b];
                    Pe[a][b] = L[4 + 2 * a + b];
                    Pn[a][b] = L[8 + 2 * a + b];
                }
        }
        atan_sum4_both<0, XM>(c, SA[0]);  // one merge test and one tier decision per axis
        atan_sum4_both<1, XM>(c, SA[1]);
        atan_sum4_both<2, XM>(c, SA[2]);
        if (FS == F_POT) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < 2; a++)
#pragma unroll
                for (int b = 0; b < 2; b++) {
                    const double sg = ((a + b) & 1) ? -1.0 : 1.0;
                    v += sg * (e[a] * n[b] * Pu[a][b] + n[a] * u[b] * Pe[a][b]
                               + e[a] * u[b] * Pn[a][b]);
                }
            v -= 0.5 * (c.se2[0] * SA[0][0] - c.se2[1] * SA[0][1]);
            v -= 0.5 * (c.sn2[0] * SA[1][0] - c.sn2[1] * SA[1][1]);
            v -= 0.5 * (c.su2[0] * SA[2][0] - c.su2[1] * SA[2][1]);
            acc[0] += prm[0] * v;
        } else {
            // E: sum_j (-1)^j n_j Gu(j) + sum_k (-1)^k u_k Gn(k) - sum_i (-1)^i e_i SAe[i]
            const double ve = n[0] * (Pu[0][0] - Pu[1][0]) - n[1] * (Pu[0][1] - Pu[1][1])
                            + u[0] * (Pn[0][0] - Pn[1][0]) - u[1] * (Pn[0][1] - Pn[1][1])
                            - (e[0] * SA[0][0] - e[1] * SA[0][1]);
            // N: sum_k (-1)^k u_k Ge(k) + sum_i (-1)^i e_i Gu(i) - sum_j (-1)^j n_j SAn[j]
            const double vn = u[0] * (Pe[0][0] - Pe[1][0]) - u[1] * (Pe[0][1] - Pe[1][1])
                            + e[0] * (Pu[0][0] - Pu[0][1]) - e[1] * (Pu[1][0] - Pu[1][1])
                            - (n[0] * SA[1][0] - n[1] * SA[1][1]);
            // U: sum_i (-1)^i e_i Gn(i) + sum_j (-1)^j n_j Ge(j) - sum_k (-1)^k u_k SAu[k]
            const double vu = e[0] * (Pn[0][0] - Pn[0][1]) - e[1] * (Pn[1][0] - Pn[1][1])
                            + n[0] * (Pe[0][0] - Pe[0][1]) - n[1] * (Pe[1][0] - Pe[1][1])
                            - (u[0] * SA[2][0] - u[1] * SA[2][1]);
            acc[0] += prm[0] * -ve;
            acc[1] += prm[0] * -vn;
            acc[2] += prm[0] * -vu;
        }
    } else {
        // second-derivative kernels: tensor components and magnetics
        double kee = 0, knn = 0, kuu = 0, ken = 0, keu = 0, knu = 0;
        constexpr bool fused = (FS == FS_TENSOR6 || FS == FS_MAG_B);
        if (XM && fused) {
            atan_sum8_two<0, 1, XM>(c, kee, knn);
            kee = -kee;
            knn = -knn;
        } else {
            if (T::ae) kee = -atan_sum8<0, XM>(c);
            if (T::an) knn = -atan_sum8<1, XM>(c);
        }
        if (T::au) {
            if (XM && T::ae && T::an) {
                // all three diagonal kernels wanted: the 8-vertex sums satisfy Laplace/Poisson
                // identically, k_ee + k_nn + k_uu = 0 outside and -4 pi inside the prism
                // (+4 pi per inverted axis), so the third one costs two additions instead of
                // four atan pair terms and an atan2. Boundary points never get here.
                const bool inside = (is_neg(e[0]) != is_neg(e[1])) && (is_neg(n[0]) != is_neg(n[1]))
                                    && (is_neg(u[0]) != is_neg(u[1]));
                const bool flipped = (is_neg(e[0]) != is_neg(n[0])) != is_neg(u[0]);
                const double trace = inside ? (flipped ? 4 * kPi : -4 * kPi) : 0.0;
                kuu = trace - (kee + knn);
            } else {
                kuu = -atan_sum8<2, XM>(c);
            }
        }
        if (XM && fused) {
            double top[3], bot[3], L[3];
            unsigned flip[3];
            log_sum8_tb<2, true>(c, top[0], bot[0], flip[0]);
            log_sum8_tb<1, true>(c, top[1], bot[1], flip[1]);
            log_sum8_tb<0, true>(c, top[2], bot[2], flip[2]);
            if (any_mixed(c)) {  // observer inside the prism's extent on some axis: redo that one
                if (log_mixed<2>(c)) log_sum8_tb<2, false>(c, top[0], bot[0], flip[0]);
                if (log_mixed<1>(c)) log_sum8_tb<1, false>(c, top[1], bot[1], flip[1]);
                if (log_mixed<0>(c)) log_sum8_tb<0, false>(c, top[2], bot[2], flip[2]);
            }
            x_log_ratio4<XM, 3>(top, bot, flip, L);
            ken = L[0];
            keu = L[1];
            knu = L[2];
        } else {
            if (T::lu) ken = log_sum8<2, XM>(c);
            if (T::ln) keu = log_sum8<1, XM>(c);
            if (T::le) knu = log_sum8<0, XM>(c);
        }
        if (cls == PAIR_FAST_CHECK) {
            // a shift is exactly zero (observer in the plane of a face): the reference's
            // outside-limit rule (K3) on top of the plain 8-vertex sum. Two zero shifts (edges,
            // vertices: the NaN rules) never reach this path.
            const PairPreds pr = make_preds(g);
            const bool face_rule = !T::mag || (mag_rules & MAG_FACE_OUTSIDE_LIMIT);
            if (face_rule) {
                kee += pr.east_face ? 4 * kPi : 0.0;
                knn += pr.north_face ? 4 * kPi : 0.0;
                kuu += pr.top_face ? 4 * kPi : 0.0;
            }
        }
        if (FS == F_EE) acc[0] += prm[0] * kee;
        else if (FS == F_NN) acc[0] += prm[0] * knn;
        else if (FS == F_UU) acc[0] += prm[0] * kuu;
        else if (FS == F_EN) acc[0] += prm[0] * ken;
        else if (FS == F_EU) acc[0] += prm[0] * keu;
        else if (FS == F_NU) acc[0] += prm[0] * knu;
        else if (FS == FS_TENSOR6) {
            acc[0] += prm[0] * kee; acc[1] += prm[0] * knn; acc[2] += prm[0] * kuu;
            acc[3] += prm[0] * ken; acc[4] += prm[0] * keu; acc[5] += prm[0] * knu;
        } else if (FS == FS_MAG_B) {
            acc[0] += prm[0] * kee + prm[1] * ken + prm[2] * keu;
            acc[1] += prm[0] * ken + prm[1] * knn + prm[2] * knu;
            acc[2] += prm[0] * keu + prm[1] * knu + prm[2] * kuu;
        } else if (FS == FS_MAG_E) acc[0] += prm[0] * kee + prm[1] * ken + prm[2] * keu;
        else if (FS == FS_MAG_N) acc[0] += prm[0] * ken + prm[1] * knn + prm[2] * knu;
        else if (FS == FS_MAG_U) acc[0] += prm[0] * keu + prm[1] * knu + prm[2] * kuu;
    }
}

// One (observer, prism) pair of kernel variant VARIANT: 0 = rule-exact direct evaluation for every
// pair (libm); 1 = merged path on CUDA libm; 2 = merged path on the library's own sequences.
template <int FS, int VARIANT>
HB_HD void prism_pair(const PairGeom& g, const double* prm, unsigned mag_rules, double* acc,
                      unsigned& flags, LayerCarry* carry = nullptr, bool reuse = false)
{
    if (VARIANT == 0) {
        prism_pair_direct<FS>(g, prm, mag_rules, acc, flags);
        return;
    }
    const int cls = classify_pair<FS>(g);
    if (cls == PAIR_EXACT) {
        prism_pair_direct<FS, (VARIANT == 2)>(g, prm, mag_rules, acc, flags);
        if (carry) carry->valid = false;
    } else {
        prism_pair_fast<FS, (VARIANT == 2)>(g, prm, acc, cls, mag_rules, &flags, carry, reuse);
        if (carry) carry->valid = true;
    }
}

}  // namespace hb

