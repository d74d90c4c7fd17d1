// hb200_math.cuh -- per-pair closed forms for right-rectangular prisms and
// point masses, written once as __host__ __device__ code.
//
// The device build (nvcc, sm_100a) is the product. The host build exists only
// so that tests/ can compile this very source with g++ into a test harness and
// compare the *same statements* with the oracle without a GPU; nothing in the
// product path ever calls the host build.
//
// What these functions replace in the reference (file:line under
// /root/reference/src/harmonica):
//   _forward/prisms/gravity.py:526-537    forward_func(...) = choclo.prism.gravity_*
//   _forward/prisms/magnetic.py:319-335   choclo.prism.magnetic_field
//   _forward/prisms/magnetic.py:384-397   choclo.prism.magnetic_{e,n,u}
//   _forward/point.py:390-398             choclo.point.gravity_*
//   _forward/point.py:324-354             potential_spherical / gravity_u_spherical
//   _equivalent_sources/cartesian.py:634-644  greens_func_cartesian
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HB_HD __device__ __forceinline__
#else
#define HB_HD inline
#endif

namespace hb {

// ---------------------------------------------------------------- field ids
enum : int {
    F_POT = 0, F_E = 1, F_N = 2, F_U = 3,
    F_EE = 4, F_NN = 5, F_UU = 6, F_EN = 7, F_EU = 8, F_NU = 9,
    FS_ACC3 = 10,     // g_e, g_n, g_u fused
    FS_TENSOR6 = 11,  // g_ee, g_nn, g_uu, g_en, g_eu, g_nu fused
    FS_MAG_B = 12,    // b_e, b_n, b_u fused
    FS_MAG_E = 13, FS_MAG_N = 14, FS_MAG_U = 15,
    FS_COUNT = 16
};

constexpr double kPi = 3.14159265358979323846;
constexpr double kG = 6.6743e-11;  // harmonica/constants.py:12 == choclo.constants

// rule switches of the magnetic kernels (parity unpinned)
constexpr unsigned MAG_NAN_ON_EDGES = 1u;
constexpr unsigned MAG_FACE_OUTSIDE_LIMIT = 2u;

// flags reported back to the host
constexpr unsigned FLAG_SINGULAR = 1u;  // an (observer, prism) pair hit a NaN rule
constexpr unsigned FLAG_ZERO_DIV = 2u;  // coincident observer / point source

template <int FS> struct Traits;
#define HB_TRAITS(FS_, NOUT_, LE_, LN_, LU_, AE_, AN_, AU_)                         \
    template <> struct Traits<FS_> {                                               \
        static constexpr int nout = NOUT_;                                         \
        static constexpr bool le = LE_, ln = LN_, lu = LU_, ae = AE_, an = AN_, au = AU_; \
        static constexpr bool mag = (FS_ >= FS_MAG_B);                             \
        static constexpr int nparam = mag ? 3 : 1;                                 \
    }
//                 nout  Le Ln Lu Ae An Au       (Lx = safe_log(x; ...), Ax = atan term with x*r)
HB_TRAITS(F_POT,      1, 1, 1, 1, 1, 1, 1);
HB_TRAITS(F_E,        1, 0, 1, 1, 1, 0, 0);
HB_TRAITS(F_N,        1, 1, 0, 1, 0, 1, 0);
HB_TRAITS(F_U,        1, 1, 1, 0, 0, 0, 1);
HB_TRAITS(F_EE,       1, 0, 0, 0, 1, 0, 0);
HB_TRAITS(F_NN,       1, 0, 0, 0, 0, 1, 0);
HB_TRAITS(F_UU,       1, 0, 0, 0, 0, 0, 1);
HB_TRAITS(F_EN,       1, 0, 0, 1, 0, 0, 0);
HB_TRAITS(F_EU,       1, 0, 1, 0, 0, 0, 0);
HB_TRAITS(F_NU,       1, 1, 0, 0, 0, 0, 0);
HB_TRAITS(FS_ACC3,    3, 1, 1, 1, 1, 1, 1);
HB_TRAITS(FS_TENSOR6, 6, 1, 1, 1, 1, 1, 1);
HB_TRAITS(FS_MAG_B,   3, 1, 1, 1, 1, 1, 1);
HB_TRAITS(FS_MAG_E,   1, 0, 1, 1, 1, 0, 0);
HB_TRAITS(FS_MAG_N,   1, 1, 0, 1, 0, 1, 0);
HB_TRAITS(FS_MAG_U,   1, 1, 1, 0, 0, 0, 1);
#undef HB_TRAITS

// ------------------------------------------------------------ bit helpers
HB_HD int hi_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
HB_HD int lo_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return (int)(b & 0xffffffff);
#endif
}
HB_HD double make_double(int hi, int lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    int64_t b = ((int64_t)hi << 32) | (uint32_t)lo;
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}
// ------------------------------------------------- exact (uncontracted) ops
// r and y^2+z^2 decide which branch of safe_log is taken (r == 0, r == -x), so
// they are computed with the reference's rounding sequence: separately rounded
// squares and sums, never fused.
HB_HD double mul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;  // host harness is compiled with -ffp-contract=off
#endif
}
HB_HD double add_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// ------------------------------------------------------------ reference rules
// choclo.prism._utils.safe_atan2
HB_HD double safe_atan2(double y, double x)
{
    double a = atan(y / x);
    double lim = (y > 0.0) ? (kPi / 2) : ((y < 0.0) ? (-kPi / 2) : 0.0);
    return (x != 0.0) ? a : lim;
}

// choclo.prism._utils.safe_log with yz2 = y*y + z*z precomputed (K2)
HB_HD double safe_log(double x, double yz2, double r)
{
    const bool neg = x < 0.0;
    const bool axis = neg && (r == -x);
    double arg = neg ? (axis ? -2.0 * x : yz2 / (r - x)) : (x + r);
    double l = log(arg);
    l = axis ? -l : l;
    return (r == 0.0) ? 0.0 : l;
}

// The same two rules on the library's own sequences (hb200_xmath.cuh, defined after this
// header): used by the rule-exact path of kernel variant 2. Values agree with the libm versions
// to ~1e-16; every branch decision (x == 0, r == 0, r == -x, signs) is identical.
HB_HD double fast_log(double a);
HB_HD double fast_atan2(double y, double x);
HB_HD double fast_rcp(double x);
template <bool XM> HB_HD double safe_atan2_t(double y, double x)
{
    if (!XM) return safe_atan2(y, x);
    const double a = fast_atan2(x < 0.0 ? -y : y, fabs(x));  // atan(y / x) for x != 0
    const double lim = (y > 0.0) ? (kPi / 2) : ((y < 0.0) ? (-kPi / 2) : 0.0);
    return (x != 0.0) ? a : lim;
}
template <bool XM> HB_HD double safe_log_t(double x, double yz2, double r)
{
    if (!XM) return safe_log(x, yz2, r);
    const bool neg = x < 0.0;
    const bool axis = neg && (r == -x);
    const double arg = neg ? (axis ? -2.0 * x : yz2 * fast_rcp(r - x)) : (x + r);
    double l = fast_log(arg);
    l = axis ? -l : l;
    return (r == 0.0) ? 0.0 : l;
}

// Geometry of one (observer, prism) pair: shifted coordinates and the partial
// sums every kernel shares. Index 0 = east/north/top, 1 = west/south/bottom
// (the reference's vertex order).
struct PairGeom {
    double se[2], sn[2], su[2];
    double se2[2], sn2[2], su2[2];
};

HB_HD void make_geom(PairGeom& g, double E, double N, double U, double w, double e, double s,
                     double n, double b, double t)
{
    g.se[0] = e - E; g.se[1] = w - E;
    g.sn[0] = n - N; g.sn[1] = s - N;
    g.su[0] = t - U; g.su[1] = b - U;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        g.se2[i] = mul_rn(g.se[i], g.se[i]);
        g.sn2[i] = mul_rn(g.sn[i], g.sn[i]);
        g.su2[i] = mul_rn(g.su[i], g.su[i]);
    }
}

// Exact singular-point predicates (choclo.prism._utils.is_point_on_*_edge,
// K6) from the shifted coordinates: a - b == 0 iff a == b and the sign of an
// exact-rounded difference is the sign of the true difference.
struct PairPreds {
    bool e_edge, n_edge, u_edge;       // on an edge parallel to easting / northing / upward
    bool east_face, north_face, top_face;
};
HB_HD PairPreds make_preds(const PairGeom& g)
{
    const bool eqE = (g.se[0] == 0.0) | (g.se[1] == 0.0);
    const bool eqN = (g.sn[0] == 0.0) | (g.sn[1] == 0.0);
    const bool eqU = (g.su[0] == 0.0) | (g.su[1] == 0.0);
    const bool inE = (g.se[1] <= 0.0) & (g.se[0] >= 0.0);
    const bool inN = (g.sn[1] <= 0.0) & (g.sn[0] >= 0.0);
    const bool inU = (g.su[1] <= 0.0) & (g.su[0] >= 0.0);
    PairPreds p;
    p.e_edge = inE & eqN & eqU;
    p.n_edge = inN & eqE & eqU;
    p.u_edge = inU & eqE & eqN;
    p.east_face = (g.se[0] == 0.0) & inN & inU;
    p.north_face = (g.sn[0] == 0.0) & inE & inU;
    p.top_face = (g.su[0] == 0.0) & inE & inN;
    return p;
}

// Which evaluation a pair gets. With h = exponent bits of the six squared shifts (non-negative
// doubles order like their bit patterns: integer pipe only):
//
//  PAIR_EXACT  the rule-exact (direct) path: (a) the MEDIAN of the three per-axis minima is zero
//    or > 2^50 times smaller than the largest square, i.e. TWO axes carry a (nearly) zero shift:
//    the observer is on (the extension of) an edge or a vertex, where the reference's on-axis
//    safe_log branch (r == -x needs y^2 + z^2 < 2^-52 x^2) and r == 0 can fire and the NaN rules
//    of the tensor / magnetic kernels apply; (b) pairs at absurd length scales (the merged
//    products reach the 16th power of a length; beyond ~1e-18 .. 1e18 m the reference's
//    formulation is used as well).
//  PAIR_FAST / PAIR_FAST_CHECK  the merged path. ONE axis may hold a zero (or arbitrarily small)
//    shift: for every vertex and every safe_log type y^2 + z^2 then contains a square of a
//    constrained axis, hence >= 2^-50 x^2: no on-axis branch, r > 0, all merged products
//    positive; the atan pair terms degenerate correctly (im = +-0 with the right sign, re != 0).
//    Observers in the plane of a prism face -- stations on flat prism tops, grids aligned with
//    the prism edges -- therefore stay on the merged path for EVERY field. What the reference
//    does there beyond the plain 8-vertex sum is a function of the shifts' zero / sign pattern
//    only (PairPreds: +4 pi on the east / north / top face for the face-normal diagonal
//    component; NaN needs two zero shifts and cannot occur here), and is applied after the merged
//    evaluation when the smallest square is zero: PAIR_FAST_CHECK.
enum : int { PAIR_FAST = 0, PAIR_FAST_CHECK = 1, PAIR_EXACT = 2 };

template <int FS> HB_HD int classify_pair(const PairGeom& g)
{
    const unsigned e0 = (unsigned)hi_word(g.se2[0]), e1 = (unsigned)hi_word(g.se2[1]);
    const unsigned n0 = (unsigned)hi_word(g.sn2[0]), n1 = (unsigned)hi_word(g.sn2[1]);
    const unsigned u0 = (unsigned)hi_word(g.su2[0]), u1 = (unsigned)hi_word(g.su2[1]);
    const unsigned he = e0 < e1 ? e0 : e1, hn = n0 < n1 ? n0 : n1, hu = u0 < u1 ? u0 : u1;
    const unsigned me = e0 < e1 ? e1 : e0, mn = n0 < n1 ? n1 : n0, mu = u0 < u1 ? u1 : u0;
    const unsigned men = me > mn ? me : mn;
    const unsigned hi = men > mu ? men : mu;
    const unsigned lo_en = he < hn ? he : hn, hi_en = he < hn ? hn : he;
    const unsigned t = hi_en < hu ? hi_en : hu;
    const unsigned med = lo_en > t ? lo_en : t;    // median of (he, hn, hu)
    const unsigned low = lo_en < hu ? lo_en : hu;  // minimum
    if ((med + (50u << 20) < hi) | (hi - 0x38800000u > 0x47600000u - 0x38800000u)) return PAIR_EXACT;
    constexpr bool has_rules = (FS >= F_EE && FS <= F_NU) || FS == FS_TENSOR6 || FS >= FS_MAG_B;
    // a square below the normal range: the shift may be exactly zero
    return (has_rules && low < 0x00100000u) ? PAIR_FAST_CHECK : PAIR_FAST;
}
template <int FS> HB_HD bool needs_exact_path(const PairGeom& g)
{
    return classify_pair<FS>(g) == PAIR_EXACT;
}

// NaN rule per field set (gravity.py:272-449 predicate sets == choclo's)
template <int FS> HB_HD bool nan_rule(const PairPreds& p, unsigned mag_rules)
{
    if (FS == F_EE) return p.n_edge | p.u_edge;
    if (FS == F_NN) return p.e_edge | p.u_edge;
    if (FS == F_UU) return p.e_edge | p.n_edge;
    if (FS == F_EN) return p.u_edge;
    if (FS == F_EU) return p.n_edge;
    if (FS == F_NU) return p.e_edge;
    if (FS >= FS_MAG_B) return (mag_rules & MAG_NAN_ON_EDGES) && (p.e_edge | p.n_edge | p.u_edge);
    return false;
}

// ------------------------------------------------------------- direct path
// Statement-for-statement the reference's 8-vertex sum (K1, K3, K4). Used for
// pairs with an exactly-zero shifted coordinate, and as the cross-check of
// the merged path. `prm` = {G*rho} for gravity, {me, mn, mu} for magnetics.
// Adds this pair's contribution to acc[0..nout).
template <int FS, bool XM = false>
HB_HD void prism_pair_direct(const PairGeom& g, const double* prm, unsigned mag_rules, double* acc,
                             unsigned& flags)
{
    typedef Traits<FS> T;
    const PairPreds pr = make_preds(g);
    double sum[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // per-vertex alternating sums
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const double en2 = add_rn(g.se2[i], g.sn2[j]);
#pragma unroll
            for (int k = 0; k < 2; k++) {
                const double e = g.se[i], n = g.sn[j], u = g.su[k];
                const double r = sqrt(add_rn(en2, g.su2[k]));
                const double sg = ((i + j + k) & 1) ? -1.0 : 1.0;
                double Le = 0, Ln = 0, Lu = 0, Ae = 0, An = 0, Au = 0;
                if (T::le) Le = safe_log_t<XM>(e, add_rn(g.sn2[j], g.su2[k]), r);
                if (T::ln) Ln = safe_log_t<XM>(n, add_rn(g.se2[i], g.su2[k]), r);
                if (T::lu) Lu = safe_log_t<XM>(u, en2, r);
                if (T::ae) Ae = safe_atan2_t<XM>(n * u, e * r);
                if (T::an) An = safe_atan2_t<XM>(e * u, n * r);
                if (T::au) Au = safe_atan2_t<XM>(e * n, u * r);
                if (FS == F_POT) {
                    sum[0] += sg * (e * n * Lu + n * u * Le + e * u * Ln - 0.5 * g.se2[i] * Ae
                                    - 0.5 * g.sn2[j] * An - 0.5 * g.su2[k] * Au);
                } else if (FS == F_E) {
                    sum[0] += sg * -(n * Lu + u * Ln - e * Ae);
                } else if (FS == F_N) {
                    sum[0] += sg * -(u * Le + e * Lu - n * An);
                } else if (FS == F_U) {
                    sum[0] += sg * -(e * Ln + n * Le - u * Au);
                } else if (FS == F_EE) {
                    sum[0] += sg * -Ae;
                } else if (FS == F_NN) {
                    sum[0] += sg * -An;
                } else if (FS == F_UU) {
                    sum[0] += sg * -Au;
                } else if (FS == F_EN) {
                    sum[0] += sg * Lu;
                } else if (FS == F_EU) {
                    sum[0] += sg * Ln;
                } else if (FS == F_NU) {
                    sum[0] += sg * Le;
                } else if (FS == FS_ACC3) {
                    sum[0] += sg * -(n * Lu + u * Ln - e * Ae);
                    sum[1] += sg * -(u * Le + e * Lu - n * An);
                    sum[2] += sg * -(e * Ln + n * Le - u * Au);
                } else if (FS == FS_TENSOR6) {
                    sum[0] += sg * -Ae; sum[1] += sg * -An; sum[2] += sg * -Au;
                    sum[3] += sg * Lu;  sum[4] += sg * Ln;  sum[5] += sg * Le;
                } else if (FS == FS_MAG_B) {
                    sum[0] += sg * (prm[0] * -Ae + prm[1] * Lu + prm[2] * Ln);
                    sum[1] += sg * (prm[0] * Lu + prm[1] * -An + prm[2] * Le);
                    sum[2] += sg * (prm[0] * Ln + prm[1] * Le + prm[2] * -Au);
                } else if (FS == FS_MAG_E) {
                    sum[0] += sg * (prm[0] * -Ae + prm[1] * Lu + prm[2] * Ln);
                } else if (FS == FS_MAG_N) {
                    sum[0] += sg * (prm[0] * Lu + prm[1] * -An + prm[2] * Le);
                } else if (FS == FS_MAG_U) {
                    sum[0] += sg * (prm[0] * Ln + prm[1] * Le + prm[2] * -Au);
                }
            }
        }
    }
    // outside-limit rule on the faces whose outward normal is +e / +n / +u (K3)
    const double four_pi = 4 * kPi;
    if (FS == F_EE && pr.east_face) sum[0] += four_pi;
    if (FS == F_NN && pr.north_face) sum[0] += four_pi;
    if (FS == F_UU && pr.top_face) sum[0] += four_pi;
    if (FS == FS_TENSOR6) {
        if (pr.east_face) sum[0] += four_pi;
        if (pr.north_face) sum[1] += four_pi;
        if (pr.top_face) sum[2] += four_pi;
    }
    if (T::mag && (mag_rules & MAG_FACE_OUTSIDE_LIMIT)) {
        if ((FS == FS_MAG_B || FS == FS_MAG_E) && pr.east_face) sum[0] += prm[0] * four_pi;
        if (FS == FS_MAG_B && pr.north_face) sum[1] += prm[1] * four_pi;
        if (FS == FS_MAG_B && pr.top_face) sum[2] += prm[2] * four_pi;
        if (FS == FS_MAG_N && pr.north_face) sum[0] += prm[1] * four_pi;
        if (FS == FS_MAG_U && pr.top_face) sum[0] += prm[2] * four_pi;
    }
    const double qnan = NAN;
    if (FS == FS_TENSOR6) {
        // each component has its own NaN predicate set
        const bool s[6] = {pr.n_edge || pr.u_edge, pr.e_edge || pr.u_edge, pr.e_edge || pr.n_edge,
                           pr.u_edge, pr.n_edge, pr.e_edge};
#pragma unroll
        for (int c = 0; c < 6; c++) {
            acc[c] += s[c] ? qnan : prm[0] * sum[c];
            if (s[c]) flags |= FLAG_SINGULAR;
        }
    } else {
        const bool sing = nan_rule<FS>(pr, mag_rules);
        if (sing) flags |= FLAG_SINGULAR;
        const double scale = T::mag ? 1.0 : prm[0];
#pragma unroll
        for (int c = 0; c < T::nout; c++) acc[c] += sing ? qnan : scale * sum[c];
    }
}

// -------------------------------------------------------------- point masses
// d2 is a sum of squares: zero means +0.0 exactly; tested on the integer pipe
HB_HD bool is_pos_zero(double x)
{
#if defined(__CUDA_ARCH__)
    return (__double2hiint(x) | __double2loint(x)) == 0;
#else
    return x == 0.0;
#endif
}
HB_HD double point_rsqrt(double d2);  // defined in hb200_xmath.cuh (MUFU.RSQ64H + one cubic step)

// choclo.point kernels (K5) without the G*mass factor; d = observer - source.
// returns kernel value(s); zero distance reported through flags.
HB_HD double point_d2(double de, double dn, double du) { return de * de + dn * dn + du * du; }

// CHECK = false: the caller looks for zero distances itself (point_kernel_cart does so only when a
// sum came out non-finite: 1 / 0 poisons it).
template <int FIELD, bool CHECK = true>
HB_HD double point_kernel(double de, double dn, double du, unsigned& flags)
{
    const double d2 = point_d2(de, dn, du);
    if (CHECK && is_pos_zero(d2)) flags |= FLAG_ZERO_DIV;
    const double inv = point_rsqrt(d2);
    if (FIELD == F_POT) return inv;
    const double inv3 = inv * inv * inv;
    if (FIELD == F_E) return -de * inv3;
    if (FIELD == F_N) return -dn * inv3;
    if (FIELD == F_U) return -du * inv3;
    const double inv5 = inv3 * inv * inv;
    if (FIELD == F_EE) return 3.0 * de * de * inv5 - inv3;
    if (FIELD == F_NN) return 3.0 * dn * dn * inv5 - inv3;
    if (FIELD == F_UU) return 3.0 * du * du * inv5 - inv3;
    if (FIELD == F_EN) return 3.0 * de * dn * inv5;
    if (FIELD == F_EU) return 3.0 * de * du * inv5;
    return 3.0 * dn * du * inv5;  // F_NU
}

}  // namespace hb
