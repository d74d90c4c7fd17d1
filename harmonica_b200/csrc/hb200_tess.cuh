// hb200_tess.cuh -- tesseroid forward model: adaptive discretisation of every
// (observer, tesseroid) pair with a per-thread stack, then 2 x 2 x 2 Gauss-Legendre point masses
// through the spherical point kernels. Written once as __host__ __device__ code (the host build is
// test infrastructure only, like hb200_math.cuh).
//
// Reference (file:line under /root/reference/src/harmonica/_forward):
//   tesseroid_gravity.py:305-339     jit_tesseroid_gravity (the pair loop)
//   _tesseroid_utils.py:136-217      _adaptive_discretization
//   _tesseroid_utils.py:220-300      _split_tesseroid, _tesseroid_dimensions, _distance_tesseroid_point
//   _tesseroid_utils.py:19-107       gauss_legendre_quadrature
//   utils.py:164-201                 distance_spherical_core
//   point.py:324-354                 potential_spherical, gravity_u_spherical
//
// The SPLIT DECISIONS (distance / size < ratio) are evaluated with the reference's expressions in
// the reference's order, so that the set of leaves is the reference's (a decision can only flip
// when distance / size equals the ratio to within the rounding of sin / cos / acos). What is
// restructured is decision-neutral and value-identical:
//   * the leaves are integrated as they are popped instead of being collected first;
//   * the observer's radians / cos / sin are hoisted out of the pair loop, the two distinct
//     cos(longitude_p - longitude) of a leaf are computed once instead of for all eight nodes;
//   * everything about a ROOT tesseroid that does not depend on the observer (its dimensions, the
//     trig of its centre, its eight quadrature nodes and their masses) is computed once per
//     tesseroid by the pack kernel with the same statements (tess_pack_record) instead of once
//     per pair: 98 % of the pairs of a typical model never split, for them the pair loop is left
//     with three cosines, nine square roots and the divisions;
//   * (kernel variants 1, 2) pairs that do split are deferred and walked later by all lanes of the warp
//     concurrently, each lane through its own list, instead of one lane at a time while the
//     others wait (the first build ran with 19.8 of 32 lanes active, profiles/).
// Only the ORDER in which a thread adds its pairs changes with the last item.
#pragma once
#include "hb200_math.cuh"
#include "hb200_trig.cuh"
#include "hb200_xmath.cuh"

namespace hb {

constexpr int kTessStack = 100;          // tesseroid_gravity.py:30  STACK_SIZE
constexpr int kTessMaxLeaves = 100000;   // tesseroid_gravity.py:31  MAX_DISCRETIZATIONS
constexpr int kTessStride = 8;           // plain record: w e s n bottom top density -
constexpr int kTessRec = 32;             // root record with the observer-independent parts
constexpr unsigned FLAG_TESS_STACK = 4u;    // "Stack Overflow. Try to increase the stack size."
constexpr unsigned FLAG_TESS_LEAVES = 8u;   // "Exceeded maximum discretizations."
constexpr unsigned FLAG_TESS_INSIDE = 16u;  // a computation point lies inside a tesseroid

// _tesseroid_utils.py:431-454 (_check_points_outside_tesseroids): strictly inside, with the
// longitude tried in [0, 360) and in [-180, 180). Python's % takes the sign of the divisor.
HB_HD double tess_pymod360(double x)
{
    double r = fmod(x, 360.0);
    if (r != 0.0 && r < 0.0) r += 360.0;
    return r;
}

HB_HD bool tess_point_inside(double lon, double lat, double rad, const double* t)
{
    const double longitude_360 = tess_pymod360(lon);
    const double longitude_180 = tess_pymod360(lon + 180) - 180;
    const bool in_lon = (t[0] < longitude_180 && longitude_180 < t[1])
                     || (t[0] < longitude_360 && longitude_360 < t[1]);
    return in_lon && t[2] < lat && lat < t[3] && t[4] < rad && rad < t[5];
}

// numpy.polynomial.legendre.leggauss(2): nodes -/+ 1/sqrt(3) (0x1.279a74590331cp-1), weights 1
constexpr double kGlqNode = 0.5773502691896257;
constexpr double kDeg2Rad = kPi / 180.0;  // np.radians multiplies by pi / 180

struct TessObs {
    double lon, lat, rad;      // degrees, degrees, metres (as given)
    double lam, cphi, sphi;    // radians(lon), cos / sin of radians(lat)
    double clam, slam;         // cos / sin of lam (fast root path only)
};

HB_HD void tess_make_obs(TessObs& o, double lon, double lat, double rad)
{
    o.lon = lon;
    o.lat = lat;
    o.rad = rad;
    o.lam = lon * kDeg2Rad;
    const double phi = lat * kDeg2Rad;
    o.cphi = cos(phi);
    o.sphi = sin(phi);
    o.clam = cos(o.lam);
    o.slam = sin(o.lam);
}

// Where the walk takes its sin / cos / acos from: CUDA's (glibc's in the host build) or the
// library's own bounded-angle sequences (hb200_trig.cuh; kernel variant 3).
struct LibmTrig {
    static HB_HD double sin_(double x) { return sin(x); }
    static HB_HD double cos_(double x) { return cos(x); }
    static HB_HD void sincos_(double x, double& s, double& c) { s = sin(x); c = cos(x); }
    static HB_HD double acos_(double x) { return acos(x); }
};
struct OwnTrig {
    static HB_HD double sin_(double x) { double s, c; fast_sincos(x, s, c); return s; }
    static HB_HD double cos_(double x) { return fast_cos(x); }
    static HB_HD void sincos_(double x, double& s, double& c) { fast_sincos(x, s, c); }
    static HB_HD double acos_(double x) { return fast_acos(x); }
};

// ---- the observer-independent parts of one tesseroid -----------------------------------------
struct TessDims {
    double l_lon, l_lat, l_rad;  // _tesseroid_dimensions
};
struct TessCentre {
    double lam, cphi, sphi, rad;  // radians(longitude), cos / sin of radians(latitude), radius
};
struct TessNodes {
    double lam[2];              // GLQ longitude nodes, radians
    double cphi[2], sphi[2];    // GLQ latitude nodes
    double rad[2];              // GLQ radial nodes
    double mass[2][2];          // [lat node][radial node]: density * a_factor * kappa (weights 1)
};

// _tesseroid_utils.py:261-279
template <class TRIG = LibmTrig>
HB_HD void tess_dims(TessDims& d, double w, double e, double s, double n, double bottom, double top)
{
    const double wr = w * kDeg2Rad, er = e * kDeg2Rad, sr = s * kDeg2Rad, nr = n * kDeg2Rad;
    const double latitude_center = (nr + sr) / 2;
    double sin_n, cos_n, sin_s, cos_s, sc, cc;
    TRIG::sincos_(nr, sin_n, cos_n);
    TRIG::sincos_(sr, sin_s, cos_s);
    d.l_lat = top * TRIG::acos_(sin_n * sin_s + cos_n * cos_s);
    TRIG::sincos_(latitude_center, sc, cc);
    d.l_lon = top * TRIG::acos_(sc * sc + cc * cc * TRIG::cos_(er - wr));
    d.l_rad = top - bottom;
}

// the point _distance_tesseroid_point (:282-300) measures to, as distance_spherical sees it
template <class TRIG = LibmTrig>
HB_HD void tess_centre(TessCentre& c, double w, double e, double s, double n, double bottom,
                       double top)
{
    c.lam = ((w + e) / 2) * kDeg2Rad;
    const double latitude_p = ((s + n) / 2) * kDeg2Rad;
    c.rad = (bottom + top) / 2;
    TRIG::sincos_(latitude_p, c.sphi, c.cphi);
}

// utils.py:164-201 between the observer and the centre
template <class TRIG = LibmTrig> HB_HD double tess_distance(const TessObs& o, const TessCentre& c)
{
    const double coslambda = TRIG::cos_(c.lam - o.lam);
    const double cospsi = c.sphi * o.sphi + c.cphi * o.cphi * coslambda;
    const double dr = o.rad - c.rad;
    return sqrt(dr * dr + 2 * o.rad * c.rad * (1 - cospsi));
}

// _tesseroid_utils.py:186-191. Returns false where numba's float division raises
// ZeroDivisionError (all three quotients are evaluated; a child so small that acos(...) == 0).
HB_HD bool tess_split_counts(double distance, const TessDims& d, double ratio, bool radial,
                             int& n_lon, int& n_lat, int& n_rad)
{
    n_lon = n_lat = n_rad = 1;
    if (d.l_lon == 0.0 || d.l_lat == 0.0 || d.l_rad == 0.0) return false;
    n_lon = (distance / d.l_lon < ratio) ? 2 : 1;
    n_lat = (distance / d.l_lat < ratio) ? 2 : 1;
    n_rad = (distance / d.l_rad < ratio && radial) ? 2 : 1;
    return true;
}

// the node coordinates and masses of gauss_legendre_quadrature (:19-107). density[k] belongs to
// the radial node k: equal for a homogeneous tesseroid; density(radius_p) of the variable-density
// quadrature (_tesseroid_variable_density.py:20-106) otherwise.
template <class TRIG = LibmTrig>
HB_HD void tess_nodes(TessNodes& q, double w, double e, double s, double n, double bottom,
                      double top, const double* density)
{
    const double a_factor = 1.0 / 8 * ((e - w) * kDeg2Rad) * ((n - s) * kDeg2Rad) * (top - bottom);
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double node = i ? kGlqNode : -kGlqNode;
        q.lam[i] = (0.5 * (e - w) * node + 0.5 * (e + w)) * kDeg2Rad;
        const double latitude_p = (0.5 * (n - s) * node + 0.5 * (n + s)) * kDeg2Rad;
        TRIG::sincos_(latitude_p, q.sphi[i], q.cphi[i]);
        q.rad[i] = 0.5 * (top - bottom) * node + 0.5 * (top + bottom);
    }
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double kappa = q.rad[k] * q.rad[k] * q.cphi[j];
            q.mass[j][k] = density[k] * a_factor * kappa;  // the three GLQ weights are 1
        }
}

// sum over the eight nodes in the reference's order (latitude, radius, longitude) with the
// kernels of point.py:324-354. FIELD: F_POT or F_U (radial component).
template <int FIELD, class TRIG = LibmTrig>
HB_HD double tess_glq_nodes(const TessObs& o, const TessNodes& q, unsigned& flags)
{
    double coslambda[2];
#pragma unroll
    for (int i = 0; i < 2; i++) coslambda[i] = TRIG::cos_(q.lam[i] - o.lam);
    double result = 0.0;
#pragma unroll
    for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double radius_p = q.rad[k];
            const double dr = o.rad - radius_p;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double cospsi = q.sphi[j] * o.sphi + q.cphi[j] * o.cphi * coslambda[i];
                const double dist = sqrt(dr * dr + 2 * o.rad * radius_p * (1 - cospsi));
                double kern;
                if (dist == 0.0) flags |= FLAG_ZERO_DIV;  // observer on a quadrature node
                if (FIELD == F_POT) {
                    kern = 1 / dist * kG;
                } else {
                    const double delta_z = o.rad - radius_p * cospsi;
                    kern = -kG * delta_z / (dist * dist * dist);
                }
                result += q.mass[j][k] * kern;
            }
        }
    return result;
}

// The decision of one pop of _adaptive_discretization (:178-191): dimensions, distance to the
// centre, split counts. Returns false where numba raises ZeroDivisionError.
template <class TRIG = LibmTrig>
HB_HD bool tess_classify(const TessObs& o, double ratio, bool radial, double w, double e, double s,
                         double n, double bottom, double top, int& n_lon, int& n_lat, int& n_rad)
{
    TessDims dims;
    tess_dims<TRIG>(dims, w, e, s, n, bottom, top);
    TessCentre centre;
    tess_centre<TRIG>(centre, w, e, s, n, bottom, top);
    const double distance = tess_distance<TRIG>(o, centre);
    return tess_split_counts(distance, dims, ratio, radial, n_lon, n_lat, n_rad);
}

// ---- the walk over the adaptive discretisation of one pair --------------------------------------
struct TessWalk {
    int stack_top;   // < 0: no pair in progress
    int n_leaves;
    double density[2];  // per radial quadrature node
};

HB_HD void tess_walk_begin(TessWalk& w, const double* tess, double density0, double density1,
                           double* stack)
{
#pragma unroll
    for (int c = 0; c < 6; c++) stack[c] = tess[c];
    w.stack_top = 0;
    w.n_leaves = 0;
    w.density[0] = density0;
    w.density[1] = density1;
}

// One pop of _adaptive_discretization (:178-216): either pushes the children (_split_tesseroid)
// or integrates the leaf into `acc`. Where the reference raises, a flag is set and the walk of
// this pair ends. `stack` holds STACK x 6 doubles and is private to the caller. (STACK /
// MAX_LEAVES are template parameters only so that the tests can provoke the overflow errors like
// the reference's do.)
template <int FIELD, int STACK = kTessStack, int MAX_LEAVES = kTessMaxLeaves, class TRIG = LibmTrig>
HB_HD void tess_walk_step(const TessObs& o, double ratio, bool radial, double* stack, TessWalk& wk,
                          double& acc, unsigned& flags)
{
    const double* q = stack + 6 * wk.stack_top;
    const double w = q[0], e = q[1], s = q[2], n = q[3], bottom = q[4], top = q[5];
    wk.stack_top -= 1;
    int n_lon, n_lat, n_rad;
    if (!tess_classify<TRIG>(o, ratio, radial, w, e, s, n, bottom, top, n_lon, n_lat, n_rad)) {
        flags |= FLAG_ZERO_DIV;
        wk.stack_top = -1;
        return;
    }
    if (n_lon * n_lat * n_rad > 1) {
        if ((wk.stack_top + 1) + n_lon * n_lat * n_rad > STACK) {
            flags |= FLAG_TESS_STACK;
            wk.stack_top = -1;
            return;
        }
        // _split_tesseroid
        const double d_lon = (e - w) / n_lon, d_lat = (n - s) / n_lat, d_rad = (top - bottom) / n_rad;
        for (int i = 0; i < n_lon; i++)
            for (int j = 0; j < n_lat; j++)
                for (int k = 0; k < n_rad; k++) {
                    wk.stack_top += 1;
                    double* c = stack + 6 * wk.stack_top;
                    c[0] = w + d_lon * i;
                    c[1] = w + d_lon * (i + 1);
                    c[2] = s + d_lat * j;
                    c[3] = s + d_lat * (j + 1);
                    c[4] = bottom + d_rad * k;
                    c[5] = bottom + d_rad * (k + 1);
                }
    } else {
        if (wk.n_leaves + 1 > MAX_LEAVES) {
            flags |= FLAG_TESS_LEAVES;
            wk.stack_top = -1;
            return;
        }
        TessNodes nodes;
        tess_nodes<TRIG>(nodes, w, e, s, n, bottom, top, wk.density);
        acc += tess_glq_nodes<FIELD, TRIG>(o, nodes, flags);
        wk.n_leaves += 1;
    }
}

// One (observer, tesseroid) pair: adds the quadrature of every leaf of the adaptive
// discretisation to `acc` in the reference's order. Returns the number of leaves.
template <int FIELD, int STACK = kTessStack, int MAX_LEAVES = kTessMaxLeaves, class TRIG = LibmTrig>
HB_HD int tess_pair(const TessObs& o, const double* tess, double density0, double density1,
                    double ratio, bool radial, double* stack, double& acc, unsigned& flags)
{
    TessWalk wk;
    tess_walk_begin(wk, tess, density0, density1, stack);
    while (wk.stack_top >= 0)
        tess_walk_step<FIELD, STACK, MAX_LEAVES, TRIG>(o, ratio, radial, stack, wk, acc, flags);
    return wk.n_leaves;
}

// ---- root record -----------------------------------------------------------------------------------
// [0..5] w e s n bottom top  [6] density  [7..9] l_lon l_lat l_rad
// [10..13] centre: lam cphi sphi rad   [14,15] node lam   [16,17] node cphi   [18,19] node sphi
// [20,21] node rad   [22..25] mass[j][k]   [26] density of the upper radial node ([6]: lower)
constexpr int kTessRho1 = 26;      // slot of the second density in a root record
constexpr int kTessRho1Fast = 30;  // ... in a fast root record
HB_HD void tess_pack_record(double* rec, const double* tess, double density0, double density1)
{
#pragma unroll
    for (int c = 0; c < 6; c++) rec[c] = tess[c];
    rec[6] = density0;
    const double density[2] = {density0, density1};
    TessDims d;
    tess_dims(d, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5]);
    rec[7] = d.l_lon; rec[8] = d.l_lat; rec[9] = d.l_rad;
    TessCentre c;
    tess_centre(c, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5]);
    rec[10] = c.lam; rec[11] = c.cphi; rec[12] = c.sphi; rec[13] = c.rad;
    TessNodes q;
    tess_nodes(q, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5], density);
    rec[14] = q.lam[0]; rec[15] = q.lam[1];
    rec[16] = q.cphi[0]; rec[17] = q.cphi[1];
    rec[18] = q.sphi[0]; rec[19] = q.sphi[1];
    rec[20] = q.rad[0]; rec[21] = q.rad[1];
    rec[22] = q.mass[0][0]; rec[23] = q.mass[0][1]; rec[24] = q.mass[1][0]; rec[25] = q.mass[1][1];
    rec[kTessRho1] = density1;
#pragma unroll
    for (int c2 = 27; c2 < kTessRec; c2++) rec[c2] = 0.0;
}

// The root pop of a pair from its record. Returns 1 when the root is a leaf (integrated into
// `acc`), 0 when it splits (the caller walks it with tess_pair / tess_walk_step, which repeat the
// root decision with the same values) and -1 where the reference raises (flag set).
template <int FIELD>
HB_HD int tess_root(const TessObs& o, const double* rec, double ratio, bool radial, double& acc,
                    unsigned& flags)
{
    TessDims dims;
    dims.l_lon = rec[7]; dims.l_lat = rec[8]; dims.l_rad = rec[9];
    TessCentre centre;
    centre.lam = rec[10]; centre.cphi = rec[11]; centre.sphi = rec[12]; centre.rad = rec[13];
    const double distance = tess_distance(o, centre);
    int n_lon, n_lat, n_rad;
    if (!tess_split_counts(distance, dims, ratio, radial, n_lon, n_lat, n_rad)) {
        flags |= FLAG_ZERO_DIV;
        return -1;
    }
    if (n_lon * n_lat * n_rad > 1) return 0;
    TessNodes q;
    q.lam[0] = rec[14]; q.lam[1] = rec[15];
    q.cphi[0] = rec[16]; q.cphi[1] = rec[17];
    q.sphi[0] = rec[18]; q.sphi[1] = rec[19];
    q.rad[0] = rec[20]; q.rad[1] = rec[21];
    q.mass[0][0] = rec[22]; q.mass[0][1] = rec[23]; q.mass[1][0] = rec[24]; q.mass[1][1] = rec[25];
    acc += tess_glq_nodes<FIELD>(o, q, flags);
    return 1;
}

// ---- fast root record (kernel variant 2) ------------------------------------------------------------
// For the pairs whose ROOT does not split (the far field: almost all pairs) nothing but arithmetic
// is left: cos(lam_p - lam) = cos lam_p cos lam + sin lam_p sin lam with both factors precomputed
// (per tesseroid here, per observer in TessObs), the split test compares the squared distance with
// precomputed (ratio * size)^2, 1 / distance is the library's reciprocal square root (hb200_xmath),
// and G (and the sign) is folded into the node masses. These change roundings in the last place
// (the cosine of the difference carries ~1.5e-16 absolute error instead of cos()'s 0.6e-16, the
// same class as the reference's own cos(psi)), not the algorithm; pairs that split are walked with
// the exact statements as before. Record:
// [0..5] w e s n bottom top  [6] density  [7..9] (ratio l_lon)^2 (ratio l_lat)^2 (ratio l_rad)^2
// (negative where that direction never splits)  [10,11] centre cos / sin lam  [12,13] centre
// cphi sphi  [14] centre rad  [15,16] node cos lam  [17,18] node sin lam  [19,20] node cphi
// [21,22] node sphi  [23,24] node rad  [25..28] G * mass[j][k]  [29] 1 if a dimension is zero
// [30] density of the upper radial node
HB_HD void tess_pack_record_fast(double* rec, const double* tess, double density0, double density1,
                                 double ratio, bool radial)
{
#pragma unroll
    for (int c = 0; c < 6; c++) rec[c] = tess[c];
    rec[6] = density0;
    const double density[2] = {density0, density1};
    TessDims d;
    tess_dims(d, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5]);
    const double t_lon = ratio * d.l_lon, t_lat = ratio * d.l_lat, t_rad = ratio * d.l_rad;
    rec[7] = t_lon * t_lon;  // NaN sizes give NaN thresholds: "distance < NaN" is false, no split
    rec[8] = t_lat * t_lat;
    rec[9] = radial ? t_rad * t_rad : -1.0;
    rec[29] = (d.l_lon == 0.0 || d.l_lat == 0.0 || d.l_rad == 0.0) ? 1.0 : 0.0;
    TessCentre c;
    tess_centre(c, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5]);
    rec[10] = cos(c.lam); rec[11] = sin(c.lam);
    rec[12] = c.cphi; rec[13] = c.sphi; rec[14] = c.rad;
    TessNodes q;
    tess_nodes(q, tess[0], tess[1], tess[2], tess[3], tess[4], tess[5], density);
    rec[15] = cos(q.lam[0]); rec[16] = cos(q.lam[1]);
    rec[17] = sin(q.lam[0]); rec[18] = sin(q.lam[1]);
    rec[19] = q.cphi[0]; rec[20] = q.cphi[1];
    rec[21] = q.sphi[0]; rec[22] = q.sphi[1];
    rec[23] = q.rad[0]; rec[24] = q.rad[1];
    rec[25] = kG * q.mass[0][0]; rec[26] = kG * q.mass[0][1];
    rec[27] = kG * q.mass[1][0]; rec[28] = kG * q.mass[1][1];
    rec[kTessRho1Fast] = density1;
    rec[31] = 0.0;
}

// Same contract as tess_root, on a fast record.
template <int FIELD>
HB_HD int tess_root_fast(const TessObs& o, const double* rec, double& acc, unsigned& flags)
{
    if (rec[29] != 0.0) {
        flags |= FLAG_ZERO_DIV;
        return -1;
    }
    const double two_r = 2 * o.rad;
    {
        const double coslambda = fma(rec[10], o.clam, rec[11] * o.slam);
        const double cospsi = fma(rec[12] * o.cphi, coslambda, rec[13] * o.sphi);
        const double dr = o.rad - rec[14];
        const double d2 = fma(two_r * rec[14], 1 - cospsi, dr * dr);
        if (d2 < rec[7] || d2 < rec[8] || d2 < rec[9]) return 0;
    }
    double coslambda[2];
#pragma unroll
    for (int i = 0; i < 2; i++) coslambda[i] = fma(rec[15 + i], o.clam, rec[17 + i] * o.slam);
    double result = 0.0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double a = rec[21 + j] * o.sphi, b = rec[19 + j] * o.cphi;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double radius_p = rec[23 + k];
            const double dr = o.rad - radius_p;
            const double dr2 = dr * dr, rr = two_r * radius_p;
            const double mass = rec[25 + 2 * j + k];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double cospsi = fma(b, coslambda[i], a);
                const double d2 = fma(rr, 1 - cospsi, dr2);
                const double inv = fast_rsqrt(d2);  // d2 == 0: not finite, looked at below
                if (FIELD == F_POT) {
                    result = fma(mass, inv, result);
                } else {
                    const double delta_z = fma(-radius_p, cospsi, o.rad);
                    result = fma(-mass * delta_z, inv * inv * inv, result);
                }
            }
        }
    }
    acc += result;
    if (!(fabs(result) <= 1.7976931348623157e308)) {
        // Not finite. An observer ON a quadrature node (numba raises ZeroDivisionError on
        // 1 / distance) makes it so; NaN input does too: tell them apart here, off the hot path
        // (eight compares + selects per pair otherwise).
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const double radius_p = rec[23 + k];
                    const double dr = o.rad - radius_p;
                    const double cospsi = fma(rec[19 + j] * o.cphi, coslambda[i], rec[21 + j] * o.sphi);
                    if (fma(two_r * radius_p, 1 - cospsi, dr * dr) == 0.0) flags |= FLAG_ZERO_DIV;
                }
    }
    return 1;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------ kernels
// plain records (the inside scan and kernel variant 0)
// density0 / density1: per radial quadrature node (the same array twice for homogeneous bodies)
__global__ void pack_tesseroids_kernel(const double* __restrict__ tesseroids,
                                       const double* density0, const double* density1, int64_t n,
                                       double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kTessStride;
#pragma unroll
    for (int c = 0; c < 6; c++) q[c] = tesseroids[j * 6 + c];
    q[6] = density0[j];
    q[7] = density1[j];
}

// root records: everything about a tesseroid that does not depend on the observer
__global__ void pack_tesseroid_records_kernel(const double* __restrict__ tesseroids,
                                              const double* density0, const double* density1,
                                              int64_t n, double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double t[6];
#pragma unroll
    for (int c = 0; c < 6; c++) t[c] = tesseroids[j * 6 + c];
    tess_pack_record(packed + j * kTessRec, t, density0[j], density1[j]);
}

__global__ void pack_tesseroid_fast_records_kernel(const double* __restrict__ tesseroids,
                                                   const double* density0, const double* density1,
                                                   int64_t n, double ratio, int radial,
                                                   double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double t[6];
#pragma unroll
    for (int c = 0; c < 6; c++) t[c] = tesseroids[j * 6 + c];
    tess_pack_record_fast(packed + j * kTessRec, t, density0[j], density1[j], ratio, radial != 0);
}

struct TessArgs {
    const double* lon;
    const double* lat;
    const double* rad;
    int64_t n_obs;
    const double* packed;
    int64_t n_src;
    int64_t chunk_len;  // tesseroids per blockIdx.y
    double* out;        // gridDim.y == 1: final [n_obs]; else partial [y][n_obs]
    double scale;       // -1e5 for g_z (tesseroid_gravity.py:222-225), 1 for the potential
    double ratio;       // distance-size ratio, tesseroid_gravity.py:33
    int radial;
    unsigned* flags;
};

constexpr int kTessBlock = 64;  // observers per CTA: the per-thread stack lives in local memory
constexpr int kTessTile = 32;   // root records per shared-memory tile (8 KB)
// Split pairs a thread collects before the warp walks them. It has to exceed the number of
// tesseroids that split for ONE observer in typical models: with 16 every lane filled its list
// inside its own neighbourhood, alone, and the walks ran with 1-3 of 32 lanes (ncu, profiles/).
constexpr int kTessDefer = 64;

// Variant 0 (first build, kept selectable): plain records, every pair walked where it is met.
// One thread owns one observer, keeps its accumulator in a register and its discretisation
// stack (4.8 KB) in local memory; the CTA walks the tesseroid records in shared-memory tiles.
template <int FIELD>
__global__ void __launch_bounds__(kTessBlock) tesseroid_kernel(const TessArgs a)
{
    __shared__ double tile[kTessBlock * kTessStride];
    double stack[kTessStack * 6];
    const int64_t i = (int64_t)blockIdx.x * kTessBlock + threadIdx.x;
    const int64_t ic = i < a.n_obs ? i : a.n_obs - 1;
    TessObs o;
    tess_make_obs(o, a.lon[ic], a.lat[ic], a.rad[ic]);
    double acc = 0.0;
    unsigned flags = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    for (int64_t t0 = begin; t0 < end; t0 += kTessBlock) {
        const int cnt = (int)((end - t0) < kTessBlock ? (end - t0) : kTessBlock);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * kTessStride; x += kTessBlock)
            tile[x] = a.packed[t0 * kTessStride + x];
        __syncthreads();
        if (i < a.n_obs)
            for (int s = 0; s < cnt; s++) {
                const double* rec = tile + s * kTessStride;
                tess_pair<FIELD>(o, rec, rec[6], rec[7], a.ratio, a.radial != 0, stack, acc, flags);
            }
    }
    if (i < a.n_obs) {
        if (gridDim.y == 1) a.out[i] = acc * a.scale;
        else a.out[(int64_t)blockIdx.y * a.n_obs + i] = acc;
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// Walk the pairs a thread has deferred: every lane goes through ITS list with ITS stack, one pop
// per trip of a single loop, so the lanes of a warp work concurrently whatever the shapes of
// their discretisation trees.
template <int FIELD, bool FAST, class TRIG>
__device__ __forceinline__ void tess_walk_deferred(const TessObs& o, const TessArgs& a,
                                                   const int* defer, int& n_defer, int64_t begin,
                                                   double* stack, double& acc, unsigned& flags)
{
    TessWalk wk;
    wk.stack_top = -1;
    wk.n_leaves = 0;
    wk.density[0] = wk.density[1] = 0.0;
    int k = 0;
    while (true) {
        if (wk.stack_top < 0) {
            if (k >= n_defer) break;
            const double* rec = a.packed + (begin + defer[k++]) * kTessRec;
            tess_walk_begin(wk, rec, rec[6], rec[FAST ? kTessRho1Fast : kTessRho1], stack);
        }
        tess_walk_step<FIELD, kTessStack, kTessMaxLeaves, TRIG>(o, a.ratio, a.radial != 0, stack, wk, acc,
                                                                flags);
    }
    n_defer = 0;
}

// Variants 1, 2 (FAST) and 3 (FAST + the library's own trig in the walks): root records +
// deferred walks. The loop over a tile is uniform (root
// decision from the record, unsplit pairs integrated at once, three cosines per pair); a pair
// that splits is only noted. When any lane of the warp has kTessDefer pairs noted, and at the
// end, all lanes walk their lists together.
template <int FIELD, bool FAST, class TRIG = LibmTrig, int MINB = 1>
__global__ void __launch_bounds__(kTessBlock, MINB) tesseroid_deferred_kernel(const TessArgs a)
{
    __shared__ double tile[kTessTile * kTessRec];
    double stack[kTessStack * 6];
    int defer[kTessDefer];  // offsets from `begin`
    int n_defer = 0;
    const int64_t i = (int64_t)blockIdx.x * kTessBlock + threadIdx.x;
    const bool live = i < a.n_obs;
    const int64_t ic = live ? i : a.n_obs - 1;
    TessObs o;
    tess_make_obs(o, a.lon[ic], a.lat[ic], a.rad[ic]);
    double acc = 0.0;
    unsigned flags = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    for (int64_t t0 = begin; t0 < end; t0 += kTessTile) {
        const int cnt = (int)((end - t0) < kTessTile ? (end - t0) : kTessTile);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * kTessRec; x += kTessBlock)
            tile[x] = a.packed[t0 * kTessRec + x];
        __syncthreads();
        for (int s = 0; s < cnt; s++) {
            int root = 1;
            if (live)
                root = FAST ? tess_root_fast<FIELD>(o, tile + s * kTessRec, acc, flags)
                            : tess_root<FIELD>(o, tile + s * kTessRec, a.ratio, a.radial != 0, acc, flags);
            if (root == 0) defer[n_defer++] = (int)(t0 - begin) + s;
            if (__any_sync(0xffffffffu, n_defer == kTessDefer))
                tess_walk_deferred<FIELD, FAST, TRIG>(o, a, defer, n_defer, begin, stack, acc, flags);
        }
    }
    tess_walk_deferred<FIELD, FAST, TRIG>(o, a, defer, n_defer, begin, stack, acc, flags);
    if (live) {
        if (gridDim.y == 1) a.out[i] = acc * a.scale;
        else a.out[(int64_t)blockIdx.y * a.n_obs + i] = acc;
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// ---- variant 6: the root pass and the walks as TWO kernels ---------------------------------------
// tesseroid_root_kernel: the arithmetic-only far field of tess_root_fast over all pairs; a pair
// whose root splits is only RECORDED (its offset in the chunk) in a per-(chunk, observer) list in
// global memory. Without the walk (trig, divisions, the 4.8 KB stack) the kernel needs half the
// registers of tesseroid_deferred_kernel: 128-thread CTAs, twice the resident warps.
// tesseroid_walk_kernel: one thread per (chunk, observer) walks the recorded pairs of ITS list in
// order (lanes = neighbouring observers: similar lists, similar trees) and writes their sum.
// reduce_partials_kernel then adds, per observer and in fixed order, the root partials and the
// walk sums of all chunks: deterministic, no atomics on data.
// A list holds kTessListCap pairs. An observer with more split roots in one chunk (next to a
// pole every tesseroid of a latitude ring is near) stops its root pass at the first pair that
// does not fit and records that offset: the walk kernel, after the listed pairs, evaluates the
// REST of the chunk for this observer pair by pair with the general statements (root decision
// included). Nothing is skipped and nothing is counted twice.
constexpr int kTessRootBlock = 128;
constexpr int kTessListCap = 64;
// Chunks of the tesseroid list are SHORT (128 tesseroids while the list has <= 32768 of them):
// the walks of one observer are then spread over n / 128 threads of the walk kernel, which is
// what bounds its tail (an observer next to a pole has hundreds of near tesseroids, all others
// ~20; with 32 long chunks the walk kernel kept the SMs busy 40 % of its run time, profiles/).
// Offsets inside a chunk fit 16 bits.
constexpr int kTessChunk = 128;
constexpr int kTessMaxChunks = 256;
constexpr int kTessObsBatch = 32768;  // observers per pass: bounds the list workspace (~0.6 GB)

// TMA bulk copy + mbarrier helpers (defined in hb200_kernels.cuh, included first by hb200_api.cu)
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count);
__device__ __forceinline__ void mbar_fence_init();
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar);
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity);

constexpr int kTessWarpTile = 16;  // root records per bulk copy (4 KB), two buffers per warp
constexpr int kTessWalkSlices = 4; // threads that share the walks of one (chunk, observer) list

// Every warp streams its own copy of the chunk's root records (TMA bulk copies, two buffers, one
// mbarrier each) and synchronises only with itself, like prism_kernel.
template <int FIELD, int MINB>
__global__ void __launch_bounds__(kTessRootBlock, MINB) tesseroid_root_kernel(const TessArgs a,
                                                                               unsigned short* list, int* count,
                                                                               int* items = nullptr,
                                                                               int* n_items = nullptr)
{
    constexpr int WARPS = kTessRootBlock / 32;
    __shared__ alignas(128) double tiles[WARPS][2][kTessWarpTile * kTessRec];
    __shared__ alignas(8) uint64_t bars[WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * kTessRootBlock + threadIdx.x;
    const bool live = i < a.n_obs;
    const int64_t ic = live ? i : a.n_obs - 1;
    TessObs o;
    tess_make_obs(o, a.lon[ic], a.lat[ic], a.rad[ic]);
    double acc = 0.0;
    unsigned flags = 0;
    int n_split = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    int resume = (int)(end - begin);  // offset from which the walk kernel takes over (none)
    bool open = live;
    unsigned short* my_list = list + ((int64_t)blockIdx.y * kTessListCap) * a.n_obs + ic;  // [chunk][k][obs]
    if (lane == 0) {
        mbar_init(&bars[warp][0], 1);
        mbar_init(&bars[warp][1], 1);
        mbar_fence_init();
    }
    __syncwarp();
    if (lane == 0 && begin < end) {
        const int cnt0 = (int)((end - begin) < kTessWarpTile ? (end - begin) : kTessWarpTile);
        bulk_load(tiles[warp][0], a.packed + begin * kTessRec, (unsigned)(cnt0 * kTessRec * sizeof(double)),
                  &bars[warp][0]);
    }
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    for (int64_t t0 = begin; t0 < end; t0 += kTessWarpTile, buf ^= 1) {
        const int cnt = (int)((end - t0) < kTessWarpTile ? (end - t0) : kTessWarpTile);
        const int64_t t1 = t0 + kTessWarpTile;
        if (lane == 0 && t1 < end) {
            const int cnt1 = (int)((end - t1) < kTessWarpTile ? (end - t1) : kTessWarpTile);
            bulk_load(tiles[warp][buf ^ 1], a.packed + t1 * kTessRec,
                      (unsigned)(cnt1 * kTessRec * sizeof(double)), &bars[warp][buf ^ 1]);
        }
        if (buf == 0) { mbar_wait(&bars[warp][0], phase0); phase0 ^= 1; }
        else { mbar_wait(&bars[warp][1], phase1); phase1 ^= 1; }
        const double* tile = tiles[warp][buf];
        for (int s = 0; s < cnt; s++) {
            if (!open) continue;
            const int root = tess_root_fast<FIELD>(o, tile + s * kTessRec, acc, flags);
            if (root == 0) {
                if (n_split < kTessListCap) {
                    my_list[(int64_t)n_split * a.n_obs] = (unsigned short)((int)(t0 - begin) + s);
                    n_split++;
                } else {  // list full: this pair and the rest of the chunk go to the walk kernel
                    resume = (int)(t0 - begin) + s;
                    open = false;
                }
            }
        }
        __syncwarp();  // the buffer just read may be refilled in the next iteration
    }
    if (live) {
        a.out[(int64_t)blockIdx.y * a.n_obs + i] = acc;  // always partial: [chunk][obs]
        count[(int64_t)blockIdx.y * a.n_obs + i] = n_split;
        count[((int64_t)gridDim.y + blockIdx.y) * a.n_obs + i] = resume;
    }
    if (items) {
        // the (chunk, observer) lists that hold work, compacted for tesseroid_coop_walk_kernel:
        // one atomic per warp; the order of the items has no influence on any result
        const bool has = live && (n_split > 0 || resume < (int)(end - begin));
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (m) {
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(n_items, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (has) items[base + __popc(m & ((1u << lane) - 1u))] = (int)blockIdx.y * kTessObsBatch + (int)i;
        }
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// The listed pairs k, k + stride, ... of one (chunk, observer) list and then the pairs resume,
// resume + stride, ... of the rest of the chunk, each walked depth-first with the caller's private
// stack exactly like the reference does (one pop per trip of a single loop).
template <int FIELD, class TRIG>
__device__ __forceinline__ void tess_walk_list_exact(const TessArgs& a, const TessObs& o,
                                                     const unsigned short* my_list, int n, int k, int resume,
                                                     int chunk_cnt, int64_t begin, int stride, double* stack,
                                                     double& acc, unsigned& flags)
{
    TessWalk wk;
    wk.stack_top = -1;
    wk.n_leaves = 0;
    wk.density[0] = wk.density[1] = 0.0;
    while (true) {  // one pop per trip, whatever the shapes of the lanes' trees
        if (wk.stack_top < 0) {
            int off;
            if (k < n) {  // the listed pairs of this slice ...
                off = my_list[(int64_t)k * a.n_obs];
                k += stride;
            } else if (resume < chunk_cnt) {  // ... then its share of the rest of the chunk
                off = resume;
                resume += stride;
            } else {
                break;
            }
            const double* rec = a.packed + (begin + off) * kTessRec;
            if (rec[29] != 0.0) {  // a zero dimension: numba's ZeroDivisionError (tess_root_fast)
                flags |= FLAG_ZERO_DIV;
                continue;
            }
            tess_walk_begin(wk, rec, rec[6], rec[kTessRho1Fast], stack);
        }
        tess_walk_step<FIELD, kTessStack, kTessMaxLeaves, TRIG>(o, a.ratio, a.radial != 0, stack, wk, acc,
                                                                flags);
    }
}

// grid = (observer blocks, chunks, kTessWalkSlices): slice z of a (chunk, observer) list takes
// every kTessWalkSlices-th listed pair and every kTessWalkSlices-th pair of the remainder, so
// the few heavy observers (next to a pole) are shared by more threads. walk_sum is
// [chunk][slice][obs]; the final reduce adds the slices in fixed order.
template <int FIELD, class TRIG>
__global__ void __launch_bounds__(kTessBlock) tesseroid_walk_kernel(const TessArgs a,
                                                                    const unsigned short* list,
                                                                    const int* count, double* walk_sum)
{
    double stack[kTessStack * 6];
    const int64_t i = (int64_t)blockIdx.x * kTessBlock + threadIdx.x;
    if (i >= a.n_obs) return;
    const int n = count[(int64_t)blockIdx.y * a.n_obs + i];
    int resume = count[((int64_t)gridDim.y + blockIdx.y) * a.n_obs + i] + (int)blockIdx.z;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int chunk_cnt = (int)((begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src) - begin);
    int k = (int)blockIdx.z;
    double acc = 0.0;
    if (k < n || resume < chunk_cnt) {
        TessObs o;
        tess_make_obs(o, a.lon[i], a.lat[i], a.rad[i]);
        unsigned flags = 0;
        const unsigned short* my_list = list + ((int64_t)blockIdx.y * kTessListCap) * a.n_obs + i;
        tess_walk_list_exact<FIELD, TRIG>(a, o, my_list, n, k, resume, chunk_cnt, begin, kTessWalkSlices,
                                          stack, acc, flags);
        if (flags && a.flags) atomicOr(a.flags, flags);
    }
    walk_sum[((int64_t)blockIdx.y * kTessWalkSlices + blockIdx.z) * a.n_obs + i] = acc;
}

// ---- cooperative walks (kernel variant 9) --------------------------------------------------------
// tesseroid_walk_kernel gives every (chunk, observer) list to one thread: the 32 lists of a warp
// have very different lengths and tree shapes, and ncu counted 8 of 32 active lanes on its
// instructions (profiles/r2_ncu_tesseroid_gz_two_kernel.txt). Here a GROUP of kCoopG = 16 lanes
// walks one list together: the group keeps a stack of nodes (bounds + the pair they belong to)
// in shared memory, every trip each lane pops one node, decides on it with the reference's
// statements, and either pushes its children back (positions from a prefix sum over the group) or
// integrates the leaf into its own accumulator; the lanes' accumulators are added in fixed order
// when the list is done. The lists that hold work are compacted by the root kernel into a work
// list; the kernel is persistent (one CTA per resident slot) and every group draws its next list
// from a global cursor, so the few long lists (observers next to a pole) do not leave the rest of
// the machine idle. The two groups of a warp run the same trip loop in lockstep. What changes
// against the reference is only the ORDER in which the leaves of a list are added; that order
// depends on nothing but the list itself (bit-reproducible under any batching of the observers).
//
// Where the reference raises (OverflowError): its depth-first stack of kTessStack nodes cannot
// overflow before depth (kTessStack - 8) / 7 = 13, and no pair can exceed kTessMaxLeaves leaves
// unless the whole list does. A list that reaches either bound (or that would overflow the
// group's stack) is thrown away and noted; tesseroid_redo_kernel walks it with ONE thread and the
// exact depth-first walk of tesseroid_walk_kernel, which reports the reference's errors.
constexpr int kCoopG = 16;       // lanes per group (measured: 4 lanes 2.4 ms, 8: 1.46, 16: 1.39, 32: 1.73 ms per batch)
constexpr int kCoopCap = 160;    // nodes per group stack (107 with the 3-D discretisation: 6 bounds per node)
constexpr int kCoopCapRadial = 107;
constexpr int kCoopDeep = 13;    // a node at this depth that wants to split sends the list to the exact walk
constexpr int kCoopBlock = 128;  // 4 warps = 8 groups
constexpr int kCoopCtasPerSm = 4;

// One row per bound (lanes read neighbouring words): w e s n, and bottom top only when the radial
// direction is discretised too (otherwise they follow from the root's, read from the record).
struct CoopStack {
    double b[6 * kCoopCapRadial];  // >= 4 * kCoopCap
    int tag[kCoopCap];             // offset of the pair's root record in the chunk | depth << 16; -1: void
};
static_assert(6 * kCoopCapRadial >= 4 * kCoopCap, "CoopStack rows");

// counters[0]: number of items; [1]: cursor of the walk kernel; [2]: number of redo items
template <int FIELD, class TRIG>
__global__ void __launch_bounds__(kCoopBlock, kCoopCtasPerSm) tesseroid_coop_walk_kernel(
    const TessArgs a, int n_chunks, const unsigned short* list, const int* count, const int* items,
    int* counters, int* redo_items, double* walk_sum)
{
    __shared__ CoopStack stacks[kCoopBlock / kCoopG];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int g = lane / kCoopG, l = lane % kCoopG;
    CoopStack& S = stacks[threadIdx.x / kCoopG];
    const int n_items = counters[0];
    const bool radial = a.radial != 0;
    const int spread = radial ? 7 : 3;  // net growth of the stack per node that splits, at most
    const int cap = radial ? kCoopCapRadial : kCoopCap;
    // Sixteen lanes popping the top of the stack need sixteen times the memory of a depth-first
    // walk (48 nodes per level). While the stack is nearly full only ONE lane pops -- a plain
    // depth-first descent, 3 nodes per level, for which `narrow` slots are kept free -- until
    // the top of the stack has drained.
    const int narrow = radial ? 14 : 18;

    // group state (the same in all lanes of a group, except acc / leaves / flags)
    int phase = 0;  // 0: draw the next list; 1: walking; 2: the work list is exhausted
    int item = 0, n = 0, k = 0, resume = 0, cnt = 0, leaves = 0, chunk_cnt = 0;
    int64_t begin = 0, obs = 0;
    const unsigned short* my_list = list;
    TessObs o;
    tess_make_obs(o, 0.0, 0.0, 1.0);
    double acc = 0.0;
    unsigned flags = 0;

    while (true) {
        {
            int drawn = 0;
            if (phase == 0 && l == 0) drawn = atomicAdd(&counters[1], 1);
            drawn = __shfl_sync(FULL, drawn, 0, kCoopG);
            if (phase == 0) {
                if (drawn < n_items) {
                    item = items[drawn];
                    const int chunk = item / kTessObsBatch;
                    obs = item % kTessObsBatch;
                    begin = (int64_t)chunk * a.chunk_len;
                    chunk_cnt = (int)((begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src) - begin);
                    n = count[(int64_t)chunk * a.n_obs + obs];
                    resume = count[((int64_t)n_chunks + chunk) * a.n_obs + obs];
                    tess_make_obs(o, a.lon[obs], a.lat[obs], a.rad[obs]);
                    my_list = list + ((int64_t)chunk * kTessListCap) * a.n_obs + obs;
                    k = 0; cnt = 0; leaves = 0; acc = 0.0;
                    phase = 1;
                } else {
                    phase = 2;
                }
            }
        }
        if (!__any_sync(FULL, phase == 1)) break;

        // feed: the next pairs of the list (then of the rest of the chunk) become root nodes
        if (phase == 1 && cnt < kCoopG) {
            const int rest = chunk_cnt - resume;
            const int roots = (n - k) + (rest > 0 ? rest : 0);
            const int m = roots < kCoopG ? roots : kCoopG;
            const int from_list = (n - k) < m ? (n - k) : m;
            if (l < m) {
                const int off = l < from_list ? (int)my_list[(int64_t)(k + l) * a.n_obs] : resume + (l - from_list);
                const double* rec = a.packed + (begin + off) * kTessRec;
                int tag = off;
                if (rec[29] != 0.0) {  // a zero dimension: numba's ZeroDivisionError (tess_root_fast)
                    flags |= FLAG_ZERO_DIV;
                    tag = -1;
                }
#pragma unroll
                for (int c = 0; c < 4; c++) S.b[c * cap + cnt + l] = rec[c];
                if (radial) {
                    S.b[4 * cap + cnt + l] = rec[4];
                    S.b[5 * cap + cnt + l] = rec[5];
                }
                S.tag[cnt + l] = tag;
            }
            k += from_list;
            resume += m - from_list;
            cnt += m;
        }
        __syncwarp();

        // pop: as many nodes as there are lanes, fewer while the stack is nearly full
        int take = 0;
        bool stuck = false;
        if (phase == 1) {
            take = cnt < kCoopG ? cnt : kCoopG;
            int wide = (cap - narrow - cnt) / spread;
            if (wide < 1) wide = 1;
            if (take > wide) take = wide;
            if (cnt + spread > cap) take = 0;  // even one pop could overflow: give the list up
            stuck = take == 0 && cnt > 0;
        }
        double w = 0, e = 0, s = 0, nn = 0, bottom = 0, top = 0;
        int tag = -1;
        if (l < take) {
            const int at = cnt - 1 - l;
            w = S.b[at]; e = S.b[cap + at]; s = S.b[2 * cap + at]; nn = S.b[3 * cap + at];
            tag = S.tag[at];
            if (radial) {
                bottom = S.b[4 * cap + at];
                top = S.b[5 * cap + at];
            } else if (tag >= 0) {
                // _split_tesseroid with n_rad = 1 hands every child bottom + 0 (exact) and
                // bottom + (top - bottom) / 1, which need not be the parent's top bit for bit:
                // the node's top is the root's after `depth` such steps (a fixed point after one
                // or two)
                const double* rec = a.packed + (begin + (tag & 0xffff)) * kTessRec;
                bottom = rec[4];
                top = rec[5];
                for (int d = tag >> 16; d > 0; d--) {
                    const double next = bottom + (top - bottom);
                    if (next == top) break;
                    top = next;
                }
            }
        }
        cnt -= take;
        __syncwarp();  // every pop has been read before a child is pushed over it

        int n_lon = 1, n_lat = 1, n_rad = 1, kids = 0;
        bool leaf = false;
        if (tag >= 0) {
            if (!tess_classify<TRIG>(o, a.ratio, radial, w, e, s, nn, bottom, top, n_lon, n_lat, n_rad)) {
                flags |= FLAG_ZERO_DIV;
            } else {
                kids = n_lon * n_lat * n_rad;
                leaf = kids == 1;
                if (leaf) kids = 0;
            }
        }
        const bool too_deep = kids > 0 && (tag >> 16) + 1 >= kCoopDeep;
        // where the children go: prefix sum over the lanes of the group
        int incl = kids;
#pragma unroll
        for (int d = 1; d < kCoopG; d <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, d, kCoopG);
            if (l >= d) incl += v;
        }
        const int pushed = __shfl_sync(FULL, incl, kCoopG - 1, kCoopG);
        const unsigned bad_lanes = __ballot_sync(FULL, too_deep || stuck);
        constexpr unsigned group_mask = kCoopG == 32 ? 0xffffffffu : ((1u << (kCoopG % 32)) - 1u);
        const bool bad = ((bad_lanes >> ((kCoopG * g) % 32)) & group_mask) != 0u;
        if (kids > 0 && !bad) {  // _split_tesseroid
            int at = cnt + incl - kids;
            const double d_lon = (e - w) / n_lon, d_lat = (nn - s) / n_lat, d_rad = (top - bottom) / n_rad;
            const int child_tag = tag + (1 << 16);
            for (int i = 0; i < n_lon; i++)
                for (int j = 0; j < n_lat; j++)
                    for (int r = 0; r < n_rad; r++, at++) {
                        S.b[at] = w + d_lon * i;
                        S.b[cap + at] = w + d_lon * (i + 1);
                        S.b[2 * cap + at] = s + d_lat * j;
                        S.b[3 * cap + at] = s + d_lat * (j + 1);
                        if (radial) {
                            S.b[4 * cap + at] = bottom + d_rad * r;
                            S.b[5 * cap + at] = bottom + d_rad * (r + 1);
                        }
                        S.tag[at] = child_tag;
                    }
        }
        cnt += pushed;
        if (leaf) {
            const double* rec = a.packed + (begin + (tag & 0xffff)) * kTessRec;
            const double density[2] = {rec[6], rec[kTessRho1Fast]};
            TessNodes nodes;
            tess_nodes<TRIG>(nodes, w, e, s, nn, bottom, top, density);
            acc += tess_glq_nodes<FIELD, TRIG>(o, nodes, flags);
            leaves += 1;
        }

        // a list is done (or given up): add the lanes' sums in fixed order
        const bool done = phase == 1 && (bad || (cnt == 0 && k >= n && resume >= chunk_cnt));
        if (__any_sync(FULL, done)) {
            double total = acc;
            int total_leaves = leaves;
#pragma unroll
            for (int d = 1; d < kCoopG; d <<= 1) {
                total += __shfl_xor_sync(FULL, total, d, kCoopG);
                total_leaves += __shfl_xor_sync(FULL, total_leaves, d, kCoopG);
            }
            if (done) {
                if (l == 0) {
                    if (bad || total_leaves > kTessMaxLeaves) redo_items[atomicAdd(&counters[2], 1)] = item;
                    else walk_sum[(int64_t)(item / kTessObsBatch) * a.n_obs + obs] = total;
                }
                phase = 0;
            }
        }
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// the lists the cooperative walk gave up (rare): one thread each, the exact depth-first walk
template <int FIELD, class TRIG>
__global__ void __launch_bounds__(kTessBlock) tesseroid_redo_kernel(const TessArgs a, int n_chunks,
                                                                    const unsigned short* list, const int* count,
                                                                    const int* counters, const int* redo_items,
                                                                    double* walk_sum)
{
    const int n_redo = counters[2];
    double stack[kTessStack * 6];
    for (int r = blockIdx.x * kTessBlock + threadIdx.x; r < n_redo; r += gridDim.x * kTessBlock) {
        const int item = redo_items[r];
        const int chunk = item / kTessObsBatch;
        const int64_t obs = item % kTessObsBatch;
        const int64_t begin = (int64_t)chunk * a.chunk_len;
        const int chunk_cnt = (int)((begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src) - begin);
        TessObs o;
        tess_make_obs(o, a.lon[obs], a.lat[obs], a.rad[obs]);
        double acc = 0.0;
        unsigned flags = 0;
        tess_walk_list_exact<FIELD, TRIG>(a, o, list + ((int64_t)chunk * kTessListCap) * a.n_obs + obs,
                                          count[(int64_t)chunk * a.n_obs + obs], 0,
                                          count[((int64_t)n_chunks + chunk) * a.n_obs + obs], chunk_cnt, begin, 1,
                                          stack, acc, flags);
        walk_sum[(int64_t)chunk * a.n_obs + obs] = acc;
        if (flags && a.flags) atomicOr(a.flags, flags);
    }
}

// check_points_outside_tesseroids as one pass: sets FLAG_TESS_INSIDE if any pair conflicts
__global__ void __launch_bounds__(128) tesseroid_inside_scan_kernel(const TessArgs a)
{
    __shared__ double tile[128 * kTessStride];
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t ic = i < a.n_obs ? i : a.n_obs - 1;
    const double lon = a.lon[ic], lat = a.lat[ic], rad = a.rad[ic];
    bool hit = false;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    for (int64_t t0 = begin; t0 < end; t0 += 128) {
        const int cnt = (int)((end - t0) < 128 ? (end - t0) : 128);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * kTessStride; x += 128)
            tile[x] = a.packed[t0 * kTessStride + x];
        __syncthreads();
        for (int s = 0; s < cnt; s++) hit |= tess_point_inside(lon, lat, rad, tile + s * kTessStride);
    }
    if (hit && i < a.n_obs) atomicOr(a.flags, FLAG_TESS_INSIDE);
}
#endif  // __CUDACC__

}  // namespace hb
