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
// restructured is decision-neutral and value-identical: the leaves are integrated as they are
// popped instead of being collected first (same order of additions), the observer's radians /
// cos / sin are hoisted out of the pair loop, and the two distinct cos(longitude_p - longitude) of
// a leaf are computed once instead of for all eight nodes.
#pragma once
#include "hb200_math.cuh"

namespace hb {

constexpr int kTessStack = 100;          // tesseroid_gravity.py:30  STACK_SIZE
constexpr int kTessMaxLeaves = 100000;   // tesseroid_gravity.py:31  MAX_DISCRETIZATIONS
constexpr int kTessStride = 8;           // w e s n bottom top density -
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
}

// _tesseroid_utils.py:19-107 with point.py:324-354. FIELD: F_POT or F_U (radial component).
template <int FIELD>
HB_HD double tess_glq(const TessObs& o, double w, double e, double s, double n, double bottom,
                      double top, double density, unsigned& flags)
{
    const double a_factor = 1.0 / 8 * ((e - w) * kDeg2Rad) * ((n - s) * kDeg2Rad) * (top - bottom);
    double coslambda[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const double node = i ? kGlqNode : -kGlqNode;
        const double longitude_p = (0.5 * (e - w) * node + 0.5 * (e + w)) * kDeg2Rad;
        coslambda[i] = cos(longitude_p - o.lam);
    }
    double result = 0.0;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const double latitude_p = (0.5 * (n - s) * (j ? kGlqNode : -kGlqNode) + 0.5 * (n + s)) * kDeg2Rad;
        const double cosphi_p = cos(latitude_p), sinphi_p = sin(latitude_p);
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const double radius_p = 0.5 * (top - bottom) * (k ? kGlqNode : -kGlqNode) + 0.5 * (top + bottom);
            const double kappa = radius_p * radius_p * cosphi_p;
            const double mass = density * a_factor * kappa;  // the three GLQ weights are 1
            const double dr = o.rad - radius_p;
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const double cospsi = sinphi_p * o.sphi + cosphi_p * o.cphi * coslambda[i];
                const double dist = sqrt(dr * dr + 2 * o.rad * radius_p * (1 - cospsi));
                double kern;
                if (dist == 0.0) flags |= FLAG_ZERO_DIV;  // observer on a quadrature node
                if (FIELD == F_POT) {
                    kern = 1 / dist * kG;
                } else {
                    const double delta_z = o.rad - radius_p * cospsi;
                    kern = -kG * delta_z / (dist * dist * dist);
                }
                result += mass * kern;
            }
        }
    }
    return result;
}

// One (observer, tesseroid) pair: adds the quadrature of every leaf of the adaptive
// discretisation to `acc` in the reference's order. `stack` holds STACK x 6 doubles and is
// private to the caller. Returns the number of leaves. (STACK / MAX_LEAVES are template
// parameters only so that the tests can provoke the overflow errors like the reference's do.)
template <int FIELD, int STACK = kTessStack, int MAX_LEAVES = kTessMaxLeaves>
HB_HD int tess_pair(const TessObs& o, const double* tess, double density, double ratio, bool radial,
                    double* stack, double& acc, unsigned& flags)
{
#pragma unroll
    for (int c = 0; c < 6; c++) stack[c] = tess[c];
    int stack_top = 0;
    int n_leaves = 0;
    while (stack_top >= 0) {
        const double* q = stack + 6 * stack_top;
        const double w = q[0], e = q[1], s = q[2], n = q[3], bottom = q[4], top = q[5];
        stack_top -= 1;
        // _tesseroid_dimensions
        const double wr = w * kDeg2Rad, er = e * kDeg2Rad, sr = s * kDeg2Rad, nr = n * kDeg2Rad;
        const double latitude_center = (nr + sr) / 2;
        const double l_lat = top * acos(sin(nr) * sin(sr) + cos(nr) * cos(sr));
        const double sc = sin(latitude_center), cc = cos(latitude_center);
        const double l_lon = top * acos(sc * sc + cc * cc * cos(er - wr));
        const double l_rad = top - bottom;
        // _distance_tesseroid_point -> distance_spherical (degrees in, centre of the tesseroid)
        const double longitude_p = ((w + e) / 2) * kDeg2Rad;
        const double latitude_p = ((s + n) / 2) * kDeg2Rad;
        const double radius_p = (bottom + top) / 2;
        const double cosphi_p = cos(latitude_p), sinphi_p = sin(latitude_p);
        const double coslambda = cos(longitude_p - o.lam);
        const double cospsi = sinphi_p * o.sphi + cosphi_p * o.cphi * coslambda;
        const double dr = o.rad - radius_p;
        const double distance = sqrt(dr * dr + 2 * o.rad * radius_p * (1 - cospsi));
        // numba evaluates all three quotients and raises ZeroDivisionError on a zero divisor
        // (a child so small that acos(...) == 0: the observer sits on a corner that every level
        // of the 3-D discretisation keeps splitting)
        if (l_lon == 0.0 || l_lat == 0.0 || l_rad == 0.0) {
            flags |= FLAG_ZERO_DIV;
            return n_leaves;
        }
        const int n_lon = (distance / l_lon < ratio) ? 2 : 1;
        const int n_lat = (distance / l_lat < ratio) ? 2 : 1;
        const int n_rad = (distance / l_rad < ratio && radial) ? 2 : 1;
        if (n_lon * n_lat * n_rad > 1) {
            if ((stack_top + 1) + n_lon * n_lat * n_rad > STACK) {
                flags |= FLAG_TESS_STACK;
                return n_leaves;
            }
            // _split_tesseroid
            const double d_lon = (e - w) / n_lon, d_lat = (n - s) / n_lat, d_rad = (top - bottom) / n_rad;
            for (int i = 0; i < n_lon; i++)
                for (int j = 0; j < n_lat; j++)
                    for (int k = 0; k < n_rad; k++) {
                        stack_top += 1;
                        double* c = stack + 6 * stack_top;
                        c[0] = w + d_lon * i;
                        c[1] = w + d_lon * (i + 1);
                        c[2] = s + d_lat * j;
                        c[3] = s + d_lat * (j + 1);
                        c[4] = bottom + d_rad * k;
                        c[5] = bottom + d_rad * (k + 1);
                    }
        } else {
            if (n_leaves + 1 > MAX_LEAVES) {
                flags |= FLAG_TESS_LEAVES;
                return n_leaves;
            }
            acc += tess_glq<FIELD>(o, w, e, s, n, bottom, top, density, flags);
            n_leaves += 1;
        }
    }
    return n_leaves;
}

#if defined(__CUDACC__)
// ------------------------------------------------------------------ kernels
__global__ void pack_tesseroids_kernel(const double* __restrict__ tesseroids,
                                       const double* __restrict__ density, int64_t n,
                                       double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kTessStride;
#pragma unroll
    for (int c = 0; c < 6; c++) q[c] = tesseroids[j * 6 + c];
    q[6] = density[j];
    q[7] = 0.0;
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

// One thread owns one observer, keeps its accumulator in a register and its discretisation
// stack (4.8 KB) in local memory; the CTA walks the tesseroid records in shared-memory tiles.
// Lanes diverge only inside tess_pair (near pairs split, far pairs do not).
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
                tess_pair<FIELD>(o, rec, rec[6], a.ratio, a.radial != 0, stack, acc, flags);
            }
    }
    if (i < a.n_obs) {
        if (gridDim.y == 1) a.out[i] = acc * a.scale;
        else a.out[(int64_t)blockIdx.y * a.n_obs + i] = acc;
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
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
