// hb200_kernels.cuh -- CUDA kernels (sm_100a) of the pairwise forward models.
//
// Work decomposition (all kernels): one thread owns one observation point and
// keeps its float64 accumulators in registers; a CTA of BLOCK observers walks
// the source list in tiles staged in shared memory as packed records; every
// lane of a warp reads the SAME record (one broadcast LDS.128 per double2), so
// the only per-pair memory traffic is shared-memory broadcast. grid.y splits
// the source list into chunks when there are too few observer CTAs to fill
// the 148 SMs; chunk partials are combined in a fixed order (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hb200_math.cuh"
#include "hb200_fast.cuh"

namespace hb {

constexpr int kBlock = 128;  // observers per CTA
constexpr int kTile = 128;   // sources per shared-memory tile

constexpr int kPrismStride = 8;   // w e s n b t G*rho skip
constexpr int kMagStride = 10;    // w e s n b t me mn mu skip
constexpr int kPointStride = 4;   // e n u weight
constexpr int kSphStride = 6;     // cos(lon) sin(lon) cos(lat) sin(lat) radius weight

struct Scales {
    double s[6];
};

// ------------------------------------------------------------ pack kernels
// prisms (P,6) AoS + density -> packed records. gravity.py:526-537 reads the
// same six columns per pair; here they are read once.
__global__ void pack_prisms_kernel(const double* __restrict__ prisms,
                                   const double* __restrict__ density, int64_t n,
                                   double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kPrismStride;
#pragma unroll
    for (int c = 0; c < 6; c++) q[c] = prisms[j * 6 + c];
    q[6] = density ? kG * density[j] : 0.0;
    q[7] = 0.0;
}

__global__ void pack_mag_kernel(const double* __restrict__ prisms, const double* __restrict__ me,
                                const double* __restrict__ mn, const double* __restrict__ mu,
                                int64_t n, double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kMagStride;
#pragma unroll
    for (int c = 0; c < 6; c++) q[c] = prisms[j * 6 + c];
    q[6] = me[j];
    q[7] = mn[j];
    q[8] = mu[j];
    q[9] = 0.0;
}

// layer.py:584-610: bounds from 1-D centres, skip rules in the reference's
// order, records emitted easting-outer / northing-inner (record index
// = j * n_north + k for easting index j, northing index k).
__global__ void pack_layer_kernel(const double* __restrict__ east_c,
                                  const double* __restrict__ north_c, int64_t n_east,
                                  int64_t n_north, const double* __restrict__ bottom,
                                  const double* __restrict__ top,
                                  const double* __restrict__ density, double thickness_threshold,
                                  double* __restrict__ packed)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_east * n_north) return;
    const int64_t j = idx / n_north, k = idx % n_north;
    const double half_e = (east_c[1] - east_c[0]) / 2;
    const double half_n = (north_c[1] - north_c[0]) / 2;
    const double rho = density[k * n_east + j];
    const double b = bottom[k * n_east + j], t = top[k * n_east + j];
    bool skip = (rho == 0.0) || isnan(rho);
    skip = skip || (t - b < thickness_threshold);
    skip = skip || isnan(t) || isnan(b);
    double* q = packed + idx * kPrismStride;
    q[0] = east_c[j] - half_e;
    q[1] = east_c[j] + half_e;
    q[2] = north_c[k] - half_n;
    q[3] = north_c[k] + half_n;
    q[4] = b;
    q[5] = t;
    q[6] = kG * rho;
    // 2: evaluated, in the same column as the previous record, which is evaluated too, with the
    // same bottom, and its south bound IS that record's north bound (bit for bit):
    // prism_kernel<.., REUSE> then keeps two vertex distances (LayerCarry, hb200_fast.cuh)
    double mark = skip ? 1.0 : 0.0;
    if (!skip && k > 0) {
        const double rho_p = density[(k - 1) * n_east + j];
        const double b_p = bottom[(k - 1) * n_east + j], t_p = top[(k - 1) * n_east + j];
        bool skip_p = (rho_p == 0.0) || isnan(rho_p);
        skip_p = skip_p || (t_p - b_p < thickness_threshold);
        skip_p = skip_p || isnan(t_p) || isnan(b_p);
        if (!skip_p && b_p == b && north_c[k - 1] + half_n == q[2]) mark = 2.0;
    }
    q[7] = mark;
}

// point sources: weight = G*mass (choclo.point: G * mass * kernel) or the EQS
// coefficient as is (utils.py:90: coeffs[j] * greens_function).
__global__ void pack_points_kernel(const double* __restrict__ pe, const double* __restrict__ pn,
                                   const double* __restrict__ pu, const double* __restrict__ w,
                                   int64_t n, int scale_by_G, double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kPointStride;
    q[0] = pe[j];
    q[1] = pn[j];
    q[2] = pu[j];
    q[3] = scale_by_G ? kG * w[j] : w[j];
}

// point.py:426-435: radians / cos / sin of the sources computed once.
__global__ void pack_points_sph_kernel(const double* __restrict__ lon,
                                       const double* __restrict__ lat,
                                       const double* __restrict__ rad,
                                       const double* __restrict__ mass, int64_t n,
                                       double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double d2r = kPi / 180.0;
    const double phi = lat[j] * d2r, lam = lon[j] * d2r;
    double* q = packed + j * kSphStride;
    q[0] = cos(lam);
    q[1] = sin(lam);
    q[2] = cos(phi);
    q[3] = sin(phi);
    q[4] = rad[j];
    q[5] = mass[j];
}

// ------------------------------------------------------------ prism kernel
struct PrismArgs {
    const double* oe;
    const double* on;
    const double* ou;
    int64_t n_obs;
    const double* packed;
    int64_t n_src;
    int64_t chunk_len;  // sources per blockIdx.y
    double* out;        // gridDim.y == 1: final [nout][n_obs]; else partial [y][nout][n_obs]
    Scales sc;
    unsigned rules;
    unsigned* flags;
};

// ---- TMA bulk-copy staging (cp.async.bulk + mbarrier; SASS: UBLKCP / SYNCS) --------------
// One elected lane per warp asks the copy engine for the next tile of packed records while the
// warp computes on the current one (two shared-memory buffers, one mbarrier each); the other
// lanes never touch global memory inside the source loop.
__device__ __forceinline__ unsigned smem_addr(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}

// Every WARP stages its own copy of the record stream (kWarpTile records per bulk copy, two
// buffers, one mbarrier each) and synchronises only with itself: there is no CTA-wide barrier in
// the source loop, so a warp whose lanes meet near-field prisms (longer sequences) never holds
// up the other three. The copies come out of L2 (the packed records of a call are read
// n_obs / 32 times in total: ~0.1 TB/s for the 500 x 500 layer, far below L2 bandwidth).
constexpr int kWarpTile = 64;
constexpr int kWarps = kBlock / 32;

// WARP_TILES = true: as described above. false: ONE copy per CTA (kTile records, elected thread
// 0, __syncthreads() closes every tile) -- kept selectable (hb200_set_tile_mode) for comparison.
// REUSE = true: records marked 2 by pack_layer_kernel take two vertex distances from the previous
// record (LayerCarry, hb200_fast.cuh); results are bit-identical to REUSE = false.
template <int FS, int VARIANT, bool WARP_TILES = true, int MINB = 4, bool REUSE = false>
__global__ void __launch_bounds__(kBlock, MINB) prism_kernel(const PrismArgs a)
{
    typedef Traits<FS> T;
    constexpr int STRIDE = T::mag ? kMagStride : kPrismStride;
    constexpr int GROUPS = WARP_TILES ? kWarps : 1;       // independent record streams per CTA
    constexpr int TILE = WARP_TILES ? kWarpTile : kTile;  // records per bulk copy
    __shared__ alignas(128) double2 tiles[GROUPS][2][TILE * STRIDE / 2];
    __shared__ alignas(8) uint64_t bars[GROUPS][2];

    const int grp = WARP_TILES ? (threadIdx.x >> 5) : 0;
    const bool leader = WARP_TILES ? ((threadIdx.x & 31) == 0) : (threadIdx.x == 0);
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t ic = i < a.n_obs ? i : a.n_obs - 1;
    const double E = a.oe[ic], N = a.on[ic], U = a.ou[ic];

    double acc[T::nout];
#pragma unroll
    for (int c = 0; c < T::nout; c++) acc[c] = 0.0;
    unsigned flags = 0;
    LayerCarry carry;
    carry.r[0] = carry.r[1] = 0.0;
    carry.valid = false;

    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    const double* src = a.packed;

    if (leader) {
        mbar_init(&bars[grp][0], 1);
        mbar_init(&bars[grp][1], 1);
        mbar_fence_init();
    }
    if (WARP_TILES) __syncwarp(); else __syncthreads();
    if (leader && begin < end) {
        const int cnt0 = (int)((end - begin) < TILE ? (end - begin) : TILE);
        bulk_load(tiles[grp][0], src + begin * STRIDE, (unsigned)(cnt0 * STRIDE * sizeof(double)),
                  &bars[grp][0]);
    }
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    for (int64_t t0 = begin; t0 < end; t0 += TILE, buf ^= 1) {
        const int cnt = (int)((end - t0) < TILE ? (end - t0) : TILE);
        // prefetch the next tile into the other buffer (every thread of the group finished
        // reading it at the barrier that closed the previous iteration)
        const int64_t t1 = t0 + TILE;
        if (leader && t1 < end) {
            const int cnt1 = (int)((end - t1) < TILE ? (end - t1) : TILE);
            bulk_load(tiles[grp][buf ^ 1], src + t1 * STRIDE, (unsigned)(cnt1 * STRIDE * sizeof(double)),
                      &bars[grp][buf ^ 1]);
        }
        if (buf == 0) { mbar_wait(&bars[grp][0], phase0); phase0 ^= 1; }
        else { mbar_wait(&bars[grp][1], phase1); phase1 ^= 1; }
        const double2* tile = tiles[grp][buf];
#pragma unroll 1
        for (int s = 0; s < cnt; s++) {
            const double2* p = tile + s * (STRIDE / 2);
            const double2 we = p[0], sn = p[1], bt = p[2], q3 = p[3];
            double prm[3];
            int mark;  // upper word of the record's last double: 0.0, 1.0 (skip) or 2.0 (reuse)
            if (T::mag) {
                const double2 q4 = p[4];
                mark = __double2hiint(q4.y);
                prm[0] = q3.x; prm[1] = q3.y; prm[2] = q4.x;
            } else {
                mark = __double2hiint(q3.y);
                prm[0] = q3.x; prm[1] = 0.0; prm[2] = 0.0;
            }
            if (mark == 0x3ff00000) {  // 1.0: skip (warp-uniform)
                if (REUSE) carry.valid = false;
                continue;
            }
            PairGeom g;
            make_geom(g, E, N, U, we.x, we.y, sn.x, sn.y, bt.x, bt.y);
            if (REUSE) prism_pair<FS, VARIANT>(g, prm, a.rules, acc, flags, &carry, carry.valid && mark == 0x40000000);
            else prism_pair<FS, VARIANT>(g, prm, a.rules, acc, flags);
        }
        // the buffer just read may be refilled in the next iteration
        if (WARP_TILES) __syncwarp(); else __syncthreads();
    }
    if (i < a.n_obs) {
        if (gridDim.y == 1) {
#pragma unroll
            for (int c = 0; c < T::nout; c++) a.out[c * a.n_obs + i] = acc[c] * a.sc.s[c];
        } else {
#pragma unroll
            for (int c = 0; c < T::nout; c++)
                a.out[((int64_t)blockIdx.y * T::nout + c) * a.n_obs + i] = acc[c];
        }
        if (flags && a.flags) atomicOr(a.flags, flags);
    }
}

// out[c][i] = scale[c] * sum_y partial[y][c][i]   (fixed order: deterministic)
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_parts, int nout,
                                       int64_t n_obs, Scales sc, double* __restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nout * n_obs) return;
    const int c = (int)(idx / n_obs);
    double s = 0.0;
    for (int y = 0; y < n_parts; y++) s += partial[(int64_t)y * nout * n_obs + idx];
    out[idx] = s * sc.s[c];
}

// ----------------------------------------------------- singular-point scan
// gravity.py:272-449 as one pass; field selects the predicate set.
__global__ void __launch_bounds__(kBlock) singular_scan_kernel(const PrismArgs a, int field)
{
    __shared__ double2 tile[kTile * kPrismStride / 2];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t ic = i < a.n_obs ? i : a.n_obs - 1;
    const double E = a.oe[ic], N = a.on[ic], U = a.ou[ic];
    bool hit = false;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    const double2* src = reinterpret_cast<const double2*>(a.packed);
    for (int64_t t0 = begin; t0 < end; t0 += kTile) {
        const int cnt = (int)((end - t0) < kTile ? (end - t0) : kTile);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * (kPrismStride / 2); x += kBlock)
            tile[x] = src[t0 * (kPrismStride / 2) + x];
        __syncthreads();
        for (int s = 0; s < cnt; s++) {
            const double2* p = tile + s * (kPrismStride / 2);
            const double2 we = p[0], sn = p[1], bt = p[2];
            PairGeom g;
            g.se[0] = we.y - E; g.se[1] = we.x - E;
            g.sn[0] = sn.y - N; g.sn[1] = sn.x - N;
            g.su[0] = bt.y - U; g.su[1] = bt.x - U;
            const PairPreds pr = make_preds(g);
            bool sing;
            switch (field) {
            case F_EE: sing = pr.n_edge | pr.u_edge; break;
            case F_NN: sing = pr.e_edge | pr.u_edge; break;
            case F_UU: sing = pr.e_edge | pr.n_edge; break;
            case F_EN: sing = pr.u_edge; break;
            case F_EU: sing = pr.n_edge; break;
            case F_NU: sing = pr.e_edge; break;
            default: sing = pr.e_edge | pr.n_edge | pr.u_edge; break;
            }
            hit |= sing;
        }
    }
    if (hit && i < a.n_obs) atomicOr(a.flags, FLAG_SINGULAR);
}

// ------------------------------------------------------------ point kernel
struct PointArgs {
    const double* oe;
    const double* on;
    const double* ou;
    int64_t n_obs;
    const double* packed;
    int64_t n_src;
    int64_t chunk_len;
    double* out;
    double scale;
    unsigned* flags;
    double gconst;  // spherical kernels: G (point masses) or 1 (equivalent sources)
};

constexpr int kPointObs = 4;  // observers per thread (amortises the record loads)

template <int FIELD>
__global__ void __launch_bounds__(kBlock) point_kernel_cart(const PointArgs a)
{
    __shared__ double2 tile[kTile * kPointStride / 2];
    double E[kPointObs], N[kPointObs], U[kPointObs], acc[kPointObs];
    int64_t idx[kPointObs];
#pragma unroll
    for (int o = 0; o < kPointObs; o++) {
        idx[o] = ((int64_t)blockIdx.x * kPointObs + o) * kBlock + threadIdx.x;
        const int64_t ic = idx[o] < a.n_obs ? idx[o] : a.n_obs - 1;
        E[o] = a.oe[ic]; N[o] = a.on[ic]; U[o] = a.ou[ic];
        acc[o] = 0.0;
    }
    unsigned flags = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    const double2* src = reinterpret_cast<const double2*>(a.packed);
    for (int64_t t0 = begin; t0 < end; t0 += kTile) {
        const int cnt = (int)((end - t0) < kTile ? (end - t0) : kTile);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * (kPointStride / 2); x += kBlock)
            tile[x] = src[t0 * (kPointStride / 2) + x];
        __syncthreads();
#pragma unroll 4
        for (int s = 0; s < cnt; s++) {
            const double2 en = tile[2 * s], uw = tile[2 * s + 1];
#pragma unroll
            for (int o = 0; o < kPointObs; o++) {
                const double k = point_kernel<FIELD, false>(E[o] - en.x, N[o] - en.y, U[o] - uw.x, flags);
                acc[o] = fma(uw.y, k, acc[o]);
            }
        }
    }
    // A zero distance (the reference's jitted loop raises ZeroDivisionError) makes 1 / distance,
    // and with it the observer's sum, non-finite for good; so does NaN input. Only such a sum is
    // looked into, off the hot path (two integer instructions per pair otherwise).
    bool suspect = false;
#pragma unroll
    for (int o = 0; o < kPointObs; o++) suspect |= !(fabs(acc[o]) <= 1.7976931348623157e308);
    if (suspect) {
        for (int64_t j = begin; j < end; j++) {
            const double2 en = src[2 * j], uw = src[2 * j + 1];
#pragma unroll
            for (int o = 0; o < kPointObs; o++) {
                const double de = E[o] - en.x, dn = N[o] - en.y, du = U[o] - uw.x;
                if (is_pos_zero(point_d2(de, dn, du))) flags |= FLAG_ZERO_DIV;
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kPointObs; o++) {
        if (idx[o] < a.n_obs) {
            if (gridDim.y == 1) a.out[idx[o]] = acc[o] * a.scale;
            else a.out[(int64_t)blockIdx.y * a.n_obs + idx[o]] = acc[o];
        }
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// point.py:324-354, 436-447 with _forward/utils.py:198-201. The reference evaluates
// cos(lon_p - lon) per pair; here cos/sin of both longitudes are computed once per point and
// combined with the angle-difference identity (two FMAs instead of a float64 cos per pair; the
// rounding differs from a direct cos by <= 1 ulp, like CUDA's cos differs from glibc's).
template <int FIELD>
__global__ void __launch_bounds__(kBlock) point_kernel_sph(const PointArgs a)
{
    __shared__ double2 tile[kTile * kSphStride / 2];
    const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const int64_t ic = i < a.n_obs ? i : a.n_obs - 1;
    const double d2r = kPi / 180.0;
    const double lam = a.oe[ic] * d2r;
    const double phi = a.on[ic] * d2r;
    const double cphi = cos(phi), sphi = sin(phi), rad = a.ou[ic];
    const double clam = cos(lam), slam = sin(lam);
    double acc = 0.0;
    unsigned flags = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    const double2* src = reinterpret_cast<const double2*>(a.packed);
    for (int64_t t0 = begin; t0 < end; t0 += kTile) {
        const int cnt = (int)((end - t0) < kTile ? (end - t0) : kTile);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * (kSphStride / 2); x += kBlock)
            tile[x] = src[t0 * (kSphStride / 2) + x];
        __syncthreads();
        for (int s = 0; s < cnt; s++) {
            const double2 q0 = tile[3 * s], q1 = tile[3 * s + 1], q2 = tile[3 * s + 2];
            const double coslambda = fma(q0.x, clam, q0.y * slam);
            const double cospsi = q1.y * sphi + q1.x * cphi * coslambda;
            const double dr = rad - q2.x;
            const double d2 = dr * dr + 2 * rad * q2.x * (1 - cospsi);
            if (d2 == 0.0) flags |= FLAG_ZERO_DIV;
            const double dist = sqrt(d2);
            double k;
            if (FIELD == F_POT) {
                k = 1 / dist * a.gconst;
            } else {
                const double delta_z = rad - q2.x * cospsi;
                k = -a.gconst * delta_z / (dist * dist * dist);
            }
            acc += q2.y * k;
        }
    }
    if (i < a.n_obs) {
        if (gridDim.y == 1) a.out[i] = acc * a.scale;
        else a.out[(int64_t)blockIdx.y * a.n_obs + i] = acc;
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// ----------------------------------------------------------- dipole kernel
// dipole.py:329-347 / :386-400 with choclo.dipole.magnetic_*:
//   B = mu0/4pi (3 (m.r) r / d^5 - m / d^3),  r = observer - dipole.
// Records: 6 doubles  e n u | m_e m_n m_u. COMP < 0: all three components.
constexpr int kDipoleStride = 6;

__global__ void pack_dipoles_kernel(const double* __restrict__ pe, const double* __restrict__ pn,
                                    const double* __restrict__ pu, const double* __restrict__ me,
                                    const double* __restrict__ mn, const double* __restrict__ mu,
                                    int64_t n, double* __restrict__ packed)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double* q = packed + j * kDipoleStride;
    q[0] = pe[j]; q[1] = pn[j]; q[2] = pu[j];
    q[3] = me[j]; q[4] = mn[j]; q[5] = mu[j];
}

constexpr int kDipoleObs = 2;

template <int COMP>
__global__ void __launch_bounds__(kBlock) dipole_kernel(const PointArgs a)
{
    constexpr int NOUT = COMP < 0 ? 3 : 1;
    __shared__ double2 tile[kTile * kDipoleStride / 2];
    double E[kDipoleObs], N[kDipoleObs], U[kDipoleObs], acc[kDipoleObs][NOUT];
    int64_t idx[kDipoleObs];
#pragma unroll
    for (int o = 0; o < kDipoleObs; o++) {
        idx[o] = ((int64_t)blockIdx.x * kDipoleObs + o) * kBlock + threadIdx.x;
        const int64_t ic = idx[o] < a.n_obs ? idx[o] : a.n_obs - 1;
        E[o] = a.oe[ic]; N[o] = a.on[ic]; U[o] = a.ou[ic];
#pragma unroll
        for (int c = 0; c < NOUT; c++) acc[o][c] = 0.0;
    }
    unsigned flags = 0;
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int64_t end = begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src;
    const double2* src = reinterpret_cast<const double2*>(a.packed);
    for (int64_t t0 = begin; t0 < end; t0 += kTile) {
        const int cnt = (int)((end - t0) < kTile ? (end - t0) : kTile);
        __syncthreads();
        for (int x = threadIdx.x; x < cnt * (kDipoleStride / 2); x += kBlock)
            tile[x] = src[t0 * (kDipoleStride / 2) + x];
        __syncthreads();
#pragma unroll 2
        for (int s = 0; s < cnt; s++) {
            const double2 q0 = tile[3 * s], q1 = tile[3 * s + 1], q2 = tile[3 * s + 2];
            const double me = q1.y, mn = q2.x, mu = q2.y;
#pragma unroll
            for (int o = 0; o < kDipoleObs; o++) {
                const double re = E[o] - q0.x, rn = N[o] - q0.y, ru = U[o] - q1.x;
                const double d2 = re * re + rn * rn + ru * ru;
                if (is_pos_zero(d2)) flags |= FLAG_ZERO_DIV;
                const double inv = point_rsqrt(d2);
                const double inv2 = inv * inv;
                const double inv3 = inv2 * inv;
                const double t = 3.0 * (me * re + mn * rn + mu * ru) * inv2;
                if (COMP < 0) {
                    acc[o][0] = fma(fma(t, re, -me), inv3, acc[o][0]);
                    acc[o][1] = fma(fma(t, rn, -mn), inv3, acc[o][1]);
                    acc[o][2] = fma(fma(t, ru, -mu), inv3, acc[o][2]);
                } else {
                    const double r = COMP == 0 ? re : COMP == 1 ? rn : ru;
                    const double m = COMP == 0 ? me : COMP == 1 ? mn : mu;
                    acc[o][0] = fma(fma(t, r, -m), inv3, acc[o][0]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kDipoleObs; o++) {
        if (idx[o] < a.n_obs) {
#pragma unroll
            for (int c = 0; c < NOUT; c++) {
                if (gridDim.y == 1) a.out[c * a.n_obs + idx[o]] = acc[o][c] * a.scale;
                else a.out[((int64_t)blockIdx.y * NOUT + c) * a.n_obs + idx[o]] = acc[o][c];
            }
        }
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// utils.py:54-74: dense Green's-function matrix, row-major (n_obs, n_src).
// HBM-write bound: 8 bytes per pair; threads run along the source index so
// stores are coalesced.
__global__ void eqs_jacobian_kernel(const double* __restrict__ oe, const double* __restrict__ on,
                                    const double* __restrict__ ou, int64_t n_obs,
                                    const double* __restrict__ pe, const double* __restrict__ pn,
                                    const double* __restrict__ pu, int64_t n_src,
                                    double* __restrict__ jac, unsigned* flags)
{
    // row blocks on grid.x (no 65535 limit), column blocks on grid.y
    const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n_src) return;
    const double se = pe[j], sn = pn[j], su = pu[j];
    const int64_t i0 = (int64_t)blockIdx.x * 16;
    bool zero = false;
#pragma unroll 4
    for (int r = 0; r < 16; r++) {
        const int64_t i = i0 + r;
        if (i >= n_obs) break;
        const double de = oe[i] - se, dn = on[i] - sn, du = ou[i] - su;
        const double d2 = de * de + dn * dn + du * du;
        zero |= d2 == 0.0;  // the reference's jitted loop raises ZeroDivisionError here
        jac[i * n_src + j] = 1.0 / sqrt(d2);
    }
    if (zero && flags) atomicOr(flags, FLAG_ZERO_DIV);
}

// ----------------------------------------------------- FP64 issue-rate probe
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace hb
