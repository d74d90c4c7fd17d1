// hb200_api.cu -- C ABI of libharmonica_b200.so (see include/harmonica_b200.h).
//
// Host entry points: upload caller-owned numpy buffers, shard the work over the
// selected B200s (by observers: disjoint output slices, no collective; or by
// sources: per-device partial fields combined on device 0 by peer copies over
// NVLink + a fixed-order reduce kernel), launch the kernels of
// hb200_kernels.cuh, download. Device entry points (*_dev) launch the same
// kernels on caller-owned device buffers, asynchronously on the caller's stream.
#include <cuda_runtime.h>
#include <cstdlib>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/harmonica_b200.h"
#include "hb200_kernels.cuh"
#include "hb200_tess.cuh"
#include "hb200_tess_leaves.cuh"

using namespace hb;

namespace {

thread_local std::string g_err;
std::mutex g_mu;
int g_variant = 2;
int g_tile_mode = 1;         // 1: one record stream per warp, vertex reuse along layer columns;
                             // 2: without the reuse; 0: one stream per CTA (hb200_set_tile_mode)
int g_source_chunks = 0;     // 0: chosen from the grid size (choose_chunks)
double g_fit_rcond = 2.220446049250313e-16;  // cutoff of the undamped fit (hb200_fit_host.cuh)
std::atomic<uint64_t> g_launches{0};  // kernels launched by this library (hb200_launch_count)

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(expr)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (expr);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(HB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                              \
    } while (0)

size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct Dev {
    int id = -1;
    int sms = 148;
    cudaStream_t st = nullptr;
    char* pool = nullptr;
    size_t pool_bytes = 0;
    size_t used = 0;
    unsigned* d_flags = nullptr;

    int ensure(size_t bytes)
    {
        used = 0;
        if (bytes <= pool_bytes) return HB200_OK;
        if (pool) cudaFree(pool);
        pool = nullptr;
        pool_bytes = 0;
        size_t want = align_up(bytes + (bytes >> 3), 1 << 20);
        cudaError_t e = cudaMalloc(&pool, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(HB200_ENOMEM, "cudaMalloc(%zu bytes) on device %d: %s", want, id,
                        cudaGetErrorString(e));
        }
        pool_bytes = want;
        return HB200_OK;
    }
    template <typename T> T* take(size_t count)
    {
        T* p = reinterpret_cast<T*>(pool + used);
        used += align_up(count * sizeof(T));
        return p;
    }
};

std::vector<Dev> g_devs;

int sm_count_current()
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

// ----------------------------------------------------------- launch helpers
struct FieldSpec {
    int fs;      // field set id (hb::F_* / FS_*)
    int slot;    // first output slot
    int nout;
};

double gravity_scale(int field)
{
    switch (field) {
    case F_POT: return 1.0;
    case F_E: case F_N: return 1e5;            // gravity.py:231-232
    case F_U: return -1e5;                     // gravity.py:228-229 then :231-232
    case F_EE: case F_NN: case F_UU: case F_EN: return 1e9;  // :234-235
    default: return -1e9;                      // g_ez, g_nz: :228-229 then :234-235
    }
}

// split a gravity field mask into fused passes
std::vector<FieldSpec> plan_gravity(uint32_t mask)
{
    std::vector<FieldSpec> plan;
    int slot_of[10], s = 0;
    for (int f = 0; f < 10; f++) slot_of[f] = (mask >> f & 1u) ? s++ : -1;
    uint32_t rest = mask;
    if ((rest & HB200_MASK_TENSOR) == HB200_MASK_TENSOR) {
        plan.push_back({FS_TENSOR6, slot_of[F_EE], 6});
        rest &= ~HB200_MASK_TENSOR;
    }
    if ((rest & HB200_MASK_ACCEL) == HB200_MASK_ACCEL) {
        plan.push_back({FS_ACC3, slot_of[F_E], 3});
        rest &= ~HB200_MASK_ACCEL;
    }
    for (int f = 0; f < 10; f++)
        if (rest >> f & 1u) plan.push_back({f, slot_of[f], 1});
    return plan;
}

int popcount(uint32_t m) { return __builtin_popcount(m); }

int choose_chunks(int64_t n_obs, int64_t n_src, int obs_per_block, int sms, int64_t* chunk_len)
{
    const int64_t obs_blocks = (n_obs + obs_per_block - 1) / obs_per_block;
    const int64_t tiles = std::max<int64_t>(1, (n_src + kTile - 1) / kTile);
    // All CTAs do equal work, so a grid of T CTAs over S resident slots runs at
    // (T/S)/ceil(T/S) of full speed (wave quantisation; ncu showed 2.2 waves = 73 %
    // for the 500x500 layer). Split the source list until T >= ~20 S (>= 95 %).
    const int64_t target = (int64_t)sms * 6 * 20;
    int64_t chunks = 1;
    if (obs_blocks < target) chunks = std::min<int64_t>(tiles, (target + obs_blocks - 1) / obs_blocks);
    chunks = std::min<int64_t>(chunks, 256);
    // pinned by the caller (hb200_set_source_chunks): the association of the partial sums then
    // does not depend on the number of observers in the call
    if (g_source_chunks > 0) chunks = std::min<int64_t>(tiles, g_source_chunks);
    int64_t tiles_per_chunk = (tiles + chunks - 1) / chunks;
    *chunk_len = tiles_per_chunk * kTile;
    chunks = (n_src + *chunk_len - 1) / *chunk_len;
    return (int)std::max<int64_t>(1, chunks);
}

// set by prism_layer_dev_impl around its passes: the records come from pack_layer_kernel, which
// marks the ones that share vertices with their predecessor (one host thread per device)
thread_local bool t_layer_records = false;

template <int FS> void launch_prism_fs(const PrismArgs& a, dim3 grid, cudaStream_t st)
{
    // the fields of prism_layer.gravity; the potential kernel is at its register limit (the carry spills)
    constexpr bool gravity = FS <= F_NU && FS != F_POT;
    if (g_variant == 0) prism_kernel<FS, 0><<<grid, kBlock, 0, st>>>(a);
    else if (g_variant == 1) prism_kernel<FS, 1><<<grid, kBlock, 0, st>>>(a);
    else if (g_tile_mode == 0) prism_kernel<FS, 2, false><<<grid, kBlock, 0, st>>>(a);
    else if (gravity && g_tile_mode == 1 && t_layer_records)
        prism_kernel<FS, 2, true, 4, gravity><<<grid, kBlock, 0, st>>>(a);  // vertex reuse along layer columns
    else prism_kernel<FS, 2, true><<<grid, kBlock, 0, st>>>(a);
}

void launch_prism_any(int fs, const PrismArgs& a, dim3 grid, cudaStream_t st)
{
    switch (fs) {
    case F_POT: launch_prism_fs<F_POT>(a, grid, st); break;
    case F_E: launch_prism_fs<F_E>(a, grid, st); break;
    case F_N: launch_prism_fs<F_N>(a, grid, st); break;
    case F_U: launch_prism_fs<F_U>(a, grid, st); break;
    case F_EE: launch_prism_fs<F_EE>(a, grid, st); break;
    case F_NN: launch_prism_fs<F_NN>(a, grid, st); break;
    case F_UU: launch_prism_fs<F_UU>(a, grid, st); break;
    case F_EN: launch_prism_fs<F_EN>(a, grid, st); break;
    case F_EU: launch_prism_fs<F_EU>(a, grid, st); break;
    case F_NU: launch_prism_fs<F_NU>(a, grid, st); break;
    case FS_ACC3: launch_prism_fs<FS_ACC3>(a, grid, st); break;
    case FS_TENSOR6: launch_prism_fs<FS_TENSOR6>(a, grid, st); break;
    case FS_MAG_B: launch_prism_fs<FS_MAG_B>(a, grid, st); break;
    case FS_MAG_E: launch_prism_fs<FS_MAG_E>(a, grid, st); break;
    case FS_MAG_N: launch_prism_fs<FS_MAG_N>(a, grid, st); break;
    case FS_MAG_U: launch_prism_fs<FS_MAG_U>(a, grid, st); break;
    }
}

// One fused pass over packed prism records. `partial` must hold
// chunks*nout*n_obs doubles when the source list is split.
int run_prism_pass(int fs, int nout, const double* oe, const double* on, const double* ou,
                   int64_t n_obs, const double* packed, int64_t n_src, const Scales& sc,
                   unsigned rules, double* out, double* partial, size_t partial_bytes,
                   unsigned* d_flags, int sms, cudaStream_t st)
{
    if (n_obs == 0) return HB200_OK;
    if (n_src == 0) {
        CU(cudaMemsetAsync(out, 0, sizeof(double) * nout * n_obs, st));
        return HB200_OK;
    }
    int64_t chunk_len = 0;
    int chunks = choose_chunks(n_obs, n_src, kBlock, sms, &chunk_len);
    if (chunks > 1 && (size_t)chunks * nout * n_obs * sizeof(double) > partial_bytes) {
        chunks = 1;
        chunk_len = n_src;
    }
    PrismArgs a;
    a.oe = oe; a.on = on; a.ou = ou; a.n_obs = n_obs;
    a.packed = packed; a.n_src = n_src; a.chunk_len = chunk_len;
    a.out = chunks > 1 ? partial : out;
    a.sc = sc; a.rules = rules; a.flags = d_flags;
    dim3 grid((unsigned)((n_obs + kBlock - 1) / kBlock), (unsigned)chunks);
    launch_prism_any(fs, a, grid, st);
    CU(cudaGetLastError());
    g_launches += chunks > 1 ? 2 : 1;
    if (chunks > 1) {
        const int64_t total = (int64_t)nout * n_obs;
        reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, chunks, nout,
                                                                                n_obs, sc, out);
        CU(cudaGetLastError());
    }
    return HB200_OK;
}

size_t partial_bytes_for(int64_t n_obs, int64_t n_src, int nout, int obs_per_block, int sms)
{
    int64_t chunk_len;
    int chunks = choose_chunks(n_obs, n_src, obs_per_block, sms, &chunk_len);
    return chunks > 1 ? align_up((size_t)chunks * nout * n_obs * sizeof(double)) : 0;
}

template <int FIELD> void launch_point_cart(const PointArgs& a, dim3 grid, cudaStream_t st)
{
    point_kernel_cart<FIELD><<<grid, kBlock, 0, st>>>(a);
}

int run_point_pass(int field, int spherical, const double* oe, const double* on, const double* ou,
                   int64_t n_obs, const double* packed, int64_t n_src, double scale, double* out,
                   double* partial, size_t partial_bytes, unsigned* d_flags, int sms,
                   cudaStream_t st, double gconst = kG)
{
    if (n_obs == 0) return HB200_OK;
    if (n_src == 0) {
        CU(cudaMemsetAsync(out, 0, sizeof(double) * n_obs, st));
        return HB200_OK;
    }
    const int opb = spherical ? kBlock : kBlock * kPointObs;
    int64_t chunk_len = 0;
    int chunks = choose_chunks(n_obs, n_src, opb, sms, &chunk_len);
    if (chunks > 1 && (size_t)chunks * n_obs * sizeof(double) > partial_bytes) {
        chunks = 1;
        chunk_len = n_src;
    }
    PointArgs a;
    a.oe = oe; a.on = on; a.ou = ou; a.n_obs = n_obs;
    a.packed = packed; a.n_src = n_src; a.chunk_len = chunk_len;
    a.out = chunks > 1 ? partial : out;
    a.scale = scale; a.flags = d_flags; a.gconst = gconst;
    dim3 grid((unsigned)((n_obs + opb - 1) / opb), (unsigned)chunks);
    if (spherical) {
        if (field == F_POT) point_kernel_sph<F_POT><<<grid, kBlock, 0, st>>>(a);
        else point_kernel_sph<F_U><<<grid, kBlock, 0, st>>>(a);
    } else {
        switch (field) {
        case F_POT: launch_point_cart<F_POT>(a, grid, st); break;
        case F_E: launch_point_cart<F_E>(a, grid, st); break;
        case F_N: launch_point_cart<F_N>(a, grid, st); break;
        case F_U: launch_point_cart<F_U>(a, grid, st); break;
        case F_EE: launch_point_cart<F_EE>(a, grid, st); break;
        case F_NN: launch_point_cart<F_NN>(a, grid, st); break;
        case F_UU: launch_point_cart<F_UU>(a, grid, st); break;
        case F_EN: launch_point_cart<F_EN>(a, grid, st); break;
        case F_EU: launch_point_cart<F_EU>(a, grid, st); break;
        case F_NU: launch_point_cart<F_NU>(a, grid, st); break;
        }
    }
    CU(cudaGetLastError());
    g_launches += chunks > 1 ? 2 : 1;
    if (chunks > 1) {
        Scales sc;
        sc.s[0] = scale;
        reduce_partials_kernel<<<(unsigned)((n_obs + 255) / 256), 256, 0, st>>>(partial, chunks, 1,
                                                                               n_obs, sc, out);
        CU(cudaGetLastError());
    }
    return HB200_OK;
}

// ---------------------------------------------------------- device-level ops
// Workspace layout for the prism ops: [packed records][chunk partials].
size_t prism_ws_bytes(int64_t n_obs, int64_t n_src, int nf, int sms)
{
    return align_up((size_t)n_src * kMagStride * sizeof(double))
         + partial_bytes_for(n_obs, n_src, std::min(nf, 6), kBlock, sms) + 256;
}

struct Ws {
    char* base;
    size_t bytes, used;
    Ws(void* p, size_t b) : base((char*)p), bytes(b), used(0) {}
    double* take(size_t nbytes)
    {
        nbytes = align_up(nbytes);
        if (used + nbytes > bytes) return nullptr;
        double* p = (double*)(base + used);
        used += nbytes;
        return p;
    }
    size_t left() const { return bytes - used; }
};

int gravity_passes(const double* oe, const double* on, const double* ou, int64_t n_obs,
                   const double* packed, int64_t n_src, uint32_t mask, bool raw, double* out,
                   unsigned* d_flags, Ws& ws, int sms, cudaStream_t st)
{
    double* partial = (double*)(ws.base + ws.used);
    const size_t partial_bytes = ws.left();
    for (const FieldSpec& p : plan_gravity(mask)) {
        Scales sc;
        for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
        if (!raw) {
            if (p.fs == FS_TENSOR6) for (int c = 0; c < 6; c++) sc.s[c] = gravity_scale(F_EE + c);
            else if (p.fs == FS_ACC3) for (int c = 0; c < 3; c++) sc.s[c] = gravity_scale(F_E + c);
            else sc.s[0] = gravity_scale(p.fs);
        }
        int rc = run_prism_pass(p.fs, p.nout, oe, on, ou, n_obs, packed, n_src, sc, 0u,
                                out + (int64_t)p.slot * n_obs, partial, partial_bytes, d_flags, sms,
                                st);
        if (rc) return rc;
    }
    return HB200_OK;
}

int prism_gravity_dev_impl(const double* oe, const double* on, const double* ou, int64_t n_obs,
                           const double* prisms, const double* density, int64_t n_prisms,
                           uint32_t mask, bool raw, double* out, unsigned* d_flags, void* wsp,
                           size_t ws_bytes, int sms, cudaStream_t st)
{
    Ws ws(wsp, ws_bytes);
    double* packed = ws.take((size_t)std::max<int64_t>(n_prisms, 1) * kPrismStride * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    if (n_prisms > 0) {
        pack_prisms_kernel<<<(unsigned)((n_prisms + 255) / 256), 256, 0, st>>>(prisms, density,
                                                                              n_prisms, packed);
        CU(cudaGetLastError());
        g_launches += 1;
    }
    return gravity_passes(oe, on, ou, n_obs, packed, n_prisms, mask, raw, out, d_flags, ws, sms, st);
}

int prism_layer_dev_impl(const double* oe, const double* on, const double* ou, int64_t n_obs,
                         const double* east_c, int64_t n_east, const double* north_c,
                         int64_t n_north, const double* bottom, const double* top,
                         const double* density, double thr, uint32_t mask, bool raw, double* out,
                         unsigned* d_flags, void* wsp, size_t ws_bytes, int sms, cudaStream_t st)
{
    const int64_t n_src = n_east * n_north;
    Ws ws(wsp, ws_bytes);
    double* packed = ws.take((size_t)std::max<int64_t>(n_src, 1) * kPrismStride * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    if (n_src > 0) {
        pack_layer_kernel<<<(unsigned)((n_src + 255) / 256), 256, 0, st>>>(
            east_c, north_c, n_east, n_north, bottom, top, density, thr, packed);
        CU(cudaGetLastError());
        g_launches += 1;
    }
    t_layer_records = true;
    const int rc = gravity_passes(oe, on, ou, n_obs, packed, n_src, mask, raw, out, d_flags, ws, sms, st);
    t_layer_records = false;
    return rc;
}

int prism_magnetic_dev_impl(const double* oe, const double* on, const double* ou, int64_t n_obs,
                            const double* prisms, const double* me, const double* mn,
                            const double* mu, int64_t n_prisms, uint32_t cmask, uint32_t rules,
                            bool raw, double* out, unsigned* d_flags, void* wsp, size_t ws_bytes,
                            int sms, cudaStream_t st)
{
    Ws ws(wsp, ws_bytes);
    double* packed = ws.take((size_t)std::max<int64_t>(n_prisms, 1) * kMagStride * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    if (n_prisms > 0) {
        pack_mag_kernel<<<(unsigned)((n_prisms + 255) / 256), 256, 0, st>>>(prisms, me, mn, mu,
                                                                           n_prisms, packed);
        CU(cudaGetLastError());
        g_launches += 1;
    }
    double* partial = (double*)(ws.base + ws.used);
    const size_t partial_bytes = ws.left();
    // magnetic.py:195-197, :271: T -> nT; mu0/4pi as choclo
    const double mu0 = 4 * kPi * 1e-7;
    const double cm = mu0 / 4 / kPi;
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = raw ? 1.0 : cm * 1e9;
    if (cmask == HB200_B_ALL)
        return run_prism_pass(FS_MAG_B, 3, oe, on, ou, n_obs, packed, n_prisms, sc, rules, out,
                              partial, partial_bytes, d_flags, sms, st);
    int slot = 0;
    for (int c = 0; c < 3; c++) {
        if (!(cmask >> c & 1u)) continue;
        int rc = run_prism_pass(FS_MAG_E + c, 1, oe, on, ou, n_obs, packed, n_prisms, sc, rules,
                                out + (int64_t)slot * n_obs, partial, partial_bytes, d_flags, sms,
                                st);
        if (rc) return rc;
        slot++;
    }
    return HB200_OK;
}

double point_scale(int field)
{
    // point.py:253-260
    return gravity_scale(field);
}

size_t point_ws_bytes(int64_t n_obs, int64_t n_src, int sms)
{
    // the spherical kernel has kBlock observers per CTA, the Cartesian one kBlock * kPointObs
    return align_up((size_t)n_src * kSphStride * sizeof(double))
         + std::max(partial_bytes_for(n_obs, n_src, 1, kBlock, sms),
                    partial_bytes_for(n_obs, n_src, 1, kBlock * kPointObs, sms)) + 256;
}

int point_gravity_dev_impl(const double* oe, const double* on, const double* ou, int64_t n_obs,
                           const double* pe, const double* pn, const double* pu, const double* w,
                           int64_t n_src, uint32_t mask, int spherical, int scale_by_G, bool raw,
                           double* out, unsigned* d_flags, void* wsp, size_t ws_bytes, int sms,
                           cudaStream_t st)
{
    Ws ws(wsp, ws_bytes);
    const int stride = spherical ? kSphStride : kPointStride;
    double* packed = ws.take((size_t)std::max<int64_t>(n_src, 1) * stride * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    if (n_src > 0) {
        const unsigned nb = (unsigned)((n_src + 255) / 256);
        if (spherical) pack_points_sph_kernel<<<nb, 256, 0, st>>>(pe, pn, pu, w, n_src, packed);
        else pack_points_kernel<<<nb, 256, 0, st>>>(pe, pn, pu, w, n_src, scale_by_G, packed);
        CU(cudaGetLastError());
        g_launches += 1;
    }
    double* partial = (double*)(ws.base + ws.used);
    const size_t partial_bytes = ws.left();
    int slot = 0;
    for (int f = 0; f < 10; f++) {
        if (!(mask >> f & 1u)) continue;
        const double scale = (raw || !scale_by_G) ? 1.0 : point_scale(f);
        int rc = run_point_pass(f, spherical, oe, on, ou, n_obs, packed, n_src, scale,
                                out + (int64_t)slot * n_obs, partial, partial_bytes, d_flags, sms,
                                st, scale_by_G ? kG : 1.0);
        if (rc) return rc;
        slot++;
    }
    return HB200_OK;
}

// dipole_magnetic: component_mask 7 = fused b, else single components one pass each
size_t dipole_ws_bytes(int64_t n_obs, int64_t n_src, int sms)
{
    return align_up((size_t)n_src * kDipoleStride * sizeof(double))
         + partial_bytes_for(n_obs, n_src, 3, kBlock * kDipoleObs, sms) + 256;
}

int dipole_magnetic_dev_impl(const double* oe, const double* on, const double* ou, int64_t n_obs,
                             const double* pe, const double* pn, const double* pu,
                             const double* me, const double* mn, const double* mu, int64_t n_src,
                             uint32_t cmask, bool raw, double* out, unsigned* d_flags, void* wsp,
                             size_t ws_bytes, int sms, cudaStream_t st)
{
    if (n_obs == 0) return HB200_OK;
    const int nf = popcount(cmask);
    if (n_src == 0) {
        CU(cudaMemsetAsync(out, 0, sizeof(double) * nf * n_obs, st));
        return HB200_OK;
    }
    Ws ws(wsp, ws_bytes);
    double* packed = ws.take((size_t)n_src * kDipoleStride * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    pack_dipoles_kernel<<<(unsigned)((n_src + 255) / 256), 256, 0, st>>>(pe, pn, pu, me, mn, mu, n_src,
                                                                        packed);
    CU(cudaGetLastError());
    g_launches += 1;
    double* partial = (double*)(ws.base + ws.used);
    const size_t partial_bytes = ws.left();
    // dipole.py:195-197, :270: T -> nT; mu0/4pi as choclo
    const double mu0 = 4 * kPi * 1e-7;
    const double scale = raw ? 1.0 : mu0 / 4 / kPi * 1e9;
    const int opb = kBlock * kDipoleObs;
    int slot = 0;
    for (int pass = 0; pass < 3; pass++) {
        int comp, nout;
        if (cmask == HB200_B_ALL) {
            if (pass > 0) break;
            comp = -1; nout = 3;
        } else {
            if (!(cmask >> pass & 1u)) continue;
            comp = pass; nout = 1;
        }
        int64_t chunk_len = 0;
        int chunks = choose_chunks(n_obs, n_src, opb, sms, &chunk_len);
        if (chunks > 1 && (size_t)chunks * nout * n_obs * sizeof(double) > partial_bytes) {
            chunks = 1;
            chunk_len = n_src;
        }
        PointArgs a;
        a.oe = oe; a.on = on; a.ou = ou; a.n_obs = n_obs;
        a.packed = packed; a.n_src = n_src; a.chunk_len = chunk_len;
        double* dst = out + (int64_t)slot * n_obs;
        a.out = chunks > 1 ? partial : dst;
        a.scale = scale; a.flags = d_flags; a.gconst = 1.0;
        dim3 grid((unsigned)((n_obs + opb - 1) / opb), (unsigned)chunks);
        if (comp < 0) dipole_kernel<-1><<<grid, kBlock, 0, st>>>(a);
        else if (comp == 0) dipole_kernel<0><<<grid, kBlock, 0, st>>>(a);
        else if (comp == 1) dipole_kernel<1><<<grid, kBlock, 0, st>>>(a);
        else dipole_kernel<2><<<grid, kBlock, 0, st>>>(a);
        CU(cudaGetLastError());
        g_launches += chunks > 1 ? 2 : 1;
        if (chunks > 1) {
            Scales sc;
            for (int c = 0; c < 6; c++) sc.s[c] = scale;
            const int64_t total = (int64_t)nout * n_obs;
            reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, chunks,
                                                                                   nout, n_obs, sc, dst);
            CU(cudaGetLastError());
        }
        slot += nout;
    }
    return HB200_OK;
}

// tesseroid_gravity: one field per pass (potential or g_z); workspace = [packed][partials]
int g_tess_variant = 9;  // 0: first build; 1: root records + deferred walks; 2: 1 + fast far field;
                         // 3: 2 + the library's own sin / cos / acos in the walks; 4, 5: register
                         // experiments; 6 (7, 8): root pass and walks as two kernels; 9: as 6 with the
                         // walks done by groups of 16 lanes from a work list (default)

// chunking of the two-kernel variant (hb200_tess.cuh): short chunks, at most kTessMaxChunks
int tess_two_kernel_chunks(int64_t n_src, int64_t* chunk_len)
{
    int64_t len = kTessChunk;
    if ((n_src + len - 1) / len > kTessMaxChunks) {
        len = (n_src + kTessMaxChunks - 1) / kTessMaxChunks;
        len = (len + kTessTile - 1) / kTessTile * kTessTile;
    }
    *chunk_len = len;
    return (int)std::max<int64_t>(1, (n_src + len - 1) / len);
}

size_t tesseroid_ws_bytes(int64_t n_obs, int64_t n_src, int sms)
{
    // two-kernel variant, one observer batch: [records][root partials + walk sums: 2 chunks x
    // batch doubles][lists: chunks x cap x batch u16][counts + resume offsets: 2 chunks x batch]
    int64_t chunk_len;
    const int chunks = tess_two_kernel_chunks(n_src, &chunk_len);
    const int64_t batch = std::min<int64_t>(std::max<int64_t>(n_obs, 1), kTessObsBatch);
    const size_t two_kernel = align_up((size_t)(1 + kTessWalkSlices) * chunks * batch * sizeof(double))
                            + align_up((size_t)chunks * kTessListCap * batch * sizeof(unsigned short))
                            + align_up((size_t)2 * chunks * batch * sizeof(int))
                            // work list and redo list of the cooperative walks, their counters
                            + 2 * align_up((size_t)chunks * batch * sizeof(int)) + align_up(4 * sizeof(int));
    return align_up((size_t)std::max<int64_t>(n_src, 1) * kTessRec * sizeof(double))
         + std::max(partial_bytes_for(n_obs, n_src, 1, kTessBlock, sms), two_kernel) + 256;
}

// density0 / density1: density at the lower / upper radial quadrature node of every tesseroid
// (the same array twice for homogeneous tesseroids)
int tesseroid_dev_impl(const double* lon, const double* lat, const double* rad, int64_t n_obs,
                       const double* tesseroids, const double* density0, const double* density1,
                       int64_t n_tess, int field, int radial, bool raw, double* out,
                       unsigned* d_flags, void* wsp, size_t ws_bytes, int sms, cudaStream_t st)
{
    if (field != F_POT && field != F_U) return fail(HB200_EINVAL, "tesseroids: potential or g_z only");
    Ws ws(wsp, ws_bytes);
    int variant = g_tess_variant;
    if (variant >= 6) {  // 16-bit offsets inside a chunk: absurdly long lists use the one-kernel variant
        int64_t cl;
        tess_two_kernel_chunks(n_tess, &cl);
        if (cl > 65535) variant = 3;
    }
    double* packed = ws.take((size_t)std::max<int64_t>(n_tess, 1) * kTessRec * sizeof(double));
    if (!packed) return fail(HB200_EINVAL, "workspace too small");
    if (n_obs == 0) return HB200_OK;
    if (n_tess == 0) {
        CU(cudaMemsetAsync(out, 0, sizeof(double) * n_obs, st));
        return HB200_OK;
    }
    if (variant == 0)
        pack_tesseroids_kernel<<<(unsigned)((n_tess + 255) / 256), 256, 0, st>>>(
            tesseroids, density0, density1, n_tess, packed);
    else if (variant == 1)
        pack_tesseroid_records_kernel<<<(unsigned)((n_tess + 127) / 128), 128, 0, st>>>(
            tesseroids, density0, density1, n_tess, packed);
    else
        pack_tesseroid_fast_records_kernel<<<(unsigned)((n_tess + 127) / 128), 128, 0, st>>>(
            tesseroids, density0, density1, n_tess, field == F_POT ? 1.0 : 2.5, radial, packed);
    CU(cudaGetLastError());
    // tesseroid_gravity.py:222-225: g_z is the downward component in mGal
    const double scale = raw ? 1.0 : (field == F_U ? -1e5 : 1.0);
    if (variant >= 6) {
        // root pass + walk pass + ordered sum (hb200_tess.cuh), in observer batches
        int64_t chunk_len = 0;
        const int chunks = tess_two_kernel_chunks(n_tess, &chunk_len);
        const int64_t batch = std::min<int64_t>(n_obs, kTessObsBatch);
        double* parts = ws.take((size_t)(1 + kTessWalkSlices) * chunks * batch * sizeof(double));
        unsigned short* list =
            (unsigned short*)ws.take((size_t)chunks * kTessListCap * batch * sizeof(unsigned short));
        int* count = (int*)ws.take((size_t)2 * chunks * batch * sizeof(int));  // counts, resume offsets
        int* items = (int*)ws.take((size_t)chunks * batch * sizeof(int));
        int* redo_items = (int*)ws.take((size_t)chunks * batch * sizeof(int));
        int* counters = (int*)ws.take(4 * sizeof(int));
        if (!parts || !list || !count || !items || !redo_items || !counters)
            return fail(HB200_EINVAL, "workspace too small");
        for (int64_t o0 = 0; o0 < n_obs; o0 += batch) {
            const int64_t nb = std::min(batch, n_obs - o0);
            TessArgs a;
            a.lon = lon + o0; a.lat = lat + o0; a.rad = rad + o0; a.n_obs = nb;
            a.packed = packed; a.n_src = n_tess; a.chunk_len = chunk_len;
            a.out = parts; a.scale = scale;
            a.ratio = field == F_POT ? 1.0 : 2.5;
            a.radial = radial; a.flags = d_flags;
            dim3 grid_r((unsigned)((nb + kTessRootBlock - 1) / kTessRootBlock), (unsigned)chunks);
            const bool coop = variant >= 9;  // walks by groups of 16 lanes (tesseroid_coop_walk_kernel)
            const int slices = coop ? 1 : kTessWalkSlices;
            dim3 grid_w((unsigned)((nb + kTessBlock - 1) / kTessBlock), (unsigned)chunks, kTessWalkSlices);
            const unsigned grid_c = (unsigned)(sms * kCoopCtasPerSm);  // persistent: one CTA per slot
            double* walk_sum = parts + (size_t)chunks * nb;
            if (coop) {  // only the lists on the work list are written
                CU(cudaMemsetAsync(counters, 0, 4 * sizeof(int), st));
                CU(cudaMemsetAsync(walk_sum, 0, (size_t)chunks * nb * sizeof(double), st));
            }
            int* items_arg = coop ? items : nullptr;
            // variants 6 / 7 / 8: the root kernel compiled for 4 / 6 / 8 resident CTAs
            if (field == F_POT) {
                if (variant == 7) tesseroid_root_kernel<F_POT, 6><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                else if (variant == 8) tesseroid_root_kernel<F_POT, 8><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                else tesseroid_root_kernel<F_POT, 4><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                if (coop) {
                    tesseroid_coop_walk_kernel<F_POT, OwnTrig><<<grid_c, kCoopBlock, 0, st>>>(
                        a, chunks, list, count, items, counters, redo_items, walk_sum);
                    tesseroid_redo_kernel<F_POT, OwnTrig><<<(unsigned)sms, kTessBlock, 0, st>>>(
                        a, chunks, list, count, counters, redo_items, walk_sum);
                }
                else tesseroid_walk_kernel<F_POT, OwnTrig><<<grid_w, kTessBlock, 0, st>>>(a, list, count, walk_sum);
            } else {
                if (variant == 7) tesseroid_root_kernel<F_U, 6><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                else if (variant == 8) tesseroid_root_kernel<F_U, 8><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                else tesseroid_root_kernel<F_U, 4><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count, items_arg, counters);
                if (coop) {
                    tesseroid_coop_walk_kernel<F_U, OwnTrig><<<grid_c, kCoopBlock, 0, st>>>(
                        a, chunks, list, count, items, counters, redo_items, walk_sum);
                    tesseroid_redo_kernel<F_U, OwnTrig><<<(unsigned)sms, kTessBlock, 0, st>>>(
                        a, chunks, list, count, counters, redo_items, walk_sum);
                }
                else tesseroid_walk_kernel<F_U, OwnTrig><<<grid_w, kTessBlock, 0, st>>>(a, list, count, walk_sum);
            }
            CU(cudaGetLastError());
            Scales sc;
            sc.s[0] = scale;
            reduce_partials_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(
                parts, (1 + slices) * chunks, 1, nb, sc, out + o0);
            CU(cudaGetLastError());
            g_launches += coop ? 4 : 3;
        }
        g_launches += 1;  // the pack kernel
        return HB200_OK;
    }
    double* partial = (double*)(ws.base + ws.used);
    const size_t partial_bytes = ws.left();
    int64_t chunk_len = 0;
    int chunks = choose_chunks(n_obs, n_tess, kTessBlock, sms, &chunk_len);
    if (chunks > 1 && (size_t)chunks * n_obs * sizeof(double) > partial_bytes) {
        chunks = 1;
        chunk_len = n_tess;
    }
    TessArgs a;
    a.lon = lon; a.lat = lat; a.rad = rad; a.n_obs = n_obs;
    a.packed = packed; a.n_src = n_tess; a.chunk_len = chunk_len;
    a.out = chunks > 1 ? partial : out;
    a.scale = scale;
    a.ratio = field == F_POT ? 1.0 : 2.5;  // tesseroid_gravity.py:33 DISTANCE_SIZE_RATII
    a.radial = radial;
    a.flags = d_flags;
    dim3 grid((unsigned)((n_obs + kTessBlock - 1) / kTessBlock), (unsigned)chunks);
    if (variant == 0) {
        if (field == F_POT) tesseroid_kernel<F_POT><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_kernel<F_U><<<grid, kTessBlock, 0, st>>>(a);
    } else if (variant == 1) {
        if (field == F_POT) tesseroid_deferred_kernel<F_POT, false><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_deferred_kernel<F_U, false><<<grid, kTessBlock, 0, st>>>(a);
    } else if (variant == 2) {
        if (field == F_POT) tesseroid_deferred_kernel<F_POT, true><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_deferred_kernel<F_U, true><<<grid, kTessBlock, 0, st>>>(a);
    } else if (variant == 3) {
        if (field == F_POT) tesseroid_deferred_kernel<F_POT, true, OwnTrig><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_deferred_kernel<F_U, true, OwnTrig><<<grid, kTessBlock, 0, st>>>(a);
    } else if (variant == 4) {  // experiment: 80 registers, 24 resident warps
        if (field == F_POT) tesseroid_deferred_kernel<F_POT, true, OwnTrig, 12><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_deferred_kernel<F_U, true, OwnTrig, 12><<<grid, kTessBlock, 0, st>>>(a);
    } else {  // experiment: 64 registers, 32 resident warps
        if (field == F_POT) tesseroid_deferred_kernel<F_POT, true, OwnTrig, 16><<<grid, kTessBlock, 0, st>>>(a);
        else tesseroid_deferred_kernel<F_U, true, OwnTrig, 16><<<grid, kTessBlock, 0, st>>>(a);
    }
    CU(cudaGetLastError());
    g_launches += chunks > 1 ? 3 : 2;
    if (chunks > 1) {
        Scales sc;
        sc.s[0] = scale;
        reduce_partials_kernel<<<(unsigned)((n_obs + 255) / 256), 256, 0, st>>>(partial, chunks, 1,
                                                                               n_obs, sc, out);
        CU(cudaGetLastError());
    }
    return HB200_OK;
}

// ------------------------------------------------------------- host sharding
int lazy_init()
{
    if (!g_devs.empty()) return HB200_OK;
    return hb200_init(nullptr, 0);
}

struct HostArray {
    const double* host;  // may be null (optional array)
    int64_t rows;        // rows (sharded along rows when `sharded`)
    int width;           // doubles per row
    bool sharded;        // follows the source shard
};

struct ShardPlan {
    int ndev;
    bool by_sources;
};

ShardPlan plan_shards(int64_t n_obs, int64_t n_src, int shard_mode, bool src_shardable)
{
    ShardPlan p;
    const int avail = (int)g_devs.size();
    const double pairs = (double)n_obs * (double)n_src;
    int want = (int)std::min<double>(avail, std::max(1.0, pairs / 2e8));
    p.by_sources = false;
    if (shard_mode == HB200_SHARD_SOURCES && src_shardable) p.by_sources = true;
    else if (shard_mode == HB200_SHARD_AUTO && src_shardable && want > 1
             && n_obs < (int64_t)4096 * want && n_src >= (int64_t)4096 * want)
        p.by_sources = true;
    if (shard_mode != HB200_SHARD_AUTO) want = avail;
    const int64_t units = p.by_sources ? n_src : n_obs;
    want = (int)std::max<int64_t>(1, std::min<int64_t>(want, units));
    p.ndev = want;
    return p;
}

// Generic host driver. `launch` runs the op on one device for the observer
// range and source range that device owns.
template <typename Launch>
int run_host_job(const double* oe, const double* on, const double* ou, int64_t n_obs,
                 std::vector<HostArray> arrays, int64_t n_src, int nf,
                 int shard_mode, bool src_shardable, const Scales& final_scales, double* out,
                 uint32_t* flags, Launch launch,
                 size_t (*ws_fn)(int64_t, int64_t, int, int))
{
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = lazy_init();
    if (rc) return rc;
    if (flags) *flags = 0;
    if (n_obs == 0) return HB200_OK;
    const ShardPlan plan = plan_shards(n_obs, n_src, shard_mode, src_shardable);
    const int nd = plan.ndev;

    struct Part {
        int64_t olo, ohi, slo, shi;
        double *d_oe, *d_on, *d_ou, *d_out;
        std::vector<double*> d_arr;
        void* ws;
        size_t ws_bytes;
    };
    std::vector<Part> parts(nd);
    for (int d = 0; d < nd; d++) {
        Part& p = parts[d];
        if (plan.by_sources) {
            p.olo = 0; p.ohi = n_obs;
            p.slo = n_src * d / nd; p.shi = n_src * (d + 1) / nd;
        } else {
            p.olo = n_obs * d / nd; p.ohi = n_obs * (d + 1) / nd;
            p.slo = 0; p.shi = n_src;
        }
    }
    // phase 1: allocate + upload + launch on every device. Uploads from pageable numpy buffers
    // block the issuing host thread, so each device gets its own host thread: all devices are
    // fed (and start computing) concurrently.
    auto feed = [&](int d) -> int {
        int rc = HB200_OK;
        Dev& dev = g_devs[d];
        Part& p = parts[d];
        CU(cudaSetDevice(dev.id));
        const int64_t no = p.ohi - p.olo, ns = p.shi - p.slo;
        size_t need = 3 * align_up(no * sizeof(double)) + align_up((size_t)nf * no * sizeof(double));
        for (const HostArray& a : arrays) {
            const int64_t rows = a.sharded ? ns : a.rows;
            need += align_up((size_t)std::max<int64_t>(rows, 1) * a.width * sizeof(double));
        }
        p.ws_bytes = ws_fn(no, ns, nf, dev.sms);
        need += align_up(p.ws_bytes);
        if (plan.by_sources && d == 0)
            need += align_up((size_t)nd * nf * n_obs * sizeof(double)) + 256;
        rc = dev.ensure(need + 4096);
        if (rc) return rc;
        p.d_oe = dev.take<double>(no);
        p.d_on = dev.take<double>(no);
        p.d_ou = dev.take<double>(no);
        p.d_out = dev.take<double>((size_t)nf * no);
        CU(cudaMemcpyAsync(p.d_oe, oe + p.olo, no * sizeof(double), cudaMemcpyHostToDevice, dev.st));
        CU(cudaMemcpyAsync(p.d_on, on + p.olo, no * sizeof(double), cudaMemcpyHostToDevice, dev.st));
        CU(cudaMemcpyAsync(p.d_ou, ou + p.olo, no * sizeof(double), cudaMemcpyHostToDevice, dev.st));
        for (const HostArray& a : arrays) {
            const int64_t rows = a.sharded ? ns : a.rows;
            double* dptr = nullptr;
            if (a.host) {
                dptr = dev.take<double>((size_t)std::max<int64_t>(rows, 1) * a.width);
                const double* src = a.host + (a.sharded ? p.slo * a.width : 0);
                if (rows > 0)
                    CU(cudaMemcpyAsync(dptr, src, (size_t)rows * a.width * sizeof(double),
                                       cudaMemcpyHostToDevice, dev.st));
            }
            p.d_arr.push_back(dptr);
        }
        p.ws = dev.take<char>(p.ws_bytes);
        CU(cudaMemsetAsync(dev.d_flags, 0, sizeof(unsigned), dev.st));
        rc = launch(dev, p.d_oe, p.d_on, p.d_ou, no, p.d_arr, ns, /*raw=*/plan.by_sources, p.d_out,
                    p.ws, p.ws_bytes);
        return rc;
    };
    if (nd == 1) {
        rc = feed(0);
        if (rc) return rc;
    } else {
        std::vector<int> rcs(nd, HB200_OK);
        std::vector<std::string> errs(nd);
        std::vector<std::thread> threads;
        for (int d = 0; d < nd; d++)
            threads.emplace_back([&, d]() {
                rcs[d] = feed(d);
                if (rcs[d]) errs[d] = g_err;  // g_err is thread-local: hand the text back
            });
        for (std::thread& t : threads) t.join();
        for (int d = 0; d < nd; d++)
            if (rcs[d]) {
                g_err = errs[d];
                return rcs[d];
            }
    }
    // phase 2: collect
    unsigned all_flags = 0;
    if (!plan.by_sources) {
        for (int d = 0; d < nd; d++) {
            Dev& dev = g_devs[d];
            Part& p = parts[d];
            CU(cudaSetDevice(dev.id));
            const int64_t no = p.ohi - p.olo;
            for (int k = 0; k < nf; k++)
                CU(cudaMemcpyAsync(out + (int64_t)k * n_obs + p.olo, p.d_out + (int64_t)k * no,
                                   no * sizeof(double), cudaMemcpyDeviceToHost, dev.st));
        }
    } else {
        // reduce-sum of the per-device partial fields on device 0: peer copies
        // (NVLink when peer access is enabled) + fixed-order reduce kernel.
        Dev& d0 = g_devs[0];
        CU(cudaSetDevice(d0.id));
        double* gathered = d0.take<double>((size_t)nd * nf * n_obs);
        for (int d = 0; d < nd; d++) {
            Dev& dev = g_devs[d];
            CU(cudaSetDevice(dev.id));
            CU(cudaStreamSynchronize(dev.st));
        }
        CU(cudaSetDevice(d0.id));
        for (int d = 0; d < nd; d++)
            CU(cudaMemcpyPeerAsync(gathered + (size_t)d * nf * n_obs, d0.id, parts[d].d_out,
                                   g_devs[d].id, (size_t)nf * n_obs * sizeof(double), d0.st));
        const int64_t total = (int64_t)nf * n_obs;
        reduce_partials_kernel<<<(unsigned)((total + 255) / 256), 256, 0, d0.st>>>(
            gathered, nd, nf, n_obs, final_scales, parts[0].d_out);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(out, parts[0].d_out, total * sizeof(double), cudaMemcpyDeviceToHost,
                           d0.st));
    }
    for (int d = 0; d < nd; d++) {
        Dev& dev = g_devs[d];
        CU(cudaSetDevice(dev.id));
        unsigned f = 0;
        CU(cudaMemcpyAsync(&f, dev.d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, dev.st));
        CU(cudaStreamSynchronize(dev.st));
        all_flags |= f;
    }
    if (flags) *flags = all_flags;
    return HB200_OK;
}

// the reference's jitted loops raise ZeroDivisionError when an observation point coincides with
// a source; the entry points without a flags argument report it as an error code
int check_zero_division(Dev& dev)
{
    unsigned f = 0;
    CU(cudaMemcpyAsync(&f, dev.d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, dev.st));
    CU(cudaStreamSynchronize(dev.st));
    if (f & FLAG_ZERO_DIV)
        return fail(HB200_EZERODIV, "division by zero: an observation point coincides with a source");
    return HB200_OK;
}

size_t ws_prism(int64_t no, int64_t ns, int nf, int sms) { return prism_ws_bytes(no, ns, nf, sms); }
size_t ws_point(int64_t no, int64_t ns, int nf, int sms)
{
    (void)nf;
    return point_ws_bytes(no, ns, sms);
}
size_t ws_dipole(int64_t no, int64_t ns, int nf, int sms)
{
    (void)nf;
    return dipole_ws_bytes(no, ns, sms);
}

size_t ws_tesseroid(int64_t no, int64_t ns, int nf, int sms)
{
    (void)nf;
    return tesseroid_ws_bytes(no, ns, sms);
}

Scales gravity_scales_for_mask(uint32_t mask)
{
    Scales sc;
    int s = 0;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    for (int f = 0; f < 10 && s < 6; f++)
        if (mask >> f & 1u) sc.s[s++] = gravity_scale(f);
    return sc;
}

}  // namespace

#include "hb200_fit_host.cuh"

// =============================================================== C ABI
extern "C" {

int hb200_version(void) { return 100; }

const char* hb200_last_error(void) { return g_err.c_str(); }

int hb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
        if (major == 10) ok++;
    }
    return ok;
}

int hb200_set_variant(int variant)
{
    if (variant < 0 || variant > 2) return fail(HB200_EINVAL, "variant must be 0, 1 or 2");
    g_variant = variant;
    return HB200_OK;
}
int hb200_get_variant(void) { return g_variant; }
int hb200_set_tesseroid_variant(int variant)
{
    if (variant < 0 || variant > 9) return fail(HB200_EINVAL, "tesseroid variant must be 0 .. 9");
    g_tess_variant = variant;
    return HB200_OK;
}
int hb200_get_tesseroid_variant(void) { return g_tess_variant; }
int hb200_set_tile_mode(int mode)
{
    if (mode < 0 || mode > 2)
        return fail(HB200_EINVAL, "tile mode must be 0 (per CTA), 1 (per warp; layers reuse vertices) or 2 (per warp, no reuse)");
    g_tile_mode = mode;
    return HB200_OK;
}
int hb200_get_tile_mode(void) { return g_tile_mode; }
int hb200_set_source_chunks(int chunks)
{
    if (chunks < 0 || chunks > 256) return fail(HB200_EINVAL, "source chunks must be 0 (auto) .. 256");
    g_source_chunks = chunks;
    return HB200_OK;
}
int hb200_get_source_chunks(void) { return g_source_chunks; }
int hb200_set_fit_rcond(double rcond)
{
    if (!(rcond >= 0.0 && rcond < 1.0)) return fail(HB200_EINVAL, "rcond must be in [0, 1)");
    g_fit_rcond = rcond;
    return HB200_OK;
}
double hb200_get_fit_rcond(void) { return g_fit_rcond; }
uint64_t hb200_launch_count(void) { return g_launches.load(); }

void hb200_shutdown(void)
{
    dense_release();
    for (Dev& d : g_devs) {
        cudaSetDevice(d.id);
        if (d.pool) cudaFree(d.pool);
        if (d.d_flags) cudaFree(d.d_flags);
        if (d.st) cudaStreamDestroy(d.st);
    }
    g_devs.clear();
}

int hb200_num_devices(void) { return (int)g_devs.size(); }

int hb200_init(const int* devices, int n_devices)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(HB200_ENODEV, "no CUDA device visible (%s); harmonica_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    std::vector<int> ids;
    if (devices && n_devices > 0) ids.assign(devices, devices + n_devices);
    else for (int d = 0; d < n; d++) ids.push_back(d);
    hb200_shutdown();
    for (int id : ids) {
        if (id < 0 || id >= n) return fail(HB200_EINVAL, "device %d out of range [0, %d)", id, n);
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, id);
        if (major != 10)
            return fail(HB200_ENODEV, "device %d is sm_%d*, this library is built for sm_100a only",
                        id, major * 10);
        Dev d;
        d.id = id;
        CU(cudaSetDevice(id));
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, id);
        CU(cudaStreamCreateWithFlags(&d.st, cudaStreamNonBlocking));
        CU(cudaMalloc(&d.d_flags, sizeof(unsigned)));
        g_devs.push_back(d);
    }
    // peer access towards device 0 (source-sharded reduce)
    for (size_t a = 1; a < g_devs.size(); a++) {
        int can = 0;
        if (g_devs[a].id == g_devs[0].id) continue;  // several shards on one device (tests)
        if (cudaDeviceCanAccessPeer(&can, g_devs[0].id, g_devs[a].id) != cudaSuccess) cudaGetLastError();
        if (can) {
            cudaSetDevice(g_devs[0].id);
            if (cudaDeviceEnablePeerAccess(g_devs[a].id, 0) != cudaSuccess) cudaGetLastError();
        }
    }
    return HB200_OK;
}

int hb200_prism_gravity(const double* easting, const double* northing, const double* upward,
                        int64_t n_obs, const double* prisms, const double* density,
                        int64_t n_prisms, uint32_t field_mask, int shard_mode, double* out,
                        uint32_t* flags)
{
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    if (n_obs < 0 || n_prisms < 0) return fail(HB200_EINVAL, "negative size");
    const int nf = popcount(field_mask);
    if (nf > 6) return fail(HB200_EINVAL, "at most 6 fields per call");
    std::vector<HostArray> arrays = {{prisms, n_prisms, 6, true}, {density, n_prisms, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return prism_gravity_dev_impl(oe, on, ou, no, arr[0], arr[1], ns, field_mask, raw, d_out,
                                      dev.d_flags, ws, wsb, dev.sms, dev.st);
    };
    return run_host_job(easting, northing, upward, n_obs, arrays, n_prisms, nf, shard_mode, true,
                        gravity_scales_for_mask(field_mask), out, flags, launch, ws_prism);
}

int hb200_prism_singular_scan(const double* easting, const double* northing,
                              const double* upward, int64_t n_obs, const double* prisms,
                              int64_t n_prisms, int field, uint32_t* flags)
{
    if (!flags) return fail(HB200_EINVAL, "flags must not be NULL");
    *flags = 0;
    if (n_obs <= 0 || n_prisms <= 0) return HB200_OK;
    std::vector<HostArray> arrays = {{prisms, n_prisms, 6, false}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        (void)raw; (void)d_out; (void)ns;
        Ws w(ws, wsb);
        double* packed = w.take((size_t)n_prisms * kPrismStride * sizeof(double));
        if (!packed) return fail(HB200_EINVAL, "workspace too small");
        pack_prisms_kernel<<<(unsigned)((n_prisms + 255) / 256), 256, 0, dev.st>>>(arr[0], nullptr,
                                                                                  n_prisms, packed);
        int64_t chunk_len;
        int chunks = choose_chunks(no, n_prisms, kBlock, dev.sms, &chunk_len);
        PrismArgs a;
        a.oe = oe; a.on = on; a.ou = ou; a.n_obs = no; a.packed = packed; a.n_src = n_prisms;
        a.chunk_len = chunk_len; a.out = nullptr; a.rules = 0; a.flags = dev.d_flags;
        dim3 grid((unsigned)((no + kBlock - 1) / kBlock), (unsigned)chunks);
        singular_scan_kernel<<<grid, kBlock, 0, dev.st>>>(a, field);
        CU(cudaGetLastError());
        g_launches += 2;
        return HB200_OK;
    };
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    std::vector<double> dummy((size_t)n_obs);
    return run_host_job(easting, northing, upward, n_obs, arrays, n_prisms, 1,
                        HB200_SHARD_OBSERVERS, false, sc, dummy.data(), flags, launch, ws_prism);
}

int hb200_prism_magnetic(const double* easting, const double* northing, const double* upward,
                         int64_t n_obs, const double* prisms, const double* mag_e,
                         const double* mag_n, const double* mag_u, int64_t n_prisms,
                         uint32_t component_mask, uint32_t rules, int shard_mode, double* out,
                         uint32_t* flags)
{
    if (!component_mask || (component_mask >> 3))
        return fail(HB200_EINVAL, "bad component_mask 0x%x", component_mask);
    const int nf = popcount(component_mask);
    std::vector<HostArray> arrays = {{prisms, n_prisms, 6, true}, {mag_e, n_prisms, 1, true},
                                     {mag_n, n_prisms, 1, true}, {mag_u, n_prisms, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return prism_magnetic_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], arr[3], ns,
                                       component_mask, rules, raw, d_out, dev.d_flags, ws, wsb,
                                       dev.sms, dev.st);
    };
    const double mu0 = 4 * kPi * 1e-7;
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = mu0 / 4 / kPi * 1e9;
    return run_host_job(easting, northing, upward, n_obs, arrays, n_prisms, nf, shard_mode, true,
                        sc, out, flags, launch, ws_prism);
}

int hb200_prism_layer_gravity(const double* easting, const double* northing,
                              const double* upward, int64_t n_obs, const double* prisms_easting,
                              int64_t n_east, const double* prisms_northing, int64_t n_north,
                              const double* bottom, const double* top, const double* density,
                              double thickness_threshold, uint32_t field_mask, int shard_mode,
                              double* out, uint32_t* flags)
{
    (void)shard_mode;  // the layer is replicated; observers are sharded
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    if (n_east < 2 || n_north < 2)
        return fail(HB200_EINVAL, "a prism layer needs at least 2 coordinates per axis");
    const int nf = popcount(field_mask);
    if (nf > 6) return fail(HB200_EINVAL, "at most 6 fields per call");
    const int64_t n_src = n_east * n_north;
    std::vector<HostArray> arrays = {{prisms_easting, n_east, 1, false},
                                     {prisms_northing, n_north, 1, false},
                                     {bottom, n_src, 1, false},
                                     {top, n_src, 1, false},
                                     {density, n_src, 1, false}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        (void)ns;
        return prism_layer_dev_impl(oe, on, ou, no, arr[0], n_east, arr[1], n_north, arr[2], arr[3],
                                    arr[4], thickness_threshold, field_mask, raw, d_out, dev.d_flags,
                                    ws, wsb, dev.sms, dev.st);
    };
    return run_host_job(easting, northing, upward, n_obs, arrays, n_src, nf,
                        HB200_SHARD_OBSERVERS, false, gravity_scales_for_mask(field_mask), out,
                        flags, launch, ws_prism);
}

int hb200_point_gravity(const double* easting, const double* northing, const double* upward,
                        int64_t n_obs, const double* src_easting, const double* src_northing,
                        const double* src_upward, const double* masses, int64_t n_src,
                        uint32_t field_mask, int spherical, int shard_mode, double* out,
                        uint32_t* flags)
{
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    if (spherical && (field_mask & ~((1u << F_POT) | (1u << F_U))))
        return fail(HB200_EINVAL, "spherical point masses: only potential and g_z exist");
    const int nf = popcount(field_mask);
    if (nf > 6) return fail(HB200_EINVAL, "at most 6 fields per call");
    std::vector<HostArray> arrays = {{src_easting, n_src, 1, true}, {src_northing, n_src, 1, true},
                                     {src_upward, n_src, 1, true}, {masses, n_src, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return point_gravity_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], arr[3], ns, field_mask,
                                      spherical, 1, raw, d_out, dev.d_flags, ws, wsb, dev.sms,
                                      dev.st);
    };
    return run_host_job(easting, northing, upward, n_obs, arrays, n_src, nf, shard_mode, true,
                        gravity_scales_for_mask(field_mask), out, flags, launch, ws_point);
}

int hb200_eqs_predict(const double* easting, const double* northing, const double* upward,
                      int64_t n_obs, const double* src_easting, const double* src_northing,
                      const double* src_upward, const double* coefs, int64_t n_src,
                      int shard_mode, double* out, uint32_t* flags)
{
    std::vector<HostArray> arrays = {{src_easting, n_src, 1, true}, {src_northing, n_src, 1, true},
                                     {src_upward, n_src, 1, true}, {coefs, n_src, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return point_gravity_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], arr[3], ns,
                                      1u << F_POT, 0, 0, raw, d_out, dev.d_flags, ws, wsb, dev.sms,
                                      dev.st);
    };
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    return run_host_job(easting, northing, upward, n_obs, arrays, n_src, 1, shard_mode, true, sc,
                        out, flags, launch, ws_point);
}

int hb200_eqs_predict_spherical(const double* longitude, const double* latitude,
                                const double* radius, int64_t n_obs, const double* src_longitude,
                                const double* src_latitude, const double* src_radius,
                                const double* coefs, int64_t n_src, int shard_mode, double* out,
                                uint32_t* flags)
{
    std::vector<HostArray> arrays = {{src_longitude, n_src, 1, true}, {src_latitude, n_src, 1, true},
                                     {src_radius, n_src, 1, true}, {coefs, n_src, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return point_gravity_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], arr[3], ns,
                                      1u << F_POT, 1, 0, raw, d_out, dev.d_flags, ws, wsb, dev.sms,
                                      dev.st);
    };
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    return run_host_job(longitude, latitude, radius, n_obs, arrays, n_src, 1, shard_mode, true, sc,
                        out, flags, launch, ws_point);
}

int hb200_dipole_magnetic(const double* easting, const double* northing, const double* upward,
                          int64_t n_obs, const double* src_easting, const double* src_northing,
                          const double* src_upward, const double* moment_e, const double* moment_n,
                          const double* moment_u, int64_t n_src, uint32_t component_mask,
                          int shard_mode, double* out, uint32_t* flags)
{
    if (!component_mask || (component_mask >> 3))
        return fail(HB200_EINVAL, "bad component_mask 0x%x", component_mask);
    const int nf = popcount(component_mask);
    std::vector<HostArray> arrays = {{src_easting, n_src, 1, true}, {src_northing, n_src, 1, true},
                                     {src_upward, n_src, 1, true}, {moment_e, n_src, 1, true},
                                     {moment_n, n_src, 1, true}, {moment_u, n_src, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return dipole_magnetic_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], arr[3], arr[4],
                                        arr[5], ns, component_mask, raw, d_out, dev.d_flags, ws, wsb,
                                        dev.sms, dev.st);
    };
    const double mu0 = 4 * kPi * 1e-7;
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = mu0 / 4 / kPi * 1e9;
    return run_host_job(easting, northing, upward, n_obs, arrays, n_src, nf, shard_mode, true, sc,
                        out, flags, launch, ws_dipole);
}

static int tesseroid_gravity_host(const double* longitude, const double* latitude,
                                  const double* radius, int64_t n_obs, const double* tesseroids,
                                  const double* density0, const double* density1,
                                  int64_t n_tesseroids, int field, int radial, int shard_mode,
                                  double* out, uint32_t* flags)
{
    if (field != HB200_POTENTIAL && field != HB200_G_Z)
        return fail(HB200_EINVAL, "tesseroid_gravity computes the potential or g_z, not field %d", field);
    if (n_obs < 0 || n_tesseroids < 0) return fail(HB200_EINVAL, "negative size");
    std::vector<HostArray> arrays = {{tesseroids, n_tesseroids, 6, true},
                                     {density0, n_tesseroids, 1, true},
                                     {density1, n_tesseroids, 1, true}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        return tesseroid_dev_impl(oe, on, ou, no, arr[0], arr[1], arr[2], ns, field, radial, raw,
                                  d_out, dev.d_flags, ws, wsb, dev.sms, dev.st);
    };
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    sc.s[0] = field == HB200_G_Z ? -1e5 : 1.0;
    return run_host_job(longitude, latitude, radius, n_obs, arrays, n_tesseroids, 1, shard_mode, true,
                        sc, out, flags, launch, ws_tesseroid);
}

int hb200_tesseroid_gravity(const double* longitude, const double* latitude, const double* radius,
                            int64_t n_obs, const double* tesseroids, const double* density,
                            int64_t n_tesseroids, int field, int radial_adaptive_discretization,
                            int shard_mode, double* out, uint32_t* flags)
{
    return tesseroid_gravity_host(longitude, latitude, radius, n_obs, tesseroids, density, density,
                                  n_tesseroids, field, radial_adaptive_discretization, shard_mode,
                                  out, flags);
}

int hb200_tesseroid_gravity_variable_density(const double* longitude, const double* latitude,
                                             const double* radius, int64_t n_obs,
                                             const double* tesseroids, const double* density_lower,
                                             const double* density_upper, int64_t n_tesseroids,
                                             int field, int shard_mode, double* out,
                                             uint32_t* flags)
{
    return tesseroid_gravity_host(longitude, latitude, radius, n_obs, tesseroids, density_lower,
                                  density_upper, n_tesseroids, field, 0, shard_mode, out, flags);
}

// density function + radial discretisation: root pass, leaf collection, callback, leaf quadrature
// (hb200_tess_leaves.cuh), in batches of computation points; device 0 of the selection
int hb200_tesseroid_gravity_density_function(const double* longitude, const double* latitude,
                                             const double* radius, int64_t n_obs,
                                             const double* tesseroids, const double* density_lower,
                                             const double* density_upper, int64_t n_tess, int field,
                                             hb200_density_fn density, void* user, double* out,
                                             uint32_t* flags)
{
    if (field != F_POT && field != F_U) return fail(HB200_EINVAL, "tesseroids: potential or g_z only");
    if (!density) return fail(HB200_EINVAL, "density function must not be NULL");
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = lazy_init();
    if (rc) return rc;
    if (flags) *flags = 0;
    if (n_obs <= 0) return HB200_OK;
    if (n_tess <= 0) {
        std::fill(out, out + n_obs, 0.0);
        return HB200_OK;
    }
    Dev& dev = g_devs[0];
    CU(cudaSetDevice(dev.id));
    cudaStream_t st = dev.st;
    int64_t chunk_len = 0;
    const int chunks = tess_two_kernel_chunks(n_tess, &chunk_len);
    if (chunk_len > 65535) return fail(HB200_EINVAL, "too many tesseroids for the density-function path");
    int64_t batch = std::min<int64_t>(n_obs, 8192);
    int cap = (int)std::min<int64_t>(1 << 24, std::max<int64_t>(1 << 16, batch * 4096));
    if (const char* env = std::getenv("HB200_LEAF_CAP"))  // tests: provoke the smaller-batch retry
        cap = std::max(64, std::atoi(env));
    const size_t need = 4 * align_up(n_obs * 8) + align_up((size_t)n_tess * 6 * 8) + 2 * align_up(n_tess * 8)
                      + align_up((size_t)n_tess * kTessRec * 8) + align_up((size_t)(chunks + 1) * batch * 8)
                      + align_up((size_t)chunks * kTessListCap * batch * sizeof(unsigned short))
                      + align_up((size_t)2 * chunks * batch * sizeof(int))
                      + align_up((size_t)cap * sizeof(int)) + align_up((size_t)6 * cap * 8)
                      + 2 * align_up((size_t)2 * cap * 8) + align_up(4 * sizeof(int));
    rc = dev.ensure(need + 4096);
    if (rc) return rc;
    double* d_lon = dev.take<double>(n_obs);
    double* d_lat = dev.take<double>(n_obs);
    double* d_rad = dev.take<double>(n_obs);
    double* d_out = dev.take<double>(n_obs);
    double* d_tess = dev.take<double>((size_t)n_tess * 6);
    double* d_rho0 = dev.take<double>(n_tess);
    double* d_rho1 = dev.take<double>(n_tess);
    double* packed = dev.take<double>((size_t)n_tess * kTessRec);
    double* parts = dev.take<double>((size_t)(chunks + 1) * batch);
    unsigned short* list = dev.take<unsigned short>((size_t)chunks * kTessListCap * batch);
    int* count = dev.take<int>((size_t)2 * chunks * batch);
    LeafBuf L;
    L.obs = dev.take<int>(cap);
    L.bounds = dev.take<double>((size_t)6 * cap);
    L.radii = dev.take<double>((size_t)2 * cap);
    double* d_rho = dev.take<double>((size_t)2 * cap);
    L.count = dev.take<int>(4);
    L.cap = cap;
    CU(cudaMemcpyAsync(d_lon, longitude, n_obs * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_lat, latitude, n_obs * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rad, radius, n_obs * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_tess, tesseroids, (size_t)n_tess * 6 * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rho0, density_lower, n_tess * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_rho1, density_upper, n_tess * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(dev.d_flags, 0, sizeof(unsigned), st));
    const double ratio = field == F_POT ? 1.0 : 2.5;  // tesseroid_gravity.py:33
    pack_tesseroid_fast_records_kernel<<<(unsigned)((n_tess + 127) / 128), 128, 0, st>>>(
        d_tess, d_rho0, d_rho1, n_tess, ratio, 1, packed);
    CU(cudaGetLastError());
    g_launches += 1;
    // tesseroid_gravity.py:222-225: g_z is the downward component in mGal
    Scales sc;
    sc.s[0] = field == F_U ? -1e5 : 1.0;
    std::vector<double> h_radii, h_rho;
    for (int64_t o0 = 0; o0 < n_obs;) {
        const int64_t nb = std::min(batch, n_obs - o0);
        TessArgs a;
        a.lon = d_lon + o0; a.lat = d_lat + o0; a.rad = d_rad + o0; a.n_obs = nb;
        a.packed = packed; a.n_src = n_tess; a.chunk_len = chunk_len;
        a.out = parts; a.scale = sc.s[0]; a.ratio = ratio; a.radial = 1; a.flags = dev.d_flags;
        dim3 grid_r((unsigned)((nb + kTessRootBlock - 1) / kTessRootBlock), (unsigned)chunks);
        dim3 grid_w((unsigned)((nb + kTessBlock - 1) / kTessBlock), (unsigned)chunks);
        CU(cudaMemsetAsync(L.count, 0, 4 * sizeof(int), st));
        if (field == F_POT) tesseroid_root_kernel<F_POT, 4><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count);
        else tesseroid_root_kernel<F_U, 4><<<grid_r, kTessRootBlock, 0, st>>>(a, list, count);
        tesseroid_collect_kernel<OwnTrig><<<grid_w, kTessBlock, 0, st>>>(a, list, count, L);
        CU(cudaGetLastError());
        g_launches += 2;
        int n_leaves = 0;
        CU(cudaMemcpyAsync(&n_leaves, L.count, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (n_leaves > cap) {  // more leaves than the buffer holds: fewer computation points per batch
            if (nb <= 1) return fail(HB200_ENOMEM, "one computation point has more than %d leaves", cap);
            batch = std::max<int64_t>(1, nb / 2);
            continue;
        }
        double* leaf_sum = parts + (size_t)chunks * nb;
        CU(cudaMemsetAsync(leaf_sum, 0, nb * 8, st));
        if (n_leaves > 0) {
            h_radii.resize((size_t)2 * n_leaves);
            h_rho.resize((size_t)2 * n_leaves);
            for (int k = 0; k < 2; k++)
                CU(cudaMemcpyAsync(h_radii.data() + (size_t)k * n_leaves, L.radii + (size_t)k * cap,
                                   (size_t)n_leaves * 8, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            density(h_radii.data(), h_rho.data(), (int64_t)2 * n_leaves, user);
            for (int k = 0; k < 2; k++)
                CU(cudaMemcpyAsync(d_rho + (size_t)k * cap, h_rho.data() + (size_t)k * n_leaves,
                                   (size_t)n_leaves * 8, cudaMemcpyHostToDevice, st));
            const unsigned blocks = (unsigned)((n_leaves + 127) / 128);
            if (field == F_POT) tesseroid_leaf_kernel<F_POT, OwnTrig><<<blocks, 128, 0, st>>>(a, L, n_leaves, d_rho, leaf_sum);
            else tesseroid_leaf_kernel<F_U, OwnTrig><<<blocks, 128, 0, st>>>(a, L, n_leaves, d_rho, leaf_sum);
            CU(cudaGetLastError());
            g_launches += 1;
        }
        reduce_partials_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, st>>>(parts, chunks + 1, 1, nb, sc,
                                                                             d_out + o0);
        CU(cudaGetLastError());
        g_launches += 1;
        CU(cudaStreamSynchronize(st));  // h_rho is reused by the next batch
        o0 += nb;
    }
    CU(cudaMemcpyAsync(out, d_out, n_obs * 8, cudaMemcpyDeviceToHost, st));
    unsigned f = 0;
    CU(cudaMemcpyAsync(&f, dev.d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (flags) *flags = f;
    return HB200_OK;
}

int hb200_tesseroid_inside_scan(const double* longitude, const double* latitude,
                                const double* radius, int64_t n_obs, const double* tesseroids,
                                int64_t n_tesseroids, uint32_t* flags)
{
    if (!flags) return fail(HB200_EINVAL, "flags must not be NULL");
    *flags = 0;
    if (n_obs <= 0 || n_tesseroids <= 0) return HB200_OK;
    std::vector<HostArray> arrays = {{tesseroids, n_tesseroids, 6, false}};
    auto launch = [=](Dev& dev, const double* oe, const double* on, const double* ou, int64_t no,
                      std::vector<double*>& arr, int64_t ns, bool raw, double* d_out, void* ws,
                      size_t wsb) {
        (void)raw; (void)d_out; (void)ns;
        Ws w(ws, wsb);
        double* packed = w.take((size_t)n_tesseroids * kTessStride * sizeof(double));
        if (!packed) return fail(HB200_EINVAL, "workspace too small");
        // the density slot of the record is not read by the scan
        pack_tesseroids_kernel<<<(unsigned)((n_tesseroids + 255) / 256), 256, 0, dev.st>>>(
            arr[0], arr[0], arr[0], n_tesseroids, packed);
        int64_t chunk_len;
        int chunks = choose_chunks(no, n_tesseroids, 128, dev.sms, &chunk_len);
        TessArgs a;
        a.lon = oe; a.lat = on; a.rad = ou; a.n_obs = no; a.packed = packed;
        a.n_src = n_tesseroids; a.chunk_len = chunk_len; a.out = nullptr; a.scale = 1.0;
        a.ratio = 0.0; a.radial = 0; a.flags = dev.d_flags;
        dim3 grid((unsigned)((no + 127) / 128), (unsigned)chunks);
        tesseroid_inside_scan_kernel<<<grid, 128, 0, dev.st>>>(a);
        CU(cudaGetLastError());
        g_launches += 2;
        return HB200_OK;
    };
    Scales sc;
    for (int c = 0; c < 6; c++) sc.s[c] = 1.0;
    std::vector<double> dummy((size_t)n_obs);
    return run_host_job(longitude, latitude, radius, n_obs, arrays, n_tesseroids, 1,
                        HB200_SHARD_OBSERVERS, false, sc, dummy.data(), flags, launch, ws_tesseroid);
}

static int eqs_jacobian_host(int spherical, const double* easting, const double* northing,
                             const double* upward, int64_t n_obs, const double* src_easting,
                             const double* src_northing, const double* src_upward, int64_t n_src,
                             double* jac)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = lazy_init();
    if (rc) return rc;
    if (n_obs <= 0 || n_src <= 0) return HB200_OK;
    Dev& dev = g_devs[0];
    CU(cudaSetDevice(dev.id));
    // rows are produced in slabs so that the device buffer stays bounded
    const int64_t slab = std::max<int64_t>(16, std::min<int64_t>(n_obs, (int64_t)(1 << 28) / n_src));
    size_t need = 3 * align_up(n_obs * 8) + 3 * align_up(n_src * 8) + align_up((size_t)slab * n_src * 8);
    rc = dev.ensure(need + 4096);
    if (rc) return rc;
    double* d_o[3];
    double* d_p[3];
    const double* ho[3] = {easting, northing, upward};
    const double* hp[3] = {src_easting, src_northing, src_upward};
    for (int c = 0; c < 3; c++) {
        d_o[c] = dev.take<double>(n_obs);
        d_p[c] = dev.take<double>(n_src);
        CU(cudaMemcpyAsync(d_o[c], ho[c], n_obs * 8, cudaMemcpyHostToDevice, dev.st));
        CU(cudaMemcpyAsync(d_p[c], hp[c], n_src * 8, cudaMemcpyHostToDevice, dev.st));
    }
    double* d_jac = dev.take<double>((size_t)slab * n_src);
    CU(cudaMemsetAsync(dev.d_flags, 0, sizeof(unsigned), dev.st));
    for (int64_t i0 = 0; i0 < n_obs; i0 += slab) {
        const int64_t rows = std::min(slab, n_obs - i0);
        const double* obs[3] = {d_o[0] + i0, d_o[1] + i0, d_o[2] + i0};
        rc = build_jacobian_dev(dev, spherical, obs, rows, d_p, n_src, d_jac);
        if (rc) return rc;
        CU(cudaMemcpyAsync(jac + i0 * n_src, d_jac, (size_t)rows * n_src * 8, cudaMemcpyDeviceToHost,
                           dev.st));
        CU(cudaStreamSynchronize(dev.st));
    }
    return check_zero_division(dev);
}

int hb200_eqs_jacobian(const double* easting, const double* northing, const double* upward,
                       int64_t n_obs, const double* src_easting, const double* src_northing,
                       const double* src_upward, int64_t n_src, double* jac)
{
    return eqs_jacobian_host(0, easting, northing, upward, n_obs, src_easting, src_northing,
                             src_upward, n_src, jac);
}

int hb200_eqs_jacobian_spherical(const double* longitude, const double* latitude,
                                 const double* radius, int64_t n_obs, const double* src_longitude,
                                 const double* src_latitude, const double* src_radius,
                                 int64_t n_src, double* jac)
{
    return eqs_jacobian_host(1, longitude, latitude, radius, n_obs, src_longitude, src_latitude,
                             src_radius, n_src, jac);
}

int hb200_eqs_fit(const double* easting, const double* northing, const double* upward,
                  int64_t n_obs, const double* src_easting, const double* src_northing,
                  const double* src_upward, int64_t n_src, const double* data,
                  const double* weights, double damping, int spherical, double* coefs,
                  int* solver_path)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = lazy_init();
    if (rc) return rc;
    if (n_obs <= 0 || n_src <= 0) return fail(HB200_EINVAL, "the fit needs data points and sources");
    Dev& dev = g_devs[0];
    CU(cudaSetDevice(dev.id));
    const size_t need = 5 * align_up(n_obs * 8) + 4 * align_up(n_src * 8)
                      + align_up((size_t)n_obs * n_src * 8);
    rc = dev.ensure(need + 4096);
    if (rc) return rc;
    double* d_o[3];
    double* d_p[3];
    const double* ho[3] = {easting, northing, upward};
    const double* hp[3] = {src_easting, src_northing, src_upward};
    for (int c = 0; c < 3; c++) {
        d_o[c] = dev.take<double>(n_obs);
        d_p[c] = dev.take<double>(n_src);
        CU(cudaMemcpyAsync(d_o[c], ho[c], n_obs * 8, cudaMemcpyHostToDevice, dev.st));
        CU(cudaMemcpyAsync(d_p[c], hp[c], n_src * 8, cudaMemcpyHostToDevice, dev.st));
    }
    double* d_data = dev.take<double>(n_obs);
    double* d_w = dev.take<double>(n_obs);
    double* d_coef = dev.take<double>(n_src);
    double* d_jac = dev.take<double>((size_t)n_obs * n_src);
    CU(cudaMemcpyAsync(d_data, data, n_obs * 8, cudaMemcpyHostToDevice, dev.st));
    if (weights) CU(cudaMemcpyAsync(d_w, weights, n_obs * 8, cudaMemcpyHostToDevice, dev.st));
    CU(cudaMemsetAsync(dev.d_flags, 0, sizeof(unsigned), dev.st));
    rc = build_jacobian_dev(dev, spherical, d_o, n_obs, d_p, n_src, d_jac);
    if (rc) return rc;
    if ((rc = check_zero_division(dev))) return rc;
    const bool damped = !std::isnan(damping);
    rc = dense_least_squares(dev, d_jac, n_obs, n_src, d_data, weights ? d_w : nullptr, damped,
                             damping, d_coef, solver_path);
    if (rc) return rc;
    CU(cudaMemcpyAsync(coefs, d_coef, n_src * 8, cudaMemcpyDeviceToHost, dev.st));
    CU(cudaStreamSynchronize(dev.st));
    return HB200_OK;
}

int hb200_eqs_fit_gb(const double* easting, const double* northing, const double* upward,
                     int64_t n_obs, const double* src_easting, const double* src_northing,
                     const double* src_upward, int64_t n_src, const double* data,
                     const double* weights, double damping, int spherical, int64_t n_windows,
                     const int64_t* src_index, const int64_t* src_offset,
                     const int64_t* data_index, const int64_t* data_offset, double* coefs,
                     double* rmse)
{
    std::lock_guard<std::mutex> lock(g_mu);
    int rc = lazy_init();
    if (rc) return rc;
    if (n_obs <= 0 || n_src <= 0 || n_windows < 0)
        return fail(HB200_EINVAL, "the fit needs data points and sources");
    int64_t max_nd = 0, max_np = 0;
    size_t max_jac = 0;
    for (int64_t w = 0; w < n_windows; w++) {
        const int64_t np = src_offset[w + 1] - src_offset[w], nd = data_offset[w + 1] - data_offset[w];
        if (np <= 0 || nd <= 0) return fail(HB200_EINVAL, "window %lld is empty", (long long)w);
        max_np = std::max(max_np, np);
        max_nd = std::max(max_nd, nd);
        max_jac = std::max(max_jac, (size_t)np * (size_t)nd);
    }
    const int64_t n_si = src_offset[n_windows], n_di = data_offset[n_windows];
    // the gather / scatter kernels trust the index lists: check them here, on the host
    for (int64_t k = 0; k < n_si; k++)
        if (src_index[k] < 0 || src_index[k] >= n_src)
            return fail(HB200_EINVAL, "source index %lld out of range", (long long)src_index[k]);
    for (int64_t k = 0; k < n_di; k++)
        if (data_index[k] < 0 || data_index[k] >= n_obs)
            return fail(HB200_EINVAL, "data index %lld out of range", (long long)data_index[k]);
    Dev& dev = g_devs[0];
    CU(cudaSetDevice(dev.id));
    const int64_t n_blocks = (n_obs + 255) / 256;
    const size_t ws_bytes = point_ws_bytes(n_obs, std::max<int64_t>(max_np, 1), dev.sms);
    const size_t need = 7 * align_up(n_obs * 8) + 4 * align_up(n_src * 8)
                      + align_up(n_si * 8) + align_up(n_di * 8) + 5 * align_up(max_nd * 8)
                      + 4 * align_up(max_np * 8) + align_up(max_jac * 8) + align_up(n_blocks * 8)
                      + align_up((n_windows + 1) * 8) + align_up(ws_bytes);
    rc = dev.ensure(need + 8192);
    if (rc) return rc;
    cudaStream_t st = dev.st;
    double* d_o[3];
    double* d_p[3];
    const double* ho[3] = {easting, northing, upward};
    const double* hp[3] = {src_easting, src_northing, src_upward};
    for (int c = 0; c < 3; c++) {
        d_o[c] = dev.take<double>(n_obs);
        d_p[c] = dev.take<double>(n_src);
        CU(cudaMemcpyAsync(d_o[c], ho[c], n_obs * 8, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_p[c], hp[c], n_src * 8, cudaMemcpyHostToDevice, st));
    }
    double* d_residue = dev.take<double>(n_obs);
    double* d_w = dev.take<double>(n_obs);
    double* d_pred = dev.take<double>(n_obs);
    double* d_coefs = dev.take<double>(n_src);
    int64_t* d_si = dev.take<int64_t>(std::max<int64_t>(n_si, 1));
    int64_t* d_di = dev.take<int64_t>(std::max<int64_t>(n_di, 1));
    double* c_o[3] = {dev.take<double>(max_nd), dev.take<double>(max_nd), dev.take<double>(max_nd)};
    double* c_res = dev.take<double>(max_nd);
    double* c_w = dev.take<double>(max_nd);
    double* c_p[3] = {dev.take<double>(max_np), dev.take<double>(max_np), dev.take<double>(max_np)};
    double* c_coef = dev.take<double>(max_np);
    double* d_jac = dev.take<double>(max_jac);
    double* d_bsum = dev.take<double>(n_blocks);
    double* d_rmse = dev.take<double>(n_windows + 1);
    void* d_ws = dev.take<char>(ws_bytes);
    CU(cudaMemcpyAsync(d_residue, data, n_obs * 8, cudaMemcpyHostToDevice, st));
    if (weights) CU(cudaMemcpyAsync(d_w, weights, n_obs * 8, cudaMemcpyHostToDevice, st));
    if (n_si) CU(cudaMemcpyAsync(d_si, src_index, n_si * 8, cudaMemcpyHostToDevice, st));
    if (n_di) CU(cudaMemcpyAsync(d_di, data_index, n_di * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(d_coefs, 0, n_src * 8, st));
    CU(cudaMemsetAsync(dev.d_flags, 0, sizeof(unsigned), st));
    // errors[0] = sqrt(mean(data^2)), gradient_boosted.py:253
    residue_update_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(d_residue, nullptr, n_obs, d_bsum);
    finish_rmse_kernel<<<1, 256, 0, st>>>(d_bsum, n_blocks, n_obs, d_rmse);
    CU(cudaGetLastError());
    g_launches += 2;
    const bool damped = !std::isnan(damping);
    for (int64_t w = 0; w < n_windows; w++) {
        const int64_t np = src_offset[w + 1] - src_offset[w], nd = data_offset[w + 1] - data_offset[w];
        const int64_t* si = d_si + src_offset[w];
        const int64_t* di = d_di + data_offset[w];
        // gradient_boosted.py:265-271: the sources, data points, residues and weights of the window
        Gather4 gp = {{d_p[0], d_p[1], d_p[2], nullptr}, {c_p[0], c_p[1], c_p[2], nullptr}};
        gather_kernel<<<blocks_for(np), 256, 0, st>>>(gp, si, np);
        Gather4 go = {{d_o[0], d_o[1], d_o[2], d_residue}, {c_o[0], c_o[1], c_o[2], c_res}};
        gather_kernel<<<blocks_for(nd), 256, 0, st>>>(go, di, nd);
        if (weights) {
            Gather4 gw = {{d_w, nullptr, nullptr, nullptr}, {c_w, nullptr, nullptr, nullptr}};
            gather_kernel<<<blocks_for(nd), 256, 0, st>>>(gw, di, nd);
            g_launches += 1;
        }
        CU(cudaGetLastError());
        g_launches += 2;
        // :273-280: Jacobian of the window and its least-squares coefficients
        rc = build_jacobian_dev(dev, spherical, c_o, nd, c_p, np, d_jac);
        if (rc) return rc;
        rc = dense_least_squares(dev, d_jac, nd, np, c_res, weights ? c_w : nullptr, damped, damping,
                                 c_coef, nullptr);
        if (rc) return rc;
        // :282-289: field of the window's sources on EVERY data point
        rc = point_gravity_dev_impl(d_o[0], d_o[1], d_o[2], n_obs, c_p[0], c_p[1], c_p[2], c_coef, np,
                                    1u << F_POT, spherical, 0, false, d_pred, dev.d_flags, d_ws,
                                    ws_bytes, dev.sms, st);
        if (rc) return rc;
        // :290-294: residue, RMSE history, coefficients
        residue_update_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(d_residue, d_pred, n_obs, d_bsum);
        finish_rmse_kernel<<<1, 256, 0, st>>>(d_bsum, n_blocks, n_obs, d_rmse + w + 1);
        scatter_add_kernel<<<blocks_for(np), 256, 0, st>>>(d_coefs, si, c_coef, np);
        CU(cudaGetLastError());
        g_launches += 3;
    }
    CU(cudaMemcpyAsync(coefs, d_coefs, n_src * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(rmse, d_rmse, (n_windows + 1) * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return check_zero_division(dev);  // a window's Jacobian or its prediction met a zero distance
}

// ---- device-buffer entry points
size_t hb200_prism_ws_bytes(int64_t n_obs, int64_t n_sources, int n_fields)
{
    return prism_ws_bytes(n_obs, n_sources, n_fields, 148);
}

size_t hb200_point_ws_bytes(int64_t n_obs, int64_t n_sources)
{
    return point_ws_bytes(n_obs, n_sources, 148);
}

size_t hb200_tesseroid_ws_bytes(int64_t n_obs, int64_t n_tesseroids)
{
    return tesseroid_ws_bytes(n_obs, n_tesseroids, sm_count_current());
}

int hb200_tesseroid_gravity_dev(const double* longitude, const double* latitude, const double* radius,
                                int64_t n_obs, const double* tesseroids, const double* density,
                                int64_t n_tesseroids, int field, int radial_adaptive_discretization,
                                double* out, uint32_t* flags_dev, void* ws, size_t ws_bytes,
                                void* stream)
{
    return tesseroid_dev_impl(longitude, latitude, radius, n_obs, tesseroids, density, density,
                              n_tesseroids, field, radial_adaptive_discretization, false, out,
                              flags_dev, ws, ws_bytes, sm_count_current(), (cudaStream_t)stream);
}

int hb200_prism_gravity_dev(const double* easting, const double* northing, const double* upward,
                            int64_t n_obs, const double* prisms, const double* density,
                            int64_t n_prisms, uint32_t field_mask, double* out,
                            uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream)
{
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    return prism_gravity_dev_impl(easting, northing, upward, n_obs, prisms, density, n_prisms,
                                  field_mask, false, out, flags_dev, ws, ws_bytes,
                                  sm_count_current(), (cudaStream_t)stream);
}

int hb200_prism_magnetic_dev(const double* easting, const double* northing,
                             const double* upward, int64_t n_obs, const double* prisms,
                             const double* mag_e, const double* mag_n, const double* mag_u,
                             int64_t n_prisms, uint32_t component_mask, uint32_t rules,
                             double* out, uint32_t* flags_dev, void* ws, size_t ws_bytes,
                             void* stream)
{
    if (!component_mask || (component_mask >> 3))
        return fail(HB200_EINVAL, "bad component_mask 0x%x", component_mask);
    return prism_magnetic_dev_impl(easting, northing, upward, n_obs, prisms, mag_e, mag_n, mag_u,
                                   n_prisms, component_mask, rules, false, out, flags_dev, ws,
                                   ws_bytes, sm_count_current(), (cudaStream_t)stream);
}

int hb200_prism_layer_gravity_dev(const double* easting, const double* northing,
                                  const double* upward, int64_t n_obs,
                                  const double* prisms_easting, int64_t n_east,
                                  const double* prisms_northing, int64_t n_north,
                                  const double* bottom, const double* top, const double* density,
                                  double thickness_threshold, uint32_t field_mask, double* out,
                                  uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream)
{
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    return prism_layer_dev_impl(easting, northing, upward, n_obs, prisms_easting, n_east,
                                prisms_northing, n_north, bottom, top, density, thickness_threshold,
                                field_mask, false, out, flags_dev, ws, ws_bytes, sm_count_current(),
                                (cudaStream_t)stream);
}

int hb200_point_gravity_dev(const double* easting, const double* northing, const double* upward,
                            int64_t n_obs, const double* src_easting, const double* src_northing,
                            const double* src_upward, const double* weights, int64_t n_src,
                            uint32_t field_mask, int spherical, int scale_by_G, double* out,
                            uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream)
{
    if (!field_mask || (field_mask >> 10)) return fail(HB200_EINVAL, "bad field_mask 0x%x", field_mask);
    return point_gravity_dev_impl(easting, northing, upward, n_obs, src_easting, src_northing,
                                  src_upward, weights, n_src, field_mask, spherical, scale_by_G,
                                  false, out, flags_dev, ws, ws_bytes, sm_count_current(),
                                  (cudaStream_t)stream);
}

int hb200_fp64_peak(int iters, double* flops, double* seconds)
{
    int dev = 0;
    CU(cudaGetDevice(&dev));
    const int sms = sm_count_current();
    const int blocks = sms * 8, threads = 256;
    double* d_out = nullptr;
    CU(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    fp64_peak_kernel<<<blocks, threads>>>(d_out, 16, 0.999999, 1e-9);  // warm-up
    CU(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-9);
    g_launches += 2;
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    CU(cudaGetLastError());
    const double n_fma = (double)blocks * threads * 8.0 * 16.0 * iters;
    if (seconds) *seconds = ms * 1e-3;
    if (flops) *flops = 2.0 * n_fma / (ms * 1e-3);
    cudaFree(d_out);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HB200_OK;
}

}  // extern "C"
