// hb200_fit_kernels.cuh -- device pieces of the equivalent-sources FIT that are not a
// library call: the spherical Jacobian, the column scaling verde applies before the solve,
// the singular-value filters, and the gather / scatter / residue steps of the gradient-boosted
// loop. The dense factorisations themselves are cuBLAS / cuSOLVER calls (hb200_api.cu).
//
// Reference: harmonica/_equivalent_sources/utils.py:54-74 (jacobian),
// cartesian.py:279-280 -> verde.base.least_squares (StandardScaler(with_mean=False) on the
// columns, then LinearRegression / Ridge without intercept), gradient_boosted.py:244-293.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hb {

// jac[i][j] = 1 / distance_spherical(obs_i, src_j), spherical.py:412-424. Both sides arrive
// as trig records (cos lon, sin lon, cos lat, sin lat, radius, -) written by
// pack_points_sph_kernel, so the inner loop has no transcendental but the square root.
__global__ void eqs_jacobian_sph_kernel(const double* __restrict__ obs, int64_t n_obs,
                                        const double* __restrict__ src, int64_t n_src,
                                        double* __restrict__ jac, unsigned* flags)
{
    // row blocks on grid.x (no 65535 limit), column blocks on grid.y
    const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= n_src) return;
    const double* q = src + j * 6;
    const double clam_p = q[0], slam_p = q[1], cphi_p = q[2], sphi_p = q[3], rad_p = q[4];
    const int64_t i0 = (int64_t)blockIdx.x * 16;
    bool zero = false;
#pragma unroll 4
    for (int r = 0; r < 16; r++) {
        const int64_t i = i0 + r;
        if (i >= n_obs) break;
        const double* o = obs + i * 6;
        const double coslambda = fma(clam_p, o[0], slam_p * o[1]);
        const double cospsi = sphi_p * o[3] + cphi_p * o[2] * coslambda;
        const double dr = o[4] - rad_p;
        const double d2 = dr * dr + 2 * o[4] * rad_p * (1 - cospsi);
        zero |= d2 == 0.0;  // the reference's jitted loop raises ZeroDivisionError here
        jac[i * n_src + j] = 1.0 / sqrt(d2);
    }
    if (zero && flags) atomicOr(flags, FLAG_ZERO_DIV);
}

// Column statistics of a row-major n x p matrix, as sklearn's StandardScaler(with_mean=False)
// computes them: scale_j = sqrt(population variance of column j); a column that is constant up
// to rounding (sklearn.preprocessing._data._is_constant_feature) gets scale 1.
// blockDim = (32, 8): 32 adjacent columns per CTA (coalesced rows), 8 row lanes reduced in smem.
__global__ void column_scale_kernel(const double* __restrict__ jac, int64_t n, int64_t p,
                                    double* __restrict__ scale)
{
    __shared__ double red[8][33];
    const int64_t j = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const bool live = j < p;
    double s = 0.0;
    if (live)
        for (int64_t i = threadIdx.y; i < n; i += 8) s += jac[i * p + j];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    double mean = 0.0;
    for (int r = 0; r < 8; r++) mean += red[r][threadIdx.x];
    mean /= (double)n;
    __syncthreads();
    double ss = 0.0;
    if (live)
        for (int64_t i = threadIdx.y; i < n; i += 8) {
            const double d = jac[i * p + j] - mean;
            ss += d * d;
        }
    red[threadIdx.y][threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.y == 0 && live) {
        double var = 0.0;
        for (int r = 0; r < 8; r++) var += red[r][threadIdx.x];
        var /= (double)n;
        const double eps = 2.220446049250313e-16;
        const double nn = (double)n;
        const double bound = nn * eps * var + (nn * mean * eps) * (nn * mean * eps);
        scale[j] = (var <= bound) ? 1.0 : sqrt(var);
    }
}

// jac[i][j] = jac[i][j] / scale[j] * sqrt(w[i]);  y[i] = data[i] * sqrt(w[i])
// (StandardScaler.transform, then sklearn's _rescale_data for sample_weight). w may be null.
__global__ void scale_system_kernel(double* __restrict__ jac, int64_t n, int64_t p,
                                    const double* __restrict__ scale,
                                    const double* __restrict__ w, const double* __restrict__ data,
                                    double* __restrict__ y)
{
    const int64_t j = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= p) return;
    const double sc = scale[j];
    const int64_t i0 = (int64_t)blockIdx.x * 16;
    for (int r = 0; r < 16; r++) {
        const int64_t i = i0 + r;
        if (i >= n) break;
        const double sw = w ? sqrt(w[i]) : 1.0;
        double v = jac[i * p + j] / sc;
        if (w) v *= sw;
        jac[i * p + j] = v;
        if (j == 0) y[i] = w ? data[i] * sw : data[i];
    }
}

__global__ void add_diagonal_kernel(double* __restrict__ a, int64_t k, double alpha)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) a[i * k + i] += alpha;
}

// out[0] = min, out[1] = max of the diagonal of a k x k matrix (one CTA): the Cholesky factor's
// pivots, whose squared ratio estimates the conditioning of the normal equations.
__global__ void diagonal_minmax_kernel(const double* __restrict__ a, int64_t k,
                                       double* __restrict__ out)
{
    __shared__ double lo[256], hi[256];
    double mn = INFINITY, mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < k; i += 256) {
        const double v = a[i * k + i];
        mn = fmin(mn, v);
        mx = fmax(mx, v);
    }
    lo[threadIdx.x] = mn;
    hi[threadIdx.x] = mx;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            lo[threadIdx.x] = fmin(lo[threadIdx.x], lo[threadIdx.x + s]);
            hi[threadIdx.x] = fmax(hi[threadIdx.x], hi[threadIdx.x + s]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[0] = lo[0];
        out[1] = hi[0];
    }
}

// t[k] *= filter(s[k]).  mode 0: truncated pseudo-inverse, 1/s above rcond * s[0], else 0
// (scipy.linalg.lstsq, cond = eps);  mode 1: ridge filter s / (s^2 + alpha) above 1e-15
// (sklearn.linear_model._ridge._solve_svd).
__global__ void singular_filter_kernel(double* __restrict__ t, const double* __restrict__ s,
                                       int64_t k, int mode, double param)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const double sv = s[i];
    double f;
    if (mode == 0) f = sv > param * s[0] ? 1.0 / sv : 0.0;
    else f = sv > 1e-15 ? sv / (sv * sv + param) : 0.0;
    t[i] *= f;
}

// coef[j] = x[j] / scale[j]  (verde.base.least_squares: regr.coef_ / scaler.scale_)
__global__ void unscale_kernel(const double* __restrict__ x, const double* __restrict__ scale,
                               int64_t p, double* __restrict__ coef)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < p) coef[j] = x[j] / scale[j];
}

// ---- gradient-boosted loop (gradient_boosted.py:262-292)
// dst_c[k] = src_c[idx[k]] for up to four arrays at once (null pairs are skipped)
struct Gather4 {
    const double* src[4];
    double* dst[4];
};

__global__ void gather_kernel(Gather4 g, const int64_t* __restrict__ idx, int64_t count)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const int64_t i = idx[k];
#pragma unroll
    for (int c = 0; c < 4; c++)
        if (g.src[c]) g.dst[c][k] = g.src[c][i];
}

// coefs[idx[k]] += chunk[k]; the indices of one window are distinct, so no atomics
__global__ void scatter_add_kernel(double* __restrict__ coefs, const int64_t* __restrict__ idx,
                                   const double* __restrict__ chunk, int64_t count)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) coefs[idx[k]] += chunk[k];
}

// residue -= predicted (predicted may be null), per-CTA sums of residue^2 in fixed order
__global__ void residue_update_kernel(double* __restrict__ residue,
                                      const double* __restrict__ predicted, int64_t n,
                                      double* __restrict__ block_sums)
{
    __shared__ double red[256];
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    double v = 0.0;
    if (i < n) {
        v = residue[i];
        if (predicted) {
            v -= predicted[i];
            residue[i] = v;
        }
    }
    red[threadIdx.x] = v * v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = red[0];
}

// rmse = sqrt(sum(block_sums) / n), one CTA, fixed order (deterministic)
__global__ void finish_rmse_kernel(const double* __restrict__ block_sums, int64_t n_blocks,
                                   int64_t n, double* __restrict__ rmse)
{
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t b = threadIdx.x; b < n_blocks; b += 256) s += block_sums[b];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 128; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) *rmse = sqrt(red[0] / (double)n);
}

}  // namespace hb
