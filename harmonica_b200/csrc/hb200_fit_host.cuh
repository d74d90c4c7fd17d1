// hb200_fit_host.cuh -- host side of the device-resident equivalent-sources FIT.
// Included by hb200_api.cu only (uses its Dev / fail / CU / point_gravity_dev_impl).
//
// What stays on the device: the Jacobian (utils.py:54-74), verde's column scaling, the
// weighted normal equations, the factorisation and the solve; for the gradient-boosted
// variant (gradient_boosted.py:244-293) also every window's gather, prediction over ALL data
// points and residue update. Only the coefficients (and the RMSE history) travel back.
//
// The factorisations are library calls: cuBLAS DSYRK/DGEMV/DGEAM and cuSOLVER DPOTRF/DPOTRS/
// DGESVD. Both libraries are bound with dlopen on first use, so the pair kernels (the product's
// hot path) keep a dependency-free .so and a fit fails loudly if the libraries are missing.
//
// Solver semantics follow verde.base.least_squares (the reference's call, cartesian.py:279-280):
//   damping given -> sklearn Ridge(alpha, fit_intercept=False), dense => "cholesky" solver:
//       (X'X + alpha I) c = X'y   (dual form (XX' + alpha I) when n_features > n_samples),
//       SVD ridge filter if the factorisation fails (sklearn's own fallback) or its pivots
//       show the normal equations to be singular to working precision;
//   damping None  -> sklearn LinearRegression => scipy.linalg.lstsq(X, y, cond = tol): the
//       minimum-norm solution from the SVD with singular values below rcond * s_max dropped.
//       rcond = g_fit_rcond (hb200_set_fit_rcond). Default: machine epsilon = cond=None of the
//       scikit-learn releases the reference's tests were written against (its undamped fits
//       must reproduce the data to 1e-3, test/test_eq_sources_spherical.py:26-67, which a
//       truncation at 1e-6 does not). scikit-learn >= 1.7 passes cond = tol = 1e-6: set
//       hb200_set_fit_rcond(1e-6) to reproduce THAT (tests/test_eqs_fit_host.py checks both).
#pragma once
#include <cublas_v2.h>
#include <cusolverDn.h>
#include <dlfcn.h>

#include <climits>
#include <cmath>

#include "hb200_fit_kernels.cuh"

namespace {

struct DenseLibs {
    void* h_blas = nullptr;
    void* h_solver = nullptr;
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t solver = nullptr;
    int device = -1;
    decltype(&cublasCreate_v2) blasCreate = nullptr;
    decltype(&cublasDestroy_v2) blasDestroy = nullptr;
    decltype(&cublasSetStream_v2) blasSetStream = nullptr;
    decltype(&cublasDsyrk_v2) dsyrk = nullptr;
    decltype(&cublasDgemv_v2) dgemv = nullptr;
    decltype(&cublasDgeam) dgeam = nullptr;
    decltype(&cusolverDnCreate) solCreate = nullptr;
    decltype(&cusolverDnDestroy) solDestroy = nullptr;
    decltype(&cusolverDnSetStream) solSetStream = nullptr;
    decltype(&cusolverDnDpotrf_bufferSize) potrfSize = nullptr;
    decltype(&cusolverDnDpotrf) potrf = nullptr;
    decltype(&cusolverDnDpotrs) potrs = nullptr;
    decltype(&cusolverDnDgesvd_bufferSize) gesvdSize = nullptr;
    decltype(&cusolverDnDgesvd) gesvd = nullptr;
};

DenseLibs g_dense;

void* open_first(const char* const* names)
{
    for (int i = 0; names[i]; i++)
        if (void* h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL)) return h;
    return nullptr;
}

template <typename F> bool bind(void* h, const char* name, F& f)
{
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

void dense_release()
{
    DenseLibs& L = g_dense;
    if (L.blas && L.blasDestroy) L.blasDestroy(L.blas);
    if (L.solver && L.solDestroy) L.solDestroy(L.solver);
    L.blas = nullptr;
    L.solver = nullptr;
    L.device = -1;
}

int dense_acquire(Dev& dev)
{
    DenseLibs& L = g_dense;
    if (!L.h_blas) {
        static const char* const blas_names[] = {"libcublas.so.12", "libcublas.so",
                                                 "/usr/local/cuda/lib64/libcublas.so.12", nullptr};
        static const char* const sol_names[] = {"libcusolver.so.11", "libcusolver.so",
                                                "/usr/local/cuda/lib64/libcusolver.so.11", nullptr};
        L.h_blas = open_first(blas_names);
        if (!L.h_blas) return fail(HB200_ECUDA, "cannot load cuBLAS for the EQS fit: %s", dlerror());
        L.h_solver = open_first(sol_names);
        if (!L.h_solver) {
            L.h_blas = nullptr;
            return fail(HB200_ECUDA, "cannot load cuSOLVER for the EQS fit: %s", dlerror());
        }
        bool ok = bind(L.h_blas, "cublasCreate_v2", L.blasCreate)
               && bind(L.h_blas, "cublasDestroy_v2", L.blasDestroy)
               && bind(L.h_blas, "cublasSetStream_v2", L.blasSetStream)
               && bind(L.h_blas, "cublasDsyrk_v2", L.dsyrk)
               && bind(L.h_blas, "cublasDgemv_v2", L.dgemv)
               && bind(L.h_blas, "cublasDgeam", L.dgeam)
               && bind(L.h_solver, "cusolverDnCreate", L.solCreate)
               && bind(L.h_solver, "cusolverDnDestroy", L.solDestroy)
               && bind(L.h_solver, "cusolverDnSetStream", L.solSetStream)
               && bind(L.h_solver, "cusolverDnDpotrf_bufferSize", L.potrfSize)
               && bind(L.h_solver, "cusolverDnDpotrf", L.potrf)
               && bind(L.h_solver, "cusolverDnDpotrs", L.potrs)
               && bind(L.h_solver, "cusolverDnDgesvd_bufferSize", L.gesvdSize)
               && bind(L.h_solver, "cusolverDnDgesvd", L.gesvd);
        if (!ok) {
            L.h_blas = L.h_solver = nullptr;
            return fail(HB200_ECUDA, "cuBLAS / cuSOLVER lack a symbol the EQS fit needs");
        }
    }
    if (L.device != dev.id) {
        dense_release();
        CU(cudaSetDevice(dev.id));
        if (L.blasCreate(&L.blas) != CUBLAS_STATUS_SUCCESS)
            return fail(HB200_ECUDA, "cublasCreate failed");
        if (L.solCreate(&L.solver) != CUSOLVER_STATUS_SUCCESS)
            return fail(HB200_ECUDA, "cusolverDnCreate failed");
        L.device = dev.id;
    }
    if (L.blasSetStream(L.blas, dev.st) != CUBLAS_STATUS_SUCCESS
        || L.solSetStream(L.solver, dev.st) != CUSOLVER_STATUS_SUCCESS)
        return fail(HB200_ECUDA, "cannot bind the dense-algebra handles to the device stream");
    return HB200_OK;
}

#define BLAS(expr)                                                                       \
    do {                                                                                 \
        cublasStatus_t s_ = (expr);                                                      \
        if (s_ != CUBLAS_STATUS_SUCCESS)                                                 \
            return fail(HB200_ECUDA, "%s failed: cuBLAS status %d", #expr, (int)s_);     \
    } while (0)
#define SOLVER(expr)                                                                     \
    do {                                                                                 \
        cusolverStatus_t s_ = (expr);                                                    \
        if (s_ != CUSOLVER_STATUS_SUCCESS)                                               \
            return fail(HB200_ECUDA, "%s failed: cuSOLVER status %d", #expr, (int)s_);   \
    } while (0)

// scratch that lives for one solve
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes)
    {
        if (p) cudaFree(p);
        p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 8);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(HB200_ENOMEM, "cudaMalloc(%zu bytes) for the EQS fit: %s", bytes,
                        cudaGetErrorString(e));
        }
        return HB200_OK;
    }
    double* d() const { return static_cast<double*>(p); }
};

unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

// Jacobian of `n` observers x `p` sources on the device, row-major. Spherical inputs go
// through trig records first (6 doubles each).
int build_jacobian_dev(Dev& dev, int spherical, const double* const obs[3], int64_t n,
                       const double* const src[3], int64_t p, double* jac)
{
    // a zero distance sets HB200_FLAG_ZERO_DIV in dev.d_flags (callers clear and read it)
    dim3 grid((unsigned)((n + 15) / 16), blocks_for(p));
    if (!spherical) {
        eqs_jacobian_kernel<<<grid, 256, 0, dev.st>>>(obs[0], obs[1], obs[2], n, src[0], src[1],
                                                     src[2], p, jac, dev.d_flags);
        CU(cudaGetLastError());
        g_launches += 1;
        return HB200_OK;
    }
    DevBuf rec;
    int rc = rec.alloc((size_t)(n + p) * kSphStride * sizeof(double));
    if (rc) return rc;
    double* orec = rec.d();
    double* srec = orec + n * kSphStride;
    // the weight slot of the record is unused here: any valid array of the right length does
    pack_points_sph_kernel<<<blocks_for(n), 256, 0, dev.st>>>(obs[0], obs[1], obs[2], obs[2], n, orec);
    pack_points_sph_kernel<<<blocks_for(p), 256, 0, dev.st>>>(src[0], src[1], src[2], src[2], p, srec);
    eqs_jacobian_sph_kernel<<<grid, 256, 0, dev.st>>>(orec, n, srec, p, jac, dev.d_flags);
    CU(cudaGetLastError());
    g_launches += 3;
    CU(cudaStreamSynchronize(dev.st));  // rec is released on return
    return HB200_OK;
}

// Solve min |W^(1/2) (J c - data)|^2 (+ alpha |c|^2 on the column-scaled system) for c.
// jac (n x p row-major, device) is overwritten; data / weights / coef are device arrays.
// *path: 0 = Cholesky, 1 = SVD pseudo-inverse (no damping), 2 = SVD ridge filter (fallback).
int dense_least_squares(Dev& dev, double* jac, int64_t n, int64_t p, const double* data,
                        const double* weights, bool damped, double alpha, double* coef, int* path)
{
    if (n <= 0 || p <= 0) return fail(HB200_EINVAL, "empty system (%lld x %lld)", (long long)n, (long long)p);
    if (n > INT_MAX || p > INT_MAX) return fail(HB200_EINVAL, "system too large for the dense solver");
    int rc = dense_acquire(dev);
    if (rc) return rc;
    DenseLibs& L = g_dense;
    cudaStream_t st = dev.st;
    const int in = (int)n, ip = (int)p;
    const double one = 1.0, zero = 0.0;

    DevBuf b_scale, b_y, b_x, b_info, b_piv;
    if ((rc = b_piv.alloc(2 * sizeof(double)))) return rc;
    double* y_scratch = b_piv.d();
    if ((rc = b_scale.alloc(p * sizeof(double))) || (rc = b_y.alloc(n * sizeof(double)))
        || (rc = b_x.alloc(std::max(n, p) * sizeof(double))) || (rc = b_info.alloc(sizeof(int))))
        return rc;
    double* scale = b_scale.d();
    double* y = b_y.d();
    double* x = b_x.d();
    int* d_info = static_cast<int*>(b_info.p);

    column_scale_kernel<<<(unsigned)((p + 31) / 32), dim3(32, 8), 0, st>>>(jac, n, p, scale);
    scale_system_kernel<<<dim3((unsigned)((n + 15) / 16), blocks_for(p)), 256, 0, st>>>(
        jac, n, p, scale, weights, data, y);
    CU(cudaGetLastError());
    g_launches += 2;

    // cuBLAS is column-major: the row-major n x p `jac` is M = J' (p x n, leading dimension p)
    bool need_svd = !damped;
    int svd_mode = 0;
    double svd_param = g_fit_rcond;
    if (path) *path = 0;
    if (damped) {
        const bool primal = p <= n;
        const int k = primal ? ip : in;
        DevBuf b_g, b_work;
        if ((rc = b_g.alloc((size_t)k * k * sizeof(double)))) return rc;
        double* g = b_g.d();
        if (primal) {
            BLAS(L.dsyrk(L.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, ip, in, &one, jac, ip, &zero, g, ip));
            BLAS(L.dgemv(L.blas, CUBLAS_OP_N, ip, in, &one, jac, ip, y, 1, &zero, x, 1));
        } else {
            BLAS(L.dsyrk(L.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, in, ip, &one, jac, ip, &zero, g, in));
            CU(cudaMemcpyAsync(x, y, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        }
        add_diagonal_kernel<<<blocks_for(k), 256, 0, st>>>(g, k, alpha);
        CU(cudaGetLastError());
        g_launches += 1;
        int lwork = 0;
        SOLVER(L.potrfSize(L.solver, CUBLAS_FILL_MODE_LOWER, k, g, k, &lwork));
        if ((rc = b_work.alloc((size_t)std::max(lwork, 1) * sizeof(double)))) return rc;
        SOLVER(L.potrf(L.solver, CUBLAS_FILL_MODE_LOWER, k, g, k, b_work.d(), lwork, d_info));
        int info = 0;
        double pivots[2] = {1.0, 1.0};
        diagonal_minmax_kernel<<<1, 256, 0, st>>>(g, k, y_scratch);
        CU(cudaGetLastError());
        g_launches += 1;
        CU(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(pivots, y_scratch, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // A factorisation that "succeeds" on numerically singular normal equations (pivot ratio
        // squared below machine epsilon) solves noise: treat it like the failed one.
        const double ratio = pivots[0] / pivots[1];
        if (info == 0 && !(ratio * ratio > 2.220446049250313e-16)) info = k + 1;
        if (info == 0) {
            SOLVER(L.potrs(L.solver, CUBLAS_FILL_MODE_LOWER, k, 1, g, k, x, k, d_info));
            if (!primal) {
                // c = X' d: M (p x n) times the dual solution
                if ((rc = b_work.alloc(n * sizeof(double)))) return rc;
                CU(cudaMemcpyAsync(b_work.d(), x, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
                BLAS(L.dgemv(L.blas, CUBLAS_OP_N, ip, in, &one, jac, ip, b_work.d(), 1, &zero, x, 1));
                CU(cudaStreamSynchronize(st));
            }
        } else if (info > 0) {
            // not positive definite in floating point: sklearn falls back to its SVD solver
            need_svd = true;
            svd_mode = 1;
            svd_param = alpha;
            if (path) *path = 2;
        } else {
            return fail(HB200_ECUDA, "cusolverDnDpotrf: illegal argument %d", -info);
        }
    }
    if (need_svd) {
        if (path && !damped) *path = 1;
        // gesvd wants rows >= columns. p >= n: factor M (p x n) in place. p < n: factor the
        // column-major n x p copy of J (= M').
        const bool tall_m = p >= n;
        const int m = tall_m ? ip : in, k = tall_m ? in : ip;
        DevBuf b_a, b_s, b_u, b_vt, b_t, b_work;
        double* a = jac;
        if (!tall_m) {
            if ((rc = b_a.alloc((size_t)n * p * sizeof(double)))) return rc;
            a = b_a.d();
            BLAS(L.dgeam(L.blas, CUBLAS_OP_T, CUBLAS_OP_N, in, ip, &one, jac, ip, &zero, a, in, a, in));
        }
        if ((rc = b_s.alloc(k * sizeof(double))) || (rc = b_u.alloc((size_t)m * k * sizeof(double)))
            || (rc = b_vt.alloc((size_t)k * k * sizeof(double))) || (rc = b_t.alloc(k * sizeof(double))))
            return rc;
        int lwork = 0;
        SOLVER(L.gesvdSize(L.solver, m, k, &lwork));
        if ((rc = b_work.alloc((size_t)std::max(lwork, 1) * sizeof(double)))) return rc;
        SOLVER(L.gesvd(L.solver, 'S', 'S', m, k, a, m, b_s.d(), b_u.d(), m, b_vt.d(), k, b_work.d(),
                       lwork, nullptr, d_info));
        int info = 0;
        CU(cudaMemcpyAsync(&info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (info != 0) return fail(HB200_ECUDA, "cusolverDnDgesvd did not converge (info %d)", info);
        double* t = b_t.d();
        if (tall_m) {
            // M = U S VT  =>  J = M' = VT' S U',  c = U S^+ (VT y)
            BLAS(L.dgemv(L.blas, CUBLAS_OP_N, k, k, &one, b_vt.d(), k, y, 1, &zero, t, 1));
            singular_filter_kernel<<<blocks_for(k), 256, 0, st>>>(t, b_s.d(), k, svd_mode, svd_param);
            BLAS(L.dgemv(L.blas, CUBLAS_OP_N, m, k, &one, b_u.d(), m, t, 1, &zero, x, 1));
        } else {
            // J = U S VT,  c = VT' S^+ (U' y)
            BLAS(L.dgemv(L.blas, CUBLAS_OP_T, m, k, &one, b_u.d(), m, y, 1, &zero, t, 1));
            singular_filter_kernel<<<blocks_for(k), 256, 0, st>>>(t, b_s.d(), k, svd_mode, svd_param);
            BLAS(L.dgemv(L.blas, CUBLAS_OP_T, k, k, &one, b_vt.d(), k, t, 1, &zero, x, 1));
        }
        CU(cudaGetLastError());
        g_launches += 1;
    }
    unscale_kernel<<<blocks_for(p), 256, 0, st>>>(x, scale, p, coef);
    CU(cudaGetLastError());
    g_launches += 1;
    CU(cudaStreamSynchronize(st));  // the scratch buffers are released on return
    return HB200_OK;
}

}  // namespace
