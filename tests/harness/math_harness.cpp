// tests/harness/math_harness.cpp -- TEST INFRASTRUCTURE.
// Compiles the product's per-pair math (harmonica_b200/csrc/hb200_math.cuh,
// hb200_fast.cuh) for the HOST so that the CPU test-suite can check the very
// statements the CUDA kernels execute against the oracle without a GPU.
// Never loaded by the product package.
#include <cstdint>
#include <cstring>

#include "../../harmonica_b200/csrc/hb200_math.cuh"
#include "../../harmonica_b200/csrc/hb200_fast.cuh"
#include "../../harmonica_b200/csrc/hb200_tess.cuh"

using namespace hb;

template <int FS>
static void pair_fs(int variant, const double* o, const double* p, const double* prm, unsigned rules,
                    double* acc, unsigned* flags)
{
    PairGeom g;
    make_geom(g, o[0], o[1], o[2], p[0], p[1], p[2], p[3], p[4], p[5]);
    unsigned f = 0;
    // the kernel's own dispatch (hb200_fast.cuh)
    if (variant == 0) prism_pair<FS, 0>(g, prm, rules, acc, f);
    else if (variant == 1) prism_pair<FS, 1>(g, prm, rules, acc, f);
    else prism_pair<FS, 2>(g, prm, rules, acc, f);
    *flags |= f;
}

static void pair_any(int fs, int variant, const double* o, const double* p, const double* prm,
                     unsigned rules, double* acc, unsigned* flags)
{
    switch (fs) {
#define C(X) case X: pair_fs<X>(variant, o, p, prm, rules, acc, flags); break;
        C(F_POT) C(F_E) C(F_N) C(F_U) C(F_EE) C(F_NN) C(F_UU) C(F_EN) C(F_EU) C(F_NU)
        C(FS_ACC3) C(FS_TENSOR6) C(FS_MAG_B) C(FS_MAG_E) C(FS_MAG_N) C(FS_MAG_U)
#undef C
    }
}

extern "C" {

// element-wise checks of the xmath sequences: op 0 rcp, 1 sqrt, 2 log, 3 atan2(a, b),
// 4 sqrt (1 ulp), 5 rsqrt, 6 log1p for |a| < 1/4 (atanh form)
void hbt_xmath(int op, int64_t n, const double* a, const double* b, double* out)
{
    for (int64_t i = 0; i < n; i++) {
        if (op == 0) out[i] = fast_rcp(a[i]);
        else if (op == 1) out[i] = fast_sqrt(a[i]);
        else if (op == 2) out[i] = fast_log(a[i]);
        else if (op == 3) out[i] = fast_atan2(a[i], b[i]);
        else if (op == 4) out[i] = fast_sqrt_1ulp(a[i]);
        else if (op == 6) out[i] = log1p_atanh(a[i]);
        else out[i] = fast_rsqrt(a[i]);
    }
}

int hbt_nout(int fs)
{
    if (fs == FS_ACC3 || fs == FS_MAG_B) return 3;
    if (fs == FS_TENSOR6) return 6;
    return 1;
}

// out[c * n_obs + i] = sum_j pair(i, j)[c]; prm is (n_prisms, 3) row-major
void hbt_prism_loop(int fs, int variant, int64_t n_obs, const double* oe, const double* on,
                    const double* ou, int64_t n_prisms, const double* prisms, const double* prm,
                    unsigned rules, double* out, unsigned* flags)
{
    const int nout = hbt_nout(fs);
    unsigned f = 0;
    for (int64_t i = 0; i < n_obs; i++) {
        double acc[6] = {0, 0, 0, 0, 0, 0};
        const double o[3] = {oe[i], on[i], ou[i]};
        for (int64_t j = 0; j < n_prisms; j++)
            pair_any(fs, variant, o, prisms + 6 * j, prm + 3 * j, rules, acc, &f);
        for (int c = 0; c < nout; c++) out[c * n_obs + i] = acc[c];
    }
    if (flags) *flags = f;
}
}

extern "C" {

// host build of the tesseroid pair function (hb200_tess.cuh), summed like the kernel does:
// out[i] = sum_j pair(i, j); counts (may be null): leaves per pair, n_obs x n_tess
void hbt_tesseroid_loop(int field, int64_t n_obs, const double* lon, const double* lat,
                        const double* rad, int64_t n_tess, const double* tesseroids,
                        const double* density, const double* density_upper, int radial, double* out,
                        int64_t* counts, unsigned* flags)
{
    double stack[kTessStack * 6];
    unsigned f = 0;
    const double ratio = field == F_POT ? 1.0 : 2.5;
    for (int64_t i = 0; i < n_obs; i++) {
        TessObs o;
        tess_make_obs(o, lon[i], lat[i], rad[i]);
        double acc = 0.0;
        for (int64_t j = 0; j < n_tess; j++) {
            int leaves;
            if (field == F_POT)
                leaves = tess_pair<F_POT>(o, tesseroids + 6 * j, density[j], density_upper[j], ratio, radial != 0, stack, acc, f);
            else
                leaves = tess_pair<F_U>(o, tesseroids + 6 * j, density[j], density_upper[j], ratio, radial != 0, stack, acc, f);
            if (counts) counts[i * n_tess + j] = leaves;
        }
        out[i] = acc;
    }
    if (flags) *flags = f;
}

// The per-thread algorithm of tesseroid_deferred_kernel on the host: root records, root decision
// from the record, unsplit pairs integrated at once, split pairs noted and walked when `defer_cap`
// of them are pending and at the end (one lane, so no warp vote). leaves: leaves per pair.
void hbt_tesseroid_loop_deferred(int field, int64_t n_obs, const double* lon, const double* lat,
                                 const double* rad, int64_t n_tess, const double* tesseroids,
                                 const double* density, const double* density_upper, int radial,
                                 int defer_cap, int fast, double* out, int64_t* counts,
                                 unsigned* flags)
{
    double stack[kTessStack * 6];
    double* records = new double[(size_t)(n_tess > 0 ? n_tess : 1) * kTessRec];
    for (int64_t j = 0; j < n_tess; j++) {
        if (fast)
            tess_pack_record_fast(records + j * kTessRec, tesseroids + 6 * j, density[j], density_upper[j],
                                  field == F_POT ? 1.0 : 2.5, radial != 0);
        else
            tess_pack_record(records + j * kTessRec, tesseroids + 6 * j, density[j], density_upper[j]);
    }
    int64_t* defer = new int64_t[defer_cap > 0 ? defer_cap : 1];
    unsigned f = 0;
    const double ratio = field == F_POT ? 1.0 : 2.5;
    for (int64_t i = 0; i < n_obs; i++) {
        TessObs o;
        tess_make_obs(o, lon[i], lat[i], rad[i]);
        double acc = 0.0;
        int n_defer = 0;
        auto walk = [&]() {
            for (int k = 0; k < n_defer; k++) {
                const double* rec = records + defer[k] * kTessRec;
                int leaves;
                const double rho1 = rec[fast ? kTessRho1Fast : kTessRho1];
                if (fast == 2) {
                    if (field == F_POT) leaves = tess_pair<F_POT, kTessStack, kTessMaxLeaves, OwnTrig>(o, rec, rec[6], rho1, ratio, radial != 0, stack, acc, f);
                    else leaves = tess_pair<F_U, kTessStack, kTessMaxLeaves, OwnTrig>(o, rec, rec[6], rho1, ratio, radial != 0, stack, acc, f);
                } else if (field == F_POT) leaves = tess_pair<F_POT>(o, rec, rec[6], rho1, ratio, radial != 0, stack, acc, f);
                else leaves = tess_pair<F_U>(o, rec, rec[6], rho1, ratio, radial != 0, stack, acc, f);
                if (counts) counts[i * n_tess + defer[k]] = leaves;
            }
            n_defer = 0;
        };
        for (int64_t j = 0; j < n_tess; j++) {
            const double* rec = records + j * kTessRec;
            int r;
            if (fast) r = field == F_POT ? tess_root_fast<F_POT>(o, rec, acc, f) : tess_root_fast<F_U>(o, rec, acc, f);
            else r = field == F_POT ? tess_root<F_POT>(o, rec, ratio, radial != 0, acc, f)
                                    : tess_root<F_U>(o, rec, ratio, radial != 0, acc, f);
            if (counts) counts[i * n_tess + j] = r == 1 ? 1 : 0;
            if (r == 0) defer[n_defer++] = j;
            if (n_defer == defer_cap) walk();
        }
        walk();
        out[i] = acc;
    }
    delete[] records;
    delete[] defer;
    if (flags) *flags = f;
}

// test/test_tesseroid.py:317-336: tiny stack / tiny leaf budget provoke the overflow flags
unsigned hbt_tesseroid_overflow(int which, const double* point, const double* tesseroid, double ratio)
{
    double stack[kTessStack * 6];
    TessObs o;
    tess_make_obs(o, point[0], point[1], point[2]);
    double acc = 0.0;
    unsigned f = 0;
    if (which == 0) tess_pair<F_POT, 2, kTessMaxLeaves>(o, tesseroid, 1.0, 1.0, ratio, false, stack, acc, f);
    else tess_pair<F_POT, kTessStack, 2>(o, tesseroid, 1.0, 1.0, ratio, false, stack, acc, f);
    return f;
}

}  // extern "C"

#include "../../harmonica_b200/csrc/hb200_trig.cuh"
extern "C" {
// op 0 sin, 1 cos (fast_sincos), 2 acos (fast_acos)
void hbt_trig(int op, int64_t n, const double* a, double* out)
{
    for (int64_t i = 0; i < n; i++) {
        if (op == 2) out[i] = hb::fast_acos(a[i]);
        else {
            double s, c;
            hb::fast_sincos(a[i], s, c);
            out[i] = op == 0 ? s : c;
        }
    }
}
}
