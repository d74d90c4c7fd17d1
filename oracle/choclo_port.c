/*
 * oracle/choclo_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE float64, no FMA contraction, glibc libm) of the
 * pairwise forward-modelling hot path of fatiando/harmonica:
 *
 *   - the per-pair closed-form kernels that live in the third-party dependency
 *     `choclo` (declared only as "choclo >= 0.1" in the reference's
 *     pyproject.toml:41; NOT vendored under /root/reference and NOT installed
 *     in this image), restated from the published algorithm (Nagy et al. 2000,
 *     2002; Fukushima 2020, the three papers the reference cites at
 *     src/harmonica/_forward/prisms/gravity.py:160-164) and from the
 *     reference's own call sites and tests;
 *   - the jitted double loops of the reference itself
 *       src/harmonica/_forward/prisms/gravity.py:524-537   (jit_prism_gravity)
 *       src/harmonica/_forward/prisms/magnetic.py:317-335   (_jit_prism_magnetic_field)
 *       src/harmonica/_forward/prisms/magnetic.py:382-397   (_jit_prism_magnetic_component)
 *       src/harmonica/_forward/prisms/layer.py:582-625      (_forward_gravity_prism_layer)
 *       src/harmonica/_forward/point.py:388-398, 426-447    (point_mass_cartesian / _spherical)
 *       src/harmonica/_equivalent_sources/utils.py:86-97    (predict)
 *     in the same loop shape: outer parallel loop over observers (OpenMP here,
 *     numba.prange there), inner serial loop over sources, one accumulator per
 *     observer living in the output array.
 *
 * PARITY PINS (see oracle/README.md and tests/test_oracle_pins.py): point
 * potential vs the reference's golden CSV, prism g_z vs the reference's two
 * doctests and the Bouguer-slab limit, Laplace identity, finite differences,
 * an mpmath quadrature of the potential, and bit-comparison against the
 * reference's UNMODIFIED wrappers+loops driven through oracle/ref_shim.py.
 * "parity unpinned" items (no reference test pins them): NaN-on-edge and +4pi
 * face rules of the magnetic kernels, the digits of choclo's mu_0. The
 * absolute values / units / signs of prism_magnetic (which the reference only
 * compares with choclo) are pinned formula-independently by a Gauss-Legendre
 * quadrature of the dipole field over the prism volume (test_oracle_pins.py);
 * the dipole field itself has a REFERENCE-HELD pin that does not go through
 * choclo: the reference's own ellipsoid_magnetic (numpy + scipy) for
 * uniformly magnetised spheres, which are exactly dipoles outside
 * (oracle/make_golden_ellipsoid.py, tests/golden/ellipsoid_sphere_magnetic.npz):
 * agreement to 1.35e-10, all of it the digits of mu_0 (scipy's CODATA 2022
 * value there, 4 pi 1e-7 here).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * arm may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HBO_G 6.6743e-11 /* choclo.constants.GRAVITATIONAL_CONST == harmonica/constants.py:12 */
#define HBO_PI 3.14159265358979323846

/* field ids shared with include/harmonica_b200.h */
enum {
    F_POT = 0, F_E = 1, F_N = 2, F_U = 3,
    F_EE = 4, F_NN = 5, F_UU = 6, F_EN = 7, F_EU = 8, F_NU = 9
};

/* ------------------------------------------------------------------ K2 ---- */
/* choclo.prism._utils.safe_atan2: atan(y/x) with the x==0 limits. */
double hbo_safe_atan2(double y, double x)
{
    if (x != 0.0) return atan(y / x);
    if (y > 0.0) return HBO_PI / 2;
    if (y < 0.0) return -HBO_PI / 2;
    return 0.0;
}

/* choclo.prism._utils.safe_log: log(x + r) evaluated without cancellation
 * (Fukushima 2020), with the r==0 and on-axis limits. */
double hbo_safe_log(double x, double y, double z, double r)
{
    if (r == 0.0) return 0.0;
    if (x < 0.0) {
        if (r == -x) return -log(-2 * x);
        return log((y * y + z * z) / (r - x));
    }
    return log(x + r);
}

/* ------------------------------------------------------------------ K3 ---- */
typedef double (*kernel_fn)(double e, double n, double u, double r);

static double k_pot(double e, double n, double u, double r)
{
    return e * n * hbo_safe_log(u, e, n, r)
         + n * u * hbo_safe_log(e, n, u, r)
         + e * u * hbo_safe_log(n, e, u, r)
         - 0.5 * (e * e) * hbo_safe_atan2(u * n, e * r)
         - 0.5 * (n * n) * hbo_safe_atan2(u * e, n * r)
         - 0.5 * (u * u) * hbo_safe_atan2(e * n, u * r);
}
static double k_e(double e, double n, double u, double r)
{
    return -(n * hbo_safe_log(u, e, n, r) + u * hbo_safe_log(n, e, u, r)
             - e * hbo_safe_atan2(n * u, e * r));
}
static double k_n(double e, double n, double u, double r)
{
    return -(u * hbo_safe_log(e, n, u, r) + e * hbo_safe_log(u, e, n, r)
             - n * hbo_safe_atan2(u * e, n * r));
}
static double k_u(double e, double n, double u, double r)
{
    return -(e * hbo_safe_log(n, e, u, r) + n * hbo_safe_log(e, n, u, r)
             - u * hbo_safe_atan2(e * n, u * r));
}
static double k_ee(double e, double n, double u, double r) { return -hbo_safe_atan2(n * u, e * r); }
static double k_nn(double e, double n, double u, double r) { return -hbo_safe_atan2(e * u, n * r); }
static double k_uu(double e, double n, double u, double r) { return -hbo_safe_atan2(e * n, u * r); }
static double k_en(double e, double n, double u, double r) { return hbo_safe_log(u, e, n, r); }
static double k_eu(double e, double n, double u, double r) { return hbo_safe_log(n, e, u, r); }
static double k_nu(double e, double n, double u, double r) { return hbo_safe_log(e, n, u, r); }

static const kernel_fn KERNELS[10] = { k_pot, k_e, k_n, k_u, k_ee, k_nn, k_uu, k_en, k_eu, k_nu };

/* ------------------------------------------------------------------ K1 ---- */
/* 8-vertex alternating sum; vertex order east/west x north/south x top/bottom. */
static double evaluate_kernel(double E, double N, double U, double w, double e, double s,
                              double n, double b, double t, kernel_fn kernel)
{
    double result = 0.0;
    for (int i = 0; i < 2; i++) {
        double se = (i == 0 ? e : w) - E;
        double se2 = se * se;
        for (int j = 0; j < 2; j++) {
            double sn = (j == 0 ? n : s) - N;
            double sn2 = sn * sn;
            for (int k = 0; k < 2; k++) {
                double su = (k == 0 ? t : b) - U;
                double su2 = su * su;
                double r = sqrt(se2 + sn2 + su2);
                double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                result += sign * kernel(se, sn, su, r);
            }
        }
    }
    return result;
}

/* ------------------------------------------------------------------ K6 ---- */
static int on_easting_edge(double E, double N, double U, double w, double e, double s, double n,
                           double b, double t)
{
    return (w <= E && E <= e) && (N == s || N == n) && (U == b || U == t);
}
static int on_northing_edge(double E, double N, double U, double w, double e, double s, double n,
                            double b, double t)
{
    return (s <= N && N <= n) && (E == w || E == e) && (U == b || U == t);
}
static int on_upward_edge(double E, double N, double U, double w, double e, double s, double n,
                          double b, double t)
{
    return (b <= U && U <= t) && (E == w || E == e) && (N == s || N == n);
}
static int on_east_face(double E, double N, double U, double w, double e, double s, double n,
                        double b, double t)
{
    (void)w;
    return E == e && (s <= N && N <= n) && (b <= U && U <= t);
}
static int on_north_face(double E, double N, double U, double w, double e, double s, double n,
                         double b, double t)
{
    (void)s;
    return N == n && (w <= E && E <= e) && (b <= U && U <= t);
}
static int on_top_face(double E, double N, double U, double w, double e, double s, double n,
                       double b, double t)
{
    (void)b;
    return U == t && (w <= E && E <= e) && (s <= N && N <= n);
}

/* Predicate sets per tensor component: harmonica gravity.py:272-449 (the
 * reference's own _any_singular_point_g_*), same sets choclo uses for NaN. */
int hbo_prism_is_singular(int field, double E, double N, double U, double w, double e, double s,
                          double n, double b, double t)
{
    switch (field) {
    case F_EE: return on_northing_edge(E, N, U, w, e, s, n, b, t) || on_upward_edge(E, N, U, w, e, s, n, b, t);
    case F_NN: return on_easting_edge(E, N, U, w, e, s, n, b, t) || on_upward_edge(E, N, U, w, e, s, n, b, t);
    case F_UU: return on_easting_edge(E, N, U, w, e, s, n, b, t) || on_northing_edge(E, N, U, w, e, s, n, b, t);
    case F_EN: return on_upward_edge(E, N, U, w, e, s, n, b, t);
    case F_EU: return on_northing_edge(E, N, U, w, e, s, n, b, t);
    case F_NU: return on_easting_edge(E, N, U, w, e, s, n, b, t);
    default: return 0;
    }
}

/* choclo.prism.gravity_{pot,e,n,u,ee,nn,uu,en,eu,nu} */
double hbo_prism_gravity(int field, double E, double N, double U, double w, double e, double s,
                         double n, double b, double t, double density)
{
    if (field >= F_EE && hbo_prism_is_singular(field, E, N, U, w, e, s, n, b, t)) return NAN;
    double result = evaluate_kernel(E, N, U, w, e, s, n, b, t, KERNELS[field]);
    /* outside limit on the face whose outward normal is the component's +axis */
    if (field == F_EE && on_east_face(E, N, U, w, e, s, n, b, t)) result += 4 * HBO_PI;
    if (field == F_NN && on_north_face(E, N, U, w, e, s, n, b, t)) result += 4 * HBO_PI;
    if (field == F_UU && on_top_face(E, N, U, w, e, s, n, b, t)) result += 4 * HBO_PI;
    return HBO_G * density * result;
}

/* ------------------------------------------------------------------ K4 ---- */
/* choclo.prism.magnetic_field; flags: bit0 = NaN on any edge/vertex,
 * bit1 = +4pi outside-limit fix-up of k_ee/k_nn/k_uu on east/north/top face.
 * Default in this project: both on (HBO_MAG_DEFAULT = 3). [parity unpinned] */
void hbo_prism_magnetic_field(double E, double N, double U, double w, double e, double s, double n,
                              double b, double t, double me, double mn, double mu, int flags,
                              double out[3])
{
    if ((flags & 1) && (on_easting_edge(E, N, U, w, e, s, n, b, t)
                        || on_northing_edge(E, N, U, w, e, s, n, b, t)
                        || on_upward_edge(E, N, U, w, e, s, n, b, t))) {
        out[0] = out[1] = out[2] = NAN;
        return;
    }
    double be = 0.0, bn = 0.0, bu = 0.0;
    for (int i = 0; i < 2; i++) {
        double se = (i == 0 ? e : w) - E;
        double se2 = se * se;
        for (int j = 0; j < 2; j++) {
            double sn = (j == 0 ? n : s) - N;
            double sn2 = sn * sn;
            for (int k = 0; k < 2; k++) {
                double su = (k == 0 ? t : b) - U;
                double su2 = su * su;
                double r = sqrt(se2 + sn2 + su2);
                double sign = ((i + j + k) & 1) ? -1.0 : 1.0;
                double ee = k_ee(se, sn, su, r), nn = k_nn(se, sn, su, r), uu = k_uu(se, sn, su, r);
                double en = k_en(se, sn, su, r), eu = k_eu(se, sn, su, r), nu = k_nu(se, sn, su, r);
                be += sign * (me * ee + mn * en + mu * eu);
                bn += sign * (me * en + mn * nn + mu * nu);
                bu += sign * (me * eu + mn * nu + mu * uu);
            }
        }
    }
    if (flags & 2) {
        if (on_east_face(E, N, U, w, e, s, n, b, t)) be += me * (4 * HBO_PI);
        if (on_north_face(E, N, U, w, e, s, n, b, t)) bn += mn * (4 * HBO_PI);
        if (on_top_face(E, N, U, w, e, s, n, b, t)) bu += mu * (4 * HBO_PI);
    }
    const double mu0 = 4 * HBO_PI * 1e-7; /* choclo.constants.VACUUM_MAGNETIC_PERMEABILITY [RECALL] */
    const double cm = mu0 / 4 / HBO_PI;
    out[0] = cm * be;
    out[1] = cm * bn;
    out[2] = cm * bu;
}

/* ------------------------------------------------------------------ K5 ---- */
/* choclo.point.gravity_*; *zero_div is set when distance == 0 (the reference
 * raises ZeroDivisionError there under numba's python error model). */
double hbo_point_gravity(int field, double E, double N, double U, double eq, double nq, double uq,
                         double mass, int* zero_div)
{
    double de = E - eq, dn = N - nq, du = U - uq;
    double d = sqrt(de * de + dn * dn + du * du);
    if (d == 0.0) { if (zero_div) *zero_div = 1; }
    double k;
    switch (field) {
    case F_POT: k = 1 / d; break;
    case F_E: k = -de / (d * d * d); break;
    case F_N: k = -dn / (d * d * d); break;
    case F_U: k = -du / (d * d * d); break;
    case F_EE: k = 3 * de * de / (d * d * d * d * d) - 1 / (d * d * d); break;
    case F_NN: k = 3 * dn * dn / (d * d * d * d * d) - 1 / (d * d * d); break;
    case F_UU: k = 3 * du * du / (d * d * d * d * d) - 1 / (d * d * d); break;
    case F_EN: k = 3 * de * dn / (d * d * d * d * d); break;
    case F_EU: k = 3 * de * du / (d * d * d * d * d); break;
    case F_NU: k = 3 * dn * du / (d * d * d * d * d); break;
    default: k = NAN;
    }
    return HBO_G * mass * k;
}

/* ------------------------------------------------------------- loops ---- */
static void set_threads(int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
}

int hbo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* gravity.py:524-537: out[i] += forward_func(obs_i, prisms[j,:], density[j]).
 * `out` must be zero-initialised by the caller (result = np.zeros, :201).
 * f32acc != 0 emulates dtype=float32 output arrays (round after every +=). */
void hbo_prism_gravity_loop(int field, int64_t n_obs, const double* oe, const double* on,
                            const double* ou, int64_t n_prisms, const double* prisms,
                            const double* density, double* out, int f32acc, int nthreads)
{
    set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        for (int64_t j = 0; j < n_prisms; j++) {
            const double* p = prisms + 6 * j;
            acc += hbo_prism_gravity(field, oe[i], on[i], ou[i], p[0], p[1], p[2], p[3], p[4], p[5],
                                     density[j]);
            if (f32acc) acc = (double)(float)acc;
        }
        out[i] = acc;
    }
}

/* gravity.py:272-449: any (observer, prism) pair singular for `field`? */
int hbo_prism_any_singular(int field, int64_t n_obs, const double* oe, const double* on,
                           const double* ou, int64_t n_prisms, const double* prisms)
{
    for (int64_t i = 0; i < n_obs; i++)
        for (int64_t j = 0; j < n_prisms; j++) {
            const double* p = prisms + 6 * j;
            if (hbo_prism_is_singular(field, oe[i], on[i], ou[i], p[0], p[1], p[2], p[3], p[4], p[5]))
                return 1;
        }
    return 0;
}

/* magnetic.py:317-335 (component < 0: all three, out = 3 arrays be|bn|bu of
 * n_obs each) and magnetic.py:382-397 (component 0,1,2: single, out = n_obs). */
void hbo_prism_magnetic_loop(int component, int64_t n_obs, const double* oe, const double* on,
                             const double* ou, int64_t n_prisms, const double* prisms,
                             const double* me, const double* mn, const double* mu, int flags,
                             double* out, int nthreads)
{
    set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_obs; i++) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, b[3];
        for (int64_t j = 0; j < n_prisms; j++) {
            const double* p = prisms + 6 * j;
            hbo_prism_magnetic_field(oe[i], on[i], ou[i], p[0], p[1], p[2], p[3], p[4], p[5], me[j],
                                     mn[j], mu[j], flags, b);
            a0 += b[0];
            a1 += b[1];
            a2 += b[2];
        }
        if (component < 0) {
            out[i] += a0;
            out[n_obs + i] += a1;
            out[2 * n_obs + i] += a2;
        } else {
            out[i] += (component == 0 ? a0 : component == 1 ? a1 : a2);
        }
    }
}

/* layer.py:582-625: bounds built on the fly; easting-outer / northing-inner
 * order; skip rules in the reference's order. bottom/top/density are
 * (n_north, n_east) C-order. */
void hbo_prism_layer_loop(int field, int64_t n_obs, const double* oe, const double* on,
                          const double* ou, int64_t n_east, int64_t n_north, const double* east_c,
                          const double* north_c, const double* bottom, const double* top,
                          const double* density, double thickness_threshold, double* out,
                          int nthreads)
{
    set_threads(nthreads);
    const double half_e = (east_c[1] - east_c[0]) / 2;
    const double half_n = (north_c[1] - north_c[0]) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        for (int64_t j = 0; j < n_east; j++) {
            double w = east_c[j] - half_e, e = east_c[j] + half_e;
            for (int64_t k = 0; k < n_north; k++) {
                double rho = density[k * n_east + j];
                if (rho == 0.0 || isnan(rho)) continue;
                double b = bottom[k * n_east + j], t = top[k * n_east + j];
                if (t - b < thickness_threshold) continue;
                if (isnan(t) || isnan(b)) continue;
                double s = north_c[k] - half_n, n = north_c[k] + half_n;
                acc += hbo_prism_gravity(field, oe[i], on[i], ou[i], w, e, s, n, b, t, rho);
            }
        }
        out[i] = acc;
    }
}

/* point.py:388-398 */
int hbo_point_cartesian_loop(int field, int64_t n_obs, const double* oe, const double* on,
                             const double* ou, int64_t n_src, const double* pe, const double* pn,
                             const double* pu, const double* mass, double* out, int nthreads)
{
    int zero_div = 0;
    set_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : zero_div)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        int zd = 0;
        for (int64_t j = 0; j < n_src; j++)
            acc += hbo_point_gravity(field, oe[i], on[i], ou[i], pe[j], pn[j], pu[j], mass[j], &zd);
        out[i] = acc;
        zero_div |= zd;
    }
    return zero_div;
}

/* point.py:324-354 + 426-447 + _forward/utils.py:198-201. Angles in degrees.
 * field: F_POT or F_U only. */
int hbo_point_spherical_loop(int field, int64_t n_obs, const double* lon, const double* lat,
                             const double* rad, int64_t n_src, const double* lon_p,
                             const double* lat_p, const double* rad_p, const double* mass,
                             double* out, double* scratch /* 3*(n_obs+n_src) */, int nthreads)
{
    const double d2r = HBO_PI / 180.0; /* np.radians multiplies by pi/180 */
    double* lam = scratch;
    double* cphi = lam + n_obs;
    double* sphi = cphi + n_obs;
    double* lam_p = sphi + n_obs;
    double* cphi_p = lam_p + n_src;
    double* sphi_p = cphi_p + n_src;
    for (int64_t i = 0; i < n_obs; i++) {
        lam[i] = lon[i] * d2r;
        double phi = lat[i] * d2r;
        cphi[i] = cos(phi);
        sphi[i] = sin(phi);
    }
    for (int64_t j = 0; j < n_src; j++) {
        lam_p[j] = lon_p[j] * d2r;
        double phi = lat_p[j] * d2r;
        cphi_p[j] = cos(phi);
        sphi_p[j] = sin(phi);
    }
    int zero_div = 0;
    set_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : zero_div)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        for (int64_t j = 0; j < n_src; j++) {
            double coslambda = cos(lam_p[j] - lam[i]);
            double cospsi = sphi_p[j] * sphi[i] + cphi_p[j] * cphi[i] * coslambda;
            double dr = rad[i] - rad_p[j];
            double dist = sqrt(dr * dr + 2 * rad[i] * rad_p[j] * (1 - cospsi));
            if (dist == 0.0) zero_div = 1;
            double k;
            if (field == F_POT) {
                k = 1 / dist * HBO_G;
            } else {
                double delta_z = rad[i] - rad_p[j] * cospsi;
                k = -HBO_G * delta_z / (dist * dist * dist);
            }
            acc += mass[j] * k;
        }
        out[i] = acc;
    }
    return zero_div;
}

/* _equivalent_sources/utils.py:86-97 with cartesian.py:641-644 and
 * _forward/utils.py:111-118: result[i] += coeffs[j] * (1 / distance). */
int hbo_eqs_predict_loop(int64_t n_obs, const double* oe, const double* on, const double* ou,
                         int64_t n_src, const double* pe, const double* pn, const double* pu,
                         const double* coefs, double* out, int nthreads)
{
    int zero_div = 0;
    set_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : zero_div)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        for (int64_t j = 0; j < n_src; j++) {
            double de = oe[i] - pe[j], dn = on[i] - pn[j], du = ou[i] - pu[j];
            double dist = sqrt(de * de + dn * dn + du * du);
            if (dist == 0.0) zero_div = 1;
            acc += coefs[j] * (1 / dist);
        }
        out[i] = acc;
    }
    return zero_div;
}

/* _equivalent_sources/utils.py:54-74: jac[i, j] = 1 / distance. */
void hbo_eqs_jacobian_loop(int64_t n_obs, const double* oe, const double* on, const double* ou,
                           int64_t n_src, const double* pe, const double* pn, const double* pu,
                           double* jac, int nthreads)
{
    set_threads(nthreads);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_obs; i++)
        for (int64_t j = 0; j < n_src; j++) {
            double de = oe[i] - pe[j], dn = on[i] - pn[j], du = ou[i] - pu[j];
            jac[i * n_src + j] = 1 / sqrt(de * de + dn * dn + du * du);
        }
}

/* ------------------------------------------------------------ dipoles ---- */
/* choclo.dipole.magnetic_field [RECALL; the dipole formula itself is the pin]:
 * B = mu0/4pi * (3 (m.r) r / d^5 - m / d^3), r = observer - dipole. Reference
 * loop: src/harmonica/_forward/dipole.py:329-347 (vector), :386-400 (component).
 * *zero_div is set for a coincident observer/dipole (the reference raises). */
void hbo_dipole_magnetic_field(double E, double N, double U, double eq, double nq, double uq,
                               double me, double mn, double mu, double out[3], int* zero_div)
{
    const double re = E - eq, rn = N - nq, ru = U - uq;
    const double d = sqrt(re * re + rn * rn + ru * ru);
    if (d == 0.0 && zero_div) *zero_div = 1;
    const double dot = me * re + mn * rn + mu * ru;
    const double mu0 = 4 * HBO_PI * 1e-7;
    const double cm = mu0 / 4 / HBO_PI;
    const double d3 = d * d * d, d5 = d3 * d * d;
    out[0] = cm * (3 * dot * re / d5 - me / d3);
    out[1] = cm * (3 * dot * rn / d5 - mn / d3);
    out[2] = cm * (3 * dot * ru / d5 - mu / d3);
}

int hbo_dipole_magnetic_loop(int component, int64_t n_obs, const double* oe, const double* on,
                             const double* ou, int64_t n_src, const double* pe, const double* pn,
                             const double* pu, const double* me, const double* mn,
                             const double* mu, double* out, int nthreads)
{
    int zero_div = 0;
    set_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : zero_div)
    for (int64_t i = 0; i < n_obs; i++) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, b[3];
        int zd = 0;
        for (int64_t j = 0; j < n_src; j++) {
            hbo_dipole_magnetic_field(oe[i], on[i], ou[i], pe[j], pn[j], pu[j], me[j], mn[j], mu[j],
                                      b, &zd);
            a0 += b[0];
            a1 += b[1];
            a2 += b[2];
        }
        zero_div |= zd;
        if (component < 0) {
            out[i] += a0;
            out[n_obs + i] += a1;
            out[2 * n_obs + i] += a2;
        } else {
            out[i] += (component == 0 ? a0 : component == 1 ? a1 : a2);
        }
    }
    return zero_div;
}

/* EquivalentSourcesSph.predict: _equivalent_sources/utils.py:86-97 with
 * greens_func_spherical (spherical.py:412-424) = 1 / distance_spherical
 * (_forward/utils.py:121-160, 198-201); angles in degrees. The per-point
 * radians/cos/sin the reference recomputes for every pair are hoisted (same
 * values). scratch: 3 * (n_obs + n_src) doubles. */
int hbo_eqs_predict_spherical_loop(int64_t n_obs, const double* lon, const double* lat,
                                   const double* rad, int64_t n_src, const double* lon_p,
                                   const double* lat_p, const double* rad_p, const double* coefs,
                                   double* out, double* scratch, int nthreads)
{
    const double d2r = HBO_PI / 180.0;
    double* lam = scratch;
    double* cphi = lam + n_obs;
    double* sphi = cphi + n_obs;
    double* lam_p = sphi + n_obs;
    double* cphi_p = lam_p + n_src;
    double* sphi_p = cphi_p + n_src;
    for (int64_t i = 0; i < n_obs; i++) {
        lam[i] = lon[i] * d2r;
        cphi[i] = cos(lat[i] * d2r);
        sphi[i] = sin(lat[i] * d2r);
    }
    for (int64_t j = 0; j < n_src; j++) {
        lam_p[j] = lon_p[j] * d2r;
        cphi_p[j] = cos(lat_p[j] * d2r);
        sphi_p[j] = sin(lat_p[j] * d2r);
    }
    int zero_div = 0;
    set_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(| : zero_div)
    for (int64_t i = 0; i < n_obs; i++) {
        double acc = out[i];
        for (int64_t j = 0; j < n_src; j++) {
            double coslambda = cos(lam_p[j] - lam[i]);
            double cospsi = sphi_p[j] * sphi[i] + cphi_p[j] * cphi[i] * coslambda;
            double dr = rad[i] - rad_p[j];
            double dist = sqrt(dr * dr + 2 * rad[i] * rad_p[j] * (1 - cospsi));
            if (dist == 0.0) zero_div = 1;
            acc += coefs[j] * (1 / dist);
        }
        out[i] = acc;
    }
    return zero_div;
}
