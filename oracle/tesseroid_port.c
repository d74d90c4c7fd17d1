/*
 * oracle/tesseroid_port.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement (IEEE float64, no FMA contraction, glibc libm, OpenMP over observers) of
 * the reference's constant-density tesseroid forward model, statement for statement:
 *
 *   jit_tesseroid_gravity          src/harmonica/_forward/tesseroid_gravity.py:236-339
 *   _adaptive_discretization       src/harmonica/_forward/_tesseroid_utils.py:136-217
 *   _split_tesseroid               src/harmonica/_forward/_tesseroid_utils.py:220-258
 *   _tesseroid_dimensions          src/harmonica/_forward/_tesseroid_utils.py:261-279
 *   _distance_tesseroid_point      src/harmonica/_forward/_tesseroid_utils.py:282-300
 *   gauss_legendre_quadrature      src/harmonica/_forward/_tesseroid_utils.py:19-107
 *   distance_spherical(_core)      src/harmonica/_forward/utils.py:121-201
 *   potential_spherical / gravity_u_spherical   src/harmonica/_forward/point.py:324-354
 *
 * with the reference's constants (tesseroid_gravity.py:30-33): STACK_SIZE = 100,
 * MAX_DISCRETIZATIONS = 100000, GLQ_DEGREES = (2, 2, 2), distance-size ratio 1 (potential) /
 * 2.5 (g_z). Like the reference it first collects the leaves of the adaptive discretisation of
 * one (observer, tesseroid) pair and then adds their quadratures to result[i] in that order.
 *
 * PARITY PIN: bit-compared in the build container against the reference's UNMODIFIED
 * tesseroid_gravity (real numba) driven through oracle/ref_shim.py; its outputs are committed
 * as tests/golden/tesseroid_*.npz (oracle/make_golden.py). Further pins: the spherical-shell
 * closed form of the reference's tests (test/test_tesseroid.py:664-770) and its discretisation
 * counts (:585-661), see tests/test_tesseroid_oracle.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define HBO_G 6.6743e-11
#define HBO_PI 3.14159265358979323846
#define STACK_SIZE 100
#define MAX_DISCRETIZATIONS 100000

#define HBO_TESS_STACK_OVERFLOW 1
#define HBO_TESS_MAX_DISCRETIZATIONS 2
#define HBO_TESS_ZERO_DIVISION 4 /* numba's float division raises ZeroDivisionError on a zero divisor */

/* numpy.polynomial.legendre.leggauss(2): nodes -/+ 1/sqrt(3), weights 1 */
static const double GLQ_NODES[2] = {-0x1.279a74590331cp-1, 0x1.279a74590331cp-1};
static const double GLQ_WEIGHTS[2] = {1.0, 1.0};

static double radians(double x) { return x * (HBO_PI / 180.0); }

/* _forward/utils.py:164-201 */
static double distance_spherical_core(double longitude, double cosphi, double sinphi, double radius,
                                      double longitude_p, double cosphi_p, double sinphi_p,
                                      double radius_p, double* cospsi_out)
{
    double coslambda = cos(longitude_p - longitude);
    double cospsi = sinphi_p * sinphi + cosphi_p * cosphi * coslambda;
    double dr = radius - radius_p;
    double dist = sqrt(dr * dr + 2 * radius * radius_p * (1 - cospsi));
    *cospsi_out = cospsi;
    return dist;
}

/* _forward/utils.py:121-160: everything in degrees */
static double distance_spherical(const double* p, const double* q)
{
    double longitude = radians(p[0]), latitude = radians(p[1]);
    double longitude_p = radians(q[0]), latitude_p = radians(q[1]);
    double cosphi_p = cos(latitude_p), sinphi_p = sin(latitude_p);
    double cosphi = cos(latitude), sinphi = sin(latitude);
    double unused;
    return distance_spherical_core(longitude, cosphi, sinphi, p[2], longitude_p, cosphi_p, sinphi_p,
                                   q[2], &unused);
}

/* _tesseroid_utils.py:261-279 */
void hbo_tesseroid_dimensions(const double* t, double* l_lon, double* l_lat, double* l_rad)
{
    double w = radians(t[0]), e = radians(t[1]), s = radians(t[2]), n = radians(t[3]);
    double bottom = t[4], top = t[5];
    double latitude_center = (n + s) / 2;
    *l_lat = top * acos(sin(n) * sin(s) + cos(n) * cos(s));
    double sc = sin(latitude_center), cc = cos(latitude_center);
    *l_lon = top * acos(sc * sc + cc * cc * cos(e - w));
    *l_rad = top - bottom;
}

/* _tesseroid_utils.py:282-300 */
double hbo_distance_tesseroid_point(const double* coordinates, const double* t)
{
    double centre[3] = {(t[0] + t[1]) / 2, (t[2] + t[3]) / 2, (t[4] + t[5]) / 2};
    return distance_spherical(coordinates, centre);
}

/* _tesseroid_utils.py:220-258; returns the new stack top */
int hbo_split_tesseroid(const double* t, int n_lon, int n_lat, int n_rad, double* stack,
                        int stack_top)
{
    double w = t[0], e = t[1], s = t[2], n = t[3], bottom = t[4], top = t[5];
    double d_lon = (e - w) / n_lon, d_lat = (n - s) / n_lat, d_rad = (top - bottom) / n_rad;
    for (int i = 0; i < n_lon; i++)
        for (int j = 0; j < n_lat; j++)
            for (int k = 0; k < n_rad; k++) {
                stack_top += 1;
                double* q = stack + 6 * stack_top;
                q[0] = w + d_lon * i;
                q[1] = w + d_lon * (i + 1);
                q[2] = s + d_lat * j;
                q[3] = s + d_lat * (j + 1);
                q[4] = bottom + d_rad * k;
                q[5] = bottom + d_rad * (k + 1);
            }
    return stack_top;
}

/* _tesseroid_utils.py:136-217. Returns the number of leaves written to `small`, or a negative
 * HBO_TESS_* code where the reference raises OverflowError. stack: stack_size x 6 doubles,
 * small: max_small x 6 doubles. */
int64_t hbo_adaptive_discretization(const double* coordinates, const double* tesseroid,
                                    double distance_size_ratio, double* stack, int stack_size,
                                    double* small, int64_t max_small, int radial_discretization)
{
    for (int c = 0; c < 6; c++) stack[c] = tesseroid[c];
    int stack_top = 0;
    int64_t n_splits = 0;
    while (stack_top >= 0) {
        double t[6];
        for (int c = 0; c < 6; c++) t[c] = stack[6 * stack_top + c];
        stack_top -= 1;
        double l_lon, l_lat, l_rad;
        hbo_tesseroid_dimensions(t, &l_lon, &l_lat, &l_rad);
        double distance = hbo_distance_tesseroid_point(coordinates, t);
        int n_lon = 1, n_lat = 1, n_rad = 1;
        /* the three quotients are all evaluated (:186-191); a zero divisor raises in numba */
        if (l_lon == 0.0 || l_lat == 0.0 || l_rad == 0.0) return -HBO_TESS_ZERO_DIVISION;
        if (distance / l_lon < distance_size_ratio) n_lon = 2;
        if (distance / l_lat < distance_size_ratio) n_lat = 2;
        if (distance / l_rad < distance_size_ratio && radial_discretization) n_rad = 2;
        if (n_lon * n_lat * n_rad > 1) {
            if ((stack_top + 1) + n_lon * n_lat * n_rad > stack_size) return -HBO_TESS_STACK_OVERFLOW;
            stack_top = hbo_split_tesseroid(t, n_lon, n_lat, n_rad, stack, stack_top);
        } else {
            if (n_splits + 1 > max_small) return -HBO_TESS_MAX_DISCRETIZATIONS;
            for (int c = 0; c < 6; c++) small[6 * n_splits + c] = t[c];
            n_splits += 1;
        }
    }
    return n_splits;
}

/* _tesseroid_utils.py:19-107 with the kernels of point.py:324-354. field 0 potential, 3 g_z
 * (radial / "upward" component, before the sign flip of tesseroid_gravity.py:222-223). */
double hbo_glq_tesseroid(int field, double longitude, double cosphi, double sinphi, double radius,
                         const double* t, double density, int* zero_div)
{
    double w = t[0], e = t[1], s = t[2], n = t[3], bottom = t[4], top = t[5];
    double a_factor = 1.0 / 8 * radians(e - w) * radians(n - s) * (top - bottom);
    double result = 0.0;
    for (int j = 0; j < 2; j++) {
        double latitude_p = radians(0.5 * (n - s) * GLQ_NODES[j] + 0.5 * (n + s));
        double cosphi_p = cos(latitude_p), sinphi_p = sin(latitude_p);
        for (int k = 0; k < 2; k++) {
            double radius_p = 0.5 * (top - bottom) * GLQ_NODES[k] + 0.5 * (top + bottom);
            double kappa = radius_p * radius_p * cosphi_p;
            for (int i = 0; i < 2; i++) {
                double longitude_p = radians(0.5 * (e - w) * GLQ_NODES[i] + 0.5 * (e + w));
                double mass = density * a_factor * kappa * GLQ_WEIGHTS[i] * GLQ_WEIGHTS[j] * GLQ_WEIGHTS[k];
                double cospsi;
                double dist = distance_spherical_core(longitude, cosphi, sinphi, radius, longitude_p,
                                                      cosphi_p, sinphi_p, radius_p, &cospsi);
                double kern;
                if (dist == 0.0) *zero_div = 1; /* 1 / distance, delta_z / distance**3 raise */
                if (field == 0) {
                    kern = 1 / dist * HBO_G;
                } else {
                    double delta_z = radius - radius_p * cospsi;
                    kern = -HBO_G * delta_z / (dist * dist * dist);
                }
                result += mass * kern;
            }
        }
    }
    return result;
}

/* tesseroid_gravity.py:305-339. out must be zero-initialised (the reference adds into result).
 * counts (may be NULL): number of leaves per (observer, tesseroid) pair, n_obs x n_tess.
 * Returns 0 or the OR of HBO_TESS_* where the reference raises OverflowError /
 * ZeroDivisionError. */
int hbo_tesseroid_loop(int field, int64_t n_obs, const double* lon, const double* lat,
                       const double* rad, int64_t n_tess, const double* tesseroids,
                       const double* density, double distance_size_ratio, int radial, double* out,
                       int64_t* counts, int nthreads)
{
    int status = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel reduction(| : status)
    {
        double* stack = (double*)malloc(sizeof(double) * 6 * STACK_SIZE);
        double* small = (double*)malloc(sizeof(double) * 6 * MAX_DISCRETIZATIONS);
#pragma omp for schedule(dynamic, 4)
        for (int64_t i = 0; i < n_obs; i++) {
            double coordinates[3] = {lon[i], lat[i], rad[i]};
            double longitude_rad = radians(lon[i]);
            double cosphi = cos(radians(lat[i])), sinphi = sin(radians(lat[i]));
            for (int64_t j = 0; j < n_tess; j++) {
                int64_t n_splits = hbo_adaptive_discretization(coordinates, tesseroids + 6 * j,
                                                               distance_size_ratio, stack, STACK_SIZE,
                                                               small, MAX_DISCRETIZATIONS, radial);
                if (n_splits < 0) {
                    status |= (int)(-n_splits);
                    n_splits = 0;
                }
                if (counts) counts[i * n_tess + j] = n_splits;
                int zero_div = 0;
                for (int64_t q = 0; q < n_splits; q++)
                    out[i] += hbo_glq_tesseroid(field, longitude_rad, cosphi, sinphi, rad[i],
                                                small + 6 * q, density[j], &zero_div);
                if (zero_div) status |= HBO_TESS_ZERO_DIVISION;
            }
        }
        free(stack);
        free(small);
    }
    return status;
}

/* ---- variable density (density is a function of the radius) ------------------------------------
 * gauss_legendre_quadrature_variable_density, _forward/_tesseroid_variable_density.py:20-106, and
 * jit_tesseroid_gravity_variable_density, _forward/tesseroid_gravity.py:342-445. The density
 * function is called back once per (latitude node, radial node) like the reference does. Serial
 * (the callback may be a Python function). */
typedef double (*hbo_density_fn)(double radius);

double hbo_glq_tesseroid_variable_density(int field, double longitude, double cosphi, double sinphi,
                                          double radius, const double* t, hbo_density_fn density,
                                          int* zero_div)
{
    double w = t[0], e = t[1], s = t[2], n = t[3], bottom = t[4], top = t[5];
    double a_factor = 1.0 / 8 * radians(e - w) * radians(n - s) * (top - bottom);
    double result = 0.0;
    for (int j = 0; j < 2; j++) {
        double latitude_p = radians(0.5 * (n - s) * GLQ_NODES[j] + 0.5 * (n + s));
        double cosphi_p = cos(latitude_p), sinphi_p = sin(latitude_p);
        for (int k = 0; k < 2; k++) {
            double radius_p = 0.5 * (top - bottom) * GLQ_NODES[k] + 0.5 * (top + bottom);
            double density_p = density(radius_p);
            double kappa = radius_p * radius_p * cosphi_p;
            for (int i = 0; i < 2; i++) {
                double longitude_p = radians(0.5 * (e - w) * GLQ_NODES[i] + 0.5 * (e + w));
                double mass = density_p * a_factor * kappa * GLQ_WEIGHTS[i] * GLQ_WEIGHTS[j] * GLQ_WEIGHTS[k];
                double cospsi;
                double dist = distance_spherical_core(longitude, cosphi, sinphi, radius, longitude_p,
                                                      cosphi_p, sinphi_p, radius_p, &cospsi);
                double kern;
                if (dist == 0.0) *zero_div = 1;
                if (field == 0) {
                    kern = 1 / dist * HBO_G;
                } else {
                    double delta_z = radius - radius_p * cospsi;
                    kern = -HBO_G * delta_z / (dist * dist * dist);
                }
                result += mass * kern;
            }
        }
    }
    return result;
}

int hbo_tesseroid_loop_variable_density(int field, int64_t n_obs, const double* lon,
                                        const double* lat, const double* rad, int64_t n_tess,
                                        const double* tesseroids, hbo_density_fn density,
                                        double distance_size_ratio, int radial, double* out)
{
    int status = 0;
    double* stack = (double*)malloc(sizeof(double) * 6 * STACK_SIZE);
    double* small = (double*)malloc(sizeof(double) * 6 * MAX_DISCRETIZATIONS);
    for (int64_t i = 0; i < n_obs; i++) {
        double coordinates[3] = {lon[i], lat[i], rad[i]};
        double longitude_rad = radians(lon[i]);
        double cosphi = cos(radians(lat[i])), sinphi = sin(radians(lat[i]));
        for (int64_t j = 0; j < n_tess; j++) {
            int64_t n_splits = hbo_adaptive_discretization(coordinates, tesseroids + 6 * j,
                                                           distance_size_ratio, stack, STACK_SIZE,
                                                           small, MAX_DISCRETIZATIONS, radial);
            if (n_splits < 0) {
                status |= (int)(-n_splits);
                n_splits = 0;
            }
            int zero_div = 0;
            for (int64_t q = 0; q < n_splits; q++)
                out[i] += hbo_glq_tesseroid_variable_density(field, longitude_rad, cosphi, sinphi, rad[i],
                                                             small + 6 * q, density, &zero_div);
            if (zero_div) status |= HBO_TESS_ZERO_DIVISION;
        }
    }
    free(stack);
    free(small);
    return status;
}
