/*
 * harmonica_b200.h -- C ABI of libharmonica_b200.so
 *
 * The drop-in boundary for harmonica's pairwise source->observer forward
 * models. The reference (fatiando/harmonica) has no FFI of its own on this
 * path: its hot loops are Numba-jitted Python. Each entry point below replaces
 * one jitted loop (plus the choclo kernel it calls) and is what the
 * reference's Python wrapper would bind through ctypes (INTEGRATION.md shows
 * the stub). Citations are file:line under /root/reference/src/harmonica.
 *
 * Conventions
 *  - plain pointers and sizes only; all arrays are C-contiguous float64
 *  - host entry points (no suffix): caller-owned HOST buffers, blocking; the
 *    library owns device memory, streams and peer access. Work is sharded over
 *    the devices selected by hb200_init() by observation points (shard_mode 1,
 *    default) or by sources with a peer reduce-sum (shard_mode 2).
 *  - *_dev entry points: caller-owned DEVICE buffers on the current device,
 *    asynchronous on `stream` (a cudaStream_t passed as void*); workspace is
 *    caller-provided (query with the matching *_ws_bytes).
 *  - outputs are field-major: out[k * n_obs + i] for the k-th requested field
 *    in ascending field-id order, ALREADY in harmonica's units and sign
 *    convention (mGal / Eotvos / nT, z down for g_z, g_ez, g_nz).
 *  - return value: HB200_OK or a negative HB200_E*; text via hb200_last_error()
 *  - *flags (may be NULL) receives an OR of HB200_FLAG_*
 */
#ifndef HARMONICA_B200_H
#define HARMONICA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB200_OK 0
#define HB200_EINVAL (-1)  /* bad argument */
#define HB200_ECUDA (-2)   /* CUDA runtime error, see hb200_last_error() */
#define HB200_ENODEV (-3)  /* no usable sm_100 device */
#define HB200_ENOMEM (-4)
#define HB200_EZERODIV (-5) /* an observation point coincides with a source (entry points without
                             * a flags argument: the Jacobians and the fits); the reference's
                             * jitted loops raise ZeroDivisionError there */

/* field ids (bit positions of field_mask). Names follow harmonica's FIELDS
 * dict, _forward/prisms/gravity.py:37-48. */
enum hb200_field {
    HB200_POTENTIAL = 0, HB200_G_E = 1, HB200_G_N = 2, HB200_G_Z = 3,
    HB200_G_EE = 4, HB200_G_NN = 5, HB200_G_ZZ = 6,
    HB200_G_EN = 7, HB200_G_EZ = 8, HB200_G_NZ = 9
};
#define HB200_MASK_ACCEL 0x00Eu  /* g_e | g_n | g_z, one fused pass */
#define HB200_MASK_TENSOR 0x3F0u /* the six tensor components, one fused pass */

/* magnetic component mask, _forward/prisms/magnetic.py:20-25 */
#define HB200_B_E 1u
#define HB200_B_N 2u
#define HB200_B_U 4u
#define HB200_B_ALL 7u /* field="b": one fused pass */

/* behaviour switches of the magnetic kernels (choclo version dependent) */
#define HB200_MAG_NAN_ON_EDGES 1u
#define HB200_MAG_FACE_OUTSIDE_LIMIT 2u
#define HB200_MAG_DEFAULT 3u

#define HB200_FLAG_SINGULAR 1u /* an observer sits on a singular point of a prism */
#define HB200_FLAG_ZERO_DIV 2u /* observer coincides with a point source */
#define HB200_FLAG_TESS_STACK 4u  /* tesseroids: the reference's "Stack Overflow" OverflowError */
#define HB200_FLAG_TESS_LEAVES 8u /* tesseroids: "Exceeded maximum discretizations" */
#define HB200_FLAG_TESS_INSIDE 16u /* tesseroids: a computation point lies inside a tesseroid */

#define HB200_SHARD_AUTO 0
#define HB200_SHARD_OBSERVERS 1
#define HB200_SHARD_SOURCES 2

/* ---- library / device management ---------------------------------------- */
int hb200_version(void);
/* number of visible CUDA devices with compute capability 10.x */
int hb200_device_count(void);
/* select the devices used by the host entry points; devices == NULL selects
 * all visible ones. Idempotent; called lazily with NULL if never called. */
int hb200_init(const int* devices, int n_devices);
int hb200_num_devices(void);
void hb200_shutdown(void);
const char* hb200_last_error(void);
/* variant 0 = rule-exact direct evaluation for every pair; 1 = merged
 * transcendentals (CUDA libm) with the direct path on pairs whose observer lies
 * on (the extension of) a prism edge or vertex; 2 (default) = as 1 with the
 * library's own division / sqrt / log / atan2 sequences (hb200_xmath.cuh) */
int hb200_set_variant(int variant);
int hb200_get_variant(void);
/* tesseroid kernels: 1 = observer-independent parts of every tesseroid precomputed into root
 * records, pairs that split deferred and walked by all lanes of a warp together; 2 = as 1 with an
 * arithmetic-only far field (cosine of the longitude difference from precomputed factors, the
 * library's reciprocal square root, squared split thresholds); 3 = as 2 with the library's own
 * bounded-angle sin / cos / acos in the walks; 4, 5 = as 3 compiled for 80 / 64 registers
 * (occupancy experiments); 6 (7, 8: other register budgets) = the far field of all pairs in one
 * kernel that lists the pairs that split, their walks in a second kernel (one thread per list);
 * 9 (default) = as 6 with the walks done by groups of 16 lanes on a shared stack, lists drawn
 * from a compacted work list by a persistent kernel; 0 = first build (every pair walked where
 * it is met) */
int hb200_set_tesseroid_variant(int variant);
int hb200_get_tesseroid_variant(void);
/* staging of the packed prism records in the prism kernels: 1 (default) = every warp streams
 * its own copy with TMA bulk copies and synchronises only with itself, and in prism_layer_gravity
 * neighbouring prisms of a layer column share two vertex distances; 2 = as 1 without that reuse;
 * 0 = one copy per CTA with a CTA-wide barrier per tile (first build, kept for comparison).
 * Values are identical. */
int hb200_set_tile_mode(int mode);
int hb200_get_tile_mode(void);
/* Reproducibility. The value of an (observer, source) pair never depends on the batch it is
 * computed in. What does depend on the call is the ASSOCIATION of the sum over sources: with few
 * observers the source list is split into chunks over grid.y (to fill the GPU) whose partial sums
 * are added in a fixed order. chunks > 0 pins the number of chunks (1 = one sequential sum per
 * observer, like the reference's loop): results are then bit-identical under any batching,
 * permutation or sharding of the observers. 0 (default): chosen from the grid size. */
int hb200_set_source_chunks(int chunks);
int hb200_get_source_chunks(void);
/* relative singular-value cutoff of the UNDAMPED equivalent-sources fit (damping = NaN): values
 * below rcond * s_max are dropped, like scipy.linalg.lstsq(cond=rcond) behind sklearn's
 * LinearRegression. Default: machine epsilon (cond=None, scikit-learn < 1.7, the behaviour the
 * reference's own tests pin); 1e-6 reproduces LinearRegression(tol=1e-6) of scikit-learn >= 1.7. */
int hb200_set_fit_rcond(double rcond);
double hb200_get_fit_rcond(void);
/* number of CUDA kernels this library has launched so far (all entry points) */
uint64_t hb200_launch_count(void);

/* ---- host-buffer entry points ------------------------------------------- */
/* replaces jit_prism_gravity, _forward/prisms/gravity.py:489-545, with the
 * sign/unit post-scaling of :227-235 folded into the epilogue. prisms is
 * (n_prisms, 6) row-major [w, e, s, n, bottom, top]. */
int hb200_prism_gravity(const double* easting, const double* northing, const double* upward,
                        int64_t n_obs, const double* prisms, const double* density,
                        int64_t n_prisms, uint32_t field_mask, int shard_mode, double* out,
                        uint32_t* flags);

/* replaces _any_singular_point_g_*, _forward/prisms/gravity.py:272-449:
 * *flags gets HB200_FLAG_SINGULAR if any pair is singular for `field`. */
int hb200_prism_singular_scan(const double* easting, const double* northing,
                              const double* upward, int64_t n_obs, const double* prisms,
                              int64_t n_prisms, int field, uint32_t* flags);

/* replaces _jit_prism_magnetic_field / _jit_prism_magnetic_component,
 * _forward/prisms/magnetic.py:275-400, output in nT (:195-197, :271). */
int hb200_prism_magnetic(const double* easting, const double* northing, const double* upward,
                         int64_t n_obs, const double* prisms, const double* mag_e,
                         const double* mag_n, const double* mag_u, int64_t n_prisms,
                         uint32_t component_mask, uint32_t rules, int shard_mode, double* out,
                         uint32_t* flags);

/* replaces _forward_gravity_prism_layer, _forward/prisms/layer.py:522-633.
 * bottom/top/density are (n_north, n_east) row-major; prisms are visited
 * easting-outer / northing-inner like the reference. */
int hb200_prism_layer_gravity(const double* easting, const double* northing,
                              const double* upward, int64_t n_obs, const double* prisms_easting,
                              int64_t n_east, const double* prisms_northing, int64_t n_north,
                              const double* bottom, const double* top, const double* density,
                              double thickness_threshold, uint32_t field_mask, int shard_mode,
                              double* out, uint32_t* flags);

/* replaces point_mass_cartesian / point_mass_spherical,
 * _forward/point.py:357-454. spherical != 0: coordinates are longitude,
 * latitude (degrees), radius; only HB200_POTENTIAL and HB200_G_Z exist. */
int hb200_point_gravity(const double* easting, const double* northing, const double* upward,
                        int64_t n_obs, const double* src_easting, const double* src_northing,
                        const double* src_upward, const double* masses, int64_t n_src,
                        uint32_t field_mask, int spherical, int shard_mode, double* out,
                        uint32_t* flags);

/* replaces predict, _equivalent_sources/utils.py:77-101 with
 * greens_func_cartesian (cartesian.py:634-644): out[i] = sum_j coefs[j] / dist */
int hb200_eqs_predict(const double* easting, const double* northing, const double* upward,
                      int64_t n_obs, const double* src_easting, const double* src_northing,
                      const double* src_upward, const double* coefs, int64_t n_src,
                      int shard_mode, double* out, uint32_t* flags);

/* replaces predict with greens_func_spherical (EquivalentSourcesSph.predict,
 * _equivalent_sources/spherical.py:219-248, 412-424): longitude, latitude in
 * degrees, radius in metres; out[i] = sum_j coefs[j] / distance_spherical */
int hb200_eqs_predict_spherical(const double* longitude, const double* latitude,
                                const double* radius, int64_t n_obs, const double* src_longitude,
                                const double* src_latitude, const double* src_radius,
                                const double* coefs, int64_t n_src, int shard_mode, double* out,
                                uint32_t* flags);

/* replaces _jit_dipole_magnetic_field_cartesian / _jit_dipole_magnetic_component_cartesian,
 * _forward/dipole.py:292-415 (choclo.dipole.magnetic_*), output in nT. component_mask as for
 * hb200_prism_magnetic (HB200_B_ALL = one fused pass). */
int hb200_dipole_magnetic(const double* easting, const double* northing, const double* upward,
                          int64_t n_obs, const double* src_easting, const double* src_northing,
                          const double* src_upward, const double* moment_e, const double* moment_n,
                          const double* moment_u, int64_t n_src, uint32_t component_mask,
                          int shard_mode, double* out, uint32_t* flags);

/* replaces jit_tesseroid_gravity, _forward/tesseroid_gravity.py:236-339 (constant densities):
 * adaptive discretisation (_tesseroid_utils.py:136-217, STACK_SIZE 100, MAX_DISCRETIZATIONS
 * 100000, distance-size ratio 1 for the potential and 2.5 for g_z) + 2x2x2 Gauss-Legendre
 * point masses (:19-107). longitude, latitude in degrees, radius in metres; tesseroids is
 * (n_tesseroids, 6) row-major [w, e, s, n, bottom, top]; field is HB200_POTENTIAL (J/kg) or
 * HB200_G_Z (downward, mGal, tesseroid_gravity.py:222-225). *flags gets HB200_FLAG_TESS_*
 * where the reference raises OverflowError and HB200_FLAG_ZERO_DIV where its jitted loop raises
 * ZeroDivisionError (a tesseroid dimension or a node distance of exactly zero). */
int hb200_tesseroid_gravity(const double* longitude, const double* latitude, const double* radius,
                            int64_t n_obs, const double* tesseroids, const double* density,
                            int64_t n_tesseroids, int field, int radial_adaptive_discretization,
                            int shard_mode, double* out, uint32_t* flags);

/* replaces jit_tesseroid_gravity_variable_density, _forward/tesseroid_gravity.py:342-445, for the
 * horizontal (default) adaptive discretisation: the leaves of a tesseroid then keep its radial
 * bounds, so the density function is only ever evaluated at the tesseroid's two radial
 * Gauss-Legendre nodes (_tesseroid_variable_density.py:55-58). The host evaluates it there
 * (after density_based_discretization, :108-157) and passes density_lower / density_upper
 * (n_tesseroids each). Everything else as hb200_tesseroid_gravity. */
int hb200_tesseroid_gravity_variable_density(const double* longitude, const double* latitude,
                                             const double* radius, int64_t n_obs,
                                             const double* tesseroids, const double* density_lower,
                                             const double* density_upper, int64_t n_tesseroids,
                                             int field, int shard_mode, double* out,
                                             uint32_t* flags);

/* replaces jit_tesseroid_gravity_variable_density, _forward/tesseroid_gravity.py:342-445, with the
 * 3-D (radial) adaptive discretisation: the leaves then have their own radial bounds, and the
 * density function is wanted at the two radial quadrature nodes of EVERY leaf
 * (_tesseroid_variable_density.py:55-58). The library walks the pairs, hands the radii of all
 * leaves of a batch of computation points to `density` (a host function: density_out[i] =
 * density(radius[i]), i < n; called a few times per batch from the calling thread) and integrates
 * the leaves with the values it gets back. density_lower / density_upper are the values at the
 * two radial nodes of every (radially pre-split) tesseroid, used for the pairs whose root does
 * not split. Runs on the first selected device. The leaves of a computation point are added with
 * float64 atomics (last-bit differences from run to run). */
typedef void (*hb200_density_fn)(const double* radius, double* density_out, int64_t n, void* user);
int hb200_tesseroid_gravity_density_function(const double* longitude, const double* latitude,
                                             const double* radius, int64_t n_obs,
                                             const double* tesseroids, const double* density_lower,
                                             const double* density_upper, int64_t n_tesseroids,
                                             int field, hb200_density_fn density, void* user,
                                             double* out, uint32_t* flags);

/* replaces _check_points_outside_tesseroids, _forward/_tesseroid_utils.py:431-454, as one
 * pass: *flags gets HB200_FLAG_TESS_INSIDE if any computation point lies strictly inside any
 * tesseroid (the host then lists the pairs for the reference's error message). */
int hb200_tesseroid_inside_scan(const double* longitude, const double* latitude,
                                const double* radius, int64_t n_obs, const double* tesseroids,
                                int64_t n_tesseroids, uint32_t* flags);

/* replaces jacobian, _equivalent_sources/utils.py:54-74: jac[i*n_src+j] = 1/dist */
int hb200_eqs_jacobian(const double* easting, const double* northing, const double* upward,
                       int64_t n_obs, const double* src_easting, const double* src_northing,
                       const double* src_upward, int64_t n_src, double* jac);

/* the same matrix with greens_func_spherical (EquivalentSourcesSph.jacobian,
 * _equivalent_sources/spherical.py:249-283, 412-424): longitude, latitude in degrees */
int hb200_eqs_jacobian_spherical(const double* longitude, const double* latitude,
                                 const double* radius, int64_t n_obs, const double* src_longitude,
                                 const double* src_latitude, const double* src_radius,
                                 int64_t n_src, double* jac);

/* replaces the body of EquivalentSources.fit, _equivalent_sources/cartesian.py:277-280
 * (EquivalentSourcesSph.fit, spherical.py:213-215): Jacobian + verde.base.least_squares,
 * entirely on the device (device 0 of hb200_init). The Jacobian never leaves HBM; its
 * columns are scaled by their standard deviation, rows by sqrt(weights) (weights may be
 * NULL), then
 *   damping given (not NaN): (X'X + damping I) c = X'y by Cholesky, like sklearn's Ridge
 *       (dual form when n_src > n_obs; SVD ridge filter if the factorisation fails);
 *   damping = NaN ("None"): minimum-norm least squares from the SVD with singular values
 *       below rcond * s_max dropped (hb200_set_fit_rcond), like scipy.linalg.lstsq behind
 *       sklearn's LinearRegression.
 * coefs (n_src) receives the unscaled coefficients. *solver_path (may be NULL): 0 Cholesky,
 * 1 SVD pseudo-inverse, 2 SVD ridge filter. Dense algebra: cuBLAS / cuSOLVER, loaded on
 * first use. */
int hb200_eqs_fit(const double* easting, const double* northing, const double* upward,
                  int64_t n_obs, const double* src_easting, const double* src_northing,
                  const double* src_upward, int64_t n_src, const double* data,
                  const double* weights, double damping, int spherical, double* coefs,
                  int* solver_path);

/* replaces EquivalentSourcesGB._gradient_boosting, _equivalent_sources/
 * gradient_boosted.py:244-293: for every window (in the given order) fit the window's sources
 * to the residue of the window's data points (as hb200_eqs_fit), predict their field on ALL
 * data points, update the residue and add the coefficients. Window w owns
 * src_index[src_offset[w] .. src_offset[w+1]) and data_index[data_offset[w] .. ), both
 * non-empty with distinct entries. coefs (n_src) receives the summed coefficients, rmse
 * (n_windows + 1) the reference's rmse_per_iteration_. Everything between the upload of the
 * inputs and the download of coefs / rmse runs on the device. */
int hb200_eqs_fit_gb(const double* easting, const double* northing, const double* upward,
                     int64_t n_obs, const double* src_easting, const double* src_northing,
                     const double* src_upward, int64_t n_src, const double* data,
                     const double* weights, double damping, int spherical, int64_t n_windows,
                     const int64_t* src_index, const int64_t* src_offset,
                     const int64_t* data_index, const int64_t* data_offset, double* coefs,
                     double* rmse);

/* ---- device-buffer entry points (current device, async on stream) -------- */
size_t hb200_prism_ws_bytes(int64_t n_obs, int64_t n_sources, int n_fields);
size_t hb200_point_ws_bytes(int64_t n_obs, int64_t n_sources);
size_t hb200_tesseroid_ws_bytes(int64_t n_obs, int64_t n_tesseroids);
int hb200_tesseroid_gravity_dev(const double* longitude, const double* latitude, const double* radius,
                                int64_t n_obs, const double* tesseroids, const double* density,
                                int64_t n_tesseroids, int field, int radial_adaptive_discretization,
                                double* out, uint32_t* flags_dev, void* ws, size_t ws_bytes,
                                void* stream);
int hb200_prism_gravity_dev(const double* easting, const double* northing, const double* upward,
                            int64_t n_obs, const double* prisms, const double* density,
                            int64_t n_prisms, uint32_t field_mask, double* out,
                            uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream);
int hb200_prism_magnetic_dev(const double* easting, const double* northing,
                             const double* upward, int64_t n_obs, const double* prisms,
                             const double* mag_e, const double* mag_n, const double* mag_u,
                             int64_t n_prisms, uint32_t component_mask, uint32_t rules,
                             double* out, uint32_t* flags_dev, void* ws, size_t ws_bytes,
                             void* stream);
int hb200_prism_layer_gravity_dev(const double* easting, const double* northing,
                                  const double* upward, int64_t n_obs,
                                  const double* prisms_easting, int64_t n_east,
                                  const double* prisms_northing, int64_t n_north,
                                  const double* bottom, const double* top, const double* density,
                                  double thickness_threshold, uint32_t field_mask, double* out,
                                  uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream);
/* weights = masses (scale_by_G != 0) or EQS coefficients (scale_by_G == 0,
 * field_mask must then be the potential bit: 1/dist) */
int hb200_point_gravity_dev(const double* easting, const double* northing, const double* upward,
                            int64_t n_obs, const double* src_easting, const double* src_northing,
                            const double* src_upward, const double* weights, int64_t n_src,
                            uint32_t field_mask, int spherical, int scale_by_G, double* out,
                            uint32_t* flags_dev, void* ws, size_t ws_bytes, void* stream);

/* FP64 FMA issue-rate microbenchmark on the current device: runs `iters`
 * dependent-chain rounds of 8 independent DFMA chains per thread on every SM
 * and returns the achieved FP64 FLOP/s (2 flop per DFMA) in *flops. */
int hb200_fp64_peak(int iters, double* flops, double* seconds);

#ifdef __cplusplus
}
#endif
#endif /* HARMONICA_B200_H */
