// hb200_tess_leaves.cuh -- tesseroids whose density is a FUNCTION of the radius, with the 3-D
// (radial) adaptive discretisation.
//
// Reference: jit_tesseroid_gravity_variable_density (_forward/tesseroid_gravity.py:342-445) calls
// density(radius_p) at the radial quadrature nodes of every LEAF of the discretisation
// (_tesseroid_variable_density.py:55-58). With the radial direction discretised too the leaves
// have their own radial bounds, so those radii are only known once a pair has been walked, and
// the density function is a Python callable that cannot be called from a kernel. Three steps per
// batch of observers:
//   1. tesseroid_root_kernel (hb200_tess.cuh): the far field of all pairs; a pair whose root does
//      not split is integrated there with the two densities of its root (evaluated by the host
//      beforehand, like for the horizontal discretisation); pairs that split are listed;
//   2. tesseroid_collect_kernel: one thread per list walks its pairs depth-first, exactly like the
//      reference (stack and leaf limits included), and APPENDS every leaf -- observer, bounds and
//      the two radii radius_p -- to a buffer; the host hands the radii to the density callback;
//   3. tesseroid_leaf_kernel: one thread per leaf integrates it with the reference's statements
//      (tess_nodes / tess_glq_nodes) and the two densities it got back, and adds it to its
//      observer's sum (float64 atomics: the leaves of an observer are added in no fixed order, the
//      one place in the library where a result is not bit-reproducible from run to run).
#pragma once
#include "hb200_tess.cuh"

namespace hb {

#if defined(__CUDACC__)
struct LeafBuf {
    int* obs;        // [cap] observer (index inside the batch)
    double* bounds;  // [6][cap] w e s n bottom top
    double* radii;   // [2][cap] radius_p of the two radial quadrature nodes
    int* count;      // leaves appended (may run past cap: the host then retries with fewer observers)
    int cap;
};

// radius_p of gauss_legendre_quadrature_variable_density (:55-57), as tess_nodes computes it
__device__ __forceinline__ double leaf_radial_node(double bottom, double top, int k)
{
    const double node = k ? kGlqNode : -kGlqNode;
    return 0.5 * (top - bottom) * node + 0.5 * (top + bottom);
}

template <class TRIG>
__global__ void __launch_bounds__(kTessBlock) tesseroid_collect_kernel(const TessArgs a,
                                                                       const unsigned short* list,
                                                                       const int* count, LeafBuf L)
{
    double stack[kTessStack * 6];
    const int64_t i = (int64_t)blockIdx.x * kTessBlock + threadIdx.x;
    if (i >= a.n_obs) return;
    const int n = count[(int64_t)blockIdx.y * a.n_obs + i];
    int resume = count[((int64_t)gridDim.y + blockIdx.y) * a.n_obs + i];
    const int64_t begin = (int64_t)blockIdx.y * a.chunk_len;
    const int chunk_cnt = (int)((begin + a.chunk_len < a.n_src ? begin + a.chunk_len : a.n_src) - begin);
    if (n == 0 && resume >= chunk_cnt) return;
    TessObs o;
    tess_make_obs(o, a.lon[i], a.lat[i], a.rad[i]);
    const unsigned short* my_list = list + ((int64_t)blockIdx.y * kTessListCap) * a.n_obs + i;
    const bool radial = a.radial != 0;
    unsigned flags = 0;
    int k = 0, stack_top = -1, n_leaves = 0;
    while (true) {
        if (stack_top < 0) {  // the next pair: the listed ones, then the rest of the chunk
            int off;
            if (k < n) off = my_list[(int64_t)(k++) * a.n_obs];
            else if (resume < chunk_cnt) off = resume++;
            else break;
            const double* rec = a.packed + (begin + off) * kTessRec;
            if (rec[29] != 0.0) {  // a zero dimension: numba's ZeroDivisionError
                flags |= FLAG_ZERO_DIV;
                continue;
            }
#pragma unroll
            for (int c = 0; c < 6; c++) stack[c] = rec[c];
            stack_top = 0;
            n_leaves = 0;
        }
        // one pop of _adaptive_discretization (_tesseroid_utils.py:178-216), as tess_walk_step
        const double* q = stack + 6 * stack_top;
        const double w = q[0], e = q[1], s = q[2], nn = q[3], bottom = q[4], top = q[5];
        stack_top -= 1;
        int n_lon, n_lat, n_rad;
        if (!tess_classify<TRIG>(o, a.ratio, radial, w, e, s, nn, bottom, top, n_lon, n_lat, n_rad)) {
            flags |= FLAG_ZERO_DIV;
            stack_top = -1;
            continue;
        }
        const int kids = n_lon * n_lat * n_rad;
        if (kids > 1) {
            if ((stack_top + 1) + kids > kTessStack) {
                flags |= FLAG_TESS_STACK;
                stack_top = -1;
                continue;
            }
            const double d_lon = (e - w) / n_lon, d_lat = (nn - s) / n_lat, d_rad = (top - bottom) / n_rad;
            for (int x = 0; x < n_lon; x++)
                for (int y = 0; y < n_lat; y++)
                    for (int z = 0; z < n_rad; z++) {
                        stack_top += 1;
                        double* c = stack + 6 * stack_top;
                        c[0] = w + d_lon * x;
                        c[1] = w + d_lon * (x + 1);
                        c[2] = s + d_lat * y;
                        c[3] = s + d_lat * (y + 1);
                        c[4] = bottom + d_rad * z;
                        c[5] = bottom + d_rad * (z + 1);
                    }
        } else {
            if (n_leaves + 1 > kTessMaxLeaves) {
                flags |= FLAG_TESS_LEAVES;
                stack_top = -1;
                continue;
            }
            n_leaves += 1;
            const int pos = atomicAdd(L.count, 1);
            if (pos < L.cap) {
                L.obs[pos] = (int)i;
                L.bounds[pos] = w;
                L.bounds[(int64_t)L.cap + pos] = e;
                L.bounds[2 * (int64_t)L.cap + pos] = s;
                L.bounds[3 * (int64_t)L.cap + pos] = nn;
                L.bounds[4 * (int64_t)L.cap + pos] = bottom;
                L.bounds[5 * (int64_t)L.cap + pos] = top;
                L.radii[pos] = leaf_radial_node(bottom, top, 0);
                L.radii[(int64_t)L.cap + pos] = leaf_radial_node(bottom, top, 1);
            }
        }
    }
    if (flags && a.flags) atomicOr(a.flags, flags);
}

// density: [2][cap], the callback's values at L.radii
template <int FIELD, class TRIG>
__global__ void __launch_bounds__(128) tesseroid_leaf_kernel(const TessArgs a, const LeafBuf L, int n_leaves,
                                                             const double* __restrict__ density,
                                                             double* leaf_sum)
{
    const int p = blockIdx.x * 128 + threadIdx.x;
    if (p >= n_leaves) return;
    const int i = L.obs[p];
    TessObs o;
    tess_make_obs(o, a.lon[i], a.lat[i], a.rad[i]);
    const int64_t cap = L.cap;
    const double rho[2] = {density[p], density[cap + p]};
    TessNodes nodes;
    tess_nodes<TRIG>(nodes, L.bounds[p], L.bounds[cap + p], L.bounds[2 * cap + p], L.bounds[3 * cap + p],
                     L.bounds[4 * cap + p], L.bounds[5 * cap + p], rho);
    unsigned flags = 0;
    const double value = tess_glq_nodes<FIELD, TRIG>(o, nodes, flags);
    atomicAdd(&leaf_sum[i], value);
    if (flags && a.flags) atomicOr(a.flags, flags);
}
#endif  // __CUDACC__

}  // namespace hb
