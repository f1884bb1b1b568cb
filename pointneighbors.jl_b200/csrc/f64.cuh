// f64.cuh -- Float64 coordinates / radius.
//
// The reference is generic in the element type: with Float64 coordinates and a Float64
// search_radius every operation of Appendix A happens in Float64.  This path provides the
// searching part of the API in Float64 -- cell list build, neighbour counts, neighbour lists
// (PrecomputedNeighborhoodSearch, arbitrary closures) and the pair geometry -- with the same
// bit-exactness contract (explicit __d*_rn intrinsics in the reference's operation order).  It
// uses straightforward kernels (one thread per point / query), not the tile machinery of the
// Float32 benchmark path; the fused SPH closures are Float32 only.
#pragma once

#include "grid.cuh"

namespace pnb {

#ifdef __CUDACC__
// floor_to_int((x - min) / cs) + 1 with Julia's saturation (src/util.jl:19-34), then the periodic
// wrap mod(c - 2, n) + 2 (src/nhs_grid.jl:619) in wrapping Int64 arithmetic.
__device__ __forceinline__ int cell_coord64(double x, double minc, double cs, int periodic, int nc)
{
    const double f = floor(__ddiv_rn(__dsub_rn(x, minc), cs));
    if (fabs(f) < 1073741824.0) {
        int c = (int)f + 1;
        if (periodic) c = floormod_i(c - 2, nc) + 2;
        return c;
    }
    if (!periodic) return f < 0.0 ? -1 : 0x7fffffff;    // NaN lands here too: out of bounds
    long long c;
    if (isnan(f) || f >= 9223372036854775808.0) c = 0x7fffffffffffffffLL;
    else if (f <= -9223372036854775808.0) c = (long long)0x8000000000000000ULL;
    else c = (long long)f;
    unsigned long long u = (unsigned long long)c + 1ULL;
    u -= 2ULL;
    long long m = (long long)u % (long long)nc;
    if (m < 0) m += nc;
    return (int)m + 2;
}

template <int ND>
__device__ __forceinline__ int point_cell64(const GridP64 &g, const double *p, int *cc)
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        cc[d] = cell_coord64(p[d], g.minc[d], g.cs[d], g.periodic, g.nc[d]);
        ok = ok && cc[d] >= 2 && cc[d] <= g.gs[d] - 1;
    }
#pragma unroll
    for (int d = ND; d < 3; d++) cc[d] = 1;
    if (!ok) return -1;
    return (cc[0] - 1) + (cc[1] - 1) * g.gs[0] + (cc[2] - 1) * g.gs[0] * g.gs[1];
}

// pos_diff = x_i - y_j, d2 left to right, periodic fix only when d2 > r2
// (src/nhs_grid.jl:547-555, src/neighborhood_search.jl:428-435), all in Float64
template <int ND>
__device__ __forceinline__ double pair_d2_64(const GridP64 &g, const double *xi, const Rec64 &yj,
                                             double *p, bool radius_test_fix)
{
    if (g.mixed) {
        // mixed precision (tut_gpu_usage.jl:45-50): Float64 subtraction, ONE conversion, then the
        // Float32 operation sequence of the Float32 path; the results are returned widened
        float q[3];
        q[0] = __double2float_rn(__dsub_rn(xi[0], yj.x));
        q[1] = ND > 1 ? __double2float_rn(__dsub_rn(xi[1], yj.y)) : 0.f;
        q[2] = ND > 2 ? __double2float_rn(__dsub_rn(xi[2], yj.z)) : 0.f;
        float d2f = dist2<ND>(q[0], q[1], q[2]);
        if (g.periodic && radius_test_fix && d2f > (float)g.r2) {
#pragma unroll
            for (int d = 0; d < ND; d++) {
                const float bs = (float)g.bsize[d];
                q[d] = __fsub_rn(q[d], __fmul_rn(bs, rintf(__fdiv_rn(q[d], bs))));
            }
            d2f = dist2<ND>(q[0], q[1], q[2]);
        }
        p[0] = (double)q[0]; p[1] = (double)q[1]; p[2] = (double)q[2];
        return (double)d2f;
    }
    p[0] = __dsub_rn(xi[0], yj.x);
    p[1] = ND > 1 ? __dsub_rn(xi[1], yj.y) : 0.0;
    p[2] = ND > 2 ? __dsub_rn(xi[2], yj.z) : 0.0;
    double d2 = __dmul_rn(p[0], p[0]);
    if (ND > 1) d2 = __dadd_rn(d2, __dmul_rn(p[1], p[1]));
    if (ND > 2) d2 = __dadd_rn(d2, __dmul_rn(p[2], p[2]));
    if (g.periodic && radius_test_fix && d2 > g.r2) {
#pragma unroll
        for (int d = 0; d < ND; d++)
            p[d] = __dsub_rn(p[d], __dmul_rn(g.bsize[d], rint(__ddiv_rn(p[d], g.bsize[d]))));
        d2 = __dmul_rn(p[0], p[0]);
        if (ND > 1) d2 = __dadd_rn(d2, __dmul_rn(p[1], p[1]));
        if (ND > 2) d2 = __dadd_rn(d2, __dmul_rn(p[2], p[2]));
    }
    return d2;
}

// One thread per query point: the 3^d stencil in CartesianIndices order, ids ascending inside a
// cell (the build leaves the cells in canonical order).
//   MODE 0: out_count[i] (int64) = number of neighbours          (count_neighbors.jl:24-27)
//   MODE 1: list_count[i] (uint32)                               (list build, count pass)
//   MODE 2: ids[offsets[i] ...] = neighbours                     (list build, fill pass)
template <int ND, int MODE>
__global__ void __launch_bounds__(128)
k_sweep_points64(GridP64 g, const uint32_t *__restrict__ cell_start,
                 const Rec64 *__restrict__ sorted, const double *__restrict__ x, int64_t n_loop,
                 const int32_t *__restrict__ points, int base, int64_t *__restrict__ out_count,
                 uint32_t *__restrict__ list_count, const int64_t *__restrict__ offsets,
                 int32_t *__restrict__ ids, int *__restrict__ err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_loop) return;
    const int64_t i = points ? (int64_t)points[t] - base : t;
    double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) xi[d] = x[i * ND + d];
    int cc[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < ND; d++) cc[d] = cell_coord64(xi[d], g.minc[d], g.cs[d], g.periodic, g.nc[d]);
    int64_t cnt = 0;
    int64_t pos = MODE == 2 ? offsets[i] : 0;
    bool oob = false;
    for (int oz = (ND > 2 ? -1 : 0); oz <= (ND > 2 ? 1 : 0); oz++)
        for (int oy = (ND > 1 ? -1 : 0); oy <= (ND > 1 ? 1 : 0); oy++)
            for (int ox = -1; ox <= 1; ox++) {
                int c0 = cc[0] + ox, c1 = cc[1] + oy, c2 = cc[2] + oz;
                if (g.periodic) {
                    c0 = floormod_i(c0 - 2, g.nc[0]) + 2;
                    if (ND > 1) c1 = floormod_i(c1 - 2, g.nc[1]) + 2;
                    if (ND > 2) c2 = floormod_i(c2 - 2, g.nc[2]) + 2;
                }
                if (c0 < 1 || c0 > g.gs[0] || c1 < 1 || c1 > g.gs[1] || c2 < 1 || c2 > g.gs[2]) {
                    oob = true;   // the safe variant's bounds check (nhs_grid.jl:530-532)
                    continue;
                }
                const int lin = (c0 - 1) + (c1 - 1) * g.gs[0] + (c2 - 1) * g.gs[0] * g.gs[1];
                const uint32_t b0 = cell_start[lin], b1 = cell_start[lin + 1];
                for (uint32_t k = b0; k < b1; k++) {
                    const Rec64 yj = sorted[k];
                    double p[3];
                    const double d2 = pair_d2_64<ND>(g, xi, yj, p, true);
                    if (d2 <= g.r2) {
                        if (MODE == 2) ids[pos++] = (int32_t)yj.id;
                        else cnt++;
                    }
                }
            }
    if (oob) atomicOr(err, 2);
    if (MODE == 0) out_count[i] = cnt;
    if (MODE == 1) list_count[i] = (uint32_t)cnt;
}
#endif  // __CUDACC__

}  // namespace pnb
