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

// Fused closures in Float64 (the reference is generic in the element type and publishes Float64
// WCSPH numbers, benchmarks/plot_benchmarks.jl:67): one thread per query point, candidates in the
// reference's order (3^d cells in CartesianIndices order, ids ascending inside a cell), every
// operation an explicit IEEE double operation in the oracle's order (pno_cl_nbody / pno_cl_wcsph
// instantiated for double) -> sums bit-identical to the Float64 oracle.  Per-point state is read
// by id straight from the caller's arrays.
struct Wcsph64P {
    double h, sound_speed, alpha, beta, epsilon, delta, kernel_norm;
};
//   KIND 0: n-body   (benchmarks/n_body.jl:38-48)            state: mass[n], out dv[nx x nd]
//   KIND 1: WCSPH    (smoothed_particle_hydrodynamics.jl:45-102)  state: v (nd+1), mass, pressure
template <int ND, int KIND>
__global__ void __launch_bounds__(128)
k_sweep_closure64(GridP64 g, const uint32_t *__restrict__ cell_start,
                  const Rec64 *__restrict__ sorted, const double *__restrict__ x, int64_t n_loop,
                  const int32_t *__restrict__ points, int base, const double *__restrict__ v_x,
                  const double *__restrict__ v_y, const double *__restrict__ mass_y,
                  const double *__restrict__ p_x, const double *__restrict__ p_y, double G,
                  Wcsph64P prm, double *__restrict__ dv, int *__restrict__ err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_loop) return;
    const int64_t i = points ? (int64_t)points[t] - base : t;
    constexpr int NS = ND + 1;
    const double sqrt_eps = 1.4901161193847656e-8;
    double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) xi[d] = x[i * ND + d];
    int cc[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < ND; d++) cc[d] = cell_coord64(xi[d], g.minc[d], g.cs[d], g.periodic, g.nc[d]);
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double va[4] = {0.0, 0.0, 0.0, 0.0}, p_a = 0.0;
    if (KIND == 1) {
#pragma unroll
        for (int k = 0; k < NS; k++) va[k] = v_x[i * NS + k];
        p_a = p_x[i];
    }
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
                    oob = true;
                    continue;
                }
                const int lin = (c0 - 1) + (c1 - 1) * g.gs[0] + (c2 - 1) * g.gs[0] * g.gs[1];
                const uint32_t b0 = cell_start[lin], b1 = cell_start[lin + 1];
                for (uint32_t kk = b0; kk < b1; kk++) {
                    const Rec64 yj = sorted[kk];
                    double p[3];
                    const double d2 = pair_d2_64<ND>(g, xi, yj, p, true);
                    if (!(d2 <= g.r2)) continue;
                    const double d = __dsqrt_rn(d2);
                    const int64_t j = yj.id;
                    if (KIND == 0) {
                        if (d < sqrt_eps) continue;
                        const double tt = __dmul_rn(-G, mass_y[j]);
                        const double d3 = __dmul_rn(__dmul_rn(d, d), d);
#pragma unroll
                        for (int k = 0; k < ND; k++)
                            acc[k] = __dadd_rn(acc[k], __ddiv_rn(__dmul_rn(tt, p[k]), d3));
                    } else {
                        double vb[4];
#pragma unroll
                        for (int k = 0; k < NS; k++) vb[k] = v_y[j * NS + k];
                        const double rho_a = va[ND], rho_b = vb[ND];
                        const double rho_mean = __dmul_rn(0.5, __dadd_rn(rho_a, rho_b));
                        const double m_b = mass_y[j], p_b = p_y[j];
                        const bool far = !(d < sqrt_eps);
                        double grad[3] = {0.0, 0.0, 0.0};
                        if (far) {
                            const double q = __ddiv_rn(d, prm.h);
                            double w = 0.0;
                            if (q < 2.0) {
                                const double t1 = __dsub_rn(1.0, __dmul_rn(q, 0.5));
                                w = __dmul_rn(__dmul_rn(-5.0, q), __dmul_rn(__dmul_rn(t1, t1), t1));
                            }
                            const double dw = __dmul_rn(__ddiv_rn(prm.kernel_norm, prm.h), w);
                            const double sg = __ddiv_rn(dw, d);
#pragma unroll
                            for (int k = 0; k < ND; k++) grad[k] = __dmul_rn(sg, p[k]);
                        }
                        const double pf = __ddiv_rn(__dmul_rn(-m_b, __dadd_rn(p_a, p_b)), __dmul_rn(rho_a, rho_b));
                        double vdiff[3] = {0.0, 0.0, 0.0};
#pragma unroll
                        for (int k = 0; k < ND; k++) vdiff[k] = __dsub_rn(va[k], vb[k]);
                        double vr = __dmul_rn(vdiff[0], p[0]);
#pragma unroll
                        for (int k = 1; k < ND; k++) vr = __dadd_rn(vr, __dmul_rn(vdiff[k], p[k]));
                        double visc = 0.0;
                        if (vr < 0.0) {
                            const double mu = __ddiv_rn(__dmul_rn(prm.h, vr),
                                                        __dadd_rn(__dmul_rn(d, d), __dmul_rn(prm.epsilon, __dmul_rn(prm.h, prm.h))));
                            const double pi_ab = __ddiv_rn(
                                __dsub_rn(__dmul_rn(__dmul_rn(prm.alpha, prm.sound_speed), mu),
                                          __dmul_rn(prm.beta, __dmul_rn(mu, mu))), rho_mean);
                            visc = __dmul_rn(m_b, pi_ab);
                        }
#pragma unroll
                        for (int k = 0; k < ND; k++)
                            acc[k] = __dadd_rn(acc[k], __dadd_rn(__dmul_rn(pf, grad[k]), __dmul_rn(visc, grad[k])));
                        double vg = __dmul_rn(vdiff[0], grad[0]);
#pragma unroll
                        for (int k = 1; k < ND; k++) vg = __dadd_rn(vg, __dmul_rn(vdiff[k], grad[k]));
                        double drho = __dmul_rn(__dmul_rn(__ddiv_rn(rho_a, rho_b), m_b), vg);
                        if (far) {
                            const double vol_b = __ddiv_rn(m_b, rho_b);
                            const double two_drho = __dmul_rn(2.0, __dsub_rn(rho_a, rho_b));
                            const double dd = __dmul_rn(d, d);
                            double pg = 0.0;
#pragma unroll
                            for (int k = 0; k < ND; k++) {
                                const double psi = __ddiv_rn(__dmul_rn(two_drho, p[k]), dd);
                                pg = (k == 0) ? __dmul_rn(psi, grad[k]) : __dadd_rn(pg, __dmul_rn(psi, grad[k]));
                            }
                            drho = __dadd_rn(drho, __dmul_rn(__dmul_rn(__dmul_rn(prm.delta, prm.h), prm.sound_speed),
                                                             __dmul_rn(pg, vol_b)));
                        }
                        acc[ND] = __dadd_rn(acc[ND], drho);
                    }
                }
            }
    if (oob) atomicOr(err, 2);
    if (KIND == 0) {
#pragma unroll
        for (int k = 0; k < ND; k++) dv[i * ND + k] = acc[k];
    } else {
#pragma unroll
        for (int k = 0; k < NS; k++) dv[i * NS + k] = acc[k];
    }
}
#endif  // __CUDACC__

}  // namespace pnb
