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
template <class T>
struct WcsphTP {
    T h, sound_speed, alpha, beta, epsilon, delta, kernel_norm;
};
using Wcsph64P = WcsphTP<double>;
// explicit IEEE operations of the element type the CLOSURE computes in: double for a Float64
// search; float for a mixed-precision search (Float64 coordinates, Float32 radius: pos_diff and
// distance arrive as Float32, so the closures of the benchmarks run in Float32 on Float32 state)
struct OpsF64 {
    using R = double;
    static __device__ __forceinline__ R mul(R a, R b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ R add(R a, R b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ R sub(R a, R b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ R div(R a, R b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ R sqrt(R a) { return __dsqrt_rn(a); }
    static __device__ __forceinline__ R sqrt_eps() { return 1.4901161193847656e-8; }
};
struct OpsF32 {
    using R = float;
    static __device__ __forceinline__ R mul(R a, R b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ R add(R a, R b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ R sub(R a, R b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ R div(R a, R b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ R sqrt(R a) { return __fsqrt_rn(a); }
    static __device__ __forceinline__ R sqrt_eps() { return 3.4526698300124393e-4f; }
};
//   KIND 0: n-body   (benchmarks/n_body.jl:38-48)            state: mass[n], out dv[nx x nd]
//   KIND 1: WCSPH    (smoothed_particle_hydrodynamics.jl:45-102)  state: v (nd+1), mass, pressure
template <int ND, int KIND, class O>
__global__ void __launch_bounds__(128)
k_sweep_closure_t(GridP64 g, const uint32_t *__restrict__ cell_start,
                  const Rec64 *__restrict__ sorted, const double *__restrict__ x, int64_t n_loop,
                  const int32_t *__restrict__ points, int base, const typename O::R *__restrict__ v_x,
                  const typename O::R *__restrict__ v_y, const typename O::R *__restrict__ mass_y,
                  const typename O::R *__restrict__ p_x, const typename O::R *__restrict__ p_y,
                  typename O::R G, WcsphTP<typename O::R> prm, typename O::R *__restrict__ dv,
                  int *__restrict__ err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_loop) return;
    const int64_t i = points ? (int64_t)points[t] - base : t;
    constexpr int NS = ND + 1;
    using R = typename O::R;
    const R sqrt_eps = O::sqrt_eps();
    double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) xi[d] = x[i * ND + d];
    int cc[3] = {1, 1, 1};
#pragma unroll
    for (int d = 0; d < ND; d++) cc[d] = cell_coord64(xi[d], g.minc[d], g.cs[d], g.periodic, g.nc[d]);
    R acc[4] = {R(0.0), R(0.0), R(0.0), R(0.0)};
    R va[4] = {R(0.0), R(0.0), R(0.0), R(0.0)}, p_a = R(0.0);
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
                    double pd[3];
                    const double d2d = pair_d2_64<ND>(g, xi, yj, pd, true);
                    if (!(d2d <= g.r2)) continue;
                    // (a mixed-precision search hands over Float32 values, widened: the casts are exact)
                    const R p[3] = {(R)pd[0], (R)pd[1], (R)pd[2]};
                    const R d = O::sqrt((R)d2d);
                    const int64_t j = yj.id;
                    if (KIND == 0) {
                        if (d < sqrt_eps) continue;
                        const R tt = O::mul(-G, mass_y[j]);
                        const R d3 = O::mul(O::mul(d, d), d);
#pragma unroll
                        for (int k = 0; k < ND; k++)
                            acc[k] = O::add(acc[k], O::div(O::mul(tt, p[k]), d3));
                    } else {
                        R vb[4];
#pragma unroll
                        for (int k = 0; k < NS; k++) vb[k] = v_y[j * NS + k];
                        const R rho_a = va[ND], rho_b = vb[ND];
                        const R rho_mean = O::mul(R(0.5), O::add(rho_a, rho_b));
                        const R m_b = mass_y[j], p_b = p_y[j];
                        const bool far = !(d < sqrt_eps);
                        R grad[3] = {R(0.0), R(0.0), R(0.0)};
                        if (far) {
                            const R q = O::div(d, prm.h);
                            R w = R(0.0);
                            if (q < R(2.0)) {
                                const R t1 = O::sub(R(1.0), O::mul(q, R(0.5)));
                                w = O::mul(O::mul(R(-5.0), q), O::mul(O::mul(t1, t1), t1));
                            }
                            const R dw = O::mul(O::div(prm.kernel_norm, prm.h), w);
                            const R sg = O::div(dw, d);
#pragma unroll
                            for (int k = 0; k < ND; k++) grad[k] = O::mul(sg, p[k]);
                        }
                        const R pf = O::div(O::mul(-m_b, O::add(p_a, p_b)), O::mul(rho_a, rho_b));
                        R vdiff[3] = {R(0.0), R(0.0), R(0.0)};
#pragma unroll
                        for (int k = 0; k < ND; k++) vdiff[k] = O::sub(va[k], vb[k]);
                        R vr = O::mul(vdiff[0], p[0]);
#pragma unroll
                        for (int k = 1; k < ND; k++) vr = O::add(vr, O::mul(vdiff[k], p[k]));
                        R visc = R(0.0);
                        if (vr < R(0.0)) {
                            const R mu = O::div(O::mul(prm.h, vr),
                                                        O::add(O::mul(d, d), O::mul(prm.epsilon, O::mul(prm.h, prm.h))));
                            const R pi_ab = O::div(
                                O::sub(O::mul(O::mul(prm.alpha, prm.sound_speed), mu),
                                          O::mul(prm.beta, O::mul(mu, mu))), rho_mean);
                            visc = O::mul(m_b, pi_ab);
                        }
#pragma unroll
                        for (int k = 0; k < ND; k++)
                            acc[k] = O::add(acc[k], O::add(O::mul(pf, grad[k]), O::mul(visc, grad[k])));
                        R vg = O::mul(vdiff[0], grad[0]);
#pragma unroll
                        for (int k = 1; k < ND; k++) vg = O::add(vg, O::mul(vdiff[k], grad[k]));
                        R drho = O::mul(O::mul(O::div(rho_a, rho_b), m_b), vg);
                        if (far) {
                            const R vol_b = O::div(m_b, rho_b);
                            const R two_drho = O::mul(R(2.0), O::sub(rho_a, rho_b));
                            const R dd = O::mul(d, d);
                            R pg = R(0.0);
#pragma unroll
                            for (int k = 0; k < ND; k++) {
                                const R psi = O::div(O::mul(two_drho, p[k]), dd);
                                pg = (k == 0) ? O::mul(psi, grad[k]) : O::add(pg, O::mul(psi, grad[k]));
                            }
                            drho = O::add(drho, O::mul(O::mul(O::mul(prm.delta, prm.h), prm.sound_speed),
                                                             O::mul(pg, vol_b)));
                        }
                        acc[ND] = O::add(acc[ND], drho);
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
