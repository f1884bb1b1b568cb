// closures.cuh -- the benchmarked closures f(i, j, pos_diff, distance) as device functors that are
// fused into the sweep kernels.
//
// Functor protocol (see sweep.cuh):
//   State                 per-point accumulators, live in registers of the lane that owns point i
//   init(st, active, i_sorted, i_id)   i_sorted >= 0: cell-ordered index (x === y fast path)
//                                      i_sorted <  0: general path, i_id indexes the x arrays
//   stage(pay, slot, gi)  copy the payload of cell-ordered candidate gi into shared slot `slot`
//   pair<ND>(st, px, py, pz, d2, j_id, pay, slot, cap)   one accepted pair, payload in shared memory
//   pair_global<ND>(st, px, py, pz, d2, j_id, gi)        same, payload read from global memory
//   finish(st, i_sorted, i_id)         write the result of point i (original numbering)
//
// All per-pair arithmetic is written with *_rn intrinsics in the operation order of the oracle
// (oracle/pn_oracle_impl.h), which follows the Julia closures; because the sweep also visits the
// candidates in the oracle's order, sums agree with the oracle's Float32 sums bit for bit.
#pragma once

#include "common.cuh"

namespace pnb {

#define PNB_SQRT_EPS_F32 3.4526698300124393e-4f  // sqrt(eps(Float32))

// benchmarks/count_neighbors.jl:24-27   n_neighbors[i] += 1
struct CountCl {
    static constexpr bool kCountOnly = true;
    static constexpr int kPayBytes = 0;
    int64_t *out;
    struct State { int cnt; };
    __device__ __forceinline__ void init(State &s, bool, int, int) const { s.cnt = 0; }
    __device__ __forceinline__ void stage(unsigned char *, int, uint32_t) const {}
    __device__ __forceinline__ void count(State &s, int c) const { s.cnt += c; }
    template <int ND>
    __device__ __forceinline__ void pair(State &, float, float, float, float, int,
                                         const unsigned char *, int, int) const {}
    template <int ND>
    __device__ __forceinline__ void pair_global(State &, float, float, float, float, int,
                                                uint32_t) const {}
    __device__ __forceinline__ void finish(State &s, int, int i_id) const { out[i_id] = (int64_t)s.cnt; }
};

// benchmarks/n_body.jl:38-48
//   distance < sqrt(eps(ELTYPE)) && return
//   dv_ = -G * mass[j] * pos_diff / distance^3 ;  dv[dim, i] += dv_[dim]
struct NBodyCl {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 4;
    const float *mass_sorted;  // mass of the neighbour points in cell order
    float negG;
    float *dv;
    int nd;
    struct State { float a[3]; };
    __device__ __forceinline__ void init(State &s, bool, int, int) const { s.a[0] = s.a[1] = s.a[2] = 0.f; }
    __device__ __forceinline__ void stage(unsigned char *pay, int slot, uint32_t gi) const
    {
        reinterpret_cast<float *>(pay)[slot] = mass_sorted[gi];
    }
    __device__ __forceinline__ void count(State &, int) const {}
    template <int ND>
    __device__ __forceinline__ void term(State &s, float px, float py, float pz, float d2, float m) const
    {
        const float d = __fsqrt_rn(d2);
        if (d < PNB_SQRT_EPS_F32) return;
        const float t = __fmul_rn(negG, m);
        const float d3 = __fmul_rn(__fmul_rn(d, d), d);
        s.a[0] = __fadd_rn(s.a[0], __fdiv_rn(__fmul_rn(t, px), d3));
        if (ND > 1) s.a[1] = __fadd_rn(s.a[1], __fdiv_rn(__fmul_rn(t, py), d3));
        if (ND > 2) s.a[2] = __fadd_rn(s.a[2], __fdiv_rn(__fmul_rn(t, pz), d3));
    }
    template <int ND>
    __device__ __forceinline__ void pair(State &s, float px, float py, float pz, float d2, int,
                                         const unsigned char *pay, int slot, int) const
    {
        term<ND>(s, px, py, pz, d2, reinterpret_cast<const float *>(pay)[slot]);
    }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float px, float py, float pz, float d2,
                                                int, uint32_t gi) const
    {
        term<ND>(s, px, py, pz, d2, __ldg(mass_sorted + gi));
    }
    __device__ __forceinline__ void finish(State &s, int, int i_id) const
    {
        for (int k = 0; k < nd; k++) dv[(int64_t)i_id * nd + k] = s.a[k];
    }
};

// WCSPH continuity + momentum (TrixiParticles.interact!, benchmarks/smoothed_particle_hydrodynamics.jl:45-102).
// Formulas: oracle/pn_oracle_impl.h pno_cl_wcsph (parity with TrixiParticles itself is unpinned).
struct WcsphCl {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 16 + 8;
    const float4 *vrho_sorted;  // (vx, vy, vz, rho) of the neighbour points, cell order
    const float2 *mp_sorted;    // (mass, pressure) of the neighbour points, cell order
    const float *v_x;           // general path: state of the points looped over, (nd+1) per point
    const float *p_x;
    pnb_wcsph_params prm;
    float *dv;
    int nd;
    struct State { float v[3]; float rho, p; float acc[4]; };

    __device__ __forceinline__ void init(State &s, bool active, int i_sorted, int i_id) const
    {
        s.acc[0] = s.acc[1] = s.acc[2] = s.acc[3] = 0.f;
        s.v[0] = s.v[1] = s.v[2] = 0.f; s.rho = 1.f; s.p = 0.f;
        if (!active) return;
        if (i_sorted >= 0) {
            const float4 a = vrho_sorted[i_sorted];
            s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.rho = a.w;
            s.p = mp_sorted[i_sorted].y;
        } else {
            const int ns = nd + 1;
            for (int k = 0; k < nd; k++) s.v[k] = v_x[(int64_t)i_id * ns + k];
            s.rho = v_x[(int64_t)i_id * ns + nd];
            s.p = p_x[i_id];
        }
    }
    __device__ __forceinline__ void stage(unsigned char *pay, int slot, uint32_t gi) const
    {
        reinterpret_cast<float4 *>(pay)[slot] = vrho_sorted[gi];
        reinterpret_cast<float2 *>(pay + 16 * 512)[slot] = mp_sorted[gi];  // plane 1 after kCap float4
    }
    __device__ __forceinline__ void count(State &, int) const {}

    template <int ND>
    __device__ __forceinline__ void term(State &s, float px, float py, float pz, float d2,
                                         float4 vb, float2 mpb) const
    {
        const float d = __fsqrt_rn(d2);
        const float rho_a = s.rho, rho_b = vb.w;
        const float m_b = mpb.x, p_b = mpb.y;
        const float h = prm.smoothing_length;
        const float pd[3] = {px, py, pz};
        const float vbv[3] = {vb.x, vb.y, vb.z};
        const bool far = !(d < PNB_SQRT_EPS_F32);
        float grad[3] = {0.f, 0.f, 0.f};
        if (far) {
            const float q = __fdiv_rn(d, h);
            float w = 0.f;
            if (q < 2.f) {
                const float t = __fsub_rn(1.f, __fmul_rn(q, 0.5f));
                w = __fmul_rn(__fmul_rn(-5.f, q), __fmul_rn(__fmul_rn(t, t), t));
            }
            const float dw = __fmul_rn(__fdiv_rn(prm.kernel_norm, h), w);
            const float sgrad = __fdiv_rn(dw, d);
#pragma unroll
            for (int k = 0; k < ND; k++) grad[k] = __fmul_rn(sgrad, pd[k]);
        }
        const float pf = __fdiv_rn(__fmul_rn(-m_b, __fadd_rn(s.p, p_b)), __fmul_rn(rho_a, rho_b));
        float vdiff[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < ND; k++) vdiff[k] = __fsub_rn(s.v[k], vbv[k]);
        float vr = __fmul_rn(vdiff[0], pd[0]);
#pragma unroll
        for (int k = 1; k < ND; k++) vr = __fadd_rn(vr, __fmul_rn(vdiff[k], pd[k]));
        float visc = 0.f;
        if (vr < 0.f) {
            const float rho_mean = __fmul_rn(0.5f, __fadd_rn(rho_a, rho_b));
            const float mu = __fdiv_rn(__fmul_rn(h, vr),
                                       __fadd_rn(__fmul_rn(d, d), __fmul_rn(prm.epsilon, __fmul_rn(h, h))));
            const float pi_ab = __fdiv_rn(
                __fsub_rn(__fmul_rn(__fmul_rn(prm.alpha, prm.sound_speed), mu),
                          __fmul_rn(prm.beta, __fmul_rn(mu, mu))),
                rho_mean);
            visc = __fmul_rn(m_b, pi_ab);
        }
#pragma unroll
        for (int k = 0; k < ND; k++)
            s.acc[k] = __fadd_rn(s.acc[k], __fadd_rn(__fmul_rn(pf, grad[k]), __fmul_rn(visc, grad[k])));
        float vg = __fmul_rn(vdiff[0], grad[0]);
#pragma unroll
        for (int k = 1; k < ND; k++) vg = __fadd_rn(vg, __fmul_rn(vdiff[k], grad[k]));
        float drho = __fmul_rn(__fmul_rn(__fdiv_rn(rho_a, rho_b), m_b), vg);
        if (far) {
            const float vol_b = __fdiv_rn(m_b, rho_b);
            const float two_drho = __fmul_rn(2.f, __fsub_rn(rho_a, rho_b));
            const float dd = __fmul_rn(d, d);
            float pg = 0.f;
#pragma unroll
            for (int k = 0; k < ND; k++) {
                const float psi = __fdiv_rn(__fmul_rn(two_drho, pd[k]), dd);
                pg = (k == 0) ? __fmul_rn(psi, grad[k]) : __fadd_rn(pg, __fmul_rn(psi, grad[k]));
            }
            drho = __fadd_rn(drho, __fmul_rn(__fmul_rn(__fmul_rn(prm.delta, h), prm.sound_speed),
                                             __fmul_rn(pg, vol_b)));
        }
        s.acc[3] = __fadd_rn(s.acc[3], drho);
    }
    template <int ND>
    __device__ __forceinline__ void pair(State &s, float px, float py, float pz, float d2, int,
                                         const unsigned char *pay, int slot, int) const
    {
        const float4 vb = reinterpret_cast<const float4 *>(pay)[slot];
        const float2 mpb = reinterpret_cast<const float2 *>(pay + 16 * 512)[slot];
        term<ND>(s, px, py, pz, d2, vb, mpb);
    }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float px, float py, float pz, float d2,
                                                int, uint32_t gi) const
    {
        term<ND>(s, px, py, pz, d2, __ldg(vrho_sorted + gi), __ldg(mp_sorted + gi));
    }
    __device__ __forceinline__ void finish(State &s, int, int i_id) const
    {
        const int ns = nd + 1;
        for (int k = 0; k < nd; k++) dv[(int64_t)i_id * ns + k] = s.acc[k];
        dv[(int64_t)i_id * ns + nd] = s.acc[3];
    }
};

}  // namespace pnb
