// closures.cuh -- the benchmarked closures f(i, j, pos_diff, distance) as device functors that are
// fused into the sweep kernels.
//
// Functor protocol (see sweep.cuh):
//   State                 per-point accumulators, live in registers of the lane that owns point i
//   init(st, active, i_sorted, i_id)   i_sorted >= 0: cell-ordered index (x === y fast path)
//                                      i_sorted <  0: general path, i_id indexes the x arrays
//   stage(pay, slot, gi, cap)  copy the payload of cell-ordered candidate gi into shared slot
//                         `slot`; payload planes are arrays of `cap` elements
//   merge(st, other)      add the partial accumulators of another warp (sweep_tiles.cuh)
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
    static constexpr int kWarpsPerCell = 2;
    static constexpr int kAccWords = 1;
    int64_t *out;
    struct State { int cnt; };
    __device__ __forceinline__ void save_acc(const State &, float *) const {}
    __device__ __forceinline__ void add_acc(State &, const float *) const {}
    __device__ __forceinline__ void seek(State &, int) const {}
    __device__ __forceinline__ void total(State &, int, bool) const {}
    __device__ __forceinline__ void flush(State &) const {}
    __device__ __forceinline__ void init(State &s, bool, int, int) const { s.cnt = 0; }
    __device__ __forceinline__ void stage(unsigned char *, int, uint32_t, int) const {}
    __device__ __forceinline__ void count(State &s, int c) const { s.cnt += c; }
    __device__ __forceinline__ void merge(State &s, const State &o) const { s.cnt += o.cnt; }
    template <int ND>
    __device__ __forceinline__ void pair(State &, float, float, float, float, int,
                                         const unsigned char *, int, int) const {}
    template <int ND>
    __device__ __forceinline__ void pair_s(State &, float, float, float, float, int, uint32_t, int,
                                           int) const {}
    template <int ND>
    __device__ __forceinline__ void pair_global(State &, float, float, float, float, int,
                                                uint32_t) const {}
    __device__ __forceinline__ void finish(State &s, int, int i_id) const { out[i_id] = (int64_t)s.cnt; }
};

// MUFU-based approximations used by the fast (default) interaction arithmetic.  They only ever
// touch the per-pair TERMS; which pairs are delivered is decided by exact arithmetic.
__device__ __forceinline__ float fast_rsqrt(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_rcp(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// benchmarks/n_body.jl:38-48
//   distance < sqrt(eps(ELTYPE)) && return
//   dv_ = -G * mass[j] * pos_diff / distance^3 ;  dv[dim, i] += dv_[dim]
// EXACT: the Julia operation sequence with IEEE sqrt/div (bit-identical terms and sums).
// fast (default): one MUFU.RSQ, FMAs; per-term relative error ~3e-7, far inside the 1e-5 bar.
template <bool EXACT>
struct NBodyClT {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 4;
    static constexpr int kWarpsPerCell = 2;
    const float *mass_sorted;  // mass of the neighbour points in cell order
    float negG;
    float *dv;
    int nd;
    struct State { float a[3]; };
    static constexpr int kAccWords = 3;
    // partial accumulators of one warp, word w of lane l at p[w * 32] (p already points at lane l)
    __device__ __forceinline__ void save_acc(const State &s, float *p) const
    {
        p[0] = s.a[0]; p[32] = s.a[1]; p[64] = s.a[2];
    }
    __device__ __forceinline__ void add_acc(State &s, const float *p) const
    {
        s.a[0] += p[0]; s.a[1] += p[32]; s.a[2] += p[64];
    }
    __device__ __forceinline__ void seek(State &, int) const {}
    __device__ __forceinline__ void total(State &, int, bool) const {}
    __device__ __forceinline__ void flush(State &) const {}
    __device__ __forceinline__ void init(State &s, bool, int, int) const { s.a[0] = s.a[1] = s.a[2] = 0.f; }
    __device__ __forceinline__ void stage(unsigned char *pay, int slot, uint32_t gi, int) const
    {
        reinterpret_cast<float *>(pay)[slot] = mass_sorted[gi];
    }
    __device__ __forceinline__ void count(State &, int) const {}
    __device__ __forceinline__ void merge(State &s, const State &o) const
    {
        s.a[0] += o.a[0]; s.a[1] += o.a[1]; s.a[2] += o.a[2];
    }
    template <int ND>
    __device__ __forceinline__ void term(State &s, float px, float py, float pz, float d2, float m) const
    {
        if (EXACT) {
            const float d = __fsqrt_rn(d2);
            if (d < PNB_SQRT_EPS_F32) return;
            const float t = __fmul_rn(negG, m);
            const float d3 = __fmul_rn(__fmul_rn(d, d), d);
            s.a[0] = __fadd_rn(s.a[0], __fdiv_rn(__fmul_rn(t, px), d3));
            if (ND > 1) s.a[1] = __fadd_rn(s.a[1], __fdiv_rn(__fmul_rn(t, py), d3));
            if (ND > 2) s.a[2] = __fadd_rn(s.a[2], __fdiv_rn(__fmul_rn(t, pz), d3));
        } else {
            if (d2 < PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32) return;
            const float inv_d = fast_rsqrt(d2);
            const float c = (negG * m) * (inv_d * inv_d * inv_d);
            s.a[0] = fmaf(c, px, s.a[0]);
            if (ND > 1) s.a[1] = fmaf(c, py, s.a[1]);
            if (ND > 2) s.a[2] = fmaf(c, pz, s.a[2]);
        }
    }
    template <int ND>
    __device__ __forceinline__ void pair(State &s, float px, float py, float pz, float d2, int,
                                         const unsigned char *pay, int slot, int) const
    {
        term<ND>(s, px, py, pz, d2, reinterpret_cast<const float *>(pay)[slot]);
    }
    // payload addressed by its 32-bit shared-window address (k_sweep_tiles)
    template <int ND>
    __device__ __forceinline__ void pair_s(State &s, float px, float py, float pz, float d2, int,
                                           uint32_t pay_sa, int slot, int) const
    {
        float m;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m) : "r"(pay_sa + 4u * (uint32_t)slot));
        term<ND>(s, px, py, pz, d2, m);
    }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float px, float py, float pz, float d2,
                                                int, uint32_t gi) const
    {
        term<ND>(s, px, py, pz, d2, __ldg(mass_sorted + gi));
    }
    static constexpr int kPayBytesTile = 4;
    __device__ __forceinline__ void stage_tile(unsigned char *pay, int slot, uint32_t gi, int cap) const
    {
        stage(pay, slot, gi, cap);
    }
    __device__ __forceinline__ void stage_async(uint32_t pay_sa, int slot, uint32_t gi, int) const
    {
        cp_async4(pay_sa + 4u * (uint32_t)slot, mass_sorted + gi);
    }
    template <int ND>
    __device__ __forceinline__ void pair_tile(State &s, float px, float py, float pz, float d2, int j,
                                              uint32_t pay_sa, int slot, int cap) const
    {
        pair_s<ND>(s, px, py, pz, d2, j, pay_sa, slot, cap);
    }
    template <int ND>
    __device__ __forceinline__ void pair_tile_pred(State &s, float px, float py, float pz, float d2,
                                                   int j, uint32_t pay_sa, int slot, int cap, bool ok) const
    {
        if (EXACT) { if (ok) pair_s<ND>(s, px, py, pz, d2, j, pay_sa, slot, cap); return; }
        float m;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(m) : "r"(pay_sa + 4u * (uint32_t)slot));
        const bool use = ok && d2 >= PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32;
        const float inv_d = fast_rsqrt(use ? d2 : 1.0f);
        float c = (negG * m) * (inv_d * inv_d * inv_d);
        c = use ? c : 0.f;
        s.a[0] = fmaf(c, px, s.a[0]);
        if (ND > 1) s.a[1] = fmaf(c, py, s.a[1]);
        if (ND > 2) s.a[2] = fmaf(c, pz, s.a[2]);
    }
    __device__ __forceinline__ void finish(State &s, int, int i_id) const
    {
        for (int k = 0; k < nd; k++) dv[(int64_t)i_id * nd + k] = s.a[k];
    }
};

// WCSPH continuity + momentum (TrixiParticles.interact!, benchmarks/smoothed_particle_hydrodynamics.jl:45-102).
// Formulas: oracle/pn_oracle_impl.h pno_cl_wcsph (parity with TrixiParticles itself is unpinned).
// EXACT: the oracle's operation sequence (IEEE div/sqrt, no FMA) -> sums identical to the oracle.
// fast (default): algebraically the same terms with one MUFU.RSQ per pair (two MUFU.RCP more for
// approaching pairs), FMAs, and 1/rho_b, m_b/rho_b precomputed once per neighbour:
//   grad_k = s p_k,  s = (sigma/h) w(q) / d
//   dv_k  += (pf + visc) s p_k
//   drho  += (m_b/rho_b) s [ rho_a (v_ab . p) + 2 delta h c (rho_a - rho_b) ]     (psi.grad = 2 (rho_a - rho_b) s)
template <bool EXACT>
struct WcsphClT {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 16 + 16;
    static constexpr int kWarpsPerCell = 4;
    const float4 *vrho_sorted;  // (vx, vy, vz, rho) of the neighbour points, cell order
    const float4 *mp_sorted;    // (mass, pressure, 1/rho, mass/rho) of the neighbour points
    const float *v_x;           // general path: state of the points looped over, (nd+1) per point
    const float *p_x;
    pnb_wcsph_params prm;
    float *dv;
    int nd;
    struct State { float v[3]; float rho, p, inv_rho, neg_inv_rho; float acc[4]; };
    static constexpr int kAccWords = 4;
    __device__ __forceinline__ void save_acc(const State &s, float *p) const
    {
        p[0] = s.acc[0]; p[32] = s.acc[1]; p[64] = s.acc[2]; p[96] = s.acc[3];
    }
    __device__ __forceinline__ void add_acc(State &s, const float *p) const
    {
        s.acc[0] += p[0]; s.acc[1] += p[32]; s.acc[2] += p[64]; s.acc[3] += p[96];
    }
    __device__ __forceinline__ void seek(State &, int) const {}
    __device__ __forceinline__ void total(State &, int, bool) const {}
    __device__ __forceinline__ void flush(State &) const {}

    __device__ __forceinline__ void init(State &s, bool active, int i_sorted, int i_id) const
    {
        s.acc[0] = s.acc[1] = s.acc[2] = s.acc[3] = 0.f;
        s.v[0] = s.v[1] = s.v[2] = 0.f; s.rho = 1.f; s.p = 0.f; s.inv_rho = 1.f; s.neg_inv_rho = -1.f;
        if (!active) return;
        if (i_sorted >= 0) {
            const float4 a = vrho_sorted[i_sorted];
            s.v[0] = a.x; s.v[1] = a.y; s.v[2] = a.z; s.rho = a.w;
            if (EXACT) {
                const float4 b = mp_sorted[i_sorted];
                s.p = b.y; s.inv_rho = b.z;
            } else {
                s.p = vp_sorted[i_sorted].y;
                s.inv_rho = __fdiv_rn(1.f, a.w);
            }
        } else {
            const int ns = nd + 1;
            for (int k = 0; k < nd; k++) s.v[k] = v_x[(int64_t)i_id * ns + k];
            s.rho = v_x[(int64_t)i_id * ns + nd];
            s.p = p_x[i_id];
            s.inv_rho = __fdiv_rn(1.f, s.rho);
        }
        s.neg_inv_rho = -s.inv_rho;
    }
    __device__ __forceinline__ void stage(unsigned char *pay, int slot, uint32_t gi, int cap) const
    {
        reinterpret_cast<float4 *>(pay)[slot] = vrho_sorted[gi];
        if (EXACT) {
            reinterpret_cast<float4 *>(pay + 16 * cap)[slot] = mp_sorted[gi];   // plane 1
        } else {
            const float2 vp = vp_sorted[gi];      // the fast term reads (.w, .y) = (m/rho, p)
            reinterpret_cast<float4 *>(pay + 16 * cap)[slot] = make_float4(0.f, vp.y, 0.f, vp.x);
        }
    }
    __device__ __forceinline__ void count(State &, int) const {}
    __device__ __forceinline__ void merge(State &s, const State &o) const
    {
        s.acc[0] += o.acc[0]; s.acc[1] += o.acc[1]; s.acc[2] += o.acc[2]; s.acc[3] += o.acc[3];
    }

    // vol_b = m_b / rho_b and p_b are all the fast term needs of (mass, pressure, 1/rho, m/rho);
    // m_b itself only appears in the viscosity of approaching pairs: m_b = vol_b * rho_b
    template <int ND>
    __device__ __forceinline__ void term_fast(State &s, float px, float py, float pz, float d2,
                                              float4 vb, float vol_b, float p_b) const
    {
        // pairs closer than sqrt(eps) (the self pair) contribute exactly zero in the reference
        if (d2 < PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32) return;
        const float h = prm.smoothing_length;
        // grad W / d = (sigma/h) w(q)/d with w = -5 q (1 - q/2)^3 and q = d/h: the d cancels,
        //   sg = (-5 sigma / h^2) t^3,  t = 1 - d / (2 h)
        const float inv_d = fast_rsqrt(d2);
        const float d = d2 * inv_d;
        const float t = fmaxf(fmaf(prm_nhalf_inv_h, d, 1.f), 0.f);   // w = 0 beyond the support q >= 2
        const float sg = (prm_k5 * t) * (t * t);
        const float rho_a = s.rho, rho_b = vb.w;
        float coef = (vol_b * (s.p + p_b)) * s.neg_inv_rho;       // -m_b (p_a + p_b) / (rho_a rho_b)
        const float vdx = s.v[0] - vb.x, vdy = s.v[1] - vb.y, vdz = s.v[2] - vb.z;
        float vr = vdx * px;
        if (ND > 1) vr = fmaf(vdy, py, vr);
        if (ND > 2) vr = fmaf(vdz, pz, vr);
        if (vr < 0.f) {
            const float mu = (h * vr) * fast_rcp(fmaf(prm.epsilon, h * h, d2));
            const float pi_ab = (prm_ac * mu - prm.beta * (mu * mu)) * fast_rcp(0.5f * (rho_a + rho_b));
            coef = fmaf(vol_b * rho_b, pi_ab, coef);
        }
        coef *= sg;
        s.acc[0] = fmaf(coef, px, s.acc[0]);
        if (ND > 1) s.acc[1] = fmaf(coef, py, s.acc[1]);
        if (ND > 2) s.acc[2] = fmaf(coef, pz, s.acc[2]);
        s.acc[3] = fmaf(vol_b * sg, fmaf(rho_a, vr, prm_dhc2 * (rho_a - rho_b)), s.acc[3]);
    }

    template <int ND>
    __device__ __forceinline__ void term(State &s, float px, float py, float pz, float d2,
                                         float4 vb, float4 mpb4) const
    {
        if (!EXACT) { term_fast<ND>(s, px, py, pz, d2, vb, mpb4.w, mpb4.y); return; }
        const float2 mpb = make_float2(mpb4.x, mpb4.y);
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
                                         const unsigned char *pay, int slot, int cap) const
    {
        const float4 vb = reinterpret_cast<const float4 *>(pay)[slot];
        const float4 mpb = reinterpret_cast<const float4 *>(pay + 16 * cap)[slot];
        term<ND>(s, px, py, pz, d2, vb, mpb);
    }
    template <int ND>
    __device__ __forceinline__ void pair_s(State &s, float px, float py, float pz, float d2, int,
                                           uint32_t pay_sa, int slot, int cap) const
    {
        float4 vb, mpb;
        const uint32_t a = pay_sa + 16u * (uint32_t)slot;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(vb.x), "=f"(vb.y), "=f"(vb.z), "=f"(vb.w) : "r"(a));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(mpb.x), "=f"(mpb.y), "=f"(mpb.z), "=f"(mpb.w) : "r"(a + 16u * (uint32_t)cap));
        term<ND>(s, px, py, pz, d2, vb, mpb);
    }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float px, float py, float pz, float d2,
                                                int, uint32_t gi) const
    {
        if (EXACT) {
            term<ND>(s, px, py, pz, d2, __ldg(vrho_sorted + gi), __ldg(mp_sorted + gi));
        } else {
            const float2 vp = __ldg(vp_sorted + gi);
            term_fast<ND>(s, px, py, pz, d2, __ldg(vrho_sorted + gi), vp.x, vp.y);
        }
    }
    // k_sweep_flat stages 24 instead of 32 bytes per candidate: (vx, vy, vz, rho) and (m/rho, p)
    static constexpr int kPayBytesTile = EXACT ? 32 : 24;
    __device__ __forceinline__ void stage_tile(unsigned char *pay, int slot, uint32_t gi, int cap) const
    {
        reinterpret_cast<float4 *>(pay)[slot] = vrho_sorted[gi];
        if (EXACT) reinterpret_cast<float4 *>(pay + 16 * cap)[slot] = mp_sorted[gi];
        else reinterpret_cast<float2 *>(pay + 16 * cap)[slot] = vp_sorted[gi];
    }
    // the same as asynchronous copies global -> shared (LDGSTS), no registers in between
    __device__ __forceinline__ void stage_async(uint32_t pay_sa, int slot, uint32_t gi, int cap) const
    {
        cp_async16(pay_sa + 16u * (uint32_t)slot, vrho_sorted + gi);
        stage_async_rest(pay_sa, slot, gi, cap);
    }
    // k_sweep_flat copies plane 0 (16 bytes per candidate) cell by cell with the TMA unit
    // (bulk_plane) and only the rest per candidate
    __device__ __forceinline__ const float4 *bulk_plane() const { return vrho_sorted; }
    __device__ __forceinline__ void stage_async_rest(uint32_t pay_sa, int slot, uint32_t gi, int cap) const
    {
        if (EXACT) cp_async16(pay_sa + 16u * (uint32_t)cap + 16u * (uint32_t)slot, mp_sorted + gi);
        else cp_async8(pay_sa + 16u * (uint32_t)cap + 8u * (uint32_t)slot, vp_sorted + gi);
    }
    // branch-free form for the drain of k_sweep_flat: the payload is loaded unconditionally (the
    // loads do not wait for the radius test), `ok` only decides whether the term counts
    template <int ND>
    __device__ __forceinline__ void pair_tile_pred(State &s, float px, float py, float pz, float d2,
                                                   int, uint32_t pay_sa, int slot, int cap, bool ok) const
    {
        if (EXACT) { if (ok) pair_s<ND>(s, px, py, pz, d2, 0, pay_sa, slot, cap); return; }
        float4 vb;
        float vol_b, p_b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(vb.x), "=f"(vb.y), "=f"(vb.z), "=f"(vb.w) : "r"(pay_sa + 16u * (uint32_t)slot));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                     : "=f"(vol_b), "=f"(p_b) : "r"(pay_sa + 16u * (uint32_t)cap + 8u * (uint32_t)slot));
        // pairs closer than sqrt(eps) (the self pair) and lanes without a hit in this round
        // contribute exactly zero: sg = 0
        const bool use = ok && d2 >= PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32;
        const float d2s = use ? d2 : 1.0f;
        const float h = prm.smoothing_length;
        const float inv_d = fast_rsqrt(d2s);
        const float d = d2s * inv_d;
        const float t = fmaxf(fmaf(prm_nhalf_inv_h, d, 1.f), 0.f);
        float sg = (prm_k5 * t) * (t * t);
        sg = use ? sg : 0.f;
        const float rho_a = s.rho, rho_b = vb.w;
        float coef = (vol_b * (s.p + p_b)) * s.neg_inv_rho;
        const float vdx = s.v[0] - vb.x, vdy = s.v[1] - vb.y, vdz = s.v[2] - vb.z;
        float vr = vdx * px;
        if (ND > 1) vr = fmaf(vdy, py, vr);
        if (ND > 2) vr = fmaf(vdz, pz, vr);
        if (vr < 0.f) {
            const float mu = (h * vr) * fast_rcp(fmaf(prm.epsilon, h * h, d2s));
            const float pi_ab = (prm_ac * mu - prm.beta * (mu * mu)) * fast_rcp(0.5f * (rho_a + rho_b));
            coef = fmaf(vol_b * rho_b, pi_ab, coef);
        }
        coef *= sg;
        s.acc[0] = fmaf(coef, px, s.acc[0]);
        if (ND > 1) s.acc[1] = fmaf(coef, py, s.acc[1]);
        if (ND > 2) s.acc[2] = fmaf(coef, pz, s.acc[2]);
        s.acc[3] = fmaf(vol_b * sg, fmaf(rho_a, vr, prm_dhc2 * (rho_a - rho_b)), s.acc[3]);
    }
    template <int ND>
    __device__ __forceinline__ void pair_tile(State &s, float px, float py, float pz, float d2, int,
                                              uint32_t pay_sa, int slot, int cap) const
    {
        if (EXACT) { pair_s<ND>(s, px, py, pz, d2, 0, pay_sa, slot, cap); return; }
        float4 vb;
        float vol_b, p_b;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(vb.x), "=f"(vb.y), "=f"(vb.z), "=f"(vb.w) : "r"(pay_sa + 16u * (uint32_t)slot));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];"
                     : "=f"(vol_b), "=f"(p_b) : "r"(pay_sa + 16u * (uint32_t)cap + 8u * (uint32_t)slot));
        term_fast<ND>(s, px, py, pz, d2, vb, vol_b, p_b);
    }
    // the fast term vanishes identically for d >= 2 h (t = 0): when the kernel support 2 h does
    // not exceed the search radius, candidates of the fp16 pre-filter band (r < d <= 1.005 r)
    // contribute exactly zero and the exact radius test of the drain is redundant
    __host__ __device__ __forceinline__ bool no_radius_test() const { return !EXACT && support_in_radius; }
    __device__ __forceinline__ void finish(State &s, int, int i_id) const
    {
        const int ns = nd + 1;
        for (int k = 0; k < nd; k++) dv[(int64_t)i_id * ns + k] = s.acc[k];
        dv[(int64_t)i_id * ns + nd] = s.acc[3];
    }
    // derived constants of the fast path, filled by the host
    float prm_nhalf_inv_h;   // -1 / (2 h)
    float prm_k5;            // -5 kernel_norm / h^2
    float prm_ac;      // alpha * c
    float prm_dhc2;    // 2 * delta * h * c
    bool support_in_radius;   // 2 h <= search_radius (see no_radius_test)
    const float2 *vp_sorted;  // fast mode: (mass / rho, pressure) of the neighbour points, cell order
                              // (mp_sorted is filled in exact mode only)
};

}  // namespace pnb
