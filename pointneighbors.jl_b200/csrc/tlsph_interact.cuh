// tlsph_interact.cuh -- the steps either side of the neighbour sweep in a TLSPH / WCSPH right-hand
// side: TrixiParticles.interact_structure_structure! over precomputed lists,
// compute_pk1_corrected! and compute_pressure! (pointwise).
//
// reference call sites: benchmarks/smoothed_particle_hydrodynamics.jl:99 (compute_pressure!),
// :121 (interact_structure_structure!), :186 (compute_pk1_corrected!); list sweep without radius
// test: src/nhs_precomputed.jl:210-247.  The arithmetic itself belongs to TrixiParticles.jl (not
// vendored): PARITY UNPINNED, the definition is the oracle's (pno_tlsph_interact,
// pno_tlsph_pk1_corrected, pno_wcsph_compute_pressure).
#pragma once

#include "closures.cuh"
#include "grid.cuh"

namespace pnb {

#ifdef __CUDACC__

// Per-particle record of the force sweep: one aligned 128-byte line per neighbour instead of 26
// scattered 4-byte gathers from six arrays.  floats: [0..2] X0, [3] m, [4..6] x, [7] m / rho0,
// [8..16] PK1c / rho0^2 (column-major), [17..25] F, [26..31] unused.
constexpr int kTlRecF4 = 8;

template <int ND>
__global__ void k_pack_tlsph_full(int64_t n, const float *__restrict__ X0,
                                  const float *__restrict__ xcur, const float *__restrict__ mass,
                                  const float *__restrict__ rho0, const float *__restrict__ pk1c,
                                  const float *__restrict__ F, float4 *__restrict__ rec)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    constexpr int NN = ND * ND;
    float b[32];
#pragma unroll
    for (int e = 0; e < 32; e++) b[e] = 0.f;
#pragma unroll
    for (int d = 0; d < ND; d++) { b[d] = __ldg(X0 + j * ND + d); b[4 + d] = __ldg(xcur + j * ND + d); }
    const float m = __ldg(mass + j), rho = __ldg(rho0 + j);
    b[3] = m;
    b[7] = __fdiv_rn(m, rho);
    const float rho2 = __fmul_rn(rho, rho);
#pragma unroll
    for (int e = 0; e < NN; e++) {
        b[8 + e] = __fdiv_rn(__ldg(pk1c + j * NN + e), rho2);
        b[17 + e] = __ldg(F + j * NN + e);
    }
#pragma unroll
    for (int q = 0; q < kTlRecF4; q++)
        rec[kTlRecF4 * j + q] = make_float4(b[4 * q], b[4 * q + 1], b[4 * q + 2], b[4 * q + 3]);
}

__device__ __forceinline__ void ldg256f(const float4 *p, float *o)
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]), "=f"(o[5]),
                   "=f"(o[6]), "=f"(o[7])
                 : "l"(p));
}

// 8 lanes per point stride the point's list (coalesced id reads); every pair gathers the
// neighbour's 128-byte record with four LDG.E.256, the 3 partial accelerations are added by a
// shuffle tree.  Algorithmic HBM bytes: 4 P (ids) + 8 N (offsets) + 104 N (inputs) + 12 N (dv).
// EXACT: the oracle's IEEE operation sequence per pair; fast (default): MUFU.RSQ + FMAs.
template <int ND, bool PER, bool EXACT>
__global__ void __launch_bounds__(256, 2)
k_tlsph_interact(GridP g, int64_t n, const int64_t *__restrict__ offsets,
                 const int32_t *__restrict__ ids, const float4 *__restrict__ rec, float h,
                 float kernel_norm, float young, float alpha, float *__restrict__ dv)
{
    constexpr int kG = 8;
    constexpr int NN = ND * ND;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kG;
    const int sub = (int)(threadIdx.x % kG);
    const bool have = i < n;
    const int64_t ii = have ? i : 0;
    float own[32];
#pragma unroll
    for (int q = 0; q < 4; q++) ldg256f(rec + kTlRecF4 * ii + 2 * q, own + 8 * q);
    const float Xi[3] = {own[0], own[1], own[2]}, xi[3] = {own[4], own[5], own[6]};
    const float m_i = own[3], vol_i = own[7];
    float Bi[NN], Fi[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) { Bi[e] = own[8 + e]; Fi[e] = own[17 + e]; }
    float acc[3] = {0.f, 0.f, 0.f};
    const float nh = __fdiv_rn(kernel_norm, h);
    const float k5 = -5.0f * kernel_norm / (h * h), nhalf_inv_h = -0.5f / h, inv_h = 1.0f / h;
    // fast mode: everything of the penalty coefficient that belongs to point i
    const float ci = (alpha * 0.5f) * vol_i * young * kernel_norm / m_i;
    const PerP pp = make_perp(g);
    const int64_t k_beg = have ? offsets[ii] : 0;
    const int len = have ? (int)(offsets[ii + 1] - k_beg) : 0;
    const int32_t *my_ids = ids + k_beg;
    int max_len = len;
#pragma unroll
    for (int o = 16; o >= kG; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    int jn = sub < len ? __ldg(my_ids + sub) : -1;
    for (int k0 = sub; k0 < max_len; k0 += kG) {
        const int jj = jn;
        float nb[32];
        {
            const int64_t j = jj >= 0 ? jj : ii;
#pragma unroll
            for (int q = 0; q < 4; q++) ldg256f(rec + kTlRecF4 * j + 2 * q, nb + 8 * q);
        }
        jn = (k0 + kG < len) ? __ldg(my_ids + k0 + kG) : -1;
        if (jj < 0) continue;
        float p[3] = {0.f, 0.f, 0.f}, cp[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < ND; d++) { p[d] = __fsub_rn(Xi[d], nb[d]); cp[d] = __fsub_rn(xi[d], nb[4 + d]); }
        float d2 = dist2<ND>(p[0], p[1], p[2]);
        d2 = maybe_periodic_fix<ND, PER>(pp, d2, p[0], p[1], p[2]);
        const float m_j = nb[3], vol_j = nb[7];
        if (EXACT) {
            const float dist = __fsqrt_rn(d2);
            if (dist < PNB_SQRT_EPS_F32) continue;
            const float q = __fdiv_rn(dist, h);
            float dw = 0.f, w = 0.f;
            if (q < 2.f) {
                const float t = __fsub_rn(1.f, __fmul_rn(q, 0.5f));
                const float t2 = __fmul_rn(t, t);
                dw = __fmul_rn(__fmul_rn(-5.f, q), __fmul_rn(t2, t));
                w = __fmul_rn(__fmul_rn(t2, t2), __fadd_rn(__fmul_rn(2.f, q), 1.f));
            }
            const float sg = __fdiv_rn(__fmul_rn(nh, dw), dist);
            float grad[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < ND; d++) grad[d] = __fmul_rn(sg, p[d]);
            float term[3] = {0.f, 0.f, 0.f}, es[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int a = 0; a < ND; a++) {
                float t = 0.f, u = 0.f;
#pragma unroll
                for (int b = 0; b < ND; b++) {
                    const float A = __fadd_rn(Bi[b * ND + a], nb[8 + b * ND + a]);
                    const float Fs = __fadd_rn(Fi[b * ND + a], nb[17 + b * ND + a]);
                    t = b == 0 ? __fmul_rn(A, grad[0]) : __fadd_rn(t, __fmul_rn(A, grad[b]));
                    u = b == 0 ? __fmul_rn(Fs, p[0]) : __fadd_rn(u, __fmul_rn(Fs, p[b]));
                }
                term[a] = __fmul_rn(m_j, t);
                es[a] = __fsub_rn(u, __fmul_rn(2.f, cp[a]));
            }
            const float cd = __fsqrt_rn(dist2<ND>(cp[0], cp[1], cp[2]));
            float ds = __fmul_rn(es[0], cp[0]);
#pragma unroll
            for (int d = 1; d < ND; d++) ds = __fadd_rn(ds, __fmul_rn(es[d], cp[d]));
            ds = __fdiv_rn(ds, cd);
            float c = __fmul_rn(__fmul_rn(__fmul_rn(alpha, 0.5f), vol_i), vol_j);
            c = __fdiv_rn(__fmul_rn(c, __fmul_rn(kernel_norm, w)), __fmul_rn(dist, dist));
            c = __fdiv_rn(__fmul_rn(__fmul_rn(c, young), ds), cd);
#pragma unroll
            for (int d = 0; d < ND; d++) {
                acc[d] = __fadd_rn(acc[d], term[d]);
                acc[d] = __fadd_rn(acc[d], __fdiv_rn(__fmul_rn(c, cp[d]), m_i));
            }
        } else {
            if (d2 < PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32) continue;
            const float inv_d = fast_rsqrt(d2);
            const float dist = d2 * inv_d;
            const float t = fmaxf(fmaf(nhalf_inv_h, dist, 1.f), 0.f);
            const float t2 = t * t;
            const float sg = (k5 * t) * t2;                       // grad W / d
            const float w = (t2 * t2) * fmaf(2.f * inv_h, dist, 1.f);
            float grad[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < ND; d++) grad[d] = sg * p[d];
            const float c2 = fmaf(cp[2], cp[2], fmaf(cp[1], cp[1], cp[0] * cp[0]));
            const float inv_cd = fast_rsqrt(c2);
            float ds = 0.f, term[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int a = 0; a < ND; a++) {
                float tt = 0.f, u = -2.f * cp[a];
#pragma unroll
                for (int b = 0; b < ND; b++) {
                    tt = fmaf(Bi[b * ND + a] + nb[8 + b * ND + a], grad[b], tt);
                    u = fmaf(Fi[b * ND + a] + nb[17 + b * ND + a], p[b], u);
                }
                term[a] = m_j * tt;
                ds = fmaf(u, cp[a], ds);
            }
            // c = alpha/2 V_i V_j W / d0^2 E delta / |x_ij| / m_i, delta = ds / |x_ij|
            const float c = ci * vol_j * w * (inv_d * inv_d) * ds * (inv_cd * inv_cd);
#pragma unroll
            for (int d = 0; d < ND; d++) acc[d] += fmaf(c, cp[d], term[d]);
        }
    }
#pragma unroll
    for (int d = 0; d < ND; d++) {
        float v = acc[d];
#pragma unroll
        for (int o = kG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[d] = v;
    }
    if (have && sub == 0) {
#pragma unroll
        for (int d = 0; d < ND; d++) dv[i * ND + d] = acc[d];
    }
}

// compute_pk1_corrected! per particle (oracle pno_tlsph_pk1_corrected): E = (F^T F - I)/2,
// S = lambda tr(E) I + 2 mu E, P = F S, out = P L; explicit IEEE operations in the oracle's order.
template <int ND>
__global__ void k_tlsph_pk1_corrected(int64_t n, const float *__restrict__ F,
                                      const float *__restrict__ L, float lambda, float mu,
                                      float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    constexpr int NN = ND * ND;
    float Fi[NN], Li[NN], E[NN], S[NN], P[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) { Fi[e] = __ldg(F + i * NN + e); Li[e] = __ldg(L + i * NN + e); }
#pragma unroll
    for (int a = 0; a < ND; a++)
#pragma unroll
        for (int b = 0; b < ND; b++) {
            float t = __fmul_rn(Fi[a * ND], Fi[b * ND]);
#pragma unroll
            for (int k = 1; k < ND; k++) t = __fadd_rn(t, __fmul_rn(Fi[a * ND + k], Fi[b * ND + k]));
            E[b * ND + a] = __fmul_rn(0.5f, __fsub_rn(t, a == b ? 1.f : 0.f));
        }
    float tr = E[0];
#pragma unroll
    for (int a = 1; a < ND; a++) tr = __fadd_rn(tr, E[a * ND + a]);
    const float two_mu = __fmul_rn(2.f, mu), ltr = __fmul_rn(lambda, tr);
#pragma unroll
    for (int a = 0; a < ND; a++)
#pragma unroll
        for (int b = 0; b < ND; b++) {
            float t = __fmul_rn(two_mu, E[b * ND + a]);
            if (a == b) t = __fadd_rn(ltr, t);
            S[b * ND + a] = t;
        }
#pragma unroll
    for (int a = 0; a < ND; a++)
#pragma unroll
        for (int b = 0; b < ND; b++) {
            float t = __fmul_rn(Fi[a], S[b * ND]);
#pragma unroll
            for (int k = 1; k < ND; k++) t = __fadd_rn(t, __fmul_rn(Fi[k * ND + a], S[b * ND + k]));
            P[b * ND + a] = t;
        }
#pragma unroll
    for (int a = 0; a < ND; a++)
#pragma unroll
        for (int b = 0; b < ND; b++) {
            float t = __fmul_rn(P[a], Li[b * ND]);
#pragma unroll
            for (int k = 1; k < ND; k++) t = __fadd_rn(t, __fmul_rn(P[k * ND + a], Li[b * ND + k]));
            out[i * NN + b * ND + a] = t;
        }
}

// compute_pressure! (ContinuityDensity + StateEquationCole): p = B ((rho / rho0)^gamma - 1) + p_bg,
// rho = row ndims + 1 of v.  gamma == 1 (the benchmark) needs no pow.
__global__ void k_wcsph_compute_pressure(int64_t n, int ns, const float *__restrict__ v, float B,
                                         float rho0, float exponent, float background,
                                         float *__restrict__ pressure)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float ratio = __fdiv_rn(__ldg(v + i * ns + (ns - 1)), rho0);
    const float pw = exponent == 1.f ? ratio : (float)pow((double)ratio, (double)exponent);
    pressure[i] = __fadd_rn(__fmul_rn(B, __fsub_rn(pw, 1.f)), background);
}

#endif  // __CUDACC__

}  // namespace pnb
