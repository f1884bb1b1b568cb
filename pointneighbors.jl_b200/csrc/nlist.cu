// nlist.cu -- PrecomputedNeighborhoodSearch on the device: neighbour lists as CSR
// (count pass, 64-bit decoupled-lookback scan, fill pass, per-list sort), exports in the
// reference's DynamicVectorOfVectors layouts, the list sweep without radius test and the TLSPH
// deformation gradient.
//
// reference: src/nhs_precomputed.jl:130-247 (initialize!, initialize_neighbor_lists!, sweep),
// src/vector_of_vectors.jl:3-31,177-212 (layout, sorteach!),
// benchmarks/smoothed_particle_hydrodynamics.jl:136-189 (TLSPH set-up).
#include <cstring>

#include "closures.cuh"
#include "sweep.cuh"

struct pnb_nlist {
    int64_t nx;        // number of lists (points of x)
    int64_t n_pairs;
    int64_t *offsets;  // [nx+1] device
    int32_t *ids;      // [n_pairs] device, 0-based
    uint32_t *counts;  // [nx] device
    int ndims;
    int *d_err;
    int *h_err;
};

namespace pnb {

// count pass: lengths of the lists (the `lengths[i] += 1` half of pushat!, vector_of_vectors.jl:83)
struct ListCountCl {
    static constexpr bool kCountOnly = true;
    static constexpr int kPayBytes = 0;
    static constexpr int kWarpsPerCell = 2;
    uint32_t *out;
    struct State { int cnt; };
    __device__ __forceinline__ void init(State &s, bool, int, int) const { s.cnt = 0; }
    __device__ __forceinline__ void stage(unsigned char *, int, uint32_t, int) const {}
    __device__ __forceinline__ void count(State &s, int c) const { s.cnt += c; }
    __device__ __forceinline__ void merge(State &s, const State &o) const { s.cnt += o.cnt; }
    template <int ND>
    __device__ __forceinline__ void pair(State &, float, float, float, float, int,
                                         const unsigned char *, int, int) const {}
    template <int ND>
    __device__ __forceinline__ void pair_global(State &, float, float, float, float, int,
                                                uint32_t) const {}
    __device__ __forceinline__ void finish(State &s, int, int i_id) const { out[i_id] = (uint32_t)s.cnt; }
};

// fill pass: pushat!(neighbor_lists, point, neighbor)  (nhs_precomputed.jl:198-201)
struct ListFillCl {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 0;
    static constexpr int kWarpsPerCell = 2;
    const int64_t *offsets;
    int32_t *ids;
    struct State { int64_t pos; };
    __device__ __forceinline__ void init(State &s, bool active, int, int i_id) const
    {
        s.pos = active ? offsets[i_id] : 0;
    }
    __device__ __forceinline__ void stage(unsigned char *, int, uint32_t, int) const {}
    __device__ __forceinline__ void count(State &, int) const {}
    template <int ND>
    __device__ __forceinline__ void pair(State &s, float, float, float, float, int j_id,
                                         const unsigned char *, int, int) const
    {
        ids[s.pos++] = j_id;
    }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float, float, float, float, int j_id,
                                                uint32_t) const
    {
        ids[s.pos++] = j_id;
    }
    __device__ __forceinline__ void finish(State &, int, int) const {}
};

template <int ND, bool PER, class CL>
static pnb_status launch_list_nd(pnb_grid *g, bool fast, const float *x, int64_t nx, const CL &cl,
                                 cudaStream_t s)
{
    if (fast) {
        const int nxc = g->p.gs[0] - 2;
        const int nyc = ND > 1 ? g->p.gs[1] - 2 : 1;
        const int nzc = ND > 2 ? g->p.gs[2] - 2 : 1;
        if (nxc <= 0 || nyc <= 0 || nzc <= 0) return PNB_OK;
        const int64_t blocks = (int64_t)div_up(nxc, kTX) * nyc * nzc;
        const size_t smem = sizeof(float4) * kCapPad + (size_t)kCap * CL::kPayBytes;
        ProfScope ps(PH_SWEEP_CELLS, s);
        k_sweep_cells<ND, PER, CL><<<(unsigned)blocks, kCellThreads, smem, s>>>(
            g->p, g->cell_start, g->sorted, cl);
        PNB_LAUNCHED();
    } else if (nx > 0) {
        ProfScope ps(PH_SWEEP_POINTS, s);
        k_sweep_points<ND, PER, CL><<<(unsigned)div_up(nx, 128), 128, 0, s>>>(
            g->p, g->cell_start, g->sorted, x, nx, nullptr, 0, cl, g->d_err);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

template <class CL>
static pnb_status launch_list(pnb_grid *g, bool fast, const float *x, int64_t nx, const CL &cl,
                              cudaStream_t s)
{
    if (g->template_search || g->n_built == 0) return PNB_OK;
    const bool per = g->p.periodic != 0;
    switch (g->p.ndims) {
        case 1: return per ? launch_list_nd<1, true>(g, fast, x, nx, cl, s) : launch_list_nd<1, false>(g, fast, x, nx, cl, s);
        case 2: return per ? launch_list_nd<2, true>(g, fast, x, nx, cl, s) : launch_list_nd<2, false>(g, fast, x, nx, cl, s);
        default: return per ? launch_list_nd<3, true>(g, fast, x, nx, cl, s) : launch_list_nd<3, false>(g, fast, x, nx, cl, s);
    }
}

// sorteach! (vector_of_vectors.jl:177-183): every list ascending.  One warp per list.
//   <= 128 neighbours (the normal case, ~108 at r = 3 spacings): bitonic network in registers,
//      element e = r * 32 + lane lives in register r of lane `lane`; partners at distance < 32
//      are exchanged with __shfl_xor_sync, larger distances are register swaps.
//   <= kSortCap: rank by counting in shared memory.   larger: odd-even transposition in global.
constexpr int kSortCap = 512;
constexpr int kSortWarps = 4;

template <int NR>
__device__ __forceinline__ void bitonic_sort_regs(int32_t (&v)[NR], int lane)
{
#pragma unroll
    for (int k = 2; k <= 32 * NR; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int jr = j >> 5;
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    if ((r & jr) == 0) {
                        const int r2 = r | jr;
                        const bool asc = (((r << 5) | lane) & k) == 0;
                        const int32_t a = v[r], c = v[r2];
                        const bool sw = (a > c) == asc;
                        v[r] = sw ? c : a;
                        v[r2] = sw ? a : c;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < NR; r++) {
                    const int32_t other = __shfl_xor_sync(0xffffffffu, v[r], j);
                    const bool asc = (((r << 5) | lane) & k) == 0;
                    const bool lower = (lane & j) == 0;
                    v[r] = (lower == asc) ? min(v[r], other) : max(v[r], other);
                }
            }
        }
    }
}

template <int NR>
__device__ __forceinline__ void sort_list_regs(int32_t *lst, int len, int lane)
{
    int32_t v[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) v[r] = (r * 32 + lane < len) ? lst[r * 32 + lane] : 0x7fffffff;
    bitonic_sort_regs<NR>(v, lane);
#pragma unroll
    for (int r = 0; r < NR; r++)
        if (r * 32 + lane < len) lst[r * 32 + lane] = v[r];
}

__global__ void __launch_bounds__(kSortWarps * 32)
k_sort_lists(int64_t nx, const int64_t *__restrict__ offsets, int32_t *__restrict__ ids)
{
    __shared__ int32_t s_buf[kSortWarps][kSortCap];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int64_t i = (int64_t)blockIdx.x * kSortWarps + warp;
    if (i >= nx) return;
    const int64_t o0 = offsets[i];
    const int len = (int)(offsets[i + 1] - o0);
    if (len <= 1) return;
    int32_t *lst = ids + o0;
    if (len <= 32) {
        sort_list_regs<1>(lst, len, lane);
    } else if (len <= 64) {
        sort_list_regs<2>(lst, len, lane);
    } else if (len <= 128) {
        sort_list_regs<4>(lst, len, lane);
    } else if (len <= kSortCap) {
        int32_t *buf = s_buf[warp];
        for (int e = lane; e < len; e += 32) buf[e] = lst[e];
        __syncwarp();
        for (int e = lane; e < len; e += 32) {
            const int32_t v = buf[e];
            int r = 0;
            int k = 0;
            for (; k + 4 <= len; k += 4) {
                r += (buf[k] < v) + (buf[k + 1] < v) + (buf[k + 2] < v) + (buf[k + 3] < v);
            }
            for (; k < len; k++) r += (buf[k] < v);
            lst[r] = v;
        }
    } else {
        // rare (> kSortCap neighbours): odd-even transposition in global memory by one warp
        for (int pass = 0; pass < len; pass++) {
            for (int e = (pass & 1) + 2 * lane; e + 1 < len; e += 64) {
                const int32_t a = lst[e], b = lst[e + 1];
                if (a > b) { lst[e] = b; lst[e + 1] = a; }
            }
            __syncwarp();
        }
    }
}

// DynamicVectorOfVectors export (vector_of_vectors.jl:3-31): one warp per list.
__global__ void k_nlist_export_dvov(int64_t nx, const int64_t *__restrict__ offsets,
                                    const int32_t *__restrict__ ids, int32_t *__restrict__ backend,
                                    int32_t *__restrict__ lengths, int max_neighbors,
                                    int transposed, int base, int *__restrict__ err)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nx) return;
    const int lane = lane_id();
    const int64_t o0 = offsets[i];
    int len = (int)(offsets[i + 1] - o0);
    if (len > max_neighbors) {
        if (lane == 0) atomicOr(err, 4);
        len = max_neighbors;
    }
    if (lane == 0) lengths[i] = len;
    for (int e = lane; e < max_neighbors; e += 32) {
        // unused slots = typemax(Int32), like the GPU sorteach! (vector_of_vectors.jl:200-204)
        const int32_t v = e < len ? ids[o0 + e] + base : 0x7fffffff;
        if (transposed) backend[(int64_t)e * nx + i] = v;      // parent is nx x max_neighbors
        else backend[i * (int64_t)max_neighbors + e] = v;      // max_neighbors x nx column-major
    }
}

__global__ void k_nlist_export_csr(int64_t nx1, int64_t n_pairs, const int64_t *__restrict__ off,
                                   const int32_t *__restrict__ ids, int64_t *__restrict__ out_off,
                                   int32_t *__restrict__ out_ids, int base)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_off && i < nx1) out_off[i] = off[i];
    if (out_ids && i < n_pairs) out_ids[i] = ids[i] + base;
}

// mapreduce_neighbor_inner(::PrecomputedNeighborhoodSearch) (nhs_precomputed.jl:210-247):
// pos_diff, d2, periodic fix, distance = sqrt(d2) -- no radius test.  One warp per point.
template <int ND, bool PER>
__global__ void k_nlist_pairs(GridP g, int64_t nx, const int64_t *__restrict__ offsets,
                              const int32_t *__restrict__ ids, const float *__restrict__ x,
                              const float *__restrict__ y, float *__restrict__ pos_diff,
                              float *__restrict__ dist)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nx) return;
    float xi[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < ND; d++) xi[d] = __ldg(x + i * ND + d);
    for (int64_t k = offsets[i] + lane_id(); k < offsets[i + 1]; k += 32) {
        const int64_t j = ids[k];
        float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __fsub_rn(xi[d], __ldg(y + j * ND + d));
        float d2 = dist2<ND>(p[0], p[1], p[2]);
        d2 = maybe_periodic_fix<ND, PER>(make_perp(g), d2, p[0], p[1], p[2]);
        if (pos_diff) {
#pragma unroll
            for (int d = 0; d < ND; d++) pos_diff[k * ND + d] = p[d];
        }
        if (dist) dist[k] = __fsqrt_rn(d2);
    }
}

// TLSPH deformation gradient (formulas: oracle pno_tlsph_deformation_grad; unpinned).
// One warp per point: lanes stride over the point's list (coalesced id reads), gather X0_j, x_j,
// m_j, rho0_j, accumulate the ND x ND outer products privately, then a shuffle tree adds the 32
// partial matrices.  HBM: 4 B per pair + 108 B per point.
template <int ND, bool PER>
__global__ void __launch_bounds__(256)
k_tlsph_defgrad(GridP g, int64_t n, const int64_t *__restrict__ offsets,
                const int32_t *__restrict__ ids, const float *__restrict__ X0,
                const float *__restrict__ xcur, const float *__restrict__ mass,
                const float *__restrict__ rho0, const float *__restrict__ L, float h,
                float kernel_norm, float *__restrict__ F)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    constexpr int NN = ND * ND;
    float Xi[3] = {0.f, 0.f, 0.f}, xi[3] = {0.f, 0.f, 0.f}, Li[NN];
#pragma unroll
    for (int d = 0; d < ND; d++) { Xi[d] = __ldg(X0 + i * ND + d); xi[d] = __ldg(xcur + i * ND + d); }
#pragma unroll
    for (int e = 0; e < NN; e++) Li[e] = __ldg(L + i * NN + e);
    float acc[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) acc[e] = 0.f;
    const float nh = __fdiv_rn(kernel_norm, h);
    for (int64_t k = offsets[i] + lane_id(); k < offsets[i + 1]; k += 32) {
        const int64_t j = ids[k];
        float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __fsub_rn(Xi[d], __ldg(X0 + j * ND + d));
        float d2 = dist2<ND>(p[0], p[1], p[2]);
        d2 = maybe_periodic_fix<ND, PER>(make_perp(g), d2, p[0], p[1], p[2]);
        const float dist = __fsqrt_rn(d2);
        if (dist < PNB_SQRT_EPS_F32) continue;
        const float q = __fdiv_rn(dist, h);
        float w = 0.f;
        if (q < 2.f) {
            const float t = __fsub_rn(1.f, __fmul_rn(q, 0.5f));
            w = __fmul_rn(__fmul_rn(-5.f, q), __fmul_rn(__fmul_rn(t, t), t));
        }
        const float sg = __fdiv_rn(__fmul_rn(nh, w), dist);
        float grad[3] = {0.f, 0.f, 0.f}, lg[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < ND; d++) grad[d] = __fmul_rn(sg, p[d]);
#pragma unroll
        for (int a = 0; a < ND; a++) {
            float t = __fmul_rn(Li[a], grad[0]);
#pragma unroll
            for (int b = 1; b < ND; b++) t = __fadd_rn(t, __fmul_rn(Li[b * ND + a], grad[b]));
            lg[a] = t;
        }
        const float nvol = -__fdiv_rn(__ldg(mass + j), __ldg(rho0 + j));
#pragma unroll
        for (int b = 0; b < ND; b++)
#pragma unroll
            for (int a = 0; a < ND; a++) {
                const float cd = __fsub_rn(xi[a], __ldg(xcur + j * ND + a));
                acc[b * ND + a] = __fadd_rn(acc[b * ND + a], __fmul_rn(__fmul_rn(nvol, cd), lg[b]));
            }
    }
#pragma unroll
    for (int e = 0; e < NN; e++) {
        float v = acc[e];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[e] = v;
    }
    if (lane_id() == 0) {
#pragma unroll
        for (int e = 0; e < NN; e++) F[i * NN + e] = acc[e];
    }
}

static pnb_status nlist_check(pnb_nlist *l, cudaStream_t s)
{
    PNB_CUDA(cudaMemcpyAsync(l->h_err, l->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    int e = *l->h_err;
    if (e == 0) return PNB_OK;
    PNB_CUDA(cudaMemsetAsync(l->d_err, 0, sizeof(int), s));
    set_error("cell list is full. Use a larger `max_points_per_cell`.");
    return PNB_ERR_LIST_FULL;
}

}  // namespace pnb

using namespace pnb;

extern "C" void pnb_nlist_destroy(pnb_nlist *l)
{
    if (!l) return;
    cudaFree(l->offsets);
    cudaFree(l->ids);
    cudaFree(l->counts);
    cudaFree(l->d_err);
    if (l->h_err) cudaFreeHost(l->h_err);
    cudaGetLastError();
    delete l;
}

extern "C" int64_t pnb_nlist_n_points(const pnb_nlist *l) { return l ? l->nx : 0; }
extern "C" int64_t pnb_nlist_n_pairs(const pnb_nlist *l) { return l ? l->n_pairs : 0; }

extern "C" pnb_status pnb_nlist_build_f32(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                          int64_t n, int sort, pnb_nlist **out, void *stream)
{
    (void)y;
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    // nhs_precomputed.jl:136-138: the precomputed search rejects inactive points
    if (!g->template_search && (!g->full_build || n != g->n_y_built)) {
        set_error("this neighborhood search does not support inactive points");
        return PNB_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    pnb_nlist *l = new pnb_nlist();
    memset(l, 0, sizeof(*l));
    l->nx = nx;
    l->ndims = g->p.ndims;
    auto fail = [&](pnb_status st) { pnb_nlist_destroy(l); return st; };
#define NL_CUDA(expr)                                                           \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #expr));             \
    } while (0)
    NL_CUDA(cudaMalloc(&l->offsets, sizeof(int64_t) * (size_t)(nx + 1)));
    NL_CUDA(cudaMalloc(&l->counts, sizeof(uint32_t) * (size_t)(nx + 8)));
    NL_CUDA(cudaMalloc(&l->d_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->d_err, 0, sizeof(int), s));
    NL_CUDA(cudaMallocHost(&l->h_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->counts, 0, sizeof(uint32_t) * (size_t)(nx + 8), s));
    const bool fast = g->full_build && x == g->y_built && nx == g->n_y_built;
    pnb_status st = PNB_OK;
    if (!sort) {
        // unsorted lists keep the visiting order: make it the reproducible one (ids ascending
        // inside every cell); sorted lists do not depend on it
        st = ensure_canonical(g, s);
        if (st != PNB_OK) return fail(st);
    }
    st = launch_list(g, fast, x, nx, ListCountCl{l->counts}, s);
    if (st != PNB_OK) return fail(st);
    st = exclusive_scan_u32_to_i64(g, l->counts, l->offsets, nx, s);
    if (st != PNB_OK) return fail(st);
    int64_t total = 0;
    NL_CUDA(cudaMemcpyAsync(&total, l->offsets + nx, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    NL_CUDA(cudaStreamSynchronize(s));
    l->n_pairs = total;
    NL_CUDA(cudaMalloc(&l->ids, sizeof(int32_t) * (size_t)(total > 0 ? total : 1)));
    st = launch_list(g, fast, x, nx, ListFillCl{l->offsets, l->ids}, s);
    if (st != PNB_OK) return fail(st);
    if (sort && nx > 0 && total > 0) {
        ProfScope ps(PH_NLIST_SORT, s);
        k_sort_lists<<<(unsigned)div_up(nx, kSortWarps), kSortWarps * 32, 0, s>>>(nx, l->offsets,
                                                                                  l->ids);
        cudaError_t e = cudaGetLastError();
        g_launch_count++;
        if (e != cudaSuccess) return fail(cuda_fail(e, "k_sort_lists"));
    }
    st = check_err_word(g, s);
    if (st != PNB_OK) return fail(st);
#undef NL_CUDA
    *out = l;
    return PNB_OK;
}

extern "C" pnb_status pnb_nlist_export_csr(const pnb_nlist *l, int64_t *offsets, int32_t *ids,
                                           int index_base, void *stream)
{
    if (!l) { set_error("list handle is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t m = (l->nx + 1) > l->n_pairs ? (l->nx + 1) : l->n_pairs;
    k_nlist_export_csr<<<(unsigned)div_up(m, 256), 256, 0, s>>>(l->nx + 1, l->n_pairs, l->offsets,
                                                                l->ids, offsets, ids, index_base);
    PNB_LAUNCHED();
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

extern "C" pnb_status pnb_nlist_export_dvov(const pnb_nlist *l_, int32_t *backend,
                                            int32_t *lengths, int32_t max_neighbors,
                                            int transposed, int index_base, void *stream)
{
    pnb_nlist *l = const_cast<pnb_nlist *>(l_);
    if (!l) { set_error("list handle is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        k_nlist_export_dvov<<<(unsigned)div_up(l->nx * 32, 256), 256, 0, s>>>(
            l->nx, l->offsets, l->ids, backend, lengths, max_neighbors, transposed, index_base,
            l->d_err);
        PNB_LAUNCHED();
    }
    return nlist_check(l, s);
}

extern "C" pnb_status pnb_nlist_pairs_f32(const pnb_nlist *l, const pnb_grid *g, const float *x,
                                          const float *y, float *pos_diff, float *distance,
                                          void *stream)
{
    if (!l || !g) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        const unsigned blocks = (unsigned)div_up(l->nx * 32, 256);
        const bool per = g->p.periodic != 0;
        ProfScope ps(PH_NLIST_SWEEP, s);
#define PAIRS(ND)                                                                                  \
    if (per) k_nlist_pairs<ND, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, x, y, pos_diff, distance); \
    else k_nlist_pairs<ND, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, x, y, pos_diff, distance)
        switch (l->ndims) {
            case 1: PAIRS(1); break;
            case 2: PAIRS(2); break;
            default: PAIRS(3); break;
        }
#undef PAIRS
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

extern "C" pnb_status pnb_tlsph_deformation_grad_f32(const pnb_nlist *l, const pnb_grid *g,
                                                     const float *X0, const float *xcur,
                                                     const float *mass, const float *rho0,
                                                     const float *L, float smoothing_length,
                                                     float kernel_norm, float *F, void *stream)
{
    if (!l || !g) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        const unsigned blocks = (unsigned)div_up(l->nx * 32, 256);
        const bool per = g->p.periodic != 0;
        ProfScope ps(PH_NLIST_SWEEP, s);
#define DEFGRAD(ND)                                                                                 \
    if (per) k_tlsph_defgrad<ND, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, X0, xcur, mass, rho0, L, smoothing_length, kernel_norm, F); \
    else k_tlsph_defgrad<ND, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, X0, xcur, mass, rho0, L, smoothing_length, kernel_norm, F)
        switch (l->ndims) {
            case 1: DEFGRAD(1); break;
            case 2: DEFGRAD(2); break;
            default: DEFGRAD(3); break;
        }
#undef DEFGRAD
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}
