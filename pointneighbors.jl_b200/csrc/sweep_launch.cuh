// sweep_launch.cuh -- host-side dispatch of the sweep kernels (shared by sweep.cu and nlist.cu).
#pragma once

#include "sweep.cuh"
#include "sweep_tiles.cuh"
#include "sweep_flat.cuh"
#include "hashgrid.cuh"

namespace pnb {

extern int g_tune_wpc;    // pnb_set_tuning: warps per cell override (0 = closure default)
extern int g_tune_half;   // pnb_set_tuning: 0 = exact Float32 test instead of the fp16 pre-filter
extern int g_tune_twoset; // pnb_set_twoset_tiles: 0 = x != y always uses the per-point kernel
extern int g_tune_left;   // pnb_set_sweep_left / PNB_SWEEP_LEFT: 0 never, 1 default, 2 always (tests)
extern int g_reserve_ctas; // pnb_set_sweep_reserve: SMs the persistent sweep leaves to other streams' kernels
extern int g_tune_flat;   // pnb_set_sweep_kernel / PNB_SWEEP_FLAT: 1 (default) k_sweep_flat except for
                          // count-only closures, 2 always k_sweep_flat, 0 always k_sweep_tiles

template <class K>
static pnb_status allow_smem(K kernel, size_t smem)
{
    // static + dynamic shared memory above 48 KB needs the opt-in; the kernels carry up to
    // ~17 KB of static shared memory, so opt in whenever the dynamic part alone exceeds 28 KB
    if (smem > 28 * 1024)
        PNB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return PNB_OK;
}

// two point sets (x != y) below this many query points use the per-point kernel: building the
// query cell list does not pay off
constexpr int64_t kTwoSetMinPoints = 4096;
constexpr double kTwoSetMinPerCell = 12.0;   // query points per occupied cell

// exclusive prefix of the tiles per row segment (a few thousand entries): one block
static __global__ void __launch_bounds__(1024)
k_flat_scan(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, int64_t n)
{
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0u;
    __syncthreads();
    for (int64_t i0 = 0; i0 < n; i0 += 1024) {
        const int64_t i = i0 + threadIdx.x;
        const uint32_t v = i < n ? in[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_w[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t base = s_run + s_w[warp];
        if (i < n) out[i] = base + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = base + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_run;
}

// The flat tile sweep over the cell layers [lay0, lay0 + n_lay) of the last used dimension
// (0-based among the valid layers; all of them for an ordinary sweep, one part of a slab for the
// overlapped multi-GPU step): tile table (3 small kernels), persistent sweep, overflow tiles.
template <int ND, bool PER, class CL, bool TWO>
static pnb_status launch_flat(pnb_grid *g, const CellsView &cand, const CellsView &qry, int64_t n_q,
                              const CL &cl, int lay0, int n_lay, cudaStream_t s, int lay1 = 0,
                              int n_lay1 = 0)
{
    const int nxc = g->p.gs[0] - 2;
    const int nyc = ND > 1 ? g->p.gs[1] - 2 : 1;
    if (ND == 1) { lay0 = 0; n_lay = 1; n_lay1 = 0; }
    if (n_lay < 0) n_lay = 0;
    if (n_lay1 < 0) n_lay1 = 0;
    if (nxc <= 0 || n_lay + n_lay1 <= 0 || n_q <= 0) return PNB_OK;
    const int n_seg_row = (int)div_up(nxc, kFSegCells);
    const int64_t rows = ND == 3 ? (int64_t)nyc * (n_lay + n_lay1) : (ND == 2 ? n_lay + n_lay1 : 1);
    const int64_t n_segs = (int64_t)n_seg_row * rows;
    // non-empty tiles: at most one per query point, and at most the weighted length / 96 + 1 per segment
    int64_t max_tiles = (n_q + (int64_t)kFMinW * nxc * rows) / kFTP + n_segs + 1;
    if (max_tiles > n_q) max_tiles = n_q;
    if (max_tiles > g->flat_tiles_cap) {
        cudaFree(g->flat_tiles); cudaFree(g->flat_ovf); cudaFree(g->flat_tabs);
        g->flat_tiles = nullptr; g->flat_ovf = nullptr; g->flat_tabs = nullptr; g->flat_tiles_cap = 0;
        const int64_t want = max_tiles + max_tiles / 8 + 64;
        PNB_CUDA(cudaMalloc(&g->flat_tiles, sizeof(FlatTile) * (size_t)want));
        PNB_CUDA(cudaMalloc(&g->flat_ovf, sizeof(int) * (size_t)want));
        PNB_CUDA(cudaMalloc(&g->flat_tabs, sizeof(uint32_t) * kTabWords * (size_t)want));
        g->flat_tiles_cap = want;
    }
    if (n_segs > g->flat_seg_cap) {
        cudaFree(g->flat_seg);
        g->flat_seg = nullptr; g->flat_seg_cap = 0;
        PNB_CUDA(cudaMalloc(&g->flat_seg, sizeof(uint32_t) * 2 * (size_t)(n_segs + 65)));
        g->flat_seg_cap = n_segs + 64;
    }
    if (!g->flat_ctl) {
        PNB_CUDA(cudaMalloc(&g->flat_ctl, sizeof(uint32_t) * (kFlatCtl + kFlatSmStates)));
        PNB_CUDA(cudaMemsetAsync(g->flat_ctl, 0, sizeof(uint32_t) * (kFlatCtl + kFlatSmStates), s));
    }
    uint32_t *seg_tiles = g->flat_seg, *seg_first = g->flat_seg + (g->flat_seg_cap + 1);
    FlatTile *tiles = reinterpret_cast<FlatTile *>(g->flat_tiles);
    constexpr int threads = kFG * flat_wpc<CL>() * 32;
    constexpr size_t smem = flat_smem_bytes<ND, CL>();
    // closures whose term is identically zero beyond the search radius (WCSPH with 2 h <= r) run
    // the variant without the exact radius test in the drain
    bool nor2 = false;
    if constexpr (has_no_radius_test<CL>::value) nor2 = cl.no_radius_test();
    auto kern = k_sweep_flat<ND, PER, CL, TWO, false>;
    if constexpr (has_no_radius_test<CL>::value) { if (nor2) kern = k_sweep_flat<ND, PER, CL, TWO, true>; }
    static int ctas_per_sm = 0;          // per instantiation
    static bool smem_allowed[2] = {false, false};   // per instantiation and variant
    if (!smem_allowed[nor2 ? 1 : 0]) {
        pnb_status st = allow_smem(kern, smem);     // (both variants share the occupancy)
        if (st != PNB_OK) return st;
        smem_allowed[nor2 ? 1 : 0] = true;
    }
    if (ctas_per_sm == 0) {
        int nb = 0;
        PNB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem));
        if (nb < 1) { set_error("k_sweep_flat does not fit an SM (%zu bytes of shared memory)", smem); return PNB_ERR_CUDA; }
        ctas_per_sm = nb;
    }
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, g->device);
    // pnb_set_sweep_reserve: SMs the persistent grid leaves to the kernels of other streams
    const unsigned n_ctas = (unsigned)(n_sm * ctas_per_sm);
    const int reserve_sms = g_reserve_ctas > 0 ? (g_reserve_ctas < n_sm / 2 ? g_reserve_ctas : n_sm / 2) : 0;
    {
        ProfScope ps(PH_SWEEP_TILES_PREP, s);
        const unsigned pb = (unsigned)div_up(n_segs, 4);
        k_flat_tiles<ND, false><<<pb, 128, 0, s>>>(g->p, qry, lay0, n_lay, lay1, n_lay1, n_seg_row, seg_tiles, seg_first,
                                                   tiles, g->flat_ctl, reserve_sms > 0 ? 0u : n_ctas);
        PNB_LAUNCHED();
        k_flat_scan<<<1, 1024, 0, s>>>(seg_tiles, seg_first, n_segs);
        PNB_LAUNCHED();
        k_flat_tiles<ND, true><<<pb, 128, 0, s>>>(g->p, qry, lay0, n_lay, lay1, n_lay1, n_seg_row, seg_tiles, seg_first,
                                                  tiles, g->flat_ctl, reserve_sms > 0 ? 0u : n_ctas);
        PNB_LAUNCHED();
    }
    {
        ProfScope ps(PH_SWEEP_TILES_PREP, s);
        k_flat_tables<ND, PER, flat_cap<CL>(), flat_nblk_max<CL>()><<<(unsigned)div_up(max_tiles, 8), 256, 0, s>>>(
            g->p, cand, qry, tiles, g->flat_ctl, g->flat_tabs);
        PNB_LAUNCHED();
    }
    {
        ProfScope ps(PH_SWEEP_CELLS, s);
        kern<<<n_ctas, threads, smem, s>>>(g->p, cand, qry, cl, g->flat_tabs, g->flat_ctl, g->flat_ovf, reserve_sms);
        PNB_LAUNCHED();
    }
    {
        ProfScope ps(PH_SWEEP_OVERFLOW, s);
        k_sweep_flat_overflow<ND, PER, CL, TWO><<<(unsigned)(n_sm * 8), 128, 0, s>>>(
            g->p, cand, qry, cl, tiles, g->flat_ctl, g->flat_ovf);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

template <int ND, bool PER, class CL>
static pnb_status launch_nd(pnb_grid *g, bool fast, bool tiles, const float *x, int64_t n_loop,
                            const int32_t *points, int base, const CL &cl, cudaStream_t s)
{
    // x != y, all points looped over: bin the query points into the grid's cells
    // (build_query_list) and run the tile kernel with queries from that copy
    bool two = !fast && tiles && points == nullptr && n_loop >= kTwoSetMinPoints &&
               g_tune_twoset != 0;
    if (two) {
        double per_cell = 0.0;
        pnb_status stq = build_query_list(g, x, n_loop, &per_cell, s);
        if (stq != PNB_OK) return stq;
        // a lane per query point of a cell: sparse query sets are better served per point
        if (per_cell < kTwoSetMinPerCell && g_tune_twoset != 2) two = false;
    }
    if (fast || two) {
        // candidates: buckets or CSR, whatever the last build wrote; queries: the same list
        // (x === y) or the CSR copy of the second point set
        const CellsView cand = cells_view(g);
        const CellsView qry = two ? CellsView{g->xq_start, g->xq_sorted, 0u} : cand;
        const int nxc = g->p.gs[0] - 2;
        const int nyc = ND > 1 ? g->p.gs[1] - 2 : 1;
        const int nzc = ND > 2 ? g->p.gs[2] - 2 : 1;
        if (nxc <= 0 || nyc <= 0 || nzc <= 0) return PNB_OK;
        const size_t smem_rows = sizeof(float4) * kCapPad + (size_t)kCap * CL::kPayBytes;
        if (!tiles) {
            pnb_status stc = ensure_csr(g, s);     // the ordered kernel walks the CSR arrays
            if (stc != PNB_OK) return stc;
            const int64_t blocks = (int64_t)div_up(nxc, kTX) * nyc * nzc;
            ProfScope ps(PH_SWEEP_CELLS, s);
            k_sweep_cells<ND, PER, CL><<<(unsigned)blocks, kCellThreads, smem_rows, s>>>(
                g->p, g->cell_start, g->sorted, cl);
            PNB_LAUNCHED();
            return PNB_OK;
        }
        // count-only closures spend their time in the test phase, where the 4-cell tiles of
        // round 1 stage less per point and need no tile table (config 3 count: 4.98 vs 5.29 ms,
        // config 1: 0.117 vs 0.132 ms): they keep k_sweep_tiles unless the flat kernel is forced
        // (measured with the table pipeline: config 3 count 4.80 + 0.07 ms flat vs 4.95 ms; config 1
        // 0.095 + 0.012 vs 0.096 ms: the fixed cost of the tile pre-pass decides on small clouds)
        if (g_tune_flat == 2 || (g_tune_flat == 1 && (!CL::kCountOnly || n_loop >= 2000000))) {
            if (two) return launch_flat<ND, PER, CL, true>(g, cand, qry, n_loop, cl, 0, ND == 2 ? nyc : nzc, s);
            return launch_flat<ND, PER, CL, false>(g, cand, qry, n_loop, cl, 0, ND == 2 ? nyc : nzc, s);
        }
        const int64_t blocks = (int64_t)div_up(nxc, kFTX) * nyc * nzc;
        if (blocks > g->ovf_cap) {
            cudaFree(g->ovf_tiles);
            g->ovf_tiles = nullptr;
            g->ovf_cap = 0;
            PNB_CUDA(cudaMalloc(&g->ovf_tiles, sizeof(int) * (size_t)blocks));
            g->ovf_cap = blocks;
        }
        if (!g->ovf_count) PNB_CUDA(cudaMalloc(&g->ovf_count, 2 * sizeof(int)));
        PNB_CUDA(cudaMemsetAsync(g->ovf_count, 0, 2 * sizeof(int), s));
        // list of the surplus points (at most kLeftMax of the >= 33 points of a cell).  Only for
        // closures with a drain worth saving (n-body, WCSPH) on clouds large enough to amortise
        // one more launch: count-only closures repeat just the test phase (config 1: 0.123 ms
        // without, 0.151 ms with the extra kernel), rank-dependent list fills would run it with one
        // thread per point (config 4 list build: +0.1 ms).
        const bool use_left = g_tune_left != 0 && !CL::kCountOnly && !needs_exact_masks<CL>::value &&
                              (n_loop >= 200000 || g_tune_left == 2);
        const int64_t left_need = n_loop / 4 + 64;
        if (use_left && left_need > g->left_cap) {
            cudaFree(g->left_ids);
            g->left_ids = nullptr;
            g->left_cap = 0;
            PNB_CUDA(cudaMalloc(&g->left_ids, sizeof(int) * (size_t)left_need));
            g->left_cap = left_need;
        }
        int *left_ids = use_left ? g->left_ids : nullptr;
        // variant: warps per cell and the fp16 pre-filter; the tuning overrides (pnb_set_tuning)
        // exist for A/B measurements of the 3-D non-periodic x === y kernels
        int wpc = CL::kWarpsPerCell;
        bool half = true;
        if (ND == 3 && !PER && !two) {
            if (g_tune_wpc == 2 || g_tune_wpc == 4) wpc = g_tune_wpc;
            if (g_tune_half == 0) half = false;
        }
        pnb_status st = PNB_OK;
#define PNB_TILES(WPC, HALF, TWO)                                                                 \
    do {                                                                                          \
        constexpr size_t smem = tiles_smem_bytes<ND, CL, HALF>();                                 \
        st = allow_smem(k_sweep_tiles<ND, PER, CL, WPC, HALF, TWO>, smem);                        \
        if (st != PNB_OK) return st;                                                              \
        ProfScope ps(PH_SWEEP_CELLS, s);                                                          \
        k_sweep_tiles<ND, PER, CL, WPC, HALF, TWO><<<(unsigned)blocks, kFTX * WPC * 32, smem, s>>>( \
            g->p, cand, qry, cl, g->ovf_tiles, g->ovf_count, left_ids);                           \
        PNB_LAUNCHED();                                                                           \
    } while (0)
        if (two) {
            PNB_TILES(CL::kWarpsPerCell, true, true);
        } else if constexpr (ND == 3 && !PER) {
            if (wpc == 2 && half) PNB_TILES(2, true, false);
            else if (wpc == 2) PNB_TILES(2, false, false);
            else if (half) PNB_TILES(4, true, false);
            else PNB_TILES(4, false, false);
        } else {
            PNB_TILES(CL::kWarpsPerCell, true, false);
        }
#undef PNB_TILES
        {
            ProfScope ps(PH_SWEEP_OVERFLOW, s);
            k_sweep_overflow<ND, PER, CL><<<148 * 2, kFTX * 32, smem_rows, s>>>(
                g->p, cand, qry, cl, g->ovf_tiles, g->ovf_count);
            PNB_LAUNCHED();
            if (left_ids) {
                k_sweep_left<ND, PER, CL><<<148 * 16, 128, 0, s>>>(g->p, cand, x, left_ids,
                                                                 g->ovf_count + 1, cl);
                PNB_LAUNCHED();
            }
        }
    } else if (n_loop > 0) {
        pnb_status stc = ensure_csr(g, s);         // the per-point kernel walks the CSR arrays
        if (stc != PNB_OK) return stc;
        ProfScope ps(PH_SWEEP_POINTS, s);
        k_sweep_points<ND, PER, CL><<<(unsigned)div_up(n_loop, 128), 128, 0, s>>>(
            g->p, g->cell_start, g->sorted, x, n_loop, points, base, cl, g->d_err);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

// tiles = true: the throughput kernel (any visiting order); false: the ordered kernel whose
// candidate order is the reference's (needed for bit-identical sums in exact mode).
template <class CL>
static pnb_status launch_sweep(pnb_grid *g, bool fast, bool tiles, const float *x, int64_t n_loop,
                               const int32_t *points, int base, const CL &cl, cudaStream_t s)
{
    if (g->template_search || g->n_built == 0) return PNB_OK;  // every neighbourhood is empty
    const bool per = g->p.periodic != 0;
    if (g->hashed) {
        // SpatialHashingCellList: always one thread per query point over the hash table
        if (n_loop <= 0) return PNB_OK;
        ProfScope ps(PH_SWEEP_POINTS, s);
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        // x === y: loop over the table's own records (cell order) instead of the input order
#define PNB_HASH(ND)                                                                               \
    if (fast && per) k_sweep_points_hash<ND, true, CL, true><<<blocks, 128, 0, s>>>(               \
        g->p, g->cell_start, g->sorted, g->hmeta, x, n_loop, points, base, cl, g->d_err);          \
    else if (fast) k_sweep_points_hash<ND, false, CL, true><<<blocks, 128, 0, s>>>(                \
        g->p, g->cell_start, g->sorted, g->hmeta, x, n_loop, points, base, cl, g->d_err);          \
    else if (per) k_sweep_points_hash<ND, true, CL, false><<<blocks, 128, 0, s>>>(                 \
        g->p, g->cell_start, g->sorted, g->hmeta, x, n_loop, points, base, cl, g->d_err);          \
    else k_sweep_points_hash<ND, false, CL, false><<<blocks, 128, 0, s>>>(                         \
        g->p, g->cell_start, g->sorted, g->hmeta, x, n_loop, points, base, cl, g->d_err)
        switch (g->p.ndims) {
            case 1: PNB_HASH(1); break;
            case 2: PNB_HASH(2); break;
            default: PNB_HASH(3); break;
        }
#undef PNB_HASH
        PNB_LAUNCHED();
        return PNB_OK;
    }
    switch (g->p.ndims) {
        case 1:
            return per ? launch_nd<1, true>(g, fast, tiles, x, n_loop, points, base, cl, s)
                       : launch_nd<1, false>(g, fast, tiles, x, n_loop, points, base, cl, s);
        case 2:
            return per ? launch_nd<2, true>(g, fast, tiles, x, n_loop, points, base, cl, s)
                       : launch_nd<2, false>(g, fast, tiles, x, n_loop, points, base, cl, s);
        default:
            return per ? launch_nd<3, true>(g, fast, tiles, x, n_loop, points, base, cl, s)
                       : launch_nd<3, false>(g, fast, tiles, x, n_loop, points, base, cl, s);
    }
}


}  // namespace pnb
