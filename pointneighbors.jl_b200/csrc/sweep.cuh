// sweep.cuh -- the 3^d neighbour-cell sweep kernels behind foreach_point_neighbor.
//
// reference: foreach_point_neighbor (src/neighborhood_search.jl:183-201) -> foreach_neighbor
// (:250-276) -> mapreduce_neighbor_inner(::GridNeighborhoodSearch) (src/nhs_grid.jl:519-575).
//
// Two kernels share the closures (functors) defined in closures.cuh:
//
//  k_sweep_cells   the fast path for x === y (the benchmarked case).  A CTA owns TX consecutive
//                  cells of one x-row of the grid; the 3^(d-1) neighbour rows are contiguous
//                  ranges of the cell-ordered coordinate array, staged row by row into shared
//                  memory (coalesced float4 copies).  One warp per cell, one lane per point of
//                  the cell: the lane keeps x_i and the closure's accumulators in registers and
//                  walks the staged candidates with broadcast shared-memory loads.  Hits are
//                  recorded as bits of a per-lane mask while testing a block of 32 candidates
//                  and the (expensive) interaction runs afterwards, once per set bit, so the
//                  interaction code is not executed under an 85 % idle mask.
//                  Candidates are visited in exactly the reference's order (neighbour cells in
//                  CartesianIndices order, ids ascending inside a cell) so floating point sums
//                  are reproducible run to run and comparable with the oracle.
//
//  k_sweep_points  the general path (x != y, or a `points` subset): one thread per query
//                  point, candidates read from the cell-ordered array through L1/L2.
#pragma once

#include "grid.cuh"

namespace pnb {

constexpr int kTX = 8;            // cells (= warps) per CTA in k_sweep_cells
constexpr int kCellThreads = kTX * 32;
constexpr int kCap = 512;         // staged candidates per chunk
constexpr int kCapPad = kCap + 32; // position buffer is read up to 31 slots past the chunk (masked)
constexpr int kSlots = kTX + 2;   // x-neighbour cells of a tile in one row

// One block of 32 staged candidates against the lane's own point: bit k of the result is set
// iff candidate k is within the search radius (same operations, same order as
// src/nhs_grid.jl:547-555).  All lanes read the same address: broadcast LDS.128.
template <int ND, bool PER>
__device__ __forceinline__ unsigned test_block(const PerP &g, const float4 *__restrict__ cp,
                                               float xi, float yi, float zi)
{
    unsigned hits = 0u;
#pragma unroll
    for (int k = 0; k < 32; k++) {
        const float4 pj = cp[k];
        float px = __fsub_rn(xi, pj.x);
        float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
        float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
        float d2 = dist2<ND>(px, py, pz);
        d2 = maybe_periodic_fix<ND, PER>(g, d2, px, py, pz);
        if (d2 <= g.r2) hits |= 1u << k;
    }
    return hits;
}

// Candidate views handed to the closures --------------------------------------------------------
// Shared-memory view: payload planes are arrays of kCap elements.
// Global view: payload planes are the cell-ordered arrays themselves.

// (q_start, q_sorted): cell-ordered query points; the same arrays as (cell_start, sorted) for
// x === y, the copy made by build_query_list for a second point set.
template <int ND, bool PER, class CL, int TX>
__device__ __forceinline__ void
sweep_tile_rows(const GridP &g, const CellsView &cand, const CellsView &qry, const CL &cl,
                int64_t tile)
{
    const float4 *__restrict__ sorted = cand.rec;
    const float4 *__restrict__ q_sorted = qry.rec;
    const bool two = q_sorted != sorted;
    constexpr int kSlotsT = TX + 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_pos = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *s_pay = smem_raw + sizeof(float4) * kCapPad;
    __shared__ uint32_t s_begin[kSlotsT];   // global begin of each slot's cell
    __shared__ uint32_t s_prefix[kSlotsT + 1];

    // ---- which tile ---------------------------------------------------------------------
    // valid (non-padding) cells are 2 .. gs-1 in every used dimension
    const int nx = g.gs[0] - 2;
    const int ny = ND > 1 ? g.gs[1] - 2 : 1;
    const int ntx = (nx + TX - 1) / TX;
    int64_t b = tile;
    const int tx = (int)(b % ntx); b /= ntx;
    const int iy = (int)(b % ny);  b /= ny;
    const int iz = (int)b;
    const int cx0 = 2 + tx * TX;
    const int cx1 = min(cx0 + TX - 1, g.gs[0] - 1);
    const int cy = ND > 1 ? 2 + iy : 1;
    const int cz = ND > 2 ? 2 + iz : 1;

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const PerP pp = make_perp(g);

    // points of this tile are one contiguous range of the cell-ordered array
    {
        uint32_t n_tile = 0;   // query points of the tile (uniform for the CTA)
        for (int cx = cx0; cx <= cx1; cx++) {
            uint32_t b0, cnt;
            cell_range(qry, linear_cell(g, cx, cy, cz), b0, cnt);
            n_tile += cnt;
        }
        if (n_tile == 0) return;
    }

    // work items of the tile: (cell, pass of 32 points); a batch gives every warp one item
    const int my_cx = cx0 + warp;
    uint32_t c_p0 = 0, c_p1 = 0;
    if (my_cx <= cx1) {
        uint32_t cnt;
        cell_range(qry, linear_cell(g, my_cx, cy, cz), c_p0, cnt);
        c_p1 = c_p0 + cnt;
    }
    const int my_passes = (int)((c_p1 - c_p0 + 31) / 32);
    int max_passes = my_passes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_passes = max(max_passes, __shfl_xor_sync(0xffffffffu, max_passes, o));
    __shared__ int s_maxpass[TX];
    if (lane == 0) s_maxpass[warp] = max_passes;
    __syncthreads();
    int n_batches = 0;
#pragma unroll
    for (int w = 0; w < TX; w++) n_batches = max(n_batches, s_maxpass[w]);

    for (int batch = 0; batch < n_batches; batch++) {
        const uint32_t i_sorted = c_p0 + (uint32_t)batch * 32u + (uint32_t)lane;
        const bool active = i_sorted < c_p1;
        float xi = 0.f, yi = 0.f, zi = 0.f;
        int i_id = 0;
        typename CL::State st;
        if (active) {
            const float4 pi = q_sorted[i_sorted];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
        }
        cl.init(st, active, two ? -1 : (int)i_sorted, i_id);
        const bool warp_active = __any_sync(0xffffffffu, active);

        // ---- neighbour rows in CartesianIndices order: dz outer, dy inner, dx = slots ------
        for (int oz = (ND > 2 ? -1 : 0); oz <= (ND > 2 ? 1 : 0); oz++) {
            for (int oy = (ND > 1 ? -1 : 0); oy <= (ND > 1 ? 1 : 0); oy++) {
                int ry = cy + oy, rz = cz + oz;
                if (PER) {
                    if (ND > 1) ry = floormod_i(ry - 2, g.nc[1]) + 2;
                    if (ND > 2) rz = floormod_i(rz - 2, g.nc[2]) + 2;
                }
                __syncthreads();   // previous row fully consumed; s_begin/s_prefix reusable
                // slot s <-> cell x = cx0 - 1 + s (wrapped when periodic)
                if (threadIdx.x < kSlotsT) {
                    int sx = cx0 - 1 + (int)threadIdx.x;
                    uint32_t b0 = 0, cnt = 0;
                    if (sx <= cx1 + 1) {
                        if (PER) sx = floormod_i(sx - 2, g.nc[0]) + 2;
                        cell_range(cand, linear_cell(g, sx, ry, rz), b0, cnt);
                    }
                    s_begin[threadIdx.x] = b0;
                    // exclusive prefix over <= 10 slots, done by the first warp
                    uint32_t incl = cnt;
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) {
                        uint32_t t = __shfl_up_sync((1u << kSlotsT) - 1u, incl, o);
                        if ((int)threadIdx.x >= o) incl += t;
                    }
                    s_prefix[threadIdx.x + 1] = incl;
                    if (threadIdx.x == 0) s_prefix[0] = 0;
                }
                __syncthreads();
                const uint32_t total = s_prefix[kSlotsT];
                // this warp's candidates in the row: slots warp .. warp+2 (cells cx-1 .. cx+1)
                const uint32_t w_q0 = s_prefix[warp], w_q1 = s_prefix[min(warp + 3, kSlotsT)];

                for (uint32_t q0 = 0; q0 < total; q0 += kCap) {
                    const uint32_t q1 = min(q0 + (uint32_t)kCap, total);
                    if (q0 > 0) __syncthreads();
                    // ---- stage chunk [q0, q1) -------------------------------------------
                    for (uint32_t q = q0 + threadIdx.x; q < q1; q += TX * 32) {
                        int s = 0;
#pragma unroll
                        for (int t = 1; t < kSlotsT; t++) s += (q >= s_prefix[t]) ? 1 : 0;
                        const uint32_t gi = s_begin[s] + (q - s_prefix[s]);
                        s_pos[q - q0] = sorted[gi];
                        cl.stage(s_pay, (int)(q - q0), gi, kCap);
                    }
                    __syncthreads();
                    // ---- test + interact -------------------------------------------------
                    // Super-blocks of 128 candidates: four 32-bit hit masks per lane are filled by
                    // fully unrolled test blocks (candidates past the warp's range are masked off,
                    // so there is no remainder loop), then drained bit by bit.
                    if (warp_active) {
                        const uint32_t a0 = max(w_q0, q0), a1 = min(w_q1, q1);
                        for (uint32_t sb = a0; sb < a1; sb += 128) {
                            unsigned m[4];
#pragma unroll
                            for (int bb = 0; bb < 4; bb++) {
                                const uint32_t blk = sb + 32u * bb;
                                m[bb] = 0u;
                                if (blk < a1) {   // warp-uniform
                                    unsigned hh = test_block<ND, PER>(pp, s_pos + (blk - q0), xi, yi, zi);
                                    const uint32_t nv = a1 - blk;
                                    if (nv < 32u) hh &= (1u << nv) - 1u;
                                    m[bb] = active ? hh : 0u;
                                }
                            }
                            if (CL::kCountOnly) {
                                cl.count(st, __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]));
                            } else {
                                unsigned any = m[0] | m[1] | m[2] | m[3];
                                while (__any_sync(0xffffffffu, any != 0u)) {
                                    if (any) {
                                        int k;
                                        if (m[0]) { k = __ffs(m[0]) - 1; m[0] &= m[0] - 1u; }
                                        else if (m[1]) { k = 31 + __ffs(m[1]); m[1] &= m[1] - 1u; }
                                        else if (m[2]) { k = 63 + __ffs(m[2]); m[2] &= m[2] - 1u; }
                                        else { k = 95 + __ffs(m[3]); m[3] &= m[3] - 1u; }
                                        const int slot = (int)(sb - q0) + k;
                                        const float4 pj = s_pos[slot];
                                        float px = __fsub_rn(xi, pj.x);
                                        float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                                        float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                                        float d2 = dist2<ND>(px, py, pz);
                                        d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                                        cl.template pair<ND>(st, px, py, pz, d2,
                                                             __float_as_int(pj.w), s_pay, slot, kCap);
                                        any = m[0] | m[1] | m[2] | m[3];
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (active) cl.finish(st, two ? -1 : (int)i_sorted, i_id);
        __syncthreads();
    }
}


// x === y fast path, ordered variant: one CTA per tile of kTX cells.
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(kCellThreads)
k_sweep_cells(GridP g, const uint32_t *__restrict__ cell_start, const float4 *__restrict__ sorted,
              CL cl)
{
    const CellsView v{cell_start, sorted, 0u};
    sweep_tile_rows<ND, PER, CL, kTX>(g, v, v, cl, (int64_t)blockIdx.x);
}

// General path: one thread per query point ------------------------------------------------------
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(128)
k_sweep_points(GridP g, const uint32_t *__restrict__ cell_start, const float4 *__restrict__ sorted,
               const float *__restrict__ x, int64_t n_loop, const int32_t *__restrict__ points,
               int base, CL cl, int *__restrict__ err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_loop) return;
    const int i_id = points ? points[t] - base : (int)t;
    float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = __ldg(x + (int64_t)i_id * ND + d);
    int cc[3];
#pragma unroll
    for (int d = 0; d < ND; d++) cc[d] = cell_coord(p[d], g.minc[d], g.cs[d], g.periodic, g.nc[d], g.off[d]);
#pragma unroll
    for (int d = ND; d < 3; d++) cc[d] = 1;
    typename CL::State st;
    cl.init(st, true, -1, i_id);
    const PerP pp = make_perp(g);
    bool oob = false;
    for (int oz = (ND > 2 ? -1 : 0); oz <= (ND > 2 ? 1 : 0); oz++)
        for (int oy = (ND > 1 ? -1 : 0); oy <= (ND > 1 ? 1 : 0); oy++)
            for (int ox = -1; ox <= 1; ox++) {
                int c0 = cc[0] + ox, c1 = cc[1] + oy, c2 = cc[2] + oz;
                if (PER) {
                    c0 = floormod_i(c0 - 2, g.nc[0]) + 2;
                    if (ND > 1) c1 = floormod_i(c1 - 2, g.nc[1]) + 2;
                    if (ND > 2) c2 = floormod_i(c2 - 2, g.nc[2]) + 2;
                }
                // the safe variant's bounds check on the neighbour cell (nhs_grid.jl:530-532)
                if (c0 < 1 || c0 > g.gs[0] || c1 < 1 || c1 > g.gs[1] || c2 < 1 || c2 > g.gs[2]) {
                    oob = true;
                    continue;
                }
                const int lin = linear_cell(g, c0, c1, c2);
                const uint32_t b0 = cell_start[lin], b1 = cell_start[lin + 1];
                for (uint32_t gi = b0; gi < b1; gi++) {
                    const float4 pj = __ldg(sorted + gi);
                    float px = __fsub_rn(p[0], pj.x);
                    float py = ND > 1 ? __fsub_rn(p[1], pj.y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(p[2], pj.z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    if (d2 <= g.r2) {
                        if (CL::kCountOnly) cl.count(st, 1);
                        else cl.template pair_global<ND>(st, px, py, pz, d2, __float_as_int(pj.w), gi);
                    }
                }
            }
    if (oob) atomicOr(err, 2);
    cl.finish(st, -1, i_id);
}

}  // namespace pnb
