// sweep_tiles.cuh -- the throughput variant of the x === y sweep (default, "fast" arithmetic mode).
//
// Why a second kernel: ncu on the row-by-row kernel (profiles/r1_wcsph_sweep_v2_*) showed the
// deferred interaction loop running with 10 of 32 lanes active.  A lane's hits inside ONE
// neighbour row depend strongly on where its point sits in the cell (near the +y face -> many
// hits in the dy = +1 rows, few in dy = -1), while its hits summed over ALL 3^(d-1) rows are
// nearly the same for every lane.  So this kernel
//   * stages the candidates of all rows of a tile at once (slot-major: for every x-column of
//     the tile's (TX+2) columns the cells of all rows are contiguous, so the 3^d neighbour cells
//     of a tile cell are ONE contiguous range of shared memory),
//   * splits that range between kWPC (2 or 4) warps per cell in an interleaved fashion (warp p takes the
//     32-candidate blocks b = p, p + kWPC, ...), so every warp sees a uniform sample of all rows,
//   * tests up to 8 blocks (256 candidates) into 8 hit masks per lane before draining them,
//   * adds the kWPC partial accumulators of a point through shared memory at the end.
// The visiting order is no longer the reference's, which only matters for the bit-identical
// "exact" mode; that mode (and any tile whose candidates exceed the staging capacity) runs the
// ordered row-by-row kernel of sweep.cuh instead.
#pragma once

#include "sweep.cuh"

namespace pnb {

constexpr int kFTX = 4;                    // cells per tile
constexpr int kFSlots = kFTX + 2;
constexpr int kFCap = 1920;                // staged candidates per tile (typical: 6*9*27 = 1458)
constexpr int kFCapPad = kFCap + 32;

__host__ __device__ constexpr int rows_of(int nd) { return nd == 3 ? 9 : (nd == 2 ? 3 : 1); }

// kWPC (warps per cell) is a property of the closure: cheap closures run best with 2 (less
// merge/synchronisation overhead, better balance: measured count 7.8 -> 6.6 ms, n-body
// 14.5 -> 12.9 ms), the latency-heavy WCSPH interaction needs the occupancy of 4
// (18.9 ms vs 22.8 ms with 2).
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(kFTX * CL::kWarpsPerCell * 32, 1024 / (kFTX * CL::kWarpsPerCell * 32))
k_sweep_tiles(GridP g, const uint32_t *__restrict__ cell_start, const float4 *__restrict__ sorted,
              CL cl, int *__restrict__ overflow_tiles, int *__restrict__ overflow_count)
{
    constexpr int NR = rows_of(ND);
    constexpr int NE = kFSlots * NR;          // staged cells per tile
    constexpr int kWPC = CL::kWarpsPerCell;
    constexpr int kFThreads = kFTX * kWPC * 32;
    // hit-mask words per lane: one drain per batch covers a whole part (729 / kWPC candidates)
    constexpr int kFMasks = kWPC == 2 ? 12 : 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_pos = reinterpret_cast<float4 *>(smem_raw);
    unsigned char *s_pay = smem_raw + sizeof(float4) * kFCapPad;
    __shared__ uint32_t s_cbeg[NE];
    __shared__ uint32_t s_cpre[NE + 1];
    __shared__ int s_maxpass[kFTX];
    __shared__ unsigned s_mask[kFMasks][kFThreads];   // hit masks of the current super-block

    const int nx = g.gs[0] - 2;
    const int ny = ND > 1 ? g.gs[1] - 2 : 1;
    const int ntx = (nx + kFTX - 1) / kFTX;
    int64_t b = blockIdx.x;
    const int tx = (int)(b % ntx); b /= ntx;
    const int iy = (int)(b % ny);  b /= ny;
    const int iz = (int)b;
    const int cx0 = 2 + tx * kFTX;
    const int cx1 = min(cx0 + kFTX - 1, g.gs[0] - 1);
    const int cy = ND > 1 ? 2 + iy : 1;
    const int cz = ND > 2 ? 2 + iz : 1;
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const PerP pp = make_perp(g);

    const uint32_t tile_p0 = cell_start[linear_cell(g, cx0, cy, cz)];
    const uint32_t tile_p1 = cell_start[linear_cell(g, cx1, cy, cz) + 1];
    if (tile_p0 == tile_p1) return;

    // ---- table of staged cells, entry e = slot * NR + row (rows in CartesianIndices order) ----
    if (warp == 0) {
        uint32_t cnt[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int e = (int)threadIdx.x * 2 + h;      // thread t owns entries 2t, 2t+1
            if (e < NE) {
                const int slot = e / NR, row = e % NR;
                int sx = cx0 - 1 + slot;
                int ry = cy + (ND > 1 ? (row % 3) - 1 : 0);
                int rz = cz + (ND > 2 ? (row / 3) - 1 : 0);
                uint32_t b0 = 0;
                if (sx <= cx1 + 1) {
                    if (PER) {
                        sx = floormod_i(sx - 2, g.nc[0]) + 2;
                        if (ND > 1) ry = floormod_i(ry - 2, g.nc[1]) + 2;
                        if (ND > 2) rz = floormod_i(rz - 2, g.nc[2]) + 2;
                    }
                    const int lin = linear_cell(g, sx, ry, rz);
                    b0 = cell_start[lin];
                    cnt[h] = cell_start[lin + 1] - b0;
                }
                s_cbeg[e] = b0;
            }
        }
        // exclusive prefix over the (<= 64) entries by the first warp
        uint32_t pair_sum = cnt[0] + cnt[1];
        uint32_t incl = pair_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        static_assert(NE <= 64, "cell table is scanned by one warp");
        const uint32_t excl = incl - pair_sum;
        const int e0 = lane * 2;
        if (e0 < NE) s_cpre[e0] = excl;
        if (e0 + 1 < NE) s_cpre[e0 + 1] = excl + cnt[0];
        if (e0 + 2 == NE || e0 + 1 == NE) s_cpre[NE] = excl + cnt[0] + (e0 + 1 < NE ? cnt[1] : 0u);
    }
    // passes per cell (cells with more than 32 points are swept in several batches)
    const int my_cell = warp / kWPC, part = warp % kWPC;
    const int my_cx = cx0 + my_cell;
    uint32_t c_p0 = 0, c_p1 = 0;
    if (my_cx <= cx1) {
        const int lin = linear_cell(g, my_cx, cy, cz);
        c_p0 = cell_start[lin];
        c_p1 = cell_start[lin + 1];
    }
    if (part == 0 && lane == 0) s_maxpass[my_cell] = (int)((c_p1 - c_p0 + 31) / 32);
    __syncthreads();
    const uint32_t total = s_cpre[NE];
    if (total > (uint32_t)kFCap) {
        // too dense for the staging buffer: hand the tile to the ordered kernel
        if (threadIdx.x == 0) overflow_tiles[atomicAdd(overflow_count, 1)] = (int)blockIdx.x;
        return;
    }
    int n_batches = 0;
#pragma unroll
    for (int w = 0; w < kFTX; w++) n_batches = max(n_batches, s_maxpass[w]);

    // ---- stage every cell of the table: positions + closure payload ---------------------------
    for (int e = warp; e < NE; e += kFTX * kWPC) {
        const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_cpre[e + 1] - d0;
        for (uint32_t k = lane; k < n; k += 32) {
            s_pos[d0 + k] = sorted[b0 + k];
            cl.stage(s_pay, (int)(d0 + k), b0 + k, kFCap);
        }
    }
    __syncthreads();

    // candidates of my cell: slots my_cell .. my_cell + 2, all rows
    const uint32_t R0 = s_cpre[my_cell * NR], R1 = s_cpre[(my_cell + 3) * NR];
    const int nblk = (int)((R1 - R0 + 31u) / 32u);

    for (int batch = 0; batch < n_batches; batch++) {
        const uint32_t i_sorted = c_p0 + (uint32_t)batch * 32u + (uint32_t)lane;
        const bool active = i_sorted < c_p1;
        float xi = 0.f, yi = 0.f, zi = 0.f;
        int i_id = 0;
        typename CL::State st;
        if (active) {
            const float4 pi = sorted[i_sorted];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
        }
        cl.init(st, active, (int)i_sorted, i_id);
        if (__any_sync(0xffffffffu, active)) {
            for (int b0 = part; b0 < nblk; b0 += kWPC * kFMasks) {
                // ---- test: up to kFMasks blocks of 32 candidates, hit masks parked in shared
                //      memory (each thread only ever touches its own words: no barrier needed)
                int rem = 0;
#pragma unroll 1
                for (int u = 0; u < kFMasks; u++) {
                    const int bb = b0 + u * kWPC;
                    unsigned hh = 0u;
                    if (bb < nblk) {   // warp-uniform
                        const uint32_t blk = R0 + 32u * (uint32_t)bb;
                        hh = test_block<ND, PER>(pp, s_pos + blk, xi, yi, zi);
                        const uint32_t nv = R1 - blk;
                        if (nv < 32u) hh &= (1u << nv) - 1u;
                        if (!active) hh = 0u;
                    }
                    s_mask[u][threadIdx.x] = hh;
                    rem += __popc(hh);
                }
                if (CL::kCountOnly) {
                    cl.count(st, rem);
                } else {
                    // ---- drain: every lane walks its own masks with a cursor, one hit per round
                    int u = -1;
                    unsigned mm = 0u;
                    while (__any_sync(0xffffffffu, rem > 0)) {
                        if (rem > 0) {
                            while (mm == 0u) mm = s_mask[++u][threadIdx.x];
                            const int k = __ffs(mm) - 1;
                            mm &= mm - 1u;
                            rem--;
                            const int slot = (int)R0 + 32 * (b0 + u * kWPC) + k;
                            const float4 pj = s_pos[slot];
                            float px = __fsub_rn(xi, pj.x);
                            float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                            float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                            float d2 = dist2<ND>(px, py, pz);
                            d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                            cl.template pair<ND>(st, px, py, pz, d2, __float_as_int(pj.w), s_pay,
                                                 slot, kFCap);
                        }
                    }
                }
            }
        }
        // ---- add the kWPC partial states of every point: parts 1.. publish, part 0 merges ------
        // (the payload region is free once every warp has finished its drain)
        __syncthreads();
        typename CL::State *s_red = reinterpret_cast<typename CL::State *>(s_pay);
        if (part > 0) s_red[((my_cell * (kWPC - 1)) + (part - 1)) * 32 + lane] = st;
        __syncthreads();
        if (part == 0 && active) {
#pragma unroll
            for (int q = 0; q < kWPC - 1; q++)
                cl.merge(st, s_red[((my_cell * (kWPC - 1)) + q) * 32 + lane]);
            cl.finish(st, (int)i_sorted, i_id);
        }
        if (batch + 1 < n_batches) {
            // the payload was overwritten by the partial states: stage it again for the next batch
            __syncthreads();
            for (int e = warp; e < NE; e += kFTX * kWPC) {
                const uint32_t bb0 = s_cbeg[e], d0 = s_cpre[e], n = s_cpre[e + 1] - d0;
                for (uint32_t k = lane; k < n; k += 32) cl.stage(s_pay, (int)(d0 + k), bb0 + k, kFCap);
            }
            __syncthreads();
        }
    }
}

// Tiles that did not fit the staging buffer of k_sweep_tiles: ordered row-by-row sweep with the
// same tile shape (kFTX cells, one warp per cell), persistent over the overflow list.
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(kFTX * 32)
k_sweep_overflow(GridP g, const uint32_t *__restrict__ cell_start,
                 const float4 *__restrict__ sorted, CL cl, const int *__restrict__ overflow_tiles,
                 const int *__restrict__ overflow_count)
{
    const int n = *overflow_count;
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        sweep_tile_rows<ND, PER, CL, kFTX>(g, cell_start, sorted, cl, (int64_t)overflow_tiles[t]);
        __syncthreads();
    }
}

}  // namespace pnb
