// sweep_tiles.cuh -- the throughput variant of the x === y sweep (default, "fast" arithmetic mode).
//
// What ncu said about the previous version (profiles/r1_wcsph_sweep_v3_ncu_summary.txt): the
// kernel is issue bound (72 % issue-active, 15.7 G warp instructions for 16.4 M points); 35 % of
// the instructions are the FP32 distance test (12 per candidate and warp) and 53 % the deferred
// interaction loop, which ran with 17 of 32 lanes because every warp drained only the hits of
// its own interleaved quarter of the candidates.  This version removes both:
//
//   * TEST in packed half precision (non-periodic grids).  Every staged candidate also gets a
//     (x, y, z) copy relative to the tile centre in units of the search radius, rounded to
//     fp16 and packed two candidates per register (candidate k and k + 16 of a 32-block).  One
//     HADD2 x3 + HMUL2 + HFMA2 x2 + HSET2 + LOP3 tests two candidates: 4 instructions per
//     candidate instead of 12.  The fp16 test is a conservative PRE-FILTER, never the decision:
//     all rounding errors together are below 0.007 r^2 (derivation in DESIGN.md 5.2), the
//     threshold is 1.0098 r^2, so no true neighbour can be missed, and every hit is re-tested
//     in the drain loop with the reference's exact Float32 operation sequence (the drain
//     computes pos_diff and d2 exactly anyway) -- the delivered neighbour SET is bit-identical.
//     About 3 % of the pre-filter hits are rejected there.  Periodic grids keep the exact
//     Float32 test (the wrap-around needs it).
//   * BALANCED DRAIN.  The kWPC warps of a cell first publish their hit masks (one word per
//     lane and 32-block) in shared memory; after a cell-local named barrier every lane knows
//     the total hit count H of its point and takes the hits of rank [part * H / kWPC,
//     (part + 1) * H / kWPC): the warps of a cell run the same number of rounds (+-1) and the
//     per-point totals vary by only a few percent, so ~26 of the 27 occupied lanes stay busy.
//
// Layout of the staged candidates: slot-major (for every x-column of the tile's (TX+2) columns
// the cells of all 3^(d-1) rows are contiguous), every slot padded to a multiple of 32, so
// the 3^d neighbour cells of a tile cell are ONE run of whole 32-blocks.
// The visiting order is not the reference's, which only matters for the bit-identical "exact"
// mode; that mode (and any tile whose candidates exceed the staging capacity) runs the ordered
// row-by-row kernel of sweep.cuh instead.
#pragma once

#include <cuda_fp16.h>

#include "sweep.cuh"

namespace pnb {

constexpr int kFTX = 4;                    // cells per tile
constexpr int kFSlots = kFTX + 2;
// Staging capacity of a tile (candidates incl. padding; typical 6 * 256) and 32-blocks per cell
// (3 slots): 1728 / 28 keeps the shared memory small enough for 2 (WCSPH) to 4 (count) CTAs per
// SM.  The neighbour-list passes declare kBigTiles: measured on the 8 M periodic cloud (27.8
// points per cell), 5 % of the tiles exceeded 1728 and cost 1.4 ms in the overflow kernel,
// more than the larger tiles cost in occupancy (count 5.2 -> 5.9 ms, n-body 10.8 -> 14.7 ms
// at 16.4 M points when every closure got them).
template <class CL, class = void>
struct wants_big_tiles { static constexpr bool value = false; };
template <class CL>
struct wants_big_tiles<CL, decltype((void)CL::kBigTiles)> { static constexpr bool value = CL::kBigTiles; };
template <class CL> __host__ __device__ constexpr int tile_cap() { return wants_big_tiles<CL>::value ? 2304 : 1728; }
template <class CL> __host__ __device__ constexpr int tile_nblk_max() { return wants_big_tiles<CL>::value ? 36 : 28; }
constexpr float kHalfSentinel = 64.0f;     // padding candidates: farther than any real one
constexpr int kLeftMax = 8;                // surplus points of a cell handed to k_sweep_left
// pre-filter thresholds in units of r^2, exactly representable in fp16.  Total rounding error of
// the fp16 distance (derivation in DESIGN.md 5.2): < 0.0069 on non-periodic grids (|u_x| <= 3,
// |u_yz| <= 1.5), < 0.011 on periodic grids (cell_size / r < 4/3: |u_x| < 4, |u_yz| < 2).
__host__ __device__ constexpr float half_thr_hi(bool per) { return per ? 1.013671875f : 1.009765625f; }
__host__ __device__ constexpr float half_thr_lo(bool per) { return per ? 0.986328125f : 0.990234375f; }

__host__ __device__ constexpr int rows_of(int nd) { return nd == 3 ? 9 : (nd == 2 ? 3 : 1); }

// dynamic shared memory of k_sweep_tiles
// closures that need the masks of the test phase to be the exact neighbour set (neighbour-list
// fill: the write position of a hit is its rank) declare kExactMasks = true
template <class CL, class = void>
struct needs_exact_masks { static constexpr bool value = false; };
template <class CL>
struct needs_exact_masks<CL, decltype((void)CL::kExactMasks)> { static constexpr bool value = CL::kExactMasks; };

template <int ND, class CL, bool HALF>
__host__ __device__ constexpr size_t tiles_smem_bytes()
{
    // mask planes: 1 = hit masks of the drain; 2 = certain hits + undecided band (exact modes)
    constexpr int planes = CL::kCountOnly ? (HALF ? 1 : 0)            // only the undecided band
                           : ((HALF && needs_exact_masks<CL>::value) ? 2 : 1);
    constexpr int kFCap = tile_cap<CL>(), kFBlocks = kFCap / 32, kFNBlkMax = tile_nblk_max<CL>();
    return sizeof(float4) * kFCap                               // exact positions + id
           + (HALF ? (size_t)kFBlocks * ND * 16 * 4 : 0)        // packed fp16 coordinates
           + (size_t)kFCap * CL::kPayBytes                      // closure payload planes
           + (size_t)planes * kFTX * kFNBlkMax * 32 * 4;
}

__device__ __forceinline__ float4 lds128(uint32_t sa)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa));
    return v;
}
__device__ __forceinline__ float lds32(uint32_t sa)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(sa));
    return v;
}

__device__ __forceinline__ void cell_barrier(int cell, int nthreads)
{
    // immediate barrier ids so that ptxas reserves kFTX + 1 barriers, not all 16
    switch (cell) {
        case 0: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
        default: asm volatile("bar.sync 4, %0;" ::"r"(nthreads) : "memory"); break;
    }
    static_assert(kFTX == 4, "one named barrier per tile cell");
}

// One 32-block in packed fp16: bit k of the result <=> candidate k passes the pre-filter
// (d2 <= thr_hi).  TWO: *sure gets the candidates with d2 <= thr_lo, which are neighbours
// whatever the rounding did (count-only closures re-test only the band in between).
template <int ND, bool TWO>
__device__ __forceinline__ unsigned test_block_half(const uint32_t *__restrict__ hb, __half2 xi,
                                                    __half2 yi, __half2 zi, __half2 thr_hi,
                                                    __half2 thr_lo, unsigned *sure)
{
    unsigned hits = 0u, in = 0u;
    const uint4 *hp = reinterpret_cast<const uint4 *>(hb);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 X = hp[q];
        uint4 Y = make_uint4(0u, 0u, 0u, 0u), Z = make_uint4(0u, 0u, 0u, 0u);
        if (ND > 1) Y = hp[4 + q];
        if (ND > 2) Z = hp[8 + q];
        const uint32_t xw[4] = {X.x, X.y, X.z, X.w};
        const uint32_t yw[4] = {Y.x, Y.y, Y.z, Y.w};
        const uint32_t zw[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const int k = q * 4 + w;
            const __half2 ex = __hsub2(xi, *reinterpret_cast<const __half2 *>(&xw[w]));
            __half2 d2 = __hmul2(ex, ex);
            if (ND > 1) {
                const __half2 ey = __hsub2(yi, *reinterpret_cast<const __half2 *>(&yw[w]));
                d2 = __hfma2(ey, ey, d2);
            }
            if (ND > 2) {
                const __half2 ez = __hsub2(zi, *reinterpret_cast<const __half2 *>(&zw[w]));
                d2 = __hfma2(ez, ez, d2);
            }
            const unsigned sel = (1u << k) | (1u << (k + 16));
            hits |= __hle2_mask(d2, thr_hi) & sel;             // 0xffff per passing half
            if (TWO) in |= __hle2_mask(d2, thr_lo) & sel;
        }
    }
    if (TWO) *sure = in;
    return hits;
}

// TWO: the query points are a second point set (x != y) whose cell-ordered copy is
// (q_start, q_sorted) (pnb::build_query_list); the candidates always come from the cell list
// (cell_start, sorted).  For x === y both pairs of pointers are the same arrays.
template <int ND, bool PER, class CL, int kWPC, bool HALF, bool TWO>
__global__ void __launch_bounds__(kFTX * kWPC * 32, 1024 / (kFTX * kWPC * 32))
k_sweep_tiles(GridP g, CellsView cand, CellsView qry, CL cl, int *__restrict__ overflow_tiles,
              int *__restrict__ overflow_count, int *__restrict__ left_ids)
{
    const float4 *__restrict__ sorted = cand.rec;
    const float4 *__restrict__ q_sorted = qry.rec;
    constexpr int NR = rows_of(ND);
    constexpr int NE = kFSlots * NR;          // staged cells per tile
    constexpr int kFThreads = kFTX * kWPC * 32;
    constexpr int kCellThreads_ = kWPC * 32;
    constexpr int kFCap = tile_cap<CL>(), kFBlocks = kFCap / 32, kFNBlkMax = tile_nblk_max<CL>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_pos = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *s_half = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float4) * kFCap);
    unsigned char *s_pay = smem_raw + sizeof(float4) * kFCap + (HALF ? (size_t)kFBlocks * ND * 64 : 0);
    unsigned *s_mask = reinterpret_cast<unsigned *>(s_pay + (size_t)kFCap * CL::kPayBytes);
    __shared__ uint32_t s_cbeg[NE];           // global begin of every staged cell
    __shared__ uint32_t s_ccnt[NE];           // its point count
    __shared__ uint32_t s_cpre[NE];           // its first staged (padded) index
    __shared__ uint32_t s_slot0[kFSlots + 1]; // first staged index of every slot (multiples of 32)
    __shared__ uint32_t s_spop[kFSlots];      // candidates in the slot (without padding)
    __shared__ int s_cnt[kFTX][kWPC][32];     // hits per lane and part
    __shared__ int s_maxpass[kFTX];

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

    {
        uint32_t n_tile = 0;   // query points of the tile (uniform for the CTA)
        for (int cx = cx0; cx <= cx1; cx++) {
            uint32_t b0, cnt;
            cell_range(qry, linear_cell(g, cx, cy, cz), b0, cnt);
            n_tile += cnt;
        }
        if (n_tile == 0) return;   // no query points in this tile
    }

    // ---- table of staged cells, entry e = slot * NR + row (rows in CartesianIndices order) ----
    if (warp == 0) {
        for (int e = lane; e < NE; e += 32) {
            const int slot = e / NR, row = e % NR;
            int sx = cx0 - 1 + slot;
            int ry = cy + (ND > 1 ? (row % 3) - 1 : 0);
            int rz = cz + (ND > 2 ? (row / 3) - 1 : 0);
            uint32_t b0 = 0, cnt = 0;
            if (sx <= cx1 + 1) {
                if (PER) {
                    sx = floormod_i(sx - 2, g.nc[0]) + 2;
                    if (ND > 1) ry = floormod_i(ry - 2, g.nc[1]) + 2;
                    if (ND > 2) rz = floormod_i(rz - 2, g.nc[2]) + 2;
                }
                cell_range(cand, linear_cell(g, sx, ry, rz), b0, cnt);
            }
            s_cbeg[e] = b0;
            s_ccnt[e] = cnt;
        }
        __syncwarp();
        if (lane < kFSlots) {
            uint32_t run = 0;
#pragma unroll
            for (int r = 0; r < NR; r++) { s_cpre[lane * NR + r] = run; run += s_ccnt[lane * NR + r]; }
            s_spop[lane] = run;
        }
        __syncwarp();
        if (lane == 0) {
            uint32_t run = 0;
#pragma unroll
            for (int sl = 0; sl < kFSlots; sl++) { s_slot0[sl] = run; run += (s_spop[sl] + 31u) & ~31u; }
            s_slot0[kFSlots] = run;
        }
        __syncwarp();
        for (int e = lane; e < NE; e += 32) s_cpre[e] += s_slot0[e / NR];
    }
    // passes per cell (cells with more than 32 points are swept in several batches)
    const int my_cell = warp / kWPC, part = warp % kWPC;
    const int my_cx = cx0 + my_cell;
    uint32_t c_p0 = 0, c_p1 = 0;
    if (my_cx <= cx1) {
        uint32_t cnt;
        cell_range(qry, linear_cell(g, my_cx, cy, cz), c_p0, cnt);
        c_p1 = c_p0 + cnt;
    }
    if (part == 0 && lane == 0) s_maxpass[my_cell] = (int)((c_p1 - c_p0 + 31) / 32);
    __syncthreads();
    // too dense for the staging buffer or the mask table: hand the tile to the ordered kernel
    bool too_big = s_slot0[kFSlots] > (uint32_t)kFCap;
#pragma unroll
    for (int w = 0; w < kFTX; w++)
        too_big = too_big || (s_slot0[w + 3] - s_slot0[w]) > (uint32_t)(kFNBlkMax * 32);
    if (too_big) {
        if (threadIdx.x == 0) overflow_tiles[atomicAdd(overflow_count, 1)] = (int)blockIdx.x;
        return;
    }
    // A cell with a few points more than a warp (33 .. 32 + kLeftMax; 4.8 % of the cells of the
    // benchmark cloud, 18 % of the tiles) would need a second batch that repeats the whole test
    // phase for a handful of active lanes: those points go to a list that k_sweep_left works
    // off (one thread per point), and the cell runs ONE batch.  overflow_count[1] = list length.
    if (left_ids != nullptr) {
        const uint32_t cnt = c_p1 - c_p0;
        if (cnt > 32u && cnt <= 32u + (uint32_t)kLeftMax) {
            if (part == 0) {
                const uint32_t nl = cnt - 32u;
                int base_l = 0;
                if (lane == 0) base_l = atomicAdd(overflow_count + 1, (int)nl);
                base_l = __shfl_sync(0xffffffffu, base_l, 0);
                if ((uint32_t)lane < nl)
                    left_ids[base_l + lane] = __float_as_int(__ldg(&q_sorted[c_p0 + 32u + (uint32_t)lane].w));
            }
            c_p1 = c_p0 + 32u;
        }
    }
    // batches of THIS cell (its warps synchronise among themselves only: named barriers)
    const int n_batches = (int)((c_p1 - c_p0 + 31u) / 32u);

    // ---- stage every cell of the table: exact positions, fp16 copies, closure payload ---------
    // tile centre (local cell c covers [minc + (c + off - 1) cs, minc + (c + off) cs))
    float org[3] = {0.f, 0.f, 0.f};
    float inv_r = 0.f;
    if (HALF) {
        org[0] = fmaf((float)(cx0 + g.off[0] + 1), g.cs[0], g.minc[0]);
        if (ND > 1) org[1] = fmaf((float)(cy + g.off[1]) - 0.5f, g.cs[1], g.minc[1]);
        if (ND > 2) org[2] = fmaf((float)(cz + g.off[2]) - 0.5f, g.cs[2], g.minc[2]);
        inv_r = __frcp_rn(g.r);
    }
    float inv_bs[3] = {0.f, 0.f, 0.f};
    if (HALF && PER) {
#pragma unroll
        for (int d = 0; d < 3; d++) inv_bs[d] = __frcp_rn(pp.bs[d]);
    }
    __half *s_half16 = reinterpret_cast<__half *>(s_half);
    for (int e = warp; e < NE; e += kFTX * kWPC) {
        const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_ccnt[e];
        // periodic grids: every candidate is staged at the image closest to the NOMINAL position
        // of its cell next to the tile (minimum image): its fp16 copy is shifted by whole
        // periods, k = rint((cell centre - y) / box size) per point -- points may lie outside
        // the box, the reference wraps cells, not coordinates.  Only the pre-filter sees the
        // shifted copy; the exact test works on the raw coordinates.
        float cen[3] = {0.f, 0.f, 0.f};
        if (HALF && PER) {
            const int slot = e / NR, row = e % NR;
            const int sx = cx0 - 1 + slot;
            const int ry = cy + (ND > 1 ? (row % 3) - 1 : 0);
            const int rz = cz + (ND > 2 ? (row / 3) - 1 : 0);
            cen[0] = fmaf((float)(sx + g.off[0]) - 0.5f, g.cs[0], g.minc[0]);
            if (ND > 1) cen[1] = fmaf((float)(ry + g.off[1]) - 0.5f, g.cs[1], g.minc[1]);
            if (ND > 2) cen[2] = fmaf((float)(rz + g.off[2]) - 0.5f, g.cs[2], g.minc[2]);
        }
        for (uint32_t k = lane; k < n; k += 32) {
            const float4 pj = sorted[b0 + k];
            const uint32_t q = d0 + k;
            s_pos[q] = pj;
            if (HALF) {
                const uint32_t hb = (q >> 5) * (ND * 32) + (q & 15u) * 2u + ((q >> 4) & 1u);
                float ux = pj.x - org[0], uy = pj.y - org[1], uz = pj.z - org[2];
                if (PER) {
                    ux = fmaf(rintf((cen[0] - pj.x) * inv_bs[0]), pp.bs[0], ux);
                    if (ND > 1) uy = fmaf(rintf((cen[1] - pj.y) * inv_bs[1]), pp.bs[1], uy);
                    if (ND > 2) uz = fmaf(rintf((cen[2] - pj.z) * inv_bs[2]), pp.bs[2], uz);
                }
                s_half16[hb] = __float2half_rn(ux * inv_r);
                if (ND > 1) s_half16[hb + 32] = __float2half_rn(uy * inv_r);
                if (ND > 2) s_half16[hb + 64] = __float2half_rn(uz * inv_r);
            }
            cl.stage(s_pay, (int)q, b0 + k, kFCap);
        }
    }
    // padding between a slot's last candidate and the next multiple of 32
    for (int sl = warp; sl < kFSlots; sl += kFTX * kWPC) {
        const uint32_t q = s_slot0[sl] + s_spop[sl] + (uint32_t)lane;
        if (q < s_slot0[sl + 1]) {
            if (HALF) {
                const uint32_t hb = (q >> 5) * (ND * 32) + (q & 15u) * 2u + ((q >> 4) & 1u);
                const __half far = __float2half_rn(kHalfSentinel);
                s_half16[hb] = far;
                if (ND > 1) s_half16[hb + 32] = far;
                if (ND > 2) s_half16[hb + 64] = far;
            }
            s_pos[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();

    // candidates of my cell: slots my_cell .. my_cell + 2 = whole blocks B0 .. B0 + nblk - 1
    const int B0 = (int)(s_slot0[my_cell] >> 5);
    const int nblk = (int)((s_slot0[my_cell + 3] - s_slot0[my_cell]) >> 5);
    unsigned *my_mask = s_mask + (size_t)my_cell * kFNBlkMax * 32 + lane;   // word of block b: my_mask[b * 32]
    // 32-bit shared-window addresses, computed once (the compiler otherwise rebuilds the window
    // base of the dynamic shared memory in every round of the drain loop)
    uint32_t pos_sa = (uint32_t)__cvta_generic_to_shared(s_pos);
    asm volatile("mov.u32 %0, %0;" : "+r"(pos_sa));    // opaque: keep it in a register, do not rematerialise
    const uint32_t pay_sa = pos_sa + (uint32_t)(sizeof(float4) * kFCap + (HALF ? (size_t)kFBlocks * ND * 64 : 0));
    const int nb4 = (nblk + kWPC - 1) / kWPC;                // blocks per share
    const int blk_lo = min(part * nb4, nblk), blk_hi = min(blk_lo + nb4, nblk);
    const uint32_t q_self0 = s_cpre[(my_cell + 1) * NR + NR / 2];
    const __half2 thr = __float2half2_rn(half_thr_hi(PER));
    const __half2 thr_lo = __float2half2_rn(half_thr_lo(PER));

    for (int batch = 0; batch < n_batches; batch++) {
        const uint32_t i_sorted = c_p0 + (uint32_t)batch * 32u + (uint32_t)lane;
        const bool active = i_sorted < c_p1;
        float xi = 0.f, yi = 0.f, zi = 0.f;
        int i_id = 0;
        __half2 hx = __float2half2_rn(0.f), hy = hx, hz = hx;
        typename CL::State st;
        if (active && !TWO) {
            // x === y: the point is one of the staged candidates of its own cell
            const uint32_t q = q_self0 + (uint32_t)batch * 32u + (uint32_t)lane;
            const float4 pi = s_pos[q];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
            if (HALF) {
                const uint32_t hb = (q >> 5) * (ND * 32) + (q & 15u) * 2u + ((q >> 4) & 1u);
                hx = __half2half2(s_half16[hb]);
                if (ND > 1) hy = __half2half2(s_half16[hb + 32]);
                if (ND > 2) hz = __half2half2(s_half16[hb + 64]);
            }
        }
        if (active && TWO) {
            const float4 pi = q_sorted[i_sorted];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
            if (HALF) {
                float ux = xi - org[0], uy = yi - org[1], uz = zi - org[2];
                if (PER) {
                    const float c0 = fmaf((float)(my_cx + g.off[0]) - 0.5f, g.cs[0], g.minc[0]);
                    ux = fmaf(rintf((c0 - xi) * inv_bs[0]), pp.bs[0], ux);
                    if (ND > 1) {
                        const float c1 = fmaf((float)(cy + g.off[1]) - 0.5f, g.cs[1], g.minc[1]);
                        uy = fmaf(rintf((c1 - yi) * inv_bs[1]), pp.bs[1], uy);
                    }
                    if (ND > 2) {
                        const float c2 = fmaf((float)(cz + g.off[2]) - 0.5f, g.cs[2], g.minc[2]);
                        uz = fmaf(rintf((c2 - zi) * inv_bs[2]), pp.bs[2], uz);
                    }
                }
                hx = __float2half2_rn(ux * inv_r);
                if (ND > 1) hy = __float2half2_rn(uy * inv_r);
                if (ND > 2) hz = __float2half2_rn(uz * inv_r);
            }
        }
        cl.init(st, active, TWO ? -1 : (int)i_sorted, i_id);

        // ---- phase 1: test my contiguous share of the cell's blocks; masks to shared memory --
        // kExact: the closure needs EXACT masks / counts (count only, list fill): certain hits
        // (below the lower threshold) go to the hit masks, the undecided band between the two
        // thresholds goes to a second mask plane and is decided below by the exact test.
        constexpr bool kExact = HALF && (CL::kCountOnly || needs_exact_masks<CL>::value);
        unsigned *my_band = CL::kCountOnly ? my_mask : my_mask + (size_t)kFTX * kFNBlkMax * 32;
        int cnt = 0, n_maybe = 0;
        for (int bb = blk_lo; bb < blk_hi; bb++) {
            unsigned hh;
            if (kExact) {
                unsigned sure;
                hh = test_block_half<ND, true>(s_half + (size_t)(B0 + bb) * (ND * 16), hx, hy, hz,
                                               thr, thr_lo, &sure);
                if (!active) { hh = 0u; sure = 0u; }
                my_band[bb * 32] = hh & ~sure;
                n_maybe += __popc(hh & ~sure);
                hh = sure;
            } else if (HALF) {
                hh = test_block_half<ND, false>(s_half + (size_t)(B0 + bb) * (ND * 16), hx, hy, hz,
                                                thr, thr_lo, nullptr);
            } else {
                hh = test_block<ND, PER>(pp, s_pos + 32 * (B0 + bb), xi, yi, zi);
                // padding slots are not candidates
                const int q0 = 32 * (B0 + bb);
                int sl = my_cell;
                if (q0 >= (int)s_slot0[my_cell + 1]) sl = my_cell + 1;
                if (q0 >= (int)s_slot0[my_cell + 2]) sl = my_cell + 2;
                const int nv = (int)(s_slot0[sl] + s_spop[sl]) - q0;
                if (nv < 32) hh &= (1u << max(nv, 0)) - 1u;
            }
            if (!active) hh = 0u;
            if (!CL::kCountOnly) my_mask[bb * 32] = hh;
            cnt += __popc(hh);
        }
        if (kExact) {
            // exact test (the reference's operation sequence, periodic fix included) of the
            // undecided candidates of my own blocks; accepted ones join the hit masks
            const int rounds = __reduce_max_sync(0xffffffffu, n_maybe);
            int bb = blk_lo - 1;
            unsigned mm = 0u;
            for (int t = 0; t < rounds; t++) {
                if (t < n_maybe) {
                    while (mm == 0u) { bb++; mm = my_band[bb * 32]; }
                    const int k = __ffs(mm) - 1;
                    mm &= mm - 1u;
                    const float4 pj = s_pos[32 * (B0 + bb) + k];
                    float px = __fsub_rn(xi, pj.x);
                    float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    if (d2 <= pp.r2) {
                        cnt++;
                        if (!CL::kCountOnly) my_mask[bb * 32] |= 1u << k;
                    }
                }
            }
        }
        if (CL::kCountOnly) {
            // the parts' counts are added through s_cnt
            s_cnt[my_cell][part][lane] = cnt;
            cell_barrier(my_cell, kCellThreads_);
            if (part == 0) {
                int tot = 0;
#pragma unroll
                for (int p = 0; p < kWPC; p++) tot += s_cnt[my_cell][p][lane];
                cl.count(st, tot);
            }
            if (batch + 1 < n_batches) cell_barrier(my_cell, kCellThreads_);
        } else {
            s_cnt[my_cell][part][lane] = cnt;
            cell_barrier(my_cell, kCellThreads_);

            // ---- phase 2: my share of the point's hits: ranks [part * Q, part * Q + n_mine) ----
            int pc[kWPC];
            int H = 0;
#pragma unroll
            for (int p = 0; p < kWPC; p++) { pc[p] = s_cnt[my_cell][p][lane]; H += pc[p]; }
            cl.total(st, H, part == 0 && active);
            const int Q = (H + kWPC - 1) / kWPC;
            int skip = part * Q;
            const int n_mine = max(0, min(Q, H - skip));
            int bb = 0;
            unsigned mm = 0u;
            if (n_mine > 0) {
                // the share that holds my first hit (hits of share p live in blocks p * nb4 ...),
                // then the word inside it
#pragma unroll
                for (int p = 0; p < kWPC - 1; p++)
                    if (bb == p * nb4 && skip >= pc[p]) { skip -= pc[p]; bb = (p + 1) * nb4; }
                mm = my_mask[bb * 32];
                int c = __popc(mm);
                while (skip >= c) { skip -= c; bb++; mm = my_mask[bb * 32]; c = __popc(mm); }
                for (; skip > 0; skip--) mm &= mm - 1u;
            }
            const int rounds = __reduce_max_sync(0xffffffffu, n_mine);
            cl.seek(st, part * Q);

            // ---- phase 3: drain, one hit per lane and round ------------------------------------
            for (int t = 0; t < rounds; t++) {
                if (t < n_mine) {
                    while (mm == 0u) { bb++; mm = my_mask[bb * 32]; }
                    const int k = __ffs(mm) - 1;
                    mm &= mm - 1u;
                    const int slot = 32 * (B0 + bb) + k;
                    const float4 pj = lds128(pos_sa + 16u * (uint32_t)slot);
                    float px = __fsub_rn(xi, pj.x);
                    float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    // the decision: the reference's exact test (the fp16 pass only pre-selects)
                    if (!HALF || kExact || d2 <= pp.r2)
                        cl.template pair_s<ND>(st, px, py, pz, d2, __float_as_int(pj.w), pay_sa, slot, kFCap);
                }
            }
            cl.flush(st);
            // ---- phase 4: add the kWPC partial accumulators of every point ---------------------
            // (the cell's mask words are free once all its warps have finished their drain)
            cell_barrier(my_cell, kCellThreads_);
            float *s_red = reinterpret_cast<float *>(s_mask + (size_t)my_cell * kFNBlkMax * 32);
            static_assert((kWPC - 1) * CL::kAccWords <= kFNBlkMax, "partials fit the mask words");
            if (part > 0) cl.save_acc(st, s_red + ((part - 1) * CL::kAccWords) * 32 + lane);
            cell_barrier(my_cell, kCellThreads_);
            if (part == 0) {
#pragma unroll
                for (int p = 0; p < kWPC - 1; p++) cl.add_acc(st, s_red + (p * CL::kAccWords) * 32 + lane);
            }
            if (batch + 1 < n_batches) cell_barrier(my_cell, kCellThreads_);
        }
        if (part == 0 && active) cl.finish(st, TWO ? -1 : (int)i_sorted, i_id);
    }
    (void)kFThreads;
}

// The surplus points of cells with 33 .. 32 + kLeftMax points (see k_sweep_tiles).  G = 8 lanes
// share a point: lane s takes candidates s, s + 8, ... of every neighbour cell (one coalesced
// 128-byte read per round), the reference's exact test, the closure's global-memory pair
// function; the partial results are merged through the closure's own save_acc / add_acc (or
// summed, for count-only closures).  Closures whose result depends on the rank of a hit
// (neighbour-list fill) keep one thread per point.
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(128)
k_sweep_left(GridP g, CellsView cand, const float *__restrict__ x,
             const int *__restrict__ left_ids, const int *__restrict__ left_count, CL cl)
{
    constexpr int G = needs_exact_masks<CL>::value ? 1 : 8;
    constexpr int kWords = CL::kAccWords > 0 ? CL::kAccWords : 1;
    __shared__ float s_red[4][kWords][32];
    const int n = *left_count;
    const PerP pp = make_perp(g);
    const int lane = lane_id(), sub = lane % G;
    const int groups_total = (int)(gridDim.x * blockDim.x) / G;
    const int gidx = (int)(blockIdx.x * blockDim.x + threadIdx.x) / G;
    // the groups of a warp hold consecutive t: the warp leaves the loop together
    for (int t = gidx; t - lane / G < n; t += groups_total) {
        const bool have = t < n;
        const int i_id = have ? left_ids[t] : left_ids[0];
        float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __ldg(x + (int64_t)i_id * ND + d);
        int cc[3];
#pragma unroll
        for (int d = 0; d < ND; d++) cc[d] = cell_coord(p[d], g.minc[d], g.cs[d], g.periodic, g.nc[d], g.off[d]);
#pragma unroll
        for (int d = ND; d < 3; d++) cc[d] = 1;
        typename CL::State st;
        cl.init(st, have, -1, i_id);
        int hits = 0;
        // the point sits in a valid cell (it came out of the cell list), so its stencil is inside
        // the padded grid.  All cell ranges first, then the candidates with several loads in flight.
        constexpr int NC = ND == 3 ? 27 : (ND == 2 ? 9 : 3);
        uint32_t cb[NC], cn[NC];
#pragma unroll
        for (int e = 0; e < NC; e++) {
            int c0 = cc[0] + (e % 3) - 1;
            int c1 = cc[1] + (ND > 1 ? (e / 3) % 3 - 1 : 0);
            int c2 = cc[2] + (ND > 2 ? e / 9 - 1 : 0);
            if (PER) {
                c0 = floormod_i(c0 - 2, g.nc[0]) + 2;
                if (ND > 1) c1 = floormod_i(c1 - 2, g.nc[1]) + 2;
                if (ND > 2) c2 = floormod_i(c2 - 2, g.nc[2]) + 2;
            }
            cell_range(cand, linear_cell(g, c0, c1, c2), cb[e], cn[e]);
            if (!have) cn[e] = 0u;
        }
        constexpr int U = G == 1 ? 8 : 4;          // loads in flight per lane
#pragma unroll 1
        for (int e = 0; e < NC; e++) {
            const uint32_t b0 = cb[e], cnt = cn[e];
            for (uint32_t k = (uint32_t)sub; k < cnt; k += (uint32_t)(G * U)) {
                float4 pj[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    pj[u] = __ldg(cand.rec + b0 + min(k + (uint32_t)(u * G), cnt - 1u));
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (k + (uint32_t)(u * G) >= cnt) break;
                    const uint32_t gi = b0 + k + (uint32_t)(u * G);
                    float px = __fsub_rn(p[0], pj[u].x);
                    float py = ND > 1 ? __fsub_rn(p[1], pj[u].y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(p[2], pj[u].z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    if (d2 <= pp.r2) {
                        if (CL::kCountOnly) hits++;
                        else cl.template pair_global<ND>(st, px, py, pz, d2, __float_as_int(pj[u].w), gi);
                    }
                }
            }
        }
        if (G > 1) {
            if (CL::kCountOnly) {
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
            } else {
                float *mine = &s_red[threadIdx.x >> 5][0][lane];
                __syncwarp();
                if (sub != 0) cl.save_acc(st, mine);
                __syncwarp();
                if (sub == 0) {
#pragma unroll
                    for (int u = 1; u < G; u++) cl.add_acc(st, mine + u);
                }
            }
        }
        if (CL::kCountOnly) cl.count(st, hits);
        if (have && sub == 0) cl.finish(st, -1, i_id);
    }
}

// Tiles that did not fit the staging buffer of k_sweep_tiles: ordered row-by-row sweep with the
// same tile shape (kFTX cells, one warp per cell), persistent over the overflow list.
template <int ND, bool PER, class CL>
__global__ void __launch_bounds__(kFTX * 32)
k_sweep_overflow(GridP g, CellsView cand, CellsView qry, CL cl,
                 const int *__restrict__ overflow_tiles, const int *__restrict__ overflow_count)
{
    const int n = *overflow_count;
    for (int t = blockIdx.x; t < n; t += gridDim.x) {
        sweep_tile_rows<ND, PER, CL, kFTX>(g, cand, qry, cl, (int64_t)overflow_tiles[t]);
        __syncthreads();
    }
}

}  // namespace pnb
