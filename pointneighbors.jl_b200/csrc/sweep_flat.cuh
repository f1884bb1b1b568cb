// sweep_flat.cuh -- the x === y / two-set tile sweep with FLAT point -> lane packing (round 2).
//
// What ncu said about k_sweep_tiles (profiles/r1_wcsph_sweep_v6_ncu_summary.txt): 9.94 G warp
// instructions at 70 % issue; 25.3 of 32 lanes active because a warp owned ONE cell (27 points on
// the benchmark cloud); cells with 33..40 points needed a second batch or a second kernel
// (k_sweep_left, 0.31 ms); the drain walked its hit masks with a divergent `while (mm == 0)`
// (5.5 % of the instructions at 2 active lanes).  This kernel removes the three of them:
//
//   * TILES ARE POINT RANGES, NOT CELL RANGES.  A tile = up to kFTP = 96 consecutive points of
//     one x-row of the cell-ordered point list (any cell boundaries), cut by a small pre-pass
//     (k_flat_tiles) at multiples of 96 of a weighted point count in which every cell weighs at
//     least 24, so that a tile spans at most 5 cells (7 staged x-columns).  The 3 lane groups of
//     a tile are FULL warps (32 of 32 lanes except in the last tile of a row segment); a lane
//     whose cell differs from its neighbour's simply walks another run of candidate blocks
//     (per-lane block base), so cells with any number of points need neither batches nor a
//     surplus kernel.
//   * PERSISTENT CTAs (2 per SM for WCSPH) fetch tiles from an atomic counter; the next tile's
//     index is requested while the current tile is processed.
//   * The drain finds its next non-empty mask word with a per-lane bitmap of non-empty blocks
//     (one FLO + LOP3, predicated) instead of a divergent loop over the words.
//
// Everything that decides a neighbour is unchanged: packed-fp16 PRE-FILTER with the conservative
// thresholds of sweep_tiles.cuh (the staged span grows from 6 to at most 7 cells, |u_x| <= 3.5:
// same binade, same error budget -- DESIGN.md 3), then the reference's exact Float32 test
// (src/nhs_grid.jl:547-555) on every pre-filter hit; rank-balanced drain over kWPC warps.
#pragma once

#include <type_traits>

#include "sweep_tiles.cuh"

namespace pnb {

constexpr int kFG = 3;                  // lane groups (warps of points) per tile
constexpr int kFTP = kFG * 32;          // points per tile
constexpr int kFMinW = kFTP / 4;        // weight floor of a cell -> a tile touches <= 5 cells
constexpr int kFSMax = 5;               // cells per tile
constexpr int kFNSL = kFSMax + 2;       // staged x-columns ("slots") per tile
constexpr int kFSegCells = 1024;        // rows longer than this are cut into segments
constexpr bool kFlatBulkRecords = false; // stage the exact records with TMA bulk copies (measured slower)
constexpr bool kFlatBulkPayload = false; // ... and payload plane 0 as well

// staged candidates per tile (incl. the padding of every slot to a multiple of 32) and 32-blocks
// per cell (3 slots).  7 slots x 9 cells x 27 points = 1701 (+ padding) on the benchmark cloud.
template <class CL> __host__ __device__ constexpr int flat_cap() { return wants_big_tiles<CL>::value ? 2560 : 1984; }
template <class CL> __host__ __device__ constexpr int flat_nblk_max() { return wants_big_tiles<CL>::value ? 36 : 28; }
// warps per lane group: every group's blocks / hits are split over this many warps
template <class CL> __host__ __device__ constexpr int flat_wpc() { return CL::kWarpsPerCell <= 2 ? 3 : 5; }

// bytes of tile-kernel payload per candidate (closures may stage less than the ordered kernel)
template <class CL, class = void>
struct pay_bytes_tile { static constexpr int value = CL::kPayBytes; };
template <class CL>
struct pay_bytes_tile<CL, decltype((void)CL::kPayBytesTile)> { static constexpr int value = CL::kPayBytesTile; };
// closures whose per-pair term vanishes identically beyond the search radius (compact kernel
// support <= r) can skip the exact radius test of the drain: no_radius_test() is a member then
template <class CL, class = void>
struct has_no_radius_test { static constexpr bool value = false; };
template <class CL>
struct has_no_radius_test<CL, decltype((void)&CL::no_radius_test)> { static constexpr bool value = true; };
template <class CL>
__device__ __forceinline__ bool skip_radius_test(const CL &cl)
{
    if constexpr (has_no_radius_test<CL>::value) return cl.no_radius_test();
    else return false;
}
// closures that stage a compact payload for the tile kernels implement stage_tile / pair_tile
template <class CL, class = void>
struct has_stage_tile { static constexpr bool value = false; };
template <class CL>
struct has_stage_tile<CL, decltype((void)CL::kPayBytesTile)> { static constexpr bool value = true; };

// closures with asynchronous staging (cp.async) and a branch-free pair function for the drain
template <class CL, class = void>
struct has_stage_async { static constexpr bool value = false; };
template <class CL>
struct has_stage_async<CL, decltype((void)&CL::stage_async)> { static constexpr bool value = true; };
template <class CL, class = void>
struct has_bulk_plane { static constexpr bool value = false; };
template <class CL>
struct has_bulk_plane<CL, decltype((void)&CL::bulk_plane)> { static constexpr bool value = true; };
template <class CL, class = void>
struct has_pair_pred { static constexpr bool value = false; };
template <class CL>
struct has_pair_pred<CL, decltype((void)&CL::template pair_tile_pred<3>)> { static constexpr bool value = true; };

struct alignas(16) FlatTile {
    uint32_t cell;   // linear index of the first cell with points of the tile
    uint32_t off;    // first point of the tile inside that cell
    uint32_t npts;   // points of the tile (1 .. kFTP)
    uint32_t pad_;
};
// control words of one sweep: [0] number of tiles, [1] tile counter of the persistent CTAs,
// [2] overflow tiles, [3] SMs that have reported in (reserved-SM election)
constexpr int kFlatCtl = 4;
constexpr int kFlatSmStates = 1024;   // per-SM state words behind the control words (pnb_set_sweep_reserve)

template <int ND, class CL>
__host__ __device__ constexpr size_t flat_smem_bytes()
{
    constexpr int planes = CL::kCountOnly ? 1 : (needs_exact_masks<CL>::value ? 2 : 1);
    constexpr int cap = flat_cap<CL>(), nbm = flat_nblk_max<CL>();
    return sizeof(float4) * cap + (size_t)(cap / 32) * ND * 64 + (size_t)cap * pay_bytes_tile<CL>::value +
           (size_t)planes * kFG * nbm * 32 * 4;
}

__device__ __forceinline__ void group_barrier(int group, int nthreads)
{
    switch (group) {
        case 0: asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"r"(nthreads) : "memory"); break;
        default: asm volatile("bar.sync 3, %0;" ::"r"(nthreads) : "memory"); break;
    }
    static_assert(kFG == 3, "one named barrier per lane group");
}

// ------------------------------------------------------------------------------------------------
// pre-pass: cut every x-row (segment) of the query cell list into tiles
// ------------------------------------------------------------------------------------------------
// One warp per row segment.  Cells weigh max(count, kFMinW); W = exclusive prefix of the weights,
// P = exclusive prefix of the counts (both in shared memory); tile t of the segment covers the
// weighted positions [96 t, 96 t + 96).  EMIT = false: seg_tiles[seg] = number of non-empty
// tiles; EMIT = true: tile records written at seg_first[seg] + rank.  z0 / nz select the cell
// layers (last used dimension) that are swept: all of them, or the layers of one slab pass.
template <int ND, bool EMIT>
__global__ void __launch_bounds__(128)
k_flat_tiles(GridP g, CellsView qry, int lay0, int n_lay, int lay1, int n_lay1, int n_seg_row,
             uint32_t *__restrict__ seg_tiles,
             const uint32_t *__restrict__ seg_first, FlatTile *__restrict__ tiles,
             uint32_t *__restrict__ ctl, uint32_t n_ctas)
{
    __shared__ uint32_t s_w[4][kFSegCells + 1];
    __shared__ uint32_t s_p[4][kFSegCells + 1];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int nx = g.gs[0] - 2;
    const int ny = ND > 1 ? g.gs[1] - 2 : 1;
    // two layer ranges: [lay0, lay0 + n_lay) and [lay1, lay1 + n_lay1) (the second may be empty)
    const int nrow_y = (ND == 2) ? n_lay + n_lay1 : ny;   // rows per layer selection
    const int nrow_z = (ND == 3) ? n_lay + n_lay1 : 1;
    const int64_t n_segs = (int64_t)n_seg_row * nrow_y * nrow_z;
    const int64_t seg = (int64_t)blockIdx.x * 4 + warp;
    if (EMIT && blockIdx.x == 0 && threadIdx.x == 0) {
        ctl[1] = n_ctas;          // tile counter: the first n_ctas tiles are taken by blockIdx
        ctl[2] = 0u;              // overflow tiles
        ctl[3] = 0u;
    }
    if (EMIT && blockIdx.x == 0)
        for (int k = threadIdx.x; k < kFlatSmStates; k += blockDim.x) ctl[kFlatCtl + k] = 0u;
    if (seg >= n_segs) return;
    int64_t b = seg;
    const int isg = (int)(b % n_seg_row); b /= n_seg_row;
    const int iy = (int)(b % nrow_y);     b /= nrow_y;
    const int iz = (int)b;
    int cy = 1, cz = 1;
    if (ND == 2) cy = 2 + (iy < n_lay ? lay0 + iy : lay1 + (iy - n_lay));
    if (ND == 3) { cy = 2 + iy; cz = 2 + (iz < n_lay ? lay0 + iz : lay1 + (iz - n_lay)); }
    const int cxa = 2 + isg * kFSegCells;                         // first cell of the segment
    const int ncell = min(kFSegCells, nx - isg * kFSegCells);     // its cells
    uint32_t *W = s_w[warp], *P = s_p[warp];
    // prefix sums, 32 cells per round
    uint32_t runw = 0, runp = 0;
    for (int c0 = 0; c0 < ncell; c0 += 32) {
        const int c = c0 + lane;
        uint32_t cnt = 0, w = 0;
        if (c < ncell) {
            uint32_t b0;
            cell_range(qry, linear_cell(g, cxa + c, cy, cz), b0, cnt);
            w = max(cnt, (uint32_t)kFMinW);
        }
        uint32_t iw = w, ip = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tw = __shfl_up_sync(0xffffffffu, iw, o), tp = __shfl_up_sync(0xffffffffu, ip, o);
            if (lane >= o) { iw += tw; ip += tp; }
        }
        if (c < ncell) { W[c] = runw + iw - w; P[c] = runp + ip - cnt; }
        runw += __shfl_sync(0xffffffffu, iw, 31);
        runp += __shfl_sync(0xffffffffu, ip, 31);
    }
    if (lane == 0) { W[ncell] = runw; P[ncell] = runp; }
    __syncwarp();
    // R(p) = number of points at weighted positions < p
    auto cell_of = [&](uint32_t p) {       // last cell c with W[c] <= p   (p < W[ncell])
        int lo = 0, hi = ncell - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (W[mid] <= p) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    const uint32_t n_t = (runw + kFTP - 1) / kFTP;
    uint32_t emitted = 0;
    for (uint32_t t0 = 0; t0 < n_t; t0 += 32) {
        const uint32_t t = t0 + lane;
        uint32_t npts = 0, cell = 0, off = 0;
        if (t < n_t && runp > 0) {
            const uint32_t pa = t * kFTP, pb = min(pa + kFTP, runw);
            int ca = cell_of(pa);
            uint32_t oa = pa - W[ca];
            const uint32_t cnt_a = P[ca + 1] - P[ca];
            uint32_t ra = P[ca] + min(oa, cnt_a);
            uint32_t rb;
            {
                const int cb = cell_of(pb - 1);
                rb = P[cb] + min(pb - W[cb], P[cb + 1] - P[cb]);
            }
            npts = rb - ra;
            if (npts > 0) {
                // first cell that really holds a point of the tile
                while (oa >= (P[ca + 1] - P[ca])) { ca++; oa = 0; }
                cell = (uint32_t)linear_cell(g, cxa + ca, cy, cz);
                off = oa;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, npts > 0);
        if (EMIT && npts > 0) {
            FlatTile ft;
            ft.cell = cell; ft.off = off; ft.npts = npts; ft.pad_ = 0u;
            tiles[seg_first[seg] + emitted + __popc(m & ((1u << lane) - 1u))] = ft;
        }
        emitted += __popc(m);
    }
    if (!EMIT && lane == 0) seg_tiles[seg] = emitted;
    if (EMIT && seg == n_segs - 1 && lane == 0) ctl[0] = seg_first[seg] + emitted;
}

// ------------------------------------------------------------------------------------------------
// pre-pass, step 2: the staging tables of every tile
// ------------------------------------------------------------------------------------------------
// What a CTA needs before it can stage a tile -- begin / count / staged position of its <= 7 x 9
// cells, the padded start of every x-column, the query cells -- is computed HERE, one warp per
// tile with all tiles in flight, and written as one 896-byte table; the persistent CTAs copy the
// table of their next tile with cp.async while they work on the current one (in round 2's first
// version warp 0 of every CTA computed it serially: 133 dependent instructions per tile with
// 14 warps waiting at a barrier, 6.7 % of the warp samples).
constexpr int kTabWords = 224;
template <int ND>
struct FlatTab {
    static constexpr int NR = rows_of(ND);
    static constexpr int NE = kFNSL * NR;
    static constexpr int oCbeg = 0, oCcnt = NE, oCpre = 2 * NE, oSlot0 = 3 * NE;
    static constexpr int oSpop = oSlot0 + kFNSL + 1, oQb = oSpop + kFNSL, oQt = oQb + kFSMax;
    static constexpr int oNsl = oQt + kFSMax + 1, oCell = oNsl + 1, oOff = oCell + 1, oNpts = oOff + 1;
    static constexpr int oBig = oNpts + 1, oEnd = oBig + 1;
    static_assert(oEnd <= kTabWords, "table layout");
};

template <int ND, bool PER, int CAP, int NBM>
__global__ void __launch_bounds__(256)
k_flat_tables(GridP g, CellsView cand, CellsView qry, const FlatTile *__restrict__ tiles,
              const uint32_t *__restrict__ ctl, uint32_t *__restrict__ tabs)
{
    using L = FlatTab<ND>;
    constexpr int NR = L::NR, NEmax = L::NE;
    __shared__ uint32_t s_all[8][kTabWords];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const uint32_t tile = blockIdx.x * 8u + (uint32_t)warp;
    if (tile >= ctl[0]) return;
    uint32_t *T = s_all[warp];
    const FlatTile ft = tiles[tile];
    const int lin0 = (int)ft.cell;
    const int cx0 = lin0 % g.gs[0] + 1;
    const int cy = ND > 1 ? (lin0 / g.gs[0]) % g.gs[1] + 1 : 1;
    const int cz = ND > 2 ? lin0 / (g.gs[0] * g.gs[1]) + 1 : 1;
    for (int k = lane; k < kTabWords; k += 32) T[k] = 0u;
    __syncwarp();
    // tile cells: lane s = cell cx0 + s
    uint32_t b0 = 0, cntq = 0;
    if (lane < kFSMax && cx0 + lane <= g.gs[0] - 1)
        cell_range(qry, linear_cell(g, cx0 + lane, cy, cz), b0, cntq);
    uint32_t avail = cntq;
    if (lane == 0) { avail -= ft.off; b0 += ft.off; }
    uint32_t incl = avail;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t excl = incl - avail;
    // cells that hold a point of the tile: exclusive prefix < npts
    const unsigned touch = __ballot_sync(0xffffffffu, lane < kFSMax && excl < ft.npts && avail > 0u);
    const int S = 32 - __clz((int)touch);            // index of the last touched cell + 1
    if (lane < kFSMax) { T[L::oQb + lane] = b0; T[L::oQt + lane] = min(excl, ft.npts); }
    const int NSL = S + 2, NE = NSL * NR;
    if (lane == 0) {
        T[L::oQt + kFSMax] = ft.npts; T[L::oNsl] = (uint32_t)NSL;
        T[L::oCell] = ft.cell; T[L::oOff] = ft.off; T[L::oNpts] = ft.npts;
    }
    for (int e = lane; e < NEmax; e += 32) {
        const int slot = e / NR, row = e % NR;
        int sx = cx0 - 1 + slot;
        int ry = cy + (ND > 1 ? (row % 3) - 1 : 0);
        int rz = cz + (ND > 2 ? (row / 3) - 1 : 0);
        uint32_t c0 = 0, cn = 0;
        if (e < NE && sx <= g.gs[0]) {
            if (PER) {
                sx = floormod_i(sx - 2, g.nc[0]) + 2;
                if (ND > 1) ry = floormod_i(ry - 2, g.nc[1]) + 2;
                if (ND > 2) rz = floormod_i(rz - 2, g.nc[2]) + 2;
            }
            cell_range(cand, linear_cell(g, sx, ry, rz), c0, cn);
        }
        T[L::oCbeg + e] = c0;
        T[L::oCcnt + e] = cn;
    }
    __syncwarp();
    if (lane < kFNSL) {
        uint32_t run = 0;
#pragma unroll
        for (int r = 0; r < NR; r++) { T[L::oCpre + lane * NR + r] = run; run += T[L::oCcnt + lane * NR + r]; }
        T[L::oSpop + lane] = run;
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t run = 0;
#pragma unroll
        for (int sl = 0; sl < kFNSL; sl++) { T[L::oSlot0 + sl] = run; run += (T[L::oSpop + sl] + 31u) & ~31u; }
        T[L::oSlot0 + kFNSL] = run;
        // too dense for the staging buffer or the mask table: the per-point kernel takes the tile
        bool too_big = run > (uint32_t)CAP;
#pragma unroll
        for (int w = 0; w < kFSMax; w++)
            if (w + 3 <= NSL) too_big = too_big || (T[L::oSlot0 + w + 3] - T[L::oSlot0 + w]) > (uint32_t)(NBM * 32);
        T[L::oBig] = too_big ? 1u : 0u;
    }
    __syncwarp();
    for (int e = lane; e < NEmax; e += 32) T[L::oCpre + e] += T[L::oSlot0 + e / NR];
    __syncwarp();
    uint32_t *out = tabs + (size_t)tile * kTabWords;
    for (int k = lane; k < kTabWords; k += 32) out[k] = T[k];
}

// ------------------------------------------------------------------------------------------------
// the sweep
// ------------------------------------------------------------------------------------------------
// resident CTAs per SM the register allocation is tuned for: what the shared memory allows
// (227 KB per SM, ~6 KB of static tables per CTA), at most 4
template <int ND, class CL>
__host__ __device__ constexpr int flat_min_blocks()
{
    constexpr int threads = kFG * flat_wpc<CL>() * 32;
    constexpr int by_smem = (int)(232448 / (flat_smem_bytes<ND, CL>() + 6144));
    constexpr int by_threads = 2048 / threads;
    constexpr int m = by_smem < by_threads ? by_smem : by_threads;
    return m < 1 ? 1 : (m > 4 ? 4 : m);
}

// NOR2: the closure's term vanishes identically beyond the search radius (skip_radius_test,
// decided on the host): the exact radius test of the drain is compiled out.
template <int ND, bool PER, class CL, bool TWO, bool NOR2 = false>
__global__ void __launch_bounds__(kFG * flat_wpc<CL>() * 32, flat_min_blocks<ND, CL>())
k_sweep_flat(GridP g, CellsView cand, CellsView qry, CL cl, const uint32_t *__restrict__ tabs,
             uint32_t *__restrict__ ctl, int *__restrict__ overflow_tiles, int reserve_sms)
{
    // A persistent grid that fills every SM keeps the kernels of other streams (the NCCL send /
    // recv of the overlapped multi-GPU step need a mostly empty SM) from running before it ends:
    // CTAs that find themselves on one of the last `reserve_sms` SMs leave at once, the others
    // take their tiles from the common counter.
    // (SM ids need not be contiguous: the first `reserve_sms` DISTINCT SMs that report in are the
    // reserved ones.  sm_state[smid]: 0 unseen, 1 being decided, 2 reserved, 3 working.)
    if (reserve_sms > 0) {
        __shared__ int s_leave;
        if (threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned *sm_state = ctl + kFlatCtl + (smid & 1023u);
            unsigned st = atomicCAS(sm_state, 0u, 1u);
            if (st == 0u) {
                const unsigned rank = atomicAdd(ctl + 3, 1u);
                st = rank < (unsigned)reserve_sms ? 2u : 3u;
                atomicExch(sm_state, st);
            } else {
                while (st < 2u) st = atomicAdd(sm_state, 0u);
            }
            s_leave = st == 2u;
        }
        __syncthreads();
        if (s_leave) return;
    }
    constexpr int kWPC = flat_wpc<CL>();
    constexpr int NR = rows_of(ND);
    constexpr int NEmax = kFNSL * NR;
    constexpr int kGroupThreads = kWPC * 32;
    constexpr int kCap = flat_cap<CL>(), kBlocks = kCap / 32, kNBlkMax = flat_nblk_max<CL>();
    constexpr int kPayB = pay_bytes_tile<CL>::value;
    constexpr bool kExact = CL::kCountOnly || needs_exact_masks<CL>::value;
    using nz_t = typename std::conditional<(kNBlkMax > 32), unsigned long long, unsigned>::type;
    const float4 *__restrict__ sorted = cand.rec;
    const float4 *__restrict__ q_sorted = qry.rec;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_pos = reinterpret_cast<float4 *>(smem_raw);
    uint32_t *s_half = reinterpret_cast<uint32_t *>(smem_raw + sizeof(float4) * kCap);
    __half *s_half16 = reinterpret_cast<__half *>(s_half);
    unsigned char *s_pay = smem_raw + sizeof(float4) * kCap + (size_t)kBlocks * ND * 64;
    unsigned *s_mask = reinterpret_cast<unsigned *>(s_pay + (size_t)kCap * kPayB);
    using TL = FlatTab<ND>;
    __shared__ __align__(16) uint32_t s_tab[2][kTabWords];   // staging tables: this tile / the next one
    __shared__ int s_cnt[kFG][kWPC][32];
    __shared__ nz_t s_nz[kFG][kWPC][32];
    __shared__ uint32_t s_tile;
    __shared__ __align__(8) unsigned long long s_mbar;      // completion of the TMA bulk copies of a tile

    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int group = warp / kWPC, part = warp % kWPC;
    const PerP pp = make_perp(g);
    const float inv_r = __frcp_rn(g.r);
    float inv_bs[3] = {0.f, 0.f, 0.f};
    if (PER) {
#pragma unroll
        for (int d = 0; d < 3; d++) inv_bs[d] = __frcp_rn(pp.bs[d]);
    }
    uint32_t pos_sa = (uint32_t)__cvta_generic_to_shared(s_pos);
    asm volatile("mov.u32 %0, %0;" : "+r"(pos_sa));    // opaque: keep it in a register
    const uint32_t pay_sa = pos_sa + (uint32_t)(sizeof(float4) * kCap + (size_t)kBlocks * ND * 64);
    const __half2 thr = __float2half2_rn(half_thr_hi(PER));
    const __half2 thr_lo = __float2half2_rn(half_thr_lo(PER));
    const uint32_t n_tiles = ctl[0];
    const uint32_t mbar_sa = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    uint32_t mbar_phase = 0u;
    if (threadIdx.x == 0) {
        mbar_init(mbar_sa, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    unsigned *my_mask = s_mask + (size_t)group * kNBlkMax * 32 + lane;   // word of block b: my_mask[b * 32]
    unsigned *my_band = CL::kCountOnly ? my_mask : my_mask + (size_t)kFG * kNBlkMax * 32;

    // ---- tile pipeline of the persistent CTA ------------------------------------------------------
    // While tile t is processed, warp 0 has already requested the index of tile t + 2 (atomic
    // counter) and started the asynchronous copy (LDGSTS) of the 896-byte staging table of tile
    // t + 1 (written by k_flat_tables) into the other table buffer: a tile's prologue is one
    // cp.async wait and one barrier.
    const uint32_t tab_sa = (uint32_t)__cvta_generic_to_shared(&s_tab[0][0]);
    auto prefetch_table = [&](uint32_t t, int buf) {         // warp 0
        const uint32_t *src = tabs + (size_t)t * kTabWords;
        const uint32_t dst = tab_sa + (uint32_t)buf * (kTabWords * 4u);
        cp_async16(dst + 16u * (uint32_t)lane, src + 4 * lane);
        if (lane < kTabWords / 4 - 32) cp_async16(dst + 16u * (uint32_t)(32 + lane), src + 4 * (32 + lane));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t cur = blockIdx.x, nxt = 0xffffffffu;            // warp 0: this tile, the one after it
    int par = 0;
    if (warp == 0) {
        // (with reserved SMs the CTAs that left do not take the tile of their blockIdx: every
        // tile index then comes from the counter, which the tile pre-pass set to 0)
        if (reserve_sms > 0) {
            if (lane == 0) cur = atomicAdd(ctl + 1, 1u);
            cur = __shfl_sync(0xffffffffu, cur, 0);
        }
        if (lane == 0) nxt = atomicAdd(ctl + 1, 1u);
        nxt = __shfl_sync(0xffffffffu, nxt, 0);
        if (cur < n_tiles) prefetch_table(cur, 0);
    }

    for (;; par ^= 1) {
        __syncthreads();                         // shared memory of the previous tile is free
        if (warp == 0) {
            if (lane == 0) s_tile = cur;
            if (cur < n_tiles) cp_async_wait_all();     // the table of this tile has landed
        }
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= n_tiles) break;
        const uint32_t *T = s_tab[par];
        const uint32_t *s_cbeg = T + TL::oCbeg, *s_ccnt = T + TL::oCcnt, *s_cpre = T + TL::oCpre;
        const uint32_t *s_slot0 = T + TL::oSlot0, *s_spop = T + TL::oSpop;
        const uint32_t *s_qb = T + TL::oQb, *s_qt = T + TL::oQt;   // per tile cell: first record, first tile point
        FlatTile ft;
        ft.cell = T[TL::oCell]; ft.off = T[TL::oOff]; ft.npts = T[TL::oNpts]; ft.pad_ = 0u;
        const int lin0 = (int)ft.cell;
        const int cx0 = lin0 % g.gs[0] + 1;
        const int cy = ND > 1 ? (lin0 / g.gs[0]) % g.gs[1] + 1 : 1;
        const int cz = ND > 2 ? lin0 / (g.gs[0] * g.gs[1]) + 1 : 1;
        uint32_t nn = 0u;
        if (warp == 0) {
            // the tile after this one: request the index behind it, copy its table (both
            // asynchronous: the counter's answer is only looked at once this tile is staged)
            if (lane == 0) nn = atomicAdd(ctl + 1, 1u);
            if (nxt < n_tiles) prefetch_table(nxt, par ^ 1);
        }
        const int NSL = (int)T[TL::oNsl];
        // too dense for the staging buffer or the mask table: hand the tile to the per-point kernel
        if (T[TL::oBig]) {
            if (threadIdx.x == 0) overflow_tiles[atomicAdd(ctl + 2, 1u)] = (int)tile;
            if (warp == 0) { cur = nxt; nxt = __shfl_sync(0xffffffffu, nn, 0); }
            continue;
        }

        // ---- stage every cell of the table: exact records, fp16 copies, closure payload ------
        // tile centre in x: the staged columns cover cells cx0 - 1 .. cx0 + S
        float org[3] = {0.f, 0.f, 0.f};
        org[0] = fmaf((float)(cx0 + g.off[0] - 2) + 0.5f * (float)NSL, g.cs[0], g.minc[0]);
        if (ND > 1) org[1] = fmaf((float)(cy + g.off[1]) - 0.5f, g.cs[1], g.minc[1]);
        if (ND > 2) org[2] = fmaf((float)(cz + g.off[2]) - 0.5f, g.cs[2], g.minc[2]);
        {
            const int NE = NSL * NR;
            // Asynchronous copies global -> shared (LDGSTS) of the exact records and the payload:
            // every warp issues the copies of all its cells back to back (no registers in
            // between, one memory round trip per tile instead of one per cell), waits for its own
            // copies and derives the fp16 copies from the records it has just staged.
            constexpr bool kAsync = has_stage_async<CL>::value || pay_bytes_tile<CL>::value == 0;
            // TMA bulk copies (cp.async.bulk + mbarrier, UBLKCP / SYNCS in the SASS) were measured
            // against the per-thread cp.async below and LOST on this access pattern: a tile is 63
            // cells of ~430 bytes per plane, and the TMA unit works through such small requests
            // one after the other.  Config 3, WCSPH: per-thread LDGSTS 11.07 ms, records through
            // TMA 11.60 ms, records + payload plane 0 through TMA 12.22 ms (n-body 7.57 / 7.67 /
            // 7.66 ms; count 4.80 / 4.75 / 4.74 ms).  The path is kept behind this constant.
            constexpr bool kBulk = kFlatBulkRecords;
            constexpr bool kBulkPay = kBulk && kFlatBulkPayload && has_bulk_plane<CL>::value;
            if (kAsync) {
                // The 16-byte planes (exact records; plane 0 of the payload) go cell by cell through
                // the TMA unit: one cp.async.bulk per staged cell and plane, issued by the lanes
                // of the last warp, completion counted in bytes by an mbarrier.  What is left per
                // candidate (8 or 4 payload bytes) stays a per-thread cp.async.
                if (kBulk && warp == kFG * kWPC - 1) {
                    uint32_t bytes = 0;
                    for (int e = lane; e < NE; e += 32) bytes += s_ccnt[e] * 16u * (kBulkPay ? 2u : 1u);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                    if (lane == 0) mbar_arrive_expect_tx(mbar_sa, bytes);
                    // (the previous tile's reads of these buffers, ordered before this point by the
                    // CTA barrier, must also be ordered before the async proxy's writes)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    for (int e = lane; e < NE; e += 32) {
                        const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_ccnt[e];
                        if (n == 0u) continue;
                        bulk_copy_g2s(pos_sa + 16u * d0, sorted + b0, 16u * n, mbar_sa);
                        if constexpr (kBulkPay) bulk_copy_g2s(pay_sa + 16u * d0, cl.bulk_plane() + b0, 16u * n, mbar_sa);
                    }
                }
                if constexpr (!kBulk) {
                    for (int e = warp; e < NE; e += kFG * kWPC) {
                        const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_ccnt[e];
                        for (uint32_t k = lane; k < n; k += 32) {
                            cp_async16(pos_sa + 16u * (d0 + k), sorted + b0 + k);
                            if constexpr (has_stage_async<CL>::value) cl.stage_async(pay_sa, (int)(d0 + k), b0 + k, kCap);
                        }
                    }
                } else if constexpr (has_stage_async<CL>::value) {
                    for (int e = warp; e < NE; e += kFG * kWPC) {
                        const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_ccnt[e];
                        for (uint32_t k = lane; k < n; k += 32) {
                            if constexpr (kBulkPay) cl.stage_async_rest(pay_sa, (int)(d0 + k), b0 + k, kCap);
                            else cl.stage_async(pay_sa, (int)(d0 + k), b0 + k, kCap);
                        }
                    }
                }
                if constexpr (kBulk) {
                    mbar_wait(mbar_sa, mbar_phase);   // the bulk copies of this tile have landed
                    mbar_phase ^= 1u;
                }
                cp_async_wait_all();
            }
            for (int e = warp; e < NE; e += kFG * kWPC) {
                const uint32_t b0 = s_cbeg[e], d0 = s_cpre[e], n = s_ccnt[e];
                // periodic grids: the fp16 copy of a candidate is shifted by whole periods to the
                // image closest to the NOMINAL position of its cell next to the tile (minimum
                // image; points may lie outside the box, the reference wraps cells, not
                // coordinates).  Only the pre-filter sees the shifted copy.
                float cen[3] = {0.f, 0.f, 0.f};
                if (PER) {
                    const int slot = e / NR, row = e % NR;
                    const int sx = cx0 - 1 + slot;
                    const int ry = cy + (ND > 1 ? (row % 3) - 1 : 0);
                    const int rz = cz + (ND > 2 ? (row / 3) - 1 : 0);
                    cen[0] = fmaf((float)(sx + g.off[0]) - 0.5f, g.cs[0], g.minc[0]);
                    if (ND > 1) cen[1] = fmaf((float)(ry + g.off[1]) - 0.5f, g.cs[1], g.minc[1]);
                    if (ND > 2) cen[2] = fmaf((float)(rz + g.off[2]) - 0.5f, g.cs[2], g.minc[2]);
                }
                for (uint32_t k = lane; k < n; k += 32) {
                    const uint32_t q = d0 + k;
                    float4 pj;
                    if (kAsync) pj = s_pos[q];                  // copied by this very thread
                    else { pj = sorted[b0 + k]; s_pos[q] = pj; }
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
                    if (!kAsync) {
                        if constexpr (has_stage_tile<CL>::value) cl.stage_tile(s_pay, (int)q, b0 + k, kCap);
                        else cl.stage(s_pay, (int)q, b0 + k, kCap);
                    }
                }
            }
            // padding between a slot's last candidate and the next multiple of 32
            for (int sl = warp; sl < NSL; sl += kFG * kWPC) {
                const uint32_t q = s_slot0[sl] + s_spop[sl] + (uint32_t)lane;
                if (q < s_slot0[sl + 1]) {
                    const uint32_t hb = (q >> 5) * (ND * 32) + (q & 15u) * 2u + ((q >> 4) & 1u);
                    const __half far = __float2half_rn(kHalfSentinel);
                    s_half16[hb] = far;
                    if (ND > 1) s_half16[hb + 32] = far;
                    if (ND > 2) s_half16[hb + 64] = far;
                    s_pos[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
        __syncthreads();
        if (warp == 0) { cur = nxt; nxt = __shfl_sync(0xffffffffu, nn, 0); }

        // ---- my point: tile point tp = group * 32 + lane -------------------------------------
        const uint32_t tp = (uint32_t)(group * 32 + lane);
        const bool active = tp < ft.npts;
        int cs_ = 0;                                           // local cell of the point
#pragma unroll
        for (int w = 1; w < kFSMax; w++) cs_ += (active && tp >= s_qt[w] && s_qt[w] < ft.npts) ? 1 : 0;
        const uint32_t k_in = tp - s_qt[cs_];                  // index inside the tile's part of the cell
        const uint32_t i_sorted = s_qb[cs_] + k_in;            // record in the query list
        const int B0 = (int)(s_slot0[cs_] >> 5);
        const int nblk = active ? (int)((s_slot0[cs_ + 3] - s_slot0[cs_]) >> 5) : 0;
        const int nblk_g = __reduce_max_sync(0xffffffffu, nblk);
        const int nb4 = (nblk_g + kWPC - 1) / kWPC;            // blocks per share (uniform in the group)
        const int blk_lo = min(part * nb4, nblk_g), blk_hi = min(blk_lo + nb4, nblk_g);
        float xi = 0.f, yi = 0.f, zi = 0.f;
        int i_id = 0;
        __half2 hx = __float2half2_rn(0.f), hy = hx, hz = hx;
        typename CL::State st;
        if (active && !TWO) {
            // x === y: the point is one of the staged candidates of its own cell
            const uint32_t q = s_cpre[(cs_ + 1) * NR + NR / 2] + (cs_ == 0 ? ft.off : 0u) + k_in;
            const float4 pi = s_pos[q];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
            const uint32_t hb = (q >> 5) * (ND * 32) + (q & 15u) * 2u + ((q >> 4) & 1u);
            hx = __half2half2(s_half16[hb]);
            if (ND > 1) hy = __half2half2(s_half16[hb + 32]);
            if (ND > 2) hz = __half2half2(s_half16[hb + 64]);
        }
        if (active && TWO) {
            const float4 pi = q_sorted[i_sorted];
            xi = pi.x; yi = pi.y; zi = pi.z;
            i_id = __float_as_int(pi.w);
            float ux = xi - org[0], uy = yi - org[1], uz = zi - org[2];
            if (PER) {
                const float c0 = fmaf((float)(cx0 + cs_ + g.off[0]) - 0.5f, g.cs[0], g.minc[0]);
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
        cl.init(st, active, TWO ? -1 : (int)i_sorted, i_id);

        // ---- phase 1: test my share of the point's blocks (block index relative to B0) --------
        // kExact: the closure needs EXACT masks / counts: certain hits (below the lower threshold)
        // go to the hit masks, the band between the two thresholds to a second plane and is
        // decided by the exact test below.
        int cnt = 0, n_maybe = 0;
        nz_t nzp = 0;
        for (int bb = blk_lo; bb < blk_hi; bb++) {
            const bool in = bb < nblk;
            const uint32_t *hb = s_half + (size_t)(B0 + (in ? bb : 0)) * (ND * 16);
            unsigned hh;
            if (kExact) {
                unsigned sure;
                hh = test_block_half<ND, true>(hb, hx, hy, hz, thr, thr_lo, &sure);
                if (!in) { hh = 0u; sure = 0u; }
                my_band[bb * 32] = hh & ~sure;
                n_maybe += __popc(hh & ~sure);
                hh = sure;
            } else {
                hh = test_block_half<ND, false>(hb, hx, hy, hz, thr, thr_lo, nullptr);
                if (!in) hh = 0u;
            }
            if (!CL::kCountOnly) my_mask[bb * 32] = hh;
            if (!kExact && hh) nzp |= (nz_t)1 << bb;
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
            if (!CL::kCountOnly)
                for (int bb2 = blk_lo; bb2 < blk_hi; bb2++)
                    if (my_mask[bb2 * 32]) nzp |= (nz_t)1 << bb2;
        }
        s_cnt[group][part][lane] = cnt;
        if (!CL::kCountOnly) s_nz[group][part][lane] = nzp;
        group_barrier(group, kGroupThreads);
        if (CL::kCountOnly) {
            if (part == 0) {
                int tot = 0;
#pragma unroll
                for (int p = 0; p < kWPC; p++) tot += s_cnt[group][p][lane];
                cl.count(st, tot);
                if (active) cl.finish(st, TWO ? -1 : (int)i_sorted, i_id);
            }
        } else {
            // ---- phase 2: my share of the point's hits: ranks [part * Q, part * Q + n_mine) -----
            int pc[kWPC];
            int H = 0;
            nz_t nz = 0;
#pragma unroll
            for (int p = 0; p < kWPC; p++) { pc[p] = s_cnt[group][p][lane]; H += pc[p]; nz |= s_nz[group][p][lane]; }
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
                // non-empty blocks after this one
                nz &= ~(((nz_t)2 << bb) - (nz_t)1);
            }
            const int rounds = __reduce_max_sync(0xffffffffu, n_mine);
            cl.seek(st, part * Q);
            constexpr bool no_r2 = !kExact && NOR2;

            // ---- phase 3: drain, one hit per lane and round -----------------------------------
            if constexpr (has_pair_pred<CL>::value && !kExact) {
                // branch-free rounds: a lane that has run out of hits re-reads its own record
                // (contribution exactly zero); record and payload loads do not wait for the test
                const uint32_t q_safe = TWO ? 0u : s_cpre[(cs_ + 1) * NR + NR / 2];
                for (int t = 0; t < rounds; t++) {
                    const bool on = t < n_mine;
                    if (on && mm == 0u) {
                        // next non-empty block of this point (there is one: t < n_mine)
                        if (sizeof(nz_t) == 8) bb = __ffsll((long long)nz) - 1;
                        else bb = __ffs((int)nz) - 1;
                        nz &= nz - (nz_t)1;
                        mm = my_mask[bb * 32];
                    }
                    const int k = __ffs(mm) - 1;
                    mm &= mm - 1u;
                    const int slot = on ? 32 * (B0 + bb) + k : (int)q_safe;
                    const float4 pj = lds128(pos_sa + 16u * (uint32_t)slot);
                    float px = __fsub_rn(xi, pj.x);
                    float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    // the decision: the reference's exact test (the fp16 pass only pre-selects)
                    const bool ok = on && (no_r2 || d2 <= pp.r2);
                    cl.template pair_tile_pred<ND>(st, px, py, pz, d2, __float_as_int(pj.w), pay_sa, slot, kCap, ok);
                }
            } else {
                for (int t = 0; t < rounds; t++) {
                    if (t < n_mine) {
                        if (mm == 0u) {
                            if (sizeof(nz_t) == 8) bb = __ffsll((long long)nz) - 1;
                            else bb = __ffs((int)nz) - 1;
                            nz &= nz - (nz_t)1;
                            mm = my_mask[bb * 32];
                        }
                        const int k = __ffs(mm) - 1;
                        mm &= mm - 1u;
                        const int slot = 32 * (B0 + bb) + k;
                        const float4 pj = lds128(pos_sa + 16u * (uint32_t)slot);
                        float px = __fsub_rn(xi, pj.x);
                        float py = ND > 1 ? __fsub_rn(yi, pj.y) : 0.f;
                        float pz = ND > 2 ? __fsub_rn(zi, pj.z) : 0.f;
                        float d2 = dist2<ND>(px, py, pz);
                        d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                        if (kExact || no_r2 || d2 <= pp.r2) {
                            if constexpr (has_stage_tile<CL>::value)
                                cl.template pair_tile<ND>(st, px, py, pz, d2, __float_as_int(pj.w), pay_sa, slot, kCap);
                            else
                                cl.template pair_s<ND>(st, px, py, pz, d2, __float_as_int(pj.w), pay_sa, slot, kCap);
                        }
                    }
                }
            }
            cl.flush(st);
            // ---- phase 4: add the kWPC partial accumulators of every point ----------------------
            // (the group's mask words are free once all its warps have finished their drain)
            group_barrier(group, kGroupThreads);
            float *s_red = reinterpret_cast<float *>(s_mask + (size_t)group * kNBlkMax * 32);
            static_assert((kWPC - 1) * CL::kAccWords <= kNBlkMax, "partials fit the mask words");
            if (part > 0) cl.save_acc(st, s_red + ((part - 1) * CL::kAccWords) * 32 + lane);
            group_barrier(group, kGroupThreads);
            if (part == 0) {
#pragma unroll
                for (int p = 0; p < kWPC - 1; p++) cl.add_acc(st, s_red + (p * CL::kAccWords) * 32 + lane);
                if (active) cl.finish(st, TWO ? -1 : (int)i_sorted, i_id);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// tiles that did not fit the staging buffer: one point at a time from global memory
// ------------------------------------------------------------------------------------------------
// G lanes share a point: lane s takes candidates s, s + G, ... of every neighbour cell, the
// reference's exact test, the closure's global-memory pair function; partial results are merged
// through the closure's save_acc / add_acc.  Closures whose result depends on the rank of a hit
// (neighbour-list fill) use one lane per point.
template <int ND, bool PER, class CL, bool TWO>
__global__ void __launch_bounds__(128)
k_sweep_flat_overflow(GridP g, CellsView cand, CellsView qry, CL cl, const FlatTile *__restrict__ tiles,
                      const uint32_t *__restrict__ ctl, const int *__restrict__ overflow_tiles)
{
    constexpr int G = needs_exact_masks<CL>::value ? 1 : 8;
    constexpr int kWords = CL::kAccWords > 0 ? CL::kAccWords : 1;
    constexpr int PPB = 128 / G;                      // points per block and round
    __shared__ float s_red[4][kWords][32];
    const int n_ovf = (int)ctl[2];
    const PerP pp = make_perp(g);
    const int lane = lane_id(), sub = lane % G;
    for (int o = blockIdx.x; o < n_ovf; o += gridDim.x) {
        const FlatTile ft = tiles[overflow_tiles[o]];
        const int lin0 = (int)ft.cell;
        const int cx0 = lin0 % g.gs[0] + 1;
        const int cy = ND > 1 ? (lin0 / g.gs[0]) % g.gs[1] + 1 : 1;
        const int cz = ND > 2 ? lin0 / (g.gs[0] * g.gs[1]) + 1 : 1;
        for (uint32_t t0 = 0; t0 < ft.npts; t0 += PPB) {
            const uint32_t tp = t0 + threadIdx.x / G;
            const bool have = tp < ft.npts;
            // cell and record of tile point tp: walk the tile's cells
            int cx = cx0;
            uint32_t b0 = 0, cntq = 0, skip = ft.off + (have ? tp : 0u);
            cell_range(qry, linear_cell(g, cx, cy, cz), b0, cntq);
            while (have && skip >= cntq) {
                skip -= cntq;
                cx++;
                cell_range(qry, linear_cell(g, cx, cy, cz), b0, cntq);
            }
            const uint32_t i_sorted = b0 + skip;
            float4 pi = make_float4(0.f, 0.f, 0.f, 0.f);
            if (have) pi = qry.rec[i_sorted];
            const int i_id = __float_as_int(pi.w);
            typename CL::State st;
            cl.init(st, have, TWO ? -1 : (int)i_sorted, i_id);
            int hits = 0;
            constexpr int NC = ND == 3 ? 27 : (ND == 2 ? 9 : 3);
            constexpr int U = G == 1 ? 8 : 4;          // loads in flight per lane
#pragma unroll 1
            for (int e = 0; e < NC; e++) {
                int c0 = cx + (e % 3) - 1;
                int c1 = cy + (ND > 1 ? (e / 3) % 3 - 1 : 0);
                int c2 = cz + (ND > 2 ? e / 9 - 1 : 0);
                if (PER) {
                    c0 = floormod_i(c0 - 2, g.nc[0]) + 2;
                    if (ND > 1) c1 = floormod_i(c1 - 2, g.nc[1]) + 2;
                    if (ND > 2) c2 = floormod_i(c2 - 2, g.nc[2]) + 2;
                }
                uint32_t cb, cn;
                cell_range(cand, linear_cell(g, c0, c1, c2), cb, cn);
                if (!have) cn = 0u;
                for (uint32_t k = (uint32_t)sub; k < cn; k += (uint32_t)(G * U)) {
                    float4 pj[U];
#pragma unroll
                    for (int u = 0; u < U; u++)
                        pj[u] = __ldg(cand.rec + cb + min(k + (uint32_t)(u * G), cn - 1u));
#pragma unroll
                    for (int u = 0; u < U; u++) {
                        if (k + (uint32_t)(u * G) >= cn) break;
                        const uint32_t gi = cb + k + (uint32_t)(u * G);
                        float px = __fsub_rn(pi.x, pj[u].x);
                        float py = ND > 1 ? __fsub_rn(pi.y, pj[u].y) : 0.f;
                        float pz = ND > 2 ? __fsub_rn(pi.z, pj[u].z) : 0.f;
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
                    for (int w = G / 2; w > 0; w >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, w);
                } else {
                    float *mine = &s_red[threadIdx.x >> 5][0][lane];
                    __syncwarp();
                    if (sub != 0) cl.save_acc(st, mine);
                    __syncwarp();
                    if (sub == 0) {
#pragma unroll
                        for (int u = 1; u < G; u++) cl.add_acc(st, mine + u);
                    }
                    __syncwarp();
                }
            }
            if (CL::kCountOnly) cl.count(st, hits);
            if (have && sub == 0) cl.finish(st, TWO ? -1 : (int)i_sorted, i_id);
        }
    }
}

}  // namespace pnb
