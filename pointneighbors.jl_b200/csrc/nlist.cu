// nlist.cu -- PrecomputedNeighborhoodSearch on the device: neighbour lists as CSR
// (count pass, 64-bit decoupled-lookback scan, fill pass, per-list sort), exports in the
// reference's DynamicVectorOfVectors layouts, the list sweep without radius test and the TLSPH
// deformation gradient.
//
// reference: src/nhs_precomputed.jl:130-247 (initialize!, initialize_neighbor_lists!, sweep),
// src/vector_of_vectors.jl:3-31,177-212 (layout, sorteach!),
// benchmarks/smoothed_particle_hydrodynamics.jl:136-189 (TLSPH set-up).
#include <mutex>
#include <cstdlib>
#include <cstring>

#include "closures.cuh"
#include "sweep.cuh"
#include "sweep_launch.cuh"
#include "f64.cuh"
#include "tlsph_interact.cuh"

struct pnb_nlist {
    int64_t nx;        // number of lists (points of x)
    int64_t n_pairs;
    int64_t max_len;   // longest list
    int64_t *offsets;  // [nx+1] device
    int32_t *ids;      // [n_pairs] device, 0-based
    uint32_t *counts;  // [nx] device
    int ndims;
    int *d_err;
    int *h_err;
    float4 *pack;      // [2 nx] per-point records of the TLSPH sweep (allocated on first use)
    float4 *pack8;     // [8 nx] 128-byte records of the TLSPH force sweep (allocated on first use)
    size_t bytes_offsets, bytes_ids, bytes_counts, bytes_pack, bytes_pack8;
};

namespace pnb {

// One spare device buffer per kind survives pnb_nlist_destroy, so that rebuilding the lists every
// step (update! of a PrecomputedNeighborhoodSearch) does not pay cudaMalloc / cudaFree of
// gigabytes each time.  Keyed by the device the buffer lives on.
// Thread-safe (mutex); a buffer is handed out again only after the device has finished with it:
// cached_free synchronises the device before parking the buffer, which costs nothing on the
// paths that call it (pnb_nlist_destroy follows blocking calls).
struct SpareBuf { void *p; size_t bytes; int device; };
static SpareBuf g_spare[6] = {{nullptr, 0, -1}, {nullptr, 0, -1}, {nullptr, 0, -1}, {nullptr, 0, -1},
                              {nullptr, 0, -1}, {nullptr, 0, -1}};
static std::mutex g_spare_mutex;

static cudaError_t cached_malloc(int kind, void **out, size_t bytes, size_t *got)
{
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_spare_mutex);
        SpareBuf &sp = g_spare[kind];
        if (sp.p && sp.device == dev && sp.bytes >= bytes && sp.bytes <= 2 * bytes + (1u << 20)) {
            *out = sp.p; *got = sp.bytes;
            sp.p = nullptr; sp.bytes = 0;
            return cudaSuccess;
        }
    }
    *got = bytes;
    return cudaMalloc(out, bytes);
}

static void cached_free(int kind, void *p, size_t bytes)
{
    if (!p) return;
    int dev = 0;
    cudaGetDevice(&dev);
    // kernels of ANY stream may still read the buffer (lists are swept on the caller's streams)
    cudaDeviceSynchronize();
    void *drop = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_spare_mutex);
        SpareBuf &sp = g_spare[kind];
        if (sp.p == nullptr || sp.bytes < bytes) {
            drop = sp.p;
            sp.p = p; sp.bytes = bytes; sp.device = dev;
        } else {
            drop = p;
        }
    }
    if (drop) cudaFree(drop);
}

// longest list -> out[0] (the reference's overflow check compares it with max_neighbors)
__global__ void k_max_count(int64_t n, const uint32_t *__restrict__ counts, unsigned int *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = i < n ? counts[i] : 0u;
    v = __reduce_max_sync(0xffffffffu, v);
    if (lane_id() == 0 && v > 0u) atomicMax(out, v);
}

// count pass: lengths of the lists (the `lengths[i] += 1` half of pushat!, vector_of_vectors.jl:83)
struct ListCountCl {
    static constexpr bool kCountOnly = true;
    static constexpr int kPayBytes = 0;
    static constexpr int kWarpsPerCell = 2;
    static constexpr int kAccWords = 1;
    static constexpr bool kBigTiles = true;
    uint32_t *out;
    struct State { int cnt; };
    __device__ __forceinline__ void save_acc(const State &, float *) const {}
    __device__ __forceinline__ void add_acc(State &, const float *) const {}
    __device__ __forceinline__ void seek(State &, int) const {}
    __device__ __forceinline__ void total(State &, int, bool) const {}
    __device__ __forceinline__ void flush(State &) const {}
    template <int ND>
    __device__ __forceinline__ void pair_s(State &, float, float, float, float, int, uint32_t, int,
                                           int) const {}
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
    static constexpr int kWarpsPerCell = 4;
    static constexpr int kAccWords = 1;
    // the position of a hit in its list is its rank among the point's hits: the tile kernel
    // must hand over exact hit masks (no fp16 pre-filter)
    static constexpr bool kExactMasks = true;
    static constexpr bool kBigTiles = true;
    const int64_t *offsets;
    int32_t *ids;
    // b0..b3: up to four ids waiting for one 16-byte store (tile kernel only)
    struct State { int64_t pos; int b0, b1, b2, b3, nb; };
    __device__ __forceinline__ void save_acc(const State &, float *) const {}
    __device__ __forceinline__ void add_acc(State &, const float *) const {}
    __device__ __forceinline__ void seek(State &s, int first_rank) const { s.pos += first_rank; }
    __device__ __forceinline__ void total(State &, int, bool) const {}
    // Every lane appends to its own list, so a warp-wide 4-byte store touches 32 different
    // sectors.  Ids are therefore collected four at a time and written with one 16-byte store
    // once the write position is 16-byte aligned (scalar stores before that and in flush()).
    template <int ND>
    __device__ __forceinline__ void pair_s(State &s, float, float, float, float, int j_id, uint32_t,
                                           int, int) const
    {
        if (s.nb == 0 && (s.pos & 3) != 0) { ids[s.pos++] = j_id; return; }
        if (s.nb == 0) s.b0 = j_id;
        else if (s.nb == 1) s.b1 = j_id;
        else if (s.nb == 2) s.b2 = j_id;
        else s.b3 = j_id;
        if (++s.nb == 4) {
            *reinterpret_cast<int4 *>(ids + s.pos) = make_int4(s.b0, s.b1, s.b2, s.b3);
            s.pos += 4;
            s.nb = 0;
        }
    }
    __device__ __forceinline__ void flush(State &s) const
    {
        if (s.nb > 0) ids[s.pos] = s.b0;
        if (s.nb > 1) ids[s.pos + 1] = s.b1;
        if (s.nb > 2) ids[s.pos + 2] = s.b2;
        s.pos += s.nb;
        s.nb = 0;
    }
    __device__ __forceinline__ void init(State &s, bool active, int, int i_id) const
    {
        s.pos = active ? offsets[i_id] : 0;
        s.b0 = s.b1 = s.b2 = s.b3 = 0;
        s.nb = 0;
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

// One-pass fill into rows of fixed capacity (the reference's own layout, max_neighbors x N,
// vector_of_vectors.jl:3-31, with the capacity taken from the longest list of the previous
// build): no count pass, no scan before the fill.  The tile kernel tells the closure the total
// number of hits of a point (total()), which becomes lengths[i]; a list longer than the capacity
// sets bit 3 of *err and is not written (the caller then falls back to the two-pass build).
struct ListFillRowsCl {
    static constexpr bool kCountOnly = false;
    static constexpr int kPayBytes = 0;
    static constexpr int kWarpsPerCell = 4;
    static constexpr int kAccWords = 1;
    static constexpr bool kExactMasks = true;
    static constexpr bool kBigTiles = true;
    int32_t *rows;       // [nx * cap]
    uint32_t *lengths;   // [nx]
    int cap;             // multiple of 4: every row starts 16-byte aligned
    int *err;
    // n >= 0: hits appended by a kernel that walks the candidates one by one (overflow / per-point
    // kernels; finish() writes the length); n == -1: the tile kernel already wrote it (total())
    struct State { int64_t pos; int b0, b1, b2, b3, nb; int n; int i; bool skip; };
    __device__ __forceinline__ void save_acc(const State &, float *) const {}
    __device__ __forceinline__ void add_acc(State &, const float *) const {}
    __device__ __forceinline__ void seek(State &s, int first_rank) const { s.pos += first_rank; }
    __device__ __forceinline__ void total(State &s, int hits, bool writer) const
    {
        s.n = -1;
        if (hits > cap) { s.skip = true; if (writer) atomicOr(err, 8); }
        if (writer) lengths[s.i] = (uint32_t)hits;
    }
    template <int ND>
    __device__ __forceinline__ void pair_s(State &s, float, float, float, float, int j_id, uint32_t,
                                           int, int) const
    {
        if (s.skip) return;
        if (s.nb == 0 && (s.pos & 3) != 0) { rows[s.pos++] = j_id; return; }
        if (s.nb == 0) s.b0 = j_id;
        else if (s.nb == 1) s.b1 = j_id;
        else if (s.nb == 2) s.b2 = j_id;
        else s.b3 = j_id;
        if (++s.nb == 4) {
            *reinterpret_cast<int4 *>(rows + s.pos) = make_int4(s.b0, s.b1, s.b2, s.b3);
            s.pos += 4;
            s.nb = 0;
        }
    }
    __device__ __forceinline__ void flush(State &s) const
    {
        if (s.nb > 0) rows[s.pos] = s.b0;
        if (s.nb > 1) rows[s.pos + 1] = s.b1;
        if (s.nb > 2) rows[s.pos + 2] = s.b2;
        s.pos += s.nb;
        s.nb = 0;
    }
    __device__ __forceinline__ void init(State &s, bool active, int, int i_id) const
    {
        s.pos = active ? (int64_t)i_id * cap : 0;
        s.b0 = s.b1 = s.b2 = s.b3 = 0;
        s.nb = 0;
        s.n = 0;
        s.i = i_id;
        s.skip = !active;
    }
    __device__ __forceinline__ void stage(unsigned char *, int, uint32_t, int) const {}
    __device__ __forceinline__ void count(State &, int) const {}
    __device__ __forceinline__ void append(State &s, int j_id) const
    {
        if (s.n < cap) rows[s.pos + s.n] = j_id;
        s.n++;
    }
    template <int ND>
    __device__ __forceinline__ void pair(State &s, float, float, float, float, int j_id,
                                         const unsigned char *, int, int) const { append(s, j_id); }
    template <int ND>
    __device__ __forceinline__ void pair_global(State &s, float, float, float, float, int j_id,
                                                uint32_t) const { append(s, j_id); }
    __device__ __forceinline__ void finish(State &s, int, int i_id) const
    {
        if (s.n < 0) return;
        if (s.n > cap) atomicOr(err, 8);
        lengths[i_id] = (uint32_t)s.n;
    }
};

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

// sorteach! fused with the compaction of the fixed-capacity rows into the CSR list: one warp per
// list reads its row, sorts (registers up to 128 entries) and writes ids[offsets[i] ...].
template <int NR>
__device__ __forceinline__ void sort_row_regs(const int32_t *src, int32_t *dst, int len, int lane)
{
    int32_t v[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) v[r] = (r * 32 + lane < len) ? src[r * 32 + lane] : 0x7fffffff;
    bitonic_sort_regs<NR>(v, lane);
#pragma unroll
    for (int r = 0; r < NR; r++)
        if (r * 32 + lane < len) dst[r * 32 + lane] = v[r];
}

__global__ void __launch_bounds__(kSortWarps * 32)
k_sort_compact_rows(int64_t nx, int cap, const int32_t *__restrict__ rows,
                    const uint32_t *__restrict__ lengths, const int64_t *__restrict__ offsets,
                    int32_t *__restrict__ ids, int sort)
{
    __shared__ int32_t s_buf[kSortWarps][kSortCap];
    const int warp = threadIdx.x >> 5, lane = lane_id();
    const int64_t i = (int64_t)blockIdx.x * kSortWarps + warp;
    if (i >= nx) return;
    const int len = (int)lengths[i];
    if (len == 0) return;
    const int32_t *src = rows + i * (int64_t)cap;
    int32_t *dst = ids + offsets[i];
    if (!sort) {
        for (int e = lane; e < len; e += 32) dst[e] = src[e];
    } else if (len <= 32) {
        sort_row_regs<1>(src, dst, len, lane);
    } else if (len <= 64) {
        sort_row_regs<2>(src, dst, len, lane);
    } else if (len <= 128) {
        sort_row_regs<4>(src, dst, len, lane);
    } else if (len <= kSortCap) {
        int32_t *buf = s_buf[warp];
        for (int e = lane; e < len; e += 32) buf[e] = src[e];
        __syncwarp();
        for (int e = lane; e < len; e += 32) {
            const int32_t v = buf[e];
            int r = 0;
            for (int k = 0; k < len; k++) r += (buf[k] < v);
            dst[r] = v;
        }
    } else {
        for (int e = lane; e < len; e += 32) dst[e] = src[e];
        __syncwarp();
        for (int pass = 0; pass < len; pass++) {
            for (int e = (pass & 1) + 2 * lane; e + 1 < len; e += 64) {
                const int32_t a = dst[e], b = dst[e + 1];
                if (a > b) { dst[e] = b; dst[e + 1] = a; }
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
// Per-point records of the TLSPH sweep: rec[2 j] = (X0_j, -m_j / rho0_j), rec[2 j + 1] = (x_j, 0):
// what a pair needs from its neighbour is ONE aligned 32-byte sector instead of eight 4-byte
// gathers from four arrays.
template <int ND>
__global__ void k_pack_tlsph(int64_t n, const float *__restrict__ X0, const float *__restrict__ xcur,
                             const float *__restrict__ mass, const float *__restrict__ rho0,
                             float4 *__restrict__ rec)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    a.x = __ldg(X0 + j * ND); b.x = __ldg(xcur + j * ND);
    if (ND > 1) { a.y = __ldg(X0 + j * ND + 1); b.y = __ldg(xcur + j * ND + 1); }
    if (ND > 2) { a.z = __ldg(X0 + j * ND + 2); b.z = __ldg(xcur + j * ND + 2); }
    a.w = -__fdiv_rn(__ldg(mass + j), __ldg(rho0 + j));
    rec[2 * j] = a;
    rec[2 * j + 1] = b;
}

// one 256-bit load (LDG.E.256, sm_100): the two halves of a 32-byte record in ONE L1 wavefront
__device__ __forceinline__ void ldg256(const float4 *p, float4 &a, float4 &b)
{
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// One warp per point: lanes stride over the point's list (coalesced id reads), gather the
// neighbour's 32-byte record, accumulate the ND x ND outer products privately, then a shuffle
// tree adds the 32 partial matrices.  HBM: 4 B per pair + 108 B per point; the
// binding resource is L2 -> SM gather bandwidth (one sector per pair).
// EXACT: the oracle's IEEE operation sequence per term; fast (default): one MUFU.RSQ + FMAs,
// grad W / d = (-5 sigma / h^2) t^3 with t = 1 - d / (2 h) (the d of w(q)/d cancels).
template <int ND, bool PER, bool EXACT>
__global__ void __launch_bounds__(256, 4)   // 64 registers: 32 warps / SM (5 blocks spill: 5.3 ms)
k_tlsph_defgrad(GridP g, int64_t n, const int64_t *__restrict__ offsets,
                const int32_t *__restrict__ ids, const float4 *__restrict__ rec,
                const float *__restrict__ L, float h, float kernel_norm, float *__restrict__ F)
{
    // kG lanes per point (4 points per warp): the per-point work (loads of L, the reduction tree,
    // the store of F) is shared by 4 points per warp instruction, and a list of ~108 neighbours
    // fills 14 rounds of 8 lanes to 96 % (one warp per point: 4 rounds of 32 lanes, 84 %).
    constexpr int kG = 8;
    constexpr int kU = 2;      // pairs per lane in flight: ids first, then records, then arithmetic
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kG;
    const int sub = (int)(threadIdx.x % kG);
    const bool have = i < n;
    const int64_t ii = have ? i : 0;
    constexpr int NN = ND * ND;
    float Li[NN];
    float4 ra, rb;
    ldg256(rec + 2 * ii, ra, rb);
    const float Xi[3] = {ra.x, ra.y, ra.z}, xi[3] = {rb.x, rb.y, rb.z};
#pragma unroll
    for (int e = 0; e < NN; e++) Li[e] = __ldg(L + ii * NN + e);
    float acc[NN];
#pragma unroll
    for (int e = 0; e < NN; e++) acc[e] = 0.f;
    const float nh = __fdiv_rn(kernel_norm, h);
    const float k5 = -5.0f * kernel_norm / (h * h), nhalf_inv_h = -0.5f / h;
    const PerP pp = make_perp(g);
    const int64_t k_beg = have ? offsets[ii] : 0;
    const int len = have ? (int)(offsets[ii + 1] - k_beg) : 0;
    const int32_t *my_ids = ids + k_beg;
    int max_len = len;
#pragma unroll
    for (int o = 16; o >= kG; o >>= 1) max_len = max(max_len, __shfl_xor_sync(0xffffffffu, max_len, o));
    // the ids of the NEXT round are fetched while the records of this round are in flight
    int jn[kU];
#pragma unroll
    for (int u = 0; u < kU; u++) jn[u] = (sub + kG * u < len) ? __ldg(my_ids + sub + kG * u) : -1;
    for (int k0 = sub; k0 < max_len; k0 += kG * kU) {
        int jj[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) jj[u] = jn[u];
        float4 ra4[kU], rb4[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const int64_t j = jj[u] >= 0 ? jj[u] : ii;
            ldg256(rec + 2 * j, ra4[u], rb4[u]);
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            const int kn = k0 + kG * kU + kG * u;
            jn[u] = kn < len ? __ldg(my_ids + kn) : -1;
        }
#pragma unroll
        for (int u = 0; u < kU; u++) {
            if (jj[u] < 0) continue;
            const float4 a = ra4[u], b = rb4[u];
            const float Xj[3] = {a.x, a.y, a.z}, xj[3] = {b.x, b.y, b.z};
            const float nvol = a.w;
            float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < ND; d++) p[d] = __fsub_rn(Xi[d], Xj[d]);
            float d2 = dist2<ND>(p[0], p[1], p[2]);
            d2 = maybe_periodic_fix<ND, PER>(pp, d2, p[0], p[1], p[2]);
            float sg;
            if (EXACT) {
                const float dist = __fsqrt_rn(d2);
                if (dist < PNB_SQRT_EPS_F32) continue;
                const float q = __fdiv_rn(dist, h);
                float w = 0.f;
                if (q < 2.f) {
                    const float t = __fsub_rn(1.f, __fmul_rn(q, 0.5f));
                    w = __fmul_rn(__fmul_rn(-5.f, q), __fmul_rn(__fmul_rn(t, t), t));
                }
                sg = __fdiv_rn(__fmul_rn(nh, w), dist);
            } else {
                if (d2 < PNB_SQRT_EPS_F32 * PNB_SQRT_EPS_F32) continue;
                const float dist = d2 * fast_rsqrt(d2);
                const float t = fmaxf(fmaf(nhalf_inv_h, dist, 1.f), 0.f);
                sg = (k5 * t) * (t * t);
            }
            float grad[3] = {0.f, 0.f, 0.f}, lg[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int d = 0; d < ND; d++) grad[d] = __fmul_rn(sg, p[d]);
#pragma unroll
            for (int a2 = 0; a2 < ND; a2++) {
                if (EXACT) {
                    float t = __fmul_rn(Li[a2], grad[0]);
#pragma unroll
                    for (int b2 = 1; b2 < ND; b2++) t = __fadd_rn(t, __fmul_rn(Li[b2 * ND + a2], grad[b2]));
                    lg[a2] = t;
                } else {
                    float t = Li[a2] * grad[0];
#pragma unroll
                    for (int b2 = 1; b2 < ND; b2++) t = fmaf(Li[b2 * ND + a2], grad[b2], t);
                    lg[a2] = t;
                }
            }
#pragma unroll
            for (int a2 = 0; a2 < ND; a2++) {
                const float cd = __fsub_rn(xi[a2], xj[a2]);
                const float nc = __fmul_rn(nvol, cd);
#pragma unroll
                for (int b2 = 0; b2 < ND; b2++) {
                    if (EXACT) acc[b2 * ND + a2] = __fadd_rn(acc[b2 * ND + a2], __fmul_rn(nc, lg[b2]));
                    else acc[b2 * ND + a2] = fmaf(nc, lg[b2], acc[b2 * ND + a2]);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < NN; e++) {
        float v = acc[e];
#pragma unroll
        for (int o = kG / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        acc[e] = v;
    }
    if (have && sub == 0) {
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
    cached_free(0, l->offsets, l->bytes_offsets);
    cached_free(1, l->ids, l->bytes_ids);
    cached_free(2, l->counts, l->bytes_counts);
    cached_free(3, l->pack, l->bytes_pack);
    cached_free(4, l->pack8, l->bytes_pack8);
    cudaFree(l->d_err);
    if (l->h_err) cudaFreeHost(l->h_err);
    cudaGetLastError();
    delete l;
}

extern "C" int64_t pnb_nlist_n_points(const pnb_nlist *l) { return l ? l->nx : 0; }
extern "C" int64_t pnb_nlist_n_pairs(const pnb_nlist *l) { return l ? l->n_pairs : 0; }
extern "C" int64_t pnb_nlist_max_length(const pnb_nlist *l) { return l ? l->max_len : 0; }

// capacity of the rows of the next one-pass build: the longest list + 12.5 % + 8, multiple of 4
static int nlist_cap_for(unsigned int longest)
{
    const int64_t c = ((int64_t)longest + longest / 8 + 8 + 3) / 4 * 4;
    return c > 4096 ? 0 : (int)c;       // very long lists: keep the two-pass build
}
// PNB_NLIST_ONE_PASS=0: always count + fill (two test passes)
static int g_nlist_one_pass = getenv("PNB_NLIST_ONE_PASS") ? atoi(getenv("PNB_NLIST_ONE_PASS")) : 1;

extern "C" pnb_status pnb_nlist_build_f32(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                          int64_t n, int sort, pnb_nlist **out, void *stream)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (g->f64) { set_error("Float64 grid handle passed to a Float32 entry point"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    // nhs_precomputed.jl:136-138: the precomputed search rejects inactive points
    if (!g->template_search && (!g->full_build || n != g->n_y_built)) {
        set_error("this neighborhood search does not support inactive points");
        return PNB_ERR_ARG;
    }
    { pnb_status sp = resolve_pending(g); if (sp != PNB_OK) return sp; }
    { pnb_status sy = check_built_y(g, y, n, (cudaStream_t)stream); if (sy != PNB_OK) return sy; }
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
    NL_CUDA(cached_malloc(0, (void **)&l->offsets, sizeof(int64_t) * (size_t)(nx + 1), &l->bytes_offsets));
    NL_CUDA(cached_malloc(2, (void **)&l->counts, sizeof(uint32_t) * (size_t)(nx + 8), &l->bytes_counts));
    NL_CUDA(cudaMalloc(&l->d_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->d_err, 0, sizeof(int), s));
    NL_CUDA(cudaMallocHost(&l->h_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->counts, 0, sizeof(uint32_t) * (size_t)(nx + 8), s));
    const bool fast = g->full_build && !g->y_refreshed && x == g->y_built && nx == g->n_y_built;
    pnb_status st = PNB_OK;
    if (!fast && g->bucket_valid) {
        // two-set and per-point list builds walk the CSR arrays
        st = ensure_csr(g, s);
        if (st != PNB_OK) return fail(st);
        g->bucket_valid = false;
    }
    if (!sort) {
        // unsorted lists keep the visiting order: make it the reproducible one (ids ascending
        // inside every cell); sorted lists do not depend on it
        st = ensure_canonical(g, s);
        if (st != PNB_OK) return fail(st);
    }
    // sorted lists do not depend on the visiting order: the tile kernel builds them; unsorted
    // lists keep the reference's order and use the ordered kernel
    const bool tiles = sort != 0;
    // ---- one test pass (repeated builds of sorted x === y lists): rows of fixed capacity taken
    // from the longest list of the previous build, then sort + compaction into the CSR list ----
    if (g_nlist_one_pass && fast && tiles && g->nl_cap_hint > 0 && nx > 0 && !g->hashed &&
        (int64_t)nx * g->nl_cap_hint < (int64_t)1 << 34) {
        const int cap = g->nl_cap_hint;
        int32_t *rows = nullptr;
        size_t bytes_rows = 0;
        // the rows are nx * cap ids (many gigabytes for large clouds): when that allocation fails
        // the much smaller two-pass build below is still possible
        if (cached_malloc(5, (void **)&rows, sizeof(int32_t) * (size_t)nx * cap, &bytes_rows) != cudaSuccess) {
            cudaGetLastError();
            rows = nullptr;
            g->nl_cap_hint = 0;
            goto two_pass;
        }
      {
        auto drop_rows = [&]() { cached_free(5, rows, bytes_rows); };
        st = launch_sweep(g, fast, tiles, x, nx, nullptr, 0, ListFillRowsCl{rows, l->counts, cap, l->d_err}, s);
        if (st == PNB_OK) st = exclusive_scan_u32_to_i64(g, l->counts, l->offsets, nx, s);
        if (st != PNB_OK) { drop_rows(); return fail(st); }
        int64_t total1 = 0;
        unsigned int longest1 = 0;
        int err1 = 0;
        k_max_count<<<(unsigned)div_up(nx, 256), 256, 0, s>>>(nx, l->counts, l->counts + nx);
        g_launch_count++;
        cudaError_t e1 = cudaMemcpyAsync(&total1, l->offsets + nx, sizeof(int64_t), cudaMemcpyDeviceToHost, s);
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(&longest1, l->counts + nx, sizeof(unsigned int), cudaMemcpyDeviceToHost, s);
        if (e1 == cudaSuccess) e1 = cudaMemcpyAsync(&err1, l->d_err, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e1 == cudaSuccess) e1 = cudaStreamSynchronize(s);
        if (e1 != cudaSuccess) { drop_rows(); return fail(cuda_fail(e1, "one-pass list build")); }
        if ((err1 & 8) == 0) {
            l->n_pairs = total1;
            l->max_len = longest1;
            e1 = cached_malloc(1, (void **)&l->ids, sizeof(int32_t) * (size_t)(total1 > 0 ? total1 : 1), &l->bytes_ids);
            if (e1 != cudaSuccess) { drop_rows(); return fail(cuda_fail(e1, "cudaMalloc ids")); }
            {
                ProfScope ps(PH_NLIST_SORT, s);
                k_sort_compact_rows<<<(unsigned)div_up(nx, kSortWarps), kSortWarps * 32, 0, s>>>(
                    nx, cap, rows, l->counts, l->offsets, l->ids, 1);
                g_launch_count++;
            }
            st = check_err_word(g, s);
            drop_rows();
            if (st != PNB_OK) return fail(st);
            g->nl_cap_hint = nlist_cap_for(longest1);
            *out = l;
            return PNB_OK;
        }
        // a list outgrew the capacity: forget the hint, build in two passes below
        drop_rows();
        g->nl_cap_hint = 0;
        NL_CUDA(cudaMemsetAsync(l->d_err, 0, sizeof(int), s));
        NL_CUDA(cudaMemsetAsync(l->counts, 0, sizeof(uint32_t) * (size_t)(nx + 8), s));
      }
    }
two_pass:
    st = launch_sweep(g, fast, tiles, x, nx, nullptr, 0, ListCountCl{l->counts}, s);
    if (st != PNB_OK) return fail(st);
    st = exclusive_scan_u32_to_i64(g, l->counts, l->offsets, nx, s);
    if (st != PNB_OK) return fail(st);
    int64_t total = 0;
    unsigned int longest = 0;
    if (nx > 0) {
        // counts[nx .. nx+7] is zeroed padding: word nx holds the maximum
        k_max_count<<<(unsigned)div_up(nx, 256), 256, 0, s>>>(nx, l->counts, l->counts + nx);
        g_launch_count++;
    }
    NL_CUDA(cudaMemcpyAsync(&total, l->offsets + nx, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    NL_CUDA(cudaMemcpyAsync(&longest, l->counts + nx, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    NL_CUDA(cudaStreamSynchronize(s));
    l->n_pairs = total;
    l->max_len = longest;
    NL_CUDA(cached_malloc(1, (void **)&l->ids, sizeof(int32_t) * (size_t)(total > 0 ? total : 1), &l->bytes_ids));
    st = launch_sweep(g, fast, tiles, x, nx, nullptr, 0, ListFillCl{l->offsets, l->ids}, s);
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
    if (fast && tiles) g->nl_cap_hint = nlist_cap_for(longest);   // the next build can fill in one pass
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

extern "C" pnb_status pnb_tlsph_deformation_grad_f32(const pnb_nlist *l_, const pnb_grid *g,
                                                     const float *X0, const float *xcur,
                                                     const float *mass, const float *rho0,
                                                     const float *L, float smoothing_length,
                                                     float kernel_norm, float *F, void *stream)
{
    if (!l_ || !g) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    pnb_nlist *l = const_cast<pnb_nlist *>(l_);
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        if (!l->pack)
            PNB_CUDA(cached_malloc(3, (void **)&l->pack, sizeof(float4) * 2 * (size_t)l->nx, &l->bytes_pack));
        const unsigned blocks = (unsigned)div_up(l->nx * 8, 256);   // 8 lanes per point
        const unsigned pblocks = (unsigned)div_up(l->nx, 256);
        const bool per = g->p.periodic != 0;
        const bool exact = pnb_get_exact_arithmetic() != 0;
        ProfScope ps(PH_NLIST_SWEEP, s);
#define DEFGRAD(ND)                                                                                 \
    do {                                                                                            \
        k_pack_tlsph<ND><<<pblocks, 256, 0, s>>>(l->nx, X0, xcur, mass, rho0, l->pack);             \
        if (per && exact) k_tlsph_defgrad<ND, true, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack, L, smoothing_length, kernel_norm, F); \
        else if (per) k_tlsph_defgrad<ND, true, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack, L, smoothing_length, kernel_norm, F); \
        else if (exact) k_tlsph_defgrad<ND, false, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack, L, smoothing_length, kernel_norm, F); \
        else k_tlsph_defgrad<ND, false, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack, L, smoothing_length, kernel_norm, F); \
    } while (0)
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

// =============================================================================================
// Float64 neighbour lists (f64.cuh): count pass, scan, fill pass in the reference's visiting
// order (the Float64 build leaves every cell in canonical order), optional per-list sort.
// The list handle and its exports are the same as for Float32.
// =============================================================================================
namespace pnb {
template <int ND>
__global__ void k_nlist_pairs64(GridP64 g, int64_t nx, const int64_t *__restrict__ offsets,
                                const int32_t *__restrict__ ids, const double *__restrict__ x,
                                const double *__restrict__ y, double *__restrict__ pos_diff,
                                double *__restrict__ dist)
{
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= nx) return;
    double xi[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) xi[d] = x[i * ND + d];
    for (int64_t k = offsets[i] + lane_id(); k < offsets[i + 1]; k += 32) {
        const int64_t j = ids[k];
        Rec64 yj;
        yj.x = y[j * ND];
        yj.y = ND > 1 ? y[j * ND + 1] : 0.0;
        yj.z = ND > 2 ? y[j * ND + 2] : 0.0;
        yj.id = j;
        double p[3];
        // nhs_precomputed.jl:230-238: the periodic fix applies when d2 > r2, there is no radius test
        const double d2 = pair_d2_64<ND>(g, xi, yj, p, true);
        if (pos_diff) {
#pragma unroll
            for (int d = 0; d < ND; d++) pos_diff[k * ND + d] = p[d];
        }
        // mixed precision: distance = sqrt of the Float32 d2, in Float32
        if (dist) dist[k] = g.mixed ? (double)__fsqrt_rn((float)d2) : __dsqrt_rn(d2);
    }
}
}  // namespace pnb

extern "C" pnb_status pnb_nlist_build_f64(pnb_grid *g, const double *x, int64_t nx, const double *y,
                                          int64_t n, int sort, pnb_nlist **out, void *stream)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (!g || !g->f64) { set_error("not a Float64 grid handle"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    if (!g->template_search && (!g->full_build || n != g->n_y_built)) {
        set_error("this neighborhood search does not support inactive points");
        return PNB_ERR_ARG;
    }
    { pnb_status sp = resolve_pending(g); if (sp != PNB_OK) return sp; }
    { pnb_status sy = check_built_y(g, y, n, (cudaStream_t)stream); if (sy != PNB_OK) return sy; }
    cudaStream_t s = (cudaStream_t)stream;
    pnb_nlist *l = new pnb_nlist();
    memset(l, 0, sizeof(*l));
    l->nx = nx;
    l->ndims = g->p64.ndims;
    auto fail = [&](pnb_status st) { pnb_nlist_destroy(l); return st; };
#define NL_CUDA(expr)                                                           \
    do {                                                                        \
        cudaError_t e__ = (expr);                                               \
        if (e__ != cudaSuccess) return fail(cuda_fail(e__, #expr));             \
    } while (0)
    NL_CUDA(cached_malloc(0, (void **)&l->offsets, sizeof(int64_t) * (size_t)(nx + 1), &l->bytes_offsets));
    NL_CUDA(cached_malloc(2, (void **)&l->counts, sizeof(uint32_t) * (size_t)(nx + 8), &l->bytes_counts));
    NL_CUDA(cudaMalloc(&l->d_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->d_err, 0, sizeof(int), s));
    NL_CUDA(cudaMallocHost(&l->h_err, sizeof(int)));
    NL_CUDA(cudaMemsetAsync(l->counts, 0, sizeof(uint32_t) * (size_t)(nx + 8), s));
    const bool live = !g->template_search && g->n_built > 0 && nx > 0;
    const unsigned blocks = (unsigned)div_up(nx > 0 ? nx : 1, 128);
#define PNB_SWEEP64(MODE, ...)                                                                    \
    switch (g->p64.ndims) {                                                                       \
        case 1: k_sweep_points64<1, MODE><<<blocks, 128, 0, s>>>(__VA_ARGS__); break;             \
        case 2: k_sweep_points64<2, MODE><<<blocks, 128, 0, s>>>(__VA_ARGS__); break;             \
        default: k_sweep_points64<3, MODE><<<blocks, 128, 0, s>>>(__VA_ARGS__); break;            \
    }
    if (live) {
        PNB_SWEEP64(1, g->p64, g->cell_start, g->sorted64, x, nx, nullptr, 0, nullptr, l->counts,
                    nullptr, nullptr, g->d_err);
        g_launch_count++;
    }
    pnb_status st = exclusive_scan_u32_to_i64(g, l->counts, l->offsets, nx, s);
    if (st != PNB_OK) return fail(st);
    int64_t total = 0;
    unsigned int longest = 0;
    if (nx > 0) {
        k_max_count<<<(unsigned)div_up(nx, 256), 256, 0, s>>>(nx, l->counts, l->counts + nx);
        g_launch_count++;
    }
    NL_CUDA(cudaMemcpyAsync(&total, l->offsets + nx, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    NL_CUDA(cudaMemcpyAsync(&longest, l->counts + nx, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    NL_CUDA(cudaStreamSynchronize(s));
    l->n_pairs = total;
    l->max_len = longest;
    NL_CUDA(cached_malloc(1, (void **)&l->ids, sizeof(int32_t) * (size_t)(total > 0 ? total : 1), &l->bytes_ids));
    if (live) {
        PNB_SWEEP64(2, g->p64, g->cell_start, g->sorted64, x, nx, nullptr, 0, nullptr, nullptr,
                    l->offsets, l->ids, g->d_err);
        g_launch_count++;
    }
#undef PNB_SWEEP64
    if (sort && nx > 0 && total > 0) {
        k_sort_lists<<<(unsigned)div_up(nx, kSortWarps), kSortWarps * 32, 0, s>>>(nx, l->offsets, l->ids);
        g_launch_count++;
    }
    st = check_err_word(g, s);
    if (st != PNB_OK) return fail(st);
#undef NL_CUDA
    *out = l;
    return PNB_OK;
}

extern "C" pnb_status pnb_nlist_pairs_f64(const pnb_nlist *l, const pnb_grid *g, const double *x,
                                          const double *y, double *pos_diff, double *distance,
                                          void *stream)
{
    if (!l || !g || !g->f64) { set_error("not a Float64 grid / list handle"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        const unsigned blocks = (unsigned)div_up(l->nx * 32, 256);
        switch (l->ndims) {
            case 1: k_nlist_pairs64<1><<<blocks, 256, 0, s>>>(g->p64, l->nx, l->offsets, l->ids, x, y, pos_diff, distance); break;
            case 2: k_nlist_pairs64<2><<<blocks, 256, 0, s>>>(g->p64, l->nx, l->offsets, l->ids, x, y, pos_diff, distance); break;
            default: k_nlist_pairs64<3><<<blocks, 256, 0, s>>>(g->p64, l->nx, l->offsets, l->ids, x, y, pos_diff, distance); break;
        }
        g_launch_count++;
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

// ---------------------------------------------------------------------------------------------
// TLSPH forces over the lists + the pointwise steps either side (tlsph_interact.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" pnb_status pnb_tlsph_interact_f32(const pnb_nlist *l_, const pnb_grid *g,
                                             const float *X0, const float *xcur,
                                             const float *mass, const float *rho0,
                                             const float *pk1_corrected, const float *F,
                                             const pnb_tlsph_params *params, float *dv,
                                             void *stream)
{
    if (!l_ || !g) { set_error("handle is NULL"); return PNB_ERR_ARG; }
    if (!params) { set_error("params is NULL"); return PNB_ERR_ARG; }
    pnb_nlist *l = const_cast<pnb_nlist *>(l_);
    cudaStream_t s = (cudaStream_t)stream;
    if (l->nx > 0) {
        if (!l->pack8)
            PNB_CUDA(cached_malloc(4, (void **)&l->pack8, sizeof(float4) * kTlRecF4 * (size_t)l->nx,
                                   &l->bytes_pack8));
        const unsigned blocks = (unsigned)div_up(l->nx * 8, 256);   // 8 lanes per point
        const unsigned pblocks = (unsigned)div_up(l->nx, 128);
        const bool per = g->p.periodic != 0;
        const bool exact = pnb_get_exact_arithmetic() != 0;
        const float h = params->smoothing_length, kn = params->kernel_norm;
        const float E = params->young_modulus, al = params->penalty_alpha;
        ProfScope ps(PH_NLIST_SWEEP, s);
#define TLFORCE(ND)                                                                                 \
    do {                                                                                            \
        k_pack_tlsph_full<ND><<<pblocks, 128, 0, s>>>(l->nx, X0, xcur, mass, rho0, pk1_corrected, F, l->pack8); \
        if (per && exact) k_tlsph_interact<ND, true, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack8, h, kn, E, al, dv); \
        else if (per) k_tlsph_interact<ND, true, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack8, h, kn, E, al, dv); \
        else if (exact) k_tlsph_interact<ND, false, true><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack8, h, kn, E, al, dv); \
        else k_tlsph_interact<ND, false, false><<<blocks, 256, 0, s>>>(g->p, l->nx, l->offsets, l->ids, l->pack8, h, kn, E, al, dv); \
    } while (0)
        switch (l->ndims) {
            case 1: TLFORCE(1); break;
            case 2: TLFORCE(2); break;
            default: TLFORCE(3); break;
        }
#undef TLFORCE
        g_launch_count++;
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

extern "C" pnb_status pnb_tlsph_pk1_corrected_f32(int ndims, int64_t n, const float *F,
                                                  const float *L, float young_modulus,
                                                  float poisson_ratio, float *pk1_corrected,
                                                  void *stream)
{
    if (ndims < 1 || ndims > 3) { set_error("`NDIMS` must be 1, 2, or 3"); return PNB_ERR_ARG; }
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // Lame parameters in Float32, one rounding per operation
    volatile float a = young_modulus * poisson_ratio;
    volatile float b1 = 1.0f + poisson_ratio;
    volatile float two_nu = 2.0f * poisson_ratio;
    volatile float b2 = 1.0f - two_nu;
    volatile float den = b1 * b2;
    volatile float lambda = a / den;
    volatile float den2 = 2.0f * b1;
    volatile float mu = young_modulus / den2;
    if (n > 0) {
        const unsigned blocks = (unsigned)div_up(n, 128);
        switch (ndims) {
            case 1: k_tlsph_pk1_corrected<1><<<blocks, 128, 0, s>>>(n, F, L, lambda, mu, pk1_corrected); break;
            case 2: k_tlsph_pk1_corrected<2><<<blocks, 128, 0, s>>>(n, F, L, lambda, mu, pk1_corrected); break;
            default: k_tlsph_pk1_corrected<3><<<blocks, 128, 0, s>>>(n, F, L, lambda, mu, pk1_corrected); break;
        }
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

extern "C" pnb_status pnb_wcsph_compute_pressure_f32(int ndims, int64_t n, const float *v,
                                                     float sound_speed, float reference_density,
                                                     float exponent, float background_pressure,
                                                     float *pressure, void *stream)
{
    if (ndims < 1 || ndims > 3) { set_error("`NDIMS` must be 1, 2, or 3"); return PNB_ERR_ARG; }
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    cudaStream_t s = (cudaStream_t)stream;
    volatile float c2 = sound_speed * sound_speed;
    volatile float rc2 = reference_density * c2;
    volatile float B = rc2 / exponent;
    if (n > 0) {
        k_wcsph_compute_pressure<<<(unsigned)div_up(n, 256), 256, 0, s>>>(
            n, ndims + 1, v, B, reference_density, exponent, background_pressure, pressure);
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}
