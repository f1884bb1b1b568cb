// common.cuh -- shared device/host definitions of libpnb200 (sm_100a only).
//
// Arithmetic contract: everything that decides a cell or a neighbour set
// is written with explicit round-to-nearest intrinsics (__fsub_rn, __fmul_rn, __fadd_rn,
// __fdiv_rn, __fsqrt_rn).  These are never contracted into FMAs and never replaced by
// approximate sequences, whatever flags the translation unit is compiled with.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pnb200.h"

namespace pnb {

// Plain-old-data copy of GridNeighborhoodSearch + FullGridCellList (+ PeriodicBox) scalars,
// passed to kernels by value.  (reference: src/nhs_grid.jl:67-75, src/cell_lists/full_grid.jl:30-35)
struct GridP {
    int ndims;
    int periodic;
    float r;           // search_radius
    float r2;          // r * r, rounded once (src/nhs_grid.jl:555)
    float minc[3];     // padded min corner
    float cs[3];       // cell_size
    int gs[3];         // allocated grid size (1 for unused dims)
    int nc[3];         // periodic cells per dim (-1 if not periodic)
    int off[3];        // window offset: local cell = global cell - off (0 for a full grid)
    float bsize[3];    // periodic box size
    float wrap_d2;     // d2 below this can never be changed by the periodic fix (see periodic_fix)
    int total_cells;   // cells of the grid; entries of the hash table (list_size) when hashed
    int hashed;        // SpatialHashingCellList: cell -> key = spatial_hash(cell, total_cells)
};

// Float64 searches (f64.cuh): the same scalars in double, and the cell-ordered record
struct GridP64 {
    int ndims;
    int periodic;
    double r, r2;
    double minc[3];     // padded min corner
    double cs[3];       // cell_size
    int gs[3];          // allocated grid size
    int nc[3];          // periodic cells per dim (-1 if not periodic)
    double bsize[3];    // periodic box size
    int total_cells;
    int mixed;          // Float64 coordinates, Float32 radius: r, r2, bsize hold Float32 values and
                        // the pair arithmetic after pos_diff = Float32.(x_i - y_j) is Float32
};

struct alignas(32) Rec64 {
    double x, y, z;
    long long id;
};

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
void set_error(const char *fmt, ...);
pnb_status cuda_fail(cudaError_t e, const char *what);
extern int64_t g_launch_count;

#define PNB_CUDA(expr)                                                   \
    do {                                                                 \
        cudaError_t e__ = (expr);                                        \
        if (e__ != cudaSuccess) return ::pnb::cuda_fail(e__, #expr);     \
    } while (0)

#define PNB_LAUNCHED()                                                   \
    do {                                                                 \
        ::pnb::g_launch_count++;                                         \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return ::pnb::cuda_fail(e__, "kernel launch"); \
    } while (0)

static inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// optional per-kernel timing with CUDA events on the launching stream (pnb_profile_*), used by
// bench.py for the roofline numbers.  Off by default: no events are recorded then.
// ---------------------------------------------------------------------------------------------
enum Phase {
    PH_BUILD_CELL_COUNT = 0,  // k_cell_hist
    PH_BUILD_SCAN,            // k_scan_lookback
    PH_BUILD_SCATTER,         // k_scatter_points
    PH_BUILD_FINALIZE,        // k_canonicalize (on demand, not part of update!)
    PH_GATHER,                // k_gather_* (payload into cell order)
    PH_SWEEP_CELLS,           // k_sweep_cells / k_sweep_tiles
    PH_SWEEP_OVERFLOW,        // k_sweep_overflow
    PH_SWEEP_POINTS,          // k_sweep_points
    PH_NLIST_SORT,            // k_sort_lists
    PH_NLIST_SWEEP,           // k_tlsph_defgrad / k_nlist_pairs
    PH_EXPORT,                // export kernels
    PH_BUILD_BUCKET,          // k_bucket_scatter (one-pass update!)
    PH_SWEEP_TILES_PREP,      // k_flat_tiles / k_flat_scan (tile table of k_sweep_flat)
    PH_COUNT_
};
struct ProfScope {
    int phase;
    cudaStream_t stream;
    void *slot;
    ProfScope(int phase, cudaStream_t s);
    ~ProfScope();
};

// ---------------------------------------------------------------------------------------------
// device arithmetic
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Julia mod(a, n), n > 0
__device__ __forceinline__ int floormod_i(int a, int n)
{
    int m = a % n;
    return m < 0 ? m + n : m;
}

// Slow path of floor_to_int (src/util.jl:19-34) + `.+ 1` + periodic wrap with Julia's wrapping
// Int64 arithmetic, for |floor| >= 2^30 or NaN.  Non-periodic: such a cell is always outside
// the grid, any out-of-range marker will do.
static __device__ __noinline__ int cell_coord_slow(float f, int periodic, int nc)
{
    if (!periodic) return f < 0.f ? -1 : 0x7fffffff;  // NaN lands here too -> out of bounds (high)
    long long c;
    if (isnan(f) || f >= 9223372036854775808.0f) c = 0x7fffffffffffffffLL;
    else if (f <= -9223372036854775808.0f) c = (long long)0x8000000000000000ULL;
    else c = (long long)f;
    unsigned long long u = (unsigned long long)c + 1ULL;   // .+ 1 (wraps)
    u = u - 2ULL;                                          // cell .- 2 (wraps)
    long long a = (long long)u;
    long long m = a % (long long)nc;
    if (m < 0) m += nc;
    return (int)m + 2;
}

// cell_coords (src/nhs_grid.jl:622-628) for one dimension:
//   floor_to_int((x - min_corner) / cell_size) + 1   (src/cell_lists/full_grid.jl:93)
//   periodic: mod(c - 2, n_cells) + 2                (src/nhs_grid.jl:619)
__device__ __forceinline__ int cell_coord(float x, float minc, float cs, int periodic, int nc,
                                          int off = 0)
{
    float q = __fdiv_rn(__fsub_rn(x, minc), cs);
    float f = floorf(q);
    if (!(fabsf(f) < 1073741824.0f)) return cell_coord_slow(f, periodic, nc);
    int c = (int)f + 1;
    if (periodic) c = floormod_i(c - 2, nc) + 2;
    return c - off;   // windowed grids (slabs) count cells from their own first layer
}

// Linear 0-based cell index (src/cell_lists/full_grid.jl:157-161), or -1 when the cell is not
// in 2:(size-1) in some dimension (check_cell_bounds, :205-213).
template <int ND>
__device__ __forceinline__ int point_cell(const GridP &g, const float *p, int *cc)
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        cc[d] = cell_coord(p[d], g.minc[d], g.cs[d], g.periodic, g.nc[d], g.off[d]);
        ok = ok && (cc[d] >= 2) && (cc[d] <= g.gs[d] - 1);
    }
#pragma unroll
    for (int d = ND; d < 3; d++) cc[d] = 1;
    if (!ok) return -1;
    return (cc[0] - 1) + (cc[1] - 1) * g.gs[0] + (cc[2] - 1) * g.gs[0] * g.gs[1];
}

__device__ __forceinline__ int linear_cell(const GridP &g, int c0, int c1, int c2)
{
    return (c0 - 1) + (c1 - 1) * g.gs[0] + (c2 - 1) * g.gs[0] * g.gs[1];
}

// pos_diff = x_i - y_j ; d2 = dot(pos_diff, pos_diff) left to right (src/nhs_grid.jl:547-549)
template <int ND>
__device__ __forceinline__ float dist2(float px, float py, float pz)
{
    float d2 = __fmul_rn(px, px);
    if (ND > 1) d2 = __fadd_rn(d2, __fmul_rn(py, py));
    if (ND > 2) d2 = __fadd_rn(d2, __fmul_rn(pz, pz));
    return d2;
}

// The scalars the distance test needs, kept in registers (passing GridP by reference into a
// non-inlined function would force the kernel parameters through local memory).
struct PerP {
    float r2;        // search_radius^2
    float wrap_d2;   // see maybe_periodic_fix
    float bs[3];     // periodic box size
};
__device__ __forceinline__ PerP make_perp(const GridP &g)
{
    PerP q;
    q.r2 = g.r2; q.wrap_d2 = g.wrap_d2;
    q.bs[0] = g.bsize[0]; q.bs[1] = g.bsize[1]; q.bs[2] = g.bsize[2];
    return q;
}

// compute_periodic_distance (src/neighborhood_search.jl:428-435), applied by the caller only
// when d2 > r2:   pos_diff -= size .* round.(pos_diff ./ size);  d2 = dot(pos_diff, pos_diff)
// Julia's round is ties-to-even == rintf.
template <int ND>
__device__ __noinline__ float periodic_fix(float bx, float by, float bz, float &px, float &py,
                                           float &pz)
{
    px = __fsub_rn(px, __fmul_rn(bx, rintf(__fdiv_rn(px, bx))));
    if (ND > 1) py = __fsub_rn(py, __fmul_rn(by, rintf(__fdiv_rn(py, by))));
    if (ND > 2) pz = __fsub_rn(pz, __fmul_rn(bz, rintf(__fdiv_rn(pz, bz))));
    return dist2<ND>(px, py, pz);
}

// The fix is the identity whenever every |pos_diff[d]| / size[d] rounds to 0, which is certain
// when d2 < (0.49 * min size)^2 =: wrap_d2 (and wrap_d2 > r2 because every periodic box has at
// least 3 cells of size >= r).  So `d2 >= wrap_d2` selects exactly the candidates that need the
// exact slow path; for all others the reference's recomputation reproduces the same bits.
template <int ND, bool PER>
__device__ __forceinline__ float maybe_periodic_fix(const PerP &q, float d2, float &px, float &py,
                                                    float &pz)
{
    if (PER && d2 >= q.wrap_d2) {
        if (d2 > q.r2) d2 = periodic_fix<ND>(q.bs[0], q.bs[1], q.bs[2], px, py, pz);
    }
    return d2;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// asynchronous global -> shared copies (LDGSTS): no register staging, all copies of a tile in flight
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t saddr, const void *gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(saddr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t saddr, const void *gptr)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory");
}
// 1-D bulk copies global -> shared by the TMA unit (cp.async.bulk, UBLKCP in the SASS), completion
// through an mbarrier's transaction count.  dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mbar_init(uint32_t mbar_sa, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_sa), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar_sa, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_sa, const void *src, uint32_t bytes, uint32_t mbar_sa)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_sa), "l"(src), "r"(bytes), "r"(mbar_sa) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar_sa, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(mbar_sa), "r"(parity) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}


// The cell list as the kernels see it.  Two layouts:
//   CSR     (K == 0): start[0 .. C] offsets, cell c = rec[start[c] .. start[c+1])
//   buckets (K  > 0): start[c] = number of points of cell c, cell c = rec[c K .. c K + start[c])
//           (what the one-pass update! writes: every cell owns K record slots)
//           The buckets may be numbered in TRANSPOSED cell order (last dimension fastest, t0 > 0 =
//           the grid sizes) when that is the order the points arrive in: the one-pass build then
//           writes neighbouring buckets from neighbouring points (DRAM row locality).
struct CellsView {
    const uint32_t *start;
    const float4 *rec;
    uint32_t K;
    uint32_t t0, t1, t2;   // grid sizes when the buckets are numbered transposed, else t0 == 0
};
// bucket number of the linear cell index lin (identity unless the numbering is transposed)
__host__ __device__ __forceinline__ uint32_t bucket_of(uint32_t lin, uint32_t t0, uint32_t t1, uint32_t t2)
{
    if (t0 == 0u) return lin;
    const uint32_t c0 = lin % t0, r = lin / t0;
    const uint32_t c1 = r % t1, c2 = r / t1;
    return c2 + t2 * (c1 + t1 * c0);
}
__device__ __forceinline__ void cell_range(const CellsView &v, int lin, uint32_t &b0, uint32_t &cnt)
{
    if (v.K) {
        const uint32_t lb = bucket_of((uint32_t)lin, v.t0, v.t1, v.t2);
        b0 = lb * v.K;
        // (a count above K only exists while an overflowed stream-ordered update! is waiting for
        // its blocking rebuild: stay inside the bucket, the results are discarded then)
        cnt = min(v.start[lb], v.K);
    } else { b0 = v.start[lin]; cnt = v.start[lin + 1] - b0; }
}

#endif  // __CUDACC__

}  // namespace pnb
