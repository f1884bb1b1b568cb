// grid.cu -- constructors (host arithmetic), the device counting sort behind initialize!/update!,
// the decoupled-lookback scan, and the cell-list exports.
//
// reference: src/cell_lists/full_grid.jl (FullGridCellList), src/nhs_grid.jl:77-129, 220-292,
// 470-477 (constructor, initialize!, update!), src/vector_of_vectors.jl (the layout exported by
// pnb_grid_export_dvov).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "grid.cuh"

namespace pnb {

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char t_err[512] = "";
int64_t g_launch_count = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

pnb_status cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error: %s (%s)", cudaGetErrorString(e), what);
    return PNB_ERR_CUDA;
}

// ---- profiling ----------------------------------------------------------------------------
struct ProfRec { int phase; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static double g_prof_ms[PH_COUNT_];
static int64_t g_prof_n[PH_COUNT_];

static cudaEvent_t prof_event()
{
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(int phase_, cudaStream_t s) : phase(phase_), stream(s), slot(nullptr)
{
    if (!g_prof_on) return;
    ProfRec *r = new ProfRec{phase, prof_event(), prof_event()};
    cudaEventRecord(r->a, stream);
    slot = r;
}
ProfScope::~ProfScope()
{
    if (!slot) return;
    ProfRec *r = static_cast<ProfRec *>(slot);
    cudaEventRecord(r->b, stream);
    g_prof_recs.push_back(*r);
    delete r;
}

static void prof_collect()
{
    for (ProfRec &r : g_prof_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            g_prof_ms[r.phase] += ms;
            g_prof_n[r.phase] += 1;
        }
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    cudaGetLastError();
}

static const char *kPhaseNames[PH_COUNT_] = {
    "k_cell_count", "k_scan_lookback", "k_scatter", "k_finalize_cells", "k_gather",
    "k_sweep_cells", "k_sweep_overflow", "k_sweep_points", "k_sort_lists", "k_nlist_sweep", "k_export"};

static const char *kDomainMsg =
    "particle coordinates are NaN or outside the domain bounds of the cell list";
static const char *kListFullMsg = "cell list is full. Use a larger `max_points_per_cell`.";
static const char *kBoundsMsg =
    "BoundsError: a neighboring cell of a query point is outside the cell grid";

pnb_status check_err_word(pnb_grid *g, cudaStream_t s)
{
    PNB_CUDA(cudaMemcpyAsync(g->h_err, g->d_err, sizeof(int), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    int e = *g->h_err;
    if (e == 0) return PNB_OK;
    PNB_CUDA(cudaMemsetAsync(g->d_err, 0, sizeof(int), s));
    if (e & 1) { set_error("%s", kDomainMsg); return PNB_ERR_DOMAIN; }
    if (e & 4) { set_error("%s", kListFullMsg); return PNB_ERR_LIST_FULL; }
    set_error("%s", kBoundsMsg);
    return PNB_ERR_BOUNDS;
}

pnb_status ensure_scratch(pnb_grid *g, int64_t bytes)
{
    if (bytes <= g->scratch_bytes) return PNB_OK;
    if (g->scratch) PNB_CUDA(cudaFree(g->scratch));
    g->scratch = nullptr;
    g->scratch_bytes = 0;
    int64_t want = bytes + bytes / 8;
    PNB_CUDA(cudaMalloc(&g->scratch, (size_t)want));
    g->scratch_bytes = want;
    return PNB_OK;
}

// ---------------------------------------------------------------------------------------------
// host arithmetic of the constructors, evaluated exactly like Julia does.
// Compiled by the host compiler without -ffast-math; every operation below is a single
// IEEE-754 float (or double) operation.
// ---------------------------------------------------------------------------------------------
static pnb_status grid_params_host(int ndims, float r, const float *min_corner,
                                   const float *max_corner, const float *box_min,
                                   const float *box_max, float *padded_min, float *padded_max,
                                   int64_t *grid_size, int64_t *n_cells, float *cell_size)
{
    if (ndims < 1 || ndims > 3) {
        set_error("`NDIMS` must be 1, 2, or 3");
        return PNB_ERR_ARG;
    }
    if (!min_corner || !max_corner) {
        set_error("min_corner and max_corner must have the same length");
        return PNB_ERR_ARG;
    }
    // full_grid.jl:66-67  min_corner .- 1001 // 1000 * search_radius
    volatile float factor = 1001.0f / 1000.0f;
    volatile float pad = factor * r;
    const bool is_template = (double)r < 2.220446049250313e-16;  // full_grid.jl:69, nhs_grid.jl:102
    for (int d = 0; d < ndims; d++) {
        volatile float mn = min_corner[d] - pad;
        volatile float mx = max_corner[d] + pad;
        if (padded_min) padded_min[d] = mn;
        if (padded_max) padded_max[d] = mx;
        if (grid_size) {
            if (is_template) grid_size[d] = 0;  // LinearIndices(ntuple(_ -> 0, NDIMS)), full_grid.jl:72
            else {
                volatile float ext = mx - mn;
                volatile float q = ext / r;            // full_grid.jl:74, evaluated in Float32
                grid_size[d] = (int64_t)std::ceil(q);
            }
        }
        if (n_cells) n_cells[d] = -1;      // nhs_grid.jl:104
        if (cell_size) cell_size[d] = r;   // nhs_grid.jl:105
    }
    if (box_min && box_max && !is_template) {
        for (int d = 0; d < ndims; d++) {
            volatile float size = box_max[d] - box_min[d];  // neighborhood_search.jl:139
            // nhs_grid.jl:117: `10eps()` is Float64, so the quotient is evaluated in Float64
            double nc = std::floor(((double)size + 10.0 * 2.220446049250313e-16) / (double)r);
            int64_t nci = (int64_t)nc;
            if (n_cells) n_cells[d] = nci;
            if (cell_size) {
                volatile float cs = size / (float)nci;      // nhs_grid.jl:118 in Float32
                cell_size[d] = cs;
            }
            if (nci < 3) {
                set_error("the `GridNeighborhoodSearch` needs at least 3 cells in each dimension "
                          "when used with periodicity. Please use no NHS for very small problems.");
                return PNB_ERR_ARG;
            }
        }
    }
    return PNB_OK;
}

}  // namespace pnb

using namespace pnb;

extern "C" int pnb_version(void) { return PNB200_VERSION; }
extern "C" const char *pnb_last_error(void) { return t_err; }
extern "C" int64_t pnb_launch_count(void) { return g_launch_count; }

extern "C" void pnb_profile_enable(int on) { g_prof_on = on != 0; }
extern "C" void pnb_profile_reset(void)
{
    prof_collect();
    for (int i = 0; i < PH_COUNT_; i++) { g_prof_ms[i] = 0.0; g_prof_n[i] = 0; }
}
extern "C" int pnb_profile_phases(void) { return PH_COUNT_; }
extern "C" const char *pnb_profile_name(int phase)
{
    return (phase >= 0 && phase < PH_COUNT_) ? kPhaseNames[phase] : "";
}
extern "C" pnb_status pnb_profile_get(int phase, double *total_ms, int64_t *launches)
{
    if (phase < 0 || phase >= PH_COUNT_) { set_error("no such phase"); return PNB_ERR_ARG; }
    prof_collect();
    if (total_ms) *total_ms = g_prof_ms[phase];
    if (launches) *launches = g_prof_n[phase];
    return PNB_OK;
}

extern "C" int pnb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" pnb_status pnb_grid_params_f32(int ndims, float search_radius, const float *min_corner,
                                          const float *max_corner, const float *box_min,
                                          const float *box_max, float *padded_min,
                                          float *padded_max, int64_t *grid_size, int64_t *n_cells,
                                          float *cell_size)
{
    return grid_params_host(ndims, search_radius, min_corner, max_corner, box_min, box_max,
                            padded_min, padded_max, grid_size, n_cells, cell_size);
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
extern "C" pnb_status pnb_grid_create_f32(int ndims, float r, const float *min_corner,
                                          const float *max_corner, const float *box_min,
                                          const float *box_max, pnb_grid **out)
{
    return pnb_grid_create_window_f32(ndims, r, min_corner, max_corner, box_min, box_max, nullptr,
                                      nullptr, out);
}

extern "C" pnb_status pnb_grid_create_window_f32(int ndims, float r, const float *min_corner,
                                                 const float *max_corner, const float *box_min,
                                                 const float *box_max, const int64_t *win_lo,
                                                 const int64_t *win_hi, pnb_grid **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    pnb_grid *g = new pnb_grid();
    memset(g, 0, sizeof(*g));
    pnb_status st = grid_params_host(ndims, r, min_corner, max_corner, box_min, box_max,
                                     g->padded_min, g->padded_max, g->grid_size, g->n_cells,
                                     g->cell_size);
    if (st != PNB_OK) { delete g; return st; }
    g->template_search = (double)r < 2.220446049250313e-16;
    GridP &p = g->p;
    p.ndims = ndims;
    p.periodic = (box_min && box_max && !g->template_search) ? 1 : 0;
    p.r = r;
    { volatile float r2 = r * r; p.r2 = r2; }
    int64_t total = 1;
    float min_size = INFINITY;
    for (int d = 0; d < 3; d++) {
        p.minc[d] = d < ndims ? g->padded_min[d] : 0.f;
        p.cs[d] = d < ndims ? g->cell_size[d] : 1.f;
        int64_t gs = d < ndims ? g->grid_size[d] : 1;
        if (g->template_search) gs = d < ndims ? 0 : 1;
        p.off[d] = 0;
        if (win_lo && win_hi && d < ndims && !g->template_search) {
            // window of the global grid (slab decomposition): global cells win_lo..win_hi
            // (1-based, inclusive, outermost layer = padding) become local cells 1..(hi-lo+1)
            if (box_min && box_max) {
                set_error("a windowed grid cannot be combined with a PeriodicBox");
                delete g;
                return PNB_ERR_ARG;
            }
            if (win_lo[d] < 1 || win_hi[d] > gs || win_hi[d] - win_lo[d] + 1 < 3) {
                set_error("grid window [%lld, %lld] is not inside 1..%lld with at least 3 cells",
                          (long long)win_lo[d], (long long)win_hi[d], (long long)gs);
                delete g;
                return PNB_ERR_ARG;
            }
            p.off[d] = (int)(win_lo[d] - 1);
            gs = win_hi[d] - win_lo[d] + 1;
        }
        if (gs > 0x7fffffff) gs = 0x7fffffff;
        p.gs[d] = (int)gs;
        p.nc[d] = d < ndims ? (int)g->n_cells[d] : -1;
        p.bsize[d] = 1.f;
        if (d < ndims && p.periodic) {
            g->box_min[d] = box_min[d];
            g->box_max[d] = box_max[d];
            volatile float size = box_max[d] - box_min[d];
            p.bsize[d] = size;
            if (size < min_size) min_size = size;
        }
        total *= gs;
        if (total > 0x7fffff00LL) {
            set_error("cell grid too large for this build: more than 2^31 cells");
            delete g;
            return PNB_ERR_ARG;
        }
    }
    p.total_cells = (int)total;
    p.wrap_d2 = p.periodic ? (0.49f * min_size) * (0.49f * min_size) : INFINITY;
    if (p.periodic && !(p.wrap_d2 > p.r2)) p.wrap_d2 = p.r2;  // cannot happen with >= 3 cells; stay exact

    cudaError_t e = cudaGetDevice(&g->device);
    if (e != cudaSuccess) { delete g; return cuda_fail(e, "cudaGetDevice"); }
    auto fail = [&](cudaError_t err, const char *what) {
        pnb_status s = cuda_fail(err, what);
        pnb_grid_destroy(g);
        return s;
    };
    int64_t C = total;
    if ((e = cudaMalloc(&g->cell_start, sizeof(uint32_t) * (size_t)(C + 1))) != cudaSuccess)
        return fail(e, "cudaMalloc cell_start");
    if ((e = cudaMalloc(&g->cell_count, sizeof(uint32_t) * (size_t)(C + 1))) != cudaSuccess)
        return fail(e, "cudaMalloc cell_count");
    if ((e = cudaMemset(g->cell_start, 0, sizeof(uint32_t) * (size_t)(C + 1))) != cudaSuccess)
        return fail(e, "cudaMemset");
    if ((e = cudaMalloc(&g->d_err, sizeof(int))) != cudaSuccess) return fail(e, "cudaMalloc err");
    if ((e = cudaMemset(g->d_err, 0, sizeof(int))) != cudaSuccess) return fail(e, "cudaMemset");
    if ((e = cudaMallocHost(&g->h_err, sizeof(int))) != cudaSuccess)
        return fail(e, "cudaMallocHost");
    if ((e = cudaMalloc(&g->scan_ticket, sizeof(unsigned int))) != cudaSuccess)
        return fail(e, "cudaMalloc ticket");
    *out = g;
    return PNB_OK;
}

extern "C" void pnb_grid_destroy(pnb_grid *g)
{
    if (!g) return;
    cudaFree(g->cell_start);
    cudaFree(g->cell_count);
    cudaFree(g->cell_points);
    cudaFree(g->ids_tmp);
    cudaFree(g->cell_rank);
    cudaFree(g->sorted);
    cudaFree(g->scan_status);
    cudaFree(g->scan_ticket);
    cudaFree(g->d_err);
    if (g->h_err) cudaFreeHost(g->h_err);
    cudaFree(g->scratch);
    cudaFree(g->ovf_tiles);
    cudaFree(g->ovf_count);
    cudaGetLastError();
    delete g;
}

extern "C" int64_t pnb_grid_total_cells(const pnb_grid *g) { return g ? g->p.total_cells : 0; }
extern "C" int64_t pnb_grid_n_points(const pnb_grid *g) { return g ? g->n_built : 0; }

// ---------------------------------------------------------------------------------------------
// kernels of the counting sort
// ---------------------------------------------------------------------------------------------
namespace pnb {

constexpr int kBuildThreads = 256;

// K_a: cell index + domain check + histogram with warp-aggregated atomics.
//   reference: the first half of initialize_grid! (src/nhs_grid.jl:271-278): cell_coords,
//   check_cell_bounds, and the `lengths[cell] += 1` of pushat_atomic!
//   (src/vector_of_vectors.jl:102).  The returned old value is the point's provisional rank.
// HBM: reads 12 B/point (coordinates, staged through shared memory so the three component
// loads of a warp are three fully coalesced 128 B requests), writes 8 B/point.
template <int ND>
__global__ void __launch_bounds__(kBuildThreads)
k_cell_count(GridP g, const float *__restrict__ y, int64_t n_idx, const int32_t *__restrict__ idx,
             int base, int2 *__restrict__ cell_rank, uint32_t *__restrict__ cell_count,
             int *__restrict__ err)
{
    __shared__ float s_xyz[kBuildThreads * ND];
    const int64_t block0 = (int64_t)blockIdx.x * kBuildThreads;
    const int64_t k = block0 + threadIdx.x;
    float p[3] = {0.f, 0.f, 0.f};
    if (idx == nullptr) {
        // coalesced tile load of kBuildThreads * ND consecutive floats
        const int64_t f0 = block0 * ND;
        const int64_t fend = n_idx * ND;
#pragma unroll
        for (int t = 0; t < ND; t++) {
            int64_t f = f0 + threadIdx.x + (int64_t)t * kBuildThreads;
            if (f < fend) s_xyz[threadIdx.x + t * kBuildThreads] = __ldg(y + f);
        }
        __syncthreads();
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = s_xyz[threadIdx.x * ND + d];
    } else if (k < n_idx) {
        const int64_t pt = (int64_t)idx[k] - base;
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __ldg(y + pt * ND + d);
    }
    bool in_range = k < n_idx;
    int cc[3];
    int lin = in_range ? point_cell<ND>(g, p, cc) : -1;
    bool valid = in_range && lin >= 0;
    if (in_range && lin < 0) atomicOr(err, 1);
    unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        unsigned m = __match_any_sync(act, lin);
        int leader = __ffs(m) - 1;
        int pos = __popc(m & ((1u << lane_id()) - 1u));
        unsigned basev = 0;
        if (lane_id() == leader) basev = atomicAdd(cell_count + lin, (unsigned)__popc(m));
        basev = __shfl_sync(m, basev, leader);
        cell_rank[k] = make_int2(lin, (int)(basev + pos));
    } else if (in_range) {
        cell_rank[k] = make_int2(-1, 0);
    }
}

// K_c: scatter ids to cell_start[cell] + rank   (the `backend[new_length, i] = value` of
// pushat_atomic!, src/vector_of_vectors.jl:109, into a CSR instead of a 100-row matrix).
__global__ void __launch_bounds__(kBuildThreads)
k_scatter(int64_t n_idx, const int32_t *__restrict__ idx, int base,
          const int2 *__restrict__ cell_rank, const uint32_t *__restrict__ cell_start,
          int32_t *__restrict__ ids_tmp)
{
    int64_t k = (int64_t)blockIdx.x * kBuildThreads + threadIdx.x;
    if (k >= n_idx) return;
    int2 cr = cell_rank[k];
    if (cr.x < 0) return;
    int32_t id = idx ? idx[k] - base : (int32_t)k;
    ids_tmp[cell_start[cr.x] + (uint32_t)cr.y] = id;
}

// K_d: one warp per cell: order the cell's ids ascending (rank by counting), write them to
// cell_points and gather the coordinates into the cell-ordered float4 array (x, y, z, id).
// Cells with <= 32 points (the normal case) stay in registers.
template <int ND>
__global__ void __launch_bounds__(256)
k_finalize_cells(int total_cells, const uint32_t *__restrict__ cell_start,
                 const int32_t *__restrict__ ids_tmp, const float *__restrict__ y,
                 int32_t *__restrict__ cell_points, float4 *__restrict__ sorted)
{
    int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= total_cells) return;
    const int lane = lane_id();
    const uint32_t s0 = cell_start[c], s1 = cell_start[c + 1];
    const int cnt = (int)(s1 - s0);
    if (cnt == 0) return;
    if (cnt <= 32) {
        int v = lane < cnt ? ids_tmp[s0 + lane] : 0x7fffffff;
        int r = 0;
        for (int k = 0; k < cnt; k++) r += (__shfl_sync(0xffffffffu, v, k) < v) ? 1 : 0;
        if (lane < cnt) {
            float px = __ldg(y + (int64_t)v * ND);
            float py = ND > 1 ? __ldg(y + (int64_t)v * ND + 1) : 0.f;
            float pz = ND > 2 ? __ldg(y + (int64_t)v * ND + 2) : 0.f;
            cell_points[s0 + r] = v;
            sorted[s0 + r] = make_float4(px, py, pz, __int_as_float(v));
        }
    } else {
        for (int e = lane; e < cnt; e += 32) {
            int v = ids_tmp[s0 + e];
            int r = 0;
            for (int k = 0; k < cnt; k++) r += (ids_tmp[s0 + k] < v) ? 1 : 0;
            float px = __ldg(y + (int64_t)v * ND);
            float py = ND > 1 ? __ldg(y + (int64_t)v * ND + 1) : 0.f;
            float pz = ND > 2 ? __ldg(y + (int64_t)v * ND + 2) : 0.f;
            cell_points[s0 + r] = v;
            sorted[s0 + r] = make_float4(px, py, pz, __int_as_float(v));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-pass exclusive scan with decoupled look-back (Merrill & Garland), uint32
//   tile = 256 threads x 8 items; status word = (flag << 62) | value, flag 1 = aggregate,
//   2 = inclusive prefix.  Tiles are taken from an atomic ticket so a tile never waits on a
//   tile that has not started.
// Replaces nothing in the reference (its 100-row cell matrix needs no offsets): this is what
// turns the histogram into CSR offsets.  HBM: reads 4 B/cell, writes 4 B/cell.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// status word: (flag << 62) | value, flag 1 = tile aggregate, 2 = inclusive prefix
constexpr unsigned long long kScanValMask = (1ULL << 62) - 1ULL;

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(const uint32_t *__restrict__ in, OutT *__restrict__ out, int64_t n,
                unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket)
{
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;

    uint32_t v[kScanItems];
    if (base + kScanItems <= n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(in + base));
        const uint4 b = __ldg(reinterpret_cast<const uint4 *>(in + base) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v[k] = (base + k < n) ? __ldg(in + base + k) : 0u;
    }
    // a tile holds 2048 items; callers guarantee the sum of one tile fits 32 bits
    uint32_t tsum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) tsum += v[k];

    // block exclusive scan of the thread sums
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, block_agg = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        uint32_t t = s_warp[w];
        if (w < warp) warp_off += t;
        block_agg += t;
    }
    const uint32_t thread_excl = warp_off + incl - tsum;

    // publish the aggregate, look back for the exclusive prefix of this tile
    if (warp == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            if (lane == 0) st_status(status, (2ULL << 62) | block_agg);
        } else {
            if (lane == 0) st_status(status + tile, (1ULL << 62) | block_agg);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t t = look - lane;
                const bool have = t >= 0;
                unsigned long long sv;
                do {
                    sv = have ? ld_status(status + t) : (2ULL << 62);
                } while (__any_sync(0xffffffffu, (sv >> 62) == 0));
                const unsigned incl_mask = __ballot_sync(0xffffffffu, (sv >> 62) == 2);
                const int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
                unsigned long long contrib = (lane <= first) ? (sv & kScanValMask) : 0ULL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                excl += contrib;
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0) st_status(status + tile, (2ULL << 62) | ((excl + block_agg) & kScanValMask));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    unsigned long long run = s_prefix + thread_excl;
    const bool owns_last = base < n && base + kScanItems >= n;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = (OutT)run;
        run += (base + k < n) ? v[k] : 0u;
    }
    // the thread that owns the last element also writes the total at out[n]
    if (owns_last) out[n] = (OutT)run;
}

template <typename OutT>
static pnb_status scan_impl(pnb_grid *g, const uint32_t *in, OutT *out, int64_t n, cudaStream_t s)
{
    if (n <= 0) {
        PNB_CUDA(cudaMemsetAsync(out, 0, sizeof(OutT), s));
        return PNB_OK;
    }
    int64_t tiles = div_up(n, kScanTile);
    if (tiles > g->scan_tiles_cap) {
        if (g->scan_status) PNB_CUDA(cudaFree(g->scan_status));
        g->scan_status = nullptr;
        g->scan_tiles_cap = 0;
        PNB_CUDA(cudaMalloc(&g->scan_status, sizeof(unsigned long long) * (size_t)tiles));
        g->scan_tiles_cap = tiles;
    }
    ProfScope ps(PH_BUILD_SCAN, s);
    PNB_CUDA(cudaMemsetAsync(g->scan_status, 0, sizeof(unsigned long long) * (size_t)tiles, s));
    PNB_CUDA(cudaMemsetAsync(g->scan_ticket, 0, sizeof(unsigned int), s));
    k_scan_lookback<OutT><<<(unsigned)tiles, kScanThreads, 0, s>>>(in, out, n, g->scan_status,
                                                                    g->scan_ticket);
    PNB_LAUNCHED();
    return PNB_OK;
}

pnb_status exclusive_scan_u32(pnb_grid *g, const uint32_t *in, uint32_t *out, int64_t n,
                              cudaStream_t s)
{
    return scan_impl<uint32_t>(g, in, out, n, s);
}

pnb_status exclusive_scan_u32_to_i64(pnb_grid *g, const uint32_t *in, int64_t *out, int64_t n,
                                     cudaStream_t s)
{
    return scan_impl<int64_t>(g, in, out, n, s);
}

static pnb_status ensure_point_capacity(pnb_grid *g, int64_t n)
{
    if (n <= g->cap_points) return PNB_OK;
    cudaFree(g->cell_points); g->cell_points = nullptr;
    cudaFree(g->ids_tmp); g->ids_tmp = nullptr;
    cudaFree(g->cell_rank); g->cell_rank = nullptr;
    cudaFree(g->sorted); g->sorted = nullptr;
    g->cap_points = 0;
    int64_t cap = n + n / 16 + 32;
    PNB_CUDA(cudaMalloc(&g->cell_points, sizeof(int32_t) * (size_t)cap));
    PNB_CUDA(cudaMalloc(&g->ids_tmp, sizeof(int32_t) * (size_t)cap));
    PNB_CUDA(cudaMalloc(&g->cell_rank, sizeof(int2) * (size_t)cap));
    PNB_CUDA(cudaMalloc(&g->sorted, sizeof(float4) * (size_t)cap));
    g->cap_points = cap;
    return PNB_OK;
}

template <int ND>
static pnb_status build_nd(pnb_grid *g, const float *y, int64_t n, const int32_t *idx,
                           int64_t n_idx, int base, cudaStream_t s)
{
    const int64_t C = g->p.total_cells;
    PNB_CUDA(cudaMemsetAsync(g->cell_count, 0, sizeof(uint32_t) * (size_t)(C + 1), s));
    if (n_idx > 0) {
        unsigned blocks = (unsigned)div_up(n_idx, kBuildThreads);
        ProfScope ps(PH_BUILD_CELL_COUNT, s);
        k_cell_count<ND><<<blocks, kBuildThreads, 0, s>>>(g->p, y, n_idx, idx, base, g->cell_rank,
                                                          g->cell_count, g->d_err);
        PNB_LAUNCHED();
    }
    pnb_status st = exclusive_scan_u32(g, g->cell_count, g->cell_start, C, s);
    if (st != PNB_OK) return st;
    if (n_idx > 0) {
        unsigned blocks = (unsigned)div_up(n_idx, kBuildThreads);
        {
            ProfScope ps(PH_BUILD_SCATTER, s);
            k_scatter<<<blocks, kBuildThreads, 0, s>>>(n_idx, idx, base, g->cell_rank,
                                                       g->cell_start, g->ids_tmp);
            PNB_LAUNCHED();
        }
        unsigned fblocks = (unsigned)div_up(C * 32, 256);
        ProfScope ps(PH_BUILD_FINALIZE, s);
        k_finalize_cells<ND><<<fblocks, 256, 0, s>>>((int)C, g->cell_start, g->ids_tmp, y,
                                                     g->cell_points, g->sorted);
        PNB_LAUNCHED();
    }
    (void)n;
    return PNB_OK;
}

}  // namespace pnb

extern "C" pnb_status pnb_grid_build_f32(pnb_grid *g, const float *y, int64_t n,
                                         const int32_t *eachindex_y, int64_t n_idx, int index_base,
                                         void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t C = g->p.total_cells;
    if (eachindex_y == nullptr) n_idx = n;
    g->built = false;
    // empty!(cell_list)  (src/cell_lists/full_grid.jl:96-105)
    if (g->template_search) {
        // nhs_grid.jl:263-267: zero search radius -> emptied cell list, nothing else
        PNB_CUDA(cudaMemsetAsync(g->cell_start, 0, sizeof(uint32_t) * (size_t)(C + 1), s));
        PNB_CUDA(cudaStreamSynchronize(s));
        g->n_built = 0;
        g->built = true;
        g->y_built = y;
        g->n_y_built = n;
        g->full_build = false;
        return PNB_OK;
    }
    if (n_idx > 0x7ffffff0LL || n > 0x7ffffff0LL) {
        set_error("more than 2^31 points are not supported (ids are Int32, full_grid.jl:177)");
        return PNB_ERR_ARG;
    }
    if (n_idx > 0 && y == nullptr) { set_error("y is NULL"); return PNB_ERR_ARG; }
    pnb_status st = ensure_point_capacity(g, n_idx);
    if (st != PNB_OK) return st;
    switch (g->p.ndims) {
        case 1: st = build_nd<1>(g, y, n, eachindex_y, n_idx, index_base, s); break;
        case 2: st = build_nd<2>(g, y, n, eachindex_y, n_idx, index_base, s); break;
        default: st = build_nd<3>(g, y, n, eachindex_y, n_idx, index_base, s); break;
    }
    if (st != PNB_OK) return st;
    st = check_err_word(g, s);  // also synchronizes: initialize!/update! are blocking calls
    if (st != PNB_OK) {
        // like the reference, a failed build leaves an unusable cell list behind
        g->n_built = 0;
        return st;
    }
    g->n_built = n_idx;
    g->y_built = y;
    g->n_y_built = n;
    g->full_build = (eachindex_y == nullptr);
    g->built = true;
    return PNB_OK;
}

// ---------------------------------------------------------------------------------------------
// pnb_point_cells
// ---------------------------------------------------------------------------------------------
namespace pnb {
template <int ND>
__global__ void k_point_cells(GridP g, const float *__restrict__ x, int64_t n,
                              int32_t *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = x[i * ND + d];
    int cc[3];
    out[i] = point_cell<ND>(g, p, cc);
}
}  // namespace pnb

extern "C" pnb_status pnb_point_cells_f32(const pnb_grid *g, const float *x, int64_t n,
                                          int32_t *out_linear, void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (g->template_search) {
        set_error("`search_radius` is not defined for this cell list");
        return PNB_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        unsigned blocks = (unsigned)div_up(n, 256);
        switch (g->p.ndims) {
            case 1: k_point_cells<1><<<blocks, 256, 0, s>>>(g->p, x, n, out_linear); break;
            case 2: k_point_cells<2><<<blocks, 256, 0, s>>>(g->p, x, n, out_linear); break;
            default: k_point_cells<3><<<blocks, 256, 0, s>>>(g->p, x, n, out_linear); break;
        }
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

// ---------------------------------------------------------------------------------------------
// exports
// ---------------------------------------------------------------------------------------------
namespace pnb {
__global__ void k_export_csr(int64_t n_cells_p1, int64_t n_pts, const uint32_t *__restrict__ cs,
                             const int32_t *__restrict__ cp, int32_t *__restrict__ out_cs,
                             int32_t *__restrict__ out_cp, int base)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (out_cs && i < n_cells_p1) out_cs[i] = (int32_t)cs[i];
    if (out_cp && i < n_pts) out_cp[i] = cp[i] + base;
}

// DynamicVectorOfVectors layout: backend[max_inner x C] column-major, lengths[C]
// (src/vector_of_vectors.jl:3-31).  One warp per cell.
__global__ void k_export_dvov(int total_cells, const uint32_t *__restrict__ cs,
                              const int32_t *__restrict__ cp, int32_t *__restrict__ backend,
                              int32_t *__restrict__ lengths, int max_inner, int base,
                              int *__restrict__ err)
{
    int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= total_cells) return;
    const uint32_t s0 = cs[c], s1 = cs[c + 1];
    int cnt = (int)(s1 - s0);
    if (cnt > max_inner) {
        if (lane_id() == 0) atomicOr(err, 4);
        cnt = max_inner;
    }
    if (lane_id() == 0) lengths[c] = cnt;
    for (int e = lane_id(); e < cnt; e += 32) backend[c * (int64_t)max_inner + e] = cp[s0 + e] + base;
}
}  // namespace pnb

extern "C" pnb_status pnb_grid_export_csr(const pnb_grid *g, int32_t *cell_start,
                                          int32_t *cell_points, int index_base, void *stream)
{
    if (!g || !g->built) { set_error("the neighborhood search has not been initialized"); return PNB_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t C1 = (int64_t)g->p.total_cells + 1;
    int64_t m = C1 > g->n_built ? C1 : g->n_built;
    k_export_csr<<<(unsigned)div_up(m, 256), 256, 0, s>>>(C1, g->n_built, g->cell_start,
                                                          g->cell_points, cell_start, cell_points,
                                                          index_base);
    PNB_LAUNCHED();
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_export_dvov(const pnb_grid *g_, int32_t *backend, int32_t *lengths,
                                           int32_t max_points_per_cell, int index_base,
                                           void *stream)
{
    pnb_grid *g = const_cast<pnb_grid *>(g_);
    if (!g || !g->built) { set_error("the neighborhood search has not been initialized"); return PNB_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    int64_t C = g->p.total_cells;
    if (C > 0) {
        k_export_dvov<<<(unsigned)div_up(C * 32, 256), 256, 0, s>>>(
            (int)C, g->cell_start, g->cell_points, backend, lengths, max_points_per_cell,
            index_base, g->d_err);
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
extern "C" pnb_status pnb_malloc(void **p, int64_t bytes)
{
    if (!p) { set_error("NULL"); return PNB_ERR_ARG; }
    PNB_CUDA(cudaMalloc(p, (size_t)(bytes > 0 ? bytes : 1)));
    return PNB_OK;
}
extern "C" pnb_status pnb_free(void *p) { PNB_CUDA(cudaFree(p)); return PNB_OK; }
extern "C" pnb_status pnb_malloc_host(void **p, int64_t bytes)
{
    if (!p) { set_error("NULL"); return PNB_ERR_ARG; }
    PNB_CUDA(cudaMallocHost(p, (size_t)(bytes > 0 ? bytes : 1)));
    return PNB_OK;
}
extern "C" pnb_status pnb_free_host(void *p) { PNB_CUDA(cudaFreeHost(p)); return PNB_OK; }
extern "C" pnb_status pnb_memcpy_h2d(void *dst, const void *src, int64_t bytes, void *stream)
{
    PNB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    PNB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PNB_OK;
}
extern "C" pnb_status pnb_memcpy_d2h(void *dst, const void *src, int64_t bytes, void *stream)
{
    PNB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    PNB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PNB_OK;
}
extern "C" pnb_status pnb_memset(void *p, int value, int64_t bytes, void *stream)
{
    PNB_CUDA(cudaMemsetAsync(p, value, (size_t)bytes, (cudaStream_t)stream));
    return PNB_OK;
}
extern "C" pnb_status pnb_stream_synchronize(void *stream)
{
    PNB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return PNB_OK;
}
