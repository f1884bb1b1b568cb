// grid.cu -- constructors (host arithmetic), the device counting sort behind initialize!/update!,
// the decoupled-lookback scan, and the cell-list exports.
//
// reference: src/cell_lists/full_grid.jl (FullGridCellList), src/nhs_grid.jl:77-129, 220-292,
// 470-477 (constructor, initialize!, update!), src/vector_of_vectors.jl (the layout exported by
// pnb_grid_export_dvov).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "grid.cuh"
#include "f64.cuh"
#include "hashgrid.cuh"

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
    "k_cell_hist", "k_scan_lookback", "k_scatter_points", "k_canonicalize", "k_gather",
    "k_sweep_cells", "k_sweep_overflow", "k_sweep_points", "k_sort_lists", "k_nlist_sweep", "k_export",
    "k_bucket_scatter", "k_flat_tiles"};

static const char *kDomainMsg =
    "particle coordinates are NaN or outside the domain bounds of the cell list";
static const char *kListFullMsg = "cell list is full. Use a larger `max_points_per_cell`.";
static const char *kBoundsMsg =
    "BoundsError: a neighboring cell of a query point is outside the cell grid";

static pnb_status translate_err_word(pnb_grid *g)
{
    int e = *(volatile int *)g->h_err;
    if (e == 0) return PNB_OK;
    *(volatile int *)g->h_err = 0;
    e &= ~8;   // bit 3 (bucket overflow) is consumed by the build itself
    if (e & 1) { set_error("%s", kDomainMsg); return PNB_ERR_DOMAIN; }
    if (e & 4) { set_error("%s", kListFullMsg); return PNB_ERR_LIST_FULL; }
    if (e == 0) return PNB_OK;
    set_error("%s", kBoundsMsg);
    return PNB_ERR_BOUNDS;
}

// After the stream of a stream-ordered update! has been synchronised: a bucket overflow means
// the cell list is incomplete -> blocking two-pass rebuild from the same coordinates (which also
// picks a larger K).  *rebuilt tells the caller that work launched on the old list is void.
static pnb_status settle_async_build(pnb_grid *g, bool *rebuilt)
{
    *rebuilt = false;
    if (!g->async_pending) return PNB_OK;
    g->async_pending = false;
    const int e = *(volatile int *)g->h_err;
    if ((e & 8) == 0) {
        g->async_chain = 0;
        if (e & 1) {       // domain error of the build: like the blocking build, the list is unusable
            *(volatile int *)g->h_err = 0;
            g->n_built = 0;
            g->built = false;
            set_error("%s", kDomainMsg);
            return PNB_ERR_DOMAIN;
        }
        return PNB_OK;
    }
    *(volatile int *)g->h_err = 0;
    g->bucket_K = 0;
    g->bcount_alt_clean = false;
    *rebuilt = true;
    if (g->async_chain > 1) {
        // several stream-ordered steps were chained without a check: the overflow may have
        // happened in any of them, their sweeps cannot be repeated
        const int chain = g->async_chain;
        g->async_chain = 0;
        pnb_grid_build_f32(g, (const float *)g->y_built, g->n_y_built, nullptr, 0, 0, (void *)g->async_stream);
        set_error("a bucket of the one-pass update! overflowed inside a chain of %d stream-ordered "
                  "steps: their results are invalid (check every step, or use the blocking update!)", chain);
        return PNB_ERR_STATE;
    }
    g->async_chain = 0;
    return pnb_grid_build_f32(g, (const float *)g->y_built, g->n_y_built, nullptr, 0, 0,
                              (void *)g->async_stream);
}

pnb_status resolve_pending(pnb_grid *g)
{
    if (!g->async_pending) return PNB_OK;
    PNB_CUDA(cudaStreamSynchronize(g->async_stream));
    bool rebuilt = false;
    return settle_async_build(g, &rebuilt);
}

static pnb_status check_err_word_settled(pnb_grid *g)
{
    if (g->async_pending) {
        bool rebuilt = false;
        pnb_status st = settle_async_build(g, &rebuilt);
        if (st != PNB_OK) return st;
        if (rebuilt) return PNB_RETRY_INTERNAL;
    }
    return translate_err_word(g);
}

pnb_status check_err_word(pnb_grid *g, cudaStream_t s)
{
    // the error word lives in mapped pinned host memory (kernels OR their bits into it through
    // d_err): one stream synchronisation, no copy
    PNB_CUDA(cudaStreamSynchronize(s));
    if (g->async_pending && g->async_stream != s) PNB_CUDA(cudaStreamSynchronize(g->async_stream));
    return check_err_word_settled(g);
}

// The reference reads neighbor_coords LIVE at sweep time (src/nhs_grid.jl:543-548) while its
// cell list is the one of the last initialize!/update!; this library sweeps cell-ordered records
// (x, y, z, id) snapshotted at that build.  Same array as at build time: the snapshot is current
// (moving y in place requires update!, src/neighborhood_search.jl:161-164).  Another array with
// the same number of points: the reference's semantics are "old cell list, coordinates of THIS
// array", so the records' coordinates are re-read from it through their ids (one gather pass);
// afterwards the snapshot belongs to the new array.  A different number of points is a
// call-order error.
__global__ void k_refresh_records(int64_t n_slots, int nd, CellsView cv, const float *__restrict__ y,
                                  float4 *__restrict__ rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    if (cv.K) {
        const uint32_t sl = (uint32_t)i;
        if ((sl & (cv.K - 1u)) >= cv.start[sl >> (31 - __clz((int)cv.K))]) return;
    }
    float4 r = rec[i];
    const int64_t id = __float_as_int(r.w);
    r.x = __ldg(y + id * nd);
    if (nd > 1) r.y = __ldg(y + id * nd + 1);
    if (nd > 2) r.z = __ldg(y + id * nd + 2);
    rec[i] = r;
}
__global__ void k_refresh_records64(int64_t n, int nd, const double *__restrict__ y,
                                    Rec64 *__restrict__ rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rec64 r = rec[i];
    r.x = y[r.id * nd];
    if (nd > 1) r.y = y[r.id * nd + 1];
    if (nd > 2) r.z = y[r.id * nd + 2];
    rec[i] = r;
}

pnb_status check_built_y(pnb_grid *g, const void *y, int64_t n, cudaStream_t s)
{
    if (y == g->y_built && n == g->n_y_built) return PNB_OK;
    { pnb_status sp = resolve_pending(g); if (sp != PNB_OK) return sp; }
    if (n != g->n_y_built) {
        set_error("neighbor_coords hold %lld points, the search was last initialized / updated "
                  "with %lld: call update! first", (long long)n, (long long)g->n_y_built);
        return PNB_ERR_STATE;
    }
    if (g->template_search || g->n_built == 0 || y == nullptr) { g->y_built = y; return PNB_OK; }
    if (g->f64) {
        k_refresh_records64<<<(unsigned)div_up(g->n_built, 256), 256, 0, s>>>(
            g->n_built, g->p64.ndims, (const double *)y, g->sorted64);
        PNB_LAUNCHED();
    } else {
        if (g->bucket_valid) {
            const int64_t slots = (int64_t)g->p.total_cells * g->bucket_K;
            k_refresh_records<<<(unsigned)div_up(slots, 256), 256, 0, s>>>(
                slots, g->p.ndims, cells_view(g), (const float *)y, g->brec);
            PNB_LAUNCHED();
        }
        if (g->csr_valid || !g->bucket_valid) {
            const CellsView csr{g->cell_start, g->sorted, 0u, 0u, 0u, 0u};
            k_refresh_records<<<(unsigned)div_up(g->n_built, 256), 256, 0, s>>>(
                g->n_built, g->p.ndims, csr, (const float *)y, g->sorted);
            PNB_LAUNCHED();
        }
    }
    g->y_built = y;
    g->y_refreshed = true;
    return PNB_OK;
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
// set while a grid is created from the corners STORED in a FullGridCellList (already padded,
// src/cell_lists/full_grid.jl:66-67): the constructor arithmetic then starts at :74
static thread_local bool t_corners_padded = false;

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
    volatile float pad = t_corners_padded ? 0.0f : factor * r;
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


pnb_status grid_alloc_common(pnb_grid *g, int64_t C)
{
    PNB_CUDA(cudaGetDevice(&g->device));
    // cell_start = alloc + 3 so that cell_start + 1 (the scan output / scatter cursor) is 16-byte
    // aligned; cell_start[0] = 0 is written here once and never again
    PNB_CUDA(cudaMalloc(&g->cell_start_alloc, sizeof(uint32_t) * (size_t)(C + 8)));
    g->cell_start = g->cell_start_alloc + 3;
    PNB_CUDA(cudaMalloc(&g->cell_count, sizeof(uint32_t) * (size_t)(C + 4)));
    PNB_CUDA(cudaMemset(g->cell_start_alloc, 0, sizeof(uint32_t) * (size_t)(C + 8)));
    PNB_CUDA(cudaMemset(g->cell_count, 0, sizeof(uint32_t) * (size_t)(C + 4)));
    PNB_CUDA(cudaHostAlloc(&g->h_err, 4 * sizeof(int), cudaHostAllocMapped));
    g->h_err[0] = g->h_err[1] = g->h_err[2] = g->h_err[3] = 0;
    PNB_CUDA(cudaHostGetDevicePointer(&g->d_err, g->h_err, 0));
    PNB_CUDA(cudaMalloc(&g->scan_ticket, sizeof(unsigned int)));
    PNB_CUDA(cudaMemset(g->scan_ticket, 0, sizeof(unsigned int)));
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
extern "C" pnb_status pnb_grid_create_padded_f32(int ndims, float r, const float *padded_min,
                                                 const float *padded_max, const float *box_min,
                                                 const float *box_max, pnb_grid **out)
{
    t_corners_padded = true;
    pnb_status st = pnb_grid_create_window_f32(ndims, r, padded_min, padded_max, box_min, box_max,
                                               nullptr, nullptr, out);
    t_corners_padded = false;
    return st;
}

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
            g->windowed = true;
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

    st = grid_alloc_common(g, total);
    if (st != PNB_OK) { pnb_grid_destroy(g); return st; }
    *out = g;
    return PNB_OK;
}

extern "C" void pnb_grid_destroy(pnb_grid *g)
{
    if (!g) return;
    cudaFree(g->cell_start_alloc);
    cudaFree(g->xq_start_alloc);
    cudaFree(g->hmeta);
    cudaFree(g->xq_sorted);
    cudaFree(g->bcount);
    cudaFree(g->bcount_alt);
    cudaFree(g->brec);
    cudaFree(g->sorted64);
    cudaFree(g->sorted64_tmp);
    cudaFree(g->d_maxcount);
    cudaFree(g->cell_count);
    cudaFree(g->cell_points);
    cudaFree(g->sorted);
    cudaFree(g->sorted_alt);
    cudaFree(g->scan_status);
    cudaFree(g->scan_ticket);
    if (g->h_err) cudaFreeHost(g->h_err);
    cudaFree(g->scratch);
    cudaFree(g->ovf_tiles);
    cudaFree(g->ovf_count);
    cudaFree(g->left_ids);
    cudaFree(g->flat_tiles);
    cudaFree(g->flat_ovf);
    cudaFree(g->flat_tabs);
    cudaFree(g->flat_seg);
    cudaFree(g->flat_ctl);
    cudaGetLastError();
    delete g;
}

extern "C" int64_t pnb_grid_total_cells(const pnb_grid *g) { return g ? g->p.total_cells : 0; }
extern "C" int64_t pnb_grid_n_points(const pnb_grid *g) { return g ? g->n_built : 0; }
extern "C" int pnb_grid_layout(const pnb_grid *g)
{
    if (!g || !g->built) return -1;
    return g->bucket_valid ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// kernels of the counting sort
//
//   update! = k_cell_hist (read 12 N) -> k_scan_lookback (8 C) -> k_scatter_points (read 12 N,
//   write 16 N): 40 N + 8 C bytes, two passes over the coordinates and nothing else.  The cell
//   index is recomputed in the second pass instead of being written and re-read, and the
//   slot of a point is taken from an atomic cursor (= cell_start itself, see below), so the
//   order of the points inside a cell is arrival order -- exactly what the reference has
//   (pushat_atomic!, src/vector_of_vectors.jl:99-109).  Consumers that need a reproducible
//   order (exports, exact arithmetic mode, neighbour-list fills) call ensure_canonical(), which
//   sorts every cell by point id once (k_canonicalize) and is skipped by the fast sweeps.
// ---------------------------------------------------------------------------------------------
namespace pnb {

// 1 = update! may use the one-pass bucket layout, 0 = always CSR, 2 = (tests) every build ends in
// the bucket layout; PNB_BUILD_LAYOUT in the environment overrides the default
static int initial_build_layout()
{
    const char *e = getenv("PNB_BUILD_LAYOUT");
    return e ? atoi(e) : 1;
}
int g_build_layout = initial_build_layout();
// numbering of the buckets: 0 = linear cell order, 1 = transposed (last dimension fastest),
// -1 (default) = chosen from the order of the input at the first build; PNB_BUCKET_ORDER overrides.
// Measured (config 3, the reference generator's order = last dimension fastest).  Round 1:
// transposed buckets made the one-pass build 5 % faster (0.128 vs 0.134 ms) but k_sweep_tiles 4 %
// slower (x-rows no longer contiguous for its register staging), so linear was the default.
// Round 2: k_sweep_flat stages cell by cell with cp.async and does not care (11.07 ms either
// way); the build goes from 0.124 to 0.112 ms (consecutive points write neighbouring buckets:
// DRAM page locality) = 64 % of the HBM peak, the tile pre-pass from 0.068 to 0.085 ms (strided
// count reads).  Windows of a slab decomposition keep the plain order (their layered gathers
// rely on it).
static int initial_bucket_order()
{
    const char *e = getenv("PNB_BUCKET_ORDER");
    return e ? atoi(e) : -1;
}
int g_bucket_order = initial_bucket_order();
int g_tune_build = 25;   // measurement variants of the build kernels (pnb_set_build_tuning)
constexpr int kBuildThreads = 256;
constexpr int kBuildPPT = 4;                                // points per thread
constexpr int kBuildTile = kBuildThreads * kBuildPPT;       // points per block

struct BuildP {
    float rcs[3];   // fl(1 / cell_size): fast quotient, verified against the exact division
};

// cell_coords for one dimension (src/nhs_grid.jl:622-628, src/cell_lists/full_grid.jl:93) with
// the division replaced by a multiplication whenever that provably gives the same floor:
//   q' = fl(t * fl(1/cs)) differs from the IEEE quotient fl(t / cs) by less than |q| 2^-22, so
//   if q' is farther than |q| 2^-21 from both neighbouring integers the two floors agree.
// Otherwise (about 1e-4 of the points) the exact path of common.cuh decides.
template <bool PER>
__device__ __forceinline__ int cell_coord_fast(float x, float minc, float cs, float rcs, int nc,
                                               int off)
{
    const float t = __fsub_rn(x, minc);
    const float q = __fmul_rn(t, rcs);
    const float f = floorf(q);
    const float fr = __fsub_rn(q, f);                      // exact
    const float delta = __fmul_rn(fabsf(q), 4.76837158203125e-7f);   // |q| 2^-21
    if (!(fr > delta && fr < __fsub_rn(1.0f, delta) && fabsf(f) < 4194304.0f))
        return cell_coord(x, minc, cs, PER ? 1 : 0, nc, off);
    int c = (int)f + 1;
    if (PER) {
        c -= 2;
        if (c < 0) c += nc; else if (c >= nc) c -= nc;
        if (c < 0 || c >= nc) c = floormod_i(c, nc);
        c += 2;
    }
    return c - off;
}

// TR: the TRANSPOSED cell number (last dimension fastest) = the bucket number when the buckets
// are numbered in that order (CellsView, common.cuh)
template <int ND, bool PER, bool TR = false>
__device__ __forceinline__ int point_cell_fast(const GridP &g, const BuildP &bp, const float *p)
{
    int cc[3] = {1, 1, 1};
    bool ok = true;
#pragma unroll
    for (int d = 0; d < ND; d++) {
        cc[d] = cell_coord_fast<PER>(p[d], g.minc[d], g.cs[d], bp.rcs[d], g.nc[d], g.off[d]);
        ok = ok && (cc[d] >= 2) && (cc[d] <= g.gs[d] - 1);
    }
    if (!ok) return -1;
    if (TR) return (cc[2] - 1) + g.gs[2] * ((cc[1] - 1) + g.gs[1] * (cc[0] - 1));
    return (cc[0] - 1) + (cc[1] - 1) * g.gs[0] + (cc[2] - 1) * g.gs[0] * g.gs[1];
}

// Coalesced load of one tile of coordinates into shared memory (float4 when the tile is full and
// the array is 16-byte aligned), then kBuildPPT strided points per thread.
template <int ND>
__device__ __forceinline__ void load_tile(const float *__restrict__ y, int64_t block0, int64_t n,
                                          float *s_xyz)
{
    const int64_t f0 = block0 * ND;
    const int64_t fend = n * ND;
    constexpr int kFloats = kBuildTile * ND;
    if (f0 + kFloats <= fend && ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
        const float4 *src = reinterpret_cast<const float4 *>(y + f0);
        float4 *dst = reinterpret_cast<float4 *>(s_xyz);
#pragma unroll
        for (int t = 0; t < (kFloats / 4 + kBuildThreads - 1) / kBuildThreads; t++) {
            const int q = (int)threadIdx.x + t * kBuildThreads;
            if (q < kFloats / 4) dst[q] = __ldg(src + q);
        }
    } else {
        for (int q = threadIdx.x; q < kFloats; q += kBuildThreads)
            if (f0 + q < fend) s_xyz[q] = __ldg(y + f0 + q);
    }
    __syncthreads();
}

// Runs of equal cells among consecutive lanes (points of a cell-sorted cloud arrive in runs):
// the first lane of a run issues one atomic for the whole run.
//   returns the run head lane; *run_len is valid on the head lane.
__device__ __forceinline__ int run_head(int lin, bool valid, int *run_len)
{
    const int lane = lane_id();
    const int prev = __shfl_up_sync(0xffffffffu, lin, 1);
    const bool head = valid && (lane == 0 || prev != lin);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    const unsigned below = (2u << lane) - 1u;                 // lanes 0..lane (lane 31: all ones)
    const int h = 31 - __clz(heads & below);
    const unsigned ends = (heads | ~act) & ~below;            // next run start or invalid lane
    const int e = ends ? (__ffs(ends) - 1) : 32;
    *run_len = e - lane;                                      // meaningful on the head lane
    return h;
}

// K_a: histogram.  reference: the `lengths[cell] += 1` half of pushat_atomic!
// (src/vector_of_vectors.jl:102) for every point of initialize_grid! (src/nhs_grid.jl:271-278),
// including check_cell_bounds (full_grid.jl:205-213) -> error word.
template <int ND, bool PER>
__global__ void __launch_bounds__(kBuildThreads)
k_cell_hist(GridP g, BuildP bp, const float *__restrict__ y, int64_t n_idx,
            const int32_t *__restrict__ idx, int base, uint32_t *__restrict__ cell_count,
            int *__restrict__ err, int variant, int err_bit)
{
    __shared__ __align__(16) float s_xyz[kBuildTile * ND];
    const int64_t block0 = (int64_t)blockIdx.x * kBuildTile;
    bool bad = false;
    if ((variant & 16) && idx == nullptr && block0 + kBuildTile <= n_idx) {
        // full tile, no staging: lane-strided points, runs of equal cells merged across the
        // lanes of a warp (one RED per run: ~2 per warp on a cell-sorted cloud)
        float p[kBuildPPT][3];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            const int64_t k = block0 + j * kBuildThreads + (int)threadIdx.x;
#pragma unroll
            for (int d = 0; d < 3; d++) p[j][d] = d < ND ? __ldg(y + k * ND + d) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            const int lin = point_cell_fast<ND, PER>(g, bp, p[j]);
            bad = bad || lin < 0;
            int run_len;
            const int h = run_head(lin, lin >= 0, &run_len);
            if (lin >= 0 && h == lane_id()) atomicAdd(cell_count + lin, (unsigned)run_len);
        }
        if (bad) atomicOr(err, err_bit);
        return;
    }
    if ((variant & 1) && idx == nullptr && block0 + kBuildTile <= n_idx &&
        ((reinterpret_cast<uintptr_t>(y) & 15) == 0)) {
        // full tile, no staging: every thread reads its kBuildPPT consecutive points straight
        // from global memory (ND 16-byte loads), no barrier
        float v[kBuildPPT * ND];
        const float4 *gp = reinterpret_cast<const float4 *>(y + block0 * ND) + (int)threadIdx.x * ND;
#pragma unroll
        for (int q = 0; q < ND; q++) {
            const float4 a = __ldg(gp + q);
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
        int lin[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            lin[j] = point_cell_fast<ND, PER>(g, bp, &v[j * ND]);
            bad = bad || lin[j] < 0;
        }
        int cur = lin[0], n = 1;
#pragma unroll
        for (int j = 1; j < kBuildPPT; j++) {
            if (lin[j] == cur) n++;
            else {
                if (cur >= 0) atomicAdd(cell_count + cur, (unsigned)n);
                cur = lin[j];
                n = 1;
            }
        }
        if (cur >= 0) atomicAdd(cell_count + cur, (unsigned)n);
        if (bad) atomicOr(err, err_bit);
        return;
    }
    if (idx == nullptr) load_tile<ND>(y, block0, n_idx, s_xyz);
    if (!(variant & 4) && idx == nullptr && block0 + kBuildTile <= n_idx) {
        // full tile: kBuildPPT CONSECUTIVE points per thread (ND conflict-free LDS.128), runs of
        // equal cells are merged inside the thread: no shuffles, no votes, one RED per run
        float v[kBuildPPT * ND];
        const float4 *sp = reinterpret_cast<const float4 *>(s_xyz) + (int)threadIdx.x * ND;
#pragma unroll
        for (int q = 0; q < ND; q++) {
            const float4 a = sp[q];
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
        int lin[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            lin[j] = point_cell_fast<ND, PER>(g, bp, &v[j * ND]);
            bad = bad || lin[j] < 0;
        }
        int cur = lin[0], n = 1;
#pragma unroll
        for (int j = 1; j < kBuildPPT; j++) {
            if (lin[j] == cur) n++;
            else {
                if (cur >= 0) atomicAdd(cell_count + cur, (unsigned)n);
                cur = lin[j];
                n = 1;
            }
        }
        if (cur >= 0) atomicAdd(cell_count + cur, (unsigned)n);
        if (bad) atomicOr(err, err_bit);
        return;
    }
    // generic path (last tile, `eachindex_y` subsets): strided points, runs merged across lanes
#pragma unroll
    for (int j = 0; j < kBuildPPT; j++) {
        const int loc = j * kBuildThreads + (int)threadIdx.x;
        const int64_t k = block0 + loc;
        const bool in_range = k < n_idx;
        float p[3] = {0.f, 0.f, 0.f};
        if (idx == nullptr) {
#pragma unroll
            for (int d = 0; d < ND; d++) p[d] = s_xyz[loc * ND + d];
        } else if (in_range) {
            const int64_t pt = (int64_t)idx[k] - base;
#pragma unroll
            for (int d = 0; d < ND; d++) p[d] = __ldg(y + pt * ND + d);
        }
        const int lin = in_range ? point_cell_fast<ND, PER>(g, bp, p) : -1;
        const bool valid = lin >= 0;
        bad = bad || (in_range && !valid);
        int run_len;
        const int h = run_head(lin, valid, &run_len);
        if (valid && h == lane_id()) atomicAdd(cell_count + lin, (unsigned)run_len);
    }
    if (bad) atomicOr(err, err_bit);
}

// K_c: scatter.  `cursor` is cell_start + 1 holding the exclusive prefix E[c] at cursor[c]; the
// atomic hands out the slots of cell c and leaves E[c] + count[c] = E[c + 1] behind, i.e. after
// this kernel cell_start[0 .. C] is the finished CSR offset array with no copy.
// (the `backend[new_length, i] = value` of pushat_atomic!, src/vector_of_vectors.jl:109, into a
// CSR of (x, y, z, id) records instead of a 100-row id matrix.)
template <int ND, bool PER>
__global__ void __launch_bounds__(kBuildThreads)
k_scatter_points(GridP g, BuildP bp, const float *__restrict__ y, int64_t n_idx,
                 const int32_t *__restrict__ idx, int base, uint32_t *__restrict__ cursor,
                 float4 *__restrict__ sorted, int variant)
{
    __shared__ __align__(16) float s_xyz[kBuildTile * ND];
    const int64_t block0 = (int64_t)blockIdx.x * kBuildTile;
    if ((variant & 8) && idx == nullptr && block0 + kBuildTile <= n_idx) {
        // full tile, no staging, no barrier: lane-strided points read with scalar loads (the
        // three loads of a warp cover the same 384 contiguous bytes), all atomics of the
        // kBuildPPT points are issued before the first record is stored
        float p[kBuildPPT][3];
        int lin[kBuildPPT], h[kBuildPPT], rl[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            const int64_t k = block0 + j * kBuildThreads + (int)threadIdx.x;
#pragma unroll
            for (int d = 0; d < 3; d++) p[j][d] = d < ND ? __ldg(y + k * ND + d) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            lin[j] = point_cell_fast<ND, PER>(g, bp, p[j]);
            h[j] = run_head(lin[j], lin[j] >= 0, &rl[j]);
        }
        unsigned basev[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++)
            basev[j] = (lin[j] >= 0 && h[j] == lane_id()) ? atomicAdd(cursor + lin[j], (unsigned)rl[j]) : 0u;
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            const unsigned b = __shfl_sync(0xffffffffu, basev[j], h[j] & 31);
            const int32_t id = (int32_t)(block0 + j * kBuildThreads + (int)threadIdx.x);
            if (lin[j] >= 0)
                sorted[b + (unsigned)(lane_id() - h[j])] = make_float4(p[j][0], p[j][1], p[j][2], __int_as_float(id));
        }
        return;
    }
    if (idx == nullptr) load_tile<ND>(y, block0, n_idx, s_xyz);
    if (!(variant & 2) && idx == nullptr && block0 + kBuildTile <= n_idx) {
        // full tile: consecutive points per thread, thread-local runs (see k_cell_hist); the
        // atomics of all runs are issued before any record is stored
        float v[kBuildPPT * ND];
        const float4 *sp = reinterpret_cast<const float4 *>(s_xyz) + (int)threadIdx.x * ND;
#pragma unroll
        for (int q = 0; q < ND; q++) {
            const float4 a = sp[q];
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
        }
        int lin[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) lin[j] = point_cell_fast<ND, PER>(g, bp, &v[j * ND]);
        // run lengths (a run = consecutive points of one cell), heads take the cursor once
        int rl[kBuildPPT];
        rl[kBuildPPT - 1] = 1;
#pragma unroll
        for (int j = kBuildPPT - 2; j >= 0; j--) rl[j] = (lin[j] == lin[j + 1]) ? rl[j + 1] + 1 : 1;
        unsigned slot[kBuildPPT];
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            const bool is_head = (j == 0) || (lin[j] != lin[j - 1]);
            slot[j] = (is_head && lin[j] >= 0) ? atomicAdd(cursor + lin[j], (unsigned)rl[j]) : 0u;
        }
#pragma unroll
        for (int j = 1; j < kBuildPPT; j++)
            if (lin[j] == lin[j - 1]) slot[j] = slot[j - 1] + 1u;
        const int32_t id0 = (int32_t)(block0 + (int64_t)threadIdx.x * kBuildPPT);
#pragma unroll
        for (int j = 0; j < kBuildPPT; j++) {
            if (lin[j] >= 0)
                sorted[slot[j]] = make_float4(v[j * ND], ND > 1 ? v[j * ND + 1] : 0.f,
                                              ND > 2 ? v[j * ND + 2] : 0.f, __int_as_float(id0 + j));
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < kBuildPPT; j++) {
        const int loc = j * kBuildThreads + (int)threadIdx.x;
        const int64_t k = block0 + loc;
        const bool in_range = k < n_idx;
        float p[3] = {0.f, 0.f, 0.f};
        int32_t id = (int32_t)k;
        if (idx == nullptr) {
#pragma unroll
            for (int d = 0; d < ND; d++) p[d] = s_xyz[loc * ND + d];
        } else if (in_range) {
            id = idx[k] - base;
#pragma unroll
            for (int d = 0; d < ND; d++) p[d] = __ldg(y + (int64_t)id * ND + d);
        }
        const int lin = in_range ? point_cell_fast<ND, PER>(g, bp, p) : -1;
        const bool valid = lin >= 0;
        int run_len;
        const int h = run_head(lin, valid, &run_len);
        unsigned basev = 0;
        if (valid && h == lane_id()) basev = atomicAdd(cursor + lin, (unsigned)run_len);
        basev = __shfl_sync(0xffffffffu, basev, h & 31);
        if (valid)
            sorted[basev + (unsigned)(lane_id() - h)] = make_float4(p[0], p[1], p[2], __int_as_float(id));
    }
}

// Canonical order: one warp per cell sorts the cell's records by point id (rank by counting) and
// writes them to `out` plus the id list `cell_points`.  Cells with <= 32 points (the normal
// case) stay in registers.
__global__ void __launch_bounds__(256)
k_canonicalize(int total_cells, const uint32_t *__restrict__ cell_start,
               const float4 *__restrict__ in, float4 *__restrict__ out,
               int32_t *__restrict__ cell_points)
{
    int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= total_cells) return;
    const int lane = lane_id();
    const uint32_t s0 = cell_start[c], s1 = cell_start[c + 1];
    const int cnt = (int)(s1 - s0);
    if (cnt == 0) return;
    if (cnt <= 32) {
        float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
        int v = 0x7fffffff;
        if (lane < cnt) { rec = in[s0 + lane]; v = __float_as_int(rec.w); }
        int r = 0;
        for (int k = 0; k < cnt; k++) r += (__shfl_sync(0xffffffffu, v, k) < v) ? 1 : 0;
        if (lane < cnt) {
            cell_points[s0 + r] = v;
            out[s0 + r] = rec;
        }
    } else {
        for (int e = lane; e < cnt; e += 32) {
            const float4 rec = in[s0 + e];
            const int v = __float_as_int(rec.w);
            int r = 0;
            for (int k = 0; k < cnt; k++) r += (__float_as_int(in[s0 + k].w) < v) ? 1 : 0;
            cell_points[s0 + r] = v;
            out[s0 + r] = rec;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// single-pass exclusive scan with decoupled look-back (Merrill & Garland), uint32 input
//   tile = 512 threads x 16 items; status word = flag(2) | epoch(14) | value(48), flag 1 = tile
//   aggregate, 2 = inclusive prefix.  A word counts only if its epoch is the current launch's,
//   so the status array is never cleared between launches; tiles are taken from an atomic
//   ticket (reset by the last tile) so a tile never waits on a tile that has not started.
//   ZERO_IN: the input is cleared after it has been read (the histogram is ready for the next
//   build without a memset).
// Replaces nothing in the reference (its 100-row cell matrix needs no offsets): this is what
// turns the histogram into CSR offsets.  HBM: reads 4 B/cell, writes 4 (+4) B/cell.
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 16;
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

constexpr unsigned long long kScanValMask = (1ULL << 48) - 1ULL;
constexpr int kScanEpochs = 1 << 14;

// flag of a status word as seen by a launch with epoch `ep` (0 = not published yet)
__device__ __forceinline__ unsigned scan_flag(unsigned long long sv, unsigned ep)
{
    return (((unsigned)(sv >> 48)) & (kScanEpochs - 1)) == ep ? (unsigned)(sv >> 62) : 0u;
}

template <typename OutT, bool ZERO_IN, bool WRITE_TOTAL>
__global__ void __launch_bounds__(kScanThreads)
k_scan_lookback(uint32_t *__restrict__ in, OutT *__restrict__ out, int64_t n,
                unsigned long long *__restrict__ status, unsigned int *__restrict__ ticket,
                unsigned epoch)
{
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_warp[kScanThreads / 32];
    __shared__ unsigned long long s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const int64_t base = (int64_t)tile * kScanTile + (int64_t)threadIdx.x * kScanItems;
    const unsigned long long ep_bits = (unsigned long long)epoch << 48;

    uint32_t v[kScanItems];
    if (base + kScanItems <= n && ((reinterpret_cast<uintptr_t>(in) & 15) == 0)) {
        uint4 *src = reinterpret_cast<uint4 *>(in + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++) {
            const uint4 a = src[q];
            v[4 * q] = a.x; v[4 * q + 1] = a.y; v[4 * q + 2] = a.z; v[4 * q + 3] = a.w;
            if (ZERO_IN) src[q] = make_uint4(0u, 0u, 0u, 0u);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            v[k] = (base + k < n) ? in[base + k] : 0u;
            if (ZERO_IN && base + k < n) in[base + k] = 0u;
        }
    }
    // a tile holds 8192 items; callers guarantee the sum of one tile fits 32 bits
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
            if (lane == 0) st_status(status, (2ULL << 62) | ep_bits | block_agg);
        } else {
            if (lane == 0) st_status(status + tile, (1ULL << 62) | ep_bits | block_agg);
            int64_t look = (int64_t)tile - 1;
            while (true) {
                const int64_t t = look - lane;
                const bool have = t >= 0;
                unsigned long long sv = 0;
                unsigned fl;
                do {
                    if (have) { sv = ld_status(status + t); fl = scan_flag(sv, epoch); }
                    else fl = 2u;
                } while (__any_sync(0xffffffffu, fl == 0u));
                const unsigned incl_mask = __ballot_sync(0xffffffffu, fl == 2u);
                const int first = incl_mask ? (__ffs(incl_mask) - 1) : 32;
                unsigned long long contrib = (have && lane <= first) ? (sv & kScanValMask) : 0ULL;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                excl += contrib;
                if (incl_mask) break;
                look -= 32;
            }
            if (lane == 0)
                st_status(status + tile, (2ULL << 62) | ep_bits | ((excl + block_agg) & kScanValMask));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    unsigned long long run = s_prefix + thread_excl;
    const bool owns_last = base < n && base + kScanItems >= n;
    if (base + kScanItems <= n && sizeof(OutT) == 4 &&
        ((reinterpret_cast<uintptr_t>(out + base) & 15) == 0)) {
        uint32_t o32[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; k++) { o32[k] = (uint32_t)run; run += v[k]; }
        uint4 *dst = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int q = 0; q < kScanItems / 4; q++)
            dst[q] = make_uint4(o32[4 * q], o32[4 * q + 1], o32[4 * q + 2], o32[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (base + k < n) out[base + k] = (OutT)run;
            run += (base + k < n) ? v[k] : 0u;
        }
    }
    // the thread that owns the last element also writes the total at out[n]
    if (WRITE_TOTAL && owns_last) out[n] = (OutT)run;
    // every ticket of this launch has been handed out once the last one is seen
    if (tile == gridDim.x - 1 && threadIdx.x == 0) *ticket = 0u;
}

template <typename OutT, bool ZERO_IN, bool WRITE_TOTAL>
static pnb_status scan_impl(pnb_grid *g, uint32_t *in, OutT *out, int64_t n, cudaStream_t s)
{
    if (n <= 0) {
        if (WRITE_TOTAL) PNB_CUDA(cudaMemsetAsync(out, 0, sizeof(OutT), s));
        return PNB_OK;
    }
    int64_t tiles = div_up(n, kScanTile);
    if (tiles > g->scan_tiles_cap) {
        if (g->scan_status) PNB_CUDA(cudaFree(g->scan_status));
        g->scan_status = nullptr;
        g->scan_tiles_cap = 0;
        PNB_CUDA(cudaMalloc(&g->scan_status, sizeof(unsigned long long) * (size_t)tiles));
        g->scan_tiles_cap = tiles;
        g->scan_epoch = 0;
    }
    ProfScope ps(PH_BUILD_SCAN, s);
    if (g->scan_epoch == 0 || g->scan_epoch + 1 >= kScanEpochs) {
        // first use or epoch wrap: clear every status word once (epoch 0 is never current)
        PNB_CUDA(cudaMemsetAsync(g->scan_status, 0,
                                 sizeof(unsigned long long) * (size_t)g->scan_tiles_cap, s));
        PNB_CUDA(cudaMemsetAsync(g->scan_ticket, 0, sizeof(unsigned int), s));
        g->scan_epoch = 0;
    }
    g->scan_epoch++;
    k_scan_lookback<OutT, ZERO_IN, WRITE_TOTAL><<<(unsigned)tiles, kScanThreads, 0, s>>>(
        in, out, n, g->scan_status, g->scan_ticket, (unsigned)g->scan_epoch);
    PNB_LAUNCHED();
    return PNB_OK;
}

pnb_status exclusive_scan_u32(pnb_grid *g, const uint32_t *in, uint32_t *out, int64_t n,
                              cudaStream_t s)
{
    return scan_impl<uint32_t, false, true>(g, const_cast<uint32_t *>(in), out, n, s);
}

pnb_status exclusive_scan_u32_to_i64(pnb_grid *g, const uint32_t *in, int64_t *out, int64_t n,
                                     cudaStream_t s)
{
    return scan_impl<int64_t, false, true>(g, const_cast<uint32_t *>(in), out, n, s);
}

// ---------------------------------------------------------------------------------------------
// one-pass update!: every cell owns K record slots ("buckets", the reference's own idea --
// its cell matrix is 100 x C, src/cell_lists/full_grid.jl:74-78 -- with K chosen from the data).
// One kernel: read the coordinates (12 N), cell index, the run head takes the cell's counter
// (ATOM.ADD returns the first free slot), records stored at cell * K + slot (16 N).  No
// histogram pass, no scan: 28 N + 4 C bytes, exactly the algorithmic minimum of SURVEY 8d.
// A cell that receives more than K points sets error bit 3; the build is then redone as CSR.
// ---------------------------------------------------------------------------------------------
// DIAG (measurement only, results invalid, instantiated only when the library is compiled with
// -DPNB_DIAG for tools/bucket_diag.py -- never in the shipped binary): 1 = non-returning atomics
// (slot from the lane run alone), 2 = no stores, 4 = no atomics at all.  DIAG 8 (valid results, the default of the launch
// below): in full tiles the lanes of a warp are grouped by cell with match.any instead of by runs
// of adjacent lanes -- one atomic and one contiguous piece of the bucket per distinct cell
// (0.1354 -> 0.1316 ms on config 3).
template <int ND, bool PER, int PPT, int DIAG = 0, bool TR = false>
__global__ void __launch_bounds__(kBuildThreads, 6)
k_bucket_scatter(GridP g, BuildP bp, const float *__restrict__ y, int64_t n_idx,
                 const int32_t *__restrict__ idx, int base, uint32_t K,
                 uint32_t *__restrict__ bcount, float4 *__restrict__ brec, int *__restrict__ err,
                 uint4 *__restrict__ clear4, int n_clear4)
{
    // the counters of the NEXT build (the other array) are cleared on the side: no memset launch
    for (int64_t i = (int64_t)blockIdx.x * kBuildThreads + threadIdx.x; i < n_clear4;
         i += (int64_t)gridDim.x * kBuildThreads)
        clear4[i] = make_uint4(0u, 0u, 0u, 0u);
    const int64_t block0 = (int64_t)blockIdx.x * (kBuildThreads * PPT);
    const int logK = 31 - __clz((int)K);            // K is a power of two
    if (idx == nullptr && block0 + (kBuildThreads * PPT) <= n_idx) {
        // full tile: no bounds checks, all atomics of a thread's points before the first store
        float p[PPT][3];
        int lin[PPT], h[PPT], rl[PPT], rk[PPT];
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            const int64_t k = block0 + j * kBuildThreads + (int)threadIdx.x;
#pragma unroll
            for (int d = 0; d < 3; d++) p[j][d] = d < ND ? __ldg(y + k * ND + d) : 0.f;
        }
        int bad = 0;
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            lin[j] = point_cell_fast<ND, PER, TR>(g, bp, p[j]);
            if (lin[j] < 0) bad |= 1;
            if (DIAG & 8) {
                // all lanes of the warp that hit the same cell form one group, adjacent or not:
                // one atomic and one contiguous piece of the bucket per distinct cell
                const unsigned m = __match_any_sync(0xffffffffu, lin[j]);
                h[j] = __ffs(m) - 1;
                rl[j] = __popc(m);
                rk[j] = __popc(m & ((1u << lane_id()) - 1u));
            } else {
                h[j] = run_head(lin[j], lin[j] >= 0, &rl[j]);
                rk[j] = lane_id() - h[j];
            }
        }
        unsigned basev[PPT];
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            if (DIAG & 4) basev[j] = 0u;
            else if (DIAG & 1) {
                basev[j] = 0u;
                if (lin[j] >= 0 && h[j] == lane_id()) atomicAdd(bcount + lin[j], (unsigned)rl[j]);
            } else
                basev[j] = (lin[j] >= 0 && h[j] == lane_id()) ? atomicAdd(bcount + lin[j], (unsigned)rl[j]) : 0u;
        }
#pragma unroll
        for (int j = 0; j < PPT; j++) {
            const unsigned b = __shfl_sync(0xffffffffu, basev[j], h[j] & 31);
            if (lin[j] >= 0) {
                unsigned slot = b + (unsigned)rk[j];
                const int32_t id = (int32_t)(block0 + j * kBuildThreads + (int)threadIdx.x);
                if (DIAG & 5) slot = (slot + (unsigned)id) & 31u;
                if (DIAG & 2) { if (slot == 0xffffffffu) bad |= 8; }
                else if (slot < K)
                    brec[((size_t)lin[j] << logK) + slot] = make_float4(p[j][0], p[j][1], p[j][2], __int_as_float(id));
                else
                    bad |= 8;
            }
        }
        if (bad) atomicOr(err, bad);
        return;
    }
    float p[PPT][3];
    int32_t id[PPT];
    bool in[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        const int64_t k = block0 + j * kBuildThreads + (int)threadIdx.x;
        in[j] = k < n_idx;
        id[j] = (int32_t)k;
        if (in[j] && idx) id[j] = idx[k] - base;
#pragma unroll
        for (int d = 0; d < 3; d++) p[j][d] = (in[j] && d < ND) ? __ldg(y + (int64_t)id[j] * ND + d) : 0.f;
    }
    int lin[PPT], h[PPT], rl[PPT];
    int bad = 0;
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        lin[j] = in[j] ? point_cell_fast<ND, PER, TR>(g, bp, p[j]) : -1;
        if (in[j] && lin[j] < 0) bad |= 1;
        h[j] = run_head(lin[j], lin[j] >= 0, &rl[j]);
    }
    unsigned basev[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++)
        basev[j] = (lin[j] >= 0 && h[j] == lane_id()) ? atomicAdd(bcount + lin[j], (unsigned)rl[j]) : 0u;
#pragma unroll
    for (int j = 0; j < PPT; j++) {
        const unsigned b = __shfl_sync(0xffffffffu, basev[j], h[j] & 31);
        if (lin[j] >= 0) {
            const unsigned slot = b + (unsigned)(lane_id() - h[j]);
            if (slot < K)
                brec[((size_t)lin[j] << logK) + slot] = make_float4(p[j][0], p[j][1], p[j][2], __int_as_float(id[j]));
            else
                bad |= 8;
        }
    }
    if (bad) atomicOr(err, bad);
}

// buckets -> CSR records (one warp per cell), after the scan of the bucket counts
__global__ void __launch_bounds__(256)
k_bucket_to_csr(int total_cells, uint32_t K, const uint32_t *__restrict__ bcount,
                const float4 *__restrict__ brec, const uint32_t *__restrict__ cell_start,
                float4 *__restrict__ sorted, uint32_t t0, uint32_t t1, uint32_t t2)
{
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= total_cells) return;
    const uint32_t lb = bucket_of((uint32_t)c, t0, t1, t2);
    const uint32_t cnt = bcount[lb], s0 = cell_start[c];
    for (uint32_t e = lane_id(); e < cnt; e += 32) sorted[s0 + e] = brec[(size_t)lb * K + e];
}

// bucket counts in linear cell order (input of the CSR scan when the buckets are transposed)
__global__ void k_bucket_counts_linear(int64_t total_cells, const uint32_t *__restrict__ bcount,
                                       uint32_t *__restrict__ out, uint32_t t0, uint32_t t1,
                                       uint32_t t2)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < total_cells) out[c] = bcount[bucket_of((uint32_t)c, t0, t1, t2)];
}

// Which dimension does the ORDER of the points step through?  Over a sample of pairs (k, k + 64)
// -- far enough apart that the systematic drift outweighs the jitter of clouds that are only
// approximately cell-sorted (test/point_cloud.jl:26-50 perturbs again after computing the sort
// keys) --: out[1] += cells moved in the first dimension, out[2] += cells moved in the last one
// (each clipped to 8: the jump at the end of a row / column says nothing).
constexpr int kProbeLag = 64;
template <int ND>
__global__ void k_order_probe(GridP g, const float *__restrict__ y, int64_t n_pairs, int64_t stride,
                              unsigned int *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int64_t k = t * stride;
    float a[3] = {0.f, 0.f, 0.f}, b[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < ND; d++) { a[d] = __ldg(y + k * ND + d); b[d] = __ldg(y + (k + kProbeLag) * ND + d); }
    int ca[3], cb[3];
    if (point_cell<ND>(g, a, ca) < 0 || point_cell<ND>(g, b, cb) < 0) return;
    const unsigned first = (unsigned)min(abs(ca[0] - cb[0]), 8);
    const unsigned last = (unsigned)min(abs(ca[ND - 1] - cb[ND - 1]), 8);
    if (first) atomicAdd(out + 1, first);
    if (last) atomicAdd(out + 2, last);
}

// fullest cell of a CSR cell list -> out[0]
__global__ void k_max_cell_count(int64_t n_cells, const uint32_t *__restrict__ start,
                                 unsigned int *__restrict__ out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned v = c < n_cells ? start[c + 1] - start[c] : 0u;
    v = __reduce_max_sync(0xffffffffu, v);
    if (lane_id() == 0 && v > 0u) atomicMax(out, v);
}

pnb_status ensure_point_capacity(pnb_grid *g, int64_t n)
{
    if (n <= g->cap_points) return PNB_OK;
    cudaFree(g->cell_points); g->cell_points = nullptr;
    cudaFree(g->sorted); g->sorted = nullptr;
    cudaFree(g->sorted_alt); g->sorted_alt = nullptr;
    g->cap_points = 0;
    int64_t cap = n + n / 16 + 32;
    PNB_CUDA(cudaMalloc(&g->sorted, sizeof(float4) * (size_t)cap));
    g->cap_points = cap;
    return PNB_OK;
}

// Sort every cell by point id (once per build, on demand): the CSR id list `cell_points` and
// the cell-ordered records become reproducible (ids ascending inside a cell), which is what the
// exports, the exact arithmetic mode and the neighbour-list fills are specified against.
// CSR offsets and records from the bucket layout (consumers other than the tile kernels: the
// exports, the ordered / per-point sweeps, the canonical order).
pnb_status ensure_csr(pnb_grid *g, cudaStream_t s)
{
    { pnb_status sp = resolve_pending(g); if (sp != PNB_OK) return sp; }
    if (g->csr_valid || !g->built || !g->bucket_valid) return PNB_OK;
    const int64_t C = g->p.total_cells;
    const uint32_t t0 = g->bucket_tr ? (uint32_t)g->p.gs[0] : 0u;
    const uint32_t t1 = (uint32_t)g->p.gs[1], t2 = (uint32_t)g->p.gs[2];
    // exclusive prefix of the bucket counts (in linear cell order) = CSR offsets, total at cell_start[C]
    pnb_status st;
    if (g->bucket_tr && C > 0) {
        k_bucket_counts_linear<<<(unsigned)div_up(C, 256), 256, 0, s>>>(C, g->bcount, g->cell_count,
                                                                       t0, t1, t2);
        PNB_LAUNCHED();
        st = scan_impl<uint32_t, true, true>(g, g->cell_count, g->cell_start, C, s);   // re-zeroes the scratch
    } else {
        st = scan_impl<uint32_t, false, true>(g, g->bcount, g->cell_start, C, s);
    }
    if (st != PNB_OK) return st;
    if (C > 0 && g->n_built > 0) {
        ProfScope ps(PH_BUILD_FINALIZE, s);
        k_bucket_to_csr<<<(unsigned)div_up(C * 32, 256), 256, 0, s>>>(
            (int)C, (uint32_t)g->bucket_K, g->bcount, g->brec, g->cell_start, g->sorted, t0, t1, t2);
        PNB_LAUNCHED();
    }
    g->csr_valid = true;
    return PNB_OK;
}

pnb_status ensure_canonical(pnb_grid *g, cudaStream_t s)
{
    { pnb_status sp = resolve_pending(g); if (sp != PNB_OK) return sp; }
    if (g->canonical || !g->built || g->template_search || g->n_built == 0) return PNB_OK;
    {
        pnb_status stc = ensure_csr(g, s);
        if (stc != PNB_OK) return stc;
        // the canonical order lives in the CSR arrays: the tile kernels switch to them
        g->bucket_valid = false;
    }
    if (!g->sorted_alt) PNB_CUDA(cudaMalloc(&g->sorted_alt, sizeof(float4) * (size_t)g->cap_points));
    if (!g->cell_points) PNB_CUDA(cudaMalloc(&g->cell_points, sizeof(int32_t) * (size_t)g->cap_points));
    const int64_t C = g->p.total_cells;
    {
        ProfScope ps(PH_BUILD_FINALIZE, s);
        k_canonicalize<<<(unsigned)div_up(C * 32, 256), 256, 0, s>>>((int)C, g->cell_start, g->sorted,
                                                                   g->sorted_alt, g->cell_points);
        PNB_LAUNCHED();
    }
    float4 *t = g->sorted; g->sorted = g->sorted_alt; g->sorted_alt = t;
    g->canonical = true;
    return PNB_OK;
}

// The counting sort into (start[0 .. C], records): used for the cell list itself (y) and for the
// cell-ordered copy of a second point set (x of a two-set sweep, build_query_list).
template <int ND, bool PER>
static pnb_status sort_into(pnb_grid *g, const float *y, const int32_t *idx, int64_t n_idx,
                            int base, uint32_t *start, float4 *records, int err_bit,
                            cudaStream_t s)
{
    const int64_t C = g->p.total_cells;
    BuildP bp;
    for (int d = 0; d < 3; d++) { volatile float rc = 1.0f / g->p.cs[d]; bp.rcs[d] = rc; }
    const unsigned blocks = (unsigned)div_up(n_idx, kBuildTile);
    // cell_count is all zero here: cleared at creation and by every scan (ZERO_IN)
    if (n_idx > 0) {
        ProfScope ps(PH_BUILD_CELL_COUNT, s);
        k_cell_hist<ND, PER><<<blocks, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, idx, base,
                                                              g->cell_count, g->d_err, g_tune_build,
                                                              err_bit);
        PNB_LAUNCHED();
    }
    // exclusive prefix E[c] -> start[c + 1]; start[0] stays 0 (set at allocation)
    pnb_status st = scan_impl<uint32_t, true, false>(g, g->cell_count, start + 1, C, s);
    if (st != PNB_OK) return st;
    if (n_idx > 0) {
        ProfScope ps(PH_BUILD_SCATTER, s);
        k_scatter_points<ND, PER><<<blocks, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, idx, base,
                                                                   start + 1, records, g_tune_build);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

template <int ND, bool PER>
static pnb_status build_nd(pnb_grid *g, const float *y, int64_t n, const int32_t *idx,
                           int64_t n_idx, int base, cudaStream_t s)
{
    (void)n;
    return sort_into<ND, PER>(g, y, idx, n_idx, base, g->cell_start, g->sorted, 1, s);
}

// Cell-ordered copy of the QUERY points of a two-set sweep (x != y): the same counting sort with
// the grid's cell arithmetic into xq_start / xq_sorted.  A query point outside the valid cells
// 2 .. size-1 has a stencil that leaves the grid: the safe variant's BoundsError
// (src/nhs_grid.jl:530-532) -> error bit 2.
__global__ void k_count_nonempty(int64_t n_cells, const uint32_t *__restrict__ start,
                                 unsigned int *__restrict__ out)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ne = c < n_cells && start[c + 1] != start[c];
    const unsigned m = __ballot_sync(0xffffffffu, ne);
    if (lane_id() == 0 && m) atomicAdd(out, (unsigned)__popc(m));
}

pnb_status build_query_list(pnb_grid *g, const float *x, int64_t nx, double *points_per_cell,
                            cudaStream_t s)
{
    const int64_t C = g->p.total_cells;
    if (!g->xq_start_alloc) {
        PNB_CUDA(cudaMalloc(&g->xq_start_alloc, sizeof(uint32_t) * (size_t)(C + 8)));
        PNB_CUDA(cudaMemsetAsync(g->xq_start_alloc, 0, sizeof(uint32_t) * (size_t)(C + 8), s));
        g->xq_start = g->xq_start_alloc + 3;
    }
    if (nx > g->xq_cap) {
        cudaFree(g->xq_sorted);
        g->xq_sorted = nullptr;
        g->xq_cap = 0;
        const int64_t cap = nx + nx / 16 + 32;
        PNB_CUDA(cudaMalloc(&g->xq_sorted, sizeof(float4) * (size_t)cap));
        g->xq_cap = cap;
    }
    const bool per = g->p.periodic != 0;
    pnb_status st;
    switch (g->p.ndims) {
        case 1: st = per ? sort_into<1, true>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s)
                         : sort_into<1, false>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s); break;
        case 2: st = per ? sort_into<2, true>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s)
                         : sort_into<2, false>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s); break;
        default: st = per ? sort_into<3, true>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s)
                          : sort_into<3, false>(g, x, nullptr, nx, 0, g->xq_start, g->xq_sorted, 2, s); break;
    }
    if (st != PNB_OK) return st;
    // query points per occupied cell: the tile kernel gives a lane to every point of a cell, so
    // it only pays off when the occupied cells are reasonably full (the caller decides)
    // (counter in device memory: thousands of atomics on the mapped host word would crawl)
    unsigned int *d_ne = g->xq_start_alloc;      // word 0 of the allocation, below xq_start[0]
    PNB_CUDA(cudaMemsetAsync(d_ne, 0, sizeof(unsigned int), s));
    k_count_nonempty<<<(unsigned)div_up(C, 256), 256, 0, s>>>(C, g->xq_start, d_ne);
    PNB_LAUNCHED();
    PNB_CUDA(cudaMemcpyAsync(g->h_err + 1, d_ne, sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    const unsigned ne = *(volatile unsigned int *)(g->h_err + 1);
    *points_per_cell = ne ? (double)nx / (double)ne : 0.0;
    return PNB_OK;
}

}  // namespace pnb

extern "C" void pnb_set_build_tuning(int variant) { pnb::g_tune_build = variant; }
extern "C" void pnb_set_build_layout(int buckets) { pnb::g_build_layout = buckets; }
extern "C" void pnb_set_bucket_order(int order) { pnb::g_bucket_order = order; }

// set while pnb_grid_build_async_f32 runs the build below
static thread_local bool t_async_build = false;

extern "C" pnb_status pnb_grid_build_f32(pnb_grid *g, const float *y, int64_t n,
                                         const int32_t *eachindex_y, int64_t n_idx, int index_base,
                                         void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (g->f64) { set_error("Float64 grid handle passed to a Float32 entry point"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t C = g->p.total_cells;
    if (eachindex_y == nullptr) n_idx = n;
    if (g->async_pending && t_async_build && g->async_stream == s && g->bucket_K > 0 &&
        eachindex_y == nullptr && n > 0) {
        // stream-ordered update! after stream-ordered update!: the error word keeps collecting
        // bits until somebody looks (settle_async_build knows the length of the chain)
        g->async_pending = false;
    }
    if (g->async_pending) {
        // a stream-ordered update! nobody looked at: its error word must not leak into this build
        PNB_CUDA(cudaStreamSynchronize(g->async_stream));
        g->async_pending = false;
        const int e = *(volatile int *)g->h_err;
        *(volatile int *)g->h_err = 0;
        if (e & 8) { g->bucket_K = 0; g->bcount_alt_clean = false; }
        if (e & 1) { set_error("%s", kDomainMsg); g->n_built = 0; g->built = false; return PNB_ERR_DOMAIN; }
    }
    g->built = false;
    g->y_refreshed = false;
    g->canonical = false;
    // empty!(cell_list)  (src/cell_lists/full_grid.jl:96-105)
    if (g->template_search) {
        // nhs_grid.jl:263-267: zero search radius -> emptied cell list, nothing else
        PNB_CUDA(cudaMemsetAsync(g->cell_start, 0, sizeof(uint32_t) * (size_t)(C + 1), s));
        if (g->hashed) PNB_CUDA(cudaMemsetAsync(g->hmeta, 0, sizeof(int4) * (size_t)C, s));
        PNB_CUDA(cudaStreamSynchronize(s));
        g->n_built = 0;
        g->built = true;
        g->y_built = y;
        g->n_y_built = n;
        g->full_build = false;
        g->bucket_valid = false;
        g->csr_valid = true;
        return PNB_OK;
    }
    if (n_idx > 0x7ffffff0LL || n > 0x7ffffff0LL) {
        set_error("more than 2^31 points are not supported (ids are Int32, full_grid.jl:177)");
        return PNB_ERR_ARG;
    }
    if (n_idx > 0 && y == nullptr) { set_error("y is NULL"); return PNB_ERR_ARG; }
    if (g->hashed) return hash_build(g, y, n, eachindex_y, n_idx, index_base, s);
    pnb_status st = ensure_point_capacity(g, n_idx);
    if (st != PNB_OK) return st;
    g->bucket_valid = false;
    g->csr_valid = false;
    // ---- one-pass build into the bucket layout (K chosen by the previous CSR build) -----------
    if (g_build_layout != 0 && g->bucket_K > 0 && n_idx > 0 && C > 0) {
        const int64_t slots = C * (int64_t)g->bucket_K;
        if (slots > g->brec_slots) {
            cudaFree(g->brec);
            g->brec = nullptr;
            g->brec_slots = 0;
            PNB_CUDA(cudaMalloc(&g->brec, sizeof(float4) * (size_t)slots));
            g->brec_slots = slots;
        }
        if (!g->bcount) {
            PNB_CUDA(cudaMalloc(&g->bcount, sizeof(uint32_t) * (size_t)(C + 8)));
            PNB_CUDA(cudaMalloc(&g->bcount_alt, sizeof(uint32_t) * (size_t)(C + 8)));
            g->bcount_alt_clean = false;
        }
        BuildP bp;
        for (int d = 0; d < 3; d++) { volatile float rc = 1.0f / g->p.cs[d]; bp.rcs[d] = rc; }
        {
            ProfScope ps(PH_BUILD_BUCKET, s);     // the clearing of the counters is part of it
            // two counter arrays: this build counts into the one the previous build cleared and
            // clears the other one for the next build
            if (!g->bcount_alt_clean)
                PNB_CUDA(cudaMemsetAsync(g->bcount_alt, 0, sizeof(uint32_t) * (size_t)(C + 8), s));
            { uint32_t *t = g->bcount; g->bcount = g->bcount_alt; g->bcount_alt = t; }
            g->bcount_alt_clean = true;
            uint4 *clear4 = reinterpret_cast<uint4 *>(g->bcount_alt);
            const int n_clear4 = (int)((C + 3) / 4);
#define PNB_BUCKET(ND, PER, PPT)                                                                   \
    do {                                                                                           \
        if (g->bucket_tr)                                                                          \
            k_bucket_scatter<ND, PER, PPT, 8, true><<<(unsigned)div_up(n_idx, kBuildThreads * PPT), \
                kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base,                \
                                       (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4);       \
        else                                                                                       \
            k_bucket_scatter<ND, PER, PPT, 8><<<(unsigned)div_up(n_idx, kBuildThreads * PPT),      \
                kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base,                \
                                       (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4);       \
    } while (0)
            const bool per = g->p.periodic != 0;
            switch (g->p.ndims) {
                case 1: if (per) PNB_BUCKET(1, true, 4); else PNB_BUCKET(1, false, 4); break;
                case 2: if (per) PNB_BUCKET(2, true, 4); else PNB_BUCKET(2, false, 4); break;
                default:
                    if (per) PNB_BUCKET(3, true, 4);
                    else if (g_tune_build & 32) PNB_BUCKET(3, false, 8);
                    else if (g_tune_build & 64) PNB_BUCKET(3, false, 2);
                    else if (g_tune_build & 128) PNB_BUCKET(3, false, 1);
                    else if (g_tune_build & 2048) {
                        // runs of adjacent lanes instead of match.any lane groups (the version
                        // before; valid results, kept for tools/match_diag.py)
                        if (g->bucket_tr)
                            k_bucket_scatter<3, false, 4, 0, true><<<(unsigned)div_up(n_idx, kBuildThreads * 4), kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4);
                        else
                            k_bucket_scatter<3, false, 4, 0><<<(unsigned)div_up(n_idx, kBuildThreads * 4), kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4);
                    }
#ifdef PNB_DIAG   /* tools/ builds only: these variants write garbage layouts */
                    else if ((g_tune_build >> 8) & 7) {
                        // measurement only (DIAG variants of the kernel; the layout is garbage)
                        const unsigned nb = (unsigned)div_up(n_idx, kBuildThreads * 4);
                        switch ((g_tune_build >> 8) & 7) {
                            case 1: k_bucket_scatter<3, false, 4, 1><<<nb, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4); break;
                            case 2: k_bucket_scatter<3, false, 4, 2><<<nb, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4); break;
                            case 3: k_bucket_scatter<3, false, 4, 3><<<nb, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4); break;
                            case 4: k_bucket_scatter<3, false, 4, 4><<<nb, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4); break;
                            default: k_bucket_scatter<3, false, 4, 6><<<nb, kBuildThreads, 0, s>>>(g->p, bp, y, n_idx, eachindex_y, index_base, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err, clear4, n_clear4); break;
                        }
                    }
#endif
                    else PNB_BUCKET(3, false, 4);
                    break;
            }
#undef PNB_BUCKET
            PNB_LAUNCHED();
        }
        if (t_async_build) {
            // stream-ordered update!: the error word is looked at by the next blocking call
            g->bucket_valid = true;
            g->n_built = n_idx;
            g->y_built = y;
            g->n_y_built = n;
            g->full_build = (eachindex_y == nullptr);
            g->built = true;
            g->async_pending = true;
            g->async_stream = s;
            g->async_chain++;
            return PNB_OK;
        }
        PNB_CUDA(cudaStreamSynchronize(s));     // initialize!/update! are blocking calls
        const int e = *(volatile int *)g->h_err;
        if ((e & 8) == 0) {
            st = check_err_word(g, s);
            if (st != PNB_OK) { g->n_built = 0; return st; }
            g->bucket_valid = true;
            g->n_built = n_idx;
            g->y_built = y;
            g->n_y_built = n;
            g->full_build = (eachindex_y == nullptr);
            g->built = true;
            return PNB_OK;
        }
        // a cell overflowed its bucket: rebuild as CSR below (which also picks a larger K);
        // a domain error found on the way is reported by that build again
        *(volatile int *)g->h_err = 0;
        g->bucket_K = 0;
        g->bcount_alt_clean = false;
    }
    switch (g->p.ndims) {
        case 1: st = g->p.periodic ? build_nd<1, true>(g, y, n, eachindex_y, n_idx, index_base, s)
                                   : build_nd<1, false>(g, y, n, eachindex_y, n_idx, index_base, s); break;
        case 2: st = g->p.periodic ? build_nd<2, true>(g, y, n, eachindex_y, n_idx, index_base, s)
                                   : build_nd<2, false>(g, y, n, eachindex_y, n_idx, index_base, s); break;
        default: st = g->p.periodic ? build_nd<3, true>(g, y, n, eachindex_y, n_idx, index_base, s)
                                    : build_nd<3, false>(g, y, n, eachindex_y, n_idx, index_base, s); break;
    }
    if (st != PNB_OK) return st;
    // fullest cell -> bucket capacity of the following one-pass builds; order of the input ->
    // numbering of the buckets (d_maxcount = {fullest, steps in the first dim, steps in the last dim})
    unsigned int fullest = 0;
    if (g_build_layout != 0 && C > 0 && n_idx > 0) {
        if (!g->d_maxcount) PNB_CUDA(cudaMalloc(&g->d_maxcount, 4 * sizeof(unsigned int)));
        PNB_CUDA(cudaMemsetAsync(g->d_maxcount, 0, 4 * sizeof(unsigned int), s));
        k_max_cell_count<<<(unsigned)div_up(C, 256), 256, 0, s>>>(C, g->cell_start, g->d_maxcount);
        PNB_LAUNCHED();
        if (g_bucket_order < 0 && eachindex_y == nullptr && n_idx >= 4096 && g->p.ndims > 1) {
            const int64_t avail = n_idx - kProbeLag;
            const int64_t n_pairs = avail < 16384 ? avail : 16384, stride = avail / n_pairs;
            const unsigned pb = (unsigned)div_up(n_pairs, 256);
            if (g->p.ndims == 2) k_order_probe<2><<<pb, 256, 0, s>>>(g->p, y, n_pairs, stride, g->d_maxcount);
            else k_order_probe<3><<<pb, 256, 0, s>>>(g->p, y, n_pairs, stride, g->d_maxcount);
            PNB_LAUNCHED();
        }
        PNB_CUDA(cudaMemcpyAsync(g->h_err + 1, g->d_maxcount, 3 * sizeof(unsigned int),
                                 cudaMemcpyDeviceToHost, s));
    }
    st = check_err_word(g, s);  // also synchronizes: initialize!/update! are blocking calls
    if (st != PNB_OK) {
        // like the reference, a failed build leaves an unusable cell list behind
        g->n_built = 0;
        return st;
    }
    if (g_build_layout != 0 && C > 0 && n_idx > 0) {
        fullest = *(volatile unsigned int *)(g->h_err + 1);
        // 25 % + 4 slots of head room, rounded up to a power of two (slot -> cell is a shift);
        // only while the slots cost at most ~4x the records themselves (sparse grids with a few
        // crowded cells stay CSR)
        int64_t K = 16;
        while (K < (int64_t)fullest + fullest / 4 + 4) K *= 2;
        // (slot numbers are 32-bit in the kernels)
        g->bucket_K = (C * K <= 4 * n_idx + (1 << 20) && C * K < 0x7fffffffLL) ? (int)K : 0;
        // buckets numbered with the LAST dimension fastest when consecutive points step through
        // the last dimension clearly more often than through the first (e.g. clouds sorted by cell
        // tuple, test/point_cloud.jl:49); windows of a slab decomposition keep the plain order
        const unsigned first_steps = *(volatile unsigned int *)(g->h_err + 2);
        const unsigned last_steps = *(volatile unsigned int *)(g->h_err + 3);
        if (g->windowed) g->bucket_tr = false;
        else if (g_bucket_order >= 0) g->bucket_tr = g_bucket_order == 1 && g->p.ndims > 1;
        else g->bucket_tr = g->p.ndims > 1 && last_steps > 2u * first_steps + 64u;
    }
    g->csr_valid = true;
    g->n_built = n_idx;
    g->y_built = y;
    g->n_y_built = n;
    g->full_build = (eachindex_y == nullptr);
    g->built = true;
    if (g_build_layout == 2 && g->bucket_K > 0) {
        // test mode: rebuild at once into the bucket layout, so that every consumer is
        // exercised on it even after a first build
        static thread_local bool nested = false;
        if (!nested) {
            nested = true;
            pnb_status st2 = pnb_grid_build_f32(g, y, n, eachindex_y, n_idx, index_base, stream);
            nested = false;
            return st2;
        }
    }
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_build_async_f32(pnb_grid *g, const float *y, int64_t n, void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    // only the one-pass bucket build can run without looking at its result; everything else
    // (first build, hashed / template searches, CSR layout forced) is the blocking build
    t_async_build = !g->f64 && !g->hashed && !g->template_search && g->bucket_K > 0 && n > 0;
    const pnb_status st = pnb_grid_build_f32(g, y, n, nullptr, 0, 0, stream);
    t_async_build = false;
    return st;
}

// ---- append to the bucket layout (overlapped multi-GPU step) -----------------------
// The points y[first .. first + n_more) join the cell list of the current one-pass build with
// ids first + k: the migrants and ghosts that arrive from the neighbouring slabs while the
// interior layers are already being swept.  One point per thread, the lanes of a warp that hit
// the same cell share one atomic.
namespace pnb {
template <int ND>
__global__ void __launch_bounds__(256)
k_bucket_append(GridP g, BuildP bp, const float *__restrict__ y, int64_t first, int64_t n_more,
                uint32_t K, uint32_t *__restrict__ bcount, float4 *__restrict__ brec,
                int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = k < n_more;
    float p[3] = {0.f, 0.f, 0.f};
    if (in) {
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __ldg(y + (first + k) * ND + d);
    }
    const int lin = in ? point_cell_fast<ND, false, false>(g, bp, p) : -1;
    int bad = (in && lin < 0) ? 1 : 0;
    const unsigned m = __match_any_sync(0xffffffffu, lin);
    const int head = __ffs(m) - 1;
    unsigned base = 0u;
    if (lin >= 0 && head == lane_id()) base = atomicAdd(bcount + lin, (unsigned)__popc(m));
    base = __shfl_sync(0xffffffffu, base, head);
    if (lin >= 0) {
        const unsigned slot = base + (unsigned)__popc(m & ((1u << lane_id()) - 1u));
        if (slot < K) brec[(size_t)lin * K + slot] = make_float4(p[0], p[1], p[2], __int_as_float((int)(first + k)));
        else bad |= 8;
    }
    if (bad) atomicOr(err, bad);
}
}  // namespace pnb

extern "C" pnb_status pnb_grid_append_f32(pnb_grid *g, const float *y, int64_t first, int64_t n_more,
                                          void *stream)
{
    if (!g || g->f64 || g->hashed || g->p.periodic) {
        set_error("pnb_grid_append_f32 needs a Float32 non-periodic FullGridCellList search");
        return PNB_ERR_ARG;
    }
    if (!g->built || !g->bucket_valid || g->bucket_tr || y != g->y_built || first != g->n_y_built ||
        !g->full_build) {
        set_error("pnb_grid_append_f32: the cell list must be the one-pass (bucket) build of the "
                  "first `first` points of the same array");
        return PNB_ERR_STATE;
    }
    if (first + n_more > 0x7ffffff0LL) { set_error("more than 2^31 points are not supported"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    if (n_more > 0) {
        BuildP bp;
        for (int d = 0; d < 3; d++) { volatile float rc = 1.0f / g->p.cs[d]; bp.rcs[d] = rc; }
        ProfScope ps(PH_BUILD_BUCKET, s);
        const unsigned nb = (unsigned)div_up(n_more, 256);
        switch (g->p.ndims) {
            case 1: k_bucket_append<1><<<nb, 256, 0, s>>>(g->p, bp, y, first, n_more, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err); break;
            case 2: k_bucket_append<2><<<nb, 256, 0, s>>>(g->p, bp, y, first, n_more, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err); break;
            default: k_bucket_append<3><<<nb, 256, 0, s>>>(g->p, bp, y, first, n_more, (uint32_t)g->bucket_K, g->bcount, g->brec, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    g->n_built += n_more;
    g->n_y_built = first + n_more;
    g->canonical = false;
    g->csr_valid = false;
    if (!g->async_pending) return check_err_word(g, s) == PNB_OK ? PNB_OK : PNB_ERR_DOMAIN;
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_check(pnb_grid *g, void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    const pnb_status st = check_err_word(g, (cudaStream_t)stream);
    return st == PNB_RETRY_INTERNAL ? PNB_OK : st;
}

// The same for a caller that has ALREADY waited (cudaEventSynchronize on an event it recorded
// behind the update! / append) -- nothing is synchronised here, so kernels launched after that
// event (the sweep) keep running.  Error bits they set are reported by the next check.
extern "C" pnb_status pnb_grid_check_settled(pnb_grid *g)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    const pnb_status st = check_err_word_settled(g);
    return st == PNB_RETRY_INTERNAL ? PNB_OK : st;
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
    if (g.hashed) {
        // SpatialHashingCellList: the 0-based table key of the point's cell (-1: not Int32)
        long long hc[3];
        bool ok = g.periodic ? hash_cell_coords<ND, true>(g, p, hc) : hash_cell_coords<ND, false>(g, p, hc);
        out[i] = ok ? (int)spatial_hash_key<ND>(hc, g.total_cells) : -1;
        return;
    }
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

extern "C" pnb_status pnb_grid_export_csr(const pnb_grid *g_, int32_t *cell_start,
                                          int32_t *cell_points, int index_base, void *stream)
{
    pnb_grid *g = const_cast<pnb_grid *>(g_);
    if (!g || !g->built) { set_error("the neighborhood search has not been initialized"); return PNB_ERR_STATE; }
    cudaStream_t s = (cudaStream_t)stream;
    if (cell_points) {
        pnb_status stc = ensure_canonical(g, s);
        if (stc != PNB_OK) return stc;
    } else {
        pnb_status stc = ensure_csr(g, s);
        if (stc != PNB_OK) return stc;
    }
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
    {
        pnb_status stc = ensure_canonical(g, s);
        if (stc != PNB_OK) return stc;
    }
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

// ---------------------------------------------------------------------------------------------
// slab exchange bookkeeping (multi-GPU)
// ---------------------------------------------------------------------------------------------
namespace pnb {
__device__ __forceinline__ void warp_append(bool flag, int32_t value, int32_t *list, int64_t cap,
                                            int32_t *counter)
{
    const unsigned m = __ballot_sync(0xffffffffu, flag);
    if (m == 0u) return;
    const int lane = lane_id();
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (flag) {
        const int64_t pos = (int64_t)base + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) list[pos] = value;
    }
}

__global__ void __launch_bounds__(256)
k_slab_classify(const float *__restrict__ coords, int64_t n, int nd, float pmin, float cs,
                long long z_lo, long long z_hi, int has_up, int has_down,
                int32_t *__restrict__ up_idx, int32_t *__restrict__ down_idx,
                int32_t *__restrict__ leave_idx, int64_t cap, int32_t *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool up = false, down = false, leave = false;
    if (i < n) {
        const float z = __ldg(coords + i * nd + (nd - 1));
        const float f = floorf(__fdiv_rn(__fsub_rn(z, pmin), cs));   // full_grid.jl:93
        // NaN / huge values stay here: the following update! reports them as a domain error
        if (fabsf(f) < 4.0e18f) {
            const long long cz = (long long)f + 1;
            up = has_up && cz >= z_hi;
            down = has_down && cz <= z_lo;
            // a point that leaves through the global bottom / top has no rank to go to: it is
            // KEPT, so that the following update! reports it exactly like the undecomposed search
            // ("particle coordinates are NaN or outside the domain bounds", full_grid.jl:211)
            leave = (cz < z_lo && has_down) || (cz > z_hi && has_up);
        }
    }
    warp_append(up, (int32_t)i, up_idx, cap, counts + 0);
    warp_append(down, (int32_t)i, down_idx, cap, counts + 1);
    warp_append(leave, (int32_t)i, leave_idx, cap, counts + 2);
}
}  // namespace pnb

extern "C" pnb_status pnb_slab_classify_f32(const float *coords, int64_t n, int ndims,
                                            float padded_min_z, float cell_size_z, int64_t z_lo,
                                            int64_t z_hi, int has_up, int has_down,
                                            int32_t *up_idx, int32_t *down_idx, int32_t *leave_idx,
                                            int64_t cap, int32_t *counts_dev, int64_t *counts,
                                            void *stream)
{
    if (!counts || !counts_dev) { set_error("counts is NULL"); return PNB_ERR_ARG; }
    if (ndims < 1 || ndims > 3) { set_error("`NDIMS` must be 1, 2, or 3"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    PNB_CUDA(cudaMemsetAsync(counts_dev, 0, 3 * sizeof(int32_t), s));
    if (n > 0) {
        k_slab_classify<<<(unsigned)div_up(n, 256), 256, 0, s>>>(
            coords, n, ndims, padded_min_z, cell_size_z, (long long)z_lo, (long long)z_hi, has_up,
            has_down, up_idx, down_idx, leave_idx, cap, counts_dev);
        PNB_LAUNCHED();
    }
    int32_t h[3] = {0, 0, 0};
    PNB_CUDA(cudaMemcpyAsync(h, counts_dev, sizeof(h), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; k++) counts[k] = h[k];
    return PNB_OK;
}

namespace pnb {

__device__ __forceinline__ long long slab_layer(float z, float pmin, float cs)
{
    const float f = floorf(__fdiv_rn(__fsub_rn(z, pmin), cs));   // full_grid.jl:93
    return fabsf(f) < 4.0e18f ? (long long)f + 1 : (long long)0x4000000000000000LL;
}

// row r of `list` -> dst[r * W ...]
__global__ void __launch_bounds__(256)
k_slab_pack(pnb_slab_arrays A, int W, const int32_t *__restrict__ list, int64_t count,
            float *__restrict__ dst)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const int64_t i = list[r];
    int col = 0;
    for (int a = 0; a < A.n_arrays; a++) {
        const int w = A.width[a];
        for (int k = 0; k < w; k++) dst[r * W + col + k] = A.ptr[a][i * w + k];
        col += w;
    }
}

// classes of the exchanged rows: 1 = becomes an owned point (migrant), 2 = ghost, 0 = dropped.
// cls[row] = class | rank << 2 with the rank taken from the class counter.
//   counters: [0] migrants, [1] ghosts
__global__ void __launch_bounds__(256)
k_slab_rows_classify(const float *__restrict__ rows, int64_t count, int W, int nd, float pmin,
                     float cs, long long lo_mig, long long hi_mig, long long ghost_layer,
                     int32_t *__restrict__ cls, int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (r < count) {
        const long long cz = slab_layer(rows[r * W + (nd - 1)], pmin, cs);
        if (cz >= lo_mig && cz <= hi_mig) c = 1;
        else if (cz == ghost_layer) c = 2;
    }
    // warp-aggregated ranks
    for (int k = 1; k <= 2; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, c == k);
        if (m == 0u) continue;
        int base = 0;
        if (lane_id() == __ffs(m) - 1) base = atomicAdd(counters + (k - 1), __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (c == k) cls[r] = k | ((base + __popc(m & ((1u << lane_id()) - 1u))) << 2);
    }
    if (r < count && c == 0) cls[r] = 0;
}

// place classified rows: migrants at n_stay + rank, ghosts at n_stay + n_mig + rank
__global__ void __launch_bounds__(256)
k_slab_rows_place(pnb_slab_arrays A, int W, const float *__restrict__ rows, int64_t count,
                  const int32_t *__restrict__ cls, int64_t n_stay,
                  const int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= count) return;
    const int c = cls[r] & 3;
    if (c == 0) return;
    const int64_t rank = cls[r] >> 2;
    const int64_t i = n_stay + (c == 2 ? (int64_t)counters[0] : 0) + rank;
    int col = 0;
    for (int a = 0; a < A.n_arrays; a++) {
        const int w = A.width[a];
        for (int k = 0; k < w; k++) A.ptr[a][i * w + k] = rows[r * W + col + k];
        col += w;
    }
}

// emigrants: holes = leaving points below n_stay, fillers = staying points of the tail
__global__ void k_slab_mark_tail(const int32_t *__restrict__ leave, int64_t n_leave, int64_t n_stay,
                                 int32_t *__restrict__ tailflag, int32_t *__restrict__ holes,
                                 int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool hole = false;
    int32_t v = 0;
    if (r < n_leave) {
        v = leave[r];
        if (v >= n_stay) tailflag[v - n_stay] = 1;
        else hole = true;
    }
    warp_append(hole, v, holes, n_leave, counters + 2);
}
__global__ void k_slab_fillers(int64_t n_tail, int64_t n_stay, const int32_t *__restrict__ tailflag,
                               int32_t *__restrict__ fillers, int32_t *__restrict__ counters)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool f = t < n_tail && tailflag[t] == 0;
    warp_append(f, (int32_t)(n_stay + t), fillers, n_tail, counters + 3);
}
__global__ void k_slab_fill(pnb_slab_arrays A, const int32_t *__restrict__ holes,
                            const int32_t *__restrict__ fillers, const int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= counters[2]) return;      // counters[2] == counters[3] by construction
    const int64_t h = holes[r], f = fillers[r];
    for (int a = 0; a < A.n_arrays; a++) {
        const int w = A.width[a];
        for (int k = 0; k < w; k++) A.ptr[a][h * w + k] = A.ptr[a][f * w + k];
    }
}
}  // namespace pnb

static int slab_row_width(const pnb_slab_arrays *A)
{
    int W = 0;
    for (int a = 0; a < A->n_arrays; a++) W += A->width[a];
    return W;
}

extern "C" pnb_status pnb_slab_pack_f32(const pnb_slab_arrays *arrays, int64_t n, int ndims,
                                        float padded_min_z, float cell_size_z, int64_t z_lo,
                                        int64_t z_hi, int has_up, int has_down, int32_t *up_idx,
                                        int32_t *down_idx, int32_t *leave_idx, int64_t cap,
                                        float *send_up, float *send_down, int32_t *counts_dev,
                                        int64_t *counts, void *stream)
{
    if (!arrays || arrays->n_arrays < 1 || arrays->n_arrays > 8 || arrays->width[0] != ndims) {
        set_error("arrays[0] must be the coordinates (width = NDIMS), at most 8 arrays");
        return PNB_ERR_ARG;
    }
    pnb_status st = pnb_slab_classify_f32(arrays->ptr[0], n, ndims, padded_min_z, cell_size_z, z_lo,
                                          z_hi, has_up, has_down, up_idx, down_idx, leave_idx, cap,
                                          counts_dev, counts, stream);
    if (st != PNB_OK) return st;
    if (counts[0] > cap || counts[1] > cap || counts[2] > cap) return PNB_OK;   // caller retries
    cudaStream_t s = (cudaStream_t)stream;
    const int W = slab_row_width(arrays);
    if (counts[0] > 0) {
        k_slab_pack<<<(unsigned)div_up(counts[0], 256), 256, 0, s>>>(*arrays, W, up_idx, counts[0], send_up);
        PNB_LAUNCHED();
    }
    if (counts[1] > 0) {
        k_slab_pack<<<(unsigned)div_up(counts[1], 256), 256, 0, s>>>(*arrays, W, down_idx, counts[1], send_down);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

// rows of `arrays` listed in list[0 .. count) -> dst (count x row_width), stream-ordered: the
// second half of pnb_slab_pack_f32 for callers that classify on another stream
extern "C" pnb_status pnb_slab_pack_rows_f32(const pnb_slab_arrays *arrays, const int32_t *list,
                                             int64_t count, float *dst, void *stream)
{
    if (!arrays || arrays->n_arrays < 1 || arrays->n_arrays > 8) { set_error("bad arrays"); return PNB_ERR_ARG; }
    if (count <= 0) return PNB_OK;
    cudaStream_t s = (cudaStream_t)stream;
    k_slab_pack<<<(unsigned)div_up(count, 256), 256, 0, s>>>(*arrays, slab_row_width(arrays), list, count, dst);
    PNB_LAUNCHED();
    return PNB_OK;
}

extern "C" pnb_status pnb_slab_unpack_f32(const pnb_slab_arrays *arrays, int64_t n, int ndims,
                                          float padded_min_z, float cell_size_z, int64_t z_lo,
                                          int64_t z_hi, int has_up, int has_down,
                                          const int32_t *leave_idx, int64_t n_leave,
                                          const float *recv_up, int64_t n_recv_up,
                                          const float *recv_down, int64_t n_recv_down,
                                          const float *send_up, int64_t n_up,
                                          const float *send_down, int64_t n_down, int32_t *scratch,
                                          int64_t *out, void *stream)
{
    if (!arrays || !out || !scratch) { set_error("NULL argument"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int W = slab_row_width(arrays);
    const int64_t n_stay = n - n_leave;
    // scratch layout (int32): counters[16] | tailflag[n_leave] | holes[n_leave] | fillers[n_leave]
    //                         | cls of recv_up, recv_down, send_up, send_down
    int32_t *counters = scratch;
    int32_t *tailflag = scratch + 16;
    int32_t *holes = tailflag + n_leave;
    int32_t *fillers = holes + n_leave;
    int32_t *cls_ru = fillers + n_leave;
    int32_t *cls_rd = cls_ru + n_recv_up;
    int32_t *cls_su = cls_rd + n_recv_down;
    int32_t *cls_sd = cls_su + n_up;
    PNB_CUDA(cudaMemsetAsync(scratch, 0, sizeof(int32_t) * (size_t)(16 + n_leave), s));
    if (n_leave > 0) {
        const unsigned b = (unsigned)div_up(n_leave, 256);
        k_slab_mark_tail<<<b, 256, 0, s>>>(leave_idx, n_leave, n_stay, tailflag, holes, counters);
        PNB_LAUNCHED();
        k_slab_fillers<<<b, 256, 0, s>>>(n_leave, n_stay, tailflag, fillers, counters);
        PNB_LAUNCHED();
        k_slab_fill<<<b, 256, 0, s>>>(*arrays, holes, fillers, counters);
        PNB_LAUNCHED();
    }
    const long long BIG = 0x3fffffffffffffffLL;
    // received from rank + 1: cz <= z_hi -> owned, cz == z_hi + 1 -> ghost
    // received from rank - 1: cz >= z_lo -> owned, cz == z_lo - 1 -> ghost
    // sent upwards / downwards and now one layer outside the slab -> my own ghosts
    struct Src { const float *rows; int64_t n; int32_t *cls; long long lo, hi, ghost; };
    const Src src[4] = {
        {recv_up, n_recv_up, cls_ru, -BIG, (long long)z_hi, (long long)z_hi + 1},
        {recv_down, n_recv_down, cls_rd, (long long)z_lo, BIG, (long long)z_lo - 1},
        {send_up, has_up ? n_up : 0, cls_su, BIG, -BIG, (long long)z_hi + 1},
        {send_down, has_down ? n_down : 0, cls_sd, BIG, -BIG, (long long)z_lo - 1}};
    for (const Src &q : src) {
        if (q.n <= 0) continue;
        k_slab_rows_classify<<<(unsigned)div_up(q.n, 256), 256, 0, s>>>(
            q.rows, q.n, W, ndims, padded_min_z, cell_size_z, q.lo, q.hi, q.ghost, q.cls, counters);
        PNB_LAUNCHED();
    }
    for (const Src &q : src) {
        if (q.n <= 0) continue;
        k_slab_rows_place<<<(unsigned)div_up(q.n, 256), 256, 0, s>>>(*arrays, W, q.rows, q.n, q.cls,
                                                                   n_stay, counters);
        PNB_LAUNCHED();
    }
    int32_t h[2] = {0, 0};
    PNB_CUDA(cudaMemcpyAsync(h, counters, sizeof(h), cudaMemcpyDeviceToHost, s));
    PNB_CUDA(cudaStreamSynchronize(s));
    out[0] = n_stay + h[0];
    out[1] = n_stay + h[0] + h[1];
    return PNB_OK;
}

// ---- slab bookkeeping of the overlapped step ----------------------------------------------------
namespace pnb {
// received rows -> arrays[n_own + r]; flag = 1 if the row now belongs to this slab (migrant),
// 0 = ghost.  counters[0] += migrants; counters[1] |= 1 if a migrant landed deeper than `depth`
// layers inside the slab (the interior layers were swept without it: the caller must repeat).
__global__ void __launch_bounds__(256)
k_slab_append(pnb_slab_arrays A, int64_t W, int nd, int64_t n_own, const float *__restrict__ rows_a,
              int64_t n_a, const float *__restrict__ rows_b, int64_t n_b, float pmin, float cs,
              long long z_lo, long long z_hi, long long depth, unsigned char *__restrict__ flags,
              int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool mig = false, deep = false;
    if (r < n_a + n_b) {
        const float *row = r < n_a ? rows_a + r * W : rows_b + (r - n_a) * W;
        const int64_t i = n_own + r;
        int col = 0;
#pragma unroll
        for (int a = 0; a < 8; a++) {              // static indices: A stays in the parameter bank
            if (a >= A.n_arrays) break;
            const int w = A.width[a];
            for (int k = 0; k < w; k++) A.ptr[a][i * w + k] = row[col + k];
            col += w;
        }
        const long long cz = slab_layer(row[nd - 1], pmin, cs);
        mig = cz >= z_lo && cz <= z_hi;
        deep = mig && cz >= z_lo + depth && cz <= z_hi - depth;
        flags[i] = mig ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, mig);
    if (m && lane_id() == __ffs(m) - 1) atomicAdd(counters + 0, __popc(m));
    if (__any_sync(0xffffffffu, deep) && lane_id() == 0) atomicOr(counters + 1, 1);
}

__global__ void k_slab_clear_flags(const int32_t *__restrict__ leave, int64_t n_leave,
                                   unsigned char *__restrict__ flags)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_leave) flags[leave[r]] = 0;
}
// holes: rows below n_new that are not owned; movers: owned rows at or above n_new
__global__ void __launch_bounds__(256)
k_slab_compact_lists(const unsigned char *__restrict__ flags, int64_t n_rows, int64_t n_new,
                     int32_t *__restrict__ holes, int32_t *__restrict__ movers, int64_t cap,
                     int32_t *__restrict__ counters)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in = i < n_rows;
    const bool own = in && flags[i] != 0;
    warp_append(in && !own && i < n_new, (int32_t)i, holes, cap, counters + 2);
    warp_append(own && i >= n_new, (int32_t)i, movers, cap, counters + 3);
}
__global__ void k_slab_compact_move(pnb_slab_arrays A, const int32_t *__restrict__ holes,
                                    const int32_t *__restrict__ movers,
                                    const int32_t *__restrict__ counters)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= counters[2]) return;        // counters[2] == counters[3] by construction
    const int64_t h = holes[r], f = movers[r];
    for (int a = 0; a < A.n_arrays; a++) {
        const int w = A.width[a];
        for (int k = 0; k < w; k++) A.ptr[a][h * w + k] = A.ptr[a][f * w + k];
    }
}
}  // namespace pnb

// Received rows appended behind the owned points (no hole filling: the ids of the owned points
// stay valid for the sweep that is already running).  counters_dev: >= 4 int32, zeroed here;
// after the call counters_dev[0] = migrants among the rows, counters_dev[1] = 1 if one of them
// lies deeper than `depth` layers inside the slab.
// row_stride: floats between two rows of recv_up / recv_down (0 = dense, the row width; the
// receive buffers of a pnb_slab_link pad rows to a multiple of 4 floats)
extern "C" pnb_status pnb_slab_append_strided_f32(const pnb_slab_arrays *arrays, int64_t n_own, int ndims,
                                                  float padded_min_z, float cell_size_z, int64_t z_lo,
                                                  int64_t z_hi, int64_t depth, const float *recv_up,
                                                  int64_t n_recv_up, const float *recv_down,
                                                  int64_t n_recv_down, int64_t row_stride, uint8_t *flags,
                                                  int32_t *counters_dev, void *stream)
{
    if (!arrays || arrays->n_arrays < 1 || arrays->n_arrays > 8 || arrays->width[0] != ndims || !flags ||
        !counters_dev) {
        set_error("pnb_slab_append_f32: bad arguments");
        return PNB_ERR_ARG;
    }
    int W = 0;
    for (int a = 0; a < arrays->n_arrays; a++) W += arrays->width[a];
    if (row_stride == 0) row_stride = W;
    if (row_stride < W) { set_error("pnb_slab_append_f32: row_stride < row width"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    PNB_CUDA(cudaMemsetAsync(counters_dev, 0, 4 * sizeof(int32_t), s));
    const int64_t n = n_recv_up + n_recv_down;
    if (n > 0) {
        k_slab_append<<<(unsigned)div_up(n, 256), 256, 0, s>>>(
            *arrays, row_stride, ndims, n_own, recv_up, n_recv_up, recv_down, n_recv_down, padded_min_z,
            cell_size_z, (long long)z_lo, (long long)z_hi, (long long)depth, flags, counters_dev);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}

extern "C" pnb_status pnb_slab_append_f32(const pnb_slab_arrays *arrays, int64_t n_own, int ndims,
                                          float padded_min_z, float cell_size_z, int64_t z_lo,
                                          int64_t z_hi, int64_t depth, const float *recv_up,
                                          int64_t n_recv_up, const float *recv_down,
                                          int64_t n_recv_down, uint8_t *flags, int32_t *counters_dev,
                                          void *stream)
{
    return pnb_slab_append_strided_f32(arrays, n_own, ndims, padded_min_z, cell_size_z, z_lo, z_hi, depth,
                                       recv_up, n_recv_up, recv_down, n_recv_down, 0, flags, counters_dev,
                                       stream);
}

// End of the step: rows [0, n_own) minus the n_leave leavers plus the n_mig migrants among the
// appended rows [n_own, n_own + n_app) become the owned points [0, n_new), n_new = n_own - n_leave
// + n_mig (every array of `arrays` is permuted alike: pass dv too if it is still needed).
// scratch: >= 2 * (n_leave + n_app) + 16 int32; counters_dev as above ([2], [3] are used here).
extern "C" pnb_status pnb_slab_compact_f32(const pnb_slab_arrays *arrays, int64_t n_own, int64_t n_app,
                                           const int32_t *leave_idx, int64_t n_leave, int64_t n_mig,
                                           uint8_t *flags, int32_t *scratch, int32_t *counters_dev,
                                           void *stream)
{
    if (!arrays || !flags || !scratch || !counters_dev) { set_error("NULL argument"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n_new = n_own - n_leave + n_mig, n_rows = n_own + n_app;
    const int64_t cap = n_leave + n_app + 8;
    if (n_leave == 0 && n_mig == 0) return PNB_OK;      // nothing leaves, nothing joins
    if (n_own > 0) PNB_CUDA(cudaMemsetAsync(flags, 1, (size_t)n_own, s));
    PNB_CUDA(cudaMemsetAsync(counters_dev + 2, 0, 2 * sizeof(int32_t), s));
    if (n_leave > 0) {
        k_slab_clear_flags<<<(unsigned)div_up(n_leave, 256), 256, 0, s>>>(leave_idx, n_leave, flags);
        PNB_LAUNCHED();
    }
    if (n_rows > 0 && (n_leave > 0 || n_mig > 0)) {
        // only the rows from min(n_new, first leaver ...) on can be holes or movers; scanning all
        // of them costs a byte per row
        k_slab_compact_lists<<<(unsigned)div_up(n_rows, 256), 256, 0, s>>>(flags, n_rows, n_new, scratch,
                                                                          scratch + cap, cap, counters_dev);
        PNB_LAUNCHED();
        k_slab_compact_move<<<(unsigned)div_up(cap, 256), 256, 0, s>>>(*arrays, scratch, scratch + cap,
                                                                       counters_dev);
        PNB_LAUNCHED();
    }
    return PNB_OK;
}


// =============================================================================================
// Float64 searches (f64.cuh)
// =============================================================================================
namespace pnb {

// constructors in Float64 (full_grid.jl:66-78, nhs_grid.jl:102-124): every operation is one
// IEEE double operation (the host compiler uses SSE2 doubles: no excess precision)
static pnb_status grid_params_host64(int ndims, double r, const double *min_corner,
                                     const double *max_corner, const double *box_min,
                                     const double *box_max, double *padded_min, double *padded_max,
                                     int64_t *grid_size, int64_t *n_cells, double *cell_size)
{
    if (ndims < 1 || ndims > 3) { set_error("`NDIMS` must be 1, 2, or 3"); return PNB_ERR_ARG; }
    if (!min_corner || !max_corner) {
        set_error("min_corner and max_corner must have the same length");
        return PNB_ERR_ARG;
    }
    volatile double factor = 1001.0 / 1000.0;       // Float64(1001 // 1000)
    volatile double pad = t_corners_padded ? 0.0 : factor * r;
    const bool is_template = r < 2.220446049250313e-16;
    for (int d = 0; d < ndims; d++) {
        volatile double mn = min_corner[d] - pad;
        volatile double mx = max_corner[d] + pad;
        if (padded_min) padded_min[d] = mn;
        if (padded_max) padded_max[d] = mx;
        if (grid_size) {
            if (is_template) grid_size[d] = 0;
            else {
                volatile double ext = mx - mn;
                volatile double q = ext / r;
                grid_size[d] = (int64_t)std::ceil(q);
            }
        }
        if (n_cells) n_cells[d] = -1;
        if (cell_size) cell_size[d] = r;
    }
    if (box_min && box_max && !is_template) {
        for (int d = 0; d < ndims; d++) {
            volatile double size = box_max[d] - box_min[d];
            const double nc = std::floor((size + 10.0 * 2.220446049250313e-16) / r);
            const int64_t nci = (int64_t)nc;
            if (n_cells) n_cells[d] = nci;
            if (cell_size) { volatile double cs = size / (double)nci; cell_size[d] = cs; }
            if (nci < 3) {
                set_error("the `GridNeighborhoodSearch` needs at least 3 cells in each dimension "
                          "when used with periodicity. Please use no NHS for very small problems.");
                return PNB_ERR_ARG;
            }
        }
    }
    return PNB_OK;
}

template <int ND>
__global__ void k_hist64(GridP64 g, const double *__restrict__ y, int64_t n_idx,
                         const int32_t *__restrict__ idx, int base,
                         uint32_t *__restrict__ cell_count, int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_idx) return;
    const int64_t id = idx ? (int64_t)idx[k] - base : k;
    double p[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = y[id * ND + d];
    int cc[3];
    const int lin = point_cell64<ND>(g, p, cc);
    if (lin < 0) { atomicOr(err, 1); return; }
    atomicAdd(cell_count + lin, 1u);
}

template <int ND>
__global__ void k_scatter64(GridP64 g, const double *__restrict__ y, int64_t n_idx,
                            const int32_t *__restrict__ idx, int base,
                            uint32_t *__restrict__ cursor, Rec64 *__restrict__ out)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_idx) return;
    const int64_t id = idx ? (int64_t)idx[k] - base : k;
    double p[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = y[id * ND + d];
    int cc[3];
    const int lin = point_cell64<ND>(g, p, cc);
    if (lin < 0) return;
    const uint32_t slot = atomicAdd(cursor + lin, 1u);
    Rec64 r;
    r.x = p[0]; r.y = p[1]; r.z = p[2]; r.id = id;
    out[slot] = r;
}

// ids ascending inside every cell (rank by counting), records + id list
__global__ void k_canonicalize64(int total_cells, const uint32_t *__restrict__ cell_start,
                                 const Rec64 *__restrict__ in, Rec64 *__restrict__ out,
                                 int32_t *__restrict__ cell_points)
{
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= total_cells) return;
    const uint32_t s0 = cell_start[c], s1 = cell_start[c + 1];
    const int cnt = (int)(s1 - s0);
    for (int e = lane_id(); e < cnt; e += 32) {
        const Rec64 rec = in[s0 + e];
        int r = 0;
        for (int k = 0; k < cnt; k++) r += (in[s0 + k].id < rec.id) ? 1 : 0;
        cell_points[s0 + r] = (int32_t)rec.id;
        out[s0 + r] = rec;
    }
}

template <int ND>
__global__ void k_point_cells64(GridP64 g, const double *__restrict__ x, int64_t n,
                                int32_t *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = x[i * ND + d];
    int cc[3];
    out[i] = point_cell64<ND>(g, p, cc);
}

}  // namespace pnb

// the part of the Float64 / mixed constructors after the host arithmetic
static pnb_status create64_common(pnb_grid *g, int ndims, double r, double r2, bool periodic,
                                  const int64_t *gsz, const int64_t *ncl, const double *bsize,
                                  pnb_grid **out)
{
    g->template_search = r < 2.220446049250313e-16;
    GridP64 &p = g->p64;
    p.ndims = ndims;
    p.periodic = periodic ? 1 : 0;
    p.r = r;
    p.r2 = r2;
    int64_t total = 1;
    for (int d = 0; d < 3; d++) {
        p.minc[d] = d < ndims ? g->padded_min64[d] : 0.0;
        p.cs[d] = d < ndims ? g->cell_size64[d] : 1.0;
        int64_t gs = d < ndims ? gsz[d] : 1;
        if (g->template_search) gs = d < ndims ? 0 : 1;
        if (gs > 0x7fffffff) gs = 0x7fffffff;
        p.gs[d] = (int)gs;
        p.nc[d] = d < ndims ? (int)ncl[d] : -1;
        p.bsize[d] = (d < ndims && periodic) ? bsize[d] : 1.0;
        g->grid_size[d] = gs;
        g->n_cells[d] = ncl[d];
        total *= gs;
        if (total > 0x7fffff00LL) {
            set_error("cell grid too large for this build: more than 2^31 cells");
            delete g;
            return PNB_ERR_ARG;
        }
    }
    p.total_cells = (int)total;
    // the shared (element-type independent) part of the Float32 scalars: exports read these
    g->p.ndims = ndims;
    g->p.periodic = p.periodic;
    g->p.total_cells = p.total_cells;
    for (int d = 0; d < 3; d++) { g->p.gs[d] = p.gs[d]; g->p.nc[d] = p.nc[d]; g->p.off[d] = 0; }
    pnb_status st = grid_alloc_common(g, total);
    if (st != PNB_OK) { pnb_grid_destroy(g); return st; }
    *out = g;
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_params_f64(int ndims, double search_radius, const double *min_corner,
                                          const double *max_corner, const double *box_min,
                                          const double *box_max, double *padded_min,
                                          double *padded_max, int64_t *grid_size, int64_t *n_cells,
                                          double *cell_size)
{
    return grid_params_host64(ndims, search_radius, min_corner, max_corner, box_min, box_max,
                              padded_min, padded_max, grid_size, n_cells, cell_size);
}

extern "C" pnb_status pnb_grid_create_f64(int ndims, double r, const double *min_corner,
                                          const double *max_corner, const double *box_min,
                                          const double *box_max, pnb_grid **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    pnb_grid *g = new pnb_grid();
    memset(g, 0, sizeof(*g));
    g->f64 = true;
    int64_t gsz[3] = {1, 1, 1}, ncl[3] = {-1, -1, -1};
    pnb_status st = grid_params_host64(ndims, r, min_corner, max_corner, box_min, box_max,
                                       g->padded_min64, g->padded_max64, gsz, ncl, g->cell_size64);
    if (st != PNB_OK) { delete g; return st; }
    double bsize[3] = {1.0, 1.0, 1.0};
    const bool periodic = box_min && box_max && !(r < 2.220446049250313e-16);
    for (int d = 0; d < ndims && periodic; d++) { volatile double size = box_max[d] - box_min[d]; bsize[d] = size; }
    volatile double r2 = r * r;
    return create64_common(g, ndims, r, r2, periodic, gsz, ncl, bsize, out);
}

// Float64 coordinates / corners with a Float32 search radius (and Float32 PeriodicBox), the
// combination docs/literate/src/tut_gpu_usage.jl:45-50 describes: Julia's promotion makes the
// padding Float32 (1001//1000 * r, full_grid.jl:66-67) added to Float64 corners, the grid size and
// the cell arithmetic Float64 (:74, :93), pos_diff = Float32.(x_i - y_j) and everything after it
// Float32 (nhs_grid.jl:547-555).
static pnb_status grid_params_host_mixed(int ndims, float r, const double *min_corner,
                                         const double *max_corner, const float *box_min,
                                         const float *box_max, double *padded_min,
                                         double *padded_max, int64_t *grid_size, int64_t *n_cells,
                                         float *cell_size)
{
    if (ndims < 1 || ndims > 3) { set_error("`NDIMS` must be 1, 2, or 3"); return PNB_ERR_ARG; }
    if (!min_corner || !max_corner) {
        set_error("min_corner and max_corner must have the same length");
        return PNB_ERR_ARG;
    }
    volatile float factor = 1001.0f / 1000.0f;
    volatile float pad = t_corners_padded ? 0.0f : factor * r;
    const bool is_template = (double)r < 2.220446049250313e-16;
    for (int d = 0; d < ndims; d++) {
        volatile double mn = min_corner[d] - (double)pad;
        volatile double mx = max_corner[d] + (double)pad;
        if (padded_min) padded_min[d] = mn;
        if (padded_max) padded_max[d] = mx;
        if (grid_size) {
            if (is_template) grid_size[d] = 0;
            else {
                volatile double ext = mx - mn;
                volatile double q = ext / (double)r;
                grid_size[d] = (int64_t)std::ceil(q);
            }
        }
        if (n_cells) n_cells[d] = -1;
        if (cell_size) cell_size[d] = r;
    }
    if (box_min && box_max && !is_template) {
        for (int d = 0; d < ndims; d++) {
            volatile float size = box_max[d] - box_min[d];
            const double nc = std::floor(((double)size + 10.0 * 2.220446049250313e-16) / (double)r);
            const int64_t nci = (int64_t)nc;
            if (n_cells) n_cells[d] = nci;
            if (cell_size) { volatile float cs = size / (float)nci; cell_size[d] = cs; }
            if (nci < 3) {
                set_error("the `GridNeighborhoodSearch` needs at least 3 cells in each dimension "
                          "when used with periodicity. Please use no NHS for very small problems.");
                return PNB_ERR_ARG;
            }
        }
    }
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_params_mixed(int ndims, float search_radius,
                                            const double *min_corner, const double *max_corner,
                                            const float *box_min, const float *box_max,
                                            double *padded_min, double *padded_max,
                                            int64_t *grid_size, int64_t *n_cells, float *cell_size)
{
    return grid_params_host_mixed(ndims, search_radius, min_corner, max_corner, box_min, box_max,
                                  padded_min, padded_max, grid_size, n_cells, cell_size);
}

extern "C" pnb_status pnb_grid_create_mixed(int ndims, float r, const double *min_corner,
                                            const double *max_corner, const float *box_min,
                                            const float *box_max, pnb_grid **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    pnb_grid *g = new pnb_grid();
    memset(g, 0, sizeof(*g));
    g->f64 = true;
    int64_t gsz[3] = {1, 1, 1}, ncl[3] = {-1, -1, -1};
    float csf[3] = {r, r, r};
    pnb_status st = grid_params_host_mixed(ndims, r, min_corner, max_corner, box_min, box_max,
                                           g->padded_min64, g->padded_max64, gsz, ncl, csf);
    if (st != PNB_OK) { delete g; return st; }
    double bsize[3] = {1.0, 1.0, 1.0};
    const bool periodic = box_min && box_max && !((double)r < 2.220446049250313e-16);
    for (int d = 0; d < 3; d++) {
        g->cell_size64[d] = (double)csf[d];
        g->cell_size[d] = csf[d];
        if (d < ndims && periodic) { volatile float size = box_max[d] - box_min[d]; bsize[d] = (double)size; }
    }
    volatile float r2f = r * r;
    g->p64.mixed = 1;
    return create64_common(g, ndims, (double)r, (double)r2f, periodic, gsz, ncl, bsize, out);
}

extern "C" pnb_status pnb_grid_create_padded_mixed(int ndims, float r, const double *padded_min,
                                                   const double *padded_max, const float *box_min,
                                                   const float *box_max, pnb_grid **out)
{
    t_corners_padded = true;
    pnb_status st = pnb_grid_create_mixed(ndims, r, padded_min, padded_max, box_min, box_max, out);
    t_corners_padded = false;
    return st;
}

extern "C" pnb_status pnb_grid_create_padded_f64(int ndims, double r, const double *padded_min,
                                                 const double *padded_max, const double *box_min,
                                                 const double *box_max, pnb_grid **out)
{
    t_corners_padded = true;
    pnb_status st = pnb_grid_create_f64(ndims, r, padded_min, padded_max, box_min, box_max, out);
    t_corners_padded = false;
    return st;
}

extern "C" pnb_status pnb_grid_build_f64(pnb_grid *g, const double *y, int64_t n,
                                         const int32_t *eachindex_y, int64_t n_idx, int index_base,
                                         void *stream)
{
    if (!g || !g->f64) { set_error("not a Float64 grid handle"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t C = g->p64.total_cells;
    if (eachindex_y == nullptr) n_idx = n;
    g->built = false;
    g->y_refreshed = false;
    g->bucket_valid = false;
    if (g->template_search) {
        PNB_CUDA(cudaMemsetAsync(g->cell_start, 0, sizeof(uint32_t) * (size_t)(C + 1), s));
        PNB_CUDA(cudaStreamSynchronize(s));
        g->n_built = 0; g->built = true; g->y_built = y; g->n_y_built = n; g->full_build = false;
        g->csr_valid = true; g->canonical = true;
        return PNB_OK;
    }
    if (n_idx > 0x7ffffff0LL || n > 0x7ffffff0LL) {
        set_error("more than 2^31 points are not supported (ids are Int32, full_grid.jl:177)");
        return PNB_ERR_ARG;
    }
    if (n_idx > 0 && y == nullptr) { set_error("y is NULL"); return PNB_ERR_ARG; }
    if (n_idx > g->cap_points) {
        cudaFree(g->cell_points); g->cell_points = nullptr;
        cudaFree(g->sorted64); g->sorted64 = nullptr;
        cudaFree(g->sorted64_tmp); g->sorted64_tmp = nullptr;
        g->cap_points = 0;
        const int64_t cap = n_idx + n_idx / 16 + 32;
        PNB_CUDA(cudaMalloc(&g->cell_points, sizeof(int32_t) * (size_t)cap));
        PNB_CUDA(cudaMalloc(&g->sorted64, sizeof(Rec64) * (size_t)cap));
        PNB_CUDA(cudaMalloc(&g->sorted64_tmp, sizeof(Rec64) * (size_t)cap));
        g->cap_points = cap;
    }
    const unsigned blocks = (unsigned)div_up(n_idx > 0 ? n_idx : 1, 256);
#define PNB_ND64(K, ...)                                                        \
    switch (g->p64.ndims) {                                                     \
        case 1: K<1><<<blocks, 256, 0, s>>>(__VA_ARGS__); break;                \
        case 2: K<2><<<blocks, 256, 0, s>>>(__VA_ARGS__); break;                \
        default: K<3><<<blocks, 256, 0, s>>>(__VA_ARGS__); break;               \
    }
    if (n_idx > 0) {
        PNB_ND64(k_hist64, g->p64, y, n_idx, eachindex_y, index_base, g->cell_count, g->d_err);
        PNB_LAUNCHED();
    }
    pnb_status st = scan_impl<uint32_t, true, false>(g, g->cell_count, g->cell_start + 1, C, s);
    if (st != PNB_OK) return st;
    if (n_idx > 0) {
        PNB_ND64(k_scatter64, g->p64, y, n_idx, eachindex_y, index_base, g->cell_start + 1,
                 g->sorted64_tmp);
        PNB_LAUNCHED();
        k_canonicalize64<<<(unsigned)div_up(C * 32, 256), 256, 0, s>>>((int)C, g->cell_start,
                                                                     g->sorted64_tmp, g->sorted64,
                                                                     g->cell_points);
        PNB_LAUNCHED();
    }
#undef PNB_ND64
    st = check_err_word(g, s);
    if (st != PNB_OK) { g->n_built = 0; return st; }
    g->n_built = n_idx;
    g->y_built = y;
    g->n_y_built = n;
    g->full_build = (eachindex_y == nullptr);
    g->built = true;
    g->csr_valid = true;
    g->canonical = true;
    return PNB_OK;
}

extern "C" pnb_status pnb_point_cells_f64(const pnb_grid *g, const double *x, int64_t n,
                                          int32_t *out_linear, void *stream)
{
    if (!g || !g->f64) { set_error("not a Float64 grid handle"); return PNB_ERR_ARG; }
    if (g->template_search) {
        set_error("`search_radius` is not defined for this cell list");
        return PNB_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (n > 0) {
        const unsigned blocks = (unsigned)div_up(n, 256);
        switch (g->p64.ndims) {
            case 1: k_point_cells64<1><<<blocks, 256, 0, s>>>(g->p64, x, n, out_linear); break;
            case 2: k_point_cells64<2><<<blocks, 256, 0, s>>>(g->p64, x, n, out_linear); break;
            default: k_point_cells64<3><<<blocks, 256, 0, s>>>(g->p64, x, n, out_linear); break;
        }
        PNB_LAUNCHED();
    }
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}

// ---- fused closures in Float64 (f64.cuh) ---------------------------------------------------------
static pnb_status closure64_precheck(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                     cudaStream_t s)
{
    if (!g || !g->f64) { set_error("not a Float64 grid handle"); return PNB_ERR_ARG; }
    if (g->p64.mixed) {
        set_error("the fused closures exist for Float32 and Float64 searches; a mixed-precision search "
                  "(Float64 coordinates, Float32 radius) delivers Float32 pos_diff / distance through its "
                  "neighbour lists");
        return PNB_ERR_ARG;
    }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    if (nx > 0 && !x) { set_error("x is NULL"); return PNB_ERR_ARG; }
    return check_built_y(g, y, n, s);
}

extern "C" pnb_status pnb_nbody_f64(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                    const int32_t *points, int64_t n_points, int index_base,
                                    const double *mass, double G, double *dv, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    pnb_status st = closure64_precheck(g, x, nx, y, n, s);
    if (st != PNB_OK) return st;
    const int nd = g->p64.ndims;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(double) * (size_t)nx * nd, s));
    const int64_t n_loop = points ? n_points : nx;
    if (!g->template_search && g->n_built > 0 && n_loop > 0) {
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        const Wcsph64P none{};
        ProfScope ps(PH_SWEEP_POINTS, s);
        switch (nd) {
            case 1: k_sweep_closure_t<1, 0, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
            case 2: k_sweep_closure_t<2, 0, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
            default: k_sweep_closure_t<3, 0, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}

extern "C" pnb_status pnb_wcsph_interact_f64(pnb_grid *g, const double *x, int64_t nx, const double *y,
                                             int64_t n, const int32_t *points, int64_t n_points,
                                             int index_base, const double *v_x, const double *v_y,
                                             const double *mass_x, const double *mass_y,
                                             const double *pressure_x, const double *pressure_y,
                                             const pnb_wcsph_params_f64 *params, double *dv, void *stream)
{
    (void)mass_x;
    cudaStream_t s = (cudaStream_t)stream;
    pnb_status st = closure64_precheck(g, x, nx, y, n, s);
    if (st != PNB_OK) return st;
    if (!params) { set_error("params is NULL"); return PNB_ERR_ARG; }
    const int nd = g->p64.ndims;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(double) * (size_t)nx * (nd + 1), s));
    const int64_t n_loop = points ? n_points : nx;
    if (!g->template_search && g->n_built > 0 && n_loop > 0) {
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        const Wcsph64P prm{params->smoothing_length, params->sound_speed, params->alpha, params->beta,
                           params->epsilon, params->delta, params->kernel_norm};
        ProfScope ps(PH_SWEEP_POINTS, s);
        switch (nd) {
            case 1: k_sweep_closure_t<1, 1, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0, prm, dv, g->d_err); break;
            case 2: k_sweep_closure_t<2, 1, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0, prm, dv, g->d_err); break;
            default: k_sweep_closure_t<3, 1, OpsF64><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0, prm, dv, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}

// The fused closures on a MIXED-precision search (Float64 coordinates, Float32 radius,
// tut_gpu_usage.jl:45-50): the closure receives Float32 pos_diff / distance (nhs_grid.jl:547-555
// under Julia's promotion rules), so with Float32 state arrays the benchmark closures compute in
// Float32 -- the Float32 instantiation of the same kernel (oracle: pno_nbody_mix / pno_wcsph_mix).
static pnb_status closure_mixed_precheck(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                         cudaStream_t s)
{
    if (!g || !g->f64 || !g->p64.mixed) {
        set_error("not a mixed-precision grid handle (pnb_grid_create_mixed)");
        return PNB_ERR_ARG;
    }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    if (nx > 0 && !x) { set_error("x is NULL"); return PNB_ERR_ARG; }
    return check_built_y(g, y, n, s);
}

extern "C" pnb_status pnb_nbody_mixed(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                      const int32_t *points, int64_t n_points, int index_base,
                                      const float *mass, float G, float *dv, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    pnb_status st = closure_mixed_precheck(g, x, nx, y, n, s);
    if (st != PNB_OK) return st;
    const int nd = g->p64.ndims;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)nx * nd, s));
    const int64_t n_loop = points ? n_points : nx;
    if (!g->template_search && g->n_built > 0 && n_loop > 0) {
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        const WcsphTP<float> none{};
        ProfScope ps(PH_SWEEP_POINTS, s);
        switch (nd) {
            case 1: k_sweep_closure_t<1, 0, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
            case 2: k_sweep_closure_t<2, 0, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
            default: k_sweep_closure_t<3, 0, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, nullptr, nullptr, mass, nullptr, nullptr, G, none, dv, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}

extern "C" pnb_status pnb_wcsph_interact_mixed(pnb_grid *g, const double *x, int64_t nx, const double *y,
                                               int64_t n, const int32_t *points, int64_t n_points,
                                               int index_base, const float *v_x, const float *v_y,
                                               const float *mass_x, const float *mass_y,
                                               const float *pressure_x, const float *pressure_y,
                                               const pnb_wcsph_params *params, float *dv, void *stream)
{
    (void)mass_x;
    cudaStream_t s = (cudaStream_t)stream;
    pnb_status st = closure_mixed_precheck(g, x, nx, y, n, s);
    if (st != PNB_OK) return st;
    if (!params) { set_error("params is NULL"); return PNB_ERR_ARG; }
    const int nd = g->p64.ndims;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)nx * (nd + 1), s));
    const int64_t n_loop = points ? n_points : nx;
    if (!g->template_search && g->n_built > 0 && n_loop > 0) {
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        const WcsphTP<float> prm{params->smoothing_length, params->sound_speed, params->alpha, params->beta,
                                 params->epsilon, params->delta, params->kernel_norm};
        ProfScope ps(PH_SWEEP_POINTS, s);
        switch (nd) {
            case 1: k_sweep_closure_t<1, 1, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0f, prm, dv, g->d_err); break;
            case 2: k_sweep_closure_t<2, 1, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0f, prm, dv, g->d_err); break;
            default: k_sweep_closure_t<3, 1, OpsF32><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, v_x, v_y, mass_y, pressure_x, pressure_y, 0.0f, prm, dv, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}

extern "C" pnb_status pnb_count_neighbors_f64(pnb_grid *g, const double *x, int64_t nx,
                                              const double *y, int64_t n, const int32_t *points,
                                              int64_t n_points, int index_base, int64_t *out,
                                              void *stream)
{
    if (!g || !g->f64) { set_error("not a Float64 grid handle"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    { pnb_status sy = check_built_y(g, y, n, (cudaStream_t)stream); if (sy != PNB_OK) return sy; }
    cudaStream_t s = (cudaStream_t)stream;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * (size_t)nx, s));
    const int64_t n_loop = points ? n_points : nx;
    if (!g->template_search && g->n_built > 0 && n_loop > 0) {
        const unsigned blocks = (unsigned)div_up(n_loop, 128);
        ProfScope ps(PH_SWEEP_POINTS, s);
        switch (g->p64.ndims) {
            case 1: k_sweep_points64<1, 0><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, out, nullptr, nullptr, nullptr, g->d_err); break;
            case 2: k_sweep_points64<2, 0><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, out, nullptr, nullptr, nullptr, g->d_err); break;
            default: k_sweep_points64<3, 0><<<blocks, 128, 0, s>>>(g->p64, g->cell_start, g->sorted64, x, n_loop, points, index_base, out, nullptr, nullptr, nullptr, g->d_err); break;
        }
        PNB_LAUNCHED();
    }
    return check_err_word(g, s);
}
