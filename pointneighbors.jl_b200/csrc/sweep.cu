// sweep.cu -- C-ABI entry points of the fused sweeps (count, n-body, WCSPH).
// reference: foreach_point_neighbor (src/neighborhood_search.jl:183-201) with the closures of
// benchmarks/count_neighbors.jl, benchmarks/n_body.jl, benchmarks/smoothed_particle_hydrodynamics.jl.
#include <cstdlib>

#include "closures.cuh"
#include "sweep.cuh"
#include "sweep_tiles.cuh"
#include "sweep_launch.cuh"

namespace pnb {

// payload of the neighbour points gathered into cell order (one coalesced pass), so that the
// staging loops of k_sweep_cells read contiguous memory.
// Slots of the cell list: CSR -> records 0 .. n-1; buckets -> C * K slots of which the first
// start[cell] of every cell are records.  The payload arrays use the same slot numbering.
__device__ __forceinline__ bool slot_valid(const CellsView &v, int64_t slot)
{
    if (v.K == 0) return true;
    const uint32_t sl = (uint32_t)slot;             // C * K < 2^32; K is a power of two
    return (sl & (v.K - 1u)) < v.start[sl >> (31 - __clz((int)v.K))];
}

// ids are the .w field of the cell-ordered records (no separate id list on this path)
__global__ void k_gather_f32(int64_t n_slots, CellsView cv, const float *__restrict__ src,
                             float *__restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_slots && slot_valid(cv, i)) dst[i] = __ldg(src + __float_as_int(__ldg(&cv.rec[i].w)));
}

// exact mode: mp = (mass, pressure, 1/rho, mass/rho); fast mode: vp = (mass/rho, pressure) only
__global__ void k_gather_wcsph(int64_t n, int nd, CellsView cv,
                               const float *__restrict__ v, const float *__restrict__ mass,
                               const float *__restrict__ pressure, float4 *__restrict__ vrho,
                               float4 *__restrict__ mp, float2 *__restrict__ vp, int64_t i0 = 0)
{
    int64_t i = i0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !slot_valid(cv, i)) return;
    const int64_t id = __float_as_int(__ldg(&cv.rec[i].w));
    const int ns = nd + 1;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nd == 3 && ((reinterpret_cast<uintptr_t>(v) & 15) == 0)) {
        a = __ldg(reinterpret_cast<const float4 *>(v) + id);   // (vx, vy, vz, rho) is one 16 B record
    } else {
        a.x = __ldg(v + id * ns);
        if (nd > 1) a.y = __ldg(v + id * ns + 1);
        if (nd > 2) a.z = __ldg(v + id * ns + 2);
        a.w = __ldg(v + id * ns + nd);
    }
    vrho[i] = a;
    const float m = __ldg(mass + id);
    // 1/rho_b and m_b/rho_b once per neighbour instead of once per pair (fast path)
    if (mp) mp[i] = make_float4(m, __ldg(pressure + id), __fdiv_rn(1.f, a.w), __fdiv_rn(m, a.w));
    else vp[i] = make_float2(__fdiv_rn(m, a.w), __ldg(pressure + id));
}

// pnb_set_exact_arithmetic: 0 (default) = fast per-pair terms (MUFU + FMA, |error| << 1e-5),
// 1 = the oracle's IEEE operation sequence (sums bit-identical to the oracle).
static int g_exact_arithmetic = 0;
// measurement overrides of the tile sweep (0 = closure default): warps per cell, fp16 pre-filter
int g_tune_wpc = 0;
int g_tune_half = -1;
int g_tune_twoset = 1;
int g_tune_left = getenv("PNB_SWEEP_LEFT") ? atoi(getenv("PNB_SWEEP_LEFT")) : 1;
int g_reserve_ctas = 0;
int g_tune_flat = getenv("PNB_SWEEP_FLAT") ? atoi(getenv("PNB_SWEEP_FLAT")) : 1;

static bool is_fast_path(const pnb_grid *g, const void *x, int64_t nx, const int32_t *points);

// number of record slots of the view the sweeps will use (payload scratch is sized by it)
static int64_t view_slots(const pnb_grid *g)
{
    return g->bucket_valid ? (int64_t)g->p.total_cells * g->bucket_K : g->n_built;
}

static pnb_status sweep_precheck(pnb_grid *g, const void *x, int64_t nx, const void *y, int64_t n,
                                 const int32_t *points, int64_t *n_loop, cudaStream_t s)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (g->f64) { set_error("Float64 grid handle passed to a Float32 entry point"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    if (nx > 0 && !x) { set_error("x is NULL"); return PNB_ERR_ARG; }
    // the reference reads neighbor_coords live (src/nhs_grid.jl:543-548): another array than the
    // one of the last build refreshes the cell-ordered snapshot (grid.cu, check_built_y)
    { pnb_status sy = check_built_y(g, y, n, s); if (sy != PNB_OK) return sy; }
    *n_loop = points ? *n_loop : nx;
    // every sweep but the x === y tile sweep walks (or may walk) the CSR arrays: settle the
    // layout BEFORE the payload is gathered into it
    if (!is_fast_path(g, x, nx, points) && g->bucket_valid) {
        pnb_status st = ensure_csr(g, s);
        if (st != PNB_OK) return st;
        g->bucket_valid = false;
    }
    return PNB_OK;
}

static bool is_fast_path(const pnb_grid *g, const void *x, int64_t nx, const int32_t *points)
{
    return points == nullptr && g->full_build && !g->y_refreshed && x == g->y_built && nx == g->n_y_built;
}

}  // namespace pnb

using namespace pnb;

// end of a sweep entry point: blocking calls synchronise and translate the error word (a
// stream-ordered update! that overflowed a bucket is rebuilt there -> PNB_RETRY_INTERNAL, the
// wrappers below repeat the sweep once); the *_async entry points return without looking
static thread_local bool t_async_sweep = false;
static pnb_status finish_sweep(pnb_grid *g, cudaStream_t s)
{
    if (t_async_sweep) return PNB_OK;
    return check_err_word(g, s);
}

extern "C" void pnb_set_exact_arithmetic(int on) { g_exact_arithmetic = on != 0; }
extern "C" int pnb_get_exact_arithmetic(void) { return g_exact_arithmetic; }
extern "C" void pnb_set_twoset_tiles(int on) { g_tune_twoset = on; }
extern "C" void pnb_set_sweep_left(int mode) { g_tune_left = mode; }
extern "C" void pnb_set_sweep_kernel(int flat) { g_tune_flat = flat; }
extern "C" void pnb_set_sweep_reserve(int ctas) { g_reserve_ctas = ctas; }
extern "C" void pnb_set_tuning(int warps_per_cell, int half_prefilter)
{
    g_tune_wpc = warps_per_cell;
    g_tune_half = half_prefilter;
}

static pnb_status count_neighbors_impl(pnb_grid *g, const float *x, int64_t nx,
                                              const float *y, int64_t n, const int32_t *points,
                                              int64_t n_points, int index_base, int64_t *out,
                                              void *stream)
{
    int64_t n_loop = n_points;
    pnb_status st = sweep_precheck(g, x, nx, y, n, points, &n_loop, (cudaStream_t)stream);
    if (st != PNB_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    // count_neighbors.jl:22  n_neighbors .= 0
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(out, 0, sizeof(int64_t) * (size_t)nx, s));
    CountCl cl{out};
    st = launch_sweep(g, is_fast_path(g, x, nx, points), true, x, n_loop, points, index_base, cl, s);
    if (st != PNB_OK) return st;
    return finish_sweep(g, s);
}

static pnb_status nbody_impl(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                    int64_t n, const int32_t *points, int64_t n_points,
                                    int index_base, const float *mass, float G, float *dv,
                                    void *stream)
{
    int64_t n_loop = n_points;
    pnb_status st = sweep_precheck(g, x, nx, y, n, points, &n_loop, (cudaStream_t)stream);
    if (st != PNB_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    const int nd = g->p.ndims;
    // n_body.jl:36  dv .= 0
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)nx * nd, s));
    if (g->template_search || g->n_built == 0) return finish_sweep(g, s);
    // bit-identical sums need the reference's visiting order with ids ascending in a cell
    if (g_exact_arithmetic && (st = ensure_canonical(g, s)) != PNB_OK) return st;
    const int64_t slots = view_slots(g);
    st = ensure_scratch(g, sizeof(float) * (size_t)slots);
    if (st != PNB_OK) return st;
    float *mass_sorted = reinterpret_cast<float *>(g->scratch);
    {
        ProfScope ps(PH_GATHER, s);
        k_gather_f32<<<(unsigned)div_up(slots, 256), 256, 0, s>>>(slots, cells_view(g), mass,
                                                               mass_sorted);
        PNB_LAUNCHED();
    }
    const bool fastp = is_fast_path(g, x, nx, points);
    if (g_exact_arithmetic)
        st = launch_sweep(g, fastp, false, x, n_loop, points, index_base, NBodyClT<true>{mass_sorted, -G, dv, nd}, s);
    else
        st = launch_sweep(g, fastp, true, x, n_loop, points, index_base, NBodyClT<false>{mass_sorted, -G, dv, nd}, s);
    if (st != PNB_OK) return st;
    return finish_sweep(g, s);
}

static pnb_status wcsph_interact_impl(pnb_grid *g, const float *x, int64_t nx,
                                             const float *y, int64_t n, const int32_t *points,
                                             int64_t n_points, int index_base, const float *v_x,
                                             const float *v_y, const float *mass_x,
                                             const float *mass_y, const float *pressure_x,
                                             const float *pressure_y,
                                             const pnb_wcsph_params *params, float *dv,
                                             void *stream)
{
    (void)mass_x;
    int64_t n_loop = n_points;
    pnb_status st = sweep_precheck(g, x, nx, y, n, points, &n_loop, (cudaStream_t)stream);
    if (st != PNB_OK) return st;
    if (!params) { set_error("params is NULL"); return PNB_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    const int nd = g->p.ndims;
    if (nx > 0) PNB_CUDA(cudaMemsetAsync(dv, 0, sizeof(float) * (size_t)nx * (nd + 1), s));
    if (g->template_search || g->n_built == 0) return finish_sweep(g, s);
    if (g_exact_arithmetic && (st = ensure_canonical(g, s)) != PNB_OK) return st;
    const int64_t nb = view_slots(g);
    const int64_t off_mp = ((int64_t)sizeof(float4) * nb + 255) / 256 * 256;
    st = ensure_scratch(g, off_mp + (int64_t)sizeof(float4) * nb);
    if (st != PNB_OK) return st;
    float4 *vrho = reinterpret_cast<float4 *>(g->scratch);
    float4 *mp = g_exact_arithmetic
        ? reinterpret_cast<float4 *>(reinterpret_cast<unsigned char *>(g->scratch) + off_mp) : nullptr;
    float2 *vp = g_exact_arithmetic
        ? nullptr : reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(g->scratch) + off_mp);
    {
        ProfScope ps(PH_GATHER, s);
        k_gather_wcsph<<<(unsigned)div_up(nb, 256), 256, 0, s>>>(nb, nd, cells_view(g), v_y,
                                                                 mass_y, pressure_y, vrho, mp, vp);
        PNB_LAUNCHED();
    }
    const bool fastp = is_fast_path(g, x, nx, points);
    const float h = params->smoothing_length;
    const float inv_h = -0.5f / h, kh = -5.0f * params->kernel_norm / (h * h);
    const float ac = params->alpha * params->sound_speed;
    const float dhc2 = 2.0f * params->delta * h * params->sound_speed;
    const bool in_radius = 2.0f * h <= g->p.r;     // kernel support inside the search radius
    if (g_exact_arithmetic)
        st = launch_sweep(g, fastp, false, x, n_loop, points, index_base,
                          WcsphClT<true>{vrho, mp, v_x, pressure_x, *params, dv, nd, inv_h, kh, ac, dhc2, false, nullptr}, s);
    else
        st = launch_sweep(g, fastp, true, x, n_loop, points, index_base,
                          WcsphClT<false>{vrho, nullptr, v_x, pressure_x, *params, dv, nd, inv_h, kh, ac, dhc2, in_radius, vp}, s);
    if (st != PNB_OK) return st;
    return finish_sweep(g, s);
}

// ---- C entry points: blocking (the reference's semantics, src/util.jl:166-170) and stream-ordered
#define PNB_RETRY_ONCE(call)                                                                  \
    do {                                                                                      \
        pnb_status st__ = (call);                                                             \
        if (st__ == PNB_RETRY_INTERNAL) st__ = (call);                                        \
        if (st__ == PNB_RETRY_INTERNAL) { set_error("cell list changed during the sweep"); st__ = PNB_ERR_STATE; } \
        return st__;                                                                          \
    } while (0)

extern "C" pnb_status pnb_count_neighbors_f32(pnb_grid *g, const float *x, int64_t nx,
                                              const float *y, int64_t n, const int32_t *points,
                                              int64_t n_points, int index_base, int64_t *out,
                                              void *stream)
{
    PNB_RETRY_ONCE(count_neighbors_impl(g, x, nx, y, n, points, n_points, index_base, out, stream));
}

extern "C" pnb_status pnb_nbody_f32(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                    int64_t n, const int32_t *points, int64_t n_points,
                                    int index_base, const float *mass, float G, float *dv,
                                    void *stream)
{
    PNB_RETRY_ONCE(nbody_impl(g, x, nx, y, n, points, n_points, index_base, mass, G, dv, stream));
}

extern "C" pnb_status pnb_wcsph_interact_f32(pnb_grid *g, const float *x, int64_t nx,
                                             const float *y, int64_t n, const int32_t *points,
                                             int64_t n_points, int index_base, const float *v_x,
                                             const float *v_y, const float *mass_x,
                                             const float *mass_y, const float *pressure_x,
                                             const float *pressure_y,
                                             const pnb_wcsph_params *params, float *dv,
                                             void *stream)
{
    PNB_RETRY_ONCE(wcsph_interact_impl(g, x, nx, y, n, points, n_points, index_base, v_x, v_y, mass_x,
                                       mass_y, pressure_x, pressure_y, params, dv, stream));
}

// Stream-ordered form of the x === y WCSPH sweep (all points): everything is enqueued on
// `stream`, nothing is synchronised and no error word is read; pnb_grid_check(g, stream) settles
// both this sweep and a preceding pnb_grid_build_async_f32.
extern "C" pnb_status pnb_wcsph_interact_async_f32(pnb_grid *g, const float *y, int64_t n,
                                                   const float *v, const float *mass,
                                                   const float *pressure,
                                                   const pnb_wcsph_params *params, float *dv,
                                                   void *stream)
{
    if (!g) { set_error("grid handle is NULL"); return PNB_ERR_ARG; }
    if (!is_fast_path(g, y, n, nullptr) || y != g->y_built) {
        set_error("the stream-ordered sweep needs the coordinates of the last update! (x === y)");
        return PNB_ERR_STATE;
    }
    t_async_sweep = true;
    const pnb_status st = wcsph_interact_impl(g, y, n, y, n, nullptr, 0, 0, v, v, mass, mass, pressure,
                                              pressure, params, dv, stream);
    t_async_sweep = false;
    return st;
}

// The WCSPH sweep of the cell layers [cz_a, cz_b] and [cz_c, cz_d] (local 1-based coordinates of
// the LAST used dimension; an empty range has first > last) only, stream-ordered: the payload of
// the layers [cz_a - 1, cz_b + 1] and [cz_c - 1, cz_d + 1] is gathered first.  The overlapped
// multi-GPU step sweeps the interior layers of a slab while the ghost layers are still in flight,
// then appends them (pnb_grid_append_f32) and sweeps the boundary layers of both sides with one
// launch.  Needs the one-pass (bucket) layout.
extern "C" pnb_status pnb_wcsph_interact_layers_async_f32(pnb_grid *g, const float *y, int64_t n,
                                                          const float *v, const float *mass,
                                                          const float *pressure,
                                                          const pnb_wcsph_params *params, float *dv,
                                                          int cz_a, int cz_b, int cz_c, int cz_d,
                                                          int mode, void *stream)
{
    // mode 0: gather the payload of the layers [cz_a - 1, cz_b + 1] and [cz_c - 1, cz_d + 1], then
    //         sweep [cz_a, cz_b] and [cz_c, cz_d];  1: ONLY gather exactly [cz_a, cz_b] and
    //         [cz_c, cz_d];  2: ONLY sweep (the payload of the neighbouring layers is in place)
    if (!g || !params) { set_error("NULL argument"); return PNB_ERR_ARG; }
    if (g->f64 || g->hashed || g->p.periodic || !g->built || !g->bucket_valid || g->bucket_tr ||
        !g->full_build || y != g->y_built || n != g->n_y_built || g_exact_arithmetic) {
        set_error("the layered sweep needs the one-pass (bucket) build of exactly these coordinates, "
                  "a non-periodic FullGridCellList and the fast arithmetic mode");
        return PNB_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nd = g->p.ndims;
    if (nd < 2) { set_error("the layered sweep needs 2 or 3 dimensions"); return PNB_ERR_ARG; }
    const int gl = g->p.gs[nd - 1];
    auto clip = [&](int &a, int &b) {
        if (mode == 1) { if (a < 1) a = 1; if (b > gl) b = gl; }
        else { if (a < 2) a = 2; if (b > gl - 1) b = gl - 1; }
    };
    clip(cz_a, cz_b);
    clip(cz_c, cz_d);
    const int64_t nb = view_slots(g);
    const int64_t off_mp = ((int64_t)sizeof(float4) * nb + 255) / 256 * 256;
    pnb_status st = ensure_scratch(g, off_mp + (int64_t)sizeof(float4) * nb);
    if (st != PNB_OK) return st;
    float4 *vrho = reinterpret_cast<float4 *>(g->scratch);
    float2 *vp = reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(g->scratch) + off_mp);
    int64_t layer_cells = 1;
    for (int d = 0; d < nd - 1; d++) layer_cells *= g->p.gs[d];
    auto gather = [&](int ga, int gb) -> pnb_status {
        if (ga < 1) ga = 1;
        if (gb > gl) gb = gl;
        if (ga > gb) return PNB_OK;
        const int64_t i0 = (int64_t)(ga - 1) * layer_cells * g->bucket_K;
        const int64_t i1 = (int64_t)gb * layer_cells * g->bucket_K;
        ProfScope ps(PH_GATHER, s);
        k_gather_wcsph<<<(unsigned)div_up(i1 - i0, 256), 256, 0, s>>>(i1, nd, cells_view(g), v, mass, pressure,
                                                                     vrho, nullptr, vp, i0);
        PNB_LAUNCHED();
        return PNB_OK;
    };
    const bool has1 = cz_a <= cz_b, has2 = cz_c <= cz_d;
    if (!has1 && !has2) return PNB_OK;
    if (mode != 2) {
        const int w = mode == 1 ? 0 : 1;          // mode 0 also gathers the neighbouring layers
        if (has1 && has2 && cz_c - w <= cz_b + w + 1 && cz_a - w <= cz_d + w + 1) {
            st = gather(min(cz_a, cz_c) - w, max(cz_b, cz_d) + w);      // the two ranges touch
        } else {
            st = has1 ? gather(cz_a - w, cz_b + w) : PNB_OK;
            if (st == PNB_OK && has2) st = gather(cz_c - w, cz_d + w);
        }
        if (st != PNB_OK) return st;
        if (mode == 1) return PNB_OK;
    }
    const float h = params->smoothing_length;
    const float inv_h = -0.5f / h, kh = -5.0f * params->kernel_norm / (h * h);
    const float ac = params->alpha * params->sound_speed;
    const float dhc2 = 2.0f * params->delta * h * params->sound_speed;
    const bool in_radius = 2.0f * h <= g->p.r;
    const WcsphClT<false> cl{vrho, nullptr, v, pressure, *params, dv, nd, inv_h, kh, ac, dhc2, in_radius, vp};
    const CellsView cv = cells_view(g);
    const int l0 = has1 ? cz_a - 2 : 0, n0 = has1 ? cz_b - cz_a + 1 : 0;
    const int l1 = has2 ? cz_c - 2 : 0, n1 = has2 ? cz_d - cz_c + 1 : 0;
    if (nd == 2) return launch_flat<2, false, WcsphClT<false>, false>(g, cv, cv, n, cl, l0, n0, s, l1, n1);
    return launch_flat<3, false, WcsphClT<false>, false>(g, cv, cv, n, cl, l0, n0, s, l1, n1);
}
