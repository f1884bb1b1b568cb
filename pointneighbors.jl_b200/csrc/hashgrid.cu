// hashgrid.cu -- SpatialHashingCellList behind a GridNeighborhoodSearch.
//
// reference: src/cell_lists/spatial_hashing.jl (table, push_cell!, spatial_hash,
// coordinates_flattened), src/nhs_grid.jl:479-513 (check_collision / check_cell_collision),
// :519-575 (the sweep), :622-638 (cell_coords without a min corner), src/gpu.jl:37-44 (adapt).
//
// The reference keeps a table of list_size point lists indexed by spatial_hash(cell), the cell
// stored by the first insertion (`coords`) and a collision flag per entry.  Here the table is the
// same CSR structure as the full grid with the hash key in place of the linear cell index
// (key_start[list_size + 1] + records (x, y, z, id) in key order, ids ascending inside a key) and
// one int4 (c1, c2, c3, flags) per key for `coords` / `collisions`.  The build is the same
// two-pass counting sort; coords / collisions are derived from the finished lists by walking
// each key's points in ascending id order, which reproduces the reference's serial insertion
// (push_cell!, spatial_hashing.jl:79-97) exactly -- including its use of 0 as the "unused"
// marker, which coincides with the flattened cell (0, 0, 0).
#include <cstring>

#include "hashgrid.cuh"

using namespace pnb;

namespace pnb {

// table keys of all points: histogram (pass 1) / scatter (pass 2); a point whose cell does not
// fit Int32 is the reference's InexactError (coordinates_flattened, :176-183) -> error bit 4
template <int ND, bool PER, bool SCATTER>
__global__ void __launch_bounds__(256)
k_hash_points(GridP g, const float *__restrict__ y, int64_t n_idx, const int32_t *__restrict__ idx,
              int base, uint32_t *__restrict__ count, const uint32_t *__restrict__ key_start,
              float4 *__restrict__ records, int *__restrict__ err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_idx) return;
    const int64_t id = idx ? (int64_t)idx[k] - base : k;
    float p[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < ND; d++) p[d] = __ldg(y + id * ND + d);
    long long cc[3];
    const bool ok = hash_cell_coords<ND, PER>(g, p, cc);
    if (!ok) { atomicOr(err, 16); return; }
    const uint32_t key = spatial_hash_key<ND>(cc, g.total_cells);
    if (!SCATTER) {
        atomicAdd(count + key, 1u);
    } else {
        const uint32_t pos = key_start[key] + atomicAdd(count + key, 1u);
        records[pos] = make_float4(p[0], p[1], p[2], __int_as_float((int)id));
    }
}

// One warp per 32 keys: a table of list_size = 2 N keys (the recommended size) has one occupied
// key in ~50, so every lane first looks at ITS key and the warp then works through the occupied
// ones together (one warp per key would launch 64 N threads to find 98 % of the keys empty).
// Per occupied key: sort the list by point id (canonical order, into `out` / `cell_points`) and
// derive coords / collisions from it.
template <int ND, bool PER>
__global__ void __launch_bounds__(256)
k_hash_finish(GridP g, const uint32_t *__restrict__ key_start, const float4 *__restrict__ in,
              float4 *__restrict__ out, int32_t *__restrict__ cell_points, int4 *__restrict__ meta)
{
    const int lane = lane_id();
    const int64_t key0 = (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32;
    const int64_t my_key = key0 + lane;
    uint32_t my_s0 = 0, my_s1 = 0;
    if (my_key < g.total_cells) { my_s0 = key_start[my_key]; my_s1 = key_start[my_key + 1]; }
    if (my_key < g.total_cells && my_s0 == my_s1) meta[my_key] = make_int4(0, 0, 0, 0);
    unsigned todo = __ballot_sync(0xffffffffu, my_s1 > my_s0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        const uint32_t s0 = __shfl_sync(0xffffffffu, my_s0, src);
        const uint32_t s1 = __shfl_sync(0xffffffffu, my_s1, src);
        const int cnt = (int)(s1 - s0);
        // ---- ids ascending inside the key (rank by counting) ----
        if (cnt <= 32) {
            float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
            int v = 0x7fffffff;
            if (lane < cnt) { rec = in[s0 + lane]; v = __float_as_int(rec.w); }
            int r = 0;
            for (int k = 0; k < cnt; k++) r += (__shfl_sync(0xffffffffu, v, k) < v) ? 1 : 0;
            if (lane < cnt) { cell_points[s0 + r] = v; out[s0 + r] = rec; }
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
        __syncwarp();
        // ---- coords / collisions from the list in ascending id order (push_cell!, :79-97):
        // `coords` is overwritten while it is still 0, i.e. it ends up as the cell F of the first
        // point whose flattened cell is not 0; every later point from another cell is a collision
        uint32_t first = 0xffffffffu;
        for (uint32_t e = s0 + lane; e < s1 && first == 0xffffffffu; e += 32) {
            const float4 r4 = out[e];
            const float p[3] = {r4.x, r4.y, r4.z};
            long long cc[3];
            hash_cell_coords<ND, PER>(g, p, cc);
            if (cc[0] != 0 || cc[1] != 0 || cc[2] != 0) first = e;
        }
        first = __reduce_min_sync(0xffffffffu, first);
        int4 m = make_int4(0, 0, 0, 0);
        if (first != 0xffffffffu) {
            const float4 r4 = out[first];
            const float p[3] = {r4.x, r4.y, r4.z};
            long long cf[3];
            hash_cell_coords<ND, PER>(g, p, cf);
            m.x = (int)cf[0]; m.y = (int)cf[1]; m.z = (int)cf[2];
            int coll = 0;
            for (uint32_t e = first + 1 + lane; e < s1 && !coll; e += 32) {
                const float4 q = out[e];
                const float pq[3] = {q.x, q.y, q.z};
                long long cc[3];
                hash_cell_coords<ND, PER>(g, pq, cc);
                if (cc[0] != cf[0] || cc[1] != cf[1] || cc[2] != cf[2]) coll = 1;
            }
            m.w = __any_sync(0xffffffffu, coll) ? 1 : 0;
        }
        if (lane == 0) meta[key0 + src] = m;
    }
}

// the table in the reference's field layout: coords = Vector{UInt128} (4 little-endian 32-bit
// words per entry, coordinates_flattened :176-183), collisions = Vector{Bool}
__global__ void k_hash_export(int64_t L, const int4 *__restrict__ meta, uint32_t *__restrict__ coords,
                              uint8_t *__restrict__ collisions)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= L) return;
    const int4 m = meta[k];
    if (coords) {
        coords[4 * k + 0] = (uint32_t)m.x;
        coords[4 * k + 1] = (uint32_t)m.y;
        coords[4 * k + 2] = (uint32_t)m.z;
        coords[4 * k + 3] = 0u;
    }
    if (collisions) collisions[k] = (uint8_t)(m.w & 1);
}

template <int ND, bool PER>
static pnb_status hash_build_nd(pnb_grid *g, const float *y, const int32_t *idx, int64_t n_idx,
                                int base, cudaStream_t s)
{
    const int64_t L = g->p.total_cells;
    const unsigned blocks = (unsigned)div_up(n_idx > 0 ? n_idx : 1, 256);
    if (n_idx > 0) {
        ProfScope ps(PH_BUILD_CELL_COUNT, s);
        k_hash_points<ND, PER, false><<<blocks, 256, 0, s>>>(g->p, y, n_idx, idx, base, g->cell_count,
                                                            nullptr, nullptr, g->d_err);
        PNB_LAUNCHED();
    }
    pnb_status st = exclusive_scan_u32(g, g->cell_count, g->cell_start, L, s);
    if (st != PNB_OK) return st;
    PNB_CUDA(cudaMemsetAsync(g->cell_count, 0, sizeof(uint32_t) * (size_t)L, s));
    if (n_idx > 0) {
        ProfScope ps(PH_BUILD_SCATTER, s);
        k_hash_points<ND, PER, true><<<blocks, 256, 0, s>>>(g->p, y, n_idx, idx, base, g->cell_count,
                                                           g->cell_start, g->sorted, g->d_err);
        PNB_LAUNCHED();
    }
    // the histogram is scratch: zero between builds
    PNB_CUDA(cudaMemsetAsync(g->cell_count, 0, sizeof(uint32_t) * (size_t)L, s));
    return PNB_OK;
}

template <int ND, bool PER>
static pnb_status hash_finish_nd(pnb_grid *g, cudaStream_t s)
{
    const int64_t L = g->p.total_cells;
    if (!g->sorted_alt) PNB_CUDA(cudaMalloc(&g->sorted_alt, sizeof(float4) * (size_t)g->cap_points));
    if (!g->cell_points) PNB_CUDA(cudaMalloc(&g->cell_points, sizeof(int32_t) * (size_t)g->cap_points));
    ProfScope ps(PH_BUILD_FINALIZE, s);
    k_hash_finish<ND, PER><<<(unsigned)div_up(div_up(L, 32) * 32, 256), 256, 0, s>>>(
        g->p, g->cell_start, g->sorted, g->sorted_alt, g->cell_points, g->hmeta);
    PNB_LAUNCHED();
    float4 *t = g->sorted; g->sorted = g->sorted_alt; g->sorted_alt = t;
    g->canonical = true;
    return PNB_OK;
}

pnb_status hash_build(pnb_grid *g, const float *y, int64_t n, const int32_t *idx, int64_t n_idx,
                      int base, cudaStream_t s)
{
    const int64_t L = g->p.total_cells;
    g->bucket_valid = false;
    g->csr_valid = false;
    pnb_status st = ensure_point_capacity(g, n_idx);
    if (st != PNB_OK) return st;
    const bool per = g->p.periodic != 0;
    switch (g->p.ndims) {
        case 1: st = per ? hash_build_nd<1, true>(g, y, idx, n_idx, base, s)
                         : hash_build_nd<1, false>(g, y, idx, n_idx, base, s); break;
        case 2: st = per ? hash_build_nd<2, true>(g, y, idx, n_idx, base, s)
                         : hash_build_nd<2, false>(g, y, idx, n_idx, base, s); break;
        default: st = per ? hash_build_nd<3, true>(g, y, idx, n_idx, base, s)
                          : hash_build_nd<3, false>(g, y, idx, n_idx, base, s); break;
    }
    if (st != PNB_OK) return st;
    // a failed build leaves an unusable cell list behind, like the reference
    PNB_CUDA(cudaStreamSynchronize(s));
    const int e = *(volatile int *)g->h_err;
    if (e & 16) {
        *(volatile int *)g->h_err = 0;
        g->n_built = 0;
        set_error("InexactError: a cell coordinate does not fit Int32 (coordinates_flattened, "
                  "src/cell_lists/spatial_hashing.jl:176-183): coordinates are NaN or too large");
        return PNB_ERR_DOMAIN;
    }
    g->csr_valid = true;
    g->n_built = n_idx;
    g->y_built = y;
    g->n_y_built = n;
    g->full_build = (idx == nullptr);
    g->built = true;
    g->canonical = false;
    // ids ascending inside every key (the serial insertion order that `coords` / `collisions` and
    // the visiting order of the sweeps are defined by) + the table's coords / collisions
    if (L > 0) {
        switch (g->p.ndims) {
            case 1: st = per ? hash_finish_nd<1, true>(g, s) : hash_finish_nd<1, false>(g, s); break;
            case 2: st = per ? hash_finish_nd<2, true>(g, s) : hash_finish_nd<2, false>(g, s); break;
            default: st = per ? hash_finish_nd<3, true>(g, s) : hash_finish_nd<3, false>(g, s); break;
        }
        if (st != PNB_OK) return st;
    }
    return check_err_word(g, s);   // also synchronizes: initialize!/update! are blocking calls
}

}  // namespace pnb

// host: spatial_hash (src/cell_lists/spatial_hashing.jl:159-174), 0-based key
extern "C" int64_t pnb_spatial_hash(int ndims, const int64_t *cell, int64_t list_size)
{
    if (!cell || list_size <= 0 || ndims < 1 || ndims > 3) return -1;
    uint64_t h = (uint64_t)cell[0] * 73856093ULL;
    if (ndims > 1) h ^= (uint64_t)cell[1] * 19349663ULL;
    if (ndims > 2) h ^= (uint64_t)cell[2] * 83492791ULL;
    int64_t m = (int64_t)h % list_size;
    return m < 0 ? m + list_size : m;
}

extern "C" pnb_status pnb_grid_create_hashed_f32(int ndims, float r, int64_t list_size,
                                                 const float *box_min, const float *box_max,
                                                 pnb_grid **out)
{
    if (!out) { set_error("out is NULL"); return PNB_ERR_ARG; }
    *out = nullptr;
    if (ndims < 1 || ndims > 3) {
        set_error("SpatialHashingCellList: NDIMS must be 1, 2 or 3 (coordinates_flattened holds at most 3 coordinates)");
        return PNB_ERR_ARG;
    }
    if (list_size < 1 || list_size > 0x7ffffff0LL) {
        set_error("SpatialHashingCellList: list_size must be in 1 .. 2^31 - 16");
        return PNB_ERR_ARG;
    }
    if (pnb_device_count() <= 0) {
        set_error("no CUDA device: libpnb200 has no CPU fallback");
        return PNB_ERR_CUDA;
    }
    pnb_grid *g = new pnb_grid();
    memset(g, 0, sizeof(*g));
    // GridNeighborhoodSearch constructor (src/nhs_grid.jl:100-126): cell_size = r, or
    // box.size ./ n_cells with the periodic rule; the corners play no role
    const float zero[3] = {0.f, 0.f, 0.f};
    float pmin[3], pmax[3];
    pnb_status st = pnb_grid_params_f32(ndims, r, zero, zero, box_min, box_max, pmin, pmax,
                                        g->grid_size, g->n_cells, g->cell_size);
    if (st != PNB_OK) { delete g; return st; }
    g->hashed = true;
    g->template_search = (double)r < 2.220446049250313e-16;
    GridP &p = g->p;
    p.ndims = ndims;
    p.hashed = 1;
    p.periodic = (box_min && box_max && !g->template_search) ? 1 : 0;
    p.r = r;
    { volatile float r2 = r * r; p.r2 = r2; }
    float min_size = INFINITY;
    for (int d = 0; d < 3; d++) {
        g->grid_size[d] = 0;
        p.minc[d] = 0.f;
        p.cs[d] = d < ndims ? g->cell_size[d] : 1.f;
        p.gs[d] = 0;
        p.off[d] = 0;
        p.nc[d] = d < ndims ? (int)g->n_cells[d] : -1;
        p.bsize[d] = 1.f;
        if (d < ndims && p.periodic) {
            g->box_min[d] = box_min[d];
            g->box_max[d] = box_max[d];
            volatile float size = box_max[d] - box_min[d];
            p.bsize[d] = size;
            if (size < min_size) min_size = size;
        }
    }
    p.total_cells = (int)list_size;
    p.wrap_d2 = p.periodic ? (0.49f * min_size) * (0.49f * min_size) : INFINITY;
    if (p.periodic && !(p.wrap_d2 > p.r2)) p.wrap_d2 = p.r2;
    st = grid_alloc_common(g, list_size);
    if (st != PNB_OK) { pnb_grid_destroy(g); return st; }
    cudaError_t e = cudaMalloc(&g->hmeta, sizeof(int4) * (size_t)list_size);
    if (e == cudaSuccess) e = cudaMemset(g->hmeta, 0, sizeof(int4) * (size_t)list_size);
    if (e != cudaSuccess) { pnb_grid_destroy(g); return cuda_fail(e, "cudaMalloc hash table"); }
    *out = g;
    return PNB_OK;
}

extern "C" pnb_status pnb_grid_export_hash_table(const pnb_grid *g, uint32_t *coords,
                                                 uint8_t *collisions, void *stream)
{
    if (!g || !g->hashed) { set_error("not a SpatialHashingCellList grid handle"); return PNB_ERR_ARG; }
    if (!g->built) {
        set_error("the neighborhood search has not been initialized (call initialize! first)");
        return PNB_ERR_STATE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t L = g->p.total_cells;
    k_hash_export<<<(unsigned)div_up(L, 256), 256, 0, s>>>(L, g->hmeta, coords, collisions);
    PNB_LAUNCHED();
    PNB_CUDA(cudaStreamSynchronize(s));
    return PNB_OK;
}
