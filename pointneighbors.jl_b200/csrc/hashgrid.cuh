// hashgrid.cuh -- device arithmetic and the sweep kernel of the SpatialHashingCellList path
// (hashgrid.cu; reference: src/cell_lists/spatial_hashing.jl, src/nhs_grid.jl:479-575).
#pragma once

#include "grid.cuh"

namespace pnb {

#ifdef __CUDACC__

// cell_coords of a point for a cell list without corners (src/nhs_grid.jl:622-638):
//   floor_to_int.(coords ./ cell_size)   (saturating, src/util.jl:19-34)
//   periodic: mod.(cell .- 2, n_cells) .+ 2 in Julia's wrapping Int64 arithmetic (:619)
// Returns false when a coordinate of the final cell does not fit Int32 (the reference's
// coordinates_flattened raises InexactError then, spatial_hashing.jl:176-183).
template <int ND, bool PER>
__device__ __forceinline__ bool hash_cell_coords(const GridP &g, const float *p, long long *cc)
{
    bool ok = true;
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (d >= ND) { cc[d] = 0; continue; }
        const float f = floorf(__fdiv_rn(p[d], g.cs[d]));
        long long c;
        if (isnan(f) || f >= 9223372036854775808.0f) c = 0x7fffffffffffffffLL;
        else if (f <= -9223372036854775808.0f) c = (long long)0x8000000000000000ULL;
        else c = (long long)f;
        if (PER) {
            const long long a = (long long)((unsigned long long)c - 2ULL);
            long long m = a % (long long)g.nc[d];
            if (m < 0) m += g.nc[d];
            c = m + 2;
        }
        ok = ok && c >= -2147483648LL && c <= 2147483647LL;
        cc[d] = c;
    }
    return ok;
}

// spatial_hash (spatial_hashing.jl:159-174) with wrapping Int64 products and Julia's floored
// mod; 0-based key (the reference adds 1)
template <int ND>
__device__ __forceinline__ uint32_t spatial_hash_key(const long long *cc, int list_size)
{
    unsigned long long h = (unsigned long long)cc[0] * 73856093ULL;
    if (ND > 1) h ^= (unsigned long long)cc[1] * 19349663ULL;
    if (ND > 2) h ^= (unsigned long long)cc[2] * 83492791ULL;
    long long m = (long long)h % (long long)list_size;
    if (m < 0) m += list_size;
    return (uint32_t)m;
}

// mapreduce_neighbor_inner for a hashed cell list (src/nhs_grid.jl:519-575): one thread per
// query point; the 3^d neighbour cells in CartesianIndices order, each hashed to its table
// entry; entries that hold (or may hold) points of other cells -- check_cell_collision,
// :501-513 -- re-derive the cell of every accepted candidate (check_collision, :486-492).
// Candidates of a key are visited in ascending id order (the build leaves them canonical).
// SELF (x === y, all points): thread t takes the t-th record of the table itself, so the lanes
// of a warp are points of the same (or a neighbouring) cell: uniform trip counts, broadcast
// loads, payload of the query read from the cell-ordered copy.
template <int ND, bool PER, class CL, bool SELF>
__global__ void __launch_bounds__(128)
k_sweep_points_hash(GridP g, const uint32_t *__restrict__ key_start,
                    const float4 *__restrict__ sorted, const int4 *__restrict__ meta,
                    const float *__restrict__ x, int64_t n_loop,
                    const int32_t *__restrict__ points, int base, CL cl, int *__restrict__ err)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_loop) return;
    int i_id;
    float p[3] = {0.f, 0.f, 0.f};
    if (SELF) {
        const float4 q = __ldg(sorted + t);
        p[0] = q.x; p[1] = q.y; p[2] = q.z;
        i_id = __float_as_int(q.w);
    } else {
        i_id = points ? points[t] - base : (int)t;
#pragma unroll
        for (int d = 0; d < ND; d++) p[d] = __ldg(x + (int64_t)i_id * ND + d);
    }
    long long cc[3];
    hash_cell_coords<ND, PER>(g, p, cc);
    typename CL::State st;
    cl.init(st, true, SELF ? (int)t : -1, i_id);
    const PerP pp = make_perp(g);
    bool inexact = false;
    for (int oz = (ND > 2 ? -1 : 0); oz <= (ND > 2 ? 1 : 0); oz++)
        for (int oy = (ND > 1 ? -1 : 0); oy <= (ND > 1 ? 1 : 0); oy++)
            for (int ox = -1; ox <= 1; ox++) {
                long long nc[3] = {(long long)((unsigned long long)cc[0] + (unsigned long long)(long long)ox),
                                   (long long)((unsigned long long)cc[1] + (unsigned long long)(long long)oy),
                                   (long long)((unsigned long long)cc[2] + (unsigned long long)(long long)oz)};
                if (PER) {
#pragma unroll
                    for (int d = 0; d < ND; d++) {
                        long long m = (long long)((unsigned long long)nc[d] - 2ULL) % (long long)g.nc[d];
                        if (m < 0) m += g.nc[d];
                        nc[d] = m + 2;
                    }
                }
                // coordinates_flattened(neighbor_cell) needs Int32 coordinates (:176-183)
                bool fits = true;
#pragma unroll
                for (int d = 0; d < ND; d++) fits = fits && nc[d] >= -2147483648LL && nc[d] <= 2147483647LL;
                if (!fits) { inexact = true; continue; }
                const uint32_t key = spatial_hash_key<ND>(nc, g.total_cells);
                const int4 m = __ldg(meta + key);
                const bool cell_collision = (m.w & 1) || m.x != (int)nc[0] || m.y != (int)nc[1] ||
                                            m.z != (int)nc[2];
                const uint32_t b0 = key_start[key], b1 = key_start[key + 1];
                for (uint32_t gi = b0; gi < b1; gi++) {
                    const float4 pj = __ldg(sorted + gi);
                    float px = __fsub_rn(p[0], pj.x);
                    float py = ND > 1 ? __fsub_rn(p[1], pj.y) : 0.f;
                    float pz = ND > 2 ? __fsub_rn(p[2], pj.z) : 0.f;
                    float d2 = dist2<ND>(px, py, pz);
                    d2 = maybe_periodic_fix<ND, PER>(pp, d2, px, py, pz);
                    if (d2 <= g.r2) {
                        if (cell_collision) {
                            const float q[3] = {pj.x, pj.y, pj.z};
                            long long jc[3];
                            hash_cell_coords<ND, PER>(g, q, jc);
                            if (jc[0] != nc[0] || jc[1] != nc[1] || jc[2] != nc[2]) continue;
                        }
                        if (CL::kCountOnly) cl.count(st, 1);
                        else cl.template pair_global<ND>(st, px, py, pz, d2, __float_as_int(pj.w), gi);
                    }
                }
            }
    if (inexact) atomicOr(err, 2);
    cl.finish(st, SELF ? (int)t : -1, i_id);
}

#endif  // __CUDACC__

}  // namespace pnb
