// grid.cuh -- the grid handle (host side) shared by the translation units of libpnb200.
#pragma once

#include "common.cuh"

struct pnb_grid {
    pnb::GridP p;
    // host copies of the constructor results (reference field names in comments)
    float padded_min[3], padded_max[3];  // cell_list.min_corner / max_corner
    int64_t grid_size[3];                // size(cell_list.linear_indices)
    int64_t n_cells[3];                  // nhs.n_cells
    float cell_size[3];                  // nhs.cell_size
    float box_min[3], box_max[3];
    bool template_search;                // search_radius < eps(): build only empties the list

    int device;

    // cell list in CSR form, device
    uint32_t *cell_start_alloc;  // allocation behind cell_start (offset so cell_start + 1 is 16 B aligned)
    uint32_t *cell_start;   // [C+1] CSR offsets; during a build cell_start + 1 is the scatter cursor
    uint32_t *cell_count;   // [C]   histogram (scratch of the build; zero between builds)
    float4 *sorted;         // [cap] (x, y, z, id bits) in cell order -- the list every sweep reads
    float4 *sorted_alt;     // [cap] second buffer of ensure_canonical (allocated on first use)
    int32_t *cell_points;   // [cap] ids, 0-based, ascending inside a cell (valid when canonical)
    int64_t cap_points;
    bool canonical;         // records inside every cell are ordered by point id (ensure_canonical)

    // bucket layout written by the one-pass update!: every cell owns bucket_K
    // record slots, bcount[c] of them are used.  Valid instead of / next to the CSR arrays.
    uint32_t *bcount;        // [C]
    float4 *brec;            // [brec_slots] = C * bucket_K
    int64_t brec_slots;
    int bucket_K;            // 0 = not chosen yet (the next CSR build picks it from the fullest cell)
    bool bucket_tr;          // buckets numbered in transposed cell order (chosen from the input order)
    bool windowed;           // window of a larger grid (slab decomposition): plain bucket order
    bool bucket_valid;       // bcount / brec describe the current build
    uint32_t *bcount_alt;    // [C] second counter array: all zero between builds; every one-pass
                             // build clears the array of the NEXT build while it runs (no memset)
    bool bcount_alt_clean;
    // stream-ordered update! (pnb_grid_build_async_f32): the one-pass build was launched but its
    // error word (domain error, bucket overflow) has not been looked at yet
    bool async_pending;
    int async_chain;         // stream-ordered builds since the error word was last looked at
    cudaStream_t async_stream;
    bool csr_valid;          // cell_start / sorted describe the current build
    unsigned int *d_maxcount;   // [1] device scratch

    // Float64 search (f64.cuh): scalars in double and the cell-ordered Float64 records; the CSR
    // offsets / ids (cell_start, cell_points) are shared with the Float32 path
    bool f64;
    pnb::GridP64 p64;
    pnb::Rec64 *sorted64;      // [cap_points], canonical order (ids ascending inside a cell)
    pnb::Rec64 *sorted64_tmp;  // scatter target before the per-cell sort
    double padded_min64[3], padded_max64[3], cell_size64[3];

    // SpatialHashingCellList (hashgrid.cu): per key (c1, c2, c3, flags) = the cell stored by the
    // first insertion (cell_list.coords) and bit 0 of flags = cell_list.collisions
    bool hashed;
    int4 *hmeta;               // [list_size]

    // cell-ordered copy of the query points of a two-set sweep (x != y), built per sweep
    uint32_t *xq_start_alloc, *xq_start;   // [C+1]
    float4 *xq_sorted;                     // [xq_cap]
    int64_t xq_cap;

    // scan workspace
    unsigned long long *scan_status;
    int64_t scan_tiles_cap;
    unsigned int *scan_ticket;  // [1]
    int scan_epoch;             // status words of older launches are ignored (no memset per scan)

    // error word in mapped pinned host memory (d_err = its device address):
    // bit0 domain, bit1 bounds, bit2 list full
    int *d_err;
    int *h_err;

    // state of the last build
    int64_t n_built;        // number of points in the list
    const void *y_built;    // pointer identity of the coordinates
    int64_t n_y_built;      // columns of y at build time
    bool built;
    bool full_build;        // eachindex_y == all
    bool y_refreshed;       // the records' coordinates were re-read from another array than the one
                            // of the build (check_built_y): points may sit in other cells than
                            // the ones they are listed in, so x === y sweeps bin x like a second set

    // capacity hint for one-pass neighbour-list builds (nlist.cu): longest list of the last build
    int nl_cap_hint;

    // per-sweep scratch (payload gathered into cell order), grown on demand
    void *scratch;
    int64_t scratch_bytes;

    // tiles of k_sweep_tiles that did not fit its staging buffer (handled by k_sweep_overflow)
    int *ovf_tiles;
    int64_t ovf_cap;
    int *ovf_count;          // [2]: overflow tiles, surplus points (k_sweep_left)
    int *left_ids;           // surplus points of cells with a few more than 32 points
    int64_t left_cap;

    // k_sweep_flat (sweep_flat.cuh): tile table of one sweep and its control words
    void *flat_tiles;        // FlatTile[flat_tiles_cap]
    int *flat_ovf;           // [flat_tiles_cap] tiles handed to k_sweep_flat_overflow
    uint32_t *flat_tabs;     // [flat_tiles_cap x 224] staging tables of the tiles (k_flat_tables)
    int64_t flat_tiles_cap;
    uint32_t *flat_seg;      // [2 * (flat_seg_cap + 1)] tiles per row segment, exclusive prefix
    int64_t flat_seg_cap;
    uint32_t *flat_ctl;      // [4] number of tiles, tile counter, overflow tiles
};

namespace pnb {
// device buffers every grid handle owns (offsets, histogram, error word, scan ticket) for C cells
pnb_status grid_alloc_common(pnb_grid *g, int64_t C);
pnb_status ensure_point_capacity(pnb_grid *g, int64_t n);
// initialize!/update! of a hashed grid (hashgrid.cu)
pnb_status hash_build(pnb_grid *g, const float *y, int64_t n, const int32_t *idx, int64_t n_idx,
                      int base, cudaStream_t s);
pnb_status ensure_scratch(pnb_grid *g, int64_t bytes);
// neighbor_coords of a sweep / list build: the array of the last initialize!/update! (snapshot
// current), or another array of the same length (the records are re-read from it: old cell
// list + live coordinates, the reference's semantics, src/nhs_grid.jl:543-548)
pnb_status check_built_y(pnb_grid *g, const void *y, int64_t n, cudaStream_t s);
pnb_status check_err_word(pnb_grid *g, cudaStream_t s);  // sync + translate the error word
// PNB_RETRY_INTERNAL from check_err_word: a stream-ordered update! had overflowed a bucket; the
// cell list has been rebuilt (blocking), the sweep that ran on it must be repeated
constexpr pnb_status PNB_RETRY_INTERNAL = (pnb_status)100;
// settle a pending stream-ordered update! before anything but the x === y tile sweeps touches
// the cell list (blocks; rebuilds after a bucket overflow)
pnb_status resolve_pending(pnb_grid *g);
pnb_status build_query_list(pnb_grid *g, const float *x, int64_t nx, double *points_per_cell,
                            cudaStream_t s);  // two-set sweeps
pnb_status ensure_csr(pnb_grid *g, cudaStream_t s);       // CSR arrays from the bucket layout
pnb_status ensure_canonical(pnb_grid *g, cudaStream_t s);
// the cell list for the tile kernels: buckets if that is what the last build wrote, else CSR
static inline pnb::CellsView cells_view(const pnb_grid *g)
{
    if (g->bucket_valid) {
        if (g->bucket_tr)
            return pnb::CellsView{g->bcount, g->brec, (uint32_t)g->bucket_K, (uint32_t)g->p.gs[0],
                                  (uint32_t)g->p.gs[1], (uint32_t)g->p.gs[2]};
        return pnb::CellsView{g->bcount, g->brec, (uint32_t)g->bucket_K, 0u, 0u, 0u};
    }
    return pnb::CellsView{g->cell_start, g->sorted, 0u, 0u, 0u, 0u};
} // ids ascending inside every cell + cell_points
pnb_status exclusive_scan_u32(pnb_grid *g, const uint32_t *in, uint32_t *out, int64_t n,
                              cudaStream_t s);
pnb_status exclusive_scan_u32_to_i64(pnb_grid *g, const uint32_t *in, int64_t *out, int64_t n,
                                     cudaStream_t s);
}  // namespace pnb
