/*
 * pnb200.h -- C ABI of libpnb200.so: a B200 (sm_100a) native fixed-radius neighbourhood search
 * that is a drop-in for the hot path of trixi-framework/PointNeighbors.jl v0.6.7.
 *
 * The reference has no FFI (it is pure Julia, dispatching on the coordinate array type,
 * src/util.jl:81-83, src/gpu.jl:10-35).  Each entry point below names the reference interface
 * it replaces (file:line relative to the reference checkout); INTEGRATION.md shows the Julia
 * `ccall` glue and the Python `ctypes` binding that sit on top.
 *
 * Conventions
 *   - plain C: opaque handles, raw device pointers, sizes; no torch / CUDA C++ types.
 *   - coordinates: NDIMS x N column-major (Julia) == N x NDIMS row-major (C): xyzxyz...
 *     (src/neighborhood_search.jl:25-26), NDIMS in {1,2,3}, element type float32.
 *   - every function returns a pnb_status; pnb_last_error() returns the message, which is the
 *     reference's own error text where the reference raises one.
 *   - `index_base` (0 or 1) selects the numbering of point ids in everything that is exported
 *     or imported (Julia passes 1, C/Python pass 0).  Internally ids are 0-based int32.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are
 *     synchronous with respect to the host unless the name ends in _async: the reference
 *     blocks in KernelAbstractions.synchronize after every launch (src/util.jl:166-170).
 *   - device pointers must belong to the CUDA device that was current when the handle was
 *     created.  There is NO CPU fallback: without a CUDA device every compute call fails with
 *     PNB_ERR_CUDA.
 */
#ifndef PNB200_H
#define PNB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB200_VERSION 100 /* 0.1.0 */

typedef enum pnb_status {
    PNB_OK = 0,
    PNB_ERR_DOMAIN = 1,    /* "particle coordinates are NaN or outside the domain bounds of the cell list" (src/cell_lists/full_grid.jl:211) */
    PNB_ERR_ARG = 2,       /* ArgumentError of a constructor (src/nhs_grid.jl:81-124, src/cell_lists/full_grid.jl:52-59) */
    PNB_ERR_LIST_FULL = 3, /* "cell list is full. Use a larger `max_points_per_cell`." (src/vector_of_vectors.jl:119) */
    PNB_ERR_BOUNDS = 4,    /* BoundsError of the safe sweep: a query point's 3^d stencil leaves the grid (src/nhs_grid.jl:530-532) */
    PNB_ERR_CUDA = 5,      /* CUDA runtime failure / no device */
    PNB_ERR_STATE = 6      /* call order violated (e.g. sweep before initialize!) */
} pnb_status;

typedef struct pnb_grid pnb_grid;   /* GridNeighborhoodSearch{NDIMS} + FullGridCellList (+ PeriodicBox) */
typedef struct pnb_nlist pnb_nlist; /* PrecomputedNeighborhoodSearch neighbour lists */

int pnb_version(void);
/* thread-local; valid until the next failing call on this thread */
const char *pnb_last_error(void);
/* number of CUDA devices visible; 0 when there is none (then nothing else works) */
int pnb_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Host-side grid arithmetic (pure function, no device needed)
 *   replaces: FullGridCellList(; min_corner, max_corner, search_radius)   src/cell_lists/full_grid.jl:48-82
 *             GridNeighborhoodSearch{NDIMS}(; search_radius, periodic_box, cell_list)  src/nhs_grid.jl:77-129
 *             PeriodicBox(; min_corner, max_corner)                      src/neighborhood_search.jl:129-143
 * box_min/box_max == NULL: no periodic box.  Outputs (any may be NULL):
 *   padded_min/padded_max[ndims] : cell_list.min_corner / max_corner after the 1.001*r padding
 *   grid_size[ndims]             : n_cells_per_dimension of the allocated (padded) grid
 *   n_cells[ndims]               : nhs.n_cells (-1 when not periodic)
 *   cell_size[ndims]             : nhs.cell_size
 * Errors: PNB_ERR_ARG with the reference's ArgumentError text.
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_grid_params_f32(int ndims, float search_radius, const float *min_corner,
                               const float *max_corner, const float *box_min, const float *box_max,
                               float *padded_min, float *padded_max, int64_t *grid_size,
                               int64_t *n_cells, float *cell_size);

/* ---------------------------------------------------------------------------------------------
 * Grid handle
 *   replaces: the GridNeighborhoodSearch / FullGridCellList objects after Adapt.adapt(backend, nhs)
 *             (src/gpu.jl:10-12) and copy_neighborhood_search (src/nhs_grid.jl:640-649).
 * min_corner/max_corner are the USER corners (unpadded).  search_radius < eps() gives the legal
 * "template / unused" search whose build only empties the cell list (src/nhs_grid.jl:263-267).
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_grid_create_f32(int ndims, float search_radius, const float *min_corner,
                               const float *max_corner, const float *box_min, const float *box_max,
                               pnb_grid **out);
/* The same from the corners STORED in a FullGridCellList (cell_list.min_corner / max_corner are
 * already padded, src/cell_lists/full_grid.jl:66-67): no second padding, the grid size is
 * ceil((max - min) / search_radius) of the stored corners (:74).  This is what the Julia glue
 * calls from Adapt.adapt_structure, so the device grid is bit-identical to the host cell list. */
pnb_status pnb_grid_create_padded_f32(int ndims, float search_radius, const float *padded_min,
                                      const float *padded_max, const float *box_min,
                                      const float *box_max, pnb_grid **out);
/* A window of the same global grid (multi-GPU slab decomposition): cells
 * win_lo[d]..win_hi[d] (global 1-based cell coordinates, inclusive; the outermost layer on each
 * side acts as the empty padding layer) with the GLOBAL cell arithmetic, so cell assignments and
 * neighbour sets are bit-identical to the undecomposed search.  NULL windows = the full grid.
 * There is no reference counterpart (the reference is single-device). */
pnb_status pnb_grid_create_window_f32(int ndims, float search_radius, const float *min_corner,
                                      const float *max_corner, const float *box_min,
                                      const float *box_max, const int64_t *win_lo,
                                      const int64_t *win_hi, pnb_grid **out);
/* Slab exchange bookkeeping in ONE pass over the owned points (multi-GPU):
 * cell layer cz = floor((z - padded_min_z) / cell_size) + 1 of the last coordinate (the reference's
 * arithmetic, src/cell_lists/full_grid.jl:93) and three compacted index lists (0-based, unordered):
 *   up_idx    points with cz >= z_hi (sent to rank + 1: migrants and its ghost layer), if has_up
 *   down_idx  points with cz <= z_lo (sent to rank - 1), if has_down
 *   leave_idx points with cz outside [z_lo, z_hi] (they leave this rank)
 * Every list holds at most `cap` entries; counts[0..2] (host) are the true lengths, so a count
 * above `cap` tells the caller to retry with larger lists.  Synchronises the stream. */
pnb_status pnb_slab_classify_f32(const float *coords, int64_t n, int ndims, float padded_min_z,
                                 float cell_size_z, int64_t z_lo, int64_t z_hi, int has_up,
                                 int has_down, int32_t *up_idx, int32_t *down_idx,
                                 int32_t *leave_idx, int64_t cap, int32_t *counts_dev,
                                 int64_t *counts, void *stream);

/* Structure-of-arrays particle state of one rank: n_arrays float arrays with a common leading
 * (capacity) dimension, array a has width[a] floats per point; array 0 = coordinates. */
typedef struct pnb_slab_arrays {
    float *ptr[8];
    int32_t width[8];
    int32_t n_arrays;
} pnb_slab_arrays;

/* Step 1 of the per-step exchange: classify the n owned points (as pnb_slab_classify_f32) and
 * gather the rows of the up / down lists into the contiguous send buffers (row = all arrays'
 * columns of one point, row_width = sum of widths).  counts = {n_up, n_down, n_leave}; if one
 * exceeds cap nothing is packed (retry with larger buffers).  leave_idx is kept for step 2. */
pnb_status pnb_slab_pack_f32(const pnb_slab_arrays *arrays, int64_t n, int ndims,
                             float padded_min_z, float cell_size_z, int64_t z_lo, int64_t z_hi,
                             int has_up, int has_down, int32_t *up_idx, int32_t *down_idx,
                             int32_t *leave_idx, int64_t cap, float *send_up, float *send_down,
                             int32_t *counts_dev, int64_t *counts, void *stream);

/* Step 2, after the rows have been exchanged: emigrants' holes are filled from the tail
 * (n_stay = n - n_leave owned points remain), received rows that belong to this slab are appended
 * as owned points, then the ghosts: received rows one layer outside the slab and the rows this
 * rank sent that are now one layer outside it.  scratch: >= 3 * n_leave + n_recv_up + n_recv_down
 * + n_up + n_down + 16 int32.  The arrays must have room for n_stay + all four row counts.
 * out = {n_own, n_local}.  Synchronises the stream. */
pnb_status pnb_slab_unpack_f32(const pnb_slab_arrays *arrays, int64_t n, int ndims,
                               float padded_min_z, float cell_size_z, int64_t z_lo, int64_t z_hi,
                               int has_up, int has_down, const int32_t *leave_idx, int64_t n_leave,
                               const float *recv_up, int64_t n_recv_up, const float *recv_down,
                               int64_t n_recv_down, const float *send_up, int64_t n_up,
                               const float *send_down, int64_t n_down, int32_t *scratch,
                               int64_t *out, void *stream);
/* Overlapped step: the received rows are appended behind the n_own owned points WITHOUT filling the
 * holes of the leavers (the ids of the owned points stay valid for the sweep that is already
 * running); flags[n_own + r] = 1 for a row that now belongs to this slab (migrant), 0 for a ghost.
 * counters_dev (>= 4 int32, zeroed by the call): [0] migrants, [1] != 0 if a migrant landed `depth`
 * or more layers inside the slab (the interior sweep missed it: repeat the step without overlap). */
pnb_status pnb_slab_append_f32(const pnb_slab_arrays *arrays, int64_t n_own, int ndims,
                               float padded_min_z, float cell_size_z, int64_t z_lo, int64_t z_hi,
                               int64_t depth, const float *recv_up, int64_t n_recv_up,
                               const float *recv_down, int64_t n_recv_down, uint8_t *flags,
                               int32_t *counters_dev, void *stream);
/* The same for received rows that are row_stride floats apart (0 = dense; the receive buffers of a
 * pnb_slab_link pad rows to pnb_slab_link_row_stride() floats). */
pnb_status pnb_slab_append_strided_f32(const pnb_slab_arrays *arrays, int64_t n_own, int ndims,
                                       float padded_min_z, float cell_size_z, int64_t z_lo,
                                       int64_t z_hi, int64_t depth, const float *recv_up,
                                       int64_t n_recv_up, const float *recv_down, int64_t n_recv_down,
                                       int64_t row_stride, uint8_t *flags, int32_t *counters_dev,
                                       void *stream);
/* End of the overlapped step: owned points [0, n_own) minus the n_leave leavers (leave_idx of
 * pnb_slab_pack_f32) plus the n_mig migrants among the n_app appended rows become the owned points
 * [0, n_own - n_leave + n_mig): holes are filled with owned rows from behind.  Every array is
 * permuted alike.  scratch: >= 2 * (n_leave + n_app + 8) int32. */
pnb_status pnb_slab_compact_f32(const pnb_slab_arrays *arrays, int64_t n_own, int64_t n_app,
                                const int32_t *leave_idx, int64_t n_leave, int64_t n_mig,
                                uint8_t *flags, int32_t *scratch, int32_t *counters_dev, void *stream);

/* The per-step row exchange of neighbouring slabs over NVLink PEER MEMORY instead of NCCL
 * (csrc/link.cu; no reference counterpart).  Every rank creates a link (its receive area), passes
 * the 64-byte handle of pnb_slab_link_export to rank - 1 and rank + 1 (any transport), and connects
 * the handles it got (NULL: no neighbour on that side).  Per step `seq` = 1, 2, 3, ... (the same
 * on all ranks):
 *   pnb_slab_link_send: ONE kernel classifies the n owned points (as pnb_slab_classify_f32), stores
 *     the rows of the leaving / boundary points straight into the neighbours' receive buffers and
 *     the leavers' indices into leave_idx (>= 2 * cap_rows int32); a second kernel publishes
 *     (count, seq).  Nothing is synchronised.
 *   pnb_slab_link_recv (same stream): waits for step `seq` of both neighbours, SYNCHRONISES the
 *     stream; rows_down / rows_up point into this rank's receive area (valid until step seq + 2 is
 *     received; rows are pnb_slab_link_row_stride() floats apart, the width rounded up to 4), counts = {n_from_down, n_from_up, n_sent_down, n_sent_up, n_leave}.
 *     PNB_ERR_LIST_FULL if a message exceeded cap_rows, PNB_ERR_STATE if a neighbour did not
 *     answer within the time limit (pnb_slab_link_set_timeout, default 120 s). */
typedef struct pnb_slab_link pnb_slab_link;
pnb_status pnb_slab_link_create(int64_t cap_rows, int width, pnb_slab_link **out);
pnb_status pnb_slab_link_export(pnb_slab_link *l, void *handle_out);
pnb_status pnb_slab_link_connect(pnb_slab_link *l, const void *handle_down, const void *handle_up);
/* two links of the same process (tests, profiles): pointers instead of handles */
pnb_status pnb_slab_link_connect_local(pnb_slab_link *l, pnb_slab_link *down, pnb_slab_link *up);
pnb_status pnb_slab_link_send(pnb_slab_link *l, const pnb_slab_arrays *arrays, int64_t n, int ndims,
                              float padded_min_z, float cell_size_z, int64_t z_lo, int64_t z_hi,
                              int32_t *leave_idx, uint64_t seq, void *stream);
pnb_status pnb_slab_link_recv(pnb_slab_link *l, uint64_t seq, const float **rows_down,
                              const float **rows_up, int64_t *counts, void *stream);
pnb_status pnb_slab_link_set_timeout(pnb_slab_link *l, double seconds);   /* default 120 */
int pnb_slab_link_row_stride(const pnb_slab_link *l);
void pnb_slab_link_destroy(pnb_slab_link *l);

/* ---------------------------------------------------------------------------------------------
 * GridNeighborhoodSearch{NDIMS}(; search_radius, periodic_box,
 *                               cell_list = SpatialHashingCellList{NDIMS}(; list_size))
 *   replaces: src/cell_lists/spatial_hashing.jl:24-61 (the table), src/nhs_grid.jl:100-126
 *             (cell_size / n_cells of the search), src/gpu.jl:37-44 (adapt)
 * The handle is a pnb_grid: initialize!/update! (pnb_grid_build_f32), every sweep, the neighbour
 * lists and the CSR / DVoV exports work on it unchanged, with the 0-based hash key
 * (spatial_hash(cell, list_size) - 1, :159-174) in place of the linear cell index and
 * pnb_grid_total_cells() == list_size.  Cells are floor(coords / cell_size) (no corners: the
 * domain is unbounded, src/nhs_grid.jl:636-638); the sweep applies the reference's collision
 * checks (src/nhs_grid.jl:479-513), so neighbour sets equal those of every other search.
 * A cell coordinate outside Int32 is the reference's InexactError (:176-183) -> PNB_ERR_DOMAIN.
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_grid_create_hashed_f32(int ndims, float search_radius, int64_t list_size,
                                      const float *box_min, const float *box_max, pnb_grid **out);
/* cell_list.coords (list_size x UInt128 as 4 little-endian uint32 words, coordinates_flattened
 * :176-183; 0 = unused entry) and cell_list.collisions (list_size x Bool) after the last build,
 * as the reference's serial insertion in ascending id order leaves them (push_cell!, :79-97).
 * Either pointer may be NULL. */
pnb_status pnb_grid_export_hash_table(const pnb_grid *g, uint32_t *coords, uint8_t *collisions,
                                      void *stream);
/* spatial_hash (src/cell_lists/spatial_hashing.jl:159-174), host, 0-based (the reference's key - 1);
 * -1 on invalid arguments */
int64_t pnb_spatial_hash(int ndims, const int64_t *cell, int64_t list_size);

void pnb_grid_destroy(pnb_grid *g);
int64_t pnb_grid_total_cells(const pnb_grid *g);
int64_t pnb_grid_n_points(const pnb_grid *g); /* points in the cell list after the last build */
/* layout the last build wrote: 0 = CSR (two-pass counting sort), 1 = buckets (one-pass update!),
 * -1 = not built.  Inspection only (the reference's cell storage is one layout,
 * src/vector_of_vectors.jl:3-31; both layouts here export to it). */
int pnb_grid_layout(const pnb_grid *g);

/* initialize!(nhs, x, y; eachindex_y) / update!(nhs, x, y; points_moving = (_, true), eachindex_y)
 *   src/nhs_grid.jl:220-225, 255-292, 470-477; src/cell_lists/full_grid.jl:96-139
 * Device counting sort: cell index + histogram, decoupled-lookback scan, scatter, per-cell id
 * sort (deterministic: ids ascending inside a cell) fused with the cell-order copy of the
 * coordinates.  y: device, n x ndims.  eachindex_y: device int32 ids (index_base) or NULL = all.
 * Returns PNB_ERR_DOMAIN if any listed point is NaN / outside the padded grid. */
pnb_status pnb_grid_build_f32(pnb_grid *g, const float *y, int64_t n, const int32_t *eachindex_y,
                              int64_t n_idx, int index_base, void *stream);

/* Stream-ordered update! (no reference counterpart: the reference blocks after every launch,
 * src/util.jl:166-170).  From the second build on (full rebuild, FullGridCellList) the one-pass
 * build is only ENQUEUED on `stream`: nothing is synchronised and the error word is not read.
 * The next blocking call on the handle settles it: a sweep reports "particle coordinates are NaN
 * or outside the domain bounds..." then, and a bucket that overflowed makes the
 * library rebuild the cell list (two-pass, blocking) and repeat that sweep, invisibly to the
 * caller.  First builds, hashed and template searches fall back to pnb_grid_build_f32.
 * pnb_grid_check synchronises `stream` and returns what a blocking update! would have returned. */
pnb_status pnb_grid_build_async_f32(pnb_grid *g, const float *y, int64_t n, void *stream);
pnb_status pnb_grid_check(pnb_grid *g, void *stream);
/* The same without any synchronisation, for a caller that has already waited for an event it
 * recorded behind the update! (and pnb_grid_append_f32): kernels launched after that event keep
 * running; error bits they set are returned by the next check. */
pnb_status pnb_grid_check_settled(pnb_grid *g);

/* cell_coords + cell_index of arbitrary points (src/nhs_grid.jl:622-628, full_grid.jl:84-94,157-161):
 * out[i] = 0-based linear cell index, or -1 when the cell is outside 2:(size-1). */
pnb_status pnb_point_cells_f32(const pnb_grid *g, const float *x, int64_t n, int32_t *out_linear,
                               void *stream);

/* The cell list in CSR form: cell_start[C+1] (int32 offsets), cell_points[n_points] (ids + index_base).
 * This replaces cell_list.cells (DynamicVectorOfVectors, src/vector_of_vectors.jl:3-31). */
pnb_status pnb_grid_export_csr(const pnb_grid *g, int32_t *cell_start, int32_t *cell_points,
                               int index_base, void *stream);
/* The cell list in the reference's own layout: backend[max_points_per_cell x C] column-major int32,
 * lengths[C] int32.  PNB_ERR_LIST_FULL if a cell holds more than max_points_per_cell points. */
pnb_status pnb_grid_export_dvov(const pnb_grid *g, int32_t *backend, int32_t *lengths,
                                int32_t max_points_per_cell, int index_base, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused sweeps = foreach_point_neighbor(f, x, y, nhs; points)  (src/neighborhood_search.jl:183-201,
 * src/nhs_grid.jl:519-575) with f one of the benchmarked closures.
 * x: device nx x ndims (points looped over), y: device n x ndims (the array the grid was built
 * from).  points: device int32 ids (index_base) or NULL = all of x.
 * When x == y (same pointer and size) and points == NULL the cell-ordered fast path runs; it
 * uses the coordinates snapshotted by the last pnb_grid_build (call update! after moving y, as
 * the reference requires, src/neighborhood_search.jl:161-164).
 * Returns PNB_ERR_BOUNDS if a query point's stencil leaves the grid.
 * ------------------------------------------------------------------------------------------- */

/* Arithmetic of the per-pair TERMS of the fused n-body / WCSPH closures (never of the neighbour
 * test, which is always the reference's exact operation sequence):
 *   0 (default) fast: MUFU rsqrt/rcp + FMA, per-term relative error ~1e-6, sums within the 1e-5 bar;
 *   1 exact: the Julia operation sequence with IEEE sqrt/div and no FMA -- sums bit-identical to
 *     the CPU oracle (candidates are visited in the reference's order). */
void pnb_set_exact_arithmetic(int on);
int pnb_get_exact_arithmetic(void);

/* Measurement overrides of the tile sweep (3-D, non-periodic): warps per cell (2 or 4, 0 = the
 * closure's default) and the fp16 pre-filter of the distance test (0 = off: exact Float32 test,
 * -1 = default on).  Results are identical for every setting; only the speed changes. */
void pnb_set_tuning(int warps_per_cell, int half_prefilter);
/* Sweeps over two point sets (x != y, all points of x): 1 (default) = x is binned into the grid's
 * cells and swept by the tile kernel when its occupied cells hold >= 12 query points on average,
 * 2 = always, 0 = never (one thread per query point).  Same results. */
void pnb_set_twoset_tiles(int on);
/* Cells with 33 .. 40 points in the tile sweep: 1 (default) = for the n-body / WCSPH closures on
 * clouds of >= 200 000 points the first 32 points run in the tile kernel and the surplus points
 * in a per-point kernel (instead of a second, nearly empty batch of the whole cell); 0 = never,
 * 2 = always when the closure allows it (tests).  Same results. */
void pnb_set_sweep_left(int mode);
/* Tile sweep kernel: 1 (default) = k_sweep_flat (tiles are ranges of 96 points, full warps,
 * persistent CTAs) for the closures with a drain (n-body, WCSPH, list fills) and
 * k_sweep_tiles of round 1 (tiles are 4 cells, one warp group per cell) for the count-only
 * closures, which it serves faster; 2 = always k_sweep_flat, 0 = always k_sweep_tiles (A/B
 * measurements, tests).  Same results. */
void pnb_set_sweep_kernel(int flat);
/* k_sweep_flat is a persistent kernel (one CTA per slot of every SM): kernels of OTHER streams --
 * the NCCL send / recv of the overlapped multi-GPU step, which need a mostly empty SM -- would not
 * run before it ends.  pnb_set_sweep_reserve(n) makes the CTAs that land on the last n SMs leave at
 * once (default 0), so that those SMs stay free. */
void pnb_set_sweep_reserve(int sms);
/* Measurement variants of the counting-sort kernels (bits: 1 histogram reads 4 consecutive points
 * per thread straight from global memory, 2 staged scatter with the lane-strided mapping,
 * 4 staged histogram with it, 8 scatter without staging, 16 histogram without staging and with
 * lane runs; default 25).  Results are identical for every setting.  (The variants of the
 * one-pass kernel that leave out its atomics or stores -- bits 8..10, invalid layouts, used by
 * tools/bucket_diag.py to attribute its time -- exist only in a library compiled with -DPNB_DIAG,
 * which _build.py does when PNB_DIAG=1 is set in the environment; the shipped binary ignores
 * those bits.) */
void pnb_set_build_tuning(int variant);
/* 1 (default): builds after the first one use the one-pass bucket layout (every cell owns K record
 * slots, K from the fullest cell of the last CSR build; a cell that overflows falls back to the
 * two-pass CSR build); 0: always the two-pass CSR build.  Same results. */
void pnb_set_build_layout(int buckets);
/* Numbering of the buckets of the one-pass layout: 0 = linear cell order, 1 = transposed (last
 * dimension fastest), -1 (default) = chosen from the order of the input at the first build
 * (transposed when that is how the points step through the cells, e.g. the reference's benchmark
 * clouds, test/point_cloud.jl:49: consecutive points then write neighbouring buckets).  Same
 * results; measured trade-off in DESIGN.md 5.1. */
void pnb_set_bucket_order(int order);

/* benchmarks/count_neighbors.jl:16-28: out[i] = number of neighbours (int64, zeroed first) */
pnb_status pnb_count_neighbors_f32(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                   int64_t n, const int32_t *points, int64_t n_points,
                                   int index_base, int64_t *out, void *stream);

/* benchmarks/n_body.jl:29-51: dv (nx x ndims) zeroed, then for every pair with
 * distance >= sqrt(eps): dv[:, i] += -G * mass[j] * pos_diff / distance^3 */
pnb_status pnb_nbody_f32(pnb_grid *g, const float *x, int64_t nx, const float *y, int64_t n,
                         const int32_t *points, int64_t n_points, int index_base,
                         const float *mass, float G, float *dv, void *stream);

/* WCSPH continuity + momentum pair interaction, TrixiParticles.interact! as configured by
 * benchmarks/smoothed_particle_hydrodynamics.jl:45-102 (WendlandC2, ContinuityDensity,
 * ArtificialViscosityMonaghan, DensityDiffusionMolteniColagrossi, StateEquationCole exponent 1).
 * v: (ndims+1) per point = velocity, density; dv: same shape, zeroed then accumulated. */
typedef struct pnb_wcsph_params {
    float smoothing_length;  /* h = search_radius / 2 */
    float sound_speed;       /* c */
    float alpha, beta, epsilon; /* Monaghan viscosity, epsilon = 0.01 */
    float delta;             /* Molteni-Colagrossi */
    float kernel_norm;       /* sigma_d / h^d (WendlandC2: 21/(16 pi h^3) in 3D, 7/(4 pi h^2) in 2D) */
} pnb_wcsph_params;
pnb_status pnb_wcsph_interact_f32(pnb_grid *g, const float *x, int64_t nx, const float *y,
                                  int64_t n, const int32_t *points, int64_t n_points,
                                  int index_base, const float *v_x, const float *v_y,
                                  const float *mass_x, const float *mass_y,
                                  const float *pressure_x, const float *pressure_y,
                                  const pnb_wcsph_params *params, float *dv, void *stream);

/* Stream-ordered form of pnb_wcsph_interact_f32 for x === y, all points (y must be the array of
 * the last update!): gather + sweep are enqueued on `stream`, nothing is synchronised;
 * pnb_grid_check settles it together with a preceding pnb_grid_build_async_f32. */
pnb_status pnb_wcsph_interact_async_f32(pnb_grid *g, const float *y, int64_t n, const float *v,
                                        const float *mass, const float *pressure,
                                        const pnb_wcsph_params *params, float *dv, void *stream);

/* The pieces of the OVERLAPPED multi-GPU step (pnb200/slabs.py; no reference
 * counterpart, the reference is single-device).  Per step and rank: pack the boundary rows, start
 * the NCCL exchange on a side stream, pnb_grid_build_async_f32 of the owned points, sweep the
 * interior cell layers (pnb_wcsph_interact_layers_async_f32) while the rows travel,
 * pnb_slab_append_f32 + pnb_grid_append_f32 of what arrived, sweep the boundary layers,
 * pnb_grid_check, pnb_slab_compact_f32.
 *   pnb_grid_append_f32: the points y[first .. first + n_more) join the bucket cell list built from
 *     y[0 .. first) (ids first + k); y must be the array of that build.
 *   pnb_wcsph_interact_layers_async_f32: sweep the cell layers [cz_a, cz_b] and [cz_c, cz_d] (local
 *     1-based cell coordinates of the last dimension, an empty range has first > last) with one
 *     launch, after gathering the payload of these layers and their neighbours (mode 0); mode 1
 *     ONLY gathers the payload of exactly these layers, mode 2 ONLY sweeps.  dv is written for the
 *     points of the swept layers only.  Nothing is synchronised.
 *   pnb_slab_pack_rows_f32: rows of `arrays` listed in `list` -> contiguous send buffer. */
pnb_status pnb_grid_append_f32(pnb_grid *g, const float *y, int64_t first, int64_t n_more, void *stream);
pnb_status pnb_slab_pack_rows_f32(const pnb_slab_arrays *arrays, const int32_t *list, int64_t count,
                                  float *dst, void *stream);
pnb_status pnb_wcsph_interact_layers_async_f32(pnb_grid *g, const float *y, int64_t n, const float *v,
                                               const float *mass, const float *pressure,
                                               const pnb_wcsph_params *params, float *dv, int cz_a,
                                               int cz_b, int cz_c, int cz_d, int mode, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The WCSPH step from HOST buffers, pipelined (the end-to-end call of a host-side caller).
 *   replaces, per step: copyto!(device arrays, host arrays) (Adapt, benchmarks/run_benchmarks.jl:97-99),
 *   update!(nhs, y, y; points_moving = (true, true)) (src/nhs_grid.jl:283-292), interact!
 *   (benchmarks/smoothed_particle_hydrodynamics.jl:99-101) and the copy of dv back to the host.
 * Three streams and double-buffered device arrays inside the handle: the host->device copy of
 * step s + 1 and the device->host copy of step s - 1 overlap the kernels of step s.
 * submit() enqueues step s and returns once step s - 1's kernels are done and checked (a domain
 * error of step s - 1 is returned by this call); dv_host of step s is valid after
 * pnb_hoststep_wait(), or once the submit of step s + 2 has returned.
 * y_host: n x ndims, v_host: n x (ndims + 1) (velocity, density), pressure_host: n, mass_host: n
 * or NULL (= unchanged since the last submit that passed it), dv_host: n x (ndims + 1).
 * Pinned host memory (pnb_malloc_host) is needed for the copies to overlap.  The grid handle
 * must not be used by other calls while steps are in flight.
 * ------------------------------------------------------------------------------------------- */
typedef struct pnb_hoststep pnb_hoststep;
pnb_status pnb_hoststep_create(pnb_grid *g, int64_t n, pnb_hoststep **out);
pnb_status pnb_hoststep_wcsph_submit(pnb_hoststep *h, const float *y_host, const float *v_host,
                                     const float *mass_host, const float *pressure_host,
                                     const pnb_wcsph_params *params, float *dv_host);
/* StateEquationCole of the system: afterwards pressure_host may be NULL in a submit and the
 * pressure is computed on the device from the density row of v (compute_pressure!,
 * benchmarks/smoothed_particle_hydrodynamics.jl:64-69, 99), saving its host-to-device copy. */
pnb_status pnb_hoststep_set_state_equation(pnb_hoststep *h, float sound_speed, float reference_density,
                                           float exponent, float background_pressure);
/* host seconds spent inside the submits so far: {H2D enqueue, wait for the previous step, update!
 * enqueue, interact! enqueue, D2H enqueue}; steps = number of submits (diagnostics of the pipeline) */
pnb_status pnb_hoststep_host_times(const pnb_hoststep *h, double *out5, int64_t *steps);
pnb_status pnb_hoststep_wait(pnb_hoststep *h);
void pnb_hoststep_destroy(pnb_hoststep *h);

/* ---------------------------------------------------------------------------------------------
 * Neighbour lists = PrecomputedNeighborhoodSearch (src/nhs_precomputed.jl:67-277)
 * ------------------------------------------------------------------------------------------- */
/* initialize_neighbor_lists! (src/nhs_precomputed.jl:187-206): count pass, scan, fill pass and a
 * per-list sort (sort != 0 -> ascending ids, the deterministic result of sorteach!,
 * src/vector_of_vectors.jl:177-212).  The lists are a device CSR owned by the handle. */
pnb_status pnb_nlist_build_f32(pnb_grid *g, const float *x, int64_t nx, const float *y, int64_t n,
                               int sort, pnb_nlist **out, void *stream);
void pnb_nlist_destroy(pnb_nlist *l);
int64_t pnb_nlist_n_points(const pnb_nlist *l);
int64_t pnb_nlist_n_pairs(const pnb_nlist *l);
/* length of the longest list: the reference raises "cell list is full..." when it exceeds
 * max_neighbors (src/vector_of_vectors.jl:114-121) */
int64_t pnb_nlist_max_length(const pnb_nlist *list);
/* offsets[nx+1] int64, ids[n_pairs] int32 (+ index_base) */
pnb_status pnb_nlist_export_csr(const pnb_nlist *l, int64_t *offsets, int32_t *ids, int index_base,
                                void *stream);
/* reference layout (src/vector_of_vectors.jl:3-31): lengths[nx]; backend max_neighbors x nx
 * column-major, or its transpose nx x max_neighbors (transpose_backend = true, :18-23).
 * Unused slots are set to typemax(Int32) like the GPU sorteach! (:200-204).
 * PNB_ERR_LIST_FULL if a list is longer than max_neighbors. */
pnb_status pnb_nlist_export_dvov(const pnb_nlist *l, int32_t *backend, int32_t *lengths,
                                 int32_t max_neighbors, int transposed, int index_base,
                                 void *stream);
/* sweep without radius test (src/nhs_precomputed.jl:210-247): what the closure receives, in
 * list order: pos_diff[n_pairs x ndims], distance[n_pairs] (either may be NULL). */
pnb_status pnb_nlist_pairs_f32(const pnb_nlist *l, const pnb_grid *g, const float *x,
                               const float *y, float *pos_diff, float *distance, void *stream);
/* TLSPH deformation gradient over the lists (TrixiParticles.calc_deformation_grad!, called at
 * benchmarks/smoothed_particle_hydrodynamics.jl:142): neighbours and kernel gradient on the
 * initial coordinates X0, current coordinates xcur; L = kernel correction matrix (ndims x ndims
 * column-major per point); F same layout. */
pnb_status pnb_tlsph_deformation_grad_f32(const pnb_nlist *l, const pnb_grid *g, const float *X0,
                                          const float *xcur, const float *mass, const float *rho0,
                                          const float *L, float smoothing_length,
                                          float kernel_norm, float *F, void *stream);

/* The other TLSPH kernel, TrixiParticles.interact_structure_structure! (called at
 * benchmarks/smoothed_particle_hydrodynamics.jl:121), over the lists (no radius test,
 * src/nhs_precomputed.jl:221-244): dv (nx x ndims) is overwritten with
 *   sum_j m_j (PK1c_i / rho_i^2 + PK1c_j / rho_j^2) gradW(X0_i - X0_j)
 *         + PenaltyForceGanzenmueller(alpha) / m_i          (pairs closer than sqrt(eps) skipped)
 * pk1_corrected, F: ndims x ndims column-major per point.  Arithmetic definition: oracle
 * pno_tlsph_interact (TrixiParticles is not vendored: parity unpinned). */
typedef struct pnb_tlsph_params {
    float smoothing_length; /* h = search_radius / 2 */
    float kernel_norm;      /* sigma_d / h^d of the Wendland C2 kernel */
    float young_modulus;    /* E */
    float penalty_alpha;    /* PenaltyForceGanzenmueller(alpha) */
} pnb_tlsph_params;
pnb_status pnb_tlsph_interact_f32(const pnb_nlist *l, const pnb_grid *g, const float *X0,
                                  const float *xcur, const float *mass, const float *rho0,
                                  const float *pk1_corrected, const float *F,
                                  const pnb_tlsph_params *params, float *dv, void *stream);
/* TrixiParticles.compute_pk1_corrected! (benchmarks/smoothed_particle_hydrodynamics.jl:186),
 * pointwise: PK1 of the St. Venant-Kirchhoff material from the deformation gradient F, times the
 * kernel correction matrix L (all ndims x ndims column-major per point). */
pnb_status pnb_tlsph_pk1_corrected_f32(int ndims, int64_t n, const float *F, const float *L,
                                       float young_modulus, float poisson_ratio,
                                       float *pk1_corrected, void *stream);
/* TrixiParticles.compute_pressure! for ContinuityDensity + StateEquationCole
 * (benchmarks/smoothed_particle_hydrodynamics.jl:64-69, 99), pointwise:
 * pressure[i] = B ((rho_i / rho0)^exponent - 1) + background, B = rho0 c^2 / exponent,
 * rho_i = v[i, ndims] (the density row of the (ndims + 1) x N state). */
pnb_status pnb_wcsph_compute_pressure_f32(int ndims, int64_t n, const float *v, float sound_speed,
                                          float reference_density, float exponent,
                                          float background_pressure, float *pressure,
                                          void *stream);

/* ---------------------------------------------------------------------------------------------
 * Device memory helpers for hosts without a CUDA binding (the Julia glue's B200Array;
 * replaces Adapt.adapt(backend, array), src/gpu.jl, benchmarks/run_benchmarks.jl:97-99)
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_malloc(void **dev_ptr, int64_t bytes);
pnb_status pnb_free(void *dev_ptr);
pnb_status pnb_malloc_host(void **host_ptr, int64_t bytes); /* pinned */
pnb_status pnb_free_host(void *host_ptr);
pnb_status pnb_memcpy_h2d(void *dst_dev, const void *src_host, int64_t bytes, void *stream);
pnb_status pnb_memcpy_d2h(void *dst_host, const void *src_dev, int64_t bytes, void *stream);
pnb_status pnb_memset(void *dev_ptr, int value, int64_t bytes, void *stream);
pnb_status pnb_stream_synchronize(void *stream);

/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
int64_t pnb_launch_count(void);

/* Optional per-kernel timing (CUDA events recorded on the launching stream around every kernel
 * of a phase).  Off by default.  pnb_profile_get synchronizes the recorded events and returns
 * the accumulated device time and launch count of one phase since the last reset. */
void pnb_profile_enable(int on);
void pnb_profile_reset(void);
int pnb_profile_phases(void);
const char *pnb_profile_name(int phase);
pnb_status pnb_profile_get(int phase, double *total_ms, int64_t *launches);

/* ---------------------------------------------------------------------------------------------
 * Float64 searches (coordinates, corners and search_radius all Float64): the searching part of
 * the API -- constructors, initialize!/update!, neighbour counts, neighbour lists, pair geometry --
 * with the reference's arithmetic evaluated in Float64 (the reference is generic in the element
 * type, src/nhs_grid.jl:60-65).  The list handle and all exports are shared with Float32; the
 * fused n-body / WCSPH closures exist in Float64 too (one thread per point, bit-exact); the TLSPH
 * closures are Float32 only.
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_grid_params_f64(int ndims, double search_radius, const double *min_corner,
                               const double *max_corner, const double *box_min,
                               const double *box_max, double *padded_min, double *padded_max,
                               int64_t *grid_size, int64_t *n_cells, double *cell_size);
pnb_status pnb_grid_create_f64(int ndims, double search_radius, const double *min_corner,
                               const double *max_corner, const double *box_min,
                               const double *box_max, pnb_grid **out);
pnb_status pnb_grid_create_padded_f64(int ndims, double search_radius, const double *padded_min,
                                      const double *padded_max, const double *box_min,
                                      const double *box_max, pnb_grid **out);
pnb_status pnb_grid_build_f64(pnb_grid *g, const double *y, int64_t n, const int32_t *eachindex_y,
                              int64_t n_idx, int index_base, void *stream);
pnb_status pnb_point_cells_f64(const pnb_grid *g, const double *x, int64_t n, int32_t *out_linear,
                               void *stream);
pnb_status pnb_count_neighbors_f64(pnb_grid *g, const double *x, int64_t nx, const double *y,
                                   int64_t n, const int32_t *points, int64_t n_points,
                                   int index_base, int64_t *out, void *stream);
/* The fused n-body / WCSPH closures in Float64 (the reference is generic in the element type and
 * publishes Float64 WCSPH numbers, benchmarks/plot_benchmarks.jl:67): every array double, every
 * operation the IEEE double operation of the Julia closure in its order, candidates visited in
 * the reference's order -> sums bit-identical to the Float64 oracle.  One thread per point. */
pnb_status pnb_nbody_f64(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                         const int32_t *points, int64_t n_points, int index_base, const double *mass,
                         double G, double *dv, void *stream);
typedef struct pnb_wcsph_params_f64 {
    double smoothing_length, sound_speed, alpha, beta, epsilon, delta, kernel_norm;
} pnb_wcsph_params_f64;
pnb_status pnb_wcsph_interact_f64(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                  const int32_t *points, int64_t n_points, int index_base,
                                  const double *v_x, const double *v_y, const double *mass_x,
                                  const double *mass_y, const double *pressure_x,
                                  const double *pressure_y, const pnb_wcsph_params_f64 *params,
                                  double *dv, void *stream);
/* The same closures on a MIXED-precision search (pnb_grid_create_mixed: Float64 coordinates,
 * Float32 radius): the closure receives Float32 pos_diff / distance (nhs_grid.jl:547-555 under
 * Julia's promotion rules), the state arrays and dv are Float32 and the closure arithmetic is the
 * Float32 operation sequence -- sums bit-identical to the mixed oracle. */
pnb_status pnb_nbody_mixed(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                           const int32_t *points, int64_t n_points, int index_base, const float *mass,
                           float G, float *dv, void *stream);
pnb_status pnb_wcsph_interact_mixed(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                                    const int32_t *points, int64_t n_points, int index_base,
                                    const float *v_x, const float *v_y, const float *mass_x,
                                    const float *mass_y, const float *pressure_x,
                                    const float *pressure_y, const pnb_wcsph_params *params, float *dv,
                                    void *stream);
pnb_status pnb_nlist_build_f64(pnb_grid *g, const double *x, int64_t nx, const double *y, int64_t n,
                               int sort, pnb_nlist **out, void *stream);
pnb_status pnb_nlist_pairs_f64(const pnb_nlist *list, const pnb_grid *g, const double *x,
                               const double *y, double *pos_diff, double *distance, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Mixed precision: Float64 coordinates and cell-list corners with a Float32 search radius (and a
 * Float32 PeriodicBox, src/nhs_grid.jl:107-110), "common in SPH" according to
 * docs/literate/src/tut_gpu_usage.jl:45-50.  Julia's promotion rules decide the arithmetic: the
 * padding is computed in Float32 and added to the Float64 corners (src/cell_lists/full_grid.jl:66-67),
 * grid size and cell coordinates are Float64 (:74, :93), pos_diff = Float32.(x_i - y_j) and
 * everything after it Float32 (src/nhs_grid.jl:547-555).  The handle is used with the _f64 entry
 * points (build, point_cells, count, neighbour lists, pair geometry: pos_diff / distance are
 * returned as doubles holding the Float32 values).
 * ------------------------------------------------------------------------------------------- */
pnb_status pnb_grid_params_mixed(int ndims, float search_radius, const double *min_corner,
                                 const double *max_corner, const float *box_min,
                                 const float *box_max, double *padded_min, double *padded_max,
                                 int64_t *grid_size, int64_t *n_cells, float *cell_size);
pnb_status pnb_grid_create_mixed(int ndims, float search_radius, const double *min_corner,
                                 const double *max_corner, const float *box_min,
                                 const float *box_max, pnb_grid **out);
/* the same from the (already padded) corners stored in a FullGridCellList, as
 * pnb_grid_create_padded_f32 does for Float32: what the Julia glue's adapt calls */
pnb_status pnb_grid_create_padded_mixed(int ndims, float search_radius, const double *padded_min,
                                        const double *padded_max, const float *box_min,
                                        const float *box_max, pnb_grid **out);

#ifdef __cplusplus
}
#endif
#endif /* PNB200_H */
