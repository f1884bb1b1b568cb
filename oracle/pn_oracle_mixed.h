/*
 * pn_oracle_mixed.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Mixed precision restatement: Float64 coordinates (and cell-list corners) with a Float32
 * search radius, the combination docs/literate/src/tut_gpu_usage.jl:45-50 describes as common in
 * SPH.  What Julia's promotion rules make of the generic code:
 *   - FullGridCellList (full_grid.jl:66-67,74): pad = Float32(1001//1000) * r in Float32, then
 *     min_corner - pad / max_corner + pad in Float64; n_cells = ceil((max - min) / r) in Float64;
 *   - cell_size = r (Float32), periodic: box is Float32 (nhs_grid.jl:107-110 demands the radius'
 *     type), n_cells by the Float64 rule, cell_size = size / n_cells in Float32 (:117-118);
 *   - cell of a point (full_grid.jl:93): (x - min_corner) / cell_size in Float64;
 *   - pair (nhs_grid.jl:547-555): pos_diff = Float32.(x_i - y_j) (Float64 subtraction, one
 *     conversion), d2 / periodic fix / `<=` / sqrt in Float32.
 */
typedef struct {
    int32_t ndims;
    int32_t periodic;
    float search_radius;
    double min_corner[3], max_corner[3];   /* padded */
    int64_t grid_size[3];
    int64_t n_cells[3];
    float cell_size[3];
    float box_min[3], box_max[3], box_size[3];
} pno_grid_mix;

int pno_grid_init_mix(pno_grid_mix *g, int ndims, float r, const double *min_corner,
                      const double *max_corner, int periodic, const float *box_min,
                      const float *box_max)
{
    memset(g, 0, sizeof(*g));
    g->ndims = ndims;
    g->search_radius = r;
    float factor = 1001.0f / 1000.0f;
    float pad = factor * r;
    for (int d = 0; d < 3; d++) { g->grid_size[d] = 1; g->n_cells[d] = -1; g->cell_size[d] = r; }
    for (int d = 0; d < ndims; d++) {
        g->min_corner[d] = min_corner[d] - (double)pad;
        g->max_corner[d] = max_corner[d] + (double)pad;
        double q = (g->max_corner[d] - g->min_corner[d]) / (double)r;
        g->grid_size[d] = (int64_t)ceil(q);
    }
    if (periodic && !((double)r < 2.220446049250313e-16)) {
        g->periodic = 1;
        for (int d = 0; d < ndims; d++) {
            g->box_min[d] = box_min[d];
            g->box_max[d] = box_max[d];
            g->box_size[d] = box_max[d] - box_min[d];
            double nc = floor(((double)g->box_size[d] + 10.0 * 2.220446049250313e-16) / (double)r);
            g->n_cells[d] = (int64_t)nc;
            g->cell_size[d] = g->box_size[d] / (float)g->n_cells[d];
        }
        for (int d = 0; d < ndims; d++)
            if (g->n_cells[d] < 3) return 2;
    }
    return 0;
}

int64_t pno_total_cells_mix(const pno_grid_mix *g)
{
    return g->grid_size[0] * g->grid_size[1] * g->grid_size[2];
}

static inline void pno_periodic_cell_mix(const pno_grid_mix *g, int64_t *cell)
{
    if (!g->periodic) return;
    for (int d = 0; d < g->ndims; d++)
        cell[d] = pno_floormod_f64(pno_wrap_add_f64(cell[d], -2), g->n_cells[d]) + 2;
}

void pno_cell_coords_mix(const pno_grid_mix *g, const double *x, int64_t *cell)
{
    for (int d = 0; d < g->ndims; d++) {
        double q = (x[d] - g->min_corner[d]) / (double)g->cell_size[d];
        cell[d] = pno_wrap_add_f64(pno_floor_to_int_f64(q), 1);
    }
    for (int d = g->ndims; d < 3; d++) cell[d] = 1;
    pno_periodic_cell_mix(g, cell);
}

static inline int64_t pno_linear_mix(const pno_grid_mix *g, const int64_t *cell)
{
    return (cell[0] - 1) + (cell[1] - 1) * g->grid_size[0] +
           (cell[2] - 1) * g->grid_size[0] * g->grid_size[1];
}

void pno_point_cells_mix(const pno_grid_mix *g, const double *x, int64_t n, int64_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        int64_t cell[3];
        pno_cell_coords_mix(g, x + i * g->ndims, cell);
        int ok = 1;
        for (int d = 0; d < g->ndims; d++)
            if (cell[d] < 2 || cell[d] > g->grid_size[d] - 1) ok = 0;
        out[i] = ok ? pno_linear_mix(g, cell) : -1;
    }
}

/* deterministic CSR build (ids ascending inside a cell); returns 1 when a point is outside */
int pno_build_csr_mix(const pno_grid_mix *g, const double *y, int64_t n, int64_t *cell_start,
                      int32_t *cell_points)
{
    int64_t C = pno_total_cells_mix(g);
    for (int64_t c = 0; c <= C; c++) cell_start[c] = 0;
    if ((double)g->search_radius < 2.220446049250313e-16) return 0;
    int64_t *lin = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    pno_point_cells_mix(g, y, n, lin);
    for (int64_t k = 0; k < n; k++) {
        if (lin[k] < 0) { free(lin); return 1; }
        cell_start[lin[k] + 1]++;
    }
    for (int64_t c = 0; c < C; c++) cell_start[c + 1] += cell_start[c];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(C > 0 ? C : 1));
    memcpy(cursor, cell_start, sizeof(int64_t) * (size_t)C);
    for (int64_t k = 0; k < n; k++) cell_points[cursor[lin[k]]++] = (int32_t)k;
    free(cursor);
    free(lin);
    return 0;
}

/* pos_diff = Float32.(x_i - y_j), d2, periodic fix (only when d2 > r2), all in Float32 */
static inline float pno_pair_mix(const pno_grid_mix *g, const double *xi, const double *yj,
                                 float *p, float r2)
{
    const int nd = g->ndims;
    for (int d = 0; d < nd; d++) p[d] = (float)(xi[d] - yj[d]);
    float d2 = p[0] * p[0];
    for (int d = 1; d < nd; d++) d2 = d2 + p[d] * p[d];
    if (g->periodic && d2 > r2) {
        for (int k = 0; k < nd; k++) {
            float q = p[k] / g->box_size[k];
            float rq = (float)nearbyint((double)q);
            float t = g->box_size[k] * rq;
            p[k] = p[k] - t;
        }
        d2 = p[0] * p[0];
        for (int k = 1; k < nd; k++) d2 = d2 + p[k] * p[k];
    }
    return d2;
}

/* neighbour lists in the reference's visiting order (two passes: ids == NULL counts), returns 4
 * when a stencil leaves the grid; brute == 1: TrivialNeighborhoodSearch over all n points */
int pno_neighbor_lists_mix(const pno_grid_mix *g, const int64_t *cell_start,
                           const int32_t *cell_points, const double *x, int64_t nx,
                           const double *y, int64_t n, int64_t *offsets, int32_t *ids, int brute)
{
    const int nd = g->ndims;
    const float r = g->search_radius;
    const float r2 = r * r;
    int rc = 0;
    if (ids == NULL) offsets[0] = 0;
    for (int64_t i = 0; i < nx; i++) {
        const double *xi = x + i * nd;
        int64_t cnt = 0;
        float p[3] = {0, 0, 0};
        if (brute) {
            for (int64_t j = 0; j < n; j++)
                if (pno_pair_mix(g, xi, y + j * nd, p, r2) <= r2) {
                    if (ids) ids[offsets[i] + cnt] = (int32_t)j;
                    cnt++;
                }
        } else {
            int64_t cell[3];
            pno_cell_coords_mix(g, xi, cell);
            int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) { lo[d] = -1; hi[d] = 1; }
            for (int o3 = lo[2]; o3 <= hi[2]; o3++)
                for (int o2 = lo[1]; o2 <= hi[1]; o2++)
                    for (int o1 = lo[0]; o1 <= hi[0]; o1++) {
                        int64_t nc[3] = {cell[0] + o1, cell[1] + o2, cell[2] + o3};
                        pno_periodic_cell_mix(g, nc);
                        int ok = 1;
                        for (int d = 0; d < nd; d++)
                            if (nc[d] < 1 || nc[d] > g->grid_size[d]) ok = 0;
                        if (!ok) { rc = 4; continue; }
                        int64_t c = pno_linear_mix(g, nc);
                        for (int64_t k = cell_start[c]; k < cell_start[c + 1]; k++) {
                            int64_t j = cell_points[k];
                            if (pno_pair_mix(g, xi, y + j * nd, p, r2) <= r2) {
                                if (ids) ids[offsets[i] + cnt] = (int32_t)j;
                                cnt++;
                            }
                        }
                    }
        }
        if (ids == NULL) offsets[i + 1] = offsets[i] + cnt;
    }
    return rc;
}

/* what the closure receives when sweeping lists (nhs_precomputed.jl:230-238): Float32 values */
void pno_list_pairs_mix(const pno_grid_mix *g, const double *x, int64_t nx, const double *y,
                        const int64_t *offsets, const int32_t *ids, float *pos_diff, float *dist)
{
    const int nd = g->ndims;
    const float r = g->search_radius;
    const float r2 = r * r;
    for (int64_t i = 0; i < nx; i++)
        for (int64_t k = offsets[i]; k < offsets[i + 1]; k++) {
            float p[3] = {0, 0, 0};
            float d2 = pno_pair_mix(g, x + i * nd, y + (int64_t)ids[k] * nd, p, r2);
            for (int d = 0; d < nd; d++) pos_diff[k * nd + d] = p[d];
            dist[k] = (float)sqrt((double)d2);
        }
}

/* The fused closures on a mixed-precision search.  foreach_point_neighbor hands the closure the
 * Float32 pos_diff / distance of nhs_grid.jl:547-555 (distance = sqrt(distance2) in Float32, :555);
 * with Float32 state arrays the closures of benchmarks/n_body.jl:38-48 and of the WCSPH benchmark
 * (smoothed_particle_hydrodynamics.jl:45-102) then compute in Float32: exactly the Float32
 * instantiation of pno_cl_nbody / pno_cl_wcsph above.  Pairs in the reference's visiting order.
 * points == NULL: all nx points.  Returns 4 when a stencil leaves the grid. */
typedef void (*pno_pair_fn_mix)(void *ctx, int64_t i, int64_t j, const float *p, float d);

static int pno_foreach_mix(const pno_grid_mix *g, const int64_t *cell_start, const int32_t *cell_points,
                           const double *x, int64_t nx, const double *y, const int64_t *points,
                           int64_t npoints, pno_pair_fn_mix f, void *ctx)
{
    const int nd = g->ndims;
    const float r = g->search_radius;
    const float r2 = r * r;
    int rc = 0;
    const int64_t n_loop = points ? npoints : nx;
    for (int64_t t = 0; t < n_loop; t++) {
        const int64_t i = points ? points[t] : t;
        const double *xi = x + i * nd;
        float p[3] = {0, 0, 0};
        int64_t cell[3];
        pno_cell_coords_mix(g, xi, cell);
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int d = 0; d < nd; d++) { lo[d] = -1; hi[d] = 1; }
        for (int o3 = lo[2]; o3 <= hi[2]; o3++)
            for (int o2 = lo[1]; o2 <= hi[1]; o2++)
                for (int o1 = lo[0]; o1 <= hi[0]; o1++) {
                    int64_t nc[3] = {cell[0] + o1, cell[1] + o2, cell[2] + o3};
                    pno_periodic_cell_mix(g, nc);
                    int ok = 1;
                    for (int d = 0; d < nd; d++)
                        if (nc[d] < 1 || nc[d] > g->grid_size[d]) ok = 0;
                    if (!ok) { rc = 4; continue; }
                    int64_t c = pno_linear_mix(g, nc);
                    for (int64_t k = cell_start[c]; k < cell_start[c + 1]; k++) {
                        int64_t j = cell_points[k];
                        float d2 = pno_pair_mix(g, xi, y + j * nd, p, r2);
                        if (d2 <= r2) f(ctx, i, j, p, (float)sqrt((double)d2));
                    }
                }
    }
    return rc;
}

int pno_nbody_mix(const pno_grid_mix *g, const int64_t *cell_start, const int32_t *cell_points,
                  const double *x, int64_t nx, const double *y, const int64_t *points, int64_t npoints,
                  const float *mass, float G, float *dv)
{
    pno_nbody_ctx_f32 c = {g->ndims, mass, G, dv, NULL, NULL};
    for (int64_t i = 0; i < nx * g->ndims; i++) dv[i] = 0;     /* n_body.jl:36 */
    return pno_foreach_mix(g, cell_start, cell_points, x, nx, y, points, npoints, pno_cl_nbody_f32, &c);
}

int pno_wcsph_mix(const pno_grid_mix *g, const int64_t *cell_start, const int32_t *cell_points,
                  const double *x, int64_t nx, const double *y, const int64_t *points, int64_t npoints,
                  const float *v_x, const float *v_y, const float *mass_x, const float *mass_y,
                  const float *pressure_x, const float *pressure_y,
                  const float *params /*h,c,alpha,beta,eps,delta,norm*/, float *dv)
{
    pno_wcsph_ctx_f32 c;
    c.nd = g->ndims;
    c.v_x = v_x; c.v_y = v_y; c.mass_x = mass_x; c.mass_y = mass_y;
    c.pressure_x = pressure_x; c.pressure_y = pressure_y;
    c.h = params[0]; c.sound_speed = params[1]; c.alpha = params[2]; c.beta = params[3];
    c.epsilon = params[4]; c.delta = params[5]; c.kernel_norm = params[6];
    c.dv = dv; c.dv64 = NULL; c.dvabs = NULL;
    for (int64_t i = 0; i < nx * (g->ndims + 1); i++) dv[i] = 0;
    return pno_foreach_mix(g, cell_start, cell_points, x, nx, y, points, npoints, pno_cl_wcsph_f32, &c);
}
