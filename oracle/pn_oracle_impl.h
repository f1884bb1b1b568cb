/*
 * pn_oracle_impl.h -- TEST INFRASTRUCTURE ONLY (never shipped, never on the product path).
 *
 * CPU restatement of the PointNeighbors.jl hot path, instantiated once per element type by
 * pn_oracle.c (REAL = float -> suffix _f32, REAL = double -> suffix _f64).
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * Conventions:
 *   - point ids are 0-based here (the reference is 1-based; python adds 1 when comparing with
 *     the reference's golden vectors);
 *   - cell coordinates are 1-based exactly like the reference (user min_corner -> cell 2);
 *   - linear cell index is 0-based: (c1-1) + (c2-1)*n1 + (c3-1)*n1*n2  (full_grid.jl:157-161);
 *   - coordinates are NDIMS x N column-major == xyzxyz... (neighborhood_search.jl:25-26).
 *
 * Must be compiled with -ffp-contract=off and without -ffast-math: every product and sum is
 * individually rounded, exactly like the Julia code.
 *
 * Parity pinning: the oracle is pinned against the reference's own golden vectors
 * (tests/golden/ json files, transcribed from /root/reference/test) by tests/test_oracle_golden.py.
 * The WCSPH / TLSPH pair arithmetic lives in TrixiParticles.jl (not vendored, compat "0.5",
 * test/Project.toml:16) and no reference test checks its values: for those two closures
 * parity is UNPINNED and the formulas below are this repo's definition.
 */

#ifndef REAL
#error "define REAL and SUF before including"
#endif

#define PNO_CAT_(a, b) a##b
#define PNO_CAT(a, b) PNO_CAT_(a, b)
#define FN(name) PNO_CAT(name, SUF)
#define GRID FN(pno_grid)

typedef struct {
    int32_t ndims;
    int32_t periodic;
    REAL search_radius;
    REAL min_corner[3];   /* padded corners (full_grid.jl:66-67) */
    REAL max_corner[3];
    int64_t grid_size[3]; /* n_cells_per_dimension incl. padding (full_grid.jl:74); 1 for unused dims */
    int64_t n_cells[3];   /* periodic cells per dim or -1 (nhs_grid.jl:104,117) */
    REAL cell_size[3];    /* nhs_grid.jl:105,118 */
    REAL box_min[3], box_max[3], box_size[3]; /* neighborhood_search.jl:129-141 */
} GRID;

/* util.jl:19-34 floor_to_int: saturating floor -> Int64 */
static inline int64_t FN(pno_floor_to_int)(REAL v)
{
    REAL rounded = (REAL)floor((double)v); /* floor is exact in either precision */
    if (isnan(rounded) || rounded >= (REAL)9223372036854775808.0) return INT64_MAX;
    if (rounded <= (REAL)-9223372036854775808.0) return INT64_MIN;
    return (int64_t)rounded;
}

/* Julia Int arithmetic wraps silently */
static inline int64_t FN(pno_wrap_add)(int64_t a, int64_t b)
{
    return (int64_t)((uint64_t)a + (uint64_t)b);
}

/* Julia mod(a, n) for n > 0: floored modulo */
static inline int64_t FN(pno_floormod)(int64_t a, int64_t n)
{
    int64_t m = a % n;
    return (m < 0) ? m + n : m;
}

/*
 * FullGridCellList ctor (full_grid.jl:48-82) + GridNeighborhoodSearch ctor (nhs_grid.jl:77-129)
 * + PeriodicBox (neighborhood_search.jl:134-140).
 * returns 0 ok, 2 = "needs at least 3 cells in each dimension" (nhs_grid.jl:120-124).
 */
int FN(pno_grid_init)(GRID *g, int ndims, REAL r, const REAL *min_corner, const REAL *max_corner,
                      int periodic, const REAL *box_min, const REAL *box_max)
{
    memset(g, 0, sizeof(*g));
    g->ndims = ndims;
    g->search_radius = r;
    /* `1001 // 1000 * search_radius`: Rational -> REAL is num/den evaluated in REAL */
    REAL factor = (REAL)1001 / (REAL)1000;
    REAL pad = factor * r;
    for (int d = 0; d < 3; d++) {
        g->grid_size[d] = 1;
        g->n_cells[d] = -1;
        g->cell_size[d] = r;
    }
    for (int d = 0; d < ndims; d++) {
        g->min_corner[d] = min_corner[d] - pad;
        g->max_corner[d] = max_corner[d] + pad;
        /* ceil.(Int, (max .- min) ./ r) evaluated in REAL */
        REAL q = (g->max_corner[d] - g->min_corner[d]) / r;
        g->grid_size[d] = (int64_t)(REAL)ceil((double)q);
    }
    g->periodic = 0;
    /* nhs_grid.jl:102: `search_radius < eps() || isnothing(periodic_box)` (Float64 eps) */
    if (periodic && !((double)r < 2.220446049250313e-16)) {
        g->periodic = 1;
        for (int d = 0; d < ndims; d++) {
            g->box_min[d] = box_min[d];
            g->box_max[d] = box_max[d];
            g->box_size[d] = box_max[d] - box_min[d];
            /* nhs_grid.jl:117: (size .+ 10eps()) / r -- promoted to Float64 whatever REAL is */
            double nc = floor(((double)g->box_size[d] + 10.0 * 2.220446049250313e-16) / (double)r);
            g->n_cells[d] = (int64_t)nc;
            /* nhs_grid.jl:118: size ./ n_cells in REAL */
            g->cell_size[d] = g->box_size[d] / (REAL)g->n_cells[d];
        }
        for (int d = 0; d < ndims; d++)
            if (g->n_cells[d] < 3) return 2;
    }
    return 0;
}

int64_t FN(pno_total_cells)(const GRID *g)
{
    return g->grid_size[0] * g->grid_size[1] * g->grid_size[2];
}

/* nhs_grid.jl:599-620 periodic_cell_index */
static inline void FN(pno_periodic_cell)(const GRID *g, int64_t *cell)
{
    if (!g->periodic) return;
    for (int d = 0; d < g->ndims; d++)
        cell[d] = FN(pno_floormod)(FN(pno_wrap_add)(cell[d], -2), g->n_cells[d]) + 2;
}

/* nhs_grid.jl:622-628 cell_coords ; full_grid.jl:84-94 nonperiodic_cell_coords */
void FN(pno_cell_coords)(const GRID *g, const REAL *x, int64_t *cell)
{
    for (int d = 0; d < g->ndims; d++) {
        REAL q = (x[d] - g->min_corner[d]) / g->cell_size[d];
        cell[d] = FN(pno_wrap_add)(FN(pno_floor_to_int)(q), 1);
    }
    for (int d = g->ndims; d < 3; d++) cell[d] = 1;
    FN(pno_periodic_cell)(g, cell);
}

/* full_grid.jl:205-213 check_cell_bounds: valid cells are 2:(size-1) */
static inline int FN(pno_cell_in_bounds)(const GRID *g, const int64_t *cell)
{
    for (int d = 0; d < g->ndims; d++)
        if (cell[d] < 2 || cell[d] > g->grid_size[d] - 1) return 0;
    return 1;
}

/* full_grid.jl:157-161 cell_index (LinearIndices, column-major), returned 0-based */
static inline int64_t FN(pno_linear)(const GRID *g, const int64_t *cell)
{
    return (cell[0] - 1) + (cell[1] - 1) * g->grid_size[0] +
           (cell[2] - 1) * g->grid_size[0] * g->grid_size[1];
}

/* linear (0-based) cell index of each point, or -1 if outside (K2's arithmetic). */
void FN(pno_point_cells)(const GRID *g, const REAL *x, int64_t n, int64_t *out)
{
    for (int64_t i = 0; i < n; i++) {
        int64_t cell[3];
        FN(pno_cell_coords)(g, x + i * g->ndims, cell);
        out[i] = FN(pno_cell_in_bounds)(g, cell) ? FN(pno_linear)(g, cell) : -1;
    }
}

/*
 * initialize_grid! (nhs_grid.jl:255-281) restated as a deterministic counting sort:
 * cell_start[C+1], cell_points[n_idx] with the ids of every cell in ascending order (the
 * reference's order inside a cell is atomic-arrival order, i.e. unspecified;
 * vector_of_vectors.jl:99-109).  idx == NULL means eachindex_y = all.
 * returns 0 ok, 1 = "particle coordinates are NaN or outside the domain bounds"
 * (full_grid.jl:211).
 */
int FN(pno_build_csr)(const GRID *g, const REAL *y, int64_t n, const int64_t *idx, int64_t n_idx,
                      int64_t *cell_start, int32_t *cell_points)
{
    int64_t C = FN(pno_total_cells)(g);
    if (idx == NULL) n_idx = n;
    for (int64_t c = 0; c <= C; c++) cell_start[c] = 0;
    /* nhs_grid.jl:263: zero radius -> emptied list, return */
    if ((double)g->search_radius < 2.220446049250313e-16) return 0;
    int64_t *lin = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_idx > 0 ? n_idx : 1));
    int rc = 0;
    for (int64_t k = 0; k < n_idx; k++) {
        int64_t p = idx ? idx[k] : k;
        int64_t cell[3];
        FN(pno_cell_coords)(g, y + p * g->ndims, cell);
        if (!FN(pno_cell_in_bounds)(g, cell)) { rc = 1; break; }
        lin[k] = FN(pno_linear)(g, cell);
        cell_start[lin[k] + 1]++;
    }
    if (rc) { free(lin); return rc; }
    for (int64_t c = 0; c < C; c++) cell_start[c + 1] += cell_start[c];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(C > 0 ? C : 1));
    memcpy(cursor, cell_start, sizeof(int64_t) * (size_t)C);
    /* ascending ids inside a cell require visiting points in ascending id order */
    if (idx == NULL) {
        for (int64_t k = 0; k < n_idx; k++) cell_points[cursor[lin[k]]++] = (int32_t)k;
    } else {
        /* idx may be unsorted: insertion keeps the cell's ids ascending */
        for (int64_t k = 0; k < n_idx; k++) {
            int64_t c = lin[k];
            int64_t pos = cursor[c]++;
            int32_t id = (int32_t)idx[k];
            while (pos > cell_start[c] && cell_points[pos - 1] > id) {
                cell_points[pos] = cell_points[pos - 1];
                pos--;
            }
            cell_points[pos] = id;
        }
    }
    free(cursor);
    free(lin);
    return 0;
}

/*
 * The reference's own data structure and build, used as the timed CPU baseline:
 * DynamicVectorOfVectors{Int32} = dense max_inner x C matrix + lengths (vector_of_vectors.jl:3-31),
 * empty! (full_grid.jl:96-105), then one atomic push per point (nhs_grid.jl:271-278,
 * full_grid.jl:128-139, vector_of_vectors.jl:93-112), parallel over points with static chunks
 * like Polyester.@batch (util.jl:133-137).
 * returns 0 ok, 1 out of domain, 3 "cell list is full" (vector_of_vectors.jl:114-121; note the
 * reference elides this check through @inbounds, full_grid.jl:136).
 */
int FN(pno_build_dvov)(const GRID *g, const REAL *y, int64_t n, int32_t max_inner,
                       int32_t *backend, int32_t *lengths)
{
    int64_t C = FN(pno_total_cells)(g);
    int rc = 0;
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < C; c++) lengths[c] = 0;
    if ((double)g->search_radius < 2.220446049250313e-16) return 0;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; p++) {
        int64_t cell[3];
        FN(pno_cell_coords)(g, y + p * g->ndims, cell);
        if (!FN(pno_cell_in_bounds)(g, cell)) {
#pragma omp atomic write
            rc = 1;
            continue;
        }
        int64_t c = FN(pno_linear)(g, cell);
        int32_t new_len;
#pragma omp atomic capture
        new_len = ++lengths[c];
        if (new_len > max_inner) {
#pragma omp atomic write
            rc = 3;
            continue;
        }
        backend[c * (int64_t)max_inner + (new_len - 1)] = (int32_t)p;
    }
    return rc;
}

/* ---------------------------------------------------------------------------------------
 * Cell-list views: CSR (oracle) or DVoV (reference layout)
 * ------------------------------------------------------------------------------------- */
typedef struct {
    const int64_t *cell_start; /* CSR, or NULL */
    const int32_t *cell_points;
    const int32_t *backend; /* DVoV, or NULL */
    const int32_t *lengths;
    int32_t max_inner;
} FN(pno_cells);

static inline int32_t FN(pno_cell_ids)(const FN(pno_cells) * v, int64_t c, const int32_t **ids)
{
    if (v->cell_start) {
        *ids = v->cell_points + v->cell_start[c];
        return (int32_t)(v->cell_start[c + 1] - v->cell_start[c]);
    }
    *ids = v->backend + c * (int64_t)v->max_inner;
    return v->lengths[c];
}

typedef void (*FN(pno_pair_fn))(void *ctx, int64_t i, int64_t j, const REAL *pos_diff, REAL d);

/* neighborhood_search.jl:423-436 compute_periodic_distance (only when d2 > r^2) */
static inline REAL FN(pno_periodic_fix)(const GRID *g, REAL *p, REAL d2, REAL r2)
{
    if (g->periodic && d2 > r2) {
        for (int k = 0; k < g->ndims; k++) {
            REAL q = p[k] / g->box_size[k];
            REAL rq = (REAL)nearbyint((double)q); /* Julia round = ties-to-even */
            REAL t = g->box_size[k] * rq;
            p[k] = p[k] - t;
        }
        d2 = p[0] * p[0];
        for (int k = 1; k < g->ndims; k++) d2 = d2 + p[k] * p[k];
    }
    return d2;
}

/*
 * mapreduce_neighbor_inner(::GridNeighborhoodSearch) (nhs_grid.jl:519-575) for one point:
 * 3^d neighbour cells in CartesianIndices order (dim 1 fastest, :577-583), each wrapped
 * (:593-597), pos_diff / dot / periodic fix / `<=` / sqrt.
 * returns 0 ok, 4 if a neighbour cell is outside the grid (the safe variant's BoundsError).
 */
static inline __attribute__((always_inline)) int
FN(pno_sweep_point)(const GRID *g, const FN(pno_cells) * cells, const REAL *xi, int64_t i,
                    const REAL *y, FN(pno_pair_fn) f, void *ctx)
{
    const int nd = g->ndims;
    const REAL r = g->search_radius;
    const REAL r2 = r * r;
    int64_t cell[3];
    FN(pno_cell_coords)(g, xi, cell);
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < nd; d++) { lo[d] = -1; hi[d] = 1; }
    for (int o3 = lo[2]; o3 <= hi[2]; o3++)
        for (int o2 = lo[1]; o2 <= hi[1]; o2++)
            for (int o1 = lo[0]; o1 <= hi[0]; o1++) {
                int64_t nc[3] = {FN(pno_wrap_add)(cell[0], o1), FN(pno_wrap_add)(cell[1], o2),
                                 FN(pno_wrap_add)(cell[2], o3)};
                FN(pno_periodic_cell)(g, nc);
                for (int d = 0; d < nd; d++)
                    if (nc[d] < 1 || nc[d] > g->grid_size[d]) return 4;
                const int32_t *ids;
                int32_t cnt = FN(pno_cell_ids)(cells, FN(pno_linear)(g, nc), &ids);
                for (int32_t k = 0; k < cnt; k++) {
                    int64_t j = ids[k];
                    const REAL *yj = y + j * nd;
                    REAL p[3] = {0, 0, 0};
                    for (int d = 0; d < nd; d++) p[d] = xi[d] - yj[d];
                    REAL d2 = p[0] * p[0];
                    for (int d = 1; d < nd; d++) d2 = d2 + p[d] * p[d];
                    d2 = FN(pno_periodic_fix)(g, p, d2, r2);
                    if (d2 <= r2) {
                        REAL dist = (REAL)sqrt((double)d2); /* correctly rounded in REAL */
                        f(ctx, i, j, p, dist);
                    }
                }
            }
    return 0;
}

/* foreach_point_neighbor (neighborhood_search.jl:183-201): loop over `points` (NULL = all). */
static inline __attribute__((always_inline)) int
FN(pno_foreach_point_neighbor)(const GRID *g, const FN(pno_cells) * cells, const REAL *x,
                               int64_t nx, const REAL *y, const int64_t *points, int64_t npoints,
                               FN(pno_pair_fn) f, void *ctx, int parallel)
{
    if (points == NULL) npoints = nx;
    int rc = 0;
    if ((double)g->search_radius < 2.220446049250313e-16) return 0;
#pragma omp parallel for schedule(static) if (parallel)
    for (int64_t k = 0; k < npoints; k++) {
        int64_t i = points ? points[k] : k;
        int e = FN(pno_sweep_point)(g, cells, x + i * g->ndims, i, y, f, ctx);
        if (e) {
#pragma omp atomic write
            rc = e;
        }
    }
    return rc;
}

static inline FN(pno_cells) FN(pno_view_csr)(const int64_t *cell_start, const int32_t *cell_points)
{
    FN(pno_cells) v = {cell_start, cell_points, NULL, NULL, 0};
    return v;
}
static inline FN(pno_cells) FN(pno_view_dvov)(const int32_t *backend, const int32_t *lengths,
                                               int32_t max_inner)
{
    FN(pno_cells) v = {NULL, NULL, backend, lengths, max_inner};
    return v;
}

/* ---------------------------------------------------------------------------------------
 * closures
 * ------------------------------------------------------------------------------------- */

/* benchmarks/count_neighbors.jl:24-27: n_neighbors[i] += 1 (Int64) */
static void FN(pno_cl_count)(void *ctx, int64_t i, int64_t j, const REAL *p, REAL d)
{
    (void)j; (void)p; (void)d;
    ((int64_t *)ctx)[i] += 1;
}

typedef struct {
    int nd;
    const REAL *mass;
    REAL G;
    REAL *dv;
    double *dv64;   /* optional: Float64 accumulation of the REAL per-pair terms */
    double *dvabs;  /* optional: sum of |term| */
} FN(pno_nbody_ctx);

/* benchmarks/n_body.jl:38-48 */
static void FN(pno_cl_nbody)(void *ctx_, int64_t i, int64_t j, const REAL *p, REAL d)
{
    FN(pno_nbody_ctx) *c = (FN(pno_nbody_ctx) *)ctx_;
#if REAL_IS_FLOAT
    const REAL sqrt_eps = 3.4526698300124393e-4f; /* sqrt(eps(Float32)) */
#else
    const REAL sqrt_eps = 1.4901161193847656e-8;  /* sqrt(eps(Float64)) */
#endif
    if (d < sqrt_eps) return;
    REAL t = (-c->G) * c->mass[j];
    REAL d3 = (d * d) * d;
    for (int k = 0; k < c->nd; k++) {
        REAL a = (t * p[k]) / d3;
        c->dv[i * c->nd + k] += a;
        if (c->dv64) c->dv64[i * c->nd + k] += (double)a;
        if (c->dvabs) c->dvabs[i * c->nd + k] += fabs((double)a);
    }
}

/*
 * WCSPH continuity + momentum pair term.  The arithmetic belongs to TrixiParticles.jl
 * (interact! called at benchmarks/smoothed_particle_hydrodynamics.jl:101; system set-up
 * :54-83): WendlandC2 kernel gradient, pressure acceleration for ContinuityDensity,
 * ArtificialViscosityMonaghan(alpha, beta, epsilon = 0.01), continuity equation and
 * DensityDiffusionMolteniColagrossi(delta).  PARITY UNPINNED (not vendored, no reference
 * test checks values) -- this is the repo's definition, see DESIGN.md section 3.
 *   v  : (nd+1) x N, rows 1..nd velocity, row nd+1 density     (:92)
 *   dv : (nd+1) x N, rows 1..nd acceleration, row nd+1 d(rho)/dt
 */
typedef struct {
    int nd;
    const REAL *v_x;   /* state of the points looped over   */
    const REAL *v_y;   /* state of the neighbour points      */
    const REAL *mass_x, *mass_y;
    const REAL *pressure_x, *pressure_y;
    REAL h, sound_speed, alpha, beta, epsilon, delta;
    REAL kernel_norm;  /* sigma_d / h^d */
    REAL *dv;
    double *dv64, *dvabs;
} FN(pno_wcsph_ctx);

static void FN(pno_cl_wcsph)(void *ctx_, int64_t i, int64_t j, const REAL *p, REAL d)
{
    FN(pno_wcsph_ctx) *c = (FN(pno_wcsph_ctx) *)ctx_;
    const int nd = c->nd;
    const int ns = nd + 1;
#if REAL_IS_FLOAT
    const REAL sqrt_eps = 3.4526698300124393e-4f;
#else
    const REAL sqrt_eps = 1.4901161193847656e-8;
#endif
    const REAL *va = c->v_x + i * ns, *vb = c->v_y + j * ns;
    REAL rho_a = va[nd], rho_b = vb[nd];
    REAL rho_mean = (REAL)0.5 * (rho_a + rho_b);
    REAL m_b = c->mass_y[j];
    REAL p_a = c->pressure_x[i], p_b = c->pressure_y[j];
    /* kernel gradient: zero for d < sqrt(eps) (self pair) */
    REAL grad[3] = {0, 0, 0};
    if (!(d < sqrt_eps)) {
        REAL q = d / c->h;
        REAL w = 0;
        if (q < (REAL)2) {
            REAL t = (REAL)1 - q * (REAL)0.5;
            w = ((REAL)-5 * q) * ((t * t) * t);
        }
        REAL dw = (c->kernel_norm / c->h) * w;
        REAL s = dw / d;
        for (int k = 0; k < nd; k++) grad[k] = s * p[k];
    }
    /* pressure acceleration (ContinuityDensity): -m_b (p_a + p_b) / (rho_a rho_b) grad */
    REAL pf = ((-m_b) * (p_a + p_b)) / (rho_a * rho_b);
    /* Monaghan artificial viscosity, only for approaching particles */
    REAL vdiff[3] = {0, 0, 0};
    REAL vr = 0, vg = 0;
    for (int k = 0; k < nd; k++) vdiff[k] = va[k] - vb[k];
    vr = vdiff[0] * p[0];
    for (int k = 1; k < nd; k++) vr = vr + vdiff[k] * p[k];
    REAL visc = 0;
    if (vr < 0) {
        REAL mu = (c->h * vr) / (d * d + c->epsilon * (c->h * c->h));
        REAL pi_ab = (c->alpha * c->sound_speed * mu - c->beta * (mu * mu)) / rho_mean;
        visc = m_b * pi_ab;
    }
    REAL term[4] = {0, 0, 0, 0};
    for (int k = 0; k < nd; k++) term[k] = pf * grad[k] + visc * grad[k];
    /* continuity: rho_a / rho_b * m_b * dot(vdiff, grad) */
    vg = vdiff[0] * grad[0];
    for (int k = 1; k < nd; k++) vg = vg + vdiff[k] * grad[k];
    REAL drho = ((rho_a / rho_b) * m_b) * vg;
    /* Molteni-Colagrossi density diffusion, skipped for d < sqrt(eps) */
    if (!(d < sqrt_eps)) {
        REAL vol_b = m_b / rho_b;
        REAL two_drho = (REAL)2 * (rho_a - rho_b);
        REAL dd = d * d;
        REAL pg = 0;
        for (int k = 0; k < nd; k++) {
            REAL psi = (two_drho * p[k]) / dd;
            pg = (k == 0) ? psi * grad[k] : pg + psi * grad[k];
        }
        drho = drho + ((c->delta * c->h) * c->sound_speed) * (pg * vol_b);
    }
    term[nd] = drho;
    for (int k = 0; k < ns; k++) {
        c->dv[i * ns + k] += term[k];
        if (c->dv64) c->dv64[i * ns + k] += (double)term[k];
        if (c->dvabs) c->dvabs[i * ns + k] += fabs((double)term[k]);
    }
}

/* ---------------------------------------------------------------------------------------
 * exported sweeps.  layout: 0 = CSR (oracle order), 1 = DVoV (reference structure, CPU baseline)
 * ------------------------------------------------------------------------------------- */
#define PNO_VIEW(v)                                                                      \
    FN(pno_cells) v = backend ? FN(pno_view_dvov)(backend, lengths, max_inner)           \
                              : FN(pno_view_csr)(cell_start, cell_points)

int FN(pno_count_neighbors)(const GRID *g, const int64_t *cell_start, const int32_t *cell_points,
                            const int32_t *backend, const int32_t *lengths, int32_t max_inner,
                            const REAL *x, int64_t nx, const REAL *y, const int64_t *points,
                            int64_t npoints, int64_t *out, int parallel)
{
    PNO_VIEW(v);
    /* count_neighbors.jl:22 `n_neighbors .= 0` */
    for (int64_t i = 0; i < nx; i++) out[i] = 0;
    return FN(pno_foreach_point_neighbor)(g, &v, x, nx, y, points, npoints, FN(pno_cl_count), out,
                                          parallel);
}

int FN(pno_nbody)(const GRID *g, const int64_t *cell_start, const int32_t *cell_points,
                  const int32_t *backend, const int32_t *lengths, int32_t max_inner, const REAL *x,
                  int64_t nx, const REAL *y, const int64_t *points, int64_t npoints,
                  const REAL *mass, REAL G, REAL *dv, double *dv64, double *dvabs, int parallel)
{
    PNO_VIEW(v);
    FN(pno_nbody_ctx) c = {g->ndims, mass, G, dv, dv64, dvabs};
    /* n_body.jl:36 `dv .= 0` */
    for (int64_t i = 0; i < nx * g->ndims; i++) {
        dv[i] = 0;
        if (dv64) dv64[i] = 0;
        if (dvabs) dvabs[i] = 0;
    }
    return FN(pno_foreach_point_neighbor)(g, &v, x, nx, y, points, npoints, FN(pno_cl_nbody), &c,
                                          parallel);
}

int FN(pno_wcsph)(const GRID *g, const int64_t *cell_start, const int32_t *cell_points,
                  const int32_t *backend, const int32_t *lengths, int32_t max_inner, const REAL *x,
                  int64_t nx, const REAL *y, const int64_t *points, int64_t npoints,
                  const REAL *v_x, const REAL *v_y, const REAL *mass_x, const REAL *mass_y,
                  const REAL *pressure_x, const REAL *pressure_y, const REAL *params /*h,c,alpha,beta,eps,delta,norm*/,
                  REAL *dv, double *dv64, double *dvabs, int parallel)
{
    PNO_VIEW(v);
    FN(pno_wcsph_ctx) c;
    c.nd = g->ndims;
    c.v_x = v_x; c.v_y = v_y; c.mass_x = mass_x; c.mass_y = mass_y;
    c.pressure_x = pressure_x; c.pressure_y = pressure_y;
    c.h = params[0]; c.sound_speed = params[1]; c.alpha = params[2]; c.beta = params[3];
    c.epsilon = params[4]; c.delta = params[5]; c.kernel_norm = params[6];
    c.dv = dv; c.dv64 = dv64; c.dvabs = dvabs;
    int ns = g->ndims + 1;
    /* interact! accumulates into dv; the benchmark passes dv = zero(v) (:95) */
    for (int64_t i = 0; i < nx * ns; i++) {
        dv[i] = 0;
        if (dv64) dv64[i] = 0;
        if (dvabs) dvabs[i] = 0;
    }
    return FN(pno_foreach_point_neighbor)(g, &v, x, nx, y, points, npoints, FN(pno_cl_wcsph), &c,
                                          parallel);
}

/* ---------------------------------------------------------------------------------------
 * neighbour lists
 * ------------------------------------------------------------------------------------- */
typedef struct {
    int64_t *counts;
    const int64_t *offsets;
    int32_t *ids;
} FN(pno_list_ctx);

static void FN(pno_cl_list_count)(void *ctx, int64_t i, int64_t j, const REAL *p, REAL d)
{
    (void)j; (void)p; (void)d;
    ((FN(pno_list_ctx) *)ctx)->counts[i] += 1;
}
static void FN(pno_cl_list_fill)(void *ctx_, int64_t i, int64_t j, const REAL *p, REAL d)
{
    (void)p; (void)d;
    FN(pno_list_ctx) *c = (FN(pno_list_ctx) *)ctx_;
    c->ids[c->offsets[i] + c->counts[i]++] = (int32_t)j;
}

static int FN(pno_cmp_i32)(const void *a, const void *b)
{
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}

/*
 * initialize_neighbor_lists! (nhs_precomputed.jl:187-206) as CSR: pass 1 (ids == NULL) fills
 * offsets[nx+1]; pass 2 fills ids[offsets[nx]] in sweep order, then sorts each list when
 * `sort` (sorteach!, vector_of_vectors.jl:177-183).
 */
int FN(pno_neighbor_lists)(const GRID *g, const int64_t *cell_start, const int32_t *cell_points,
                           const REAL *x, int64_t nx, const REAL *y, int64_t *offsets,
                           int32_t *ids, int sort)
{
    FN(pno_cells) v = FN(pno_view_csr)(cell_start, cell_points);
    int64_t *counts = (int64_t *)calloc((size_t)(nx > 0 ? nx : 1), sizeof(int64_t));
    FN(pno_list_ctx) c = {counts, offsets, ids};
    int rc;
    if (ids == NULL) {
        rc = FN(pno_foreach_point_neighbor)(g, &v, x, nx, y, NULL, nx, FN(pno_cl_list_count), &c, 1);
        offsets[0] = 0;
        for (int64_t i = 0; i < nx; i++) offsets[i + 1] = offsets[i] + counts[i];
    } else {
        rc = FN(pno_foreach_point_neighbor)(g, &v, x, nx, y, NULL, nx, FN(pno_cl_list_fill), &c, 1);
        if (sort) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < nx; i++)
                qsort(ids + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), sizeof(int32_t),
                      FN(pno_cmp_i32));
        }
    }
    free(counts);
    return rc;
}

/*
 * TrivialNeighborhoodSearch (nhs_trivial.jl:55 eachneighbor = all points) through the generic
 * inner loop (neighborhood_search.jl:390-421): brute force, used exactly like the reference's
 * tests use it (test/neighborhood_search.jl:209-222).  Two passes like pno_neighbor_lists.
 */
int FN(pno_trivial_lists)(int ndims, REAL r, int periodic, const REAL *box_min,
                          const REAL *box_max, const REAL *x, int64_t nx, const REAL *y, int64_t n,
                          int64_t *offsets, int32_t *ids)
{
    GRID g;
    memset(&g, 0, sizeof(g));
    g.ndims = ndims;
    g.periodic = periodic;
    g.search_radius = r;
    for (int d = 0; d < ndims && periodic; d++) g.box_size[d] = box_max[d] - box_min[d];
    const REAL r2 = r * r;
    if (ids == NULL) offsets[0] = 0;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++) {
        int64_t cnt = 0;
        for (int64_t j = 0; j < n; j++) {
            REAL p[3] = {0, 0, 0};
            for (int d = 0; d < ndims; d++) p[d] = x[i * ndims + d] - y[j * ndims + d];
            REAL d2 = p[0] * p[0];
            for (int d = 1; d < ndims; d++) d2 = d2 + p[d] * p[d];
            d2 = FN(pno_periodic_fix)(&g, p, d2, r2);
            if (d2 <= r2) {
                if (ids) ids[offsets[i] + cnt] = (int32_t)j;
                cnt++;
            }
        }
        if (ids == NULL) offsets[i + 1] = cnt;
    }
    if (ids == NULL)
        for (int64_t i = 0; i < nx; i++) offsets[i + 1] += offsets[i];
    return 0;
}

/*
 * Sweep over precomputed lists (nhs_precomputed.jl:210-247): pos_diff, d2, periodic fix,
 * d = sqrt(d2), f -- NO radius test.  Outputs pos_diff (nd per pair) and distance per pair in
 * list order so tests can pin what the closure receives.
 */
void FN(pno_list_pairs)(int ndims, REAL r, int periodic, const REAL *box_min, const REAL *box_max,
                        const REAL *x, int64_t nx, const REAL *y, const int64_t *offsets,
                        const int32_t *ids, REAL *pos_diff, REAL *dist)
{
    GRID g;
    memset(&g, 0, sizeof(g));
    g.ndims = ndims;
    g.periodic = periodic;
    for (int d = 0; d < ndims && periodic; d++) g.box_size[d] = box_max[d] - box_min[d];
    const REAL r2 = r * r;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; i++)
        for (int64_t k = offsets[i]; k < offsets[i + 1]; k++) {
            int64_t j = ids[k];
            REAL p[3] = {0, 0, 0};
            for (int d = 0; d < ndims; d++) p[d] = x[i * ndims + d] - y[j * ndims + d];
            REAL d2 = p[0] * p[0];
            for (int d = 1; d < ndims; d++) d2 = d2 + p[d] * p[d];
            d2 = FN(pno_periodic_fix)(&g, p, d2, r2);
            for (int d = 0; d < ndims; d++) pos_diff[k * ndims + d] = p[d];
            dist[k] = (REAL)sqrt((double)d2);
        }
}

/*
 * TLSPH deformation gradient over precomputed lists (TrixiParticles calc_deformation_grad!,
 * called at benchmarks/smoothed_particle_hydrodynamics.jl:142).  PARITY UNPINNED, repo's
 * definition:
 *   F_i = sum_j  -(m0_j / rho0_j) * (x_i - x_j)_current (outer) (L_i * gradW(X_i - X_j))
 * with neighbours, gradW and the periodic fix evaluated on the INITIAL coordinates X0.
 * F and L are nd x nd column-major per point.
 */
void FN(pno_tlsph_deformation_grad)(int ndims, REAL r, int periodic, const REAL *box_min,
                                    const REAL *box_max, const REAL *X0, const REAL *xcur,
                                    int64_t n, const int64_t *offsets, const int32_t *ids,
                                    const REAL *mass, const REAL *rho0, const REAL *L, REAL h,
                                    REAL kernel_norm, REAL *F, double *F64, double *Fabs)
{
    GRID g;
    memset(&g, 0, sizeof(g));
    g.ndims = ndims;
    g.periodic = periodic;
    for (int d = 0; d < ndims && periodic; d++) g.box_size[d] = box_max[d] - box_min[d];
    const REAL r2 = r * r;
    const int nd = ndims, nn = ndims * ndims;
#if REAL_IS_FLOAT
    const REAL sqrt_eps = 3.4526698300124393e-4f;
#else
    const REAL sqrt_eps = 1.4901161193847656e-8;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        REAL acc[9] = {0};
        double acc64[9] = {0}, accabs[9] = {0};
        for (int64_t k = offsets[i]; k < offsets[i + 1]; k++) {
            int64_t j = ids[k];
            REAL p[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) p[d] = X0[i * nd + d] - X0[j * nd + d];
            REAL d2 = p[0] * p[0];
            for (int d = 1; d < nd; d++) d2 = d2 + p[d] * p[d];
            d2 = FN(pno_periodic_fix)(&g, p, d2, r2);
            REAL dist = (REAL)sqrt((double)d2);
            if (dist < sqrt_eps) continue;
            REAL q = dist / h, w = 0;
            if (q < (REAL)2) {
                REAL t = (REAL)1 - q * (REAL)0.5;
                w = ((REAL)-5 * q) * ((t * t) * t);
            }
            REAL s = ((kernel_norm / h) * w) / dist;
            REAL grad[3] = {0, 0, 0}, lg[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) grad[d] = s * p[d];
            /* lg = L_i * grad (column-major L) */
            for (int a = 0; a < nd; a++) {
                REAL t = L[i * nn + a] * grad[0];
                for (int b = 1; b < nd; b++) t = t + L[i * nn + b * nd + a] * grad[b];
                lg[a] = t;
            }
            REAL vol = mass[j] / rho0[j];
            for (int b = 0; b < nd; b++)
                for (int a = 0; a < nd; a++) {
                    REAL cd = xcur[i * nd + a] - xcur[j * nd + a];
                    REAL term = ((-vol) * cd) * lg[b];
                    acc[b * nd + a] += term;
                    acc64[b * nd + a] += (double)term;
                    accabs[b * nd + a] += fabs((double)term);
                }
        }
        for (int e = 0; e < nn; e++) {
            F[i * nn + e] = acc[e];
            if (F64) F64[i * nn + e] = acc64[e];
            if (Fabs) Fabs[i * nn + e] = accabs[e];
        }
    }
}

/* ---------------------------------------------------------------------------------------
 * The rest of the TLSPH / WCSPH right-hand side either side of the neighbour sweep
 * (SURVEY.md 8f rank 2).  All of it is TrixiParticles.jl arithmetic (not vendored, compat "0.5",
 * no reference test checks values): PARITY UNPINNED, the formulas below are this repo's
 * definition.
 * ------------------------------------------------------------------------------------- */

/*
 * TrixiParticles.compute_pk1_corrected! (called at benchmarks/smoothed_particle_hydrodynamics.jl:186)
 * per particle, nd x nd column-major matrices:
 *   E = (F^T F - I) / 2,  S = lambda tr(E) I + 2 mu E  (St. Venant-Kirchhoff),  P = F S,
 *   pk1_corrected = P L,   lambda = E nu / ((1 + nu)(1 - 2 nu)),  mu = E / (2 (1 + nu)).
 */
void FN(pno_tlsph_pk1_corrected)(int nd, int64_t n, const REAL *F, const REAL *L, REAL young,
                                 REAL nu, REAL *out)
{
    const int nn = nd * nd;
    const REAL lambda = (young * nu) / (((REAL)1 + nu) * ((REAL)1 - (REAL)2 * nu));
    const REAL mu = young / ((REAL)2 * ((REAL)1 + nu));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        const REAL *Fi = F + i * nn, *Li = L + i * nn;
        REAL E[9], S[9], P[9];
        REAL tr = 0;
        for (int a = 0; a < nd; a++)
            for (int b = 0; b < nd; b++) {
                REAL t = Fi[a * nd + 0] * Fi[b * nd + 0];        /* (F^T F)[a,b] = sum_k F[k,a] F[k,b] */
                for (int k = 1; k < nd; k++) t = t + Fi[a * nd + k] * Fi[b * nd + k];
                E[b * nd + a] = (REAL)0.5 * (t - (a == b ? (REAL)1 : (REAL)0));
            }
        for (int a = 0; a < nd; a++) tr = (a == 0) ? E[0] : tr + E[a * nd + a];
        for (int a = 0; a < nd; a++)
            for (int b = 0; b < nd; b++) {
                REAL t = ((REAL)2 * mu) * E[b * nd + a];
                if (a == b) t = lambda * tr + t;
                S[b * nd + a] = t;
            }
        for (int a = 0; a < nd; a++)
            for (int b = 0; b < nd; b++) {
                REAL t = Fi[0 * nd + a] * S[b * nd + 0];          /* (F S)[a,b] */
                for (int k = 1; k < nd; k++) t = t + Fi[k * nd + a] * S[b * nd + k];
                P[b * nd + a] = t;
            }
        for (int a = 0; a < nd; a++)
            for (int b = 0; b < nd; b++) {
                REAL t = P[0 * nd + a] * Li[b * nd + 0];          /* (P L)[a,b] */
                for (int k = 1; k < nd; k++) t = t + P[k * nd + a] * Li[b * nd + k];
                out[i * nn + b * nd + a] = t;
            }
    }
}

/*
 * TrixiParticles.interact_structure_structure! (benchmarks/smoothed_particle_hydrodynamics.jl:121)
 * over precomputed lists, neighbours / kernel on the INITIAL coordinates X0 (no radius test,
 * nhs_precomputed.jl:221-244):
 *   dv_i += m_j (PK1c_i / rho_i^2 + PK1c_j / rho_j^2) gradW(X_ij)
 *   PenaltyForceGanzenmueller(alpha): with x_ij the CURRENT difference,
 *     eps = (F_i + F_j) X_ij - 2 x_ij,  delta = eps . x_ij / |x_ij|,
 *     f = alpha/2 V_i V_j W(|X_ij|) / |X_ij|^2 E delta x_ij / |x_ij|,  dv_i += f / m_i
 * pairs with |X_ij| < sqrt(eps) are skipped.  params = {h, kernel_norm, young_modulus, alpha}.
 */
void FN(pno_tlsph_interact)(int ndims, REAL r, int periodic, const REAL *box_min,
                            const REAL *box_max, const REAL *X0, const REAL *xcur, int64_t n,
                            const int64_t *offsets, const int32_t *ids, const REAL *mass,
                            const REAL *rho0, const REAL *pk1c, const REAL *F, const REAL *params,
                            REAL *dv, double *dv64, double *dvabs)
{
    GRID g;
    memset(&g, 0, sizeof(g));
    g.ndims = ndims;
    g.periodic = periodic;
    for (int d = 0; d < ndims && periodic; d++) g.box_size[d] = box_max[d] - box_min[d];
    const REAL r2 = r * r;
    const int nd = ndims, nn = ndims * ndims;
    const REAL h = params[0], kernel_norm = params[1], young = params[2], alpha = params[3];
#if REAL_IS_FLOAT
    const REAL sqrt_eps = 3.4526698300124393e-4f;
#else
    const REAL sqrt_eps = 1.4901161193847656e-8;
#endif
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        REAL acc[3] = {0, 0, 0};
        double acc64[3] = {0, 0, 0}, accabs[3] = {0, 0, 0};
        const REAL rho_i2 = rho0[i] * rho0[i];
        const REAL vol_i = mass[i] / rho0[i];
        for (int64_t k = offsets[i]; k < offsets[i + 1]; k++) {
            int64_t j = ids[k];
            REAL p[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) p[d] = X0[i * nd + d] - X0[j * nd + d];
            REAL d2 = p[0] * p[0];
            for (int d = 1; d < nd; d++) d2 = d2 + p[d] * p[d];
            d2 = FN(pno_periodic_fix)(&g, p, d2, r2);
            REAL dist = (REAL)sqrt((double)d2);
            if (dist < sqrt_eps) continue;
            REAL q = dist / h, dw = 0, w = 0;
            if (q < (REAL)2) {
                REAL t = (REAL)1 - q * (REAL)0.5;
                dw = ((REAL)-5 * q) * ((t * t) * t);
                w = ((t * t) * (t * t)) * ((REAL)2 * q + (REAL)1);
            }
            REAL sg = ((kernel_norm / h) * dw) / dist;
            REAL grad[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) grad[d] = sg * p[d];
            const REAL rho_j2 = rho0[j] * rho0[j];
            const REAL m_j = mass[j];
            REAL term[3] = {0, 0, 0}, pen[3] = {0, 0, 0};
            for (int a = 0; a < nd; a++) {
                REAL t = 0;
                for (int b = 0; b < nd; b++) {
                    REAL A = pk1c[i * nn + b * nd + a] / rho_i2 + pk1c[j * nn + b * nd + a] / rho_j2;
                    t = (b == 0) ? A * grad[0] : t + A * grad[b];
                }
                term[a] = m_j * t;
            }
            /* penalty force */
            REAL cp[3] = {0, 0, 0}, es[3] = {0, 0, 0};
            for (int d = 0; d < nd; d++) cp[d] = xcur[i * nd + d] - xcur[j * nd + d];
            REAL c2 = cp[0] * cp[0];
            for (int d = 1; d < nd; d++) c2 = c2 + cp[d] * cp[d];
            REAL cd = (REAL)sqrt((double)c2);
            for (int a = 0; a < nd; a++) {
                REAL t = 0;
                for (int b = 0; b < nd; b++) {
                    REAL Fs = F[i * nn + b * nd + a] + F[j * nn + b * nd + a];
                    t = (b == 0) ? Fs * p[0] : t + Fs * p[b];
                }
                es[a] = t - (REAL)2 * cp[a];
            }
            REAL ds = es[0] * cp[0];
            for (int d = 1; d < nd; d++) ds = ds + es[d] * cp[d];
            ds = ds / cd;
            REAL vol_j = m_j / rho0[j];
            REAL c = ((alpha * (REAL)0.5) * vol_i) * vol_j;
            c = (c * (kernel_norm * w)) / (dist * dist);
            c = ((c * young) * ds) / cd;
            for (int d = 0; d < nd; d++) pen[d] = (c * cp[d]) / mass[i];
            for (int d = 0; d < nd; d++) {
                acc[d] += term[d];
                acc[d] += pen[d];
                acc64[d] += (double)term[d] + (double)pen[d];
                accabs[d] += fabs((double)term[d]) + fabs((double)pen[d]);
            }
        }
        for (int d = 0; d < nd; d++) {
            dv[i * nd + d] = acc[d];
            if (dv64) dv64[i * nd + d] = acc64[d];
            if (dvabs) dvabs[i * nd + d] = accabs[d];
        }
    }
}

/*
 * TrixiParticles.compute_pressure! for ContinuityDensity + StateEquationCole
 * (benchmarks/smoothed_particle_hydrodynamics.jl:64-69,99): density = last row of v,
 *   p = B ((rho / rho0)^gamma - 1) + p_background,  B = rho0 c^2 / gamma.
 */
void FN(pno_wcsph_compute_pressure)(int nd, int64_t n, const REAL *v, REAL sound_speed, REAL rho0,
                                    REAL exponent, REAL background, REAL *pressure)
{
    const REAL B = (rho0 * (sound_speed * sound_speed)) / exponent;
    for (int64_t i = 0; i < n; i++) {
        REAL ratio = v[i * (nd + 1) + nd] / rho0;
        REAL pw = (exponent == (REAL)1) ? ratio : (REAL)pow((double)ratio, (double)exponent);
        pressure[i] = B * (pw - (REAL)1) + background;
    }
}

/* K_ref: candidate tests the reference performs = sum_i sum_{3^d cells} |cell| */
int64_t FN(pno_candidate_tests)(const GRID *g, const int64_t *cell_start, const REAL *x, int64_t nx)
{
    int64_t total = 0;
    const int nd = g->ndims;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int64_t i = 0; i < nx; i++) {
        int64_t cell[3];
        FN(pno_cell_coords)(g, x + i * nd, cell);
        int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (int d = 0; d < nd; d++) { lo[d] = -1; hi[d] = 1; }
        for (int o3 = lo[2]; o3 <= hi[2]; o3++)
            for (int o2 = lo[1]; o2 <= hi[1]; o2++)
                for (int o1 = lo[0]; o1 <= hi[0]; o1++) {
                    int64_t nc[3] = {cell[0] + o1, cell[1] + o2, cell[2] + o3};
                    FN(pno_periodic_cell)(g, nc);
                    int ok = 1;
                    for (int d = 0; d < nd; d++)
                        if (nc[d] < 1 || nc[d] > g->grid_size[d]) ok = 0;
                    if (!ok) continue;
                    int64_t c = FN(pno_linear)(g, nc);
                    total += cell_start[c + 1] - cell_start[c];
                }
    }
    return total;
}

/* neighborhood_search.jl:439-449 periodic_coords (used by the face-rounding KAT) */
void FN(pno_periodic_coords)(int ndims, const REAL *box_min, const REAL *box_max, const REAL *x,
                             REAL *out)
{
    for (int d = 0; d < ndims; d++) {
        REAL size = box_max[d] - box_min[d];
        REAL off = (REAL)floor((double)((x[d] - box_min[d]) / size));
        REAL c = x[d] - off * size;
        c = (c > box_min[d]) ? c : box_min[d]; /* max.(c, min_corner) */
        c = (c < box_max[d]) ? c : box_max[d]; /* min.(.., max_corner) */
        out[d] = c;
    }
}


/* ---------------------------------------------------------------------------------------
 * SpatialHashingCellList (cell_lists/spatial_hashing.jl) behind a GridNeighborhoodSearch
 * (SURVEY.md 8f rank 3).  The table has list_size entries (keys are 0-based here); per key:
 * the list of point ids, the cell coordinates stored by the first insertion (`coords`, the
 * reference flattens them into a UInt128 with 0 = "unused", spatial_hashing.jl:56,176-183) and
 * the collision flag.  Serial insertion order (push_cell!, :79-97) = ascending ids: the
 * reference's ParallelUpdate does the same with atomics (:99-118), where the stored cell of a
 * colliding key depends on the arrival order.
 * ------------------------------------------------------------------------------------- */

/* GridNeighborhoodSearch ctor (nhs_grid.jl:100-126) for a cell list without corners */
int FN(pno_hash_grid_init)(GRID *g, int ndims, REAL r, int periodic, const REAL *box_min,
                           const REAL *box_max)
{
    REAL zero[3] = {0, 0, 0};
    int rc = FN(pno_grid_init)(g, ndims, r, zero, zero, periodic, box_min, box_max);
    for (int d = 0; d < 3; d++) { g->min_corner[d] = 0; g->max_corner[d] = 0; g->grid_size[d] = 0; }
    return rc;
}

/* spatial_hash (spatial_hashing.jl:159-174), Julia Int64 products wrap; returned 0-based */
int64_t FN(pno_spatial_hash)(int ndims, const int64_t *cell, int64_t list_size)
{
    uint64_t h = (uint64_t)cell[0] * 73856093ULL;
    if (ndims > 1) h ^= (uint64_t)cell[1] * 19349663ULL;
    if (ndims > 2) h ^= (uint64_t)cell[2] * 83492791ULL;
    return FN(pno_floormod)((int64_t)h, list_size);
}

/* cell_coords (nhs_grid.jl:622-638): floor_to_int.(coords ./ cell_size), then the periodic wrap */
void FN(pno_hash_cell_coords)(const GRID *g, const REAL *x, int64_t *cell)
{
    for (int d = 0; d < g->ndims; d++) cell[d] = FN(pno_floor_to_int)(x[d] / g->cell_size[d]);
    for (int d = g->ndims; d < 3; d++) cell[d] = 0;
    FN(pno_periodic_cell)(g, cell);
}

/*
 * initialize_grid! (nhs_grid.jl:255-281) with push_cell! (spatial_hashing.jl:79-97), points
 * visited in the order of idx (NULL = 0..n-1).  key_start[L+1] / key_points[n_idx] = the lists in
 * insertion order, coords[3 L] (0 for unused dims), collisions[L].
 * returns 0 ok, 5 = a cell coordinate does not fit Int32 (InexactError, :176-183).
 */
int FN(pno_hash_build)(const GRID *g, int64_t list_size, const REAL *y, int64_t n,
                       const int64_t *idx, int64_t n_idx, int64_t *key_start, int32_t *key_points,
                       int32_t *coords, uint8_t *collisions)
{
    const int64_t L = list_size;
    if (idx == NULL) n_idx = n;
    for (int64_t k = 0; k <= L; k++) key_start[k] = 0;
    memset(coords, 0, sizeof(int32_t) * 3 * (size_t)L);
    memset(collisions, 0, (size_t)L);
    if ((double)g->search_radius < 2.220446049250313e-16) return 0;
    int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_idx > 0 ? n_idx : 1));
    for (int64_t k = 0; k < n_idx; k++) {
        int64_t p = idx ? idx[k] : k, cell[3];
        FN(pno_hash_cell_coords)(g, y + p * g->ndims, cell);
        for (int d = 0; d < g->ndims; d++)
            if (cell[d] > INT32_MAX || cell[d] < INT32_MIN) { free(key); return 5; }
        key[k] = FN(pno_spatial_hash)(g->ndims, cell, L);
        key_start[key[k] + 1]++;
        int32_t *cc = coords + 3 * key[k];
        if (cc[0] == 0 && cc[1] == 0 && cc[2] == 0) {         /* typemin(UInt128): unused */
            for (int d = 0; d < 3; d++) cc[d] = (int32_t)cell[d];
        } else if (cc[0] != cell[0] || cc[1] != cell[1] || cc[2] != cell[2]) {
            collisions[key[k]] = 1;
        }
    }
    for (int64_t k = 0; k < L; k++) key_start[k + 1] += key_start[k];
    int64_t *cursor = (int64_t *)malloc(sizeof(int64_t) * (size_t)(L > 0 ? L : 1));
    memcpy(cursor, key_start, sizeof(int64_t) * (size_t)L);
    for (int64_t k = 0; k < n_idx; k++) key_points[cursor[key[k]]++] = (int32_t)(idx ? idx[k] : k);
    free(cursor);
    free(key);
    return 0;
}

/* mapreduce_neighbor_inner (nhs_grid.jl:519-575) with the hashing hooks check_cell_collision
 * (:501-513) and check_collision (:486-492) for one point */
static inline void
FN(pno_hash_sweep_point)(const GRID *g, int64_t L, const int64_t *key_start,
                         const int32_t *key_points, const int32_t *coords,
                         const uint8_t *collisions, const REAL *xi, int64_t i, const REAL *y,
                         FN(pno_pair_fn) f, void *ctx)
{
    const int nd = g->ndims;
    const REAL r = g->search_radius;
    const REAL r2 = r * r;
    int64_t cell[3];
    FN(pno_hash_cell_coords)(g, xi, cell);
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < nd; d++) { lo[d] = -1; hi[d] = 1; }
    for (int o3 = lo[2]; o3 <= hi[2]; o3++)
        for (int o2 = lo[1]; o2 <= hi[1]; o2++)
            for (int o1 = lo[0]; o1 <= hi[0]; o1++) {
                int64_t nc[3] = {FN(pno_wrap_add)(cell[0], o1), FN(pno_wrap_add)(cell[1], o2),
                                 FN(pno_wrap_add)(cell[2], o3)};
                FN(pno_periodic_cell)(g, nc);
                const int64_t key = FN(pno_spatial_hash)(nd, nc, L);
                const int32_t *cc = coords + 3 * key;
                const int cell_collision = collisions[key] || cc[0] != (int32_t)nc[0] ||
                                           cc[1] != (int32_t)nc[1] || cc[2] != (int32_t)nc[2];
                for (int64_t k = key_start[key]; k < key_start[key + 1]; k++) {
                    int64_t j = key_points[k];
                    const REAL *yj = y + j * nd;
                    REAL p[3] = {0, 0, 0};
                    for (int d = 0; d < nd; d++) p[d] = xi[d] - yj[d];
                    REAL d2 = p[0] * p[0];
                    for (int d = 1; d < nd; d++) d2 = d2 + p[d] * p[d];
                    d2 = FN(pno_periodic_fix)(g, p, d2, r2);
                    if (d2 <= r2) {
                        if (cell_collision) {
                            int64_t jc[3];
                            FN(pno_hash_cell_coords)(g, yj, jc);
                            if (jc[0] != nc[0] || jc[1] != nc[1] || jc[2] != nc[2]) continue;
                        }
                        f(ctx, i, j, p, (REAL)sqrt((double)d2));
                    }
                }
            }
}

static void FN(pno_hash_foreach)(const GRID *g, int64_t L, const int64_t *key_start,
                                 const int32_t *key_points, const int32_t *coords,
                                 const uint8_t *collisions, const REAL *x, int64_t nx,
                                 const REAL *y, const int64_t *points, int64_t npoints,
                                 FN(pno_pair_fn) f, void *ctx)
{
    if (points == NULL) npoints = nx;
    if ((double)g->search_radius < 2.220446049250313e-16) return;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < npoints; k++) {
        int64_t i = points ? points[k] : k;
        FN(pno_hash_sweep_point)(g, L, key_start, key_points, coords, collisions,
                                 x + i * g->ndims, i, y, f, ctx);
    }
}

/* foreach_point_neighbor with the count closure (count_neighbors.jl:24-27) on the hashed list */
void FN(pno_hash_count_neighbors)(const GRID *g, int64_t L, const int64_t *key_start,
                                  const int32_t *key_points, const int32_t *coords,
                                  const uint8_t *collisions, const REAL *x, int64_t nx,
                                  const REAL *y, const int64_t *points, int64_t npoints,
                                  int64_t *out)
{
    for (int64_t i = 0; i < nx; i++) out[i] = 0;
    FN(pno_hash_foreach)(g, L, key_start, key_points, coords, collisions, x, nx, y, points,
                         npoints, FN(pno_cl_count), out);
}

/* neighbour lists in sweep order (two passes like pno_neighbor_lists), optionally sorted */
void FN(pno_hash_neighbor_lists)(const GRID *g, int64_t L, const int64_t *key_start,
                                 const int32_t *key_points, const int32_t *coords,
                                 const uint8_t *collisions, const REAL *x, int64_t nx,
                                 const REAL *y, int64_t *offsets, int32_t *ids, int sort)
{
    int64_t *counts = (int64_t *)calloc((size_t)(nx > 0 ? nx : 1), sizeof(int64_t));
    FN(pno_list_ctx) c = {counts, offsets, ids};
    if (ids == NULL) {
        FN(pno_hash_foreach)(g, L, key_start, key_points, coords, collisions, x, nx, y, NULL, nx,
                             FN(pno_cl_list_count), &c);
        offsets[0] = 0;
        for (int64_t i = 0; i < nx; i++) offsets[i + 1] = offsets[i] + counts[i];
    } else {
        FN(pno_hash_foreach)(g, L, key_start, key_points, coords, collisions, x, nx, y, NULL, nx,
                             FN(pno_cl_list_fill), &c);
        if (sort) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < nx; i++)
                qsort(ids + offsets[i], (size_t)(offsets[i + 1] - offsets[i]), sizeof(int32_t),
                      FN(pno_cmp_i32));
        }
    }
    free(counts);
}

/* n-body / WCSPH closures over the hashed list (same closures as the full grid) */
void FN(pno_hash_nbody)(const GRID *g, int64_t L, const int64_t *key_start,
                        const int32_t *key_points, const int32_t *coords,
                        const uint8_t *collisions, const REAL *x, int64_t nx, const REAL *y,
                        const REAL *mass, REAL G, REAL *dv, double *dv64, double *dvabs)
{
    const int nd = g->ndims;
    for (int64_t k = 0; k < nx * nd; k++) {
        dv[k] = 0;
        if (dv64) dv64[k] = 0.0;
        if (dvabs) dvabs[k] = 0.0;
    }
    FN(pno_nbody_ctx) c = {nd, mass, G, dv, dv64, dvabs};
    FN(pno_hash_foreach)(g, L, key_start, key_points, coords, collisions, x, nx, y, NULL, nx,
                         FN(pno_cl_nbody), &c);
}

#undef PNO_VIEW
#undef GRID
#undef FN
#undef PNO_CAT
#undef PNO_CAT_
