/* szo.c — CPU ORACLE for the Subzero.jl floe-interaction hot path.  TEST INFRASTRUCTURE.
 *
 * A plain-C, FP64, scalar restatement of the reference algorithm, function by function,
 * with the reference file:line each block follows.  It exists to CHECK the CUDA product
 * (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference) and is never
 * linked, imported or executed by the product path.
 *
 * Parity status: the reference is 100 % Julia and cannot run in this environment (no julia
 * binary), and its polygon clipping lives in GeometryOps.jl 0.1.x which is not vendored.
 * The oracle is pinned against every golden value the reference's own tests hold for this
 * path (tests/test_reference_golden.py: test_collisions.jl:50-150,190-363,
 * test_coupling.jl:464-640, test_update_floe.jl:2-42).  Below the digits those tests print,
 * and for exactly degenerate polygon configurations, parity is UNPINNED (see szo_geom.h).
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).  The OpenMP
 * loops mirror the reference's Threads.@threads sites (collisions.jl:745,
 * update_floe.jl:475); the coupling loop is serial in the reference (coupling.jl:1498) but
 * is threaded here over floes for the CPU-baseline timing (results are per-floe, so the
 * thread count does not change them).
 */
#define SZ_ORACLE_BUILD 1
#include "../include/subzero_b200.h"
#include "szo_geom.h"
#include "szo_points.h"

#include <stdio.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NROWF 7 /* floeidx xforce yforce xpoint ypoint torque overlap, floe.jl:102-110 */

typedef struct {
    double *r; /* n x 7 */
    int n, cap;
} rowlist;

typedef struct {
    int64_t *v;
    int n, cap;
} intlist;

typedef struct cellrec {
    int ix, iy;          /* 0-based cell (grid line) indices after the periodic shift */
    int64_t floe;
    double tx, ty, dx, dy;
    int64_t npts;
} cellrec;

struct sz_handle {
    sz_config cfg;
    char err[512];
    /* grid + fields */
    int Nx, Ny;
    double x0, xf, y0, yf, dx, dy;
    double *ocn_u, *ocn_v, *ocn_hflx, *atm_u, *atm_v;
    /* domain */
    int have_domain;
    int32_t kind[4];
    double val[4], wu[4], wv[4], rect[4][4];
    szo_pt wall_ring[4][5];
    int n_topo;
    szo_pt **topo_ring;
    int *topo_np;
    double *topo_cx, *topo_cy, *topo_rmax;
    /* floes */
    int64_t n, n_init, cap;
    double *cx, *cy, *height, *area, *mass, *rmax, *moment, *alpha, *u, *v, *xi;
    double *fxOA, *fyOA, *trqOA, *hflx, *overarea, *cfx, *cfy, *ctrq;
    double *p_dxdt, *p_dydt, *p_dudt, *p_dvdt, *p_dxidt, *p_dalphadt;
    double *stress_accum, *stress_instant, *strain; /* 4 per floe */
    int32_t *status;
    int64_t *id, *ghost_id;
    intlist *ghosts;   /* 0-based indices */
    intlist *fuse_idx; /* 0-based */
    szo_pt **ring;
    int *npts;
    double **mcx, **mcy;
    int *nmc;
    rowlist *rows;
    uint32_t *warn;
    /* last-step pair lists (0-based) */
    int64_t *cand, *pairs, *overlap, *fuse;
    int64_t n_cand, n_pairs, n_overlap, n_fuse, n_domain_pairs, n_clip_fail;
    double ms[8];
    /* two-way coupling */
    double *ocn_temp, *atm_temp, *taux, *tauy, *sifrac;
    struct cellrec *crec;      /* registry sorted by (cell, floe) after step_coupling */
    int64_t n_crec;
    struct cellrec **frec;     /* per-floe records of the running coupling step */
    int *nfrec;
    /* halo lists (slab decomposition) */
    int n_lists;
    int64_t *hl_off, *hl_idx;
};

#define DFIELDS(X)                                                                               \
    X(cx) X(cy) X(height) X(area) X(mass) X(rmax) X(moment) X(alpha) X(u) X(v) X(xi) X(fxOA)     \
    X(fyOA) X(trqOA) X(hflx) X(overarea) X(cfx) X(cfy) X(ctrq) X(p_dxdt) X(p_dydt) X(p_dudt)     \
    X(p_dvdt) X(p_dxidt) X(p_dalphadt)

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int32_t fail(sz_handle *h, int32_t code, const char *msg) {
    if (h) snprintf(h->err, sizeof(h->err), "%s", msg);
    return code;
}

/* ---- config ----------------------------------------------------------------------- */
void szo_default_config(sz_config *c) {
    memset(c, 0, sizeof(*c));
    /* Constants(), simulation.jl:5-18 */
    c->rho_o = 1027.0; c->rho_a = 1.2; c->Cd_io = 3e-3; c->Cd_ia = 1e-3; c->Cd_ao = 1.25e-3;
    c->f = 1.4e-4; c->turn_theta = 15.0 * M_PI / 180.0; c->L = 2.93e5; c->k = 2.14;
    c->nu = 0.3; c->mu = 0.2; c->E = 6e6;
    /* CollisionSettings(), process_settings.jl:183-187 */
    c->floe_floe_max_overlap = 0.55; c->floe_domain_max_overlap = 0.75;
    /* FloeSettings(), process_settings.jl:20-32; DecayAreaScaledCalculator λ, stress_calculators.jl:82 */
    c->rho_i = 920.0; c->max_floe_height = 10.0; c->maximum_xi = 1e-5; c->stress_lambda = 0.2;
    c->coupling_dd = 1; c->two_way_coupling_on = 0; c->dt = 10; c->device = 0;
    c->max_regions_per_pair = 4; c->max_pairs_per_floe = 24;
}

const char *szo_version(void) { return "subzero-b200 oracle 0.1 (CPU restatement, test infrastructure)"; }
const char *szo_last_error(sz_handle *h) { return h ? h->err : "null handle"; }

int32_t szo_create(const sz_config *cfg, sz_handle **out) {
    if (!cfg || !out) return SZ_ERR_INVALID;
    sz_handle *h = (sz_handle *)calloc(1, sizeof(sz_handle));
    if (!h) return SZ_ERR_NOMEM;
    h->cfg = *cfg;
#ifdef _OPENMP
    if (cfg->threads > 0) omp_set_num_threads(cfg->threads);
#endif
    *out = h;
    return SZ_OK;
}

static void free_floe_slot(sz_handle *h, int64_t i) {
    free(h->ring[i]); h->ring[i] = NULL;
    free(h->mcx[i]); free(h->mcy[i]); h->mcx[i] = h->mcy[i] = NULL;
    free(h->rows[i].r); memset(&h->rows[i], 0, sizeof(rowlist));
    free(h->ghosts[i].v); memset(&h->ghosts[i], 0, sizeof(intlist));
    free(h->fuse_idx[i].v); memset(&h->fuse_idx[i], 0, sizeof(intlist));
}

static void free_floes(sz_handle *h) {
    for (int64_t i = 0; i < h->n; ++i) free_floe_slot(h, i);
#define X(f) free(h->f); h->f = NULL;
    DFIELDS(X)
#undef X
    free(h->stress_accum); free(h->stress_instant); free(h->strain);
    free(h->status); free(h->id); free(h->ghost_id); free(h->ghosts); free(h->fuse_idx);
    free(h->ring); free(h->npts); free(h->mcx); free(h->mcy); free(h->nmc); free(h->rows);
    free(h->warn);
    h->stress_accum = h->stress_instant = h->strain = NULL;
    h->status = NULL; h->id = h->ghost_id = NULL; h->ghosts = h->fuse_idx = NULL;
    h->ring = NULL; h->npts = NULL; h->mcx = h->mcy = NULL; h->nmc = NULL; h->rows = NULL;
    h->warn = NULL;
    h->n = h->n_init = h->cap = 0;
}

void szo_destroy(sz_handle *h) {
    if (!h) return;
    free_floes(h);
    free(h->ocn_u); free(h->ocn_v); free(h->ocn_hflx); free(h->atm_u); free(h->atm_v);
    free(h->ocn_temp); free(h->atm_temp); free(h->taux); free(h->tauy); free(h->sifrac); free(h->crec);
    for (int k = 0; k < h->n_topo; ++k) free(h->topo_ring[k]);
    free(h->topo_ring); free(h->topo_np); free(h->topo_cx); free(h->topo_cy); free(h->topo_rmax);
    free(h->cand); free(h->pairs); free(h->overlap); free(h->fuse);
    free(h->hl_off); free(h->hl_idx);
    free(h);
}

static void reserve_floes(sz_handle *h, int64_t want) {
    if (want <= h->cap) return;
    int64_t nc = h->cap ? h->cap : 16;
    while (nc < want) nc *= 2;
#define X(f) h->f = (double *)realloc(h->f, sizeof(double) * (size_t)nc);
    DFIELDS(X)
#undef X
    h->stress_accum = (double *)realloc(h->stress_accum, sizeof(double) * 4 * (size_t)nc);
    h->stress_instant = (double *)realloc(h->stress_instant, sizeof(double) * 4 * (size_t)nc);
    h->strain = (double *)realloc(h->strain, sizeof(double) * 4 * (size_t)nc);
    h->status = (int32_t *)realloc(h->status, sizeof(int32_t) * (size_t)nc);
    h->id = (int64_t *)realloc(h->id, sizeof(int64_t) * (size_t)nc);
    h->ghost_id = (int64_t *)realloc(h->ghost_id, sizeof(int64_t) * (size_t)nc);
    h->warn = (uint32_t *)realloc(h->warn, sizeof(uint32_t) * (size_t)nc);
#define G(f, T)                                                                \
    h->f = (T *)realloc(h->f, sizeof(T) * (size_t)nc);                         \
    memset(h->f + h->cap, 0, sizeof(T) * (size_t)(nc - h->cap));
    G(ghosts, intlist) G(fuse_idx, intlist) G(ring, szo_pt *) G(npts, int) G(mcx, double *)
    G(mcy, double *) G(nmc, int) G(rows, rowlist)
#undef G
    h->cap = nc;
}

static void intlist_push(intlist *l, int64_t v) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 4;
        l->v = (int64_t *)realloc(l->v, sizeof(int64_t) * (size_t)l->cap);
    }
    l->v[l->n++] = v;
}

/* ---- model description ---------------------------------------------------------------- */
int32_t szo_set_grid(sz_handle *h, int32_t Nx, int32_t Ny, double x0, double xf, double y0, double yf) {
    if (!h || Nx < 1 || Ny < 1 || !(xf > x0) || !(yf > y0)) return fail(h, SZ_ERR_INVALID, "set_grid: bad extent");
    h->Nx = Nx; h->Ny = Ny; h->x0 = x0; h->xf = xf; h->y0 = y0; h->yf = yf;
    /* grids.jl:180-211: Δx = (xf - x0)/Nx */
    h->dx = (xf - x0) / Nx; h->dy = (yf - y0) / Ny;
    return SZ_OK;
}

static double *dupfield(const double *src, size_t n) {
    double *d = (double *)malloc(sizeof(double) * n);
    if (src) memcpy(d, src, sizeof(double) * n);
    else memset(d, 0, sizeof(double) * n);
    return d;
}

int32_t szo_set_fields(sz_handle *h, const double *ou, const double *ov, const double *oh,
                       const double *au, const double *av) {
    if (!h || h->Nx == 0) return fail(h, SZ_ERR_INVALID, "set_fields before set_grid");
    size_t n = (size_t)(h->Nx + 1) * (size_t)(h->Ny + 1);
    free(h->ocn_u); free(h->ocn_v); free(h->ocn_hflx); free(h->atm_u); free(h->atm_v);
    h->ocn_u = dupfield(ou, n); h->ocn_v = dupfield(ov, n); h->ocn_hflx = dupfield(oh, n);
    h->atm_u = dupfield(au, n); h->atm_v = dupfield(av, n);
    free(h->ocn_temp); free(h->atm_temp); free(h->taux); free(h->tauy); free(h->sifrac);
    h->ocn_temp = dupfield(NULL, n); h->atm_temp = dupfield(NULL, n);
    h->taux = dupfield(NULL, n); h->tauy = dupfield(NULL, n); h->sifrac = dupfield(NULL, n);
    return SZ_OK;
}

int32_t szo_set_temperatures(sz_handle *h, const double *ot, const double *at) {
    if (!h || !h->ocn_u) return fail(h, SZ_ERR_INVALID, "set_temperatures before set_fields");
    size_t n = (size_t)(h->Nx + 1) * (size_t)(h->Ny + 1);
    free(h->ocn_temp); free(h->atm_temp);
    h->ocn_temp = dupfield(ot, n); h->atm_temp = dupfield(at, n);
    return SZ_OK;
}

int32_t szo_get_ocean_fields(sz_handle *h, double *tx, double *ty, double *si, double *hf) {
    if (!h || !h->ocn_u) return fail(h, SZ_ERR_INVALID, "get_ocean_fields before set_fields");
    size_t n = (size_t)(h->Nx + 1) * (size_t)(h->Ny + 1);
    if (tx) memcpy(tx, h->taux, sizeof(double) * n);
    if (ty) memcpy(ty, h->tauy, sizeof(double) * n);
    if (si) memcpy(si, h->sifrac, sizeof(double) * n);
    if (hf) memcpy(hf, h->ocn_hflx, sizeof(double) * n);
    return SZ_OK;
}

int32_t szo_get_cell_floes(sz_handle *h, int64_t *n, int64_t *cell_xy, int64_t *floe, double *vals) {
    if (!h || !n) return SZ_ERR_INVALID;
    *n = h->n_crec;
    for (int64_t k = 0; k < h->n_crec && cell_xy && floe && vals; ++k) {
        const cellrec *r = &h->crec[k];
        cell_xy[2 * k] = r->ix + 1; cell_xy[2 * k + 1] = r->iy + 1;
        floe[k] = r->floe + 1;
        vals[5 * k] = r->tx; vals[5 * k + 1] = r->ty; vals[5 * k + 2] = (double)r->npts;
        vals[5 * k + 3] = r->dx; vals[5 * k + 4] = r->dy;
    }
    return SZ_OK;
}

/* _make_bounding_box_polygon, floe_utils.jl:104-108: (xmin,ymin),(xmin,ymax),(xmax,ymax),(xmax,ymin) */
static void make_wall_ring(szo_pt r[5], const double rect[4]) {
    double xmin = rect[0], xmax = rect[1], ymin = rect[2], ymax = rect[3];
    r[0] = (szo_pt){xmin, ymin}; r[1] = (szo_pt){xmin, ymax}; r[2] = (szo_pt){xmax, ymax};
    r[3] = (szo_pt){xmax, ymin}; r[4] = r[0];
}

int32_t szo_set_domain(sz_handle *h, const int32_t kinds[4], const double vals[4], const double uv[8],
                       const double rect[16], int32_t n_topo, const int64_t *toff, const double *txy,
                       const double *tcent, const double *trmax) {
    if (!h || !kinds || !vals || !rect) return fail(h, SZ_ERR_INVALID, "set_domain: null argument");
    /* domains.jl:11-33: periodic walls must come in opposite pairs */
    if ((kinds[0] == SZ_BOUNDARY_PERIODIC) != (kinds[1] == SZ_BOUNDARY_PERIODIC) ||
        (kinds[2] == SZ_BOUNDARY_PERIODIC) != (kinds[3] == SZ_BOUNDARY_PERIODIC))
        return fail(h, SZ_ERR_INVALID, "set_domain: periodic boundaries must be paired");
    if (!(vals[0] > vals[1]) || !(vals[2] > vals[3]))
        return fail(h, SZ_ERR_INVALID, "set_domain: north <= south or east <= west");
    for (int w = 0; w < 4; ++w) {
        h->kind[w] = kinds[w]; h->val[w] = vals[w];
        h->wu[w] = uv ? uv[2 * w] : 0.0; h->wv[w] = uv ? uv[2 * w + 1] : 0.0;
        memcpy(h->rect[w], rect + 4 * w, sizeof(double) * 4);
        make_wall_ring(h->wall_ring[w], h->rect[w]);
    }
    for (int k = 0; k < h->n_topo; ++k) free(h->topo_ring[k]);
    free(h->topo_ring); free(h->topo_np); free(h->topo_cx); free(h->topo_cy); free(h->topo_rmax);
    h->n_topo = n_topo;
    h->topo_ring = (szo_pt **)calloc((size_t)(n_topo > 0 ? n_topo : 1), sizeof(szo_pt *));
    h->topo_np = (int *)calloc((size_t)(n_topo > 0 ? n_topo : 1), sizeof(int));
    h->topo_cx = (double *)calloc((size_t)(n_topo > 0 ? n_topo : 1), sizeof(double));
    h->topo_cy = (double *)calloc((size_t)(n_topo > 0 ? n_topo : 1), sizeof(double));
    h->topo_rmax = (double *)calloc((size_t)(n_topo > 0 ? n_topo : 1), sizeof(double));
    for (int k = 0; k < n_topo; ++k) {
        int np = (int)(toff[k + 1] - toff[k]);
        h->topo_np[k] = np;
        h->topo_ring[k] = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)np);
        memcpy(h->topo_ring[k], txy + 2 * toff[k], sizeof(szo_pt) * (size_t)np);
        h->topo_cx[k] = tcent[2 * k]; h->topo_cy[k] = tcent[2 * k + 1]; h->topo_rmax[k] = trmax[k];
    }
    h->have_domain = 1;
    return SZ_OK;
}

int32_t szo_get_domain(sz_handle *h, double vals[4], double rect[16]) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "get_domain before set_domain");
    for (int w = 0; w < 4; ++w) {
        vals[w] = h->val[w];
        memcpy(rect + 4 * w, h->rect[w], sizeof(double) * 4);
    }
    return SZ_OK;
}

/* ---- floe state ------------------------------------------------------------------------ */
int32_t szo_upload_floes(sz_handle *h, const sz_floe_soa *s) {
    if (!h || !s || s->n < 0 || s->n_init < 0 || s->n_init > s->n) return fail(h, SZ_ERR_INVALID, "upload_floes: bad sizes");
    if (!s->centroid_x || !s->centroid_y || !s->area || !s->rmax || !s->vert_offsets || !s->vert_xy)
        return fail(h, SZ_ERR_INVALID, "upload_floes: geometry arrays are required");
    free_floes(h);
    reserve_floes(h, s->n + 8);
    h->n = s->n; h->n_init = s->n_init;
#define CP(dst, src) for (int64_t i = 0; i < s->n; ++i) h->dst[i] = s->src ? s->src[i] : 0.0;
    CP(cx, centroid_x) CP(cy, centroid_y) CP(height, height) CP(area, area) CP(mass, mass)
    CP(rmax, rmax) CP(moment, moment) CP(alpha, alpha) CP(u, u) CP(v, v) CP(xi, xi) CP(fxOA, fxOA)
    CP(fyOA, fyOA) CP(trqOA, trqOA) CP(hflx, hflx_factor) CP(overarea, overarea)
    CP(ctrq, collision_trq) CP(p_dxdt, p_dxdt) CP(p_dydt, p_dydt) CP(p_dudt, p_dudt)
    CP(p_dvdt, p_dvdt) CP(p_dxidt, p_dxidt) CP(p_dalphadt, p_dalphadt)
#undef CP
    for (int64_t i = 0; i < s->n; ++i) {
        h->cfx[i] = s->collision_force ? s->collision_force[2 * i] : 0.0;
        h->cfy[i] = s->collision_force ? s->collision_force[2 * i + 1] : 0.0;
        for (int k = 0; k < 4; ++k) {
            h->stress_accum[4 * i + k] = s->stress_accum ? s->stress_accum[4 * i + k] : 0.0;
            h->stress_instant[4 * i + k] = s->stress_instant ? s->stress_instant[4 * i + k] : 0.0;
            h->strain[4 * i + k] = s->strain ? s->strain[4 * i + k] : 0.0;
        }
        h->status[i] = s->status_tag ? s->status_tag[i] : SZ_STATUS_ACTIVE;
        h->id[i] = s->id ? s->id[i] : i + 1;
        h->ghost_id[i] = s->ghost_id ? s->ghost_id[i] : 0;
        h->warn[i] = 0;
        int np = (int)(s->vert_offsets[i + 1] - s->vert_offsets[i]);
        if (np < 4) return fail(h, SZ_ERR_INVALID, "upload_floes: a ring needs >= 4 points (closed)");
        h->npts[i] = np;
        h->ring[i] = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)np);
        memcpy(h->ring[i], s->vert_xy + 2 * s->vert_offsets[i], sizeof(szo_pt) * (size_t)np);
        if (h->ring[i][0].x != h->ring[i][np - 1].x || h->ring[i][0].y != h->ring[i][np - 1].y)
            return fail(h, SZ_ERR_INVALID, "upload_floes: rings must be closed");
        int nm = s->mc_offsets ? (int)(s->mc_offsets[i + 1] - s->mc_offsets[i]) : 0;
        h->nmc[i] = nm;
        if (nm > 0) {
            h->mcx[i] = (double *)malloc(sizeof(double) * (size_t)nm);
            h->mcy[i] = (double *)malloc(sizeof(double) * (size_t)nm);
            memcpy(h->mcx[i], s->mc_x + s->mc_offsets[i], sizeof(double) * (size_t)nm);
            memcpy(h->mcy[i], s->mc_y + s->mc_offsets[i], sizeof(double) * (size_t)nm);
        }
        if (s->ghost_offsets)
            for (int64_t g = s->ghost_offsets[i]; g < s->ghost_offsets[i + 1]; ++g)
                intlist_push(&h->ghosts[i], s->ghost_index[g] - 1);
    }
    h->n_cand = h->n_pairs = h->n_overlap = h->n_fuse = 0;
    return SZ_OK;
}

int32_t szo_upload_state(sz_handle *h, const sz_floe_soa *s) {
    if (!h || !s) return SZ_ERR_INVALID;
    if (s->n != h->n || s->n_init != h->n_init || h->n != h->n_init) return fail(h, SZ_ERR_INVALID, "upload_state: floe list differs from the resident store");
#define CP(dst, src) for (int64_t i = 0; i < s->n; ++i) h->dst[i] = s->src ? s->src[i] : 0.0;
    CP(cx, centroid_x) CP(cy, centroid_y) CP(height, height) CP(area, area) CP(mass, mass)
    CP(rmax, rmax) CP(moment, moment) CP(alpha, alpha) CP(u, u) CP(v, v) CP(xi, xi) CP(fxOA, fxOA)
    CP(fyOA, fyOA) CP(trqOA, trqOA) CP(hflx, hflx_factor) CP(overarea, overarea)
    CP(ctrq, collision_trq) CP(p_dxdt, p_dxdt) CP(p_dydt, p_dydt) CP(p_dudt, p_dudt)
    CP(p_dvdt, p_dvdt) CP(p_dxidt, p_dxidt) CP(p_dalphadt, p_dalphadt)
#undef CP
    int64_t vo = 0;
    for (int64_t i = 0; i < s->n; ++i) {
        h->cfx[i] = s->collision_force ? s->collision_force[2 * i] : 0.0;
        h->cfy[i] = s->collision_force ? s->collision_force[2 * i + 1] : 0.0;
        for (int k = 0; k < 4; ++k) {
            h->stress_accum[4 * i + k] = s->stress_accum ? s->stress_accum[4 * i + k] : 0.0;
            h->stress_instant[4 * i + k] = s->stress_instant ? s->stress_instant[4 * i + k] : 0.0;
            h->strain[4 * i + k] = s->strain ? s->strain[4 * i + k] : 0.0;
        }
        if (s->status_tag) h->status[i] = s->status_tag[i];
        memcpy(h->ring[i], s->vert_xy + 2 * vo, sizeof(szo_pt) * (size_t)h->npts[i]);
        vo += h->npts[i];
    }
    return SZ_OK;
}

int32_t szo_get_counts(sz_handle *h, sz_counts *c) {
    if (!h || !c) return SZ_ERR_INVALID;
    memset(c, 0, sizeof(*c));
    c->n_init = h->n_init; c->n_total = h->n;
    for (int64_t i = 0; i < h->n; ++i) {
        c->n_vertices += h->npts[i];
        c->n_ghost_links += h->ghosts[i].n;
        c->n_rows += h->rows[i].n;
    }
    for (int64_t i = 0; i < h->n_init; ++i) c->n_mc += h->nmc[i];
    c->n_candidates = h->n_cand; c->n_pairs = h->n_pairs; c->n_overlap = h->n_overlap;
    c->n_fuse = h->n_fuse; c->n_domain_pairs = h->n_domain_pairs; c->n_clip_fail = h->n_clip_fail;
    return SZ_OK;
}

int32_t szo_download_floes(sz_handle *h, sz_floe_soa *s) {
    if (!h || !s) return SZ_ERR_INVALID;
    s->n = h->n; s->n_init = h->n_init;
#define CP(dst, src) if (s->dst) for (int64_t i = 0; i < h->n; ++i) s->dst[i] = h->src[i];
    CP(centroid_x, cx) CP(centroid_y, cy) CP(height, height) CP(area, area) CP(mass, mass)
    CP(rmax, rmax) CP(moment, moment) CP(alpha, alpha) CP(u, u) CP(v, v) CP(xi, xi) CP(fxOA, fxOA)
    CP(fyOA, fyOA) CP(trqOA, trqOA) CP(hflx_factor, hflx) CP(overarea, overarea)
    CP(collision_trq, ctrq) CP(p_dxdt, p_dxdt) CP(p_dydt, p_dydt) CP(p_dudt, p_dudt)
    CP(p_dvdt, p_dvdt) CP(p_dxidt, p_dxidt) CP(p_dalphadt, p_dalphadt) CP(status_tag, status)
    CP(id, id) CP(ghost_id, ghost_id)
#undef CP
    int64_t vo = 0, mo = 0, go = 0;
    for (int64_t i = 0; i < h->n; ++i) {
        if (s->collision_force) { s->collision_force[2 * i] = h->cfx[i]; s->collision_force[2 * i + 1] = h->cfy[i]; }
        for (int k = 0; k < 4; ++k) {
            if (s->stress_accum) s->stress_accum[4 * i + k] = h->stress_accum[4 * i + k];
            if (s->stress_instant) s->stress_instant[4 * i + k] = h->stress_instant[4 * i + k];
            if (s->strain) s->strain[4 * i + k] = h->strain[4 * i + k];
        }
        if (s->vert_offsets) s->vert_offsets[i] = vo;
        if (s->vert_xy) memcpy(s->vert_xy + 2 * vo, h->ring[i], sizeof(szo_pt) * (size_t)h->npts[i]);
        vo += h->npts[i];
        if (i < h->n_init) {
            if (s->mc_offsets) s->mc_offsets[i] = mo;
            if (s->mc_x && h->nmc[i]) memcpy(s->mc_x + mo, h->mcx[i], sizeof(double) * (size_t)h->nmc[i]);
            if (s->mc_y && h->nmc[i]) memcpy(s->mc_y + mo, h->mcy[i], sizeof(double) * (size_t)h->nmc[i]);
            mo += h->nmc[i];
        } else if (s->mc_offsets) s->mc_offsets[i] = mo;
        if (s->ghost_offsets) s->ghost_offsets[i] = go;
        for (int g = 0; g < h->ghosts[i].n; ++g) {
            if (s->ghost_index) s->ghost_index[go] = h->ghosts[i].v[g] + 1;
            go++;
        }
    }
    if (s->vert_offsets) s->vert_offsets[h->n] = vo;
    if (s->mc_offsets) s->mc_offsets[h->n] = mo;
    if (s->ghost_offsets) s->ghost_offsets[h->n] = go;
    return SZ_OK;
}

/* ---- a2: ghosts (collisions.jl:881-1174) ------------------------------------------------- */
/* deepcopy_floe, floe_utils.jl:120-161 (Monte-Carlo points are not needed by ghosts on this
 * path: coupling runs after the ghosts are deleted, simulation.jl:138-161) */
static int64_t push_copy(sz_handle *h, int64_t src) {
    reserve_floes(h, h->n + 1);
    int64_t d = h->n++;
#define X(f) h->f[d] = h->f[src];
    DFIELDS(X)
#undef X
    for (int k = 0; k < 4; ++k) {
        h->stress_accum[4 * d + k] = h->stress_accum[4 * src + k];
        h->stress_instant[4 * d + k] = h->stress_instant[4 * src + k];
        h->strain[4 * d + k] = h->strain[4 * src + k];
    }
    h->status[d] = h->status[src]; h->id[d] = h->id[src]; h->ghost_id[d] = h->ghost_id[src];
    h->warn[d] = 0;
    h->npts[d] = h->npts[src];
    h->ring[d] = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)h->npts[src]);
    memcpy(h->ring[d], h->ring[src], sizeof(szo_pt) * (size_t)h->npts[src]);
    h->nmc[d] = 0; h->mcx[d] = h->mcy[d] = NULL;
    memset(&h->rows[d], 0, sizeof(rowlist));
    memset(&h->fuse_idx[d], 0, sizeof(intlist));
    memset(&h->ghosts[d], 0, sizeof(intlist));
    for (int g = 0; g < h->ghosts[src].n; ++g) intlist_push(&h->ghosts[d], h->ghosts[src].v[g]);
    return d;
}

/* _translate_floe!, floe_utils.jl:66-72 */
static void translate_floe(sz_handle *h, int64_t i, double dx, double dy) {
    h->cx[i] += dx; h->cy[i] += dy;
    for (int k = 0; k < h->npts[i]; ++k) { h->ring[i][k].x += dx; h->ring[i][k].y += dy; }
}

/* ghosts_on_bounds!, collisions.jl:881-901 */
static void ghosts_on_bounds(sz_handle *h, int64_t e, int wall, double tx, double ty, szo_regions *R) {
    int64_t nfloes = h->n;
    if (szo_clip(h->ring[e], h->npts[e], h->wall_ring[wall], 5, R) > 0) {
        int ng = h->ghosts[e].n;
        for (int g = 0; g < ng; ++g) push_copy(h, h->ghosts[e].v[g]);
        push_copy(h, e);
        for (int64_t i = nfloes; i < h->n; ++i) translate_floe(h, i, tx, ty);
    }
}

/* find_ghosts! (collisions.jl:925-952 E/W, :976-1003 N/S); axis 0 = x (east/west) */
static void find_ghosts(sz_handle *h, int64_t e, int axis, szo_regions *R) {
    int wmax = axis == 0 ? 2 : 0, wmin = axis == 0 ? 3 : 1;
    double L = h->val[wmax] - h->val[wmin];
    int64_t nfloes = h->n;
    double c = axis == 0 ? h->cx[e] : h->cy[e];
    if (c - h->rmax[e] < h->val[wmin]) ghosts_on_bounds(h, e, wmin, axis == 0 ? L : 0.0, axis == 0 ? 0.0 : L, R);
    else if (c + h->rmax[e] > h->val[wmax]) ghosts_on_bounds(h, e, wmax, axis == 0 ? -L : 0.0, axis == 0 ? 0.0 : -L, R);
    int64_t nn = h->n;
    if (nn > nfloes) {
        c = axis == 0 ? h->cx[e] : h->cy[e];
        if (c < h->val[wmin]) {
            translate_floe(h, e, axis == 0 ? L : 0.0, axis == 0 ? 0.0 : L);
            translate_floe(h, nn - 1, axis == 0 ? -L : 0.0, axis == 0 ? 0.0 : -L);
        } else if (h->val[wmax] < c) {
            translate_floe(h, e, axis == 0 ? -L : 0.0, axis == 0 ? 0.0 : -L);
            translate_floe(h, nn - 1, axis == 0 ? L : 0.0, axis == 0 ? 0.0 : L);
        }
    }
}

/* add_floe_ghosts!, collisions.jl:1017-1047 */
static void add_floe_ghosts(sz_handle *h, int axis, szo_regions *R) {
    int64_t nfloes = h->n, n0 = h->n;
    for (int64_t i = 0; i < n0; ++i) {
        if (h->status[i] == SZ_STATUS_ACTIVE && h->ghost_id[i] == 0) {
            find_ghosts(h, i, axis, R);
            int64_t nn = h->n;
            if (nn > nfloes) {
                int64_t ng = nn - nfloes;
                for (int64_t k = 0; k < ng; ++k) {
                    h->ghost_id[nfloes + k] = (k + 1) + h->ghosts[i].n;
                    h->ghosts[nfloes + k].n = 0; /* empty!.(floes.ghosts[new]) */
                }
                for (int64_t k = 0; k < ng; ++k) intlist_push(&h->ghosts[i], nfloes + k);
                nfloes += ng;
            }
        }
    }
}

int32_t szo_add_ghosts(sz_handle *h, int64_t *n_total) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "add_ghosts before set_domain");
    double t0 = now_ms();
    szo_regions R;
    szo_regions_init(&R);
    int ew = h->kind[2] == SZ_BOUNDARY_PERIODIC, ns = h->kind[0] == SZ_BOUNDARY_PERIODIC;
    if (ew) add_floe_ghosts(h, 0, &R); /* collisions.jl:1171 */
    if (ns) add_floe_ghosts(h, 1, &R); /* collisions.jl:1172 */
    szo_regions_free(&R);
    if (n_total) *n_total = h->n;
    h->ms[0] = now_ms() - t0;
    return SZ_OK;
}

/* simulation.jl:138-144 */
int32_t szo_remove_ghosts(sz_handle *h) {
    if (!h) return SZ_ERR_INVALID;
    for (int64_t i = h->n_init; i < h->n; ++i) free_floe_slot(h, i);
    h->n = h->n_init;
    for (int64_t i = 0; i < h->n; ++i) h->ghosts[i].n = 0;
    return SZ_OK;
}

/* ---- a8-a11: per-contact forces ------------------------------------------------------------ */
/* which_vertices_match_points, floe_utils.jl:331-352 (indices 0-based here) */
static int match_vertices(const szo_pt *ip, int nip, const szo_pt *reg, int nr, int *idx) {
    int m = 0, npoints = nip;
    if (nip > 0 && ip[0].x == ip[nip - 1].x && ip[0].y == ip[nip - 1].y) npoints -= 1;
    for (int i = 0; i < npoints; ++i) {
        double min_dist = INFINITY;
        int min_vert = 0;
        for (int j = 0; j < nr; ++j) {
            double dx = reg[j].x - ip[i].x, dy = reg[j].y - ip[i].y;
            double dist = sqrt(sqrt(dx * dx + dy * dy)); /* sqrt(GO.distance(..)) */
            if (dist < min_dist) { min_dist = dist; min_vert = j; }
        }
        if (min_dist < 1.0) idx[m++] = min_vert;
    }
    for (int a = 1; a < m; ++a) { /* sort! */
        int v = idx[a], b = a - 1;
        while (b >= 0 && idx[b] > v) { idx[b + 1] = idx[b]; --b; }
        idx[b + 1] = v;
    }
    return m;
}

/* test hook: which_vertices_match_points on caller data (golden values of test_floe_utils.jl:76-137); indices
 * come back 1-based like the reference's */
int32_t szo_test_match_vertices(const double *pts_xy, int32_t npts, const double *ring_xy, int32_t nring, int32_t *idx_out) {
    if (!pts_xy || !ring_xy || !idx_out || npts < 1 || nring < 1) return SZ_ERR_INVALID;
    int *idx = (int *)malloc(sizeof(int) * (size_t)npts);
    int m = match_vertices((const szo_pt *)pts_xy, npts, (const szo_pt *)ring_xy, nring, idx);
    for (int k = 0; k < m; ++k) idx_out[k] = idx[k] + 1;
    free(idx);
    return m;
}

/* _many_intersect_normal_force!, collisions.jl:78-119 */
static double many_intersect_normal(double dir[2], const szo_pt *reg, int nr, const szo_pt *P, int npp, double ff) {
    double x1 = 0, y1 = 0, dl = 0, Fx = 0, Fy = 0;
    int n_pts = 0;
    for (int i = 0; i < nr; ++i) {
        double x2 = reg[i].x, y2 = reg[i].y;
        if (i == 0) { x1 = x2; y1 = y2; continue; }
        double xmid = 0.5 * (x2 + x1), ymid = 0.5 * (y2 + y1);
        double dist = szo_point_ring_distance((szo_pt){xmid, ymid}, P, npp);
        if (dist < 1e-8) {
            double dx = x2 - x1, dy = y2 - y1;
            double mag = sqrt(dx * dx + dy * dy);
            double xt = xmid + (-dy / (100 * mag));
            double yt = ymid + (dx / (100 * mag));
            int in_region = szo_point_coveredby((szo_pt){xt, yt}, reg, nr);
            double fs = (in_region ? 1.0 : -1.0) * ff;
            Fx = Fx + fs * (-dy); Fy = Fy + fs * dx;
            dl += mag;
            n_pts += 1;
        }
        x1 = x2; y1 = y2;
    }
    if (0 < n_pts && n_pts < nr - 1) {
        dl /= n_pts;
        if (dl > 0.1) {
            double nf = sqrt(Fx * Fx + Fy * Fy);
            dir[0] = Fx / nf; dir[1] = Fy / nf;
        }
    }
    return dl;
}

/* calc_normal_force, collisions.jl:30-70 */
static double normal_force(const szo_pt *P, int npp, const szo_pt *Q, int nqp, const szo_pt *reg, int nr,
                           double area, const szo_pt *ip, int nip, double ff, double force[2],
                           szo_regions *scratch, int64_t *clipfail) {
    double dir[2] = {0, 0}, dl = 0;
    int *idx = (int *)malloc(sizeof(int) * (size_t)(nip > 0 ? nip : 1));
    int m = match_vertices(ip, nip, reg, nr, idx);
    if (m == 2) {
        double dx = reg[idx[1]].x - reg[idx[0]].x, dy = reg[idx[1]].y - reg[idx[0]].y;
        dl = sqrt(dx * dx + dy * dy);
        if (dl > 0.1) { dir[0] = -dy / dl; dir[1] = dx / dl; }
    } else if (m != 0) {
        dl = many_intersect_normal(dir, reg, nr, P, npp, ff);
    }
    free(idx);
    if (dl > 0.1) {
        szo_pt *P2 = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)npp);
        for (int k = 0; k < npp; ++k) { P2[k].x = P[k].x + dir[0]; P2[k].y = P[k].y + dir[1]; }
        szo_clip(P2, npp, Q, nqp, scratch);
        if (scratch->failed && clipfail) (*clipfail)++;
        for (int r = 0; r < scratch->nreg; ++r) {
            const szo_pt *nr_ = scratch->pts + scratch->off[r];
            int nn = scratch->off[r + 1] - scratch->off[r];
            if (szo_rings_intersect(nr_, nn, reg, nr) && szo_ring_area(nr_, nn) / area > 1) {
                dir[0] *= -1; dir[1] *= -1;
            }
        }
        free(P2);
    }
    force[0] = dir[0] * area * ff; force[1] = dir[1] * area * ff;
    return dl;
}

typedef struct {
    int n;
    double force[2], fpoint[2], overlap, dl;
} contact;

/* calc_elastic_forces, collisions.jl:149-188.  regions/areas are compacted in place. */
static int elastic_forces(const szo_pt *P, int npp, const szo_pt *Q, int nqp, szo_regions *R, double *areas,
                          double ff, contact **out, szo_regions *scratch, int64_t *clipfail) {
    szo_pt *ip = NULL;
    int nip = szo_intersection_points(P, npp, Q, nqp, &ip);
    int ncontact = 0;
    int *keep = (int *)malloc(sizeof(int) * (size_t)(R->nreg > 0 ? R->nreg : 1));
    if (nip >= 2) {
        int n1 = npp - 1, n2 = nqp - 1;
        double min_area = (double)((n1 < n2 ? n1 : n2) * 100) / 1.75;
        for (int i = 0; i < R->nreg; ++i)
            if (!(areas[i] < min_area)) keep[ncontact++] = i;
    }
    contact *c = (contact *)calloc((size_t)(ncontact > 0 ? ncontact : 1), sizeof(contact));
    for (int k = 0; k < ncontact; ++k) {
        int r = keep[k];
        c[k].overlap = areas[r];
        if (areas[r] != 0) {
            const szo_pt *reg = R->pts + R->off[r];
            int nr = R->off[r + 1] - R->off[r];
            szo_pt ce = szo_ring_centroid(reg, nr);
            c[k].fpoint[0] = ce.x; c[k].fpoint[1] = ce.y;
            c[k].dl = normal_force(P, npp, Q, nqp, reg, nr, areas[r], ip, nip, ff, c[k].force, scratch, clipfail);
        }
    }
    free(keep);
    free(ip);
    *out = c;
    return ncontact;
}

/* calc_friction_forces, collisions.jl:243-283; (iu,iv,...) from _get_velocity :206-214 */
static void friction_force(const sz_handle *h, int64_t i, double ju0, double jv0, double jxi, double jcx, double jcy,
                           const contact *c, double out[2]) {
    double G = h->cfg.E / (2 * (1 + h->cfg.nu));
    double px = c->fpoint[0], py = c->fpoint[1];
    double nnorm = sqrt(c->force[0] * c->force[0] + c->force[1] * c->force[1]);
    double iu = h->u[i] + h->xi[i] * (px - h->cx[i]);
    double iv = h->v[i] + h->xi[i] * (py - h->cy[i]);
    double ju = ju0 + jxi * (px - jcx);
    double jv = jv0 + jxi * (py - jcy);
    double udiff = iu - ju, vdiff = iv - jv;
    double vnorm = sqrt(udiff * udiff + vdiff * vdiff);
    double xdir = 0, ydir = 0;
    if (udiff != 0 || vdiff != 0) { xdir = udiff / vnorm; ydir = vdiff / vnorm; }
    double dot_dir = xdir * udiff + ydir * vdiff;
    double dt = (double)h->cfg.dt;
    double xf = G * c->dl * dt * nnorm * xdir * -dot_dir;
    double yf = G * c->dl * dt * nnorm * ydir * -dot_dir;
    double nf = sqrt(xf * xf + yf * yf);
    if (nf > h->cfg.mu * nnorm) {
        xf = -h->cfg.mu * nnorm * xdir;
        yf = -h->cfg.mu * nnorm * ydir;
    }
    out[0] = xf; out[1] = yf;
}

/* add_interactions!, collisions.jl:285-309 (one row) */
static void add_row(sz_handle *h, int64_t i, double idx, double fx, double fy, double px, double py, double ov) {
    if (fx != 0 || fy != 0) {
        rowlist *l = &h->rows[i];
        if (l->n == l->cap) {
            l->cap = l->cap ? 2 * l->cap : 4;
            l->r = (double *)realloc(l->r, sizeof(double) * NROWF * (size_t)l->cap);
        }
        double *r = l->r + NROWF * l->n++;
        r[0] = idx; r[1] = fx; r[2] = fy; r[3] = px; r[4] = py; r[5] = 0.0; r[6] = ov;
        h->overarea[i] += ov;
    }
}

/* floe_floe_interaction!, collisions.jl:347-408.  Returns 1 if total overlap area > 0. */
static int floe_floe_interaction(sz_handle *h, int64_t i, int64_t j, szo_regions *R, szo_regions *scratch,
                                 int *fused, int64_t *clipfail) {
    const szo_pt *P = h->ring[i], *Q = h->ring[j];
    int npp = h->npts[i], nqp = h->npts[j];
    szo_clip(P, npp, Q, nqp, R);
    if (R->failed) (*clipfail)++;
    double *areas = (double *)malloc(sizeof(double) * (size_t)(R->nreg > 0 ? R->nreg : 1));
    double total = 0;
    for (int r = 0; r < R->nreg; ++r) {
        areas[r] = szo_ring_area(R->pts + R->off[r], R->off[r + 1] - R->off[r]);
        total += areas[r];
    }
    *fused = 0;
    int overl = total > 0;
    if (overl) {
        if (fmax(total / h->area[i], total / h->area[j]) > h->cfg.floe_floe_max_overlap) {
            h->status[i] = SZ_STATUS_FUSE;
            intlist_push(&h->fuse_idx[i], j);
            *fused = 1;
        } else {
            double ih = h->height[i], ir = sqrt(h->area[i]), jh = h->height[j], jr = sqrt(h->area[j]);
            double ff = (ir > 1e5 || jr > 1e5) ? h->cfg.E * fmin(ih, jh) / fmin(ir, jr)
                                               : h->cfg.E * (ih * jh) / (ih * jr + jh * ir);
            contact *c = NULL;
            int np = elastic_forces(P, npp, Q, nqp, R, areas, ff, &c, scratch, clipfail);
            for (int k = 0; k < np; ++k) {
                double fr[2];
                friction_force(h, i, h->u[j], h->v[j], h->xi[j], h->cx[j], h->cy[j], &c[k], fr);
                add_row(h, i, (double)(j + 1), c[k].force[0] + fr[0], c[k].force[1] + fr[1], c[k].fpoint[0],
                        c[k].fpoint[1], c[k].overlap);
            }
            free(c);
        }
    }
    free(areas);
    return overl;
}

/* floe_domain_element_interaction!, collisions.jl:427-557; elem: 0..3 walls, 4+k topography */
static void floe_element_interaction(sz_handle *h, int64_t i, int elem, szo_regions *R, szo_regions *scratch,
                                     int64_t *clipfail) {
    int is_wall = elem < 4;
    int kind = is_wall ? h->kind[elem] : SZ_BOUNDARY_COLLISION;
    if (kind == SZ_BOUNDARY_PERIODIC) return; /* :459-468 */
    const szo_pt *Q = is_wall ? h->wall_ring[elem] : h->topo_ring[elem - 4];
    int nqp = is_wall ? 5 : h->topo_np[elem - 4];
    const szo_pt *P = h->ring[i];
    int npp = h->npts[i];
    szo_clip(P, npp, Q, nqp, R);
    if (R->failed) (*clipfail)++;
    double *areas = (double *)malloc(sizeof(double) * (size_t)(R->nreg > 0 ? R->nreg : 1));
    double max_area = 0, sum = 0;
    for (int r = 0; r < R->nreg; ++r) {
        areas[r] = szo_ring_area(R->pts + R->off[r], R->off[r + 1] - R->off[r]);
        sum += areas[r];
        if (areas[r] > max_area) max_area = areas[r];
    }
    if (kind == SZ_BOUNDARY_OPEN) { /* :427-441 */
        if (sum > 0) h->status[i] = SZ_STATUS_REMOVE;
        free(areas);
        return;
    }
    if (max_area > 0) { /* :522-555 */
        if (max_area / h->area[i] > h->cfg.floe_domain_max_overlap) {
            h->status[i] = SZ_STATUS_REMOVE;
        } else {
            double ff = h->cfg.E * h->height[i] / sqrt(h->area[i]);
            contact *c = NULL;
            int np = elastic_forces(P, npp, Q, nqp, R, areas, ff, &c, scratch, clipfail);
            /* _normal_direction_correct!, boundaries.jl:37-40,73-76,110-113,147-150 */
            for (int k = 0; k < np && is_wall; ++k) {
                if (elem == 0 && c[k].fpoint[1] >= h->val[0]) c[k].force[0] = 0.0;
                if (elem == 1 && c[k].fpoint[1] <= h->val[1]) c[k].force[0] = 0.0;
                if (elem == 2 && c[k].fpoint[0] >= h->val[2]) c[k].force[1] = 0.0;
                if (elem == 3 && c[k].fpoint[0] <= h->val[3]) c[k].force[1] = 0.0;
            }
            double ju = 0, jv = 0; /* boundaries.jl:522,565; topography.jl:76 */
            if (is_wall && kind == SZ_BOUNDARY_MOVING) { ju = h->wu[elem]; jv = h->wv[elem]; }
            for (int k = 0; k < np; ++k) {
                double fr[2];
                friction_force(h, i, ju, jv, 0.0, 0.0, 0.0, &c[k], fr);
                add_row(h, i, (double)(-(elem + 1)), c[k].force[0] + fr[0], c[k].force[1] + fr[1],
                        c[k].fpoint[0], c[k].fpoint[1], c[k].overlap);
            }
            free(c);
        }
    }
    free(areas);
}

/* floe_domain_interaction!, collisions.jl:594-662 */
static void floe_domain_interaction(sz_handle *h, int64_t i, szo_regions *R, szo_regions *scratch, int64_t *clipfail,
                                    int64_t *ndom) {
    double cx = h->cx[i], cy = h->cy[i], r = h->rmax[i];
    if (cy + r > h->val[0]) { floe_element_interaction(h, i, 0, R, scratch, clipfail); (*ndom)++; }
    if (cy - r < h->val[1]) { floe_element_interaction(h, i, 1, R, scratch, clipfail); (*ndom)++; }
    if (cx + r > h->val[2]) { floe_element_interaction(h, i, 2, R, scratch, clipfail); (*ndom)++; }
    if (cx - r < h->val[3]) { floe_element_interaction(h, i, 3, R, scratch, clipfail); (*ndom)++; }
    for (int k = 0; k < h->n_topo; ++k) {
        double dx = h->topo_cx[k] - cx, dy = h->topo_cy[k] - cy, rr = h->topo_rmax[k] + r;
        if (dx * dx + dy * dy < rr * rr) { floe_element_interaction(h, i, 4 + k, R, scratch, clipfail); (*ndom)++; }
    }
}

/* ---- a3/a4: broad phase + image-pair filter ------------------------------------------------ */
/* potential_interaction, collisions.jl:705-710 */
static inline int potential_interaction(const sz_handle *h, int64_t i, int64_t j) {
    double dx = h->cx[i] - h->cx[j], dy = h->cy[i] - h->cy[j], rr = h->rmax[i] + h->rmax[j];
    return dx * dx + dy * dy < rr * rr;
}

static int cmp_i64(const void *a, const void *b) {
    int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
    return (x > y) - (x < y);
}

/* All (i<j) passing the circle test, lexicographically sorted.  A uniform grid with cell
 * edge >= 2 rmax_max finds exactly the set of the reference's O(N^2) loop
 * (collisions.jl:745-763); with SZO_BRUTE_FORCE in the environment the literal double loop
 * is used instead (tests compare the two). */
static void build_candidates(sz_handle *h) {
    int64_t n = h->n, cap = 16 + 8 * n, m = 0;
    int64_t *out = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)cap);
    if (getenv("SZO_BRUTE_FORCE") || n < 64) {
        for (int64_t i = 0; i < n; ++i)
            for (int64_t j = i + 1; j < n; ++j)
                if (potential_interaction(h, i, j)) {
                    if (m == cap) { cap *= 2; out = (int64_t *)realloc(out, sizeof(int64_t) * 2 * (size_t)cap); }
                    out[2 * m] = i; out[2 * m + 1] = j; m++;
                }
    } else {
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY, rm = 0;
        for (int64_t i = 0; i < n; ++i) {
            xmin = fmin(xmin, h->cx[i]); xmax = fmax(xmax, h->cx[i]);
            ymin = fmin(ymin, h->cy[i]); ymax = fmax(ymax, h->cy[i]);
            rm = fmax(rm, h->rmax[i]);
        }
        double cs = 2.0 * rm * (1.0 + 1e-9) + 1e-9;
        int64_t gx = (int64_t)floor((xmax - xmin) / cs) + 1, gy = (int64_t)floor((ymax - ymin) / cs) + 1;
        while (gx * gy > 4 * n + 64) { cs *= 1.5; gx = (int64_t)floor((xmax - xmin) / cs) + 1; gy = (int64_t)floor((ymax - ymin) / cs) + 1; }
        int64_t *start = (int64_t *)calloc((size_t)(gx * gy + 1), sizeof(int64_t));
        int64_t *cell = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        int64_t *items = (int64_t *)malloc(sizeof(int64_t) * (size_t)n);
        for (int64_t i = 0; i < n; ++i) {
            int64_t ix = (int64_t)floor((h->cx[i] - xmin) / cs), iy = (int64_t)floor((h->cy[i] - ymin) / cs);
            cell[i] = iy * gx + ix;
            start[cell[i] + 1]++;
        }
        for (int64_t c = 0; c < gx * gy; ++c) start[c + 1] += start[c];
        int64_t *fill = (int64_t *)malloc(sizeof(int64_t) * (size_t)(gx * gy));
        memcpy(fill, start, sizeof(int64_t) * (size_t)(gx * gy));
        for (int64_t i = 0; i < n; ++i) items[fill[cell[i]]++] = i;
        int64_t *tmp = (int64_t *)malloc(sizeof(int64_t) * 256);
        int64_t tcap = 256;
        for (int64_t i = 0; i < n; ++i) {
            int64_t ix = cell[i] % gx, iy = cell[i] / gx, t = 0;
            for (int64_t yy = iy - 1; yy <= iy + 1; ++yy) {
                if (yy < 0 || yy >= gy) continue;
                for (int64_t xx = ix - 1; xx <= ix + 1; ++xx) {
                    if (xx < 0 || xx >= gx) continue;
                    int64_t c = yy * gx + xx;
                    for (int64_t k = start[c]; k < start[c + 1]; ++k) {
                        int64_t j = items[k];
                        if (j > i && potential_interaction(h, i, j)) {
                            if (t == tcap) { tcap *= 2; tmp = (int64_t *)realloc(tmp, sizeof(int64_t) * (size_t)tcap); }
                            tmp[t++] = j;
                        }
                    }
                }
            }
            qsort(tmp, (size_t)t, sizeof(int64_t), cmp_i64);
            for (int64_t k = 0; k < t; ++k) {
                if (m == cap) { cap *= 2; out = (int64_t *)realloc(out, sizeof(int64_t) * 2 * (size_t)cap); }
                out[2 * m] = i; out[2 * m + 1] = tmp[k]; m++;
            }
        }
        free(tmp); free(fill); free(items); free(cell); free(start);
    }
    free(h->cand);
    h->cand = out;
    h->n_cand = m;
}

/* collide_pairs Dict, collisions.jl:743,751-775, in serial (i,j) order. */
typedef struct { int64_t k1, k2, g1, g2; int used; } dict_ent;

static void filter_pairs(sz_handle *h) {
    int64_t m = h->n_cand, np = 0;
    int64_t *out = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(m > 0 ? m : 1));
    size_t tsz = 64;
    while (tsz < (size_t)(2 * m + 8)) tsz <<= 1;
    dict_ent *T = (dict_ent *)calloc(tsz, sizeof(dict_ent));
    for (int64_t p = 0; p < m; ++p) {
        int64_t i = h->cand[2 * p], j = h->cand[2 * p + 1];
        int64_t k1, k2, g1, g2;
        if (h->id[i] > h->id[j]) { k1 = h->id[i]; k2 = h->id[j]; g1 = h->ghost_id[i]; g2 = h->ghost_id[j]; }
        else { k1 = h->id[j]; k2 = h->id[i]; g1 = h->ghost_id[j]; g2 = h->ghost_id[i]; }
        if (k1 == k2) continue;
        uint64_t hsh = ((uint64_t)k1 * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)k2 * 0xC2B2AE3D27D4EB4Full);
        size_t s = (size_t)(hsh >> 17) & (tsz - 1);
        while (T[s].used && !(T[s].k1 == k1 && T[s].k2 == k2)) s = (s + 1) & (tsz - 1);
        if (!T[s].used) { T[s].used = 1; T[s].k1 = k1; T[s].k2 = k2; T[s].g1 = g1; T[s].g2 = g2; } /* get! */
        int a = (g1 == T[s].g1), b = (g2 == T[s].g2);
        if ((a && b) || (a != b)) { out[2 * np] = i; out[2 * np + 1] = j; np++; }
    }
    free(T);
    free(h->pairs);
    h->pairs = out;
    h->n_pairs = np;
}

/* calc_torque!, collisions.jl:673-686 */
static void calc_torque(sz_handle *h, int64_t i) {
    rowlist *l = &h->rows[i];
    for (int k = 0; k < l->n; ++k) {
        double *r = l->r + NROWF * k;
        double xp = r[3] - h->cx[i], yp = r[4] - h->cy[i];
        r[5] = xp * r[2] - yp * r[1];
    }
}

/* _update_boundary!, boundaries.jl:526-544 */
static void update_boundaries(sz_handle *h) {
    for (int w = 0; w < 4; ++w) {
        if (h->kind[w] != SZ_BOUNDARY_MOVING) continue;
        if (w < 2) {
            double d = h->wv[w] * h->cfg.dt;
            h->rect[w][2] += d; h->rect[w][3] += d; h->val[w] += d;
        } else {
            double d = h->wu[w] * h->cfg.dt;
            h->rect[w][0] += d; h->rect[w][1] += d; h->val[w] += d;
        }
        make_wall_ring(h->wall_ring[w], h->rect[w]);
    }
}

/* timestep_collisions!, collisions.jl:734-864 */
int32_t szo_step_collisions(sz_handle *h) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "step_collisions before set_domain");
    int64_t n = h->n;
    double t0 = now_ms();
    for (int64_t i = 0; i < n; ++i) { /* :747-749 */
        h->cfx[i] = h->cfy[i] = 0.0; h->ctrq[i] = 0.0; h->rows[i].n = 0; h->fuse_idx[i].n = 0;
    }
    build_candidates(h);
    filter_pairs(h);
    double t1 = now_ms();
    /* group the filtered pairs by i so the loop is the reference's `for i` (:745) */
    int64_t *first = (int64_t *)calloc((size_t)(n + 1), sizeof(int64_t));
    for (int64_t p = 0; p < h->n_pairs; ++p) first[h->pairs[2 * p] + 1]++;
    for (int64_t i = 0; i < n; ++i) first[i + 1] += first[i];
    char *ovl = (char *)calloc((size_t)(h->n_pairs > 0 ? h->n_pairs : 1), 1);
    char *fus = (char *)calloc((size_t)(h->n_pairs > 0 ? h->n_pairs : 1), 1);
    int64_t clipfail = 0, ndom = 0;
#pragma omp parallel reduction(+ : clipfail, ndom)
    {
        szo_regions R, S;
        szo_regions_init(&R);
        szo_regions_init(&S);
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            for (int64_t p = first[i]; p < first[i + 1]; ++p) {
                int fused = 0;
                ovl[p] = (char)floe_floe_interaction(h, i, h->pairs[2 * p + 1], &R, &S, &fused, &clipfail);
                fus[p] = (char)fused;
            }
            floe_domain_interaction(h, i, &R, &S, &clipfail, &ndom); /* :788-794 */
        }
        szo_regions_free(&R);
        szo_regions_free(&S);
    }
    h->n_clip_fail = clipfail;
    h->n_domain_pairs = ndom;
    free(h->overlap); free(h->fuse);
    h->overlap = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(h->n_pairs > 0 ? h->n_pairs : 1));
    h->fuse = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(h->n_pairs > 0 ? h->n_pairs : 1));
    h->n_overlap = h->n_fuse = 0;
    for (int64_t p = 0; p < h->n_pairs; ++p) {
        if (ovl[p]) { h->overlap[2 * h->n_overlap] = h->pairs[2 * p]; h->overlap[2 * h->n_overlap + 1] = h->pairs[2 * p + 1]; h->n_overlap++; }
        if (fus[p]) { h->fuse[2 * h->n_fuse] = h->pairs[2 * p]; h->fuse[2 * h->n_fuse + 1] = h->pairs[2 * p + 1]; h->n_fuse++; }
    }
    free(ovl); free(fus); free(first);
    double t2 = now_ms();
    update_boundaries(h); /* :797 */
    for (int64_t i = 0; i < n; ++i) {
        if (h->status[i] == SZ_STATUS_FUSE) { /* :801-806 */
            int nf = h->fuse_idx[i].n;
            for (int k = 0; k < nf; ++k) {
                int64_t idx = h->fuse_idx[i].v[k];
                h->status[idx] = SZ_STATUS_FUSE;
                intlist_push(&h->fuse_idx[idx], i);
            }
        }
        int ni = h->rows[i].n; /* :808-827 */
        for (int k = 0; k < ni; ++k) {
            double *r = h->rows[i].r + NROWF * k;
            double jf = r[0];
            if (jf <= (double)n && jf > (double)(i + 1)) {
                int64_t j = (int64_t)jf - 1;
                double fx = r[1], fy = r[2], px = r[3], py = r[4], ov = r[6];
                int before = h->rows[j].n;
                add_row(h, j, (double)(i + 1), fx, fy, px, py, ov);
                if (h->rows[j].n > before) {
                    double *rj = h->rows[j].r + NROWF * (h->rows[j].n - 1);
                    rj[1] *= -1; rj[2] *= -1;
                }
            }
        }
    }
    for (int64_t i = 0; i < h->n_init; ++i) { /* :830-862 */
        for (int g = 0; g < h->ghosts[i].n; ++g) {
            int64_t gi = h->ghosts[i].v[g];
            int gnp = h->rows[gi].n;
            double sx = h->cx[gi] - h->cx[i], sy = h->cy[gi] - h->cy[i];
            for (int k = 0; k < gnp; ++k) {
                double *r = h->rows[gi].r + NROWF * k;
                r[3] -= sx; r[4] -= sy;
            }
            for (int k = 0; k < gnp; ++k) {
                double *r = h->rows[gi].r + NROWF * k;
                double f0 = r[0], fx = r[1], fy = r[2], px = r[3], py = r[4], ov = r[6];
                add_row(h, i, f0, fx, fy, px, py, ov);
            }
        }
        calc_torque(h, i);
        double sx = 0, sy = 0, st = 0;
        for (int k = 0; k < h->rows[i].n; ++k) {
            double *r = h->rows[i].r + NROWF * k;
            sx += r[1]; sy += r[2]; st += r[5];
        }
        h->cfx[i] += sx; h->cfy[i] += sy; h->ctrq[i] += st;
    }
    double t3 = now_ms();
    h->ms[1] = t1 - t0; h->ms[2] = t2 - t1; h->ms[3] = t3 - t2;
    return SZ_OK;
}

/* ---- a15-a18: one-way coupling (coupling.jl:1486-1589) --------------------------------------- */
/* Interpolations.linear_interpolation on the knot window of find_interp_knots
 * (coupling.jl:702-797,845-902) == bilinear on the full lattice; periodic axes use grid
 * lines 1..N and wrap (line N+1 is never read, coupling.jl:722-745). */
static inline double bilinear(const double *F, int Nx, int Ny, int i0, int i1, int j0, int j1, double wx, double wy) {
    (void)Ny;
    size_t s = (size_t)(Nx + 1);
    double f00 = F[i0 + s * j0], f10 = F[i1 + s * j0], f01 = F[i0 + s * j1], f11 = F[i1 + s * j1];
    return (1 - wy) * ((1 - wx) * f00 + wx * f10) + wy * ((1 - wx) * f01 + wx * f11);
}

/* shift_cell_idx, coupling.jl:1155-1182 (1-based idx, nlines = N + 1) */
static int shift_cell_idx(int idx, int nlines, int periodic) {
    if (!periodic) return idx;
    int ncells = nlines - 1;
    return idx < 1 ? (idx + ncells) : (ncells < idx ? (idx - ncells) : idx);
}

/* floe_to_grid_info! + add_point!, coupling.jl:1329-1454, on the per-floe record list (the reference merges
 * into a cell's LAST entry when it belongs to the same floe; floes are visited one after the other, so that is
 * "one entry per (cell, floe)") */
static void floe_to_grid_info(sz_handle *h, cellrec **list, int *n, int *cap, int64_t floe, int xidx, int yidx, double tx,
                              double ty) {
    int per_x = h->kind[2] == SZ_BOUNDARY_PERIODIC, per_y = h->kind[0] == SZ_BOUNDARY_PERIODIC;
    int sx = shift_cell_idx(xidx, h->Nx + 1, per_x), sy = shift_cell_idx(yidx, h->Ny + 1, per_y);
    double ddx = (sx - xidx) * h->dx, ddy = (sy - yidx) * h->dy;
    for (int k = 0; k < *n; ++k)
        if ((*list)[k].ix == sx - 1 && (*list)[k].iy == sy - 1) {
            (*list)[k].tx += -tx; (*list)[k].ty += -ty; (*list)[k].npts += 1;
            return;
        }
    if (*n == *cap) {
        *cap = *cap ? 2 * *cap : 4;
        *list = (cellrec *)realloc(*list, sizeof(cellrec) * (size_t)*cap);
    }
    cellrec *r = &(*list)[(*n)++];
    r->ix = sx - 1; r->iy = sy - 1; r->floe = floe; r->tx = -tx; r->ty = -ty; r->dx = ddx; r->dy = ddy; r->npts = 1;
}

/* in_bounds, coupling.jl:494-597: a periodic axis accepts every coordinate, a non-periodic one the closed grid extent */
static inline int in_bounds(const sz_handle *h, double x, double y, int per_x, int per_y) {
    return (per_x || (h->x0 <= x && x <= h->xf)) && (per_y || (h->y0 <= y && y <= h->yf));
}
/* find_center_cell_index, coupling.jl:466-470 (1-based like the reference; may lie outside 1..N+1) */
static inline void center_cell_index(const sz_handle *h, double x, double y, int *xidx, int *yidx) {
    *xidx = (int)floor((x - h->x0) / h->dx + 0.5) + 1;
    *yidx = (int)floor((y - h->y0) / h->dy + 0.5) + 1;
}
/* test hooks for the truth tables of test_coupling.jl:165-195 */
int32_t szo_test_in_bounds(sz_handle *h, double x, double y, int32_t per_x, int32_t per_y) {
    return h ? in_bounds(h, x, y, per_x, per_y) : SZ_ERR_INVALID;
}
int32_t szo_test_find_center_cell_index(sz_handle *h, double x, double y, int32_t out[2]) {
    if (!h || !out) return SZ_ERR_INVALID;
    int xi, yi;
    center_cell_index(h, x, y, &xi, &yi);
    out[0] = xi;
    out[1] = yi;
    return SZ_OK;
}

static void coupling_one_floe(sz_handle *h, int64_t i) {
    const sz_config *c = &h->cfg;
    cellrec *rl = NULL;
    int nrl = 0, caprl = 0;
    int per_x = h->kind[2] == SZ_BOUNDARY_PERIODIC, per_y = h->kind[0] == SZ_BOUNDARY_PERIODIC;
    double a = h->alpha[i];
    double tot_x = 0, tot_y = 0, tot_trq = 0, tot_hflx = 0;
    int npoints = 0;
    /* first pass: count in-bounds points (calc_subfloe_values!, coupling.jl:627-657) */
    int nm = h->nmc[i];
    double *X = (double *)malloc(sizeof(double) * 2 * (size_t)(nm > 0 ? nm : 1)), *Y = X + nm;
    for (int k = 0; k < nm; ++k) {
        double px = cos(a) * h->mcx[i][k] - sin(a) * h->mcy[i][k];
        double py = sin(a) * h->mcx[i][k] + cos(a) * h->mcy[i][k];
        double x = px + h->cx[i], y = py + h->cy[i];
        if (in_bounds(h, x, y, per_x, per_y)) { X[npoints] = x; Y[npoints] = y; npoints++; }
    }
    if (npoints == 0) {
        h->status[i] = SZ_STATUS_REMOVE; /* coupling.jl:1507-1508 */
        h->frec[i] = NULL;
        h->nfrec[i] = 0;
        free(X);
        return;
    }
    double ma_ratio = h->mass[i] / h->area[i];
    double xcor = ma_ratio * c->f * h->v[i], ycor = ma_ratio * c->f * h->u[i];
    tot_x = npoints * xcor; tot_y = -npoints * ycor; /* :1522-1525 */
    for (int k = 0; k < npoints; ++k) {
        double x = X[k], y = Y[k];
        double xc = x - h->cx[i], yc = y - h->cy[i];
        double th = atan2(yc, xc), rad = sqrt(xc * xc + yc * yc);
        double up = h->u[i] - h->xi[i] * rad * sin(th);
        double vp = h->v[i] + h->xi[i] * rad * cos(th);
        /* lattice cell + weights */
        double gx = (x - h->x0) / h->dx, gy = (y - h->y0) / h->dy;
        double fx = floor(gx), fy = floor(gy);
        long ci = (long)fx, cj = (long)fy;
        double wx = gx - fx, wy = gy - fy;
        int i0, i1, j0, j1;
        if (per_x) { i0 = (int)(((ci % h->Nx) + h->Nx) % h->Nx); i1 = (i0 + 1) % h->Nx; }
        else { if (ci >= h->Nx) { ci = h->Nx - 1; wx = 1.0; } if (ci < 0) { ci = 0; wx = 0.0; } i0 = (int)ci; i1 = i0 + 1; }
        if (per_y) { j0 = (int)(((cj % h->Ny) + h->Ny) % h->Ny); j1 = (j0 + 1) % h->Ny; }
        else { if (cj >= h->Ny) { cj = h->Ny - 1; wy = 1.0; } if (cj < 0) { cj = 0; wy = 0.0; } j0 = (int)cj; j1 = j0 + 1; }
        double uatm = bilinear(h->atm_u, h->Nx, h->Ny, i0, i1, j0, j1, wx, wy);
        double vatm = bilinear(h->atm_v, h->Nx, h->Ny, i0, i1, j0, j1, wx, wy);
        double uocn = bilinear(h->ocn_u, h->Nx, h->Ny, i0, i1, j0, j1, wx, wy);
        double vocn = bilinear(h->ocn_v, h->Nx, h->Ny, i0, i1, j0, j1, wx, wy);
        double hfl = bilinear(h->ocn_hflx, h->Nx, h->Ny, i0, i1, j0, j1, wx, wy);
        /* calc_atmosphere_forcing, coupling.jl:1212-1232 */
        double dua = uatm - up, dva = vatm - vp;
        double na = sqrt(dua * dua + dva * dva);
        double tax = c->rho_a * c->Cd_ia * na * dua, tay = c->rho_a * c->Cd_ia * na * dva;
        /* calc_ocean_forcing!, coupling.jl:1277-1299 */
        double duo = uocn - up, dvo = vocn - vp;
        double no = sqrt(duo * duo + dvo * dvo);
        double tox = c->rho_o * c->Cd_io * no * (cos(c->turn_theta) * duo - sin(c->turn_theta) * dvo);
        double toy = c->rho_o * c->Cd_io * no * (sin(c->turn_theta) * duo + cos(c->turn_theta) * dvo);
        double tpx = -ma_ratio * c->f * vocn, tpy = ma_ratio * c->f * uocn;
        double tx = tax + tpx + tox, ty = tay + tpy + toy;
        double trq = (-tx * sin(th) + ty * cos(th)) * rad;
        tot_x += tx; tot_y += ty; tot_trq += trq; tot_hflx += hfl;
        {
            int xidx, yidx;
            center_cell_index(h, x, y, &xidx, &yidx);
            floe_to_grid_info(h, &rl, &nrl, &caprl, i, xidx, yidx, tox, toy);
        }
    }
    h->frec[i] = rl;
    h->nfrec[i] = nrl;
    h->fxOA[i] = tot_x / npoints * h->area[i]; /* :1583-1586 */
    h->fyOA[i] = tot_y / npoints * h->area[i];
    h->trqOA[i] = tot_trq / npoints * h->area[i];
    h->hflx[i] = tot_hflx / npoints;
    free(X);
}

static int cmp_cellrec(const void *a, const void *b) {
    const cellrec *x = (const cellrec *)a, *y = (const cellrec *)b;
    if (x->iy != y->iy) return x->iy < y->iy ? -1 : 1;
    if (x->ix != y->ix) return x->ix < y->ix ? -1 : 1;
    return (x->floe > y->floe) - (x->floe < y->floe);
}

/* center_cell_coords + check_cell_bounds, coupling.jl:931-1140 (ix, iy 1-based): xmin, xmax, ymin, ymax */
static void center_cell_coords(const sz_handle *h, int ix, int iy, double out[4]) {
    int per_x = h->kind[2] == SZ_BOUNDARY_PERIODIC, per_y = h->kind[0] == SZ_BOUNDARY_PERIODIC;
    double xmin = (ix - 1.5) * h->dx + h->x0, xmax = xmin + h->dx;
    double ymin = (iy - 1.5) * h->dy + h->y0, ymax = ymin + h->dy;
    if (!per_x) {
        xmin = xmin < h->x0 ? h->x0 : (xmin > h->xf ? h->xf : xmin);
        xmax = xmax > h->xf ? h->xf : (xmax < h->x0 ? h->x0 : xmax);
    }
    if (!per_y) {
        ymin = ymin < h->y0 ? h->y0 : (ymin > h->yf ? h->yf : ymin);
        ymax = ymax > h->yf ? h->yf : (ymax < h->y0 ? h->y0 : ymax);
    }
    out[0] = xmin; out[1] = xmax; out[2] = ymin; out[3] = ymax;
}

/* calc_two_way_coupling!, coupling.jl:1617-1680 */
static void two_way_coupling(sz_handle *h) {
    const sz_config *c = &h->cfg;
    int nx1 = h->Nx + 1, ny1 = h->Ny + 1;
    double cell_area = h->dx * h->dy;
    /* first record of every cell in the (cell, floe)-sorted registry */
    int64_t *first = (int64_t *)calloc((size_t)nx1 * ny1 + 1, sizeof(int64_t));
    for (int64_t k = 0; k < h->n_crec; ++k) first[(size_t)h->crec[k].iy * nx1 + h->crec[k].ix + 1]++;
    for (size_t q = 0; q < (size_t)nx1 * ny1; ++q) first[q + 1] += first[q];
#pragma omp parallel
    {
        szo_regions R;
        szo_regions_init(&R);
        szo_pt *T = NULL;
        int capT = 0;
#pragma omp for schedule(dynamic, 16)
        for (int q = 0; q < nx1 * ny1; ++q) {
            int ix = q % nx1, iy = q / nx1;
            double tx = 0, ty = 0, si = 0;
            if (first[q + 1] > first[q]) {
                double b[4];
                szo_pt cell[5];
                center_cell_coords(h, ix + 1, iy + 1, b);
                make_wall_ring(cell, b);
                for (int64_t k = first[q]; k < first[q + 1]; ++k) {
                    const cellrec *r = &h->crec[k];
                    int np = h->npts[r->floe];
                    if (np > capT) { capT = 2 * np; T = (szo_pt *)realloc(T, sizeof(szo_pt) * (size_t)capT); }
                    for (int v = 0; v < np; ++v) { T[v].x = h->ring[r->floe][v].x + r->dx; T[v].y = h->ring[r->floe][v].y + r->dy; }
                    szo_clip(cell, 5, T, np, &R);
                    double a = 0;
                    for (int g = 0; g < R.nreg; ++g) a += szo_ring_area(R.pts + R.off[g], R.off[g + 1] - R.off[g]);
                    if (a > 0) {
                        tx += (r->tx / (double)r->npts) * a;
                        ty += (r->ty / (double)r->npts) * a;
                        si += a;
                    }
                }
                if (si > 0) {
                    tx /= si; ty /= si;
                    si /= cell_area;
                }
            }
            double du = h->atm_u[q] - h->ocn_u[q], dv = h->atm_v[q] - h->ocn_v[q];
            double ocn_frac = 1 - si, norm = sqrt(du * du + dv * dv);
            tx += c->rho_a * c->Cd_ao * ocn_frac * norm * du;
            ty += c->rho_a * c->Cd_ao * ocn_frac * norm * dv;
            h->taux[q] = tx; h->tauy[q] = ty; h->sifrac[q] = si;
            h->ocn_hflx[q] = c->dt * c->k / (c->rho_i * c->L) * (h->ocn_temp[q] - h->atm_temp[q]);
        }
        szo_regions_free(&R);
        free(T);
    }
    free(first);
}

int32_t szo_step_coupling(sz_handle *h) {
    if (!h || !h->have_domain || !h->ocn_u) return fail(h, SZ_ERR_INVALID, "step_coupling before set_domain/set_fields");
    if (h->n != h->n_init) return fail(h, SZ_ERR_INVALID, "step_coupling with ghosts present (call remove_ghosts)");
    double t0 = now_ms();
    h->frec = (cellrec **)calloc((size_t)(h->n > 0 ? h->n : 1), sizeof(cellrec *));
    h->nfrec = (int *)calloc((size_t)(h->n > 0 ? h->n : 1), sizeof(int));
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < h->n; ++i) coupling_one_floe(h, i);
    /* grid.floe_locations / ocean.scells: gather the per-floe lists, sort by (cell, floe) */
    int64_t tot = 0;
    for (int64_t i = 0; i < h->n; ++i) tot += h->nfrec[i];
    free(h->crec);
    h->crec = (cellrec *)malloc(sizeof(cellrec) * (size_t)(tot > 0 ? tot : 1));
    h->n_crec = 0;
    for (int64_t i = 0; i < h->n; ++i) {
        for (int k = 0; k < h->nfrec[i]; ++k) h->crec[h->n_crec++] = h->frec[i][k];
        free(h->frec[i]);
    }
    free(h->frec); free(h->nfrec);
    h->frec = NULL; h->nfrec = NULL;
    qsort(h->crec, (size_t)h->n_crec, sizeof(cellrec), cmp_cellrec);
    if (h->cfg.two_way_coupling_on) two_way_coupling(h);
    h->ms[4] = now_ms() - t0;
    return SZ_OK;
}

/* ---- test hooks (oracle only; pin the index semantics against test_coupling.jl:276-462) ------------------- */
int32_t szo_test_center_cell_coords(sz_handle *h, int32_t ix, int32_t iy, double out[4]) {
    if (!h || !h->have_domain || h->Nx == 0) return SZ_ERR_INVALID;
    center_cell_coords(h, ix, iy, out);
    return SZ_OK;
}
/* feeds (xidx, yidx, tau) sequences of ONE floe through floe_to_grid_info! and leaves the result in the registry */
int32_t szo_test_floe_to_grid(sz_handle *h, int64_t floe, int32_t n, const int32_t *xidx, const int32_t *yidx,
                              const double *tx, const double *ty) {
    if (!h || !h->have_domain || h->Nx == 0) return SZ_ERR_INVALID;
    cellrec *rl = NULL;
    int nrl = 0, cap = 0;
    for (int k = 0; k < n; ++k) floe_to_grid_info(h, &rl, &nrl, &cap, floe - 1, xidx[k], yidx[k], tx[k], ty[k]);
    free(h->crec);
    h->crec = rl;
    h->n_crec = nrl;
    qsort(h->crec, (size_t)h->n_crec, sizeof(cellrec), cmp_cellrec);
    return SZ_OK;
}

/* ---- a19: state update (update_floe.jl:392-551) ----------------------------------------------- */
static void update_one_floe(sz_handle *h, int64_t i) {
    const sz_config *c = &h->cfg;
    double dt = (double)c->dt;
    uint32_t warn = 0;
    double cfx = h->cfx[i], cfy = h->cfy[i], ctrq = h->ctrq[i];
    /* calc_stress!, :392-414 */
    double s11 = 0, s12 = 0, s22 = 0;
    double xi_ = h->cx[i], yi_ = h->cy[i];
    int nr = h->rows[i].n;
    if (nr > 0) {
        for (int k = 0; k < nr; ++k) {
            double *r = h->rows[i].r + NROWF * k;
            s11 += (r[3] - xi_) * r[1];
            s12 += (r[4] - yi_) * r[1] + (r[3] - xi_) * r[2];
            s22 += (r[4] - yi_) * r[2];
        }
        s12 *= 0.5;
        double inv = 1 / (h->area[i] * h->height[i]);
        s11 *= inv; s12 *= inv; s22 *= inv;
    }
    double st[4] = {s11, s12, s12, s22};
    double lam = c->stress_lambda; /* stress_calculators.jl:118-122 */
    for (int k = 0; k < 4; ++k) {
        h->stress_accum[4 * i + k] = (1 - lam) * h->stress_accum[4 * i + k] + lam * st[k];
        h->stress_instant[4 * i + k] = st[k];
    }
    if (h->height[i] > c->max_floe_height) { h->height[i] = c->max_floe_height; warn |= SZ_WARN_HEIGHT_CAPPED; } /* :482-485 */
    while (fmax(fabs(cfx), fabs(cfy)) > h->mass[i] / (5 * dt)) { /* :487-491 */
        cfx = cfx / 10; cfy = cfy / 10; ctrq = ctrq / 10; warn |= SZ_WARN_FORCE_SCALED;
    }
    double hh = h->height[i]; /* :494-500 */
    double dh = h->hflx[i] / hh;
    double hfrac = (hh + dh) / hh;
    h->mass[i] *= hfrac; h->moment[i] *= hfrac; h->height[i] -= dh;
    hh = h->height[i];
    double Dx = 1.5 * dt * h->u[i] - 0.5 * dt * h->p_dxdt[i]; /* :503-506 */
    double Dy = 1.5 * dt * h->v[i] - 0.5 * dt * h->p_dydt[i];
    double Da = 1.5 * dt * h->xi[i] - 0.5 * dt * h->p_dalphadt[i];
    h->alpha[i] += Da;
    { /* _move_floe! / _move_poly, floe_utils.jl:74-93: p -> R p + ((R(-c) + c) + Δ) */
        double cx = h->cx[i], cy = h->cy[i], sn = sin(Da), cs = cos(Da);
        double tx = ((cs * (-cx) - sn * (-cy)) + cx) + Dx;
        double ty = ((sn * (-cx) + cs * (-cy)) + cy) + Dy;
        for (int k = 0; k < h->npts[i]; ++k) {
            double x = h->ring[i][k].x, y = h->ring[i][k].y;
            h->ring[i][k].x = (cs * x - sn * y) + tx;
            h->ring[i][k].y = (sn * x + cs * y) + ty;
        }
        h->cx[i] += Dx; h->cy[i] += Dy;
    }
    h->p_dxdt[i] = h->u[i]; h->p_dydt[i] = h->v[i]; h->p_dalphadt[i] = h->xi[i]; /* :509-511 */
    double dudt = (h->fxOA[i] + cfx) / h->mass[i]; /* :514-531 */
    double dvdt = (h->fyOA[i] + cfy) / h->mass[i];
    double frac = 1;
    double au = fabs(dt * dudt), av = fabs(dt * dvdt), lim = hh / 2;
    double sgu = (dudt > 0) - (dudt < 0), sgv = (dvdt > 0) - (dvdt < 0);
    if (au > lim && av > lim) {
        double f1 = (sgu * hh / (2 * dt)) / dudt, f2 = (sgv * hh / (2 * dt)) / dvdt;
        frac = f1 < f2 ? f1 : f2;
    } else if (au > lim && av < lim) frac = (sgu * hh / (2 * dt)) / dudt;
    else if (au < lim && av > lim) frac = (sgv * hh / (2 * dt)) / dvdt;
    if (frac != 1) { dudt = frac * dudt; dvdt = frac * dvdt; warn |= SZ_WARN_VELOCITY_LIMITED; }
    h->u[i] += 1.5 * dt * dudt - 0.5 * dt * h->p_dudt[i]; /* :532-535 */
    h->v[i] += 1.5 * dt * dvdt - 0.5 * dt * h->p_dvdt[i];
    h->p_dudt[i] = dudt; h->p_dvdt[i] = dvdt;
    double dxidt = (h->trqOA[i] + ctrq) / h->moment[i]; /* :537-545 */
    dxidt = frac * dxidt;
    double xi = h->xi[i] + 1.5 * dt * dxidt - 0.5 * dt * h->p_dxidt[i];
    if (fabs(xi) > c->maximum_xi) { xi = ((xi > 0) - (xi < 0)) * c->maximum_xi; warn |= SZ_WARN_XI_CLAMPED; }
    h->xi[i] = xi; h->p_dxidt[i] = dxidt;
    { /* calc_strain!, :425-453 (uses floe.u in the v terms, as the reference does) */
        double e11 = 0, e12 = 0, e22 = 0, x1 = 0, y1 = 0;
        for (int k = 0; k < h->npts[i]; ++k) {
            double x2 = h->ring[i][k].x + (-h->cx[i]), y2 = h->ring[i][k].y + (-h->cy[i]);
            if (k == 0) { x1 = x2; y1 = y2; continue; }
            double xd = x2 - x1, yd = y2 - y1;
            double r1 = sqrt(x1 * x1 + y1 * y1), r2 = sqrt(x2 * x2 + y2 * y2);
            double t1 = atan2(y1, x1), t2 = atan2(y2, x2);
            double u1 = h->u[i] - h->xi[i] * r1 * sin(t1), u2 = h->u[i] - h->xi[i] * r2 * sin(t2);
            double v1 = h->u[i] + h->xi[i] * r1 * cos(t1), v2 = h->u[i] + h->xi[i] * r2 * cos(t2);
            double ud = u2 - u1, vd = v2 - v1;
            e11 += ud * yd; e12 += ud * xd + vd * yd; e22 += vd * xd;
            x1 = x2; y1 = y2;
        }
        e12 *= 0.5;
        double den = 2 * h->area[i];
        h->strain[4 * i + 0] = e11 / den; h->strain[4 * i + 1] = e12 / den;
        h->strain[4 * i + 2] = e12 / den; h->strain[4 * i + 3] = e22 / den;
    }
    h->warn[i] = warn;
}

int32_t szo_step_floe_properties(sz_handle *h, int64_t tstep) {
    (void)tstep;
    if (!h) return SZ_ERR_INVALID;
    if (h->n != h->n_init) return fail(h, SZ_ERR_INVALID, "step_floe_properties with ghosts present");
    double t0 = now_ms();
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < h->n; ++i) update_one_floe(h, i);
    h->ms[5] = now_ms() - t0;
    return SZ_OK;
}

int32_t szo_step(sz_handle *h, int64_t tstep, int32_t do_coupling) {
    double t0 = now_ms();
    int32_t rc;
    if ((rc = szo_add_ghosts(h, NULL)) != SZ_OK) return rc;
    if ((rc = szo_step_collisions(h)) != SZ_OK) return rc;
    if ((rc = szo_remove_ghosts(h)) != SZ_OK) return rc;
    if (do_coupling) { if ((rc = szo_step_coupling(h)) != SZ_OK) return rc; } else h->ms[4] = 0;
    if ((rc = szo_step_floe_properties(h, tstep)) != SZ_OK) return rc;
    h->ms[6] = now_ms() - t0;
    return SZ_OK;
}

/* the host-buffer form of the timestep: no overlap to exploit on the CPU, the three calls back to back */
/* an overlap hint for the CUDA product; the oracle's step does its coupling in order */
/* sz_step_host_partial: NULL input fields keep the resident value, NULL output fields are not written */
/* the upload half of szo_step_host_partial (also used by the slab restatement, which exchanges the halo in between) */
int32_t szo_upload_partial(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in) {
    if (!h) return SZ_ERR_INVALID;
    if (in) {
        if (in->n != h->n || in->n_init != h->n_init) return fail(h, SZ_ERR_INVALID, "step_host_partial: floe count differs from the resident store");
#define UP(src, dst) if (in->src) for (int64_t i = 0; i < h->n; ++i) h->dst[i] = in->src[i];
        UP(centroid_x, cx) UP(centroid_y, cy) UP(height, height) UP(area, area) UP(mass, mass) UP(rmax, rmax) UP(moment, moment)
        UP(alpha, alpha) UP(u, u) UP(v, v) UP(xi, xi) UP(overarea, overarea) UP(p_dxdt, p_dxdt) UP(p_dydt, p_dydt) UP(p_dudt, p_dudt)
        UP(p_dvdt, p_dvdt) UP(p_dxidt, p_dxidt) UP(p_dalphadt, p_dalphadt) UP(status_tag, status)
        if (!do_coupling) { UP(fxOA, fxOA) UP(fyOA, fyOA) UP(trqOA, trqOA) UP(hflx_factor, hflx) }
#undef UP
        for (int64_t i = 0; i < h->n; ++i)
            for (int k = 0; k < 4; ++k) {
                if (in->stress_accum) h->stress_accum[4 * i + k] = in->stress_accum[4 * i + k];
                if (in->stress_instant) h->stress_instant[4 * i + k] = in->stress_instant[4 * i + k];
                if (in->strain) h->strain[4 * i + k] = in->strain[4 * i + k];
            }
        if (in->vert_xy) {
            int64_t vo = 0;
            for (int64_t i = 0; i < h->n; ++i) {
                memcpy(h->ring[i], in->vert_xy + 2 * vo, sizeof(szo_pt) * (size_t)h->npts[i]);
                vo += h->npts[i];
            }
        }
    }
    return SZ_OK;
}

int32_t szo_step_host_partial(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out) {
    if (!h || !out) return SZ_ERR_INVALID;
    int32_t rc = szo_upload_partial(h, do_coupling, in);
    if (rc != SZ_OK) return rc;
    rc = szo_step(h, tstep, do_coupling);
    if (rc != SZ_OK) return rc;
    sz_floe_soa o = *out;
    o.mc_x = o.mc_y = NULL;
    o.ghost_index = NULL;
    rc = szo_download_floes(h, &o);
    out->n = o.n;
    out->n_init = o.n_init;
    return rc;
}

int32_t szo_coupling_begin(sz_handle *h) { return h ? SZ_OK : SZ_ERR_INVALID; }
int32_t szo_upload_state_begin(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in) {
    (void)do_coupling;
    return szo_upload_state(h, in);
}
int32_t szo_step_host(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out) {
    int32_t rc;
    if (!h || !out) return SZ_ERR_INVALID;
    if (in && (rc = szo_upload_state(h, in)) != SZ_OK) return rc; /* in == NULL: szo_upload_state_begin did it */
    if ((rc = szo_step(h, tstep, do_coupling)) != SZ_OK) return rc;
    double *mx = out->mc_x, *my = out->mc_y;
    out->mc_x = out->mc_y = NULL; /* Monte-Carlo points are not transferred */
    rc = szo_download_floes(h, out);
    out->mc_x = mx;
    out->mc_y = my;
    return rc;
}


/* ---- services for the host-side processes ------------------------------ */
/* simplification.jl:98-116, welding.jl:119-150: potential_interaction, then sum(GO.area, intersect_polys(pi, pj)) */
int32_t szo_pair_overlap_areas(sz_handle *h, int64_t n_pairs, const int64_t *pairs, double *areas, uint8_t *interacts) {
    if (!h || n_pairs < 0 || (n_pairs > 0 && (!pairs || !areas))) return SZ_ERR_INVALID;
    for (int64_t k = 0; k < n_pairs; ++k)
        if (pairs[2 * k] < 1 || pairs[2 * k] > h->n || pairs[2 * k + 1] < 1 || pairs[2 * k + 1] > h->n)
            return fail(h, SZ_ERR_INVALID, "pair_overlap_areas: floe index out of range");
#pragma omp parallel
    {
        szo_regions R;
        szo_regions_init(&R);
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < n_pairs; ++k) {
            int64_t i = pairs[2 * k] - 1, j = pairs[2 * k + 1] - 1;
            int pi = potential_interaction(h, i, j);
            double a = 0.0;
            if (pi) {
                szo_clip(h->ring[i], h->npts[i], h->ring[j], h->npts[j], &R);
                for (int g = 0; g < R.nreg; ++g) a += szo_ring_area(R.pts + R.off[g], R.off[g + 1] - R.off[g]);
            }
            areas[k] = a;
            if (interacts) interacts[k] = (uint8_t)pi;
        }
        szo_regions_free(&R);
    }
    return SZ_OK;
}

/* ---- generate_subfloe_points, coupling.jl:172-208 (Monte Carlo), :235-321 (sub-grid) ---------- */
typedef struct { double *x, *y; int64_t n, cap; } ptlist;
static void pt_push(ptlist *l, double x, double y) {
    if (l->n == l->cap) {
        l->cap = l->cap ? 2 * l->cap : 256;
        l->x = (double *)realloc(l->x, sizeof(double) * (size_t)l->cap);
        l->y = (double *)realloc(l->y, sizeof(double) * (size_t)l->cap);
    }
    l->x[l->n] = x; l->y[l->n] = y; l->n++;
}

/* points of floe f in the body frame; returns the status tag */
static int generate_points_one(sz_handle *h, const sz_points_generator *g, int64_t f, ptlist *out) {
    const int np = h->npts[f];
    szo_pt *r = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)np);
    double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    for (int k = 0; k < np; ++k) { /* _translate_poly(poly, -cx, -cy), coupling.jl:189,261; GI.extent */
        r[k].x = h->ring[f][k].x + (-h->cx[f]);
        r[k].y = h->ring[f][k].y + (-h->cy[f]);
        xmin = fmin(xmin, r[k].x); xmax = fmax(xmax, r[k].x);
        ymin = fmin(ymin, r[k].y); ymax = fmax(ymax, r[k].y);
    }
    const double Dx = xmax - xmin, Dy = ymax - ymin;
    int status = SZ_STATUS_ACTIVE;
    out->n = 0;
    if (g->kind == SZ_POINTS_MONTE_CARLO) { /* :172-208 */
        const int n = g->npoints;
        int count = 1, used = 0;
        double err = 1.0;
        while (err > g->err) {
            if (count > 10) {
                err = 0.0;
                status = SZ_STATUS_REMOVE;
            } else {
                int64_t in = 0;
                for (int j = 0; j < n; ++j) {
                    szo_pt p = {xmin + Dx * szo_uniform(g->seed, h->id[f], count, j, 0), ymin + Dy * szo_uniform(g->seed, h->id[f], count, j, 1)};
                    in += szo_point_coveredby(p, r, np);
                }
                err = fabs((double)in / (double)n * (Dx * Dy) - h->area[f]) / h->area[f];
                used = count;
                count += 1;
            }
        }
        for (int j = 0; used > 0 && j < n; ++j) { /* mc_x[mc_in]: the draws of the last attempt, in draw order */
            szo_pt p = {xmin + Dx * szo_uniform(g->seed, h->id[f], used, j, 0), ymin + Dy * szo_uniform(g->seed, h->id[f], used, j, 1)};
            if (szo_point_coveredby(p, r, np)) pt_push(out, p.x, p.y);
        }
        if (out->n == 0) status = SZ_STATUS_REMOVE; /* :203-205 */
    } else { /* :235-321 */
        const double dg = g->delta_g;
        double x1 = r[0].x, y1 = r[0].y;
        for (int i = 1; i < np; ++i) {
            double x2 = r[i].x, y2 = r[i].y;
            double dx = x2 - x1, dy = y2 - y1;
            double l = sqrt(dx * dx + dy * dy);
            pt_push(out, x1, y1);
            const double x2u = x2, y2u = y2;
            if (l <= 2 * dg) {
                if (l > dg) pt_push(out, x1 + dx / 2, y1 + dy / 2);
            } else {
                if (dx == 0) {
                    double sg = (double)((dy > 0) - (dy < 0));
                    y1 += dg / 2 * sg;
                    y2 -= dg / 2 * sg;
                } else if (dy == 0) {
                    double sg = (double)((dx > 0) - (dx < 0));
                    x1 += dg / 2 * sg;
                    x2 -= dg / 2 * sg;
                } else { /* "shift points to still be on the line": x_shift is positive whatever the edge's direction */
                    double m = dy / dx;
                    double xs = sqrt(dg * dg / (4 * (1 + m * m)));
                    double ys = m * xs;
                    x1 += xs; x2 -= xs; y1 += ys; y2 -= ys;
                }
                l = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
                int64_t ne = (int64_t)ceil(l / dg) + 1;
                for (int64_t k = 0; k < ne; ++k) pt_push(out, szo_range_elem(x1, x2, k, ne), szo_range_elem(y1, y2, k, ne));
            }
            x1 = x2u; y1 = y2u;
        }
        int64_t nx = (int64_t)ceil((xmax - xmin) / dg), ny = (int64_t)ceil((ymax - ymin) / dg);
        const int xs1 = nx < 3, ys1 = ny < 3;
        if (xs1) nx = 1;
        if (ys1) ny = 1;
        for (int64_t k = 0; k < nx * ny; ++k) { /* repeat(x, ny), repeat(y, inner = nx) */
            szo_pt p = {xs1 ? 0.0 : szo_range_elem(xmin + dg / 2, xmax - dg / 2, k % nx, nx),
                        ys1 ? 0.0 : szo_range_elem(ymin + dg / 2, ymax - dg / 2, k / nx, ny)};
            if (szo_point_coveredby(p, r, np)) pt_push(out, p.x, p.y);
        }
    }
    free(r);
    return status;
}

int32_t szo_generate_subfloe_points(sz_handle *h, const sz_points_generator *g, int64_t n_floes, const int64_t *floes,
                                    int64_t *offsets, double *x, double *y, int64_t cap_points, int32_t *status, int32_t install) {
    if (!h || !g || !offsets || n_floes < 0) return SZ_ERR_INVALID;
    if (g->kind != SZ_POINTS_MONTE_CARLO && g->kind != SZ_POINTS_SUB_GRID) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: unknown generator");
    if (g->kind == SZ_POINTS_MONTE_CARLO ? g->npoints < 1 : !(g->delta_g > 0)) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: bad generator parameters");
    if (!floes && n_floes != h->n_init) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: floes == NULL means all n_init floes");
    if (install && floes) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: install needs the whole list (floes == NULL)");
    for (int64_t k = 0; floes && k < n_floes; ++k)
        if (floes[k] < 1 || floes[k] > h->n) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: floe index out of range");
    ptlist *lists = (ptlist *)calloc((size_t)(n_floes > 0 ? n_floes : 1), sizeof(ptlist));
    int *st = (int *)malloc(sizeof(int) * (size_t)(n_floes > 0 ? n_floes : 1));
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < n_floes; ++k) st[k] = generate_points_one(h, g, floes ? floes[k] - 1 : k, &lists[k]);
    offsets[0] = 0;
    for (int64_t k = 0; k < n_floes; ++k) offsets[k + 1] = offsets[k] + lists[k].n;
    int32_t rc = SZ_OK;
    if (x && y) {
        if (offsets[n_floes] > cap_points) rc = fail(h, SZ_ERR_CAPACITY, "generate_subfloe_points: output arrays too small");
        else
            for (int64_t k = 0; k < n_floes; ++k) {
                memcpy(x + offsets[k], lists[k].x, sizeof(double) * (size_t)lists[k].n);
                memcpy(y + offsets[k], lists[k].y, sizeof(double) * (size_t)lists[k].n);
            }
    }
    if (rc == SZ_OK && install)
        for (int64_t k = 0; k < n_floes; ++k) {
            free(h->mcx[k]); free(h->mcy[k]);
            h->nmc[k] = lists[k].n;
            h->mcx[k] = (double *)malloc(sizeof(double) * (size_t)(lists[k].n > 0 ? lists[k].n : 1));
            h->mcy[k] = (double *)malloc(sizeof(double) * (size_t)(lists[k].n > 0 ? lists[k].n : 1));
            memcpy(h->mcx[k], lists[k].x, sizeof(double) * (size_t)lists[k].n);
            memcpy(h->mcy[k], lists[k].y, sizeof(double) * (size_t)lists[k].n);
            if (st[k] == SZ_STATUS_REMOVE) h->status[k] = SZ_STATUS_REMOVE;
        }
    for (int64_t k = 0; k < n_floes; ++k) {
        if (status) status[k] = st[k];
        free(lists[k].x); free(lists[k].y);
    }
    free(lists); free(st);
    return rc;
}

/* calc_eulerian_data!, output.jl:794-919.  Topography (:826-829: cell_poly_list = cell box minus the topography polygons,
 * floe areas and si_frac measured on that list) is evaluated WITHOUT a polygon-difference operator, from intersections
 * only:  area(floe ∩ (cell ∖ topo)) = area(floe ∩ cell) − Σ_k area((floe ∩ cell) ∩ topo_k)  and
 * area(cell ∖ topo) = area(cell) − Σ_k area(cell ∩ topo_k), which is exact for topography elements that do not overlap one
 * another (they are separate land masses; overlapping elements would be subtracted twice).  A remainder below 1e-12 of
 * the uncut area counts as zero (the reference's list would be empty / the floe's piece absent). */
int32_t szo_eulerian_data(sz_handle *h, int32_t nx, int32_t ny, const double *xg, const double *yg, int32_t n_out,
                          const int32_t *kinds, double *data) {
    if (!h || nx < 1 || ny < 1 || !xg || !yg || n_out < 0 || (n_out > 0 && (!kinds || !data))) return SZ_ERR_INVALID;
    for (int k = 0; k < n_out; ++k)
        if (kinds[k] < 0 || kinds[k] >= SZ_GRID_NKINDS) return fail(h, SZ_ERR_INVALID, "eulerian_data: unknown output kind");
    const double dx = xg[1] - xg[0], dy = yg[1] - yg[0];      /* :796-797 */
    const double cell_rmax = sqrt(dx * dx + dy * dy);         /* :798 */
#pragma omp parallel
    {
        szo_regions R, R2;
        szo_regions_init(&R);
        szo_regions_init(&R2);
        int capf = 64, nf;
        int64_t *fidx = (int64_t *)malloc(sizeof(int64_t) * (size_t)capf);
        double *pic = (double *)malloc(sizeof(double) * (size_t)capf);
        int *tk = (int *)malloc(sizeof(int) * (size_t)(h->n_topo > 0 ? h->n_topo : 1));
#pragma omp for schedule(dynamic, 4) collapse(2)
        for (int i = 0; i < ny; ++i) {
            for (int j = 0; j < nx; ++j) {
                const double xc = xg[j] + 0.5 * dx, yc = yg[i] + 0.5 * dy; /* :799-802 */
                double b[4] = {xg[j], xg[j + 1], yg[i], yg[i + 1]};
                szo_pt cell[5];
                make_wall_ring(cell, b);                      /* _make_bounding_box_polygon, :825 */
                /* topography elements that can reach this cell, and the cell's area without them (:826-829) */
                int ntk = 0;
                double cell_area = szo_ring_area(cell, 5), cut = 0.0;
                for (int k = 0; k < h->n_topo; ++k) {
                    double ddx = xc - h->topo_cx[k], ddy = yc - h->topo_cy[k];
                    if (!(sqrt(ddx * ddx + ddy * ddy) < h->topo_rmax[k] + cell_rmax)) continue;
                    szo_clip(cell, 5, h->topo_ring[k], h->topo_np[k], &R);
                    double a = 0.0;
                    for (int g = 0; g < R.nreg; ++g) a += szo_ring_area(R.pts + R.off[g], R.off[g + 1] - R.off[g]);
                    tk[ntk++] = k;
                    cut += a;
                }
                if (!(cut > 0)) ntk = 0;                              /* nothing to cut out of this cell */
                double cell_free = cell_area - cut;
                if (ntk > 0 && !(cell_free > 1e-12 * cell_area)) {   /* length(cell_poly_list) == 0, :831-834 */
                    for (int k = 0; k < n_out; ++k) data[(size_t)j + (size_t)nx * ((size_t)i + (size_t)ny * (size_t)k)] = 0.0;
                    continue;
                }
                nf = 0;
                for (int64_t f = 0; f < h->n; ++f) {          /* mask :808-818, areas :838-846 */
                    double ddx = xc - h->cx[f], ddy = yc - h->cy[f];
                    double pint = sqrt(ddx * ddx + ddy * ddy) - (h->rmax[f] + cell_rmax);
                    if (!(pint < 0)) continue;
                    szo_clip(h->ring[f], h->npts[f], cell, 5, &R);
                    double a = 0.0;
                    for (int g = 0; g < R.nreg; ++g) a += szo_ring_area(R.pts + R.off[g], R.off[g + 1] - R.off[g]);
                    if (ntk > 0 && a > 0) {
                        double sub = 0.0;
                        for (int q = 0; q < ntk; ++q)
                            for (int g = 0; g < R.nreg; ++g) {
                                szo_clip(R.pts + R.off[g], R.off[g + 1] - R.off[g], h->topo_ring[tk[q]], h->topo_np[tk[q]], &R2);
                                for (int g2 = 0; g2 < R2.nreg; ++g2) sub += szo_ring_area(R2.pts + R2.off[g2], R2.off[g2 + 1] - R2.off[g2]);
                            }
                        a = (a - sub > 1e-12 * a) ? a - sub : 0.0;
                    }
                    if (a > 0) {                              /* :848-849 */
                        if (nf == capf) {
                            capf *= 2;
                            fidx = (int64_t *)realloc(fidx, sizeof(int64_t) * (size_t)capf);
                            pic = (double *)realloc(pic, sizeof(double) * (size_t)capf);
                        }
                        fidx[nf] = f;
                        pic[nf] = a;
                        nf++;
                    }
                }
                double area_tot = 0.0, mass_tot = 0.0;        /* :854-856 */
                for (int q = 0; q < nf; ++q) {
                    area_tot += pic[q];
                    mass_tot += h->mass[fidx[q]] * (pic[q] / h->area[fidx[q]]);
                }
                double acc[SZ_GRID_NKINDS];
                for (int k = 0; k < SZ_GRID_NKINDS; ++k) acc[k] = 0.0;
                if (mass_tot > 0) {                           /* :858-905 */
                    double over = 0.0;
                    for (int q = 0; q < nf; ++q) {
                        int64_t f = fidx[q];
                        double ma = (pic[q] / h->area[f]) * (h->mass[f] / mass_tot);
                        const double *sa = h->stress_accum + 4 * f, *st = h->strain + 4 * f;
                        acc[SZ_GRID_U] += h->u[f] * ma;
                        acc[SZ_GRID_V] += h->v[f] * ma;
                        acc[SZ_GRID_DUDT] += h->p_dudt[f] * ma;
                        acc[SZ_GRID_DVDT] += h->p_dvdt[f] * ma;
                        acc[SZ_GRID_HEIGHT] += h->height[f] * ma;
                        acc[SZ_GRID_STRESS_XX] += sa[0] * ma;  /* s[1,1] */
                        acc[SZ_GRID_STRESS_YX] += sa[2] * ma;  /* s[1,2] */
                        acc[SZ_GRID_STRESS_XY] += sa[1] * ma;  /* s[2,1] */
                        acc[SZ_GRID_STRESS_YY] += sa[3] * ma;  /* s[2,2] */
                        acc[SZ_GRID_STRAIN_UX] += st[0] * ma;
                        acc[SZ_GRID_STRAIN_VX] += st[2] * ma;
                        acc[SZ_GRID_STRAIN_UY] += st[1] * ma;
                        acc[SZ_GRID_STRAIN_VY] += st[3] * ma;
                        over += h->overarea[f];
                    }
                    acc[SZ_GRID_SI_FRAC] = area_tot / cell_free;
                    acc[SZ_GRID_OVERAREA] = over / (double)nf;
                    acc[SZ_GRID_MASS] = mass_tot;
                    acc[SZ_GRID_AREA] = area_tot;
                    {   /* maximum(eigvals([xx yx; xy yy])), zeroed when |.| > 1e8 (:884-893) */
                        double xx = acc[SZ_GRID_STRESS_XX], yx = acc[SZ_GRID_STRESS_YX], xy = acc[SZ_GRID_STRESS_XY],
                               yy = acc[SZ_GRID_STRESS_YY];
                        double hm = 0.5 * (xx + yy), hd = 0.5 * (xx - yy), disc = hd * hd + yx * xy;
                        double e = disc > 0 ? hm + sqrt(disc) : hm;
                        if (fabs(e) > 1e8) e = 0.0;
                        acc[SZ_GRID_STRESS_EIG] = e;
                    }
                }
                for (int k = 0; k < n_out; ++k) data[(size_t)j + (size_t)nx * ((size_t)i + (size_t)ny * (size_t)k)] = acc[kinds[k]];
            }
        }
        free(fidx);
        free(pic);
        free(tk);
        szo_regions_free(&R);
        szo_regions_free(&R2);
    }
    return SZ_OK;
}

/* ---- results ----------------------------------------------------------------------------------- */
int32_t szo_get_interactions(sz_handle *h, int64_t *offsets, double *rows) {
    if (!h || !offsets) return SZ_ERR_INVALID;
    int64_t o = 0;
    for (int64_t i = 0; i < h->n; ++i) {
        offsets[i] = o;
        if (rows && h->rows[i].n) memcpy(rows + NROWF * o, h->rows[i].r, sizeof(double) * NROWF * (size_t)h->rows[i].n);
        o += h->rows[i].n;
    }
    offsets[h->n] = o;
    return SZ_OK;
}

int32_t szo_set_interactions(sz_handle *h, const int64_t *offsets, const double *rows) {
    if (!h || !offsets) return SZ_ERR_INVALID;
    for (int64_t i = 0; i < h->n; ++i) {
        int cnt = (int)(offsets[i + 1] - offsets[i]);
        rowlist *l = &h->rows[i];
        if (cnt > l->cap) { l->cap = cnt; l->r = (double *)realloc(l->r, sizeof(double) * NROWF * (size_t)cnt); }
        l->n = cnt;
        if (cnt) memcpy(l->r, rows + NROWF * offsets[i], sizeof(double) * NROWF * (size_t)cnt);
    }
    return SZ_OK;
}

int32_t szo_get_pairs(sz_handle *h, int32_t which, int64_t *pairs) {
    if (!h || !pairs) return SZ_ERR_INVALID;
    const int64_t *src; int64_t m;
    switch (which) {
    case 0: src = h->cand; m = h->n_cand; break;
    case 1: src = h->pairs; m = h->n_pairs; break;
    case 2: src = h->overlap; m = h->n_overlap; break;
    case 3: src = h->fuse; m = h->n_fuse; break;
    default: return fail(h, SZ_ERR_INVALID, "get_pairs: which must be 0..3");
    }
    for (int64_t k = 0; k < 2 * m; ++k) pairs[k] = src[k] + 1;
    return SZ_OK;
}

int32_t szo_get_warnings(sz_handle *h, uint32_t *bits) {
    if (!h || !bits) return SZ_ERR_INVALID;
    for (int64_t i = 0; i < h->n_init; ++i) bits[i] = h->warn[i];
    return SZ_OK;
}

int32_t szo_get_timings(sz_handle *h, double ms[8]) {
    if (!h || !ms) return SZ_ERR_INVALID;
    memcpy(ms, h->ms, sizeof(double) * 8);
    return SZ_OK;
}

/* ---- halo exchange (same layout as the product: 8 doubles per floe, then the ring points) ------------- */
int32_t szo_halo_configure(sz_handle *h, int32_t n_lists, const int64_t *off, const int64_t *idx) {
    if (!h || n_lists < 0 || (n_lists > 0 && (!off || !idx))) return SZ_ERR_INVALID;
    free(h->hl_off); free(h->hl_idx);
    h->n_lists = n_lists;
    int64_t tot = n_lists > 0 ? off[n_lists] : 0;
    h->hl_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_lists + 1));
    h->hl_idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(tot > 0 ? tot : 1));
    h->hl_off[0] = 0;
    for (int k = 0; k < n_lists; ++k) h->hl_off[k + 1] = off[k + 1];
    for (int64_t k = 0; k < tot; ++k) {
        if (idx[k] < 1 || idx[k] > h->n_init) return fail(h, SZ_ERR_INVALID, "halo_configure: index out of range");
        h->hl_idx[k] = idx[k] - 1;
    }
    return SZ_OK;
}

int32_t szo_halo_bytes(sz_handle *h, int32_t list, int64_t *bytes) {
    if (!h || !bytes || list < 0 || list >= h->n_lists) return SZ_ERR_INVALID;
    int64_t b = 0;
    for (int64_t k = h->hl_off[list]; k < h->hl_off[list + 1]; ++k) b += 64 + 16 * (int64_t)h->npts[h->hl_idx[k]];
    *bytes = b;
    return SZ_OK;
}

int32_t szo_halo_pack(sz_handle *h, int32_t list, void *dst, int64_t cap) {
    int64_t need;
    if (!h || !dst || szo_halo_bytes(h, list, &need) != SZ_OK) return SZ_ERR_INVALID;
    if (need > cap) return fail(h, SZ_ERR_CAPACITY, "halo_pack: buffer too small");
    int64_t n = h->hl_off[list + 1] - h->hl_off[list];
    double *rec = (double *)dst, *vx = rec + 8 * n;
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = h->hl_idx[h->hl_off[list] + k];
        double *r = rec + 8 * k;
        r[0] = h->cx[i]; r[1] = h->cy[i]; r[2] = h->u[i]; r[3] = h->v[i]; r[4] = h->xi[i];
        r[5] = h->height[i]; r[6] = (double)h->status[i]; r[7] = h->alpha[i];
        memcpy(vx, h->ring[i], sizeof(szo_pt) * (size_t)h->npts[i]);
        vx += 2 * h->npts[i];
    }
    return SZ_OK;
}

int32_t szo_halo_unpack(sz_handle *h, int32_t list, const void *src, int64_t bytes) {
    int64_t need;
    if (!h || !src || szo_halo_bytes(h, list, &need) != SZ_OK) return SZ_ERR_INVALID;
    if (need != bytes) return fail(h, SZ_ERR_INVALID, "halo_unpack: size differs from the configured list");
    int64_t n = h->hl_off[list + 1] - h->hl_off[list];
    const double *rec = (const double *)src, *vx = rec + 8 * n;
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = h->hl_idx[h->hl_off[list] + k];
        const double *r = rec + 8 * k;
        h->cx[i] = r[0]; h->cy[i] = r[1]; h->u[i] = r[2]; h->v[i] = r[3]; h->xi[i] = r[4];
        h->height[i] = r[5]; h->status[i] = (int32_t)r[6]; h->alpha[i] = r[7];
        memcpy(h->ring[i], vx, sizeof(szo_pt) * (size_t)h->npts[i]);
        vx += 2 * h->npts[i];
    }
    return SZ_OK;
}

/* host buffers: nothing to order, `stream` is ignored */
int32_t szo_halo_pack_on(sz_handle *h, int32_t list, void *dst, int64_t cap, void *stream) {
    (void)stream;
    return szo_halo_pack(h, list, dst, cap);
}
int32_t szo_halo_unpack_on(sz_handle *h, int32_t list, const void *src, int64_t bytes, void *stream) {
    (void)stream;
    return szo_halo_unpack(h, list, src, bytes);
}

int32_t szo_clip_polygons(sz_handle *h, const double *p_xy, int32_t np, const double *q_xy, int32_t nq,
                          int32_t cap_regions, int32_t cap_points, int32_t *out_offsets, double *out_xy,
                          double *out_areas) {
    (void)h;
    szo_regions R;
    szo_regions_init(&R);
    int n = szo_clip((const szo_pt *)p_xy, np, (const szo_pt *)q_xy, nq, &R);
    if (n > cap_regions || (n > 0 && R.off[n] > cap_points)) { szo_regions_free(&R); return SZ_ERR_CAPACITY; }
    out_offsets[0] = 0;
    for (int r = 0; r < n; ++r) {
        out_offsets[r + 1] = R.off[r + 1];
        if (out_areas) out_areas[r] = szo_ring_area(R.pts + R.off[r], R.off[r + 1] - R.off[r]);
    }
    if (n > 0) memcpy(out_xy, R.pts, sizeof(szo_pt) * (size_t)R.off[n]);
    int failed = R.failed;
    szo_regions_free(&R);
    return failed ? -100 : n;
}
