// sz_api.cu — the C ABI of include/subzero_b200.h on top of the sm_100a kernels.
//
// Host side only: handle lifetime, device-store allocation / growth, host<->device marshalling
// of the reference's Floe struct-of-arrays, stream / event plumbing and the capacity-overflow
// retry loop.  No numerical work of the hot path happens here and there is no CPU fallback: when
// no CUDA device is usable sz_create fails with SZ_ERR_CUDA.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <unistd.h>

#include "sz_common.cuh"
#include "sz_slab_backend.h"

#define NEV 8
#define SZ_GRAPH_MAX_FLOES 32768  // sz_step replays a CUDA graph up to this field size (see step_impl)

struct FloeArr {
    void **ptr;
    size_t elem;  // bytes per floe
};

// grow-only device scratch of the service calls (sz_pair_overlap_areas, sz_eulerian_data)
struct SvcBuf {
    int2 *pairs; double *area; unsigned char *inter; int *big;
    double *xg, *yg, *data, *rec_area;
    int *rec_count, *rec_off, *cell_start, *rec_floe, *rec_cell, *val_in, *val_out;
    unsigned long long *key_in, *key_out;
    unsigned char *sort_tmp;
    double *cell_free; unsigned char *cell_topo; size_t cap_cfree, cap_ctopo;
    size_t cap_pairs, cap_area, cap_inter, cap_big, cap_xg, cap_yg, cap_data, cap_ra, cap_rc, cap_ro, cap_cs, cap_rf, cap_rcell,
        cap_vi, cap_vo, cap_ki, cap_ko, cap_sort;
};

// Host buffers of one sz_step_host call (see enqueue_uploads / enqueue_downloads below).
struct HostIO {
    const sz_floe_soa *in;
    sz_floe_soa *out;
};

// slab decomposition (sz_slab_*): this rank's exchange lists, its receive arena and the partners' arenas
struct SlabState {
    bool on;
    int rank, n_partners;
    SlabDev dev;
    int *d_send_idx, *d_recv_idx;
    long long *d_send_voff, *d_recv_voff;
    unsigned char *d_owned;
    double *d_refx, *d_refy;
    unsigned char *arena;      // flags + double-buffered staging, written by the partners
    size_t arena_bytes;
    // partners' arenas mapped from other processes, kept across rebuilds: (pid, arena generation) -> mapping.  An arena
    // lives until it is too small; the mapping of a replaced arena is closed when its successor is connected, and the
    // owner frees a replaced arena only one rebuild later (cudaFree before the importers closed is undefined)
    struct IpcMap { int pid; unsigned long long gen; void *ptr; };
    std::vector<IpcMap> ipc_open;
    unsigned long long arena_gen;
    std::vector<unsigned char *> retired;
    std::vector<long long> send_bytes;  // per partner
    int epoch;      // epoch the next step consumes
    int pushed;     // last epoch published to the partners
    int max_send, max_recv;
    long long send_bytes_total;
    long long cap_lists, cap_owned;  // capacities of the grow-only list arrays
};

// the step in flight between step_enqueue and step_finish
struct StepCur {
    int32_t do_coupling;
    HostIO io;
    bool has_io;
    bool keep_ghosts;       // repeat from the collisions, the ghosts of the failed attempt stay
    bool coupling_only;     // repeat from the (two-way) coupling: the collisions of this step are complete
    bool slab_host_mode;    // slab rank stepping on host arrays: publish at the start, not behind the update
    bool partial;           // sz_step_host_partial: NULL input fields keep their device-resident values
    int attempt;
};

struct sz_handle {
    sz_config cfg;
    SlabState slab;
    StepCur cur;
    SvcBuf svc;
    char err[512];
    Launch L;
    cudaStream_t stream2;      // coupling runs here, concurrently with the collision kernels (sz_step)
    cudaEvent_t ev_fork, ev_join, ev_c0, ev_c1;
    // sz_step_host: host -> device copies on stream_up and device -> host copies on stream_dn overlap the kernels
    cudaStream_t stream_up, stream_dn;
    cudaEvent_t ev_up[4], ev_dn[3], ev_up_start, ev_dn_end, ev_halo;
    double2 *d_cf_dn;  // [n][2] staging of collision_force in the host layout
    // sz_step replays a captured CUDA graph of the whole timestep (39 launches, two streams); `gen` counts every
    // change of a device pointer / parameter baked into the kernel nodes and forces a new capture
    cudaGraphExec_t gexec;
    unsigned long long gen, gkey_gen;
    int gkey_coupling, gkey_floes, gkey_pairs, graph_launches;
    bool graph_off;
    int graph_max_floes;
    double2 *mc_spare;              // second Monte-Carlo array of a slab rank (rebuilds gather into it and swap)
    long long mc_spare_cap, mc_off_cap;
    double *rb_tx, *rb_ty;          // grow-only scratch of a rebuild: migrants' points, per-floe source offsets
    double2 *rb_extra;
    long long *rb_src;
    long long rb_extra_cap, rb_src_cap;
    bool next_partial;              // the step being set up is a sz_step_host_partial
    unsigned long long tables_gen;  // finish_host_tables: the offset tables were last written for this floe list ...
    const void *tables_ptr[3];      // ... into these caller arrays
    long long tables_mid[2];
    int up_pending;  // sz_upload_state_begin ran: 1 + do_coupling, 0 = none
    bool cpl_prelaunched;  // sz_coupling_begin started this step's coupling kernel on stream2
    int cf_cap;
    Params P;
    bool have_grid, have_fields, have_domain, have_floes;
    DomainDev hD;
    Store S;
    StepBuf B;
    CouplingBuf CB;
    int n_crec_host;
    std::vector<FloeArr> floe_arrays;   // persistent per-floe arrays (grown with a copy)
    std::vector<void **> floe_scratch;  // int [cap_floes+1] scratch arrays (grown without)
    int n_init, n_total, n_verts, n_verts_init;
    long long n_mc;
    int n_topo, topo_verts_n;
    size_t field_n;
    Counters *h_cnt;  // pinned mirror
    Counters last;    // counters of the last collision step
    Counters last_ok_collisions;  // ... kept while a coupling-only repair repeats the rest of the step
    int n_rows_host;
    cudaEvent_t ev[NEV];
    double ms[8];
    // halo lists (slab decomposition)
    int n_lists;
    std::vector<long long> hl_off, hl_bytes;
    int *d_hl_idx;
    long long *d_hl_voff;
    // host mirrors of the static tables of the last upload (fast download path)
    std::vector<int> h_vstart, h_vcount;
    std::vector<long long> h_mc_off;
    bool ghosts_uploaded;
};

#define CK(expr)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                     __FILE__, __LINE__);                                                                \
            return SZ_ERR_CUDA;                                                                          \
        }                                                                                                \
    } while (0)

static int32_t fail(sz_handle *h, int32_t code, const char *msg) {
    if (h) snprintf(h->err, sizeof(h->err), "%s", msg);
    return code;
}

template <typename T>
static cudaError_t dalloc(T **p, size_t count) {
    *p = nullptr;
    return cudaMalloc((void **)p, sizeof(T) * std::max<size_t>(count, 1));
}
template <typename T>
static void dfree(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

extern "C" void sz_default_config(sz_config *c) {
    memset(c, 0, sizeof(*c));
    // Constants(), simulation.jl:5-18
    c->rho_o = 1027.0; c->rho_a = 1.2; c->Cd_io = 3e-3; c->Cd_ia = 1e-3; c->Cd_ao = 1.25e-3;
    c->f = 1.4e-4; c->turn_theta = 15.0 * 3.14159265358979323846 / 180.0; c->L = 2.93e5; c->k = 2.14;
    c->nu = 0.3; c->mu = 0.2; c->E = 6e6;
    // CollisionSettings(), process_settings.jl:183-187
    c->floe_floe_max_overlap = 0.55; c->floe_domain_max_overlap = 0.75;
    // FloeSettings(), process_settings.jl:20-32; DecayAreaScaledCalculator, stress_calculators.jl:82
    c->rho_i = 920.0; c->max_floe_height = 10.0; c->maximum_xi = 1e-5; c->stress_lambda = 0.2;
    c->coupling_dd = 1; c->two_way_coupling_on = 0; c->dt = 10; c->device = 0;
    c->max_regions_per_pair = 4; c->max_pairs_per_floe = 24;
}

extern "C" const char *sz_version(void) { return "subzero-b200 0.1 (CUDA sm_100a)"; }
extern "C" const char *sz_last_error(sz_handle *h) { return h ? h->err : "null handle"; }

static void register_arrays(sz_handle *h) {
    Store &S = h->S;
    StepBuf &B = h->B;
    auto add = [&](void **p, size_t e) { h->floe_arrays.push_back({p, e}); };
#define D1(f) add((void **)&S.f, sizeof(double));
    D1(cx) D1(cy) D1(height) D1(area) D1(mass) D1(rmax) D1(moment) D1(alpha) D1(u) D1(v) D1(xi) D1(fxOA) D1(fyOA)
    D1(trqOA) D1(hflx) D1(overarea) D1(cfx) D1(cfy) D1(ctrq) D1(p_dxdt) D1(p_dydt) D1(p_dudt) D1(p_dvdt) D1(p_dxidt)
    D1(p_dalphadt)
#undef D1
    add((void **)&S.stress_accum, 4 * sizeof(double));
    add((void **)&S.stress_instant, 4 * sizeof(double));
    add((void **)&S.strain, 4 * sizeof(double));
    add((void **)&S.status, sizeof(int));
    add((void **)&S.id, sizeof(long long));
    add((void **)&S.ghost_id, sizeof(long long));
    add((void **)&S.parent, sizeof(int));
    add((void **)&S.nghost, sizeof(int));
    add((void **)&S.ghost_slot, SZ_MAX_GHOSTS * sizeof(int));
    add((void **)&S.warn, sizeof(uint32_t));
    add((void **)&S.cpl_remove, sizeof(unsigned char));
    add((void **)&S.mc_r, sizeof(double));
    add((void **)&S.vstart, sizeof(int));
    add((void **)&S.vcount, sizeof(int));
    void **scr[] = {(void **)&B.cell_of, (void **)&B.cell_items, (void **)&B.up_count, (void **)&B.up_off,
                    (void **)&B.low_count, (void **)&B.low_off, (void **)&B.dom_count, (void **)&B.dom_off,
                    (void **)&B.row_pre, (void **)&B.row_count, (void **)&B.row_off, (void **)&B.g_flag,
                    (void **)&B.g_cnt, (void **)&B.g_off, (void **)&B.g_vcnt, (void **)&B.g_voff, (void **)&B.g_list};
    for (void **p : scr) h->floe_scratch.push_back(p);
}

// grow every per-floe array to new_cap floes, keeping the first `keep` entries
static int32_t grow_floes(sz_handle *h, int new_cap, int keep) {
    h->gen++;
    for (FloeArr &a : h->floe_arrays) {
        void *np = nullptr;
        CK(cudaMalloc(&np, a.elem * (size_t)std::max(new_cap, 1)));
        if (keep > 0 && *a.ptr) CK(cudaMemcpyAsync(np, *a.ptr, a.elem * (size_t)keep, cudaMemcpyDeviceToDevice, h->L.stream));
        CK(cudaStreamSynchronize(h->L.stream));
        if (*a.ptr) cudaFree(*a.ptr);
        *a.ptr = np;
    }
    for (void **p : h->floe_scratch) {
        void *np = nullptr;
        CK(cudaMalloc(&np, sizeof(int) * ((size_t)new_cap + 2)));
        if (p == (void **)&h->B.row_off) {
            CK(cudaMemsetAsync(np, 0, sizeof(int) * ((size_t)new_cap + 2), h->L.stream));
            if (keep > 0 && *p) CK(cudaMemcpyAsync(np, *p, sizeof(int) * ((size_t)keep + 1), cudaMemcpyDeviceToDevice, h->L.stream));
            CK(cudaStreamSynchronize(h->L.stream));
        }
        if (*p) cudaFree(*p);
        *p = np;
    }
    h->S.cap_floes = new_cap;
    // grid cells and scan scratch follow the floe capacity
    int cells = 4 * new_cap + 64;
    dfree(h->B.cell_count); dfree(h->B.cell_start); dfree(h->B.cell_fill); dfree(h->B.scan_block); dfree(h->B.cell_circ);
    CK(dalloc(&h->B.cell_circ, 2 * (size_t)new_cap + 2));
    CK(dalloc(&h->B.cell_count, (size_t)cells + 2));
    CK(dalloc(&h->B.cell_start, (size_t)cells + 2));
    CK(dalloc(&h->B.cell_fill, (size_t)cells + 2));
    CK(dalloc(&h->B.scan_block, (size_t)cells / 4096 + 64));  // also holds three floe-length scans side by side (scan_excl3)
    h->B.cap_cells = cells;
    dfree(h->B.lb_desc); dfree(h->B.lb_ticket); dfree(h->B.nb_scratch);
    h->B.lb_stride = cells / 4096 + 8;  // tiles of the longest scan (cells >= floes)
    CK(dalloc(&h->B.lb_desc, (size_t)12 * h->B.lb_stride));  // 4 scan slots (cells, neighbour counts, rows, ghosts) x 3 arrays
    CK(dalloc(&h->B.lb_ticket, 16));
    CK(dalloc(&h->B.nb_scratch, (size_t)20 * new_cap));  // NB_K rows
    CK(cudaMemset(h->B.lb_desc, 0, sizeof(unsigned long long) * 12 * h->B.lb_stride));
    CK(cudaMemset(h->B.lb_ticket, 0, sizeof(int) * 16));
    return SZ_OK;
}

static int32_t grow_verts(sz_handle *h, int new_cap, int keep) {
    h->gen++;
    double2 *np = nullptr;
    CK(dalloc(&np, (size_t)new_cap));
    if (keep > 0 && h->S.verts) CK(cudaMemcpy(np, h->S.verts, sizeof(double2) * (size_t)keep, cudaMemcpyDeviceToDevice));
    dfree(h->S.verts);
    h->S.verts = np;
    h->S.cap_verts = new_cap;
    return SZ_OK;
}

static int32_t set_pair_cap(sz_handle *h, int cap_pairs, int cap_dom) {
    h->gen++;
    StepBuf &B = h->B;
    dfree(B.pair_i); dfree(B.pair_j); dfree(B.low_pair); dfree(B.keep); dfree(B.dom_floe); dfree(B.dom_elem);
    dfree(B.item_nrows); dfree(B.item_row0); dfree(B.item_flags); dfree(B.large_items); dfree(B.mid_items);
    dfree(B.order); dfree(B.order_cls);
    CK(dalloc(&B.pair_i, (size_t)cap_pairs));
    CK(dalloc(&B.pair_j, (size_t)cap_pairs));
    CK(dalloc(&B.low_pair, (size_t)cap_pairs));
    CK(dalloc(&B.keep, (size_t)cap_pairs));
    CK(dalloc(&B.dom_floe, (size_t)cap_dom));
    CK(dalloc(&B.dom_elem, (size_t)cap_dom));
    size_t items = (size_t)cap_pairs + cap_dom;
    CK(dalloc(&B.item_nrows, items));
    CK(dalloc(&B.item_row0, items));
    CK(dalloc(&B.item_flags, items));
    CK(dalloc(&B.large_items, items));
    CK(dalloc(&B.mid_items, items));
    CK(dalloc(&B.order, items));
    CK(dalloc(&B.order_cls, items));
    if (!B.class_count) {
        CK(dalloc(&B.class_count, 3 * 128));
        CK(cudaMemset(B.class_count, 0, sizeof(int) * 3 * 128));
        B.class_base = B.class_count + 128;
        B.class_cursor = B.class_count + 256;
    }
    B.cap_pairs = cap_pairs;
    B.cap_dom = cap_dom;
    return SZ_OK;
}
static int32_t set_pool_cap(sz_handle *h, int cap) {
    h->gen++;
    dfree(h->B.pool); dfree(h->B.force_items); dfree(h->B.force_meta); dfree(h->B.force_pts);
    CK(dalloc(&h->B.pool, (size_t)cap * NPOOL));
    CK(dalloc(&h->B.force_items, (size_t)cap));
    CK(dalloc(&h->B.force_meta, (size_t)cap));
    CK(dalloc(&h->B.force_pts, (size_t)cap * 28));  // TN_PRE_PTS records, see sz_narrow_thread.cuh
    h->B.cap_pool = cap;
    h->B.cap_force = cap;
    return SZ_OK;
}
static int32_t set_row_cap(sz_handle *h, int cap) {
    h->gen++;
    dfree(h->B.rows);
    CK(dalloc(&h->B.rows, (size_t)cap * NCOL));
    h->B.cap_rows = cap;
    return SZ_OK;
}
static int32_t set_crec_cap(sz_handle *h, int cap) {
    h->gen++;
    CouplingBuf &C = h->CB;
    dfree(C.rec_cell); dfree(C.rec_floe); dfree(C.rec_npts); dfree(C.rec_t); dfree(C.rec_d); dfree(C.rec_area); dfree(C.perm); dfree(C.big_recs);
    CK(dalloc(&C.rec_cell, (size_t)cap)); CK(dalloc(&C.rec_floe, (size_t)cap)); CK(dalloc(&C.rec_npts, (size_t)cap));
    CK(dalloc(&C.rec_t, (size_t)cap)); CK(dalloc(&C.rec_d, (size_t)cap)); CK(dalloc(&C.rec_area, (size_t)cap));
    CK(dalloc(&C.perm, (size_t)cap)); CK(dalloc(&C.big_recs, (size_t)cap));
    C.cap_crec = cap;
    return SZ_OK;
}
static int32_t set_spill_cap(sz_handle *h, int cap) {
    h->gen++;
    CouplingBuf &C = h->CB;
    dfree(C.sp_cell); dfree(C.sp_n); dfree(C.sp_sd); dfree(C.sp_t);
    CK(dalloc(&C.sp_cell, (size_t)cap)); CK(dalloc(&C.sp_n, (size_t)cap)); CK(dalloc(&C.sp_sd, (size_t)cap)); CK(dalloc(&C.sp_t, (size_t)cap));
    C.cap_spill = cap;
    return SZ_OK;
}
static int32_t set_fuse_cap(sz_handle *h, int cap) {
    h->gen++;
    dfree(h->B.fuse_pairs);
    CK(dalloc(&h->B.fuse_pairs, (size_t)cap));
    h->B.cap_fuse = cap;
    return SZ_OK;
}

extern "C" int32_t sz_create(const sz_config *cfg, sz_handle **out) {
    if (!cfg || !out) return SZ_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) return SZ_ERR_CUDA;
    if (cudaSetDevice(cfg->device) != cudaSuccess) return SZ_ERR_CUDA;
    sz_handle *h = new (std::nothrow) sz_handle();
    if (!h) return SZ_ERR_NOMEM;
    memset(&h->S, 0, sizeof(h->S));
    memset(&h->B, 0, sizeof(h->B));
    memset(&h->CB, 0, sizeof(h->CB));
    memset(&h->svc, 0, sizeof(h->svc));
    h->n_crec_host = 0;
    h->slab.on = false; h->slab.rank = 0; h->slab.n_partners = 0;
    memset(&h->slab.dev, 0, sizeof(h->slab.dev));
    h->slab.d_send_idx = h->slab.d_recv_idx = nullptr; h->slab.d_send_voff = h->slab.d_recv_voff = nullptr;
    h->slab.d_owned = nullptr; h->slab.d_refx = h->slab.d_refy = nullptr; h->slab.arena = nullptr; h->slab.arena_bytes = 0;
    h->slab.epoch = 1; h->slab.pushed = 0; h->slab.max_send = h->slab.max_recv = 0; h->slab.send_bytes_total = 0;
    h->slab.arena_gen = 0;
    h->slab.cap_lists = h->slab.cap_owned = 0;
    memset(&h->cur, 0, sizeof(h->cur));
    memset(&h->hD, 0, sizeof(h->hD));
    memset(&h->P, 0, sizeof(h->P));
    memset(&h->last, 0, sizeof(h->last));
    memset(h->ms, 0, sizeof(h->ms));
    h->err[0] = 0;
    h->cfg = *cfg;
    h->P.cfg = *cfg;
    h->have_grid = h->have_fields = h->have_domain = h->have_floes = false;
    h->n_init = h->n_total = h->n_verts = h->n_verts_init = 0;
    h->n_lists = 0; h->d_hl_idx = nullptr; h->d_hl_voff = nullptr;
    h->n_mc = 0; h->n_topo = 0; h->topo_verts_n = 0; h->field_n = 0; h->n_rows_host = 0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    h->L.sms = prop.multiProcessorCount;
    h->L.maxv_large = 1024;
    h->L.maxx_large = 256;
    h->L.coupling_blocks_per_sm = 0;
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&h->L.stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&h->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    cudaEventCreate(&h->ev_c0);
    cudaEventCreate(&h->ev_c1);
    h->d_cf_dn = nullptr;
    h->gexec = nullptr;
    h->up_pending = 0;
    h->cpl_prelaunched = false;
    h->gen = 1;
    h->gkey_gen = 0;
    h->gkey_coupling = h->gkey_floes = h->gkey_pairs = -1;
    h->graph_launches = 0;
    h->graph_off = getenv("SZ_NO_GRAPH") != nullptr;
    h->L.capturing = false;
    h->L.no_phase_events = getenv("SZ_GRAPH_NO_EVENTS") != nullptr;
    h->L.chain_v2 = getenv("SZ_CHAIN_V1") == nullptr;
    h->L.pdl = 0;
    h->mc_spare = nullptr;
    h->mc_spare_cap = h->mc_off_cap = 0;
    h->rb_tx = h->rb_ty = nullptr; h->rb_extra = nullptr; h->rb_src = nullptr;
    h->rb_extra_cap = h->rb_src_cap = 0;
    h->next_partial = false;
    h->tables_gen = 0;
    h->tables_ptr[0] = h->tables_ptr[1] = h->tables_ptr[2] = nullptr;
    h->graph_max_floes = getenv("SZ_GRAPH_MAX_FLOES") ? atoi(getenv("SZ_GRAPH_MAX_FLOES")) : SZ_GRAPH_MAX_FLOES;
    h->cf_cap = 0;
    if (cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->stream_dn, cudaStreamNonBlocking) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    for (int k = 0; k < 4; ++k) cudaEventCreate(&h->ev_up[k]);
    for (int k = 0; k < 3; ++k) cudaEventCreate(&h->ev_dn[k]);
    cudaEventCreate(&h->ev_up_start);
    cudaEventCreate(&h->ev_dn_end);
    cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming);
    for (int k = 0; k < NEV; ++k) cudaEventCreate(&h->ev[k]);
    if (cudaMallocHost((void **)&h->h_cnt, sizeof(Counters)) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    memset(h->h_cnt, 0, sizeof(Counters));
    if (dalloc(&h->S.cnt, 1) != cudaSuccess || dalloc(&h->S.dom, 1) != cudaSuccess) { delete h; return SZ_ERR_CUDA; }
    cudaMemset(h->S.cnt, 0, sizeof(Counters));
    cudaMemset(h->S.dom, 0, sizeof(DomainDev));
    if (szk_configure(h->L) != 0 || szk_services_configure(h->L) != 0) { delete h; return SZ_ERR_CUDA; }
    register_arrays(h);
    *out = h;
    return SZ_OK;
}

extern "C" void sz_destroy(sz_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->L.stream);
    for (FloeArr &a : h->floe_arrays) if (*a.ptr) cudaFree(*a.ptr);
    for (void **p : h->floe_scratch) if (*p) cudaFree(*p);
    Store &S = h->S;
    StepBuf &B = h->B;
    dfree(S.verts); dfree(S.mc_off); dfree(S.mc); dfree(S.topo_vstart); dfree(S.topo_vcount); dfree(S.topo_verts);
    dfree(S.topo_cx); dfree(S.topo_cy); dfree(S.topo_rmax); dfree(S.ocn_u); dfree(S.ocn_v); dfree(S.ocn_hflx);
    dfree(S.atm_u); dfree(S.atm_v); dfree(S.fields8); dfree(S.cnt); dfree(S.dom);
    dfree(B.cell_count); dfree(B.cell_start); dfree(B.cell_fill); dfree(B.scan_block); dfree(B.cell_circ); dfree(B.pair_i); dfree(B.pair_j);
    dfree(B.lb_desc); dfree(B.lb_ticket); dfree(B.nb_scratch);
    dfree(B.low_pair); dfree(B.keep); dfree(B.dom_floe); dfree(B.dom_elem); dfree(B.item_nrows); dfree(B.item_row0);
    dfree(B.item_flags); dfree(B.large_items); dfree(B.mid_items); dfree(B.order); dfree(B.order_cls); dfree(B.force_items); dfree(B.force_meta); dfree(B.force_pts); dfree(B.class_count); dfree(B.pool); dfree(B.rows); dfree(B.fuse_pairs);
    dfree(h->d_hl_idx); dfree(h->d_hl_voff); dfree(h->mc_spare); dfree(h->rb_tx); dfree(h->rb_ty); dfree(h->rb_extra); dfree(h->rb_src);
    for (auto &m : h->slab.ipc_open) cudaIpcCloseMemHandle(m.ptr);
    h->slab.ipc_open.clear();
    for (unsigned char *p : h->slab.retired) cudaFree(p);
    h->slab.retired.clear();
    {
        SlabState &Bs = h->slab;
        dfree(Bs.d_send_idx); dfree(Bs.d_recv_idx); dfree(Bs.d_send_voff); dfree(Bs.d_recv_voff); dfree(Bs.d_owned);
        dfree(Bs.d_refx); dfree(Bs.d_refy); dfree(Bs.arena);
    }
    {
        SvcBuf &V = h->svc;
        dfree(V.pairs); dfree(V.area); dfree(V.inter); dfree(V.big); dfree(V.xg); dfree(V.yg); dfree(V.data); dfree(V.rec_area);
        dfree(V.rec_count); dfree(V.rec_off); dfree(V.cell_start); dfree(V.rec_floe); dfree(V.rec_cell); dfree(V.val_in);
        dfree(V.val_out); dfree(V.key_in); dfree(V.key_out); dfree(V.sort_tmp); dfree(V.cell_free); dfree(V.cell_topo);
    }
    {
        CouplingBuf &C = h->CB;
        dfree(C.rec_cell); dfree(C.rec_floe); dfree(C.rec_npts); dfree(C.rec_t); dfree(C.rec_d); dfree(C.rec_area);
        dfree(C.cell_count); dfree(C.cell_start); dfree(C.cell_fill); dfree(C.perm); dfree(C.big_recs); dfree(C.scan_block);
        dfree(C.sp_cell); dfree(C.sp_n); dfree(C.sp_sd); dfree(C.sp_t);
        dfree(S.ocn_temp); dfree(S.atm_temp); dfree(S.taux); dfree(S.tauy); dfree(S.sifrac);
    }
    if (h->h_cnt) cudaFreeHost(h->h_cnt);
    for (int k = 0; k < NEV; ++k) cudaEventDestroy(h->ev[k]);
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    cudaStreamDestroy(h->L.stream);
    cudaStreamDestroy(h->stream2);
    cudaStreamDestroy(h->stream_up);
    cudaStreamDestroy(h->stream_dn);
    for (int k = 0; k < 4; ++k) cudaEventDestroy(h->ev_up[k]);
    for (int k = 0; k < 3; ++k) cudaEventDestroy(h->ev_dn[k]);
    dfree(h->d_cf_dn);
    cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); cudaEventDestroy(h->ev_c0); cudaEventDestroy(h->ev_c1);
    delete h;
}

// ---- model description ---------------------------------------------------------------------------
extern "C" int32_t sz_set_grid(sz_handle *h, int32_t Nx, int32_t Ny, double x0, double xf, double y0, double yf) {
    if (!h || Nx < 1 || Ny < 1 || !(xf > x0) || !(yf > y0)) return fail(h, SZ_ERR_INVALID, "set_grid: bad extent");
    h->P.Nx = Nx; h->P.Ny = Ny; h->P.x0 = x0; h->P.xf = xf; h->P.y0 = y0; h->P.yf = yf;
    h->P.dx = (xf - x0) / Nx;  // grids.jl:180-211
    h->P.dy = (yf - y0) / Ny;
    h->have_grid = true;
    h->have_fields = false;
    h->gen++;
    return SZ_OK;
}

extern "C" int32_t sz_set_fields(sz_handle *h, const double *ou, const double *ov, const double *oh, const double *au,
                                 const double *av) {
    if (!h || !h->have_grid) return fail(h, SZ_ERR_INVALID, "set_fields before set_grid");
    cudaSetDevice(h->cfg.device);
    size_t n = (size_t)(h->P.Nx + 1) * (size_t)(h->P.Ny + 1);
    Store &S = h->S;
    if (n != h->field_n) {
        dfree(S.ocn_u); dfree(S.ocn_v); dfree(S.ocn_hflx); dfree(S.atm_u); dfree(S.atm_v);
        CK(dalloc(&S.ocn_u, n)); CK(dalloc(&S.ocn_v, n)); CK(dalloc(&S.ocn_hflx, n));
        CK(dalloc(&S.atm_u, n)); CK(dalloc(&S.atm_v, n));
        dfree(S.fields8);
        CK(dalloc(&S.fields8, 8 * n));
        dfree(S.ocn_temp); dfree(S.atm_temp); dfree(S.taux); dfree(S.tauy); dfree(S.sifrac);
        CK(dalloc(&S.ocn_temp, n)); CK(dalloc(&S.atm_temp, n)); CK(dalloc(&S.taux, n)); CK(dalloc(&S.tauy, n)); CK(dalloc(&S.sifrac, n));
        double *z[5] = {S.ocn_temp, S.atm_temp, S.taux, S.tauy, S.sifrac};
        for (int k = 0; k < 5; ++k) CK(cudaMemsetAsync(z[k], 0, sizeof(double) * n, h->L.stream));
        CouplingBuf &C = h->CB;
        dfree(C.cell_count); dfree(C.cell_start); dfree(C.cell_fill); dfree(C.scan_block);
        CK(dalloc(&C.cell_count, n + 2)); CK(dalloc(&C.cell_start, n + 2)); CK(dalloc(&C.cell_fill, n + 2));
        CK(dalloc(&C.scan_block, n / 4096 + 8));
        C.cap_cells = (int)n;
        h->field_n = n;
    }
    const double *src[5] = {ou, ov, oh, au, av};
    double *dst[5] = {S.ocn_u, S.ocn_v, S.ocn_hflx, S.atm_u, S.atm_v};
    for (int k = 0; k < 5; ++k) {
        if (src[k]) CK(cudaMemcpyAsync(dst[k], src[k], sizeof(double) * n, cudaMemcpyHostToDevice, h->L.stream));
        else CK(cudaMemsetAsync(dst[k], 0, sizeof(double) * n, h->L.stream));
    }
    {
        auto nonzero = [&](const double *p) { if (!p) return false; for (size_t k = 0; k < n; ++k) if (p[k] != 0.0) return true; return false; };
        h->P.atm_nonzero = nonzero(au) || nonzero(av);
        h->P.hflx_nonzero = nonzero(oh) || h->cfg.two_way_coupling_on;
    }
    h->gen++;
    szk_pack_fields(h->L, S, (int)n);
    CK(cudaStreamSynchronize(h->L.stream));
    CK(cudaGetLastError());
    h->have_fields = true;
    return SZ_OK;
}

extern "C" int32_t sz_set_temperatures(sz_handle *h, const double *ot, const double *at) {
    if (!h || !h->have_fields) return fail(h, SZ_ERR_INVALID, "set_temperatures before set_fields");
    cudaSetDevice(h->cfg.device);
    size_t n = h->field_n;
    const double *src[2] = {ot, at};
    double *dst[2] = {h->S.ocn_temp, h->S.atm_temp};
    for (int k = 0; k < 2; ++k) {
        if (src[k]) CK(cudaMemcpyAsync(dst[k], src[k], sizeof(double) * n, cudaMemcpyHostToDevice, h->L.stream));
        else CK(cudaMemsetAsync(dst[k], 0, sizeof(double) * n, h->L.stream));
    }
    CK(cudaStreamSynchronize(h->L.stream));
    return SZ_OK;
}

extern "C" int32_t sz_get_ocean_fields(sz_handle *h, double *tx, double *ty, double *si, double *hf) {
    if (!h || !h->have_fields) return fail(h, SZ_ERR_INVALID, "get_ocean_fields before set_fields");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    size_t n = h->field_n;
    double *dst[4] = {tx, ty, si, hf};
    const double *src[4] = {h->S.taux, h->S.tauy, h->S.sifrac, h->S.ocn_hflx};
    for (int k = 0; k < 4; ++k)
        if (dst[k]) CK(cudaMemcpy(dst[k], src[k], sizeof(double) * n, cudaMemcpyDeviceToHost));
    return SZ_OK;
}

extern "C" int32_t sz_get_cell_floes(sz_handle *h, int64_t *n, int64_t *cell_xy, int64_t *floe, double *vals) {
    if (!h || !n) return SZ_ERR_INVALID;
    *n = h->n_crec_host;
    if (!cell_xy || !floe || !vals || h->n_crec_host == 0) return SZ_OK;
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    const int m = h->n_crec_host, nx1 = h->P.Nx + 1;
    std::vector<int> perm(m), rc(m), rf(m), rn(m);
    std::vector<double2> rt(m), rd(m);
    CouplingBuf &C = h->CB;
    CK(cudaMemcpy(perm.data(), C.perm, sizeof(int) * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rc.data(), C.rec_cell, sizeof(int) * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rf.data(), C.rec_floe, sizeof(int) * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rn.data(), C.rec_npts, sizeof(int) * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rt.data(), C.rec_t, sizeof(double2) * m, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(rd.data(), C.rec_d, sizeof(double2) * m, cudaMemcpyDeviceToHost));
    for (int k = 0; k < m; ++k) {
        int r = perm[k];
        cell_xy[2 * k] = rc[r] % nx1 + 1;
        cell_xy[2 * k + 1] = rc[r] / nx1 + 1;
        floe[k] = rf[r] + 1;
        vals[5 * k] = rt[r].x; vals[5 * k + 1] = rt[r].y; vals[5 * k + 2] = (double)rn[r];
        vals[5 * k + 3] = rd[r].x; vals[5 * k + 4] = rd[r].y;
    }
    return SZ_OK;
}

extern "C" int32_t sz_set_domain(sz_handle *h, const int32_t kinds[4], const double vals[4], const double uv[8],
                                 const double rect[16], int32_t n_topo, const int64_t *toff, const double *txy,
                                 const double *tcent, const double *trmax) {
    if (!h || !kinds || !vals || !rect) return fail(h, SZ_ERR_INVALID, "set_domain: null argument");
    // domains.jl:11-33
    if ((kinds[0] == SZ_BOUNDARY_PERIODIC) != (kinds[1] == SZ_BOUNDARY_PERIODIC) ||
        (kinds[2] == SZ_BOUNDARY_PERIODIC) != (kinds[3] == SZ_BOUNDARY_PERIODIC))
        return fail(h, SZ_ERR_INVALID, "set_domain: periodic boundaries must be paired");
    if (!(vals[0] > vals[1]) || !(vals[2] > vals[3])) return fail(h, SZ_ERR_INVALID, "set_domain: north <= south or east <= west");
    if (n_topo < 0 || (n_topo > 0 && (!toff || !txy || !tcent || !trmax))) return fail(h, SZ_ERR_INVALID, "set_domain: topography arrays missing");
    cudaSetDevice(h->cfg.device);
    DomainDev &D = h->hD;
    for (int w = 0; w < 4; ++w) {
        D.kind[w] = kinds[w];
        D.val[w] = vals[w];
        D.wu[w] = uv ? uv[2 * w] : 0.0;
        D.wv[w] = uv ? uv[2 * w + 1] : 0.0;
        for (int k = 0; k < 4; ++k) D.rect[w][k] = rect[4 * w + k];
    }
    D.n_topo = n_topo;
    h->P.per_x = kinds[2] == SZ_BOUNDARY_PERIODIC;
    h->P.per_y = kinds[0] == SZ_BOUNDARY_PERIODIC;
    Store &S = h->S;
    dfree(S.topo_vstart); dfree(S.topo_vcount); dfree(S.topo_verts); dfree(S.topo_cx); dfree(S.topo_cy); dfree(S.topo_rmax);
    if (n_topo > 0) {
        std::vector<int> vs(n_topo), vc(n_topo);
        std::vector<double> cx(n_topo), cy(n_topo);
        for (int k = 0; k < n_topo; ++k) {
            vs[k] = (int)toff[k];
            vc[k] = (int)(toff[k + 1] - toff[k]);
            if (vc[k] < 4) return fail(h, SZ_ERR_INVALID, "set_domain: a topography ring needs >= 4 points (closed)");
            if (vc[k] > h->L.maxv_large) return fail(h, SZ_ERR_UNSUPPORTED, "set_domain: topography ring exceeds 1024 points");
            cx[k] = tcent[2 * k];
            cy[k] = tcent[2 * k + 1];
        }
        size_t nv = (size_t)toff[n_topo];
        CK(dalloc(&S.topo_vstart, n_topo)); CK(dalloc(&S.topo_vcount, n_topo)); CK(dalloc(&S.topo_verts, nv));
        CK(dalloc(&S.topo_cx, n_topo)); CK(dalloc(&S.topo_cy, n_topo)); CK(dalloc(&S.topo_rmax, n_topo));
        CK(cudaMemcpy(S.topo_vstart, vs.data(), sizeof(int) * n_topo, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.topo_vcount, vc.data(), sizeof(int) * n_topo, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.topo_verts, txy, sizeof(double2) * nv, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.topo_cx, cx.data(), sizeof(double) * n_topo, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.topo_cy, cy.data(), sizeof(double) * n_topo, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(S.topo_rmax, trmax, sizeof(double) * n_topo, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(S.dom, &D, sizeof(DomainDev), cudaMemcpyHostToDevice));
    h->n_topo = n_topo;
    h->have_domain = true;
    h->gen++;
    return SZ_OK;
}

extern "C" int32_t sz_get_domain(sz_handle *h, double vals[4], double rect[16]) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "get_domain before set_domain");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    CK(cudaMemcpy(&h->hD, h->S.dom, sizeof(DomainDev), cudaMemcpyDeviceToHost));
    for (int w = 0; w < 4; ++w) {
        vals[w] = h->hD.val[w];
        for (int k = 0; k < 4; ++k) rect[4 * w + k] = h->hD.rect[w][k];
    }
    return SZ_OK;
}

// ---- floe state -----------------------------------------------------------------------------------
static int32_t sync_counters(sz_handle *h) {
    CK(cudaMemcpyAsync(h->h_cnt, h->S.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, h->L.stream));
    CK(cudaStreamSynchronize(h->L.stream));
    h->n_total = h->h_cnt->n_total;
    h->n_verts = h->h_cnt->n_verts;
    return SZ_OK;
}

// every per-floe double array of the SoA (a NULL source zero-fills)
static int32_t upload_scalars(sz_handle *h, const sz_floe_soa *s, int n) {
    Store &S = h->S;
    cudaStream_t st = h->L.stream;
    if (n == 0) return SZ_OK;
    auto up = [&](double *dst, const double *src, size_t w) -> cudaError_t {
        if (src) return cudaMemcpyAsync(dst, src, sizeof(double) * w * n, cudaMemcpyHostToDevice, st);
        return cudaMemsetAsync(dst, 0, sizeof(double) * w * n, st);
    };
    CK(up(S.cx, s->centroid_x, 1)); CK(up(S.cy, s->centroid_y, 1)); CK(up(S.height, s->height, 1));
    CK(up(S.area, s->area, 1)); CK(up(S.mass, s->mass, 1)); CK(up(S.rmax, s->rmax, 1)); CK(up(S.moment, s->moment, 1));
    CK(up(S.alpha, s->alpha, 1)); CK(up(S.u, s->u, 1)); CK(up(S.v, s->v, 1)); CK(up(S.xi, s->xi, 1));
    CK(up(S.fxOA, s->fxOA, 1)); CK(up(S.fyOA, s->fyOA, 1)); CK(up(S.trqOA, s->trqOA, 1));
    CK(up(S.hflx, s->hflx_factor, 1)); CK(up(S.overarea, s->overarea, 1)); CK(up(S.ctrq, s->collision_trq, 1));
    CK(up(S.p_dxdt, s->p_dxdt, 1)); CK(up(S.p_dydt, s->p_dydt, 1)); CK(up(S.p_dudt, s->p_dudt, 1));
    CK(up(S.p_dvdt, s->p_dvdt, 1)); CK(up(S.p_dxidt, s->p_dxidt, 1)); CK(up(S.p_dalphadt, s->p_dalphadt, 1));
    CK(up(S.stress_accum, s->stress_accum, 4)); CK(up(S.stress_instant, s->stress_instant, 4)); CK(up(S.strain, s->strain, 4));
    // collision_force is [n][2] on the host, two columns on the device
    if (s->collision_force) {  // [n][2] on the host -> two device columns
        CK(cudaMemcpyAsync(h->B.cell_circ, s->collision_force, sizeof(double2) * n, cudaMemcpyHostToDevice, st));
        szk_deinterleave(h->L, h->B.cell_circ, S.cfx, S.cfy, n);
    } else {
        CK(cudaMemsetAsync(S.cfx, 0, sizeof(double) * n, st));
        CK(cudaMemsetAsync(S.cfy, 0, sizeof(double) * n, st));
    }
    if (s->status_tag) CK(cudaMemcpyAsync(S.status, s->status_tag, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    return SZ_OK;
}

// mc_src == NULL: Monte-Carlo points come from s->mc_x / s->mc_y (layout s->mc_offsets).
// mc_src != NULL (sz_slab_rebuild): the points of floe i are already on the device — at offset mc_src[i] of the resident
// array when mc_src[i] >= 0 — or arrive in the COMPACT host arrays s->mc_x / s->mc_y at offset -1 - mc_src[i]
// (n_extra points in total: migrants); the new array is gathered on the device.
struct DbgTimer {  // SZ_SLAB_DEBUG: wall clock of the stages of an upload on stderr
    bool on;
    std::chrono::steady_clock::time_point t0;
    int dev;
    explicit DbgTimer(int d) : on(getenv("SZ_SLAB_DEBUG") != nullptr), t0(std::chrono::steady_clock::now()), dev(d) {}
    void lap(const char *what) {
        if (!on) return;
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[upload dev %d]   %-26s %8.2f ms\n", dev, what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static int32_t upload_floes_impl(sz_handle *h, const sz_floe_soa *s, const int64_t *mc_src, int64_t n_extra) {
    if (!h || !s || s->n < 0 || s->n_init < 0 || s->n_init > s->n) return fail(h, SZ_ERR_INVALID, "upload_floes: bad sizes");
    DbgTimer dbg(h->cfg.device);
    if (s->n > 0 && (!s->centroid_x || !s->centroid_y || !s->area || !s->rmax || !s->vert_offsets || !s->vert_xy))
        return fail(h, SZ_ERR_INVALID, "upload_floes: geometry arrays are required");
    if (s->n > (1ll << 28)) return fail(h, SZ_ERR_UNSUPPORTED, "upload_floes: too many floes");
    cudaSetDevice(h->cfg.device);
    const int n = (int)s->n, n_init = (int)s->n_init;
    const long long V = n > 0 ? s->vert_offsets[n] : 0;
    const long long M = (s->mc_offsets && n > 0) ? s->mc_offsets[n] : 0;
    if (V > (1ll << 30)) return fail(h, SZ_ERR_UNSUPPORTED, "upload_floes: too many vertices");
    std::vector<int> vstart(n), vcount(n), parent(n, -1), nghost(n, 0), gslot((size_t)n * SZ_MAX_GHOSTS, 0);
    for (int i = 0; i < n; ++i) {
        long long a = s->vert_offsets[i], b = s->vert_offsets[i + 1];
        if (b - a < 4) return fail(h, SZ_ERR_INVALID, "upload_floes: a ring needs >= 4 points (closed)");
        if (b - a > h->L.maxv_large) return fail(h, SZ_ERR_UNSUPPORTED, "upload_floes: a ring exceeds 1024 points");
        const double *r = s->vert_xy + 2 * a;
        if (r[0] != r[2 * (b - a - 1)] || r[1] != r[2 * (b - a - 1) + 1]) return fail(h, SZ_ERR_INVALID, "upload_floes: rings must be closed");
        vstart[i] = (int)a;
        vcount[i] = (int)(b - a);
    }
    if (s->ghost_offsets) {
        for (int i = 0; i < n; ++i) {
            long long a = s->ghost_offsets[i], b = s->ghost_offsets[i + 1];
            if (b - a > SZ_MAX_GHOSTS) return fail(h, SZ_ERR_INVALID, "upload_floes: a floe has more than 3 ghosts");
            nghost[i] = (int)(b - a);
            for (long long g = a; g < b; ++g) {
                long long gi = s->ghost_index[g] - 1;
                if (gi < 0 || gi >= n) return fail(h, SZ_ERR_INVALID, "upload_floes: ghost index out of range");
                gslot[(size_t)i * SZ_MAX_GHOSTS + (g - a)] = (int)gi;
                parent[gi] = i;
            }
        }
    }
    // capacities
    int want_cap = h->cfg.floe_capacity > 0 ? (int)h->cfg.floe_capacity
                                            : n + std::max(64, std::min(3 * n_init, n_init / 2 + 4096));
    if (want_cap < n) want_cap = n;
    // Hysteresis: a list that grew by a few floes (every slab rebuild changes the halo a little) keeps its buffers as long
    // as half of the ghost headroom is left — re-allocating ~60 arrays costs tens of milliseconds (and far more while
    // peers have this device's memory mapped); an overflowing ghost pass still grows them on demand.
    const int min_cap = h->cfg.floe_capacity > 0 ? want_cap : n + (want_cap - n) / 2;
    if (min_cap > h->S.cap_floes || !h->S.cx) {
        int32_t rc = grow_floes(h, want_cap, 0);
        if (rc) return rc;
    }
    const long long per_floe = n_init > 0 ? (V / std::max(n, 1) + 1) : 0;
    long long want_v = V + per_floe * (h->S.cap_floes - n) + 64, min_v = V + per_floe * ((want_cap - n) / 2) + 64;
    if (min_v > h->S.cap_verts || !h->S.verts) {
        int32_t rc = grow_verts(h, (int)std::min<long long>(std::max(want_v, min_v), 1ll << 30), 0);
        if (rc) return rc;
    }
    Store &S = h->S;
    double2 *old_mc = nullptr;
    long long old_cap = 0;
    if (mc_src) {  // gather into the spare array, then swap: no allocation per rebuild once both have their headroom
        old_mc = S.mc;
        old_cap = S.cap_mc;
        if (h->mc_spare_cap < M || !h->mc_spare) {
            dfree(h->mc_spare);
            const long long cap = M + M / 16 + 4096;
            CK(dalloc(&h->mc_spare, (size_t)cap));
            h->mc_spare_cap = cap;
        }
        S.mc = h->mc_spare;
        S.cap_mc = h->mc_spare_cap;
        h->mc_spare = nullptr;
        h->mc_spare_cap = 0;
        h->gen++;
        dbg.lap("Monte-Carlo spare array");
    } else if (M > S.cap_mc || !S.mc) {
        dfree(S.mc);
        CK(dalloc(&S.mc, (size_t)M));
        S.cap_mc = M;
    }
    dbg.lap("host tables + capacities");
    if (h->mc_off_cap < (long long)n_init + 2 || !S.mc_off) {
        dfree(S.mc_off);
        h->mc_off_cap = (long long)n_init + n_init / 8 + 1024;
        CK(dalloc(&S.mc_off, (size_t)h->mc_off_cap));
    }
    StepBuf &B = h->B;
    int ppf = h->cfg.max_pairs_per_floe > 0 ? h->cfg.max_pairs_per_floe : 24;
    // sized from the floe CAPACITY, not from n: a list that grows by a few floes (slab rebuilds) must not re-allocate them
    long long want_pairs = (long long)ppf * S.cap_floes / 2 + 1024, want_domc = (long long)S.cap_floes / 2 + 4096 + 4ll * S.cap_floes;
    if (h->n_topo > 0) want_domc += (long long)S.cap_floes;
    want_pairs = std::min<long long>(want_pairs, 1ll << 30);
    want_domc = std::min<long long>(want_domc, 1ll << 30);
    if (want_pairs > B.cap_pairs || want_domc > B.cap_dom || !B.pair_i) {
        int32_t rc = set_pair_cap(h, (int)std::max<long long>(want_pairs, B.cap_pairs), (int)std::max<long long>(want_domc, B.cap_dom));
        if (rc) return rc;
    }
    long long want_pool = want_pairs / 2 + want_domc + 1024;
    if (want_pool > B.cap_pool || !B.pool) { int32_t rc = set_pool_cap(h, (int)want_pool); if (rc) return rc; }
    long long want_rows = 2 * want_pool;
    if (want_rows > B.cap_rows || !B.rows) { int32_t rc = set_row_cap(h, (int)std::min<long long>(want_rows, 1ll << 28)); if (rc) return rc; }
    if (!B.fuse_pairs) { int32_t rc = set_fuse_cap(h, std::max(1024, S.cap_floes)); if (rc) return rc; }
    if (h->cfg.two_way_coupling_on && h->CB.cap_crec < 6 * n_init + 4096) { int32_t rc = set_crec_cap(h, 6 * n_init + 4096); if (rc) return rc; }
    // copies
    cudaStream_t st = h->L.stream;
    { int32_t rc = upload_scalars(h, s, n); if (rc) return rc; }
    std::vector<int> status(n, SZ_STATUS_ACTIVE);
    std::vector<long long> id(n), gid(n, 0);
    for (int i = 0; i < n; ++i) {
        id[i] = s->id ? s->id[i] : i + 1;
        if (s->ghost_id) gid[i] = s->ghost_id[i];
    }
    if (n > 0) {
        if (!s->status_tag) CK(cudaMemcpyAsync(S.status, status.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.id, id.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.ghost_id, gid.data(), sizeof(long long) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.parent, parent.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.nghost, nghost.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.ghost_slot, gslot.data(), sizeof(int) * SZ_MAX_GHOSTS * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.vstart, vstart.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(S.vcount, vcount.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(S.warn, 0, sizeof(uint32_t) * n, st));
        CK(cudaMemcpyAsync(S.verts, s->vert_xy, sizeof(double2) * (size_t)V, cudaMemcpyHostToDevice, st));
    }
    std::vector<long long> mo((size_t)n_init + 1, 0);
    if (s->mc_offsets) for (int i = 0; i <= n_init; ++i) mo[i] = s->mc_offsets[i];
    CK(cudaMemcpyAsync(S.mc_off, mo.data(), sizeof(long long) * ((size_t)n_init + 1), cudaMemcpyHostToDevice, st));
    long long Mi = mo[n_init];
    dbg.lap("per-floe arrays + rings");
    if (mc_src) {
        // grow-only scratch: no allocation in a rebuild once the sizes have settled
        if (h->rb_extra_cap < n_extra) {
            dfree(h->rb_tx); dfree(h->rb_ty); dfree(h->rb_extra);
            h->rb_extra_cap = n_extra + n_extra / 2 + 65536;
            CK(dalloc(&h->rb_tx, (size_t)h->rb_extra_cap)); CK(dalloc(&h->rb_ty, (size_t)h->rb_extra_cap));
            CK(dalloc(&h->rb_extra, (size_t)h->rb_extra_cap));
        }
        if (h->rb_src_cap < (long long)n_init + 1) {
            dfree(h->rb_src);
            h->rb_src_cap = (long long)n_init + n_init / 8 + 1024;
            CK(dalloc(&h->rb_src, (size_t)h->rb_src_cap));
        }
        if (n_extra > 0) {
            CK(cudaMemcpyAsync(h->rb_tx, s->mc_x, sizeof(double) * n_extra, cudaMemcpyHostToDevice, st));
            CK(cudaMemcpyAsync(h->rb_ty, s->mc_y, sizeof(double) * n_extra, cudaMemcpyHostToDevice, st));
            szk_interleave(h->L, h->rb_tx, h->rb_ty, h->rb_extra, n_extra);
        }
        std::vector<long long> hs((size_t)n_init + 1, 0);
        for (int i = 0; i < n_init; ++i) hs[i] = mc_src[i];
        CK(cudaMemcpyAsync(h->rb_src, hs.data(), sizeof(long long) * ((size_t)n_init + 1), cudaMemcpyHostToDevice, st));
        szk_mc_regather(h->L, S.mc, S.mc_off, old_mc, h->rb_extra, h->rb_src, n_init);
        CK(cudaStreamSynchronize(st));
        CK(cudaGetLastError());
        h->mc_spare = old_mc;  // the next rebuild gathers into it
        h->mc_spare_cap = old_cap;
        dbg.lap("Monte-Carlo regather");
    } else if (Mi > 0) {
        double *tx = nullptr, *ty = nullptr;
        CK(dalloc(&tx, (size_t)Mi)); CK(dalloc(&ty, (size_t)Mi));
        CK(cudaMemcpyAsync(tx, s->mc_x, sizeof(double) * Mi, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ty, s->mc_y, sizeof(double) * Mi, cudaMemcpyHostToDevice, st));
        szk_interleave(h->L, tx, ty, S.mc, Mi);
        CK(cudaStreamSynchronize(st));
        cudaFree(tx); cudaFree(ty);
    }
    CK(cudaMemsetAsync(B.row_off, 0, sizeof(int) * ((size_t)S.cap_floes + 2), st));
    S.n_init = n_init;
    szk_mc_radius(h->L, S);
    szk_set_counts(h->L, S, n, (int)V);
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    h->n_init = n_init; h->n_total = n; h->n_verts = (int)V; h->n_mc = Mi;
    h->n_verts_init = n_init > 0 ? (int)s->vert_offsets[n_init] : 0;
    h->n_rows_host = 0;
    memset(&h->last, 0, sizeof(h->last));
    h->h_vstart = vstart;
    h->h_vcount = vcount;
    h->h_mc_off = mo;
    dbg.lap("rest");
    h->have_floes = true;
    h->slab.on = false;  // a new floe list: the halo lists must be configured again
    h->gen++;
    return SZ_OK;
}

extern "C" int32_t sz_upload_floes(sz_handle *h, const sz_floe_soa *s) { return upload_floes_impl(h, s, nullptr, 0); }

// Refresh the dynamic state of the resident floes (same floe list and ring sizes as the last
// sz_upload_floes); Monte-Carlo points, ids and ghost links stay resident.
extern "C" int32_t sz_upload_state(sz_handle *h, const sz_floe_soa *s) {
    if (!h || !s) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "upload_state before upload_floes");
    if (s->n != h->n_total || s->n_init != h->n_init) return fail(h, SZ_ERR_INVALID, "upload_state: floe count differs from the resident store");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "upload_state with ghosts present");
    if (!s->centroid_x || !s->centroid_y || !s->area || !s->rmax || !s->vert_xy) return fail(h, SZ_ERR_INVALID, "upload_state: geometry arrays are required");
    cudaSetDevice(h->cfg.device);
    int32_t rc = upload_scalars(h, s, h->n_total);
    if (rc) return rc;
    if (s->vert_offsets && s->vert_offsets[h->n_total] != h->n_verts) return fail(h, SZ_ERR_INVALID, "upload_state: vertex count differs from the resident store");
    if (h->n_verts > 0) CK(cudaMemcpyAsync(h->S.verts, s->vert_xy, sizeof(double2) * (size_t)h->n_verts, cudaMemcpyHostToDevice, h->L.stream));
    CK(cudaStreamSynchronize(h->L.stream));
    return SZ_OK;
}

extern "C" int32_t sz_get_counts(sz_handle *h, sz_counts *c) {
    if (!h || !c) return SZ_ERR_INVALID;
    memset(c, 0, sizeof(*c));
    c->n_init = h->n_init;
    c->n_total = h->n_total;
    c->n_vertices = h->n_verts;
    c->n_mc = h->n_mc;
    if (h->have_floes && h->n_total > 0) {
        cudaSetDevice(h->cfg.device);
        std::vector<int> ng(h->n_total);
        CK(cudaMemcpy(ng.data(), h->S.nghost, sizeof(int) * h->n_total, cudaMemcpyDeviceToHost));
        for (int v : ng) c->n_ghost_links += v;
        int nr = 0;
        CK(cudaMemcpy(&nr, h->B.row_off + h->n_total, sizeof(int), cudaMemcpyDeviceToHost));
        c->n_rows = nr;
    }
    c->n_candidates = h->last.n_cand;
    c->n_pairs = h->last.n_kept;
    c->n_overlap = h->last.n_overlap;
    c->n_fuse = h->last.n_fuse;
    c->n_domain_pairs = h->last.n_domchecks;
    c->n_clip_fail = h->last.n_clipfail;
    return SZ_OK;
}

extern "C" int32_t sz_download_floes(sz_handle *h, sz_floe_soa *s) {
    if (!h || !s) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "download_floes before upload_floes");
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->L.stream;
    CK(cudaStreamSynchronize(st));
    const int n = h->n_total, n_init = h->n_init;
    Store &S = h->S;
    s->n = n;
    s->n_init = n_init;
    if (n == 0) return SZ_OK;
    auto dn = [&](double *dst, const double *src, size_t w) -> cudaError_t {
        if (!dst) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, sizeof(double) * w * n, cudaMemcpyDeviceToHost, st);
    };
    CK(dn(s->centroid_x, S.cx, 1)); CK(dn(s->centroid_y, S.cy, 1)); CK(dn(s->height, S.height, 1));
    CK(dn(s->area, S.area, 1)); CK(dn(s->mass, S.mass, 1)); CK(dn(s->rmax, S.rmax, 1)); CK(dn(s->moment, S.moment, 1));
    CK(dn(s->alpha, S.alpha, 1)); CK(dn(s->u, S.u, 1)); CK(dn(s->v, S.v, 1)); CK(dn(s->xi, S.xi, 1));
    CK(dn(s->fxOA, S.fxOA, 1)); CK(dn(s->fyOA, S.fyOA, 1)); CK(dn(s->trqOA, S.trqOA, 1));
    CK(dn(s->hflx_factor, S.hflx, 1)); CK(dn(s->overarea, S.overarea, 1)); CK(dn(s->collision_trq, S.ctrq, 1));
    CK(dn(s->p_dxdt, S.p_dxdt, 1)); CK(dn(s->p_dydt, S.p_dydt, 1)); CK(dn(s->p_dudt, S.p_dudt, 1));
    CK(dn(s->p_dvdt, S.p_dvdt, 1)); CK(dn(s->p_dxidt, S.p_dxidt, 1)); CK(dn(s->p_dalphadt, S.p_dalphadt, 1));
    CK(dn(s->stress_accum, S.stress_accum, 4)); CK(dn(s->stress_instant, S.stress_instant, 4)); CK(dn(s->strain, S.strain, 4));
    if (s->collision_force) {  // two device columns -> [n][2] on the host (cell_circ is scratch between steps)
        szk_interleave(h->L, S.cfx, S.cfy, h->B.cell_circ, n);
        CK(cudaMemcpyAsync(s->collision_force, h->B.cell_circ, sizeof(double2) * n, cudaMemcpyDeviceToHost, st));
    }
    if (s->status_tag) CK(cudaMemcpyAsync(s->status_tag, S.status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    if (s->id) CK(cudaMemcpyAsync(s->id, S.id, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    if (s->ghost_id) CK(cudaMemcpyAsync(s->ghost_id, S.ghost_id, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    const bool no_ghosts = n == n_init && (int)h->h_vcount.size() == n;
    std::vector<int> vstart_d, vcount_d, nghost, gslot;
    const int *vstart = h->h_vstart.data(), *vcount = h->h_vcount.data();
    if (!no_ghosts) {  // ghosts resident: ring table and ghost links live on the device
        vstart_d.resize(n); vcount_d.resize(n); nghost.resize(n); gslot.resize((size_t)n * SZ_MAX_GHOSTS);
        CK(cudaMemcpyAsync(vstart_d.data(), S.vstart, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(vcount_d.data(), S.vcount, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(nghost.data(), S.nghost, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(gslot.data(), S.ghost_slot, sizeof(int) * SZ_MAX_GHOSTS * n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        vstart = vstart_d.data();
        vcount = vcount_d.data();
    }
    // rings: the device keeps them in floe order (ghost rings appended), so one contiguous copy
    long long vo = 0;
    bool contiguous = true;
    if (no_ghosts) {
        vo = h->n_verts_init;
        if (s->vert_offsets) {
            long long o = 0;
            for (int i = 0; i < n; ++i) { s->vert_offsets[i] = o; o += vcount[i]; }
            s->vert_offsets[n] = o;
        }
    } else {
        for (int i = 0; i < n; ++i) {
            if (vstart[i] != vo) contiguous = false;
            if (s->vert_offsets) s->vert_offsets[i] = vo;
            vo += vcount[i];
        }
        if (s->vert_offsets) s->vert_offsets[n] = vo;
    }
    if (s->vert_xy) {
        if (contiguous) {
            CK(cudaMemcpyAsync(s->vert_xy, S.verts, sizeof(double2) * (size_t)vo, cudaMemcpyDeviceToHost, st));
        } else {
            long long o = 0;
            for (int i = 0; i < n; ++i) {
                CK(cudaMemcpyAsync(s->vert_xy + 2 * o, S.verts + vstart[i], sizeof(double2) * vcount[i], cudaMemcpyDeviceToHost, st));
                o += vcount[i];
            }
        }
    }
    if (s->mc_offsets)
        for (int i = 0; i <= n; ++i) s->mc_offsets[i] = h->h_mc_off[std::min(i, n_init)];
    CK(cudaStreamSynchronize(st));
    if ((s->mc_x || s->mc_y) && h->n_mc > 0) {
        double *tx = nullptr, *ty = nullptr;
        CK(dalloc(&tx, (size_t)h->n_mc)); CK(dalloc(&ty, (size_t)h->n_mc));
        szk_deinterleave(h->L, S.mc, tx, ty, h->n_mc);
        if (s->mc_x) CK(cudaMemcpyAsync(s->mc_x, tx, sizeof(double) * h->n_mc, cudaMemcpyDeviceToHost, st));
        if (s->mc_y) CK(cudaMemcpyAsync(s->mc_y, ty, sizeof(double) * h->n_mc, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(tx); cudaFree(ty);
    }
    if (no_ghosts) {
        if (s->ghost_offsets) memset(s->ghost_offsets, 0, sizeof(int64_t) * ((size_t)n + 1));
        return SZ_OK;
    }
    long long go = 0;
    for (int i = 0; i < n; ++i) {
        if (s->ghost_offsets) s->ghost_offsets[i] = go;
        for (int g = 0; g < nghost[i]; ++g) {
            if (s->ghost_index) s->ghost_index[go] = gslot[(size_t)i * SZ_MAX_GHOSTS + g] + 1;
            go++;
        }
    }
    if (s->ghost_offsets) s->ghost_offsets[n] = go;
    return SZ_OK;
}

// ---- the hot path ---------------------------------------------------------------------------------------
static const char *errbits(uint32_t e, char *buf, size_t len) {
    snprintf(buf, len, "device buffer overflow:%s%s%s%s%s%s%s%s%s", (e & ERR_PAIR_CAP) ? " pairs" : "",
             (e & ERR_POOL_CAP) ? " contact-pool" : "", (e & ERR_ROW_CAP) ? " rows" : "", (e & ERR_GHOST_CAP) ? " ghost-floes" : "",
             (e & ERR_VERT_CAP) ? " ghost-vertices" : "", (e & ERR_FUSE_CAP) ? " fuse-pairs" : "",
             (e & ERR_POLY_TOO_LARGE) ? " polygon-workspace(>1024 points or >256 crossings)" : "",
             (e & ERR_DOM_CAP) ? " domain-items" : "", (e & ERR_GHOST_SLOTS) ? " ghost-slots(>3 images)" : "");
    return buf;
}

// grow whatever overflowed; returns SZ_OK if a retry makes sense
static int32_t handle_overflow(sz_handle *h, const Counters &c) {
    uint32_t e = c.error;
    char buf[256];
    if (e & ERR_SLAB_TIMEOUT) {
        szk_clear_error(h->L, h->S);
        return fail(h, SZ_ERR_INVALID, "slab: a neighbour rank did not publish (or acknowledge) its boundary floes in time — did every rank make the same sz_slab_* call?");
    }
    if (e & (ERR_POLY_TOO_LARGE | ERR_GHOST_SLOTS)) return fail(h, (e & ERR_POLY_TOO_LARGE) ? SZ_ERR_UNSUPPORTED : SZ_ERR_CAPACITY, errbits(e, buf, sizeof(buf)));
    int32_t rc = SZ_OK;
    if (e & (ERR_PAIR_CAP | ERR_DOM_CAP)) {
        long long wp = (e & ERR_PAIR_CAP) ? (long long)c.want_pairs * 5 / 4 + 1024 : h->B.cap_pairs;
        long long wd = (e & ERR_DOM_CAP) ? (long long)c.want_dom * 5 / 4 + 1024 : h->B.cap_dom;
        if (wp > (1ll << 30) || wd > (1ll << 30)) return fail(h, SZ_ERR_CAPACITY, errbits(e, buf, sizeof(buf)));
        if ((rc = set_pair_cap(h, (int)wp, (int)wd))) return rc;
    }
    if (e & ERR_POOL_CAP) {
        long long w = std::max<long long>(c.n_pool, h->B.cap_pool) * 2 + 1024;
        if ((rc = set_pool_cap(h, (int)std::min<long long>(w, 1ll << 28)))) return rc;
    }
    if (e & ERR_ROW_CAP) {
        long long w = (long long)c.want_rows * 5 / 4 + 1024;
        if ((rc = set_row_cap(h, (int)std::min<long long>(w, 1ll << 28)))) return rc;
    }
    if (e & ERR_FUSE_CAP) {
        if ((rc = set_fuse_cap(h, std::max(c.n_fuse, h->B.cap_fuse) * 2 + 1024))) return rc;
    }
    if (e & ERR_CELL_TABLE) return fail(h, SZ_ERR_UNSUPPORTED, "two-way coupling: a floe touches more than 2080 grid cells");
    if (e & ERR_SPILL_CAP) {  // every floe wider than the shared-memory table claims a block of 2048 entries
        if ((rc = set_spill_cap(h, std::max(c.n_spill, h->CB.cap_spill) + 8 * 2048))) return rc;
    }
    if (e & ERR_CREC_CAP) {
        if ((rc = set_crec_cap(h, std::max(c.n_crec, h->CB.cap_crec) * 2 + 1024))) return rc;
    }
    if (e & ERR_GHOST_CAP) {
        if ((rc = grow_floes(h, c.want_floes * 5 / 4 + 64, h->n_total))) return rc;
    }
    if (e & ERR_VERT_CAP) {
        if ((rc = grow_verts(h, c.want_verts * 5 / 4 + 64, h->n_verts))) return rc;
    }
    szk_clear_error(h->L, h->S);
    return SZ_OK;
}

static int pairs_hint(sz_handle *h) {
    long long g = h->last.n_cand > 0 ? (long long)h->last.n_cand * 2 : 8ll * h->n_total;
    long long q = 1024;  // a power of two: the grid sizes baked into the captured graph change rarely
    while (q < g) q <<= 1;
    return (int)std::min<long long>(q, h->B.cap_pairs);
}
static int floes_hint(sz_handle *h) { return std::min(h->S.cap_floes, h->n_total + h->n_total / 4 + 64); }

static void enqueue_ghosts(sz_handle *h) {
    // collisions.jl:1171-1172: east/west pass, then north/south pass
    if (h->hD.kind[2] == SZ_BOUNDARY_PERIODIC) szk_ghost_pass(h->L, h->S, h->B, 0, floes_hint(h));
    if (h->hD.kind[0] == SZ_BOUNDARY_PERIODIC) szk_ghost_pass(h->L, h->S, h->B, 1, floes_hint(h));
}

// one-way: the streaming kernel; two-way: the same integration with the floe -> cell registry, the registry
// sort, the floe ∩ cell areas and the per-cell pass (coupling.jl:1705-1738)
static void enqueue_coupling(sz_handle *h, const Launch &L) {
    if (!h->cfg.two_way_coupling_on) {
        szk_coupling(L, h->S, h->P);
        return;
    }
    szk_coupling_reg(L, h->S, h->CB, h->P);
    szk_cells_sort_and_clip(L, h->S, h->CB, h->P, h->CB.cap_crec);
    szk_cells_final(L, h->S, h->CB, h->P);
}

static double ev_ms(sz_handle *h, int a, int b) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[a], h->ev[b]) != cudaSuccess) {  // not recorded (graph replay without phase events)
        cudaGetLastError();
        return 0.0;
    }
    return (double)ms;
}

extern "C" int32_t sz_add_ghosts(sz_handle *h, int64_t *n_total) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "add_ghosts before set_domain");
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "add_ghosts before upload_floes");
    cudaSetDevice(h->cfg.device);
    for (int axis = 0; axis < 2; ++axis) {
        if (h->hD.kind[axis == 0 ? 2 : 0] != SZ_BOUNDARY_PERIODIC) continue;
        for (int attempt = 0;; ++attempt) {
            cudaEventRecord(h->ev[0], h->L.stream);
            szk_ghost_pass(h->L, h->S, h->B, axis, floes_hint(h));
            cudaEventRecord(h->ev[1], h->L.stream);
            int32_t rc = sync_counters(h);
            if (rc) return rc;
            CK(cudaGetLastError());
            if (!h->h_cnt->error) break;
            if (attempt >= 4) return fail(h, SZ_ERR_CAPACITY, "add_ghosts: capacity retry limit");
            if ((rc = handle_overflow(h, *h->h_cnt))) return rc;
        }
        h->ms[0] = (axis == 0 ? 0.0 : h->ms[0]) + ev_ms(h, 0, 1);
    }
    if (n_total) *n_total = h->n_total;
    return SZ_OK;
}

extern "C" int32_t sz_remove_ghosts(sz_handle *h) {
    if (!h || !h->have_floes) return fail(h, SZ_ERR_INVALID, "remove_ghosts before upload_floes");
    cudaSetDevice(h->cfg.device);
    szk_remove_ghosts(h->L, h->S, h->n_verts_init);
    int32_t rc = sync_counters(h);
    if (rc) return rc;
    CK(cudaGetLastError());
    return SZ_OK;
}

extern "C" int32_t sz_step_collisions(sz_handle *h) {
    if (!h || !h->have_domain) return fail(h, SZ_ERR_INVALID, "step_collisions before set_domain");
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "step_collisions before upload_floes");
    cudaSetDevice(h->cfg.device);
    for (int attempt = 0;; ++attempt) {
        cudaEventRecord(h->ev[0], h->L.stream);
        szk_collisions(h->L, h->S, h->B, h->P, floes_hint(h), pairs_hint(h), &h->ev[1]);
        int32_t rc = sync_counters(h);
        if (rc) return rc;
        CK(cudaGetLastError());
        if (!h->h_cnt->error) break;
        if (attempt >= 6) return fail(h, SZ_ERR_CAPACITY, "step_collisions: capacity retry limit");
        if ((rc = handle_overflow(h, *h->h_cnt))) return rc;
    }
    h->last = *h->h_cnt;
    h->ms[1] = ev_ms(h, 0, 1);
    h->ms[2] = ev_ms(h, 1, 2);
    h->ms[3] = ev_ms(h, 2, 3);
    return SZ_OK;
}

extern "C" int32_t sz_step_coupling(sz_handle *h) {
    if (!h || !h->have_domain || !h->have_fields) return fail(h, SZ_ERR_INVALID, "step_coupling before set_domain/set_fields");
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "step_coupling before upload_floes");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "step_coupling with ghosts present (call remove_ghosts)");
    cudaSetDevice(h->cfg.device);
    for (int attempt = 0;; ++attempt) {
        cudaEventRecord(h->ev[0], h->L.stream);
        enqueue_coupling(h, h->L);
        szk_apply_coupling_tags(h->L, h->S);
        cudaEventRecord(h->ev[1], h->L.stream);
        int32_t rc = sync_counters(h);
        if (rc) return rc;
        CK(cudaGetLastError());
        if (!h->h_cnt->error) break;
        if (attempt >= 4) return fail(h, SZ_ERR_CAPACITY, "step_coupling: capacity retry limit");
        if ((rc = handle_overflow(h, *h->h_cnt))) return rc;
    }
    h->n_crec_host = h->cfg.two_way_coupling_on ? h->h_cnt->n_crec : 0;
    h->ms[4] = ev_ms(h, 0, 1);
    return SZ_OK;
}

extern "C" int32_t sz_step_floe_properties(sz_handle *h, int64_t tstep) {
    (void)tstep;
    if (!h || !h->have_floes) return fail(h, SZ_ERR_INVALID, "step_floe_properties before upload_floes");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "step_floe_properties with ghosts present");
    cudaSetDevice(h->cfg.device);
    cudaEventRecord(h->ev[0], h->L.stream);
    szk_update(h->L, h->S, h->B, h->P);
    cudaEventRecord(h->ev[1], h->L.stream);
    CK(cudaStreamSynchronize(h->L.stream));
    CK(cudaGetLastError());
    h->ms[5] = ev_ms(h, 0, 1);
    return SZ_OK;
}

// ---- sz_step / sz_step_host ----------------------------------------------------------------------------
// Host buffers of one sz_step_host call.  The uploads are enqueued on stream_up in the order the kernels first
// need them and signal four events; the kernels wait for the group they read (or overwrite):
//   group 0  alpha, centroid, u, v, xi, area, mass  -> coupling.  First, because the
//            coupling kernel is forked at once: it must run beside the (cheap) broad phase, not beside the
//            latency-bound narrow phase, where the two slow each other down by more than the overlap gains
//   group 1  rmax, status                                                            -> broad phase
//   group 2  ring coordinates, height                                                -> narrow phase
//   group 3  moment, overarea, the AB2 history, stress / strain tensors               -> row assembly, update
// collision_force / collision_trq are NOT uploaded: timestep_collisions! zeroes them before anything reads them
// (collisions.jl:747-749, k_step_reset), they are outputs of every step.  fxOA, fyOA, trqOA, hflx_factor are outputs of
// a step that runs the coupling (coupling.jl:1583-1586) and are uploaded (group 3) only when it does not.
// The downloads go to stream_dn as soon as the producing kernel is done (collision totals after the row
// assembly, coupling outputs after the join, the rest after the update).

static int32_t enqueue_uploads(sz_handle *h, const sz_floe_soa *s, bool coupling_runs, bool partial = false) {
    Store &S = h->S;
    const int n = h->n_total;
    cudaStream_t st = h->stream_up;
    auto up = [&](double *dst, const double *src, size_t w) -> cudaError_t {
        if (src) return cudaMemcpyAsync(dst, src, sizeof(double) * w * n, cudaMemcpyHostToDevice, st);
        if (partial) return cudaSuccess;  // sz_step_host_partial: the device-resident value stands
        return cudaMemsetAsync(dst, 0, sizeof(double) * w * n, st);
    };
    CK(cudaEventRecord(h->ev_up_start, st));
    // group 0
    CK(up(S.alpha, s->alpha, 1)); CK(up(S.cx, s->centroid_x, 1)); CK(up(S.cy, s->centroid_y, 1)); CK(up(S.u, s->u, 1));
    CK(up(S.v, s->v, 1)); CK(up(S.xi, s->xi, 1)); CK(up(S.area, s->area, 1)); CK(up(S.mass, s->mass, 1));
    CK(cudaEventRecord(h->ev_up[0], st));
    // group 1
    CK(up(S.rmax, s->rmax, 1));
    if (s->status_tag) CK(cudaMemcpyAsync(S.status, s->status_tag, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    CK(cudaEventRecord(h->ev_up[1], st));
    // group 2
    if (h->n_verts > 0 && s->vert_xy) CK(cudaMemcpyAsync(S.verts, s->vert_xy, sizeof(double2) * (size_t)h->n_verts, cudaMemcpyHostToDevice, st));
    CK(up(S.height, s->height, 1));
    CK(cudaEventRecord(h->ev_up[2], st));
    // group 3
    if (!coupling_runs) {  // held between coupling steps (simulation.jl:151-154): inputs of the state update
        CK(up(S.fxOA, s->fxOA, 1)); CK(up(S.fyOA, s->fyOA, 1)); CK(up(S.trqOA, s->trqOA, 1)); CK(up(S.hflx, s->hflx_factor, 1));
    }
    CK(up(S.moment, s->moment, 1)); CK(up(S.overarea, s->overarea, 1));
    CK(up(S.p_dxdt, s->p_dxdt, 1)); CK(up(S.p_dydt, s->p_dydt, 1)); CK(up(S.p_dudt, s->p_dudt, 1));
    CK(up(S.p_dvdt, s->p_dvdt, 1)); CK(up(S.p_dxidt, s->p_dxidt, 1)); CK(up(S.p_dalphadt, s->p_dalphadt, 1));
    CK(up(S.stress_accum, s->stress_accum, 4)); CK(up(S.stress_instant, s->stress_instant, 4)); CK(up(S.strain, s->strain, 4));
    CK(cudaEventRecord(h->ev_up[3], st));
    return SZ_OK;
}

// stage: 0 after the row assembly, 1 after the coupling join, 2 after the state update
static int32_t enqueue_downloads(sz_handle *h, sz_floe_soa *s, int stage) {
    Store &S = h->S;
    const int n = h->n_init;
    cudaStream_t st = h->stream_dn;
    if (stage == 0 && s->collision_force) szk_interleave(h->L, S.cfx, S.cfy, h->d_cf_dn, n);  // [n][2] host layout
    CK(cudaEventRecord(h->ev_dn[stage], h->L.stream));
    CK(cudaStreamWaitEvent(st, h->ev_dn[stage], 0));
    auto dn = [&](double *dst, const double *src, size_t w) -> cudaError_t {
        if (!dst) return cudaSuccess;
        return cudaMemcpyAsync(dst, src, sizeof(double) * w * n, cudaMemcpyDeviceToHost, st);
    };
    if (stage == 0) {
        if (s->collision_force) CK(cudaMemcpyAsync(s->collision_force, h->d_cf_dn, sizeof(double2) * n, cudaMemcpyDeviceToHost, st));
        CK(dn(s->collision_trq, S.ctrq, 1)); CK(dn(s->overarea, S.overarea, 1));
        CK(dn(s->area, S.area, 1)); CK(dn(s->rmax, S.rmax, 1));  // not changed by a step
        if (s->id) CK(cudaMemcpyAsync(s->id, S.id, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
        if (s->ghost_id) CK(cudaMemcpyAsync(s->ghost_id, S.ghost_id, sizeof(long long) * n, cudaMemcpyDeviceToHost, st));
    } else if (stage == 1) {
        CK(dn(s->fxOA, S.fxOA, 1)); CK(dn(s->fyOA, S.fyOA, 1)); CK(dn(s->trqOA, S.trqOA, 1)); CK(dn(s->hflx_factor, S.hflx, 1));
        if (s->status_tag) CK(cudaMemcpyAsync(s->status_tag, S.status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));  // final after the tags
    } else {
        if (s->vert_xy && h->n_verts_init > 0)
            CK(cudaMemcpyAsync(s->vert_xy, S.verts, sizeof(double2) * (size_t)h->n_verts_init, cudaMemcpyDeviceToHost, st));
        CK(dn(s->centroid_x, S.cx, 1)); CK(dn(s->centroid_y, S.cy, 1)); CK(dn(s->alpha, S.alpha, 1));
        CK(dn(s->u, S.u, 1)); CK(dn(s->v, S.v, 1)); CK(dn(s->xi, S.xi, 1));
        CK(dn(s->height, S.height, 1)); CK(dn(s->mass, S.mass, 1)); CK(dn(s->moment, S.moment, 1));
        CK(dn(s->p_dxdt, S.p_dxdt, 1)); CK(dn(s->p_dydt, S.p_dydt, 1)); CK(dn(s->p_dudt, S.p_dudt, 1));
        CK(dn(s->p_dvdt, S.p_dvdt, 1)); CK(dn(s->p_dxidt, S.p_dxidt, 1)); CK(dn(s->p_dalphadt, S.p_dalphadt, 1));
        CK(dn(s->stress_accum, S.stress_accum, 4)); CK(dn(s->stress_instant, S.stress_instant, 4)); CK(dn(s->strain, S.strain, 4));
    }
    return SZ_OK;
}

// ---- slab hooks (sz_slab_*): what a slab rank adds to a timestep -----------------------------------------------------
// before the first kernel that reads a halo copy: wait for the partners' records of this epoch and scatter them
static void slab_unpack(sz_handle *h) {
    if (!h->slab.on) return;
    szk_slab_unpack(h->L, h->S, h->slab.dev, h->slab.epoch, h->slab.max_recv);
}
// publish my boundary floes as epoch `e` (and measure the displacement of the owned floes)
static void slab_push(sz_handle *h, int e) {
    if (!h->slab.on) return;
    szk_slab_push(h->L, h->S, h->slab.dev, e, h->slab.max_send);
}

// One whole timestep enqueued back to back; a single host synchronisation at the end (step_finish).
// everything a timestep enqueues, from add_ghosts! to the read-back of the counters (no host synchronisation)
static int32_t step_enqueue_direct(sz_handle *h) {
    cudaStream_t st = h->L.stream;
    const StepCur &cur = h->cur;
    const int32_t do_coupling = cur.do_coupling;
    const HostIO *io = cur.has_io ? &cur.io : nullptr;
    const bool periodic = h->hD.kind[2] == SZ_BOUNDARY_PERIODIC || h->hD.kind[0] == SZ_BOUNDARY_PERIODIC;
    const bool fork = do_coupling && !h->cfg.two_way_coupling_on;
    auto fork_coupling = [&]() {
        // fork: coupling only needs the floe state after add_ghosts! wrapped parents into the domain; it
        // reads nothing the collision kernels write, so it runs beside them on a low-priority stream and
        // fills the issue slots the latency-bound narrow phase leaves idle
        cudaEventRecord(h->ev_fork, st);
        cudaStreamWaitEvent(h->stream2, h->ev_fork, 0);
        Launch L2 = h->L;
        L2.stream = h->stream2;
        // measured (profiles/README.md): capping coupling to 2-4 resident blocks per SM so that it runs beside the
        // narrow phase slows the latter more than the overlap gains; it fills the GPU during the broad phase
        L2.coupling_blocks_per_sm = 0;
        sz_record(L2, h->ev_c0, h->stream2);
        szk_coupling(L2, h->S, h->P);
        sz_record(L2, h->ev_c1, h->stream2);
        cudaEventRecord(h->ev_join, h->stream2);
    };
    if (!cur.coupling_only) {
        if (io) {
            // add_ghosts! copies every scalar of a parent into its ghost (collisions.jl:1017-1047): with periodic
            // walls the step starts when all uploads have landed
            for (int g = 0; g < (periodic ? 4 : 1); ++g) CK(cudaStreamWaitEvent(st, h->ev_up[g], 0));  // non-periodic: group 0
        }
        sz_record(h->L, h->ev[0], st);
        bool forked = h->cpl_prelaunched;
        if (h->slab.on) {
            // the coupling of an owned floe reads only that floe's own state: without a periodic wall (add_ghosts! wraps
            // parents first) it starts before this rank waits for its neighbours' records
            if (fork && !forked && !periodic) {
                fork_coupling();
                forked = true;
            }
            // (host arrays every step: step_setup has already published what the host uploaded)
            // the halo update lands on top of the uploaded (stale) copies: everything it overwrites (centroid, velocities,
            // status, height, alpha, rings) is in upload groups 0-2; group 3 (17 MB of AB2 history and tensors) may still be
            // in flight
            if (io) CK(cudaStreamWaitEvent(st, h->ev_up[2], 0));
            slab_unpack(h);
        }
        if (!cur.keep_ghosts) enqueue_ghosts(h);
        sz_record(h->L, h->ev[1], st);
        if (fork && !forked) fork_coupling();
        // the ghost count of this step is not known on the host: size grids from the capacity-bounded hint
        cudaEvent_t waits[2] = {h->ev_up[2], h->ev_up[3]};
        if (io) CK(cudaStreamWaitEvent(st, h->ev_up[1], 0));
        // PDL pays when the chain has the GPU to itself (collisions-only step, 100 k floes: 0.83 ms).  Beside the
        // overlapped coupling the faster chain only starves the low-priority coupling kernel, which the update then
        // waits for (r3d: 1.17 ms against 1.09 ms) — so: on for steps without coupling (SZ_PDL=0 / 1 forces it)
        {
            static const char *env = getenv("SZ_PDL");
            h->L.pdl = env ? atoi(env) : (fork ? 0 : 1);
        }
        szk_collisions(h->L, h->S, h->B, h->P, floes_hint(h), pairs_hint(h), &h->ev[2], io ? waits : nullptr);
        szk_remove_ghosts(h->L, h->S, h->n_verts_init);
        sz_record(h->L, h->ev[5], st);
        if (io) { int32_t rc2 = enqueue_downloads(h, io->out, 0); if (rc2) return rc2; }
    }
    if (fork) {
        cudaStreamWaitEvent(st, h->ev_join, 0);  // join
        szk_apply_coupling_tags(h->L, h->S);
    } else if (do_coupling) {
        // two-way coupling writes ocean.hflx_factor for the NEXT step: it must not run ahead of a collision
        // phase that may still overflow and be repeated, so it stays in order on this stream
        sz_record(h->L, h->ev_c0, st);
        enqueue_coupling(h, h->L);
        sz_record(h->L, h->ev_c1, st);
        szk_apply_coupling_tags(h->L, h->S);
    }
    if (io) { int32_t rc2 = enqueue_downloads(h, io->out, 1); if (rc2) return rc2; }
    sz_record(h->L, h->ev[6], st);
    szk_update(h->L, h->S, h->B, h->P);
    sz_record(h->L, h->ev[7], st);
    if (io) { int32_t rc2 = enqueue_downloads(h, io->out, 2); if (rc2) return rc2; }
    if (h->slab.on && !cur.slab_host_mode) slab_push(h, h->slab.epoch + 1);  // device-resident: publish behind the update
    CK(cudaMemcpyAsync(h->h_cnt, h->S.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    return SZ_OK;
}

static int32_t step_enqueue(sz_handle *h) {
    cudaStream_t st = h->L.stream;
    const StepCur &cur = h->cur;
    // Device-resident steps replay a captured CUDA graph of the ~39 launches on two streams: most of them are
    // small dependent kernels (scans, checks, the broad phase) whose launch gaps then shrink; the graph is
    // captured again whenever a pointer or parameter baked into its nodes changed (h->gen) or a grid-size hint
    // moved to another bucket.  sz_step_host (host copies with caller pointers) enqueues directly, and so do LARGE
    // fields: measured (tools/ab_small.sh) the graph gives +22 % at 1 k floes and +12 % at 10 k, but -7 % at 100 k,
    // where the step is bound by three long kernels (the captured nodes keep their stream priorities: checked with
    // cudaGraphKernelNodeGetAttribute; the external timing-event nodes in the chain are the suspected cost).
    // Slab ranks (epoch-numbered push / unpack kernels) and coupling-only repairs launch directly.
    if (!cur.has_io && !h->graph_off && !h->cpl_prelaunched && !h->slab.on && !cur.coupling_only && h->n_init <= h->graph_max_floes) {
        const int fh = floes_hint(h), ph = pairs_hint(h);
        const bool fresh = !cur.keep_ghosts && h->gexec && h->gkey_gen == h->gen && h->gkey_coupling == cur.do_coupling &&
                           h->gkey_floes == fh && h->gkey_pairs == ph;
        if (!fresh) {
            if (h->gexec) cudaGraphExecDestroy(h->gexec);
            h->gexec = nullptr;
            const long long before = szk_launch_count(false);
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            h->L.capturing = true;
            const int32_t crc = step_enqueue_direct(h);
            h->L.capturing = false;
            cudaGraph_t g = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(st, &g);
            const long long captured = szk_launch_count(false) - before;
            szk_count_launches((int)-captured);  // nothing has run yet
            if (crc == SZ_OK && ce == cudaSuccess && g && cudaGraphInstantiate(&h->gexec, g, 0) == cudaSuccess) {
                h->gkey_gen = cur.keep_ghosts ? 0 : h->gen;  // a repair attempt's graph (no ghost pass) is not reused
                h->gkey_coupling = cur.do_coupling; h->gkey_floes = fh; h->gkey_pairs = ph;
                h->graph_launches = (int)captured;
            } else {
                h->gexec = nullptr;
                h->graph_off = true;  // fall back to direct launches for the life of the handle
                cudaGetLastError();
            }
            if (g) cudaGraphDestroy(g);
        }
        if (h->gexec) {
            CK(cudaGraphLaunch(h->gexec, st));
            szk_count_launches(h->graph_launches);
            return SZ_OK;
        }
    }
    return step_enqueue_direct(h);
}

// Wait for the step in flight; a buffer overflow is repaired and the step repeated:
//  * collision-phase overflow: FROM THE COLLISIONS — the ghosts of the failed attempt are complete and stay (every
//    kernel after the overflow returned at once, including the ghost removal).  Running add_ghosts! again would start
//    from parents the first pass has already wrapped into the domain and can number the images of a corner floe in
//    another order than a clean step does (rows and totals then differ from the reference in their last bit);
//  * the ghost pass itself overflowed: nothing of it was committed for that axis; start over;
//  * the floe -> cell registry of the two-way coupling overflowed: the collisions of this step are COMPLETE (rows
//    added to overarea, moving walls advanced, ghosts removed): only coupling, tags and update are repeated.
static int32_t step_finish(sz_handle *h) {
    cudaStream_t st = h->L.stream;
    StepCur &cur = h->cur;
    const bool io = cur.has_io;
    for (;;) {
        CK(cudaStreamSynchronize(st));
        if (io) {
            CK(cudaEventRecord(h->ev_dn_end, h->stream_dn));
            CK(cudaStreamSynchronize(h->stream_dn));
            if (getenv("SZ_DEBUG_HOSTIO")) {
                float t[13] = {0};
                cudaEventElapsedTime(&t[12], h->ev_up_start, h->ev[1]);
                for (int g = 0; g < 4; ++g) cudaEventElapsedTime(&t[g], h->ev_up_start, h->ev_up[g]);
                cudaEventElapsedTime(&t[4], h->ev_up_start, h->ev[0]);
                cudaEventElapsedTime(&t[5], h->ev_up_start, h->ev[2]);
                cudaEventElapsedTime(&t[6], h->ev_up_start, h->ev[3]);
                cudaEventElapsedTime(&t[7], h->ev_up_start, h->ev[5]);
                cudaEventElapsedTime(&t[8], h->ev_up_start, h->ev[7]);
                cudaEventElapsedTime(&t[9], h->ev_up_start, h->ev_dn_end);
                cudaEventElapsedTime(&t[10], h->ev_up_start, h->ev_c0);
                cudaEventElapsedTime(&t[11], h->ev_up_start, h->ev_c1);
                fprintf(stderr, "[dev %d] hostio ms: up groups %.3f %.3f %.3f %.3f | step start %.3f ghosts end %.3f broad end %.3f narrow end %.3f rows end %.3f update end %.3f | coupling %.3f..%.3f | last download %.3f\n",
                        h->cfg.device, t[0], t[1], t[2], t[3], t[4], t[12], t[5], t[6], t[7], t[8], t[10], t[11], t[9]);
            }
        }
        CK(cudaGetLastError());
        if (!h->h_cnt->error) break;
        // every kernel after the overflow returned at once
        Counters c = *h->h_cnt;
        h->n_total = c.n_total;
        h->n_verts = c.n_verts;
        if (getenv("SZ_DEBUG_RETRY")) fprintf(stderr, "[dev %d] step repeated: error bits 0x%x (attempt %d)\n", h->cfg.device, c.error, cur.attempt);
        if (cur.attempt >= 6) return fail(h, SZ_ERR_CAPACITY, "step: capacity retry limit");
        int32_t rc = handle_overflow(h, c);
        if (rc) return rc;
        const uint32_t ghost_bits = ERR_GHOST_CAP | ERR_VERT_CAP | ERR_GHOST_SLOTS;
        const uint32_t coupling_bits = ERR_CREC_CAP | ERR_CELL_TABLE | ERR_SPILL_CAP;
        if (cur.coupling_only || (c.error & ~coupling_bits) == 0) {
            if (!cur.coupling_only) h->last_ok_collisions = c;
            cur.coupling_only = true;
        } else if (c.error & ghost_bits) {
            szk_remove_ghosts(h->L, h->S, h->n_verts_init);
            cur.keep_ghosts = false;
        } else {
            cur.keep_ghosts = true;
        }
        CK(cudaStreamSynchronize(st));
        h->n_total = h->n_init;
        h->n_verts = h->n_verts_init;
        cur.attempt++;
        if ((rc = step_enqueue(h))) return rc;
    }
    const int32_t do_coupling = cur.do_coupling;
    if (!cur.coupling_only) h->last = *h->h_cnt;
    else {  // the collision counters of the first attempt stand; take the registry size of the repeated coupling
        Counters keep = h->last_ok_collisions;
        keep.n_crec = h->h_cnt->n_crec;
        keep.slab_disp = h->h_cnt->slab_disp;
        h->last = keep;
    }
    h->n_total = h->n_init;
    h->n_verts = h->n_verts_init;
    h->cpl_prelaunched = false;
    if (h->slab.on) {
        if (cur.slab_host_mode) h->slab.pushed = h->slab.epoch;
        else h->slab.pushed = h->slab.epoch + 1;
        h->slab.epoch += 1;
    }
    if (getenv("SZ_DEBUG_COUNTS"))
        fprintf(stderr, "counts: cand %d kept %d dom %d order %d force %d mid %d large %d overlap %d rows %d\n", h->last.n_cand,
                h->last.n_kept, h->last.n_dom, h->last.n_order, h->last.n_force, h->last.n_mid, h->last.n_large, h->last.n_overlap,
                h->last.n_rows);
    if (do_coupling) h->n_crec_host = h->cfg.two_way_coupling_on ? h->h_cnt->n_crec : 0;
    h->ms[0] = ev_ms(h, 0, 1);
    h->ms[1] = ev_ms(h, 1, 2);
    h->ms[2] = ev_ms(h, 2, 3);
    h->ms[3] = ev_ms(h, 3, 5);
    h->ms[4] = 0.0;  // coupling kernel time on its own stream (overlaps the collision phases)
    if (do_coupling) {
        float cms = 0.f;
        if (cudaEventElapsedTime(&cms, h->ev_c0, h->ev_c1) != cudaSuccess) {
            cudaGetLastError();
            cms = 0.f;
        }
        h->ms[4] = (double)cms;
    }
    h->ms[5] = ev_ms(h, 6, 7);
    h->ms[6] = ev_ms(h, 0, 7);
    h->ms[7] = (double)szk_launch_count(true);  // kernels launched since the previous sz_step returned (incl. halo pack/unpack)
    return SZ_OK;
}

// First half of a step: the uploads and — for a slab rank stepping on host arrays — the publication of the uploaded
// boundary floes.  A process that drives several ranks runs this on ALL of them before any rank enqueues the kernel
// that waits for its neighbours (step_enqueue): nothing that may synchronise the device (allocations, pageable copies)
// then stands between a waiting kernel and the publication it waits for.
static int32_t step_setup(sz_handle *h, int32_t do_coupling, const HostIO *io, bool slab_host_mode) {
    StepCur &cur = h->cur;
    cur.do_coupling = do_coupling;
    cur.has_io = io != nullptr;
    if (io) cur.io = *io;
    cur.keep_ghosts = cur.coupling_only = false;
    cur.slab_host_mode = slab_host_mode;
    cur.partial = h->next_partial;
    h->next_partial = false;
    cur.attempt = 0;
    if (h->slab.on && slab_host_mode && h->slab.pushed >= h->slab.epoch) h->slab.epoch = h->slab.pushed + 1;  // a fresh epoch for the re-publication
    if (io && io->in) {
        int32_t rc = enqueue_uploads(h, io->in, do_coupling != 0, cur.partial);
        if (rc) return rc;
    } else if (io && cur.partial) {  // nothing to upload: the events the kernels wait for are recorded at once
        CK(cudaEventRecord(h->ev_up_start, h->stream_up));
        for (int g = 0; g < 4; ++g) CK(cudaEventRecord(h->ev_up[g], h->stream_up));
    }
    // host arrays: what the host uploaded is what the neighbours must see.  Device-resident: the previous step published
    // behind its update — unless that step ran on host arrays (or this is the first step after a refresh-less rebuild)
    if (h->slab.on && (slab_host_mode || h->slab.pushed < h->slab.epoch)) {
        if (io && io->in) CK(cudaStreamWaitEvent(h->L.stream, h->ev_up[2], 0));  // what k_slab_push reads is in groups 0-2
        slab_push(h, h->slab.epoch);
        h->slab.pushed = h->slab.epoch;
    }
    return SZ_OK;
}

static int32_t step_begin(sz_handle *h, int32_t do_coupling, const HostIO *io, bool slab_host_mode) {
    int32_t rc = step_setup(h, do_coupling, io, slab_host_mode);
    if (rc) return rc;
    return step_enqueue(h);
}

static int32_t step_impl(sz_handle *h, int32_t do_coupling, const HostIO *io) {
    int32_t rc = step_begin(h, do_coupling, io, false);
    if (rc) return rc;
    return step_finish(h);
}

extern "C" int32_t sz_coupling_begin(sz_handle *h) {
    if (!h) return SZ_ERR_INVALID;
    if (!h->have_domain || !h->have_floes || !h->have_fields) return fail(h, SZ_ERR_INVALID, "coupling_begin before set_domain/set_fields/upload_floes");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "coupling_begin with ghosts present");
    if (h->cfg.two_way_coupling_on || h->P.per_x || h->P.per_y || h->cpl_prelaunched) return SZ_OK;  // the order matters there
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->L.stream;
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(h->stream2, h->ev_fork, 0));
    if (h->up_pending) {  // sz_upload_state_begin: the coupling inputs are upload group 0
        CK(cudaStreamWaitEvent(h->stream2, h->ev_up[0], 0));
    }
    Launch L2 = h->L;
    L2.stream = h->stream2;
    L2.coupling_blocks_per_sm = 0;
    sz_record(L2, h->ev_c0, h->stream2);
    szk_coupling(L2, h->S, h->P);
    sz_record(L2, h->ev_c1, h->stream2);
    CK(cudaEventRecord(h->ev_join, h->stream2));
    CK(cudaGetLastError());
    h->cpl_prelaunched = true;
    return SZ_OK;
}

extern "C" int32_t sz_step(sz_handle *h, int64_t tstep, int32_t do_coupling) {
    (void)tstep;
    if (!h || !h->have_domain || !h->have_floes) return fail(h, SZ_ERR_INVALID, "step before set_domain/upload_floes");
    if (do_coupling && !h->have_fields) return fail(h, SZ_ERR_INVALID, "step with coupling before set_fields");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "step with ghosts present (call remove_ghosts)");
    if (h->cpl_prelaunched && !do_coupling) return fail(h, SZ_ERR_INVALID, "step without coupling after sz_coupling_begin");
    if (h->slab.on) return fail(h, SZ_ERR_INVALID, "this handle is a slab rank: step it with sz_slab_step");
    cudaSetDevice(h->cfg.device);
    return step_impl(h, do_coupling, nullptr);
}

// sz_upload_state + sz_step + sz_download_floes in ONE call: the host <-> device copies run on their own
// streams beside the kernels (see HostIO).  `in` and `out` may point to the same arrays.
static int32_t check_host_state(sz_handle *h, const char *who, int32_t do_coupling, const sz_floe_soa *in) {
    char msg[160];
    if (!h->have_domain || !h->have_floes) { snprintf(msg, sizeof(msg), "%s before set_domain/upload_floes", who); return fail(h, SZ_ERR_INVALID, msg); }
    if (do_coupling && !h->have_fields) { snprintf(msg, sizeof(msg), "%s with coupling before set_fields", who); return fail(h, SZ_ERR_INVALID, msg); }
    if (h->n_total != h->n_init) { snprintf(msg, sizeof(msg), "%s with ghosts present (call remove_ghosts)", who); return fail(h, SZ_ERR_INVALID, msg); }
    if (in) {
        if (in->n != h->n_total || in->n_init != h->n_init) { snprintf(msg, sizeof(msg), "%s: floe count differs from the resident store", who); return fail(h, SZ_ERR_INVALID, msg); }
        if (!in->centroid_x || !in->centroid_y || !in->area || !in->rmax || !in->vert_xy) { snprintf(msg, sizeof(msg), "%s: geometry arrays are required", who); return fail(h, SZ_ERR_INVALID, msg); }
        if (in->vert_offsets && in->vert_offsets[h->n_total] != h->n_verts) { snprintf(msg, sizeof(msg), "%s: vertex count differs from the resident store", who); return fail(h, SZ_ERR_INVALID, msg); }
    }
    return SZ_OK;
}

extern "C" int32_t sz_upload_state_begin(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in) {
    if (!h || !in) return SZ_ERR_INVALID;
    int32_t rc = check_host_state(h, "upload_state_begin", do_coupling, in);
    if (rc) return rc;
    if (h->up_pending) return fail(h, SZ_ERR_INVALID, "upload_state_begin: an upload is already pending");
    cudaSetDevice(h->cfg.device);
    rc = enqueue_uploads(h, in, do_coupling != 0);
    if (rc) {
        cudaStreamSynchronize(h->stream_up);
        return rc;
    }
    h->up_pending = 1 + (do_coupling != 0);
    return SZ_OK;
}

// sz_upload_state + sz_step + sz_download_floes in ONE call: the host <-> device copies run on their own
// streams beside the kernels (see HostIO).  `in` and `out` may point to the same arrays; in == NULL: the uploads were
// enqueued by sz_upload_state_begin.
static int32_t prepare_step_host(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out) {
    if (!h || !out) return SZ_ERR_INVALID;
    {
        int32_t rc = check_host_state(h, "step_host", do_coupling, in);
        if (rc) return rc;
    }
    if (!in && h->up_pending != 1 + (do_coupling != 0)) return fail(h, SZ_ERR_INVALID, "step_host without input arrays needs sz_upload_state_begin with the same do_coupling");
    if (in && h->up_pending) return fail(h, SZ_ERR_INVALID, "step_host: an upload of sz_upload_state_begin is pending (pass in = NULL)");
    h->up_pending = 0;
    if (h->cpl_prelaunched && !do_coupling) return fail(h, SZ_ERR_INVALID, "step_host without coupling after sz_coupling_begin");
    if ((int)h->h_vcount.size() != h->n_init) return fail(h, SZ_ERR_INVALID, "step_host: no ring table of the resident floes");
    cudaSetDevice(h->cfg.device);
    const int n = h->n_init;
    if (h->cf_cap < n) {
        dfree(h->d_cf_dn);
        CK(dalloc(&h->d_cf_dn, (size_t)n));
        h->cf_cap = n;
    }
    return SZ_OK;
}

// host-side tables of the download (same as sz_download_floes without ghosts).  They only change with the floe list:
// a caller that passes the same arrays every step (the shim does) gets them written once — at 100 k floes the three
// loops cost 0.15 ms of host time per step, behind the last download.
static void finish_host_tables(sz_handle *h, sz_floe_soa *out) {
    const int n = h->n_init;
    out->n = n;
    out->n_init = n;
    if (h->tables_gen == h->gen && h->tables_ptr[0] == out->vert_offsets && h->tables_ptr[1] == out->mc_offsets &&
        h->tables_ptr[2] == out->ghost_offsets) {
        // same arrays as last time; spot-check that the caller did not recycle the addresses for fresh arrays
        const bool v_ok = !out->vert_offsets || (out->vert_offsets[n] == h->n_verts_init && out->vert_offsets[n / 2] == h->tables_mid[0]);
        const bool m_ok = !out->mc_offsets || (out->mc_offsets[n] == h->h_mc_off[n] && out->mc_offsets[n / 2] == h->tables_mid[1]);
        if (v_ok && m_ok) return;
    }
    if (out->vert_offsets) {
        long long o = 0;
        for (int i = 0; i < n; ++i) { out->vert_offsets[i] = o; o += h->h_vcount[i]; }
        out->vert_offsets[n] = o;
    }
    if (out->mc_offsets) for (int i = 0; i <= n; ++i) out->mc_offsets[i] = h->h_mc_off[i];
    if (out->ghost_offsets) memset(out->ghost_offsets, 0, sizeof(int64_t) * ((size_t)n + 1));
    h->tables_gen = h->gen;
    h->tables_mid[0] = out->vert_offsets ? out->vert_offsets[n / 2] : 0;
    h->tables_mid[1] = out->mc_offsets ? out->mc_offsets[n / 2] : 0;
    h->tables_ptr[0] = out->vert_offsets; h->tables_ptr[1] = out->mc_offsets; h->tables_ptr[2] = out->ghost_offsets;
}

extern "C" int32_t sz_step_host(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out) {
    (void)tstep;
    if (h && h->slab.on) return fail(h, SZ_ERR_INVALID, "this handle is a slab rank: step it with sz_slab_step_host");
    int32_t rc = prepare_step_host(h, do_coupling, in, out);
    if (rc) return rc;
    HostIO io = {in, out};
    rc = step_impl(h, do_coupling, &io);
    if (rc) {
        cudaStreamSynchronize(h->stream_up);
        cudaStreamSynchronize(h->stream_dn);
        return rc;
    }
    finish_host_tables(h, out);
    return SZ_OK;
}

extern "C" int32_t sz_step_host_partial(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out) {
    (void)tstep;
    if (!h || !out) return SZ_ERR_INVALID;
    if (h->slab.on) return fail(h, SZ_ERR_INVALID, "this handle is a slab rank: step it with sz_slab_step_host");
    if (!h->have_domain || !h->have_floes) return fail(h, SZ_ERR_INVALID, "step_host_partial before set_domain/upload_floes");
    if (do_coupling && !h->have_fields) return fail(h, SZ_ERR_INVALID, "step_host_partial with coupling before set_fields");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "step_host_partial with ghosts present (call remove_ghosts)");
    if (h->up_pending || h->cpl_prelaunched) return fail(h, SZ_ERR_INVALID, "step_host_partial: an upload or a coupling of the split calls is pending");
    if (in && (in->n != h->n_total || in->n_init != h->n_init)) return fail(h, SZ_ERR_INVALID, "step_host_partial: floe count differs from the resident store");
    if ((int)h->h_vcount.size() != h->n_init) return fail(h, SZ_ERR_INVALID, "step_host_partial: no ring table of the resident floes");
    cudaSetDevice(h->cfg.device);
    const int n = h->n_init;
    if (h->cf_cap < n) {
        dfree(h->d_cf_dn);
        CK(dalloc(&h->d_cf_dn, (size_t)n));
        h->cf_cap = n;
    }
    HostIO io = {in, out};
    h->next_partial = true;
    int32_t rc = step_impl(h, do_coupling, &io);
    h->next_partial = false;
    if (rc) {
        cudaStreamSynchronize(h->stream_up);
        cudaStreamSynchronize(h->stream_dn);
        return rc;
    }
    finish_host_tables(h, out);
    return SZ_OK;
}

// ---- slab backend (sz_slab_backend.h): a rank's side of the peer-memory halo update ---------------------------------
struct WireData {
    int32_t pid, device, rank, pad;
    unsigned long long base, gen;
    cudaIpcMemHandle_t ipc;
    long long off_ready, off_ack, off_stage[2], bytes;
};
static_assert(sizeof(WireData) <= sizeof(SlabWire), "SlabWire too small");

// start of a (re)build or the end of the slab: nothing is in flight any more.  final: also unmap the partners' arenas.
int32_t szb_release_peers(sz_handle *h, int32_t final) {
    if (!h) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    cudaStreamSynchronize(h->L.stream);
    if (final) {
        for (auto &m : h->slab.ipc_open) cudaIpcCloseMemHandle(m.ptr);
        h->slab.ipc_open.clear();
    }
    for (int k = 0; k < SZ_SLAB_MAX_PARTNERS; ++k) {
        h->slab.dev.p[k].r_stage[0] = h->slab.dev.p[k].r_stage[1] = nullptr;
        h->slab.dev.p[k].r_ready = h->slab.dev.p[k].r_ack = nullptr;
    }
    return SZ_OK;
}

static void slab_free(sz_handle *h) {
    h->slab.on = false;  // the arena and the (grow-only) list arrays stay
}

int32_t szb_configure(sz_handle *h, const SlabLists *l, SlabWire *wire_out) {
    if (!h || !l || l->n_partners < 0 || (l->n_partners > 0 && !wire_out)) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "slab: configure before upload_floes");
    if (l->n_partners > SZ_SLAB_MAX_PARTNERS) return fail(h, SZ_ERR_UNSUPPORTED, "slab: more than 16 exchange partners");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "slab: configure with ghosts present");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    szb_release_peers(h, 0);
    slab_free(h);
    SlabState &B = h->slab;
    const int np = l->n_partners, n = h->n_init;
    const long long ns = np ? l->send_off[np] : 0, nr = np ? l->recv_off[np] : 0;
    std::vector<int> sidx((size_t)std::max<long long>(ns, 1)), ridx((size_t)std::max<long long>(nr, 1));
    std::vector<long long> svoff((size_t)std::max<long long>(ns, 1)), rvoff((size_t)std::max<long long>(nr, 1));
    std::vector<long long> sbytes(np, 0), rbytes(np, 0);
    memset(&B.dev, 0, sizeof(B.dev));
    B.max_send = B.max_recv = 0;
    B.send_bytes_total = 0;
    for (int p = 0; p < np; ++p) {
        long long v = 0;
        for (long long k = l->send_off[p]; k < l->send_off[p + 1]; ++k) {
            if (l->send_idx[k] < 0 || l->send_idx[k] >= n) return fail(h, SZ_ERR_INVALID, "slab: send index out of range");
            sidx[k] = (int)l->send_idx[k];
            svoff[k] = v;
            v += h->h_vcount[sidx[k]];
        }
        sbytes[p] = 64 * (l->send_off[p + 1] - l->send_off[p]) + 16 * v;
        v = 0;
        for (long long k = l->recv_off[p]; k < l->recv_off[p + 1]; ++k) {
            if (l->recv_idx[k] < 0 || l->recv_idx[k] >= n) return fail(h, SZ_ERR_INVALID, "slab: receive index out of range");
            ridx[k] = (int)l->recv_idx[k];
            rvoff[k] = v;
            v += h->h_vcount[ridx[k]];
        }
        rbytes[p] = 64 * (l->recv_off[p + 1] - l->recv_off[p]) + 16 * v;
        SlabPartnerDev &d = B.dev.p[p];
        d.send_off = (int)l->send_off[p]; d.send_n = (int)(l->send_off[p + 1] - l->send_off[p]);
        d.recv_off = (int)l->recv_off[p]; d.recv_n = (int)(l->recv_off[p + 1] - l->recv_off[p]);
        B.max_send = std::max(B.max_send, d.send_n);
        B.max_recv = std::max(B.max_recv, d.recv_n);
        B.send_bytes_total += sbytes[p];
    }
    B.send_bytes = sbytes;
    // arena: [0, 4096) flags — ready[16] | ack[16] | push_count[16] | unpack_count[16] — then two halves per partner
    const size_t FLAGS = 4096;
    std::vector<size_t> off0(np), off1(np);
    size_t total = FLAGS;
    for (int p = 0; p < np; ++p) {
        size_t b = ((size_t)rbytes[p] + 255) / 256 * 256 + 256;
        off0[p] = total; total += b;
        off1[p] = total; total += b;
    }
    // arenas replaced at the PREVIOUS rebuild: every importer has switched (and closed) since
    for (unsigned char *p : B.retired) cudaFree(p);
    B.retired.clear();
    if (total > B.arena_bytes || !B.arena) {
        if (B.arena) B.retired.push_back(B.arena);
        B.arena = nullptr;
        const size_t cap = total + total / 2 + (1u << 20);  // headroom: the lists change a little at every rebuild
        CK(cudaMalloc((void **)&B.arena, cap));
        B.arena_bytes = cap;
        B.arena_gen++;
    }
    CK(cudaMemset(B.arena, 0, FLAGS));
    if (B.cap_lists < std::max(ns, nr) || !B.d_send_idx) {
        dfree(B.d_send_idx); dfree(B.d_recv_idx); dfree(B.d_send_voff); dfree(B.d_recv_voff);
        B.cap_lists = std::max(ns, nr) * 2 + 4096;
        CK(dalloc(&B.d_send_idx, (size_t)B.cap_lists)); CK(dalloc(&B.d_recv_idx, (size_t)B.cap_lists));
        CK(dalloc(&B.d_send_voff, (size_t)B.cap_lists)); CK(dalloc(&B.d_recv_voff, (size_t)B.cap_lists));
    }
    if (B.cap_owned < n || !B.d_owned) {
        dfree(B.d_owned); dfree(B.d_refx); dfree(B.d_refy);
        B.cap_owned = (long long)n + n / 8 + 1024;
        CK(dalloc(&B.d_owned, (size_t)B.cap_owned)); CK(dalloc(&B.d_refx, (size_t)B.cap_owned)); CK(dalloc(&B.d_refy, (size_t)B.cap_owned));
    }
    if (ns > 0) {
        CK(cudaMemcpy(B.d_send_idx, sidx.data(), sizeof(int) * ns, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B.d_send_voff, svoff.data(), sizeof(long long) * ns, cudaMemcpyHostToDevice));
    }
    if (nr > 0) {
        CK(cudaMemcpy(B.d_recv_idx, ridx.data(), sizeof(int) * nr, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B.d_recv_voff, rvoff.data(), sizeof(long long) * nr, cudaMemcpyHostToDevice));
    }
    if (n > 0) {
        CK(cudaMemcpy(B.d_owned, l->owned, (size_t)n, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(B.d_refx, h->S.cx, sizeof(double) * n, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(B.d_refy, h->S.cy, sizeof(double) * n, cudaMemcpyDeviceToDevice));
    }
    szk_slab_reset_disp(h->L, h->S);
    CK(cudaStreamSynchronize(h->L.stream));
    h->h_cnt->slab_disp = 0ull;
    int *flags = (int *)B.arena;
    B.dev.n_partners = np;
    B.dev.send_idx = B.d_send_idx; B.dev.recv_idx = B.d_recv_idx; B.dev.send_voff = B.d_send_voff; B.dev.recv_voff = B.d_recv_voff;
    B.dev.push_count = flags + 32; B.dev.unpack_count = flags + 48;
    B.dev.owned = B.d_owned; B.dev.refx = B.d_refx; B.dev.refy = B.d_refy;
    B.dev.period_x = l->period_x; B.dev.period_y = l->period_y;
    {
        const char *ts = getenv("SZ_SLAB_TIMEOUT_S");
        double sec = ts ? atof(ts) : 20.0;
        if (!(sec > 0.0)) sec = 20.0;
        B.dev.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    cudaIpcMemHandle_t ipc;
    memset(&ipc, 0, sizeof(ipc));
    if (cudaIpcGetMemHandle(&ipc, B.arena) != cudaSuccess) {  // same-process partners do not need it
        cudaGetLastError();
        memset(&ipc, 0, sizeof(ipc));
    }
    for (int p = 0; p < np; ++p) {
        SlabPartnerDev &d = B.dev.p[p];
        d.l_stage[0] = (const double *)(B.arena + off0[p]);
        d.l_stage[1] = (const double *)(B.arena + off1[p]);
        d.l_ready = flags + p;
        d.l_ack = flags + 16 + p;
        WireData w;
        memset(&w, 0, sizeof(w));
        w.pid = (int32_t)getpid(); w.device = h->cfg.device; w.rank = l->rank;
        w.base = (unsigned long long)(uintptr_t)B.arena;
        w.gen = B.arena_gen;
        w.ipc = ipc;
        w.off_ready = (long long)(sizeof(int) * p); w.off_ack = (long long)(sizeof(int) * (16 + p));
        w.off_stage[0] = (long long)off0[p]; w.off_stage[1] = (long long)off1[p];
        w.bytes = rbytes[p];
        memset(&wire_out[p], 0, sizeof(SlabWire));
        memcpy(&wire_out[p], &w, sizeof(w));
    }
    B.rank = l->rank;
    B.n_partners = np;
    B.epoch = 1;
    B.pushed = 0;
    B.on = true;
    h->gen++;
    return SZ_OK;
}

int32_t szb_connect(sz_handle *h, const SlabWire *peer_wire) {
    if (!h || !h->slab.on) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    SlabState &B = h->slab;
    std::vector<std::pair<unsigned long long, unsigned char *>> mapped;  // one mapping per partner process arena
    for (int p = 0; p < B.n_partners; ++p) {
        WireData w;
        memcpy(&w, &peer_wire[p], sizeof(w));
        if (w.bytes != B.send_bytes[p]) return fail(h, SZ_ERR_INVALID, "slab: a partner expects a different message size (lists disagree)");
        unsigned char *base = nullptr;
        if (w.pid == (int32_t)getpid()) {
            base = (unsigned char *)(uintptr_t)w.base;
            if (w.device != h->cfg.device) {
                int can = 0;
                CK(cudaDeviceCanAccessPeer(&can, h->cfg.device, w.device));
                if (!can) return fail(h, SZ_ERR_UNSUPPORTED, "slab: no peer access between two devices of the decomposition");
                cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                cudaGetLastError();
            }
        } else {
            void *m = nullptr;
            for (size_t q = 0; q < B.ipc_open.size();) {
                if (B.ipc_open[q].pid != w.pid) { ++q; continue; }
                if (B.ipc_open[q].gen == w.gen) { m = B.ipc_open[q].ptr; ++q; continue; }
                cudaIpcCloseMemHandle(B.ipc_open[q].ptr);  // that process replaced its arena
                B.ipc_open.erase(B.ipc_open.begin() + q);
            }
            if (!m) {
                CK(cudaIpcOpenMemHandle(&m, w.ipc, cudaIpcMemLazyEnablePeerAccess));
                B.ipc_open.push_back({w.pid, w.gen, m});
            }
            base = (unsigned char *)m;
        }
        SlabPartnerDev &d = B.dev.p[p];
        d.r_stage[0] = (double *)(base + w.off_stage[0]);
        d.r_stage[1] = (double *)(base + w.off_stage[1]);
        d.r_ready = (int *)(base + w.off_ready);
        d.r_ack = (int *)(base + w.off_ack);
    }
    // first publication: the partners' first step consumes epoch 1
    szk_slab_push(h->L, h->S, B.dev, 1, B.max_send);
    CK(cudaStreamSynchronize(h->L.stream));
    CK(cudaGetLastError());
    szk_launch_count(true);
    B.pushed = 1;
    B.epoch = 1;
    return SZ_OK;
}

int32_t szb_step_begin(sz_handle *h) {
    if (!h || !h->slab.on) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    return step_enqueue(h);
}

int32_t szb_step_publish(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in, sz_floe_soa *out, int32_t host_mode) {
    (void)tstep;
    if (!h || !h->slab.on) return fail(h, SZ_ERR_INVALID, "slab step on a handle without halo lists");
    if (!h->have_domain || !h->have_floes) return fail(h, SZ_ERR_INVALID, "step before set_domain/upload_floes");
    if (do_coupling && !h->have_fields) return fail(h, SZ_ERR_INVALID, "step with coupling before set_fields");
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "step with ghosts present (call remove_ghosts)");
    cudaSetDevice(h->cfg.device);
    if (host_mode == 2) {  // sz_slab_step_host_partial: NULL fields of `in` keep the resident values, NULL fields of `out` stay behind
        if (!out) return SZ_ERR_INVALID;
        if (h->up_pending || h->cpl_prelaunched) return fail(h, SZ_ERR_INVALID, "slab_step_host_partial: an upload or a coupling of the split calls is pending");
        if (in && (in->n != h->n_total || in->n_init != h->n_init)) return fail(h, SZ_ERR_INVALID, "slab_step_host_partial: floe count differs from the resident store");
        if ((int)h->h_vcount.size() != h->n_init) return fail(h, SZ_ERR_INVALID, "slab_step_host_partial: no ring table of the resident floes");
        const int n = h->n_init;
        if (h->cf_cap < n) {
            dfree(h->d_cf_dn);
            CK(dalloc(&h->d_cf_dn, (size_t)n));
            h->cf_cap = n;
        }
        HostIO io = {in, out};
        h->next_partial = true;
        int32_t rc = step_setup(h, do_coupling, &io, true);
        h->next_partial = false;
        return rc;
    }
    if (host_mode) {
        int32_t rc = prepare_step_host(h, do_coupling, in, out);
        if (rc) return rc;
        HostIO io = {in, out};
        return step_setup(h, do_coupling, &io, true);
    }
    return step_setup(h, do_coupling, nullptr, false);
}

int32_t szb_step_end(sz_handle *h, int32_t host_mode) {
    if (!h) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    int32_t rc = step_finish(h);
    if (host_mode) {
        if (rc) {
            cudaStreamSynchronize(h->stream_up);
            cudaStreamSynchronize(h->stream_dn);
            return rc;
        }
        finish_host_tables(h, h->cur.io.out);
    }
    return rc;
}

int32_t szb_refresh_publish(sz_handle *h) {
    if (!h || !h->slab.on) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    if (h->slab.pushed < h->slab.epoch) {
        slab_push(h, h->slab.epoch);
        h->slab.pushed = h->slab.epoch;
    }
    CK(cudaGetLastError());
    return SZ_OK;
}

int32_t szb_refresh_consume(sz_handle *h) {
    if (!h || !h->slab.on) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    slab_unpack(h);  // the next step consumes the same epoch again: idempotent
    CK(cudaStreamSynchronize(h->L.stream));
    CK(cudaGetLastError());
    return SZ_OK;
}

double szb_max_displacement(sz_handle *h) {
    if (!h) return 0.0;
    double d;
    unsigned long long e = h->h_cnt->slab_disp;
    memcpy(&d, &e, sizeof(d));
    return d;
}

int32_t szb_upload_floes_resident_mc(sz_handle *h, const sz_floe_soa *s, const int64_t *mc_src, int64_t n_extra) {
    return upload_floes_impl(h, s, mc_src, n_extra);
}

int32_t szb_fetch_mc(sz_handle *h, int64_t off, int64_t n, double *x, double *y) {
    if (!h || n < 0 || off < 0 || off + n > h->n_mc) return SZ_ERR_INVALID;
    if (n == 0) return SZ_OK;
    cudaSetDevice(h->cfg.device);
    std::vector<double2> t((size_t)n);
    CK(cudaMemcpy(t.data(), h->S.mc + off, sizeof(double2) * (size_t)n, cudaMemcpyDeviceToHost));
    for (int64_t k = 0; k < n; ++k) { x[k] = t[k].x; y[k] = t[k].y; }
    return SZ_OK;
}

int32_t szb_host_alloc(size_t bytes, void **out) {
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? SZ_OK : SZ_ERR_NOMEM;
}
void szb_host_free(void *p) { if (p) cudaFreeHost(p); }

int32_t szb_mc_offsets(sz_handle *h, int64_t *off) {
    if (!h || !off) return SZ_ERR_INVALID;
    for (size_t i = 0; i < h->h_mc_off.size(); ++i) off[i] = h->h_mc_off[i];
    return SZ_OK;
}

// ---- results -----------------------------------------------------------------------------------------------
extern "C" int32_t sz_get_interactions(sz_handle *h, int64_t *offsets, double *rows) {
    if (!h || !offsets) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "get_interactions before upload_floes");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    int n = h->n_total;
    std::vector<int> off((size_t)n + 1);
    CK(cudaMemcpy(off.data(), h->B.row_off, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    for (int i = 0; i <= n; ++i) offsets[i] = off[i];
    if (rows && off[n] > 0) CK(cudaMemcpy(rows, h->B.rows, sizeof(double) * NCOL * (size_t)off[n], cudaMemcpyDeviceToHost));
    return SZ_OK;
}

extern "C" int32_t sz_set_interactions(sz_handle *h, const int64_t *offsets, const double *rows) {
    if (!h || !offsets) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "set_interactions before upload_floes");
    cudaSetDevice(h->cfg.device);
    int n = h->n_total;
    long long tot = offsets[n];
    if (tot > h->B.cap_rows) {
        int32_t rc = set_row_cap(h, (int)tot + 1024);
        if (rc) return rc;
    }
    std::vector<int> off((size_t)n + 1);
    for (int i = 0; i <= n; ++i) off[i] = (int)offsets[i];
    CK(cudaMemcpy(h->B.row_off, off.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    if (tot > 0) {
        if (!rows) return fail(h, SZ_ERR_INVALID, "set_interactions: rows missing");
        CK(cudaMemcpy(h->B.rows, rows, sizeof(double) * NCOL * (size_t)tot, cudaMemcpyHostToDevice));
    }
    return SZ_OK;
}

extern "C" int32_t sz_get_pairs(sz_handle *h, int32_t which, int64_t *pairs) {
    if (!h || !pairs) return SZ_ERR_INVALID;
    if (which < 0 || which > 3) return fail(h, SZ_ERR_INVALID, "get_pairs: which must be 0..3");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    int np = h->last.n_cand;
    if (np == 0) return SZ_OK;
    std::vector<int> pi(np), pj(np);
    std::vector<unsigned char> keep(np);
    std::vector<uint32_t> fl(np);
    CK(cudaMemcpy(pi.data(), h->B.pair_i, sizeof(int) * np, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pj.data(), h->B.pair_j, sizeof(int) * np, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(keep.data(), h->B.keep, np, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(fl.data(), h->B.item_flags, sizeof(uint32_t) * np, cudaMemcpyDeviceToHost));
    long long m = 0;
    for (int p = 0; p < np; ++p) {
        bool take = which == 0 || (which == 1 && keep[p]) || (which == 2 && keep[p] && (fl[p] & IT_OVERLAP)) ||
                    (which == 3 && keep[p] && (fl[p] & IT_FUSE));
        if (take) {
            pairs[2 * m] = pi[p] + 1;
            pairs[2 * m + 1] = pj[p] + 1;
            m++;
        }
    }
    return SZ_OK;
}

extern "C" int32_t sz_get_warnings(sz_handle *h, uint32_t *bits) {
    if (!h || !bits) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "get_warnings before upload_floes");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    if (h->n_init > 0) CK(cudaMemcpy(bits, h->S.warn, sizeof(uint32_t) * h->n_init, cudaMemcpyDeviceToHost));
    return SZ_OK;
}

extern "C" int32_t sz_get_timings(sz_handle *h, double ms[8]) {
    if (!h || !ms) return SZ_ERR_INVALID;
    memcpy(ms, h->ms, sizeof(double) * 8);
    return SZ_OK;
}

// ---- halo exchange ----------------------------------------------------------------------------------------
extern "C" int32_t sz_halo_configure(sz_handle *h, int32_t n_lists, const int64_t *off, const int64_t *idx) {
    if (!h || n_lists < 0 || (n_lists > 0 && (!off || !idx))) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "halo_configure before upload_floes");
    cudaSetDevice(h->cfg.device);
    CK(cudaStreamSynchronize(h->L.stream));
    long long tot = n_lists > 0 ? off[n_lists] : 0;
    std::vector<int> vcount(std::max(h->n_init, 1)), lidx((size_t)std::max<long long>(tot, 1));
    std::vector<long long> voff((size_t)std::max<long long>(tot, 1));
    if (h->n_init > 0) CK(cudaMemcpy(vcount.data(), h->S.vcount, sizeof(int) * h->n_init, cudaMemcpyDeviceToHost));
    h->hl_off.assign(off, off + n_lists + 1);
    h->hl_bytes.assign(n_lists, 0);
    for (int l = 0; l < n_lists; ++l) {
        long long v = 0;
        for (long long k = off[l]; k < off[l + 1]; ++k) {
            if (idx[k] < 1 || idx[k] > h->n_init) return fail(h, SZ_ERR_INVALID, "halo_configure: index out of range");
            lidx[k] = (int)(idx[k] - 1);
            voff[k] = v;
            v += vcount[lidx[k]];
        }
        h->hl_bytes[l] = 64 * (off[l + 1] - off[l]) + 16 * v;
    }
    dfree(h->d_hl_idx); dfree(h->d_hl_voff);
    CK(dalloc(&h->d_hl_idx, (size_t)tot)); CK(dalloc(&h->d_hl_voff, (size_t)tot));
    if (tot > 0) {
        CK(cudaMemcpy(h->d_hl_idx, lidx.data(), sizeof(int) * tot, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_hl_voff, voff.data(), sizeof(long long) * tot, cudaMemcpyHostToDevice));
    }
    h->n_lists = n_lists;
    return SZ_OK;
}

extern "C" int32_t sz_halo_bytes(sz_handle *h, int32_t list, int64_t *bytes) {
    if (!h || !bytes || list < 0 || list >= h->n_lists) return SZ_ERR_INVALID;
    *bytes = h->hl_bytes[list];
    return SZ_OK;
}

static int32_t halo_move(sz_handle *h, int32_t list, void *buf, int64_t bytes, bool pack, bool on_stream = false,
                         cudaStream_t user = nullptr) {
    if (!h || !buf || list < 0 || list >= h->n_lists) return SZ_ERR_INVALID;
    if (h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "halo exchange with ghosts present");
    if (pack ? bytes < h->hl_bytes[list] : bytes != h->hl_bytes[list]) return fail(h, pack ? SZ_ERR_CAPACITY : SZ_ERR_INVALID, "halo buffer size does not match the configured list");
    cudaSetDevice(h->cfg.device);
    long long a = h->hl_off[list], n = h->hl_off[list + 1] - a;
    if (on_stream) {
        // every earlier call on this handle has synchronised: the store is complete.  The kernel runs on the
        // caller's stream; after an unpack the handle's stream waits for it (device-side), so the next sz_step is
        // ordered behind the halo update without a host synchronisation
        Launch Lu = h->L;
        Lu.stream = user;
        if (h->up_pending) CK(cudaStreamWaitEvent(user, h->ev_up[3], 0));  // sz_upload_state_begin: the uploads land first
        szk_halo(Lu, h->S, h->d_hl_idx + a, h->d_hl_voff + a, (int)n, (double *)buf, pack);
        if (!pack) {
            CK(cudaEventRecord(h->ev_halo, user));
            CK(cudaStreamWaitEvent(h->L.stream, h->ev_halo, 0));
        }
        CK(cudaGetLastError());
        return SZ_OK;
    }
    szk_halo(h->L, h->S, h->d_hl_idx + a, h->d_hl_voff + a, (int)n, (double *)buf, pack);
    CK(cudaStreamSynchronize(h->L.stream));  // the caller's communication runs on its own stream
    CK(cudaGetLastError());
    return SZ_OK;
}
extern "C" int32_t sz_halo_pack_on(sz_handle *h, int32_t list, void *dst, int64_t cap, void *stream) {
    return halo_move(h, list, dst, cap, true, true, (cudaStream_t)stream);
}
extern "C" int32_t sz_halo_unpack_on(sz_handle *h, int32_t list, const void *src, int64_t bytes, void *stream) {
    return halo_move(h, list, (void *)src, bytes, false, true, (cudaStream_t)stream);
}
extern "C" int32_t sz_halo_pack(sz_handle *h, int32_t list, void *dst, int64_t cap) { return halo_move(h, list, dst, cap, true); }
extern "C" int32_t sz_halo_unpack(sz_handle *h, int32_t list, const void *src, int64_t bytes) { return halo_move(h, list, (void *)src, bytes, false); }

// ---- services for the host-side processes ------------------------------------------------
// grow-only device scratch of the service calls
template <typename T>
static cudaError_t ensure(T *&p, size_t &cap, size_t need) {
    if (need <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void **)&p, sizeof(T) * std::max<size_t>(need + need / 4, 64));
    if (e == cudaSuccess) cap = std::max<size_t>(need + need / 4, 64);
    return e;
}

extern "C" int32_t sz_pair_overlap_areas(sz_handle *h, int64_t n_pairs, const int64_t *pairs, double *areas, uint8_t *interacts) {
    if (!h || n_pairs < 0 || (n_pairs > 0 && (!pairs || !areas))) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "pair_overlap_areas before upload_floes");
    if (n_pairs > (1ll << 30)) return fail(h, SZ_ERR_UNSUPPORTED, "pair_overlap_areas: too many pairs");
    if (n_pairs == 0) return SZ_OK;
    cudaSetDevice(h->cfg.device);
    const int n = (int)n_pairs;
    std::vector<int2> hp((size_t)n);
    for (int k = 0; k < n; ++k) {
        int64_t i = pairs[2 * k], j = pairs[2 * k + 1];
        if (i < 1 || i > h->n_total || j < 1 || j > h->n_total) return fail(h, SZ_ERR_INVALID, "pair_overlap_areas: floe index out of range");
        hp[k] = make_int2((int)(i - 1), (int)(j - 1));
    }
    SvcBuf &V = h->svc;
    CK(ensure(V.pairs, V.cap_pairs, (size_t)n)); CK(ensure(V.area, V.cap_area, (size_t)n)); CK(ensure(V.inter, V.cap_inter, (size_t)n));
    CK(ensure(V.big, V.cap_big, (size_t)n + 1));
    cudaStream_t st = h->L.stream;
    CK(cudaMemcpyAsync(V.pairs, hp.data(), sizeof(int2) * n, cudaMemcpyHostToDevice, st));
    szk_pair_areas(h->L, h->S, V.pairs, n, V.area, V.inter, V.big + 1, V.big);
    CK(cudaMemcpyAsync(areas, V.area, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    if (interacts) CK(cudaMemcpyAsync(interacts, V.inter, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->h_cnt, h->S.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (h->h_cnt->error) {
        char buf[256];
        uint32_t e = h->h_cnt->error;
        szk_clear_error(h->L, h->S);
        return fail(h, SZ_ERR_UNSUPPORTED, errbits(e, buf, sizeof(buf)));
    }
    return SZ_OK;
}

extern "C" int32_t sz_eulerian_data(sz_handle *h, int32_t nx, int32_t ny, const double *xg, const double *yg, int32_t n_out,
                                    const int32_t *kinds, double *data) {
    if (!h || nx < 1 || ny < 1 || !xg || !yg || n_out < 0 || (n_out > 0 && (!kinds || !data))) return SZ_ERR_INVALID;
    if (!h->have_floes || !h->have_domain) return fail(h, SZ_ERR_INVALID, "eulerian_data before set_domain/upload_floes");
    if (n_out > SZ_GRID_NKINDS) return fail(h, SZ_ERR_INVALID, "eulerian_data: more outputs than kinds");
    for (int k = 0; k < n_out; ++k)
        if (kinds[k] < 0 || kinds[k] >= SZ_GRID_NKINDS) return fail(h, SZ_ERR_INVALID, "eulerian_data: unknown output kind");
    if ((long long)nx * ny > (1ll << 28)) return fail(h, SZ_ERR_UNSUPPORTED, "eulerian_data: too many cells");
    if (n_out == 0) return SZ_OK;
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->L.stream;
    SvcBuf &V = h->svc;
    const int nf = h->n_total, ncell = nx * ny;
    CK(ensure(V.xg, V.cap_xg, (size_t)nx + 1)); CK(ensure(V.yg, V.cap_yg, (size_t)ny + 1));
    CK(ensure(V.rec_count, V.cap_rc, (size_t)nf + 2)); CK(ensure(V.rec_off, V.cap_ro, (size_t)nf + 2));
    CK(ensure(V.cell_start, V.cap_cs, (size_t)ncell + 2)); CK(ensure(V.data, V.cap_data, (size_t)ncell * n_out));
    CK(ensure(V.big, V.cap_big, 2));
    if (h->n_topo > 0) { CK(ensure(V.cell_free, V.cap_cfree, (size_t)ncell)); CK(ensure(V.cell_topo, V.cap_ctopo, (size_t)ncell)); }
    CK(cudaMemcpyAsync(V.xg, xg, sizeof(double) * (nx + 1), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(V.yg, yg, sizeof(double) * (ny + 1), cudaMemcpyHostToDevice, st));
    const double dx = xg[1] - xg[0], dy = yg[1] - yg[0];  // output.jl:796-797
    int n_rec = 0;
    if (nf > 0) {
        szk_eul_count(h->L, h->S, nf, nx, ny, V.xg, V.yg, dx, dy, V.rec_count, V.rec_off);
        CK(cudaMemcpyAsync(&n_rec, V.rec_off + nf, sizeof(int), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
    }
    SzkEulArgs A;
    memset(&A, 0, sizeof(A));
    if (n_rec > 0) {
        size_t nr = (size_t)n_rec;
        CK(ensure(V.rec_floe, V.cap_rf, nr)); CK(ensure(V.rec_cell, V.cap_rcell, nr)); CK(ensure(V.rec_area, V.cap_ra, nr));
        CK(ensure(V.key_in, V.cap_ki, nr)); CK(ensure(V.key_out, V.cap_ko, nr)); CK(ensure(V.val_in, V.cap_vi, nr));
        CK(ensure(V.val_out, V.cap_vo, nr)); CK(ensure(V.big, V.cap_big, nr + 1));
        A.sort_bytes = szk_eul_sort_bytes(n_rec);
        CK(ensure(V.sort_tmp, V.cap_sort, A.sort_bytes));
    }
    A.n_floes = nf; A.nx = nx; A.ny = ny; A.n_rec = n_rec; A.n_out = n_out; A.d_xg = V.xg; A.d_yg = V.yg; A.dx = dx; A.dy = dy;
    A.rec_count = V.rec_count; A.rec_off = V.rec_off; A.rec_floe = V.rec_floe; A.rec_cell = V.rec_cell; A.val_in = V.val_in;
    A.val_out = V.val_out; A.cell_start = V.cell_start; A.big = V.big + 1; A.n_big = V.big; A.rec_area = V.rec_area;
    A.key_in = V.key_in; A.key_out = V.key_out; A.sort_tmp = V.sort_tmp; A.kinds = kinds; A.d_data = V.data;
    A.cell_free = V.cell_free; A.cell_topo = V.cell_topo; A.n_topo = h->n_topo;
    if (szk_eul_run(h->L, h->S, A) != 0) return fail(h, SZ_ERR_CUDA, "eulerian_data: sort failed");
    CK(cudaMemcpyAsync(data, V.data, sizeof(double) * (size_t)ncell * n_out, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->h_cnt, h->S.cnt, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    if (h->h_cnt->error) {
        char buf[256];
        uint32_t e = h->h_cnt->error;
        szk_clear_error(h->L, h->S);
        return fail(h, SZ_ERR_UNSUPPORTED, errbits(e, buf, sizeof(buf)));
    }
    return SZ_OK;
}

// generate_subfloe_points (coupling.jl:172-208, :235-321) for floes of the resident list
extern "C" int32_t sz_generate_subfloe_points(sz_handle *h, const sz_points_generator *g, int64_t n_floes, const int64_t *floes,
                                              int64_t *offsets, double *x, double *y, int64_t cap_points, int32_t *status,
                                              int32_t install) {
    if (!h || !g || !offsets || n_floes < 0) return SZ_ERR_INVALID;
    if (!h->have_floes) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points before upload_floes");
    if (g->kind != SZ_POINTS_MONTE_CARLO && g->kind != SZ_POINTS_SUB_GRID) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: unknown generator");
    if (g->kind == SZ_POINTS_MONTE_CARLO ? g->npoints < 1 : !(g->delta_g > 0)) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: bad generator parameters");
    if (!floes && n_floes != h->n_init) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: floes == NULL means all n_init floes");
    if (install && floes) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: install needs the whole list (floes == NULL)");
    if (install && h->n_total != h->n_init) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: install with ghosts present");
    if (n_floes > (1ll << 28)) return fail(h, SZ_ERR_UNSUPPORTED, "generate_subfloe_points: too many floes");
    cudaSetDevice(h->cfg.device);
    cudaStream_t st = h->L.stream;
    const int n = (int)n_floes;
    offsets[0] = 0;
    if (n == 0) return SZ_OK;
    int *d_floes = nullptr, *d_count = nullptr, *d_attempt = nullptr, *d_status = nullptr, *d_off = nullptr, *d_scan = nullptr;
    double2 *d_out = nullptr;
    double *d_x = nullptr, *d_y = nullptr;
    long long *d_off64 = nullptr;
    int32_t rc = SZ_OK;
    std::vector<int> hidx, hcount((size_t)n), hstatus((size_t)n);
    auto cleanup = [&]() {
        cudaFree(d_floes); cudaFree(d_count); cudaFree(d_attempt); cudaFree(d_status); cudaFree(d_off); cudaFree(d_scan);
        cudaFree(d_out); cudaFree(d_x); cudaFree(d_y); cudaFree(d_off64);
    };
#define PCK(expr)                                                                                                      \
    do {                                                                                                               \
        cudaError_t _e = (expr);                                                                                       \
        if (_e != cudaSuccess) {                                                                                       \
            snprintf(h->err, sizeof(h->err), "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            cleanup();                                                                                                 \
            return SZ_ERR_CUDA;                                                                                        \
        }                                                                                                              \
    } while (0)
    if (floes) {
        hidx.resize((size_t)n);
        for (int k = 0; k < n; ++k) {
            if (floes[k] < 1 || floes[k] > h->n_total) return fail(h, SZ_ERR_INVALID, "generate_subfloe_points: floe index out of range");
            hidx[k] = (int)(floes[k] - 1);
        }
        PCK(dalloc(&d_floes, (size_t)n));
        PCK(cudaMemcpyAsync(d_floes, hidx.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    }
    PCK(dalloc(&d_count, (size_t)n)); PCK(dalloc(&d_attempt, (size_t)n)); PCK(dalloc(&d_status, (size_t)n));
    szk_points_count(h->L, h->S, *g, d_floes, n, d_count, d_attempt, d_status);
    PCK(cudaMemcpyAsync(hcount.data(), d_count, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    PCK(cudaMemcpyAsync(hstatus.data(), d_status, sizeof(int) * n, cudaMemcpyDeviceToHost, st));
    PCK(cudaStreamSynchronize(st));
    PCK(cudaGetLastError());
    std::vector<int> hoff((size_t)n + 1, 0);
    for (int k = 0; k < n; ++k) {
        offsets[k + 1] = offsets[k] + hcount[k];
        if (offsets[k + 1] > (1ll << 30)) { cleanup(); return fail(h, SZ_ERR_UNSUPPORTED, "generate_subfloe_points: more than 2^30 points"); }
        hoff[k + 1] = (int)offsets[k + 1];
        if (status) status[k] = hstatus[k];
    }
    const long long M = offsets[n];
    const bool want_points = x && y;
    if (want_points && M > cap_points) { cleanup(); return fail(h, SZ_ERR_CAPACITY, "generate_subfloe_points: output arrays too small"); }
    if ((want_points || install) && M > 0) {
        PCK(dalloc(&d_off, (size_t)n + 1));
        PCK(cudaMemcpyAsync(d_off, hoff.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
        PCK(dalloc(&d_out, (size_t)M));
        szk_points_write(h->L, h->S, *g, d_floes, n, d_attempt, d_off, d_out);
        if (want_points) {
            PCK(dalloc(&d_x, (size_t)M)); PCK(dalloc(&d_y, (size_t)M));
            szk_deinterleave(h->L, d_out, d_x, d_y, M);
            PCK(cudaMemcpyAsync(x, d_x, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
            PCK(cudaMemcpyAsync(y, d_y, sizeof(double) * M, cudaMemcpyDeviceToHost, st));
        }
        PCK(cudaStreamSynchronize(st));
        PCK(cudaGetLastError());
    }
    if (install) {  // the generated points become the store's Monte-Carlo points (floe.jl:37-38), `remove` tags are applied
        Store &S = h->S;
        std::vector<long long> mo((size_t)n + 1);
        for (int k = 0; k <= n; ++k) mo[k] = offsets[k];
        cudaFree(S.mc);
        S.mc = d_out;
        d_out = nullptr;
        if (!S.mc) PCK(dalloc(&S.mc, 1));
        S.cap_mc = M;
        PCK(cudaMemcpy(S.mc_off, mo.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice));
        for (int k = 0; k < n; ++k) hstatus[k] = hstatus[k] == SZ_STATUS_REMOVE ? 1 : 0;
        PCK(cudaMemcpy(d_status, hstatus.data(), sizeof(int) * n, cudaMemcpyHostToDevice));
        szk_apply_remove_flags(h->L, S, d_status, n);
        szk_mc_radius(h->L, S);
        PCK(cudaStreamSynchronize(st));
        h->h_mc_off = mo;
        h->n_mc = M;
        h->gen++;
    }
#undef PCK
    cleanup();
    return rc;
}

extern "C" int32_t sz_clip_polygons(sz_handle *h, const double *p_xy, int32_t np, const double *q_xy, int32_t nq,
                                    int32_t cap_regions, int32_t cap_points, int32_t *out_offsets, double *out_xy,
                                    double *out_areas) {
    if (!h || !p_xy || !q_xy || !out_offsets || !out_xy || np < 1 || nq < 1) return SZ_ERR_INVALID;
    cudaSetDevice(h->cfg.device);
    return szk_debug_clip(h->L, p_xy, np, q_xy, nq, cap_regions, cap_points, out_offsets, out_xy, out_areas);
}
