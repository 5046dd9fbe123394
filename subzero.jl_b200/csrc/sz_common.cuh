// sz_common.cuh — device store layout and launch helpers shared by all kernels.
//
// Data layout in HBM: one SoA of per-floe scalars (a1: floe.jl:24-77),
// CSR-packed ring vertices (double2, closed rings), CSR-packed body-frame Monte-Carlo
// points (double2).  Ghost floes are appended behind the n_init parents exactly like the
// reference appends them to its StructArray (collisions.jl:881-1047).  All sizes that
// change inside a step (n_total, vertex count, pair counts) live in device memory
// (`Counters`) so a whole timestep is enqueued without host synchronisation; kernels are
// launched on grids sized from host-side capacities and stride over the device-side counts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/subzero_b200.h"

#define SZ_MAX_GHOSTS 3  // a parent has at most 3 periodic images (collisions.jl:891-897)

// Interaction-row columns, floe.jl:102-110
enum { COL_IDX = 0, COL_FX, COL_FY, COL_PX, COL_PY, COL_TRQ, COL_OV, NCOL = 7 };
// contact-pool row: fx, fy, px, py, overlap
#define NPOOL 5

// per work-item flags written by the narrow phase
enum : uint32_t {
    IT_OVERLAP = 1u,    // total overlap area > 0
    IT_FUSE = 2u,       // floe_floe_max_overlap exceeded (collisions.jl:366-368)
    IT_REMOVE = 4u,     // open wall hit / floe_domain_max_overlap exceeded (:438,:524)
    IT_CLIPFAIL = 8u,   // region trace met a degenerate configuration
    IT_NEEDLARGE = 16u, // exceeded the shared-memory budget of the small kernel
    IT_DONE = 32u,
};

// device-side error bits (Counters::error); the host maps them to SZ_ERR_CAPACITY and
// grows the named buffer before retrying the step
enum : uint32_t {
    ERR_PAIR_CAP = 1u,
    ERR_POOL_CAP = 2u,
    ERR_ROW_CAP = 4u,
    ERR_GHOST_CAP = 8u,
    ERR_VERT_CAP = 16u,
    ERR_FUSE_CAP = 32u,
    ERR_POLY_TOO_LARGE = 64u,
    ERR_DOM_CAP = 128u,
    ERR_GHOST_SLOTS = 256u,
    ERR_CREC_CAP = 512u,    // floe -> cell registry records
    ERR_CELL_TABLE = 1024u, // one floe touches more than 32 + 2048 grid cells
    ERR_SPILL_CAP = 4096u,     // spill table of the floe -> cell registry (floes touching more than 32 cells)
    ERR_SLAB_TIMEOUT = 2048u,  // a slab neighbour did not publish / acknowledge its halo records in time
};

struct Counters {
    int n_total;   // floes incl. ghosts
    int n_verts;   // ring points in use
    int n_cand;    // pairs (i<j) passing the circle test
    int n_dom;     // (floe, element) work items
    int n_rows;    // interaction rows over all floes
    int n_fuse;    // fuse pairs this step
    int n_pool;    // contact rows in the pool
    int n_clipfail;
    int n_kept;    // pairs after the image filter
    int n_overlap;
    int n_large;   // work items deferred to the large-polygon kernel
    int n_mid;     // work items the thread-per-item kernel handed to the warp-per-item kernel
    int n_crec;    // records of the floe -> cell registry (coupling)
    int n_spill;   // entries claimed in the registry's spill table
    int n_gflag;   // floes on the compact list of a ghost pass
    int n_cbig;    // ... whose ring needs the warp clip kernel
    int n_ccells;  // number of grid cells (scan length of the registry sort)
    int n_order;   // work items of the thread-per-item kernels (class-sorted)
    int n_force;   // ... of which need contact forces (phase 1)
    uint32_t error;
    int want_pairs, want_pool, want_rows, want_floes, want_verts, want_dom, want_fuse;  // sizes asked for on overflow
    int gnx, gny;
    int n_cells;
    int n_domchecks;  // (floe, element) checks incl. periodic walls (diagnostic)
    // broad-phase bounding box, order-preserving uint64 encodings (k_bbox)
    unsigned long long bb[5];  // min x, min y, max x, max y, max rmax
    double gx0, gy0, cell;
    // slab decomposition: largest distance an owned floe travelled since the halo lists were built (order-preserving
    // encoding of a double, atomicMax in k_slab_push); never reset by a step
    unsigned long long slab_disp;
};

struct DomainDev {
    int kind[4];
    double val[4], wu[4], wv[4];
    double rect[4][4];  // xmin, xmax, ymin, ymax
    int n_topo;
    int pad;
};

struct Params {
    sz_config cfg;
    int Nx, Ny;
    double x0, xf, y0, yf, dx, dy;
    int per_x, per_y;  // periodic east/west, north/south (domain kinds)
    int atm_nonzero, hflx_nonzero;  // false: the atmosphere / heat-flux fields are identically zero (coupling skips them)
};

// Everything a kernel needs, passed by value.
struct Store {
    int n_init, cap_floes, cap_verts;
    long long cap_mc;
    // per-floe scalars
    double *cx, *cy, *height, *area, *mass, *rmax, *moment, *alpha, *u, *v, *xi;
    double *fxOA, *fyOA, *trqOA, *hflx, *overarea, *cfx, *cfy, *ctrq;
    double *p_dxdt, *p_dydt, *p_dudt, *p_dvdt, *p_dxidt, *p_dalphadt;
    double *stress_accum, *stress_instant, *strain;  // 4 per floe
    int *status;
    long long *id, *ghost_id;
    int *parent;      // -1 for a parent floe, else index of its parent
    int *nghost;      // number of ghosts of a parent
    int *ghost_slot;  // [cap][SZ_MAX_GHOSTS]
    uint32_t *warn;
    unsigned char *cpl_remove;  // coupling found no in-bounds Monte-Carlo point (coupling.jl:1507)
    double *mc_r;               // largest |p| over the floe's sub-floe points (the sub-grid generator's shifted edge points may
                                // lie outside the ring): what the coupling's "strictly inside the grid" fast path must clear
    // CSR geometry
    int *vstart, *vcount;  // ring of floe f = verts[vstart[f] .. vstart[f]+vcount[f]), closed
    double2 *verts;
    long long *mc_off;  // [n_init+1]
    double2 *mc;        // body-frame Monte-Carlo points
    // topography
    int *topo_vstart, *topo_vcount;
    double2 *topo_verts;
    double *topo_cx, *topo_cy, *topo_rmax;
    // fields (Nx+1)x(Ny+1), [ix + (Nx+1) iy]
    double *ocn_u, *ocn_v, *ocn_hflx, *atm_u, *atm_v;
    double *fields8;  // the five fields interleaved per node (8 doubles, see sz_kernels_fp.cu)
    double *ocn_temp, *atm_temp, *taux, *tauy, *sifrac;  // two-way coupling: inputs and per-cell outputs
    Counters *cnt;
    DomainDev *dom;
};

// Work buffers of one collision step.
struct StepBuf {
    int cap_pairs, cap_dom, cap_rows, cap_fuse, cap_pool, cap_cells, cap_force;
    // uniform grid
    int *cell_of;     // [cap_floes]
    int *cell_count;  // [cap_cells+1]
    int *cell_start;  // [cap_cells+1]
    int *cell_fill;   // [cap_cells]
    int *cell_items;  // [cap_floes]
    double2 *cell_circ;  // [cap_floes][2] cell-sorted (cx, cy), (rmax, index)
    // neighbour lists
    int *up_count, *up_off;    // [cap_floes+1] pairs (i, j>i)
    int *low_count, *low_off;  // [cap_floes+1] pairs (i<j, j) seen from j
    int *pair_i, *pair_j;      // [cap_pairs] sorted (i, j)
    int *low_pair;             // [cap_pairs] per floe j: pair indices p of (i<j, j), i ascending
    unsigned char *keep;       // [cap_pairs] survives the image filter
    // domain work items
    int *dom_count, *dom_off;  // [cap_floes+1]
    int *dom_floe, *dom_elem;  // [cap_dom]
    // narrow-phase output, item w = pair p (w < cap_pairs) or cap_pairs + q
    int *item_nrows;     // contact rows written
    int *item_row0;      // first pool row
    uint32_t *item_flags;
    double *pool;        // [cap_pool][NPOOL] = fx, fy, px, py, overlap
    int *large_items;    // [cap_pairs + cap_dom] work list of the large-polygon kernel
    int *mid_items;      // [cap_pairs + cap_dom] work list of the warp-per-item kernel
    int *order;          // [cap_pairs + cap_dom] items sorted by (edges of P, edges of Q)
    unsigned char *order_cls;  // ... and their class (the ring sizes the warp lays its shared memory out for)
    int *force_items;    // [cap_force] items that need contact forces (cap_force == cap_pool)
    int4 *force_meta;    // [cap_force] region table of clip #1
    double2 *force_pts;  // [TN_PRE_PTS][cap_force] regions and crossing points of clip #1
    int *class_count, *class_base, *class_cursor;  // [128] counting sort of the work items
    // per-floe rows
    int *row_pre, *row_count, *row_off;  // [cap_floes+1]
    double *rows;                        // [cap_rows][7]
    int2 *fuse_pairs;                    // [cap_fuse]
    // ghost scratch
    int *g_flag, *g_cnt, *g_off, *g_vcnt, *g_voff;  // [cap_floes+1]
    int *g_list;                                    // [cap_floes+1] floes flagged by the current ghost pass (unordered)
    // scan scratch
    int *scan_block;
    // single-pass scans (look-back descriptors / tickets of the step's three scans) and the parked neighbour indices
    unsigned long long *lb_desc;
    int *lb_ticket;
    int lb_stride;
    int *nb_scratch;  // [NB_K][cap_floes]
};

// floe -> cell registry (grid.floe_locations / ocean.scells, coupling.jl:1329-1454) of one coupling step
struct CouplingBuf {
    int cap_crec, cap_cells, cap_spill;
    // per-floe tables of floes touching more than 32 cells (blocks of 2048 entries claimed on demand)
    int *sp_cell, *sp_n;
    int2 *sp_sd;
    double2 *sp_t;
    int *rec_cell, *rec_floe, *rec_npts;  // [cap_crec] unsorted records
    double2 *rec_t, *rec_d;               // sum of -tau_ocn, periodic shift (dx, dy)
    double *rec_area;                     // area of floe ∩ cell box
    int *cell_count, *cell_start, *cell_fill;  // [cells + 1] counting sort by cell
    int *perm;                            // [cap_crec] record indices sorted by (cell, floe)
    int *big_recs;                        // records whose ring needs the warp kernel
    int *scan_block;
};

// ---- slab decomposition: per-step halo update over peer memory (sz_slab_*) ---------------------------
// One entry per exchange partner.  "r_" pointers live in the PARTNER's receive arena (peer-mapped: same process, or
// a cudaIpc mapping), "l_" pointers in this rank's own arena.  A message is n records of 8 doubles (cx, cy, u, v, xi,
// height, status, alpha) followed by the ring points of those floes; arenas are double-buffered by epoch parity.
struct SlabPartnerDev {
    int send_off, send_n, recv_off, recv_n;  // segments of the send / receive lists
    double *r_stage[2];        // where my records for this partner go
    int *r_ready;              // partner's flag: my records of epoch e are complete
    int *r_ack;                // partner's flag: I have consumed its records of epoch e
    const double *l_stage[2];  // where the partner's records arrive
    int *l_ready, *l_ack;      // written by the partner
};
struct SlabDev {
    int n_partners;
    SlabPartnerDev p[SZ_SLAB_MAX_PARTNERS];
    const int *send_idx, *recv_idx;          // local floe indices, per partner segment
    const long long *send_voff, *recv_voff;  // first ring point of each listed floe inside its message
    int *push_count, *unpack_count;          // [n_partners] block counters (last block raises the flag)
    const unsigned char *owned;              // [n_init] 1 = this rank owns the floe
    const double *refx, *refy;               // centroids when the lists were built
    double period_x, period_y;               // 0 = not periodic
    unsigned long long timeout_ns;           // give up waiting for a neighbour's flag after this long (ERR_SLAB_TIMEOUT)
};

__host__ __device__ inline int sz_div_up(long long a, int b) { return (int)((a + b - 1) / b); }

struct Launch {
    cudaStream_t stream;
    int sms;  // multiprocessor count: persistent grids are sized in multiples of it
    int maxv_large, maxx_large;  // workspace of the large-polygon kernels
    int coupling_blocks_per_sm;  // 0 = fill the GPU; > 0 = persistent grid of that many blocks per SM
    bool capturing;              // the stream is being captured into a CUDA graph (sz_step)
    int pdl;                     // programmatic dependent launch in the collision chain (see sz_pdl below)
    bool chain_v2;               // fused broad-phase / row chain with single-pass scans (SZ_CHAIN_V1 selects the old one)
    bool no_phase_events;        // experiment (SZ_GRAPH_NO_EVENTS): no timing-event nodes inside a captured graph
};

// Timing events: inside a stream capture they must be recorded as EXTERNAL event nodes to stay usable with
// cudaEventElapsedTime after the graph has run.
inline void sz_record(const Launch &L, cudaEvent_t e, cudaStream_t s) {
    if (L.capturing && L.no_phase_events) return;
    cudaEventRecordWithFlags(e, s, L.capturing ? cudaEventRecordExternal : cudaEventRecordDefault);
}

// ---- host-callable launchers (sz_kernels.cu) -------------------------------------------------
// one periodic axis of add_ghosts! (0 = east/west, 1 = north/south)
void szk_ghost_pass(const Launch &L, const Store &S, const StepBuf &B, int axis, int n_floes_hint);
void szk_remove_ghosts(const Launch &L, const Store &S, int n_verts_init);
// ev (optional, 3 events): recorded after the broad phase, the narrow phase and the row assembly
// waits (optional, 2 events): the stream waits for [0] before the narrow phase and for [1] before the row assembly
void szk_collisions(const Launch &L, const Store &S, const StepBuf &B, const Params &P, int n_floes_hint,
                    int n_pairs_hint, cudaEvent_t *ev, const cudaEvent_t *waits = nullptr);
int szk_configure(const Launch &L);
void szk_halo(const Launch &L, const Store &S, const int *idx, const long long *voff, int n, double *buf, bool pack);
// slab data plane: push my boundary floes into the partners' arenas (+ displacement of the owned floes) / wait for
// the partners' records of `epoch` and scatter them into the store
void szk_slab_push(const Launch &L, const Store &S, const SlabDev &D, int epoch, int max_send);
void szk_slab_unpack(const Launch &L, const Store &S, const SlabDev &D, int epoch, int max_recv);
void szk_slab_reset_disp(const Launch &L, const Store &S);
// Monte-Carlo points of a re-built floe list: segment i comes from the old device array (src[i] >= 0: offset) or from
// `extra` (src[i] < 0: offset -1 - src[i])
void szk_mc_regather(const Launch &L, double2 *dst, const long long *dst_off, const double2 *old_mc, const double2 *extra,
                     const long long *src, int n);
void szk_pack_fields(const Launch &L, const Store &S, int n_nodes);
void szk_coupling(const Launch &L, const Store &S, const Params &P);
void szk_apply_coupling_tags(const Launch &L, const Store &S);
void szk_mc_radius(const Launch &L, const Store &S);  // S.mc_r from the resident points (after every change of them)
void szk_apply_remove_flags(const Launch &L, const Store &S, const int *flags, int n);  // status.tag = remove where flags[i] != 0
void szk_coupling_reg(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P);
void szk_cells_final(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P);
void szk_cells_sort_and_clip(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P, int n_rec_hint);
void szk_update(const Launch &L, const Store &S, const StepBuf &B, const Params &P);
void szk_interleave(const Launch &L, const double *x, const double *y, double2 *out, long long n);
void szk_deinterleave(const Launch &L, const double2 *in, double *x, double *y, long long n);
void szk_set_counts(const Launch &L, const Store &S, int n_total, int n_verts);
void szk_clear_error(const Launch &L, const Store &S);
int szk_debug_clip(const Launch &L, const double *p_xy, int np, const double *q_xy, int nq, int cap_regions,
                   int cap_points, int *out_offsets, double *out_xy, double *out_areas);
size_t szk_large_smem(int maxv, int maxx);
// ---- services (sz_services.cu) ------------------------------------------------------------------
int szk_services_configure(const Launch &L);
void szk_pair_areas(const Launch &L, const Store &S, const int2 *pairs, int n, double *area, unsigned char *inter, int *big,
                    int *n_big);
void szk_eul_count(const Launch &L, const Store &S, int n_floes, int nx, int ny, const double *d_xg, const double *d_yg, double dx,
                   double dy, int *rec_count, int *rec_off);
size_t szk_eul_sort_bytes(int n_rec);
struct SzkEulArgs {
    int n_floes, nx, ny, n_rec, n_out;
    const double *d_xg, *d_yg;
    double dx, dy;
    int *rec_count, *rec_off, *rec_floe, *rec_cell, *val_in, *val_out, *cell_start, *big, *n_big;
    double *rec_area;
    unsigned long long *key_in, *key_out;
    void *sort_tmp;
    size_t sort_bytes;
    const int *kinds;  // host
    double *d_data;
    double *cell_free;         // [ncell] topography: free area of every cell
    unsigned char *cell_topo;  // [ncell]
    int n_topo;
};

int szk_eul_run(const Launch &L, const Store &S, const SzkEulArgs &A);
// sub-floe point generation: count pass (points per floe, accepted Monte-Carlo attempt, status), write pass
void szk_points_count(const Launch &L, const Store &S, const sz_points_generator &g, const int *floes, int n, int *count, int *attempt,
                      int *status);
void szk_points_write(const Launch &L, const Store &S, const sz_points_generator &g, const int *floes, int n, int *attempt, const int *off,
                      double2 *out);
long long szk_launch_count(bool reset);
void szk_count_launches(int n);

// ---- programmatic dependent launch (PDL) for the chains of small dependent kernels --------------------------------
// Every kernel of the collision chain starts with sz_pdl(): it lets the NEXT kernel of the stream be launched at once
// (its blocks become resident and run up to their own sz_pdl() while this grid is still working) and then waits until
// the PREVIOUS grid has completed and its writes are visible.  Launched without the attribute both instructions are
// no-ops, so the same kernels serve the plain launches (CUDA-graph capture, the v1 chain, SZ_NO_PDL=1).
#ifdef __CUDACC__
__device__ __forceinline__ void sz_pdl() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline void sz_launch_pdl(bool on, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = on ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
