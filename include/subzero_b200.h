/* subzero_b200.h — C ABI of the B200-native floe-interaction hot path.
 *
 * This is the drop-in boundary for the four calls inside the reference's
 * `timestep_sim!` (src/simulation_components/simulation.jl:94-220):
 *
 *   add_ghosts!(floes, domain)                      simulation.jl:102   -> sz_add_ghosts
 *   timestep_collisions!(floes, n_init, domain,     simulation.jl:109   -> sz_step_collisions
 *        consts, Δt, collision_settings, spinlock)
 *   (ghost deletion loop)                           simulation.jl:138   -> sz_remove_ghosts
 *   timestep_coupling!(model, Δt, consts,           simulation.jl:155   -> sz_step_coupling
 *        coupling_settings, floe_settings)
 *   timestep_floe_properties!(floes, tstep, Δt,     simulation.jl:165   -> sz_step_floe_properties
 *        floe_settings)
 *
 * The reference has no FFI of its own (it is 100 % Julia); a Julia host binds these
 * entry points with `ccall`.  Everything crossing the ABI is a
 * plain pointer, a size or a POD struct.  All pointers are caller-owned HOST memory,
 * read or written synchronously and never retained.  Every function returns an int32
 * status (0 = SZ_OK, negative = error; sz_last_error gives the text).  A handle is not
 * thread-safe: one host thread per handle.
 *
 * Indices: floe indices inside `interactions` rows and the pair getters are 1-BASED
 * (they are stored in the Float64 `floeidx` column exactly like the reference,
 * collisions.jl:297); domain elements are -1..-4 (N,S,E,W) and -(4+k) for topography
 * element k (collisions.jl:612-654).
 *
 * The same header is compiled with -DSZ_ORACLE_BUILD by oracle/ (test infrastructure);
 * the exported names then carry the prefix `szo_` instead of `sz_`.
 */
#ifndef SUBZERO_B200_H
#define SUBZERO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef SZ_ORACLE_BUILD
#define SZ_FN(name) szo_##name
#else
#define SZ_FN(name) sz_##name
#endif

#define SZ_OK 0
#define SZ_ERR_INVALID (-1)     /* bad argument / call order */
#define SZ_ERR_CUDA (-2)        /* CUDA runtime failure */
#define SZ_ERR_CAPACITY (-3)    /* a device buffer overflowed (pairs, regions, rows, ghosts) */
#define SZ_ERR_UNSUPPORTED (-4) /* outside the workspace (a ring of > 1024 points, a floe in > 2080 grid cells) */
#define SZ_ERR_NOMEM (-5)

/* Status tags, src/simulation_components/floe.jl:8-12 */
#define SZ_STATUS_ACTIVE 1
#define SZ_STATUS_REMOVE 2
#define SZ_STATUS_FUSE 3

/* Boundary kinds, src/simulation_components/domain_components/boundaries.jl:153,240,327,415 */
#define SZ_BOUNDARY_OPEN 0
#define SZ_BOUNDARY_PERIODIC 1
#define SZ_BOUNDARY_COLLISION 2
#define SZ_BOUNDARY_MOVING 3

/* Wall order everywhere in this ABI: 0 = North, 1 = South, 2 = East, 3 = West
 * (element ids -1, -2, -3, -4 of collisions.jl:612-642). */

/* Per-floe warning bits returned by sz_get_warnings; the host re-emits the reference's
 * @warn messages (update_floe.jl:483,488,528,541). */
#define SZ_WARN_HEIGHT_CAPPED 1u
#define SZ_WARN_FORCE_SCALED 2u
#define SZ_WARN_VELOCITY_LIMITED 4u
#define SZ_WARN_XI_CLAMPED 8u

typedef struct sz_handle sz_handle;

typedef struct sz_config {
    /* Constants, simulation.jl:5-18 */
    double rho_o, rho_a, Cd_io, Cd_ia, Cd_ao, f, turn_theta, L, k, nu, mu, E;
    /* CollisionSettings, process_settings.jl:183-229 */
    double floe_floe_max_overlap, floe_domain_max_overlap;
    /* FloeSettings subset, process_settings.jl:20-100; stress_calculators.jl:81-92 (λ) */
    double rho_i, max_floe_height, maximum_xi, stress_lambda;
    /* CouplingSettings, process_settings.jl:133-167 */
    int32_t coupling_dd;         /* Δd knot buffer; does not change a bilinear result */
    int32_t two_way_coupling_on; /* coupling.jl:1617-1680: ice/atmosphere stress on the ocean per grid cell */
    /* Simulation.Δt (Int seconds), simulation.jl:49-81 */
    int32_t dt;
    /* execution */
    int32_t device;               /* CUDA ordinal (ignored by the oracle build) */
    int32_t max_regions_per_pair; /* overlap regions kept per pair; 0 -> 4 */
    int32_t max_pairs_per_floe;   /* capacity hint for the candidate list; 0 -> 24 */
    int32_t threads;              /* oracle build only: OpenMP threads, 0 -> all */
    int32_t reserved0;
    int64_t floe_capacity;        /* 0 -> sized at upload (n + ghosts headroom) */
} sz_config;

/* Structure-of-arrays view of a floe list (a1: floe.jl:24-77).  Used for upload (read)
 * and download (written; caller allocates using sz_get_counts).  Any pointer may be NULL
 * on download to skip that field.  2x2 tensors are 4 doubles per floe in Julia's
 * column-major order (11, 21, 12, 22). */
typedef struct sz_floe_soa {
    int64_t n;      /* floes in the arrays: parents first, then ghosts (if any) */
    int64_t n_init; /* number of non-ghost floes */
    double *centroid_x, *centroid_y;
    double *height, *area, *mass, *rmax, *moment;
    double *alpha, *u, *v, *xi;
    double *fxOA, *fyOA, *trqOA, *hflx_factor, *overarea;
    double *collision_force; /* [n][2] */
    double *collision_trq;
    double *stress_accum, *stress_instant, *strain; /* [n][4] */
    double *p_dxdt, *p_dydt, *p_dudt, *p_dvdt, *p_dxidt, *p_dalphadt;
    int32_t *status_tag;
    int64_t *id, *ghost_id;
    int64_t *ghost_offsets; /* [n+1] CSR of each floe's `ghosts` list; NULL = no ghosts */
    int64_t *ghost_index;   /* 1-based indices into this floe list */
    int64_t *vert_offsets;  /* [n+1], in points; ring i = points [off[i], off[i+1]) closed */
    double *vert_xy;        /* [V][2] interleaved x,y */
    int64_t *mc_offsets;    /* [n+1]; x/y_subfloe_points (body frame, floe.jl:37-38) */
    double *mc_x, *mc_y;
} sz_floe_soa;

typedef struct sz_counts {
    int64_t n_init;        /* non-ghost floes */
    int64_t n_total;       /* floes incl. ghosts currently in the store */
    int64_t n_vertices;    /* ring points over n_total floes */
    int64_t n_mc;          /* Monte-Carlo points over n_init floes */
    int64_t n_ghost_links; /* total length of all `ghosts` lists */
    int64_t n_candidates;  /* floe pairs (i<j) passing the bounding-circle test */
    int64_t n_pairs;       /* candidates left after the id / ghost-image filter */
    int64_t n_overlap;     /* pairs with total overlap area > 0 */
    int64_t n_fuse;        /* (i,j) pairs tagged for fusion this step */
    int64_t n_rows;        /* interaction rows over n_total floes */
    int64_t n_domain_pairs;/* (floe, element) checks issued */
    int64_t n_clip_fail;   /* pairs whose region trace hit a degenerate configuration */
} sz_counts;

/* ---- lifetime ---------------------------------------------------------------------- */
void SZ_FN(default_config)(sz_config *cfg); /* reference defaults (Constants(), *Settings()) */
int32_t SZ_FN(create)(const sz_config *cfg, sz_handle **out);
void SZ_FN(destroy)(sz_handle *h);
const char *SZ_FN(last_error)(sz_handle *h);
const char *SZ_FN(version)(void);

/* ---- model description (a20) ---------------------------------------------------------- */
/* RegRectilinearGrid, grids.jl:106-116 */
int32_t SZ_FN(set_grid)(sz_handle *h, int32_t Nx, int32_t Ny, double x0, double xf, double y0,
                        double yf);
/* Ocean u, v, hflx_factor (oceans.jl:74-99) and Atmos u, v (atmos.jl:4-16): column-major
 * (Nx+1) x (Ny+1), element [ix + (Nx+1)*iy] == Julia A[ix+1, iy+1]. */
int32_t SZ_FN(set_fields)(sz_handle *h, const double *ocean_u, const double *ocean_v,
                          const double *ocean_hflx, const double *atmos_u,
                          const double *atmos_v);
/* Ocean.temp and Atmos.temp (oceans.jl:74-99, atmos.jl:4-16), same layout; only read by two-way coupling
 * (ocean.hflx_factor = dt k / (rho_i L) (T_ocean - T_atmos), coupling.jl:1676-1677).  NULL = zeros. */
int32_t SZ_FN(set_temperatures)(sz_handle *h, const double *ocean_temp, const double *atmos_temp);
/* Two-way coupling results (coupling.jl:1617-1680), each (Nx+1) x (Ny+1) like the inputs: ocean.tau_x,
 * ocean.tau_y, ocean.si_frac and the updated ocean.hflx_factor (which the next coupling step interpolates).
 * Any pointer may be NULL. */
int32_t SZ_FN(get_ocean_fields)(sz_handle *h, double *tau_x, double *tau_y, double *si_frac, double *hflx_factor);
/* The floe -> grid-cell registry of the last coupling step (grid.floe_locations + ocean.scells,
 * coupling.jl:1329-1454), sorted by (cell y, cell x, floe): cell_xy [n][2] 1-based cell indices, floe [n]
 * 1-based, vals [n][5] = sum of -tau_x_ocn, sum of -tau_y_ocn, npoints, dx, dy (the periodic shift that moves
 * the floe into the cell).  Call with NULL arrays to get *n. */
int32_t SZ_FN(get_cell_floes)(sz_handle *h, int64_t *n, int64_t *cell_xy, int64_t *floe, double *vals);
/* Domain, domains.jl:4-34.  rect[w] = {xmin, xmax, ymin, ymax} of wall w's rectangle
 * (boundaries.jl:29-33,65-69,102-106,139-143); uv[w] = MovingBoundary velocity.  Topography:
 * closed rings in CSR form plus centroid / rmax per element (topography.jl:5-9). */
int32_t SZ_FN(set_domain)(sz_handle *h, const int32_t kinds[4], const double vals[4],
                          const double uv[8], const double rect[16], int32_t n_topo,
                          const int64_t *topo_offsets, const double *topo_xy,
                          const double *topo_centroid, const double *topo_rmax);
/* Current wall positions (moving walls advance in sz_step_collisions, collisions.jl:565-571) */
int32_t SZ_FN(get_domain)(sz_handle *h, double vals[4], double rect[16]);

/* ---- floe state ---------------------------------------------------------------------------- */
int32_t SZ_FN(upload_floes)(sz_handle *h, const sz_floe_soa *floes);
/* Refresh the DYNAMIC state of the resident floes from the host: every per-floe scalar and the
 * ring coordinates of the same floe list (same n, n_init and ring sizes as the last
 * sz_upload_floes, no ghosts present).  Monte-Carlo points, ids and ghost links stay resident.
 * This is what the Julia shim calls after a host process changed floe state in place without
 * changing the floe list (simulation.jl:121-214). */
int32_t SZ_FN(upload_state)(sz_handle *h, const sz_floe_soa *floes);
int32_t SZ_FN(get_counts)(sz_handle *h, sz_counts *out);
int32_t SZ_FN(download_floes)(sz_handle *h, sz_floe_soa *floes);

/* ---- the hot path ------------------------------------------------------------------------- */
int32_t SZ_FN(add_ghosts)(sz_handle *h, int64_t *n_total);        /* collisions.jl:1060-1174 */
int32_t SZ_FN(step_collisions)(sz_handle *h);                     /* collisions.jl:734-864 */
int32_t SZ_FN(remove_ghosts)(sz_handle *h);                       /* simulation.jl:138-144 */
int32_t SZ_FN(step_coupling)(sz_handle *h);                       /* coupling.jl:1705-1738 */
int32_t SZ_FN(step_floe_properties)(sz_handle *h, int64_t tstep); /* update_floe.jl:469-551 */
/* One fused timestep without intermediate host synchronisation: add_ghosts, collisions,
 * remove_ghosts, coupling (iff do_coupling != 0), floe properties. */
int32_t SZ_FN(step)(sz_handle *h, int64_t tstep, int32_t do_coupling);
/* The same timestep on HOST buffers: sz_upload_state(in) + sz_step + sz_download_floes(out) as one call.  This
 * is what the Julia shim calls when host processes (fracture, ridging, welding, writers: simulation.jl:121-214)
 * touch the floe state every step.  The product overlaps the copies with the kernels: uploads are ordered by
 * first use (centroids and radii first, then rings, ...) and every kernel waits only for the arrays it reads;
 * results are copied back as soon as the kernel producing them is done.  `in` and `out` may alias (in-place
 * update of the host arrays); pinned host memory is needed for the overlap, pageable memory works but
 * serialises.  Monte-Carlo points and ghost lists are not transferred (mc_x / mc_y / ghost_index untouched), nor
 * are the fields the step overwrites before reading them: collision_force / collision_trq (collisions.jl:747-749)
 * and, when do_coupling != 0, fxOA / fyOA / trqOA / hflx_factor (coupling.jl:1583-1586). */
int32_t SZ_FN(step_host)(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in,
                         sz_floe_soa *out);
/* sz_step_host for a shim that knows which fields its host processes touched this step (simulation.jl:121-214: most steps
 * none of them runs): a NULL pointer in `in` KEEPS the device-resident value of that field (sz_step_host zero-fills it;
 * vert_xy NULL = the rings did not change on the host), a NULL pointer in `out` skips that download.  in == NULL: nothing
 * is uploaded at all.  Same results as sz_step_host for the fields that are exchanged. */
int32_t SZ_FN(step_host_partial)(sz_handle *h, int64_t tstep, int32_t do_coupling, const sz_floe_soa *in,
                                 sz_floe_soa *out);
/* The upload half of sz_step_host on its own, for a slab rank: sz_upload_state_begin(in) enqueues the uploads and
 * returns at once, the halo exchange follows (sz_halo_pack_on waits on the device for the uploads, so the halo
 * update lands on top of the uploaded copies), then sz_step_host(h, tstep, do_coupling, NULL, out) runs the step
 * and the overlapped downloads.  `in` must stay valid until that call returns; do_coupling must be the same. */
int32_t SZ_FN(upload_state_begin)(sz_handle *h, int32_t do_coupling, const sz_floe_soa *in);

/* ---- results ------------------------------------------------------------------------------- */
/* interactions: offsets[n_total+1] and rows[n_rows][7] =
 * (floeidx, xforce, yforce, xpoint, ypoint, torque, overlap), floe.jl:102-110.  Row order
 * per floe follows the reference: own pairs (j ascending, regions in clip order), walls
 * N,S,E,W, topography, mirrored rows, ghost rows.  Valid until the next step_collisions;
 * after remove_ghosts only the first n_init floes are reported. */
int32_t SZ_FN(get_interactions)(sz_handle *h, int64_t *offsets, double *rows);
/* Seed the rows from the host (offsets[n_total+1], rows[C][7]); only needed when
 * step_floe_properties is called without a preceding step_collisions on this handle
 * (calc_stress! reads rows 1:num_inters, update_floe.jl:392-414). */
int32_t SZ_FN(set_interactions)(sz_handle *h, const int64_t *offsets, const double *rows);
/* (i,j) 1-based, lexicographically sorted; which = 0 candidates, 1 filtered pairs,
 * 2 overlap pairs (total area > 0), 3 fuse pairs. `pairs` holds 2*count int64. */
int32_t SZ_FN(get_pairs)(sz_handle *h, int32_t which, int64_t *pairs);
int32_t SZ_FN(get_warnings)(sz_handle *h, uint32_t *bits); /* [n_init] */
/* Device (or oracle wall-clock) time of the phases of the last step, milliseconds:
 * [0] ghosts [1] broad phase [2] narrow phase [3] row assembly+reduction [4] coupling
 * [5] floe properties [6] total [7] kernels launched by the last sz_step */
int32_t SZ_FN(get_timings)(sz_handle *h, double ms[8]);

/* ---- slab decomposition: halo exchange of floe state (new, the reference is single-process) ----
 * One handle per GPU owns the floes of one spatial slab plus copies ("halo floes") of the neighbours'
 * floes that can touch them.  The floe LIST of a handle is fixed between rebuilds; every step the owner of a
 * floe sends its DYNAMIC state to the ranks that hold a copy.  A halo list is a set of local floe indices in
 * an order both sides agree on (ascending global index).  sz_halo_pack gathers, for the floes of one list,
 * 8 doubles per floe (centroid x, y, u, v, xi, height, status tag, alpha) followed by all their ring points
 * (x, y interleaved) into a caller-provided buffer in the memory space the handle computes in (a CUDA device
 * pointer for the product, host memory for the oracle build); the caller moves it (ncclSend/ncclRecv,
 * torch.distributed) and sz_halo_unpack scatters it into the floes of the matching list on the other side. */
int32_t SZ_FN(halo_configure)(sz_handle *h, int32_t n_lists, const int64_t *list_offsets,
                              const int64_t *floe_index /* 1-based local indices */);
int32_t SZ_FN(halo_bytes)(sz_handle *h, int32_t list, int64_t *bytes);
int32_t SZ_FN(halo_pack)(sz_handle *h, int32_t list, void *dst, int64_t capacity_bytes);
int32_t SZ_FN(halo_unpack)(sz_handle *h, int32_t list, const void *src, int64_t bytes);
/* The same two calls ordered on the CALLER's stream instead of returning after a host synchronisation:
 * `stream` is the cudaStream_t the caller's communication is enqueued on (ncclSend/ncclRecv, torch.distributed's
 * current stream).  pack -> send/recv -> unpack -> sz_step then run back to back on the device: sz_halo_unpack_on
 * makes the handle's own stream wait for the unpack kernel, nothing blocks the host.  The caller must not touch
 * the buffers from another stream in between.  The oracle build (host buffers) ignores `stream`. */
/* A slab rank may start the one-way coupling of the coming sz_step BEFORE the halo exchange: the coupling of an owned
 * floe reads only that floe's own state, so it runs beside pack / send / recv / unpack instead of after them
 * (the results of the halo copies are never used).  No-op (returns SZ_OK) where the order matters: with a periodic
 * boundary on this handle (add_ghosts! wraps parents first) or with two-way coupling.  The next sz_step /
 * sz_step_host(do_coupling != 0) joins it instead of launching its own. */
int32_t SZ_FN(coupling_begin)(sz_handle *h);
int32_t SZ_FN(halo_pack_on)(sz_handle *h, int32_t list, void *dst, int64_t capacity_bytes, void *stream);
int32_t SZ_FN(halo_unpack_on)(sz_handle *h, int32_t list, const void *src, int64_t bytes, void *stream);

/* ---- slab decomposition INSIDE the library ("multi-GPU is internal to one handle"; §8(e)) -------
 * A sz_slab drives the `n_local` of `world` slab ranks that live in this process:
 *   n_local == world  one host process (the Julia shim) drives every GPU of the box;
 *   n_local == 1      one process per GPU (torchrun / MPI): `alltoallv` moves the set-up / rebuild messages.
 * Each rank is a sz_handle on its own device holding its owned floes plus halo copies of the neighbours' floes,
 * sorted by GLOBAL floe index, so pair orientation, row order and the canonical image pair are those of one handle
 * and owned results are bit-identical to the single-GPU run (the reference is single-process: collisions.jl:734-864
 * sees one floe list).  The library does the partition (sz_slab_build), the per-step halo update, the staleness
 * test and the rebuild with migration of ownership (sz_slab_rebuild).
 *
 * Per-step data plane (CUDA build): no NCCL call and no pack / unpack round trip.  After the state update ONE
 * kernel per rank writes the 8 doubles + ring points of its boundary floes straight into the neighbours' receive
 * arenas over NVLink (peer-mapped memory in one process, cudaIpc handles between processes) and raises a flag
 * there; the neighbour's next step begins with a kernel that waits for the flag and scatters the records into its
 * store.  Arenas are double-buffered by step parity, so no rank waits for a neighbour's acknowledgement on the
 * critical path.  The oracle build moves the same records through host memory / `alltoallv`.
 *
 * alltoallv (MPI_Alltoallv semantics on bytes, collective over all `world` processes; only used when
 * n_local == 1): send block d = bytes [send_off[d], send_off[d+1]) of `send` goes to rank d, the block from rank s
 * arrives at [recv_off[s], recv_off[s+1]) of `recv`.  Return 0 on success. */
typedef struct sz_slab sz_slab;
typedef int32_t (*sz_alltoallv_fn)(void *ctx, const void *send, const int64_t *send_off, void *recv,
                                   const int64_t *recv_off);
#define SZ_SLAB_MAX_PARTNERS 16
int32_t SZ_FN(slab_create)(const sz_config *cfg, int32_t world, int32_t rank_first, int32_t n_local,
                           const int32_t *devices /* [n_local] CUDA ordinals */, double skin,
                           sz_alltoallv_fn alltoallv, void *ctx, sz_slab **out);
void SZ_FN(slab_destroy)(sz_slab *s);
const char *SZ_FN(slab_last_error)(sz_slab *s);
/* The handle of local rank k (0 <= k < n_local) for every per-handle call of this header (timings, counts,
 * downloads, interactions ...).  Owned by the slab. */
int32_t SZ_FN(slab_handle)(sz_slab *s, int32_t k, sz_handle **out);
/* Model description, replicated on every rank (same arguments as the per-handle calls). */
int32_t SZ_FN(slab_set_grid)(sz_slab *s, int32_t Nx, int32_t Ny, double x0, double xf, double y0, double yf);
int32_t SZ_FN(slab_set_fields)(sz_slab *s, const double *ocean_u, const double *ocean_v, const double *ocean_hflx,
                               const double *atmos_u, const double *atmos_v);
int32_t SZ_FN(slab_set_domain)(sz_slab *s, const int32_t kinds[4], const double vals[4], const double uv[8],
                               const double rect[16], int32_t n_topo, const int64_t *topo_offsets,
                               const double *topo_xy, const double *topo_centroid, const double *topo_rmax);
/* Slab boundaries in x, edges[world+1] ascending; slab r = [edges[r], edges[r+1]).  Without this call
 * sz_slab_build takes equal-count quantiles of the centroids it is given (n_local == world only). */
int32_t SZ_FN(slab_set_edges)(sz_slab *s, const double *edges);
/* Collective.  floes[k] / gidx[k]: the floes local rank k brings (full records incl. Monte-Carlo points) and their
 * 0-based global indices (unique over all ranks; ids must be unique too, NULL id = gidx + 1).  ANY initial
 * distribution works: a single-process host hands its whole list to local rank 0 and n = 0 to the others; every
 * floe migrates to the slab its centroid lies in, then the halo lists are built. */
int32_t SZ_FN(slab_build)(sz_slab *s, const sz_floe_soa *const *floes, const int64_t *const *gidx);
/* Local list of rank k after a build / rebuild: n floes (owned + halo, ascending global index). */
int32_t SZ_FN(slab_local_count)(sz_slab *s, int32_t k, int64_t *n, int64_t *n_owned);
int32_t SZ_FN(slab_local_index)(sz_slab *s, int32_t k, int64_t *gidx /* [n] */, int32_t *owner /* [n] */);
/* One timestep on every local rank (halo update + sz_step).  All ranks of the decomposition must make the same
 * sequence of sz_slab_step / sz_slab_step_host / sz_slab_rebuild calls. */
int32_t SZ_FN(slab_step)(sz_slab *s, int64_t tstep, int32_t do_coupling);
/* The same on host arrays of each rank's LOCAL list (layout of sz_slab_local_index): upload, halo update on top
 * of the uploaded (stale) halo copies, step, download.  in[k] == out[k] is allowed. */
int32_t SZ_FN(slab_step_host)(sz_slab *s, int64_t tstep, int32_t do_coupling, const sz_floe_soa *const *in,
                              sz_floe_soa *const *out);
/* sz_step_host_partial on every local rank: in == NULL or in[k] == NULL or a NULL field = the device-resident value stands,
 * a NULL field of out[k] is not downloaded.  The boundary floes are published AFTER the uploads, so a field the host
 * changed reaches the neighbours' halo copies in the same step. */
int32_t SZ_FN(slab_step_host_partial)(sz_slab *s, int64_t tstep, int32_t do_coupling, const sz_floe_soa *const *in,
                                      sz_floe_soa *const *out);
/* Largest distance an owned floe of the local ranks travelled since the lists were built (periodic wrap taken out);
 * the lists are valid while it stays below skin / 2.  No device synchronisation (the step reads it back). */
int32_t SZ_FN(slab_max_displacement)(sz_slab *s, double *metres);
/* Collective: migrate floes whose centroid left their slab (full record incl. Monte-Carlo points) and renew the
 * halo lists.  With n_local == world sz_slab_step calls it by itself when the displacement exceeds skin / 2. */
int32_t SZ_FN(slab_rebuild)(sz_slab *s);
/* Collective: bring every halo copy up to date with its owner's CURRENT state without stepping (between steps a copy
 * holds what its holder's own update made of it; only owned floes are results).  For host-side consumers of whole
 * local lists and for transport checks. */
int32_t SZ_FN(slab_refresh_halo)(sz_slab *s);
/* bytes pushed to the neighbours per step by local rank k, number of halo copies it holds, rebuilds so far */
int32_t SZ_FN(slab_stats)(sz_slab *s, int32_t k, int64_t *send_bytes, int64_t *halo_floes, int64_t *rebuilds);

/* ---- geometry service (test hook; also what SURVEY §8(f) rank 2 reuses) ---------------------- */
/* Clip two closed rings; regions are written as consecutive closed rings into out_xy
 * (capacity cap_points points), region r = points [out_offsets[r], out_offsets[r+1]).
 * Returns the number of regions (>= 0) or a negative status. */
int32_t SZ_FN(clip_polygons)(sz_handle *h, const double *p_xy, int32_t np, const double *q_xy,
                             int32_t nq, int32_t cap_regions, int32_t cap_points,
                             int32_t *out_offsets, double *out_xy, double *out_areas);

/* ---- services for the host-side processes ------------------------------------ */
/* Rank 2 — batched overlap query.  For every ordered pair (i, j) of `pairs` ([n_pairs][2], 1-based indices into
 * the resident floe list, ghosts included when present):
 *   interacts[k] = potential_interaction(centroid_i, centroid_j, rmax_i, rmax_j)          collisions.jl:705-710
 *   areas[k]     = sum(GO.area, intersect_polys(poly_i, poly_j); init = 0.0)  (0 when !interacts)
 * which is what smooth_floes! (simplification.jl:98-116), timestep_welding! (welding.jl:119-150) and the ridge /
 * raft validity test (ridge_raft.jl:706-753) evaluate pair by pair on the host.  The candidate list itself comes
 * from sz_get_pairs(h, 0, ...).  `interacts` may be NULL. */
int32_t SZ_FN(pair_overlap_areas)(sz_handle *h, int64_t n_pairs, const int64_t *pairs, double *areas,
                                  uint8_t *interacts);

/* Rank 3 — Eulerian gridded output, calc_eulerian_data! (output.jl:794-919): floe data averaged on the cells of
 * a GridOutputWriter grid.  xg [nx+1] / yg [ny+1] are the writer's grid lines; `kinds` [n_out] selects the
 * outputs (SZ_GRID_*, the symbols of output.jl:859-905); data is [nx][ny][n_out] in Julia's column-major order,
 * data[j + nx*(i + ny*k)] == writer.data[j+1, i+1, k+1] (x index first).  All floes in the store take part
 * (the reference calls it between add_ghosts! and timestep_collisions!, simulation.jl:102-105, so ghosts are
 * included when present).  Topography (cell_poly_list = cell minus the topography polygons, output.jl:826-829) is
 * evaluated from intersections only: area(floe ∩ (cell ∖ topo)) = area(floe ∩ cell) − Σ_k area((floe ∩ cell) ∩ topo_k),
 * area(cell ∖ topo) = area(cell) − Σ_k area(cell ∩ topo_k) — exact for topography elements that do not overlap each
 * other; a remainder below 1e-12 of the uncut area counts as zero (empty cell_poly_list: all outputs 0, :831-834). */
#define SZ_GRID_U 0
#define SZ_GRID_V 1
#define SZ_GRID_DUDT 2
#define SZ_GRID_DVDT 3
#define SZ_GRID_SI_FRAC 4
#define SZ_GRID_OVERAREA 5
#define SZ_GRID_MASS 6
#define SZ_GRID_AREA 7
#define SZ_GRID_HEIGHT 8
#define SZ_GRID_STRESS_XX 9
#define SZ_GRID_STRESS_YX 10
#define SZ_GRID_STRESS_XY 11
#define SZ_GRID_STRESS_YY 12
#define SZ_GRID_STRESS_EIG 13
#define SZ_GRID_STRAIN_UX 14
#define SZ_GRID_STRAIN_VX 15
#define SZ_GRID_STRAIN_UY 16
#define SZ_GRID_STRAIN_VY 17
#define SZ_GRID_NKINDS 18
int32_t SZ_FN(eulerian_data)(sz_handle *h, int32_t nx, int32_t ny, const double *xg, const double *yg,
                             int32_t n_out, const int32_t *kinds, double *data);

/* Rank 4 — sub-floe point generation (generate_subfloe_points, coupling.jl:172-208 Monte Carlo, :235-321 sub-grid),
 * what replace_floe! (update_floe.jl:55-66) calls for every floe a host process created.  For each listed floe of the
 * resident list the points are generated in the body frame (ring translated by -centroid, coupling.jl:189 / :261):
 *
 * SZ_POINTS_MONTE_CARLO  up to 10 attempts of `npoints` uniform draws in the ring's bounding box, kept when
 *     GO.coveredby(point, ring); an attempt is accepted when |kept / npoints * box area - area| / area <= err;
 *     after 10 failures the 10th attempt's points are kept and the floe is tagged `remove` (:182-185), as it is when no
 *     point was kept (:203-205).  Julia's Xoshiro stream cannot be reproduced, so the draws come from a counter-based
 *     generator — u(seed, floe id, attempt, draw, axis), splitmix64 finaliser, 53 bits — identical in this library and
 *     in the oracle: counts, `remove` semantics and point order are exact against the oracle, the DISTRIBUTION is what
 *     is held against the reference (statistical parity).
 * SZ_POINTS_SUB_GRID     deterministic: every vertex, edge points every <= delta_g (with the reference's shift of
 *     delta_g / 2 along the edge, including its always-positive x shift), then the interior lattice points kept when
 *     coveredby.  range(a, b, length = n) is restated as: endpoints exact, element i = a + i (b - a) / (n - 1)
 *     evaluated in double-double and rounded once (Julia's TwicePrecision ranges give the same value to <= 1 ulp).
 *
 * floes: 1-based indices into the resident list (NULL = all n_floes = n_init floes in order).  offsets[n_floes + 1] and
 * status[n_floes] (SZ_STATUS_ACTIVE / SZ_STATUS_REMOVE; may be NULL) are always written; x / y (capacity cap_points
 * points) may be NULL to query the sizes first.  install != 0 (all floes only): the generated points also replace the
 * resident Monte-Carlo points of the store, and `remove` tags are applied to status.tag. */
#define SZ_POINTS_MONTE_CARLO 0
#define SZ_POINTS_SUB_GRID 1
typedef struct sz_points_generator {
    int32_t kind;
    int32_t npoints;  /* Monte Carlo: draws per attempt (MonteCarloPointsGenerator.npoints, default 1000) */
    double err;       /* Monte Carlo: accepted relative area error (default 0.1) */
    double delta_g;   /* sub-grid: point spacing (SubGridPointsGenerator.Δg) */
    uint64_t seed;    /* Monte Carlo */
} sz_points_generator;
int32_t SZ_FN(generate_subfloe_points)(sz_handle *h, const sz_points_generator *gen, int64_t n_floes, const int64_t *floes,
                                       int64_t *offsets, double *x, double *y, int64_t cap_points, int32_t *status,
                                       int32_t install);

#ifdef __cplusplus
}
#endif
#endif /* SUBZERO_B200_H */
