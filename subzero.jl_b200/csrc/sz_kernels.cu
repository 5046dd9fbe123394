// sz_kernels.cu — the sm_100a kernels of the floe-interaction hot path and their launchers.
//
//   K8  ghosts      add_ghosts!                 collisions.jl:881-1174
//   K1  broad phase potential_interaction       collisions.jl:705-710,745-763  (uniform grid)
//   K2  image filter collide_pairs Dict         collisions.jl:743,751-775
//   K3  narrow phase floe_floe_interaction!     collisions.jl:347-408 (warp per pair)
//   K4  domain       floe_domain_interaction!   collisions.jl:427-662
//   K5  rows         mirror/ghost rows, torque, totals  collisions.jl:799-862 (segmented, no float atomics)
//   K6  coupling     calc_one_way_coupling!     coupling.jl:1486-1589
//   K7  update       timestep_floe_properties!  update_floe.jl:392-551
//
// Compiled with -fmad=false (see sz_geom.cuh).  Every kernel strides over device-side counts
// (Counters) and returns at once when Counters::error is set, so an overflowing step leaves
// the floe state untouched and the host can grow the buffer and run the step again.
#include "sz_geom.cuh"
#include "sz_narrow_thread.cuh"

#define TPB 256

// kernels launched since the last szk_launch_count(true) (bench.py reports it as gpu_launches)
static long long g_launch_count = 0;
long long szk_launch_count(bool reset) {
    long long v = g_launch_count;
    if (reset) g_launch_count = 0;
    return v;
}
void szk_count_launches(int n) { g_launch_count += n; }

// ---- small helpers -----------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long enc_f64(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long e) {
    unsigned long long u = (e >> 63) ? (e ^ 0x8000000000000000ull) : ~e;
    return __longlong_as_double((long long)u);
}
__device__ __forceinline__ bool potential_interaction(double xi, double yi, double ri, double xj, double yj,
                                                      double rj) {
    double dx = xi - xj, dy = yi - yj, rr = ri + rj;  // collisions.jl:705-710
    return dx * dx + dy * dy < rr * rr;
}

// ---- exclusive scan (3 launches; lengths live on the device) --------------------------------------
#define SCAN_T 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_T * SCAN_ITEMS)

__device__ __forceinline__ int block_excl_scan_512(int v, int *total) {
    __shared__ int warp_sums[16];
    __shared__ int blk_total;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FULLMASK, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int s = lane < 16 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            int y = __shfl_up_sync(FULLMASK, s, o);
            if (lane >= o) s += y;
        }
        if (lane < 16) warp_sums[lane] = s;  // inclusive
        if (lane == 15) blk_total = s;
    }
    __syncthreads();
    int off = wid ? warp_sums[wid - 1] : 0;
    *total = blk_total;
    int r = off + x - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_tiles(const int *__restrict__ in, int *__restrict__ out,
                                                       const int *len_ptr, int len_add, int *block_sums,
                                                       const Counters *cnt) {
    if (cnt->error) return;
    int len = *len_ptr + len_add;
    int base = blockIdx.x * SCAN_TILE;
    if (base >= len) return;
    int v[SCAN_ITEMS], s = 0;
    int i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (i0 + k < len) ? in[i0 + k] : 0;
        s += v[k];
    }
    int tot;
    int pre = block_excl_scan_512(s, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < len) out[i0 + k] = pre;
        pre += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_sums(int *block_sums, const int *len_ptr, int len_add,
                                                      const Counters *cnt) {
    if (cnt->error) return;
    int len = *len_ptr + len_add;
    int ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    int carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_T) {
        int i = base + threadIdx.x;
        int v = i < ntiles ? block_sums[i] : 0;
        int tot;
        int pre = block_excl_scan_512(v, &tot);
        if (i < ntiles) block_sums[i] = carry + pre;
        carry += tot;
    }
    if (threadIdx.x == 0) block_sums[ntiles] = carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int *out, const int *len_ptr, int len_add,
                                                     const int *block_sums, int *total_out, const Counters *cnt) {
    if (cnt->error) return;
    int len = *len_ptr + len_add;
    int ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    int base = blockIdx.x * SCAN_TILE;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[len] = block_sums[ntiles];
        if (total_out) *total_out = block_sums[ntiles];
    }
    if (base >= len) return;
    int off = block_sums[blockIdx.x];
    int i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (i0 + k < len) out[i0 + k] += off;
}

// out[0..len] = exclusive prefix sums of in[0..len), len = *len_ptr + len_add <= max_len
static void scan_excl(const Launch &L, const Store &S, const int *in, int *out, const int *len_ptr, int len_add,
                      int max_len, int *scratch, int *total_out) {
    int tiles = sz_div_up((long long)max_len, SCAN_TILE);
    if (tiles < 1) tiles = 1;
    k_scan_tiles<<<tiles, SCAN_T, 0, L.stream>>>(in, out, len_ptr, len_add, scratch, S.cnt);
    k_scan_sums<<<1, SCAN_T, 0, L.stream>>>(scratch, len_ptr, len_add, S.cnt);
    k_scan_add<<<tiles, SCAN_T, 0, L.stream>>>(out, len_ptr, len_add, scratch, total_out, S.cnt);
    g_launch_count += 3;
}

// three scans of equal length in one set of launches (blockIdx.y selects the array): the neighbour counts of the
// broad phase (own pairs, mirrored pairs, domain items) — 3 launches instead of 9 on the critical chain of small kernels
struct Scan3 {
    const int *in[3];
    int *out[3];
    int *total_out[3];
};
__global__ void __launch_bounds__(SCAN_T) k_scan3_tiles(Scan3 a, const int *len_ptr, int len_add, int *block_sums, int stride,
                                                        const Counters *cnt) {
    if (cnt->error) return;
    const int y = blockIdx.y, len = *len_ptr + len_add, base = blockIdx.x * SCAN_TILE;
    if (base >= len) return;
    const int *in = a.in[y];
    int *out = a.out[y];
    int v[SCAN_ITEMS], s = 0;
    const int i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (i0 + k < len) ? in[i0 + k] : 0;
        s += v[k];
    }
    int tot;
    int pre = block_excl_scan_512(s, &tot);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < len) out[i0 + k] = pre;
        pre += v[k];
    }
    if (threadIdx.x == 0) block_sums[y * stride + blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_T) k_scan3_sums(int *block_sums, int stride, const int *len_ptr, int len_add, const Counters *cnt) {
    if (cnt->error) return;
    int *bs = block_sums + blockIdx.x * stride;
    const int len = *len_ptr + len_add, ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    int carry = 0;
    for (int base = 0; base < ntiles; base += SCAN_T) {
        int i = base + threadIdx.x;
        int v = i < ntiles ? bs[i] : 0;
        int tot;
        int pre = block_excl_scan_512(v, &tot);
        if (i < ntiles) bs[i] = carry + pre;
        carry += tot;
    }
    if (threadIdx.x == 0) bs[ntiles] = carry;
}
__global__ void __launch_bounds__(SCAN_T) k_scan3_add(Scan3 a, const int *len_ptr, int len_add, const int *block_sums, int stride,
                                                      const Counters *cnt) {
    if (cnt->error) return;
    const int y = blockIdx.y, len = *len_ptr + len_add, ntiles = (len + SCAN_TILE - 1) / SCAN_TILE, base = blockIdx.x * SCAN_TILE;
    const int *bs = block_sums + y * stride;
    int *out = a.out[y];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        out[len] = bs[ntiles];
        if (a.total_out[y]) *a.total_out[y] = bs[ntiles];
    }
    if (base >= len) return;
    const int off = bs[blockIdx.x], i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (i0 + k < len) out[i0 + k] += off;
}
static void scan_excl3(const Launch &L, const Store &S, const Scan3 &a, const int *len_ptr, int len_add, int max_len, int *scratch) {
    int tiles = sz_div_up((long long)max_len, SCAN_TILE);
    if (tiles < 1) tiles = 1;
    const int stride = tiles + 2;
    k_scan3_tiles<<<dim3(tiles, 3), SCAN_T, 0, L.stream>>>(a, len_ptr, len_add, scratch, stride, S.cnt);
    k_scan3_sums<<<3, SCAN_T, 0, L.stream>>>(scratch, stride, len_ptr, len_add, S.cnt);
    k_scan3_add<<<dim3(tiles, 3), SCAN_T, 0, L.stream>>>(a, len_ptr, len_add, scratch, stride, S.cnt);
    g_launch_count += 3;
}

// ---- single-pass exclusive scan (decoupled look-back): ONE launch instead of three ------------------------------------
// Tiles take a ticket (forward progress: a tile only ever waits for tiles that already run), publish
// (flag << 62 | value) in one 64-bit word — flag 1 = the tile's own sum, 2 = inclusive prefix up to and including the
// tile — and one thread walks back over its predecessors.  Up to three arrays of equal length per launch
// (blockIdx.y).  The descriptors and tickets of all scans of a step are zeroed by k_grid_zero at its start.
#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
#define LB_VAL ((1ull << 62) - 1)
struct ScanLB {
    const int *in[3];
    int *out[3];
    int *total_out[3];
    unsigned long long *desc;  // [3][stride]
    int *ticket;               // [3]
    int stride;
};
__global__ void __launch_bounds__(SCAN_T) k_scan_lb(ScanLB a, const int *len_ptr, int len_add, const Counters *cnt) {
    sz_pdl();
    if (cnt->error) return;
    __shared__ int s_tile, s_excl;
    const int y = blockIdx.y;
    if (threadIdx.x == 0) s_tile = atomicAdd(&a.ticket[y], 1);
    __syncthreads();
    const int tile = s_tile, len = *len_ptr + len_add, ntiles = (len + SCAN_TILE - 1) / SCAN_TILE;
    const int *in = a.in[y];
    int *out = a.out[y];
    if (len == 0) {
        if (tile == 0 && threadIdx.x == 0) {
            out[0] = 0;
            if (a.total_out[y]) *a.total_out[y] = 0;
        }
        return;
    }
    if (tile >= ntiles) return;
    unsigned long long *desc = a.desc + (size_t)y * a.stride;
    const int base = tile * SCAN_TILE, i0 = base + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (i0 + k < len) ? in[i0 + k] : 0;
        sum += v[k];
    }
    int tot;
    int pre = block_excl_scan_512(sum, &tot);
    if (threadIdx.x == 0) {
        int excl = 0;
        if (tile == 0) {
            atomicExch(&desc[0], LB_INC | (unsigned long long)tot);
        } else {
            atomicExch(&desc[tile], LB_AGG | (unsigned long long)tot);
            for (int t = tile - 1; t >= 0; --t) {
                unsigned long long d;
                do {
                    d = *(volatile unsigned long long *)&desc[t];
                } while ((d >> 62) == 0);
                excl += (int)(d & LB_VAL);
                if ((d >> 62) == 2) break;
            }
            atomicExch(&desc[tile], LB_INC | (unsigned long long)(excl + tot));
        }
        s_excl = excl;
        if (tile == ntiles - 1) {
            out[len] = excl + tot;
            if (a.total_out[y]) *a.total_out[y] = excl + tot;
        }
    }
    __syncthreads();
    pre += s_excl;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (i0 + k < len) out[i0 + k] = pre;
        pre += v[k];
    }
}
// slot: which of the step's scans (each has its own descriptors): 0 cells, 1 neighbour counts (3 arrays), 2 rows
#define LB_SLOTS 3       // zeroed by k_grid_zero
#define LB_SLOT_GHOST 3  // zeroed by k_ghost_flag2
static void scan_lb(const Launch &L, const Store &S, const StepBuf &B, int slot, int narr, const int *const *in, int *const *out,
                    int *const *total_out, const int *len_ptr, int len_add, int max_len) {
    ScanLB a;
    for (int k = 0; k < 3; ++k) {
        a.in[k] = k < narr ? in[k] : nullptr;
        a.out[k] = k < narr ? out[k] : nullptr;
        a.total_out[k] = k < narr ? total_out[k] : nullptr;
    }
    a.stride = B.lb_stride;
    a.desc = B.lb_desc + (size_t)slot * 3 * B.lb_stride;
    a.ticket = B.lb_ticket + slot * 3;
    int tiles = sz_div_up((long long)max_len, SCAN_TILE);
    if (tiles < 1) tiles = 1;
    sz_launch_pdl(L.chain_v2 && L.pdl != 0, k_scan_lb, dim3(tiles, narr), dim3(SCAN_T), 0, L.stream, a, len_ptr, len_add, (const Counters *)S.cnt);
    g_launch_count += 1;
}

static inline int grid_for(const Launch &L, long long work_items, int per_block) {
    long long b = (work_items + per_block - 1) / per_block;
    long long cap = (long long)L.sms * 32;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// ---- misc small kernels ----------------------------------------------------------------------------
__global__ void k_interleave(const double *__restrict__ x, const double *__restrict__ y, double2 *__restrict__ out,
                             long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = make_double2(x[i], y[i]);
}
__global__ void k_deinterleave(const double2 *__restrict__ in, double *__restrict__ x, double *__restrict__ y,
                               long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double2 p = in[i];
        x[i] = p.x;
        y[i] = p.y;
    }
}
void szk_interleave(const Launch &L, const double *x, const double *y, double2 *out, long long n) {
    if (n > 0) k_interleave<<<grid_for(L, n, TPB), TPB, 0, L.stream>>>(x, y, out, n);
}
void szk_deinterleave(const Launch &L, const double2 *in, double *x, double *y, long long n) {
    if (n > 0) k_deinterleave<<<grid_for(L, n, TPB), TPB, 0, L.stream>>>(in, x, y, n);
}

__global__ void k_set_counts(Counters *cnt, int n_total, int n_verts) {
    cnt->n_total = n_total;
    cnt->n_verts = n_verts;
    cnt->error = 0;
    cnt->bb[0] = cnt->bb[1] = ~0ull;  // the fused chain only accumulates into the box (k_bbox2) and resets it behind its last reader
    cnt->bb[2] = cnt->bb[3] = cnt->bb[4] = 0ull;
    cnt->n_gflag = 0;
}
void szk_set_counts(const Launch &L, const Store &S, int n_total, int n_verts) {
    k_set_counts<<<1, 1, 0, L.stream>>>(S.cnt, n_total, n_verts);
}
__global__ void k_clear_error(Counters *cnt) {
    cnt->error = 0;
    cnt->n_gflag = 0;  // a ghost pass that was abandoned half-way leaves its list behind
}
void szk_clear_error(const Launch &L, const Store &S) { k_clear_error<<<1, 1, 0, L.stream>>>(S.cnt); }

// ---- K8: ghosts (collisions.jl:881-1174) -------------------------------------------------------------
// One pass per periodic axis (E-W first, then N-S: collisions.jl:1171-1172).  Each active
// parent decides on its own (the reference's loop carries no dependency between parents other
// than the append position, which a prefix sum reproduces):
//   flag   c - r < min wall  ->  test the min wall, ghost translated by +L   (elseif: max wall, -L)
//   clip   intersect_polys(poly, wall.poly) non-empty                         (:889)
//   write  copies of the parent's existing ghosts, then of the parent (:891-895); ghost_id =
//          running number (:1034-1040); parent swapped with its last ghost when its centroid
//          lies outside the domain (:943-949)
struct GhostAxis {
    int axis, wmin, wmax;
};

__global__ void k_ghost_flag(Store S, StepBuf B, GhostAxis A) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const DomainDev *D = S.dom;
    int n0 = cnt->n_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) {
        int flag = 0;
        if (S.status[i] == SZ_STATUS_ACTIVE && S.ghost_id[i] == 0) {
            double c = A.axis == 0 ? S.cx[i] : S.cy[i], r = S.rmax[i];
            if (c - r < D->val[A.wmin]) flag = 1;
            else if (c + r > D->val[A.wmax]) flag = 2;
        }
        B.g_flag[i] = flag;
        B.g_cnt[i] = 0;
        B.g_vcnt[i] = 0;
    }
}

__device__ __forceinline__ void stage_wall_ring(double2 *Q, const DomainDev *D, int wall, int lane) {
    // _make_bounding_box_polygon, floe_utils.jl:104-108
    if (lane < 5) {
        double xmin = D->rect[wall][0], xmax = D->rect[wall][1], ymin = D->rect[wall][2], ymax = D->rect[wall][3];
        double2 p;
        switch (lane) {
        case 1: p = make_double2(xmin, ymax); break;
        case 2: p = make_double2(xmax, ymax); break;
        case 3: p = make_double2(xmax, ymin); break;
        default: p = make_double2(xmin, ymin); break;
        }
        Q[lane] = p;
    }
}

__global__ void k_ghost_clip(Store S, StepBuf B, GhostAxis A, int maxv, int maxx, int large) {
    extern __shared__ __align__(16) unsigned char smem[];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    int n0 = cnt->n_total;
    for (int i = blockIdx.x * wpb + wib; i < n0; i += gridDim.x * wpb) {
        int flag = B.g_flag[i];
        if (flag == 0) continue;
        if (large ? !(flag & 4) : (flag & 4)) continue;
        int side = flag & 3;
        int npp = S.vcount[i];
        bool defer = npp > w.maxv;
        int nreg = 0, status = CLIP_OK;
        if (!defer) {
            const double2 *gP = S.verts + S.vstart[i];
            for (int k = lane; k < npp; k += 32) w.P[k] = gP[k];
            stage_wall_ring(w.Q, S.dom, side == 1 ? A.wmin : A.wmax, lane);
            __syncwarp();
            nreg = warp_clip(w, w.P, npp, w.Q, 5, w.R1, w.rs1, w.re1, status);
            defer = status == CLIP_OVERFLOW;
        }
        if (lane == 0) {
            if (defer) {
                if (large) atomicOr(&cnt->error, ERR_POLY_TOO_LARGE);
                else B.g_flag[i] = side | 4;
            } else if (nreg > 0) {
                int ng = S.nghost[i], vc = npp;
                for (int k = 0; k < ng; ++k) vc += S.vcount[S.ghost_slot[i * SZ_MAX_GHOSTS + k]];
                B.g_flag[i] = side;
                B.g_cnt[i] = ng + 1;
                B.g_vcnt[i] = vc;
                if (2 * ng + 1 > SZ_MAX_GHOSTS) atomicOr(&cnt->error, ERR_GHOST_SLOTS);
            } else {
                B.g_flag[i] = 0;
            }
        }
        __syncwarp();
    }
}

__global__ void k_ghost_check(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n0 = cnt->n_total;
    int add = B.g_off[n0], vadd = B.g_voff[n0];
    if (n0 + add > S.cap_floes) {
        cnt->error |= ERR_GHOST_CAP;
        cnt->want_floes = n0 + add;
    }
    if (cnt->n_verts + vadd > S.cap_verts) {
        cnt->error |= ERR_VERT_CAP;
        cnt->want_verts = cnt->n_verts + vadd;
    }
}

// deepcopy_floe, floe_utils.jl:120-161 (Monte-Carlo points are not copied: coupling runs after
// the ghosts are deleted, simulation.jl:138-161)
__device__ void copy_floe_scalars(const Store &S, int src, int dst) {
#define CP(f) S.f[dst] = S.f[src];
    CP(cx) CP(cy) CP(height) CP(area) CP(mass) CP(rmax) CP(moment) CP(alpha) CP(u) CP(v) CP(xi) CP(fxOA) CP(fyOA)
    CP(trqOA) CP(hflx) CP(overarea) CP(cfx) CP(cfy) CP(ctrq) CP(p_dxdt) CP(p_dydt) CP(p_dudt) CP(p_dvdt)
    CP(p_dxidt) CP(p_dalphadt) CP(status) CP(id)
#undef CP
    for (int k = 0; k < 4; ++k) {
        S.stress_accum[4 * dst + k] = S.stress_accum[4 * src + k];
        S.stress_instant[4 * dst + k] = S.stress_instant[4 * src + k];
        S.strain[4 * dst + k] = S.strain[4 * src + k];
    }
    S.warn[dst] = 0;
    S.nghost[dst] = 0;
}

__global__ void k_ghost_write(Store S, StepBuf B, GhostAxis A) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const DomainDev *D = S.dom;
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    int n0 = cnt->n_total, v0 = cnt->n_verts;
    double Lp = D->val[A.wmax] - D->val[A.wmin];
    for (int i = blockIdx.x * wpb + wib; i < n0; i += gridDim.x * wpb) {
        int cntg = B.g_cnt[i];
        if (cntg == 0) continue;
        int side = B.g_flag[i] & 3;
        double t = side == 1 ? Lp : -Lp;
        double tx = A.axis == 0 ? t : 0.0, ty = A.axis == 0 ? 0.0 : t;
        int base = n0 + B.g_off[i], vbase = v0 + B.g_voff[i];
        int ng = S.nghost[i];
        for (int k = 0; k < cntg; ++k) {
            int src = k < ng ? S.ghost_slot[i * SZ_MAX_GHOSTS + k] : i, dst = base + k;
            int nv = S.vcount[src], vs = S.vstart[src];
            if (lane == 0) {
                copy_floe_scalars(S, src, dst);
                S.cx[dst] += tx;  // _translate_floe!, floe_utils.jl:66-72
                S.cy[dst] += ty;
                S.ghost_id[dst] = (long long)((k + 1) + ng);  // collisions.jl:1036-1038
                S.parent[dst] = i;
                S.vstart[dst] = vbase;
                S.vcount[dst] = nv;
            }
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vs + q];
                p.x += tx;
                p.y += ty;
                S.verts[vbase + q] = p;
            }
            vbase += nv;
        }
        __syncwarp();
        // parent / last-ghost swap, collisions.jl:943-949
        double c = A.axis == 0 ? S.cx[i] : S.cy[i];
        double sw = 0.0;
        if (c < D->val[A.wmin]) sw = Lp;
        else if (D->val[A.wmax] < c) sw = -Lp;
        __syncwarp();
        if (sw != 0.0) {
            double sx = A.axis == 0 ? sw : 0.0, sy = A.axis == 0 ? 0.0 : sw;
            int last = base + cntg - 1;
            int vs = S.vstart[i], nv = S.vcount[i];
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vs + q];
                p.x += sx;
                p.y += sy;
                S.verts[vs + q] = p;
            }
            int vl = v0 + B.g_voff[i] + B.g_vcnt[i] - nv;  // the parent's copy is the last ring written
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vl + q];
                p.x += -sx;
                p.y += -sy;
                S.verts[vl + q] = p;
            }
            if (lane == 0) {
                S.cx[i] += sx;
                S.cy[i] += sy;
                S.cx[last] += -sx;
                S.cy[last] += -sy;
            }
        }
        if (lane == 0) {
            for (int k = 0; k < cntg; ++k) S.ghost_slot[i * SZ_MAX_GHOSTS + ng + k] = base + k;
            S.nghost[i] = ng + cntg;
        }
        __syncwarp();
    }
}

__global__ void k_ghost_commit(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n0 = cnt->n_total;
    cnt->n_verts += B.g_voff[n0];
    cnt->n_total = n0 + B.g_off[n0];
}

// ---- the ghost pass on a COMPACT list of the flagged floes (v2) -------------------------------------------------------
// Only the floes near a periodic wall take part (a few thousand of 250 k): the flag kernel appends them to a list
// (order irrelevant: every result is written at the floe's own index, the append positions come from prefix sums over
// the floe index as before), the clip and write kernels walk that list instead of striding over every floe, and the two
// prefix sums are one look-back launch.  6 launches per axis instead of 12.
__global__ void k_ghost_flag2(Store S, StepBuf B, GhostAxis A) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const DomainDev *D = S.dom;
    const int n0 = cnt->n_total;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int k = t0; k < 3 * B.lb_stride; k += nt) B.lb_desc[(size_t)LB_SLOT_GHOST * 3 * B.lb_stride + k] = 0ull;
    if (t0 < 3) B.lb_ticket[LB_SLOT_GHOST * 3 + t0] = 0;
    for (int i = t0; i < n0; i += nt) {
        int flag = 0;
        if (S.status[i] == SZ_STATUS_ACTIVE && S.ghost_id[i] == 0) {
            double c = A.axis == 0 ? S.cx[i] : S.cy[i], r = S.rmax[i];
            if (c - r < D->val[A.wmin]) flag = 1;
            else if (c + r > D->val[A.wmax]) flag = 2;
        }
        B.g_flag[i] = flag;
        B.g_cnt[i] = 0;
        B.g_vcnt[i] = 0;
        if (flag) B.g_list[atomicAdd(&cnt->n_gflag, 1)] = i;
    }
}

__global__ void k_ghost_clip2(Store S, StepBuf B, GhostAxis A, int maxv, int maxx, int large) {
    extern __shared__ __align__(16) unsigned char smem[];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    const int nl = cnt->n_gflag;
    for (int e = blockIdx.x * wpb + wib; e < nl; e += gridDim.x * wpb) {
        const int i = B.g_list[e];
        int flag = B.g_flag[i];
        if (flag == 0) continue;
        if (large ? !(flag & 4) : (flag & 4)) continue;
        int side = flag & 3;
        int npp = S.vcount[i];
        bool defer = npp > w.maxv;
        int nreg = 0, status = CLIP_OK;
        if (!defer) {
            const double2 *gP = S.verts + S.vstart[i];
            for (int k = lane; k < npp; k += 32) w.P[k] = gP[k];
            stage_wall_ring(w.Q, S.dom, side == 1 ? A.wmin : A.wmax, lane);
            __syncwarp();
            nreg = warp_clip(w, w.P, npp, w.Q, 5, w.R1, w.rs1, w.re1, status);
            defer = status == CLIP_OVERFLOW;
        }
        if (lane == 0) {
            if (defer) {
                if (large) atomicOr(&cnt->error, ERR_POLY_TOO_LARGE);
                else B.g_flag[i] = side | 4;
            } else if (nreg > 0) {
                int ng = S.nghost[i], vc = npp;
                for (int k = 0; k < ng; ++k) vc += S.vcount[S.ghost_slot[i * SZ_MAX_GHOSTS + k]];
                B.g_flag[i] = side;
                B.g_cnt[i] = ng + 1;
                B.g_vcnt[i] = vc;
                if (2 * ng + 1 > SZ_MAX_GHOSTS) atomicOr(&cnt->error, ERR_GHOST_SLOTS);
            } else {
                B.g_flag[i] = 0;
            }
        }
        __syncwarp();
    }
}

// k_ghost_check + k_ghost_write over the list
__global__ void k_ghost_write2(Store S, StepBuf B, GhostAxis A) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const DomainDev *D = S.dom;
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int n0 = cnt->n_total, v0 = cnt->n_verts;
    {
        const int add = B.g_off[n0], vadd = B.g_voff[n0];
        if (n0 + add > S.cap_floes || v0 + vadd > S.cap_verts) {  // every thread backs off, one reports
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                if (n0 + add > S.cap_floes) { atomicOr(&cnt->error, ERR_GHOST_CAP); cnt->want_floes = n0 + add; }
                if (v0 + vadd > S.cap_verts) { atomicOr(&cnt->error, ERR_VERT_CAP); cnt->want_verts = v0 + vadd; }
            }
            return;
        }
    }
    const double Lp = D->val[A.wmax] - D->val[A.wmin];
    const int nl = cnt->n_gflag;
    for (int e = blockIdx.x * wpb + wib; e < nl; e += gridDim.x * wpb) {
        const int i = B.g_list[e];
        int cntg = B.g_cnt[i];
        if (cntg == 0) continue;
        int side = B.g_flag[i] & 3;
        double t = side == 1 ? Lp : -Lp;
        double tx = A.axis == 0 ? t : 0.0, ty = A.axis == 0 ? 0.0 : t;
        int base = n0 + B.g_off[i], vbase = v0 + B.g_voff[i];
        int ng = S.nghost[i];
        for (int k = 0; k < cntg; ++k) {
            int src = k < ng ? S.ghost_slot[i * SZ_MAX_GHOSTS + k] : i, dst = base + k;
            int nv = S.vcount[src], vs = S.vstart[src];
            if (lane == 0) {
                copy_floe_scalars(S, src, dst);
                S.cx[dst] += tx;  // _translate_floe!, floe_utils.jl:66-72
                S.cy[dst] += ty;
                S.ghost_id[dst] = (long long)((k + 1) + ng);  // collisions.jl:1036-1038
                S.parent[dst] = i;
                S.vstart[dst] = vbase;
                S.vcount[dst] = nv;
            }
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vs + q];
                p.x += tx;
                p.y += ty;
                S.verts[vbase + q] = p;
            }
            vbase += nv;
        }
        __syncwarp();
        // parent / last-ghost swap, collisions.jl:943-949
        double c = A.axis == 0 ? S.cx[i] : S.cy[i];
        double sw = 0.0;
        if (c < D->val[A.wmin]) sw = Lp;
        else if (D->val[A.wmax] < c) sw = -Lp;
        __syncwarp();
        if (sw != 0.0) {
            double sx = A.axis == 0 ? sw : 0.0, sy = A.axis == 0 ? 0.0 : sw;
            int last = base + cntg - 1;
            int vs = S.vstart[i], nv = S.vcount[i];
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vs + q];
                p.x += sx;
                p.y += sy;
                S.verts[vs + q] = p;
            }
            int vl = v0 + B.g_voff[i] + B.g_vcnt[i] - nv;  // the parent's copy is the last ring written
            for (int q = lane; q < nv; q += 32) {
                double2 p = S.verts[vl + q];
                p.x += -sx;
                p.y += -sy;
                S.verts[vl + q] = p;
            }
            if (lane == 0) {
                S.cx[i] += sx;
                S.cy[i] += sy;
                S.cx[last] += -sx;
                S.cy[last] += -sy;
            }
        }
        if (lane == 0) {
            for (int k = 0; k < cntg; ++k) S.ghost_slot[i * SZ_MAX_GHOSTS + ng + k] = base + k;
            S.nghost[i] = ng + cntg;
        }
        __syncwarp();
    }
}

__global__ void k_ghost_commit2(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    cnt->n_gflag = 0;  // for the next pass, error or not
    if (cnt->error) return;
    int n0 = cnt->n_total;
    cnt->n_verts += B.g_voff[n0];
    cnt->n_total = n0 + B.g_off[n0];
}

static void ghost_pass_v2(const Launch &L, const Store &S, const StepBuf &B, int axis, int n_hint) {
    GhostAxis A;
    A.axis = axis;
    A.wmax = axis == 0 ? 2 : 0;
    A.wmin = axis == 0 ? 3 : 1;
    const int maxv_s = 32, maxx_s = 16, wpb = 4;
    const int glist = grid_for(L, n_hint / 16 + 256, wpb);  // the list holds a small fraction of the floes; the loops stride anyway
    k_ghost_flag2<<<grid_for(L, n_hint, TPB), TPB, 0, L.stream>>>(S, B, A);
    k_ghost_clip2<<<glist, wpb * 32, wpb * ws_bytes(maxv_s, maxx_s), L.stream>>>(S, B, A, maxv_s, maxx_s, 0);
    k_ghost_clip2<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), L.stream>>>(S, B, A, L.maxv_large, L.maxx_large, 1);
    {
        const int *in[2] = {B.g_cnt, B.g_vcnt};
        int *out[2] = {B.g_off, B.g_voff}, *tot[2] = {nullptr, nullptr};
        scan_lb(L, S, B, LB_SLOT_GHOST, 2, in, out, tot, &S.cnt->n_total, 0, S.cap_floes);
    }
    k_ghost_write2<<<grid_for(L, n_hint / 16 + 256, 8), 256, 0, L.stream>>>(S, B, A);
    k_ghost_commit2<<<1, 1, 0, L.stream>>>(S, B);
    g_launch_count += 5;
}

static void ghost_pass(const Launch &L, const Store &S, const StepBuf &B, int axis, int n_hint) {
    GhostAxis A;
    A.axis = axis;
    A.wmax = axis == 0 ? 2 : 0;
    A.wmin = axis == 0 ? 3 : 1;
    const int maxv_s = 32, maxx_s = 16, wpb = 4;
    k_ghost_flag<<<grid_for(L, n_hint, TPB), TPB, 0, L.stream>>>(S, B, A);
    k_ghost_clip<<<grid_for(L, n_hint, wpb), wpb * 32, wpb * ws_bytes(maxv_s, maxx_s), L.stream>>>(S, B, A, maxv_s,
                                                                                                   maxx_s, 0);
    k_ghost_clip<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), L.stream>>>(S, B, A, L.maxv_large, L.maxx_large,
                                                                              1);
    scan_excl(L, S, B.g_cnt, B.g_off, &S.cnt->n_total, 0, S.cap_floes, B.scan_block, nullptr);
    scan_excl(L, S, B.g_vcnt, B.g_voff, &S.cnt->n_total, 0, S.cap_floes, B.scan_block, nullptr);
    k_ghost_check<<<1, 1, 0, L.stream>>>(S, B);
    k_ghost_write<<<grid_for(L, n_hint, 8), 256, 0, L.stream>>>(S, B, A);
    k_ghost_commit<<<1, 1, 0, L.stream>>>(S, B);
    g_launch_count += 6;
}

void szk_ghost_pass(const Launch &L, const Store &S, const StepBuf &B, int axis, int n_hint) {
    if (L.chain_v2) ghost_pass_v2(L, S, B, axis, n_hint);
    else ghost_pass(L, S, B, axis, n_hint);
}

// simulation.jl:138-144
__global__ void k_remove_ghosts(Store S, int n_verts_init) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S.n_init; i += gridDim.x * blockDim.x) S.nghost[i] = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->n_total = S.n_init;
        cnt->n_verts = n_verts_init;
    }
}
void szk_remove_ghosts(const Launch &L, const Store &S, int n_verts_init) {
    k_remove_ghosts<<<grid_for(L, S.n_init, TPB), TPB, 0, L.stream>>>(S, n_verts_init);
    g_launch_count += 1;
}

// ---- K1: broad phase -----------------------------------------------------------------------------------
__global__ void k_step_reset(Store S) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        S.cfx[i] = 0.0;  // collisions.jl:747-749
        S.cfy[i] = 0.0;
        S.ctrq[i] = 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->n_cand = cnt->n_dom = cnt->n_rows = cnt->n_fuse = cnt->n_pool = 0;
        cnt->n_clipfail = cnt->n_kept = cnt->n_overlap = cnt->n_large = cnt->n_mid = 0;
        cnt->n_domchecks = 0;
        cnt->bb[0] = cnt->bb[1] = ~0ull;
        cnt->bb[2] = cnt->bb[3] = cnt->bb[4] = 0ull;
    }
}

__global__ void k_bbox(Store S) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY, rm = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = S.cx[i], y = S.cy[i];
        xmin = fmin(xmin, x);
        xmax = fmax(xmax, x);
        ymin = fmin(ymin, y);
        ymax = fmax(ymax, y);
        rm = fmax(rm, S.rmax[i]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(FULLMASK, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(FULLMASK, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(FULLMASK, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(FULLMASK, ymax, o));
        rm = fmax(rm, __shfl_xor_sync(FULLMASK, rm, o));
    }
    if (lane_id() == 0 && xmin <= xmax) {
        atomicMin(&cnt->bb[0], enc_f64(xmin));
        atomicMin(&cnt->bb[1], enc_f64(ymin));
        atomicMax(&cnt->bb[2], enc_f64(xmax));
        atomicMax(&cnt->bb[3], enc_f64(ymax));
        atomicMax(&cnt->bb[4], enc_f64(rm));
    }
}

// cell edge >= 2 rmax_max, so every pair passing the circle test lies in adjacent cells; the
// grid only proposes pairs, the exact predicate decides
__global__ void k_grid_setup(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if (cnt->n_total == 0) {
        cnt->gx0 = cnt->gy0 = 0.0;
        cnt->cell = 1.0;
        cnt->gnx = cnt->gny = 1;
        return;
    }
    double xmin = dec_f64(cnt->bb[0]), ymin = dec_f64(cnt->bb[1]);
    double xmax = dec_f64(cnt->bb[2]), ymax = dec_f64(cnt->bb[3]), rm = dec_f64(cnt->bb[4]);
    double cs = 2.0 * rm * (1.0 + 1e-6) + 1e-6;
    double ex = xmax - xmin, ey = ymax - ymin;
    double fx = floor(ex / cs) + 1.0, fy = floor(ey / cs) + 1.0;
    while (fx * fy > (double)B.cap_cells) {
        cs *= 1.5;
        fx = floor(ex / cs) + 1.0;
        fy = floor(ey / cs) + 1.0;
    }
    cnt->gx0 = xmin;
    cnt->gy0 = ymin;
    cnt->cell = cs;
    cnt->gnx = (int)fx;
    cnt->gny = (int)fy;
}

__global__ void k_cell_zero(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int nc = cnt->gnx * cnt->gny;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= nc; c += gridDim.x * blockDim.x) {
        B.cell_count[c] = 0;
        if (c < nc) B.cell_fill[c] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->n_cells = nc;
}

__global__ void k_cell_count(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total, gnx = cnt->gnx, gny = cnt->gny;
    double gx0 = cnt->gx0, gy0 = cnt->gy0, cs = cnt->cell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int ix = (int)floor((S.cx[i] - gx0) / cs), iy = (int)floor((S.cy[i] - gy0) / cs);
        ix = min(max(ix, 0), gnx - 1);
        iy = min(max(iy, 0), gny - 1);
        int c = iy * gnx + ix;
        B.cell_of[i] = c;
        atomicAdd(&B.cell_count[c], 1);
    }
}

// counting-sort scatter; next to the floe index each slot gets a packed copy (cx, cy, rmax, index)
// so the neighbour search below reads contiguous 32-byte records cell by cell instead of gathering
// three scalars per candidate
__global__ void k_cell_fill(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int c = B.cell_of[i];
        int slot = B.cell_start[c] + atomicAdd(&B.cell_fill[c], 1);
        B.cell_items[slot] = i;  // order inside a cell is irrelevant: lists are sorted below
        B.cell_circ[2 * slot] = make_double2(S.cx[i], S.cy[i]);
        B.cell_circ[2 * slot + 1] = make_double2(S.rmax[i], __longlong_as_double((long long)i));
    }
}

// floe_domain_interaction! triggers, collisions.jl:608-660: bit k of the mask = element k hit
// (walls N,S,E,W then topography is handled separately)
__device__ __forceinline__ int wall_mask(const DomainDev *D, double cx, double cy, double r) {
    int m = 0;
    if (cy + r > D->val[0]) m |= 1;
    if (cy - r < D->val[1]) m |= 2;
    if (cx + r > D->val[2]) m |= 4;
    if (cx - r < D->val[3]) m |= 8;
    return m;
}

// one thread per SORTED slot: neighbouring threads are spatial neighbours and walk the same three
// contiguous runs of records (a row of three cells is contiguous in the cell-sorted order)
template <bool WRITE>
__global__ void k_neighbours(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const DomainDev *D = S.dom;
    int n = cnt->n_total, gnx = cnt->gnx, gny = cnt->gny;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const double2 me0 = B.cell_circ[2 * t], me1 = B.cell_circ[2 * t + 1];
        const int i = (int)__double_as_longlong(me1.y);
        int c = B.cell_of[i], ix = c % gnx, iy = c / gnx;
        double xi = me0.x, yi = me0.y, ri = me1.x;
        int up = 0, low = 0;
        int ub = 0, lb = 0;
        if (WRITE) {
            ub = B.up_off[i];
            lb = B.low_off[i];
        }
        const int x_lo = max(ix - 1, 0), x_hi = min(ix + 1, gnx - 1);
        for (int yy = max(iy - 1, 0); yy <= min(iy + 1, gny - 1); ++yy) {
            for (int k = B.cell_start[yy * gnx + x_lo], ke = B.cell_start[yy * gnx + x_hi + 1]; k < ke; ++k) {
                if (k == t) continue;
                const double2 o0 = B.cell_circ[2 * k], o1 = B.cell_circ[2 * k + 1];
                if (!potential_interaction(xi, yi, ri, o0.x, o0.y, o1.x)) continue;
                const int j = (int)__double_as_longlong(o1.y);
                if (j > i) {
                    if (WRITE) B.pair_j[ub + up] = j;
                    up++;
                } else {
                    if (WRITE) B.low_pair[lb + low] = j;
                    low++;
                }
            }
        }
        int wm = wall_mask(D, xi, yi, ri);
        int dc = 0, checks = __popc(wm);
        int db = WRITE ? B.dom_off[i] : 0;
        for (int wl = 0; wl < 4; ++wl)
            if ((wm >> wl) & 1) {
                if (D->kind[wl] != SZ_BOUNDARY_PERIODIC) {  // collisions.jl:459-468: periodic walls do nothing
                    if (WRITE) {
                        B.dom_floe[db + dc] = i;
                        B.dom_elem[db + dc] = wl;
                    }
                    dc++;
                }
            }
        for (int k = 0; k < D->n_topo; ++k)
            if (potential_interaction(S.topo_cx[k], S.topo_cy[k], S.topo_rmax[k], xi, yi, ri)) {  // :650
                if (WRITE) {
                    B.dom_floe[db + dc] = i;
                    B.dom_elem[db + dc] = 4 + k;
                }
                dc++;
                checks++;
            }
        if (!WRITE) {
            B.up_count[i] = up;
            B.low_count[i] = low;
            B.dom_count[i] = dc;
            if (checks) atomicAdd(&cnt->n_domchecks, checks);
        } else {
            // ascending j (own pairs) and ascending i (mirrored rows): the reference's loop order
            for (int a = 1; a < up; ++a) {
                int v = B.pair_j[ub + a], b = a - 1;
                while (b >= 0 && B.pair_j[ub + b] > v) {
                    B.pair_j[ub + b + 1] = B.pair_j[ub + b];
                    --b;
                }
                B.pair_j[ub + b + 1] = v;
            }
            for (int a = 0; a < up; ++a) B.pair_i[ub + a] = i;
            for (int a = 1; a < low; ++a) {
                int v = B.low_pair[lb + a], b = a - 1;
                while (b >= 0 && B.low_pair[lb + b] > v) {
                    B.low_pair[lb + b + 1] = B.low_pair[lb + b];
                    --b;
                }
                B.low_pair[lb + b + 1] = v;
            }
        }
    }
}

__global__ void k_pair_check(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if (cnt->n_cand > B.cap_pairs) {
        cnt->error |= ERR_PAIR_CAP;
        cnt->want_pairs = cnt->n_cand;
    }
    if (cnt->n_dom > B.cap_dom) {
        cnt->error |= ERR_DOM_CAP;
        cnt->want_dom = cnt->n_dom;
    }
}

__device__ __forceinline__ int find_pair(const StepBuf &B, int lo, int hi) {
    int a = B.up_off[lo], b = B.up_off[lo + 1];
    while (a < b) {
        int m = (a + b) >> 1;
        int v = B.pair_j[m];
        if (v < hi) a = m + 1;
        else b = m;
    }
    return (a < B.up_off[lo + 1] && B.pair_j[a] == hi) ? a : -1;
}

__global__ void k_low_link(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        for (int t = B.low_off[j], te = B.low_off[j + 1]; t < te; ++t) B.low_pair[t] = find_pair(B, B.low_pair[t], j);
}

// ---- K2: image-pair filter (collide_pairs Dict, collisions.jl:743,751-775) ------------------------------
// The reference keeps, per id pair, the ghost ids of the FIRST (i, j) met in its serial loop; a
// later image pair runs iff at least one of its ghost ids matches.  "First in the serial loop"
// is the smallest pair index p, found by looking up every image combination of the two parents
// in the sorted pair list — no hash table, no atomics, same answer on every run.
__global__ void k_filter(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int np = cnt->n_cand, nkept = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        int i = B.pair_i[p], j = B.pair_j[p];
        long long idi = S.id[i], idj = S.id[j];
        unsigned char keep = 0;
        if (idi != idj) {
            int ri = S.parent[i] >= 0 ? S.parent[i] : i, rj = S.parent[j] >= 0 ? S.parent[j] : j;
            int ngi = S.nghost[ri], ngj = S.nghost[rj];
            if (ngi == 0 && ngj == 0) keep = 1;
            else {
                int pmin = p;
                for (int a = -1; a < ngi; ++a) {
                    int fa = a < 0 ? ri : S.ghost_slot[ri * SZ_MAX_GHOSTS + a];
                    for (int b = -1; b < ngj; ++b) {
                        int fb = b < 0 ? rj : S.ghost_slot[rj * SZ_MAX_GHOSTS + b];
                        int q = find_pair(B, min(fa, fb), max(fa, fb));
                        if (q >= 0 && q < pmin) pmin = q;
                    }
                }
                int ci = B.pair_i[pmin], cj = B.pair_j[pmin];
                long long g1, g2, G1, G2;
                if (idi > idj) { g1 = S.ghost_id[i]; g2 = S.ghost_id[j]; }
                else { g1 = S.ghost_id[j]; g2 = S.ghost_id[i]; }
                if (S.id[ci] > S.id[cj]) { G1 = S.ghost_id[ci]; G2 = S.ghost_id[cj]; }
                else { G1 = S.ghost_id[cj]; G2 = S.ghost_id[ci]; }
                bool ma = g1 == G1, mb = g2 == G2;
                keep = (ma && mb) || (ma != mb);
            }
        }
        B.keep[p] = keep;
        nkept += keep;
    }
    // n_kept (a diagnostic count): one atomic per warp (the loop above is not warp-uniform in its trip count)
#pragma unroll
    for (int o = 16; o; o >>= 1) nkept += __shfl_xor_sync(FULLMASK, nkept, o);
    if (lane_id() == 0 && nkept) atomicAdd(&cnt->n_kept, nkept);
}

// ---- K3/K4: narrow phase -------------------------------------------------------------------------------------
__device__ void narrow_item(const Ws &w, const Store &S, const StepBuf &B, const Params &P, int slot, bool large) {
    const int lane = lane_id();
    Counters *cnt = S.cnt;
    const DomainDev *D = S.dom;
    const bool is_pair = slot < B.cap_pairs;
    int fi, fj = -1, elem = -1;
    if (is_pair) {
        fi = B.pair_i[slot];
        fj = B.pair_j[slot];
    } else {
        int q = slot - B.cap_pairs;
        fi = B.dom_floe[q];
        elem = B.dom_elem[q];
    }
    const int npp = S.vcount[fi];
    int nqp, kind = SZ_BOUNDARY_COLLISION;
    const double2 *gQ = nullptr;
    if (is_pair) {
        nqp = S.vcount[fj];
        gQ = S.verts + S.vstart[fj];
    } else if (elem < 4) {
        nqp = 5;
        kind = D->kind[elem];
    } else {
        nqp = S.topo_vcount[elem - 4];
        gQ = S.topo_verts + S.topo_vstart[elem - 4];
    }
    int status = CLIP_OK;
    uint32_t flags = 0;
    int nrows = 0;
    if (npp > w.maxv || nqp > w.maxv) status = CLIP_OVERFLOW;
    if (status == CLIP_OK) {
        const double2 *gP = S.verts + S.vstart[fi];
        for (int k = lane; k < npp; k += 32) w.P[k] = gP[k];
        if (gQ) {
            for (int k = lane; k < nqp; k += 32) w.Q[k] = gQ[k];
        } else {
            stage_wall_ring(w.Q, D, elem, lane);
        }
        __syncwarp();
        int nreg = warp_clip(w, w.P, npp, w.Q, nqp, w.R1, w.rs1, w.re1, status);
        if (status == CLIP_FAIL) {
            flags |= IT_CLIPFAIL;
            status = CLIP_OK;
        }
        if (status == CLIP_OK) {
            for (int r = lane; r < nreg; r += 32) w.area1[r] = ring_area_seq(w.R1 + w.rs1[r], w.re1[r] - w.rs1[r]);
            __syncwarp();
            double total = 0.0, max_area = 0.0;
            for (int r = 0; r < nreg; ++r) {
                double a = w.area1[r];
                total += a;
                if (a > max_area) max_area = a;
            }
            const double ai = S.area[fi], hi = S.height[fi];
            int ncontact = 0;
            double ju = 0.0, jv = 0.0, jxi = 0.0, jcx = 0.0, jcy = 0.0;
            if (is_pair) {
                if (total > 0) {  // collisions.jl:364-405
                    flags |= IT_OVERLAP;
                    const double aj = S.area[fj];
                    if (fmax(total / ai, total / aj) > P.cfg.floe_floe_max_overlap) {
                        flags |= IT_FUSE;
                    } else {
                        const double hj = S.height[fj];
                        double ir = sqrt(ai), jr = sqrt(aj);
                        double ff = (ir > 1e5 || jr > 1e5) ? P.cfg.E * fmin(hi, hj) / fmin(ir, jr)
                                                           : P.cfg.E * (hi * hj) / (hi * jr + hj * ir);
                        ncontact = warp_elastic_forces(w, w.P, npp, w.Q, nqp, nreg, ff, status, flags);
                        ju = S.u[fj];
                        jv = S.v[fj];
                        jxi = S.xi[fj];
                        jcx = S.cx[fj];
                        jcy = S.cy[fj];
                    }
                }
            } else if (kind == SZ_BOUNDARY_OPEN) {  // collisions.jl:427-441
                if (total > 0) flags |= IT_OVERLAP | IT_REMOVE;
            } else if (max_area > 0) {  // collisions.jl:522-555
                flags |= IT_OVERLAP;
                if (max_area / ai > P.cfg.floe_domain_max_overlap) {
                    flags |= IT_REMOVE;
                } else {
                    double ff = P.cfg.E * hi / sqrt(ai);
                    ncontact = warp_elastic_forces(w, w.P, npp, w.Q, nqp, nreg, ff, status, flags);
                    if (elem < 4 && kind == SZ_BOUNDARY_MOVING) {  // boundaries.jl:522
                        ju = D->wu[elem];
                        jv = D->wv[elem];
                    }
                }
            }
            if (status == CLIP_OK && lane == 0 && ncontact > 0) {
                const double iu = S.u[fi], iv = S.v[fi], ixi = S.xi[fi], icx = S.cx[fi], icy = S.cy[fi];
                for (int k = 0; k < ncontact; ++k) {
                    double *c = w.ct + 6 * k;
                    if (!is_pair && elem < 4) {  // _normal_direction_correct!, boundaries.jl:37-40,73-76,110-113,147-150
                        if (elem == 0 && c[3] >= D->val[0]) c[0] = 0.0;
                        if (elem == 1 && c[3] <= D->val[1]) c[0] = 0.0;
                        if (elem == 2 && c[2] >= D->val[2]) c[1] = 0.0;
                        if (elem == 3 && c[2] <= D->val[3]) c[1] = 0.0;
                    }
                    double fr[2];
                    friction_force(P.cfg.E, P.cfg.nu, P.cfg.mu, (double)P.cfg.dt, iu, iv, ixi, icx, icy, ju, jv, jxi,
                                   jcx, jcy, c, fr);
                    double fx = c[0] + fr[0], fy = c[1] + fr[1];
                    if (fx != 0 || fy != 0) {  // add_interactions!, collisions.jl:288
                        double *o = w.ct + 6 * nrows;
                        double px = c[2], py = c[3], ov = c[4];
                        o[0] = fx;
                        o[1] = fy;
                        o[2] = px;
                        o[3] = py;
                        o[4] = ov;
                        nrows++;
                    }
                }
            }
        }
    }
    if (lane == 0) {
        if (status == CLIP_OVERFLOW) {
            if (large) {
                atomicOr(&cnt->error, ERR_POLY_TOO_LARGE);
            } else {
                int s = atomicAdd(&cnt->n_large, 1);
                B.large_items[s] = slot;
            }
            B.item_nrows[slot] = 0;
            B.item_flags[slot] = IT_NEEDLARGE;
        } else {
            int row0 = 0;
            if (nrows > 0) {
                row0 = atomicAdd(&cnt->n_pool, nrows);
                if (row0 + nrows > B.cap_pool) {
                    atomicOr(&cnt->error, ERR_POOL_CAP);
                    nrows = 0;
                } else {
                    for (int k = 0; k < nrows; ++k)
                        for (int q = 0; q < NPOOL; ++q) B.pool[(size_t)(row0 + k) * NPOOL + q] = w.ct[6 * k + q];
                }
            }
            B.item_nrows[slot] = nrows;
            B.item_row0[slot] = row0;
            B.item_flags[slot] = flags | IT_DONE;
            if (flags & IT_CLIPFAIL) atomicAdd(&cnt->n_clipfail, 1);
            if (is_pair && (flags & IT_OVERLAP)) atomicAdd(&cnt->n_overlap, 1);
            if (flags & IT_FUSE) {
                int s = atomicAdd(&cnt->n_fuse, 1);
                if (s < B.cap_fuse) B.fuse_pairs[s] = make_int2(fi, fj);
                else atomicOr(&cnt->error, ERR_FUSE_CAP);
            }
        }
    }
    __syncwarp();
}

__global__ void k_narrow(Store S, StepBuf B, Params P, int maxv, int maxx, int large) {
    sz_pdl();
    extern __shared__ __align__(16) unsigned char smem[];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    const int nl = large ? cnt->n_large : cnt->n_mid;
    const int *list = large ? B.large_items : B.mid_items;
    for (int it = blockIdx.x * wpb + wib; it < nl; it += gridDim.x * wpb) narrow_item(w, S, B, P, list[it], large != 0);
}

__global__ void k_pool_check(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->n_pool > B.cap_pool) cnt->want_pool = cnt->n_pool;
    if (cnt->n_fuse > B.cap_fuse) cnt->want_fuse = cnt->n_fuse;
}

// ---- K5: status, rows, totals ------------------------------------------------------------------------------------
// status.tag after the pair loop and floe_domain_interaction! (collisions.jl:366-368,438,524),
// in the reference's order: fuse from own pairs first, then a domain removal overrides it.
__global__ void k_status(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = S.status[i];
        for (int p = B.up_off[i], pe = B.up_off[i + 1]; p < pe; ++p)
            if (B.keep[p] && (B.item_flags[p] & IT_FUSE)) s = SZ_STATUS_FUSE;
        for (int q = B.dom_off[i], qe = B.dom_off[i + 1]; q < qe; ++q)
            if (B.item_flags[B.cap_pairs + q] & IT_REMOVE) s = SZ_STATUS_REMOVE;
        S.status[i] = s;
    }
}

// serial fuse propagation, collisions.jl:799-806: ascending i hands its tag to the partners in
// fuse_idx; since partners recorded by the pair loop are always j > i the serial result is the
// closure of "i fused => j fused" over the fuse pairs, computed here as a fixed point.
__global__ void k_fuse_propagate(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int nf = min(cnt->n_fuse, B.cap_fuse);
    if (nf == 0) return;
    __shared__ int changed;
    do {
        __syncthreads();
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        for (int k = threadIdx.x; k < nf; k += blockDim.x) {
            int2 pr = B.fuse_pairs[k];
            if (S.status[pr.x] == SZ_STATUS_FUSE && S.status[pr.y] != SZ_STATUS_FUSE) {
                S.status[pr.y] = SZ_STATUS_FUSE;
                changed = 1;
            }
        }
        __syncthreads();
    } while (changed);
}

__global__ void k_row_count(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        int c = 0;
        for (int p = B.up_off[f], pe = B.up_off[f + 1]; p < pe; ++p)
            if (B.keep[p]) c += B.item_nrows[p];
        for (int q = B.dom_off[f], qe = B.dom_off[f + 1]; q < qe; ++q) c += B.item_nrows[B.cap_pairs + q];
        for (int t = B.low_off[f], te = B.low_off[f + 1]; t < te; ++t) {
            int p = B.low_pair[t];
            if (B.keep[p]) c += B.item_nrows[p];
        }
        B.row_pre[f] = c;
    }
}

__global__ void k_row_total(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        int c = B.row_pre[f];
        if (f < S.n_init)
            for (int g = 0, ng = S.nghost[f]; g < ng; ++g) c += B.row_pre[S.ghost_slot[f * SZ_MAX_GHOSTS + g]];
        B.row_count[f] = c;
    }
}

__global__ void k_row_check(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if (cnt->n_rows > B.cap_rows) {
        cnt->error |= ERR_ROW_CAP;
        cnt->want_rows = cnt->n_rows;
    }
}

// rows of floe f in the reference's order: own pairs (j ascending, regions in clip order), walls
// N,S,E,W, topography (collisions.jl:776-794), mirrored rows (i ascending, :808-827).  (sx, sy)
// is the ghost->parent shift of the contact points (:835-838).  The per-floe totals (collisions.jl:830-862,
// 673-686) are accumulated while the rows are written, in row order: the same bits as a second pass over the
// rows, without reading them back.
struct RowAcc {
    double oa, sx, sy, st;
};
__device__ __forceinline__ void emit_row(double *r, double idx, double fx, double fy, double px, double py, double ov, bool sums,
                                         double cx, double cy, RowAcc &acc) {
    double trq = 0.0;
    acc.oa += ov;
    if (sums) {
        double xp = px - cx, yp = py - cy;
        trq = xp * fy - yp * fx;
        acc.sx += fx;
        acc.sy += fy;
        acc.st += trq;
    }
    r[COL_IDX] = idx;
    r[COL_FX] = fx;
    r[COL_FY] = fy;
    r[COL_PX] = px;
    r[COL_PY] = py;
    r[COL_TRQ] = trq;
    r[COL_OV] = ov;
}
__device__ int emit_base_rows(const Store &S, const StepBuf &B, int f, double *dst, double sx, double sy, bool shift, bool sums,
                              double cx, double cy, RowAcc &acc) {
    int n = 0;
    for (int p = B.up_off[f], pe = B.up_off[f + 1]; p < pe; ++p) {
        if (!B.keep[p]) continue;
        const double *src = B.pool + (size_t)B.item_row0[p] * NPOOL;
        double idx = (double)(B.pair_j[p] + 1);
        for (int k = 0, nk = B.item_nrows[p]; k < nk; ++k, ++n) {
            const double *c = src + k * NPOOL;
            emit_row(dst + (size_t)n * NCOL, idx, c[0], c[1], shift ? c[2] - sx : c[2], shift ? c[3] - sy : c[3], c[4], sums, cx, cy, acc);
        }
    }
    for (int q = B.dom_off[f], qe = B.dom_off[f + 1]; q < qe; ++q) {
        int slot = B.cap_pairs + q;
        const double *src = B.pool + (size_t)B.item_row0[slot] * NPOOL;
        double idx = (double)(-(B.dom_elem[q] + 1));
        for (int k = 0, nk = B.item_nrows[slot]; k < nk; ++k, ++n) {
            const double *c = src + k * NPOOL;
            emit_row(dst + (size_t)n * NCOL, idx, c[0], c[1], shift ? c[2] - sx : c[2], shift ? c[3] - sy : c[3], c[4], sums, cx, cy, acc);
        }
    }
    for (int t = B.low_off[f], te = B.low_off[f + 1]; t < te; ++t) {
        int p = B.low_pair[t];
        if (!B.keep[p]) continue;
        const double *src = B.pool + (size_t)B.item_row0[p] * NPOOL;
        double idx = (double)(B.pair_i[p] + 1);
        for (int k = 0, nk = B.item_nrows[p]; k < nk; ++k, ++n) {
            const double *c = src + k * NPOOL;
            emit_row(dst + (size_t)n * NCOL, idx, -c[0], -c[1], shift ? c[2] - sx : c[2], shift ? c[3] - sy : c[3], c[4], sums, cx, cy, acc);
        }
    }
    return n;
}

// one thread per floe: its rows are written and summed in row order, so collision_force /
// collision_trq / overarea are the same bits on every run (collisions.jl:830-862, 673-686)
__global__ void k_row_write(Store S, StepBuf B) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        double *dst = B.rows + (size_t)B.row_off[f] * NCOL;
        const int par = S.parent[f];
        const bool sums = f < S.n_init;
        const double cx = S.cx[f], cy = S.cy[f];
        RowAcc acc = {S.overarea[f], 0.0, 0.0, 0.0};
        int nr;
        if (f >= S.n_init && par >= 0) {
            nr = emit_base_rows(S, B, f, dst, cx - S.cx[par], cy - S.cy[par], true, sums, cx, cy, acc);
        } else {
            nr = emit_base_rows(S, B, f, dst, 0.0, 0.0, false, sums, cx, cy, acc);
        }
        if (sums) {
            for (int g = 0, ng = S.nghost[f]; g < ng; ++g) {
                int gi = S.ghost_slot[f * SZ_MAX_GHOSTS + g];
                nr += emit_base_rows(S, B, gi, dst + (size_t)nr * NCOL, S.cx[gi] - cx, S.cy[gi] - cy, true, sums, cx, cy, acc);
            }
        }
        S.overarea[f] = acc.oa;
        if (sums) {
            S.cfx[f] += acc.sx;
            S.cfy[f] += acc.sy;
            S.ctrq[f] += acc.st;
        }
    }
}

// update_boundaries!, collisions.jl:565-571; _update_boundary!, boundaries.jl:526-544
__global__ void k_update_boundaries(Store S, Params P) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    DomainDev *D = S.dom;
    for (int wl = 0; wl < 4; ++wl) {
        if (D->kind[wl] != SZ_BOUNDARY_MOVING) continue;
        if (wl < 2) {
            double d = D->wv[wl] * P.cfg.dt;
            D->rect[wl][2] += d;
            D->rect[wl][3] += d;
            D->val[wl] += d;
        } else {
            double d = D->wu[wl] * P.cfg.dt;
            D->rect[wl][0] += d;
            D->rect[wl][1] += d;
            D->val[wl] += d;
        }
    }
}

// ---- the fused chain (v2): 9 + 5 launches instead of 19 + 10 around the narrow phase ---------------------------------------
// k_bbox2: k_step_reset's per-floe zeroing + k_bbox.  The bounding-box words and the step counters are reset by
// k_cell_count2 (after the last reader of the box) for the NEXT collision step; k_set_counts initialises them.
__global__ void k_bbox2(Store S) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    int n = cnt->n_total;
    double xmin = INFINITY, ymin = INFINITY, xmax = -INFINITY, ymax = -INFINITY, rm = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        S.cfx[i] = 0.0;  // collisions.jl:747-749
        S.cfy[i] = 0.0;
        S.ctrq[i] = 0.0;
        double x = S.cx[i], y = S.cy[i];
        xmin = fmin(xmin, x);
        xmax = fmax(xmax, x);
        ymin = fmin(ymin, y);
        ymax = fmax(ymax, y);
        rm = fmax(rm, S.rmax[i]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(FULLMASK, xmin, o));
        ymin = fmin(ymin, __shfl_xor_sync(FULLMASK, ymin, o));
        xmax = fmax(xmax, __shfl_xor_sync(FULLMASK, xmax, o));
        ymax = fmax(ymax, __shfl_xor_sync(FULLMASK, ymax, o));
        rm = fmax(rm, __shfl_xor_sync(FULLMASK, rm, o));
    }
    if (lane_id() == 0 && xmin <= xmax) {
        atomicMin(&cnt->bb[0], enc_f64(xmin));
        atomicMin(&cnt->bb[1], enc_f64(ymin));
        atomicMax(&cnt->bb[2], enc_f64(xmax));
        atomicMax(&cnt->bb[3], enc_f64(ymax));
        atomicMax(&cnt->bb[4], enc_f64(rm));
    }
}

struct GridGeom {
    double gx0, gy0, cell;
    int gnx, gny;
};
// k_grid_setup's arithmetic, evaluated redundantly by every thread that needs it (a handful of flops)
__device__ __forceinline__ GridGeom grid_geom(const Counters *cnt, int cap_cells) {
    GridGeom g;
    if (cnt->n_total == 0) {
        g.gx0 = g.gy0 = 0.0;
        g.cell = 1.0;
        g.gnx = g.gny = 1;
        return g;
    }
    double xmin = dec_f64(cnt->bb[0]), ymin = dec_f64(cnt->bb[1]);
    double xmax = dec_f64(cnt->bb[2]), ymax = dec_f64(cnt->bb[3]), rm = dec_f64(cnt->bb[4]);
    double cs = 2.0 * rm * (1.0 + 1e-6) + 1e-6;
    double ex = xmax - xmin, ey = ymax - ymin;
    double fx = floor(ex / cs) + 1.0, fy = floor(ey / cs) + 1.0;
    while (fx * fy > (double)cap_cells) {
        cs *= 1.5;
        fx = floor(ex / cs) + 1.0;
        fy = floor(ey / cs) + 1.0;
    }
    g.gx0 = xmin;
    g.gy0 = ymin;
    g.cell = cs;
    g.gnx = (int)fx;
    g.gny = (int)fy;
    return g;
}

// k_grid_setup + k_cell_zero + the reset of every look-back scan of this step
__global__ void k_grid_zero(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const GridGeom g = grid_geom(cnt, B.cap_cells);
    const int nc = g.gnx * g.gny;
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int c = t0; c <= nc; c += nt) {
        B.cell_count[c] = 0;
        if (c < nc) B.cell_fill[c] = 0;
    }
    for (int k = t0; k < LB_SLOTS * 3 * B.lb_stride; k += nt) B.lb_desc[k] = 0ull;
    if (t0 < LB_SLOTS * 3) B.lb_ticket[t0] = 0;
    if (t0 == 0) {
        cnt->gx0 = g.gx0;
        cnt->gy0 = g.gy0;
        cnt->cell = g.cell;
        cnt->gnx = g.gnx;
        cnt->gny = g.gny;
        cnt->n_cells = nc;
    }
}

// k_cell_count; thread 0 then resets what k_step_reset used to reset (the box has no reader left in this step)
__global__ void k_cell_count2(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int n = cnt->n_total, gnx = cnt->gnx, gny = cnt->gny;
    const double gx0 = cnt->gx0, gy0 = cnt->gy0, cs = cnt->cell;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int ix = (int)floor((S.cx[i] - gx0) / cs), iy = (int)floor((S.cy[i] - gy0) / cs);
        ix = min(max(ix, 0), gnx - 1);
        iy = min(max(iy, 0), gny - 1);
        int c = iy * gnx + ix;
        B.cell_of[i] = c;
        atomicAdd(&B.cell_count[c], 1);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->n_cand = cnt->n_dom = cnt->n_rows = cnt->n_fuse = cnt->n_pool = 0;
        cnt->n_clipfail = cnt->n_kept = cnt->n_overlap = cnt->n_large = cnt->n_mid = 0;
        cnt->n_domchecks = 0;
        cnt->bb[0] = cnt->bb[1] = ~0ull;  // for the next collision step (k_bbox2 only accumulates)
        cnt->bb[2] = cnt->bb[3] = cnt->bb[4] = 0ull;
    }
}

// ONE neighbour search: the count pass also parks the indices it finds (up to NB_K per floe, in the cell-sorted
// slot order so that a warp writes 32 consecutive ints per k); the write pass only reads them back, splits them into
// j > i / j < i, sorts and stores — floes with more than NB_K candidates search again.
#define NB_K 20
template <bool WRITE>
__global__ void k_neighbours2(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if (WRITE && (cnt->n_cand > B.cap_pairs || cnt->n_dom > B.cap_dom)) {  // k_pair_check: every thread backs off, one reports
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (cnt->n_cand > B.cap_pairs) { atomicOr(&cnt->error, ERR_PAIR_CAP); cnt->want_pairs = cnt->n_cand; }
            if (cnt->n_dom > B.cap_dom) { atomicOr(&cnt->error, ERR_DOM_CAP); cnt->want_dom = cnt->n_dom; }
        }
        return;
    }
    const DomainDev *D = S.dom;
    const int n = cnt->n_total, gnx = cnt->gnx, gny = cnt->gny, cap = S.cap_floes;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const double2 me0 = B.cell_circ[2 * t], me1 = B.cell_circ[2 * t + 1];
        const int i = (int)__double_as_longlong(me1.y);
        const double xi = me0.x, yi = me0.y, ri = me1.x;
        int up = 0, low = 0, ub = 0, lb = 0, db = 0;
        bool research = true, sorted = false;
        if (WRITE) {
            ub = B.up_off[i];
            lb = B.low_off[i];
            db = B.dom_off[i];
            const int tot = (B.up_off[i + 1] - ub) + (B.low_off[i + 1] - lb);
            if (tot <= NB_K) {
                // sorted in thread-local storage (one sort of the whole list: j < i first, then j > i), stored once
                research = false;
                sorted = true;
                int nb[NB_K];
                for (int k = 0; k < tot; ++k) {
                    const int v = B.nb_scratch[(size_t)k * cap + t];
                    int b = k - 1;
                    while (b >= 0 && nb[b] > v) {
                        nb[b + 1] = nb[b];
                        --b;
                    }
                    nb[b + 1] = v;
                }
                low = B.low_off[i + 1] - lb;
                up = tot - low;
                for (int k = 0; k < low; ++k) B.low_pair[lb + k] = nb[k];
                for (int k = 0; k < up; ++k) {
                    B.pair_j[ub + k] = nb[low + k];
                    B.pair_i[ub + k] = i;
                }
            }
        }
        if (research) {
            const int c = B.cell_of[i], ix = c % gnx, iy = c / gnx;
            const int x_lo = max(ix - 1, 0), x_hi = min(ix + 1, gnx - 1);
            int found = 0;
            for (int yy = max(iy - 1, 0); yy <= min(iy + 1, gny - 1); ++yy) {
                for (int k = B.cell_start[yy * gnx + x_lo], ke = B.cell_start[yy * gnx + x_hi + 1]; k < ke; ++k) {
                    if (k == t) continue;
                    const double2 o0 = B.cell_circ[2 * k], o1 = B.cell_circ[2 * k + 1];
                    if (!potential_interaction(xi, yi, ri, o0.x, o0.y, o1.x)) continue;
                    const int j = (int)__double_as_longlong(o1.y);
                    if (!WRITE && found < NB_K) B.nb_scratch[(size_t)found * cap + t] = j;
                    found++;
                    if (j > i) {
                        if (WRITE) B.pair_j[ub + up] = j;
                        up++;
                    } else {
                        if (WRITE) B.low_pair[lb + low] = j;
                        low++;
                    }
                }
            }
        }
        const int wm = wall_mask(D, xi, yi, ri);
        int dc = 0, checks = __popc(wm);
        for (int wl = 0; wl < 4; ++wl)
            if ((wm >> wl) & 1) {
                if (D->kind[wl] != SZ_BOUNDARY_PERIODIC) {  // collisions.jl:459-468: periodic walls do nothing
                    if (WRITE) {
                        B.dom_floe[db + dc] = i;
                        B.dom_elem[db + dc] = wl;
                    }
                    dc++;
                }
            }
        for (int k = 0; k < D->n_topo; ++k)
            if (potential_interaction(S.topo_cx[k], S.topo_cy[k], S.topo_rmax[k], xi, yi, ri)) {  // :650
                if (WRITE) {
                    B.dom_floe[db + dc] = i;
                    B.dom_elem[db + dc] = 4 + k;
                }
                dc++;
                checks++;
            }
        if (!WRITE) {
            B.up_count[i] = up;
            B.low_count[i] = low;
            B.dom_count[i] = dc;
            if (checks) atomicAdd(&cnt->n_domchecks, checks);
        } else if (!sorted) {
            // ascending j (own pairs) and ascending i (mirrored rows): the reference's loop order
            for (int a = 1; a < up; ++a) {
                int v = B.pair_j[ub + a], b = a - 1;
                while (b >= 0 && B.pair_j[ub + b] > v) {
                    B.pair_j[ub + b + 1] = B.pair_j[ub + b];
                    --b;
                }
                B.pair_j[ub + b + 1] = v;
            }
            for (int a = 0; a < up; ++a) B.pair_i[ub + a] = i;
            for (int a = 1; a < low; ++a) {
                int v = B.low_pair[lb + a], b = a - 1;
                while (b >= 0 && B.low_pair[lb + b] > v) {
                    B.low_pair[lb + b + 1] = B.low_pair[lb + b];
                    --b;
                }
                B.low_pair[lb + b + 1] = v;
            }
        }
    }
}

// k_low_link (blocks [0, nb_link)) and k_filter (the rest) in one launch: both only read the sorted own-pair lists
__global__ void k_link_filter(Store S, StepBuf B, int nb_link) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if ((int)blockIdx.x < nb_link) {
        const int n = cnt->n_total;
        for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += nb_link * blockDim.x)
            for (int t = B.low_off[j], te = B.low_off[j + 1]; t < te; ++t) B.low_pair[t] = find_pair(B, B.low_pair[t], j);
        return;
    }
    const int np = cnt->n_cand, nb = gridDim.x - nb_link;
    int nkept = 0;
    for (int p = (blockIdx.x - nb_link) * blockDim.x + threadIdx.x; p < np; p += nb * blockDim.x) {
        int i = B.pair_i[p], j = B.pair_j[p];
        long long idi = S.id[i], idj = S.id[j];
        unsigned char keep = 0;
        if (idi != idj) {
            int ri = S.parent[i] >= 0 ? S.parent[i] : i, rj = S.parent[j] >= 0 ? S.parent[j] : j;
            int ngi = S.nghost[ri], ngj = S.nghost[rj];
            if (ngi == 0 && ngj == 0) keep = 1;
            else {
                int pmin = p;
                for (int a = -1; a < ngi; ++a) {
                    int fa = a < 0 ? ri : S.ghost_slot[ri * SZ_MAX_GHOSTS + a];
                    for (int b = -1; b < ngj; ++b) {
                        int fb = b < 0 ? rj : S.ghost_slot[rj * SZ_MAX_GHOSTS + b];
                        int q = find_pair(B, min(fa, fb), max(fa, fb));
                        if (q >= 0 && q < pmin) pmin = q;
                    }
                }
                int ci = B.pair_i[pmin], cj = B.pair_j[pmin];
                long long g1, g2, G1, G2;
                if (idi > idj) { g1 = S.ghost_id[i]; g2 = S.ghost_id[j]; }
                else { g1 = S.ghost_id[j]; g2 = S.ghost_id[i]; }
                if (S.id[ci] > S.id[cj]) { G1 = S.ghost_id[ci]; G2 = S.ghost_id[cj]; }
                else { G1 = S.ghost_id[cj]; G2 = S.ghost_id[ci]; }
                bool ma = g1 == G1, mb = g2 == G2;
                keep = (ma && mb) || (ma != mb);
            }
        }
        B.keep[p] = keep;
        nkept += keep;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) nkept += __shfl_xor_sync(FULLMASK, nkept, o);
    if (lane_id() == 0 && nkept) atomicAdd(&cnt->n_kept, nkept);
}

// k_pool_check + k_status + k_row_count
__global__ void k_status_rowcount(Store S, StepBuf B) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (cnt->n_pool > B.cap_pool) cnt->want_pool = cnt->n_pool;
        if (cnt->n_fuse > B.cap_fuse) cnt->want_fuse = cnt->n_fuse;
    }
    if (cnt->error) return;
    int n = cnt->n_total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        int s = S.status[i], c = 0;
        for (int p = B.up_off[i], pe = B.up_off[i + 1]; p < pe; ++p)
            if (B.keep[p]) {
                if (B.item_flags[p] & IT_FUSE) s = SZ_STATUS_FUSE;
                c += B.item_nrows[p];
            }
        for (int q = B.dom_off[i], qe = B.dom_off[i + 1]; q < qe; ++q) {
            if (B.item_flags[B.cap_pairs + q] & IT_REMOVE) s = SZ_STATUS_REMOVE;
            c += B.item_nrows[B.cap_pairs + q];
        }
        for (int t = B.low_off[i], te = B.low_off[i + 1]; t < te; ++t) {
            int p = B.low_pair[t];
            if (B.keep[p]) c += B.item_nrows[p];
        }
        S.status[i] = s;
        B.row_pre[i] = c;
    }
}

// k_row_check + k_row_write + k_update_boundaries
__global__ void k_row_write2(Store S, StepBuf B, Params P) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    if (cnt->n_rows > B.cap_rows) {  // every thread backs off, one reports
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicOr(&cnt->error, ERR_ROW_CAP);
            cnt->want_rows = cnt->n_rows;
        }
        return;
    }
    int n = cnt->n_total;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) {
        double *dst = B.rows + (size_t)B.row_off[f] * NCOL;
        const int par = S.parent[f];
        const bool sums = f < S.n_init;
        const double cx = S.cx[f], cy = S.cy[f];
        RowAcc acc = {S.overarea[f], 0.0, 0.0, 0.0};
        int nr;
        if (f >= S.n_init && par >= 0) {
            nr = emit_base_rows(S, B, f, dst, cx - S.cx[par], cy - S.cy[par], true, sums, cx, cy, acc);
        } else {
            nr = emit_base_rows(S, B, f, dst, 0.0, 0.0, false, sums, cx, cy, acc);
        }
        if (sums) {
            for (int g = 0, ng = S.nghost[f]; g < ng; ++g) {
                int gi = S.ghost_slot[f * SZ_MAX_GHOSTS + g];
                nr += emit_base_rows(S, B, gi, dst + (size_t)nr * NCOL, S.cx[gi] - cx, S.cy[gi] - cy, true, sums, cx, cy, acc);
            }
        }
        S.overarea[f] = acc.oa;
        if (sums) {
            S.cfx[f] += acc.sx;
            S.cfy[f] += acc.sy;
            S.ctrq[f] += acc.st;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // update_boundaries!, collisions.jl:565-571
        DomainDev *D = S.dom;
        for (int wl = 0; wl < 4; ++wl) {
            if (D->kind[wl] != SZ_BOUNDARY_MOVING) continue;
            if (wl < 2) {
                double d = D->wv[wl] * P.cfg.dt;
                D->rect[wl][2] += d;
                D->rect[wl][3] += d;
                D->val[wl] += d;
            } else {
                double d = D->wu[wl] * P.cfg.dt;
                D->rect[wl][0] += d;
                D->rect[wl][1] += d;
                D->val[wl] += d;
            }
        }
    }
}

static void collisions_v2(const Launch &L, const Store &S, const StepBuf &B, const Params &P, int n_hint, int pairs_hint,
                          cudaEvent_t *ev, const cudaEvent_t *waits) {
    cudaStream_t st = L.stream;
    const bool pdl = L.pdl != 0;
    const int gf = grid_for(L, n_hint, TPB);
    sz_launch_pdl(false, k_bbox2, dim3(gf), dim3(TPB), 0, st, S);
    sz_launch_pdl(pdl, k_grid_zero, dim3(grid_for(L, B.cap_cells, TPB)), dim3(TPB), 0, st, S, B);
    sz_launch_pdl(pdl, k_cell_count2, dim3(gf), dim3(TPB), 0, st, S, B);
    {
        const int *in[1] = {B.cell_count};
        int *out[1] = {B.cell_start}, *tot[1] = {nullptr};
        scan_lb(L, S, B, 0, 1, in, out, tot, &S.cnt->n_cells, 0, B.cap_cells);
    }
    sz_launch_pdl(pdl, k_cell_fill, dim3(gf), dim3(TPB), 0, st, S, B);
    sz_launch_pdl(pdl, k_neighbours2<false>, dim3(sz_div_up(n_hint, 128)), dim3(128), 0, st, S, B);
    {
        const int *in[3] = {B.up_count, B.low_count, B.dom_count};
        int *out[3] = {B.up_off, B.low_off, B.dom_off}, *tot[3] = {&S.cnt->n_cand, nullptr, &S.cnt->n_dom};
        scan_lb(L, S, B, 1, 3, in, out, tot, &S.cnt->n_total, 0, S.cap_floes);
    }
    sz_launch_pdl(pdl, k_neighbours2<true>, dim3(sz_div_up(n_hint, 128)), dim3(128), 0, st, S, B);
    const int gp = grid_for(L, pairs_hint, TPB);
    sz_launch_pdl(pdl, k_link_filter, dim3(gf + gp), dim3(TPB), 0, st, S, B, gf);
    if (ev) sz_record(L, ev[0], st);
    if (waits) cudaStreamWaitEvent(st, waits[0], 0);  // sz_step_host: rings and height have landed
    const int maxv_s = 32, maxx_s = 16, wpb = 4;
    int gi = grid_for(L, (long long)pairs_hint + n_hint / 8 + 64, 256);
    sz_launch_pdl(pdl, k_item_count, dim3(gi), dim3(256), 0, st, S, B);
    sz_launch_pdl(pdl, k_class_scan, dim3(1), dim3(32), 0, st, S, B);
    sz_launch_pdl(pdl, k_item_scatter, dim3(gi), dim3(256), 0, st, S, B);
    sz_launch_pdl(pdl, k_narrow_ab<0>, dim3(3 * L.sms), dim3(TN_NT), TN_SMEM_A, st, S, B, P);
    sz_launch_pdl(pdl, k_narrow_ab<1>, dim3(3 * L.sms), dim3(TN_NT), TN_SMEM_B, st, S, B, P);
    sz_launch_pdl(pdl, k_narrow, dim3(L.sms * 4), dim3(wpb * 32), wpb * ws_bytes(maxv_s, maxx_s), st, S, B, P, maxv_s, maxx_s, 0);
    sz_launch_pdl(pdl, k_narrow, dim3(L.sms), dim3(32), ws_bytes(L.maxv_large, L.maxx_large), st, S, B, P, L.maxv_large, L.maxx_large, 1);
    if (ev) sz_record(L, ev[1], st);
    if (waits) cudaStreamWaitEvent(st, waits[1], 0);  // sz_step_host: everything else (overarea is accumulated by k_row_write)
    sz_launch_pdl(pdl, k_status_rowcount, dim3(gf), dim3(TPB), 0, st, S, B);
    sz_launch_pdl(pdl, k_fuse_propagate, dim3(1), dim3(1024), 0, st, S, B);
    sz_launch_pdl(pdl, k_row_total, dim3(gf), dim3(TPB), 0, st, S, B);
    {
        const int *in[1] = {B.row_count};
        int *out[1] = {B.row_off}, *tot[1] = {&S.cnt->n_rows};
        scan_lb(L, S, B, 2, 1, in, out, tot, &S.cnt->n_total, 0, S.cap_floes);
    }
    sz_launch_pdl(pdl, k_row_write2, dim3(sz_div_up(n_hint, 128)), dim3(128), 0, st, S, B, P);
    if (ev) sz_record(L, ev[2], st);
    g_launch_count += 18;  // + the three look-back scans counted in scan_lb
}

void szk_collisions(const Launch &L, const Store &S, const StepBuf &B, const Params &P, int n_hint, int pairs_hint,
                    cudaEvent_t *ev, const cudaEvent_t *waits) {
    if (L.chain_v2) {
        collisions_v2(L, S, B, P, n_hint, pairs_hint, ev, waits);
        return;
    }
    cudaStream_t st = L.stream;
    int gf = grid_for(L, n_hint, TPB);
    k_step_reset<<<gf, TPB, 0, st>>>(S);
    k_bbox<<<gf, TPB, 0, st>>>(S);
    k_grid_setup<<<1, 1, 0, st>>>(S, B);
    k_cell_zero<<<grid_for(L, B.cap_cells, TPB), TPB, 0, st>>>(S, B);
    k_cell_count<<<gf, TPB, 0, st>>>(S, B);
    scan_excl(L, S, B.cell_count, B.cell_start, &S.cnt->n_cells, 0, B.cap_cells, B.scan_block, nullptr);
    k_cell_fill<<<gf, TPB, 0, st>>>(S, B);
    k_neighbours<false><<<sz_div_up(n_hint, 128), 128, 0, st>>>(S, B);
    {
        Scan3 a = {{B.up_count, B.low_count, B.dom_count}, {B.up_off, B.low_off, B.dom_off}, {&S.cnt->n_cand, nullptr, &S.cnt->n_dom}};
        scan_excl3(L, S, a, &S.cnt->n_total, 0, S.cap_floes, B.scan_block);
    }
    k_pair_check<<<1, 1, 0, st>>>(S, B);
    k_neighbours<true><<<sz_div_up(n_hint, 128), 128, 0, st>>>(S, B);
    k_low_link<<<gf, TPB, 0, st>>>(S, B);
    int gp = grid_for(L, pairs_hint, TPB);
    k_filter<<<gp, TPB, 0, st>>>(S, B);
    if (ev) sz_record(L, ev[0], st);
    if (waits) cudaStreamWaitEvent(st, waits[0], 0);  // sz_step_host: rings and height have landed
    const int maxv_s = 32, maxx_s = 16, wpb = 4;
    // thread-per-item fast path (small polygons), then warp-per-item for what it handed on, then the
    // large-polygon workspace for what that one handed on
    int gi = grid_for(L, (long long)pairs_hint + n_hint / 8 + 64, 256);
    k_item_count<<<gi, 256, 0, st>>>(S, B);
    k_class_scan<<<1, 32, 0, st>>>(S, B);
    k_item_scatter<<<gi, 256, 0, st>>>(S, B);
    k_narrow_ab<0><<<3 * L.sms, TN_NT, TN_SMEM_A, st>>>(S, B, P);
    k_narrow_ab<1><<<3 * L.sms, TN_NT, TN_SMEM_B, st>>>(S, B, P);
    k_narrow<<<L.sms * 4, wpb * 32, wpb * ws_bytes(maxv_s, maxx_s), st>>>(S, B, P, maxv_s, maxx_s, 0);
    k_narrow<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), st>>>(S, B, P, L.maxv_large, L.maxx_large, 1);
    k_pool_check<<<1, 1, 0, st>>>(S, B);
    if (ev) sz_record(L, ev[1], st);
    if (waits) cudaStreamWaitEvent(st, waits[1], 0);  // sz_step_host: everything else (overarea is accumulated by k_row_write)
    k_status<<<gf, TPB, 0, st>>>(S, B);
    k_fuse_propagate<<<1, 1024, 0, st>>>(S, B);
    k_row_count<<<gf, TPB, 0, st>>>(S, B);
    k_row_total<<<gf, TPB, 0, st>>>(S, B);
    scan_excl(L, S, B.row_count, B.row_off, &S.cnt->n_total, 0, S.cap_floes, B.scan_block, &S.cnt->n_rows);
    k_row_check<<<1, 1, 0, st>>>(S, B);
    k_row_write<<<sz_div_up(n_hint, 128), 128, 0, st>>>(S, B);
    k_update_boundaries<<<1, 1, 0, st>>>(S, P);
    if (ev) sz_record(L, ev[2], st);
    g_launch_count += 26;  // + the scans counted in scan_excl / scan_excl3
}

// ---- two-way coupling: sort the registry by (cell, floe), floe ∩ cell-box areas ------------------------------
__global__ void k_crec_zero(Store S, CouplingBuf CB, int ncell) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= ncell; c += gridDim.x * blockDim.x) {
        CB.cell_count[c] = 0;
        CB.cell_fill[c] = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        cnt->n_ccells = ncell;
        cnt->n_cbig = 0;
        if (cnt->n_crec > CB.cap_crec) cnt->n_crec = CB.cap_crec;
    }
}
__global__ void k_crec_count(Store S, CouplingBuf CB) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt->n_crec; r += gridDim.x * blockDim.x)
        atomicAdd(&CB.cell_count[CB.rec_cell[r]], 1);
}
__global__ void k_crec_fill(Store S, CouplingBuf CB) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < cnt->n_crec; r += gridDim.x * blockDim.x) {
        int c = CB.rec_cell[r];
        CB.perm[CB.cell_start[c] + atomicAdd(&CB.cell_fill[c], 1)] = r;
    }
}
// ascending floe index inside every cell: the order in which the reference's serial floe loop filled the cell
__global__ void k_crec_sort(Store S, CouplingBuf CB, int ncell) {
    if (S.cnt->error) return;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        int a = CB.cell_start[c], b = CB.cell_start[c + 1];
        for (int i = a + 1; i < b; ++i) {
            int v = CB.perm[i], fv = CB.rec_floe[v], j = i - 1;
            while (j >= a && CB.rec_floe[CB.perm[j]] > fv) {
                CB.perm[j + 1] = CB.perm[j];
                --j;
            }
            CB.perm[j + 1] = v;
        }
    }
}

// center_cell_coords + check_cell_bounds, coupling.jl:931-1140 (0-based cell index): xmin, xmax, ymin, ymax
__device__ __forceinline__ void center_cell_box(const Params &P, int ix, int iy, double b[4]) {
    double xmin = ((double)(ix + 1) - 1.5) * P.dx + P.x0, xmax = xmin + P.dx;
    double ymin = ((double)(iy + 1) - 1.5) * P.dy + P.y0, ymax = ymin + P.dy;
    if (!P.per_x) {
        xmin = xmin < P.x0 ? P.x0 : (xmin > P.xf ? P.xf : xmin);
        xmax = xmax > P.xf ? P.xf : (xmax < P.x0 ? P.x0 : xmax);
    }
    if (!P.per_y) {
        ymin = ymin < P.y0 ? P.y0 : (ymin > P.yf ? P.yf : ymin);
        ymax = ymax > P.yf ? P.yf : (ymax < P.y0 ? P.y0 : ymax);
    }
    b[0] = xmin; b[1] = xmax; b[2] = ymin; b[3] = ymax;
}

// floe_area_in_cell = sum(area.(intersect_polys(cell_poly, translated floe_poly))), coupling.jl:1652-1660.
// One thread per record (rings of <= 10 edges); larger rings go to the warp kernel below.
__global__ void __launch_bounds__(TN_NT, 2) k_crec_area(Store S, CouplingBuf CB, Params P) {
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const TSp sP = tsp(threadIdx.x), sQ = sP + TN_MAXV * TN_NT, sR = sQ + TN_MAXV * TN_NT;
    for (int r = blockIdx.x * TN_NT + threadIdx.x; r < cnt->n_crec; r += gridDim.x * TN_NT) {
        const int f = CB.rec_floe[r], cell = CB.rec_cell[r], nq = S.vcount[f];
        bool big = nq > TN_MAXV;
        double area = 0.0;
        if (!big) {
            double b[4];
            center_cell_box(P, cell % (P.Nx + 1), cell / (P.Nx + 1), b);
            sP[0 * TN_NT] = make_double2(b[0], b[2]);
            sP[1 * TN_NT] = make_double2(b[0], b[3]);
            sP[2 * TN_NT] = make_double2(b[1], b[3]);
            sP[3 * TN_NT] = make_double2(b[1], b[2]);
            sP[4 * TN_NT] = make_double2(b[0], b[2]);
            const double2 d = CB.rec_d[r];
            const double2 *gQ = S.verts + S.vstart[f];
            for (int k = 0; k < nq; ++k) {  // _translate_poly, floe_utils.jl:60-64
                double2 v = gQ[k];
                sQ[k * TN_NT] = make_double2(v.x + d.x, v.y + d.y);
            }
            const unsigned long long cr = t_clip<false>(tring(sP, 5), tring(sQ, nq), sR, TN_RCAP, tsp(TSP_NONE));
            if (TC_STATUS(cr) != TN_OK) big = true;
            else
                for (int g = 0; g < TC_NREG(cr); ++g) area += t_area(tring(sR + TC_RS(cr, g) * TN_NT, TC_RE(cr, g) - TC_RS(cr, g)));
        }
        if (big) CB.big_recs[atomicAdd(&cnt->n_cbig, 1)] = r;
        else CB.rec_area[r] = area;
    }
}

__global__ void k_crec_area_warp(Store S, CouplingBuf CB, Params P, int maxv, int maxx) {
    extern __shared__ __align__(16) unsigned char smem[];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    for (int it = blockIdx.x * wpb + wib; it < cnt->n_cbig; it += gridDim.x * wpb) {
        const int r = CB.big_recs[it], f = CB.rec_floe[r], cell = CB.rec_cell[r], nq = S.vcount[f];
        if (nq > w.maxv) {
            if (lane == 0) atomicOr(&cnt->error, ERR_POLY_TOO_LARGE);
            continue;
        }
        double b[4];
        center_cell_box(P, cell % (P.Nx + 1), cell / (P.Nx + 1), b);
        if (lane < 5) {
            double2 p;
            switch (lane) {
            case 1: p = make_double2(b[0], b[3]); break;
            case 2: p = make_double2(b[1], b[3]); break;
            case 3: p = make_double2(b[1], b[2]); break;
            default: p = make_double2(b[0], b[2]); break;
            }
            w.P[lane] = p;
        }
        const double2 d = CB.rec_d[r];
        const double2 *gQ = S.verts + S.vstart[f];
        for (int k = lane; k < nq; k += 32) w.Q[k] = make_double2(gQ[k].x + d.x, gQ[k].y + d.y);
        __syncwarp();
        int status;
        int nreg = warp_clip(w, w.P, 5, w.Q, nq, w.R1, w.rs1, w.re1, status);
        if (status == CLIP_OVERFLOW) {
            if (lane == 0) atomicOr(&cnt->error, ERR_POLY_TOO_LARGE);
            continue;
        }
        double area = 0.0;
        for (int g = 0; g < nreg; ++g) area += ring_area_seq(w.R1 + w.rs1[g], w.re1[g] - w.rs1[g]);
        if (lane == 0) CB.rec_area[r] = area;
        __syncwarp();
    }
}

void szk_cells_sort_and_clip(const Launch &L, const Store &S, const CouplingBuf &CB, const Params &P, int n_rec_hint) {
    cudaStream_t st = L.stream;
    const int ncell = (P.Nx + 1) * (P.Ny + 1);
    int gr = grid_for(L, n_rec_hint, TPB), gc = grid_for(L, ncell + 1, TPB);
    k_crec_zero<<<gc, TPB, 0, st>>>(S, CB, ncell);
    k_crec_count<<<gr, TPB, 0, st>>>(S, CB);
    scan_excl(L, S, CB.cell_count, CB.cell_start, &S.cnt->n_ccells, 0, CB.cap_cells, CB.scan_block, nullptr);
    k_crec_fill<<<gr, TPB, 0, st>>>(S, CB);
    k_crec_sort<<<gc, TPB, 0, st>>>(S, CB, ncell);
    k_crec_area<<<2 * L.sms, TN_NT, TN_SMEM_C, st>>>(S, CB, P);
    k_crec_area_warp<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), st>>>(S, CB, P, L.maxv_large, L.maxx_large);
    g_launch_count += 6;
}

// ---- geometry service / test hook -------------------------------------------------------------------------------------------
__global__ void k_debug_clip(const double2 *gP, int npp, const double2 *gQ, int nqp, int maxv, int maxx, int cap_regions,
                             int cap_points, int *out_offsets, double2 *out_xy, double *out_areas, int *out_n) {
    extern __shared__ __align__(16) unsigned char smem[];
    Ws w = ws_carve(smem, maxv, maxx);
    const int lane = lane_id();
    if (npp > maxv || nqp > maxv) {
        if (lane == 0) *out_n = SZ_ERR_CAPACITY;
        return;
    }
    for (int k = lane; k < npp; k += 32) w.P[k] = gP[k];
    for (int k = lane; k < nqp; k += 32) w.Q[k] = gQ[k];
    __syncwarp();
    int status;
    int nreg = warp_clip(w, w.P, npp, w.Q, nqp, w.R1, w.rs1, w.re1, status);
    if (lane == 0) {
        if (status == CLIP_OVERFLOW) *out_n = SZ_ERR_CAPACITY;
        else if (status == CLIP_FAIL) *out_n = -100;
        else {
            int o = 0, ok = nreg <= cap_regions;
            out_offsets[0] = 0;
            for (int r = 0; r < nreg && ok; ++r) {
                int a = w.rs1[r], b = w.re1[r];
                if (o + (b - a) > cap_points) { ok = 0; break; }
                for (int k = a; k < b; ++k) out_xy[o++] = w.R1[k];
                out_offsets[r + 1] = o;
                out_areas[r] = ring_area_seq(w.R1 + a, b - a);
            }
            *out_n = ok ? nreg : SZ_ERR_CAPACITY;
        }
    }
}

size_t szk_large_smem(int maxv, int maxx) { return ws_bytes(maxv, maxx); }

int szk_debug_clip(const Launch &L, const double *p_xy, int np, const double *q_xy, int nq, int cap_regions,
                   int cap_points, int *out_offsets, double *out_xy, double *out_areas) {
    double2 *dP = nullptr, *dQ = nullptr, *dxy = nullptr;
    double *dar = nullptr;
    int *doff = nullptr, *dn = nullptr;
    int rc = SZ_ERR_CUDA, n = 0;
    if (cudaMalloc(&dP, sizeof(double2) * np) != cudaSuccess) goto done;
    if (cudaMalloc(&dQ, sizeof(double2) * nq) != cudaSuccess) goto done;
    if (cudaMalloc(&dxy, sizeof(double2) * (cap_points + 1)) != cudaSuccess) goto done;
    if (cudaMalloc(&dar, sizeof(double) * (cap_regions + 1)) != cudaSuccess) goto done;
    if (cudaMalloc(&doff, sizeof(int) * (cap_regions + 2)) != cudaSuccess) goto done;
    if (cudaMalloc(&dn, sizeof(int)) != cudaSuccess) goto done;
    cudaMemcpyAsync(dP, p_xy, sizeof(double2) * np, cudaMemcpyHostToDevice, L.stream);
    cudaMemcpyAsync(dQ, q_xy, sizeof(double2) * nq, cudaMemcpyHostToDevice, L.stream);
    k_debug_clip<<<1, 32, ws_bytes(L.maxv_large, L.maxx_large), L.stream>>>(dP, np, dQ, nq, L.maxv_large, L.maxx_large,
                                                                           cap_regions, cap_points, doff, dxy, dar, dn);
    cudaMemcpyAsync(&n, dn, sizeof(int), cudaMemcpyDeviceToHost, L.stream);
    if (cudaStreamSynchronize(L.stream) != cudaSuccess) goto done;
    rc = n;
    if (n > 0) {
        cudaMemcpy(out_offsets, doff, sizeof(int) * (n + 1), cudaMemcpyDeviceToHost);
        cudaMemcpy(out_xy, dxy, sizeof(double2) * out_offsets[n], cudaMemcpyDeviceToHost);
        if (out_areas) cudaMemcpy(out_areas, dar, sizeof(double) * n, cudaMemcpyDeviceToHost);
    } else if (n == 0) {
        out_offsets[0] = 0;
    }
done:
    cudaFree(dP); cudaFree(dQ); cudaFree(dxy); cudaFree(dar); cudaFree(doff); cudaFree(dn);
    return rc;
}

// opt in to > 48 KB dynamic shared memory for the large-polygon kernels (called once per handle)
int szk_configure(const Launch &L) {
    size_t lb = ws_bytes(L.maxv_large, L.maxx_large);
    if (cudaFuncSetAttribute(k_narrow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_ghost_clip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_ghost_clip2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_debug_clip, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_narrow_ab<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM_A) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_crec_area, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM_C) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_crec_area_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_narrow_ab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM_B) != cudaSuccess) return -1;
    return 0;
}
