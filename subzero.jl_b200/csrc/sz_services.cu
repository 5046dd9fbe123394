// sz_services.cu — device services for the HOST-side processes of the reference.
//
//   rank 2  batched overlap query: potential_interaction + sum(GO.area, intersect_polys(poly_i, poly_j)) for a
//           caller-given pair list — the pair test of smooth_floes! (simplification.jl:98-116), timestep_welding!
//           (welding.jl:119-150) and the ridge/raft validity test (ridge_raft.jl:706-753)
//   rank 3  calc_eulerian_data! (output.jl:794-919): floe data averaged on the cells of a GridOutputWriter
//
// Both reuse the clippers of the narrow phase (thread per item for rings of <= 10 edges, warp per item for the
// rest), compiled with -fmad=false like sz_kernels.cu: areas are bit-identical to the oracle's, so a host decision
// such as `intersect_area / area_j > floe_floe_max_overlap` cannot flip.
#include <cub/device/device_radix_sort.cuh>

#include "sz_common.cuh"

// the clippers are header-defined __device__ functions without `static`: give this translation unit its own
// copies (the library is built without relocatable device code, so nothing is shared across TUs anyway)
namespace {
#include "sz_narrow_thread.cuh"
}

void szk_count_launches(int n);

__device__ __forceinline__ bool sv_potential_interaction(double xi, double yi, double ri, double xj, double yj, double rj) {
    double dx = xi - xj, dy = yi - yj, rr = ri + rj;  // collisions.jl:705-710
    return dx * dx + dy * dy < rr * rr;
}

static inline int sv_grid(const Launch &L, long long items, int per_block) {
    long long b = (items + per_block - 1) / per_block, cap = (long long)L.sms * 32;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---- rank 2: pair overlap areas ------------------------------------------------------------------------------
struct PairQuery {
    const int2 *pairs;  // 0-based (i, j)
    int n;
    double *area;
    unsigned char *inter;
    int *big, *n_big;  // pairs the warp kernel must take
};

__global__ void __launch_bounds__(TN_NT, 2) k_pair_area(Store S, PairQuery Q) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *base = (double2 *)smem + threadIdx.x;
    double2 *sP = base, *sQ = sP + TN_MAXV * TN_NT, *sR = sQ + TN_MAXV * TN_NT;
    for (int k = blockIdx.x * TN_NT + threadIdx.x; k < Q.n; k += gridDim.x * TN_NT) {
        const int2 pr = Q.pairs[k];
        const bool pot = sv_potential_interaction(S.cx[pr.x], S.cy[pr.x], S.rmax[pr.x], S.cx[pr.y], S.cy[pr.y], S.rmax[pr.y]);
        if (Q.inter) Q.inter[k] = pot ? 1 : 0;
        double area = 0.0;
        bool big = false;
        if (pot) {
            const int np = S.vcount[pr.x], nq = S.vcount[pr.y];
            big = np > TN_MAXV || nq > TN_MAXV;
            if (!big) {
                const double2 *gP = S.verts + S.vstart[pr.x], *gQ = S.verts + S.vstart[pr.y];
                for (int v = 0; v < np; ++v) sP[v * TN_NT] = gP[v];
                for (int v = 0; v < nq; ++v) sQ[v * TN_NT] = gQ[v];
                const unsigned long long cr = t_clip<false>(tring(sP, np), tring(sQ, nq), sR, TN_RCAP, nullptr);
                if (TC_STATUS(cr) != TN_OK) big = true;
                else
                    for (int g = 0; g < TC_NREG(cr); ++g) area += t_area(tring(sR + TC_RS(cr, g) * TN_NT, TC_RE(cr, g) - TC_RS(cr, g)));
            }
        }
        if (big) Q.big[atomicAdd(Q.n_big, 1)] = k;
        else Q.area[k] = area;
    }
}

__global__ void k_pair_area_warp(Store S, PairQuery Q, int maxv, int maxx) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    const int nb = *Q.n_big;
    for (int it = blockIdx.x * wpb + wib; it < nb; it += gridDim.x * wpb) {
        const int k = Q.big[it];
        const int2 pr = Q.pairs[k];
        const int np = S.vcount[pr.x], nq = S.vcount[pr.y];
        const double2 *gP = S.verts + S.vstart[pr.x], *gQ = S.verts + S.vstart[pr.y];
        for (int v = lane; v < np; v += 32) w.P[v] = gP[v];
        for (int v = lane; v < nq; v += 32) w.Q[v] = gQ[v];
        __syncwarp();
        int status;
        int nreg = warp_clip(w, w.P, np, w.Q, nq, w.R1, w.rs1, w.re1, status);
        double area = 0.0;
        if (status == CLIP_OVERFLOW) {
            if (lane == 0) atomicOr(&S.cnt->error, ERR_POLY_TOO_LARGE);
        } else {
            for (int g = 0; g < nreg; ++g) area += ring_area_seq(w.R1 + w.rs1[g], w.re1[g] - w.rs1[g]);
        }
        if (lane == 0) Q.area[k] = area;
        __syncwarp();
    }
}

// pairs: device [n] int2 0-based; area [n], inter [n] or null; scratch: big [n], n_big [1] (zeroed here)
void szk_pair_areas(const Launch &L, const Store &S, const int2 *pairs, int n, double *area, unsigned char *inter, int *big,
                    int *n_big) {
    PairQuery Q = {pairs, n, area, inter, big, n_big};
    cudaMemsetAsync(n_big, 0, sizeof(int), L.stream);
    k_pair_area<<<2 * L.sms, TN_NT, TN_SMEM_C, L.stream>>>(S, Q);
    k_pair_area_warp<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), L.stream>>>(S, Q, L.maxv_large, L.maxx_large);
    szk_count_launches(2);
}

// ---- rank 3: Eulerian gridded output ---------------------------------------------------------------------------
struct EulGrid {
    int nx, ny, n_floes;
    const double *xg, *yg;  // device copies of the writer's grid lines
    double dx, dy;
};
struct EulBuf {
    int *rec_count, *rec_off;  // [n_floes + 1]
    int n_rec;                 // host-known after the count pass
    int *rec_floe, *rec_cell;  // [n_rec] in floe order
    double *rec_area;
    unsigned long long *key_in, *key_out;  // (cell << 32) | floe
    int *val_in, *val_out;                 // record indices, sorted by (cell, floe) in val_out
    int *cell_start;                       // [ncell + 1] into val_out
    int *big, *n_big;
};

// the cells whose box overlaps the floe's bounding box: every other cell has pic_area == 0 and is dropped by
// `floeidx[pic_area .> 0]` (output.jl:848); the reference's circle mask (:808-818) is a superset of these
template <bool WRITE>
__global__ void k_eul_records(Store S, EulGrid G, EulBuf B) {
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < G.n_floes; f += gridDim.x * blockDim.x) {
        const double2 *r = S.verts + S.vstart[f];
        const int nv = S.vcount[f];
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int v = 0; v < nv; ++v) {
            double2 p = r[v];
            xmin = fmin(xmin, p.x); xmax = fmax(xmax, p.x);
            ymin = fmin(ymin, p.y); ymax = fmax(ymax, p.y);
        }
        // conservative index window from the (uniform) spacing, then the exact test against the stored lines
        int j0 = (int)fmax(0.0, floor((xmin - G.xg[0]) / G.dx) - 1.0), j1 = (int)fmin((double)(G.nx - 1), floor((xmax - G.xg[0]) / G.dx) + 1.0);
        int i0 = (int)fmax(0.0, floor((ymin - G.yg[0]) / G.dy) - 1.0), i1 = (int)fmin((double)(G.ny - 1), floor((ymax - G.yg[0]) / G.dy) + 1.0);
        int c = 0, o = WRITE ? B.rec_off[f] : 0;
        for (int i = i0; i <= i1; ++i) {
            if (!(ymax > G.yg[i] && ymin < G.yg[i + 1])) continue;
            for (int j = j0; j <= j1; ++j) {
                if (!(xmax > G.xg[j] && xmin < G.xg[j + 1])) continue;
                if (WRITE) {
                    const int cell = j + G.nx * i;
                    B.rec_floe[o + c] = f;
                    B.rec_cell[o + c] = cell;
                    B.key_in[o + c] = ((unsigned long long)cell << 32) | (unsigned)f;
                    B.val_in[o + c] = o + c;
                }
                c++;
            }
        }
        if (!WRITE) B.rec_count[f] = c;
    }
}

// one block: exclusive scan of rec_count (n_floes is small next to the clipping work)
__global__ void __launch_bounds__(1024) k_eul_scan(EulGrid G, EulBuf B) {
    __shared__ int part[1024];
    const int n = G.n_floes, per = (n + 1023) / 1024, a = threadIdx.x * per, b = min(n, a + per);
    int s = 0;
    for (int k = a; k < b; ++k) s += B.rec_count[k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int k = 0; k < 1024; ++k) {
            int v = part[k];
            part[k] = run;
            run += v;
        }
        B.rec_off[n] = run;
    }
    __syncthreads();
    int run = part[threadIdx.x];
    for (int k = a; k < b; ++k) {
        B.rec_off[k] = run;
        run += B.rec_count[k];
    }
}

__device__ __forceinline__ void eul_box(const EulGrid &G, int cell, double b[4]) {
    const int j = cell % G.nx, i = cell / G.nx;
    b[0] = G.xg[j]; b[1] = G.xg[j + 1]; b[2] = G.yg[i]; b[3] = G.yg[i + 1];
}

// pic_area = sum(GO.area, intersect_polys(floe_poly, cell_poly)), output.jl:845
__global__ void __launch_bounds__(TN_NT, 2) k_eul_area(Store S, EulGrid G, EulBuf B) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2 *base = (double2 *)smem + threadIdx.x;
    double2 *sP = base, *sQ = sP + TN_MAXV * TN_NT, *sR = sQ + TN_MAXV * TN_NT;
    for (int r = blockIdx.x * TN_NT + threadIdx.x; r < B.n_rec; r += gridDim.x * TN_NT) {
        const int f = B.rec_floe[r], np = S.vcount[f];
        bool big = np > TN_MAXV;
        double area = 0.0;
        if (!big) {
            double b[4];
            eul_box(G, B.rec_cell[r], b);
            const double2 *gP = S.verts + S.vstart[f];
            for (int v = 0; v < np; ++v) sP[v * TN_NT] = gP[v];
            sQ[0 * TN_NT] = make_double2(b[0], b[2]);  // _make_bounding_box_polygon, floe_utils.jl:104-108
            sQ[1 * TN_NT] = make_double2(b[0], b[3]);
            sQ[2 * TN_NT] = make_double2(b[1], b[3]);
            sQ[3 * TN_NT] = make_double2(b[1], b[2]);
            sQ[4 * TN_NT] = make_double2(b[0], b[2]);
            const unsigned long long cr = t_clip<false>(tring(sP, np), tring(sQ, 5), sR, TN_RCAP, nullptr);
            if (TC_STATUS(cr) != TN_OK) big = true;
            else
                for (int g = 0; g < TC_NREG(cr); ++g) area += t_area(tring(sR + TC_RS(cr, g) * TN_NT, TC_RE(cr, g) - TC_RS(cr, g)));
        }
        if (big) B.big[atomicAdd(B.n_big, 1)] = r;
        else B.rec_area[r] = area;
    }
}

__global__ void k_eul_area_warp(Store S, EulGrid G, EulBuf B, int maxv, int maxx) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    const int nb = *B.n_big;
    for (int it = blockIdx.x * wpb + wib; it < nb; it += gridDim.x * wpb) {
        const int r = B.big[it], f = B.rec_floe[r], np = S.vcount[f];
        double b[4];
        eul_box(G, B.rec_cell[r], b);
        const double2 *gP = S.verts + S.vstart[f];
        for (int v = lane; v < np; v += 32) w.P[v] = gP[v];
        if (lane < 5) {
            double2 p;
            switch (lane) {
            case 1: p = make_double2(b[0], b[3]); break;
            case 2: p = make_double2(b[1], b[3]); break;
            case 3: p = make_double2(b[1], b[2]); break;
            default: p = make_double2(b[0], b[2]); break;
            }
            w.Q[lane] = p;
        }
        __syncwarp();
        int status;
        int nreg = warp_clip(w, w.P, np, w.Q, 5, w.R1, w.rs1, w.re1, status);
        double area = 0.0;
        if (status == CLIP_OVERFLOW) {
            if (lane == 0) atomicOr(&S.cnt->error, ERR_POLY_TOO_LARGE);
        } else {
            for (int g = 0; g < nreg; ++g) area += ring_area_seq(w.R1 + w.rs1[g], w.re1[g] - w.rs1[g]);
        }
        if (lane == 0) B.rec_area[r] = area;
        __syncwarp();
    }
}

// first sorted record of every cell (binary search on the sorted keys)
__global__ void k_eul_cell_start(EulGrid G, EulBuf B) {
    const int ncell = G.nx * G.ny;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c <= ncell; c += gridDim.x * blockDim.x) {
        const unsigned long long key = (unsigned long long)c << 32;
        int lo = 0, hi = B.n_rec;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (B.key_out[mid] < key) lo = mid + 1;
            else hi = mid;
        }
        B.cell_start[c] = lo;
    }
}

struct EulOut {
    int n_out;
    int kinds[SZ_GRID_NKINDS];
    double *data;  // [nx][ny][n_out], data[j + nx (i + ny k)]
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULLMASK, v, o);
    return v;
}

// One warp per cell; the records of a cell are in floe order, lanes take them round-robin and the partial sums are
// combined with a fixed shuffle tree: the result does not depend on scheduling (output.jl:848-905).
__global__ void k_eul_cells(Store S, EulGrid G, EulBuf B, EulOut O) {
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    const int ncell = G.nx * G.ny;
    for (int c = blockIdx.x * wpb + wib; c < ncell; c += gridDim.x * wpb) {
        const int a = B.cell_start[c], b = B.cell_start[c + 1];
        double area_tot = 0.0, mass_tot = 0.0, over = 0.0;
        int cnt = 0;
        for (int q = a + lane; q < b; q += 32) {
            const int r = B.val_out[q];
            const double pic = B.rec_area[r];
            if (pic > 0) {
                const int f = B.rec_floe[r];
                area_tot += pic;
                mass_tot += S.mass[f] * (pic / S.area[f]);
                over += S.overarea[f];
                cnt++;
            }
        }
        area_tot = warp_sum(area_tot);
        mass_tot = warp_sum(mass_tot);
        over = warp_sum(over);
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULLMASK, cnt, o);
        double acc[SZ_GRID_NKINDS];
#pragma unroll
        for (int k = 0; k < SZ_GRID_NKINDS; ++k) acc[k] = 0.0;
        if (mass_tot > 0) {
            for (int q = a + lane; q < b; q += 32) {
                const int r = B.val_out[q];
                const double pic = B.rec_area[r];
                if (pic > 0) {
                    const int f = B.rec_floe[r];
                    const double ma = (pic / S.area[f]) * (S.mass[f] / mass_tot);
                    const double *sa = S.stress_accum + 4 * (size_t)f, *st = S.strain + 4 * (size_t)f;
                    acc[SZ_GRID_U] += S.u[f] * ma;
                    acc[SZ_GRID_V] += S.v[f] * ma;
                    acc[SZ_GRID_DUDT] += S.p_dudt[f] * ma;
                    acc[SZ_GRID_DVDT] += S.p_dvdt[f] * ma;
                    acc[SZ_GRID_HEIGHT] += S.height[f] * ma;
                    acc[SZ_GRID_STRESS_XX] += sa[0] * ma;
                    acc[SZ_GRID_STRESS_YX] += sa[2] * ma;
                    acc[SZ_GRID_STRESS_XY] += sa[1] * ma;
                    acc[SZ_GRID_STRESS_YY] += sa[3] * ma;
                    acc[SZ_GRID_STRAIN_UX] += st[0] * ma;
                    acc[SZ_GRID_STRAIN_VX] += st[2] * ma;
                    acc[SZ_GRID_STRAIN_UY] += st[1] * ma;
                    acc[SZ_GRID_STRAIN_VY] += st[3] * ma;
                }
            }
#pragma unroll
            for (int k = 0; k < SZ_GRID_NKINDS; ++k) acc[k] = warp_sum(acc[k]);
            double bx[4];
            eul_box(G, c, bx);
            // GO.area of the cell ring (xmin,ymin),(xmin,ymax),(xmax,ymax),(xmax,ymin): the shoelace sum, not dx*dy
            const double2 q0 = make_double2(bx[0], bx[2]), q1 = make_double2(bx[0], bx[3]), q2 = make_double2(bx[1], bx[3]),
                          q3 = make_double2(bx[1], bx[2]);
            double a2 = 0.0;
            a2 += q0.x * q1.y - q0.y * q1.x;
            a2 += q1.x * q2.y - q1.y * q2.x;
            a2 += q2.x * q3.y - q2.y * q3.x;
            a2 += q3.x * q0.y - q3.y * q0.x;
            acc[SZ_GRID_SI_FRAC] = area_tot / fabs(a2 / 2.0);
            acc[SZ_GRID_OVERAREA] = over / (double)cnt;
            acc[SZ_GRID_MASS] = mass_tot;
            acc[SZ_GRID_AREA] = area_tot;
            const double xx = acc[SZ_GRID_STRESS_XX], yx = acc[SZ_GRID_STRESS_YX], xy = acc[SZ_GRID_STRESS_XY], yy = acc[SZ_GRID_STRESS_YY];
            const double hm = 0.5 * (xx + yy), hd = 0.5 * (xx - yy), disc = hd * hd + yx * xy;
            double e = disc > 0 ? hm + sqrt(disc) : hm;  // maximum(eigvals([xx yx; xy yy]))
            if (fabs(e) > 1e8) e = 0.0;
            acc[SZ_GRID_STRESS_EIG] = e;
        }
        if (lane == 0) {
            const int j = c % G.nx, i = c / G.nx;
            for (int k = 0; k < O.n_out; ++k) {
                double v = 0.0;
#pragma unroll
                for (int m = 0; m < SZ_GRID_NKINDS; ++m)
                    if (O.kinds[k] == m) v = acc[m];
                O.data[(size_t)j + (size_t)G.nx * ((size_t)i + (size_t)G.ny * (size_t)k)] = v;
            }
        }
    }
}

// ---- host drivers (called from sz_api.cu) ------------------------------------------------------------------------
// pass 1: count the (floe, cell) records; *n_rec is read back by the caller after a synchronisation
void szk_eul_count(const Launch &L, const Store &S, int n_floes, int nx, int ny, const double *d_xg, const double *d_yg, double dx,
                   double dy, int *rec_count, int *rec_off) {
    EulGrid G = {nx, ny, n_floes, d_xg, d_yg, dx, dy};
    EulBuf B = {};
    B.rec_count = rec_count;
    B.rec_off = rec_off;
    k_eul_records<false><<<sv_grid(L, n_floes, 128), 128, 0, L.stream>>>(S, G, B);
    k_eul_scan<<<1, 1024, 0, L.stream>>>(G, B);
    szk_count_launches(2);
}

size_t szk_eul_sort_bytes(int n_rec) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long *)nullptr, (unsigned long long *)nullptr,
                                    (const int *)nullptr, (int *)nullptr, n_rec);
    return bytes;
}

// pass 2: records, areas, (cell, floe) sort, per-cell reduction
int szk_eul_run(const Launch &L, const Store &S, const SzkEulArgs &A) {
    EulGrid G = {A.nx, A.ny, A.n_floes, A.d_xg, A.d_yg, A.dx, A.dy};
    EulBuf B = {};
    B.rec_count = A.rec_count; B.rec_off = A.rec_off; B.n_rec = A.n_rec; B.rec_floe = A.rec_floe; B.rec_cell = A.rec_cell;
    B.rec_area = A.rec_area; B.key_in = A.key_in; B.key_out = A.key_out; B.val_in = A.val_in; B.val_out = A.val_out;
    B.cell_start = A.cell_start; B.big = A.big; B.n_big = A.n_big;
    EulOut O;
    O.n_out = A.n_out;
    for (int k = 0; k < SZ_GRID_NKINDS; ++k) O.kinds[k] = k < A.n_out ? A.kinds[k] : -1;
    O.data = A.d_data;
    cudaStream_t st = L.stream;
    const int ncell = A.nx * A.ny;
    if (A.n_rec > 0) {
        cudaMemsetAsync(A.n_big, 0, sizeof(int), st);
        k_eul_records<true><<<sv_grid(L, A.n_floes, 128), 128, 0, st>>>(S, G, B);
        k_eul_area<<<2 * L.sms, TN_NT, TN_SMEM_C, st>>>(S, G, B);
        k_eul_area_warp<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), st>>>(S, G, B, L.maxv_large, L.maxx_large);
        size_t bytes = A.sort_bytes;
        int bits = 32;
        for (int c = ncell; c > 0; c >>= 1) bits++;
        if (cub::DeviceRadixSort::SortPairs(A.sort_tmp, bytes, A.key_in, A.key_out, A.val_in, A.val_out, A.n_rec, 0, bits > 64 ? 64 : bits,
                                            st) != cudaSuccess)
            return -1;
        szk_count_launches(6);
    }
    k_eul_cell_start<<<sv_grid(L, ncell + 1, 256), 256, 0, st>>>(G, B);
    k_eul_cells<<<sv_grid(L, ncell, 4), 128, 0, st>>>(S, G, B, O);
    szk_count_launches(2);
    return 0;
}

int szk_services_configure(const Launch &L) {
    size_t lb = ws_bytes(L.maxv_large, L.maxx_large);
    if (cudaFuncSetAttribute(k_pair_area, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM_C) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_eul_area, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TN_SMEM_C) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_pair_area_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(k_eul_area_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    return 0;
}
