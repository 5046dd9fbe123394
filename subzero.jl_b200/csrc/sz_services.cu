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
    const TSp sP = tsp(threadIdx.x), sQ = sP + TN_MAXV * TN_NT, sR = sQ + TN_MAXV * TN_NT;
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
                const unsigned long long cr = t_clip<false>(tring(sP, np), tring(sQ, nq), sR, TN_RCAP, tsp(TSP_NONE));
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
    // topography (output.jl:826-829): free area of every cell (cell minus topography), 1 where something was cut out
    double *cell_free;
    unsigned char *cell_topo;
    int n_topo;
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
    const TSp sP = tsp(threadIdx.x), sQ = sP + TN_MAXV * TN_NT, sR = sQ + TN_MAXV * TN_NT;
    for (int r = blockIdx.x * TN_NT + threadIdx.x; r < B.n_rec; r += gridDim.x * TN_NT) {
        const int f = B.rec_floe[r], np = S.vcount[f];
        bool big = np > TN_MAXV || (G.n_topo > 0 && G.cell_topo[B.rec_cell[r]]);
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
            const unsigned long long cr = t_clip<false>(tring(sP, np), tring(sQ, 5), sR, TN_RCAP, tsp(TSP_NONE));
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
            if (G.n_topo > 0 && G.cell_topo[B.rec_cell[r]] && area > 0) {
                // area(floe ∩ (cell ∖ topo)) = area(floe ∩ cell) − Σ_k area((floe ∩ cell) ∩ topo_k): every region of clip #1
                // against every topography element that reaches the cell (same circle test, same order as the oracle)
                const double cell_rmax = sqrt(G.dx * G.dx + G.dy * G.dy);
                const double xc = b[0] + 0.5 * G.dx, yc = b[2] + 0.5 * G.dy;
                double sub = 0.0;
                for (int k = 0; k < G.n_topo; ++k) {
                    const double ddx = xc - S.topo_cx[k], ddy = yc - S.topo_cy[k];
                    if (!(sqrt(ddx * ddx + ddy * ddy) < S.topo_rmax[k] + cell_rmax)) continue;
                    const int nq = S.topo_vcount[k];
                    const double2 *gQ = S.topo_verts + S.topo_vstart[k];
                    __syncwarp();
                    for (int v = lane; v < nq; v += 32) w.Q[v] = gQ[v];
                    for (int g = 0; g < nreg; ++g) {
                        const int a0 = w.rs1[g], nr = w.re1[g] - a0;
                        if (nr > w.maxv) {
                            if (lane == 0) atomicOr(&S.cnt->error, ERR_POLY_TOO_LARGE);
                            continue;
                        }
                        __syncwarp();
                        for (int v = lane; v < nr; v += 32) w.P2[v] = w.R1[a0 + v];
                        __syncwarp();
                        int st2;
                        const int n2 = warp_clip(w, w.P2, nr, w.Q, nq, w.R2, w.rs2, w.re2, st2);
                        if (st2 == CLIP_OVERFLOW) {
                            if (lane == 0) atomicOr(&S.cnt->error, ERR_POLY_TOO_LARGE);
                            continue;
                        }
                        for (int g2 = 0; g2 < n2; ++g2) sub += ring_area_seq(w.R2 + w.rs2[g2], w.re2[g2] - w.rs2[g2]);
                    }
                }
                area = (area - sub > 1e-12 * area) ? area - sub : 0.0;
            }
        }
        if (lane == 0) B.rec_area[r] = area;
        __syncwarp();
    }
}

// free area of every cell: area(cell) − Σ_k area(cell ∩ topo_k) (output.jl:826-829 without a difference operator);
// cell_topo = 1 where something was cut out; a cell that is (all but 1e-12) covered gets free area 0 => all outputs 0
__global__ void k_eul_cell_topo(Store S, EulGrid G, int maxv, int maxx) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = lane_id(), wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
    Ws w = ws_carve(smem + (size_t)wib * ws_bytes(maxv, maxx), maxv, maxx);
    const int ncell = G.nx * G.ny;
    const double cell_rmax = sqrt(G.dx * G.dx + G.dy * G.dy);
    for (int c = blockIdx.x * wpb + wib; c < ncell; c += gridDim.x * wpb) {
        double b[4];
        eul_box(G, c, b);
        const double2 q0 = make_double2(b[0], b[2]), q1 = make_double2(b[0], b[3]), q2 = make_double2(b[1], b[3]), q3 = make_double2(b[1], b[2]);
        double a2 = 0.0;
        a2 += q0.x * q1.y - q0.y * q1.x;
        a2 += q1.x * q2.y - q1.y * q2.x;
        a2 += q2.x * q3.y - q2.y * q3.x;
        a2 += q3.x * q0.y - q3.y * q0.x;
        const double cell_area = fabs(a2 / 2.0);
        const double xc = b[0] + 0.5 * G.dx, yc = b[2] + 0.5 * G.dy;
        double cut = 0.0;
        for (int k = 0; k < G.n_topo; ++k) {
            const double ddx = xc - S.topo_cx[k], ddy = yc - S.topo_cy[k];
            if (!(sqrt(ddx * ddx + ddy * ddy) < S.topo_rmax[k] + cell_rmax)) continue;
            const int nq = S.topo_vcount[k];
            const double2 *gQ = S.topo_verts + S.topo_vstart[k];
            __syncwarp();
            if (lane == 0) { w.P[0] = q0; w.P[1] = q1; w.P[2] = q2; w.P[3] = q3; w.P[4] = q0; }
            for (int v = lane; v < nq; v += 32) w.Q[v] = gQ[v];
            __syncwarp();
            int status;
            const int nreg = warp_clip(w, w.P, 5, w.Q, nq, w.R1, w.rs1, w.re1, status);
            if (status == CLIP_OVERFLOW) {
                if (lane == 0) atomicOr(&S.cnt->error, ERR_POLY_TOO_LARGE);
                continue;
            }
            for (int g = 0; g < nreg; ++g) cut += ring_area_seq(w.R1 + w.rs1[g], w.re1[g] - w.rs1[g]);
        }
        if (lane == 0) {
            const bool any = cut > 0;
            double fr = cell_area - cut;
            if (any && !(fr > 1e-12 * cell_area)) fr = 0.0;
            G.cell_free[c] = fr;
            G.cell_topo[c] = any ? 1 : 0;
        }
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
        const bool cell_gone = G.n_topo > 0 && G.cell_topo[c] && !(G.cell_free[c] > 0.0);  // length(cell_poly_list) == 0, output.jl:831-834
        if (mass_tot > 0 && !cell_gone) {
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
            acc[SZ_GRID_SI_FRAC] = area_tot / (G.n_topo > 0 ? G.cell_free[c] : fabs(a2 / 2.0));
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
    EulGrid G = {nx, ny, n_floes, d_xg, d_yg, dx, dy, nullptr, nullptr, 0};
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
    EulGrid G = {A.nx, A.ny, A.n_floes, A.d_xg, A.d_yg, A.dx, A.dy, A.cell_free, A.cell_topo, A.n_topo};
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
    if (A.n_topo > 0) {
        k_eul_cell_topo<<<L.sms, 32, ws_bytes(L.maxv_large, L.maxx_large), st>>>(S, G, L.maxv_large, L.maxx_large);
        szk_count_launches(1);
    }
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
    if (cudaFuncSetAttribute(k_eul_cell_topo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lb) != cudaSuccess) return -1;
    return 0;
}

// ---- rank 4: sub-floe point generation (generate_subfloe_points, coupling.jl:172-208, :235-321) ----------------------
// One block per floe.  Two passes with the same arithmetic: COUNT (how many points, which Monte-Carlo attempt is
// accepted, status) and WRITE (the points, compacted in the reference's order with block-wide prefix sums), with a
// device scan of the counts in between, so the output is one contiguous CSR array.  The ring is read from global
// memory and translated by -centroid on the fly (_translate_poly: x + (-cx), the same bits as the reference).
#define PG_NT 128

__device__ __forceinline__ unsigned long long pg_sm64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// u(seed, floe id, attempt, draw, axis) in [0, 1): the counter-based generator of include/subzero_b200.h
__device__ __forceinline__ double pg_uniform(unsigned long long seed, long long id, int attempt, long long draw, int axis) {
    unsigned long long z = pg_sm64(seed ^ ((unsigned long long)id * 0x9E3779B97F4A7C15ull));
    z = pg_sm64(z ^ (((unsigned long long)attempt << 40) | (unsigned long long)draw));
    z = pg_sm64(z ^ (unsigned long long)axis);
    return (double)(z >> 11) * 0x1.0p-53;
}
// element i of range(a, b, length = n): endpoints exact, interior points in double-double, rounded once
__device__ __forceinline__ double pg_range_elem(double a, double b, long long i, long long n) {
    if (i == 0 || n < 2) return a;
    if (i == n - 1) return b;
    double dh = b - a, bb = dh - b, dl = (b - (dh - bb)) + (-a - bb);
    const double c = (double)i, m = (double)(n - 1);
    double ph = dh * c, pl = __fma_rn(dh, c, -ph) + dl * c;
    double s = ph + pl;
    pl = pl - (s - ph);
    ph = s;
    double qh = ph / m, th = qh * m, tl = __fma_rn(qh, m, -th);
    double ql = (((ph - th) - tl) + pl) / m;
    double sh = a + qh, t = sh - a, sl = (a - (sh - t)) + (qh - t);
    return sh + (sl + ql);
}
// GO.coveredby(point, ring) on the translated ring: interior or boundary (one thread, edges in order)
__device__ bool pg_coveredby(double2 p, const double2 *__restrict__ g, int n, double cx, double cy) {
    bool in = false;
    double2 a = make_double2(g[0].x + (-cx), g[0].y + (-cy));
    for (int k = 0; k + 1 < n; ++k) {
        const double2 b = make_double2(g[k + 1].x + (-cx), g[k + 1].y + (-cy));
        if (orient2d(a, b, p) == 0.0 && p.x >= fmin(a.x, b.x) && p.x <= fmax(a.x, b.x) && p.y >= fmin(a.y, b.y) && p.y <= fmax(a.y, b.y))
            return true;
        if ((a.y > p.y) != (b.y > p.y)) {
            const double xi = a.x + (p.y - a.y) / (b.y - a.y) * (b.x - a.x);
            if (p.x < xi) in = !in;
        }
        a = b;
    }
    return in;
}
// block-wide exclusive prefix sum of one int per thread (PG_NT threads); *total = sum over the block
__device__ int pg_block_scan(int v, int *total) {
    __shared__ int ws[PG_NT / 32], tot;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) ws[w] = x;
    __syncthreads();
    int off = 0;
    for (int k = 0; k < w; ++k) off += ws[k];
    if (threadIdx.x == PG_NT - 1) tot = off + x;
    __syncthreads();
    *total = tot;
    return off + x - v;
}

struct PgArgs {
    sz_points_generator g;
    const int *floes;   // 0-based indices, nullptr = identity
    int n;
    int *count, *attempt, *status;  // [n]
    const int *off;     // [n + 1] (write pass)
    double2 *out;
};

// edge i (1 <= i < np) of the sub-grid generator: points it emits (coupling.jl:248-288); WRITE: store them at dst
template <bool WRITE>
__device__ int pg_edge(const double2 *__restrict__ g, int i, double cx, double cy, double dg, double2 *dst) {
    double x1 = g[i - 1].x + (-cx), y1 = g[i - 1].y + (-cy), x2 = g[i].x + (-cx), y2 = g[i].y + (-cy);
    double dx = x2 - x1, dy = y2 - y1;
    double l = sqrt(dx * dx + dy * dy);
    int n = 1;
    if (WRITE) dst[0] = make_double2(x1, y1);
    if (l <= 2 * dg) {
        if (l > dg) {
            if (WRITE) dst[1] = make_double2(x1 + dx / 2, y1 + dy / 2);
            n = 2;
        }
        return n;
    }
    if (dx == 0) {
        const double sg = (double)((dy > 0) - (dy < 0));
        y1 += dg / 2 * sg;
        y2 -= dg / 2 * sg;
    } else if (dy == 0) {
        const double sg = (double)((dx > 0) - (dx < 0));
        x1 += dg / 2 * sg;
        x2 -= dg / 2 * sg;
    } else {  // the reference's x shift is positive whatever the direction of the edge
        const double m = dy / dx;
        const double xs = sqrt(dg * dg / (4 * (1 + m * m)));
        const double ys = m * xs;
        x1 += xs; x2 -= xs; y1 += ys; y2 -= ys;
    }
    l = sqrt((x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1));
    const long long ne = (long long)ceil(l / dg) + 1;
    if (WRITE)
        for (long long k = 0; k < ne; ++k) dst[1 + k] = make_double2(pg_range_elem(x1, x2, k, ne), pg_range_elem(y1, y2, k, ne));
    return 1 + (int)ne;
}

template <bool WRITE>
__global__ void __launch_bounds__(PG_NT) k_points(Store S, PgArgs A) {
    __shared__ double red[4][PG_NT / 32];
    __shared__ int s_in;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int q = blockIdx.x; q < A.n; q += gridDim.x) {
        const int f = A.floes ? A.floes[q] : q;
        const double2 *__restrict__ g = S.verts + S.vstart[f];
        const int np = S.vcount[f];
        const double cx = S.cx[f], cy = S.cy[f];
        // GI.extent of the translated ring
        double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
        for (int k = tid; k < np; k += PG_NT) {
            const double x = g[k].x + (-cx), y = g[k].y + (-cy);
            xmin = fmin(xmin, x); xmax = fmax(xmax, x); ymin = fmin(ymin, y); ymax = fmax(ymax, y);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
            xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = fmin(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
            ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        }
        __syncthreads();
        if (lane == 0) { red[0][w] = xmin; red[1][w] = xmax; red[2][w] = ymin; red[3][w] = ymax; }
        __syncthreads();
        for (int k = 0; k < PG_NT / 32; ++k) {
            xmin = fmin(xmin, red[0][k]); xmax = fmax(xmax, red[1][k]); ymin = fmin(ymin, red[2][k]); ymax = fmax(ymax, red[3][k]);
        }
        const double Dx = xmax - xmin, Dy = ymax - ymin;
        double2 *dst = WRITE ? A.out + A.off[q] : nullptr;
        if (A.g.kind == SZ_POINTS_MONTE_CARLO) {
            const int n = A.g.npoints;
            const long long id = S.id[f];
            if (!WRITE) {
                const double area = S.area[f];
                int count = 1, used = 0, kept = 0, status = SZ_STATUS_ACTIVE;
                double err = 1.0;
                while (err > A.g.err) {  // warp- and block-uniform
                    if (count > 10) {
                        err = 0.0;
                        status = SZ_STATUS_REMOVE;
                    } else {
                        int in = 0;
                        for (int j = tid; j < n; j += PG_NT) {
                            const double2 p = make_double2(xmin + Dx * pg_uniform(A.g.seed, id, count, j, 0),
                                                           ymin + Dy * pg_uniform(A.g.seed, id, count, j, 1));
                            in += pg_coveredby(p, g, np, cx, cy) ? 1 : 0;
                        }
                        int tot;
                        pg_block_scan(in, &tot);
                        kept = tot;
                        err = fabs((double)tot / (double)n * (Dx * Dy) - area) / area;
                        used = count;
                        count += 1;
                    }
                }
                if (kept == 0) status = SZ_STATUS_REMOVE;
                if (tid == 0) { A.count[q] = kept; A.attempt[q] = used; A.status[q] = status; }
            } else {
                const int used = A.attempt[q];
                int base = 0;
                for (int j0 = 0; used > 0 && j0 < n; j0 += PG_NT) {  // draw order is kept: chunk by chunk, scan inside a chunk
                    const int j = j0 + tid;
                    double2 p = make_double2(0.0, 0.0);
                    bool in = false;
                    if (j < n) {
                        p = make_double2(xmin + Dx * pg_uniform(A.g.seed, id, used, j, 0), ymin + Dy * pg_uniform(A.g.seed, id, used, j, 1));
                        in = pg_coveredby(p, g, np, cx, cy);
                    }
                    int tot;
                    const int pos = pg_block_scan(in ? 1 : 0, &tot);
                    if (in) dst[base + pos] = p;
                    base += tot;
                }
            }
        } else {
            const double dg = A.g.delta_g;
            int base = 0;
            // every vertex + the points on its edge, edges in ring order
            for (int i0 = 1; i0 < np; i0 += PG_NT) {
                const int i = i0 + tid;
                const int ne = i < np ? pg_edge<false>(g, i, cx, cy, dg, nullptr) : 0;
                int tot;
                const int pos = pg_block_scan(ne, &tot);
                if (WRITE && i < np) pg_edge<true>(g, i, cx, cy, dg, dst + base + pos);
                base += tot;
            }
            // interior lattice (:289-318)
            long long nx = (long long)ceil((xmax - xmin) / dg), ny = (long long)ceil((ymax - ymin) / dg);
            const bool xs1 = nx < 3, ys1 = ny < 3;
            if (xs1) nx = 1;
            if (ys1) ny = 1;
            const double xa = xmin + dg / 2, xb = xmax - dg / 2, ya = ymin + dg / 2, yb = ymax - dg / 2;
            for (long long k0 = 0; k0 < nx * ny; k0 += PG_NT) {
                const long long k = k0 + tid;
                double2 p = make_double2(0.0, 0.0);
                bool in = false;
                if (k < nx * ny) {
                    p = make_double2(xs1 ? 0.0 : pg_range_elem(xa, xb, k % nx, nx), ys1 ? 0.0 : pg_range_elem(ya, yb, k / nx, ny));
                    in = pg_coveredby(p, g, np, cx, cy);
                }
                int tot;
                const int pos = pg_block_scan(in ? 1 : 0, &tot);
                if (WRITE && in) dst[base + pos] = p;
                base += tot;
            }
            if (!WRITE && tid == 0) { A.count[q] = base; A.attempt[q] = 0; A.status[q] = SZ_STATUS_ACTIVE; }
        }
        __syncthreads();
        (void)s_in;
    }
}

void szk_points_count(const Launch &L, const Store &S, const sz_points_generator &g, const int *floes, int n, int *count, int *attempt,
                      int *status) {
    if (n <= 0) return;
    PgArgs A = {g, floes, n, count, attempt, status, nullptr, nullptr};
    k_points<false><<<sv_grid(L, n, 1), PG_NT, 0, L.stream>>>(S, A);
    szk_count_launches(1);
}
void szk_points_write(const Launch &L, const Store &S, const sz_points_generator &g, const int *floes, int n, int *attempt, const int *off,
                      double2 *out) {
    if (n <= 0) return;
    PgArgs A = {g, floes, n, nullptr, attempt, nullptr, off, out};
    k_points<true><<<sv_grid(L, n, 1), PG_NT, 0, L.stream>>>(S, A);
    szk_count_launches(1);
}
