// sz_geom.cuh — warp-cooperative FP64 polygon geometry for the narrow phase (K3/K4) and the
// ghost trigger (K8).  One warp owns one (polygon, polygon) work item; both rings are staged in
// shared memory; edge-pair predicates are evaluated one pair per lane and compacted with
// ballots; reductions whose result feeds a DISCRETE decision are either order-independent
// (min, any, parity) or evaluated in one canonical sequential order, so the product makes the
// same decisions as the definition the parity tests check against (intersect_polys /
// GeometryOps semantics, floe_utils.jl:55).  Every translation
// unit that includes this file is compiled with -fmad=false: no contraction of a*b+c, matching
// Julia's unfused arithmetic.
//
// Degeneracies (vertex on an edge, collinear overlapping edges) are resolved by a symbolic
// perturbation: Q is treated as translated by the infinitesimal vector (eps, eps^2).  A rigid
// motion keeps every side decision mutually consistent; an orientation that evaluates to
// exactly 0 takes the sign of its first non-vanishing eps term.
#pragma once
#include "sz_common.cuh"

// keep the warp kernels small: see the instruction-cache note in sz_narrow_thread.cuh
#ifndef SZ_ROLLED
#define SZ_ROLLED _Pragma("unroll 1")
#endif

#define FULLMASK 0xffffffffu

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

__device__ __forceinline__ double orient2d(double2 a, double2 b, double2 c) {
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}
// side of a P point w.r.t. the perturbed Q edge c->d (true = left)
__device__ __forceinline__ bool side_q(double o, double2 c, double2 d) {
    if (o != 0.0) return o > 0.0;
    if (d.y != c.y) return d.y > c.y;
    return d.x <= c.x;
}
// side of a perturbed Q point w.r.t. the P edge a->b (true = left)
__device__ __forceinline__ bool side_p(double o, double2 a, double2 b) {
    if (o != 0.0) return o > 0.0;
    if (b.y != a.y) return b.y < a.y;
    return b.x >= a.x;
}

// Twice the signed shoelace area, canonical sequential order (every lane computes the same).
__device__ __forceinline__ double ring_area2_seq(const double2 *r, int n) {
    double a = 0.0;
    SZ_ROLLED
    for (int k = 0; k + 1 < n; ++k) a += r[k].x * r[k + 1].y - r[k].y * r[k + 1].x;
    return a;
}
__device__ __forceinline__ double ring_area_seq(const double2 *r, int n) { return fabs(ring_area2_seq(r, n) / 2.0); }
__device__ __forceinline__ double2 ring_centroid_seq(const double2 *r, int n) {
    double a = 0.0, cx = 0.0, cy = 0.0;
    SZ_ROLLED
    for (int k = 0; k + 1 < n; ++k) {
        double c = r[k].x * r[k + 1].y - r[k].y * r[k + 1].x;
        a += c;
        cx += (r[k].x + r[k + 1].x) * c;
        cy += (r[k].y + r[k + 1].y) * c;
    }
    a /= 2.0;
    return make_double2(cx / (6.0 * a), cy / (6.0 * a));
}
// orientation only (sign of the shoelace sum): lane-strided partial sums are enough
__device__ __forceinline__ bool warp_ring_is_ccw(const double2 *r, int n) {
    double a = 0.0;
    SZ_ROLLED
    for (int k = lane_id(); k + 1 < n; k += 32) a += r[k].x * r[k + 1].y - r[k].y * r[k + 1].x;
#pragma unroll
    for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(FULLMASK, a, o);
    return a > 0.0;
}

__device__ __forceinline__ double point_segment_distance(double2 p, double2 a, double2 b) {
    double dx = b.x - a.x, dy = b.y - a.y;
    double l2 = dx * dx + dy * dy;
    double t = 0.0;
    if (l2 > 0.0) {
        t = ((p.x - a.x) * dx + (p.y - a.y) * dy) / l2;
        if (t < 0.0) t = 0.0;
        if (t > 1.0) t = 1.0;
    }
    double qx = a.x + t * dx - p.x, qy = a.y + t * dy - p.y;
    return sqrt(qx * qx + qy * qy);
}

// Closed-segment intersection (endpoints and collinear overlaps count): 0, 1 or 2 points.
__device__ __forceinline__ int segment_intersection(double2 a, double2 b, double2 c, double2 d, double2 &p0,
                                                    double2 &p1) {
    double o1 = orient2d(c, d, a), o2 = orient2d(c, d, b);
    double o3 = orient2d(a, b, c), o4 = orient2d(a, b, d);
    if (o1 == 0.0 && o2 == 0.0) {
        bool usex = fabs(b.x - a.x) >= fabs(b.y - a.y);
        double a0 = usex ? a.x : a.y, a1 = usex ? b.x : b.y;
        double c0 = usex ? c.x : c.y, c1 = usex ? d.x : d.y;
        double2 lo1 = a0 <= a1 ? a : b, hi1 = a0 <= a1 ? b : a;
        double2 lo2 = c0 <= c1 ? c : d, hi2 = c0 <= c1 ? d : c;
        double l1 = fmin(a0, a1), h1 = fmax(a0, a1), l2 = fmin(c0, c1), h2 = fmax(c0, c1);
        double2 lo = l1 >= l2 ? lo1 : lo2, hi = h1 <= h2 ? hi1 : hi2;
        double lv = fmax(l1, l2), hv = fmin(h1, h2);
        if (lv > hv) return 0;
        p0 = lo;
        if (lv == hv) return 1;
        p1 = hi;
        return 2;
    }
    if ((o1 > 0.0 && o2 > 0.0) || (o1 < 0.0 && o2 < 0.0)) return 0;
    if ((o3 > 0.0 && o4 > 0.0) || (o3 < 0.0 && o4 < 0.0)) return 0;
    if (o3 == 0.0 && o4 == 0.0) return 0;
    if (o1 == 0.0) p0 = a;
    else if (o2 == 0.0) p0 = b;
    else if (o3 == 0.0) p0 = c;
    else if (o4 == 0.0) p0 = d;
    else {
        double t = o1 / (o1 - o2);
        p0.x = a.x + t * (b.x - a.x);
        p0.y = a.y + t * (b.y - a.y);
    }
    return 1;
}

// ---- warp-parallel point predicates (order-independent reductions) -------------------------
// |GO.signed_distance(point, ring)|, collisions.jl:91
__device__ __forceinline__ double warp_point_ring_distance(double2 p, const double2 *r, int n) {
    double best = INFINITY;
    SZ_ROLLED
    for (int k = lane_id(); k + 1 < n; k += 32) best = fmin(best, point_segment_distance(p, r[k], r[k + 1]));
#pragma unroll
    for (int o = 16; o; o >>= 1) best = fmin(best, __shfl_xor_sync(FULLMASK, best, o));
    return best;
}
// GO.coveredby(point, ring): interior or boundary, collisions.jl:99
__device__ __forceinline__ bool warp_point_coveredby(double2 p, const double2 *r, int n) {
    bool onb = false, in = false;
    SZ_ROLLED
    for (int k = lane_id(); k + 1 < n; k += 32) {
        double2 a = r[k], b = r[k + 1];
        if (orient2d(a, b, p) == 0.0 && p.x >= fmin(a.x, b.x) && p.x <= fmax(a.x, b.x) && p.y >= fmin(a.y, b.y) &&
            p.y <= fmax(a.y, b.y))
            onb = true;
        if ((a.y > p.y) != (b.y > p.y)) {
            double xi = a.x + (p.y - a.y) / (b.y - a.y) * (b.x - a.x);
            if (p.x < xi) in = !in;
        }
    }
    unsigned par = __ballot_sync(FULLMASK, in);
    bool anyb = __any_sync(FULLMASK, onb);
    return anyb || (__popc(par) & 1);
}
// P point inside perturbed ring Q
__device__ __forceinline__ bool warp_point_in_ring_q(double2 p, const double2 *r, int n) {
    bool in = false;
    SZ_ROLLED
    for (int k = lane_id(); k + 1 < n; k += 32) {
        double2 c = r[k], d = r[k + 1];
        if (c.y < p.y && p.y <= d.y) {
            if (side_q(orient2d(c, d, p), c, d)) in = !in;
        } else if (d.y < p.y && p.y <= c.y) {
            if (!side_q(orient2d(c, d, p), c, d)) in = !in;
        }
    }
    return __popc(__ballot_sync(FULLMASK, in)) & 1;
}
// perturbed Q point inside ring P
__device__ __forceinline__ bool warp_point_in_ring_p(double2 q, const double2 *r, int n) {
    bool in = false;
    SZ_ROLLED
    for (int k = lane_id(); k + 1 < n; k += 32) {
        double2 a = r[k], b = r[k + 1];
        if (a.y <= q.y && q.y < b.y) {
            if (side_p(orient2d(a, b, q), a, b)) in = !in;
        } else if (b.y <= q.y && q.y < a.y) {
            if (!side_p(orient2d(a, b, q), a, b)) in = !in;
        }
    }
    return __popc(__ballot_sync(FULLMASK, in)) & 1;
}
// GO.intersects(ringA, ringB), collisions.jl:64
__device__ __forceinline__ bool warp_rings_intersect(const double2 *A, int na, const double2 *B, int nb) {
    int ea = na - 1, eb = nb - 1, tot = ea * eb;
    bool hit = false;
    SZ_ROLLED
    for (int idx = lane_id(); idx < tot; idx += 32) {
        int e = idx / eb, f = idx - e * eb;
        double2 p0, p1;
        if (segment_intersection(A[e], A[e + 1], B[f], B[f + 1], p0, p1) > 0) hit = true;
    }
    if (__any_sync(FULLMASK, hit)) return true;
    if (warp_point_coveredby(A[0], B, nb)) return true;
    if (warp_point_coveredby(B[0], A, na)) return true;
    return false;
}

// ---- per-warp workspace, carved from dynamic shared memory ------------------------------------
struct Ws {
    int maxv, maxx, rcap, maxreg, maxip;
    double2 *P, *Q, *P2, *R1, *R2, *xp, *ip;
    double *xt, *xs, *area1, *ct;  // ct: contacts [maxreg][6] = fx, fy, px, py, overlap, dl
    short *xe, *xf, *rankP, *rankQ, *ordP, *ordQ;
    short *rs1, *re1, *rs2, *re2, *minrank, *keepr, *ipidx;
    unsigned char *xentry, *xvis, *ipdup;
};

__host__ __device__ inline int ws_rcap(int maxv, int maxx) { return 2 * maxv + 2 * maxx + 8; }
__host__ __device__ inline int ws_maxreg(int maxx) { return maxx / 2 + 1; }
__host__ __device__ inline int ws_maxip(int maxx) { return 2 * maxx + 8; }
__host__ __device__ inline size_t ws_bytes(int maxv, int maxx) {
    size_t rcap = ws_rcap(maxv, maxx), maxreg = ws_maxreg(maxx), maxip = ws_maxip(maxx);
    size_t b = 16 * (3 * (size_t)maxv + 2 * rcap + maxx + maxip);
    b += 8 * (2 * (size_t)maxx + maxreg + 6 * maxreg);
    b += 2 * (6 * (size_t)maxx + 6 * maxreg + maxip);
    b += 2 * (size_t)maxx + maxip;
    return (b + 15) & ~(size_t)15;
}
__device__ inline Ws ws_carve(unsigned char *base, int maxv, int maxx) {
    Ws w;
    w.maxv = maxv;
    w.maxx = maxx;
    w.rcap = ws_rcap(maxv, maxx);
    w.maxreg = ws_maxreg(maxx);
    w.maxip = ws_maxip(maxx);
    double2 *d2 = (double2 *)base;
    w.P = d2; d2 += maxv;
    w.Q = d2; d2 += maxv;
    w.P2 = d2; d2 += maxv;
    w.R1 = d2; d2 += w.rcap;
    w.R2 = d2; d2 += w.rcap;
    w.xp = d2; d2 += maxx;
    w.ip = d2; d2 += w.maxip;
    double *d = (double *)d2;
    w.xt = d; d += maxx;
    w.xs = d; d += maxx;
    w.area1 = d; d += w.maxreg;
    w.ct = d; d += 6 * w.maxreg;
    short *s = (short *)d;
    w.xe = s; s += maxx;
    w.xf = s; s += maxx;
    w.rankP = s; s += maxx;
    w.rankQ = s; s += maxx;
    w.ordP = s; s += maxx;
    w.ordQ = s; s += maxx;
    w.rs1 = s; s += w.maxreg;
    w.re1 = s; s += w.maxreg;
    w.rs2 = s; s += w.maxreg;
    w.re2 = s; s += w.maxreg;
    w.minrank = s; s += w.maxreg;
    w.keepr = s; s += w.maxreg;
    w.ipidx = s; s += w.maxip;
    unsigned char *c = (unsigned char *)s;
    w.xentry = c; c += maxx;
    w.xvis = c; c += maxx;
    w.ipdup = c;
    return w;
}

enum { CLIP_OK = 0, CLIP_FAIL = 1, CLIP_OVERFLOW = 2 };

// intersect_polys (floe_utils.jl:55): clip ring P (npp points, closed) against ring Q (nqp
// points).  Result: nreg closed rings inside `R` (region r = R[rs[r] .. re[r])), ordered by
// their first crossing along P (the order test_collisions.jl:64-81,135-150 pin).  All lanes
// return the same (nreg, status).  Weiler-Atherton trace:
//  1. crossings: P-edge (a,b) and Q-edge (c,d) cross iff a,b lie on different sides of cd and
//     c,d on different sides of ab; point = a + t (b - a), t = o1 / (o1 - o2);
//  2. a crossing is an ENTRY (P goes into Q) iff b lies on Q's interior side;
//  3. from each entry in P order: follow P to the next crossing (an exit), then Q (forward if P
//     and Q have the same orientation, else backward) to the next crossing, until closed;
//  4. no crossings: P inside Q -> P; Q inside P -> Q; else nothing.
// ranking parameters of a crossing and the perturbation tie-break: see szo_rank_params / szo_sym_before in
// oracle/szo_geom.h (the same expressions, hence the same bits)
__device__ __forceinline__ void rank_params(double2 a, double2 b, double2 c, double2 d, double o1, double o2, double o3, double o4,
                                            double &t, double &s) {
    t = o1 / (o1 - o2);
    s = o3 / (o3 - o4);
    if (o1 == 0.0 || o2 == 0.0) {
        const double2 wv = o1 == 0.0 ? a : b;
        const double vx = d.x - c.x, vy = d.y - c.y;
        s = ((wv.x - c.x) * vx + (wv.y - c.y) * vy) / (vx * vx + vy * vy);
    }
    if (o3 == 0.0 || o4 == 0.0) {
        const double2 wv = o3 == 0.0 ? c : d;
        const double ux = b.x - a.x, uy = b.y - a.y;
        t = ((wv.x - a.x) * ux + (wv.y - a.y) * uy) / (ux * ux + uy * uy);
    }
}
__device__ __noinline__ bool sym_before(const double2 *P, const double2 *Q, int em, int fm, int ek, int fk, bool along_p, bool m_lt_k) {
    double ct[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int e = i == 0 ? em : ek, f = i == 0 ? fm : fk;
        const double ux = P[e + 1].x - P[e].x, uy = P[e + 1].y - P[e].y;
        const double vx = Q[f + 1].x - Q[f].x, vy = Q[f + 1].y - Q[f].y;
        const double dot = ux * vx + uy * vy;
        const double crs = along_p ? vx * uy - vy * ux : ux * vy - uy * vx;  // v x u along P, u x v along Q
        if (crs == 0.0) return m_lt_k;
        ct[i] = dot / crs;
    }
    if (ct[0] == ct[1]) return m_lt_k;
    if (along_p) {
        const double ux = P[em + 1].x - P[em].x, uy = P[em + 1].y - P[em].y;  // the shared P edge
        const bool sg = uy != 0.0 ? (uy > 0.0) : (ux < 0.0);                  // sign(eps u_y - eps^2 u_x) > 0
        return sg ? ct[0] > ct[1] : ct[0] < ct[1];
    }
    const double vx = Q[fm + 1].x - Q[fm].x, vy = Q[fm + 1].y - Q[fm].y;  // the shared Q edge
    const bool sg = vy != 0.0 ? (vy > 0.0) : (vx < 0.0);
    return sg ? ct[0] < ct[1] : ct[0] > ct[1];
}

__device__ int warp_clip(const Ws &w, const double2 *P, int npp, const double2 *Q, int nqp, double2 *R, short *rs,
                         short *re, int &status) {
    const int RCAP = w.rcap, MAXREG = w.maxreg, MAXX = w.maxx;
    const int lane = lane_id();
    const int np = npp - 1, nq = nqp - 1;
    status = CLIP_OK;
    if (np < 3 || nq < 3) return 0;
    const bool p_ccw = warp_ring_is_ccw(P, npp), q_ccw = warp_ring_is_ccw(Q, nqp);
    const bool same = (p_ccw == q_ccw);
    // 1. crossings, one edge pair per lane, compacted in (e, f) order
    int K = 0;
    const int tot = np * nq;
    SZ_ROLLED
    for (int base = 0; base < tot; base += 32) {
        int idx = base + lane;
        bool hit = false;
        int e = 0, f = 0;
        double t = 0, s = 0;
        double2 xp = make_double2(0.0, 0.0);
        bool ent = false;
        if (idx < tot) {
            e = idx / nq;
            f = idx - e * nq;
            double2 a = P[e], b = P[e + 1];
            double2 c = Q[f], d = Q[f + 1];
            double o1 = orient2d(c, d, a), o2 = orient2d(c, d, b);
            bool sa = side_q(o1, c, d), sb = side_q(o2, c, d);
            if (sa != sb) {
                double o3 = orient2d(a, b, c), o4 = orient2d(a, b, d);
                bool sc = side_p(o3, a, b), sd = side_p(o4, a, b);
                if (sc != sd) {
                    hit = true;
                    const double t0 = o1 / (o1 - o2);
                    rank_params(a, b, c, d, o1, o2, o3, o4, t, s);  // ranking parameters; the point uses t0
                    xp.x = a.x + t0 * (b.x - a.x);
                    xp.y = a.y + t0 * (b.y - a.y);
                    ent = (sb == q_ccw);
                }
            }
        }
        unsigned m = __ballot_sync(FULLMASK, hit);
        if (hit) {
            int pos = K + __popc(m & lanemask_lt());
            if (pos < MAXX) {
                w.xe[pos] = (short)e;
                w.xf[pos] = (short)f;
                w.xt[pos] = t;
                w.xs[pos] = s;
                w.xp[pos] = xp;
                w.xentry[pos] = ent;
                w.xvis[pos] = 0;
            }
        }
        K += __popc(m);
    }
    if (K > MAXX) {
        status = CLIP_OVERFLOW;
        return 0;
    }
    __syncwarp();
    if (K == 0) {
        // containment
        bool pin = warp_point_in_ring_q(P[0], Q, nqp);
        bool qin = pin ? false : warp_point_in_ring_p(Q[0], P, npp);
        const double2 *src = pin ? P : (qin ? Q : nullptr);
        int ns = pin ? npp : nqp;
        if (!src) return 0;
        if (ns > RCAP) {
            status = CLIP_OVERFLOW;
            return 0;
        }
        SZ_ROLLED
        for (int k = lane; k < ns; k += 32) R[k] = src[k];
        if (lane == 0) {
            rs[0] = 0;
            re[0] = (short)ns;
        }
        __syncwarp();
        return 1;
    }
    // 2. ranks along P (e, t, index) and along Q (f, s, index)
    int nentry = 0;
    SZ_ROLLED
    for (int k = lane; k < K; k += 32) {
        int rp = 0, rq = 0;
        int ek = w.xe[k], fk = w.xf[k];
        double tk = w.xt[k], sk = w.xs[k];
        SZ_ROLLED
        for (int m = 0; m < K; ++m) {
            if (m == k) continue;
            int em = w.xe[m], fm = w.xf[m];
            double tm = w.xt[m], sm = w.xs[m];
            if (em < ek || (em == ek && (tm < tk || (tm == tk && sym_before(P, Q, em, fm, ek, fk, true, m < k))))) rp++;
            if (fm < fk || (fm == fk && (sm < sk || (sm == sk && sym_before(P, Q, em, fm, ek, fk, false, m < k))))) rq++;
        }
        w.rankP[k] = (short)rp;
        w.rankQ[k] = (short)rq;
        w.ordP[rp] = (short)k;
        w.ordQ[rq] = (short)k;
        nentry += w.xentry[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) nentry += __shfl_xor_sync(FULLMASK, nentry, o);
    __syncwarp();
    int nreg = 0;
    int ok = ((K & 1) == 0) && (2 * nentry == K);
    // 3. trace (sequential, lane 0)
    if (lane == 0 && ok) {
        int npts = 0;
        SZ_ROLLED
        for (int r = 0; ok && r < K; ++r) {
            int startk = w.ordP[r];
            if (!w.xentry[startk] || w.xvis[startk]) continue;
            int start = npts, cur = startk, mr = K, guard = 0;
#define SZ_PUSH(pt)                                                                 \
    do {                                                                            \
        double2 _p = (pt);                                                          \
        if (!(npts > start && R[npts - 1].x == _p.x && R[npts - 1].y == _p.y)) {    \
            if (npts >= RCAP - 1) {                                                 \
                ok = 0;                                                             \
                status = CLIP_OVERFLOW;                                             \
            } else                                                                  \
                R[npts++] = _p;                                                     \
        }                                                                           \
    } while (0)
            SZ_ROLLED
            while (ok) {
                if (w.xvis[cur]) { ok = 0; break; }
                w.xvis[cur] = 1;
                if (w.rankP[cur] < mr) mr = w.rankP[cur];
                SZ_PUSH(w.xp[cur]);
                int rn = (w.rankP[cur] + 1) % K, nx = w.ordP[rn];
                int cnt = w.xe[nx] - w.xe[cur] + (rn == 0 ? np : 0);
                SZ_ROLLED
                for (int k = 0; k < cnt && ok; ++k) SZ_PUSH(P[(w.xe[cur] + 1 + k) % np]);
                if (!ok) break;
                if (w.xentry[nx] || w.xvis[nx]) { ok = 0; break; }
                w.xvis[nx] = 1;
                if (w.rankP[nx] < mr) mr = w.rankP[nx];
                SZ_PUSH(w.xp[nx]);
                int nn;
                if (same) {
                    int rq = (w.rankQ[nx] + 1) % K;
                    nn = w.ordQ[rq];
                    cnt = w.xf[nn] - w.xf[nx] + (rq == 0 ? nq : 0);
                    SZ_ROLLED
                    for (int k = 0; k < cnt && ok; ++k) SZ_PUSH(Q[(w.xf[nx] + 1 + k) % nq]);
                } else {
                    int rq = (w.rankQ[nx] - 1 + K) % K;
                    nn = w.ordQ[rq];
                    cnt = w.xf[nx] - w.xf[nn] + (w.rankQ[nx] == 0 ? nq : 0);
                    SZ_ROLLED
                    for (int k = 0; k < cnt && ok; ++k) SZ_PUSH(Q[(w.xf[nx] - k + nq) % nq]);
                }
                if (!ok) break;
                if (!w.xentry[nn]) { ok = 0; break; }
                if (nn == startk) break;
                cur = nn;
                if (++guard > K) { ok = 0; break; }
            }
#undef SZ_PUSH
            if (!ok) break;
            if (npts - start > 1 && R[npts - 1].x == R[start].x && R[npts - 1].y == R[start].y) npts--;
            if (npts - start < 3) { npts = start; continue; }
            R[npts] = R[start];
            npts++;
            if (ring_area2_seq(R + start, npts - start) == 0.0) { npts = start; continue; }
            if (nreg >= MAXREG) { ok = 0; status = CLIP_OVERFLOW; break; }
            // stable insertion by first-crossing rank
            int pos = nreg;
            SZ_ROLLED
            while (pos > 0 && w.minrank[pos - 1] > mr) {
                w.minrank[pos] = w.minrank[pos - 1];
                rs[pos] = rs[pos - 1];
                re[pos] = re[pos - 1];
                --pos;
            }
            w.minrank[pos] = (short)mr;
            rs[pos] = (short)start;
            re[pos] = (short)npts;
            nreg++;
        }
        if (!ok) {
            nreg = 0;
            if (status == CLIP_OK) status = CLIP_FAIL;
        }
    } else if (lane == 0 && !ok) {
        status = CLIP_FAIL;
    }
    nreg = __shfl_sync(FULLMASK, nreg, 0);
    status = __shfl_sync(FULLMASK, status, 0);
    __syncwarp();
    return nreg;
}

// GO.intersection_points(P, Q), collisions.jl:156: every edge-edge intersection point in
// (edge of P, edge of Q) order, de-duplicated (first occurrence kept).  Result in w.ip.
__device__ int warp_intersection_points(const Ws &w, const double2 *P, int npp, const double2 *Q, int nqp,
                                        int &status) {
    const int lane = lane_id();
    const int np = npp - 1, nq = nqp - 1, tot = np * nq;
    int n = 0;
    SZ_ROLLED
    for (int base = 0; base < tot; base += 32) {
        int idx = base + lane, c = 0;
        double2 p0 = make_double2(0.0, 0.0), p1 = p0;
        if (idx < tot) {
            int e = idx / nq, f = idx - e * nq;
            c = segment_intersection(P[e], P[e + 1], Q[f], Q[f + 1], p0, p1);
        }
        unsigned b1 = __ballot_sync(FULLMASK, c >= 1), b2 = __ballot_sync(FULLMASK, c == 2);
        int pos = n + __popc(b1 & lanemask_lt()) + __popc(b2 & lanemask_lt());
        if (c >= 1 && pos < w.maxip) w.ip[pos] = p0;
        if (c == 2 && pos + 1 < w.maxip) w.ip[pos + 1] = p1;
        n += __popc(b1) + __popc(b2);
    }
    if (n > w.maxip) {
        status = CLIP_OVERFLOW;
        return 0;
    }
    __syncwarp();
    SZ_ROLLED
    for (int k = lane; k < n; k += 32) {
        double2 p = w.ip[k];
        bool dup = false;
        SZ_ROLLED
        for (int m = 0; m < k && !dup; ++m) dup = (w.ip[m].x == p.x && w.ip[m].y == p.y);
        w.ipdup[k] = dup;
    }
    __syncwarp();
    int out = 0;
    SZ_ROLLED
    for (int base = 0; base < n; base += 32) {
        int k = base + lane;
        bool keep = k < n && !w.ipdup[k];
        double2 p = keep ? w.ip[k] : make_double2(0.0, 0.0);
        unsigned m = __ballot_sync(FULLMASK, keep);
        __syncwarp();
        if (keep) w.ip[out + __popc(m & lanemask_lt())] = p;
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

// which_vertices_match_points, floe_utils.jl:331-352 (0-based, sorted, duplicates kept).
__device__ int warp_match_vertices(const Ws &w, int nip, const double2 *reg, int nr) {
    const int lane = lane_id();
    int npoints = nip;
    if (nip > 0 && w.ip[0].x == w.ip[nip - 1].x && w.ip[0].y == w.ip[nip - 1].y) npoints -= 1;
    int m = 0;
    SZ_ROLLED
    for (int base = 0; base < npoints; base += 32) {
        int i = base + lane;
        bool hit = false;
        int min_vert = 0;
        if (i < npoints) {
            double2 p = w.ip[i];
            double min_dist = INFINITY;
            SZ_ROLLED
            for (int j = 0; j < nr; ++j) {
                double dx = reg[j].x - p.x, dy = reg[j].y - p.y;
                double dist = sqrt(sqrt(dx * dx + dy * dy));  // sqrt(GO.distance(..)), floe_utils.jl:341
                if (dist < min_dist) {
                    min_dist = dist;
                    min_vert = j;
                }
            }
            hit = min_dist < 1.0;
        }
        unsigned b = __ballot_sync(FULLMASK, hit);
        if (hit) w.ipidx[m + __popc(b & lanemask_lt())] = (short)min_vert;
        m += __popc(b);
    }
    __syncwarp();
    if (lane == 0) {
        SZ_ROLLED
        for (int a = 1; a < m; ++a) {
            short v = w.ipidx[a];
            int b = a - 1;
            SZ_ROLLED
            while (b >= 0 && w.ipidx[b] > v) {
                w.ipidx[b + 1] = w.ipidx[b];
                --b;
            }
            w.ipidx[b + 1] = v;
        }
    }
    __syncwarp();
    return m;
}

// _many_intersect_normal_force!, collisions.jl:78-119.  Sequential over the region's edges (the
// sums feed a direction), warp-parallel inside each edge's point predicates.
__device__ double warp_many_intersect_normal(double dir[2], const double2 *reg, int nr, const double2 *P, int npp,
                                             double ff) {
    double x1 = 0, y1 = 0, dl = 0, Fx = 0, Fy = 0;
    int n_pts = 0;
    SZ_ROLLED
    for (int i = 0; i < nr; ++i) {
        double x2 = reg[i].x, y2 = reg[i].y;
        if (i == 0) {
            x1 = x2;
            y1 = y2;
            continue;
        }
        double xmid = 0.5 * (x2 + x1), ymid = 0.5 * (y2 + y1);
        double dist = warp_point_ring_distance(make_double2(xmid, ymid), P, npp);
        if (dist < 1e-8) {
            double dx = x2 - x1, dy = y2 - y1;
            double mag = sqrt(dx * dx + dy * dy);
            double xt = xmid + (-dy / (100 * mag));
            double yt = ymid + (dx / (100 * mag));
            bool in_region = warp_point_coveredby(make_double2(xt, yt), reg, nr);
            double fs = (in_region ? 1.0 : -1.0) * ff;
            Fx = Fx + fs * (-dy);
            Fy = Fy + fs * dx;
            dl += mag;
            n_pts += 1;
        }
        x1 = x2;
        y1 = y2;
    }
    if (0 < n_pts && n_pts < nr - 1) {
        dl /= n_pts;
        if (dl > 0.1) {
            double nf = sqrt(Fx * Fx + Fy * Fy);
            dir[0] = Fx / nf;
            dir[1] = Fy / nf;
        }
    }
    return dl;
}

// calc_normal_force, collisions.jl:30-70.  Returns Δl; force = dir * area * force_factor.
__device__ double warp_normal_force(const Ws &w, const double2 *P, int npp, const double2 *Q, int nqp,
                                    const double2 *reg, int nr, double area, int nip, double ff, double force[2],
                                    int &status, uint32_t &flags) {
    double dir[2] = {0.0, 0.0}, dl = 0.0;
    int m = warp_match_vertices(w, nip, reg, nr);
    if (m == 2) {
        int i0 = w.ipidx[0], i1 = w.ipidx[1];
        double dx = reg[i1].x - reg[i0].x, dy = reg[i1].y - reg[i0].y;
        dl = sqrt(dx * dx + dy * dy);
        if (dl > 0.1) {
            dir[0] = -dy / dl;
            dir[1] = dx / dl;
        }
    } else if (m != 0) {
        dl = warp_many_intersect_normal(dir, reg, nr, P, npp, ff);
    }
    if (dl > 0.1) {
        SZ_ROLLED
        for (int k = lane_id(); k < npp; k += 32) w.P2[k] = make_double2(P[k].x + dir[0], P[k].y + dir[1]);
        __syncwarp();
        int st2;
        int nreg2 = warp_clip(w, w.P2, npp, Q, nqp, w.R2, w.rs2, w.re2, st2);
        if (st2 == CLIP_OVERFLOW) {
            status = CLIP_OVERFLOW;
            return 0.0;
        }
        if (st2 == CLIP_FAIL) flags |= IT_CLIPFAIL;
        SZ_ROLLED
        for (int r = 0; r < nreg2; ++r) {
            const double2 *nr_ = w.R2 + w.rs2[r];
            int nn = w.re2[r] - w.rs2[r];
            if (warp_rings_intersect(nr_, nn, reg, nr) && ring_area_seq(nr_, nn) / area > 1) {
                dir[0] *= -1;
                dir[1] *= -1;
            }
        }
    }
    force[0] = dir[0] * area * ff;
    force[1] = dir[1] * area * ff;
    return dl;
}

// calc_elastic_forces, collisions.jl:149-188.  Regions of clip #1 are in w.R1 / rs1 / re1 with
// areas in w.area1.  Contacts are written to w.ct (fx, fy, px, py, overlap, Δl); returns their
// number.
__device__ int warp_elastic_forces(const Ws &w, const double2 *P, int npp, const double2 *Q, int nqp, int nreg,
                                   double ff, int &status, uint32_t &flags) {
    const int lane = lane_id();
    int nip = warp_intersection_points(w, P, npp, Q, nqp, status);
    if (status == CLIP_OVERFLOW) return 0;
    int ncontact = 0;
    if (nip >= 2) {
        int n1 = npp - 1, n2 = nqp - 1;
        double min_area = (double)((n1 < n2 ? n1 : n2) * 100) / 1.75;
        SZ_ROLLED
        for (int r = 0; r < nreg; ++r)
            if (!(w.area1[r] < min_area)) {
                if (lane == 0) w.keepr[ncontact] = (short)r;
                ncontact++;
            }
    }
    __syncwarp();
    SZ_ROLLED
    for (int k = 0; k < ncontact; ++k) {
        int r = w.keepr[k];
        double ov = w.area1[r];
        double force[2] = {0.0, 0.0}, fp[2] = {0.0, 0.0}, dl = 0.0;
        if (ov != 0) {
            const double2 *reg = w.R1 + w.rs1[r];
            int nr = w.re1[r] - w.rs1[r];
            double2 ce = ring_centroid_seq(reg, nr);
            fp[0] = ce.x;
            fp[1] = ce.y;
            dl = warp_normal_force(w, P, npp, Q, nqp, reg, nr, ov, nip, ff, force, status, flags);
            if (status == CLIP_OVERFLOW) return 0;
        }
        if (lane == 0) {
            double *c = w.ct + 6 * k;
            c[0] = force[0];
            c[1] = force[1];
            c[2] = fp[0];
            c[3] = fp[1];
            c[4] = ov;
            c[5] = dl;
        }
    }
    __syncwarp();
    return ncontact;
}

// calc_friction_forces, collisions.jl:243-283 with _get_velocity (:206-214): contact-point
// velocity is (u + xi (x - cx), v + xi (y - cy)) as the reference writes it.
__device__ __forceinline__ void friction_force(double E, double nu, double mu, double dt, double iu0, double iv0,
                                               double ixi, double icx, double icy, double ju0, double jv0,
                                               double jxi, double jcx, double jcy, const double *c, double out[2]) {
    double G = E / (2 * (1 + nu));
    double px = c[2], py = c[3];
    double nnorm = sqrt(c[0] * c[0] + c[1] * c[1]);
    double iu = iu0 + ixi * (px - icx);
    double iv = iv0 + ixi * (py - icy);
    double ju = ju0 + jxi * (px - jcx);
    double jv = jv0 + jxi * (py - jcy);
    double udiff = iu - ju, vdiff = iv - jv;
    double vnorm = sqrt(udiff * udiff + vdiff * vdiff);
    double xdir = 0, ydir = 0;
    if (udiff != 0 || vdiff != 0) {
        xdir = udiff / vnorm;
        ydir = vdiff / vnorm;
    }
    double dot_dir = xdir * udiff + ydir * vdiff;
    double xf = G * c[5] * dt * nnorm * xdir * -dot_dir;
    double yf = G * c[5] * dt * nnorm * ydir * -dot_dir;
    double nf = sqrt(xf * xf + yf * yf);
    if (nf > mu * nnorm) {
        xf = -mu * nnorm * xdir;
        yf = -mu * nnorm * ydir;
    }
    out[0] = xf;
    out[1] = yf;
}
