// sz_narrow_thread.cuh — thread-per-item narrow phase for SMALL polygons (K3/K4 fast path).
//
// ncu on the warp-per-pair kernel (profiles/r1a_*) showed 3.4 k warp instructions per pair at
// 13/32 active lanes and `no_inst` (instruction fetch) as the top stall: a 7-gon pair has ~36 edge
// pairs, far too little to feed 32 lanes, while the sequential parts (trace, ranks, sorts) run on
// one lane.  Voronoi-packed fields are almost entirely such pairs, so they get this kernel: one
// THREAD owns one (polygon, polygon) item, a warp works on 32 items at once, and the rings live
// in shared memory laid out [point][thread] (a warp reading "its point k" touches 32 consecutive
// double2 = the 4-wavefront minimum, no bank conflicts).  The arithmetic is the same sequence of
// unfused FP64 operations as the warp kernel and the definition the parity tests check, so the
// results are bit-identical; anything that exceeds the fixed capacities below (more than 10 edges,
// more than 4 crossings, more than 2 regions, a degenerate trace) is handed to the warp kernel.
#pragma once
#include "sz_geom.cuh"

#ifndef TN_FN
#define TN_FN __noinline__  // measured: forcing these inline triples the kernel time (3.4 active lanes per instruction instead of 10)
#endif
#define TN_NT 128    // threads per block
#define TN_MAXV 14   // ring points incl. the closing point (<= 13 edges)
#define TN_MAXX 4    // crossings per clip
#define TN_MAXREG 2  // regions per clip
#define TN_RCAP 24   // region points of clip #1 AND clip #2 together (they share one buffer)
#define TN_RCAP_A 16 // phase 0 (clip #1 only): at most this many region rows (what phase 1 can take over); the few
                     // regions with more points go to the warp kernel
#define TN_MAXIP 4   // intersection points = the crossing points of clip #1 (<= TN_MAXX)
#define TN_MAXC 24   // edge pairs whose P edge straddles the Q edge's line

enum { TN_OK = 0, TN_DEFER = 1 };

// Instruction-cache pressure (profiles/README.md, r1k): the L1.5 instruction cache holds 32 KB; with the compiler's
// default 4x unrolling of these runtime-trip-count loops k_narrow_ab<1> was 165 KB of SASS and 25 % of its stall
// samples were `no_instructions`.  The loops below are kept rolled (measured one by one: match-vertices -5 %,
// clip trace / ranking -7 %; the ring gather of thread_item and the pass-1 orientation loop are better unrolled),
// and the rare paths (degenerate crossings, many-intersect normals, containment in the flip test) are not
// instantiated here at all: those items go to the warp kernel.  Together: narrow phase 1.17 -> 0.81 ms.
#define TN_ROLLED _Pragma("unroll 1")
#define U_MATCH TN_ROLLED
#define U_RINGS TN_ROLLED
#define U_CLIP TN_ROLLED
#define U_SMALL TN_ROLLED

// a ring in [point][thread] shared memory, optionally translated on the fly (P2 = P + dir is
// never stored: reading P[k] + dir yields the same bits every time)
// A position in the block's dynamic shared memory, in double2 units from its base.  Rings used to be passed to the
// __noinline__ helpers as generic pointers: every read inside them was a generic LD.E.128 behind a 64-bit address
// computation and a descriptor move (ncu r2t: 2.06 M of them per launch of k_narrow_ab<1> against 0.2 M LDS).  Through
// this 32-bit index the compiler knows the address space and emits LDS / STS with an immediate row offset.
extern __shared__ __align__(16) double2 tn_sm[];
#define TSP_NONE 0xffffffffu
struct TSp {
    unsigned a;
    __device__ __forceinline__ double2 &operator[](int k) const { return tn_sm[a + (unsigned)k]; }
    __device__ __forceinline__ TSp operator+(int k) const { TSp r; r.a = a + (unsigned)k; return r; }
    __device__ __forceinline__ bool none() const { return a == TSP_NONE; }
};
__device__ __forceinline__ TSp tsp(unsigned a) { TSp r; r.a = a; return r; }

struct TRing {
    TSp b;
    int n;
    double sx, sy;
    bool shifted;
};
__device__ __forceinline__ double2 tget(const TRing r, int k) { return r.b[k * TN_NT]; }
// the translated ring P2 = P + dir of calc_normal_force's second clip (collisions.jl:59-61)
template <bool SHIFT>
__device__ __forceinline__ double2 tgets(const TRing r, int k) {
    double2 v = r.b[k * TN_NT];
    if (SHIFT) {
        v.x = v.x + r.sx;
        v.y = v.y + r.sy;
    }
    return v;
}
__device__ __forceinline__ TRing tring(TSp b, int n) {
    TRing r;
    r.b = b;
    r.n = n;
    r.sx = r.sy = 0.0;
    r.shifted = false;
    return r;
}

__device__ TN_FN double t_area2(const TRing r) {
    double a = 0.0;
    double2 p = tget(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 q = tget(r, k + 1);
        a += p.x * q.y - p.y * q.x;
        p = q;
    }
    return a;
}
__device__ TN_FN double t_area2s(const TRing r) {
    double a = 0.0;
    double2 p = tgets<true>(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 q = tgets<true>(r, k + 1);
        a += p.x * q.y - p.y * q.x;
        p = q;
    }
    return a;
}
__device__ __forceinline__ double t_area(const TRing r) { return fabs(t_area2(r) / 2.0); }
__device__ TN_FN double2 t_centroid(const TRing r) {
    double a = 0.0, cx = 0.0, cy = 0.0;
    double2 p = tget(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 q = tget(r, k + 1);
        double c = p.x * q.y - p.y * q.x;
        a += c;
        cx += (p.x + q.x) * c;
        cy += (p.y + q.y) * c;
        p = q;
    }
    a /= 2.0;
    return make_double2(cx / (6.0 * a), cy / (6.0 * a));
}
__device__ TN_FN bool t_point_in_ring_q(double2 p, const TRing r) {
    bool in = false;
    double2 c = tget(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 d = tget(r, k + 1);
        if (c.y < p.y && p.y <= d.y) {
            if (side_q(orient2d(c, d, p), c, d)) in = !in;
        } else if (d.y < p.y && p.y <= c.y) {
            if (!side_q(orient2d(c, d, p), c, d)) in = !in;
        }
        c = d;
    }
    return in;
}
__device__ TN_FN bool t_point_in_ring_p(double2 q, const TRing r) {
    bool in = false;
    double2 a = tget(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 b = tget(r, k + 1);
        if (a.y <= q.y && q.y < b.y) {
            if (side_p(orient2d(a, b, q), a, b)) in = !in;
        } else if (b.y <= q.y && q.y < a.y) {
            if (!side_p(orient2d(a, b, q), a, b)) in = !in;
        }
        a = b;
    }
    return in;
}
__device__ TN_FN bool t_point_in_ring_ps(double2 q, const TRing r) {
    bool in = false;
    double2 a = tgets<true>(r, 0);
    U_SMALL
    for (int k = 0; k + 1 < r.n; ++k) {
        double2 b = tgets<true>(r, k + 1);
        if (a.y <= q.y && q.y < b.y) {
            if (side_p(orient2d(a, b, q), a, b)) in = !in;
        } else if (b.y <= q.y && q.y < a.y) {
            if (!side_p(orient2d(a, b, q), a, b)) in = !in;
        }
        a = b;
    }
    return in;
}
// NOTE on control flow in this file: no `return` / `break` out of nested loops.  ncu (profiles/r1f)
// showed 4-5 active lanes in the region trace: a divergent lane that leaves through an early return only
// reconverges at the function exit.  Failures set a flag, loops run to a structured exit.
__device__ TN_FN bool t_rings_intersect(const TRing A, const TRing B, bool &nohit) {
    bool hit = false;
    U_RINGS
    for (int e = 0; e + 1 < A.n && !hit; ++e) {
        double2 a = tget(A, e), b = tget(A, e + 1);
        double2 c = tget(B, 0);
        U_RINGS
        for (int f = 0; f + 1 < B.n; ++f) {
            double2 d = tget(B, f + 1), p0, p1;
            hit |= segment_intersection(a, b, c, d, p0, p1) > 0;
            c = d;
        }
    }
    if (!hit) nohit = true;  // containment or disjoint (GO.intersects then tests coveredby): rare, warp kernel
    return hit;
}

// intersect_polys for one thread; regions go to R ([point][thread]) as closed rings
// [rs[r], re[r]), ordered by first crossing along P.
//
// Crossing search: the side of every P vertex w.r.t. every Q edge is evaluated ONCE (np x nq
// orientation values, kept as one bit mask per Q edge); an edge pair (e, f) can only cross where
// the bit changes between vertex e and e+1, and only there the four orientations are evaluated
// in full (identical expressions, hence identical bits, as the exhaustive (e, f) loop).
// `xp_out` / `generic`: when no orientation value was exactly zero the crossing points ARE
// GO.intersection_points(P, Q) (closed-segment intersection == proper crossing), in the same
// (e, f) order; the caller then skips the separate 4 np nq pass.
// The result comes back PACKED in one 64-bit value (a __noinline__ function returns it in registers; reference or
// pointer outputs would live in local memory): number of regions, status, the two region ranges, the number of
// crossings and whether they are GO.intersection_points.
#define TC_PACK(nreg, status, rs0, re0, rs1, re1, K, gen)                                                          \
    ((unsigned long long)(nreg) | ((unsigned long long)(status) << 8) | ((unsigned long long)(rs0) << 16) |          \
     ((unsigned long long)(re0) << 24) | ((unsigned long long)(rs1) << 32) | ((unsigned long long)(re1) << 40) |     \
     ((unsigned long long)(K) << 48) | ((unsigned long long)(gen) << 56))
#define TC_NREG(v) ((int)((v)&0xffu))
#define TC_STATUS(v) ((int)(((v) >> 8) & 0xffu))
#define TC_RS(v, r) ((int)(((v) >> (16 + 16 * (r))) & 0xffu))
#define TC_RE(v, r) ((int)(((v) >> (24 + 16 * (r))) & 0xffu))
#define TC_K(v) ((int)(((v) >> 48) & 0xffu))
#define TC_GENERIC(v) ((bool)(((v) >> 56) & 1u))
template <bool SHIFT>
__device__ TN_FN unsigned long long t_clip(const TRing P, const TRing Q, TSp R, int rcap, TSp xp_out) {
    static_assert(TN_MAXREG == 2, "the packed result holds two regions");
    const int np = P.n - 1, nq = Q.n - 1;
    if (np < 3 || nq < 3) return TC_PACK(0, TN_OK, 0, 0, 0, 0, 0, 0);
    const bool q_ccw = t_area2(Q) > 0.0;
    const bool same = (SHIFT ? t_area2s(P) : t_area2(P)) > 0.0 == q_ccw;
    // The tables of the (at most TN_MAXX = 4) crossings live in REGISTERS: small integers packed 8 bits per crossing,
    // flags as bit masks, the doubles in arrays that are only ever indexed by unrolled compile-time constants (a
    // dynamic index goes through X4D below).  As dynamically indexed local arrays they cost a ~30-cycle local-memory
    // load per access in the divergent trace loop: 22 % of the stall samples of k_narrow_ab<0> in ncu r1n.
    static_assert(TN_MAXX == 4, "the packed crossing tables hold 4 entries");
    unsigned xe_p = 0, xf_p = 0, rankP_p = 0, rankQ_p = 0, ordP_p = 0, ordQ_p = 0, xent_m = 0, xvis_m = 0;
    double xt[TN_MAXX], xs[TN_MAXX];
    double2 xp[TN_MAXX];
#define X8(pack, k) (((pack) >> (8 * (k))) & 0xffu)
#define X4D(arr, k) ((k) == 0 ? arr[0] : ((k) == 1 ? arr[1] : ((k) == 2 ? arr[2] : arr[3])))
    // pass 1 (uniform across the warp): one orientation per (P vertex, Q edge); the (e, f) pairs whose
    // side bit changes from vertex e to e+1 are appended to a short candidate list in (e, f) order
    unsigned char ce[TN_MAXC], cf[TN_MAXC];
    int nc = 0;
    bool anyzero = false, fail = false;
    {
        unsigned first = 0, prev = 0;
        for (int v = 0; v <= np; ++v) {
            unsigned cur;
            if (v < np) {
                double2 pv = tgets<SHIFT>(P, v);
                double2 c = tget(Q, 0);
                cur = 0;
                        for (int f = 0; f < nq; ++f) {
                    double2 d = tget(Q, f + 1);
                    double o = orient2d(c, d, pv);
                    anyzero |= (o == 0.0);
                    cur |= (unsigned)side_q(o, c, d) << f;
                    c = d;
                }
                if (v == 0) first = cur;
            } else {
                cur = first;  // the closing point is vertex 0
            }
            if (v > 0) {
                unsigned chg = prev ^ cur;
                while (chg) {
                    int f = __ffs(chg) - 1;
                    chg &= chg - 1;
                    if (nc == TN_MAXC) {
                        fail = true;
                    } else {
                        ce[nc] = (unsigned char)(v - 1);
                        cf[nc] = (unsigned char)f;
                        nc++;
                    }
                }
            }
            prev = cur;
        }
    }
    // pass 2a: which candidates are crossings (the Q edge must straddle the P edge's line as well)
    int K = 0;
    U_CLIP
    for (int k = 0; k < nc; ++k) {
        const int e = ce[k], f = cf[k];
        double2 a = tgets<SHIFT>(P, e), b = tgets<SHIFT>(P, e + 1);
        double o3 = orient2d(a, b, tget(Q, f)), o4 = orient2d(a, b, tget(Q, f + 1));
        anyzero |= (o3 == 0.0) | (o4 == 0.0);
        if (side_p(o3, a, b) != side_p(o4, a, b)) {
            if (K == TN_MAXX) {
                fail = true;
            } else {
                xe_p |= (unsigned)e << (8 * K);
                xf_p |= (unsigned)f << (8 * K);
                K++;
            }
        }
    }
    // pass 2b (lanes in step again): parameters and point of every crossing
#pragma unroll
    for (int k = 0; k < TN_MAXX; ++k) {
        xt[k] = xs[k] = 0.0;
        xp[k] = make_double2(0.0, 0.0);
        if (k < K) {
            const int e = X8(xe_p, k), f = X8(xf_p, k);
            double2 a = tgets<SHIFT>(P, e), b = tgets<SHIFT>(P, e + 1);
            double2 c = tget(Q, f), d = tget(Q, f + 1);
            double o1 = orient2d(c, d, a), o2 = orient2d(c, d, b);
            double o3 = orient2d(a, b, c), o4 = orient2d(a, b, d);
            double t = o1 / (o1 - o2);
            xt[k] = t;
            xs[k] = o3 / (o3 - o4);
            xp[k] = make_double2(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y));
            xent_m |= (unsigned)(side_q(o2, c, d) == q_ccw) << k;
        }
    }
    if (fail) return TC_PACK(0, TN_DEFER, 0, 0, 0, 0, 0, 0);
    // a vertex exactly on the other ring's edge: two crossings may coincide there and their order along that edge is
    // decided by the perturbation (sym_before in sz_geom.cuh), which only the warp kernel implements
    if (anyzero && K > 1) return TC_PACK(0, TN_DEFER, 0, 0, 0, 0, 0, 0);
    bool gen = false;
    if (!xp_out.none()) {
        bool dup = false;
#pragma unroll
        for (int k = 0; k < TN_MAXX; ++k) {
            if (k < K) {
                xp_out[k * TN_NT] = xp[k];
#pragma unroll
                for (int m = 0; m < k; ++m) dup |= (xp[m].x == xp[k].x && xp[m].y == xp[k].y);
            }
        }
        gen = !anyzero && !dup;
    }
    if (K == 0) {
        bool pin = t_point_in_ring_q(tgets<SHIFT>(P, 0), Q);
        bool qin = pin ? false : (SHIFT ? t_point_in_ring_ps(tget(Q, 0), P) : t_point_in_ring_p(tget(Q, 0), P));
        if (!pin && !qin) return TC_PACK(0, TN_OK, 0, 0, 0, 0, 0, gen);
        const TRing src = pin ? P : Q;
        if (src.n > rcap) return TC_PACK(0, TN_DEFER, 0, 0, 0, 0, 0, 0);
        U_CLIP
        for (int k = 0; k < src.n; ++k) R[k * TN_NT] = pin ? tgets<SHIFT>(src, k) : tget(src, k);
        return TC_PACK(1, TN_OK, 0, src.n, 0, 0, 0, gen);
    }
#pragma unroll
    for (int k = 0; k < TN_MAXX; ++k) {
        if (k < K) {
            int rp = 0, rq = 0;
            const int ek = X8(xe_p, k), fk = X8(xf_p, k);
#pragma unroll
            for (int m = 0; m < TN_MAXX; ++m) {
                if (m == k || m >= K) continue;
                const int em = X8(xe_p, m), fm = X8(xf_p, m);
                if (em < ek || (em == ek && (xt[m] < xt[k] || (xt[m] == xt[k] && m < k)))) rp++;
                if (fm < fk || (fm == fk && (xs[m] < xs[k] || (xs[m] == xs[k] && m < k)))) rq++;
            }
            rankP_p |= (unsigned)rp << (8 * k);
            rankQ_p |= (unsigned)rq << (8 * k);
            ordP_p |= (unsigned)k << (8 * rp);
            ordQ_p |= (unsigned)k << (8 * rq);
        }
    }
    const int nentry = __popc(xent_m);
    if ((K & 1) || 2 * nentry != K) return TC_PACK(0, TN_DEFER, 0, 0, 0, 0, 0, 0);  // the warp kernel records the degenerate trace
    int nreg = 0, npts = 0, rs0 = 0, re0 = 0, rs1 = 0, re1 = 0, mr0 = 0;
#define TN_PUSH(pt)                                                                                  \
    do {                                                                                             \
        double2 _p = (pt);                                                                           \
        double2 _l = npts > start ? R[(npts - 1) * TN_NT] : make_double2(0.0, 0.0);                  \
        if (!(npts > start && _l.x == _p.x && _l.y == _p.y)) {                                       \
            if (npts >= rcap - 1) fail = true;                                                       \
            else R[(npts++) * TN_NT] = _p;                                                           \
        }                                                                                            \
    } while (0)
    U_CLIP
    for (int r = 0; r < K && !fail; ++r) {
        const int startk = X8(ordP_p, r);
        if (((xent_m & ~xvis_m) >> startk) & 1u) {
            int start = npts, cur = startk, mr = K, guard = 0;
            bool closed = false;
            while (!closed && !fail) {
                if ((xvis_m >> cur) & 1u) {
                    fail = true;
                } else {
                    xvis_m |= 1u << cur;
                    const int rpc = X8(rankP_p, cur), xec = X8(xe_p, cur);
                    if (rpc < mr) mr = rpc;
                    TN_PUSH(X4D(xp, cur));
                    const int rn = rpc + 1 == K ? 0 : rpc + 1, nx = X8(ordP_p, rn);
                    int cnt = (int)X8(xe_p, nx) - xec + (rn == 0 ? np : 0);
                    U_CLIP
                    for (int k = 0, v = xec + 1; k < cnt; ++k, ++v) {
                        if (v >= np) v -= np;
                        TN_PUSH(tgets<SHIFT>(P, v));
                    }
                    if (((xent_m | xvis_m) >> nx) & 1u) {
                        fail = true;
                    } else {
                        xvis_m |= 1u << nx;
                        const int rpn = X8(rankP_p, nx), rqn = X8(rankQ_p, nx), xfn = X8(xf_p, nx);
                        if (rpn < mr) mr = rpn;
                        TN_PUSH(X4D(xp, nx));
                        int nn;
                        if (same) {
                            int rq = rqn + 1 == K ? 0 : rqn + 1;
                            nn = X8(ordQ_p, rq);
                            cnt = (int)X8(xf_p, nn) - xfn + (rq == 0 ? nq : 0);
                            U_CLIP
                            for (int k = 0, v = xfn + 1; k < cnt; ++k, ++v) {
                                if (v >= nq) v -= nq;
                                TN_PUSH(tget(Q, v));
                            }
                        } else {
                            int rq = rqn == 0 ? K - 1 : rqn - 1;
                            nn = X8(ordQ_p, rq);
                            cnt = xfn - (int)X8(xf_p, nn) + (rqn == 0 ? nq : 0);
                            U_CLIP
                            for (int k = 0, v = xfn; k < cnt; ++k, --v) {
                                if (v < 0) v += nq;
                                TN_PUSH(tget(Q, v));
                            }
                        }
                        if (!((xent_m >> nn) & 1u)) fail = true;
                        else if (nn == startk) closed = true;
                        else {
                            cur = nn;
                            if (++guard > K) fail = true;
                        }
                    }
                }
            }
            if (!fail) {
                double2 f0 = R[start * TN_NT], l0 = R[(npts - 1) * TN_NT];
                if (npts - start > 1 && l0.x == f0.x && l0.y == f0.y) npts--;
                bool keep = npts - start >= 3;
                if (keep) {
                    R[npts * TN_NT] = f0;
                    npts++;
                    keep = t_area2(tring(R + start * TN_NT, npts - start)) != 0.0;
                }
                if (!keep) {
                    npts = start;
                } else if (nreg >= TN_MAXREG) {
                    fail = true;
                } else {  // regions ordered by their first crossing along P (two at most here)
                    if (nreg == 0) {
                        rs0 = start; re0 = npts; mr0 = mr;
                    } else if (mr0 > mr) {
                        rs1 = rs0; re1 = re0;
                        rs0 = start; re0 = npts; mr0 = mr;
                    } else {
                        rs1 = start; re1 = npts;
                    }
                    nreg++;
                }
            }
        }
    }
#undef TN_PUSH
#undef X8
#undef X4D
    if (fail) return TC_PACK(0, TN_DEFER, 0, 0, 0, 0, 0, 0);
    return TC_PACK(nreg, TN_OK, rs0, re0, rs1, re1, K, gen);
}

// which_vertices_match_points, floe_utils.jl:331-352.
// The reference takes, for every intersection point, the FIRST vertex with the smallest dist = sqrt(sqrt(d2)) and
// keeps it if dist < 1.  No square root is needed for that: sqrt is monotone and correctly rounded, so
//   * dist < 1  <=>  d2 < 1  (the largest double below 1 has a square root below 1), and
//   * the first vertex with the smallest dist is the first vertex with the smallest d2 unless another vertex has
//     a d2 within a few ulp above the minimum (then the two roots may tie and the earlier index wins).  That
//     near-tie (never seen on the bench fields) sets `rare`: the item is left to the warp kernel, which evaluates
//     the roots.  ncu r1k: the two divergent DSQRT expansions were 7 % of the instructions of k_narrow_ab<1>.
__device__ TN_FN int t_match_vertices(TSp ip, int nip, const TRing reg, int *idx, bool &rare) {
    int m = 0, npoints = nip;
    if (nip > 0) {
        double2 f = ip[0], l = ip[(nip - 1) * TN_NT];
        if (f.x == l.x && f.y == l.y) npoints -= 1;
    }
    U_MATCH
    for (int i = 0; i < npoints; ++i) {
        double2 p = ip[i * TN_NT];
        double m2 = INFINITY, s2 = INFINITY;  // smallest and second smallest DISTINCT squared distance
        int min_vert = 0;
        U_MATCH
        for (int j = 0; j < reg.n; ++j) {
            double2 v = tget(reg, j);
            double dx = v.x - p.x, dy = v.y - p.y;
            double d2 = dx * dx + dy * dy;
            if (d2 < m2) {
                s2 = m2;
                m2 = d2;
                min_vert = j;
            } else if (d2 > m2 && d2 < s2) {
                s2 = d2;
            }
        }
        rare |= s2 <= m2 * (1.0 + 1e-12);
        if (m2 < 1.0) idx[m++] = min_vert;
    }
    U_MATCH
    for (int a = 1; a < m; ++a) {
        int v = idx[a], b = a - 1;
        while (b >= 0 && idx[b] > v) {
            idx[b + 1] = idx[b];
            --b;
        }
        idx[b + 1] = v;
    }
    return m;
}

struct TWs {
    TSp P, Q, R1, R2, ip;  // [cap][TN_NT], already offset by the thread index
    int rcap;                       // rows of the region buffer (clip #1 and clip #2 share it)
    int r2cap;                      // R2 = the part of the region buffer clip #1 left free
};

// calc_normal_force, collisions.jl:30-70
// (m, i0, i1, rare): the result of t_match_vertices for this region — evaluated for ALL regions before the first
// second clip, because in k_narrow_ab<1> the crossing points share their rows with the regions of clip #2
// (one call site: inlined, so that `force` and `status` stay in registers)
__device__ __forceinline__ double t_normal_force(const TWs w, const TRing P, const TRing Q, const TRing reg,
                                              double area, int m, int i0, int i1, bool rare, double ff, double force[2], int &status) {
    double dir[2] = {0.0, 0.0}, dl = 0.0;
    // m not in {0, 2}: _many_intersect_normal_force! (collisions.jl:78-119) — rare, like a near-tie above: the item
    // goes to the warp kernel (no early return: a divergent return only reconverges at the function exit)
    bool defer = rare || (m != 2 && m != 0);
    if (m == 2 && !defer) {
        double2 v0 = tget(reg, i0), v1 = tget(reg, i1);
        double dx = v1.x - v0.x, dy = v1.y - v0.y;
        dl = sqrt(dx * dx + dy * dy);
        if (dl > 0.1) {
            dir[0] = -dy / dl;
            dir[1] = dx / dl;
        }
    }
    if (dl > 0.1) {
        TRing P2 = P;
        P2.shifted = true;
        P2.sx = dir[0];
        P2.sy = dir[1];
        const unsigned long long c2 = t_clip<true>(P2, Q, w.R2, w.r2cap, tsp(TSP_NONE));
        defer |= TC_STATUS(c2) != TN_OK;
        const int nreg2 = defer ? 0 : TC_NREG(c2);
#pragma unroll
        for (int r = 0; r < TN_MAXREG; ++r) {
            if (r >= nreg2) continue;
            TRing nr = tring(w.R2 + TC_RS(c2, r) * TN_NT, TC_RE(c2, r) - TC_RS(c2, r));
            bool nohit = false;
            if (t_rings_intersect(nr, reg, nohit) && t_area(nr) / area > 1) {
                dir[0] *= -1;
                dir[1] *= -1;
            }
            defer |= nohit;
        }
    }
    if (defer) status = TN_DEFER;
    force[0] = dir[0] * area * ff;
    force[1] = dir[1] * area * ff;
    return dl;
}

// One work item, one thread.  phase 0 (k_narrow_a): clip #1, areas and the fuse / remove decisions; an
// item that needs contact forces is handed to phase 1 (k_narrow_b), which repeats clip #1 (cheap next to the
// force part) and computes the rows.  Returns TI_DONE, TI_WARP (the warp kernel must take it) or TI_FORCES.
enum { TI_DONE = 0, TI_WARP = 1, TI_FORCES = 2 };
// what phase 0 hands to phase 1 (through global memory, SoA over the force list): the regions of clip #1,
// the crossing points and whether they are GO.intersection_points
struct TPre {
    int nreg, rs[TN_MAXREG], re[TN_MAXREG], K1, used, np, nq;
    bool generic;
};
template <int PHASE>
__device__ int thread_item(TWs w, const Store &S, const StepBuf &B, const Params &P, int slot, TPre &pre) {
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
    if (npp > TN_MAXV || nqp > TN_MAXV) return TI_WARP;
    {
        const double2 *gP = S.verts + S.vstart[fi];
        for (int k = 0; k < npp; ++k) w.P[k * TN_NT] = gP[k];
        if (gQ) {
                for (int k = 0; k < nqp; ++k) w.Q[k * TN_NT] = gQ[k];
        } else {  // _make_bounding_box_polygon, floe_utils.jl:104-108
            double xmin = D->rect[elem][0], xmax = D->rect[elem][1], ymin = D->rect[elem][2], ymax = D->rect[elem][3];
            w.Q[0 * TN_NT] = make_double2(xmin, ymin);
            w.Q[1 * TN_NT] = make_double2(xmin, ymax);
            w.Q[2 * TN_NT] = make_double2(xmax, ymax);
            w.Q[3 * TN_NT] = make_double2(xmax, ymin);
            w.Q[4 * TN_NT] = make_double2(xmin, ymin);
        }
    }
    const TRing Pr = tring(w.P, npp), Qr = tring(w.Q, nqp);
    int status = TN_OK;
    uint32_t flags = 0;
    int rs1[TN_MAXREG], re1[TN_MAXREG];
    double area1[TN_MAXREG];
    int K1 = 0, nreg, used = 0;
    bool generic = false;
    if (PHASE == 0) {
        const unsigned long long c1 = t_clip<false>(Pr, Qr, w.R1, w.rcap, w.ip);
        if (TC_STATUS(c1) != TN_OK) return TI_WARP;
        nreg = TC_NREG(c1);
        K1 = TC_K(c1);
        generic = TC_GENERIC(c1);
#pragma unroll
        for (int r = 0; r < TN_MAXREG; ++r) {
            rs1[r] = TC_RS(c1, r);
            re1[r] = TC_RE(c1, r);
        }
#pragma unroll
        for (int r = 0; r < TN_MAXREG; ++r)
            if (r < nreg) used = max(used, re1[r]);
        pre.nreg = nreg;
        pre.np = npp;
        pre.nq = nqp;
        pre.K1 = K1;
        pre.generic = generic;
        pre.used = used;
        for (int r = 0; r < TN_MAXREG; ++r) {
            pre.rs[r] = r < nreg ? rs1[r] : 0;
            pre.re[r] = r < nreg ? re1[r] : 0;
        }
    } else {  // the regions and crossing points were loaded into w.R1 / w.ip by the caller
        nreg = pre.nreg;
        K1 = pre.K1;
        generic = pre.generic;
        used = pre.used;
        for (int r = 0; r < TN_MAXREG; ++r) {
            rs1[r] = pre.rs[r];
            re1[r] = pre.re[r];
        }
    }
    w.R2 = w.R1 + used * TN_NT;
    w.r2cap = w.rcap - used;
    double total = 0.0, max_area = 0.0;
#pragma unroll
    for (int r = 0; r < TN_MAXREG; ++r) {  // constant indices: rs1 / re1 / area1 stay in registers
        area1[r] = 0.0;
        if (r < nreg) {
            area1[r] = t_area(tring(w.R1 + rs1[r] * TN_NT, re1[r] - rs1[r]));
            total += area1[r];
            if (area1[r] > max_area) max_area = area1[r];
        }
    }
    const double ai = S.area[fi], hi = S.height[fi];
    bool forces = false;
    double ff = 0.0, ju = 0.0, jv = 0.0, jxi = 0.0, jcx = 0.0, jcy = 0.0;
    if (is_pair) {
        if (total > 0) {  // collisions.jl:364-405
            flags |= IT_OVERLAP;
            const double aj = S.area[fj];
            if (fmax(total / ai, total / aj) > P.cfg.floe_floe_max_overlap) {
                flags |= IT_FUSE;
            } else {
                const double hj = S.height[fj];
                double ir = sqrt(ai), jr = sqrt(aj);
                ff = (ir > 1e5 || jr > 1e5) ? P.cfg.E * fmin(hi, hj) / fmin(ir, jr) : P.cfg.E * (hi * hj) / (hi * jr + hj * ir);
                forces = true;
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
            ff = P.cfg.E * hi / sqrt(ai);
            forces = true;
            if (elem < 4 && kind == SZ_BOUNDARY_MOVING) {
                ju = D->wu[elem];
                jv = D->wv[elem];
            }
        }
    }
    if (forces && PHASE == 0) return TI_FORCES;
    double rows[TN_MAXREG][NPOOL];
    int nrows = 0;
    if (forces) {
        // calc_elastic_forces, collisions.jl:149-188
        // GO.intersection_points: the crossing points of clip #1 when the configuration is generic
        // GO.intersection_points are the crossing points of clip #1 when the configuration is generic; the rare
        // degenerate configurations (an orientation exactly zero, duplicate points) go to the warp kernel
        if (!generic) return TI_WARP;
        const int nip = K1;
        if (nip >= 2) {
            int n1 = npp - 1, n2 = nqp - 1;
            double min_area = (double)((n1 < n2 ? n1 : n2) * 100) / 1.75;
            const double iu = S.u[fi], iv = S.v[fi], ixi = S.xi[fi], icx = S.cx[fi], icy = S.cy[fi];
            // which_vertices_match_points for every region first (the crossing points are dead afterwards)
            int mm[TN_MAXREG], mi0[TN_MAXREG], mi1[TN_MAXREG];
            bool rare = false;
#pragma unroll
            for (int r = 0; r < TN_MAXREG; ++r) {
                mm[r] = mi0[r] = mi1[r] = 0;
                if (r < nreg && !(area1[r] < min_area) && area1[r] != 0) {
                    int idx[TN_MAXIP];
                    mm[r] = t_match_vertices(w.ip, nip, tring(w.R1 + rs1[r] * TN_NT, re1[r] - rs1[r]), idx, rare);
                    mi0[r] = idx[0];
                    mi1[r] = idx[1];
                }
            }
#pragma unroll
            for (int r = 0; r < TN_MAXREG; ++r) {
                if (r >= nreg) continue;
                if (area1[r] < min_area) continue;
                double c[6] = {0.0, 0.0, 0.0, 0.0, area1[r], 0.0};
                if (area1[r] != 0) {
                    TRing reg = tring(w.R1 + rs1[r] * TN_NT, re1[r] - rs1[r]);
                    double2 ce = t_centroid(reg);
                    c[2] = ce.x;
                    c[3] = ce.y;
                    double force[2];
                    c[5] = t_normal_force(w, Pr, Qr, reg, area1[r], mm[r], mi0[r], mi1[r], rare, ff, force, status);
                    if (status != TN_OK) return TI_WARP;
                    c[0] = force[0];
                    c[1] = force[1];
                }
                if (!is_pair && elem < 4) {  // _normal_direction_correct!, boundaries.jl:37-40,73-76,110-113,147-150
                    if (elem == 0 && c[3] >= D->val[0]) c[0] = 0.0;
                    if (elem == 1 && c[3] <= D->val[1]) c[0] = 0.0;
                    if (elem == 2 && c[2] >= D->val[2]) c[1] = 0.0;
                    if (elem == 3 && c[2] <= D->val[3]) c[1] = 0.0;
                }
                double fr[2];
                friction_force(P.cfg.E, P.cfg.nu, P.cfg.mu, (double)P.cfg.dt, iu, iv, ixi, icx, icy, ju, jv, jxi, jcx, jcy, c,
                               fr);
                double fx = c[0] + fr[0], fy = c[1] + fr[1];
                if (fx != 0 || fy != 0) {  // add_interactions!, collisions.jl:288
                    // (constant first index: the two staged rows stay in registers)
                    if (nrows == 0) {
                        rows[0][0] = fx; rows[0][1] = fy; rows[0][2] = c[2]; rows[0][3] = c[3]; rows[0][4] = c[4];
                    } else {
                        rows[1][0] = fx; rows[1][1] = fy; rows[1][2] = c[2]; rows[1][3] = c[3]; rows[1][4] = c[4];
                    }
                    nrows++;
                }
            }
        }
    }
    int row0 = 0;
    if (nrows > 0) {
        row0 = atomicAdd(&cnt->n_pool, nrows);
        if (row0 + nrows > B.cap_pool) {
            atomicOr(&cnt->error, ERR_POOL_CAP);
            nrows = 0;
        } else {
#pragma unroll
            for (int k = 0; k < TN_MAXREG; ++k)
                if (k < nrows) {
#pragma unroll
                    for (int q = 0; q < NPOOL; ++q) B.pool[(size_t)(row0 + k) * NPOOL + q] = rows[k][q];
                }
        }
    }
    B.item_nrows[slot] = nrows;
    B.item_row0[slot] = row0;
    B.item_flags[slot] = flags | IT_DONE;
    if (is_pair && (flags & IT_OVERLAP)) atomicAdd(&cnt->n_overlap, 1);
    if (flags & IT_FUSE) {
        int s = atomicAdd(&cnt->n_fuse, 1);
        if (s < B.cap_fuse) B.fuse_pairs[s] = make_int2(fi, fj);
        else atomicOr(&cnt->error, ERR_FUSE_CAP);
    }
    return TI_DONE;
}

#define TN_ROWS_B 36  // rows per thread of k_narrow_ab<0/1>: three blocks per SM, split per warp (see the kernel)
#define TN_SMEM_A (sizeof(double2) * TN_NT * TN_ROWS_B)                             // k_narrow_ab<0>
#define TN_SMEM_B (sizeof(double2) * TN_NT * TN_ROWS_B)                             // k_narrow_ab<1>
#define TN_SMEM_C (sizeof(double2) * TN_NT * (2 * TN_MAXV + TN_RCAP))               // single-clip kernels (P, Q, R)
#define TN_NCE (TN_MAXV - 3)      // edge counts 3 .. TN_MAXV - 1
#define TN_NCLASS 128             // >= TN_NCE^2 classes: (edges of P - 3) * TN_NCE + (edges of Q - 3)
static_assert(TN_NCE * TN_NCE <= TN_NCLASS, "class table too small");

// ---- work-item ordering -----------------------------------------------------------------------------
// A warp of the thread-per-item kernels runs 32 items in lockstep, so its speed is set by the LARGEST
// rings among them (ncu r1e: 10 of 32 lanes active).  Items are therefore counting-sorted by the pair
// (edge count of P, edge count of Q): all lanes of a warp then walk identical loop trip counts.
__device__ __forceinline__ int item_slot(const StepBuf &B, int it, int np) { return it < np ? it : B.cap_pairs + (it - np); }
// -1: nothing to do (filtered pair), TN_NCLASS: too large for the thread kernels, else the class
__device__ __forceinline__ int item_class(const Store &S, const StepBuf &B, int it, int np) {
    int ep, eq;
    if (it < np) {
        if (!B.keep[it]) return -1;
        ep = S.vcount[B.pair_i[it]] - 1;
        eq = S.vcount[B.pair_j[it]] - 1;
    } else {
        int q = it - np, elem = B.dom_elem[q];
        ep = S.vcount[B.dom_floe[q]] - 1;
        eq = elem < 4 ? 4 : S.topo_vcount[elem - 4] - 1;
    }
    if (ep + 1 > TN_MAXV || eq + 1 > TN_MAXV) return TN_NCLASS;
    return (ep - 3) * TN_NCE + (eq - 3);
}

__global__ void __launch_bounds__(256) k_item_count(Store S, StepBuf B) {
    sz_pdl();
    __shared__ int hist[TN_NCLASS];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    for (int k = threadIdx.x; k < TN_NCLASS; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const int np = cnt->n_cand, total = np + cnt->n_dom;
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
        int c = item_class(S, B, it, np), slot = item_slot(B, it, np);
        if (c < 0) {
            B.item_nrows[slot] = 0;
            B.item_flags[slot] = 0;
        } else if (c == TN_NCLASS) {
            B.mid_items[atomicAdd(&cnt->n_mid, 1)] = slot;
            B.item_nrows[slot] = 0;
            B.item_flags[slot] = IT_NEEDLARGE;
        } else {
            atomicAdd(&hist[c], 1);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < TN_NCLASS; k += blockDim.x)
        if (hist[k]) atomicAdd(&B.class_count[k], hist[k]);
}

__global__ void k_class_scan(Store S, StepBuf B) {
    sz_pdl();  // one warp, four classes per lane
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int lane = threadIdx.x & 31;
    int c[4], tot = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c[k] = B.class_count[4 * lane + k];
        tot += c[k];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
    }
    int ex = incl - tot;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        B.class_base[4 * lane + k] = ex;
        B.class_cursor[4 * lane + k] = ex;
        B.class_count[4 * lane + k] = 0;  // ready for the next step
        ex += c[k];
    }
    if (lane == 31) {
        cnt->n_order = incl;
        cnt->n_force = 0;
    }
}

__global__ void __launch_bounds__(256) k_item_scatter(Store S, StepBuf B) {
    sz_pdl();
    __shared__ int hist[TN_NCLASS], base[TN_NCLASS];
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const int np = cnt->n_cand, total = np + cnt->n_dom;
    // block-strided chunks: rank inside the block first, then one global reservation per class and block
    for (int start = blockIdx.x * blockDim.x; start < total; start += gridDim.x * blockDim.x) {
        for (int k = threadIdx.x; k < TN_NCLASS; k += blockDim.x) hist[k] = 0;
        __syncthreads();
        int it = start + threadIdx.x, c = -1, rank = 0;
        if (it < total) {
            c = item_class(S, B, it, np);
            if (c >= 0 && c < TN_NCLASS) rank = atomicAdd(&hist[c], 1);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < TN_NCLASS; k += blockDim.x)
            if (hist[k]) base[k] = atomicAdd(&B.class_cursor[k], hist[k]);
        __syncthreads();
        if (c >= 0 && c < TN_NCLASS) {
            B.order[base[c] + rank] = item_slot(B, it, np);
            B.order_cls[base[c] + rank] = (unsigned char)c;
        }
        __syncthreads();
    }
}

// pre-clip records of the force list, SoA: point k of record b at [k * cap_force + b]
#define TN_PRE_PTS (TN_RCAP + TN_MAXX)

template <int PHASE>
__global__ void __launch_bounds__(TN_NT, 3) k_narrow_ab(Store S, StepBuf B, Params P) {
    sz_pdl();
    Counters *cnt = S.cnt;
    if (cnt->error) return;
    const TSp base = tsp(threadIdx.x);
    TWs w;
    w.P = base;
    w.Q = w.R1 = w.R2 = w.ip = base;
    w.rcap = 0;
    w.r2cap = 0;
    // With worst-case capacities (14 + 14 ring rows, 24 region rows, 4 crossing points) only two blocks would fit on
    // an SM.  The items are sorted by ring size, so every WARP lays out its 32 columns of the 36 rows for the largest
    // rings among its own items: P | Q | regions (what is left: 22 rows for two hexagons in phase 1).  Phase 0 keeps the
    // 4 crossing points in the last rows; in phase 1 they alias the last region rows and are consumed
    // (which_vertices_match_points of every region) before the first second clip may write there.  Regions that do not
    // fit go to the warp kernel.
    const int total = PHASE == 0 ? cnt->n_order : min(cnt->n_force, B.cap_force);
    const int *list = PHASE == 0 ? B.order : B.force_items;
    const int lane = threadIdx.x & 31;
    const size_t cf = (size_t)B.cap_force;
    for (int base_it = blockIdx.x * TN_NT + (threadIdx.x & ~31); base_it < total; base_it += gridDim.x * TN_NT) {
        const int it = base_it + lane;
        int rc = TI_DONE, slot = -1;
        TPre pre;
        if (it < total) {
            slot = list[it];
            if (PHASE == 1) {
                const int4 m = B.force_meta[it];
                pre.nreg = m.x & 0xff;
                pre.generic = (m.x >> 8) & 1;
                pre.K1 = (m.x >> 16) & 0xff;
                pre.used = m.w & 0xff;
                pre.np = (m.w >> 8) & 0xff;
                pre.nq = (m.w >> 16) & 0xff;
                pre.rs[0] = m.y & 0xffff; pre.re[0] = m.y >> 16;
                pre.rs[1] = m.z & 0xffff; pre.re[1] = m.z >> 16;
            }
        }
        if (PHASE == 0) {
            const int cls = it < total ? (int)B.order_cls[it] : -1;
            int npm = cls >= 0 ? cls / TN_NCE + 4 : 0, nqm = cls >= 0 ? cls % TN_NCE + 4 : 0;  // ring points = edges + 1
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                npm = max(npm, __shfl_xor_sync(0xffffffffu, npm, o));
                nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, o));
            }
            w.Q = w.P + npm * TN_NT;
            w.R1 = w.Q + nqm * TN_NT;
            w.R2 = w.R1;
            w.rcap = min(TN_ROWS_B - TN_MAXIP - npm - nqm, TN_RCAP_A);
            w.ip = w.P + (TN_ROWS_B - TN_MAXIP) * TN_NT;
        }
        if (PHASE == 1) {
            int npm = it < total ? pre.np : 0, nqm = it < total ? pre.nq : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                npm = max(npm, __shfl_xor_sync(0xffffffffu, npm, o));
                nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, o));
            }
            w.Q = w.P + npm * TN_NT;
            w.R1 = w.Q + nqm * TN_NT;
            w.R2 = w.R1;
            w.rcap = TN_ROWS_B - npm - nqm;  // clip #1's regions + the crossing points must fit: checked per item below
            w.ip = w.P + (TN_ROWS_B - TN_MAXIP) * TN_NT;  // the last rows of the region buffer: read before clip #2 writes there
        }
        if (it < total) {
            bool fits = true;
            if (PHASE == 1) {
                fits = pre.used + TN_MAXIP <= w.rcap;  // (a warp of 13-gons leaves 8 region rows)
                if (fits) {
                    for (int k = 0; k < pre.used; ++k) w.R1[k * TN_NT] = B.force_pts[k * cf + it];
                    for (int k = 0; k < pre.K1; ++k) w.ip[k * TN_NT] = B.force_pts[(TN_RCAP + k) * cf + it];
                }
            }
            rc = fits ? thread_item<PHASE>(w, S, B, P, slot, pre) : TI_WARP;
        }
        if (rc == TI_WARP) {
            B.mid_items[atomicAdd(&cnt->n_mid, 1)] = slot;
            B.item_nrows[slot] = 0;
            B.item_flags[slot] = IT_NEEDLARGE;
        }
        if (PHASE == 0) {
            // warp-aggregated append: the force list keeps the (class-sorted) order warp by warp
            unsigned m = __ballot_sync(0xffffffffu, rc == TI_FORCES);
            if (m) {
                int b0 = 0;
                if (lane == 0) b0 = atomicAdd(&cnt->n_force, __popc(m));
                b0 = __shfl_sync(0xffffffffu, b0, 0);
                if (rc == TI_FORCES) {
                    const int b = b0 + __popc(m & ((1u << lane) - 1));
                    if (b < B.cap_force) {
                        B.force_items[b] = slot;
                        B.force_meta[b] = make_int4(pre.nreg | ((int)pre.generic << 8) | (pre.K1 << 16), pre.rs[0] | (pre.re[0] << 16),
                                                    pre.rs[1] | (pre.re[1] << 16), pre.used | (pre.np << 8) | (pre.nq << 16));
                        for (int k = 0; k < pre.used; ++k) B.force_pts[k * cf + b] = w.R1[k * TN_NT];
                        for (int k = 0; k < pre.K1; ++k) B.force_pts[(TN_RCAP + k) * cf + b] = w.ip[k * TN_NT];
                    } else {
                        atomicOr(&cnt->error, ERR_POOL_CAP);  // the force list shares the contact pool's capacity
                        atomicMax(&cnt->n_pool, b + 1);
                    }
                }
            }
        }
    }
}
