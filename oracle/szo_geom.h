/* szo_geom.h — CPU ORACLE geometry primitives.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build or call anything under oracle/.  The product (subzero.jl_b200/csrc)
 * never includes this file.
 *
 * These restate the third-party primitives the reference's hot path calls through
 * GeometryOps.jl (compat "0.1", /root/reference/Project.toml:38 — NOT vendored, so the
 * exact arithmetic is unavailable).  What is pinned by the
 * reference's own golden tests: areas, centroids, region ORDER for two regions
 * (test/test_physical_processes/test_collisions.jl:50-81,124-150).  Everything below
 * that level (ring start vertex, sub-1e-2 digits, exactly degenerate inputs) is
 * "parity unpinned": this file is then the definition, and the CUDA kernels follow it
 * operation for operation (no FMA contraction on either side).
 *
 * Call sites restated:
 *   intersect_polys            src/floe_utils.jl:55        -> szo_clip
 *   GO.area                    collisions.jl:360,515,64    -> szo_ring_area
 *   GO.centroid                collisions.jl:178           -> szo_ring_centroid
 *   GO.intersection_points     collisions.jl:156           -> szo_intersection_points
 *   GO.intersects              collisions.jl:64            -> szo_rings_intersect
 *   GO.signed_distance (abs)   collisions.jl:91            -> szo_point_ring_distance
 *   GO.coveredby               collisions.jl:99            -> szo_point_coveredby
 */
#ifndef SZO_GEOM_H
#define SZO_GEOM_H

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double x, y;
} szo_pt;

/* A list of closed rings (first point repeated at the end). */
typedef struct {
    int nreg;
    int *off; /* nreg + 1 */
    szo_pt *pts;
    int cap_reg, cap_pts;
    int failed; /* 1 if the trace met an inconsistent (degenerate) configuration */
} szo_regions;

static inline void szo_regions_init(szo_regions *r) { memset(r, 0, sizeof(*r)); }
static inline void szo_regions_free(szo_regions *r) {
    free(r->off);
    free(r->pts);
    memset(r, 0, sizeof(*r));
}
static inline void szo_regions_clear(szo_regions *r) {
    r->nreg = 0;
    r->failed = 0;
}
static inline void szo_regions_reserve(szo_regions *r, int nreg, int npts) {
    if (nreg + 1 > r->cap_reg) {
        r->cap_reg = 2 * (nreg + 1);
        r->off = (int *)realloc(r->off, sizeof(int) * (size_t)r->cap_reg);
    }
    if (npts > r->cap_pts) {
        r->cap_pts = 2 * npts;
        r->pts = (szo_pt *)realloc(r->pts, sizeof(szo_pt) * (size_t)r->cap_pts);
    }
    if (r->nreg == 0 && r->off) r->off[0] = 0;
}

/* (b-a) x (c-a); sign > 0 means c is left of a->b. Plain floating point on purpose:
 * the CUDA side evaluates the identical expression. */
static inline double szo_orient(szo_pt a, szo_pt b, szo_pt c) {
    return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x);
}

/* Twice the signed shoelace area of a closed ring of n points (Appendix B). */
static inline double szo_ring_area2(const szo_pt *r, int n) {
    double a = 0.0;
    for (int k = 0; k + 1 < n; ++k) a += r[k].x * r[k + 1].y - r[k].y * r[k + 1].x;
    return a;
}
/* GO.area of a polygon without holes = |signed area|. */
static inline double szo_ring_area(const szo_pt *r, int n) { return fabs(szo_ring_area2(r, n) / 2.0); }

/* Area-weighted centroid of a closed ring: sum (p_k + p_k+1) a_k / (6 A). */
static inline szo_pt szo_ring_centroid(const szo_pt *r, int n) {
    double a = 0.0, cx = 0.0, cy = 0.0;
    for (int k = 0; k + 1 < n; ++k) {
        double c = r[k].x * r[k + 1].y - r[k].y * r[k + 1].x;
        a += c;
        cx += (r[k].x + r[k + 1].x) * c;
        cy += (r[k].y + r[k + 1].y) * c;
    }
    a /= 2.0;
    szo_pt o = {cx / (6.0 * a), cy / (6.0 * a)};
    return o;
}

/* Even-odd crossing test with the half-open rule; boundary points are NOT special-cased. */
static inline int szo_point_in_ring(szo_pt p, const szo_pt *r, int n) {
    int in = 0;
    for (int k = 0; k + 1 < n; ++k) {
        szo_pt a = r[k], b = r[k + 1];
        if ((a.y > p.y) != (b.y > p.y)) {
            double xi = a.x + (p.y - a.y) / (b.y - a.y) * (b.x - a.x);
            if (p.x < xi) in = !in;
        }
    }
    return in;
}

static inline double szo_point_segment_distance(szo_pt p, szo_pt a, szo_pt b) {
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

/* |GO.signed_distance(point, polygon)|: distance to the nearest boundary segment. */
static inline double szo_point_ring_distance(szo_pt p, const szo_pt *r, int n) {
    double best = INFINITY;
    for (int k = 0; k + 1 < n; ++k) {
        double d = szo_point_segment_distance(p, r[k], r[k + 1]);
        if (d < best) best = d;
    }
    return best;
}

/* GO.coveredby(point, polygon): interior or boundary. */
static inline int szo_point_coveredby(szo_pt p, const szo_pt *r, int n) {
    for (int k = 0; k + 1 < n; ++k) {
        szo_pt a = r[k], b = r[k + 1];
        if (szo_orient(a, b, p) == 0.0 && p.x >= fmin(a.x, b.x) && p.x <= fmax(a.x, b.x) &&
            p.y >= fmin(a.y, b.y) && p.y <= fmax(a.y, b.y))
            return 1;
    }
    return szo_point_in_ring(p, r, n);
}

/* Closed-segment intersection (endpoints and collinear overlaps count).  Writes 0, 1 or 2
 * points.  The single-point formula a + t (b - a), t = o1 / (o1 - o2), is the same one
 * the clipper uses, so crossing points and intersection points are bit-identical. */
static inline int szo_segment_intersection(szo_pt a, szo_pt b, szo_pt c, szo_pt d, szo_pt out[2]) {
    double o1 = szo_orient(c, d, a), o2 = szo_orient(c, d, b);
    double o3 = szo_orient(a, b, c), o4 = szo_orient(a, b, d);
    if (o1 == 0.0 && o2 == 0.0) { /* collinear: overlap of the two intervals */
        int usex = fabs(b.x - a.x) >= fabs(b.y - a.y);
        double a0 = usex ? a.x : a.y, a1 = usex ? b.x : b.y;
        double c0 = usex ? c.x : c.y, c1 = usex ? d.x : d.y;
        szo_pt lo1 = a0 <= a1 ? a : b, hi1 = a0 <= a1 ? b : a;
        szo_pt lo2 = c0 <= c1 ? c : d, hi2 = c0 <= c1 ? d : c;
        double l1 = fmin(a0, a1), h1 = fmax(a0, a1), l2 = fmin(c0, c1), h2 = fmax(c0, c1);
        szo_pt lo = l1 >= l2 ? lo1 : lo2, hi = h1 <= h2 ? hi1 : hi2;
        double lv = fmax(l1, l2), hv = fmin(h1, h2);
        if (lv > hv) return 0;
        out[0] = lo;
        if (lv == hv) return 1;
        out[1] = hi;
        return 2;
    }
    if ((o1 > 0.0 && o2 > 0.0) || (o1 < 0.0 && o2 < 0.0)) return 0;
    if ((o3 > 0.0 && o4 > 0.0) || (o3 < 0.0 && o4 < 0.0)) return 0;
    if (o3 == 0.0 && o4 == 0.0) return 0; /* c,d on line ab but a,b not both on cd: impossible */
    if (o1 == 0.0) out[0] = a;
    else if (o2 == 0.0) out[0] = b;
    else if (o3 == 0.0) out[0] = c;
    else if (o4 == 0.0) out[0] = d;
    else {
        double t = o1 / (o1 - o2);
        out[0].x = a.x + t * (b.x - a.x);
        out[0].y = a.y + t * (b.y - a.y);
    }
    return 1;
}

/* GO.intersection_points(p1, p2): all edge-edge intersection points, de-duplicated,
 * in (edge of P, edge of Q) lexicographic discovery order.  Returns the count; *out is
 * malloc'ed (caller frees). */
static inline int szo_intersection_points(const szo_pt *P, int npp, const szo_pt *Q, int nqp,
                                          szo_pt **out) {
    int cap = 16, n = 0;
    szo_pt *v = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)cap);
    for (int e = 0; e + 1 < npp; ++e)
        for (int f = 0; f + 1 < nqp; ++f) {
            szo_pt tmp[2];
            int c = szo_segment_intersection(P[e], P[e + 1], Q[f], Q[f + 1], tmp);
            for (int k = 0; k < c; ++k) {
                int dup = 0;
                for (int m = 0; m < n && !dup; ++m) dup = (v[m].x == tmp[k].x && v[m].y == tmp[k].y);
                if (dup) continue;
                if (n == cap) {
                    cap *= 2;
                    v = (szo_pt *)realloc(v, sizeof(szo_pt) * (size_t)cap);
                }
                v[n++] = tmp[k];
            }
        }
    *out = v;
    return n;
}

/* GO.intersects(polyA, polyB): the two closed regions share at least one point. */
static inline int szo_rings_intersect(const szo_pt *A, int na, const szo_pt *B, int nb) {
    szo_pt tmp[2];
    for (int e = 0; e + 1 < na; ++e)
        for (int f = 0; f + 1 < nb; ++f)
            if (szo_segment_intersection(A[e], A[e + 1], B[f], B[f + 1], tmp) > 0) return 1;
    if (szo_point_coveredby(A[0], B, nb)) return 1;
    if (szo_point_coveredby(B[0], A, na)) return 1;
    return 0;
}

/* ---- polygon ∩ polygon ------------------------------------------------------------
 * Weiler–Atherton style trace for two simple rings without holes.
 *  0. degeneracies (vertex on an edge, collinear overlapping edges — every hand-made
 *     rectangle test of the reference has them) are resolved by a symbolic perturbation:
 *     Q is treated as translated by the infinitesimal vector (eps, eps^2).  That is a
 *     rigid motion, so all side decisions are mutually consistent; an orientation that
 *     evaluates to exactly 0 takes the sign of its first non-vanishing eps term
 *     (szo_side_q / szo_side_p).  Coordinates are never perturbed.
 *  1. crossings: P-edge (a,b) and Q-edge (c,d) cross iff a,b are on different sides of cd
 *     and c,d are on different sides of ab.  Point = a + t (b - a), t = o1 / (o1 - o2).
 *  2. a crossing is an ENTRY (P goes into Q) iff b lies on Q's interior side.
 *  3. from each entry in P order: follow P forward to the next crossing (an exit), then
 *     follow Q (forward if P and Q have the same orientation, else backward) to the next
 *     crossing along Q (an entry); repeat until the start is met.
 *  4. regions are ordered by their first crossing (entry OR exit) along P from P's first
 *     stored vertex — the order the reference's two-region golden tests pin
 *     (test_collisions.jl:64-81,135-150).
 *  5. no crossings: P inside Q -> P; Q inside P -> Q; else nothing.
 * Rings with fewer than 3 distinct points or zero area are dropped. */

/* side of a P point w.r.t. the perturbed Q edge c->d: sign of
 * o + (d.y - c.y) eps - (d.x - c.x) eps^2, o = orient(c, d, point).  1 = left. */
static inline int szo_side_q(double o, szo_pt c, szo_pt d) {
    if (o != 0.0) return o > 0.0;
    if (d.y != c.y) return d.y > c.y;
    return d.x <= c.x;
}
/* side of a perturbed Q point w.r.t. the P edge a->b: sign of
 * o - (b.y - a.y) eps + (b.x - a.x) eps^2, o = orient(a, b, point).  1 = left. */
static inline int szo_side_p(double o, szo_pt a, szo_pt b) {
    if (o != 0.0) return o > 0.0;
    if (b.y != a.y) return b.y < a.y;
    return b.x >= a.x;
}
/* is the P point p inside the perturbed ring Q (n points, closed)?  Crossing count of the
 * ray towards -x... evaluated with the same perturbation: p is compared as p - (eps, eps^2). */
static inline int szo_point_in_ring_q(szo_pt p, const szo_pt *r, int n) {
    int in = 0;
    for (int k = 0; k + 1 < n; ++k) {
        szo_pt c = r[k], d = r[k + 1];
        if (c.y < p.y && p.y <= d.y) { /* upward edge spans the (lowered) point */
            if (szo_side_q(szo_orient(c, d, p), c, d)) in = !in;
        } else if (d.y < p.y && p.y <= c.y) { /* downward edge */
            if (!szo_side_q(szo_orient(c, d, p), c, d)) in = !in;
        }
    }
    return in;
}
/* is the perturbed Q point q inside ring P?  q is compared as q + (eps, eps^2). */
static inline int szo_point_in_ring_p(szo_pt q, const szo_pt *r, int n) {
    int in = 0;
    for (int k = 0; k + 1 < n; ++k) {
        szo_pt a = r[k], b = r[k + 1];
        if (a.y <= q.y && q.y < b.y) {
            if (szo_side_p(szo_orient(a, b, q), a, b)) in = !in;
        } else if (b.y <= q.y && q.y < a.y) {
            if (!szo_side_p(szo_orient(a, b, q), a, b)) in = !in;
        }
    }
    return in;
}
typedef struct {
    int e, f;
    double t, s;   /* RANKING parameters along P / Q (szo_rank_params); the point itself is a + t0 (b - a) */
    szo_pt p;
    int entry, rankP, rankQ, visited;
} szo_xing;

/* Parameters that order the crossings along P (t) and along Q (s).  In general position they are the intersection
 * parameters o1 / (o1 - o2) and o3 / (o3 - o4).  When the crossing sits exactly AT a vertex — a P vertex lying exactly on
 * the Q edge's line (o1 == 0 or o2 == 0), a Q vertex exactly on the P edge's line (o3 == 0 or o4 == 0) — the two P (Q)
 * edges that meet in that vertex both cross there, and their parameters along the OTHER polygon's edge are two different
 * roundings of one number: they are replaced by ONE expression of the vertex itself (its projection on the edge), so that
 * the two compare equal and the tie is broken by the perturbation (szo_sym_before) instead of by rounding noise. */
static inline void szo_rank_params(szo_pt a, szo_pt b, szo_pt c, szo_pt d, double o1, double o2, double o3, double o4,
                                   double *t, double *s) {
    *t = o1 / (o1 - o2);
    *s = o3 / (o3 - o4);
    if (o1 == 0.0 || o2 == 0.0) {
        szo_pt w = o1 == 0.0 ? a : b;
        double vx = d.x - c.x, vy = d.y - c.y;
        *s = ((w.x - c.x) * vx + (w.y - c.y) * vy) / (vx * vx + vy * vy);
    }
    if (o3 == 0.0 || o4 == 0.0) {
        szo_pt w = o3 == 0.0 ? c : d;
        double ux = b.x - a.x, uy = b.y - a.y;
        *t = ((w.x - a.x) * ux + (w.y - a.y) * uy) / (ux * ux + uy * uy);
    }
}
/* Two crossings on the same edge with the same ranking parameter — they sit at ONE vertex w of the other ring that lies
 * exactly on this edge's line: which comes first once Q is translated by delta = (eps, eps^2)?
 * Along a Q edge (direction v; the two P edges u_m, u_k meet in w):  the P line w + tau u meets the shifted Q line at
 * s = s0 + [(delta x v) cot(u, v) - delta . v] / |v|^2, cot(u, v) = (u . v) / (u x v), so
 *     s_m < s_k  <=>  sign(delta x v) cot(u_m, v) < sign(delta x v) cot(u_k, v),   delta x v = eps v_y - eps^2 v_x.
 * Along a P edge (direction u; the two Q edges v_m, v_k meet in the shifted vertex w + delta):
 *     t = t0 + [delta . u - (delta x u) cot(v, u)] / |u|^2   =>   t_m < t_k  <=>  sign(delta x u) cot(v_m, u) > sign(delta x u) cot(v_k, u).
 * One division per crossing and no term that is mathematically common to both sides, so equal first-order shifts (a
 * horizontal or vertical edge) cannot be decided by rounding noise.  Returns 1 when crossing m comes before crossing k. */
static inline int szo_sym_before(const szo_pt *P, const szo_pt *Q, int em, int fm, int ek, int fk, int along_p, int m_lt_k) {
    double cot[2];
    const int e[2] = {em, ek}, f[2] = {fm, fk};
    for (int i = 0; i < 2; ++i) {
        double ux = P[e[i] + 1].x - P[e[i]].x, uy = P[e[i] + 1].y - P[e[i]].y;
        double vx = Q[f[i] + 1].x - Q[f[i]].x, vy = Q[f[i] + 1].y - Q[f[i]].y;
        double dot = ux * vx + uy * vy;
        double crs = along_p ? vx * uy - vy * ux : ux * vy - uy * vx; /* v x u along P, u x v along Q */
        if (crs == 0.0) return m_lt_k;
        cot[i] = dot / crs;
    }
    if (cot[0] == cot[1]) return m_lt_k;
    if (along_p) {
        double ux = P[em + 1].x - P[em].x, uy = P[em + 1].y - P[em].y; /* the shared P edge */
        int sg = uy != 0.0 ? (uy > 0.0) : (ux < 0.0);                  /* sign(eps u_y - eps^2 u_x) > 0 */
        return sg ? cot[0] > cot[1] : cot[0] < cot[1];
    } else {
        double vx = Q[fm + 1].x - Q[fm].x, vy = Q[fm + 1].y - Q[fm].y; /* the shared Q edge */
        int sg = vy != 0.0 ? (vy > 0.0) : (vx < 0.0);
        return sg ? cot[0] < cot[1] : cot[0] > cot[1];
    }
}

static inline void szo_push_pt(szo_regions *R, int *n, szo_pt p, int start) {
    if (*n > start && R->pts[*n - 1].x == p.x && R->pts[*n - 1].y == p.y) return;
    szo_regions_reserve(R, R->nreg + 1, *n + 2);
    R->pts[(*n)++] = p;
}

static inline int szo_clip(const szo_pt *P, int npp, const szo_pt *Q, int nqp, szo_regions *R) {
    int np = npp - 1, nq = nqp - 1;
    szo_regions_clear(R);
    szo_regions_reserve(R, 1, 8);
    R->off[0] = 0;
    if (np < 3 || nq < 3) return 0;
    double sP = szo_ring_area2(P, npp), sQ = szo_ring_area2(Q, nqp);
    int q_ccw = sQ > 0.0;
    int same = (sP > 0.0) == (sQ > 0.0);
    int K = 0, cap = 16;
    szo_xing *X = (szo_xing *)malloc(sizeof(szo_xing) * (size_t)cap);
    for (int e = 0; e < np; ++e) {
        szo_pt a = P[e], b = P[e + 1];
        for (int f = 0; f < nq; ++f) {
            szo_pt c = Q[f], d = Q[f + 1];
            double o1 = szo_orient(c, d, a), o2 = szo_orient(c, d, b);
            int sa = szo_side_q(o1, c, d), sb = szo_side_q(o2, c, d);
            if (sa == sb) continue;
            double o3 = szo_orient(a, b, c), o4 = szo_orient(a, b, d);
            int sc = szo_side_p(o3, a, b), sd = szo_side_p(o4, a, b);
            if (sc == sd) continue;
            if (K == cap) {
                cap *= 2;
                X = (szo_xing *)realloc(X, sizeof(szo_xing) * (size_t)cap);
            }
            szo_xing *x = &X[K++];
            x->e = e;
            x->f = f;
            double t0 = o1 / (o1 - o2);
            szo_rank_params(a, b, c, d, o1, o2, o3, o4, &x->t, &x->s);
            x->p.x = a.x + t0 * (b.x - a.x);
            x->p.y = a.y + t0 * (b.y - a.y);
            x->entry = (sb == q_ccw);
            x->visited = 0;
        }
    }
    if (K == 0) {
        const szo_pt *src = NULL;
        int ns = 0;
        if (szo_point_in_ring_q(P[0], Q, nqp)) {
            src = P;
            ns = npp;
        } else if (szo_point_in_ring_p(Q[0], P, npp)) {
            src = Q;
            ns = nqp;
        }
        if (src) {
            szo_regions_reserve(R, 1, ns);
            memcpy(R->pts, src, sizeof(szo_pt) * (size_t)ns);
            R->nreg = 1;
            R->off[1] = ns;
        }
        free(X);
        return R->nreg;
    }
    /* ranks along P (e, t, index) and along Q (f, s, index) */
    int *ordP = (int *)malloc(sizeof(int) * (size_t)K * 2), *ordQ = ordP + K;
    int nentry = 0;
    for (int k = 0; k < K; ++k) {
        int rp = 0, rq = 0;
        for (int m = 0; m < K; ++m) {
            if (m == k) continue;
            if (X[m].e < X[k].e || (X[m].e == X[k].e && (X[m].t < X[k].t || (X[m].t == X[k].t &&
                    szo_sym_before(P, Q, X[m].e, X[m].f, X[k].e, X[k].f, 1, m < k))))) rp++;
            if (X[m].f < X[k].f || (X[m].f == X[k].f && (X[m].s < X[k].s || (X[m].s == X[k].s &&
                    szo_sym_before(P, Q, X[m].e, X[m].f, X[k].e, X[k].f, 0, m < k))))) rq++;
        }
        X[k].rankP = rp;
        X[k].rankQ = rq;
        ordP[rp] = k;
        ordQ[rq] = k;
        nentry += X[k].entry;
    }
    int ok = (K % 2 == 0) && (2 * nentry == K);
    /* trace; keep (first crossing rank, ring) and order the rings afterwards */
    int *minrank = (int *)malloc(sizeof(int) * (size_t)(K + 1));
    int npts = 0;
    for (int r = 0; ok && r < K; ++r) {
        int startk = ordP[r];
        if (!X[startk].entry || X[startk].visited) continue;
        int start = npts, cur = startk, mr = K, guard = 0;
        while (1) {
            if (X[cur].visited) { ok = 0; break; }
            X[cur].visited = 1;
            if (X[cur].rankP < mr) mr = X[cur].rankP;
            szo_push_pt(R, &npts, X[cur].p, start);
            int rn = (X[cur].rankP + 1) % K, nx = ordP[rn];
            int cnt = X[nx].e - X[cur].e + (rn == 0 ? np : 0);
            for (int k = 0; k < cnt; ++k) szo_push_pt(R, &npts, P[(X[cur].e + 1 + k) % np], start);
            if (X[nx].entry || X[nx].visited) { ok = 0; break; }
            X[nx].visited = 1;
            if (X[nx].rankP < mr) mr = X[nx].rankP;
            szo_push_pt(R, &npts, X[nx].p, start);
            int nn;
            if (same) {
                int rq = (X[nx].rankQ + 1) % K;
                nn = ordQ[rq];
                cnt = X[nn].f - X[nx].f + (rq == 0 ? nq : 0);
                for (int k = 0; k < cnt; ++k) szo_push_pt(R, &npts, Q[(X[nx].f + 1 + k) % nq], start);
            } else {
                int rq = (X[nx].rankQ - 1 + K) % K;
                nn = ordQ[rq];
                cnt = X[nx].f - X[nn].f + (X[nx].rankQ == 0 ? nq : 0);
                for (int k = 0; k < cnt; ++k) szo_push_pt(R, &npts, Q[(X[nx].f - k + nq) % nq], start);
            }
            if (!X[nn].entry) { ok = 0; break; }
            if (nn == startk) break;
            cur = nn;
            if (++guard > K) { ok = 0; break; }
        }
        if (!ok) break;
        /* drop a trailing duplicate of the first point, then close */
        if (npts - start > 1 && R->pts[npts - 1].x == R->pts[start].x && R->pts[npts - 1].y == R->pts[start].y) npts--;
        if (npts - start < 3) {
            npts = start;
            continue;
        }
        szo_regions_reserve(R, R->nreg + 1, npts + 2);
        R->pts[npts] = R->pts[start];
        npts++;
        if (szo_ring_area2(R->pts + start, npts - start) == 0.0) {
            npts = start;
            continue;
        }
        minrank[R->nreg] = mr;
        R->nreg++;
        R->off[R->nreg] = npts;
    }
    if (!ok) {
        R->nreg = 0;
        R->failed = 1;
    } else if (R->nreg > 1) {
        /* stable selection by first-crossing rank: rebuild in order */
        int n = R->nreg;
        int *perm = (int *)malloc(sizeof(int) * (size_t)n);
        for (int i = 0; i < n; ++i) perm[i] = i;
        for (int i = 1; i < n; ++i) {
            int v = perm[i], j = i - 1;
            while (j >= 0 && minrank[perm[j]] > minrank[v]) {
                perm[j + 1] = perm[j];
                --j;
            }
            perm[j + 1] = v;
        }
        int sorted = 1;
        for (int i = 0; i < n; ++i) sorted &= (perm[i] == i);
        if (!sorted) {
            szo_pt *np_ = (szo_pt *)malloc(sizeof(szo_pt) * (size_t)R->cap_pts);
            int *no = (int *)malloc(sizeof(int) * (size_t)R->cap_reg);
            int w = 0;
            no[0] = 0;
            for (int i = 0; i < n; ++i) {
                int a = R->off[perm[i]], b = R->off[perm[i] + 1];
                memcpy(np_ + w, R->pts + a, sizeof(szo_pt) * (size_t)(b - a));
                w += b - a;
                no[i + 1] = w;
            }
            free(R->pts);
            free(R->off);
            R->pts = np_;
            R->off = no;
        }
        free(perm);
    }
    free(minrank);
    free(ordP);
    free(X);
    return R->nreg;
}

#endif /* SZO_GEOM_H */
