/* szo_points.h — oracle restatement of generate_subfloe_points (coupling.jl:172-208, :235-321).  TEST INFRASTRUCTURE.
 *
 * The random draws of the Monte-Carlo generator come from the counter-based generator the C ABI header specifies
 * (Julia's Xoshiro stream cannot be reproduced outside Julia); range(a, b, length = n) is restated as documented there. */
#ifndef SZO_POINTS_H
#define SZO_POINTS_H
#include <math.h>
#include <stdint.h>

static inline uint64_t szo_sm64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* u(seed, floe id, attempt, draw, axis) in [0, 1) */
static inline double szo_uniform(uint64_t seed, int64_t id, int attempt, int64_t draw, int axis) {
    uint64_t z = szo_sm64(seed ^ ((uint64_t)id * 0x9E3779B97F4A7C15ull));
    z = szo_sm64(z ^ (((uint64_t)attempt << 40) | (uint64_t)draw));
    z = szo_sm64(z ^ (uint64_t)axis);
    return (double)(z >> 11) * 0x1.0p-53;
}

/* element i of range(a, b, length = n) */
static inline double szo_range_elem(double a, double b, int64_t i, int64_t n) {
    if (i == 0 || n < 2) return a;
    if (i == n - 1) return b;
    /* d = b - a as a double-double */
    double dh = b - a, bb = dh - b, dl = (b - (dh - bb)) + (-a - bb);
    /* p = d * i */
    const double c = (double)i, m = (double)(n - 1);
    double ph = dh * c, pl = fma(dh, c, -ph) + dl * c;
    double s = ph + pl;
    pl = pl - (s - ph);
    ph = s;
    /* q = p / m */
    double qh = ph / m, th = qh * m, tl = fma(qh, m, -th);
    double ql = (((ph - th) - tl) + pl) / m;
    /* a + q */
    double sh = a + qh, t = sh - a, sl = (a - (sh - t)) + (qh - t);
    return sh + (sl + ql);
}
#endif
