"""Shared comparison helpers of the parity tests (CUDA product vs CPU oracle, same inputs)."""
import numpy as np

STATE_FIELDS = ("centroid_x", "centroid_y", "height", "area", "mass", "rmax", "moment", "alpha", "u", "v", "xi",
                "fxOA", "fyOA", "trqOA", "hflx_factor", "overarea", "collision_force", "collision_trq",
                "stress_accum", "stress_instant", "strain", "p_dxdt", "p_dydt", "p_dudt", "p_dvdt", "p_dxidt",
                "p_dalphadt", "vert_xy")
INT_FIELDS = ("status_tag", "id", "ghost_id", "ghost_offsets", "ghost_index", "vert_offsets", "mc_offsets")

RTOL = 1e-9  # BASELINE.json north_star: per-floe force, torque and updated state within 1e-9 relative


def rel_err(x, y):
    """max |x - y| / max(|y|, rms(y)): relative to the value, floored by the field's typical
    magnitude so that sums which cancel to ~0 are judged against their terms."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    if x.shape != y.shape:
        return np.inf
    if x.size == 0:
        return 0.0
    scale = max(float(np.sqrt(np.mean(y * y))), 1e-300)
    return float(np.max(np.abs(x - y) / np.maximum(np.abs(y), scale)))


def compare_state(a, b, tol=RTOL, exact=(), skip=()):
    """a: product FloeArrays, b: oracle FloeArrays.  Returns a list of failure strings."""
    bad = []
    if a.n != b.n or a.n_init != b.n_init:
        return ["n %d/%d vs %d/%d" % (a.n, a.n_init, b.n, b.n_init)]
    for name in INT_FIELDS:
        x, y = np.asarray(getattr(a, name)), np.asarray(getattr(b, name))
        if x.shape != y.shape or not np.array_equal(x, y):
            bad.append("%s differs (int)" % name)
    for name in STATE_FIELDS:
        if name in skip:
            continue
        x, y = np.asarray(getattr(a, name)), np.asarray(getattr(b, name))
        if x.shape != y.shape:
            bad.append("%s shape %s vs %s" % (name, x.shape, y.shape))
            continue
        if name in exact:
            if not np.array_equal(x, y):
                k = np.argmax(np.abs(x - y).reshape(len(x), -1).max(axis=1)) if x.size else 0
                bad.append("%s not bit-equal (rel %.3e, worst row %d: %s vs %s)" % (name, rel_err(x, y), k, x[k], y[k]))
        else:
            e = rel_err(x, y)
            if not e < tol:
                k = np.argmax(np.abs(x - y).reshape(len(x), -1).max(axis=1)) if x.size else 0
                bad.append("%s rel err %.3e > %.1e (worst row %d: %s vs %s)" % (name, e, tol, k, x[k], y[k]))
    return bad


def compare_collision_outputs(hg, ho):
    """Pair sets and interaction rows must be bit-exact (integer / index work + unfused FP64)."""
    bad = []
    cg, co = hg.counts(), ho.counts()
    for k in ("n_total", "n_candidates", "n_pairs", "n_overlap", "n_fuse", "n_rows", "n_domain_pairs"):
        if cg[k] != co[k]:
            bad.append("count %s: %d vs %d" % (k, cg[k], co[k]))
    for which, name in enumerate(("candidates", "filtered", "overlap", "fuse")):
        pg, po = hg.pairs(which), ho.pairs(which)
        if pg.shape != po.shape or not np.array_equal(pg, po):
            sg, so = set(map(tuple, pg.tolist())), set(map(tuple, po.tolist()))
            bad.append("%s pairs differ: %d vs %d, only-gpu %s only-oracle %s" %
                       (name, len(pg), len(po), sorted(sg - so)[:5], sorted(so - sg)[:5]))
    og, rg = hg.interactions()
    oo, ro = ho.interactions()
    if og.shape != oo.shape or not np.array_equal(og, oo):
        d = np.nonzero(np.diff(og) != np.diff(oo))[0] if og.shape == oo.shape else []
        bad.append("row offsets differ (first floes %s)" % list(d[:5]))
    elif not np.array_equal(rg, ro):
        d = np.nonzero(np.any(rg != ro, axis=1))[0]
        k = d[0]
        fl = int(np.searchsorted(og, k, side="right") - 1)
        bad.append("rows differ in %d of %d rows (rel %.3e); first row %d (floe %d): %s vs %s" %
                   (len(d), len(rg), rel_err(rg, ro), k, fl, rg[k], ro[k]))
    if cg["n_clip_fail"] or co["n_clip_fail"]:
        bad.append("clip failures: %d vs %d" % (cg["n_clip_fail"], co["n_clip_fail"]))
    return bad
