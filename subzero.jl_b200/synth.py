"""Synthetic Voronoi-packed floe fields.

The reference builds such fields with VoronoiCells.jl + Xoshiro (floe.jl:548-634), whose random
streams cannot be reproduced outside Julia, so shapes and Monte-Carlo points are INPUTS here:
`scipy.spatial.Voronoi` of N seeds ~ U[0, L)^2 (numpy PCG64(seed)), every cell scaled about its
centroid by `scale` (0.99 = "packed", 1.01 = "dense contacts": each Voronoi neighbour pair
overlaps by a sliver), 1000 bounding-box draws per floe kept when inside (coupling.jl:194-201).

Everything is vectorised (CSR arrays, no per-floe Python objects) so 1e5-1e6 floes are
generated in seconds to a few minutes; results are cached under $SZ_SYNTH_CACHE (default
/tmp/subzero_b200_synth).
"""
import math
import os

import numpy as np

from . import capi, host


def _replicate(pts, L, margin, periodic):
    """Seeds plus their images within `margin` of the box: translated copies (periodic walls)
    or mirrored copies (collision walls: cells end exactly on the wall)."""
    out = [pts]
    x, y = pts[:, 0], pts[:, 1]
    for sx in (-1, 0, 1):
        for sy in (-1, 0, 1):
            if sx == 0 and sy == 0:
                continue
            m = np.ones(len(pts), dtype=bool)
            if sx == -1:
                m &= (x > L - margin) if periodic else (x < margin)
            if sx == 1:
                m &= (x < margin) if periodic else (x > L - margin)
            if sy == -1:
                m &= (y > L - margin) if periodic else (y < margin)
            if sy == 1:
                m &= (y < margin) if periodic else (y > L - margin)
            q = pts[m].copy()
            if periodic:
                q[:, 0] += sx * L
                q[:, 1] += sy * L
            else:
                if sx == -1:
                    q[:, 0] = -q[:, 0]
                if sx == 1:
                    q[:, 0] = 2 * L - q[:, 0]
                if sy == -1:
                    q[:, 1] = -q[:, 1]
                if sy == 1:
                    q[:, 1] = 2 * L - q[:, 1]
            out.append(q)
    return np.concatenate(out)


def voronoi_rings(n, L, seed, periodic, scale):
    """-> (offsets[n+1], xy[V,2]) closed clockwise rings of the n scaled Voronoi cells."""
    from scipy.spatial import Voronoi
    rng = np.random.Generator(np.random.PCG64(seed))
    pts = rng.random((n, 2)) * L
    margin = min(L, 8.0 * L / math.sqrt(n))
    allp = _replicate(pts, L, margin, periodic) if n > 1 else pts
    if len(allp) < 4:
        raise ValueError("need at least 4 seeds")
    vor = Voronoi(allp)
    regs = [vor.regions[vor.point_region[i]] for i in range(n)]
    cnt = np.fromiter((len(r) for r in regs), dtype=np.int64, count=n)
    flat = np.fromiter((v for r in regs for v in r), dtype=np.int64, count=int(cnt.sum()))
    if (flat < 0).any():
        raise RuntimeError("unbounded Voronoi cell: increase the replication margin")
    cell = np.repeat(np.arange(n), cnt)
    v = vor.vertices[flat]
    ang = np.arctan2(v[:, 1] - pts[cell, 1], v[:, 0] - pts[cell, 0])
    order = np.lexsort((-ang, cell))  # clockwise around the seed
    v = v[order]
    start = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    # area-weighted centroid of each (open) ring
    nxt = np.arange(len(v)) + 1
    last = start + cnt - 1
    nxt[last] = start
    cr = v[:, 0] * v[nxt, 1] - v[:, 1] * v[nxt, 0]
    a2 = np.add.reduceat(cr, start)
    cx = np.add.reduceat((v[:, 0] + v[nxt, 0]) * cr, start) / (3.0 * a2)
    cy = np.add.reduceat((v[:, 1] + v[nxt, 1]) * cr, start) / (3.0 * a2)
    c = np.stack([cx, cy], axis=1)[cell]
    v = c + scale * (v - c)
    # close the rings
    offs = np.concatenate([[0], np.cumsum(cnt + 1)])
    xy = np.empty((offs[-1], 2))
    dst = np.arange(len(v)) + cell  # each ring shifts by its index (one closing point per earlier ring)
    xy[dst] = v
    xy[offs[1:] - 1] = v[start]
    return offs.astype(np.int64), xy


def ring_properties(offs, xy, height, rho_i):
    """area, centroid, rmax, moment (floe_utils.jl:273-313) of closed CSR rings, vectorised."""
    n = len(offs) - 1
    cnt = np.diff(offs)
    seg = np.ones(len(xy), dtype=bool)
    seg[offs[1:] - 1] = False  # the closing point starts no edge
    i0 = np.nonzero(seg)[0]
    i1 = i0 + 1
    estart = offs[:-1] - np.arange(n)  # edge offsets (one edge fewer than points per ring)
    x0, y0, x1, y1 = xy[i0, 0], xy[i0, 1], xy[i1, 0], xy[i1, 1]
    cr = x0 * y1 - y0 * x1
    a2 = np.add.reduceat(cr, estart)
    area = np.abs(a2) / 2.0
    cx = np.add.reduceat((x0 + x1) * cr, estart) / (3.0 * a2)
    cy = np.add.reduceat((y0 + y1) * cr, estart) / (3.0 * a2)
    cell = np.repeat(np.arange(n), cnt)
    d2 = (xy[:, 0] - cx[cell]) ** 2 + (xy[:, 1] - cy[cell]) ** 2
    rmax = np.sqrt(np.maximum.reduceat(d2, offs[:-1]))
    ce = cell[i0]
    # _calc_moment_inertia as written in the reference (second centroid subtraction inside wi)
    X0, Y0, X1, Y1 = x0 - cx[ce], y0 - cy[ce], x1 - cx[ce], y1 - cy[ce]
    wi = (X0 - cx[ce]) * (Y1 - cy[ce]) - (X1 - cx[ce]) * (Y0 - cy[ce])
    Ixx = np.add.reduceat(wi * (Y0 ** 2 + Y0 * Y1 + Y1 ** 2), estart) / 12.0
    Iyy = np.add.reduceat(wi * (X0 ** 2 + X0 * X1 + X1 ** 2), estart) / 12.0
    moment = np.abs(Ixx + Iyy) * height * rho_i
    return area, cx, cy, rmax, moment


def mc_points_convex(offs, xy, cx, cy, npoints, seed, chunk=2048):
    """Monte-Carlo sub-floe points (coupling.jl:172-208) for CONVEX rings: `npoints` uniform
    draws in each body-frame bounding box, kept when inside.  -> (mc_offsets, mc_x, mc_y)."""
    n = len(offs) - 1
    cnt = np.diff(offs) - 1  # edges
    rng = np.random.Generator(np.random.PCG64(seed))
    xs, ys, counts = [], [], np.zeros(n, dtype=np.int64)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        m = b - a
        ne = int(cnt[a:b].max())
        idx = offs[a:b, None] + np.minimum(np.arange(ne + 1)[None, :], cnt[a:b, None])  # padded with the closing point
        rx = xy[idx, 0] - cx[a:b, None]
        ry = xy[idx, 1] - cy[a:b, None]
        xmin, xmax, ymin, ymax = rx.min(1), rx.max(1), ry.min(1), ry.max(1)
        px = xmin[:, None] + (xmax - xmin)[:, None] * rng.random((m, npoints))
        py = ymin[:, None] + (ymax - ymin)[:, None] * rng.random((m, npoints))
        inside = np.ones((m, npoints), dtype=bool)
        for e in range(ne):
            ex, ey = rx[:, e + 1] - rx[:, e], ry[:, e + 1] - ry[:, e]
            crs = ex[:, None] * (py - ry[:, e][:, None]) - ey[:, None] * (px - rx[:, e][:, None])
            inside &= crs <= 0.0  # clockwise ring: interior on the right (degenerate padded edges give 0)
        counts[a:b] = inside.sum(1)
        xs.append(px[inside])
        ys.append(py[inside])
    moffs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    return moffs, np.concatenate(xs), np.concatenate(ys)


class Field:
    """A synthetic configuration: FloeArrays + grid / ocean / atmos / domain / constants."""
    pass


def make_field(n, scale=1.01, walls="collision", seed=None, npoints=1000, hmean=0.25, flow="random",
               cache=True, rho_i=920.0):
    """walls: 'collision' (4 collision walls), 'periodic' (doubly periodic), 'shear'
    (periodic E-W, collision N-S; examples/shear_flow.jl).  flow: 'random' (u, v ~ U(-0.1, 0.1)) or
    'converging'."""
    seed = n if seed is None else seed
    key = "n%d_s%g_%s_seed%d_mc%d_%s.npz" % (n, scale, walls, seed, npoints, flow)
    cdir = os.environ.get("SZ_SYNTH_CACHE", "/tmp/subzero_b200_synth")
    path = os.path.join(cdir, key)
    L = math.ceil(math.sqrt(n * 4e6) / 1e4) * 1e4
    per_x = walls in ("periodic", "shear")
    per_y = walls == "periodic"
    d = None
    if cache and os.path.exists(path):
        try:
            d = dict(np.load(path))
        except Exception:
            d = None
    if d is None:
        offs, xy = voronoi_rings(n, L, seed, walls == "periodic", scale)
        area, cx, cy, rmax, moment = ring_properties(offs, xy, hmean, rho_i)
        moffs, mx, my = mc_points_convex(offs, xy, cx, cy, npoints, seed + 2)
        rng = np.random.Generator(np.random.PCG64(seed + 1))
        if flow == "converging":
            u = -0.2 * (cx - L / 2) / L
            v = -0.2 * (cy - L / 2) / L
        else:
            u = rng.uniform(-0.1, 0.1, n)
            v = rng.uniform(-0.1, 0.1, n)
        d = dict(offs=offs, xy=xy, area=area, cx=cx, cy=cy, rmax=rmax, moment=moment, moffs=moffs, mx=mx, my=my,
                 u=u, v=v)
        if cache:
            os.makedirs(cdir, exist_ok=True)
            tmp = path + ".tmp%d.npz" % os.getpid()
            np.savez(tmp, **d)
            os.replace(tmp, path)
    fa = capi.FloeArrays(n)
    fa.centroid_x, fa.centroid_y = d["cx"], d["cy"]
    fa.area, fa.rmax, fa.moment = d["area"], d["rmax"], d["moment"]
    fa.height = np.full(n, hmean)
    fa.mass = d["area"] * hmean * rho_i
    fa.u, fa.v = d["u"], d["v"]
    fa.vert_offsets, fa.vert_xy = d["offs"], d["xy"]
    fa.mc_offsets, fa.mc_x, fa.mc_y = d["moffs"], d["mx"], d["my"]
    # a floe whose MC generation failed is tagged for removal (coupling.jl:189-205)
    fa.status_tag[np.diff(d["moffs"]) == 0] = capi.STATUS_REMOVE
    f = Field()
    f.n, f.L, f.floes = n, L, fa
    f.grid = host.RegRectilinearGrid(0.0, L, 0.0, L, dx=1e4, dy=1e4)
    g = f.grid
    # ocean: triangular shear profile u(y) 0 -> 0.5 -> 0 m/s (examples/shear_flow.jl:15-18), v = 0
    yl = np.linspace(g.y0, g.yf, g.Ny + 1)
    prof = 0.5 * (1.0 - np.abs(2.0 * (yl - g.y0) / (g.yf - g.y0) - 1.0))
    f.ocean = host.Ocean(g, np.repeat(prof[None, :], g.Nx + 1, axis=0), 0.0, 0.0)
    f.atmos = host.Atmos(g, 0.0, 0.0, 0.0)
    B = lambda per: host.PeriodicBoundary if per else host.CollisionBoundary
    f.domain = host.Domain(B(per_y)(host.North, g), B(per_y)(host.South, g), B(per_x)(host.East, g),
                           B(per_x)(host.West, g))
    sq = np.sqrt(d["area"])
    f.consts = host.Constants(E=1.5e3 * (sq.mean() + sq.min()))  # examples/uniform_flow.jl:40
    f.walls, f.scale, f.seed, f.npoints = walls, scale, seed, npoints
    return f


def tiled_model(tile, world, walls="collision"):
    """Grid / ocean / atmosphere / domain / constants of `world` tiles side by side in x (weak scaling:
    rank r owns the tile shifted by r L).  walls: 'collision' or 'shear' (periodic east/west)."""
    f = Field()
    L = tile.L
    f.L, f.n, f.walls = L, tile.n * world, walls
    f.grid = host.RegRectilinearGrid(0.0, world * L, 0.0, L, dx=1e4, dy=1e4)
    g = f.grid
    yl = np.linspace(g.y0, g.yf, g.Ny + 1)
    prof = 0.5 * (1.0 - np.abs(2.0 * (yl - g.y0) / (g.yf - g.y0) - 1.0))
    f.ocean = host.Ocean(g, np.repeat(prof[None, :], g.Nx + 1, axis=0), 0.0, 0.0)
    f.atmos = host.Atmos(g, 0.0, 0.0, 0.0)
    EW = host.PeriodicBoundary if walls == "shear" else host.CollisionBoundary
    f.domain = host.Domain(host.CollisionBoundary(host.North, g), host.CollisionBoundary(host.South, g),
                           EW(host.East, g), EW(host.West, g))
    f.consts = tile.consts
    f.floes = None
    return f


def setup_handle(field, lib=None, dt=10, floes=None, **overrides):
    """A handle with grid, fields, domain and the floes of `field` (or `floes`) uploaded."""
    h = host._make_handle(lib, field.consts, dt, None, None, None, **overrides)
    g = field.grid
    h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
    h.set_fields(field.ocean.u, field.ocean.v, field.ocean.hflx_factor, field.atmos.u, field.atmos.v)
    field.domain.push(h)
    fl = field.floes if floes is None else floes
    if fl is not None:
        h.upload_floes(fl)
    return h
