"""Test fields beyond the synthetic Voronoi generator: the reference's own floe shapes
(test/inputs/floe_shapes.jld2 -> tests/golden/floe_shapes.npz: 462 non-convex rings of 7-591
vertices) placed as a real field, as examples/many_floes.jl and test_conservation.jl:159-187 do."""
import os

import numpy as np

from subzero_jl_b200 import capi, host, synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def fixture_shape_field(scale=1.04, walls="collision", npoints=100, nmax=None, seed=5):
    d = np.load(os.path.join(GOLD, "floe_shapes.npz"))
    offs, xy = d["offsets"], d["xy"]
    n = len(offs) - 1 if nmax is None else min(nmax, len(offs) - 1)
    rng = np.random.default_rng(seed)
    floes = []
    for i in range(n):
        ring = host.close_ring(xy[offs[i]:offs[i + 1]])
        c = host.ring_centroid(ring)
        ring = c + scale * (ring - c)
        ring[-1] = ring[0]
        fs = host.FloeSettings(mc_npoints=npoints)
        f = host.Floe([ring.tolist()], 0.25, 0.0, floe_settings=fs, rng=rng)
        f.u, f.v = float(rng.uniform(-0.1, 0.1)), float(rng.uniform(-0.1, 0.1))
        f.xi = float(rng.uniform(-1e-6, 1e-6))
        f.id = i + 1
        floes.append(f)
    fld = synth.Field()
    fld.floes = host.FloeField(floes)
    fld.n = n
    L = 7e4
    fld.L = 2 * L
    fld.grid = host.RegRectilinearGrid(-L, L, -L, L, dx=1e4, dy=1e4)
    g = fld.grid
    yl = np.linspace(g.y0, g.yf, g.Ny + 1)
    prof = 0.5 * (1.0 - np.abs(2.0 * (yl - g.y0) / (g.yf - g.y0) - 1.0))
    fld.ocean = host.Ocean(g, np.repeat(prof[None, :], g.Nx + 1, axis=0), 0.05, 0.0)
    fld.atmos = host.Atmos(g, 2.0, -1.0, 0.0)
    per_x = walls in ("periodic", "shear")
    per_y = walls == "periodic"
    B = lambda per: host.PeriodicBoundary if per else host.CollisionBoundary
    fld.domain = host.Domain(B(per_y)(host.North, g), B(per_y)(host.South, g), B(per_x)(host.East, g),
                             B(per_x)(host.West, g))
    sq = np.sqrt(fld.floes.area)
    fld.consts = host.Constants(E=1.5e3 * (sq.mean() + sq.min()))
    return fld


def perturb_state(fa, seed=3):
    """Non-trivial dynamic state so that every term of the state update is exercised."""
    rng = np.random.default_rng(seed)
    n = fa.n
    fa.xi = rng.uniform(-5e-6, 5e-6, n)
    fa.alpha = rng.uniform(-0.3, 0.3, n)
    fa.p_dxdt = fa.u + rng.uniform(-0.01, 0.01, n)
    fa.p_dydt = fa.v + rng.uniform(-0.01, 0.01, n)
    fa.p_dudt = rng.uniform(-1e-5, 1e-5, n)
    fa.p_dvdt = rng.uniform(-1e-5, 1e-5, n)
    fa.p_dxidt = rng.uniform(-1e-10, 1e-10, n)
    fa.p_dalphadt = fa.xi + rng.uniform(-1e-7, 1e-7, n)
    fa.stress_accum = rng.uniform(-1e3, 1e3, (n, 4))
    fa.overarea = rng.uniform(0, 1e3, n)
    fa.hflx_factor = rng.uniform(-1e-4, 1e-4, n)
    return fa
