"""Host-side mirror of the reference's API for the floe-interaction hot path.

Same names, argument meaning and error behaviour as Subzero.jl for THIS path only, so the
parity tests read like the reference's own tests (test/test_physical_processes/*.jl):

  Constants, CollisionSettings, CouplingSettings, FloeSettings      simulation.jl:5-18,
                                                                    process_settings.jl:20-229
  RegRectilinearGrid, Ocean, Atmos                                   grids.jl:106-211, oceans.jl, atmos.jl
  North/South/East/West, Open/Periodic/Collision/MovingBoundary,
  TopographyElement, initialize_topography_field, Domain             domain_components/*.jl
  Floe, initialize_floe_field, Model, Simulation                     floe.jl, model.jl, simulation.jl
  add_ghosts, timestep_collisions, timestep_coupling,
  timestep_floe_properties, timestep_sim, run                        the four replaced calls

Everything numerical on the hot path is executed by the C-ABI library passed as `backend`
(default: the CUDA product library; there is no CPU fallback).  Host-only, one-time work
(floe construction: area, centroid, moment, Monte-Carlo points) is plain numpy.
Fracture, ridging, welding, simplification and I/O are out of scope.
"""
import math
import warnings

import numpy as np

from . import capi

# ---------------------------------------------------------------------------------------
# settings (defaults and clamping follow the reference constructors)
# ---------------------------------------------------------------------------------------


class Constants:
    """simulation.jl:5-18"""

    def __init__(self, rho_o=1027.0, rho_a=1.2, Cd_io=3e-3, Cd_ia=1e-3, Cd_ao=1.25e-3, f=1.4e-4,
                 turn_theta=15 * math.pi / 180, L=2.93e5, k=2.14, nu=0.3, mu=0.2, E=6e6):
        self.rho_o, self.rho_a, self.Cd_io, self.Cd_ia, self.Cd_ao = rho_o, rho_a, Cd_io, Cd_ia, Cd_ao
        self.f, self.turn_theta, self.L, self.k, self.nu, self.mu, self.E = f, turn_theta, L, k, nu, mu, E


class CollisionSettings:
    """process_settings.jl:183-229: overlaps are clamped to [0, 1] with a warning."""

    def __init__(self, collisions_on=True, floe_floe_max_overlap=0.55, floe_domain_max_overlap=0.75):
        def clamp(v, name):
            if v > 1:
                warnings.warn("The %s can't be greater than 1. Setting to 1." % name)
                return 1.0
            if v < 0:
                warnings.warn("The %s can't be less than 0. Setting to 0." % name)
                return 0.0
            return float(v)
        self.collisions_on = bool(collisions_on)
        self.floe_floe_max_overlap = clamp(floe_floe_max_overlap, "floe_floe_max_overlap")
        self.floe_domain_max_overlap = clamp(floe_domain_max_overlap, "floe_domain_max_overlap")


class CouplingSettings:
    """process_settings.jl:133-167: Δd < 0 is reset to 0 with a warning."""

    def __init__(self, coupling_on=True, dt=10, dd=1, two_way_coupling_on=False):
        if dd < 0:
            warnings.warn("Δd must be at least 0. Setting to 0.")
            dd = 0
        self.coupling_on, self.dt, self.dd = bool(coupling_on), int(dt), int(dd)
        self.two_way_coupling_on = bool(two_way_coupling_on)


class FloeSettings:
    """process_settings.jl:20-100 (subset read by the hot path) + stress_calculators.jl:81-92"""

    def __init__(self, rho_i=920.0, min_floe_area=0.0, min_floe_height=0.1, max_floe_height=10.0,
                 min_aspect_ratio=0.05, maximum_xi=1e-5, stress_lambda=0.2, mc_npoints=1000):
        self.rho_i, self.min_floe_area, self.min_floe_height = rho_i, min_floe_area, min_floe_height
        self.max_floe_height, self.min_aspect_ratio, self.maximum_xi = max_floe_height, min_aspect_ratio, maximum_xi
        if stress_lambda < 0 or stress_lambda > 1:
            warnings.warn("λ must be between 0 and 1. Resetting to 0.2.")
            stress_lambda = 0.2
        self.stress_lambda = stress_lambda
        self.mc_npoints = mc_npoints


# ---------------------------------------------------------------------------------------
# grid / ocean / atmosphere
# ---------------------------------------------------------------------------------------


class RegRectilinearGrid:
    """grids.jl:106-116,180-211.  Give (Nx, Ny) or (dx, dy)."""

    def __init__(self, x0, xf, y0, yf, dx=None, dy=None, Nx=None, Ny=None):
        if Nx is None:
            Nx = int(math.floor((xf - x0) / dx))
            Ny = int(math.floor((yf - y0) / dy))
            xf, yf = x0 + Nx * dx, y0 + Ny * dy
        self.Nx, self.Ny = int(Nx), int(Ny)
        self.x0, self.xf, self.y0, self.yf = float(x0), float(xf), float(y0), float(yf)
        self.dx, self.dy = (self.xf - self.x0) / self.Nx, (self.yf - self.y0) / self.Ny


def _field(grid, v):
    shape = (grid.Nx + 1, grid.Ny + 1)
    if np.isscalar(v):
        return np.full(shape, float(v))
    v = np.asarray(v, dtype=np.float64)
    if v.shape != shape:
        raise ValueError("field must be (Nx+1, Ny+1) = %s indexed [x, y], got %s" % (shape, v.shape))
    return v.copy()


class Ocean:
    """oceans.jl:74-99,207-222: u, v, temp, hflx_factor on grid lines, indexed [x, y]."""

    def __init__(self, grid, u=0.0, v=0.0, temp=0.0):
        self.u, self.v, self.temp = _field(grid, u), _field(grid, v), _field(grid, temp)
        self.hflx_factor = np.zeros_like(self.u)
        # two-way coupling outputs (oceans.jl:74-99)
        self.tau_x, self.tau_y, self.si_frac = np.zeros_like(self.u), np.zeros_like(self.u), np.zeros_like(self.u)


class Atmos:
    """atmos.jl:4-16"""

    def __init__(self, grid, u=0.0, v=0.0, temp=0.0):
        self.u, self.v, self.temp = _field(grid, u), _field(grid, v), _field(grid, temp)


# ---------------------------------------------------------------------------------------
# domain
# ---------------------------------------------------------------------------------------

North, South, East, West = 0, 1, 2, 3  # wall order of the C ABI (element ids -1..-4)


def _boundary_rect(direction, x0, xf, y0, yf):
    """_boundary_info_from_extent, boundaries.jl:29-33,65-69,102-106,139-143 -> (rect, val)"""
    hx, hy = (xf - x0) / 2, (yf - y0) / 2
    if direction == North:
        return (x0 - hx, xf + hx, yf, yf + hy), yf
    if direction == South:
        return (x0 - hx, xf + hx, y0 - hy, y0), y0
    if direction == East:
        return (xf, xf + hx, y0 - hy, yf + hy), xf
    if direction == West:
        return (x0 - hx, x0, y0 - hy, yf + hy), x0
    raise ValueError("direction must be North, South, East or West")


class _Boundary:
    kind = None

    def __init__(self, direction, grid=None, x0=None, xf=None, y0=None, yf=None, u=0.0, v=0.0):
        if grid is not None:
            x0, xf, y0, yf = grid.x0, grid.xf, grid.y0, grid.yf
        elif None in (x0, xf, y0, yf):
            raise ValueError("To create a boundary, either provide a grid or x0, xf, y0, AND yf.")
        self.direction = direction
        self.rect, self.val = _boundary_rect(direction, x0, xf, y0, yf)
        self.u, self.v = float(u), float(v)


class OpenBoundary(_Boundary):
    kind = capi.BOUNDARY_OPEN


class PeriodicBoundary(_Boundary):
    kind = capi.BOUNDARY_PERIODIC


class CollisionBoundary(_Boundary):
    kind = capi.BOUNDARY_COLLISION


class MovingBoundary(_Boundary):
    kind = capi.BOUNDARY_MOVING

    def __init__(self, direction, grid=None, u=0.0, v=0.0, **kw):
        if u == 0 and v == 0:
            warnings.warn("MovingBoundary velocities are both zero. Boundary will not move.")
        super().__init__(direction, grid, u=u, v=v, **kw)


def ring_area2(r):
    return float(np.sum(r[:-1, 0] * r[1:, 1] - r[:-1, 1] * r[1:, 0]))


def ring_centroid(r):
    c = r[:-1, 0] * r[1:, 1] - r[:-1, 1] * r[1:, 0]
    a = c.sum() / 2.0
    return np.array([((r[:-1, 0] + r[1:, 0]) * c).sum() / (6 * a), ((r[:-1, 1] + r[1:, 1]) * c).sum() / (6 * a)])


def close_ring(coords):
    """valid_ringvec!, floe_utils.jl:10-17: drop repeated neighbours, close the ring."""
    r = np.asarray(coords, dtype=np.float64).reshape(-1, 2)
    keep = np.ones(len(r), dtype=bool)
    keep[:-1] = np.any(r[:-1] != r[1:], axis=1)
    r = r[keep]
    if np.any(r[0] != r[-1]):
        r = np.vstack([r, r[:1]])
    if len(r) <= 3:
        raise ValueError("Polygon needs at least 3 distinct points.")
    return np.ascontiguousarray(r)


def calc_max_radius(ring, cent):
    """floe_utils.jl:301-313"""
    d = ring - np.asarray(cent)
    return float(np.sqrt(np.max(d[:, 0] ** 2 + d[:, 1] ** 2)))


def calc_moment_inertia(ring, cent, height, rho_i=920.0):
    """_calc_moment_inertia, floe_utils.jl:273-298 (restated as written, including the second
    centroid subtraction inside `wi`)."""
    xc, yc = cent
    x = ring[:, 0] - xc
    y = ring[:, 1] - yc
    x1, y1, x2, y2 = x[:-1], y[:-1], x[1:], y[1:]
    wi = (x1 - xc) * (y2 - yc) - (x2 - xc) * (y1 - yc)
    Ixx = np.sum(wi * (y1 ** 2 + y1 * y2 + y2 ** 2)) * (1 / 12)
    Iyy = np.sum(wi * (x1 ** 2 + x1 * x2 + x2 ** 2)) * (1 / 12)
    return abs(Ixx + Iyy) * height * rho_i


class TopographyElement:
    """topography.jl:5-9,66-74"""

    def __init__(self, coords):
        c = coords[0] if np.ndim(coords[0][0]) else coords  # PolyVec -> exterior ring (rmholes!)
        self.ring = close_ring(c)
        self.centroid = ring_centroid(self.ring)
        self.rmax = calc_max_radius(self.ring, self.centroid)


def initialize_topography_field(coords):
    return [TopographyElement(c) for c in coords]


class Domain:
    """domains.jl:4-34: periodic walls must be paired; north > south; east > west."""

    def __init__(self, north, south, east, west, topography=()):
        if (north.kind == capi.BOUNDARY_PERIODIC) != (south.kind == capi.BOUNDARY_PERIODIC) or \
           (east.kind == capi.BOUNDARY_PERIODIC) != (west.kind == capi.BOUNDARY_PERIODIC):
            raise ValueError("If a boundary is periodic, its opposite boundary must also be periodic.")
        if north.val < south.val:
            raise ValueError("North boundary value is less than south boundary value.")
        if east.val < west.val:
            raise ValueError("East boundary value is less than west boundary value.")
        self.north, self.south, self.east, self.west = north, south, east, west
        self.topography = list(topography)

    @property
    def walls(self):
        return [self.north, self.south, self.east, self.west]

    def push(self, h):
        w = self.walls
        h.set_domain([b.kind for b in w], [b.val for b in w], [[b.u, b.v] for b in w],
                     [b.rect for b in w], [t.ring for t in self.topography],
                     np.array([t.centroid for t in self.topography]).reshape(-1, 2),
                     np.array([t.rmax for t in self.topography]))

    def pull(self, h):
        vals, rect = h.get_domain()
        for b, v, r in zip(self.walls, vals, rect):
            b.val, b.rect = float(v), tuple(r)


# ---------------------------------------------------------------------------------------
# floes
# ---------------------------------------------------------------------------------------


def points_in_ring(px, py, ring):
    """Vectorised even-odd test (host-side helper for MC point generation)."""
    inside = np.zeros(px.shape, dtype=bool)
    for k in range(len(ring) - 1):
        ax, ay = ring[k]
        bx, by = ring[k + 1]
        if ay == by:
            continue
        cond = (ay > py) != (by > py)
        xi = ax + (py - ay) / (by - ay) * (bx - ax)
        inside ^= cond & (px < xi)
    return inside


def mc_points(ring, centroid, npoints, rng):
    r = ring - centroid
    xmin, ymin = r.min(axis=0)
    xmax, ymax = r.max(axis=0)
    px = xmin + (xmax - xmin) * rng.random(npoints)
    py = ymin + (ymax - ymin) * rng.random(npoints)
    m = points_in_ring(px, py, r)
    return px[m], py[m]


class Floe:
    """floe.jl:24-77,144-242.  Host-side record of one floe (used to build a FloeField)."""

    def __init__(self, coords, hmean, dh=0.0, floe_settings=None, rng=None, **kw):
        fs = floe_settings or FloeSettings()
        rng = rng or np.random.default_rng(0)
        c = coords[0] if np.ndim(coords[0][0]) else coords
        self.ring = close_ring(c)
        self.centroid = ring_centroid(self.ring)
        h = hmean + (-1) ** int(rng.integers(0, 2)) * rng.random() * dh
        self.height = min(max(h, fs.min_floe_height), fs.max_floe_height)
        self.area = abs(ring_area2(self.ring)) / 2.0
        self.mass = self.area * self.height * fs.rho_i
        self.moment = calc_moment_inertia(self.ring, self.centroid, self.height, fs.rho_i)
        self.rmax = calc_max_radius(self.ring, self.centroid)
        self.x_subfloe_points, self.y_subfloe_points = mc_points(self.ring, self.centroid, fs.mc_npoints, rng)
        self.status_tag = capi.STATUS_ACTIVE if len(self.x_subfloe_points) else capi.STATUS_REMOVE
        self.alpha = self.u = self.v = self.xi = 0.0
        self.id = 0
        self.ghost_id = 0
        for k, v in kw.items():
            setattr(self, k, v)


class FloeField(capi.FloeArrays):
    """StructArray{Floe} stand-in: the SoA the C ABI consumes plus per-floe `interactions`,
    `num_inters`, `fuse_idx` and `ghosts` views rebuilt after each collision step."""

    def __init__(self, floes):
        n = len(floes)
        super().__init__(n)
        offs, moffs = [0], [0]
        for i, f in enumerate(floes):
            self.centroid_x[i], self.centroid_y[i] = f.centroid
            for name in ("height", "area", "mass", "rmax", "moment", "alpha", "u", "v", "xi"):
                getattr(self, name)[i] = getattr(f, name)
            for name in ("fxOA", "fyOA", "trqOA", "hflx_factor", "overarea", "collision_trq", "p_dxdt",
                         "p_dydt", "p_dudt", "p_dvdt", "p_dxidt", "p_dalphadt"):
                if hasattr(f, name):
                    getattr(self, name)[i] = getattr(f, name)
            self.status_tag[i] = f.status_tag
            self.id[i] = f.id if f.id else i + 1
            self.ghost_id[i] = f.ghost_id
            offs.append(offs[-1] + len(f.ring))
            moffs.append(moffs[-1] + len(f.x_subfloe_points))
        self.vert_offsets = np.array(offs, dtype=np.int64)
        self.vert_xy = np.concatenate([f.ring for f in floes]) if n else np.zeros((0, 2))
        self.mc_offsets = np.array(moffs, dtype=np.int64)
        self.mc_x = np.concatenate([np.asarray(f.x_subfloe_points, dtype=np.float64) for f in floes]) if n else np.zeros(0)
        self.mc_y = np.concatenate([np.asarray(f.y_subfloe_points, dtype=np.float64) for f in floes]) if n else np.zeros(0)
        self.interactions = [np.zeros((0, 7)) for _ in range(n)]
        self.num_inters = np.zeros(n, dtype=np.int64)
        self.fuse_idx = [[] for _ in range(n)]
        self.warnings = np.zeros(n, dtype=np.uint32)

    def __len__(self):
        return self.n


def initialize_floe_field(coords, domain=None, hmean=0.25, dh=0.0, floe_settings=None, rng=None, **kw):
    """floe.jl:361-411 (from coordinates): ids are 1..n in order."""
    rng = rng or np.random.default_rng(0)
    floes = []
    for k, c in enumerate(coords):
        f = Floe(c, hmean, dh, floe_settings=floe_settings, rng=rng, **kw)
        f.id = k + 1
        floes.append(f)
    return FloeField(floes)


class Model:
    """model.jl:47-120"""

    def __init__(self, grid, ocean, atmos, domain, floes):
        for b, lo, hi in ((domain.north, grid.y0, grid.yf), (domain.south, grid.y0, grid.yf),
                          (domain.east, grid.x0, grid.xf), (domain.west, grid.x0, grid.xf)):
            if not (lo <= b.val <= hi):
                raise ValueError("Domain does not fit within grid.")
        self.grid, self.ocean, self.atmos, self.domain, self.floes = grid, ocean, atmos, domain, floes


# ---------------------------------------------------------------------------------------
# backend plumbing
# ---------------------------------------------------------------------------------------


def _make_handle(backend, consts, dt, collision_settings=None, coupling_settings=None, floe_settings=None,
                 **overrides):
    lib = backend or capi.product()
    cs = collision_settings or CollisionSettings()
    cp = coupling_settings or CouplingSettings()
    fs = floe_settings or FloeSettings()
    cfg = lib.default_config_struct()
    for name in ("rho_o", "rho_a", "Cd_io", "Cd_ia", "Cd_ao", "f", "turn_theta", "L", "k", "nu", "mu", "E"):
        setattr(cfg, name, getattr(consts, name))
    cfg.floe_floe_max_overlap, cfg.floe_domain_max_overlap = cs.floe_floe_max_overlap, cs.floe_domain_max_overlap
    cfg.rho_i, cfg.max_floe_height, cfg.maximum_xi, cfg.stress_lambda = fs.rho_i, fs.max_floe_height, fs.maximum_xi, fs.stress_lambda
    cfg.coupling_dd, cfg.two_way_coupling_on, cfg.dt = cp.dd, int(cp.two_way_coupling_on), int(dt)
    return capi.Handle(lib, cfg, **overrides)


def _default_grid_for(domain):
    return RegRectilinearGrid(domain.west.val, domain.east.val, domain.south.val, domain.north.val, Nx=1, Ny=1)


def _pull_collision_results(h, floes):
    fa = h.download_floes()
    floes.adopt(fa)
    offs, rows = h.interactions()
    n = fa.n
    floes.interactions = [rows[offs[i]:offs[i + 1]].copy() for i in range(n)]
    floes.num_inters = np.diff(offs)
    # status.fuse_idx, collisions.jl:368 then the serial propagation :799-806
    fuse = h.pairs(3)
    fidx = [[] for _ in range(n)]
    for i, j in fuse:
        fidx[i - 1].append(int(j))
    for i in range(n):
        if floes.status_tag[i] == capi.STATUS_FUSE:
            for idx in list(fidx[i]):
                fidx[idx - 1].append(i + 1)
    floes.fuse_idx = fidx


def add_ghosts(floes, domain, backend=None):
    """add_ghosts!(floes, domain), collisions.jl:1060-1174"""
    h = _make_handle(backend, Constants(), 10)
    g = _default_grid_for(domain)
    h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
    domain.push(h)
    h.upload_floes(floes)
    h.add_ghosts()
    floes.adopt(h.download_floes())
    n = floes.n
    floes.interactions = [np.zeros((0, 7)) for _ in range(n)]
    floes.num_inters = np.zeros(n, dtype=np.int64)
    floes.fuse_idx = [[] for _ in range(n)]
    h.close()
    return floes


def timestep_collisions(floes, n_init_floes, domain, consts, dt, collision_settings=None, spinlock=None,
                        backend=None, **overrides):
    """timestep_collisions!(floes, n_init, domain, consts, Δt, collision_settings, spinlock),
    collisions.jl:734-864.  `spinlock` is accepted and ignored (the device path is lock-free
    and deterministic)."""
    h = _make_handle(backend, consts, dt, collision_settings, **overrides)
    g = _default_grid_for(domain)
    h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
    domain.push(h)
    floes.n_init = int(n_init_floes)
    h.upload_floes(floes)
    h.step_collisions()
    _pull_collision_results(h, floes)
    domain.pull(h)
    h.last_counts = h.counts()
    floes.last_counts = h.last_counts
    h.close()
    return floes


def floe_floe_interaction(ifloe, i, jfloe, j, consts, dt, max_overlap, backend=None):
    """floe_floe_interaction!, collisions.jl:347-408, on a 2-floe list far from any wall.
    Returns the FloeField [ifloe, jfloe]; rows of ifloe are `.interactions[0]`."""
    assert (i, j) == (1, 2)
    ff = FloeField([ifloe, jfloe])
    ff.id[:] = (1, 2)
    big = 1e9
    grid = RegRectilinearGrid(-big, big, -big, big, Nx=1, Ny=1)
    dom = Domain(PeriodicBoundary(North, grid), PeriodicBoundary(South, grid),
                 PeriodicBoundary(East, grid), PeriodicBoundary(West, grid))
    cs = CollisionSettings(floe_floe_max_overlap=max_overlap)
    return timestep_collisions(ff, 2, dom, consts, dt, cs, backend=backend)


def floe_domain_interaction(floe, domain, consts, dt, max_overlap, backend=None):
    """floe_domain_interaction!, collisions.jl:594-662, for a single floe."""
    ff = FloeField([floe])
    cs = CollisionSettings(floe_domain_max_overlap=max_overlap)
    return timestep_collisions(ff, 1, domain, consts, dt, cs, backend=backend)


def timestep_coupling(model, dt, consts, coupling_settings=None, floe_settings=None, backend=None):
    """timestep_coupling!(model, Δt, consts, coupling_settings, floe_settings), coupling.jl:1705-1738"""
    h = _make_handle(backend, consts, dt, None, coupling_settings, floe_settings)
    g = model.grid
    h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
    h.set_fields(model.ocean.u, model.ocean.v, model.ocean.hflx_factor, model.atmos.u, model.atmos.v)
    h.set_temperatures(model.ocean.temp, model.atmos.temp)
    model.domain.push(h)
    h.upload_floes(model.floes)
    h.step_coupling()
    model.floes.adopt(h.download_floes())
    if coupling_settings is not None and coupling_settings.two_way_coupling_on:
        oc = model.ocean
        oc.tau_x, oc.tau_y, oc.si_frac, oc.hflx_factor = h.ocean_fields()
    model.cell_floes = h.cell_floes()  # grid.floe_locations / ocean.scells as (cell, floe, values)
    h.close()
    return model


def timestep_floe_properties(floes, tstep, dt, floe_settings=None, consts=None, backend=None):
    """timestep_floe_properties!(floes, tstep, Δt, floe_settings), update_floe.jl:469-551.
    Uses `floes.interactions` (rows 1:num_inters) for calc_stress!."""
    h = _make_handle(backend, consts or Constants(), dt, None, None, floe_settings)
    h.upload_floes(floes)
    h.set_interactions(floes.interactions)
    h.step_floe_properties(tstep)
    keep_rows = floes.interactions
    floes.adopt(h.download_floes())
    floes.interactions = keep_rows
    floes.warnings = h.warnings()
    h.close()
    return floes


# ---------------------------------------------------------------------------------------
# Simulation: persistent device-resident state
# ---------------------------------------------------------------------------------------


class GridOutputWriter:
    """GridOutputWriter(Δtout, grid, dims; outputs) of output.jl:560-640: the averaging grid (xg, yg: dims[1] x
    dims[2] cells over the model grid's extent) and the list of outputs; `data` is [dims[1], dims[2], n_outputs]."""

    def __init__(self, dtout, grid, dims, outputs=None):
        self.dtout = int(dtout)
        self.outputs = list(outputs) if outputs is not None else list(capi.GRID_OUTPUTS)
        for o in self.outputs:
            if o not in capi.GRID_OUTPUTS:
                raise ValueError("unknown grid output %r" % (o,))
        self.xg = np.linspace(grid.x0, grid.xf, int(dims[0]) + 1)
        self.yg = np.linspace(grid.y0, grid.yf, int(dims[1]) + 1)
        self.data = np.zeros((int(dims[0]), int(dims[1]), len(self.outputs)))


def calc_eulerian_data(floes, topography, writer, domain=None, backend=None):
    """calc_eulerian_data!(floes, topography, writer), output.jl:794-919: floe data averaged on the writer's grid.
    Topography (cell polygons minus topography) is outside the device path's scope."""
    if topography is not None and len(topography) > 0:
        raise capi.SubzeroError(-4, "calc_eulerian_data with topography stays on the host")
    h = _make_handle(backend, Constants(), 10)
    if domain is None:
        pad = 1e6
        x0, xf, y0, yf = writer.xg[0] - pad, writer.xg[-1] + pad, writer.yg[0] - pad, writer.yg[-1] + pad
        domain = Domain(*[OpenBoundary(d, x0=x0, xf=xf, y0=y0, yf=yf) for d in (North, South, East, West)])
    g = _default_grid_for(domain)
    h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
    domain.push(h)
    h.upload_floes(floes)
    kinds = [capi.GRID_OUTPUTS.index(o) for o in writer.outputs]
    writer.data[...] = h.eulerian_data(writer.xg, writer.yg, kinds)
    h.close()
    return writer.data


class Simulation:
    """simulation.jl:49-81.  Only the fields the hot path reads are kept; the host processes
    (fracture, ridging, welding, simplification, writers) are out of scope and therefore OFF.
    The floe store stays on the device between steps; `sync_host()` downloads it."""

    def __init__(self, model, consts=None, dt=10, n_dt=7500, collision_settings=None, coupling_settings=None,
                 floe_settings=None, verbose=False, name="sim", backend=None, **overrides):
        self.model = model
        self.consts = consts or Constants()
        self.dt, self.n_dt, self.verbose, self.name = int(dt), int(n_dt), verbose, name
        self.collision_settings = collision_settings or CollisionSettings()
        self.coupling_settings = coupling_settings or CouplingSettings()
        self.floe_settings = floe_settings or FloeSettings()
        self.h = _make_handle(backend, self.consts, self.dt, self.collision_settings, self.coupling_settings,
                              self.floe_settings, **overrides)
        g = model.grid
        self.h.set_grid(g.Nx, g.Ny, g.x0, g.xf, g.y0, g.yf)
        self.h.set_fields(model.ocean.u, model.ocean.v, model.ocean.hflx_factor, model.atmos.u, model.atmos.v)
        self.h.set_temperatures(model.ocean.temp, model.atmos.temp)
        model.domain.push(self.h)
        self.h.upload_floes(model.floes)
        self._resident = True

    def upload(self):
        self.h.upload_floes(self.model.floes)

    def sync_host(self):
        self.model.floes.adopt(self.h.download_floes())
        self.model.domain.pull(self.h)
        if self.coupling_settings.two_way_coupling_on:
            oc = self.model.ocean
            oc.tau_x, oc.tau_y, oc.si_frac, oc.hflx_factor = self.h.ocean_fields()
        return self.model.floes

    def close(self):
        self.h.close()


def timestep_sim(sim, tstep, start_tstep=0):
    """timestep_sim!(sim, tstep), simulation.jl:94-220 restricted to the replaced calls:
    add_ghosts! -> timestep_collisions! -> ghost removal -> timestep_coupling! (every
    coupling Δt steps) -> timestep_floe_properties!."""
    if sim.verbose and tstep % 50 == 0:
        print(tstep, " timesteps")
    cp = sim.coupling_settings
    do_cpl = cp.coupling_on and tstep % cp.dt == 0
    if sim.collision_settings.collisions_on:
        sim.h.step(tstep, do_cpl)
    else:
        # add_ghosts! runs whether or not collisions are on (simulation.jl:100-102): it is what wraps a parent that
        # drifted out of a periodic domain back inside (collisions.jl:943-949); the ghosts are deleted again at :138-144
        sim.h.add_ghosts()
        sim.h.remove_ghosts()
        if do_cpl:
            sim.h.step_coupling()
        sim.h.step_floe_properties(tstep)


def run(sim, start_tstep=0):
    """run!(sim), simulation.jl:287-297"""
    tstep = start_tstep
    while tstep <= start_tstep + sim.n_dt:
        timestep_sim(sim, tstep, start_tstep)
        tstep += 1
    return sim.sync_host()
