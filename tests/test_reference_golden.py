"""Golden vectors transcribed from the reference's own tests for the hot path.

Every case runs twice: against the CPU oracle (`-m "not gpu"`, pins the oracle) and against
the CUDA product through the C ABI (`-m gpu`).  Sources:
  /root/reference/test/test_physical_processes/test_collisions.jl:39-363
  /root/reference/test/test_physical_processes/test_coupling.jl:464-640
  /root/reference/test/test_physical_processes/test_update_floe.jl:2-42
  /root/reference/test/test_floe_utils.jl:65-71
Tolerances are the reference's own `atol`s.
"""
import json
import os

import numpy as np
import pytest

from subzero_jl_b200 import capi, host
from subzero_jl_b200.host import (CollisionBoundary, Constants, Domain, East, Floe, North, OpenBoundary,
                                  PeriodicBoundary, RegRectilinearGrid, South, West)

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FLOEIDX, XFORCE, YFORCE, XPOINT, YPOINT, TORQUE, OVERLAP = range(7)


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def lib(request):
    if request.param == "oracle":
        return request.getfixturevalue("oracle_lib")
    return request.getfixturevalue("product_lib")


def translate(coords, dx, dy):
    return [[[p[0] + dx, p[1] + dy] for p in coords[0]]]


Lx = Ly = 1e5
GRID = RegRectilinearGrid(-Lx, Lx, -Ly, Ly, dx=1e4, dy=1e4)


def domains():
    pb = lambda d: PeriodicBoundary(d, GRID)
    cb = lambda d: CollisionBoundary(d, GRID)
    ob = lambda d: OpenBoundary(d, GRID)
    topos = host.initialize_topography_field(
        [[[[1e4, 0.0], [0.0, 1e4], [1e4, 2e4], [2e4, 1e4], [1e4, 0.0]]]])
    return dict(
        topo=Domain(pb(North), pb(South), cb(East), ob(West), topography=topos),
        collision=Domain(cb(North), cb(South), cb(East), cb(West)),
        open=Domain(ob(North), ob(South), ob(East), ob(West)),
        ew=Domain(ob(North), ob(South), pb(East), pb(West)),
        ns=Domain(pb(North), pb(South), ob(East), ob(West)),
        double=Domain(pb(North), pb(South), pb(East), pb(West)),
    )


# --------------------------------------------------------------------------------------
# test_collisions.jl:39-103  "Floe-Floe Interactions"
# --------------------------------------------------------------------------------------
TRI = [[[0.0, 0.0], [1e4, 3e4], [2e4, 0], [0.0, 0.0]]]
CORNER_RECT = [[[0.0, 2.5e4], [0.0, 2.9e4], [2e4, 2.9e4], [2e4, 2.5e4], [0.0, 2.5e4]]]
MIDDLE_RECT = [[[1.8e4, 2.7e4], [1.8e4, 2.8e4], [2.1e4, 2.8e4], [2.1e4, 2.7e4], [1.8e4, 2.7e4]]]
CSHAPE = [[[0.5e4, 2.7e4], [0.5e4, 3.5e4], [1.5e4, 3.5e4], [1.5e4, 2.7e4], [1.25e4, 2.7e4],
           [1.25e4, 3e4], [1e4, 3e4], [1e4, 2.7e4], [0.5e4, 2.7e4]]]


def torque_of(ff, i, k):
    """calc_torque!, collisions.jl:673-686 on row k of floe i."""
    r = ff.interactions[i][k]
    return (r[XPOINT] - ff.centroid_x[i]) * r[YFORCE] - (r[YPOINT] - ff.centroid_y[i]) * r[XFORCE]


def test_tri_tip_into_rectangle(lib):  # :50-62
    tri, rect = Floe(TRI, 0.25), Floe(CORNER_RECT, 0.25)
    tri.u, rect.v = 0.1, -0.1
    ff = host.floe_floe_interaction(tri, 1, rect, 2, Constants(), 10, 0.55, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 1 and r[0, FLOEIDX] == 2
    assert r[0, XFORCE] == pytest.approx(-64613382.47, abs=1e-2)
    assert r[0, YFORCE] == pytest.approx(-521498991.51, abs=1e-2)
    assert r[0, XPOINT] == pytest.approx(10000.00, abs=1e-2)
    assert r[0, YPOINT] == pytest.approx(26555.55, abs=1e-2)
    assert r[0, OVERLAP] == pytest.approx(8000000, abs=1e-2)
    assert ff.status_tag[0] != capi.STATUS_FUSE and ff.status_tag[1] != capi.STATUS_FUSE
    assert ff.fuse_idx[0] == []
    assert r[0, TORQUE] == pytest.approx(1069710443203.99, abs=1e-2)
    assert torque_of(ff, 0, 0) == pytest.approx(1069710443203.99, abs=1e-2)
    # mirrored row on the rectangle (collisions.jl:808-827): equal and opposite
    m = ff.interactions[1]
    assert len(m) == 1 and m[0, FLOEIDX] == 1
    assert m[0, XFORCE] == -r[0, XFORCE] and m[0, YFORCE] == -r[0, YFORCE]


def test_cshape_two_regions(lib):  # :64-81, pins the region order of the clipper
    c, rect = Floe(CSHAPE, 0.25), Floe(CORNER_RECT, 0.25)
    c.u, rect.v = 0.3, -0.1
    ff = host.floe_floe_interaction(c, 1, rect, 2, Constants(), 10, 0.55, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 2
    assert r[0, XFORCE] == pytest.approx(-163013665.41, abs=1e-2)
    assert r[1, XFORCE] == pytest.approx(-81506832.70, abs=1e-2)
    assert r[0, YFORCE] == pytest.approx(804819565.60, abs=1e-2)
    assert r[1, YFORCE] == pytest.approx(402409782.80, abs=1e-2)
    assert r[0, XPOINT] == pytest.approx(7500.00, abs=1e-2)
    assert r[1, XPOINT] == pytest.approx(13750.00, abs=1e-2)
    assert r[0, YPOINT] == pytest.approx(28000.00, abs=1e-2)
    assert r[1, YPOINT] == pytest.approx(28000.00, abs=1e-2)
    assert r[0, OVERLAP] == pytest.approx(10000000, abs=1e-2)
    assert r[1, OVERLAP] == pytest.approx(5000000, abs=1e-2)
    assert r[0, TORQUE] == pytest.approx(-2439177121266.03, abs=1e-2)
    assert r[1, TORQUE] == pytest.approx(1295472581868.05, abs=1e-2)


def test_overlap_over_55_percent_fuses(lib):  # :83-96
    a, b = Floe(CORNER_RECT, 0.25), Floe(translate(CORNER_RECT, 0.5e4, 0.0), 0.25)
    a.v = b.v = -0.1
    ff = host.floe_floe_interaction(a, 1, b, 2, Constants(), 10, 0.55, backend=lib)
    assert ff.status_tag[0] == capi.STATUS_FUSE
    assert ff.fuse_idx[0][0] == 2
    assert len(ff.interactions[0]) == 0
    a, b = Floe(CORNER_RECT, 0.25), Floe(MIDDLE_RECT, 0.25)
    a.v = -0.1
    ff = host.floe_floe_interaction(a, 1, b, 2, Constants(), 10, 0.55, backend=lib)
    assert ff.status_tag[0] == capi.STATUS_FUSE and ff.fuse_idx[0][0] == 2
    # serial propagation, collisions.jl:799-806: partner tagged, duplicate entry on the first floe
    assert ff.status_tag[1] == capi.STATUS_FUSE
    assert ff.fuse_idx[1] == [1] and ff.fuse_idx[0] == [2, 2]


def test_sliver_overlap_gives_no_force(lib):  # :98-102
    a, b = Floe(CORNER_RECT, 0.25), Floe(translate(CORNER_RECT, 1.9999999e4, 0.0), 0.25)
    a.v = b.v = -0.1
    ff = host.floe_floe_interaction(b, 1, a, 2, Constants(), 10, 0.55, backend=lib)
    assert len(ff.interactions[0]) == 0


# --------------------------------------------------------------------------------------
# test_collisions.jl:105-188  "Floe Boundary Interactions"
# --------------------------------------------------------------------------------------
EAST_SMALL = [[[9.5e4, 0.0], [9e4, 0.5e4], [10e4, 2.5e4], [10.05e4, 2e4], [9.5e4, 0.0]]]
EAST_LARGE = [[[9e4, -7e4], [9e4, -5e4], [1.4e5, -5e4], [1.4e5, -7e4], [9e4, -7e4]]]
WEST = [[[-9.75e4, 7e4], [-9.75e4, 5e4], [-10.05e4, 5e4], [-10.05e4, 7e4], [-9.75e4, 7e4]]]
NORTH_C = [[[5e4, 9.75e4], [5e4, 10.05e4], [7e4, 10.05e4], [7e4, 9.75e4], [5e4, 9.75e4]]]
CSHAPE_E = [[[9.5e4, 7e4], [9.5e4, 9e4], [1.05e5, 9e4], [1.05e5, 8.5e4], [9.9e4, 8.5e4], [9.9e4, 8e4],
             [1.05e5, 8e4], [1.05e5, 7e4], [9.5e4, 7e4]]]
TOPO_OVERLAP = [[[-0.5e4, 0.0], [-0.5e4, 0.75e4], [0.5e4, 0.75e4], [0.5e4, 0.0], [-0.5e4, 0.0]]]
CORNER = [[[9.5e4, 7e4], [9e4, 7.5e4], [10e4, 1.05e5], [10.05e4, 9.5e4], [9.5e4, 7e4]]]


def test_floe_east_collision_wall_one_region(lib):  # :124-133
    f = Floe(EAST_SMALL, 0.25)
    f.u, f.v = 0.5, 0.25
    ff = host.floe_domain_interaction(f, domains()["topo"], Constants(), 10, 0.75, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 1 and r[0, FLOEIDX] == -3
    assert r[0, XFORCE] == pytest.approx(-311304795.629, abs=1e-3)
    assert r[0, YFORCE] == pytest.approx(-23618874.648, abs=1e-3)
    assert r[0, OVERLAP] == pytest.approx(1704545.454, abs=1e-3)
    assert r[0, XPOINT] == pytest.approx(100166.666, abs=1e-3)
    assert r[0, YPOINT] == pytest.approx(21060.606, abs=1e-3)


def test_cshape_east_wall_two_regions(lib):  # :135-150
    f = Floe(CSHAPE_E, 0.25)
    f.v = -0.1
    ff = host.floe_domain_interaction(f, domains()["topo"], Constants(), 10, 0.75, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 2 and r[0, FLOEIDX] == -3 and r[1, FLOEIDX] == -3
    assert r[0, XFORCE] == pytest.approx(-2876118708.17, abs=1e-2)
    assert r[1, XFORCE] == pytest.approx(-5752237416.35, abs=1e-2)
    assert r[0, YFORCE] == pytest.approx(575223741.63, abs=1e-2)
    assert r[1, YFORCE] == pytest.approx(1150447483.27, abs=1e-2)
    assert r[0, XPOINT] == pytest.approx(102500, abs=1e-2)
    assert r[1, XPOINT] == pytest.approx(102500, abs=1e-2)
    assert r[0, YPOINT] == pytest.approx(87500, abs=1e-2)
    assert r[1, YPOINT] == pytest.approx(75000, abs=1e-2)
    assert r[0, OVERLAP] == pytest.approx(25000000, abs=1e-2)
    assert r[1, OVERLAP] == pytest.approx(50000000, abs=1e-2)


def test_wall_overlap_removal_and_other_kinds(lib):  # :152-187
    d = domains()
    f = Floe(EAST_LARGE, 0.25)
    f.u, f.v = -0.4, 0.2
    ff = host.floe_domain_interaction(f, d["topo"], Constants(), 10, 0.75, backend=lib)
    assert len(ff.interactions[0]) == 0 and ff.status_tag[0] == capi.STATUS_REMOVE
    f = Floe(EAST_LARGE, 0.25)
    f.u, f.v = -0.4, 0.2
    ff = host.floe_domain_interaction(f, d["topo"], Constants(), 10, 1.0, backend=lib)
    assert len(ff.interactions[0]) > 0 and ff.num_inters[0] > 0
    ff = host.floe_domain_interaction(Floe(WEST, 0.25), d["topo"], Constants(), 10, 0.75, backend=lib)
    assert ff.status_tag[0] == capi.STATUS_REMOVE  # open boundary
    ff = host.floe_domain_interaction(Floe(NORTH_C, 0.25), d["topo"], Constants(), 10, 0.75, backend=lib)
    assert ff.status_tag[0] == capi.STATUS_ACTIVE and len(ff.interactions[0]) == 0  # periodic wall
    ff = host.floe_domain_interaction(Floe(TOPO_OVERLAP, 0.25), d["topo"], Constants(), 10, 0.75, backend=lib)
    r = ff.interactions[0]
    assert r[0, XFORCE] < 0 and r[0, YFORCE] < 0 and r[0, FLOEIDX] == -5
    ff = host.floe_domain_interaction(Floe(CORNER, 0.25), d["collision"], Constants(), 10, 0.75, backend=lib)
    r = ff.interactions[0]
    assert len(r) >= 2 and np.all(r[:, XFORCE] <= 0) and np.all(r[:, YFORCE] <= 0)


# --------------------------------------------------------------------------------------
# test_collisions.jl:190-259  "Add Ghosts"
# --------------------------------------------------------------------------------------
C1 = [[[9.9e4, 9.9e4], [9.9e4, 1.02e5], [1.02e5, 1.02e5], [1.02e5, 9.9e4], [9.9e4, 9.9e4]]]
C2 = [[[-1.01e5, 7e4], [-1.01e5, 8e4], [-8e4, 8e4], [-8e4, 7e4], [-1.01e5, 7e4]]]
C3 = [[[-2e4, 9.5e4], [-2e4, 1.1e5], [-1e4, 1.1e5], [-1e4, 9.5e4], [-2e4, 9.5e4]]]
C4 = [[[0.0, 0.0], [0.0, 2e4], [2e4, 2e4], [2e4, 0.0], [0.0, 0.0]]]
COORD_LIST = [C1, C2, C3, C4]


def ring_eq(ff, i, coords):
    return np.array_equal(ff.coords(i), np.asarray(coords[0], dtype=np.float64))


def test_add_ghosts_nonperiodic(lib):  # :204-207
    ff = host.initialize_floe_field(COORD_LIST, hmean=0.5)
    host.add_ghosts(ff, domains()["open"], backend=lib)
    assert ff.n == 4 and all(ring_eq(ff, i, COORD_LIST[i]) for i in range(4))


def test_add_ghosts_east_west(lib):  # :209-222
    ff = host.initialize_floe_field(COORD_LIST, hmean=0.5)
    host.add_ghosts(ff, domains()["ew"], backend=lib)
    assert -1e5 < ff.centroid_x[0] < 1e5 and -1e5 < ff.centroid_y[1] < 1e5
    assert ring_eq(ff, 0, translate(C1, -2e5, 0.0))
    assert all(ring_eq(ff, i, COORD_LIST[i]) for i in (1, 2, 3))
    assert ring_eq(ff, 4, C1)
    assert ring_eq(ff, 5, translate(C2, 2e5, 0.0))
    assert list(ff.id) == [1, 2, 3, 4, 1, 2]
    assert list(ff.ghost_id) == [0, 0, 0, 0, 1, 1]
    assert ff.ghosts(0) == [5] and ff.ghosts(1) == [6]
    assert all(ff.ghosts(i) == [] for i in range(2, 6))


def test_add_ghosts_north_south(lib):  # :224-238
    ff = host.initialize_floe_field(COORD_LIST, hmean=0.5)
    host.add_ghosts(ff, domains()["ns"], backend=lib)
    assert -1e5 < ff.centroid_y[0] < 1e5 and -1e5 < ff.centroid_y[2] < 1e5
    assert ring_eq(ff, 0, translate(C1, 0.0, -2e5)) and ring_eq(ff, 2, translate(C3, 0.0, -2e5))
    assert ring_eq(ff, 1, C2) and ring_eq(ff, 3, C4)
    assert ring_eq(ff, 4, C1) and ring_eq(ff, 5, C3)
    assert list(ff.id) == [1, 2, 3, 4, 1, 3]
    assert list(ff.ghost_id) == [0, 0, 0, 0, 1, 1]
    assert ff.ghosts(0) == [5] and ff.ghosts(2) == [6]
    assert all(ff.ghosts(i) == [] for i in (1, 3, 4, 5))


def test_add_ghosts_doubly_periodic(lib):  # :240-258
    ff = host.initialize_floe_field(COORD_LIST, hmean=0.5)
    host.add_ghosts(ff, domains()["double"], backend=lib)
    assert -1e5 < ff.centroid_x[0] < 1e5 and -1e5 < ff.centroid_y[0] < 1e5
    assert ring_eq(ff, 0, translate(C1, -2e5, -2e5))
    assert ring_eq(ff, 2, translate(C3, 0.0, -2e5))
    assert ring_eq(ff, 1, C2) and ring_eq(ff, 3, C4)
    assert ring_eq(ff, 4, C1)
    assert ring_eq(ff, 5, translate(C2, 2e5, 0.0))
    assert ring_eq(ff, 6, translate(C1, 0.0, -2e5))
    assert ring_eq(ff, 7, translate(C1, -2e5, 0.0))
    assert ring_eq(ff, 8, C3)
    assert list(ff.id) == [1, 2, 3, 4, 1, 2, 1, 1, 3]
    assert list(ff.ghost_id) == [0, 0, 0, 0, 1, 1, 2, 3, 1]
    assert ff.ghosts(0) == [5, 7, 8] and ff.ghosts(1) == [6] and ff.ghosts(2) == [9]
    assert all(ff.ghosts(i) == [] for i in range(3, 9))


# --------------------------------------------------------------------------------------
# test_collisions.jl:260-363  "Ghost Collisions" (bitwise parent/ghost equivalence)
# --------------------------------------------------------------------------------------
def splitdims(m):
    m = np.asarray(m, dtype=np.float64)
    return [[float(m[0, k]), float(m[1, k])] for k in range(m.shape[1])]


LSHAPE = [splitdims([[Lx / 2, Lx / 2, 3 * Lx / 4, 3 * Lx / 4, Lx + 10000, Lx + 10000],
                     [Ly / 2, Ly + 10000, Ly + 10000, 3 * Ly / 4, 3 * Ly / 4, Ly / 2]])]
_TH = np.arange(0, 2 * np.pi + 1e-12, np.pi / 50)
_R = Ly / 4 + 1000
OVAL = [[[float(_R * np.cos(t) + (Lx - 1)), float(_R * np.sin(t) + (Ly - 1))] for t in _TH]]
TALL_RECT = [splitdims([[5 * Lx / 8 + 1000, 5 * Lx / 8 + 1000, 3 * Lx / 4 + 1000, 3 * Lx / 4 + 1000],
                        [3 * Ly / 4, 5 * Ly / 4, 5 * Ly / 4, 3 * Ly / 4]])]
LONG_RECT = [splitdims([[-5 * Lx / 4, -5 * Lx / 4, -(3 * Lx / 4 - 1000), -(3 * Lx / 4 - 1000)],
                        [-7 * Lx / 8, -(3 * Lx / 4 - 1000), -(3 * Lx / 4 - 1000), -7 * Lx / 8]])]
SMALL_CORNER_RECT = [[[-1.1e5, -1.1e5], [-1.1e5, -9.5e4], [-9.5e4, -9.5e4], [-9.5e4, -1.1e5], [-1.1e5, -1.1e5]]]
LARGE_TRI = [[[-1e5, -1e5], [-1e5, 1e5], [1e5, -1e5], [-1e5, -1e5]]]
SOUTH_BOUND_RECT = [[[-9.8e4, -1.1e5], [-9.8e4, -9.5e4], [9.8e4, -9.5e4], [9.8e4, -1.1e5], [-9.8e4, -1.1e5]]]


def collide(ff, n_init, dom, lib, **kw):
    return host.timestep_collisions(ff, n_init, dom, Constants(), 10, host.CollisionSettings(), backend=lib, **kw)


def test_ghost_parent_parent(lib):  # :287-303
    dom = domains()["double"]
    ff = host.initialize_floe_field([LSHAPE, OVAL], hmean=0.5)
    collide(ff, 2, dom, lib)
    fx, fy = abs(ff.collision_force[0, 0]), abs(ff.collision_force[0, 1])
    t1, t2 = ff.collision_trq[0], ff.collision_trq[1]
    assert fx > 0 and fy > 0
    host.add_ghosts(ff, dom, backend=lib)
    assert ff.n == 8
    collide(ff, 2, dom, lib)
    assert fx == abs(ff.collision_force[0, 0]) == abs(ff.collision_force[1, 0])
    assert fy == abs(ff.collision_force[0, 1]) == abs(ff.collision_force[1, 1])
    assert t1 == ff.collision_trq[0] and t2 == ff.collision_trq[1]
    assert np.all(ff.collision_force[2:] == 0) and np.all(ff.collision_trq[2:] == 0)


def test_ghost_ghost(lib):  # :305-325
    dom = domains()["double"]
    ff = host.initialize_floe_field([TALL_RECT, LONG_RECT], hmean=0.5)
    tr = host.initialize_floe_field([translate(TALL_RECT, 0.0, -2 * Ly), translate(LONG_RECT, 2 * Lx, 0.0)], hmean=0.5)
    collide(tr, 2, dom, lib)
    fx, fy = abs(tr.collision_force[0, 0]), abs(tr.collision_force[0, 1])
    t1, t2 = tr.collision_trq[0], tr.collision_trq[1]
    assert fx > 0 or fy > 0
    host.add_ghosts(ff, dom, backend=lib)
    collide(ff, 2, dom, lib)
    assert [fx, fx] == list(np.abs(ff.collision_force[:2, 0]))
    assert [fy, fy] == list(np.abs(ff.collision_force[:2, 1]))
    assert t1 == ff.collision_trq[0] and t2 == ff.collision_trq[1]
    cols = [0, 1, 2, 3, 4, 6]
    assert np.array_equal(ff.interactions[0][:, cols], ff.interactions[3][:, cols])
    assert np.array_equal(ff.interactions[1][:, cols], ff.interactions[2][:, cols])


def test_ghost_parent(lib):  # :327-343
    dom = domains()["double"]
    up = translate(LONG_RECT, 0.0, 1.615 * Ly)
    ff = host.initialize_floe_field([TALL_RECT, up], hmean=0.5)
    tr = host.initialize_floe_field([translate(TALL_RECT, -2 * Lx, 0.0), up], hmean=0.5)
    collide(tr, 2, dom, lib)
    fx, fy = abs(tr.collision_force[0, 0]), abs(tr.collision_force[0, 1])
    t1, t2 = tr.collision_trq[0], tr.collision_trq[1]
    assert fx > 0 or fy > 0
    host.add_ghosts(ff, dom, backend=lib)
    collide(ff, 2, dom, lib)
    assert [fx, fx] == list(np.abs(ff.collision_force[:2, 0]))
    assert [fy, fy] == list(np.abs(ff.collision_force[:2, 1]))
    assert t1 == ff.collision_trq[0] and t2 == ff.collision_trq[1]
    cols = [0, 1, 2, 3, 4, 6]
    assert np.array_equal(ff.interactions[1][:, cols], ff.interactions[2][:, cols])
    assert len(ff.interactions[3]) == 0


def test_parent_and_ghosts_hit_same_floe(lib):  # :345-362
    dom = domains()["double"]
    ff = host.initialize_floe_field([SMALL_CORNER_RECT, LARGE_TRI], hmean=0.5)
    host.add_ghosts(ff, dom, backend=lib)
    assert ff.n == 5
    collide(ff, 2, dom, lib)
    a, b = ff.interactions[0], ff.interactions[1]
    assert len(a) == 3 and len(b) == 3
    assert a[0, XFORCE] != a[1, XFORCE] and a[0, XFORCE] != a[2, XFORCE]
    assert a[0, YFORCE] != a[1, YFORCE] and a[0, YFORCE] != a[2, YFORCE]
    ff = host.initialize_floe_field([SMALL_CORNER_RECT, SOUTH_BOUND_RECT], hmean=0.5)
    host.add_ghosts(ff, dom, backend=lib)
    assert ff.n == 6
    collide(ff, 2, dom, lib)
    a, b = ff.interactions[0], ff.interactions[1]
    assert len(a) == 2 and len(b) == 2
    assert a[0, XPOINT] != a[1, XPOINT]
    assert a[0, YPOINT] == a[1, YPOINT]


# --------------------------------------------------------------------------------------
# test_coupling.jl:464-640  "OA Forcings" (values originate from the MATLAB model)
# --------------------------------------------------------------------------------------
def oa_floe():
    f = Floe([[[-1.75e4, 5e4], [-1.75e4, 7e4], [-1.25e4, 7e4], [-1.25e4, 5e4], [-1.75e4, 5e4]]], 0.25)
    mc = json.load(open(os.path.join(GOLD, "test_mc_points.json")))
    f.x_subfloe_points, f.y_subfloe_points = np.array(mc["X"]), np.array(mc["Y"])
    return f


def nonuniform_fields():
    xl = np.arange(GRID.Nx + 1) * GRID.dx + GRID.x0
    yl = np.arange(GRID.Ny + 1) * GRID.dy + GRID.y0
    xgrid, ygrid = np.meshgrid(xl, yl)  # [row = y, col = x], output.jl:775-779
    psi = 0.5e4 * (np.sin(4 * (np.pi / 4e5) * xgrid) * np.sin(4 * (np.pi / 4e5) * ygrid))
    u = np.zeros_like(xgrid)
    u[1:, :] = -1e-4 * (psi[1:, :] - psi[:-1, :])
    v = np.zeros_like(ygrid)
    v[:, 1:] = 1e-4 * (psi[:, 1:] - psi[:, :-1])
    return u.T.copy(), v.T.copy()


OA_CASES = [
    # ocean (u,v), atmos (u,v), floe (u,v), dd, expected (fx/A, fy/A, trq/A), atol
    (("c", 1.0, 0.0), ("c", 0.0, 0.0), (0.0, 0.0), 2, (2.9760, 0.8296, -523.9212), (1e-3, 1e-3, 1e-3)),
    (("c", 0.0, 1.0), ("c", 0.0, 0.0), (0.0, 0.0), 2, (-0.8296, 2.9760, 239.3141), (1e-3, 1e-3, 1e-3)),
    (("c", 0.0, 0.0), ("c", 0.0, 0.0), (0.25, 0.1), 2, (-0.1756, -0.1419, 29.0465), (1e-3, 1e-3, 1e-1)),
    (("c", 0.0, 0.0), ("c", -1.0, -0.5), (0.0, 0.0), 2, (-0.0013, -6.7082e-4, 0.2276), (1e-3, 1e-3, 1e-3)),
    (("n",), ("c", 0.0, 0.0), (0.0, 0.0), 1, (-0.0182, 0.0392, 23.6399), (1e-3, 1e-3, 1e-3)),
    (("n",), ("n",), (0.5, -0.5), 1, (-1.6300, 1.1240, 523.2361), (1e-3, 1e-3, 2e-1)),
]


@pytest.mark.parametrize("case", range(len(OA_CASES)))
def test_oa_forcings(lib, case):
    ocn, atm, (fu, fv), dd, expect, atol = OA_CASES[case]
    nu, nv = nonuniform_fields()
    ocean = host.Ocean(GRID, *(ocn[1:] if ocn[0] == "c" else (nu, nv)))
    atmos = host.Atmos(GRID, *(atm[1:] if atm[0] == "c" else (nu, nv)))
    cb = lambda d: CollisionBoundary(d, GRID)
    dom = Domain(cb(North), cb(South), cb(East), cb(West))
    f = oa_floe()
    f.u, f.v = fu, fv
    area = f.area
    consts = Constants(E=1.5e3 * (np.sqrt(area) + np.sqrt(area)))
    model = host.Model(GRID, ocean, atmos, dom, host.FloeField([f]))
    host.timestep_coupling(model, 10, consts, host.CouplingSettings(dd=dd), host.FloeSettings(), backend=lib)
    fl = model.floes
    assert fl.fxOA[0] / area == pytest.approx(expect[0], abs=atol[0])
    assert fl.fyOA[0] / area == pytest.approx(expect[1], abs=atol[1])
    assert fl.trqOA[0] / area == pytest.approx(expect[2], abs=atol[2])


# --------------------------------------------------------------------------------------
# test_update_floe.jl:2-42  calc_stress! / calc_strain!
# --------------------------------------------------------------------------------------
def test_stress_strain(lib):
    d = json.load(open(os.path.join(GOLD, "stress_strain.json")))
    stress_hist = [[-4971.252, 17483.052, 17483.052, -57097.458], [4028.520, 9502.886, 9502.886, -205199.791]]
    strains = [[-0.0372, 0, 0, .9310], [7.419, 0, 0, -6.987]]
    for i in range(2):
        f = Floe(d["coords"][i], d["height"][i], 0.0)
        f.u, f.v, f.xi = d["u"][i], d["v"][i], d["ξ"][i]
        ff = host.FloeField([f])
        ff.interactions = [np.array(d["interactions"][i], dtype=np.float64)]
        ff.stress_instant[0] = np.array(d["last_stress"][i]).T.ravel()
        # Δt = 0 freezes the rigid move and the AB2 velocity update, so stress_instant and
        # strain are exactly what calc_stress!/calc_strain! give on the stored state.
        host.timestep_floe_properties(ff, 1, 0, host.FloeSettings(), backend=lib)
        assert ff.stress_instant[0] == pytest.approx(stress_hist[i], abs=1e-3)
        assert ff.strain[0] * 1e6 == pytest.approx(strains[i], abs=1e-3)
        assert np.array_equal(ff.coords(0), np.asarray(d["coords"][i][0]))


# --------------------------------------------------------------------------------------
# test_floe_utils.jl:65-71 (host helper)
# --------------------------------------------------------------------------------------
def test_moment_of_inertia_helper():
    tri = np.array([[0, 1], [0, 0], [1, 0], [0, 1]], dtype=np.float64) * 6.67
    assert host.calc_moment_inertia(tri, host.ring_centroid(tri), 0.5) == pytest.approx(50581.145, abs=1e-3)


# ---- which_vertices_match_points (floe_utils.jl:331-352), test_floe_utils.jl:76-137 ------------------------------
MATCH_CASES = [
    # (points = ring of polygon 1, ring of polygon 2, expected 1-based vertex indices of polygon 2)
    ([[0.0, 0.0], [0.0, 20.0], [20.0, 20.0], [20.0, 0.0], [0.0, 0.0]],
     [[20.0, 0.0], [20.0, 20.0], [40.0, 20.0], [40.0, 0.0], [20.0, 0.0]], [1, 2]),                               # :76-83
    ([[0.0, 0.0], [0.0, 20.0], [20.0, 20.0], [20.0, 10.0], [20.0, 0.0], [0.0, 0.0]],
     [[40.0, 20.0], [40.0, 0.0], [20.0, 0.0], [20.0, 10.0], [20.0, 20.0], [40.0, 20.0]], [3, 4, 5]),             # :90-97
    ([[0.0, 0.0], [0.0, 20.0], [20.0, 20.0], [20.0, 18.0], [20.0, 15.0], [20.0, 0.0], [0.0, 0.0]],
     [[20.0, 18.0], [20.0, 20.0], [40.0, 20.0], [40.0, 0.0], [20.0, 0.0], [20.0, 15.0], [20.0, 18.0]], [1, 2, 5, 6]),  # :110-117
    ([[0.0, 0.0], [0.0, 20.0], [20.0, 20.0], [5.0, 5.0], [0.0, 0.0]],
     [[0.0, 0.0], [5.0, 5.0], [20.0, 20.0], [20.0, 0.0], [0.0, 0.0]], [1, 2, 3]),                                # :129-136
]


@pytest.mark.parametrize("case", MATCH_CASES, ids=["two_shared", "three_shared", "four_shared", "triangle_shared"])
def test_which_vertices_match_points(oracle_lib, case):
    """Pins the oracle's restatement (the CUDA kernels are held to the oracle bit for bit on the rows these indices feed)."""
    import ctypes as C
    pts, ring, want = case
    fn = oracle_lib.dll.szo_test_match_vertices
    fn.restype = C.c_int32
    fn.argtypes = [C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_double), C.c_int32, C.POINTER(C.c_int32)]
    p = np.ascontiguousarray(pts, dtype=np.float64)
    r = np.ascontiguousarray(ring, dtype=np.float64)
    out = np.zeros(len(p), dtype=np.int32)
    m = fn(p.ctypes.data_as(C.POINTER(C.c_double)), len(p), r.ctypes.data_as(C.POINTER(C.c_double)), len(r),
           out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert out[:m].tolist() == want


# ---- coupling helper truth tables, test_coupling.jl:165-195 (grid x in [-10, 10] dx 2, y in [-8, 8] dy 4) ------------
def _helper_handle(oracle_lib):
    h = capi.Handle(oracle_lib)
    h.set_grid(10, 4, -10.0, 10.0, -8.0, 8.0)
    return h


def test_find_center_cell_index_table(oracle_lib):  # :167-178
    import ctypes as C
    h = _helper_handle(oracle_lib)
    fn = oracle_lib.dll.szo_test_find_center_cell_index
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_int32)]
    xpoints = [-10.5, -10, -10, -6.5, -6, -4, 10, 10.5, 12]
    ypoints = [0.0, 6.0, -8.0, 4.5, 0.0, 5.0, -8.0, 0.0, 0.0]
    xidx = [1, 1, 1, 3, 3, 4, 11, 11, 12]
    yidx = [3, 5, 1, 4, 3, 4, 1, 3, 3]
    out = (C.c_int32 * 2)()
    for x, y, ix, iy in zip(xpoints, ypoints, xidx, yidx):
        assert fn(h.h, x, y, out) == 0
        assert (out[0], out[1]) == (ix, iy), (x, y)


def test_in_bounds_truth_tables(oracle_lib):  # :180-195; in_bounds(x, y, grid, north/south kind, east/west kind)
    import ctypes as C
    h = _helper_handle(oracle_lib)
    fn = oracle_lib.dll.szo_test_in_bounds
    fn.restype = C.c_int32
    fn.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32, C.c_int32]
    x = [-12, -10, -8, -6, 0, 4, 4, 10, 12, 12]
    y = [5, -6, 4, 10, -10, 8, -8, -6, 4, 10]
    open_open = [False, True, True, False, False, True, True, True, False, False]
    periodic_open = [False, True, True, True, True, True, True, True, False, False]   # N-S periodic, E-W open
    open_periodic = [True, True, True, False, False, True, True, True, True, False]   # N-S open, E-W periodic
    for k in range(len(x)):
        assert bool(fn(h.h, x[k], y[k], 0, 0)) == open_open[k]
        assert bool(fn(h.h, x[k], y[k], 1, 0)) == open_periodic[k]   # per_x = E-W periodic
        assert bool(fn(h.h, x[k], y[k], 0, 1)) == periodic_open[k]   # per_y = N-S periodic
        assert bool(fn(h.h, x[k], y[k], 1, 1))


# ---- hand-derived known answers for what NO reference test reaches (round-1 verdict: "pin what is unpinned") -----------
# The reference cannot run here (no Julia on the build image or on the GPU box: profiles/r2/r2a_julia_probe.txt), so these
# cases are worked out by hand from the reference's formulas (collisions.jl:30-119,149-188,243-283) plus ONE stated
# assumption each about GeometryOps 0.1.x, the un-vendored clipper.  They pin the oracle and the CUDA kernels on
# configurations where a shared misreading would otherwise be invisible; the assumption is what remains unverified.
COMB = [[[0.0, 2.7e4], [0.0, 3.5e4], [5e4, 3.5e4], [5e4, 2.7e4], [4e4, 2.7e4], [4e4, 3e4], [3e4, 3e4], [3e4, 2.7e4],
         [2e4, 2.7e4], [2e4, 3e4], [1e4, 3e4], [1e4, 2.7e4], [0.0, 2.7e4]]]
COMB_RECT = [[[-1e4, 2.5e4], [-1e4, 2.9e4], [6e4, 2.9e4], [6e4, 2.5e4], [-1e4, 2.5e4]]]


def test_three_regions_are_ordered_by_first_crossing_along_polygon_one(lib):
    """A comb with three teeth dipping into a long rectangle: three overlap regions of 1e4 x 0.2e4 m each.
    ASSUMPTION (GeometryOps): output rings are started at the first not-yet-visited crossing met walking along polygon 1
    from its first vertex — the rule that the reference's own two-region answer pins (test_collisions.jl:64-81).  Walking
    the comb from (0, 2.7e4): exit at x = 0 (tooth 1), entry / exit at x = 5e4 / 4e4 (tooth 3), entry / exit at
    x = 3e4 / 2e4 (tooth 2), entry at x = 1e4 (tooth 1): regions in the order tooth 1, tooth 3, tooth 2."""
    c, rect = Floe(COMB, 0.25), Floe(COMB_RECT, 0.25)
    rect.v = -0.1
    ff = host.floe_floe_interaction(c, 1, rect, 2, Constants(), 10, 0.55, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 3
    assert np.allclose(r[:, XPOINT], [0.5e4, 4.5e4, 2.5e4], rtol=0, atol=1e-6)   # contact point = region centroid (:178)
    assert np.allclose(r[:, YPOINT], [2.8e4] * 3, rtol=0, atol=1e-6)
    assert np.allclose(r[:, OVERLAP], [2e7] * 3, rtol=0, atol=1e-3)
    # each tooth has two crossing points on y = 2.9e4, 1e4 apart: normal force along -y ... flipped to +y because moving the
    # comb DOWN would deepen the overlap (:59-67); magnitude = area * force_factor (:69), force_factor of :375-379
    ai, aj = c.area, rect.area
    ffac = 6e6 * (0.25 * 0.25) / (0.25 * np.sqrt(aj) + 0.25 * np.sqrt(ai))
    assert np.allclose(np.abs(r[:, XFORCE]), 0.0, atol=1e-6 * 2e7 * ffac)  # friction (the rectangle moves in y only) has no x part
    # normal part +y; friction: relative velocity (0, -0.1) - 0 at the contact => along -y on the comb, capped by mu |N|
    normal = 2e7 * ffac
    G = 6e6 / (2 * (1 + 0.3))
    fric = min(G * 1e4 * 10 * normal * 0.1, 0.2 * normal)   # :253-278 with dl = 1e4, dt = 10, |dv| = 0.1
    assert np.allclose(r[:, YFORCE], [normal - fric] * 3, rtol=1e-12)


def test_exactly_shared_edge_is_no_overlap(lib):
    """Two rectangles that share one whole edge, and two that touch in one corner.
    ASSUMPTION (GeometryOps): `intersection(...; target = PolygonTrait)` returns no polygon when the interiors do not
    overlap, so no row, no fuse tag, overarea unchanged (collisions.jl:356-364: total area 0)."""
    a = Floe(CORNER_RECT, 0.25)
    for dx, dy in ((2e4, 0.0), (2e4, 0.4e4)):
        b = Floe(translate(CORNER_RECT, dx, dy), 0.25)
        a.v, b.v = -0.1, 0.2
        ff = host.floe_floe_interaction(a, 1, b, 2, Constants(), 10, 0.55, backend=lib)
        assert len(ff.interactions[0]) == 0 and len(ff.interactions[1]) == 0
        assert ff.status_tag[0] == capi.STATUS_ACTIVE and ff.status_tag[1] == capi.STATUS_ACTIVE
        assert ff.overarea[0] == 0.0 and ff.overarea[1] == 0.0


def test_collinear_overlapping_edges_take_the_many_intersection_path(lib):
    """Two equal rectangles shifted by half their length along x: the top and bottom edges overlap COLLINEARLY, the
    overlap region is [1e4, 2e4] x [2.5e4, 2.9e4] (50 % of either area: no fusing).
    ASSUMPTION (GeometryOps): `intersection_points` returns the end points of the two collinear overlaps, de-duplicated:
    the four corners of the region, so m = 4 and `_many_intersect_normal_force!` runs (collisions.jl:50-52,78-119).
    By hand: three of the region's four edges lie on rectangle 1 (top 1e4, right 0.4e4, bottom 1e4; the left edge is
    interior to it), each contributes its inward normal times its length: (0, -1e4) + (-0.4e4, 0) + (0, 1e4) =
    (-0.4e4, 0) => direction (-1, 0), dl = (1e4 + 0.4e4 + 1e4) / 3 = 8000; moving rectangle 1 by (-1, 0) shrinks the
    overlap, so no flip (:59-67).  Equal velocities: no friction."""
    a, b = Floe(CORNER_RECT, 0.25), Floe(translate(CORNER_RECT, 1e4, 0.0), 0.25)
    a.v = b.v = -0.1
    ff = host.floe_floe_interaction(a, 1, b, 2, Constants(), 10, 0.55, backend=lib)
    r = ff.interactions[0]
    assert len(r) == 1 and ff.status_tag[0] == capi.STATUS_ACTIVE
    ffac = 6e6 * (0.25 * 0.25) / (2 * 0.25 * np.sqrt(8e7))
    assert r[0, OVERLAP] == pytest.approx(4e7, abs=1e-3)
    assert r[0, XPOINT] == pytest.approx(1.5e4, abs=1e-6) and r[0, YPOINT] == pytest.approx(2.7e4, abs=1e-6)
    assert r[0, XFORCE] == pytest.approx(-4e7 * ffac, rel=1e-12) and abs(r[0, YFORCE]) <= 1e-6 * 4e7 * ffac
    m = ff.interactions[1]
    assert len(m) == 1 and m[0, XFORCE] == -r[0, XFORCE]


# A vertex of one ring lying EXACTLY on an edge of the other (floes cut by the domain wall, writer cells that end on the
# wall): the two ring edges that meet in the vertex both cross the other ring's edge at that point, and the order of the
# two coincident crossings along that edge must come from the perturbation, not from rounding — otherwise the trace walks
# all the way round the other polygon and returns "cell + floe" (found in round 2 through the gridded output: cell areas
# of 2.0; in the collision step the same mis-trace against a wall rectangle tagged floes of an exactly packed field for
# removal).  Rings taken from synth.make_field(200 / 2000, scale = 1.0, collision walls).
TOUCH_CASES = [
    # floe ring, cell box (xmin, xmax, ymin, ymax)
    ([[0.0, 1.1378211036468638e+04], [9.3351619109063267e+02, 1.0247514716945245e+04], [2.2737367544323206e-13, 9.5258943684869191e+03],
      [0.0, 1.1378211036468638e+04]], (0.0, 7500.0, 7500.0, 15000.0)),
    ([[3.6053980784869505e+04, 2.6135673163098036e+03], [3.7328998928432207e+04, 2.7849060063884131e+03],
      [3.7461388812657264e+04, 2.7112724623584427e+03], [3.7564493061034278e+04, 2.5345229846931297e+03],
      [3.8203156761207458e+04, 6.8212102632969618e-13], [3.4632370257829061e+04, 0.0], [3.6053980784869505e+04, 2.6135673163098036e+03]],
     (22500.0, 45000.0, 0.0, 22500.0)),
]


@pytest.mark.parametrize("case", range(len(TOUCH_CASES)))
def test_vertex_exactly_on_the_other_rings_edge(case, lib):
    ring, (x0, x1, y0, y1) = TOUCH_CASES[case]
    ring = np.array(ring)
    box = np.array([[x0, y0], [x0, y1], [x1, y1], [x1, y0], [x0, y0]])
    a2 = np.sum(ring[:-1, 0] * ring[1:, 1] - ring[1:, 0] * ring[:-1, 1])
    h = capi.Handle(lib)
    for p, q in ((ring, box), (box, ring)):
        regs, areas = h.clip_polygons(p, q)
        assert len(regs) == 1 and areas[0] == pytest.approx(abs(a2) / 2, rel=1e-12)   # the floe lies inside the box
    h.close()


def test_exactly_packed_field_touches_the_walls_without_overlap(lib):
    """Voronoi cells cut by the collision walls (scale 1.0: the floes tile the domain exactly, vertices ON the walls): no
    wall contact, nobody removed or fused, and the gridded area of a writer grid that ends on the walls is the domain's."""
    from subzero_jl_b200 import synth
    for n in (200, 2000):
        f = synth.make_field(n, scale=1.0, walls="collision", npoints=10, cache=False)
        h = synth.setup_handle(f, lib)
        st0 = f.floes.status_tag.copy()
        h.step_collisions()
        fa = h.download_floes(mc=False)
        offs, rows = h.interactions()
        assert np.array_equal(fa.status_tag, st0) and not (rows[:, 0] < 0).any()
        xg = np.linspace(0.0, f.L, 5)
        d = h.eulerian_data(xg, xg, [capi.GRID_OUTPUTS.index("area_grid"), capi.GRID_OUTPUTS.index("si_frac_grid")])
        assert d[..., 0].sum() == pytest.approx(f.L ** 2, rel=1e-12) and d[..., 1].max() < 1.0 + 1e-9
        h.close()
