"""Services for the host-side processes: batched pair overlap areas and the Eulerian
gridded output.  CPU tests pin the oracle on hand-computed cases (the reference's own tests hold no numeric golden for
either: "parity unpinned" beyond the clip/area goldens of test_reference_golden.py); GPU tests compare the CUDA
product with the oracle — areas bit-exact, gridded averages within 1e-9."""
import numpy as np
import pytest

import fields
from parity_util import rel_err
from subzero_jl_b200 import capi, host, synth

KINDS = list(range(len(capi.GRID_OUTPUTS)))
TOL = 1e-9  # gridded averages are floating-point reductions: the north_star tolerance for per-floe state


def grid_err(dg, do, k):
    """Relative error of output k, judged against the magnitude of its tensor: the off-diagonal strain of a rigid
    motion is pure rounding noise (1e-21 next to 1e-5 diagonals) in the product and the oracle alike."""
    group = range(9, 14) if 9 <= k <= 13 else (range(14, 18) if k >= 14 else (k,))
    scale = max(float(np.sqrt(np.mean(do[..., g] ** 2))) for g in group)
    return float(np.max(np.abs(dg[..., k] - do[..., k]) / np.maximum(np.abs(do[..., k]), max(scale, 1e-300))))


def square(x0, y0, s):
    return [[x0, y0], [x0, y0 + s], [x0 + s, y0 + s], [x0 + s, y0], [x0, y0]]


def two_square_field():
    grid = host.RegRectilinearGrid(0.0, 4e4, 0.0, 4e4, dx=1e4, dy=1e4)
    dom = host.Domain(*[host.OpenBoundary(d, grid) for d in (host.North, host.South, host.East, host.West)])
    fl = host.initialize_floe_field([square(5e3, 5e3, 1e4), square(1.2e4, 7e3, 1e4), square(3.1e4, 3.1e4, 5e3)], dom, hmean=0.5,
                                    rng=np.random.default_rng(1))
    return grid, dom, fl


def oracle_handle(oracle_lib, grid, dom, fl):
    h = capi.Handle(oracle_lib)
    h.set_grid(grid.Nx, grid.Ny, grid.x0, grid.xf, grid.y0, grid.yf)
    dom.push(h)
    h.upload_floes(fl)
    return h


def test_oracle_pair_overlap_known_answers(oracle_lib):
    grid, dom, fl = two_square_field()
    h = oracle_handle(oracle_lib, grid, dom, fl)
    areas, inter = h.pair_overlap_areas([[1, 2], [2, 1], [1, 3], [3, 3]])
    # squares [5,15]x[5,15] km and [12,22]x[7,17] km overlap in [12,15]x[7,15] km = 3 km x 8 km
    assert areas[0] == pytest.approx(3e3 * 8e3, rel=1e-12) and areas[1] == pytest.approx(3e3 * 8e3, rel=1e-12)
    assert inter[0] and inter[1]
    assert areas[2] == 0.0 and not inter[2]
    assert areas[3] == pytest.approx(25e6, rel=1e-12) and inter[3]   # a floe with itself: its own area
    with pytest.raises(capi.SubzeroError):
        h.pair_overlap_areas([[1, 4]])


def test_oracle_eulerian_known_answers(oracle_lib):
    grid, dom, fl = two_square_field()
    fl.u[:] = [1.0, 3.0, -2.0]
    fl.v[:] = [0.5, 0.5, 4.0]
    fl.overarea[:] = [2.0, 4.0, 6.0]
    fl.stress_accum[:] = np.array([[1.0, 0.5, 0.5, 3.0], [1.0, 0.5, 0.5, 3.0], [2.0, 0.0, 0.0, -1.0]])
    h = oracle_handle(oracle_lib, grid, dom, fl)
    xg = np.linspace(0, 4e4, 3)   # 2 x 2 cells of 20 km
    data = h.eulerian_data(xg, xg, KINDS)
    k = {n: i for i, n in enumerate(capi.GRID_OUTPUTS)}
    # cell (0,0) = [0,20]^2 km: floe 1 entirely (100 km^2), floe 2 over [12,20]x[7,17] = 80 km^2
    a1, a2 = 100e6, 80e6
    m1, m2 = fl.mass[0], fl.mass[1]
    mt = m1 * 1.0 + m2 * 0.8
    w1, w2 = 1.0 * m1 / mt, 0.8 * m2 / mt
    c = data[0, 0]
    assert c[k["area_grid"]] == pytest.approx(a1 + a2, rel=1e-12)
    assert c[k["mass_grid"]] == pytest.approx(mt, rel=1e-12)
    assert c[k["si_frac_grid"]] == pytest.approx((a1 + a2) / 400e6, rel=1e-12)
    assert c[k["u_grid"]] == pytest.approx(1.0 * w1 + 3.0 * w2, rel=1e-12)
    assert c[k["v_grid"]] == pytest.approx(0.5 * (w1 + w2), rel=1e-12)
    assert c[k["overarea_grid"]] == pytest.approx(3.0)
    sxx, sxy, syy = 1.0 * (w1 + w2), 0.5 * (w1 + w2), 3.0 * (w1 + w2)
    assert c[k["stress_xx_grid"]] == pytest.approx(sxx) and c[k["stress_yy_grid"]] == pytest.approx(syy)
    assert c[k["stress_eig_grid"]] == pytest.approx(max(np.linalg.eigvals([[sxx, sxy], [sxy, syy]]).real), rel=1e-12)
    # cell (1,0) (x in [20,40], y in [0,20]): floe 2 over [20,22]x[7,17] = 20 km^2 only -> its own values
    c = data[1, 0]
    assert c[k["area_grid"]] == pytest.approx(20e6, rel=1e-12) and c[k["u_grid"]] == pytest.approx(3.0)
    assert c[k["mass_grid"]] == pytest.approx(m2 * 0.2, rel=1e-12)
    # cell (1,1): floe 3 entirely; cell (0,1): empty -> zeros (output.jl:912-916)
    assert data[1, 1, k["u_grid"]] == pytest.approx(-2.0) and data[1, 1, k["area_grid"]] == pytest.approx(25e6)
    assert data[1, 1, k["stress_eig_grid"]] == pytest.approx(2.0)
    assert np.all(data[0, 1] == 0.0)
    # the host mirror of the reference call
    w = host.GridOutputWriter(100, grid, (2, 2), outputs=["u_grid", "area_grid"])
    host.calc_eulerian_data(fl, None, w, backend=oracle_lib)
    assert np.allclose(w.data[..., 0], data[..., k["u_grid"]]) and np.allclose(w.data[..., 1], data[..., k["area_grid"]])


def test_oracle_eulerian_conserves_area_and_mass(oracle_lib):
    f = synth.make_field(500, scale=0.99, walls="collision", npoints=20, cache=False)
    fields.perturb_state(f.floes)
    h = synth.setup_handle(f, oracle_lib)
    xg = np.linspace(0.0, f.L, 8)
    yg = np.linspace(0.0, f.L, 5)
    data = h.eulerian_data(xg, yg, [capi.GRID_OUTPUTS.index("area_grid"), capi.GRID_OUTPUTS.index("mass_grid")])
    fa = h.download_floes(mc=False)
    assert data[..., 0].sum() == pytest.approx(fa.area.sum(), rel=1e-9)   # the cells tile the domain
    assert data[..., 1].sum() == pytest.approx(fa.mass.sum(), rel=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(400, "collision", 1.01), (3000, "periodic", 1.02), (3000, "shear", 0.99)],
                         ids=lambda c: "n%d_%s_%g" % c)
def test_pair_overlap_areas_bit_exact(cfg, product_lib, oracle_lib):
    n, walls, scale = cfg
    f = synth.make_field(n, scale=scale, walls=walls, npoints=20, cache=False)
    hg, ho = synth.setup_handle(f, product_lib), synth.setup_handle(f, oracle_lib)
    for h in (hg, ho):
        h.add_ghosts()
        h.step_collisions()
    cand = hg.pairs(0)
    assert np.array_equal(cand, ho.pairs(0)) and len(cand) > n
    rng = np.random.default_rng(n)
    nt = hg.counts()["n_total"]
    extra = rng.integers(1, nt + 1, size=(2000, 2))          # mostly non-interacting pairs, some i == j
    pairs = np.concatenate([cand, cand[:, ::-1], extra])
    ag, ig = hg.pair_overlap_areas(pairs)
    ao, io = ho.pair_overlap_areas(pairs)
    assert np.array_equal(ig, io)
    assert np.array_equal(ag, ao), "areas differ: rel %.3e" % rel_err(ag, ao)
    assert ig[:len(cand)].all() and (ag[:len(cand)] > 0).sum() == hg.counts()["n_overlap"] or walls != "collision"
    assert np.all(ag[~ig] == 0.0)


@pytest.mark.gpu
def test_services_on_nonconvex_fixture_shapes(product_lib, oracle_lib):
    """The reference's own floe shapes (non-convex, up to 591 vertices): multi-region clips and the warp kernels."""
    f = fields.fixture_shape_field(scale=1.04, walls="collision")
    hg, ho = synth.setup_handle(f, product_lib), synth.setup_handle(f, oracle_lib)
    for h in (hg, ho):
        h.step_collisions()
    cand = hg.pairs(0)
    ag, ig = hg.pair_overlap_areas(cand)
    ao, io = ho.pair_overlap_areas(cand)
    assert np.array_equal(ig, io) and np.array_equal(ag, ao) and (ag > 0).sum() > 50
    g = f.grid
    xg, yg = np.linspace(g.x0, g.xf, 8), np.linspace(g.y0, g.yf, 6)
    dg, do = hg.eulerian_data(xg, yg, KINDS), ho.eulerian_data(xg, yg, KINDS)
    for k, name in enumerate(capi.GRID_OUTPUTS):
        assert grid_err(dg, do, k) < TOL, name


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [(300, "collision", (2, 2)), (3000, "periodic", (10, 5)), (3000, "shear", (40, 40)),
                                 (20000, "collision", (64, 48))], ids=lambda c: "n%d_%s_%dx%d" % (c[0], c[1], c[2][0], c[2][1]))
def test_eulerian_data_matches_oracle(cfg, product_lib, oracle_lib):
    n, walls, dims = cfg
    f = synth.make_field(n, scale=1.01, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    hg, ho = synth.setup_handle(f, product_lib), synth.setup_handle(f, oracle_lib)
    for h in (hg, ho):
        h.step(0, True)          # non-trivial stress, strain, overarea, AB2 history
        h.add_ghosts()           # the reference writes between add_ghosts! and timestep_collisions!
    xg = np.linspace(0.0, f.L, dims[0] + 1)
    yg = np.linspace(0.0, f.L, dims[1] + 1)
    dg = hg.eulerian_data(xg, yg, KINDS)
    do = ho.eulerian_data(xg, yg, KINDS)
    assert dg.shape == do.shape == (dims[0], dims[1], len(KINDS))
    for k, name in enumerate(capi.GRID_OUTPUTS):
        e = grid_err(dg, do, k)
        assert e < TOL, (name, e)
    assert np.array_equal(dg[..., 7] > 0, do[..., 7] > 0)
    # a subset of outputs in another order
    sub = [7, 0, 13]
    assert np.array_equal(hg.eulerian_data(xg, yg, sub), dg[..., sub])
    # deterministic
    assert np.array_equal(hg.eulerian_data(xg, yg, KINDS), dg)


def topo_field():
    f = synth.make_field(200, scale=1.0, walls="collision", npoints=10, cache=False)
    topo = host.initialize_topography_field([[square(0.45 * f.L, 0.45 * f.L, 2e3)]])
    g = f.grid
    f.domain = host.Domain(*[host.CollisionBoundary(d, g) for d in (host.North, host.South, host.East, host.West)], topography=topo)
    return f


def hand_topo_case(lib):
    """Hand-computed: grid of 2 x 1 cells of 10 km; topography = the right half of cell A, [5, 10] x [0, 10] km; one floe
    [2, 12] x [4, 6] km (20 km^2).  Cell A: the floe's piece [2, 10] x [4, 6] = 16 km^2, of which [5, 10] x [4, 6] = 10 km^2 lie
    under the topography -> 6 km^2 on a free cell area of 50 km^2 (output.jl:826-846: cell_poly_list = cell minus
    topography); cell B: 4 km^2 on 100 km^2.  A second set-up covers cell A completely: every output of A is 0 (:831-834)."""
    km = 1e3
    grid = host.RegRectilinearGrid(0.0, 20 * km, 0.0, 10 * km, dx=10 * km, dy=10 * km)
    floe = host.Floe([[[2 * km, 4 * km], [2 * km, 6 * km], [12 * km, 6 * km], [12 * km, 4 * km], [2 * km, 4 * km]]], 0.5, 0.0,
                     rng=np.random.default_rng(1))
    floe.u, floe.v = 0.3, -0.1
    out = {}
    for name, topo_ring in (("half", [[5 * km, 0.0], [5 * km, 10 * km], [10 * km, 10 * km], [10 * km, 0.0], [5 * km, 0.0]]),
                            ("all", [[-1 * km, -1 * km], [-1 * km, 11 * km], [10 * km, 11 * km], [10 * km, -1 * km], [-1 * km, -1 * km]])):
        topo = host.initialize_topography_field([[topo_ring]])
        dom = host.Domain(*[host.OpenBoundary(d, grid) for d in (host.North, host.South, host.East, host.West)], topography=topo)
        h = capi.Handle(lib)
        h.set_grid(grid.Nx, grid.Ny, grid.x0, grid.xf, grid.y0, grid.yf)
        dom.push(h)
        ff = host.FloeField([floe])
        h.upload_floes(ff)
        kinds = [capi.GRID_OUTPUTS.index(k) for k in ("area_grid", "si_frac_grid", "mass_grid", "u_grid", "v_grid")]
        out[name] = (h.eulerian_data(np.array([0.0, 10 * km, 20 * km]), np.array([0.0, 10 * km]), kinds), float(ff.mass[0]))
        h.close()
    d, mass = out["half"]
    assert d[0, 0, 0] == pytest.approx(6e6, rel=1e-12) and d[1, 0, 0] == pytest.approx(4e6, rel=1e-12)
    assert d[0, 0, 1] == pytest.approx(6e6 / 50e6, rel=1e-12) and d[1, 0, 1] == pytest.approx(0.04, rel=1e-12)
    assert d[0, 0, 2] == pytest.approx(mass * 6 / 20, rel=1e-12) and d[1, 0, 2] == pytest.approx(mass * 4 / 20, rel=1e-12)
    assert d[0, 0, 3] == pytest.approx(0.3, rel=1e-12) and d[1, 0, 4] == pytest.approx(-0.1, rel=1e-12)  # one floe: its own velocity
    d, mass = out["all"]
    assert np.all(d[0, 0, :] == 0.0)
    assert d[1, 0, 0] == pytest.approx(4e6, rel=1e-12) and d[1, 0, 1] == pytest.approx(0.04, rel=1e-12)


def test_oracle_eulerian_topography_hand_case(oracle_lib):
    hand_topo_case(oracle_lib)


def test_oracle_eulerian_topography_conserves_the_ice_outside_the_land(oracle_lib):
    """Voronoi field with two floes' rings turned into topography: the gridded area is the floe area minus what lies under
    the land, and never exceeds the free cell area."""
    f = topo_field()
    h = synth.setup_handle(f, oracle_lib)
    xg = yg = np.linspace(0.0, f.L, 5)
    d = h.eulerian_data(xg, yg, [capi.GRID_OUTPUTS.index("area_grid"), capi.GRID_OUTPUTS.index("si_frac_grid")])
    h0 = synth.setup_handle(synth.make_field(200, scale=1.0, walls="collision", npoints=10, cache=False), oracle_lib)
    d0 = h0.eulerian_data(xg, yg, [capi.GRID_OUTPUTS.index("area_grid"), capi.GRID_OUTPUTS.index("si_frac_grid")])
    assert d[..., 0].sum() < d0[..., 0].sum() and d[..., 0].sum() > 0.95 * d0[..., 0].sum()   # a 4 km x 4 km island in 28 km x 28 km
    assert np.all(d[..., 1] <= 1.0 + 1e-9) and np.all(d[..., 0] <= d0[..., 0] + 1e-6)


@pytest.mark.gpu
def test_eulerian_data_topography_hand_case_cuda(product_lib):
    hand_topo_case(product_lib)


@pytest.mark.gpu
def test_eulerian_data_with_topography_matches_oracle(product_lib, oracle_lib):
    f = topo_field()
    fields.perturb_state(f.floes)
    hg, ho = synth.setup_handle(f, product_lib), synth.setup_handle(f, oracle_lib)
    for h in (hg, ho):
        h.step(0, True)
    for dims in ((4, 4), (9, 7), (28, 28)):
        xg, yg = np.linspace(0.0, f.L, dims[0] + 1), np.linspace(0.0, f.L, dims[1] + 1)
        dg, do = hg.eulerian_data(xg, yg, KINDS), ho.eulerian_data(xg, yg, KINDS)
        for k, name in enumerate(capi.GRID_OUTPUTS):
            assert grid_err(dg, do, k) < TOL, (dims, name)
        assert np.array_equal(dg[..., 7] > 0, do[..., 7] > 0)
