"""CPU-side checks (`-m "not gpu"`): the oracle against itself (grid vs literal O(N^2) loop,
thread-count independence, conservation properties), the host mirror of the reference API, and
that the product C-ABI library loads and exports every symbol include/subzero_b200.h declares
(no compute calls: there is no GPU here)."""
import os
import re
import warnings

import numpy as np
import pytest

import fields
from parity_util import STATE_FIELDS, compare_collision_outputs, compare_state, rel_err
from subzero_jl_b200 import capi, host, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_bound_and_exported():
    hdr = open(os.path.join(ROOT, "include", "subzero_b200.h")).read()
    names = set(re.findall(r"SZ_FN\((\w+)\)\(", hdr))
    assert names == set(capi.Library.SIGS), names ^ set(capi.Library.SIGS)
    lib = capi.product()  # dlopen + getattr of every symbol; raises if one is missing
    assert sorted(lib.exported()) == sorted("sz_" + n for n in names)
    assert lib.version().decode().startswith("subzero-b200")
    cfg = lib.default_config_struct()
    assert cfg.E == 6e6 and cfg.dt == 10 and cfg.floe_floe_max_overlap == 0.55 and cfg.stress_lambda == 0.2


def test_product_fails_loudly_without_gpu():
    import ctypes
    lib = capi.product()
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.SubzeroError):
        capi.Handle(lib)  # sz_create -> SZ_ERR_CUDA: no CPU fallback exists


def test_oracle_default_config_matches_product(oracle_lib):
    a, b = oracle_lib.default_config_struct(), capi.product().default_config_struct()
    for name, _ in capi.Config._fields_:
        assert getattr(a, name) == getattr(b, name), name


@pytest.mark.parametrize("walls", ["collision", "periodic", "shear"])
def test_grid_broad_phase_equals_all_pairs_loop(walls, oracle_lib):
    """collisions.jl:745-763 is a literal double loop; the oracle's uniform grid must emit the
    same sorted candidate list."""
    f = synth.make_field(700, scale=1.02, walls=walls, npoints=20, cache=False)
    a, b = synth.setup_handle(f, oracle_lib), synth.setup_handle(f, oracle_lib)
    a.add_ghosts()
    b.add_ghosts()
    a.step_collisions()
    os.environ["SZO_BRUTE_FORCE"] = "1"
    try:
        b.step_collisions()
    finally:
        del os.environ["SZO_BRUTE_FORCE"]
    assert not compare_collision_outputs(a, b)
    p = a.pairs(0)
    assert len(p) > 700 and np.all(p[:, 0] < p[:, 1])


def test_oracle_thread_count_independent(oracle_lib):
    f = synth.make_field(1500, scale=1.01, walls="periodic", npoints=40, cache=False)
    fields.perturb_state(f.floes)
    a, b = synth.setup_handle(f, oracle_lib, threads=1), synth.setup_handle(f, oracle_lib, threads=4)
    a.step(0, True)
    b.step(0, True)
    assert not compare_state(a.download_floes(), b.download_floes(), exact=STATE_FIELDS)


def test_newtons_third_law_and_row_bookkeeping(oracle_lib):
    f = synth.make_field(1200, scale=1.02, walls="collision", npoints=20, cache=False)
    h = synth.setup_handle(f, oracle_lib)
    h.step_collisions()
    offs, rows = h.interactions()
    c = h.counts()
    ff = rows[:, 0] > 0
    tot = np.abs(rows[ff, 1]).sum()
    assert abs(rows[ff, 1].sum()) <= 1e-9 * tot and abs(rows[ff, 2].sum()) <= 1e-9 * tot
    fa = h.download_floes()
    owner = np.repeat(np.arange(c["n_total"]), np.diff(offs))
    fx = np.bincount(owner, weights=rows[:, 1], minlength=c["n_total"])
    trq = np.bincount(owner, weights=rows[:, 5], minlength=c["n_total"])
    assert rel_err(fa.collision_force[:, 0], fx) < 1e-12 and rel_err(fa.collision_trq, trq) < 1e-12
    # wall rows exist and carry the reference's element ids (collisions.jl:612-642)
    ids = set(rows[~ff, 0].astype(int).tolist())
    assert ids and ids <= {-1, -2, -3, -4}


def test_fixture_shape_field_runs_clean(oracle_lib):
    f = fields.fixture_shape_field(scale=1.04)
    h = synth.setup_handle(f, oracle_lib)
    h.step(0, True)
    c = h.counts()
    assert c["n_clip_fail"] == 0 and c["n_overlap"] > 100
    fa = h.download_floes()
    assert np.all(np.isfinite(fa.vert_xy)) and np.all(np.isfinite(fa.u))


def test_coupling_outputs_are_held_between_coupling_steps(oracle_lib):
    """simulation.jl:151-154: fxOA/fyOA/trqOA only change when tstep % coupling.dt == 0."""
    f = synth.make_field(200, scale=0.9, walls="periodic", npoints=50, cache=False)
    model = host.Model(f.grid, f.ocean, f.atmos, f.domain, f.floes)
    sim = host.Simulation(model, consts=f.consts, dt=10, n_dt=3, coupling_settings=host.CouplingSettings(dt=2),
                          backend=oracle_lib)
    host.timestep_sim(sim, 0)
    a = sim.sync_host().fxOA.copy()
    assert np.any(a != 0)
    host.timestep_sim(sim, 1)
    b = sim.sync_host().fxOA.copy()
    host.timestep_sim(sim, 2)
    c = sim.sync_host().fxOA.copy()
    assert np.array_equal(a, b) and not np.array_equal(b, c)


def test_oracle_step_host_equals_three_calls(oracle_lib):
    f = synth.make_field(400, scale=1.01, walls="periodic", npoints=60, cache=False)
    fields.perturb_state(f.floes)
    ha, hb = synth.setup_handle(f, oracle_lib), synth.setup_handle(f, oracle_lib)
    fa, fb = ha.download_floes(mc=False), hb.download_floes(mc=False)
    for t in range(3):
        ha.upload_state(fa)
        ha.step(t, True)
        ha.download_floes(into=fa, mc=False)
        hb.step_host(fb, t, True)
        assert not compare_state(fb, fa, exact=STATE_FIELDS)


def test_settings_clamp_like_the_reference():
    # test_process_settings.jl:22-93
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        cs = host.CollisionSettings(floe_floe_max_overlap=1.5, floe_domain_max_overlap=-0.2)
        cp = host.CouplingSettings(dd=-3)
        assert cs.floe_floe_max_overlap == 1.0 and cs.floe_domain_max_overlap == 0.0 and cp.dd == 0
        assert len(w) == 3


def test_domain_validation_like_the_reference(oracle_lib):
    g = host.RegRectilinearGrid(0, 1e5, 0, 1e5, dx=1e4, dy=1e4)
    P, C = host.PeriodicBoundary, host.CollisionBoundary
    with pytest.raises(ValueError):
        host.Domain(P(host.North, g), C(host.South, g), C(host.East, g), C(host.West, g))  # domains.jl:11-33
    h = capi.Handle(oracle_lib)
    with pytest.raises(capi.SubzeroError):
        h.set_domain([1, 2, 2, 2], [1e5, 0, 1e5, 0], np.zeros(8), np.zeros(16))
    with pytest.raises(capi.SubzeroError):
        h.step_collisions()  # before set_domain


def test_synth_generator_properties():
    f = synth.make_field(500, scale=1.0, walls="periodic", npoints=100, cache=False)
    fa = f.floes
    assert abs(fa.area.sum() - f.L ** 2) < 1e-6 * f.L ** 2  # Voronoi cells tile the periodic box
    for i in (0, 17, 499):
        r = fa.ring(i)
        assert np.array_equal(r[0], r[-1]) and host.ring_area2(r) < 0  # closed, clockwise
        assert fa.area[i] == pytest.approx(abs(host.ring_area2(r)) / 2, rel=1e-12)
        assert fa.rmax[i] == pytest.approx(host.calc_max_radius(r, fa.centroid(i)), rel=1e-12)
        m = host.points_in_ring(fa.mc_x[fa.mc_offsets[i]:fa.mc_offsets[i + 1]],
                                fa.mc_y[fa.mc_offsets[i]:fa.mc_offsets[i + 1]], r - fa.centroid(i))
        assert m.all()


# ---- find_interp_knots (coupling.jl:702-797): the knot window never changes a bilinear result ------------------------------
def find_interp_knots(point_idx, ncells, glines, L, dd, periodic):
    """Python restatement of both methods (1-based indices like the reference)."""
    glines = np.asarray(glines, dtype=float)
    lo, hi = min(point_idx) - (dd + 1), max(point_idx) + (dd + 1)
    if not periodic:
        lo, hi = max(lo, 1), min(hi, ncells + 1)
        idx = np.arange(lo, hi + 1)
        return glines[idx - 1], idx
    low = np.arange(0)
    high = np.arange(0)
    if lo < 1 and hi > ncells:
        low, high, inb = np.arange(lo + ncells, ncells + 1), np.arange(1, hi - ncells + 1), np.arange(1, ncells + 1)
    elif lo < 1:
        low, inb = np.arange(lo + ncells, ncells + 1), np.arange(1, hi + 1)
    elif hi > ncells:
        high, inb = np.arange(1, hi - ncells + 1), np.arange(lo, ncells + 1)
    else:
        inb = np.arange(lo, hi + 1)
    idx = np.concatenate([low, inb, high])
    knots = np.concatenate([glines[low - 1] - L, glines[inb - 1], glines[high - 1] + L])
    return knots, idx


def test_find_interp_knots_reference_values():
    """test_coupling.jl:197-274, all nine cases."""
    g = np.arange(0.0, 81.0, 10.0)
    r = lambda a, b, c=1: np.arange(a, b + (1 if c > 0 else -1), c)
    cases = [
        ([4], 2, False, r(0, 60, 10), r(1, 7)), ([4], 2, True, r(0, 60, 10), r(1, 7)),
        ([0, 1], 2, False, r(0, 30, 10), r(1, 4)), ([0, 1], 2, True, r(-40, 30, 10), np.concatenate([r(5, 8), r(1, 4)])),
        ([8, 9], 1, False, r(50, 80, 10), r(6, 9)), ([8, 9], 1, True, r(50, 100, 10), np.concatenate([r(6, 8), r(1, 3)])),
        (list(range(1, 9)), 2, False, r(0, 80, 10), r(1, 9)),
        (list(range(1, 9)), 2, True, r(-30, 100, 10), np.concatenate([r(6, 8), r(1, 8), r(1, 3)])),
        (list(range(0, 10)), 2, False, r(0, 80, 10), r(1, 9)),
    ]
    for pidx, dd, per, knots, kidx in cases:
        k, i = find_interp_knots(pidx, 8, g, 80.0, dd, per)
        assert np.array_equal(k, knots) and np.array_equal(i, kidx), (pidx, dd, per)


@pytest.mark.parametrize("periodic", [False, True])
def test_knot_buffer_does_not_change_the_bilinear_result(periodic):
    """mc_interpolation (coupling.jl:845-902) interpolates on the window of knots around the floe; the library (and the
    oracle) interpolate on the full grid, wrapping periodic axes on lines 1..N.  For every buffer dd in {0, 1, 3} the
    windowed result equals the full-grid one exactly — which is why sz_config.coupling_dd is carried but not used."""
    rng = np.random.default_rng(5)
    N, dx = 8, 10.0
    g = np.arange(0.0, N * dx + 1, dx)
    field = rng.normal(size=N + 1)
    if periodic:
        field[N] = field[0]  # the reference never reads line N+1 on a periodic axis; make the comparison well defined
    xs = rng.uniform(-15.0, 95.0, 400) if periodic else rng.uniform(0.0, 80.0, 400)

    def full(x):
        gx = (x - g[0]) / dx
        i0 = int(np.floor(gx))
        w = gx - i0
        if periodic:
            a, b = i0 % N, (i0 + 1) % N
        else:
            if i0 >= N:
                i0, w = N - 1, 1.0
            a, b = i0, i0 + 1
        return field[a] + w * (field[b] - field[a])

    for dd in (0, 1, 3):
        for x in xs:
            nearest = int(np.floor((x - g[0]) / dx + 0.5)) + 1  # find_center_cell_index-style nearest grid line, 1-based
            knots, idx = find_interp_knots([nearest], N, g, N * dx, dd, periodic)
            vals = field[idx - 1]
            k = np.searchsorted(knots, x, side="right") - 1
            k = min(max(k, 0), len(knots) - 2)
            w = (x - knots[k]) / (knots[k + 1] - knots[k])
            assert abs((vals[k] + w * (vals[k + 1] - vals[k])) - full(x)) < 1e-12, (dd, x)


def check_step_host_partial(lib):
    """sz_step_host_partial: only the named fields cross the boundary; the others keep their device-resident values.
    Same bits as the device-resident sz_step (nothing uploaded) and as sz_step_host (a host edit of u and the status)."""
    f = synth.make_field(600, scale=1.01, walls="periodic", npoints=40, cache=False)
    fields.perturb_state(f.floes)
    ha, hb = synth.setup_handle(f, lib), synth.setup_handle(f, lib)
    fb = hb.download_floes(mc=False)
    down = ("centroid_x", "centroid_y", "alpha", "u", "v", "xi", "collision_force", "collision_trq", "status_tag", "fxOA")
    for t in range(2):
        ha.step(t, True)
        hb.step_host_partial(fb, t, True, upload=(), download=down)
    fa = ha.download_floes(mc=False)
    for name in down:
        assert np.array_equal(getattr(fa, name), getattr(fb, name)), name
    # a host process edits u and tags a floe: only those two fields go up
    fa_full = ha.download_floes(mc=False)
    fa_full.u[:50] *= 0.5
    fa_full.status_tag[3] = capi.STATUS_REMOVE
    fb.u[:50] *= 0.5
    fb.status_tag[3] = capi.STATUS_REMOVE
    ha.step_host(fa_full, 2, True)
    hb.step_host_partial(fb, 2, True, upload=("u", "status_tag"), download=down)
    for name in down:
        assert np.array_equal(getattr(fa_full, name), getattr(fb, name)), name


def test_step_host_partial_oracle(oracle_lib):
    check_step_host_partial(oracle_lib)


@pytest.mark.gpu
def test_step_host_partial_cuda(product_lib):
    check_step_host_partial(product_lib)
