"""Slab decomposition inside the library (sz_slab_*): the owned floes of every rank must be
BIT-IDENTICAL to the single-rank run — same pair orientation, candidate order, row order and canonical image
pair, because the local lists are sorted by global index.

CPU (`not gpu`): the product's slab host logic (csrc/sz_slab.cpp) compiled against the oracle's ABI — ranks of one
process, and real 2-process gloo runs where the library's set-up / rebuild messages travel through its alltoallv
callback.  GPU: the CUDA library — ranks emulated on one device (the peer-memory push / unpack kernels, flags and
double-buffered arenas are the same code), and on two devices when the box has them."""
import os
import sys

import numpy as np
import pytest

import fields
from parity_util import STATE_FIELDS, compare_state
from subzero_jl_b200 import capi, slab, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def single_rank_reference(f, lib, steps, coupling=True, floes=None):
    h = synth.setup_handle(f, lib, floes=floes)
    for t in range(steps):
        h.step(t, coupling)
    return h.download_floes(mc=False)


def check_against_single(f, lib, s, ref, n_total=None):
    """Every floe is owned exactly once and its state equals the single-rank run bit for bit."""
    n_total = f.floes.n if n_total is None else n_total
    seen = np.zeros(n_total, dtype=bool)
    for k in range(s.n_local):
        g, own = s.owned_state(k)
        assert not seen[g].any()
        seen[g] = True
        bad = compare_state(own, slab.extract(ref, g), exact=STATE_FIELDS, skip=())
        bad = [b for b in bad if not b.startswith(("mc_offsets", "ghost_"))]
        assert not bad, "rank %d: %s" % (s.rank_first + k, "\n".join(bad))
    if s.n_local == s.world:
        assert seen.all()
    return seen


def check_halo_copies_equal_owners(s):
    """After sz_slab_refresh_halo every halo copy carries its owner's current dynamic state bit for bit."""
    s.refresh_halo()
    state, index = [], []
    for k in range(s.n_local):
        index.append(s.local_index(k))
        state.append(s.handles[k].download_floes(mc=False))
    n_copies = 0
    for k in range(s.n_local):
        g, o = index[k]
        for src in range(s.n_local):
            if src == k:
                continue
            sel = np.nonzero(o == src)[0]
            if len(sel) == 0:
                continue
            gs, _ = index[src]
            pos = np.searchsorted(gs, g[sel])
            assert np.array_equal(gs[pos], g[sel])
            a, b = slab.extract(state[k], sel), slab.extract(state[src], pos)
            for name in ("centroid_x", "centroid_y", "u", "v", "xi", "alpha", "height", "status_tag", "vert_xy"):
                assert np.array_equal(getattr(a, name), getattr(b, name)), (k, src, name)
            n_copies += len(sel)
    assert n_copies > 0


def run_slab(f, lib, world, steps, skin=200.0, devices=None, coupling=True, host=False):
    s = slab.Slab(lib, f, world, skin=skin, devices=devices)
    s.build(f.floes)
    if host:
        arrays = [s.handles[k].download_floes(mc=False) for k in range(world)]
    for t in range(steps):
        if host:
            s.step_host(arrays, t, coupling)
        else:
            s.step(t, coupling)
    return s


@pytest.mark.parametrize("walls,world", [("collision", 2), ("collision", 3), ("periodic", 2), ("periodic", 4), ("shear", 3)])
def test_emulated_ranks_match_single_rank_oracle(walls, world, oracle_lib):
    f = synth.make_field(1600, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    s = run_slab(f, oracle_lib, world, 3)
    counts = [s.local_index(k) for k in range(world)]
    assert sum(int((o == k).sum()) for k, (g, o) in enumerate(counts)) == f.floes.n
    assert all(len(g) < f.floes.n for g, o in counts)          # a real decomposition ...
    assert all((o != k).any() for k, (g, o) in enumerate(counts))  # ... with halo copies
    check_against_single(f, oracle_lib, s, single_rank_reference(f, oracle_lib, 3))
    assert not s.stale()
    check_halo_copies_equal_owners(s)
    s.step(3, True)  # the refresh does not disturb the exchange protocol
    check_against_single(f, oracle_lib, s, single_rank_reference(f, oracle_lib, 4))


@pytest.mark.parametrize("walls,world", [("periodic", 2), ("collision", 3)])
def test_emulated_ranks_through_host_arrays_oracle(walls, world, oracle_lib):
    f = synth.make_field(1200, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    s = run_slab(f, oracle_lib, world, 3, host=True)
    check_against_single(f, oracle_lib, s, single_rank_reference(f, oracle_lib, 3))


def run_partial(f, lib, world):
    """Three sz_slab_step_host_partial steps: the host halves every velocity and uploads ONLY u, v on the first step (the
    neighbours' halo copies must see it), later steps upload the status only; positions, velocities and forces come back."""
    s = slab.Slab(lib, f, world, skin=200.0)
    s.build(f.floes)
    arrays = [s.handles[k].download_floes(mc=False) for k in range(world)]
    for a in arrays:
        a.u *= 0.5
        a.v *= 0.5
    down = ("centroid_x", "centroid_y", "alpha", "u", "v", "xi", "collision_force", "collision_trq", "fxOA", "fyOA", "trqOA", "status_tag")
    for a in arrays:
        a.height[:] = -1.0  # not exchanged in either direction: must neither reach the device nor be overwritten
    s.step_host_partial(arrays, 0, True, upload=("u", "v"), download=down)
    s.step_host_partial(arrays, 1, True, upload=("status_tag",), download=down)
    s.step_host_partial(arrays, 2, True, upload=(), download=down)
    for k, a in enumerate(arrays):
        assert np.all(a.height == -1.0)
        dev = s.handles[k].download_floes(mc=False)
        assert np.all(dev.height > 0.0)
        own = s.local_index(k)[1] == k
        for name in down:
            assert np.array_equal(getattr(a, name)[own], getattr(dev, name)[own]), name
    return s


@pytest.mark.parametrize("walls,world", [("periodic", 2), ("collision", 3)])
def test_partial_host_arrays_oracle(walls, world, oracle_lib):
    f = synth.make_field(1200, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    s = run_partial(f, oracle_lib, world)
    f.floes.u *= 0.5
    f.floes.v *= 0.5
    check_against_single(f, oracle_lib, s, single_rank_reference(f, oracle_lib, 3))


@pytest.mark.gpu
def test_partial_host_arrays_on_one_gpu(product_lib):
    f = synth.make_field(3000, scale=1.01, walls="periodic", npoints=40, cache=False)
    fields.perturb_state(f.floes)
    s = run_partial(f, product_lib, 3)
    f.floes.u *= 0.5
    f.floes.v *= 0.5
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 3))
    s.step(3, True)  # a device-resident step behind a host step re-publishes
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 4))


@pytest.mark.parametrize("walls,world", [("periodic", 3), ("collision", 2)])
def test_rebuild_migrates_ownership_and_keeps_results(walls, world, oracle_lib):
    """Fast floes + a small skin: the halo lists go stale, the library rebuilds by itself (ownership migrates to the
    slab the centroid moved into, halo lists are renewed) and the owned results still equal the single-rank run."""
    f = synth.make_field(1200, scale=1.02, walls=walls, npoints=20, cache=False)
    fields.perturb_state(f.floes)
    f.floes.u = f.floes.u * 40.0  # up to 4 m/s: 40 m per step
    f.floes.v = f.floes.v * 40.0
    s = slab.Slab(oracle_lib, f, world, skin=60.0)
    s.build(f.floes)
    owner0 = [s.local_index(k) for k in range(world)]
    for t in range(8):
        s.step(t, True)
    assert s.stats()["rebuilds"] >= 2
    owner1 = [s.local_index(k) for k in range(world)]
    assert any(not np.array_equal(a[0][a[1] == k], b[0][b[1] == k]) for k, (a, b) in enumerate(zip(owner0, owner1)))  # something migrated
    check_against_single(f, oracle_lib, s, single_rank_reference(f, oracle_lib, 8))


def test_any_initial_distribution_and_explicit_edges(oracle_lib):
    """sz_slab_build accepts any initial distribution (here: odd / even global indices on two ranks) and explicit
    slab edges; a rank may even start with nothing."""
    f = synth.make_field(900, scale=1.02, walls="collision", npoints=20, cache=False)
    fields.perturb_state(f.floes)
    ref = single_rank_reference(f, oracle_lib, 2)
    for parts in ([np.arange(0, 900, 2), np.arange(1, 900, 2), None], [None, None, np.arange(900)]):
        s = slab.Slab(oracle_lib, f, 3, skin=200.0)
        s.set_edges([-np.inf, 0.3 * f.L, 0.55 * f.L, np.inf])
        s.build([slab.extract(f.floes, p) if p is not None else None for p in parts], parts)
        for t in range(2):
            s.step(t, True)
        check_against_single(f, oracle_lib, s, ref)
        g, o = s.local_index(1)
        own = g[o == 1]
        cx = f.floes.centroid_x[own]
        assert ((cx >= 0.3 * f.L) & (cx < 0.55 * f.L)).all()


def _spawn(target, world, port, *args):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=target, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, bad in res:
        assert not bad, "rank %d: %s" % (rank, "\n".join(bad))


def test_two_process_gloo_halo_exchange():
    import slab_worker
    _spawn(slab_worker.gloo_worker, 2, 29500 + (os.getpid() % 2000))


@pytest.mark.parametrize("walls", ["collision", "shear"])
def test_two_process_gloo_weak_scaling_tiles(walls):
    """bench.py's N > 1 construction: every rank generates its own tile and hands it to sz_slab_build; the library
    finds the neighbours' boundary floes; owned results equal the single-rank run of all tiles."""
    import slab_worker
    _spawn(slab_worker.tile_worker, 2, 31500 + (os.getpid() % 2000) + (7 if walls == "shear" else 0), walls)


# ---- CUDA ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("walls,world", [("collision", 2), ("periodic", 3), ("shear", 4)])
def test_emulated_ranks_on_one_gpu(walls, world, product_lib):
    """k ranks on one device: local-list construction, peer-memory push / flag / unpack kernels and double-buffered
    arenas reproduce the single-handle CUDA run bit for bit."""
    f = synth.make_field(3000, scale=1.01, walls=walls, npoints=40, cache=False)
    fields.perturb_state(f.floes)
    s = run_slab(f, product_lib, world, 4)
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 4))
    assert s.stats()["send_bytes_per_step"] > 0
    check_halo_copies_equal_owners(s)
    s.step(4, True)  # the refresh does not disturb the epoch protocol
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 5))


@pytest.mark.gpu
def test_emulated_ranks_through_host_arrays_on_one_gpu(product_lib):
    """sz_slab_step_host on CUDA: uploads, publication of the uploaded state, halo update on top of the stale copies,
    step, overlapped downloads — bit-identical to the single-handle device-resident run; then device-resident steps
    and host steps mixed (the epoch protocol re-publishes)."""
    f = synth.make_field(3000, scale=1.01, walls="periodic", npoints=40, cache=False)
    fields.perturb_state(f.floes)
    s = run_slab(f, product_lib, 3, 3, host=True)
    ref = single_rank_reference(f, product_lib, 3)
    check_against_single(f, product_lib, s, ref)
    s.step(3, True)
    arrays = [s.handles[k].download_floes(mc=False) for k in range(3)]
    s.step_host(arrays, 4, True)
    s.step(5, True)
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 6))


@pytest.mark.gpu
@pytest.mark.parametrize("walls,world", [("periodic", 3), ("collision", 2)])
def test_rebuild_on_gpu_keeps_monte_carlo_points_resident(walls, world, product_lib):
    """Device-side displacement test + rebuild on CUDA: Monte-Carlo points of floes that stay are re-gathered on the
    device, only migrants travel; owned results equal the single-handle run (coupling included, 1e-9)."""
    f = synth.make_field(2500, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    f.floes.u = f.floes.u * 40.0
    f.floes.v = f.floes.v * 40.0
    s = slab.Slab(product_lib, f, world, skin=60.0)
    s.build(f.floes)
    for t in range(8):
        s.step(t, True)
    assert s.stats()["rebuilds"] >= 2
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 8))


@pytest.mark.gpu
def test_two_gpus_one_process_bit_identical(product_lib):
    """The real transport: two devices of one box driven by ONE host process (what the Julia shim does) — peer-mapped
    arenas over NVLink.  Skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    f = synth.make_field(20000, scale=1.01, walls="shear", npoints=60, cache=False)
    fields.perturb_state(f.floes)
    s = run_slab(f, product_lib, 2, 5, skin=500.0, devices=[0, 1])
    check_against_single(f, product_lib, s, single_rank_reference(f, product_lib, 5))


@pytest.mark.gpu
def test_two_gpus_two_processes_bit_identical():
    """One process per GPU (torchrun's layout): cudaIpc-mapped arenas, the set-up messages over gloo.  Skipped on a
    one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import slab_worker
    _spawn(slab_worker.cuda_worker, 2, 33500 + (os.getpid() % 2000))
