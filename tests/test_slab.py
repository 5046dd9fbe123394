"""Slab decomposition: the owned floes of every rank must be BIT-IDENTICAL to the
single-rank run — same pair orientation, candidate order, row order and canonical image pair,
because the local lists are sorted by global index.  CPU: ranks emulated in one process and a real
2-process gloo run, both on the oracle; GPU: ranks emulated on one device with the CUDA pack/unpack."""
import os
import sys

import numpy as np
import pytest

import fields
from parity_util import STATE_FIELDS, compare_state
from subzero_jl_b200 import capi, slab, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXACT = tuple(n for n in STATE_FIELDS if n not in ("fxOA", "fyOA", "trqOA", "hflx_factor"))


def make_handle(f, lib, **kw):
    h = synth.setup_handle(f, lib, **kw)
    return h


def run_decomposed(f, lib, world, steps, device="cpu", walls_period=None, coupling=True):
    import torch
    period = f.L if walls_period else None
    ranks = slab.partition_global(f.floes, world, period, skin=200.0, period_y=f.L if f.walls == "periodic" else None)
    for r in ranks:
        h = make_handle(f, lib)  # grid, fields, domain of the GLOBAL model; floes replaced below
        r.attach(h)
        r.make_buffers(torch.device(device))
    for t in range(steps):
        if coupling:
            for r in ranks:
                r.h.coupling_begin()   # CUDA, no periodic wall on the rank: the coupling runs beside the exchange
        slab.exchange_local(ranks)
        for r in ranks:
            r.h.step(t, coupling)
    return ranks


def run_decomposed_host(f, lib, world, steps, device="cpu", walls_period=None):
    """The same decomposition driven through HOST arrays every step: sz_upload_state_begin -> halo exchange ->
    sz_step_host(in = NULL): the halo update must land on top of the (stale) uploaded halo copies."""
    import torch
    period = f.L if walls_period else None
    ranks = slab.partition_global(f.floes, world, period, skin=200.0, period_y=f.L if f.walls == "periodic" else None)
    host = []
    for r in ranks:
        r.attach(make_handle(f, lib))
        r.make_buffers(torch.device(device))
        host.append(r.h.download_floes(mc=False))
    for t in range(steps):
        for r, fa in zip(ranks, host):
            r.h.upload_state_begin(fa, True)
            r.h.coupling_begin()
        slab.exchange_local(ranks)
        for r, fa in zip(ranks, host):
            r.h.step_host(None, t, True, out=fa)
    return ranks


def check_against_single(f, lib, ranks, steps, coupling=True):
    h = make_handle(f, lib)
    for t in range(steps):
        h.step(t, coupling)
    ref = h.download_floes(mc=False)
    seen = np.zeros(f.floes.n, dtype=bool)
    for r in ranks:
        g, own = r.owned_state()
        assert not seen[g].any()
        seen[g] = True
        want = slab.extract(ref, g)
        bad = compare_state(own, want, exact=STATE_FIELDS, skip=())
        bad = [b for b in bad if not b.startswith(("mc_offsets", "ghost_"))]
        assert not bad, "rank %d: %s" % (r.rank, "\n".join(bad))
        assert not r.stale()
    assert seen.all()


@pytest.mark.parametrize("walls,world", [("collision", 2), ("collision", 3), ("periodic", 2), ("periodic", 4), ("shear", 3)])
def test_emulated_ranks_match_single_rank_oracle(walls, world, oracle_lib):
    f = synth.make_field(1600, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    ranks = run_decomposed(f, oracle_lib, world, 3, walls_period=walls in ("periodic", "shear"))
    assert sum(int(r.owned.sum()) for r in ranks) == f.floes.n
    assert all(r.local.n < f.floes.n for r in ranks)
    check_against_single(f, oracle_lib, ranks, 3)


@pytest.mark.parametrize("walls,world", [("periodic", 2), ("collision", 3)])
def test_emulated_ranks_through_host_arrays_oracle(walls, world, oracle_lib):
    f = synth.make_field(1200, scale=1.02, walls=walls, npoints=30, cache=False)
    fields.perturb_state(f.floes)
    ranks = run_decomposed_host(f, oracle_lib, world, 3, walls_period=walls in ("periodic", "shear"))
    check_against_single(f, oracle_lib, ranks, 3)


@pytest.mark.parametrize("walls,world", [("periodic", 3), ("collision", 2)])
def test_rebuild_migrates_ownership_and_keeps_results(walls, world, oracle_lib):
    """Fast floes + a small skin: the halo lists go stale, every rank rebuilds (ownership migrates to the
    slab the centroid moved into, halo lists are renewed) and the owned results still equal the single-rank run."""
    import torch
    f = synth.make_field(1200, scale=1.02, walls=walls, npoints=20, cache=False)
    fields.perturb_state(f.floes)
    f.floes.u = f.floes.u * 40.0  # up to 4 m/s: 40 m per step
    f.floes.v = f.floes.v * 40.0
    period = f.L if walls == "periodic" else None
    ranks = slab.partition_global(f.floes, world, period, skin=60.0, period_y=period)
    for r in ranks:
        r.attach(make_handle(f, oracle_lib))
        r.make_buffers(torch.device("cpu"))
    owner0 = [r.gidx[r.owned].copy() for r in ranks]
    rebuilds = 0
    for t in range(8):
        if any(r.stale() for r in ranks):
            slab.rebuild_local(ranks)
            rebuilds += 1
        slab.exchange_local(ranks)
        for r in ranks:
            r.h.step(t, True)
    assert rebuilds >= 2
    assert any(not np.array_equal(o, r.gidx[r.owned]) for o, r in zip(owner0, ranks))  # something migrated
    h = make_handle(f, oracle_lib)
    for t in range(8):
        h.step(t, True)
    ref = h.download_floes(mc=False)
    seen = np.zeros(f.floes.n, dtype=bool)
    for r in ranks:
        g, own = r.owned_state()
        seen[g] = True
        bad = compare_state(own, slab.extract(ref, g), exact=STATE_FIELDS)
        bad = [b for b in bad if not b.startswith(("mc_offsets", "ghost_"))]
        assert not bad, "rank %d: %s" % (r.rank, "\n".join(bad))
    assert seen.all()


def test_two_process_gloo_halo_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    import slab_worker
    procs = [ctx.Process(target=slab_worker.gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, bad in res:
        assert not bad, "rank %d: %s" % (rank, "\n".join(bad))


@pytest.mark.parametrize("walls", ["collision", "shear"])
def test_two_process_gloo_weak_scaling_tiles(walls):
    """bench.py's N > 1 construction: every rank generates its own tile and learns the neighbours'
    boundary floes at set-up; owned results equal the single-rank run of all tiles."""
    import torch.multiprocessing as mp
    import slab_worker
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + (7 if walls == "shear" else 0)
    procs = [ctx.Process(target=slab_worker.tile_worker, args=(r, 2, port, q, walls)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, bad in res:
        assert not bad, "rank %d: %s" % (rank, "\n".join(bad))


@pytest.mark.gpu
def test_emulated_ranks_through_host_arrays_on_one_gpu(product_lib):
    """sz_upload_state_begin -> stream-ordered halo exchange -> sz_step_host(in = NULL) on CUDA: bit-identical to the
    single-handle device-resident run."""
    f = synth.make_field(3000, scale=1.01, walls="periodic", npoints=40, cache=False)
    fields.perturb_state(f.floes)
    ranks = run_decomposed_host(f, product_lib, 3, 3, device="cuda", walls_period=True)
    check_against_single(f, product_lib, ranks, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("walls,world", [("collision", 2), ("periodic", 3)])
def test_emulated_ranks_on_one_gpu(walls, world, product_lib):
    """The CUDA pack / unpack kernels and the local-list construction: k ranks emulated on one device
    must reproduce the single-handle CUDA run bit for bit (collisions) / to 1e-9 (everything)."""
    f = synth.make_field(3000, scale=1.01, walls=walls, npoints=40, cache=False)
    fields.perturb_state(f.floes)
    ranks = run_decomposed(f, product_lib, world, 3, device="cuda", walls_period=walls == "periodic")
    check_against_single(f, product_lib, ranks, 3)
