"""Workers of tests/test_slab.py's multi-process tests (importable by a spawned process: sets up the import
paths itself).  One rank per process: the library's set-up / rebuild messages go through its alltoallv callback
(torch.distributed point-to-point on gloo, slab._make_transport)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import szload  # noqa: E402,F401


def _compare_owned(s, ref, extra_skip=()):
    from parity_util import STATE_FIELDS, compare_state
    from subzero_jl_b200 import slab
    g, own = s.owned_state(0)
    bad = compare_state(own, slab.extract(ref, g), exact=STATE_FIELDS, skip=())
    return [b for b in bad if not b.startswith(("mc_offsets", "ghost_") + tuple(extra_skip))]


def gloo_worker(rank, world, port, q):
    """Global list known everywhere; rank r brings the floes with gidx % world == r (ANY distribution works).  A forced
    rebuild in the middle migrates ownership through the callback."""
    try:
        import numpy as np
        import torch.distributed as dist
        import fields
        from oracle import szo
        from subzero_jl_b200 import slab, synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = szo.oracle()
        f = synth.make_field(700, scale=1.02, walls="periodic", npoints=20, cache=False)
        fields.perturb_state(f.floes)
        s = slab.Slab(lib, f, world, rank=rank, skin=200.0, threads=1)
        s.set_edges(np.concatenate([[0.0], np.quantile(f.floes.centroid_x, np.arange(1, world) / world), [f.L]]))
        mine = np.arange(rank, f.floes.n, world)
        s.build([slab.extract(f.floes, mine)], [mine])
        for t in range(3):
            if t == 2:
                s.rebuild()  # collective: migration + fresh halo lists
            s.step(t, True)
        # host arrays of the local list (sz_slab_step_host), then a field-masked step (sz_slab_step_host_partial): the
        # uploaded boundary floes reach the other PROCESS before its step
        arrays = [s.handles[0].download_floes(mc=False)]
        s.step_host(arrays, 3, True)
        s.step_host_partial(arrays, 4, True, upload=("status_tag",), download=("centroid_x", "centroid_y", "u", "v", "xi", "alpha", "status_tag"))
        h = synth.setup_handle(f, lib, threads=1)
        for t in range(5):
            h.step(t, True)
        bad = _compare_owned(s, h.download_floes(mc=False))
        g, o = s.local_index(0)
        own = o == rank
        dev = s.handles[0].download_floes(mc=False)
        for name in ("centroid_x", "u", "alpha"):
            if not np.array_equal(getattr(arrays[0], name)[own], getattr(dev, name)[own]):
                bad.append("masked download of %s differs from the device state" % name)
        if not (o != rank).any() or len(g) >= f.floes.n:
            bad.append("no decomposition happened")
        if s.stale():
            bad.append("halo lists went stale")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, ["exception: %s\n%s" % (e, traceback.format_exc())]))


def _tiles(world, n, walls):
    import numpy as np
    import fields
    from subzero_jl_b200 import slab, synth
    tiles = [synth.make_field(n, scale=1.02, walls="collision", npoints=20, cache=False, seed=1000 + r) for r in range(world)]
    for r, t in enumerate(tiles):
        fields.perturb_state(t.floes, seed=r)
        slab.shift_x(t.floes, r * t.L)
        t.floes.id = r * n + np.arange(1, n + 1, dtype=np.int64)  # ids must be unique over all tiles (collisions.jl:751-758)
    gfield = synth.tiled_model(tiles[0], world, walls)
    allf = slab.concat([t.floes for t in tiles])
    return tiles, gfield, allf


def tile_worker(rank, world, port, q, walls):
    """Weak-scaling tiles: every rank generates its own tile; compare with one rank holding all tiles."""
    try:
        import numpy as np
        import torch.distributed as dist
        from oracle import szo
        from subzero_jl_b200 import slab, synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = szo.oracle()
        n = 500
        tiles, gfield, allf = _tiles(world, n, walls)
        s = slab.Slab(lib, gfield, world, rank=rank, skin=200.0, threads=1)
        s.set_edges(slab.tile_edges(world, tiles[0].L, walls == "shear"))
        s.build([tiles[rank].floes], [rank * n + np.arange(n, dtype=np.int64)])
        for t in range(3):
            s.step(t, True)
        hs = synth.setup_handle(gfield, lib, threads=1, floes=allf)
        for t in range(3):
            hs.step(t, True)
        bad = _compare_owned(s, hs.download_floes(mc=False))
        g, o = s.local_index(0)
        if s.stale():
            bad.append("halo lists went stale")
        if len(g) >= allf.n or not (o != rank).any():
            bad.append("no decomposition happened")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, ["exception: %s\n%s" % (e, traceback.format_exc())]))


def cuda_worker(rank, world, port, q):
    """One process per GPU on the CUDA library: arenas mapped with cudaIpc, flags over NVLink; a forced rebuild in the
    middle.  Compared with a single-handle run on this rank's own device."""
    try:
        import numpy as np
        import torch
        import torch.distributed as dist
        import fields
        from subzero_jl_b200 import capi, slab, synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(rank)
        lib = capi.product()
        f = synth.make_field(20000, scale=1.01, walls="shear", npoints=60, cache=False)
        fields.perturb_state(f.floes)
        s = slab.Slab(lib, f, world, rank=rank, skin=500.0, devices=[rank], device=rank)
        s.set_edges(np.concatenate([[0.0], np.quantile(f.floes.centroid_x, np.arange(1, world) / world), [f.L]]))
        mine = np.arange(rank, f.floes.n, world)
        s.build([slab.extract(f.floes, mine)], [mine])
        for t in range(6):
            if t == 3:
                s.rebuild()
            s.step(t, True)
        h = synth.setup_handle(f, lib, device=rank)
        for t in range(6):
            h.step(t, True)
        bad = _compare_owned(s, h.download_floes(mc=False))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, ["exception: %s\n%s" % (e, traceback.format_exc())]))
