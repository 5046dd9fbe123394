"""Worker of tests/test_slab.py::test_two_process_gloo_halo_exchange (importable by a spawned
process: sets up the import paths itself)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import szload  # noqa: E402,F401


def gloo_worker(rank, world, port, q):
    try:
        import torch
        import torch.distributed as dist
        import fields
        from oracle import szo
        from parity_util import STATE_FIELDS, compare_state
        from subzero_jl_b200 import slab, synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = szo.oracle()
        f = synth.make_field(700, scale=1.02, walls="periodic", npoints=20, cache=False)
        fields.perturb_state(f.floes)
        me = slab.partition_global(f.floes, world, f.L, skin=200.0, period_y=f.L)[rank]
        me.attach(synth.setup_handle(f, lib, threads=1))
        me.make_buffers(torch.device("cpu"))
        for t in range(3):
            if t == 2:
                slab.rebuild(me)  # collective: migration + fresh halo lists through all_gather_object
            me.exchange()
            me.h.step(t, True)
        h = synth.setup_handle(f, lib, threads=1)
        for t in range(3):
            h.step(t, True)
        ref = h.download_floes(mc=False)
        g, own = me.owned_state()
        bad = compare_state(own, slab.extract(ref, g), exact=STATE_FIELDS)
        bad = [b for b in bad if not b.startswith(("mc_offsets", "ghost_"))]
        if me.stale():
            bad.append("halo lists went stale")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, ["exception: %s\n%s" % (e, traceback.format_exc())]))


def tile_worker(rank, world, port, q, walls):
    """Weak-scaling tiles: every rank generates its own tile; compare with one rank holding all tiles."""
    try:
        import numpy as np
        import torch
        import torch.distributed as dist
        import fields
        from oracle import szo
        from parity_util import STATE_FIELDS, compare_state
        from subzero_jl_b200 import host, slab, synth
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        lib = szo.oracle()
        n = 500
        tiles = [synth.make_field(n, scale=1.02, walls="collision", npoints=20, cache=False, seed=1000 + r) for r in range(world)]
        for r, t in enumerate(tiles):
            fields.perturb_state(t.floes, seed=r)
            slab.shift_x(t.floes, r * t.L)
        L = tiles[0].L
        gfield = synth.tiled_model(tiles[0], world, walls)
        me = slab.partition_tiles(tiles[rank].floes, rank, world, L, world * L if walls == "shear" else None, skin=200.0)
        h = synth.setup_handle(gfield, lib, threads=1, floes=me.local)
        me.attach(h)
        me.make_buffers(torch.device("cpu"))
        for t in range(3):
            me.exchange()
            me.h.step(t, True)
        allf = slab.concat([t.floes for t in tiles])
        allf.id = np.arange(1, allf.n + 1, dtype=np.int64)
        hs = synth.setup_handle(gfield, lib, threads=1, floes=allf)
        for t in range(3):
            hs.step(t, True)
        ref = hs.download_floes(mc=False)
        g, own = me.owned_state()
        bad = compare_state(own, slab.extract(ref, g), exact=STATE_FIELDS, skip=())
        bad = [b for b in bad if not b.startswith(("mc_offsets", "ghost_", "id "))]
        if me.stale():
            bad.append("halo lists went stale")
        if me.local.n >= allf.n:
            bad.append("no decomposition happened")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, bad))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, ["exception: %s\n%s" % (e, traceback.format_exc())]))
