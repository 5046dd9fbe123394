"""Soak: CUDA product vs CPU oracle on many random fields (seeds, scales, wall kinds, flows), bit-exact collision outputs
and 1e-9 state.  Not part of the test suite (minutes of oracle time); run on the GPU box after kernel changes:
    python tools/soak.py [n_cases] [n_floes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import szload  # noqa
import fields
from parity_util import compare_collision_outputs, compare_state
from subzero_jl_b200 import capi, synth
from oracle import szo

ncase = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
prod, orc = capi.product(), szo.oracle()
rng = np.random.default_rng(2024)
bad_total = 0
t0 = time.time()
for k in range(ncase):
    scale = float(rng.choice([0.995, 1.0, 1.005, 1.01, 1.02, 1.04, 1.08, 1.3, 1.7, 2.5]))
    walls = str(rng.choice(["collision", "periodic", "shear"]))
    flow = str(rng.choice(["random", "converging"]))
    seed = int(rng.integers(1, 10**6))
    f = synth.make_field(n, scale=scale, walls=walls, flow=flow, npoints=40, seed=seed, cache=False)
    fields.perturb_state(f.floes, seed=seed % 97)
    hg, ho = synth.setup_handle(f, prod), synth.setup_handle(f, orc, threads=os.cpu_count())
    bad = []
    for step in range(3):   # three steps from the SAME state (re-uploaded): contacts develop, floes cross the walls
        for h in (hg, ho):
            h.add_ghosts(); h.step_collisions()
        b = compare_collision_outputs(hg, ho)
        if scale == 1.0:
            b = [x for x in b if not x.startswith("clip failures")]
        bad += ["step %d: %s" % (step, x) for x in b]
        for h in (hg, ho):
            h.remove_ghosts(); h.step_coupling(); h.step_floe_properties(step)
        bad += ["step %d: %s" % (step, x) for x in compare_state(hg.download_floes(), ho.download_floes())]
        hg.upload_floes(ho.download_floes())
    c = ho.counts()
    print("case %2d n=%d scale=%.3f %-9s %-10s seed=%6d overlap=%6d fuse=%4d -> %s" % (k, n, scale, walls, flow, seed, c["n_overlap"], c["n_fuse"], "OK" if not bad else "MISMATCH"), flush=True)
    for x in bad[:4]:
        print("    ", x)
    bad_total += bool(bad)
    hg.close(); ho.close()
print("soak: %d cases, %d mismatching, %.0f s" % (ncase, bad_total, time.time() - t0))
sys.exit(1 if bad_total else 0)
