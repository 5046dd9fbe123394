"""Where does an sz_step_host call spend its time?  (development probe, run on the GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import szload  # noqa
import torch
from subzero_jl_b200 import capi, synth
sys.path.insert(0, ROOT)
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
f = synth.make_field(n, scale=1.01, walls="collision", npoints=1000, seed=n)
h = synth.setup_handle(f, capi.product())
for t in range(3):
    h.step(t, True)
fa = bench.pin_floe_arrays(h.download_floes(mc=False))
for mode in ("host", "separate", "host"):
    rows = []
    for t in range(12):
        t0 = time.perf_counter()
        if mode == "host":
            h.step_host(fa, t, True)
        else:
            h.upload_state(fa); t1 = time.perf_counter(); h.step(t, True); t2 = time.perf_counter(); h.download_floes(into=fa, mc=False)
        t3 = time.perf_counter()
        ms = h.timings_raw()
        rows.append([1e3 * (t3 - t0)] + list(ms[:7]) + ([1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)] if mode != "host" else []))
    r = np.median(np.array(rows[2:]), axis=0)
    print(mode, "wall %.3f | ghosts %.3f broad %.3f narrow %.3f rows %.3f coupling %.3f update %.3f device %.3f" % tuple(r[:8]),
          ("| up %.3f step %.3f down %.3f" % tuple(r[8:])) if mode != "host" else "")
