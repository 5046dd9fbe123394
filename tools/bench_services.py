"""Measure the host-process services on the bench workload: one JSON line each.

    python tools/bench_services.py [--floes 100000] [--cells 64] [--cpu-floes 20000]

Times are wall clock around the C-ABI call (host pair list / grid lines in, host results out: these are
service calls, their results are consumed on the host), median of 5 after 2 warm-up calls; the CPU oracle is
timed on a smaller field of the same generator and scaled by the floe count (both services are O(N))."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import szload  # noqa
from subzero_jl_b200 import capi, synth


def med(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def setup(n, lib, **kw):
    f = synth.make_field(n, scale=1.01, walls="collision", npoints=50, seed=n)
    h = synth.setup_handle(f, lib, **kw)
    h.step(0, True)
    return f, h


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--floes", type=int, default=100000)
    ap.add_argument("--cells", type=int, default=64)
    ap.add_argument("--cpu-floes", type=int, default=20000)
    a = ap.parse_args()
    from oracle import szo
    f, h = setup(a.floes, capi.product())
    fo, ho = setup(a.cpu_floes, szo.oracle(), threads=os.cpu_count())
    h.add_ghosts(); h.step_collisions(); ho.add_ghosts(); ho.step_collisions()
    pairs, pairs_o = h.pairs(0), ho.pairs(0)
    tg = med(lambda: h.pair_overlap_areas(pairs))
    to = med(lambda: ho.pair_overlap_areas(pairs_o), n=3, warm=1)
    print(json.dumps({"service": "sz_pair_overlap_areas", "floes": a.floes, "pairs": int(len(pairs)), "ms": 1e3 * tg,
                      "pairs_per_s": len(pairs) / tg, "h2d_bytes": int(pairs.nbytes), "d2h_bytes": int(len(pairs) * 9),
                      "cpu_oracle": {"floes": a.cpu_floes, "pairs": int(len(pairs_o)), "ms": 1e3 * to, "pairs_per_s": len(pairs_o) / to,
                                     "cores": os.cpu_count()}, "speedup_pairs_per_s": (len(pairs) / tg) / (len(pairs_o) / to)}))
    kinds = list(range(len(capi.GRID_OUTPUTS)))
    for cells in (a.cells, 10):
        xg, yg = np.linspace(0, f.L, cells + 1), np.linspace(0, f.L, cells + 1)
        xo, yo = np.linspace(0, fo.L, cells + 1), np.linspace(0, fo.L, cells + 1)
        tg = med(lambda: h.eulerian_data(xg, yg, kinds))
        to = med(lambda: ho.eulerian_data(xo, yo, kinds), n=1, warm=0)
        print(json.dumps({"service": "sz_eulerian_data", "floes": a.floes, "cells": [cells, cells], "outputs": len(kinds), "ms": 1e3 * tg,
                          "floes_per_s": a.floes / tg,
                          "cpu_oracle": {"floes": a.cpu_floes, "ms": 1e3 * to, "floes_per_s": a.cpu_floes / to, "cores": os.cpu_count(),
                                         "note": "the oracle keeps the reference's O(cells x floes) candidate mask"},
                          "speedup_floes_per_s": (a.floes / tg) / (a.cpu_floes / to)}))


if __name__ == "__main__":
    main()
