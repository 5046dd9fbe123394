#!/usr/bin/env python
"""Phase times of the default bench field with and without the overlapped coupling (how much the collision chain pays
for sharing the SMs with k_coupling).  One JSON line.  usage: python tools/phase_probe.py [steps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import szload  # noqa: F401,E402
from subzero_jl_b200 import capi, synth  # noqa: E402

NAMES = ("ghosts", "broad", "narrow", "rows", "coupling", "update", "step_device")


def run(h, steps, coupling):
    for t in range(10):
        h.step(t, coupling)
    acc = np.zeros(8)
    for t in range(steps):
        h.step(10 + t, coupling)
        acc += h.timings_raw()
    return {k: round(float(acc[i] / steps), 4) for i, k in enumerate(NAMES)}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    f = synth.make_field(100000, scale=1.01, walls="collision", npoints=1000, seed=100000)
    h = synth.setup_handle(f, capi.product(), device=0)
    out = {"with_coupling": run(h, steps, True), "collisions_only": run(h, steps, False)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
