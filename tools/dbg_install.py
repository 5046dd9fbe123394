import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import szload
from subzero_jl_b200 import capi, synth
from oracle import szo
lib = capi.product()
f = synth.make_field(400, scale=0.98, walls="collision", npoints=30, cache=False)
h = synth.setup_handle(f, lib)
c0 = np.diff(f.floes.mc_offsets)
offs, x, y, status = h.generate_subfloe_points(capi.POINTS_SUB_GRID, delta_g=400.0, install=True)
h.step_coupling(); A = h.download_floes(mc=False)
f2 = synth.make_field(400, scale=0.98, walls="collision", npoints=30, cache=False)
f2.floes.mc_offsets, f2.floes.mc_x, f2.floes.mc_y = offs, x, y
h2 = synth.setup_handle(f2, lib)
h2.step_coupling(); B = h2.download_floes(mc=False)
ho = synth.setup_handle(f2, szo.oracle())
ho.step_coupling(); O = ho.download_floes(mc=False)
i = 255
print("orig mc count of 255:", c0[i], "new:", np.diff(offs)[i], "status", f.floes.status_tag[i], f2.floes.status_tag[i])
print("install:", A.fxOA[i], A.fyOA[i], A.trqOA[i], A.status_tag[i])
print("upload :", B.fxOA[i], B.fyOA[i], B.trqOA[i], B.status_tag[i])
print("oracle :", O.fxOA[i], O.fyOA[i], O.trqOA[i], O.status_tag[i])
print("cx", f.floes.centroid_x[i], f.floes.centroid_y[i], f.floes.rmax[i], "L", f.L)
for name in ("fxOA", "fyOA", "trqOA", "hflx_factor"):
    print(name, "install vs oracle", np.abs(getattr(A, name) - getattr(O, name)).max(), "upload vs oracle", np.abs(getattr(B, name) - getattr(O, name)).max())
