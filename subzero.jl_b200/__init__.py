"""subzero.jl_b200 — B200-native (sm_100a) floe-interaction hot path of Subzero.jl.

Contents (only what the hot path needs):
  csrc/      hand-written CUDA kernels + the C-ABI shared library (include/subzero_b200.h)
  capi.py    ctypes binding of that C ABI (what a Julia `ccall` shim binds)
  host.py    host-side mirror of the reference's API for this path (Simulation / Model /
             Constants / CollisionSettings / timestep_sim! ...), used by tests and bench
  synth.py   synthetic Voronoi-packed floe fields
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
