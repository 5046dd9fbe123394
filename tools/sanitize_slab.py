"""One emulated 3-rank slab step (periodic field, CUDA pack / unpack kernels) for compute-sanitizer."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import szload  # noqa: E402,F401
import fields  # noqa: E402
import test_slab  # noqa: E402
from subzero_jl_b200 import capi, synth  # noqa: E402

f = synth.make_field(600, scale=1.01, walls="periodic", npoints=30, cache=False)
fields.perturb_state(f.floes)
lib = capi.product()
ranks = test_slab.run_decomposed(f, lib, 3, 2, device="cuda", walls_period=True)
test_slab.check_against_single(f, lib, ranks, 2)
print("slab step ok")
