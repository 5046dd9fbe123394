"""A few emulated 3-rank slab steps (periodic field, the library's push / unpack kernels) for compute-sanitizer.
NB: compute-sanitizer is closed on this pool (profiles/r2/r2a_compute_sanitizer_closed.txt); kept for other boxes."""
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
s = test_slab.run_slab(f, lib, 3, 2)
test_slab.check_against_single(f, lib, s, test_slab.single_rank_reference(f, lib, 2))
print("slab step ok")
