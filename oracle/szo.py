"""ctypes binding of the CPU oracle (oracle/libszo.so, prefix `szo_`).

TEST INFRASTRUCTURE: import this only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports it.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import szload  # noqa: E402,F401  (registers subzero_jl_b200)
from subzero_jl_b200 import capi  # noqa: E402

LIB = os.path.join(HERE, "libszo.so")
_oracle = None


def build(force=False):
    root = os.path.dirname(HERE)
    src = [os.path.join(HERE, f) for f in ("szo.c", "szo_geom.h", "Makefile")] + [
        os.path.join(root, "include", "subzero_b200.h"),
        os.path.join(root, "subzero.jl_b200", "csrc", "sz_slab.cpp"),  # the product's slab host logic on the oracle's ABI
        os.path.join(root, "subzero.jl_b200", "csrc", "sz_slab_backend.h")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(LIB):
            build()
        _oracle = capi.Library(LIB, "szo_")
    return _oracle
