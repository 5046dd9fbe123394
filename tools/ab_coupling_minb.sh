#!/bin/bash
# A/B of k_coupling_bulk's register target (__launch_bounds__(128, CP_MINB)): builds one library per value into
# build_ab/ (git-ignored, travels with gpurun) and prints the command that times them on the GPU box.
#   tools/ab_coupling_minb.sh 4 5 7 8
# Round-2 result at 100 k floes (step / coupling ms): 4 -> 1.110 / 0.657, 5 -> 1.086 / 0.492, 6 -> 1.079 / 0.469 (shipped),
# 7 -> 1.091 / 0.474, 8 -> 1.116 / 0.469 (profiles/r2/r3h_*, r3i_*).
set -e
cd "$(dirname "$0")/../subzero.jl_b200/csrc"
make -s
mkdir -p ../../build_ab
ARCH="-gencode arch=compute_100a,code=sm_100a"
for m in "$@"; do
    nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DCP_MINB=$m -c sz_kernels_fp.cu -o /tmp/sz_fp_minb$m.o
    nvcc $ARCH -shared -o ../../build_ab/libsz_minb$m.so sz_kernels.o /tmp/sz_fp_minb$m.o sz_services.o sz_api.o sz_slab.o
    echo "built build_ab/libsz_minb$m.so"
done
echo "gpurun -- 'for m in $*; do SZ_B200_LIB=\$PWD/build_ab/libsz_minb\$m.so python tools/phase_probe.py 100; done'"
