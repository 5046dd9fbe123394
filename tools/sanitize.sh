#!/bin/bash
# compute-sanitizer memcheck + racecheck on smoke() and on an emulated 3-rank slab step.
# usage (GPU box): bash tools/sanitize.sh <tag>   -> gpurun_out/<tag>_{memcheck,racecheck}_{smoke,slab}.log
tag=${1:-san}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $CS --tool $tool --log-file gpurun_out/${tag}_${tool}_smoke.log \
      python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_${tool}_smoke.out 2>&1
  echo "$tool smoke rc=$?"
  timeout 1200 $CS --tool $tool --log-file gpurun_out/${tag}_${tool}_slab.log \
      python tools/sanitize_slab.py > gpurun_out/${tag}_${tool}_slab.out 2>&1
  echo "$tool slab rc=$?"
done
tail -n 3 gpurun_out/${tag}_*_*.log
