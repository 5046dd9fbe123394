#!/bin/bash
# graph vs direct launches at several field sizes
for n in 1000 10000 100000; do for g in 0 1; do
  if [ $g = 1 ]; then export SZ_NO_GRAPH=1; else unset SZ_NO_GRAPH; fi
  python bench.py --floes $n --steps 100 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); print('floes %7d %s steps/s %8.1f dev %.3f ms launches/step %d' % ($n, 'direct' if os.environ.get('SZ_NO_GRAPH') else 'graph ', d['value'], d['device_ms_per_step'], d['gpu_launches'] // d['steps']))"
done; done
