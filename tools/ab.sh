#!/bin/bash
# A/B: bench the in-tree library against other builds of it (paths relative to the repo root)
for lib in "" "$@"; do
  echo "== ${lib:-libsubzero_b200.so}"
  SZ_B200_LIB=${lib:+$PWD/$lib} SZ_DEBUG_COUNTS=${SZ_DEBUG_COUNTS:-} python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline 2> /tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['roofline']['phase_ms']; print('steps/s %.1f  dev %.3f | broad %.3f narrow %.3f rows %.3f coupling %.3f update %.3f' % (d['value'], d['device_ms_per_step'], p['broad'], p['narrow'], p['rows'], p['coupling'], p['update']))"
  tail -1 /tmp/ab.err
done
