#!/bin/bash
# quick A/B of library builds: scripts/gpu_variants.sh lib1.so lib2.so ...
for lib in "" "$@"; do
  echo "=== ${lib:-default}"
  D2DX_LIB=$lib python bench.py --horizon 2000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-secondary 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('VALUE %.4g steps/s kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done
