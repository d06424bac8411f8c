#!/bin/bash
export D2DX_ROLLOUT_WS=1
python -m pytest tests/test_gpu_rollout.py tests/test_gpu_edges.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -4
for ws in 0 1; do
D2DX_ROLLOUT_WS=$ws python bench.py --horizon 2000 --steps 3 --warmup 2 --no-cpu --no-e2e --no-secondary 2>&1 | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('WS=$ws VALUE %.4g steps/s kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))"
done
