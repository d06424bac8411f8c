#!/bin/bash
# round 2, first GPU check: new collocation kernels (pairs, peer emulation), then a short core-only bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_colloc.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1; echo "colloc tests rc=$?"; tail -25 gpurun_out/r2_t1.log
timeout 600 python bench.py --steps 2 --warmup 3 --core-only --no-cpu > gpurun_out/r2_b1.json 2> gpurun_out/r2_b1.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_b1.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_b1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'roofline', d['roofline']['frac'])
for k, v in (d.get('secondary') or {}).items():
    print(k, json.dumps(v)[:400])
print(d['checks'])
PY
