#!/bin/bash
# round 2 multi-GPU check ($1 = number of GPUs): the >= 2-GPU parity test, then the bench at N with the core metrics
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi_n$N.log 2>&1; echo "multi test rc=$?"; tail -15 gpurun_out/r2_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'ms', d['ms_per_step'], d['e2e'] and d['e2e']['ms_per_step'])
for k, v in (d.get('secondary') or {}).items():
    print(k, json.dumps(v)[:330])
print(d['checks'])
PY
