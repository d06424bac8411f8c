#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --horizon 2000 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_sec.json 2> gpurun_out/bench_sec.err; tail -3 gpurun_out/bench_sec.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sec.json"))
print("VALUE %.4g" % d["value"]); print(json.dumps(d["secondary"], indent=1))
PY
CMD="python bench.py --scenarios 227328 --horizon 100 --steps 1 --warmup 1 --chunks 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_colloc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:colloc_kernel -s 4 -c 1 -o gpurun_out/prof_colloc_r1 $CMD > gpurun_out/ncu_colloc.log 2>&1
tail -2 gpurun_out/ncu_colloc.log
$CMD > gpurun_out/plain_form.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_formation -s 2 -c 1 -o gpurun_out/prof_formation_r1 $CMD > gpurun_out/ncu_form.log 2>&1
tail -2 gpurun_out/ncu_form.log
