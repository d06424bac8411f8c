#!/bin/bash
# usage: scripts/gpu_check.sh TAG [full]   -- GPU tests, a short bench, the launch list and one ncu --set full capture
TAG=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --horizon 2000 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -3 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${TAG}.json"))
print("VALUE %.4g steps/s  e2e %.4g  kernel_ms %.3f  roofline frac %.3f (peak %.2f)  clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["roofline"]["peak"], d["clocks"]))
PY
if [ "$2" != "noncu" ]; then
CMD="python bench.py --scenarios 227328 --horizon 200 --steps 1 --warmup 1 --chunks 2 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 && ncu --set full --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --import-source on -k regex:rollout_dfff -s 2 -c 1 -o gpurun_out/prof_rollout_${TAG} $CMD > gpurun_out/ncu_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_${TAG}.log
fi
