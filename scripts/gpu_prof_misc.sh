#!/bin/bash
# full captures of the tracker and pure-pursuit rollout kernels from the bench command (after it ran once without ncu)
mkdir -p gpurun_out
TAG=${1:-r1}
CMD="python bench.py --scenarios 227328 --horizon 100 --steps 1 --warmup 1 --chunks 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_misc.log 2>&1 || exit 1
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
ncu --set full --metrics $M --clock-control none --import-source on -k regex:rollout_tracker -s 2 -c 1 -o gpurun_out/prof_tracker_$TAG $CMD > gpurun_out/ncu_misc1.log 2>&1; tail -1 gpurun_out/ncu_misc1.log | cut -c1-160
ncu --set full --metrics $M --clock-control none --import-source on -k regex:rollout_pursuit -s 2 -c 1 -o gpurun_out/prof_pursuit_$TAG $CMD > gpurun_out/ncu_misc2.log 2>&1; tail -1 gpurun_out/ncu_misc2.log | cut -c1-160
ls -la gpurun_out | grep "prof_.*_$TAG"
