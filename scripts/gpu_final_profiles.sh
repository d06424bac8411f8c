#!/bin/bash
# final evidence of the round: launch list of the default bench command, full captures of the three hot kernels
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log | cut -c1-200
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_final.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:rollout_dfff -s 12 -c 1 -o gpurun_out/prof_rollout_final $CMD > gpurun_out/ncu_final1.log 2>&1
tail -1 gpurun_out/ncu_final1.log | cut -c1-200
$CMD > gpurun_out/plain_final2.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:colloc_kernel -s 4 -c 1 -o gpurun_out/prof_colloc_final $CMD > gpurun_out/ncu_final2.log 2>&1
tail -1 gpurun_out/ncu_final2.log | cut -c1-200
$CMD > gpurun_out/plain_final3.log 2>&1 && ncu --set full --metrics $M --clock-control none --import-source on -k regex:rollout_formation -s 2 -c 1 -o gpurun_out/prof_formation_final $CMD > gpurun_out/ncu_final3.log 2>&1
tail -1 gpurun_out/ncu_final3.log | cut -c1-200
ls -la gpurun_out | grep final
