set -x
mkdir -p gpurun_out
python bench.py --horizon 1000 --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; tail -3 gpurun_out/bench_small.err
cat gpurun_out/bench_small.json
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -3 gpurun_out/bench_full.err
cat gpurun_out/bench_full.json
CMD="python bench.py --scenarios 227328 --horizon 200 --steps 1 --warmup 1 --chunks 2 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__inst_executed_pipe_fp64.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --import-source on -k regex:rollout_dfff -s 2 -c 1 -o gpurun_out/prof_rollout_r1 $CMD > gpurun_out/ncu2.log 2>&1
tail -5 gpurun_out/ncu2.log
ls -la gpurun_out
