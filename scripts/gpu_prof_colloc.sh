#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --scenarios 227328 --horizon 100 --steps 1 --warmup 1 --chunks 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_colloc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:colloc_kernel -s 4 -c 1 -o gpurun_out/prof_colloc_r1b $CMD > gpurun_out/ncu_colloc.log 2>&1
tail -2 gpurun_out/ncu_colloc.log
