#!/bin/bash
# full capture of one collocation launch: $1 = number of colloc_kernel launches to skip (4: lean C3 batch, 30: C4 all-pairs batch)
mkdir -p gpurun_out
SKIP=${1:-4}; TAG=${2:-colloc}
CMD="python bench.py --scenarios 227328 --horizon 100 --steps 1 --warmup 1 --chunks 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:colloc_kernel -s $SKIP -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
