#!/bin/bash
export D2DX_ROLLOUT_WS=1
CMD="python bench.py --scenarios 227328 --horizon 200 --steps 1 --warmup 1 --chunks 2 --no-e2e --no-cpu --no-secondary"
$CMD > gpurun_out/plain_ws.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rollout_dfff -s 2 -c 1 -o gpurun_out/prof_rollout_ws $CMD > gpurun_out/ncu_ws.log 2>&1
tail -1 gpurun_out/ncu_ws.log
