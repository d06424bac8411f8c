#!/bin/bash
# ncu --set full capture of one kernel: $1 = prof_target name, $2 = kernel regex, $3 = launches to skip, $4 = tag
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
T=$1; K=$2; S=${3:-3}; TAG=${4:-$1}
python scripts/prof_target.py $T > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
cat gpurun_out/plain_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_$TAG python scripts/prof_target.py $T > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
