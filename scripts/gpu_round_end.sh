#!/bin/bash
# round-end evidence at HEAD: GPU tests, smoke, the default bench line, the reference arm, the launch list of the bench command
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_end.json 2> gpurun_out/bench_end.err; tail -2 gpurun_out/bench_end.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_end_reference.json 2>> gpurun_out/bench_end.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_end.csv python bench.py --no-cpu > gpurun_out/ncu_launches_end.log 2>&1
tail -1 gpurun_out/ncu_launches_end.log | cut -c1-200
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_end.json"))
print("value %.4g e2e %.4g frac %.3f cpu %.3g launches %s clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"]))
r = json.load(open("gpurun_out/bench_end_reference.json"))
print("reference arm:", r["value"], r["unit"], r["cpu_baseline"])
PY
