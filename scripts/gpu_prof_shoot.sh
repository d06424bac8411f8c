#!/bin/bash
# ncu capture of the planner-solve kernels (run after the tests passed without ncu).  Usage: scripts/gpu_prof_shoot.sh TAG
TAG=${1:-r1}
mkdir -p gpurun_out build
cat > build/shoot_prof.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200.collocation import CollocationProblem, CostSpec
from d2d_b200 import shooting
from d2d_b200.engine import get_engine
eng = get_engine()
# (a) evaluation throughput: 16384 problems on the C3 grid
Ps, Ns = 16384, 1001
prob = CollocationProblem(1, Ns, 0.02, cost=CostSpec(vsp=12., kvel=1.), multi=False)
nlp = shooting.ShootingNLP(prob, np.zeros((3, 1)), np.array([0., 30., np.pi]).reshape(3, 1), (-0.52, 0.52), (9., 14.), P=Ps)
th = eng.to_device(np.random.default_rng(0).uniform(-1., 1., (Ps, nlp.n)))
for _ in range(3):
    nlp.launch(th)
torch.cuda.synchronize()
del nlp, th
# (b) a population solve without graph capture: 2048 problems, N = 101, 60 ticks
P, N = 2048, 101
prob = CollocationProblem(1, N, 0.1, cost=CostSpec(vsp=12., kvel=1.), multi=False)
rng = np.random.default_rng(1)
p1 = np.stack([rng.uniform(-10, 10, P), rng.uniform(28, 40, P), np.pi + rng.uniform(-0.5, 0.5, P)], 1).reshape(P, 3, 1)
nlp = shooting.ShootingNLP(prob, np.zeros((3, 1)), p1, (-0.52, 0.52), (9., 14.), P=P)
shooting.solve(nlp, nlp.theta_of(np.full((1, N), 0.1), np.full((1, N), 12.)), use_graph=False, ticks_per_check=20, max_ticks=60)
torch.cuda.synchronize()
PY
python build/shoot_prof.py && \
ncu --set full --clock-control none --import-source on -k regex:'shoot_|al_lbfgs' --launch-skip 3 -c 3 -o gpurun_out/${TAG}_shoot python build/shoot_prof.py > gpurun_out/${TAG}_shoot_ncu.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'shoot_|al_lbfgs' --launch-skip 120 -c 6 -o gpurun_out/${TAG}_solve python build/shoot_prof.py > gpurun_out/${TAG}_solve_ncu.log 2>&1
for f in shoot solve; do ncu -i gpurun_out/${TAG}_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null; done
ls -la gpurun_out | tail -8
