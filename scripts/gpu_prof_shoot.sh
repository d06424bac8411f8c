#!/bin/bash
# ncu capture of the shooting kernels (run after the tests passed without ncu).  Usage: scripts/gpu_prof_shoot.sh TAG
TAG=${1:-r1}
mkdir -p gpurun_out
cat > /tmp/shoot_prof.py <<'PY'
import sys, numpy as np
sys.path.insert(0, "drone-sim-python_b200")
import torch
from d2d_b200.collocation import CollocationProblem, CostSpec
from d2d_b200.shooting import ShootingNLP
from d2d_b200.engine import get_engine
eng = get_engine()
Ps, Ns = 16384, 1001
prob = CollocationProblem(1, Ns, 0.02, cost=CostSpec(vsp=12., kvel=1.), multi=False)
nlp = ShootingNLP(prob, np.zeros((3, 1)), np.array([0., 30., np.pi]).reshape(3, 1), (-0.52, 0.52), (9., 14.), P=Ps)
th = eng.to_device(np.random.default_rng(0).uniform(-1., 1., (Ps, nlp.n)))
for _ in range(3):
    nlp.launch(th)
torch.cuda.synchronize()
PY
python /tmp/shoot_prof.py && \
ncu --set full --clock-control none --import-source on -k regex:shoot_ -c 6 -o gpurun_out/${TAG}_shoot python /tmp/shoot_prof.py > gpurun_out/${TAG}_shoot_ncu.log 2>&1
ncu -i gpurun_out/${TAG}_shoot.ncu-rep --page raw --csv > gpurun_out/${TAG}_shoot_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5
