#!/bin/bash
# usage (under gpurun): bash scripts/ab_variants.sh base cubic fold ...   -> ODE kernel ms for 4 x 262144 evaluations per variant
for v in "$@"; do
  if [ "$v" = base ]; then unset HMX_LIB; else export HMX_LIB=scripts/_variants/libhamilton_$v.so; fi
  ms=$(python scripts/quick_perf.py 262144 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ode_ms'], d['iters'], d['trips'])")
  par=$(python - <<PY
import sys; sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as orc
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine
wl = workloads.four_condition_workload(96, seed=20260101)
eng = HamiltonEngine(wl.config, device=0)
ll = eng.loglik_batch(wl.theta, wl.cond_id)
ref = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id)
print("%.2e" % (np.abs(ll - ref) / np.abs(ref)).max())
PY
)
  echo "variant $v: ode_ms iters trips = $ms ; max rel vs oracle (384 evals) = $par"
done
