"""One K1L launch of 4000 evaluations (profiling target)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine

wl = workloads.four_condition_workload(1000, seed=99)
eng = HamiltonEngine(wl.config, 0)
eng.set_latency_kernel(-1)
for _ in range(2):
    eng.loglik_batch(wl.theta, wl.cond_id)
eng.close()
