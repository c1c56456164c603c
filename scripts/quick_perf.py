"""Ad-hoc GPU timing of the hot path (development aid; bench.py is the contract)."""
import sys, time, json
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine

def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    solver = dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=4.0) if "--t500" in sys.argv else None
    sizes = [int(x) for x in args] or [250, 1250, 9472]
    wl0 = workloads.four_condition_workload(8, solver=solver)
    if solver is not None:  # observation rows rescaled onto the 500-step grid
        for k in range(4):
            for o in range(5):
                wl0.config.cond[k].idx_sparse[o] = [68, 136, 226, 339, 475][o]
    eng = HamiltonEngine(wl0.config, 0)
    print("fp64 peak TFLOP/s:", eng.measure_fp64_peak())
    eng.set_timing(True)
    for n in sizes:
        wl = workloads.four_condition_workload(n, solver=solver)
        for rep in range(3):
            t0 = time.perf_counter()
            logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
            t1 = time.perf_counter()
            c = eng.counters()
        M = logL.size
        print(json.dumps(dict(evals=M, wall_s=round(t1 - t0, 4), ode_ms=round(c["last_ode_kernel_ms"], 2),
                              evals_per_s=round(M / (t1 - t0), 1), kernel_evals_per_s=round(M / (c["last_ode_kernel_ms"] * 1e-3), 1),
                              iters=c["newton_iters"], trips=c["q_evals"], trips_per_eval=round(c["q_evals"] / M, 1),
                              nonfinite=int((st & 1).sum()), maxiter=int(((st & 2) > 0).sum()), singular=int(((st & 4) > 0).sum()))))
    eng.close()

if __name__ == "__main__":
    main()
