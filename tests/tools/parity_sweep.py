"""Parity distribution of the GPU log-likelihoods against the CPU oracle over many prior draws
("check 1e-9 parity on 10^4 theta from each condition's prior").

    python tests/tools/parity_sweep.py [n_per_condition=10000] [out=gpurun_out/parity_sweep.json] [--variants]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402  (checker)
from tmcmc202601_b200 import workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine  # noqa: E402


VARIANTS = {
    # name: solver overrides on top of workloads.PRODUCTION_SOLVER (every template instance of the ODE kernel is covered)
    "production": {},
    "hill_n1": {"n_hill": 1.0},
    "hill_n2": {"n_hill": 2.0},
    "hill_n3": {"n_hill": 3.0},
    "hill_n2.5_pow": {"n_hill": 2.5},
    "hill_off": {"K_hill": 0.0},
    "alpha3": {"alpha_const": 3.0},
    "alpha3_hill_off": {"alpha_const": 3.0, "K_hill": 0.0},
    "t500_c100_a100": None,  # the surrogate configuration (dt=1e-5, 500 steps, c = alpha = 100, phi_init 0.2)
}


def sweep(n, solver, t500=False):
    wl = workloads.four_condition_workload(n, seed=777, solver=solver)
    if t500:
        for k in range(4):
            for o in range(5):
                wl.config.cond[k].idx_sparse[o] = [68, 136, 226, 339, 475][o]
    eng = HamiltonEngine(wl.config, 0)
    t0 = time.perf_counter()
    logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
    t_gpu = time.perf_counter() - t0
    eng.close()
    t0 = time.perf_counter()
    ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
    t_cpu = time.perf_counter() - t0
    rel = np.abs(logL - ref) / np.maximum(np.abs(ref), 1e-300)
    return {"n_per_condition": n, "evaluations": int(logL.size), "gpu_s": t_gpu, "oracle_s": t_cpu,
            "rel_max": float(rel.max()), "rel_p999": float(np.quantile(rel, 0.999)), "rel_p99": float(np.quantile(rel, 0.99)),
            "rel_median": float(np.median(rel)), "n_above_1e-12": int((rel > 1e-12).sum()), "n_above_1e-9": int((rel > 1e-9).sum()),
            "iteration_counts_identical_fraction": float(np.mean(it.astype(np.int64) == rit)),
            "iteration_total_gpu": int(it.astype(np.int64).sum()), "iteration_total_oracle": int(rit.sum()),
            "nonfinite_flags_equal": bool(np.array_equal(st & 1, rst & 1)), "n_sentinel": int((logL == -1e20).sum()),
            "per_condition_rel_max": [float(rel[wl.cond_id == c].max()) for c in range(4)]}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else 10000
    out = Path(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ROOT / "gpurun_out" / "parity_sweep.json"
    if "--variants" in sys.argv:
        res = {}
        for name, over in VARIANTS.items():
            solver = dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=4.0) if over is None else dict(workloads.PRODUCTION_SOLVER, **over)
            res[name] = sweep(n, solver, t500=over is None)
            print(name, json.dumps({k: res[name][k] for k in ("evaluations", "rel_max", "rel_p999", "n_above_1e-9",
                                                               "iteration_counts_identical_fraction", "nonfinite_flags_equal")}), flush=True)
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps(res, indent=1))
        return
    res = sweep(n, None)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(res, indent=1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
