"""Parity distribution of the GPU log-likelihoods against the CPU oracle over many prior draws
("check 1e-9 parity on 10^4 theta from each condition's prior").

    python tests/tools/parity_sweep.py [n_per_condition=10000] [out=gpurun_out/parity_sweep.json]
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    out = Path(sys.argv[2]) if len(sys.argv) > 2 else ROOT / "gpurun_out" / "parity_sweep.json"
    wl = workloads.four_condition_workload(n, seed=777)
    eng = HamiltonEngine(wl.config, 0)
    t0 = time.perf_counter()
    logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
    t_cpu = time.perf_counter() - t0
    rel = np.abs(logL - ref) / np.maximum(np.abs(ref), 1e-300)
    res = {"n_per_condition": n, "evaluations": int(logL.size), "gpu_s": t_gpu, "oracle_s": t_cpu,
           "rel_max": float(rel.max()), "rel_p999": float(np.quantile(rel, 0.999)), "rel_p99": float(np.quantile(rel, 0.99)),
           "rel_median": float(np.median(rel)), "n_above_1e-12": int((rel > 1e-12).sum()), "n_above_1e-9": int((rel > 1e-9).sum()),
           "iteration_counts_identical_fraction": float(np.mean(it.astype(np.int64) == rit)),
           "iteration_total_gpu": int(it.astype(np.int64).sum()), "iteration_total_oracle": int(rit.sum()),
           "nonfinite_flags_equal": bool(np.array_equal(st & 1, rst & 1)), "n_sentinel": int((logL == -1e20).sum()),
           "per_condition_rel_max": [float(rel[wl.cond_id == c].max()) for c in range(4)]}
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(res, indent=1))
    print(json.dumps(res))
    eng.close()


if __name__ == "__main__":
    main()
