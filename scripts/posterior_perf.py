"""Throughput of the ODE kernel on posterior-like particles (development aid): the stored reference posterior of the
production run, jittered, against prior draws of the same condition."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tmcmc202601_b200 import _abi, workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
    cd = workloads.load_conditions()["Dysbiotic_HOBIC_legacy_1k30"]
    cfg = _abi.make_config([workloads.cond_struct(cd, phi_init=0.02, sigma_obs=cd["sigma_scalar"])], **workloads.PRODUCTION_SOLVER)
    post = np.load(ROOT / "tests" / "golden" / "posterior_dh_1k30_samples.npz")["samples"]
    rng = np.random.default_rng(0)
    b = np.array(cd["bounds"])
    sets = {
        "posterior": np.clip(post[rng.integers(0, len(post), n)] + rng.normal(size=(n, 20)) * 0.02, b[:, 0], b[:, 1]),
        "prior": rng.uniform(b[:, 0], b[:, 1], size=(n, 20)),
    }
    eng = HamiltonEngine(cfg, 0)
    eng.set_timing(True)
    for name, th in sets.items():
        for _ in range(2):
            t0 = time.perf_counter()
            ll, st, it = eng.loglik_batch(th, None, return_status=True)
            dt = time.perf_counter() - t0
            c = eng.counters()
        print(json.dumps({"set": name, "evals": n, "wall_s": round(dt, 3), "kernel_ms": round(c["last_ode_kernel_ms"], 1),
                          "kernel_evals_per_s": round(n / c["last_ode_kernel_ms"] * 1e3), "trips_per_eval": round(c["q_evals"] / n, 1),
                          "iters_per_eval": round(c["newton_iters"] / n, 1), "iters_min": int(it.min()), "iters_max": int(it.max()),
                          "logL_median": float(np.median(ll))}))
    eng.close()


if __name__ == "__main__":
    main()
