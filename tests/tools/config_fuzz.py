"""Parity fuzz over solver configurations around the reference's operating envelope (GPU vs oracle):

    python tests/tools/config_fuzz.py [n_configs=20] [n_per_condition=100] [out=gpurun_out/config_fuzz.json]

dt, c, alpha, Kp1, K_hill, n_hill, eta, eta_phi, phi_init and the horizon are drawn at random; every configuration is
evaluated for 4 x n prior draws on the device (K1 and, for half of the configurations, K1L) and by the oracle."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as orc  # noqa: E402  (checker)
from tmcmc202601_b200 import workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine  # noqa: E402


def main():
    n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    out = Path(sys.argv[3]) if len(sys.argv) > 3 else ROOT / "gpurun_out" / "config_fuzz.json"
    rng = np.random.default_rng(2026)
    res = []
    for k in range(n_cfg):
        T = int(rng.choice([300, 500, 800, 1200]))
        solver = dict(
            dt=float(10 ** rng.uniform(-5.0, -3.7)), maxtimestep=T, c_const=float(rng.choice([25.0, 50.0, 100.0])),
            alpha_const=float(rng.choice([0.0, 0.0, 3.0, 100.0])), Kp1=float(10 ** rng.uniform(-4.5, -3.5)),
            K_hill=float(rng.choice([0.0, 0.02, 0.05, 0.1])), n_hill=float(rng.choice([1.0, 2.0, 2.5, 3.0, 4.0])),
            eta_vec=list(rng.uniform(0.7, 1.4, 5)), eta_phi_vec=list(rng.uniform(0.7, 1.4, 5)),
        )
        wl = workloads.four_condition_workload(n, seed=1000 + k, solver=solver)
        rows = np.unique(np.linspace(T // 8, T, 5, dtype=int))
        for c in range(4):
            for o in range(5):
                wl.config.cond[c].idx_sparse[o] = int(rows[min(o, len(rows) - 1)])
            if k % 3 == 0:  # scalar small phi_init instead of the experimental composition
                for i in range(5):
                    wl.config.cond[c].phi_init[i] = 0.02
        eng = HamiltonEngine(wl.config, 0)
        if k % 2 == 1:
            eng.set_latency_kernel(10 ** 9)
        logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        eng.close()
        ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
        ok = ref > -1e19
        rel = np.abs(logL[ok] - ref[ok]) / np.maximum(np.abs(ref[ok]), 1e-300)
        r = {"config": {kk: (vv if not isinstance(vv, list) else [round(x, 3) for x in vv]) for kk, vv in solver.items()},
             "kernel": "K1L" if k % 2 == 1 else "K1", "evaluations": int(logL.size), "rel_max": float(rel.max()) if rel.size else 0.0,
             "n_above_1e-9": int((rel > 1e-9).sum()), "sentinels_gpu": int((logL <= -1e19).sum()), "sentinels_oracle": int((~ok).sum()),
             "sentinels_equal": bool(np.array_equal(logL <= -1e19, ~ok)), "nonfinite_flags_equal": bool(np.array_equal(st & 1, rst & 1)),
             "iteration_counts_identical_fraction": float(np.mean(it.astype(np.int64) == rit))}
        res.append(r)
        print(json.dumps(r), flush=True)
    out.parent.mkdir(parents=True, exist_ok=True)
    out.write_text(json.dumps(res, indent=1))
    worst = max(r["rel_max"] for r in res)
    print(f"worst rel_max over {n_cfg} configurations: {worst:.3e}; above 1e-9: {sum(r['n_above_1e-9'] for r in res)}")


if __name__ == "__main__":
    main()
