"""K1L (eight lanes per solve) against K1 (one lane per solve): results and launch latency (development aid)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine


def run(wl, coop, reps=3):
    os.environ["HMX_COOP_MAX_EVALS"] = "100000000" if coop else "0"
    eng = HamiltonEngine(wl.config, 0)
    eng.set_timing(True)
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        out = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True, return_obs=True)
        best = min(best, time.perf_counter() - t0)
    c = eng.counters()
    eng.close()
    return out, best, c


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [250, 1250, 2500]
    for n in sizes:
        wl = workloads.four_condition_workload(n, seed=99)
        a, ta, ca = run(wl, False)
        b, tb, cb = run(wl, True)
        rel = np.abs(a[0] - b[0]) / np.maximum(np.abs(a[0]), 1e-300)
        print(json.dumps({"evals": int(a[0].size), "lane_ms": round(ta * 1e3, 2), "coop_ms": round(tb * 1e3, 2), "speedup": round(ta / tb, 2),
                          "logL_rel_max": float(rel.max()), "status_equal": bool(np.array_equal(a[1], b[1])),
                          "iters_identical_fraction": float(np.mean(a[2] == b[2])), "iters_total": [int(ca["newton_iters"]), int(cb["newton_iters"])],
                          "trips_total": [int(ca["q_evals"]), int(cb["q_evals"])],
                          "obs_state_rel_max": float(np.max(np.abs(a[3] - b[3]) / np.maximum(np.abs(a[3]), 1e-300)))}), flush=True)


if __name__ == "__main__":
    main()
