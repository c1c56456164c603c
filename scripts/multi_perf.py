"""Throughput of the host-facing multi-GPU handle (hmx_multi_*) and of the gradient / full-TSM entry points (development aid).

    python scripts/multi_perf.py [n_per_condition=65536]
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from tmcmc202601_b200 import workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine, MultiEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ndev = torch.cuda.device_count()
wl = workloads.four_condition_workload(n, seed=5)
th = torch.from_numpy(wl.theta).pin_memory().numpy()
out = {"devices": ndev, "evaluations": int(th.shape[0])}
for devs in ([0], list(range(ndev))):
    m = MultiEngine(wl.config, devs)
    m.loglik_batch(th[:4096], wl.cond_id[:4096])
    t0 = time.perf_counter()
    ll = m.loglik_batch(th, wl.cond_id)
    dt = time.perf_counter() - t0
    out[f"multi_{len(devs)}gpu_evals_per_s"] = th.shape[0] / dt
    m.close()
eng = HamiltonEngine(wl.config, 0)
sub = slice(0, min(th.shape[0], 32768))
eng.loglik_grad_batch(th[:1024], wl.cond_id[:1024])
t0 = time.perf_counter()
llg, g, st = eng.loglik_grad_batch(th[sub], wl.cond_id[sub])
out["value_and_grad_per_s"] = llg.size / (time.perf_counter() - t0)
t0 = time.perf_counter()
x0, s2, st2 = eng.solve_tsm_batch(th[:8192], wl.cond_id[:8192], want_x0=False)
out["full_tsm_per_s"] = 8192 / (time.perf_counter() - t0)
t0 = time.perf_counter()
eng.loglik_batch(th[sub], wl.cond_id[sub])
out["loglik_only_per_s_same_batch"] = llg.size / (time.perf_counter() - t0)
eng.close()
print(json.dumps(out))
