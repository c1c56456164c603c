"""Time of everything in hmx_loglik_batch that is not the ODE kernel (observation kernel K1c, scheduling kernels), and the cost
of the all-rows sigma2 guard (development aid):  [HMX_LIB=...] python scripts/obs_perf.py"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tmcmc202601_b200 import workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
wl = workloads.four_condition_workload(n, seed=20260101)
eng = HamiltonEngine(wl.config, device=0)
eng.use_torch_stream()
eng.set_timing(True)
th = torch.from_numpy(wl.theta).cuda()
cid = torch.from_numpy(wl.cond_id).cuda()
ll = torch.empty(th.shape[0], dtype=torch.float64, device="cuda")


def timed(k=3, m=None):
    tot, ode = [], []
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.loglik_batch_device(th[:m], cid[:m], ll[:m])
        e1.record()
        torch.cuda.synchronize()
        tot.append(e0.elapsed_time(e1))
        ode.append(eng.counters()["last_ode_kernel_ms"])
    return min(tot), min(ode)


timed(1)
t, o = timed()
out = {"evals": th.shape[0], "call_ms": t, "ode_ms": o, "rest_ms": t - o}
m = 65536
t2, o2 = timed(2, m)
eng.set_sig2_guard(True)
timed(1, m)
t3, o3 = timed(2, m)
out.update({"guard_evals": m, "guard_off_ms": t2, "guard_on_ms": t3, "guard_factor": t3 / t2})
print(json.dumps(out))
