"""Small run through every round-2 kernel for compute-sanitizer (development aid):
    compute-sanitizer --tool memcheck python scripts/sanitize_target.py
Short trajectories (60 steps) so that the instrumented run stays within a minute."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tmcmc202601_b200 import _abi, workloads  # noqa: E402
from tmcmc202601_b200.engine import HamiltonEngine, MultiEngine  # noqa: E402

solver = dict(workloads.PRODUCTION_SOLVER, maxtimestep=60)
wl = workloads.four_condition_workload(40, seed=3, solver=solver, interleave=True)
for k in range(4):
    for o, r in enumerate((5, 12, 30, 44, 60)):
        wl.config.cond[k].idx_sparse[o] = r
eng = HamiltonEngine(wl.config, 0)
ll, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
x0, s2, x1, ll2, st2 = eng.tsm_obs_rows(wl.theta, wl.cond_id)
llg, grad, ds, stg = eng.loglik_grad_batch(wl.theta, wl.cond_id, want_dstate=True)
xf, s2f, stf = eng.solve_tsm_batch(wl.theta[:16], wl.cond_id[:16])
eng.set_sig2_guard(True)
ll3 = eng.loglik_batch(wl.theta, wl.cond_id)
eng.set_sig2_guard(False)
g, stt, itt = eng.trajectory_rows(wl.theta[:8], np.array([0, 1, 59, 60], dtype=np.int32), wl.cond_id[:8])
gl, stl, itl = eng.legacy_trajectory_rows(_abi.make_legacy_config(n_species=4, maxtimestep=40, active_species=[0, 2, 3]), np.random.default_rng(0).uniform(0, 3, (50, 14)))
gl2, _, _ = eng.legacy_trajectory_rows(_abi.make_legacy_config(n_species=4, maxtimestep=40), np.random.default_rng(1).uniform(0, 3, (7, 14)), np.array([0, 7, 40], dtype=np.int32))
w = np.random.default_rng(2).random(70001)
w /= w.sum()
cs = eng.cumsum(w)
idx = eng.resample(w, 0.37)
ref = np.cumsum(w)
ref[-1] = 1.0
assert np.array_equal(cs, ref)
mean, cov = eng.weighted_cov(wl.theta, np.full(wl.theta.shape[0], 1.0 / wl.theta.shape[0]))
pop = eng.device_array(wl.theta.shape, wl.theta)
assert np.array_equal(eng.loglik_batch(pop, wl.cond_id), ll)
pop.free()
m = MultiEngine(wl.config, [0, 0])
assert np.array_equal(m.loglik_batch(wl.theta, wl.cond_id), ll)
m.close()
os.environ["HMX_FORCE_REFERENCE_LOOP"] = "1"
e2 = HamiltonEngine(wl.config, 0)
llr = e2.loglik_batch(wl.theta, wl.cond_id)
e2.close()
assert np.array_equal(ll, ll2) and np.array_equal(ll, llg) and np.array_equal(ll, ll3) and np.allclose(ll, llr, rtol=1e-9)
eng.close()
print("sanitize target ok:", ll[:3], grad[0, :3])
