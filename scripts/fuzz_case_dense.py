import sys, os
import numpy as np
sys.path.insert(0, '.')
from oracle import oracle as orc
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine
solver = dict(dt=0.00011848512631842148, maxtimestep=500, c_const=50.0, alpha_const=0.0, Kp1=6.89941793513431e-05, K_hill=0.0, n_hill=2.0)
rng = np.random.default_rng(2026)
# reproduce the configuration's eta vectors exactly as config_fuzz.py draws them is not needed: use the printed (rounded) ones
solver["eta_vec"] = [0.858, 1.068, 1.002, 1.164, 0.709]
solver["eta_phi_vec"] = [1.013, 0.956, 0.837, 1.116, 1.005]
for seed in range(1000, 1030):
    wl = workloads.four_condition_workload(100, seed=seed, solver=solver)
    T = 500
    rows = np.unique(np.linspace(T // 8, T, 5, dtype=int))
    for c in range(4):
        for o in range(5):
            wl.config.cond[c].idx_sparse[o] = int(rows[min(o, len(rows) - 1)])
    ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
    out = {}
    for name, env in (("K1", "0"), ("K1r", "1")):
        os.environ["HMX_FORCE_REFERENCE_LOOP"] = env
        eng = HamiltonEngine(wl.config, 0)
        out[name] = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        eng.close()
    r1 = np.abs(out["K1"][0] - ref) / np.abs(ref)
    r2 = np.abs(out["K1r"][0] - ref) / np.abs(ref)
    j = int(np.argmax(r1))
    print(seed, "K1 max rel %.2e (status %d, oracle %d)  K1r at same eval %.2e  K1r max %.2e" % (r1.max(), out["K1"][1][j], rst[j], r2[j], r2.max()))
