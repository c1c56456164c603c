"""K1L, the eight-lanes-per-solve kernel of the latency regime (csrc/hmx_coop.cuh), against K1 (one lane per solve) and
the oracle, for every template instance (alpha x Hill variants), with trajectories and observation states."""
import numpy as np
import pytest

from oracle import oracle as orc
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine

pytestmark = pytest.mark.gpu

VARIANTS = {
    "production": {}, "hill_n1": {"n_hill": 1.0}, "hill_n2": {"n_hill": 2.0}, "hill_n2.5_pow": {"n_hill": 2.5}, "hill_off": {"K_hill": 0.0},
    "alpha3": {"alpha_const": 3.0}, "alpha3_hill_off": {"alpha_const": 3.0, "K_hill": 0.0}, "alpha3_pow": {"alpha_const": 3.0, "n_hill": 2.5},
}


def _both(wl, monkeypatch, fn):
    out = []
    for limit in ("0", "100000000"):  # K1 only / K1L for every batch
        monkeypatch.setenv("HMX_COOP_MAX_EVALS", limit)
        eng = HamiltonEngine(wl.config, 0)
        try:
            out.append(fn(eng))
        finally:
            eng.close()
    monkeypatch.delenv("HMX_COOP_MAX_EVALS")
    return out


@pytest.mark.parametrize("name", list(VARIANTS))
def test_coop_kernel_equals_lane_kernel_and_oracle(built, monkeypatch, name):
    solver = dict(workloads.PRODUCTION_SOLVER, maxtimestep=600, **VARIANTS[name])
    wl = workloads.four_condition_workload(45, seed=5, solver=solver, interleave=True)
    for k in range(4):
        for o, v in enumerate([60, 150, 290, 440, 599]):
            wl.config.cond[k].idx_sparse[o] = v
    lane, coop = _both(wl, monkeypatch, lambda e: e.loglik_batch(wl.theta, wl.cond_id, return_status=True, return_obs=True))
    rel = np.abs(lane[0] - coop[0]) / np.maximum(np.abs(lane[0]), 1e-300)
    assert rel.max() < 1e-10, rel.max()
    assert np.array_equal(lane[1], coop[1])
    assert np.mean(lane[2] == coop[2]) > 0.97
    assert np.allclose(lane[3], coop[3], rtol=1e-10, atol=1e-300)
    ref = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id)
    assert np.max(np.abs(coop[0] - ref) / np.abs(ref)) < 1e-9


def test_coop_kernel_trajectories_and_work_queue(built, monkeypatch):
    """Whole trajectories and row subsets from K1L; more evaluations than resident groups exercises the work counter."""
    solver = dict(workloads.PRODUCTION_SOLVER, maxtimestep=40)
    wl = workloads.four_condition_workload(2200, seed=8, solver=solver)  # 8800 evaluations > 7104 resident groups
    for k in range(4):
        for o, v in enumerate([5, 12, 20, 30, 39]):
            wl.config.cond[k].idx_sparse[o] = v
    lane, coop = _both(wl, monkeypatch, lambda e: (e.loglik_batch(wl.theta, wl.cond_id, return_status=True),
                                                   e.trajectory_batch(wl.theta[:300], wl.cond_id[:300])[0],
                                                   e.trajectory_rows(wl.theta[:300], np.array([0, 1, 17, 40]), wl.cond_id[:300])[0]))
    assert np.allclose(lane[0][0], coop[0][0], rtol=1e-10) and np.array_equal(lane[0][1], coop[0][1])
    assert np.allclose(lane[1], coop[1], rtol=1e-10, atol=1e-300)
    assert np.array_equal(coop[2], coop[1][:, [0, 1, 17, 40]])
    assert np.array_equal(coop[1][:, 0, :5], np.tile(np.array([c.phi_init[:] for c in wl.config.cond[:4]])[wl.cond_id[:300]], 1))


def test_latency_kernel_switch(built):
    """hmx_set_latency_kernel(-1): small batches take K1L and large ones K1; off by default (bitwise batch independence)."""
    wl = workloads.four_condition_workload(2000, seed=3, solver=dict(workloads.PRODUCTION_SOLVER, maxtimestep=30))
    for k in range(4):
        for o, v in enumerate([3, 8, 14, 22, 29]):
            wl.config.cond[k].idx_sparse[o] = v
    eng = HamiltonEngine(wl.config, 0)
    big = eng.loglik_batch(wl.theta, wl.cond_id)            # 8000 evaluations: K1
    small_default = eng.loglik_batch(wl.theta[:500], wl.cond_id[:500])
    assert np.array_equal(big[:500], small_default)         # default: one kernel, values independent of the batch
    eng.set_latency_kernel(-1)
    small = eng.loglik_batch(wl.theta[:500], wl.cond_id[:500])  # K1L
    assert np.array_equal(eng.loglik_batch(wl.theta, wl.cond_id), big)  # above the limit: still K1
    eng.close()
    assert np.allclose(big[:500], small, rtol=1e-10)
