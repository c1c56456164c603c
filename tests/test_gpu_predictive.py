"""Batched trajectory consumers on the B200: hmx_trajectory_rows against hmx_trajectory_batch, the
surrogate training set against vectors of the live reference generator, posterior-predictive bands."""
import time

import numpy as np
import pytest

from helpers import GOLDEN
from tmcmc202601_b200 import predictive, workloads
from tmcmc202601_b200.solver import BiofilmNewtonSolver5S

pytestmark = pytest.mark.gpu


def test_trajectory_rows_equal_rows_of_the_full_trajectory(built):
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    b = np.array(cd["bounds"])
    th = np.random.default_rng(1).uniform(b[:, 0], b[:, 1], size=(37, 20))
    s = BiofilmNewtonSolver5S(**dict(workloads.PRODUCTION_SOLVER, maxtimestep=300, phi_init=cd["phi_init"]))
    full = s.run_deterministic_batch(th)
    for rows in (np.array([0, 1, 2, 299, 300]), np.arange(0, 301, 7), np.array([300]), np.array([0])):
        part, st = s.run_deterministic_batch(th, rows=rows, return_status=True)
        assert part.shape == (37, len(rows), 12) and np.array_equal(part, full[:, rows])
        assert (st & 1).sum() == 0
    with pytest.raises(Exception):
        s.run_deterministic_batch(th, rows=np.array([5, 5]))
    with pytest.raises(Exception):
        s.run_deterministic_batch(th, rows=np.array([0, 301]))


def test_training_dataset_equals_the_live_reference_generator(built):
    k = np.load(GOLDEN / "training_dataset_kat.npz")
    for name, mf, post in (("uniform", 0.0, None), ("map", 0.4, None), ("mix", 0.3, k["posterior"])):
        th, phi, t, b = predictive.generate_dataset(14, k["bounds_in"], seed=11, maxtimestep=60, dt=1e-5, n_time_out=13, theta_map=k["theta_map"],
                                                    map_frac=mf, map_std_frac=0.1, posterior_samples=post, posterior_frac=0.5)
        assert np.array_equal(th, k[f"{name}_theta"]) and np.array_equal(b, k[f"{name}_bounds"])
        assert np.allclose(phi, k[f"{name}_phi"], rtol=1e-9, atol=1e-14), name
    # the reference generator's default size: 10^4 solves at T = 500, one numba solve at a time there
    t0 = time.perf_counter()
    th, phi, t, _ = predictive.generate_dataset(10000, k["bounds_in"], seed=42, theta_map=k["theta_map"], map_frac=0.3)
    wall = time.perf_counter() - t0
    print(f"10^4-sample training set (T=500, 100 time points): {wall:.2f} s")
    assert th.shape == (10000, 20) and phi.shape == (10000, 100, 5) and t.shape == (100,) and np.isfinite(phi).all()
    assert wall < 60.0


def test_posterior_predictive_bands(built):
    s = np.load(GOLDEN / "posterior_dh_1k30_samples.npz")["samples"]
    cd = workloads.load_conditions()["Dysbiotic_HOBIC_legacy_1k30"]
    out = predictive.posterior_predictive_bands(s, dict(workloads.PRODUCTION_SOLVER, phi_init=0.02), n_draws=1500, n_time_out=126)
    q = out["phibar_q"]
    assert q.shape == (3, len(out["rows"]), 5) and out["n_used"] == 1500 and out["n_failed"] == 0
    assert np.all(q[0] <= q[1] + 1e-15) and np.all(q[1] <= q[2] + 1e-15)
    assert np.all(out["total_q"][2] <= 1.0) and out["t"][0] == 0.0 and abs(out["t"][-1] - 0.25) < 1e-12
    # the 90 % band of the calibrated posterior brackets most observations within the observation noise
    rows = out["rows"]
    near = [int(np.argmin(np.abs(rows - i))) for i in cd["idx_sparse"]]
    data = np.array(cd["data"])
    inside = (data >= q[0][near] - 2 * cd["sigma_scalar"]) & (data <= q[2][near] + 2 * cd["sigma_scalar"])
    assert inside.mean() > 0.9


@pytest.mark.gpu
def test_trajectory_rows_accepts_a_device_row_list(built):
    """hmx_trajectory_rows takes its row list from host or device memory (round-1 ABI rejected a device list at run time)."""
    import torch

    from tmcmc202601_b200 import _abi, workloads
    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(4, seed=2, solver=dict(workloads.PRODUCTION_SOLVER, maxtimestep=80))
    for k in range(4):
        for o, r in enumerate((5, 12, 30, 44, 60)):
            wl.config.cond[k].idx_sparse[o] = r
    eng = HamiltonEngine(wl.config, device=0)
    try:
        rows = np.array([0, 3, 40, 80], dtype=np.int32)
        g_host, _, _ = eng.trajectory_rows(wl.theta, rows, wl.cond_id)
        eng.use_torch_stream()
        rows_d = torch.from_numpy(rows).cuda()
        g = np.empty_like(g_host)
        rc = eng._lib.hmx_trajectory_rows(eng._ctx, wl.theta.ctypes.data, wl.cond_id.ctypes.data, wl.theta.shape[0], rows_d.data_ptr(), rows.size,
                                          g.ctypes.data, None, None)
        assert rc == 0 and np.array_equal(g, g_host)
    finally:
        eng.close()
