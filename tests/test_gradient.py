"""d logL / d theta for the NUTS / HMC engine (data_5species/main/tmcmc_nuts_engine.py:202).

The reference has no analytic gradient of this path (its JAX model is a different solver), so the chain of evidence is:
  1. oracle/hamilton_oracle_grad.c (dense exact Jacobian, independent of the product's block-structured code)
     == central finite differences of the reference-pinned oracle likelihood, at a tight Newton tolerance  (<= 1e-6);
  2. the product's arithmetic (hmx_grad.cuh: sensitivity_step / direction_exact, compiled for the host) == that oracle (<= 1e-9);
  3. (-m gpu) hmx_loglik_grad_batch through the C ABI == that oracle (<= 1e-9), values == hmx_loglik_batch bit for bit."""
import ctypes as C

import numpy as np
import pytest

from helpers import load_emu
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, workloads

KW = dict(dt=1e-4, maxtimestep=2500, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0)


def _setup(key, seed, tsm, **over):
    cd = workloads.load_conditions()[key]
    b = np.array(cd["bounds"])
    th = np.random.default_rng(seed).uniform(b[:, 0], b[:, 1])
    kw = dict(KW, **over)
    cfg = orc.make_solver_cfg(phi_init=cd["phi_init"], **kw)
    cc = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], cd["sigma_obs"], cov_rel=0.005, active_theta_indices=cd["active_indices"],
                           tsm_variance=tsm)
    return cd, th, cfg, cc, kw


def test_exact_jacobian_is_the_derivative_of_the_residual():
    """The dense exact Jacobian of the checker against central differences of the (reference-pinned) oracle residual,
    and the statement that the reference's Newton matrix is NOT that derivative."""
    cd, th, cfg, cc, _ = _setup("Dysbiotic_HOBIC", 0, False)
    _, g, *_ = orc.run_deterministic(cfg, th)
    for t in (5, 700, 2400):  # (at t = 1 phi_0 ~ 1e-7 and its barrier ~ 1e14 swamps a finite difference of Q_5)
        Jx, Jp = orc.exact_jacobians(cfg, th, g[t], g[t - 1])
        fd = np.zeros((12, 12))
        fdp = np.zeros((12, 12))
        for j in range(12):
            h = 1e-6 * max(abs(g[t][j]), 1e-3)
            e = np.zeros(12)
            e[j] = h
            fd[:, j] = (orc.compute_Q(cfg, th, g[t] + e, g[t - 1]) - orc.compute_Q(cfg, th, g[t] - e, g[t - 1])) / (2 * h)
            fdp[:, j] = (orc.compute_Q(cfg, th, g[t], g[t - 1] + e) - orc.compute_Q(cfg, th, g[t], g[t - 1] - e)) / (2 * h)
        assert np.abs(Jx - fd).max() < 2e-5 * np.abs(Jx).max()
        dense_p = np.zeros((12, 12))
        for i in range(5):
            dense_p[i, i], dense_p[i, 6 + i] = Jp[i]
            dense_p[6 + i, i], dense_p[6 + i, 6 + i] = Jp[6 + i]
        dense_p[5, 5] = Jp[5, 0]
        assert np.abs(dense_p - fdp).max() < 2e-5 * np.abs(dense_p).max()
        K_ref = orc.compute_J(cfg, th, g[t], g[t - 1])
        assert np.abs(K_ref - Jx).max() > 1e-3 * np.abs(Jx).max()  # the reference's matrix is an approximation


@pytest.mark.parametrize("key,tsm", [("Dysbiotic_HOBIC", False), ("Dysbiotic_Static", True)])
def test_oracle_gradient_equals_central_differences_of_the_oracle_likelihood(key, tsm):
    cd, th, cfg, cc, _ = _setup(key, 1, tsm, eps=1e-13, max_newton_iter=200)
    ll, grad = orc.loglik_grad(cfg, cc, th)
    assert ll == orc.evaluate(cfg, cc, th)
    h = 1e-5 * np.maximum(np.abs(th), 0.1)
    pts = np.concatenate([th[None, :] + np.diag(h), th[None, :] - np.diag(h)])
    vals = orc.evaluate_batch(cfg, [cc], pts)  # the 40 difference points on all host threads
    fd = (vals[:20] - vals[20:]) / (2 * h)
    err = np.abs(grad - fd).max() / np.abs(fd).max()
    print(f"{key} tsm={tsm}: max |grad - fd| / max |fd| = {err:.2e}")
    # frozen variance: exact to the differences' own accuracy without the TSM variance, to O(variance / sigma_obs^2) with it
    assert err < (1e-6 if not tsm else 1e-5)


def _emu_sens(hcfg, th, g, k, rows):
    emu = load_emu()
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    out = np.zeros((rows.size, 12))
    assert emu.emu_sensitivities(C.byref(hcfg), th.ctypes.data_as(C.c_void_p), np.ascontiguousarray(g).ctypes.data_as(C.c_void_p), k,
                                 rows.ctypes.data_as(C.c_void_p), rows.size, out.ctypes.data_as(C.c_void_p)) == 0
    return out


@pytest.mark.parametrize("over", [dict(), dict(K_hill=0.0), dict(n_hill=2.5), dict(alpha_const=3.0)], ids=["hill4", "nohill", "hillpow", "alpha"])
def test_structured_sensitivity_recursion_matches_the_dense_oracle_on_the_host(over):
    """hmx_grad.cuh (block-structured exact solve) vs hamilton_oracle_grad.c (dense exact Jacobian + dgesv): the state
    sensitivities dg[idx]/dtheta_k for every parameter."""
    cd, th, cfg, cc, kw = _setup("Dysbiotic_HOBIC", 2, False, maxtimestep=360, **over)
    cc.n_obs = 3
    for o, r in enumerate((60, 113, 339)):
        cc.idx_sparse[o] = r
    _, g, *_ = orc.run_deterministic(cfg, th)
    _, grad, ds = orc.loglik_grad(cfg, cc, th, want_dstate=True)
    hcfg = _abi.make_config([workloads.cond_struct(cd, tsm_variance=False)], **kw)
    worst = 0.0
    for k in range(20):
        S = _emu_sens(hcfg, th, g, k, [60, 113, 339])
        ref = ds[:, :, k]
        worst = max(worst, float(np.abs(S - ref).max() / max(np.abs(ref).max(), 1e-300)))
    print(f"{over}: structured vs dense sensitivities, worst scaled error {worst:.2e}")
    assert worst < 1e-9


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_gradient_matches_oracle_and_values_match_loglik_batch(built):
    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(6, seed=31)
    eng = HamiltonEngine(wl.config, device=0)
    try:
        ll, grad, ds, st = eng.loglik_grad_batch(wl.theta, wl.cond_id, want_dstate=True)
        assert np.array_equal(ll, eng.loglik_batch(wl.theta, wl.cond_id))
        pairs = orc.from_hmx_config(wl.config)
        worst_g = worst_s = 0.0
        for i in range(wl.theta.shape[0]):
            scfg, cc = pairs[int(wl.cond_id[i])]
            llo, go, dso = orc.loglik_grad(scfg, cc, wl.theta[i], want_dstate=True)
            assert abs(ll[i] - llo) <= 1e-9 * abs(llo)
            worst_g = max(worst_g, float(np.abs(grad[i] - go).max() / np.abs(go).max()))
            worst_s = max(worst_s, float(np.abs(ds[i, : cc.n_obs] - dso).max() / np.abs(dso).max()))
        print(f"GPU gradient vs oracle: grad {worst_g:.2e}, state sensitivities {worst_s:.2e}")
        assert worst_g < 1e-9 and worst_s < 1e-9
        # parameters the model does not depend on (b_i with alpha = 0) have exactly zero gradient
        assert not grad[:, [3, 4, 8, 9, 15]].any()
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_gradient_equals_central_differences_at_a_tight_newton_tolerance(built):
    """End to end on the device: gradient vs central differences of hmx_loglik_batch itself (eps = 1e-13)."""
    from tmcmc202601_b200.engine import HamiltonEngine

    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    cfg = _abi.make_config([workloads.cond_struct(cd, tsm_variance=False)], eps=1e-13, max_newton_iter=200, **KW)
    eng = HamiltonEngine(cfg, device=0)
    try:
        b = np.array(cd["bounds"])
        th = np.random.default_rng(5).uniform(b[:, 0], b[:, 1], size=(3, 20))
        ll, grad, st = eng.loglik_grad_batch(th)
        for i in range(3):
            h = 1e-5 * np.maximum(np.abs(th[i]), 0.1)
            tp = th[i][None, :] + np.diag(h)
            tm = th[i][None, :] - np.diag(h)
            fd = (eng.loglik_batch(tp) - eng.loglik_batch(tm)) / (2 * h)
            err = np.abs(grad[i] - fd).max() / np.abs(fd).max()
            print(f"theta {i}: max |grad - fd| / max |fd| = {err:.2e}")
            assert err < 1e-6
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_gradient_multi_chunk_and_device_pointers(built, monkeypatch):
    import torch

    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(10, seed=8, solver=dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=2.0))
    for k in range(4):
        for o, r in enumerate((68, 136, 226, 339, 475)):
            wl.config.cond[k].idx_sparse[o] = r
    eng = HamiltonEngine(wl.config, device=0)
    monkeypatch.setenv("HMX_CHUNK_EVALS", "16")
    eng2 = HamiltonEngine(wl.config, device=0)  # 40 evaluations in chunks of 16
    try:
        ll, grad, st = eng.loglik_grad_batch(wl.theta, wl.cond_id)
        ll2, grad2, st2 = eng2.loglik_grad_batch(wl.theta, wl.cond_id)
        assert np.array_equal(ll, ll2) and np.array_equal(grad, grad2)
        eng.use_torch_stream()
        th, cid = torch.from_numpy(wl.theta).cuda(), torch.from_numpy(wl.cond_id).cuda()
        lld = torch.empty(ll.size, dtype=torch.float64, device="cuda")
        gd = torch.empty((ll.size, 20), dtype=torch.float64, device="cuda")
        eng._check(eng._lib.hmx_loglik_grad_batch(eng._ctx, th.data_ptr(), cid.data_ptr(), ll.size, lld.data_ptr(), gd.data_ptr(), None, None), "grad")
        torch.cuda.synchronize()
        assert np.array_equal(gd.cpu().numpy(), grad) and np.array_equal(lld.cpu().numpy(), ll)
    finally:
        eng.close()
        eng2.close()
