"""The CPU oracle against the committed golden vectors of the LIVE reference (tests/golden, written by
make_golden.py from numba 0.65 / numpy 2.3.5) and against first-principles properties of the model."""
import json

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import oracle as orc


def _solver_cases():
    k = np.load(GOLDEN / "solver_kat.npz")
    for i, name in enumerate(k["names"]):
        rows = k["rows"][i]
        rows = rows[rows >= 0]
        yield str(name), json.loads(str(k["kwargs"][i])), k["theta"][i], rows, k["g_rows"][i][: len(rows)]


@pytest.mark.parametrize("case", list(_solver_cases()), ids=lambda c: c[0])
def test_oracle_trajectory_matches_live_reference(case):
    name, kw, theta, rows, g_ref = case
    _, g, its, trials, st = orc.run_deterministic(orc.make_solver_cfg(**kw), theta)
    rel = np.abs(g[rows] - g_ref) / np.maximum(np.abs(g_ref), 1e-300)
    assert rel.max() < 1e-9, f"{name}: max rel {rel.max():.2e}"  # observed <= 1e-11 (MANIFEST.json)
    assert np.isfinite(g).all()


def test_oracle_evaluator_matches_live_reference():
    """LogLikelihoodEvaluator.__call__ of the reference (TSM variance ON) and the sig=0 variant."""
    from tmcmc202601_b200 import workloads

    k = np.load(GOLDEN / "evaluator_kat.npz")
    conds = workloads.load_conditions()
    kw = json.loads(str(k["solver_kwargs"]))
    worst = 0.0
    for i, key in enumerate(k["cond"]):
        cd = conds[str(key)]
        sig = cd["sigma_scalar"] if str(key).endswith("legacy_1k30") else cd["sigma_obs"]
        cfg = orc.make_solver_cfg(phi_init=cd["phi_init"], **kw)
        cc = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], sig, cov_rel=0.005, active_theta_indices=cd["active_indices"])
        cc0 = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], sig, cov_rel=0.005, active_theta_indices=cd["active_indices"], tsm_variance=False)
        ll = orc.evaluate(cfg, cc, k["theta_full"][i])
        ll0 = orc.evaluate(cfg, cc0, k["theta_full"][i])
        worst = max(worst, abs(ll - k["logL"][i]) / abs(k["logL"][i]), abs(ll0 - k["logL_det"][i]) / abs(k["logL_det"][i]))
    assert worst < 1e-11, worst
    # the survey's probe value for the production MAP
    j = list(k["names"]).index("Dysbiotic_HOBIC_legacy_1k30#2")
    assert abs(k["logL"][j] - 11.787791376773674) < 1e-9


def test_oracle_resample_bit_exact_with_live_reference():
    k = np.load(GOLDEN / "stage_kat.npz")
    for c in range(int(k["resample_n"])):
        idx = orc.resample(k[f"resample_w{c}"], float(k["resample_u"][c]))
        assert np.array_equal(idx, k[f"resample_idx{c}"])


def test_resample_matches_numpy_semantics():
    rng = np.random.default_rng(0)
    for N in (1, 2, 3, 33, 1000, 4097):
        w = rng.random(N) ** 4
        w /= w.sum()
        u = rng.random()
        cs = np.cumsum(w)
        cs[-1] = 1.0
        ref = np.clip(np.searchsorted(cs, (u + np.arange(N)) / N), 0, N - 1)
        assert np.array_equal(orc.resample(w, u), ref)
    # degenerate weights: everything on one particle
    w = np.zeros(50)
    w[17] = 1.0
    assert np.all(orc.resample(w, 0.3) == 17)


def test_stage_arithmetic_matches_numpy():
    rng = np.random.default_rng(1)
    for N, scale in ((10, 1.0), (1000, 5.0), (5000, 50.0)):
        logL = rng.normal(size=N) * scale
        for beta in (0.0, 0.3, 0.97):
            lo, hi = 0.0, 1.0 - beta
            ess_lo = None
            for _ in range(25):  # TM:617-650 written out with NumPy
                mid = 0.5 * (lo + hi)
                w = np.exp(mid * (logL - logL.max()))
                ess = 1.0 / np.sum((w / np.sum(w)) ** 2)
                if ess >= 0.5 * N:
                    lo, ess_lo = mid, ess
                else:
                    hi = mid
            db_ref = min(max(lo, 0.02), 0.5, 1.0 - beta)
            db, ess = orc.beta_step(logL, beta)
            assert abs(db - db_ref) <= 1e-12 * max(db_ref, 1e-300)
            if ess_lo is not None:
                assert abs(ess - ess_lo) <= 1e-10 * ess_lo
            lw = db * logL
            w_ref = np.exp(lw - lw.max())
            S = w_ref.sum()
            w, ls, flag = orc.weights(logL, db)
            assert flag == 0 and np.allclose(w, w_ref / S, rtol=1e-13, atol=0)
            assert abs(ls - (np.log(S) - np.log(N) + lw.max())) < 1e-12
    theta = rng.normal(size=(2000, 20)) * rng.random(20)
    w = rng.random(2000)
    w /= w.sum()
    mean, cov = orc.weighted_cov(theta, w)
    ref = np.cov(theta.T, aweights=w)
    assert np.allclose(cov, ref, rtol=1e-11, atol=1e-15)
    assert np.allclose(mean, np.average(theta, axis=0, weights=w), rtol=1e-13)
    # 1-parameter edge case (TM:771-775)
    _, c1 = orc.weighted_cov(theta[:, :1], w)
    assert np.allclose(c1, np.cov(theta[:, :1].T, aweights=w), rtol=1e-11)


def test_weights_fallback_to_uniform():
    w, ls, flag = orc.weights(np.full(8, -np.inf), 0.5)
    assert flag == 1 and np.allclose(w, 1.0 / 8)


def test_dgesv_restatement_against_lapack():
    rng = np.random.default_rng(2)
    import ctypes as C

    for n in (1, 2, 5, 12):
        for _ in range(20):
            A = rng.normal(size=(n, n)) + np.diag(rng.normal(size=n))
            b = rng.normal(size=(3, n))  # column-major [n, 3]
            Ac, bc = A.copy(), b.copy()
            info = orc.lib().orc_dgesv(C.c_int(n), Ac.ctypes.data_as(C.c_void_p), C.c_int(3), bc.ctypes.data_as(C.c_void_p))
            assert info == 0
            x = np.linalg.solve(A, b.T).T
            assert np.allclose(bc, x, rtol=1e-9, atol=1e-11)
    Z = np.zeros((3, 3))
    assert orc.lib().orc_dgesv(C.c_int(3), Z.ctypes.data_as(C.c_void_p), C.c_int(1), np.ones(3).ctypes.data_as(C.c_void_p)) > 0


def test_residual_algebra_properties():
    """5-species analogue of the reference's own unit test (tmcmc/program2602/test_hamilton_model_consistency.py:90-225):
    psi-equations do not depend on gamma, Q11 is the volume constraint, the Hill gate only touches species 5."""
    rng = np.random.default_rng(3)
    cfg = orc.make_solver_cfg(K_hill=0.05, n_hill=4.0)
    cfg0 = orc.make_solver_cfg(K_hill=0.0)
    for _ in range(20):
        theta = rng.uniform(0, 3, 20)
        g_old = np.concatenate([rng.uniform(0.01, 0.15, 5), [0.3], rng.uniform(0.5, 0.99, 5), [0.0]])
        g_new = g_old + rng.normal(size=12) * 1e-3
        Q = orc.compute_Q(cfg, theta, g_new, g_old)
        g2 = g_new.copy()
        g2[11] += 7.5
        Q2 = orc.compute_Q(cfg, theta, g2, g_old)
        assert np.allclose(Q2[6:11], Q[6:11], rtol=0, atol=0)  # psi rows: no gamma
        assert np.allclose(Q2[:6] - Q[:6], 7.5, rtol=1e-12)  # gamma enters phi rows with 1/eta = 1
        assert abs(Q[11] - (g_new[:6].sum() - 1.0)) < 1e-15
        Qn = orc.compute_Q(cfg0, theta, g_new, g_old)
        assert np.array_equal(Q[[0, 1, 2, 3, 5, 6, 7, 8, 9, 11]], Qn[[0, 1, 2, 3, 5, 6, 7, 8, 9, 11]])
    # clipping: values outside [1e-12, 1-1e-12] are evaluated at the bound
    g_old = np.concatenate([np.full(5, 0.1), [0.5], np.full(5, 0.9), [0.0]])
    g_a, g_b = g_old.copy(), g_old.copy()
    g_a[2], g_b[2] = -0.5, 1e-12
    assert np.array_equal(orc.compute_Q(cfg, np.ones(20), g_a, g_old), orc.compute_Q(cfg, np.ones(20), g_b, g_old))


def test_nonfinite_theta_gives_sentinel():
    cfg = orc.make_solver_cfg(maxtimestep=50)
    cc = orc.make_cond_cfg(np.full((1, 5), 0.1), [40], 0.1)
    th = np.ones(20)
    th[3] = np.nan  # b1 is unused with alpha = 0 ...
    th[0] = np.nan  # ... a11 is not
    ll, obs, st, its, tr = orc.evaluate(cfg, cc, th, want_obs=True)
    assert ll == -1e20 and (st & orc.ST_NONFINITE)


def test_active_species_masks_and_vector_phi_init_against_live_reference():
    """Solver variants outside the main fixture set (tests/golden/make_golden_masked.py): inactive species start at
    phi = psi = 0 and are held at the clip floor by the barrier (S5:563-575)."""
    k = np.load(GOLDEN / "solver_masked_kat.npz")
    for name in ("four_species", "three_species_vec", "all_vec"):
        cfg = orc.make_solver_cfg(phi_init=k[f"{name}_phi_init"], active_species=list(k[f"{name}_active"]), dt=1e-4, maxtimestep=400,
                                  c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0)
        _, g, *_ = orc.run_deterministic(cfg, k[f"{name}_theta"])
        ref = k[f"{name}_g"]
        assert np.array_equal(g[0], ref[0])
        rel = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300)
        assert rel.max() < 1e-9, (name, rel.max())
