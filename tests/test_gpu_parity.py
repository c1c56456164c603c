"""Parity tests proper: the CUDA path, called through the C ABI (libhamilton.so), against the CPU oracle on
the same seeded inputs, against the committed golden vectors of the LIVE reference, and -- at sizes the
oracle cannot cover -- through size-independent properties.

Bar (BASELINE.json north_star): trajectories and log-likelihoods within 1e-9 relative in FP64."""
import json

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, workloads

pytestmark = pytest.mark.gpu

REL = 1e-9


@pytest.fixture(scope="module")
def engines(built):
    from tmcmc202601_b200.engine import HamiltonEngine

    made = []

    def make(cfg):
        e = HamiltonEngine(cfg, device=0)
        made.append(e)
        return e

    yield make
    for e in made:
        e.close()


def _rel(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def test_four_condition_batch_matches_oracle(engines):
    wl = workloads.four_condition_workload(48, seed=20260101, interleave=True)
    eng = engines(wl.config)
    logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
    ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
    rel = _rel(logL, ref)
    print(f"logL rel error: max {rel.max():.2e}, p99 {np.quantile(rel, 0.99):.2e}, median {np.median(rel):.2e}")
    assert rel.max() < REL
    # same quasi-Newton iterate sequence: iteration counts agree evaluation by evaluation except where a
    # line-search / convergence comparison sits within rounding of its threshold (FMA contraction differs
    # between nvcc and the oracle's -ffp-contract=off build); such flips move a count by a few units.
    it64 = it.astype(np.int64)
    same = float(np.mean(it64 == rit))
    print(f"iteration counts identical for {100 * same:.1f}% of evaluations, total {it64.sum()} vs oracle {rit.sum()}")
    assert same > 0.8 and abs(int(it64.sum()) - int(rit.sum())) <= 1e-3 * rit.sum()
    assert np.array_equal(st & 1, rst & 1)
    c = eng.counters()
    assert c["evals"] == logL.size and c["newton_iters"] == int(it64.sum()) and c["q_evals"] >= c["newton_iters"]


def test_golden_evaluator_values_of_live_reference(engines):
    k = np.load(GOLDEN / "evaluator_kat.npz")
    conds = workloads.load_conditions()
    kw = json.loads(str(k["solver_kwargs"]))
    for key in sorted(set(map(str, k["cond"]))):
        cd = conds[key]
        sig = cd["sigma_scalar"] if key.endswith("legacy_1k30") else cd["sigma_obs"]
        structs = [workloads.cond_struct(cd, sigma_obs=sig), workloads.cond_struct(cd, sigma_obs=sig, tsm_variance=False)]
        eng = engines(_abi.make_config(structs, **kw))
        sel = [i for i, c in enumerate(k["cond"]) if str(c) == key]
        th = k["theta_full"][sel]
        ll = eng.loglik_batch(th, np.zeros(len(sel), np.int32))
        ll0 = eng.loglik_batch(th, np.ones(len(sel), np.int32))
        assert _rel(ll, k["logL"][sel]).max() < REL, key
        assert _rel(ll0, k["logL_det"][sel]).max() < REL, key


def test_golden_trajectories_of_live_reference(engines):
    k = np.load(GOLDEN / "solver_kat.npz")
    from tmcmc202601_b200.solver import BiofilmNewtonSolver5S

    worst = 0.0
    for i, name in enumerate(k["names"]):
        kw = json.loads(str(k["kwargs"][i]))
        rows = k["rows"][i]
        rows = rows[rows >= 0]
        s = BiofilmNewtonSolver5S(**kw)
        t_arr, g = s.run_deterministic(k["theta"][i])
        assert g.shape == (kw["maxtimestep"] + 1, 12) and t_arr.shape == (kw["maxtimestep"] + 1,)
        assert t_arr[0] == 0.0 and t_arr[-1] == pytest.approx(kw["maxtimestep"] * kw["dt"])
        rel = _rel(g[rows], k["g_rows"][i][: len(rows)])
        worst = max(worst, rel.max())
        assert rel.max() < REL, (str(name), rel.max())
        s._engine.close()
    print(f"trajectory rows vs live numba reference: worst rel {worst:.2e}")


@pytest.mark.parametrize("extra", [
    dict(workloads.T500_SOLVER), dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=2.0),
    dict(workloads.PRODUCTION_SOLVER, n_hill=1.0), dict(workloads.PRODUCTION_SOLVER, n_hill=2.5),
    dict(workloads.PRODUCTION_SOLVER, n_hill=3.0, alpha_const=2.0),
    dict(workloads.PRODUCTION_SOLVER, eta_vec=[1.0, 1.3, 0.8, 1.1, 0.9], eta_phi_vec=[1.2, 0.7, 1.0, 1.5, 0.95]),
    dict(workloads.PRODUCTION_SOLVER, maxtimestep=300, max_newton_iter=7),
], ids=["t500", "t500_hill", "n1", "n2.5", "n3_alpha", "eta", "maxiter7"])
def test_solver_variants_match_oracle(engines, extra):
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    T = extra["maxtimestep"]
    idx = np.unique(np.clip((cd["idx_sparse"] * T) // 2500, 1, T))
    hc = _abi.make_cond(cd["data"][: len(idx)], idx, cd["sigma_obs"][: len(idx)], 0.2 if T == 500 else cd["phi_init"],
                        active_theta_indices=cd["active_indices"])
    cfg = _abi.make_config([hc], **extra)
    eng = engines(cfg)
    th = workloads.sample_prior(cd["bounds"], 24, np.random.default_rng(3))
    logL, st, it = eng.loglik_batch(th, None, return_status=True)
    ref, rst, rit, _ = orc.evaluate_hmx(cfg, th, None, details=True)
    assert _rel(logL, ref).max() < REL
    assert abs(int(it.astype(np.int64).sum()) - int(rit.sum())) <= 2e-3 * rit.sum()
    g, _, _ = eng.trajectory_batch(th[:3])
    (ocfg, _), = orc.from_hmx_config(cfg)
    for j in range(3):
        _, go, *_ = orc.run_deterministic(ocfg, th[j])
        assert _rel(g[j], go).max() < REL


def test_obs_state_output_equals_trajectory_rows(engines):
    wl = workloads.single_condition_workload("Commensal_HOBIC", 6, seed=5)
    eng = engines(wl.config)
    logL, obs = eng.loglik_batch(wl.theta, wl.cond_id, return_obs=True)
    g, _, _ = eng.trajectory_batch(wl.theta, wl.cond_id)
    idx = np.array(wl.config.cond[0].idx_sparse[:5])
    assert np.array_equal(obs, g[:, idx, :])  # same kernel arithmetic -> bitwise


def test_use_absolute_volume_and_weights(engines):
    cd = workloads.load_conditions()["Dysbiotic_Static"]
    from tmcmc202601_b200.evaluator import build_likelihood_weights

    W = build_likelihood_weights(5, 5)
    structs = [workloads.cond_struct(cd, weights=W), workloads.cond_struct(cd, use_absolute_volume=True)]
    cfg = _abi.make_config(structs, **workloads.PRODUCTION_SOLVER)
    eng = engines(cfg)
    th = workloads.sample_prior(cd["bounds"], 10, np.random.default_rng(8))
    th = np.vstack([th, th])
    cid = np.repeat([0, 1], 10).astype(np.int32)
    ref = orc.evaluate_hmx(cfg, th, cid)
    assert _rel(eng.loglik_batch(th, cid), ref).max() < REL


def test_edge_cases_empty_single_nonfinite(engines):
    wl = workloads.single_condition_workload("Dysbiotic_HOBIC", 4, seed=1)
    eng = engines(wl.config)
    assert eng.loglik_batch(np.zeros((0, 20))).shape == (0,)
    one = eng.loglik_batch(wl.theta[:1])
    assert one.shape == (1,) and one[0] == eng.loglik_batch(wl.theta)[0]
    bad = wl.theta.copy()
    bad[1, 0] = np.nan
    bad[2, 14] = np.inf
    ll, st, _ = eng.loglik_batch(bad, return_status=True)
    assert ll[1] == -1e20 and ll[2] == -1e20 and (st[1] & 1) and (st[2] & 1)  # the reference's sentinel, never an error
    assert ll[0] == one[0] and not (st[0] & 1) and np.isfinite(ll[3])
    with pytest.raises(ValueError):
        eng.loglik_batch(wl.theta, np.full(4, 7, np.int32))
    with pytest.raises(ValueError):
        eng.loglik_batch(np.zeros((3, 19)))


def test_device_pointer_path_is_bitwise_equal_to_host_path(engines):
    import torch

    wl = workloads.four_condition_workload(40, seed=77)
    eng = engines(wl.config)
    host = eng.loglik_batch(wl.theta, wl.cond_id)
    dev = torch.device("cuda", 0)
    th = torch.from_numpy(wl.theta).to(dev)
    cid = torch.from_numpy(wl.cond_id).to(dev)
    out = torch.empty(th.shape[0], dtype=torch.float64, device=dev)
    eng.use_torch_stream()
    eng.loglik_batch_device(th, cid, out)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host)
    # and deterministic launch to launch
    assert np.array_equal(eng.loglik_batch(wl.theta, wl.cond_id), host)


def test_large_batch_is_order_independent(engines):
    """Beyond oracle sizes: 60k evaluations (more than one resident wave, dynamic work fetch) must give every
    duplicate of an input the bit-identical result wherever it lands in the batch."""
    wl = workloads.four_condition_workload(64, seed=9)
    eng = engines(wl.config)
    base = eng.loglik_batch(wl.theta, wl.cond_id)
    rng = np.random.default_rng(0)
    pick = rng.integers(0, wl.theta.shape[0], size=60000)
    big = eng.loglik_batch(wl.theta[pick], wl.cond_id[pick])
    assert np.array_equal(big, base[pick])
    assert np.isfinite(big).all() and (big > -1e19).all()


def test_multi_chunk_batches_equal_single_chunk(engines, monkeypatch):
    """Batches above the chunk size (4 Mi evaluations by default) are processed in pieces with staging-buffer
    reuse; HMX_CHUNK_EVALS shrinks the chunk so the path is exercised at test size."""
    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(50, seed=21, interleave=True)
    ref = engines(wl.config).loglik_batch(wl.theta, wl.cond_id, return_status=True, return_obs=True)
    monkeypatch.setenv("HMX_CHUNK_EVALS", "37")
    eng = HamiltonEngine(wl.config, device=0)
    try:
        got = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True, return_obs=True)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
        c = eng.counters()
        assert c["newton_iters"] == int(ref[2].astype(np.int64).sum())  # counters accumulate over the chunks
        g1, _, _ = eng.trajectory_batch(wl.theta[:80], wl.cond_id[:80])
    finally:
        eng.close()
    monkeypatch.delenv("HMX_CHUNK_EVALS")
    g0, _, _ = engines(wl.config).trajectory_batch(wl.theta[:80], wl.cond_id[:80])
    assert np.array_equal(g0, g1)


def test_parity_distribution_over_prior_draws(engines):
    """10^4 prior draws (2500 per condition): the whole distribution of the relative logL error must sit below
    the 1e-9 bar, not just a handful of cases; iteration counts may differ only where a threshold comparison
    is within rounding (a few units on a minority of evaluations)."""
    wl = workloads.four_condition_workload(2500, seed=777)
    eng = engines(wl.config)
    logL, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
    ref, rst, rit, _ = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id, details=True)
    rel = _rel(logL, ref)
    print(f"rel err over {rel.size} evaluations: max {rel.max():.2e} p99.9 {np.quantile(rel, 0.999):.2e} median {np.median(rel):.2e}; "
          f"identical iteration counts {100 * np.mean(it.astype(np.int64) == rit):.1f}%")
    assert rel.max() < REL
    assert np.quantile(rel, 0.999) < 1e-10
    assert np.array_equal(st & 1, rst & 1)


def test_active_species_masks_against_live_reference_trajectories(built):
    """BiofilmNewtonSolver5S(active_species=..., phi_init=vector) on the device vs the live reference solver."""
    from tmcmc202601_b200.solver import BiofilmNewtonSolver5S

    k = np.load(GOLDEN / "solver_masked_kat.npz")
    for name in ("four_species", "three_species_vec", "all_vec"):
        s = BiofilmNewtonSolver5S(dt=1e-4, maxtimestep=400, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0,
                                  phi_init=k[f"{name}_phi_init"], active_species=list(k[f"{name}_active"]))
        t, g = s.run_deterministic(k[f"{name}_theta"])
        ref = k[f"{name}_g"]
        assert np.array_equal(g[0], ref[0])
        rel = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300)
        assert rel.max() < 1e-9, (name, rel.max())


def test_deviations_above_the_bar_only_where_the_reference_itself_is_ill_conditioned(engines):
    """Outside the production settings (here dt = 1.18e-4, c = 50, non-unit eta: a configuration from the fuzz tool) some
    time steps exhaust their 50 Newton iterations without contracting, and the result of such an evaluation depends on
    rounding: the ORACLE's own value moves by up to 4e-5 when theta is perturbed by two ulps.  Parity at 1e-9 is undefined
    there.  The statement that can be tested: wherever the GPU is further than 1e-9 from the oracle, the oracle is at least
    1e-10 away from itself under a 2-ulp perturbation of its input -- and everywhere else the bar holds."""
    solver = dict(dt=0.00011848512631842148, maxtimestep=500, c_const=50.0, alpha_const=0.0, Kp1=6.89941793513431e-05, K_hill=0.0, n_hill=2.0,
                  eta_vec=[0.858, 1.068, 1.002, 1.164, 0.709], eta_phi_vec=[1.013, 0.956, 0.837, 1.116, 1.005])
    n_far = 0
    for seed in (1001, 1021, 1027):
        wl = workloads.four_condition_workload(100, seed=seed, solver=solver)
        rows = np.unique(np.linspace(500 // 8, 500, 5, dtype=int))
        for c in range(4):
            for o in range(5):
                wl.config.cond[c].idx_sparse[o] = int(rows[o])
        eng = engines(wl.config)
        ll = eng.loglik_batch(wl.theta, wl.cond_id)
        ref = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id)
        sens = np.zeros_like(ref)
        for k in range(3):  # three independent 2-ulp perturbations of every theta
            sgn = np.sign(np.random.default_rng(k).standard_normal(wl.theta.shape))
            sens = np.maximum(sens, _rel(orc.evaluate_hmx(wl.config, wl.theta * (1.0 + 4e-16 * sgn), wl.cond_id), ref))
        rel = _rel(ll, ref)
        far = rel > REL
        n_far += int(far.sum())
        print(f"seed {seed}: {int(far.sum())} of {rel.size} evaluations above 1e-9 (max {rel.max():.1e}); oracle self-sensitivity there >= {sens[far].min() if far.any() else 0:.1e}")
        assert np.all(sens[far] > 1e-10), "a deviation above the bar where the reference is well conditioned"
        assert np.all(rel[sens < 1e-13] < REL)
    assert n_far >= 1  # the configuration was chosen because it has such evaluations
