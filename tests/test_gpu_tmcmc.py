"""End-to-end TMCMC on the GPU: the toy trace of the live reference with GPU stage kernels, and a short real
run (Dysbiotic HOBIC constants, ODE likelihood on the device) through the reference-shaped API."""
import numpy as np
import pytest

from helpers import GOLDEN, toy_loglik
from tmcmc202601_b200 import tmcmc, workloads

pytestmark = pytest.mark.gpu


def test_toy_trace_with_gpu_stage_kernels(built):
    k = np.load(GOLDEN / "stage_kat.npz")
    bounds = [tuple(map(float, b)) for b in k["toy_bounds"]]
    res = tmcmc.run_TMCMC(toy_loglik, bounds, n_particles=int(k["toy_n_particles"]), n_stages=int(k["toy_n_stages"]),
                          seed=int(k["toy_seed"]), n_jobs=1, min_delta_beta=0.02, max_delta_beta=0.5,
                          n_mutation_steps=int(k["toy_n_mutation_steps"]))
    assert np.allclose(res.beta_schedule, k["toy_beta_schedule"], rtol=1e-10)
    assert np.allclose(res.acc_rate_history, k["toy_acc_rate"], rtol=1e-12)
    assert abs(res.log_evidence - float(k["toy_log_evidence"])) < 1e-9
    assert np.allclose(res.samples, k["toy_samples"], rtol=1e-8, atol=1e-10)


def test_short_real_tmcmc_run(built):
    from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"])
    theta_base = np.full(20, 1.5)

    def make_evaluator(theta_linearization=None):
        ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], cd["active_indices"], theta_base, cd["data"], cd["idx_sparse"],
                                    cd["sigma_obs"], cov_rel=0.005, theta_linearization=theta_linearization)
        return ev

    chains, logL, MAP, conv, diag = tmcmc.run_multi_chain_TMCMC(
        "Dysbiotic_HOBIC", make_evaluator, [tuple(b) for b in cd["bounds"]], theta_base, cd["active_indices"],
        n_particles=256, n_stages=30, n_chains=1, seed=42, min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=2)
    assert all(conv) and chains[0].shape == (256, 20)
    b = np.array(cd["bounds"])
    assert np.all(chains[0] >= b[:, 0]) and np.all(chains[0] <= b[:, 1])
    assert np.isfinite(logL[0]).all() and logL[0].max() > -50.0
    s = diag["stage_summaries"][0]
    assert s[-1]["beta_next"] == 1.0 and all(r["linearization_enabled"] == 0 for r in s)
    assert diag["likelihood_health_total"]["n_calls"] >= 256 * (1 + len(s))
    # the tempered population must have concentrated: mean logL of the final population beats the prior's
    ev = make_evaluator(theta_base)
    prior_ll = ev.batch(workloads.sample_prior(cd["bounds"], 256, np.random.default_rng(0)))
    assert logL[0].mean() > np.median(prior_ll)


def test_production_run_posterior_matches_reference_run(built):
    """BASELINE.json configs[1]: the stored production run (_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30, N=1000, 2 chains,
    seed 42, K=5; 63 h on the reference) re-run through the drop-in API; posterior summaries must agree with the
    reference's samples within Monte-Carlo error (tests/golden/posterior_dh_1k30.npz)."""
    import sys

    sys.path.insert(0, str(GOLDEN.parents[1] / "scripts"))
    import run_production_tmcmc as rp

    out, samples, ll = rp.run_condition("Dysbiotic_HOBIC_legacy_1k30", 1000, 2)
    cmp_ = rp.compare_with_reference(samples, ll)
    print(f"wall {out['wall_s']:.2f} s for {out['likelihood_evaluations']} evaluations; {cmp_}")
    assert all(out["converged"]) and samples.shape == (2000, 20)
    assert cmp_["z_mean_max"] < 4.5  # 20 parameters, n_eff ~ 200 on both sides
    assert 0.75 < cmp_["std_ratio_min"] and cmp_["std_ratio_max"] < 1.33
    assert cmp_["quantile_shift_over_std_max"] < 0.5
    assert abs(out["logL_max"] - cmp_["logL_max_ref"]) < 1.5 and abs(out["logL_mean"] - cmp_["logL_mean_ref"]) < 0.5
    assert out["wall_s"] < 120.0


def test_merged_four_condition_run_equals_sequential(built):
    """BASELINE.json configs[2]: the four conditions x 2 chains run concurrently with merged launches; every
    chain must equal the chain run on its own (bitwise), and the whole batch must finish in seconds."""
    import sys

    sys.path.insert(0, str(GOLDEN.parents[1] / "scripts"))
    import run_production_tmcmc as rp

    summary, out = rp.run_merged(workloads.CONDITION_KEYS, 200, 2)
    assert all(all(o["converged"]) for o in out) and summary["merged_launches"] < 40
    solo, samples, ll = rp.run_condition("Commensal_HOBIC", 200, 2)
    k = workloads.CONDITION_KEYS.index("Commensal_HOBIC")
    assert np.array_equal(np.concatenate(out[k]["chains"], axis=0), samples)
    assert np.array_equal(np.concatenate(out[k]["logL"]), ll)
    print(summary)


def test_estimate_writes_a_reference_shaped_run_directory(built, tmp_path):
    """SURVEY 8(f)2 on the device: `rundir.estimate` end to end (device-side mutation), and fit_metrics.json at the
    stored reference MAP / mean equal to the reference's own file (trajectories from hmx_trajectory_batch)."""
    import json

    from tmcmc202601_b200 import rundir

    res = rundir.estimate("Dysbiotic_HOBIC_legacy_1k30", tmp_path / "run", n_particles=400, n_chains=2, device_mutation=True)
    out = tmp_path / "run"
    schema = json.loads((GOLDEN / "rundir_schema.json").read_text())
    have = {p.name for p in out.iterdir()}
    assert set(schema["_files"]) - {"figures", "figures_pub", "diagnostics_tables", "log_tmcmc.txt"} <= have
    assert np.load(out / "samples.npy").shape == (800, 20) and all(res["converged"])
    fm = json.loads((out / "fit_metrics.json").read_text())
    assert set(schema["fit_metrics.json"]) <= set(fm) and set(schema["fit_metrics.json"]["MAP"]) <= set(fm["MAP"])
    assert fm["MAP"]["n_obs"] == 6 and fm["MAP"]["rmse_total"] < 0.5

    g = np.load(GOLDEN / "rundir_dh_1k30.npz")
    cd = workloads.load_conditions()["Dysbiotic_HOBIC_legacy_1k30"]
    ref = rundir.fit_metrics_for(g["theta_MAP_full"], g["theta_mean_full"], dict(workloads.PRODUCTION_SOLVER, phi_init=0.02), cd["data"],
                                 cd["idx_sparse"])
    # the reference builds this file with a solver WITHOUT the Hill gate (MAIN:3452-3459): reproduced by default
    from oracle import oracle as orc
    from tmcmc202601_b200 import posterior

    ocfg = orc.make_solver_cfg(phi_init=0.02, **dict(workloads.PRODUCTION_SOLVER, K_hill=0.0, n_hill=2.0))
    for k, th in (("MAP", g["theta_MAP_full"]), ("Mean", g["theta_mean_full"])):
        _, go, *_ = orc.run_deterministic(ocfg, th)
        want = posterior.compute_fit_metrics(None, go, [0, 1, 2, 3, 4], cd["data"], cd["idx_sparse"])
        for m in ("rmse_per_species", "mae_per_species", "rmse_total", "mae_total", "max_abs"):
            assert np.allclose(ref[k][m], want[m], rtol=1e-9, atol=0), (k, m)
            # the stored file was written by an older revision of the reference solver: same quirk, 1e-5-level drift
            assert np.allclose(ref[k][m], g[f"fit_{k}_{m}"], rtol=2e-4, atol=0), (k, m)
    gated = rundir.fit_metrics_for(g["theta_MAP_full"], g["theta_mean_full"], dict(workloads.PRODUCTION_SOLVER, phi_init=0.02), cd["data"],
                                   cd["idx_sparse"], with_hill_gate=True)
    assert gated["MAP"]["rmse_total"] < 0.6 * ref["MAP"]["rmse_total"]  # the calibrated model fits far better than the plotted one
