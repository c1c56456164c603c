"""Run-directory writer against the reference's stored production run
(_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30 -> tests/golden/rundir_dh_1k30.npz, rundir_schema.json): same file names, same
JSON keys, and -- fed with the reference's own samples -- the same numbers in parameter_summary.csv,
mcmc_diagnostics.json, theta_MAP.json, theta_mean.json and credible_intervals.json."""
import csv
import json

import numpy as np

from helpers import GOLDEN, OracleEngine, OracleStageOps
from tmcmc202601_b200 import posterior, rundir, workloads
from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator


def _golden():
    g = np.load(GOLDEN / "rundir_dh_1k30.npz")
    s = np.load(GOLDEN / "posterior_dh_1k30_samples.npz")
    schema = json.loads((GOLDEN / "rundir_schema.json").read_text())
    return g, s, schema


def test_rhat_and_ess_equal_the_reference_run():
    g, s, _ = _golden()
    chains = [s["samples"][:1000], s["samples"][1000:]]
    d = posterior.compute_mcmc_diagnostics(chains, [g["logL"][:1000], g["logL"][1000:]], list(g["names"]))
    assert np.allclose(d["rhat"], g["rhat"], rtol=1e-13, atol=0)
    assert np.allclose(d["ess"], g["ess"], rtol=1e-12, atol=0)
    assert d["converged"] and d["n_samples_total"] == 2000 and d["n_chains"] == 2
    assert np.isnan(posterior.compute_gelman_rubin_rhat(chains[:1])).all()
    # white noise: ESS ~ n; AR(1) with rho = 0.9: ESS ~ n (1 - rho) / (1 + rho)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(4000, 1))
    assert posterior.compute_effective_sample_size(x)[0] > 3000
    y = np.zeros(4000)
    for i in range(1, 4000):
        y[i] = 0.9 * y[i - 1] + rng.normal()
    assert 100 < posterior.compute_effective_sample_size(y[:, None])[0] < 450


def test_run_directory_matches_reference_layout_and_numbers(tmp_path):
    g, s, schema = _golden()
    cd = workloads.load_conditions()["Dysbiotic_HOBIC_legacy_1k30"]
    samples, logL = s["samples"], g["logL"]
    chains, lls = [samples[:1000], samples[1000:]], [logL[:1000], logL[1000:]]
    diag = {"beta_schedules": [[0.0, 0.3, 1.0], [0.0, 0.4, 1.0]], "acc_rate_histories": [[0.5, 0.3], [0.4, 0.3]], "stage_summaries": [[], []],
            "likelihood_health_histories": [{}, {}], "timing_breakdown_s": [None, None], "n_rom_evaluations": [1, 1],
            "n_fom_evaluations": [0, 0], "log_evidence": [3.0, 3.2]}
    theta_base = np.full(20, 1.5)
    res = rundir.summarize_run(chains, lls, diag, theta_base, list(range(20)), [tuple(b) for b in cd["bounds"]], [True, True], 226864.8,
                               hdi_credibility=float(s["credibility"]), sigma_obs=cd["sigma_scalar"])
    cfg = {k: None for k in schema["config.json"]}
    out = rundir.write_run_directory(tmp_path / "run", res, cfg, cd["data"], cd["idx_sparse"], np.array(cd["days"]))
    # file names: everything the reference run holds, except its logs / figures / tables and fit_metrics (GPU test)
    ref_files = set(schema["_files"]) - {"figures", "figures_pub", "diagnostics_tables", "log_tmcmc.txt", "fit_metrics.json"}
    assert ref_files <= {p.name for p in out.iterdir()}
    for extra in ("chains", "tmcmc_diagnostics", "prior_bounds.json", "model_evidence.json", "sigma_info.json"):
        assert (out / extra).exists()
    # array layouts
    for name, shape in schema["_npy_shapes"].items():
        assert list(np.load(out / name).shape) == shape, name
    assert np.array_equal(np.load(out / "chains" / "chain_1.npy"), samples[1000:])
    # JSON keys
    for name, keys in schema.items():
        if name.startswith("_") or name in ("fit_metrics.json",):
            continue
        mine = json.loads((out / name).read_text())
        assert set(keys) <= set(mine), (name, set(keys) - set(mine))
    # numbers
    with open(out / "parameter_summary.csv") as fh:
        rows = list(csv.DictReader(fh))
    assert list(rows[0].keys()) == list(g["csv_header"])
    assert [r["name"] for r in rows] == list(g["names"]) and [int(r["index"]) for r in rows] == list(g["index"])
    for col, want in (("MAP", g["MAP"]), ("mean", g["mean"]), ("rhat", g["rhat"]), ("ess", g["ess"]), ("median", s["median"]),
                      ("hdi_lower", s["hdi_lower"]), ("hdi_upper", s["hdi_upper"]), ("et_lower", s["et_lower"]), ("et_upper", s["et_upper"])):
        got = np.array([float(r[col]) for r in rows])
        assert np.allclose(got, want, rtol=1e-12, atol=0), col
    tm = json.loads((out / "theta_MAP.json").read_text())
    assert np.allclose(tm["theta_full"], g["theta_MAP_full"], rtol=0, atol=0) and tm["active_indices"] == list(range(20))
    assert np.allclose(json.loads((out / "theta_mean.json").read_text())["theta_full"], g["theta_mean_full"], rtol=1e-14)
    ev = json.loads((out / "model_evidence.json").read_text())
    assert ev["method"] == "harmonic_mean" and np.isfinite(ev["log_evidence"]) and ev["tmcmc"]["log_evidence"] == 3.1
    assert json.loads((out / "sigma_info.json").read_text()) == {"type": "scalar", "value": cd["sigma_scalar"]}
    pb = json.loads((out / "prior_bounds.json").read_text())
    assert pb["bounds"][18]["name"] == "a35" and pb["bounds"][18]["max"] == cd["bounds"][18][1]


def test_estimate_end_to_end_on_the_oracle_engine(tmp_path):
    """The caller-side driver: locked parameters (Commensal Static: 9 locks), replicate sigma matrix, device mutation."""

    def factory(sk, active, theta_base, cd, sigma, cov_rel, theta_lin):
        rows = np.array([4, 8, 12, 16, 19])

        def eng(cfg):
            for o, v in enumerate(rows):
                cfg.cond[0].idx_sparse[o] = cfg.cond[1].idx_sparse[o] = int(v)
            return OracleEngine(cfg)

        return LogLikelihoodEvaluator(dict(sk, maxtimestep=20), [0, 1, 2, 3, 4], active, theta_base, cd["data"], rows, sigma, cov_rel=cov_rel,
                                      theta_linearization=theta_lin, engine_factory=eng)

    res = rundir.estimate("Commensal_Static", tmp_path / "cs", n_particles=60, n_chains=2, n_mutation_steps=2, seed=3, with_fit_metrics=False,
                          evaluator_factory=factory, stage_ops=OracleStageOps(), device_mutation=True)
    out = tmp_path / "cs"
    cd = workloads.load_conditions()["Commensal_Static"]
    assert np.load(out / "samples.npy").shape == (120, 20) and np.load(out / "logL.npy").shape == (120,)
    assert len(res["param_names"]) == len(cd["active_indices"]) == 20 - len(cd["locked"])
    with open(out / "parameter_summary.csv") as fh:
        rows = list(csv.DictReader(fh))
    assert len(rows) == len(cd["active_indices"]) and [int(r["index"]) for r in rows] == list(cd["active_indices"])
    tm = json.loads((out / "theta_MAP.json").read_text())
    assert len(tm["theta_sub"]) == len(cd["active_indices"]) and len(tm["theta_full"]) == 20
    assert all(tm["theta_full"][i] == 0.0 for i in cd["locked"])  # locked parameters stay at zero (bounds (0, 0))
    assert json.loads((out / "sigma_info.json").read_text())["type"] == "heteroscedastic_matrix"
    assert np.load(out / "sigma_matrix.npy").shape == (5, 5)
    cfg = json.loads((out / "config.json").read_text())
    assert cfg["condition"] == "Commensal" and cfg["cultivation"] == "Static" and cfg["metadata"]["device_mutation"] is True
    bs = json.loads((out / "tmcmc_diagnostics" / "beta_schedule.json").read_text())
    assert set(bs) == {"chain_0", "chain_1"} and bs["chain_0"][-1] == 1.0
