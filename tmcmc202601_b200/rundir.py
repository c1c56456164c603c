"""Run-directory writer: the files `run_estimation` + `main` of the reference leave behind
(data_5species/main/estimate_reduced_nishioka.py:2904-3430, 3690-3717), with the same names, keys and array
layouts, so Stage-2 scripts and plotters consume a GPU run unchanged:

    config.json  data.npy  idx_sparse.npy  t_days.npy  samples.npy  logL.npy  chains/chain_<i>.npy
    tmcmc_diagnostics/{beta_schedule,acceptance_rates,stage_summaries,linearization_history,solver_health,
                       timing_breakdown,rom_info}.json
    sigma_info.json (+ sigma_matrix.npy)  prior_bounds.json  theta_MAP.json  theta_mean.json
    mcmc_diagnostics.json  credible_intervals.json  parameter_summary.csv  model_evidence.json
    results_summary.json  fit_metrics.json

`estimate(...)` is the caller-side driver for the conditions shipped in tmcmc202601_b200/data (the reference's
`run_estimation` for `--condition X --cultivation Y`): multi-chain TMCMC on the GPU, summaries, and the directory.
Host-side work throughout; the likelihood, the stage arithmetic, (optionally) the mutation and the MAP / mean
trajectories for fit_metrics.json run on the device.
"""

from __future__ import annotations

import csv
import json
import time
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import posterior, tmcmc, workloads
from .evaluator import LogLikelihoodEvaluator

# MAIN:2323-2346
PARAM_NAMES = [
    "a11", "a12", "a22", "b1", "b2", "a33", "a34", "a44", "b3", "b4", "a13", "a14", "a23", "a24", "a55", "b5",
    "a15", "a25", "a35", "a45", "log_tau_relax", "E0_Einf_ratio",
]
SPECIES_NAMES = ["S. oralis", "A. naeslundii", "V. dispar", "F. nucleatum", "P. gingivalis"]


def to_jsonable(obj):
    """NumPy-safe conversion (utils/io.py: arrays -> lists, NumPy scalars -> Python scalars)."""
    if isinstance(obj, dict):
        return {str(k): to_jsonable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_jsonable(v) for v in obj]
    if isinstance(obj, np.ndarray):
        return to_jsonable(obj.tolist())
    if isinstance(obj, (np.floating,)):
        return float(obj)
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.bool_,)):
        return bool(obj)
    if isinstance(obj, Path):
        return str(obj)
    return obj


def save_json(path: Path, payload) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with path.open("w", encoding="utf-8") as f:
        json.dump(to_jsonable(payload), f, indent=2, ensure_ascii=False)


def save_npy(path: Path, arr) -> None:
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    np.save(path, np.asarray(arr))


def estimate_log_evidence_harmonic_mean(logL: np.ndarray, stabilize: bool = True):
    """Harmonic-mean estimate over the upper half of the log-likelihoods (MAIN:464-508)."""
    logL = np.asarray(logL, dtype=np.float64)
    sel = logL[logL >= np.median(logL)] if stabilize else logL
    log_ev = -(np.logaddexp.reduce(-sel) - np.log(len(sel)))
    return float(log_ev), float(np.std(sel) / np.sqrt(len(sel)))


def compute_model_evidence(chains, logL, prior_bounds, active_indices, tmcmc_log_evidence: Optional[Sequence[float]] = None) -> Dict[str, Any]:
    """model_evidence.json (MAIN:551-625).  The reference has the harmonic mean as its working estimator (its
    thermodynamic branch needs per-stage means the sampler never records); the TMCMC estimate sum_j log S_j that
    run_TMCMC accumulates anyway is reported next to it."""
    all_logL = np.concatenate([np.asarray(l, dtype=np.float64) for l in logL])
    log_ev, se = estimate_log_evidence_harmonic_mean(all_logL, stabilize=True)
    out: Dict[str, Any] = {"method": "harmonic_mean", "log_evidence": log_ev, "log_evidence_se": se,
                           "harmonic_mean": {"log_evidence": log_ev, "se": se}}
    if tmcmc_log_evidence is not None:
        out["tmcmc"] = {"log_evidence_per_chain": [float(x) for x in tmcmc_log_evidence],
                        "log_evidence": float(np.mean(tmcmc_log_evidence))}
    act = [prior_bounds[i] for i in active_indices]
    out["log_prior_volume"] = float(sum(np.log(b[1] - b[0]) for b in act if b[1] > b[0]))
    out["n_params"] = len(active_indices)
    out["max_logL"] = float(np.max(all_logL))
    out["BIC"] = float(-2 * out["max_logL"] + len(active_indices) * np.log(len(all_logL)))
    return out


def summarize_run(chains, logL, diagnostics, theta_base, active_indices, prior_bounds, converged, elapsed,
                  hdi_credibility: float = 0.95, sigma_obs=None) -> Dict[str, Any]:
    """The `results` dict run_estimation returns (MAIN:2300-2440)."""
    samples = np.concatenate(chains, axis=0)
    logL_all = np.concatenate(logL, axis=0)
    summ = tmcmc.compute_MAP_with_uncertainty(samples, logL_all)
    MAP, mean = summ["MAP"], summ["mean"]
    theta_base = np.asarray(theta_base, dtype=np.float64)
    if MAP.shape[0] == len(theta_base):
        MAP_full, mean_full = MAP.copy(), mean.copy()
    else:
        MAP_full, mean_full = theta_base.copy(), theta_base.copy()
        MAP_full[active_indices] = MAP
        mean_full[active_indices] = mean
    names = [PARAM_NAMES[i] for i in active_indices if i < len(PARAM_NAMES)]
    diag = posterior.compute_mcmc_diagnostics(chains, logL, names)
    ci = posterior.compute_credible_intervals(samples, hdi_credibility)
    ev = compute_model_evidence(chains, logL, prior_bounds, active_indices,
                                tmcmc_log_evidence=(diagnostics or {}).get("log_evidence"))
    return {
        "samples": samples, "logL": logL_all, "chains": list(chains), "MAP": MAP, "mean": mean, "theta_MAP_full": MAP_full,
        "theta_mean_full": mean_full, "elapsed_time": float(elapsed), "converged": [bool(c) for c in converged],
        "diagnostics": diagnostics, "mcmc_diagnostics": diag, "rhat": diag["rhat"], "ess": diag["ess"],
        "credible_intervals": ci, "evidence": ev, "active_indices": list(active_indices), "param_names": names,
        "prior_bounds": [tuple(b) for b in prior_bounds], "sigma_obs": sigma_obs,
    }


def _active(arr, active_indices, n_full):
    arr = np.asarray(arr)
    return arr[active_indices] if (len(arr) == n_full and len(active_indices) < n_full) else arr


def write_run_directory(output_dir, results: Dict[str, Any], config: Dict[str, Any], data, idx_sparse, t_days,
                        fit_metrics: Optional[Dict[str, Any]] = None) -> Path:
    """Write everything listed in the module docstring.  `results` as returned by summarize_run()."""
    out = Path(output_dir)
    out.mkdir(parents=True, exist_ok=True)
    save_npy(out / "data.npy", data)
    save_npy(out / "idx_sparse.npy", idx_sparse)
    save_npy(out / "t_days.npy", t_days)
    save_json(out / "config.json", config)
    save_npy(out / "samples.npy", results["samples"])
    save_npy(out / "logL.npy", results["logL"])
    for ci_, chain in enumerate(results.get("chains") or []):
        save_npy(out / "chains" / f"chain_{ci_}.npy", chain)

    dd = results.get("diagnostics") or {}
    if dd:
        ddir = out / "tmcmc_diagnostics"
        if "beta_schedules" in dd:
            save_json(ddir / "beta_schedule.json", {f"chain_{i}": [float(b) for b in bs] for i, bs in enumerate(dd["beta_schedules"])})
        if "acc_rate_histories" in dd:
            save_json(ddir / "acceptance_rates.json", {f"chain_{i}": [float(a) for a in ar] for i, ar in enumerate(dd["acc_rate_histories"])})
        if "stage_summaries" in dd:
            save_json(ddir / "stage_summaries.json", dd["stage_summaries"])
        if "theta0_history" in dd:
            save_json(ddir / "linearization_history.json",
                      {f"chain_{i}": [[float(x) for x in th] for th in hist] for i, hist in enumerate(dd["theta0_history"])})
        if "likelihood_health_histories" in dd:
            save_json(ddir / "solver_health.json", dd["likelihood_health_histories"])
        if "timing_breakdown_s" in dd:
            save_json(ddir / "timing_breakdown.json", dd["timing_breakdown_s"])
        rom = {k: dd[k] for k in ("n_rom_evaluations", "n_fom_evaluations", "rom_error_histories") if k in dd}
        if rom:
            save_json(ddir / "rom_info.json", rom)

    sig = results.get("sigma_obs")
    if sig is not None:
        if isinstance(sig, np.ndarray) and sig.ndim >= 1 and sig.size > 1:
            save_npy(out / "sigma_matrix.npy", sig)
            save_json(out / "sigma_info.json", {"type": "heteroscedastic_matrix", "shape": list(sig.shape), "min": float(sig.min()),
                                                "max": float(sig.max()), "mean": float(sig.mean())})
        else:
            save_json(out / "sigma_info.json", {"type": "scalar", "value": float(np.asarray(sig).reshape(-1)[0])})

    act = list(results["active_indices"])
    names = list(results["param_names"])
    pb = results["prior_bounds"]
    save_json(out / "prior_bounds.json", {
        "active_indices": act, "param_names": names,
        "bounds": [{"name": names[i], "index": int(act[i]), "min": float(pb[act[i]][0]), "max": float(pb[act[i]][1])}
                   for i in range(len(names))],
    })
    n_full = len(PARAM_NAMES[:20]) if len(results["theta_MAP_full"]) <= 20 else len(PARAM_NAMES)
    MAP_a, mean_a = _active(results["MAP"], act, n_full), _active(results["mean"], act, n_full)
    save_json(out / "theta_MAP.json", {"theta_sub": MAP_a.tolist(), "theta_full": np.asarray(results["theta_MAP_full"]).tolist(),
                                       "active_indices": act})
    save_json(out / "theta_mean.json", {"theta_sub": mean_a.tolist(), "theta_full": np.asarray(results["theta_mean_full"]).tolist(),
                                        "active_indices": act})
    md = results["mcmc_diagnostics"]
    save_json(out / "mcmc_diagnostics.json", {k: md[k] for k in ("rhat", "ess", "rhat_max", "ess_min", "converged_rhat", "converged_ess",
                                                                 "converged", "n_samples_total", "n_chains", "per_parameter")})
    ci = results["credible_intervals"]
    save_json(out / "credible_intervals.json", {"credibility": ci["credibility"], "hdi_lower": ci["hdi_lower"], "hdi_upper": ci["hdi_upper"],
                                                "et_lower": ci["et_lower"], "et_upper": ci["et_upper"], "median": ci["median"],
                                                "param_names": names})
    cols = {
        "name": names, "index": act, "MAP": MAP_a, "mean": mean_a, "median": _active(ci["median"], act, n_full),
        "hdi_lower": _active(ci["hdi_lower"], act, n_full), "hdi_upper": _active(ci["hdi_upper"], act, n_full),
        "et_lower": _active(ci["et_lower"], act, n_full), "et_upper": _active(ci["et_upper"], act, n_full),
        "rhat": _active(md["rhat"], act, n_full), "ess": _active(md["ess"], act, n_full),
    }
    bad = {k: len(v) for k, v in cols.items() if len(v) != len(act)}
    if bad:
        raise ValueError(f"Cannot create parameter summary: array length mismatches {bad}")
    with (out / "parameter_summary.csv").open("w", newline="") as fh:
        wr = csv.writer(fh)
        wr.writerow(list(cols.keys()))
        for r in range(len(act)):
            wr.writerow([cols[k][r] if k in ("name", "index") else repr(float(cols[k][r])) for k in cols])
    if results.get("evidence") is not None:
        save_json(out / "model_evidence.json", results["evidence"])
    save_json(out / "results_summary.json", {"elapsed_time": results["elapsed_time"], "converged": results["converged"],
                                             "MAP": np.asarray(results["MAP"]).tolist(), "mean": np.asarray(results["mean"]).tolist()})
    if fit_metrics is not None:
        save_json(out / "fit_metrics.json", fit_metrics)
    return out


def fit_metrics_for(theta_MAP_full, theta_mean_full, solver_kwargs: Dict[str, Any], data, idx_sparse, device: int = 0,
                    with_hill_gate: bool = False) -> Dict[str, Any]:
    """fit_metrics.json (MAIN:3690-3717): trajectories at the MAP and the posterior mean (one batched GPU call).

    Reference quirk kept by default: the solver `main` builds for this file and the plots (MAIN:3452-3459) receives
    dt, maxtimestep, c_const, alpha_const, phi_init and Kp1 only, i.e. it runs WITHOUT the Hill gate (K_hill = 0) even
    when the calibration used one.  `with_hill_gate=True` evaluates the calibrated model instead."""
    from .solver import BiofilmNewtonSolver5S

    keep = ("dt", "maxtimestep", "c_const", "alpha_const", "phi_init", "Kp1") + (("K_hill", "n_hill") if with_hill_gate else ())
    solver = BiofilmNewtonSolver5S(**{k: v for k, v in solver_kwargs.items() if k in keep}, device=device)
    th = np.stack([np.asarray(theta_MAP_full, dtype=np.float64)[:20], np.asarray(theta_mean_full, dtype=np.float64)[:20]])
    g = solver.run_deterministic_batch(th)
    t = np.arange(g.shape[1]) * solver.dt
    out = {}
    for name, traj in (("MAP", g[0]), ("Mean", g[1])):
        m = posterior.compute_fit_metrics(t, traj, [0, 1, 2, 3, 4], data, idx_sparse)
        m["estimate_type"] = name
        out[name] = m
    out["species_names"] = list(SPECIES_NAMES)
    return out


def estimate(condition_key: str, output_dir, n_particles: int = 1000, n_stages: int = 30, n_chains: int = 2, seed: int = 42,
             n_mutation_steps: int = 5, min_delta_beta: float = 0.02, max_delta_beta: float = 0.5, hdi_credibility: float = 0.95,
             cov_rel: float = 0.005, device_mutation: bool = False, device: int = 0, solver_overrides: Optional[Dict[str, Any]] = None,
             with_fit_metrics: bool = True, evaluator_factory=None, stage_ops=None) -> Dict[str, Any]:
    """One condition end to end: multi-chain TMCMC on the GPU + summaries + the run directory.
    `condition_key` is an entry of tmcmc202601_b200/data/heine_conditions.json (e.g. "Dysbiotic_HOBIC")."""
    cd = workloads.load_conditions()[condition_key]
    legacy = condition_key.endswith("legacy_1k30")
    sk = dict(workloads.PRODUCTION_SOLVER, **(solver_overrides or {}))
    phi_init = 0.02 if legacy else cd["phi_init"]
    sigma = cd["sigma_scalar"] if legacy else cd["sigma_obs"]
    theta_base = np.full(20, 1.5)  # MAIN:1937-1939
    theta_base[cd["locked"]] = 0.0
    active = list(cd["active_indices"])
    bounds = [tuple(b) for b in cd["bounds"]]
    evs: List[Any] = []

    def make_evaluator(theta_linearization=None):
        if evaluator_factory is not None:
            ev = evaluator_factory(dict(sk, phi_init=phi_init), active, theta_base, cd, sigma, cov_rel, theta_linearization)
        else:
            ev = LogLikelihoodEvaluator(dict(sk, phi_init=phi_init), [0, 1, 2, 3, 4], active, theta_base, cd["data"], cd["idx_sparse"], sigma,
                                        cov_rel=cov_rel, theta_linearization=theta_linearization, device=device)
        evs.append(ev)
        return ev

    tag = f"{cd['condition']}_{cd['cultivation']}"
    t0 = time.time()
    chains, logL, MAP, conv, diag = tmcmc.run_multi_chain_TMCMC(
        tag, make_evaluator, bounds, theta_base, active, n_particles=n_particles, n_stages=n_stages, n_chains=n_chains, seed=seed,
        min_delta_beta=min_delta_beta, max_delta_beta=max_delta_beta, n_mutation_steps=n_mutation_steps, device_mutation=device_mutation,
        stage_ops=stage_ops)
    elapsed = time.time() - t0
    sig_arr = np.asarray(sigma, dtype=np.float64)
    results = summarize_run(chains, logL, diag, theta_base, active, bounds, conv, elapsed, hdi_credibility,
                            sigma_obs=(sig_arr if sig_arr.size > 1 else float(sig_arr.reshape(-1)[0])))
    pi = np.asarray(phi_init, dtype=np.float64)
    days = list(cd.get("days", []))
    config = {
        "condition": cd["condition"], "cultivation": cd["cultivation"], "dt": sk["dt"], "maxtimestep": sk["maxtimestep"],
        "c_const": sk["c_const"], "Kp1": sk["Kp1"], "K_hill": sk["K_hill"], "n_hill": sk["n_hill"], "alpha_const": sk["alpha_const"],
        "phi_init": pi.tolist() if pi.size > 1 else float(pi.reshape(-1)[0]), "phi_init_is_array": bool(pi.size > 1),
        "use_exp_init": bool(pi.size > 1), "day_scale": None,
        "sigma_obs": "replicate_matrix" if sig_arr.size > 1 else float(sig_arr.reshape(-1)[0]), "replicate_sigma": bool(sig_arr.size > 1),
        "sigma_scale": 1.0, "cov_rel": cov_rel, "prior_min": None, "prior_max": None, "prior_decay_max": None,
        "n_particles": n_particles, "n_stages": n_stages, "n_chains": n_chains, "seed": seed, "hdi_credibility": hdi_credibility,
        "compute_evidence": True,
        "metadata": {"condition": cd["condition"], "cultivation": cd["cultivation"], "days": days, "n_timepoints": len(days),
                     "idx_sparse": [int(i) for i in cd["idx_sparse"]], "backend": "tmcmc202601_b200 (libhamilton, sm_100a)",
                     "device_mutation": bool(device_mutation)},
    }
    fm = None
    if with_fit_metrics:
        fm = fit_metrics_for(results["theta_MAP_full"], results["theta_mean_full"], dict(sk, phi_init=phi_init), cd["data"], cd["idx_sparse"],
                             device=device)
    write_run_directory(output_dir, results, config, cd["data"], cd["idx_sparse"], np.asarray(days), fit_metrics=fm)
    for ev in evs:
        try:
            ev.close()
        except Exception:
            pass
    results["output_dir"] = str(output_dir)
    results["fit_metrics"] = fm
    return results
