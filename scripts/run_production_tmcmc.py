"""Full TMCMC runs through the reference-shaped API on the GPU (BASELINE.json configs[0..2]).

  python scripts/run_production_tmcmc.py [--particles 1000] [--chains 2] [--out gpurun_out/tmcmc_runs.json]

* `Dysbiotic_HOBIC_legacy_1k30` reproduces the stored production run of the reference
  (_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30: N=1000, 2 chains, seed 42, K=5 mutation steps, 6 observation rows,
  phi_init=0.02, scalar sigma; 226,865 s on 8 CPU processes) and compares the posterior with its summaries
  (tests/golden/posterior_dh_1k30.npz).
* the four current-default conditions (exp-init, 5 rows, replicate sigma) are the `--batch all` workload.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from tmcmc202601_b200 import tmcmc, workloads  # noqa: E402
from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator  # noqa: E402


def run_condition(key, n_particles, n_chains, seed=42, n_mutation_steps=5, n_stages=30, device_mutation=False, latency_kernel=False):
    cd = workloads.load_conditions()[key]
    legacy = key.endswith("legacy_1k30")
    sk = dict(workloads.PRODUCTION_SOLVER, phi_init=0.02 if legacy else cd["phi_init"])
    sigma = cd["sigma_scalar"] if legacy else cd["sigma_obs"]
    theta_base = np.full(20, 1.5)  # PRIOR_BOUNDS_DEFAULT mean (MAIN:1937-1939); locked -> 0
    theta_base[cd["locked"]] = 0.0
    active = cd["active_indices"]
    bounds = [tuple(b) for b in cd["bounds"]]
    evs = []

    def make_evaluator(theta_linearization=None):
        ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], active, theta_base, cd["data"], cd["idx_sparse"], sigma, cov_rel=0.005,
                                    theta_linearization=theta_linearization, latency_kernel=latency_kernel)
        evs.append(ev)
        return ev

    tag = f"{cd['condition']}_{cd['cultivation']}"
    t0 = time.perf_counter()
    chains, logL, MAP, conv, diag = tmcmc.run_multi_chain_TMCMC(
        tag, make_evaluator, bounds, theta_base, active, n_particles=n_particles, n_stages=n_stages, n_chains=n_chains,
        seed=seed, min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=n_mutation_steps, device_mutation=device_mutation)
    wall = time.perf_counter() - t0
    samples = np.concatenate(chains, axis=0)
    ll = np.concatenate(logL)
    summ = tmcmc.compute_MAP_with_uncertainty(samples, ll)
    n_evals = sum(ev.health.n_calls for ev in evs)
    gpu_s = sum(ev.timing.get_s("tsm.solve_tsm") for ev in evs)
    out = {
        "condition": key, "device_mutation": bool(device_mutation), "latency_kernel": bool(latency_kernel), "n_particles": n_particles, "n_chains": n_chains, "seed": seed, "n_mutation_steps": n_mutation_steps,
        "wall_s": wall, "likelihood_evaluations": int(n_evals), "likelihood_wall_s": gpu_s, "evals_per_s_in_run": n_evals / gpu_s,
        "converged": [bool(c) for c in conv], "stages": [len(s) for s in diag["stage_summaries"]],
        "beta_schedules": diag["beta_schedules"], "acc_rates": diag["acc_rate_histories"], "log_evidence": diag["log_evidence"],
        "logL_max": float(ll.max()), "logL_mean": float(ll.mean()), "MAP": summ["MAP"].tolist(), "mean": summ["mean"].tolist(),
        "std": summ["std"].tolist(), "ci_lower": summ["ci_lower"].tolist(), "ci_upper": summ["ci_upper"].tolist(),
        "Rhat_max": float(np.nanmax(diag["Rhat_reference"])) if n_chains > 1 else None,
    }
    for ev in evs:
        ev.close()
    return out, samples, ll


def compare_with_reference(samples, ll):
    ref = np.load(ROOT / "tests" / "golden" / "posterior_dh_1k30.npz")
    n_eff = 200.0  # conservative effective sample size for 2 x 1000 correlated TMCMC particles on both sides
    z = np.abs(samples.mean(0) - ref["mean"]) / np.sqrt((samples.std(0, ddof=1) ** 2 + ref["std"] ** 2) / n_eff)
    ratio = samples.std(0, ddof=1) / ref["std"]
    q = np.quantile(samples, [0.05, 0.5, 0.95], axis=0)
    dq = np.abs(q - ref["q"][[0, 2, 4]]) / ref["std"]
    return {"z_mean_max": float(z.max()), "std_ratio_min": float(ratio.min()), "std_ratio_max": float(ratio.max()),
            "quantile_shift_over_std_max": float(dq.max()), "logL_max_ref": float(ref["logL_max"]), "logL_mean_ref": float(ref["logL_mean"]),
            "reference_elapsed_s": 226864.8}


def run_merged(keys, n_particles, n_chains, seed=42, n_mutation_steps=5, n_stages=30, device_mutation=False, merge_launches=True):
    """All conditions x chains concurrently, likelihood launches merged (tmcmc202601_b200.batch)."""
    from tmcmc202601_b200 import batch

    conds = workloads.load_conditions()
    jobs = []
    for key in keys:
        cd = conds[key]
        legacy = key.endswith("legacy_1k30")  # the stored production run: phi_init = 0.02, scalar sigma, 6 observation rows
        base = np.full(20, 1.5)
        base[cd["locked"]] = 0.0
        jobs.append(dict(model_tag=f"{cd['condition']}_{cd['cultivation']}", data=cd["data"], idx_sparse=cd["idx_sparse"],
                         sigma_obs=cd["sigma_scalar"] if legacy else cd["sigma_obs"], phi_init=0.02 if legacy else cd["phi_init"],
                         prior_bounds=cd["bounds"], theta_base=base, active_indices=cd["active_indices"]))
    out, stats = batch.run_batch_TMCMC(jobs, workloads.PRODUCTION_SOLVER, n_particles=n_particles, n_stages=n_stages,
                                       n_chains=n_chains, seed=seed, n_mutation_steps=n_mutation_steps, device_mutation=device_mutation,
                                       merge_launches=merge_launches)
    summary = {"mode": "merged" if merge_launches else "own stream per chain", "device_mutation": bool(device_mutation), "conditions": list(keys), "n_particles": n_particles, "n_chains": n_chains, **stats,
               "per_condition": [{"condition": k, "converged": o["converged"], "stages": [len(r.stage_summary) for r in o["results"]],
                                  "logL_max": float(max(l.max() for l in o["logL"]))} for k, o in zip(keys, out)]}
    return summary, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--merged", action="store_true", help="run the four current-default conditions concurrently")
    ap.add_argument("--particles", type=int, default=1000)
    ap.add_argument("--chains", type=int, default=2)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "tmcmc_runs.json"))
    ap.add_argument("--only", default=None)
    ap.add_argument("--device-mutation", action="store_true", help="proposals / accept-reject on the GPU (Philox stream)")
    ap.add_argument("--latency-kernel", action="store_true", help="K1L (eight lanes per solve) for production-sized launches")
    args = ap.parse_args()
    keys = ["Dysbiotic_HOBIC_legacy_1k30"] + workloads.CONDITION_KEYS
    if args.only:
        keys = [k for k in keys if k in args.only.split(",")]
    results = []
    if args.merged:
        summary, _ = run_merged(workloads.CONDITION_KEYS, args.particles, args.chains, device_mutation=args.device_mutation)
        print(json.dumps(summary), flush=True)
        results.append(summary)
        keys = []
    for k in keys:
        out, samples, ll = run_condition(k, args.particles, args.chains, device_mutation=args.device_mutation, latency_kernel=args.latency_kernel)
        if k.endswith("legacy_1k30") and args.particles >= 500:
            out["vs_reference_run"] = compare_with_reference(samples, ll)
        print(json.dumps({kk: out[kk] for kk in ("condition", "wall_s", "likelihood_evaluations", "stages", "logL_max", "Rhat_max") if kk in out}
                         | ({"vs_reference_run": out["vs_reference_run"]} if "vs_reference_run" in out else {})), flush=True)
        results.append(out)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
