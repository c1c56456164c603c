#!/usr/bin/env python
"""Generate the committed golden fixtures from the LIVE reference (keisuke58/Tmcmc202601).

Run in the BUILD CONTAINER only (needs /root/reference + numba):

    python tests/golden/make_golden.py            # ~5 minutes (the reference evaluator is 5 s / call)

Outputs (committed):
    tmcmc202601_b200/data/heine_conditions.json   observation constants + prior bounds from the reference loader
    tests/golden/solver_kat.npz                   numba run_deterministic rows for seeded theta / configs
    tests/golden/evaluator_kat.npz                LogLikelihoodEvaluator.__call__ values (TSM variance ON) + sig=0 values
    tests/golden/stage_kat.npz                    systematic_resample / run_TMCMC toy-likelihood traces
    tests/golden/MANIFEST.json                    versions + max deviation of the C oracle from each live vector

The reference is only *called* here (public API); nothing is copied from it.
"""

from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402
from oracle import reference_live as rl  # noqa: E402

OUT = Path(__file__).resolve().parent
PKG_DATA = ROOT / "tmcmc202601_b200" / "data"

THETA_MAP_ROUNDED = np.array(  # the rounded production MAP
    [1.974, 2.024, 1.037, 1.833, 0.3156, 0.0477, 1.554, 4.089, 2.023, 0.3458, 2.290, 0.899, 2.910,
     -0.388, 0.067, 0.413, 0.131, 1.099, 26.64, 3.968]
)

BASE_KW = dict(dt=1e-4, maxtimestep=2500, c_const=25.0, alpha_const=0.0, phi_init=0.02, Kp1=1e-4,
               K_hill=0.05, n_hill=4.0)

ROWS = sorted(set([0, 1, 2, 3, 5, 10, 20, 50, 100, 113, 200, 339, 500, 679, 900, 1131, 1400, 1696, 2000,
                   2375, 2499, 2500]))


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def main():
    t_start = time.time()
    m = rl.bootstrap()
    import numba
    import scipy

    manifest = {
        "generated_by": "tests/golden/make_golden.py",
        "versions": {"numpy": np.__version__, "numba": numba.__version__, "scipy": scipy.__version__},
        "oracle_vs_live": {},
    }

    # ------------------------------------------------------------------ conditions
    conds = {}
    for c, k in rl.CONDITIONS:
        d = rl.load_condition(c, k)
        conds[f"{c}_{k}"] = d
    legacy_dir = rl.REF / "_runs" / "Dysbiotic_HOBIC_K0.05_n4.0_1k30"
    legacy_cfg = json.loads((legacy_dir / "config.json").read_text())
    legacy = dict(
        condition="Dysbiotic", cultivation="HOBIC", data=np.load(legacy_dir / "data.npy"),
        idx_sparse=np.load(legacy_dir / "idx_sparse.npy") if (legacy_dir / "idx_sparse.npy").exists()
        else np.array([113, 339, 679, 1131, 1696, 2375]),
        sigma_scalar=float(legacy_cfg.get("sigma_obs", 0.23209876543209873)),
        phi_init=np.full(5, 0.02), bounds=conds["Dysbiotic_HOBIC"]["bounds"], locked=[], days=[1, 3, 6, 10, 15, 21],
    )
    legacy["sigma_obs"] = np.full((legacy["data"].shape[0], 5), legacy["sigma_scalar"])
    # The stored run predates the current model_config/prior_bounds.json (its samples reach 30 in a23/a35 and
    # 20 in a45); the run directory does not record its bounds, so they are inferred from the extent of its
    # 2000 posterior samples, snapped outward to the grid of bound values the reference uses.
    grid = np.array([-3.0, -1.0, -0.5, 0.0, 0.5, 1.0, 2.5, 3.0, 5.0, 10.0, 15.0, 20.0, 30.0])
    s_leg = np.load(legacy_dir / "samples.npy")
    legacy["bounds"] = [(float(grid[grid <= lo + 1e-9].max()), float(grid[grid >= hi - 1e-9].min()))
                        for lo, hi in zip(s_leg.min(0), s_leg.max(0))]
    legacy["bounds_note"] = "inferred from the sample extent of _runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30/samples.npy"
    conds["Dysbiotic_HOBIC_legacy_1k30"] = legacy

    def ser(d):
        return {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}

    PKG_DATA.mkdir(parents=True, exist_ok=True)
    (PKG_DATA / "heine_conditions.json").write_text(
        json.dumps({"source": "reference loader outputs (MAIN:1644-1867, EV:134-187, nishioka_model.py:152-252), "
                              "CLI defaults: use_exp_init, normalize, replicate sigma, dt=1e-4, maxtimestep=2500",
                    "conditions": {k: ser(v) for k, v in conds.items()}}, indent=1)
    )

    # ------------------------------------------------------------------ solver KATs
    S = m["BiofilmNewtonSolver5S"]
    cases = []
    rng = np.random.default_rng(20260101)
    dh = conds["Dysbiotic_HOBIC"]
    bounds_dh = np.array(dh["bounds"])

    def add_case(name, kw, theta):
        cases.append((name, dict(kw), np.asarray(theta, float)))

    add_case("map_phi0.02_n4", BASE_KW, THETA_MAP_ROUNDED)
    for n in (1.0, 2.0, 3.0, 2.5):
        add_case(f"map_phi0.02_n{n}", dict(BASE_KW, n_hill=n), THETA_MAP_ROUNDED)
    add_case("map_phi0.02_nohill", dict(BASE_KW, K_hill=0.0), THETA_MAP_ROUNDED)
    for key, cd in conds.items():
        if key.endswith("legacy_1k30"):
            continue
        b = np.array(cd["bounds"])
        kw = dict(BASE_KW, phi_init=cd["phi_init"].tolist())
        for j in range(3):
            add_case(f"{key}_prior{j}", kw, rng.uniform(b[:, 0], b[:, 1]))
    kw500 = dict(dt=1e-5, maxtimestep=500, c_const=100.0, alpha_const=100.0, phi_init=0.2, Kp1=1e-4,
                 K_hill=0.0, n_hill=2.0)  # deeponet/surrogate_tmcmc.py:353-365 configuration
    add_case("t500_map", kw500, THETA_MAP_ROUNDED)
    for j in range(2):
        add_case(f"t500_prior{j}", kw500, rng.uniform(bounds_dh[:, 0], bounds_dh[:, 1]))
    add_case("t500_hill_alpha", dict(kw500, K_hill=0.05, n_hill=2.0), rng.uniform(bounds_dh[:, 0], bounds_dh[:, 1]))

    names, kws, thetas, rows_l, g_rows = [], [], [], [], []
    worst = 0.0
    for name, kw, theta in cases:
        _, g_ref = S(**kw).run_deterministic(theta)
        T = kw["maxtimestep"]
        rows = np.array([r for r in ROWS if r <= T] + ([T] if T not in ROWS else []))
        rows = np.unique(rows)
        _, g_orc, its, tr, st = orc.run_deterministic(orc.make_solver_cfg(**kw), theta)
        e = relerr(g_orc, g_ref)
        worst = max(worst, e)
        print(f"solver {name:36s} its={its:6d} trials={tr:6d} st={st} oracle-vs-live max rel {e:.2e}")
        names.append(name)
        kws.append(json.dumps(kw))
        thetas.append(theta)
        pad = np.full(len(ROWS) + 1, -1)
        pad[: len(rows)] = rows
        rows_l.append(pad)
        gr = np.full((len(ROWS) + 1, 12), np.nan)
        gr[: len(rows)] = g_ref[rows]
        g_rows.append(gr)
    manifest["oracle_vs_live"]["solver_full_trajectory_max_rel"] = worst
    np.savez_compressed(OUT / "solver_kat.npz", names=np.array(names), kwargs=np.array(kws),
                        theta=np.array(thetas), rows=np.array(rows_l), g_rows=np.array(g_rows))

    # ------------------------------------------------------------------ evaluator KATs (TSM variance ON)
    ev_names, ev_theta_sub, ev_theta_full, ev_ll, ev_ll_det, ev_keys = [], [], [], [], [], []
    worst = worst_det = 0.0
    for key, cd in conds.items():
        kw = dict(BASE_KW, phi_init=np.asarray(cd["phi_init"]).tolist())
        sig = cd["sigma_scalar"] if key.endswith("legacy_1k30") else cd["sigma_obs"]
        ev, theta_base, active = rl.make_evaluator(cd, kw, sigma_obs=sig)
        b = np.array(cd["bounds"])
        thetas_sub = [rng.uniform(b[:, 0], b[:, 1]) for _ in range(2)]
        if not cd["locked"]:
            thetas_sub.append(THETA_MAP_ROUNDED)
        cc = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], sig, cov_rel=0.005, active_theta_indices=active)
        cc_det = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], sig, cov_rel=0.005,
                                   active_theta_indices=active, tsm_variance=False)
        cfg = orc.make_solver_cfg(**kw)
        for j, ts in enumerate(thetas_sub):
            t0 = time.time()
            ll = float(ev(ts))
            # the evaluator's own theta_sub -> theta_full map, EV:635-637
            full = theta_base.copy()
            for i, idx in enumerate(active):
                full[idx] = ts[i]
            _, g_ref = ev.solver.run_deterministic(full)
            mu = g_ref[cd["idx_sparse"], 0:5] * g_ref[cd["idx_sparse"], 6:11]
            ll_det = float(m["log_likelihood_sparse"](mu, np.zeros_like(mu), cd["data"], sig, rho=0.0))
            ll_o = orc.evaluate(cfg, cc, full)
            ll_o_det = orc.evaluate(cfg, cc_det, full)
            e, ed = abs(ll_o - ll) / abs(ll), abs(ll_o_det - ll_det) / abs(ll_det)
            worst, worst_det = max(worst, e), max(worst_det, ed)
            print(f"evaluator {key:28s} #{j} live={ll:.15g} oracle={ll_o:.15g} rel={e:.2e} | det live={ll_det:.15g} "
                  f"rel={ed:.2e}  ({time.time()-t0:.1f}s)")
            ev_names.append(f"{key}#{j}")
            ev_keys.append(key)
            ev_theta_sub.append(ts)
            ev_theta_full.append(full)
            ev_ll.append(ll)
            ev_ll_det.append(ll_det)
    manifest["oracle_vs_live"]["evaluator_logL_max_rel"] = worst
    manifest["oracle_vs_live"]["evaluator_logL_det_only_max_rel"] = worst_det
    np.savez_compressed(OUT / "evaluator_kat.npz", names=np.array(ev_names), cond=np.array(ev_keys),
                        theta_sub=np.array(ev_theta_sub), theta_full=np.array(ev_theta_full),
                        logL=np.array(ev_ll), logL_det=np.array(ev_ll_det),
                        solver_kwargs=json.dumps({k: v for k, v in BASE_KW.items() if k != "phi_init"}))

    # ------------------------------------------------------------------ stage arithmetic KATs
    tm = m["tmcmc"]
    rs_w, rs_u, rs_idx = [], [], []
    for seed, N in ((1, 7), (2, 64), (3, 1000), (4, 1000), (5, 4096), (6, 30000)):
        r = np.random.default_rng(seed)
        ll = r.normal(size=N) * (3.0 if seed % 2 else 30.0)
        w = np.exp(ll - ll.max())
        w /= w.sum()
        r2 = np.random.default_rng(1000 + seed)
        u = np.random.default_rng(1000 + seed).random()
        idx = tm.systematic_resample(w, r2)
        rs_w.append(w)
        rs_u.append(u)
        rs_idx.append(np.asarray(idx, np.int64))
        assert np.array_equal(orc.resample(w, u), idx), "oracle resample differs from live reference"
    manifest["oracle_vs_live"]["resample_bit_exact_cases"] = len(rs_w)

    # toy-likelihood run of the live run_TMCMC: pins beta schedule, evidence, acceptance and the RNG stream usage
    def toy_loglik(theta):
        mu = np.array([0.3, -0.2, 0.8])
        s = np.array([0.1, 0.25, 0.05])
        return float(-0.5 * np.sum(((theta - mu) / s) ** 2))

    toy_bounds = [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)]
    res = tm.run_TMCMC(toy_loglik, toy_bounds, n_particles=300, n_stages=30, seed=1234, n_jobs=1,
                       min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=3)
    ss = res.stage_summary
    np.savez_compressed(
        OUT / "stage_kat.npz",
        resample_n=len(rs_w), resample_u=np.array(rs_u),
        **{f"resample_w{k}": a for k, a in enumerate(rs_w)}, **{f"resample_idx{k}": a for k, a in enumerate(rs_idx)},
        toy_bounds=np.array(toy_bounds), toy_seed=1234, toy_n_particles=300, toy_n_stages=30, toy_n_mutation_steps=3,
        toy_beta_schedule=np.array(res.beta_schedule), toy_samples=res.samples, toy_logL=res.logL_values,
        toy_theta_MAP=res.theta_MAP, toy_log_evidence=res.log_evidence,
        toy_acc_rate=np.array(res.acc_rate_history),
        toy_delta_beta=np.array([r["delta_beta"] for r in ss]),
        toy_ess=np.array([np.nan if r["ess"] is None else r["ess"] for r in ss]),
    )
    manifest["toy_tmcmc"] = {"stages": len(ss), "final_beta": res.beta_schedule[-1], "log_evidence": res.log_evidence}

    # ------------------------------------------------------------------ posterior summaries of the stored production run
    # _runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30: N=1000, 2 chains, seed 42, 63 h on 8 CPU processes (results_summary.json)
    s_ref = np.load(legacy_dir / "samples.npy")
    l_ref = np.load(legacy_dir / "logL.npy")
    np.savez_compressed(
        OUT / "posterior_dh_1k30.npz", n=s_ref.shape[0], mean=s_ref.mean(0), std=s_ref.std(0, ddof=1),
        q=np.quantile(s_ref, [0.05, 0.25, 0.5, 0.75, 0.95], axis=0), logL_mean=l_ref.mean(), logL_std=l_ref.std(),
        logL_max=l_ref.max(), logL_min=l_ref.min(), MAP=s_ref[np.argmax(l_ref)],
        chain_means=np.stack([s_ref[:1000].mean(0), s_ref[1000:].mean(0)]),
    )
    # the reference's own credible intervals for those samples (written by its run-directory code, MAIN:378-456)
    import csv as _csv

    with open(legacy_dir / "parameter_summary.csv") as fh:
        rows = list(_csv.DictReader(fh))
    np.savez_compressed(
        OUT / "posterior_dh_1k30_samples.npz", samples=s_ref.astype(np.float64),
        hdi_lower=np.array([float(r["hdi_lower"]) for r in rows]), hdi_upper=np.array([float(r["hdi_upper"]) for r in rows]),
        et_lower=np.array([float(r["et_lower"]) for r in rows]), et_upper=np.array([float(r["et_upper"]) for r in rows]),
        median=np.array([float(r["median"]) for r in rows]), mean=np.array([float(r["mean"]) for r in rows]),
        credibility=float(legacy_cfg.get("hdi_credibility", 0.95)),
    )
    manifest["production_run"] = {"samples": list(s_ref.shape), "logL_max": float(l_ref.max()), "elapsed_s": 226864.8}

    manifest["wall_s"] = round(time.time() - t_start, 1)
    (OUT / "MANIFEST.json").write_text(json.dumps(manifest, indent=1))
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
