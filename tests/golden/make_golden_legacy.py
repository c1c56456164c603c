#!/usr/bin/env python
"""Golden vectors for the legacy 4-species solver from the LIVE reference.

Run in the BUILD CONTAINER only (needs /root/reference + numba), ~1 minute:

    python tests/golden/make_golden_legacy.py

* numba path: `BiofilmNewtonSolver(**kw).run_deterministic(theta)` (S4:1005-1028) for all species active, the M1 / M2
  sub-models (active_species [0,1] / [2,3]), a three-species mask, and the production-style settings (dt=1e-4, c=25, alpha=0).
* alpha_schedule: the reference honours it only in its pure-Python fallback (`use_numba=False`), whose Jacobian is a
  finite difference (S4:962-975) and whose Newton tolerance is the constant eps (S4:999) -- a different iterate
  sequence that converges to the same implicit-Euler solution within the Newton tolerance.  Those trajectories are
  stored for a LOOSE comparison (the tight pin of the scheduled solve is the C oracle, hamilton_oracle4.c).

Output (committed): tests/golden/legacy4_kat.npz.  The reference is only *called*; nothing is copied from it.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, "/root/reference/tmcmc/program2602")

from oracle import oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent
ROWS = np.array([0, 1, 2, 5, 10, 50, 100, 250, 400, 499, 500])


def main():
    from improved1207_paper_jit import BiofilmNewtonSolver, get_theta_true

    t0 = time.time()
    th0 = get_theta_true()
    rng = np.random.default_rng(20260303)
    cases = []
    for name, kw in [
        ("all_active_default", dict()),
        ("M1_species01", dict(active_species=[0, 1])),
        ("M2_species23", dict(active_species=[2, 3])),
        ("three_species_alpha3", dict(active_species=[0, 2, 3], alpha_const=3.0)),
        ("production_style", dict(dt=1e-4, maxtimestep=500, c_const=25.0, alpha_const=0.0, phi_init=0.05)),
        ("eta_nonunit", dict(eta_vec=[1.0, 1.5, 0.8, 1.2], eta_phi_vec=[1.0, 0.9, 1.1, 1.3])),
    ]:
        for j, th in enumerate([th0, th0 * rng.uniform(0.5, 1.5, 14), rng.uniform(0.0, 3.0, 14)]):
            cases.append((f"{name}#{j}", kw, th))
    names, kws, thetas, g_rows = [], [], [], []
    worst = 0.0
    for name, kw, th in cases:
        _, g = BiofilmNewtonSolver(**kw).run_deterministic(th)
        _, go, its, _, st = orc.run_deterministic4(orc.make_cfg4(**kw), th)
        e = float((np.abs(go - g) / np.maximum(np.abs(g), 1e-300))[g != 0].max())
        worst = max(worst, e)
        print(f"{name:32s} its={its:5d} st={st} oracle-vs-live max rel {e:.2e}", flush=True)
        names.append(name)
        kws.append(json.dumps(kw))
        thetas.append(th)
        g_rows.append(g[ROWS])
    # alpha schedule through the reference's pure-Python fallback (slow: finite-difference Jacobian in Python)
    sched_kw = dict(maxtimestep=120, alpha_const=100.0, alpha_schedule={"alpha_before": 0.0, "alpha_after": 100.0, "switch_step": 60})
    s = BiofilmNewtonSolver(use_numba=False, **sched_kw)
    _, g_sched = s.run_deterministic(th0)
    alphas = np.array([s.alpha((n + 1) * s.dt) for n in range(s.maxtimestep)])
    s_frac = BiofilmNewtonSolver(use_numba=False, maxtimestep=120, alpha_schedule={"alpha_before": 5.0, "alpha_after": 50.0, "switch_frac": 0.25})
    alphas_frac = np.array([s_frac.alpha((n + 1) * s_frac.dt) for n in range(120)])
    s_time = BiofilmNewtonSolver(use_numba=False, maxtimestep=120, alpha_schedule={"alpha_after": 7.0, "switch_time": 3.3e-4})
    alphas_time = np.array([s_time.alpha((n + 1) * s_time.dt) for n in range(120)])
    np.savez_compressed(
        OUT / "legacy4_kat.npz", names=np.array(names), kwargs=np.array(kws), theta=np.array(thetas), rows=ROWS,
        g_rows=np.array(g_rows), oracle_vs_live_max_rel=worst,
        sched_kwargs=json.dumps(sched_kw), sched_theta=th0, sched_g=g_sched, sched_alpha_per_step=alphas,
        sched_frac_alpha_per_step=alphas_frac, sched_time_alpha_per_step=alphas_time,
    )
    print(json.dumps({"cases": len(names), "oracle_vs_live_max_rel": worst, "wall_s": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()
