#!/usr/bin/env python
"""Golden vectors for the TSM variance from the LIVE reference.

Run in the BUILD CONTAINER only (needs /root/reference + numba), ~2 minutes:

    python tests/golden/make_golden_tsm.py

For every condition (the four loader outputs + the legacy production configuration) and a few theta it calls
the reference's `BiofilmTSM_Analytical.solve_tsm(theta_full)` (-> BiofilmTSM.solve_tsm, S4:1077-1160: complex-step
sensitivities at ALL 2500 rows) and stores what the evaluator consumes (EV:645-745):

    x0[idx_sparse], sig2[idx_sparse], tsm._last_x1[idx_sparse] (active columns), tsm._last_var_act,
    the evaluator's logL, and -- for the whole-trajectory guard of EV:657-667 -- max |sig2| and the finiteness of
    sig2 over all T+1 rows, plus x1 / sig2 at a few extra rows off the observation grid.

Output (committed): tests/golden/tsm_kat.npz.  The reference is only *called*; nothing is copied from it.
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
EXTRA_ROWS = np.array([1, 2, 50, 1000, 2500])

THETA_MAP_ROUNDED = np.array(
    [1.974, 2.024, 1.037, 1.833, 0.3156, 0.0477, 1.554, 4.089, 2.023, 0.3458, 2.290, 0.899, 2.910,
     -0.388, 0.067, 0.413, 0.131, 1.099, 26.64, 3.968]
)
BASE_KW = dict(dt=1e-4, maxtimestep=2500, c_const=25.0, alpha_const=0.0, phi_init=0.02, Kp1=1e-4,
               K_hill=0.05, n_hill=4.0)


def main():
    t_start = time.time()
    rl.bootstrap()
    conds = json.loads((ROOT / "tmcmc202601_b200" / "data" / "heine_conditions.json").read_text())["conditions"]
    rng = np.random.default_rng(20260202)
    out = dict(names=[], cond=[], theta_full=[], active=[], idx_sparse=[], n_obs=[], logL=[], x0=[], sig2=[], x1=[],
               var_act=[], sig2_all_max=[], sig2_all_finite=[], x0_extra=[], sig2_extra=[], x1_extra=[])
    worst = dict(x1=0.0, sig2=0.0)
    for key, cd in conds.items():
        cd = {k: (np.asarray(v) if isinstance(v, list) and k not in ("bounds", "locked", "days") else v) for k, v in cd.items()}
        kw = dict(BASE_KW, phi_init=np.asarray(cd["phi_init"], float).tolist())
        legacy = key.endswith("legacy_1k30")
        sig = cd["sigma_scalar"] if legacy else np.asarray(cd["sigma_obs"], float)
        cd["data"] = np.asarray(cd["data"], float)
        cd["idx_sparse"] = np.asarray(cd["idx_sparse"], int)
        ev, theta_base, active = rl.make_evaluator(cd, kw, sigma_obs=sig)
        b = np.array(cd["bounds"])
        thetas_sub = [rng.uniform(b[:, 0], b[:, 1])]
        if not cd["locked"]:
            thetas_sub.append(THETA_MAP_ROUNDED)
        cfg = orc.make_solver_cfg(**kw)
        idx = cd["idx_sparse"]
        for j, ts in enumerate(thetas_sub):
            t0 = time.time()
            ll = float(ev(ts))
            full = theta_base.copy()
            for i, a in enumerate(active):
                full[a] = ts[i]
            t_arr, x0, sig2 = ev.tsm.solve_tsm(full)
            x1 = np.asarray(ev.tsm._last_x1)
            var_act = np.asarray(ev.tsm._last_var_act)
            act = np.asarray(ev.tsm._last_active_idx, dtype=int)
            assert list(act) == [a for a in active if a < 20]
            # oracle pin: orc.sensitivity_row against the live x1 at the observation rows
            for r in idx:
                t = max(int(r), 1)
                x1o = orc.sensitivity_row(cfg, full, x0[t], x0[t - 1], sum(1 << int(k) for k in act))
                d = np.abs(x1o[:, act] - x1[t]).max() / max(np.abs(x1[t]).max(), 1e-300)
                worst["x1"] = max(worst["x1"], float(d))
                s2o = (x1o[:, act] ** 2 * var_act[None, :]).sum(axis=1) + 1e-12
                worst["sig2"] = max(worst["sig2"], float((np.abs(s2o - sig2[t]) / np.abs(sig2[t])).max()))
            n_obs = len(idx)
            pad = 8 - n_obs
            x1_rows = np.zeros((8, 12, 20))
            x1_rows[:n_obs][:, :, act] = x1[idx]
            x1_extra = np.zeros((len(EXTRA_ROWS), 12, 20))
            x1_extra[:, :, act] = x1[EXTRA_ROWS]
            out["names"].append(f"{key}#{j}")
            out["cond"].append(key)
            out["theta_full"].append(full)
            out["active"].append(np.pad(act, (0, 20 - len(act)), constant_values=-1))
            out["idx_sparse"].append(np.pad(idx, (0, pad), constant_values=-1))
            out["n_obs"].append(n_obs)
            out["logL"].append(ll)
            out["x0"].append(np.pad(x0[idx], ((0, pad), (0, 0))))
            out["sig2"].append(np.pad(sig2[idx], ((0, pad), (0, 0))))
            out["x1"].append(x1_rows)
            out["var_act"].append(np.pad(var_act, (0, 20 - len(var_act))))
            out["sig2_all_max"].append(float(np.abs(sig2).max()))
            out["sig2_all_finite"].append(bool(np.all(np.isfinite(sig2))))
            out["x0_extra"].append(x0[EXTRA_ROWS])
            out["sig2_extra"].append(sig2[EXTRA_ROWS])
            out["x1_extra"].append(x1_extra)
            print(f"{key:30s} #{j} logL={ll:.15g} max|sig2| all rows={np.abs(sig2).max():.3e} "
                  f"oracle-vs-live x1 {worst['x1']:.2e} sig2 {worst['sig2']:.2e} ({time.time()-t0:.1f}s)", flush=True)
    np.savez_compressed(
        OUT / "tsm_kat.npz", **{k: np.array(v) for k, v in out.items()}, extra_rows=EXTRA_ROWS,
        solver_kwargs=json.dumps({k: v for k, v in BASE_KW.items() if k != "phi_init"}),
        oracle_vs_live_x1_max_rel=worst["x1"], oracle_vs_live_sig2_max_rel=worst["sig2"],
    )
    print(json.dumps({"cases": len(out["names"]), "oracle_vs_live": worst, "wall_s": round(time.time() - t_start, 1)}))


if __name__ == "__main__":
    main()
