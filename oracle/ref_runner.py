"""The reference AS SHIPPED (oracle/_ref, see make_ref.py) on a product-side batch.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: importable from tests/ and from bench.py's cpu_baseline / --impl reference
legs.  It drives the reference's own public API and stock code path -- `core.evaluator.LogLikelihoodEvaluator`
(EV:381-778: numba ODE solve + the Python complex-step TSM loop at all 2500 rows) mapped over particles by
`core.tmcmc.evaluate_particles_parallel` (TM:247-312, one task per particle on a ProcessPoolExecutor) -- with
none of this repository's kernels or oracle arithmetic on that path.
"""

from __future__ import annotations

import os
import time
from pathlib import Path

import numpy as np

from . import make_ref

_api = None


def available() -> bool:
    return make_ref.available()


def _load():
    global _api
    if _api is None:
        if not available():
            raise RuntimeError("oracle/_ref is absent: run `python oracle/make_ref.py` where /root/reference exists")
        _api = make_ref.import_reference(make_ref.OUT)
    return _api


class ReferenceArm:
    """One reference evaluator per condition of an HmxConfig, built like MAIN:2181-2239 builds it."""

    def __init__(self, hcfg, n_jobs=None):
        LogLikelihoodEvaluator, self._parallel = _load()
        self.n_jobs = n_jobs or os.cpu_count() or 1
        self.evaluators, self.active = [], []
        for k in range(hcfg.n_cond):
            hc = hcfg.cond[k]
            n = hc.n_obs
            active = [j for j in range(20) if (hc.active_theta_mask >> j) & 1]
            kw = dict(dt=hcfg.dt, maxtimestep=hcfg.n_steps, c_const=hcfg.c_const, alpha_const=hcfg.alpha_const,
                      phi_init=list(hc.phi_init), Kp1=hcfg.Kp1, K_hill=hcfg.K_hill, n_hill=hcfg.n_hill)
            w = np.array(hc.weights[: n * 5]).reshape(n, 5)
            ev = LogLikelihoodEvaluator(
                solver_kwargs=kw, active_species=[0, 1, 2, 3, 4], active_indices=active, theta_base=np.zeros(20),
                data=np.array(hc.data[: n * 5]).reshape(n, 5), idx_sparse=np.array(hc.idx_sparse[:n]),
                sigma_obs=np.array(hc.sigma_obs[: n * 5]).reshape(n, 5), cov_rel=hc.cov_rel, rho=0.0,
                theta_linearization=np.full(20, 1e6),  # never np.allclose to a particle: the full TSM path, as in production
                paper_mode=False, use_absolute_volume=bool(hc.use_absolute_volume), weights=None if np.all(w == 1.0) else w,
            )
            self.evaluators.append(ev)
            self.active.append(active)

    def warm_up(self, theta_full, cond_id):
        """One call per condition in THIS process: numba compiles here, forked workers inherit the compiled code."""
        t0 = time.perf_counter()
        for c in range(len(self.evaluators)):
            sel = np.nonzero(cond_id == c)[0]
            if sel.size:
                self.evaluators[c](theta_full[sel[0]][self.active[c]])
        return time.perf_counter() - t0

    def evaluate(self, theta_full, cond_id):
        """logL[n] through the reference's evaluate_particles_parallel, one call per condition present."""
        theta_full = np.asarray(theta_full, dtype=np.float64)
        out = np.empty(theta_full.shape[0])
        for c in range(len(self.evaluators)):
            sel = np.nonzero(cond_id == c)[0]
            if sel.size == 0:
                continue
            sub = theta_full[sel][:, self.active[c]]  # EV:635-637 scatters theta_sub[i] to full[active[i]]
            out[sel] = self._parallel(self.evaluators[c], sub, n_jobs=min(self.n_jobs, sel.size) if sel.size > 1 else 1)
        return out

    # ---- the reference WITHOUT its Python sensitivity loop -------------------------------------------
    def _ode_only_one(self, c, theta_full):
        """The reference's numba solve + its own log_likelihood_sparse with sig = 0 (EV:568, 698-745 with the variance off)."""
        import sys

        ev = self.evaluators[c]
        lls = sys.modules[type(ev).__module__].log_likelihood_sparse
        _, x0 = ev.solver.run_deterministic(np.asarray(theta_full, dtype=np.float64))
        if not np.all(np.isfinite(x0)):
            return -1e20
        idx = np.asarray(ev.idx_sparse)
        ns = (x0.shape[1] - 2) // 2
        other = x0[idx, -1][:, None] if ev.use_absolute_volume else x0[idx][:, ns + 1:2 * ns + 1]
        mu = x0[idx][:, :ns] * other
        return float(lls(mu, np.zeros_like(mu), ev.data, ev.sigma_obs, rho=0.0, weights=ev.weights))

    def evaluate_ode_only(self, theta_full, cond_id):
        """logL[n] with the TSM variance off, particles mapped over forked worker processes (the compiled numba code of the
        parent is inherited, so call warm_up_ode_only first)."""
        import multiprocessing as mp

        global _ODE_ARM
        theta_full = np.asarray(theta_full, dtype=np.float64)
        _ODE_ARM = self
        tasks = [(int(c), theta_full[i]) for i, c in enumerate(cond_id)]
        n = min(self.n_jobs, len(tasks))
        if n <= 1:
            return np.array([_ode_task(t) for t in tasks])
        with mp.get_context("fork").Pool(n) as pool:
            return np.array(pool.map(_ode_task, tasks, chunksize=max(1, len(tasks) // (4 * n))))

    def warm_up_ode_only(self, theta_full, cond_id):
        t0 = time.perf_counter()
        for c in range(len(self.evaluators)):
            sel = np.nonzero(cond_id == c)[0]
            if sel.size:
                self._ode_only_one(c, theta_full[sel[0]])
        return time.perf_counter() - t0


_ODE_ARM = None


def _ode_task(task):
    c, th = task
    return _ODE_ARM._ode_only_one(c, th)
