"""LogLikelihoodEvaluator & friends -- the reference's evaluator surface (data_5species/core/evaluator.py)
on top of hmx_loglik_batch.

The constructor takes the reference's arguments (EV:381-398).  `__call__(theta_sub) -> float` is the
reference's single-particle interface; `batch(theta[N, n]) -> logL[N]` is the entry
`evaluate_particles_parallel` dispatches to (one kernel launch for the whole population).  Failed solves
return -1e20, never raise (EV:655,667,753).

What stays on the host, exactly as in the reference:
  * theta_sub -> theta_full:  full[active_indices[i]] = theta_sub[i]  (EV:635-637), including the
    column shift this produces when locks are present and 20-dimensional particles are passed
    (SURVEY Appendix B-1).  `strict_mapping=True` opts into the intended mapping instead.
  * the np.allclose(theta, theta_linearization) shortcut of BiofilmTSM_Analytical.solve_tsm
    (demo_analytical_tsm_with_linearization_jit.py:364-401): such particles are evaluated without the
    TSM variance (sig2 = 0).  (The reference additionally mixes in the *previous* call's sensitivities
    through the stale `_last_x1` attribute, EV:723-740; that order-dependent artefact is not reproduced.)
The linear-ROM branch (TSMA:409-456) is out of scope: it stays off in the production runs
(log_tmcmc.txt:59,108), so `enable_linearization(True)` makes `compute_ROM_error` report +inf and the
TMCMC driver switches it off again, like production.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional

import numpy as np

from . import _abi
from .engine import HamiltonEngine
from .utils import LikelihoodHealthCounter, TimingStats, timed

_shared_engines: Dict[int, HamiltonEngine] = {}


def build_species_sigma(sigma_global: float, n_species: int = 5, vd_species_idx: int = 2, vd_factor: float = 2.0) -> np.ndarray:
    """Per-species sigma with V. dispar inflated (EV:100-131)."""
    s = np.full(n_species, sigma_global, dtype=np.float64)
    if vd_species_idx < n_species:
        s[vd_species_idx] *= vd_factor
    return s


def build_replicate_sigma(replicate_csv: str, condition: str, cultivation: str, days: list, species_map: dict,
                          n_species: int = 5, min_sigma: float = 0.005) -> np.ndarray:
    """Per-(day, species) sigma = IQR / 1.35 of the replicate percentages, floored at min_sigma, in fraction
    space (EV:134-187).  Setup-time host work; needs the reference's fig3_species_distribution_replicates.csv."""
    import pandas as pd

    df = pd.read_csv(replicate_csv)
    dc = df[(df["condition"] == condition) & (df["cultivation"] == cultivation)]
    sigma = np.full((len(days), n_species), min_sigma, dtype=np.float64)
    for i, day in enumerate(days):
        for name, idx in species_map.items():
            vals = dc[(dc["species"] == name) & (dc["day"] == day)]["distribution_pct"].values
            if len(vals) >= 3:
                iqr = np.percentile(vals, 75) - np.percentile(vals, 25)
                sigma[i, idx] = max(iqr / 1.35, min_sigma * 100.0) / 100.0
    return sigma


def build_likelihood_weights(n_obs: int, n_species: int, pg_species_idx: int = 4, n_late: int = 2,
                             lambda_pg: float = 5.0, lambda_late: float = 3.0) -> np.ndarray:
    """Weight matrix: column pg_species_idx x lambda_pg, last n_late rows x lambda_late (EV:190-234)."""
    W = np.ones((n_obs, n_species), dtype=np.float64)
    if pg_species_idx < n_species:
        W[:, pg_species_idx] *= lambda_pg
    if n_late > 0:
        W[-n_late:, :] *= lambda_late
    return W


def log_likelihood_sparse(mu, sig, data, sigma_obs, rho: float = 0.0, health=None, weights=None, device: int = 0) -> float:
    """EV:237-371 (rho = 0 branch) evaluated on the GPU through hmx_log_likelihood_sparse."""
    if abs(rho) > 1e-9:
        raise NotImplementedError("rho != 0 is never used by the reference (rho=0.0 hard-coded at MAIN:2221)")
    data = np.asarray(data, dtype=np.float64)
    if data.shape[1] != 5:
        raise ValueError("5-species data expected")
    eng = _shared_engines.get(device)
    if eng is None:
        cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.2, tsm_variance=False)
        eng = _shared_engines[device] = HamiltonEngine(_abi.make_config([cond]), device=device)
    out = eng.log_likelihood_sparse(mu, sig, data, sigma_obs, weights)
    if health is not None:
        var_raw = np.asarray(sig, dtype=float) + _abi.broadcast_sigma(sigma_obs, data.shape[0]) ** 2
        health["n_var_raw_nonfinite"] = int(health.get("n_var_raw_nonfinite", 0)) + int((~np.isfinite(var_raw)).sum())
        health["n_var_raw_negative"] = int(health.get("n_var_raw_negative", 0)) + int((np.isfinite(var_raw) & (var_raw < 0)).sum())
        health["n_var_total_clipped"] = int(health.get("n_var_total_clipped", 0)) + int(((~np.isfinite(var_raw)) | (var_raw <= 1e-20)).sum())
    return out


class LogLikelihoodEvaluator:
    """Drop-in for core.evaluator.LogLikelihoodEvaluator (5-species path)."""

    def __init__(self, solver_kwargs: Dict[str, Any], active_species: List[int], active_indices: List[int],
                 theta_base: np.ndarray, data: np.ndarray, idx_sparse: np.ndarray, sigma_obs, cov_rel: float,
                 rho: float = 0.0, theta_linearization: Optional[np.ndarray] = None, use_analytical: bool = True,
                 paper_mode: bool = False, debug_logger=None, use_absolute_volume: bool = False,
                 weights: Optional[np.ndarray] = None, device: int = 0, tsm_variance: bool = True,
                 strict_mapping: bool = False, engine_factory=None, latency_kernel: bool = False, sig2_guard_all_rows: bool = False):
        if abs(rho) > 1e-9:
            raise NotImplementedError("rho != 0 is never used by the reference (MAIN:2221)")
        self.active_species = list(active_species)
        if not ((len(self.active_species) == 5) or (max(self.active_species) >= 4)):
            raise NotImplementedError("only the 5-species solver is on the hot path (SURVEY 2.1 #2)")
        self.active_indices = list(active_indices)
        self.theta_base = np.asarray(theta_base, dtype=np.float64).copy()
        self.data = np.asarray(data, dtype=np.float64)
        self.idx_sparse = np.asarray(idx_sparse)
        self.sigma_obs = sigma_obs
        self.cov_rel = float(cov_rel)
        self.rho = rho
        self.n_species = len(self.active_species)
        self.solver_kwargs = dict(solver_kwargs)
        self.debug_logger = debug_logger
        self.use_absolute_volume = bool(use_absolute_volume)
        self.weights = weights
        self.strict_mapping = bool(strict_mapping)
        self.device = int(device)

        self.call_count = 0
        self.fom_call_count = 0
        self.timing = TimingStats()
        self.health = LikelihoodHealthCounter()
        self.theta_history: list = []
        self.logL_history: list = []
        self.ve_prior = None

        if np.any(self.idx_sparse < 0) or np.any(self.idx_sparse > int(self.solver_kwargs.get("maxtimestep", 500))):
            raise IndexError(f"Invalid idx_sparse {self.idx_sparse}: valid range [0, maxtimestep] (EV:682-697)")

        kw = dict(self.solver_kwargs)
        phi_init = kw.pop("phi_init", 0.2)
        # BiofilmTSM.active_idx is fixed at construction time from the active_indices passed here (S4:1075)
        tsm_active = [int(i) for i in self.active_indices if int(i) < _abi.NTHETA]
        conds = [
            _abi.make_cond(self.data, self.idx_sparse, sigma_obs, phi_init, cov_rel=self.cov_rel,
                           active_theta_indices=tsm_active, weights=weights,
                           use_absolute_volume=self.use_absolute_volume, tsm_variance=tsm_variance),
            # condition 1: same constants without the TSM variance (np.allclose shortcut, TSMA:364-401)
            _abi.make_cond(self.data, self.idx_sparse, sigma_obs, phi_init, cov_rel=self.cov_rel,
                           active_theta_indices=tsm_active, weights=weights,
                           use_absolute_volume=self.use_absolute_volume, tsm_variance=False),
        ]
        # engine_factory is the test seam (CPU suite injects an oracle-backed stand-in); the product default is the GPU
        make_engine = engine_factory or (lambda cfg: HamiltonEngine(cfg, device=self.device))
        self._engine = make_engine(_abi.make_config(conds, active_species=self.active_species, **kw))
        if sig2_guard_all_rows and hasattr(self._engine, "set_sig2_guard"):
            # EV:657-667 scans sig2 at all T + 1 rows; the default checks the state at every row and sig2 at the observation rows
            self._engine.set_sig2_guard(True)
        if latency_kernel and hasattr(self._engine, "set_latency_kernel"):
            self._engine.set_latency_kernel(-1)  # K1L for production-sized launches (see include/hmx.h)

        if theta_linearization is None:
            theta_linearization = self.theta_base.copy()
        self._theta_linearization = np.asarray(theta_linearization, dtype=np.float64).copy()
        self._linearization_enabled = False

    # -- linearization hooks used by run_TMCMC (TM:557-562,1436-1441,1561) ---------------------------------
    def update_linearization_point(self, theta_new_full: np.ndarray):
        self._theta_linearization = np.asarray(theta_new_full, dtype=np.float64).copy()
        self.theta_history, self.logL_history, self.call_count = [], [], 0  # EV:508-511

    def enable_linearization(self, enable: bool = True):
        self._linearization_enabled = bool(enable)

    def get_linearization_point(self) -> np.ndarray:
        return self._theta_linearization.copy()

    def compute_ROM_error(self, theta_full: np.ndarray) -> float:
        """With linearization off the TSM mean IS the full solve, so the observable-space error is 0 (EV:534-627).
        The linear ROM is not implemented: requesting it reports +inf so the driver keeps it off."""
        self.fom_call_count += 1
        return float("inf") if self._linearization_enabled else 0.0

    def get_health(self) -> Dict[str, int]:
        return self.health.to_dict()

    def get_MAP(self):
        if len(self.logL_history) == 0:
            raise ValueError("No evaluations yet")
        i = int(np.argmax(self.logL_history))
        return self.theta_history[i], self.logL_history[i]

    # -- evaluation ------------------------------------------------------------------------------------------
    def _full_theta(self, theta_sub: np.ndarray) -> np.ndarray:
        """theta_sub[N, n] -> theta_full[N, len(theta_base)] (EV:635-637)."""
        theta_sub = np.atleast_2d(np.asarray(theta_sub, dtype=np.float64))
        N = theta_sub.shape[0]
        full = np.tile(self.theta_base[None, :], (N, 1))
        if self.strict_mapping and theta_sub.shape[1] == full.shape[1]:
            for idx in self.active_indices:
                full[:, idx] = theta_sub[:, idx]
        else:
            for i, idx in enumerate(self.active_indices):
                full[:, idx] = theta_sub[:, i]
        return full

    def batch(self, theta_sub: np.ndarray, record_history: bool = False) -> np.ndarray:
        theta_sub = np.atleast_2d(np.asarray(theta_sub, dtype=np.float64))
        N = theta_sub.shape[0]
        self.call_count += N
        self.health.n_calls += N
        full = self._full_theta(theta_sub)
        ode = np.ascontiguousarray(full[:, :20])
        lin = self._theta_linearization[:20]
        at_lin = np.all(np.abs(ode - lin[None, :]) <= (1e-8 + 1e-5 * np.abs(lin[None, :])), axis=1)  # np.allclose
        cond_id = at_lin.astype(np.int32)
        with timed(self.timing, "tsm.solve_tsm"):
            logL, status, _ = self._engine.loglik_batch(ode, cond_id, return_status=True)
        self.health.n_output_nonfinite += int(((status & _abi.ST_NONFINITE) != 0).sum())
        self.health.n_max_newton_iter += int(((status & _abi.ST_MAXITER) != 0).sum())
        self.health.n_singular_direction += int(((status & _abi.ST_SINGULAR) != 0).sum())
        if self.ve_prior is not None:
            logL = logL + np.array([self.ve_prior.log_prior(t) for t in theta_sub])
        if record_history:
            for t, l in zip(theta_sub, logL):
                self.theta_history.append(t.copy())
                self.logL_history.append(float(l))
        return logL

    def __call__(self, theta_sub: np.ndarray) -> float:
        return float(self.batch(np.asarray(theta_sub, dtype=np.float64)[None, :], record_history=True)[0])

    def value_and_grad(self, theta_sub: np.ndarray):
        """(logL[N], dlogL/dtheta_sub[N, n]) for theta_sub[N, n]: what the reference's NUTS / HMC engine obtains from
        `jax.value_and_grad(log_likelihood)` (main/tmcmc_nuts_engine.py:202), here by forward sensitivities on the GPU
        (hmx_loglik_grad_batch).  The gradient of the full 20-vector is pulled back through the theta_sub -> theta_full
        scatter of `_full_theta` (a column that lands on several full entries collects all of them; a column that reaches
        no ODE parameter gets 0).  Failed solves: logL = -1e20, zero gradient."""
        theta_sub = np.atleast_2d(np.asarray(theta_sub, dtype=np.float64))
        N, n = theta_sub.shape
        self.call_count += N
        self.health.n_calls += N
        full = self._full_theta(theta_sub)
        ode = np.ascontiguousarray(full[:, :20])
        lin = self._theta_linearization[:20]
        at_lin = np.all(np.abs(ode - lin[None, :]) <= (1e-8 + 1e-5 * np.abs(lin[None, :])), axis=1)
        with timed(self.timing, "tsm.solve_tsm"):
            logL, g_full, status = self._engine.loglik_grad_batch(ode, at_lin.astype(np.int32))
        self.health.n_output_nonfinite += int(((status & _abi.ST_NONFINITE) != 0).sum())
        grad = np.zeros((N, n))
        amap = self.mutation_map(n)
        for j, a in enumerate(amap):
            if 0 <= a < _abi.NTHETA:
                grad[:, j] += g_full[:, a]
        if self.ve_prior is not None:
            logL = logL + np.array([self.ve_prior.log_prior(t) for t in theta_sub])
            for ve_idx, pos in self.ve_prior._sub_positions.items():
                grad[:, pos] -= (theta_sub[:, pos] - self.ve_prior.mu[ve_idx]) / self.ve_prior.sigma[ve_idx] ** 2
        return logL, grad

    def mutation_map(self, d: int) -> List[int]:
        """Position of theta_sub[j] in the full vector, exactly as _full_theta scatters it (-1: column unused)."""
        if self.strict_mapping and d == self.theta_base.size:
            act = set(int(i) for i in self.active_indices)
            return [j if j in act else -1 for j in range(d)]
        if len(self.active_indices) > d:
            raise ValueError("theta_sub has fewer columns than active_indices")
        return [int(i) for i in self.active_indices]

    def mutate_device(self, theta_res: np.ndarray, logL_res: np.ndarray, factor: np.ndarray, lows, highs, K: int,
                      beta_next: float, seed: int, sequence: int = 0, index_offset: int = 0):
        """K-step MH mutation of the resampled population on the GPU (hmx_mutate; reference TM:849-962 with Philox
        instead of PCG64).  Returns (theta[N,d], logL[N], stats)."""
        if self.ve_prior is not None:
            raise NotImplementedError("the viscoelastic prior is host-side arithmetic; use the host mutation path")
        theta_res = np.ascontiguousarray(theta_res, dtype=np.float64)
        N, d = theta_res.shape
        mut = _abi.make_mutation(factor, lows, highs, K, beta_next, seed, sequence=sequence, index_offset=index_offset,
                                 theta_base=self.theta_base[:_abi.NTHETA], active_map=self.mutation_map(d),
                                 theta_lin=self._theta_linearization[:_abi.NTHETA], cond_default=0, cond_at_lin=1)
        with timed(self.timing, "tsm.solve_tsm"):
            theta, logL, st = self._engine.mutate(mut, theta_res, logL_res)
        self.call_count += int(st["n_candidates"])
        self.health.n_calls += N * int(K)
        self.health.n_output_nonfinite += int(st["n_nonfinite"])
        self.health.n_max_newton_iter += int(st["n_maxiter"])
        self.health.n_singular_direction += int(st["n_singular"])
        return theta, logL, st

    def close(self):
        self._engine.close()


class ViscoelasticPrior:
    """Gaussian penalty on the two SLS parameters theta[20], theta[21] (EV:796-836); host-side scalar work."""

    def __init__(self, active_indices, mu=None, sigma=None):
        self.ve_indices = [20, 21]
        self.active_indices = list(active_indices)
        self.mu = mu or {20: 1.0, 21: 3.0}
        self.sigma = sigma or {20: 0.5, 21: 1.0}
        self._sub_positions = {i: self.active_indices.index(i) for i in self.ve_indices if i in self.active_indices}

    def log_prior(self, theta_sub) -> float:
        lp = 0.0
        for ve_idx, pos in self._sub_positions.items():
            lp -= 0.5 * ((theta_sub[pos] - self.mu[ve_idx]) / self.sigma[ve_idx]) ** 2
        return lp
