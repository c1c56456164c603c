"""BiofilmTSM -- the reference's TSM-ROM wrapper (tmcmc/program2602/improved1207_paper_jit.py:1055-1160 and its
`BiofilmTSM_Analytical` subclass with linearization disabled, the production setting) on a B200.

    tsm = BiofilmTSM(BiofilmNewtonSolver5S(**solver_kwargs), active_theta_indices=active, cov_rel=0.005)
    t_arr, x0, sigma2 = tsm.solve_tsm(theta20)       # (T+1,), (T+1, 12), (T+1, 12): every time level
    tsm._last_x1[idx_sparse]                          # via solve_tsm_rows(theta, idx_sparse)

`solve_tsm` returns the variance at EVERY time level like the reference (hmx_solve_tsm_batch: one residual and
n_active direction solves per level); the likelihood path does not need it (hmx_loglik_batch evaluates the
observation rows only).  The reference's `_last_x1` side channel holds 2501 x 12 x n_active doubles; here the
sensitivities are available at requested rows through `solve_tsm_rows`."""

from __future__ import annotations

import numpy as np

from . import _abi
from .engine import HamiltonEngine


class BiofilmTSM:
    def __init__(self, solver, active_theta_indices=None, cov_rel: float = 0.005, use_complex_step: bool = True, device: int = 0):
        self.solver = solver
        self.cov_rel = float(cov_rel)
        self.use_complex_step = bool(use_complex_step)  # kept for signature compatibility: dQ/dtheta is exact in closed form
        if active_theta_indices is None:
            active_theta_indices = np.arange(_abi.NTHETA, dtype=int)
        self.active_idx = np.array([int(k) for k in active_theta_indices if int(k) < _abi.NTHETA], dtype=int)
        self.device = int(device)
        self._engines = {}

    def _engine(self, rows) -> HamiltonEngine:
        key = tuple(int(r) for r in rows)
        if key not in self._engines:
            cond = _abi.make_cond(np.zeros((len(key), 5)), list(key), 1.0, self.solver.phi_init, cov_rel=self.cov_rel,
                                  active_theta_indices=self.active_idx.tolist(), tsm_variance=True)
            self._engines[key] = HamiltonEngine(_abi.make_config([cond], **self.solver.solver_kwargs()), device=self.device)
        return self._engines[key]

    def solve_tsm_batch(self, theta):
        """x0[N,T+1,12], sigma2[N,T+1,12], status[N] for theta[N,20]."""
        theta = np.ascontiguousarray(np.atleast_2d(np.real(theta)), dtype=np.float64)[:, : _abi.NTHETA]
        return self._engine(()).solve_tsm_batch(theta)

    def solve_tsm(self, theta):
        x0, s2, _ = self.solve_tsm_batch(np.asarray(theta, dtype=np.float64)[None, :])
        self._last_theta = np.real(np.asarray(theta, dtype=np.float64))
        self._last_active_idx = self.active_idx
        self._last_var_act = (self.cov_rel * self._last_theta) ** 2
        self._last_var_act = self._last_var_act[self.active_idx]
        self._last_t_arr = self.solver.time_array()
        return self._last_t_arr, x0[0], s2[0]

    def solve_tsm_rows(self, theta, rows):
        """(x0[rows], sigma2[rows], x1[rows][:, :, active_idx]) for one theta: the rows of `solve_tsm`'s outputs and of the
        reference's `_last_x1` (strictly ascending `rows`, at most 8 per call)."""
        rows = [int(r) for r in rows]
        x0, s2, x1, _, _ = self._engine(rows).tsm_obs_rows(np.asarray(theta, dtype=np.float64)[None, : _abi.NTHETA])
        n = len(rows)
        return x0[0, :n], s2[0, :n], x1[0, :n][:, :, self.active_idx]

    def close(self):
        for e in self._engines.values():
            e.close()
        self._engines = {}
