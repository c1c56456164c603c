"""BiofilmNewtonSolver5S -- same constructor and methods as the reference class
(tmcmc/program2602/improved_5species_jit.py:440-581), backed by hmx_trajectory_batch on a B200.

    solver = BiofilmNewtonSolver5S(dt=1e-4, maxtimestep=2500, c_const=25.0, alpha_const=0.0,
                                   phi_init=0.02, Kp1=1e-4, K_hill=0.05, n_hill=4.0)
    t_arr, g_arr = solver.run_deterministic(theta)        # (T+1,), (T+1, 12)
    g_batch = solver.run_deterministic_batch(theta[N,20]) # (N, T+1, 12)

There is no NumPy/numba fallback: `use_numba` is accepted for signature compatibility and ignored.
"""

from __future__ import annotations

import numpy as np

from . import _abi
from .engine import HamiltonEngine


class BiofilmNewtonSolver5S:
    THETA_NAMES = [
        "a11", "a12", "a22", "b1", "b2", "a33", "a34", "a44", "b3", "b4",
        "a13", "a14", "a23", "a24", "a55", "b5", "a15", "a25", "a35", "a45",
    ]

    def __init__(self, dt=1e-5, maxtimestep=500, eps=1e-6, Kp1=1e-4, eta_vec=None, eta_phi_vec=None,
                 c_const=100.0, alpha_const=100.0, alpha_schedule=None, phi_init=0.2, active_species=None,
                 use_numba=True, max_newton_iter=50, K_hill=0.0, n_hill=2.0, device=0):
        self.dt, self.maxtimestep, self.eps, self.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
        self.c_const, self.alpha_const = float(c_const), float(alpha_const)
        self.alpha_schedule = alpha_schedule  # stored and never read by the 5-species reference solver either (S5:474, 488)
        self.phi_init = _abi.broadcast_phi_init(phi_init)
        self.max_newton_iter = int(max_newton_iter)
        self.K_hill, self.n_hill = float(K_hill), float(n_hill)
        self.active_species = [0, 1, 2, 3, 4] if active_species is None else list(active_species)
        self.Eta_vec = np.ones(5) if eta_vec is None else np.asarray(eta_vec, dtype=float)
        self.Eta_phi_vec = np.ones(5) if eta_phi_vec is None else np.asarray(eta_phi_vec, dtype=float)
        self.use_numba = False  # nothing is JIT-compiled here; kept for attribute compatibility
        self.device = int(device)
        self._engine = None

    def solver_kwargs(self) -> dict:
        return dict(dt=self.dt, maxtimestep=self.maxtimestep, eps=self.eps, Kp1=self.Kp1, eta_vec=self.Eta_vec,
                    eta_phi_vec=self.Eta_phi_vec, c_const=self.c_const, alpha_const=self.alpha_const,
                    max_newton_iter=self.max_newton_iter, K_hill=self.K_hill, n_hill=self.n_hill,
                    active_species=self.active_species)

    def _eng(self) -> HamiltonEngine:
        if self._engine is None:
            # a trajectory-only context: one dummy condition carrying the initial state, no observation rows
            cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, self.phi_init, tsm_variance=False)
            self._engine = HamiltonEngine(_abi.make_config([cond], **self.solver_kwargs()), device=self.device)
        return self._engine

    def theta_to_matrices(self, theta):
        """theta (20,) -> symmetric A (5,5), b (5,).  Index map of S5:513-559 (pure indexing, host side)."""
        theta = np.asarray(theta)
        A = np.zeros((5, 5), dtype=theta.dtype)
        b = np.zeros(5, dtype=theta.dtype)
        pairs = {0: (0, 0), 1: (0, 1), 2: (1, 1), 5: (2, 2), 6: (2, 3), 7: (3, 3), 10: (0, 2), 11: (0, 3),
                 12: (1, 2), 13: (1, 3), 14: (4, 4), 16: (0, 4), 17: (1, 4), 18: (2, 4), 19: (3, 4)}
        for k, (i, j) in pairs.items():
            A[i, j] = theta[k]
            A[j, i] = theta[k]
        for k, i in {3: 0, 4: 1, 8: 2, 9: 3, 15: 4}.items():
            b[i] = theta[k]
        return A, b

    def time_array(self) -> np.ndarray:
        t_arr = np.empty(self.maxtimestep + 1)
        t_arr[0] = 0.0
        t_arr[1:] = (np.arange(self.maxtimestep) + 1) * self.dt  # tt = (step + 1) * dt, S5:860
        return t_arr

    def run_deterministic_batch(self, theta, rows=None, return_status=False):
        """g[N, T+1, 12] for theta[N, 20]; with `rows` (ascending time levels) only g[N, len(rows), 12] leaves the GPU."""
        theta = np.ascontiguousarray(np.atleast_2d(np.real(theta)), dtype=np.float64)
        if rows is None:
            g, st, _ = self._eng().trajectory_batch(theta[:, :20])
        else:
            g, st, _ = self._eng().trajectory_rows(theta[:, :20], rows)
        return (g, st) if return_status else g

    def run_deterministic(self, theta):
        g = self.run_deterministic_batch(np.asarray(theta, dtype=np.float64)[None, :])[0]
        return self.time_array(), g

    solve = run_deterministic  # alias used by the TSM wrapper (S5:646)


class BiofilmNewtonSolver:
    """The legacy 4-species solver -- same constructor and methods as the reference class
    (tmcmc/program2602/improved1207_paper_jit.py:715-1053), backed by hmx_legacy_trajectory_rows on a B200.

        solver = BiofilmNewtonSolver(dt=1e-5, maxtimestep=500, c_const=100.0, alpha_const=100.0, phi_init=0.2,
                                     active_species=[0, 1], alpha_schedule={"alpha_before": 0.0, "alpha_after": 100.0, "switch_frac": 0.5})
        t_arr, g_arr = solver.run_deterministic(theta14)          # (T+1,), (T+1, 10)

    State g = [phi1..phi4, phi0, psi1..psi4, gamma]; theta = a11,a12,a22,b1,b2, a33,a34,a44,b3,b4, a13,a14,a23,a24.
    `alpha_schedule` is honoured (the reference's numba integrator ignores it, S4:1024; only its pure-Python fallback
    evaluates `alpha(t)`): the step that produces t = n dt uses `self.alpha(t)`.  `max_newton_iter` is accepted and the
    time loop uses 50 like the reference's numba loop (S4:685) unless `honour_max_newton_iter=True`."""

    THETA_NAMES = ["a11", "a12", "a22", "b1", "b2", "a33", "a34", "a44", "b3", "b4", "a13", "a14", "a23", "a24"]

    def __init__(self, dt=1e-5, maxtimestep=500, eps=1e-6, Kp1=1e-4, eta_vec=None, eta_phi_vec=None, c_const=100.0,
                 alpha_const=100.0, alpha_schedule=None, phi_init=0.2, active_species=None, use_numba=True, max_newton_iter=50,
                 device=0, honour_max_newton_iter=False):
        self.dt, self.maxtimestep, self.eps, self.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
        self.c_const, self.alpha_const = float(c_const), float(alpha_const)
        self.alpha_schedule = alpha_schedule
        self.phi_init = float(phi_init)
        self.max_newton_iter = int(max_newton_iter)
        self.active_species = [0, 1, 2, 3] if active_species is None else list(active_species)
        self.active_mask = np.zeros(4, dtype=np.int64)
        for i in self.active_species:
            if 0 <= i < 4:
                self.active_mask[i] = 1
        self.Eta_vec = np.ones(4) if eta_vec is None else np.asarray(eta_vec, dtype=float)
        self.Eta_phi_vec = np.ones(4) if eta_phi_vec is None else np.asarray(eta_phi_vec, dtype=float)
        self.use_numba = False  # nothing is JIT-compiled here; kept for attribute compatibility
        self.device = int(device)
        self._iter_cap = self.max_newton_iter if honour_max_newton_iter else 50
        self._engine = None

    def c(self, t: float) -> float:
        return float(self.c_const)

    def alpha(self, t: float) -> float:
        """Piecewise-constant antibiotic strength at time t: the decision rules of S4:783-822 (switch_time, then
        switch_frac of the total duration, then switch_step on the time grid); a malformed schedule means alpha_const."""
        sched = self.alpha_schedule
        if not sched:
            return self.alpha_const
        try:
            before = float(sched.get("alpha_before", self.alpha_const))
            after = float(sched.get("alpha_after", self.alpha_const))
            if "switch_time" in sched:
                return after if float(t) >= float(sched["switch_time"]) else before
            if "switch_frac" in sched:
                frac = min(max(float(sched["switch_frac"]), 0.0), 1.0)
                return after if float(t) >= frac * float(self.maxtimestep) * self.dt else before
            if "switch_step" in sched:
                step = int(np.floor(float(t) / self.dt + 1e-12))
                return after if step >= int(sched["switch_step"]) else before
        except Exception:
            return self.alpha_const
        return self.alpha_const

    def _switch_level(self):
        """(switch_level, alpha_before, alpha_after) equivalent to `alpha((n) dt)` for the levels n = 1..T."""
        if not self.alpha_schedule:
            return -1, self.alpha_const, self.alpha_const
        levels = np.arange(1, self.maxtimestep + 1)
        a = np.array([self.alpha(n * self.dt) for n in levels])
        first, last = a[0], a[-1]
        change = np.nonzero(a != first)[0]
        if change.size == 0:
            return 0, first, first  # constant over the run (possibly != alpha_const)
        k = int(change[0])
        if not np.all(a[k:] == last):
            raise ValueError("alpha_schedule does not reduce to one switch on the time grid")
        return int(levels[k]), float(first), float(last)

    def theta_to_matrices(self, theta):
        """theta (14,) -> symmetric A (4,4), b (4,): the index map of S4:824-855."""
        theta = np.asarray(theta)
        A = np.zeros((4, 4), dtype=theta.dtype)
        b = np.zeros(4, dtype=theta.dtype)
        for k, (i, j) in {0: (0, 0), 1: (0, 1), 2: (1, 1), 5: (2, 2), 6: (2, 3), 7: (3, 3), 10: (0, 2), 11: (0, 3), 12: (1, 2), 13: (1, 3)}.items():
            A[i, j] = theta[k]
            A[j, i] = theta[k]
        for k, i in {3: 0, 4: 1, 8: 2, 9: 3}.items():
            b[i] = theta[k]
        return A, b

    def get_initial_state(self) -> np.ndarray:
        g = np.zeros(10)
        for i in range(4):
            g[i] = self.phi_init if i in self.active_species else 0.0
        g[4] = 1.0 - np.sum(g[0:4])
        for i in range(4):
            g[5 + i] = 0.999 if i in self.active_species else 0.0
        return g

    def _legacy_config(self) -> "_abi.HmxLegacyConfig":
        lvl, before, after = self._switch_level()
        return _abi.make_legacy_config(
            n_species=4, dt=self.dt, maxtimestep=self.maxtimestep, eps=self.eps, Kp1=self.Kp1, eta_vec=self.Eta_vec,
            eta_phi_vec=self.Eta_phi_vec, c_const=self.c_const, alpha_const=self.alpha_const, phi_init=self.phi_init,
            active_species=[i for i in self.active_species if 0 <= i < 4], max_newton_iter=self._iter_cap,
            alpha_switch_level=lvl, alpha_before=before, alpha_after=after)

    def _eng(self) -> HamiltonEngine:
        if self._engine is None:
            # any context will do: the legacy call carries its own solver settings; this one has no observation rows
            cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.2, tsm_variance=False)
            self._engine = HamiltonEngine(_abi.make_config([cond]), device=self.device)
        return self._engine

    def time_array(self) -> np.ndarray:
        t_arr = np.empty(self.maxtimestep + 1)
        t_arr[0] = 0.0
        t_arr[1:] = (np.arange(self.maxtimestep) + 1) * self.dt
        return t_arr

    def run_deterministic_batch(self, theta, rows=None, return_status=False):
        """g[N, T+1, 10] for theta[N, 14]; with `rows` only those time levels leave the GPU."""
        theta = np.ascontiguousarray(np.atleast_2d(np.real(theta)), dtype=np.float64)
        g, st, _ = self._eng().legacy_trajectory_rows(self._legacy_config(), theta[:, :14], rows)
        return (g, st) if return_status else g

    def run_deterministic(self, theta, show_progress: bool = False):
        g = self.run_deterministic_batch(np.asarray(theta, dtype=np.float64)[None, :])[0]
        return self.time_array(), g

    solve = run_deterministic
