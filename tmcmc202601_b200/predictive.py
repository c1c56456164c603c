"""Batched trajectory consumers: the callers of the reference that solve the ODE for many
parameter vectors and keep (a down-sampled view of) the trajectories -- posterior-predictive bands and the
surrogate's training set -- on top of hmx_trajectory_rows (one launch per batch, only the selected time levels
leave the GPU).

  * posterior_predictive_bands     plot_manager.plot_posterior_predictive_band / FEM/posterior_ci_0d.py: quantile
                                   bands of phi_i psi_i over posterior draws
  * generate_dataset               deeponet/generate_training_data.py:144-333 with the same sampling stream
                                   (prior / MAP-centred / posterior mixtures, failure rejection, np.linspace
                                   down-sampling), 10^4-5x10^4 solves at T = 500 in seconds instead of hours
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np

from .solver import BiofilmNewtonSolver5S


def posterior_predictive_bands(samples: np.ndarray, solver_kwargs: Dict[str, Any], n_draws: Optional[int] = None, n_time_out: int = 251,
                               quantiles: Sequence[float] = (0.05, 0.5, 0.95), seed: int = 0, device: int = 0, solver=None) -> Dict[str, Any]:
    """Quantile bands of the observable phi_i psi_i and of the total living volume over posterior draws.

    samples[n, >=20] posterior particles; n_draws of them (all when None) are drawn without replacement.
    Returns {"t", "rows", "phibar_q"[len(q), n_time, 5], "total_q"[len(q), n_time], "mean"[n_time, 5], "n_used", "n_failed"}."""
    samples = np.asarray(samples, dtype=np.float64)
    rng = np.random.default_rng(seed)
    if n_draws is not None and n_draws < len(samples):
        samples = samples[rng.choice(len(samples), size=int(n_draws), replace=False)]
    solver = solver or BiofilmNewtonSolver5S(**solver_kwargs, device=device)
    T = solver.maxtimestep
    rows = np.unique(np.linspace(0, T, min(int(n_time_out), T + 1), dtype=int))
    g, st = solver.run_deterministic_batch(samples[:, :20], rows=rows, return_status=True)
    ok = (st & 1) == 0  # non-finite trajectories are dropped, as the plotters' try/except does
    g = g[ok]
    phibar = g[:, :, 0:5] * g[:, :, 6:11]
    q = np.asarray(quantiles, dtype=np.float64)
    return {
        "t": solver.time_array()[rows], "rows": rows, "quantiles": q, "phibar_q": np.quantile(phibar, q, axis=0),
        "total_q": np.quantile(phibar.sum(axis=2), q, axis=0), "mean": phibar.mean(axis=0), "n_used": int(ok.sum()),
        "n_failed": int((~ok).sum()),
    }


def sample_theta(bounds: np.ndarray, rng: np.random.Generator, theta_map: Optional[np.ndarray] = None, map_std_frac: float = 0.1) -> np.ndarray:
    """One parameter vector, drawn component by component in the reference's order (generate_training_data.py:52-73):
    locked -> low; MAP-centred -> clip(N(theta_MAP_i, (frac * range_i)^2)); otherwise U(low_i, high_i)."""
    theta = np.zeros(20)
    for i in range(20):
        lo, hi = bounds[i]
        if abs(hi - lo) < 1e-12:
            theta[i] = lo
        elif theta_map is not None:
            theta[i] = np.clip(rng.normal(theta_map[i], map_std_frac * (hi - lo)), lo, hi)
        else:
            theta[i] = rng.uniform(lo, hi)
    return theta


def expand_bounds_to_include(bounds: np.ndarray, theta_map: Optional[np.ndarray]) -> np.ndarray:
    """Widen a bound that excludes theta_MAP by 20 % of max(range, |theta|/2, 0.5) (generate_training_data.py:186-203)."""
    b = np.array(bounds, dtype=np.float64, copy=True)
    if theta_map is None:
        return b
    for i in range(20):
        lo, hi = b[i]
        if abs(hi - lo) < 1e-12:
            continue
        val = theta_map[i]
        margin = 0.2 * max(hi - lo, abs(val) * 0.5, 0.5)
        if val < lo:
            b[i, 0] = val - margin
        if val > hi:
            b[i, 1] = val + margin
    return b


def _usable(phi: np.ndarray) -> np.ndarray:
    """run_single's acceptance test (generate_training_data.py:76-88), per trajectory of phi[N, n_time, 5]."""
    flat = phi.reshape(phi.shape[0], -1)
    return np.isfinite(flat).all(axis=1) & ~(flat < -0.5).any(axis=1) & ~(flat > 1.5).any(axis=1)


def generate_dataset(n_samples: int, bounds, seed: int = 42, maxtimestep: int = 500, dt: float = 1e-5, n_time_out: int = 100,
                     theta_map: Optional[np.ndarray] = None, map_frac: float = 0.3, map_std_frac: float = 0.1, expand_bounds: bool = True,
                     posterior_samples: Optional[np.ndarray] = None, posterior_frac: float = 0.5, device: int = 0, solver=None,
                     jit_warmup_draw: bool = True, block: int = 2048) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """(theta[N,20], phi[N,n_time_out,5], t_out, bounds) -- the surrogate training set of the reference
    (deeponet/generate_training_data.py:144-333), same random stream and same category order (posterior draws first,
    then MAP-centred, then uniform; a failed solve is redrawn in place), but the solves run as GPU batches.

    The usability test of a trajectory (finite, phi within [-0.5, 1.5]) looks at all T+1 levels like the reference's;
    candidates are solved in blocks of `block` and a failure discards at most the rest of its block; `jit_warmup_draw` keeps
    the reference's one discarded draw ("Warming up JIT", :249-251) so that the stream lines up."""
    bounds = np.array(bounds, dtype=np.float64, copy=True)
    if expand_bounds:
        bounds = expand_bounds_to_include(bounds, theta_map)
    rng = np.random.default_rng(seed)
    tm = theta_map if map_frac > 0 else None
    n_post = int(n_samples * posterior_frac) if posterior_samples is not None else 0
    n_rem = n_samples - n_post
    if tm is not None and n_rem > 0:
        n_map = int(n_rem * map_frac)
    elif posterior_samples is not None:
        n_map = 0
    elif tm is not None:
        n_map = int(n_samples * map_frac)
    else:
        n_map = 0
    solver = solver or BiofilmNewtonSolver5S(dt=dt, maxtimestep=maxtimestep, eps=1e-6, Kp1=1e-4, c_const=100.0, alpha_const=100.0,
                                             phi_init=0.2, K_hill=0.05, n_hill=4.0, max_newton_iter=50, device=device)
    if jit_warmup_draw:
        sample_theta(bounds, rng)
    rows = np.linspace(0, maxtimestep, n_time_out, dtype=int)
    if np.any(np.diff(rows) <= 0):
        raise ValueError("n_time_out must not exceed maxtimestep + 1")
    t_out = solver.time_array()[rows]

    thetas, phis = [], []
    n_post_done = n_map_done = post_idx = n_failed = 0
    while len(thetas) < n_samples:
        # draw the rest of the set assuming every solve succeeds, remembering the stream position before each draw
        cand, states, kinds = [], [], []
        pd, md, pi = n_post_done, n_map_done, post_idx
        for _ in range(n_samples - len(thetas)):
            states.append(rng.bit_generator.state)
            use_post = posterior_samples is not None and pd < n_post
            use_map = (not use_post) and tm is not None and md < n_map
            if use_post:
                th = np.array(posterior_samples[pi % len(posterior_samples)][:20], dtype=np.float64)
                pi += 1
                th = np.clip(th + rng.normal(0, 0.01, size=20), bounds[:, 0], bounds[:, 1])
                pd += 1
            else:
                th = sample_theta(bounds, rng, theta_map=tm if use_map else None, map_std_frac=map_std_frac)
                md += 1 if use_map else 0
            cand.append(th)
            kinds.append(0 if use_post else (1 if use_map else 2))
        cand = np.array(cand)
        # Solve block by block and stop at the first unusable trajectory: a failure costs at most one block of solves.  The
        # usability test scans ALL maxtimestep + 1 levels like the reference (generate_training_data.py:266-272), so the whole
        # trajectory of a block comes back (48 KB per solve at 500 steps) and is down-sampled here.
        phi_blocks, upto = [], len(cand)
        for lo in range(0, len(cand), block):
            g, st = solver.run_deterministic_batch(cand[lo:lo + block], return_status=True)
            good = _usable(g[:, :, :5]) & ((st & 1) == 0)
            phi_blocks.append(g[:, rows, :5])
            bad = np.nonzero(~good)[0]
            if bad.size:
                upto = lo + int(bad[0])
                break
        phi = np.concatenate(phi_blocks, axis=0)
        bad = np.array([upto]) if upto < len(cand) else np.array([], dtype=int)
        for k in range(upto):
            thetas.append(cand[k])
            phis.append(phi[k])
            n_post_done += kinds[k] == 0
            n_map_done += kinds[k] == 1
            post_idx += kinds[k] == 0
        if bad.size:
            # the failed draw consumed its random numbers (and its posterior slot index) but not its category quota
            n_failed += 1
            post_idx += kinds[upto] == 0
            if upto + 1 < len(states):
                rng.bit_generator.state = states[upto + 1]
            # (a failure on the very last candidate leaves the stream after its draw, which is where it already is)
    return np.array(thetas), np.array(phis), t_out, bounds
