"""TMCMC with random-walk / HMC / NUTS mutation -- the reference's gradient engine
(data_5species/main/tmcmc_nuts_engine.py: `tmcmc_engine`, `hmc_step`, `nuts_step`, dual averaging) with the
population advanced AS A BATCH: every leapfrog step of all particles is one `value_and_grad` call, i.e. one
hmx_loglik_grad_batch launch (forward sensitivities on the GPU) instead of one `jax.value_and_grad` call per particle.

    from tmcmc202601_b200.nuts import tmcmc_engine
    out = tmcmc_engine(evaluator, prior_bounds[20, 2], mutation="nuts", n_particles=200, seed=42)

`target` is anything with `value_and_grad(theta[N, d]) -> (logL[N], grad[N, d])` (LogLikelihoodEvaluator has it) or a
plain callable with that signature.  Per particle the algorithms are the reference's, statement by statement:
  * leapfrog with the position clipped into the prior box after the drift                       (:20-28)
  * HMC: n_leapfrog steps, momentum flip, Metropolis test on the Hamiltonian                    (:125-148)
  * NUTS: tree doubling up to max_depth, 2^depth leapfrogs per doubling from the chosen end, slice variable, divergence at
    dH > 1000, progressive candidate sampling, U-turn test on the end points                    (:34-93)
  * dual-averaging step size during the warm-up stages (NUTS)                                   (:96-122)
  * stage loop: 50-step ESS bisection, multinomial resampling, tempered target beta * logL      (:216-249)
The random numbers come from NumPy's Generator instead of jax.random (JAX is not a dependency here), so parity with the
reference is statistical.  Returned dict: the reference's keys (:333-348).
"""

from __future__ import annotations

import time
from typing import Callable, Optional

import numpy as np


def _vg(target) -> Callable:
    return target.value_and_grad if hasattr(target, "value_and_grad") else target


def dual_averaging_init(step_size_init=1.0, target_accept=0.65, gamma=0.05, t0=10, kappa=0.75, eps_max_factor=5.0) -> dict:
    return {"log_eps": np.log(step_size_init), "log_eps_bar": np.log(step_size_init), "H_bar": 0.0, "mu": np.log(10 * step_size_init),
            "target_accept": target_accept, "gamma": gamma, "t0": t0, "kappa": kappa, "m": 0, "eps_max": step_size_init * eps_max_factor}


def dual_averaging_update(state: dict, accept_prob: float) -> dict:
    m = state["m"] + 1
    eta = 1.0 / (m + state["t0"])
    H_bar = (1 - eta) * state["H_bar"] + eta * (state["target_accept"] - accept_prob)
    log_eps = min(state["mu"] - np.sqrt(m) / state["gamma"] * H_bar, np.log(state["eps_max"]))
    mk = m ** (-state["kappa"])
    log_eps_bar = min(mk * log_eps + (1 - mk) * state["log_eps_bar"], np.log(state["eps_max"]))
    return {**state, "log_eps": log_eps, "log_eps_bar": log_eps_bar, "H_bar": H_bar, "m": m}


def hmc_step_batch(rng, theta, vg, step_size, n_leapfrog, lo, hi):
    """One HMC transition for every row of theta[N, d] (reference hmc_step, :125-148).  Returns (theta_new, accepted[N], logp_new,
    n_gradient_batches)."""
    N, d = theta.shape
    p = rng.standard_normal((N, d))
    logp0, g = vg(theta)
    H0 = -logp0 + 0.5 * np.sum(p * p, axis=1)
    q = theta.copy()
    p = p + 0.5 * step_size * g
    for _ in range(n_leapfrog - 1):
        q = np.clip(q + step_size * p, lo, hi)
        _, g = vg(q)
        p = p + step_size * g
    q = np.clip(q + step_size * p, lo, hi)
    logp1, g = vg(q)
    p = -(p + 0.5 * step_size * g)
    H1 = -logp1 + 0.5 * np.sum(p * p, axis=1)
    with np.errstate(divide="ignore"):
        acc = np.log(rng.random(N)) < (H0 - H1)
    acc &= np.isfinite(H1)
    return np.where(acc[:, None], q, theta), acc, np.where(acc, logp1, logp0), n_leapfrog + 1


def nuts_step_batch(rng, theta, vg, step_size, lo, hi, max_depth=6):
    """One NUTS transition per row (reference nuts_step, :34-93), all particles in lock-step over the tree depth; particles
    whose tree has stopped drop out of the batch.  Returns (theta_new, accepted[N], logp_new[N], n_leapfrog[N])."""
    N, d = theta.shape
    logp0, g0 = vg(theta)
    p0 = rng.standard_normal((N, d))
    H0 = -logp0 + 0.5 * np.sum(p0 * p0, axis=1)
    with np.errstate(divide="ignore"):
        log_u = np.log(rng.random(N)) + logp0 - 0.5 * np.sum(p0 * p0, axis=1)
    qm, pm, gm = theta.copy(), p0.copy(), g0.copy()
    qp, pp, gp = theta.copy(), p0.copy(), g0.copy()
    q_prop, logp_prop = theta.copy(), logp0.copy()
    n_valid = np.ones(N, dtype=np.int64)
    going = np.ones(N, dtype=bool)
    n_lf = np.zeros(N, dtype=np.int64)
    for depth in range(max_depth):
        act = np.nonzero(going)[0]
        if act.size == 0:
            break
        direction = np.where(rng.random(act.size) < 0.5, 1.0, -1.0)
        fwd = direction > 0
        q = np.where(fwd[:, None], qp[act], qm[act])
        p = np.where(fwd[:, None], pp[act], pm[act])
        g = np.where(fwd[:, None], gp[act], gm[act])
        q_cand, logp_cand = q.copy(), logp0[act].copy()
        n_sub = np.zeros(act.size, dtype=np.int64)
        diverged = np.zeros(act.size, dtype=bool)
        alive = np.ones(act.size, dtype=bool)  # still inside this doubling (not diverged)
        for _ in range(2 ** depth):
            sel = np.nonzero(alive)[0]
            if sel.size == 0:
                break
            dsel = direction[sel][:, None]
            ps = dsel * p[sel] + 0.5 * step_size * g[sel]  # leapfrog on (q, direction * p)
            qs = np.clip(q[sel] + step_size * ps, lo, hi)
            lps, gs = vg(qs)
            ps = ps + 0.5 * step_size * gs
            p_new = dsel * ps
            q[sel], p[sel], g[sel] = qs, p_new, gs
            n_lf[act[sel]] += 1
            H_new = -lps + 0.5 * np.sum(p_new * p_new, axis=1)
            div = ~(H_new - H0[act[sel]] <= 1000.0)  # also catches NaN
            diverged[sel[div]] = True
            alive[sel[div]] = False
            ok = ~div & (-H_new > log_u[act[sel]])
            idx = sel[ok]
            n_sub[idx] += 1
            take = rng.random(idx.size) < 1.0 / np.maximum(n_sub[idx], 1)
            q_cand[idx[take]] = qs[ok][take]
            logp_cand[idx[take]] = lps[ok][take]
        keep = ~diverged
        # the chosen end of the trajectory moves (not after a divergence: the reference breaks out before the update)
        upd_f, upd_b = keep & fwd, keep & ~fwd
        qp[act[upd_f]], pp[act[upd_f]], gp[act[upd_f]] = q[upd_f], p[upd_f], g[upd_f]
        qm[act[upd_b]], pm[act[upd_b]], gm[act[upd_b]] = q[upd_b], p[upd_b], g[upd_b]
        has = keep & (n_sub > 0)
        prob = n_sub / np.maximum(n_valid[act] + n_sub, 1)
        swap = has & (rng.random(act.size) < prob)
        q_prop[act[swap]] = q_cand[swap]
        logp_prop[act[swap]] = logp_cand[swap]
        n_valid[act[keep]] += n_sub[keep]
        dq = qp[act] - qm[act]
        uturn = (np.sum(dq * pm[act], axis=1) < 0) | (np.sum(dq * pp[act], axis=1) < 0)
        going[act] = keep & ~uturn
    accepted = np.any(q_prop != theta, axis=1)
    return q_prop, accepted, logp_prop, n_lf


def tmcmc_engine(target, prior_bounds, mutation: str = "nuts", n_particles: int = 200, max_stages: int = 30, target_ess_ratio: float = 0.5,
                 hmc_step_size: float = 0.01, hmc_n_leapfrog: int = 10, nuts_max_depth: int = 6, warmup_stages: int = 3, seed: int = 42,
                 label: Optional[str] = None, verbose: bool = False, log_prior_fn=None) -> dict:
    """Reference `tmcmc_engine` (:151-348) with batched mutations.  `log_prior_fn` (the GNN prior hook) is not supported."""
    if log_prior_fn is not None:
        raise NotImplementedError("the GNN log-prior hook of the reference engine is outside the likelihood hot path")
    if mutation not in ("rw", "hmc", "nuts"):
        raise ValueError("mutation must be 'rw', 'hmc' or 'nuts'")
    label = label or f"{mutation.upper()}-TMCMC"
    vg = _vg(target)
    rng = np.random.default_rng(seed)
    prior_bounds = np.asarray(prior_bounds, dtype=np.float64)
    d = prior_bounds.shape[0]
    lo, hi = prior_bounds[:, 0], prior_bounds[:, 1]
    free = np.abs(hi - lo) > 1e-12
    free_dims = np.nonzero(free)[0]
    particles = np.where(free[None, :], rng.uniform(lo, np.where(free, hi, lo + 1.0), size=(n_particles, d)), lo[None, :])
    t0 = time.time()
    logL = vg(particles)[0]
    scales = np.where(hi - lo < 1e-12, 1.0, hi - lo)
    init_eps = hmc_step_size * float(np.mean(scales[free]))
    da = dual_averaging_init(step_size_init=init_eps, target_accept=0.65)
    beta, betas = 0.0, [0.0]
    stage_times, accept_rates, ess_history, n_lf_history, eps_history = [], [], [], [], []
    stage = 0
    while beta < 1.0 and stage < max_stages:
        stage += 1
        ts = time.time()

        def ess_of(db):
            w = np.exp(np.clip(db * logL - logL.max(), -500, 500))
            return w.sum() ** 2 / np.sum(w * w)

        a, b = 0.0, 1.0 - beta
        for _ in range(50):
            mid = 0.5 * (a + b)
            if ess_of(mid) > target_ess_ratio * n_particles:
                a = mid
            else:
                b = mid
        delta = a if a >= 1e-6 else 1.0 - beta
        delta = min(delta, 1.0 - beta)
        beta_new = min(beta + delta, 1.0)
        w = np.exp(np.clip((beta_new - beta) * logL - np.max((beta_new - beta) * logL), -500, 500))
        ess_val = w.sum() ** 2 / np.sum(w * w)
        idx = rng.choice(n_particles, size=n_particles, p=w / w.sum())
        particles, logL = particles[idx].copy(), logL[idx].copy()

        def tempered(theta):
            v, g = vg(theta)
            return beta_new * v, beta_new * g

        eps = float(np.exp(da["log_eps"])) if mutation == "nuts" else init_eps
        n_lf_stage = 0
        if mutation == "rw":
            cov = np.atleast_2d(np.cov(particles[:, free_dims].T)) * 0.04
            prop = particles.copy()
            prop[:, free_dims] += rng.multivariate_normal(np.zeros(free_dims.size), cov, size=n_particles)
            inside = np.all((prop[:, free_dims] >= lo[free_dims]) & (prop[:, free_dims] <= hi[free_dims]), axis=1)
            ll_new = np.full(n_particles, -np.inf)
            if inside.any():
                ll_new[inside] = vg(prop[inside])[0]
            with np.errstate(divide="ignore", invalid="ignore"):
                acc = inside & (np.log(rng.random(n_particles)) < beta_new * (ll_new - logL))
            particles[acc], logL[acc] = prop[acc], ll_new[acc]
        elif mutation == "hmc":
            new, acc, logp, nb = hmc_step_batch(rng, particles, tempered, eps, hmc_n_leapfrog, lo, hi)
            particles[acc], logL[acc] = new[acc], logp[acc] / beta_new
            n_lf_stage = hmc_n_leapfrog * n_particles
        else:
            new, acc, logp, nlf = nuts_step_batch(rng, particles, tempered, eps, lo, hi, max_depth=nuts_max_depth)
            particles[acc], logL[acc] = new[acc], logp[acc] / beta_new
            n_lf_stage = int(nlf.sum())
            if stage <= warmup_stages:
                da = dual_averaging_update(da, float(np.mean(acc)))
        beta = beta_new
        betas.append(beta)
        stage_times.append(time.time() - ts)
        accept_rates.append(float(np.mean(acc)))
        ess_history.append(float(ess_val))
        n_lf_history.append(n_lf_stage)
        eps_history.append(eps)
        if verbose:
            print(f"  Stage {stage:2d}: beta={beta:.4f}, accept={accept_rates[-1]:.2f}, ESS={ess_val:.0f}, "
                  f"logL=[{logL.min():.1f},{logL.max():.1f}], {stage_times[-1]:.1f}s")
    best = int(np.argmax(logL))
    return {"label": label, "mutation": mutation, "samples": particles, "log_likelihoods": logL, "betas": np.array(betas),
            "theta_MAP": particles[best], "total_time": time.time() - t0, "stage_times": stage_times, "accept_rates": accept_rates,
            "ess_history": ess_history, "n_stages": len(betas) - 1, "n_leapfrog_history": n_lf_history, "eps_history": eps_history,
            "final_eps": float(np.exp(da["log_eps"])) if mutation == "nuts" else eps}
