"""TMCMC engine with the reference's call surface (data_5species/core/tmcmc.py), stage arithmetic on the GPU.

Names, arguments, return types and the order of random draws follow the reference so existing drivers
(`run_estimation`, MAIN:2280-2299) keep working:

    run_TMCMC, run_multi_chain_TMCMC, TMCMCResult, systematic_resample, evaluate_particles_parallel,
    reflect_into_bounds, compute_MAP_with_uncertainty, choose_subset_size, should_do_fom_check

What runs where
  * likelihood of the whole population / of all N*K proposals: ONE `evaluator.batch(...)` call ->
    hmx_loglik_batch (the reference: one pickled future per particle in a process pool, TM:289-303);
  * delta-beta bisection, weights + evidence, systematic resampling, weighted covariance: `StageOps`
    -> hmx_beta_step / hmx_weights / hmx_resample / hmx_weighted_cov (TM:613-659, 686-701, 186-198, 766);
  * proposal generation, reflection and the MH accept/reject are host NumPy, vectorised over particles with
    the SAME draw order as the reference's per-particle Python loops (TM:870-945): d standard normals per
    (particle, step) in particle-major order, then one uniform per evaluated proposal in the same order.
    Quirk kept on purpose: all K proposals of a particle are drawn around its
    resampled state, then accepted sequentially.

`StageOps` is the seam used by the tests: the product default (`GpuStageOps`) calls the CUDA library and
raises if it is unavailable; the CPU test-suite injects an oracle-backed implementation to exercise the
host control flow without a GPU.  The product never falls back on its own.
"""

from __future__ import annotations

import logging
import threading
import time
import zlib
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

from .utils import TimingStats

logger = logging.getLogger(__name__)

# constants of tmcmc/program2602/config.py:29-86 (TMCMC_DEFAULTS, PROPOSAL_DEFAULTS, ...)
DEFAULT_N_PARTICLES = 500
DEFAULT_N_STAGES = 30
DEFAULT_TARGET_ESS_RATIO = 0.5
DEFAULT_MIN_DELTA_BETA = 0.02
MAX_DELTA_BETA = 0.05
DEFAULT_UPDATE_LINEARIZATION_INTERVAL = 3
DEFAULT_N_MUTATION_STEPS = 1
DEFAULT_LINEARIZATION_THRESHOLD = 0.75
MAX_LINEARIZATION_UPDATES = 5
MUTATION_SCALE_FACTOR = 2.0
OPTIMAL_SCALE_FACTOR = 2.38**2
COVARIANCE_NUGGET_BASE = 1e-8
COVARIANCE_NUGGET_SCALE = 1e-6
BETA_CONVERGENCE_THRESHOLD = 1.0
THETA_CONVERGENCE_THRESHOLD = 1e-3
ROM_ERROR_THRESHOLD = 0.01
MAX_THETA0_STEP_NORM = 0.75
MAX_LINEARIZATION_SUBUPDATES_PER_EVENT = 3
MIN_ACCEPTANCE_RATE = 0.01  # DebugConfig.min_acceptance_rate


@dataclass
class TMCMCResult:
    """Same fields as the reference dataclass (TM:160-183)."""

    samples: np.ndarray
    logL_values: np.ndarray
    theta_MAP: np.ndarray
    beta_schedule: List[float]
    converged: bool
    theta0_history: Optional[List[np.ndarray]] = None
    n_linearization_updates: int = 0
    final_MAP: Optional[np.ndarray] = None
    rom_error_pre_history: Optional[List[float]] = None
    rom_error_history: Optional[List[float]] = None
    acc_rate_history: Optional[List[float]] = None
    n_rom_evaluations: int = 0
    n_fom_evaluations: int = 0
    wall_time_s: float = 0.0
    timing_breakdown_s: Optional[Dict[str, float]] = None
    likelihood_health: Optional[Dict[str, int]] = None
    stage_summary: Optional[List[Dict[str, Any]]] = None
    log_evidence: float = 0.0


# --------------------------------------------------------------------------------------------- stage ops
class GpuStageOps:
    """Per-stage arithmetic through libhamilton (K2-K5).  One shared context per device."""

    _engines: Dict[int, Any] = {}
    _engines_lock = threading.Lock()

    def __init__(self, engine=None, device: int = 0):
        if engine is None:
            with GpuStageOps._engines_lock:  # chains of run_batch_TMCMC construct stage ops from several threads
                engine = GpuStageOps._engines.get(device)
                if engine is None:
                    from . import _abi
                    from .engine import HamiltonEngine

                    cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.2, tsm_variance=False)
                    engine = GpuStageOps._engines[device] = HamiltonEngine(_abi.make_config([cond]), device=device)
        self.engine = engine

    def beta_step(self, logL, beta, target_ess_ratio, min_delta_beta, max_delta_beta):
        db, ess = self.engine.beta_step(logL, beta, target_ess_ratio, min_delta_beta, max_delta_beta)
        return db, (None if np.isnan(ess) else ess)

    def weights(self, logL, delta_beta):
        w, log_Sj = self.engine.weights(logL, delta_beta)
        return w, (None if np.isnan(log_Sj) else log_Sj)

    def resample(self, w, u):
        return self.engine.resample(w, u)

    def weighted_cov(self, theta, w):
        return self.engine.weighted_cov(theta, w)[1]


_default_stage_ops: Optional[Any] = None


def get_stage_ops(evaluator=None, log_likelihood=None):
    """The GPU context of the evaluator (or of the likelihood callable, when that is an evaluator of this package) when it
    has one -- stage kernels then run on the SAME device as the likelihood -- else a stage-only context on that object's
    `.device`, else (nothing known) on device 0."""
    for obj in (evaluator, log_likelihood):
        eng = getattr(obj, "_engine", None)
        if eng is not None and hasattr(eng, "beta_step"):
            return GpuStageOps(engine=eng)
    for obj in (evaluator, log_likelihood):
        dev = getattr(obj, "device", None)
        if isinstance(dev, int):
            return GpuStageOps(device=dev)
    global _default_stage_ops
    if _default_stage_ops is None:
        _default_stage_ops = GpuStageOps()
    return _default_stage_ops


# --------------------------------------------------------------------------------------------- helpers
def compute_MAP_with_uncertainty(samples: np.ndarray, logL: np.ndarray) -> Dict[str, np.ndarray]:
    """Posterior summaries (TM:123-138): mean, std(ddof=1), MAP = argmax logL, 2.5 / 97.5 percentiles."""
    return {
        "mean": samples.mean(axis=0),
        "std": samples.std(axis=0, ddof=1),
        "MAP": samples[np.argmax(logL)],
        "ci_lower": np.percentile(samples, 2.5, axis=0),
        "ci_upper": np.percentile(samples, 97.5, axis=0),
    }


def systematic_resample(weights: np.ndarray, rng: np.random.Generator, stage_ops=None) -> np.ndarray:
    """Systematic resampling (TM:186-198): one uniform from `rng`, cumsum + searchsorted on the GPU (bit-exact)."""
    ops = stage_ops or get_stage_ops()
    return ops.resample(np.asarray(weights, dtype=np.float64), rng.random())


def reflect_into_bounds(x, low, high):
    """Fold x into [low, high] by reflection; width <= 0 clips (TM:201-226).  Works on scalars and arrays.

    Same arithmetic as the reference's scalar function, `low + fold((x - low) % (2 width))`, but the modulo
    (exact, hence the identity for 0 <= x - low < 2 width) is only evaluated where it can change the value."""
    scalar = np.ndim(x) == 0
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    low = np.asarray(low, dtype=np.float64)
    high = np.asarray(high, dtype=np.float64)
    if x.ndim == 2 and x.shape[0] > 32768:  # cache-sized row blocks: the passes below are memory bound
        out = np.empty_like(x)
        for a in range(0, x.shape[0], 16384):
            out[a:a + 16384] = reflect_into_bounds(x[a:a + 16384], low, high)
        return out
    width = np.broadcast_to(high - low, x.shape)
    lowb = np.broadcast_to(low, x.shape)
    y = x - lowb
    two_w = 2.0 * width
    far = (y < 0.0) | (y >= two_w)
    far &= width > 0
    if far.any():
        y[far] = np.mod(y[far], two_w[far])
    over = y > width
    over &= width > 0
    if over.any():
        y[over] = two_w[over] - y[over]
    y += lowb
    degenerate = width <= 0
    if degenerate.any():
        y[degenerate] = np.clip(x, lowb, np.broadcast_to(high, x.shape))[degenerate]
    return float(y[0]) if scalar else y


def evaluate_particles_parallel(log_likelihood: Callable, theta: np.ndarray, n_jobs: Optional[int] = None,
                                use_threads: bool = False) -> np.ndarray:
    """Population likelihood (TM:247-312).  Evaluators exposing `.batch` get ONE batched GPU call; a plain
    callable is mapped sequentially.  n_jobs / use_threads are accepted and ignored (no process pool)."""
    theta = np.asarray(theta, dtype=np.float64)
    if theta.shape[0] == 0:
        return np.array([])
    batch = getattr(log_likelihood, "batch", None)
    if callable(batch):
        return np.asarray(batch(theta), dtype=np.float64)
    out = np.empty(theta.shape[0])
    for i, t in enumerate(theta):
        try:
            out[i] = log_likelihood(t)
        except Exception as e:  # the reference maps worker failures to -1e20 (TM:304-306)
            logger.warning("Particle %d evaluation failed: %s", i, e)
            out[i] = -1e20
    return out


def _validate_tmcmc_inputs(log_likelihood, prior_bounds, n_particles, n_stages, target_ess_ratio, evaluator,
                           theta_base_full, active_indices) -> None:
    """Same checks and exception types as data_5species/utils/validation.py:14-82."""
    if not callable(log_likelihood):
        raise TypeError("log_likelihood must be callable")
    if not isinstance(prior_bounds, list) or len(prior_bounds) == 0:
        raise ValueError("prior_bounds must be a non-empty list")
    for i, (low, high) in enumerate(prior_bounds):
        if not isinstance(low, (int, float)) or not isinstance(high, (int, float)):
            raise TypeError(f"prior_bounds[{i}] must be numeric tuple")
        if low > high:
            raise ValueError(f"prior_bounds[{i}]: lower bound must be <= upper bound")
    if n_particles <= 0:
        raise ValueError(f"n_particles must be > 0, got {n_particles}")
    if n_stages <= 0:
        raise ValueError(f"n_stages must be > 0, got {n_stages}")
    if not (0 < target_ess_ratio <= 1):
        raise ValueError(f"target_ess_ratio must be in (0, 1], got {target_ess_ratio}")
    if evaluator is not None:
        if theta_base_full is None:
            raise ValueError("theta_base_full must be provided when evaluator is provided")
        if active_indices is None:
            raise ValueError("active_indices must be provided when evaluator is provided")
        if not isinstance(theta_base_full, np.ndarray):
            raise TypeError("theta_base_full must be numpy array")
        if not isinstance(active_indices, list):
            raise TypeError("active_indices must be list")


def _mvn_factor(cov: np.ndarray) -> np.ndarray:
    """The factor Generator.multivariate_normal(method='svd') multiplies standard normals with: (u*sqrt(s))."""
    u, s, _ = np.linalg.svd(cov)
    return u * np.sqrt(s)


# --------------------------------------------------------------------------------------------- run_TMCMC
def run_TMCMC(
    log_likelihood: Callable,
    prior_bounds: List[Tuple[float, float]],
    n_particles: int = DEFAULT_N_PARTICLES,
    n_stages: int = DEFAULT_N_STAGES,
    target_ess_ratio: float = DEFAULT_TARGET_ESS_RATIO,
    min_delta_beta: float = DEFAULT_MIN_DELTA_BETA,
    max_delta_beta: float = MAX_DELTA_BETA,
    logL_scale: float = 1.0,
    seed: Optional[int] = None,
    model_name: str = "",
    evaluator: Optional[Any] = None,
    theta_base_full: Optional[np.ndarray] = None,
    active_indices: Optional[List[int]] = None,
    update_linearization_interval: int = DEFAULT_UPDATE_LINEARIZATION_INTERVAL,
    n_mutation_steps: int = DEFAULT_N_MUTATION_STEPS,
    use_observation_based_update: bool = True,
    linearization_threshold: float = DEFAULT_LINEARIZATION_THRESHOLD,
    linearization_enable_rom_threshold: float = 0.05,
    debug_logger: Optional[Any] = None,
    force_beta_one: bool = False,
    n_jobs: Optional[int] = None,
    use_threads: bool = False,
    gnn_prior: Optional[Any] = None,
    beta_schedule: str = "ess",
    target_cov: float = 1.0,
    proposal_type: str = "gaussian",
    dreamzs_gamma: Optional[float] = None,
    dreamzs_snooker_prob: float = 0.1,
    stage_ops: Optional[Any] = None,
    device_mutation: bool = False,
) -> TMCMCResult:
    """One TMCMC chain (TM:393-1842).  Extra keywords: `stage_ops` selects the stage-arithmetic backend (default:
    the GPU); `device_mutation=True` runs proposal generation, reflection, likelihood and accept/reject of every
    stage on the GPU (`log_likelihood.mutate_device`, Philox random numbers: same algorithm and constants, a
    different random stream, so results agree with the host path statistically, not bitwise)."""
    _validate_tmcmc_inputs(log_likelihood, prior_bounds, n_particles, n_stages, target_ess_ratio, evaluator,
                           theta_base_full, active_indices)
    if proposal_type != "gaussian":
        raise NotImplementedError("only the Gaussian MH proposal is on the production path (TM:899-900)")
    if beta_schedule not in ("ess", "cov"):
        raise ValueError("beta_schedule must be 'ess' or 'cov'")
    ops = stage_ops or get_stage_ops(evaluator, log_likelihood)
    if device_mutation:
        if not callable(getattr(log_likelihood, "mutate_device", None)):
            raise ValueError("device_mutation=True needs a likelihood with .mutate_device (LogLikelihoodEvaluator)")
        if gnn_prior is not None:
            raise NotImplementedError("device_mutation supports the uniform box prior only")
    # CoV(w) <= c  <=>  ESS >= N / (1 + c^2): the "cov" schedule (TM:636-643) is the ESS bisection with this ratio
    ess_ratio = target_ess_ratio if beta_schedule == "ess" else 1.0 / (1.0 + float(target_cov) ** 2)

    rng = np.random.default_rng(seed)
    # key of the device-side Philox stream: a function of `seed` only, drawn from a separate generator so that the
    # host stream (initial population, resampling offsets, subset draws) is the same with and without device_mutation
    mut_seed = int(np.random.default_rng([0x6D757461, int(seed) & 0xFFFFFFFF if seed is not None else 0]).integers(0, 2**63)) \
        if seed is not None else int(np.random.default_rng().integers(0, 2**63))
    mut_sequence = 0
    t_start = time.perf_counter()
    n_params = len(prior_bounds)
    lows = np.array([b[0] for b in prior_bounds], dtype=np.float64)
    highs = np.array([b[1] for b in prior_bounds], dtype=np.float64)

    def log_prior_batch(th: np.ndarray) -> np.ndarray:
        th = np.atleast_2d(th)
        inside = np.all((th >= lows) & (th <= highs), axis=1)
        lp = np.where(inside, 0.0, -np.inf)
        if gnn_prior is not None:
            lp = np.array([gnn_prior.log_density(t) if ok else -np.inf for t, ok in zip(th, inside)])
        return lp

    # initial population: rng.uniform(low_j, high_j) in row-major order (TM:513-515)
    if gnn_prior is not None:
        theta = np.array([gnn_prior.sample(rng, prior_bounds) for _ in range(n_particles)], dtype=np.float64)
    else:
        theta = rng.uniform(lows, highs, size=(n_particles, n_params))
    logL = evaluate_particles_parallel(log_likelihood, theta, n_jobs=n_jobs, use_threads=use_threads)

    beta = 0.0
    beta_sched = [beta]
    theta0_history: List[np.ndarray] = []
    n_linearization_updates = 0
    rom_error_pre_history: List[float] = []
    rom_error_history: List[float] = []
    acc_rate_history: List[float] = []
    log_evidence = 0.0
    theta_MAP_posterior_history: List[np.ndarray] = []
    stage_summary: List[Dict[str, Any]] = []

    initial_rom_count = initial_fom_count = 0
    if evaluator is not None:
        initial_rom_count = evaluator.call_count
        initial_fom_count = evaluator.fom_call_count
        th0 = evaluator.get_linearization_point()
        if th0 is not None:
            theta0_history.append(th0.copy())
    min_acc = MIN_ACCEPTANCE_RATE
    if debug_logger is not None and hasattr(debug_logger, "config"):
        min_acc = getattr(debug_logger.config, "min_acceptance_rate", MIN_ACCEPTANCE_RATE)

    for stage in range(1, n_stages + 1):
        rom_error_pre_stage = rom_error_post_stage = delta_theta0_stage = None

        # 1. delta-beta by ESS bisection (TM:613-659)
        delta_beta, ess_at_lo = ops.beta_step(logL, beta, ess_ratio, min_delta_beta, float(max_delta_beta))
        beta_next = min(beta + delta_beta, 1.0)
        if force_beta_one and stage == n_stages and beta_next < 1.0:
            beta_next = 1.0
            delta_beta = 1.0 - beta

        # 2. weights + evidence (TM:686-701)
        w, log_Sj = ops.weights(logL, beta_next - beta)
        if log_Sj is not None and np.isfinite(log_Sj):
            log_evidence += log_Sj

        theta_before = theta.copy()
        logL_before = logL.copy()
        w_before = w.copy()
        log_post_before = log_prior_batch(theta_before) + beta_next * logL_before

        idx = ops.resample(w, rng.random())  # systematic_resample(w, rng), TM:742
        theta = theta[idx]
        logL = logL[idx]
        theta_after_resample = theta.copy()
        logL_after_resample = logL.copy()

        # 3. proposal covariance (TM:766-811)
        cov_base = np.asarray(ops.weighted_cov(theta_before, w_before), dtype=np.float64).reshape(n_params, n_params)
        tempered_scale = (OPTIMAL_SCALE_FACTOR / n_params) / max(beta_next, 0.1)
        adaptive_scale_factor = MUTATION_SCALE_FACTOR
        if acc_rate_history:
            adaptive_scale_factor = MUTATION_SCALE_FACTOR * ((1.0 / 9.0) + (8.0 / 9.0) * float(acc_rate_history[-1]))
        cov = cov_base * (adaptive_scale_factor * tempered_scale)
        nugget = COVARIANCE_NUGGET_BASE + COVARIANCE_NUGGET_SCALE * (np.trace(cov_base) / n_params)
        cov = cov + nugget * np.eye(n_params)

        def mutate(cov_matrix: np.ndarray, steps: int):
            """K-step mutation of the whole population (TM:849-962)."""
            nonlocal mut_sequence
            K = int(max(1, steps))
            base_theta = theta_after_resample
            base_logL = logL_after_resample
            if device_mutation:
                th_new, ll_new, st = log_likelihood.mutate_device(base_theta, base_logL, _mvn_factor(cov_matrix), lows, highs, K,
                                                                  beta_next, mut_seed, sequence=mut_sequence)
                mut_sequence += 1
                total = int(st["n_candidates"])
                n_acc = int(st["n_accepted"])
                return th_new, ll_new, (n_acc / total if total > 0 else 0.0), n_acc, total
            # proposals: mean + z @ (u sqrt(s))^T with z drawn exactly like N*K successive multivariate_normal calls
            factor = _mvn_factor(cov_matrix)
            z = rng.standard_normal((n_particles * K, n_params))
            props = np.repeat(base_theta, K, axis=0) + z @ factor.T
            props = reflect_into_bounds(props, lows[None, :], highs[None, :])
            lp_props = log_prior_batch(props)
            ok = np.isfinite(lp_props)
            ll_props = np.full(n_particles * K, np.nan)
            if ok.any():
                ll_props[ok] = evaluate_particles_parallel(log_likelihood, props[ok], n_jobs=n_jobs, use_threads=use_threads)
            props = props.reshape(n_particles, K, n_params)
            ll_props = ll_props.reshape(n_particles, K)
            ok2 = ok.reshape(n_particles, K)
            lp2 = lp_props.reshape(n_particles, K)
            cur_theta = base_theta.copy()
            cur_logL = base_logL.copy()
            cur_lp = log_prior_batch(cur_theta)
            total = int(ok.sum())
            draws = np.isfinite(ll_props) & ok2  # one uniform per evaluated proposal with finite logL (TM:931-939)
            u = np.full(n_particles * K, np.nan)
            u[draws.reshape(-1)] = rng.random(int(draws.sum()))
            u = u.reshape(n_particles, K)
            n_acc = 0
            for k in range(K):
                with np.errstate(divide="ignore", invalid="ignore"):
                    log_ratio = (lp2[:, k] + beta_next * ll_props[:, k]) - (cur_lp + beta_next * cur_logL)
                    acc = draws[:, k] & (np.log(u[:, k]) < log_ratio)
                cur_theta[acc] = props[acc, k]
                cur_logL[acc] = ll_props[acc, k]
                cur_lp[acc] = lp2[acc, k]
                n_acc += int(acc.sum())
            rate = n_acc / total if total > 0 else 0.0
            return cur_theta, cur_logL, rate, n_acc, total

        adaptive_n_steps = n_mutation_steps
        if acc_rate_history:
            prev_acc = float(acc_rate_history[-1])
            if prev_acc > 0.01:
                with np.errstate(divide="ignore"):  # prev_acc == 1 -> log(0) = -inf -> one step, as in the reference
                    adaptive_n_steps = min(n_mutation_steps * 3, max(1, int(np.ceil(np.log(0.01) / np.log(1.0 - prev_acc)))))
        theta, logL, acc_rate, acc, total_proposals = mutate(cov, adaptive_n_steps)

        # recovery when the mutation is stuck (TM:998-1033)
        if acc_rate < min_acc:
            theta, logL, acc_rate, acc, total_proposals = mutate(cov * 0.3, max(1, n_mutation_steps // 2))
        if acc_rate < min_acc:
            theta_after_resample = theta_after_resample + rng.normal(loc=0.0, scale=1e-3, size=theta_after_resample.shape)
            theta_after_resample = reflect_into_bounds(theta_after_resample, lows[None, :], highs[None, :])
            logL_after_resample = evaluate_particles_parallel(log_likelihood, theta_after_resample, n_jobs=n_jobs, use_threads=use_threads)
            theta, logL, acc_rate, acc, total_proposals = mutate(cov * 0.1, 1)
        if acc_rate < min_acc:
            raise RuntimeError(
                f"TMCMC mutation stuck: acc_rate={acc_rate:.4f} < {min_acc:.4f} after recovery attempts. "
                f"Stage={stage}, beta_next={beta_next:.4f}."
            )

        # 4. linearization hooks: the protocol run_TMCMC drives on its evaluator (TM:557-562, 1436-1441, 1561).  The evaluators
        # of this package have no linear ROM -- compute_ROM_error is 0 while linearization is off and +inf if someone switches it
        # on -- so the reference's bookkeeping reduces to: remember the MAP of the tempered posterior (it is what the run
        # reports as theta_MAP), probe the ROM error at it, and leave the linearization point alone.  An evaluator that DOES
        # report a finite error above the threshold gets the reference's remedy in its plainest form: move the point towards
        # the MAP in bounded steps and re-evaluate the population after each move.  The reference's observation-based
        # re-scoring of a random particle subset (TM:1130-1260) only matters with a non-zero ROM error and is not carried.
        if evaluator is not None and theta_base_full is not None and active_indices is not None:
            if (beta_next > 0.5 and stage % update_linearization_interval == 0) or stage == n_stages:
                theta_MAP_posterior_history.append(theta_before[int(np.argmax(log_post_before))].copy())
                best = theta[int(np.argmax(log_prior_batch(theta) + beta_next * logL))]
                theta_full_MAP = theta_base_full.copy()
                for i, ai in enumerate(active_indices):
                    theta_full_MAP[ai] = best[i]
                theta0 = evaluator.get_linearization_point()
                if theta0 is not None:
                    delta_theta0_stage = float(np.linalg.norm(theta_full_MAP - theta0))
                if theta0 is None or delta_theta0_stage >= THETA_CONVERGENCE_THRESHOLD:
                    err = evaluator.compute_ROM_error(theta_full_MAP)
                    rom_error_pre_stage = float(err)
                    rom_error_pre_history.append(err)
                    moves = 0
                    while (theta0 is not None and np.isfinite(err) and err >= ROM_ERROR_THRESHOLD and moves < MAX_LINEARIZATION_SUBUPDATES_PER_EVENT
                           and n_linearization_updates < MAX_LINEARIZATION_UPDATES):
                        gap = theta_full_MAP - theta0
                        dist_ = float(np.linalg.norm(gap))
                        if not np.isfinite(dist_) or dist_ <= 1e-12:
                            break
                        theta0 = theta0 + min(1.0, MAX_THETA0_STEP_NORM / dist_) * gap
                        evaluator.update_linearization_point(theta0)
                        theta0_history.append(theta0.copy())
                        n_linearization_updates += 1
                        moves += 1
                        logL = evaluate_particles_parallel(log_likelihood, theta, n_jobs=n_jobs, use_threads=use_threads)  # TM:1587-1591
                        err = evaluator.compute_ROM_error(theta_full_MAP)
                        rom_error_history.append(err)
                        rom_error_post_stage = float(err)

        acc_rate_history.append(acc_rate)
        stage_summary.append({
            "stage": int(stage), "beta": float(beta), "beta_next": float(beta_next), "delta_beta": float(delta_beta),
            "ess": float(ess_at_lo) if ess_at_lo is not None else None,
            "ess_target": float(target_ess_ratio * n_particles), "acc_rate": float(acc_rate),
            "logL_min": float(np.min(logL)) if len(logL) > 0 else None,
            "logL_max": float(np.max(logL)) if len(logL) > 0 else None,
            "linearization_enabled": int(bool(getattr(evaluator, "_linearization_enabled", False))) if evaluator is not None else 0,
            "rom_error_pre": rom_error_pre_stage, "rom_error_post": rom_error_post_stage, "delta_theta0": delta_theta0_stage,
            "n_accepted": int(acc), "n_proposals": int(total_proposals),
        })
        logger.info("[TMCMC] %s stage %d: beta %.4f -> %.4f, acc=%.3f", model_name, stage, beta, beta_next, acc_rate)
        beta = beta_next
        beta_sched.append(beta)
        if beta >= BETA_CONVERGENCE_THRESHOLD:
            break

    # final MAP (TM:1730-1752)
    if theta_MAP_posterior_history:
        theta_MAP = theta_MAP_posterior_history[-1].copy()
    else:
        theta_MAP = theta[int(np.argmax(log_prior_batch(theta) + beta * logL))]

    final_MAP = None
    if theta0_history:
        final_MAP = theta0_history[-1].copy()
    elif evaluator is not None:
        final_MAP = evaluator.get_linearization_point()
    n_rom = n_fom = 0
    if evaluator is not None:
        n_rom = evaluator.call_count - initial_rom_count
        n_fom = evaluator.fom_call_count - initial_fom_count
    wall = float(time.perf_counter() - t_start)
    timing_breakdown = None
    if evaluator is not None and isinstance(getattr(evaluator, "timing", None), TimingStats):
        tsm_s = float(evaluator.timing.get_s("tsm.solve_tsm"))
        fom_s = float(evaluator.timing.get_s("fom.run_deterministic"))
        timing_breakdown = {"tmcmc_total_s": wall, "tsm_s": tsm_s, "fom_s": fom_s, "tmcmc_overhead_s": float(max(0.0, wall - tsm_s - fom_s))}
    health = None
    if evaluator is not None and hasattr(evaluator, "get_health"):
        try:
            health = evaluator.get_health()
        except Exception:
            health = None
    return TMCMCResult(
        samples=theta, logL_values=logL, theta_MAP=theta_MAP, beta_schedule=beta_sched,
        converged=(beta >= BETA_CONVERGENCE_THRESHOLD), theta0_history=theta0_history or None,
        n_linearization_updates=n_linearization_updates, final_MAP=final_MAP,
        rom_error_pre_history=rom_error_pre_history or None, rom_error_history=rom_error_history or None,
        acc_rate_history=acc_rate_history or None, n_rom_evaluations=n_rom, n_fom_evaluations=n_fom, wall_time_s=wall,
        timing_breakdown_s=timing_breakdown, likelihood_health=health, stage_summary=stage_summary or None,
        log_evidence=log_evidence,
    )


def _stable_hash_int(text: str) -> int:
    """crc32-based stable hash used for chain seeds (TM:2097-2101)."""
    return int(zlib.crc32(text.encode("utf-8")) & 0x7FFFFFFF)


def _rhat_ess_reference(chains: List[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Reference-only split-free R-hat / ESS indicators over chains (the reference delegates to
    tmcmc/program2602/mcmc_diagnostics.py, host post-processing outside the hot path)."""
    x = np.stack([c[: min(len(c) for c in chains)] for c in chains], axis=0)  # (m, n, d)
    m, n, _ = x.shape
    if m < 2 or n < 2:
        return np.full(x.shape[2], np.nan), np.full(x.shape[2], float(m * n))
    W = x.var(axis=1, ddof=1).mean(axis=0)
    B = n * x.mean(axis=1).var(axis=0, ddof=1)
    var_hat = (n - 1) / n * W + B / n
    with np.errstate(divide="ignore", invalid="ignore"):
        rhat = np.sqrt(var_hat / W)
        ess = m * n * np.minimum(var_hat / np.where(B > 0, B, np.nan), 1.0)
    return rhat, np.where(np.isfinite(ess), ess, float(m * n))


def run_multi_chain_TMCMC(
    model_tag: str,
    make_evaluator: Callable,
    prior_bounds: List[Tuple[float, float]],
    theta_base_full: np.ndarray,
    active_indices: List[int],
    theta_linearization_init: Optional[np.ndarray] = None,
    n_particles: int = 2000,
    n_stages: int = 30,
    target_ess_ratio: float = 0.5,
    min_delta_beta: float = 0.05,
    max_delta_beta: float = 0.2,
    logL_scale: float = 1.0,
    n_chains: int = 1,
    update_linearization_interval: int = 3,
    n_mutation_steps: int = DEFAULT_N_MUTATION_STEPS,
    use_observation_based_update: bool = True,
    linearization_threshold: float = DEFAULT_LINEARIZATION_THRESHOLD,
    linearization_enable_rom_threshold: float = 0.05,
    debug_config: Optional[Any] = None,
    seed: Optional[int] = None,
    force_beta_one: bool = False,
    n_jobs: Optional[int] = None,
    use_threads: bool = False,
    gnn_prior: Optional[Any] = None,
    stage_ops: Optional[Any] = None,
    device_mutation: bool = False,
):
    """Sequential chains with the reference's seeding and MAP hand-over (TM:2104-2381).
    Returns (all_samples, all_logL, global_MAP, converged_flags, diagnostics)."""
    if theta_linearization_init is None:
        theta_linearization_init = theta_base_full.copy()
    all_samples, all_logL, converged_flags, results, all_MAPs = [], [], [], [], []
    global_MAP = theta_linearization_init.copy()
    for chain_idx in range(n_chains):
        chain_seed = (int(seed or 0) + _stable_hash_int(model_tag) % (2**31) + n_chains * 1000 + chain_idx) % (2**31)
        current_lin = theta_linearization_init.copy() if chain_idx == 0 else global_MAP.copy()
        evaluator = make_evaluator(theta_linearization=current_lin)
        res = run_TMCMC(
            log_likelihood=evaluator, prior_bounds=prior_bounds, n_particles=n_particles, n_stages=n_stages,
            target_ess_ratio=target_ess_ratio, min_delta_beta=min_delta_beta, max_delta_beta=max_delta_beta,
            logL_scale=logL_scale, seed=chain_seed, model_name=f"{model_tag}_chain{chain_idx + 1}", n_jobs=n_jobs,
            use_threads=use_threads, evaluator=evaluator, theta_base_full=theta_base_full, active_indices=active_indices,
            update_linearization_interval=update_linearization_interval, n_mutation_steps=n_mutation_steps,
            use_observation_based_update=use_observation_based_update, linearization_threshold=linearization_threshold,
            linearization_enable_rom_threshold=linearization_enable_rom_threshold, force_beta_one=force_beta_one,
            gnn_prior=gnn_prior, stage_ops=stage_ops, device_mutation=device_mutation,
        )
        all_samples.append(res.samples)
        all_logL.append(res.logL_values)
        converged_flags.append(res.converged)
        results.append(res)
        if res.final_MAP is not None:
            all_MAPs.append(res.final_MAP.copy())
            global_MAP = all_MAPs[-1].copy()
    best = max(range(len(all_logL)), key=lambda c: float(np.max(all_logL[c])))
    global_MAP = all_samples[best][int(np.argmax(all_logL[best]))]
    rhat, ess = _rhat_ess_reference(all_samples)
    diagnostics = {
        "Rhat_reference": rhat, "ESS_reference": ess, "converged_chains": sum(converged_flags), "total_chains": n_chains,
        "MAP_global": global_MAP, "beta_schedules": [r.beta_schedule for r in results],
        "theta0_history": [r.theta0_history for r in results if r.theta0_history is not None],
        "total_linearization_updates": sum(r.n_linearization_updates for r in results),
        "rom_error_pre_histories": [r.rom_error_pre_history for r in results if r.rom_error_pre_history is not None],
        "rom_error_histories": [r.rom_error_history for r in results if r.rom_error_history is not None],
        "acc_rate_histories": [r.acc_rate_history for r in results if r.acc_rate_history is not None],
        "n_rom_evaluations": [r.n_rom_evaluations for r in results],
        "n_fom_evaluations": [r.n_fom_evaluations for r in results],
        "tmcmc_wall_time_s": [float(r.wall_time_s) for r in results],
        "timing_breakdown_s": [r.timing_breakdown_s for r in results],
        "likelihood_health_histories": [r.likelihood_health for r in results if r.likelihood_health is not None],
        "stage_summaries": [r.stage_summary for r in results if r.stage_summary is not None],
        "log_evidence": [r.log_evidence for r in results],
        "note": "R-hat/ESS are reference indicators only; TMCMC convergence is judged by beta = 1.0.",
    }
    total: Dict[str, int] = {}
    for h in diagnostics["likelihood_health_histories"]:
        for k, v in h.items():
            total[k] = int(total.get(k, 0) + int(v))
    if total:
        diagnostics["likelihood_health_total"] = total
    return all_samples, all_logL, global_MAP, converged_flags, diagnostics
