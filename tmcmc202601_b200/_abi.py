"""ctypes binding of the C ABI declared in include/hmx.h (libhamilton.so).

There is no CPU fallback: `load()` raises if the CUDA library has not been built, and `hmx_create`
fails when no sm_100 GPU is present.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Iterable, Optional, Sequence

import numpy as np

ABI_VERSION = 1
NSTATE = 12
NTHETA = 20
MAX_OBS = 8
MAX_COND = 8

ST_NONFINITE = 1
ST_MAXITER = 2
ST_SINGULAR = 4
ST_BAD_COND = 8

LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libhamilton.so"


class HmxCond(C.Structure):
    _fields_ = [
        ("phi_init", C.c_double * 5),
        ("n_obs", C.c_int32),
        ("idx_sparse", C.c_int32 * MAX_OBS),
        ("data", C.c_double * (MAX_OBS * 5)),
        ("sigma_obs", C.c_double * (MAX_OBS * 5)),
        ("weights", C.c_double * (MAX_OBS * 5)),
        ("cov_rel", C.c_double),
        ("active_theta_mask", C.c_uint32),
        ("use_absolute_volume", C.c_int32),
        ("tsm_variance", C.c_int32),
    ]


class HmxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("dt", C.c_double),
        ("n_steps", C.c_int32),
        ("eps", C.c_double),
        ("Kp1", C.c_double),
        ("c_const", C.c_double),
        ("alpha_const", C.c_double),
        ("K_hill", C.c_double),
        ("n_hill", C.c_double),
        ("max_newton_iter", C.c_int32),
        ("eta", C.c_double * 5),
        ("eta_phi", C.c_double * 5),
        ("active_species_mask", C.c_int32),
        ("n_cond", C.c_int32),
        ("cond", HmxCond * MAX_COND),
    ]


class HmxCounters(C.Structure):
    _fields_ = [
        ("evals", C.c_int64),
        ("newton_iters", C.c_int64),
        ("q_evals", C.c_int64),
        ("kernel_launches", C.c_int64),
        ("last_ode_kernel_ms", C.c_double),
    ]


class HmxLegacyConfig(C.Structure):
    _fields_ = [
        ("n_species", C.c_int32), ("dt", C.c_double), ("n_steps", C.c_int32), ("eps", C.c_double), ("Kp1", C.c_double),
        ("c_const", C.c_double), ("alpha_const", C.c_double), ("max_newton_iter", C.c_int32), ("eta", C.c_double * 5),
        ("eta_phi", C.c_double * 5), ("phi_init", C.c_double * 5), ("active_species_mask", C.c_int32),
        ("alpha_switch_level", C.c_int32), ("alpha_before", C.c_double), ("alpha_after", C.c_double),
    ]


def make_legacy_config(n_species=4, dt=1e-5, maxtimestep=500, eps=1e-6, Kp1=1e-4, eta_vec=None, eta_phi_vec=None, c_const=100.0,
                       alpha_const=100.0, phi_init=0.2, active_species=None, max_newton_iter=50, alpha_switch_level=-1,
                       alpha_before=None, alpha_after=None) -> HmxLegacyConfig:
    """Keyword names and defaults of BiofilmNewtonSolver.__init__ (S4:737-781); the alpha schedule resolved to a level."""
    if n_species not in (4, 5):
        raise ValueError("n_species must be 4 or 5")
    c = HmxLegacyConfig()
    c.n_species = n_species
    c.dt, c.n_steps, c.eps, c.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
    c.c_const, c.alpha_const, c.max_newton_iter = float(c_const), float(alpha_const), int(max_newton_iter)
    eta = np.ones(n_species) if eta_vec is None else np.asarray(eta_vec, dtype=float)
    etaphi = np.ones(n_species) if eta_phi_vec is None else np.asarray(eta_phi_vec, dtype=float)
    pi = np.asarray(phi_init, dtype=float)
    pi = np.full(n_species, float(pi.flat[0])) if pi.ndim == 0 or pi.size == 1 else np.resize(pi, n_species)
    for i in range(n_species):
        c.eta[i], c.eta_phi[i], c.phi_init[i] = eta[i], etaphi[i], pi[i]
    act = list(range(n_species)) if active_species is None else list(active_species)
    c.active_species_mask = sum(1 << i for i in set(act) if 0 <= i < n_species)
    c.alpha_switch_level = int(alpha_switch_level)
    c.alpha_before = float(alpha_const if alpha_before is None else alpha_before)
    c.alpha_after = float(alpha_const if alpha_after is None else alpha_after)
    return c


MUT_MAX_DIM = 32
MUT_MAX_GROUPS = 64


class HmxMutation(C.Structure):
    _fields_ = [
        ("d", C.c_int32),
        ("K", C.c_int32),
        ("factor", C.c_double * (MUT_MAX_DIM * MUT_MAX_DIM)),
        ("low", C.c_double * MUT_MAX_DIM),
        ("high", C.c_double * MUT_MAX_DIM),
        ("theta_base", C.c_double * NTHETA),
        ("map", C.c_int32 * MUT_MAX_DIM),
        ("n_map", C.c_int32),
        ("theta_lin", C.c_double * NTHETA),
        ("cond_default", C.c_int32),
        ("cond_at_lin", C.c_int32),
        ("beta", C.c_double),
        ("seed", C.c_uint64),
        ("sequence", C.c_uint64),
        ("index_offset", C.c_uint64),
    ]


class HmxMutationStats(C.Structure):
    _fields_ = [
        ("n_accepted", C.c_int64),
        ("n_candidates", C.c_int64),
        ("n_nonfinite", C.c_int64),
        ("n_maxiter", C.c_int64),
        ("n_singular", C.c_int64),
    ]


def make_mutation(factor, lows, highs, K, beta, seed, sequence=0, index_offset=0, theta_base=None, active_map=None,
                  theta_lin=None, cond_default=0, cond_at_lin=-1) -> HmxMutation:
    """Fill an hmx_mutation.  `active_map[j]` is the position of theta_sub[j] in the 20-vector (default: identity);
    entries >= 20 (the viscoelastic parameters) do not reach the ODE."""
    factor = np.ascontiguousarray(factor, dtype=np.float64)
    d = factor.shape[0]
    if factor.shape != (d, d) or not 1 <= d <= MUT_MAX_DIM:
        raise ValueError(f"factor must be square with 1 <= d <= {MUT_MAX_DIM}")
    lows = np.asarray(lows, dtype=np.float64).reshape(-1)
    highs = np.asarray(highs, dtype=np.float64).reshape(-1)
    if lows.size != d or highs.size != d:
        raise ValueError("bounds and factor disagree on the dimension")
    m = HmxMutation()
    m.d, m.K = d, int(K)
    for i in range(d * d):
        m.factor[i] = factor.flat[i]
    for j in range(d):
        m.low[j], m.high[j] = lows[j], highs[j]
    base = np.zeros(NTHETA) if theta_base is None else np.asarray(theta_base, dtype=np.float64).reshape(-1)[:NTHETA]
    amap = list(range(min(d, NTHETA))) if active_map is None else [int(a) for a in active_map][:d]
    for k in range(NTHETA):
        m.theta_base[k] = base[k] if k < base.size else 0.0
    for j, a in enumerate(amap):
        m.map[j] = a
    m.n_map = len(amap)
    lin = base if theta_lin is None else np.asarray(theta_lin, dtype=np.float64).reshape(-1)[:NTHETA]
    for k in range(NTHETA):
        m.theta_lin[k] = lin[k] if k < lin.size else 0.0
    m.cond_default, m.cond_at_lin = int(cond_default), int(cond_at_lin)
    m.beta = float(beta)
    m.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    m.sequence = int(sequence)
    m.index_offset = int(index_offset)
    return m


EXPORTS = [
    "hmx_create", "hmx_destroy", "hmx_last_error", "hmx_set_stream", "hmx_synchronize", "hmx_abi_version",
    "hmx_loglik_batch", "hmx_trajectory_batch", "hmx_beta_step", "hmx_weights", "hmx_resample",
    "hmx_weighted_cov", "hmx_weighted_moments", "hmx_weighted_scatter", "hmx_log_likelihood_sparse", "hmx_get_counters",
    "hmx_set_timing",
    "hmx_measure_fp64_peak",
    "hmx_mutate", "hmx_mutate_groups", "hmx_trajectory_rows", "hmx_set_latency_kernel", "hmx_cumsum",
    "hmx_tsm_obs_rows", "hmx_legacy_trajectory_rows", "hmx_loglik_grad_batch", "hmx_solve_tsm_batch", "hmx_set_sig2_guard",
    "hmx_multi_create", "hmx_multi_destroy", "hmx_multi_last_error", "hmx_multi_device_count", "hmx_multi_context",
    "hmx_multi_loglik_batch", "hmx_multi_weighted_cov",
    "hmx_population_create", "hmx_population_destroy", "hmx_population_upload", "hmx_population_download",
]


def broadcast_phi_init(phi_init) -> np.ndarray:
    """Scalar / array phi_init exactly as BiofilmNewtonSolver5S.__init__ broadcasts it (S5:489-495)."""
    pi = np.asarray(phi_init, dtype=float)
    if pi.ndim == 0 or pi.size == 1:
        return np.full(5, float(pi.flat[0]))
    pi = np.asarray(pi, dtype=float)
    return np.resize(pi, 5) if pi.size != 5 else pi.reshape(5).copy()


def broadcast_sigma(sigma_obs, n_obs: int) -> np.ndarray:
    """sigma_obs scalar / (5,) / (n_obs,5) -> (n_obs,5), as log_likelihood_sparse does (EV:286-290,309)."""
    sig = np.atleast_1d(np.asarray(sigma_obs, dtype=np.float64))
    if sig.ndim == 1 and sig.size == 1:
        sig = np.full(5, sig[0])
    if sig.ndim == 1:
        sig = np.tile(sig[None, :], (n_obs, 1))
    if sig.shape != (n_obs, 5):
        raise ValueError(f"sigma_obs shape {sig.shape} incompatible with data ({n_obs},5)")
    return sig


def make_cond(
    data,
    idx_sparse,
    sigma_obs,
    phi_init,
    cov_rel: float = 0.005,
    active_theta_indices: Optional[Iterable[int]] = None,
    weights=None,
    use_absolute_volume: bool = False,
    tsm_variance: bool = True,
) -> HmxCond:
    data = np.ascontiguousarray(data, dtype=np.float64)
    if data.ndim != 2 or data.shape[1] != 5:
        raise ValueError("data must have shape (n_obs, 5)")
    n_obs = data.shape[0]
    if n_obs > MAX_OBS:
        raise ValueError(f"at most {MAX_OBS} observation rows are supported")
    idx = np.asarray(idx_sparse, dtype=np.int64).reshape(-1)
    if idx.size != n_obs:
        raise ValueError("idx_sparse and data disagree on n_obs")
    sig = broadcast_sigma(sigma_obs, n_obs)
    w = np.ones((n_obs, 5)) if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
    if w.shape != (n_obs, 5):
        raise ValueError("weights must have shape (n_obs, 5)")
    cc = HmxCond()
    pi = broadcast_phi_init(phi_init)
    for i in range(5):
        cc.phi_init[i] = pi[i]
    cc.n_obs = n_obs
    for i in range(n_obs):
        cc.idx_sparse[i] = int(idx[i])
    for i in range(n_obs * 5):
        cc.data[i] = data.flat[i]
        cc.sigma_obs[i] = sig.flat[i]
        cc.weights[i] = w.flat[i]
    cc.cov_rel = float(cov_rel)
    act = range(NTHETA) if active_theta_indices is None else active_theta_indices
    cc.active_theta_mask = sum(1 << int(k) for k in set(act) if 0 <= int(k) < NTHETA)
    cc.use_absolute_volume = int(bool(use_absolute_volume))
    cc.tsm_variance = int(bool(tsm_variance))
    return cc


def make_config(
    conds: Sequence[HmxCond],
    dt=1e-5,
    maxtimestep=500,
    eps=1e-6,
    Kp1=1e-4,
    eta_vec=None,
    eta_phi_vec=None,
    c_const=100.0,
    alpha_const=100.0,
    max_newton_iter=50,
    K_hill=0.0,
    n_hill=2.0,
    active_species=None,
) -> HmxConfig:
    """Keyword names and defaults of BiofilmNewtonSolver5S.__init__ (S5:464-481); phi_init lives per condition."""
    if not 1 <= len(conds) <= MAX_COND:
        raise ValueError(f"between 1 and {MAX_COND} conditions")
    cfg = HmxConfig()
    cfg.abi_version = ABI_VERSION
    cfg.dt, cfg.n_steps, cfg.eps, cfg.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
    cfg.c_const, cfg.alpha_const = float(c_const), float(alpha_const)
    cfg.K_hill, cfg.n_hill, cfg.max_newton_iter = float(K_hill), float(n_hill), int(max_newton_iter)
    eta = np.ones(5) if eta_vec is None else np.asarray(eta_vec, dtype=float)
    etaphi = np.ones(5) if eta_phi_vec is None else np.asarray(eta_phi_vec, dtype=float)
    for i in range(5):
        cfg.eta[i], cfg.eta_phi[i] = eta[i], etaphi[i]
    act = [0, 1, 2, 3, 4] if active_species is None else list(active_species)
    cfg.active_species_mask = sum(1 << i for i in act if 0 <= i < 5)
    cfg.n_cond = len(conds)
    for k, c in enumerate(conds):
        cfg.cond[k] = c
    return cfg


_lib = None


def load() -> C.CDLL:
    """Load libhamilton.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    import os

    lib_path = Path(os.environ.get("HMX_LIB", str(LIB_PATH)))  # HMX_LIB: kernel-variant experiments only
    if not lib_path.exists():
        raise RuntimeError(
            f"{lib_path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
            "g.build()').  This package has no CPU fallback."
        )
    lib = C.CDLL(str(lib_path))
    dp, ip, up, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int64)
    vp = C.c_void_p
    lib.hmx_create.argtypes = [C.POINTER(vp), C.POINTER(HmxConfig), C.c_int32]
    lib.hmx_destroy.argtypes = [vp]
    lib.hmx_destroy.restype = None
    lib.hmx_last_error.argtypes = [vp]
    lib.hmx_last_error.restype = C.c_char_p
    lib.hmx_set_stream.argtypes = [vp, vp]
    lib.hmx_synchronize.argtypes = [vp]
    # array arguments are raw addresses (host or device), hence c_void_p everywhere
    lib.hmx_loglik_batch.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp]
    lib.hmx_trajectory_batch.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp]
    lib.hmx_population_create.argtypes = [vp, C.c_int64, C.c_int32, C.POINTER(vp)]
    lib.hmx_population_destroy.argtypes = [vp, vp]
    lib.hmx_population_upload.argtypes = [vp, vp, vp, C.c_int64]
    lib.hmx_population_download.argtypes = [vp, vp, vp, C.c_int64]
    lib.hmx_legacy_trajectory_rows.argtypes = [vp, C.POINTER(HmxLegacyConfig), vp, C.c_int64, vp, C.c_int32, vp, vp, vp]
    lib.hmx_multi_create.argtypes = [C.POINTER(vp), C.POINTER(HmxConfig), vp, C.c_int32]
    lib.hmx_multi_destroy.argtypes = [vp]
    lib.hmx_multi_destroy.restype = None
    lib.hmx_multi_last_error.argtypes = [vp]
    lib.hmx_multi_last_error.restype = C.c_char_p
    lib.hmx_multi_device_count.argtypes = [vp]
    lib.hmx_multi_context.argtypes = [vp, C.c_int32]
    lib.hmx_multi_context.restype = vp
    lib.hmx_multi_loglik_batch.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp]
    lib.hmx_multi_weighted_cov.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
    lib.hmx_solve_tsm_batch.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp]
    lib.hmx_set_sig2_guard.argtypes = [vp, C.c_int32]
    lib.hmx_loglik_grad_batch.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp]
    lib.hmx_tsm_obs_rows.argtypes = [vp, vp, vp, C.c_int64, vp, vp, vp, vp, vp]
    lib.hmx_trajectory_rows.argtypes = [vp, vp, vp, C.c_int64, vp, C.c_int32, vp, vp, vp]
    lib.hmx_beta_step.argtypes = [vp, vp, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, vp, vp]
    lib.hmx_weights.argtypes = [vp, vp, C.c_int64, C.c_double, vp, vp]
    lib.hmx_resample.argtypes = [vp, vp, C.c_int64, C.c_double, vp]
    lib.hmx_cumsum.argtypes = [vp, vp, C.c_int64, vp]
    lib.hmx_weighted_cov.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
    lib.hmx_weighted_moments.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp]
    lib.hmx_weighted_scatter.argtypes = [vp, vp, vp, C.c_int64, C.c_int32, vp, vp]
    lib.hmx_log_likelihood_sparse.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int32, vp]
    lib.hmx_get_counters.argtypes = [vp, C.POINTER(HmxCounters)]
    lib.hmx_set_timing.argtypes = [vp, C.c_int32]
    lib.hmx_set_latency_kernel.argtypes = [vp, C.c_int64]
    lib.hmx_measure_fp64_peak.argtypes = [vp, dp]
    lib.hmx_mutate.argtypes = [vp, C.POINTER(HmxMutation), vp, vp, C.c_int64, vp, vp, vp, vp, C.POINTER(HmxMutationStats)]
    lib.hmx_mutate_groups.argtypes = [vp, C.c_int32, C.POINTER(HmxMutation), lp, vp, vp, vp, vp, vp, vp, C.POINTER(HmxMutationStats)]
    del ip, up, lp
    if lib.hmx_abi_version() != ABI_VERSION:
        raise RuntimeError("libhamilton.so ABI version mismatch; rebuild")
    _lib = lib
    return lib
