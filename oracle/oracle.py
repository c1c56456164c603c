"""ctypes front-end of the CPU oracle (oracle/hamilton_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (tmcmc202601_b200) never imports it.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libhamilton_oracle.so"

MAX_OBS = 8
NTHETA = 20
NSTATE = 12

ST_NONFINITE = 1
ST_MAXITER = 2
ST_SINGULAR = 4


class SolverCfg(C.Structure):
    _fields_ = [
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
        ("phi_init", C.c_double * 5),
        ("active_mask", C.c_int32),
    ]


class CondCfg(C.Structure):
    _fields_ = [
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


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only; a couple of seconds)."""
    srcs = [_HERE / "hamilton_oracle.c", _HERE / "hamilton_oracle4.c", _HERE / "hamilton_oracle_grad.c", _HERE / "hamilton_oracle.h"]
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < max(p.stat().st_mtime for p in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        dp = C.POINTER(C.c_double)
        _lib.orc_evaluate.restype = C.c_double
        _lib.orc_log_likelihood_sparse.restype = C.c_double
        _lib.orc_newton_step.restype = C.c_int
        _lib.orc_dgesv.restype = C.c_int
        _lib.orc_weights.restype = C.c_int
        _lib.orc_loglik_grad.restype = C.c_double
        del dp
    return _lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_solver_cfg(
    dt=1e-4,
    maxtimestep=2500,
    eps=1e-6,
    Kp1=1e-4,
    c_const=25.0,
    alpha_const=0.0,
    K_hill=0.05,
    n_hill=4.0,
    max_newton_iter=50,
    eta_vec=None,
    eta_phi_vec=None,
    phi_init=0.02,
    active_species=None,
) -> SolverCfg:
    """Same keyword names as BiofilmNewtonSolver5S.__init__ (S5:464-511)."""
    cfg = SolverCfg()
    cfg.dt, cfg.n_steps, cfg.eps, cfg.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
    cfg.c_const, cfg.alpha_const = float(c_const), float(alpha_const)
    cfg.K_hill, cfg.n_hill, cfg.max_newton_iter = float(K_hill), float(n_hill), int(max_newton_iter)
    eta = np.ones(5) if eta_vec is None else _f64(eta_vec)
    etaphi = np.ones(5) if eta_phi_vec is None else _f64(eta_phi_vec)
    pi = np.asarray(phi_init, dtype=float)
    pi = np.full(5, float(pi.flat[0])) if pi.ndim == 0 or pi.size == 1 else np.resize(pi, 5)
    for i in range(5):
        cfg.eta[i], cfg.eta_phi[i], cfg.phi_init[i] = eta[i], etaphi[i], pi[i]
    act = [0, 1, 2, 3, 4] if active_species is None else list(active_species)
    cfg.active_mask = sum(1 << i for i in act if 0 <= i < 5)
    return cfg


def make_cond_cfg(
    data,
    idx_sparse,
    sigma_obs,
    cov_rel=0.005,
    active_theta_indices=None,
    weights=None,
    use_absolute_volume=False,
    tsm_variance=True,
) -> CondCfg:
    """Observation constants of one condition, broadcast the way log_likelihood_sparse does (EV:286-290)."""
    data = _f64(data)
    n_obs = data.shape[0]
    assert data.shape == (n_obs, 5) and n_obs <= MAX_OBS
    sig = np.atleast_1d(np.asarray(sigma_obs, dtype=np.float64))
    if sig.ndim == 1 and sig.size == 1:
        sig = np.full(5, sig[0])
    if sig.ndim == 1:
        sig = np.tile(sig[None, :], (n_obs, 1))
    w = np.ones((n_obs, 5)) if weights is None else _f64(weights)
    cc = CondCfg()
    cc.n_obs = n_obs
    for i, v in enumerate(np.asarray(idx_sparse, dtype=np.int64)):
        cc.idx_sparse[i] = int(v)
    for i in range(n_obs * 5):
        cc.data[i] = data.flat[i]
        cc.sigma_obs[i] = sig.flat[i]
        cc.weights[i] = w.flat[i]
    cc.cov_rel = float(cov_rel)
    act = range(NTHETA) if active_theta_indices is None else active_theta_indices
    cc.active_theta_mask = sum(1 << int(k) for k in act if int(k) < NTHETA)
    cc.use_absolute_volume = int(bool(use_absolute_volume))
    cc.tsm_variance = int(bool(tsm_variance))
    return cc


def run_deterministic(cfg: SolverCfg, theta):
    """-> (t_arr[T+1], g_arr[T+1,12], n_iters, n_trials, status); S5:561-581."""
    theta = _f64(theta)
    T = cfg.n_steps
    g = np.empty((T + 1, NSTATE))
    its, tr, st = C.c_int64(0), C.c_int64(0), C.c_uint32(0)
    lib().orc_run_deterministic(C.byref(cfg), _dptr(theta), _dptr(g), C.byref(its), C.byref(tr), C.byref(st))
    t_arr = np.concatenate([[0.0], (np.arange(T) + 1) * cfg.dt])
    return t_arr, g, its.value, tr.value, st.value


def compute_Q(cfg: SolverCfg, theta, g_new, g_old):
    A = np.zeros(25)
    b = np.zeros(5)
    theta = _f64(theta)
    lib().orc_theta_to_matrices(_dptr(theta), _dptr(A), _dptr(b))
    g_new, g_old = _f64(g_new).copy(), _f64(g_old)
    phi, psi = g_new[0:5].copy(), g_new[6:11].copy()
    phio, psio = g_old[0:5].copy(), g_old[6:11].copy()
    Q = np.zeros(12)
    lib().orc_compute_Q(
        _dptr(phi), C.c_double(g_new[5]), _dptr(psi), C.c_double(g_new[11]), _dptr(phio),
        C.c_double(g_old[5]), _dptr(psio), C.byref(cfg), _dptr(A), _dptr(b), _dptr(Q),
    )
    return Q


def compute_J(cfg: SolverCfg, theta, g_new, g_old):
    A = np.zeros(25)
    b = np.zeros(5)
    theta = _f64(theta)
    lib().orc_theta_to_matrices(_dptr(theta), _dptr(A), _dptr(b))
    g_new, g_old = _f64(g_new).copy(), _f64(g_old)
    phi, psi = g_new[0:5].copy(), g_new[6:11].copy()
    phio, psio = g_old[0:5].copy(), g_old[6:11].copy()
    K = np.zeros((12, 12))
    lib().orc_compute_J(
        _dptr(phi), C.c_double(g_new[5]), _dptr(psi), C.c_double(g_new[11]), _dptr(phio),
        _dptr(psio), C.byref(cfg), _dptr(A), _dptr(b), _dptr(K),
    )
    return K


def sensitivity_row(cfg: SolverCfg, theta, g_new, g_old, active_theta_mask=(1 << NTHETA) - 1):
    theta, g_new, g_old = _f64(theta), _f64(g_new), _f64(g_old)
    x1 = np.zeros((NSTATE, NTHETA))
    st = C.c_uint32(0)
    lib().orc_sensitivity_row(
        C.byref(cfg), _dptr(theta), _dptr(g_new), _dptr(g_old), C.c_uint32(active_theta_mask),
        _dptr(x1), C.byref(st),
    )
    return x1


def evaluate(cfg: SolverCfg, cond: CondCfg, theta, want_obs=False):
    theta = _f64(theta)
    obs = np.zeros((cond.n_obs, NSTATE))
    its, tr, st = C.c_int64(0), C.c_int64(0), C.c_uint32(0)
    ll = lib().orc_evaluate(
        C.byref(cfg), C.byref(cond), _dptr(theta), _dptr(obs), C.byref(st), C.byref(its), C.byref(tr)
    )
    if want_obs:
        return ll, obs, st.value, its.value, tr.value
    return ll


def evaluate_batch(cfg: SolverCfg, conds, theta, cond_id=None, n_threads=None, details=False):
    """logL[N] for theta[N,20]; conds is a list of CondCfg, cond_id[N] int32 or None (=0)."""
    theta = _f64(theta)
    N = theta.shape[0]
    arr = (CondCfg * len(conds))(*conds)
    logL = np.empty(N)
    status = np.zeros(N, dtype=np.uint32)
    its = np.zeros(N, dtype=np.int64)
    tr = np.zeros(N, dtype=np.int64)
    cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
    nt = n_threads or os.cpu_count() or 1
    lib().orc_evaluate_batch(
        C.byref(cfg), arr, C.c_int(len(conds)), _dptr(theta),
        None if cid is None else cid.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int64(N),
        _dptr(logL), status.ctypes.data_as(C.POINTER(C.c_uint32)),
        its.ctypes.data_as(C.POINTER(C.c_int64)), tr.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int(nt),
    )
    if details:
        return logL, status, its, tr
    return logL


def log_likelihood_sparse(mu, sig, data, sigma_obs, weights=None):
    data = _f64(data)
    n_obs = data.shape[0]
    cc = make_cond_cfg(data, np.zeros(n_obs, dtype=int), sigma_obs, weights=weights)
    mu, sig = _f64(mu), _f64(np.broadcast_to(sig, data.shape))
    so = np.array(cc.sigma_obs[: n_obs * 5])
    w = np.array(cc.weights[: n_obs * 5])
    return lib().orc_log_likelihood_sparse(_dptr(mu), _dptr(sig), _dptr(data), _dptr(so), _dptr(w), C.c_int(n_obs))


def beta_step(logL, beta, target_ess_ratio=0.5, min_db=0.02, max_db=0.5):
    logL = _f64(logL)
    db, ess = C.c_double(0), C.c_double(0)
    lib().orc_beta_step(
        _dptr(logL), C.c_int64(logL.size), C.c_double(beta), C.c_double(target_ess_ratio),
        C.c_double(min_db), C.c_double(max_db), C.byref(db), C.byref(ess),
    )
    return db.value, ess.value


def weights(logL, delta_beta):
    logL = _f64(logL)
    w = np.empty_like(logL)
    ls = C.c_double(0)
    flag = lib().orc_weights(_dptr(logL), C.c_int64(logL.size), C.c_double(delta_beta), _dptr(w), C.byref(ls))
    return w, ls.value, flag


def resample(w, u):
    w = _f64(w)
    idx = np.empty(w.size, dtype=np.int64)
    lib().orc_resample(_dptr(w), C.c_int64(w.size), C.c_double(u), idx.ctypes.data_as(C.POINTER(C.c_int64)))
    return idx


def weighted_cov(theta, w):
    theta, w = _f64(theta), _f64(w)
    N, d = theta.shape
    mean, cov = np.empty(d), np.empty((d, d))
    lib().orc_weighted_cov(_dptr(theta), _dptr(w), C.c_int64(N), C.c_int(d), _dptr(mean), _dptr(cov))
    return mean, cov


# ---- likelihood gradient (hamilton_oracle_grad.c) --------------------------------------------------------------------
def loglik_grad(cfg: SolverCfg, cond: CondCfg, theta, want_dstate=False):
    """(logL, grad[20][, dstate[n_obs,12,20]]): forward sensitivities with the exact Jacobian, variance frozen."""
    theta = _f64(theta)
    grad = np.zeros(NTHETA)
    ds = np.zeros((cond.n_obs, NSTATE, NTHETA))
    st = C.c_uint32(0)
    ll = lib().orc_loglik_grad(C.byref(cfg), C.byref(cond), _dptr(theta), _dptr(grad), _dptr(ds), C.byref(st))
    return (ll, grad, ds) if want_dstate else (ll, grad)


def exact_jacobians(cfg: SolverCfg, theta, g_new, g_old):
    theta, g_new, g_old = _f64(theta), _f64(g_new), _f64(g_old)
    Jx, Jp = np.zeros((12, 12)), np.zeros((11, 2))
    lib().orc_exact_jacobians(C.byref(cfg), _dptr(theta), _dptr(g_new), _dptr(g_old), _dptr(Jx), _dptr(Jp))
    return Jx, Jp


# ---- legacy 4-species solver (hamilton_oracle4.c) ------------------------------------------------------------------
class Cfg4(C.Structure):
    _fields_ = [
        ("dt", C.c_double), ("n_steps", C.c_int32), ("eps", C.c_double), ("Kp1", C.c_double), ("c_const", C.c_double),
        ("alpha_const", C.c_double), ("max_newton_iter", C.c_int32), ("eta", C.c_double * 4), ("eta_phi", C.c_double * 4),
        ("phi_init", C.c_double), ("active_mask", C.c_int32), ("alpha_switch_level", C.c_int32),
        ("alpha_before", C.c_double), ("alpha_after", C.c_double),
    ]


def make_cfg4(dt=1e-5, maxtimestep=500, eps=1e-6, Kp1=1e-4, eta_vec=None, eta_phi_vec=None, c_const=100.0, alpha_const=100.0,
              phi_init=0.2, active_species=None, max_newton_iter=50, alpha_switch_level=-1, alpha_before=None, alpha_after=None) -> Cfg4:
    """Keyword names of BiofilmNewtonSolver.__init__ (S4:737-781); the schedule already resolved to a switch level."""
    c = Cfg4()
    c.dt, c.n_steps, c.eps, c.Kp1 = float(dt), int(maxtimestep), float(eps), float(Kp1)
    c.c_const, c.alpha_const, c.max_newton_iter = float(c_const), float(alpha_const), int(max_newton_iter)
    eta = np.ones(4) if eta_vec is None else _f64(eta_vec)
    etaphi = np.ones(4) if eta_phi_vec is None else _f64(eta_phi_vec)
    for i in range(4):
        c.eta[i], c.eta_phi[i] = eta[i], etaphi[i]
    c.phi_init = float(phi_init)
    act = [0, 1, 2, 3] if active_species is None else list(active_species)
    c.active_mask = sum(1 << i for i in act if 0 <= i < 4)
    c.alpha_switch_level = int(alpha_switch_level)
    c.alpha_before = float(alpha_const if alpha_before is None else alpha_before)
    c.alpha_after = float(alpha_const if alpha_after is None else alpha_after)
    return c


def run_deterministic4(cfg: Cfg4, theta):
    """-> (t_arr[T+1], g_arr[T+1,10], n_iters, n_trials, status); S4:1005-1028."""
    theta = _f64(theta)
    T = cfg.n_steps
    g = np.empty((T + 1, 10))
    its, tr, st = C.c_int64(0), C.c_int64(0), C.c_uint32(0)
    lib().orc4_run_deterministic(C.byref(cfg), _dptr(theta), _dptr(g), C.byref(its), C.byref(tr), C.byref(st))
    t_arr = np.concatenate([[0.0], (np.arange(T) + 1) * cfg.dt])
    return t_arr, g, its.value, tr.value, st.value


def compute_Q4(cfg: Cfg4, theta, g_new, g_old, alpha=None):
    A, b = np.zeros(16), np.zeros(4)
    theta = _f64(theta)
    lib().orc4_theta_to_matrices(_dptr(theta), _dptr(A), _dptr(b))
    g_new, g_old = _f64(g_new).copy(), _f64(g_old)
    phi, psi = g_new[0:4].copy(), g_new[5:9].copy()
    Q = np.zeros(10)
    lib().orc4_compute_Q(_dptr(phi), C.c_double(g_new[4]), _dptr(psi), C.c_double(g_new[9]), _dptr(g_old[0:4].copy()),
                         C.c_double(g_old[4]), _dptr(g_old[5:9].copy()), C.byref(cfg),
                         C.c_double(cfg.alpha_const if alpha is None else alpha), _dptr(A), _dptr(b), _dptr(Q))
    return Q


def compute_J4(cfg: Cfg4, theta, g_new, g_old, alpha=None):
    A, b = np.zeros(16), np.zeros(4)
    theta = _f64(theta)
    lib().orc4_theta_to_matrices(_dptr(theta), _dptr(A), _dptr(b))
    g_new, g_old = _f64(g_new).copy(), _f64(g_old)
    phi, psi = g_new[0:4].copy(), g_new[5:9].copy()
    K = np.zeros((10, 10))
    lib().orc4_compute_J(_dptr(phi), C.c_double(g_new[4]), _dptr(psi), _dptr(g_old[0:4].copy()), _dptr(g_old[5:9].copy()),
                         C.byref(cfg), C.c_double(cfg.alpha_const if alpha is None else alpha), _dptr(A), _dptr(b), _dptr(K))
    return K


def from_hmx_config(hcfg):
    """(SolverCfg, [CondCfg...]) equivalent to a product-side HmxConfig (tmcmc202601_b200/_abi.py).

    The oracle's solver struct carries ONE phi_init, the product config one per condition; this returns a
    list of (SolverCfg, CondCfg) pairs, one per condition.
    """
    out = []
    for k in range(hcfg.n_cond):
        hc = hcfg.cond[k]
        scfg = make_solver_cfg(
            dt=hcfg.dt, maxtimestep=hcfg.n_steps, eps=hcfg.eps, Kp1=hcfg.Kp1, c_const=hcfg.c_const,
            alpha_const=hcfg.alpha_const, K_hill=hcfg.K_hill, n_hill=hcfg.n_hill,
            max_newton_iter=hcfg.max_newton_iter, eta_vec=list(hcfg.eta), eta_phi_vec=list(hcfg.eta_phi),
            phi_init=list(hc.phi_init), active_species=[i for i in range(5) if (hcfg.active_species_mask >> i) & 1],
        )
        n = hc.n_obs
        cc = make_cond_cfg(
            np.array(hc.data[: n * 5]).reshape(n, 5), list(hc.idx_sparse[:n]), np.array(hc.sigma_obs[: n * 5]).reshape(n, 5),
            cov_rel=hc.cov_rel, active_theta_indices=[j for j in range(NTHETA) if (hc.active_theta_mask >> j) & 1],
            weights=np.array(hc.weights[: n * 5]).reshape(n, 5), use_absolute_volume=bool(hc.use_absolute_volume),
            tsm_variance=bool(hc.tsm_variance),
        )
        out.append((scfg, cc))
    return out


def evaluate_hmx(hcfg, theta, cond_id=None, n_threads=None, details=False):
    """Oracle evaluation of a product-side batch: logL[N] for theta[N,20], cond_id[N] (None = 0)."""
    theta = _f64(theta)
    N = theta.shape[0]
    cid = np.zeros(N, dtype=np.int32) if cond_id is None else np.asarray(cond_id, dtype=np.int32)
    pairs = from_hmx_config(hcfg)
    logL = np.empty(N)
    status = np.zeros(N, dtype=np.uint32)
    its = np.zeros(N, dtype=np.int64)
    tr = np.zeros(N, dtype=np.int64)
    for k, (scfg, cc) in enumerate(pairs):
        sel = np.nonzero(cid == k)[0]
        if sel.size == 0:
            continue
        l, s, i, t = evaluate_batch(scfg, [cc], theta[sel], None, n_threads=n_threads, details=True)
        logL[sel], status[sel], its[sel], tr[sel] = l, s, i, t
    if details:
        return logL, status, its, tr
    return logL
