"""Shared test helpers (tests only)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from oracle import oracle as orc

ROOT = Path(__file__).resolve().parents[1]
GOLDEN = ROOT / "tests" / "golden"


class OracleStageOps:
    """Stage arithmetic through the CPU oracle: lets `pytest -m 'not gpu'` exercise the host control flow of
    tmcmc202601_b200.tmcmc.run_TMCMC.  Same interface as tmcmc202601_b200.tmcmc.GpuStageOps."""

    def beta_step(self, logL, beta, target_ess_ratio, min_delta_beta, max_delta_beta):
        db, ess = orc.beta_step(logL, beta, target_ess_ratio, min_delta_beta, max_delta_beta)
        return db, (None if np.isnan(ess) else ess)

    def weights(self, logL, delta_beta):
        w, log_Sj, flag = orc.weights(logL, delta_beta)
        return w, (None if flag else log_Sj)

    def resample(self, w, u):
        return orc.resample(w, u)

    def weighted_cov(self, theta, w):
        return orc.weighted_cov(theta, w)[1]


def load_emu():
    import __graft_entry__ as g

    return C.CDLL(str(g.build_host_emulation()))


def dptr(a):
    return a.ctypes.data_as(C.c_void_p)


def toy_loglik(theta):
    """The toy Gaussian likelihood tests/golden/make_golden.py ran through the live reference run_TMCMC."""
    mu = np.array([0.3, -0.2, 0.8])
    s = np.array([0.1, 0.25, 0.05])
    return float(-0.5 * np.sum(((theta - mu) / s) ** 2))


class OracleEngine:
    """Stand-in for tmcmc202601_b200.engine.HamiltonEngine backed by the CPU oracle -- CPU tests only
    (injected through LogLikelihoodEvaluator(engine_factory=...)); the product never constructs it."""

    def __init__(self, config, n_threads=None):
        self.config = config
        self.n_threads = n_threads
        self.calls = []

    def loglik_batch(self, theta, cond_id=None, return_status=False, return_obs=False):
        self.calls.append((np.array(theta, copy=True), None if cond_id is None else np.array(cond_id, copy=True)))
        logL, st, its, _ = orc.evaluate_hmx(self.config, theta, cond_id, n_threads=self.n_threads, details=True)
        return (logL, st.astype(np.uint32), its.astype(np.uint32)) if return_status else logL

    def mutate(self, mutation, theta_res, logL_res, return_proposals=False):
        """hmx_mutate with the kernels' own host-compiled arithmetic (tests/csrc/host_emu.cpp) around the oracle likelihood."""
        emu = load_emu()
        theta_res = np.ascontiguousarray(theta_res, dtype=np.float64)
        logL_res = np.ascontiguousarray(logL_res, dtype=np.float64)
        N, d = theta_res.shape
        K = mutation.K
        props = np.empty((N * K, d))
        full = np.empty((N * K, 20))
        cond = np.empty(N * K, dtype=np.int32)
        inside = np.empty(N * K, dtype=np.uint8)
        emu.emu_mutation_propose(C.byref(mutation), dptr(theta_res), C.c_longlong(N), dptr(props), dptr(full), dptr(cond), dptr(inside))
        ll, st, _ = self.loglik_batch(full, cond, return_status=True)
        theta_out, logL_out = np.empty_like(theta_res), np.empty_like(logL_res)
        counts = np.zeros(2, dtype=np.int64)
        emu.emu_mutation_accept(C.byref(mutation), dptr(props), dptr(np.ascontiguousarray(ll)), dptr(inside), C.c_longlong(N),
                                dptr(theta_res), dptr(logL_res), dptr(theta_out), dptr(logL_out), dptr(counts))
        stats = {"n_accepted": int(counts[0]), "n_candidates": int(counts[1]), "n_nonfinite": int(((st & 1) != 0).sum()),
                 "n_maxiter": int(((st & 2) != 0).sum()), "n_singular": int(((st & 4) != 0).sum())}
        return (theta_out, logL_out, stats, props, ll) if return_proposals else (theta_out, logL_out, stats)

    def mutate_groups(self, mutations, theta_list, logL_list):
        return [self.mutate(m, t, l) for m, t, l in zip(mutations, theta_list, logL_list)]

    def weighted_moments(self, theta, w):
        return np.concatenate([[w.sum(), (w * w).sum()], (theta * w[:, None]).sum(axis=0)])

    def weighted_scatter(self, theta, w, mean):
        x = theta - mean[None, :]
        return x.T @ (x * w[:, None])

    def close(self):
        pass
