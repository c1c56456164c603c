"""HamiltonEngine -- Python handle on one libhamilton context (one GPU).

Arrays may be NumPy arrays (host pointers: the library stages them and synchronises) or CUDA torch
tensors (device pointers: the call only enqueues work on the current torch stream).  There is no CPU
fallback; construction raises when the library or a B200 is missing.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _abi


class HmxError(RuntimeError):
    pass


def _is_torch(x) -> bool:
    """Device-resident argument: a CUDA torch tensor or a DeviceArray of this module (anything with data_ptr())."""
    return hasattr(x, "data_ptr")


def _addr(x) -> Optional[int]:
    if x is None:
        return None
    if _is_torch(x):
        return x.data_ptr()
    return x.ctypes.data


def _arg(x, dtype=np.float64):
    """NumPy input -> contiguous array of `dtype`; device-resident input passes through."""
    if x is None or _is_torch(x):
        return x
    return np.ascontiguousarray(x, dtype=dtype)


class DeviceArray:
    """Row-major float64 array living on the engine's GPU (hmx_population_create): upload once, pass to any engine
    call in place of a NumPy array.  Row slices (`a[lo:hi]`) are views; the owner frees the memory."""

    def __init__(self, engine: "HamiltonEngine", ptr: int, shape, owner: bool = True):
        self._engine, self._ptr, self.shape, self._owner = engine, int(ptr), tuple(int(v) for v in shape), owner

    def data_ptr(self) -> int:
        return self._ptr

    @property
    def size(self) -> int:
        return int(np.prod(self.shape)) if self.shape else 1

    def __getitem__(self, sl):
        if not isinstance(sl, slice) or sl.step not in (None, 1):
            raise TypeError("DeviceArray supports contiguous row slices only")
        lo, hi, _ = sl.indices(self.shape[0])
        row = int(np.prod(self.shape[1:])) if len(self.shape) > 1 else 1
        return DeviceArray(self._engine, self._ptr + lo * row * 8, (max(hi - lo, 0),) + self.shape[1:], owner=False)

    def upload(self, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        if host.size != self.size:
            raise ValueError(f"upload of {host.size} values into a DeviceArray of {self.size}")
        e = self._engine
        e._check(e._lib.hmx_population_upload(e._ctx, self._ptr, host.ctypes.data, host.size), "hmx_population_upload")
        self._keepalive = host  # the copy is asynchronous: keep the source alive until the next synchronising call
        return self

    def numpy(self) -> np.ndarray:
        out = np.empty(self.shape, dtype=np.float64)
        e = self._engine
        e._check(e._lib.hmx_population_download(e._ctx, out.ctypes.data, self._ptr, out.size), "hmx_population_download")
        return out

    def free(self):
        if self._owner and self._ptr and self._engine._ctx:
            e = self._engine
            e._check(e._lib.hmx_population_destroy(e._ctx, self._ptr), "hmx_population_destroy")
        self._ptr = 0


class HamiltonEngine:
    def __init__(self, config: _abi.HmxConfig, device: int = 0):
        self._lib = _abi.load()
        self._ctx = C.c_void_p()
        self.config = config
        self.device = int(device)
        rc = self._lib.hmx_create(C.byref(self._ctx), C.byref(config), self.device)
        if rc != 0:
            msg = self._lib.hmx_last_error(None)
            self._ctx = C.c_void_p()
            raise HmxError(f"hmx_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n_cond = config.n_cond
        self.n_steps = config.n_steps
        self.n_obs_max = max(config.cond[k].n_obs for k in range(config.n_cond))

    # -- plumbing ---------------------------------------------------------------------------------------
    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.hmx_last_error(self._ctx)
            raise HmxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self._ctx:
            self._lib.hmx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_torch_stream(self):
        """Enqueue on torch's current CUDA stream (needed before passing CUDA tensors)."""
        import torch

        h = int(torch.cuda.current_stream(self.device).cuda_stream)
        # torch's default stream is the legacy default stream, whose handle is 0 == NULL == "use the context's own stream" in
        # hmx_set_stream; cudaStreamLegacy (handle 1) names it explicitly, so the library's work is ordered with torch's
        self._check(self._lib.hmx_set_stream(self._ctx, C.c_void_p(h if h != 0 else 1)), "hmx_set_stream")

    def synchronize(self):
        self._check(self._lib.hmx_synchronize(self._ctx), "hmx_synchronize")

    def device_array(self, shape, host=None) -> DeviceArray:
        """A context-owned device array (hmx_population_create), optionally filled from `host`."""
        shape = tuple(int(v) for v in (shape if isinstance(shape, (tuple, list)) else (shape,)))
        rows = shape[0]
        cols = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        p = C.c_void_p()
        self._check(self._lib.hmx_population_create(self._ctx, rows, max(cols, 1), C.byref(p)), "hmx_population_create")
        a = DeviceArray(self, p.value or 0, shape)
        return a.upload(host) if host is not None else a

    # -- hot path ---------------------------------------------------------------------------------------
    def loglik_batch(self, theta, cond_id=None, return_status=False, return_obs=False):
        """logL[N] for FULL 20-vectors theta[N,20] (hmx_loglik_batch).  NumPy in -> NumPy out."""
        theta = _arg(theta)  # a DeviceArray / CUDA tensor (uploaded once per stage) is used in place
        if len(theta.shape) != 2 or theta.shape[1] != _abi.NTHETA:
            raise ValueError("theta must have shape (N, 20)")
        N = theta.shape[0]
        cid = _arg(cond_id, np.int32)
        if cid is not None and not _is_torch(cid) and (cid.shape != (N,) or (N and (cid.min() < 0 or cid.max() >= self.n_cond))):
            raise ValueError("cond_id must have shape (N,) with values in [0, n_cond)")
        logL = np.empty(N, dtype=np.float64)
        status = np.zeros(N, dtype=np.uint32) if return_status else None
        iters = np.zeros(N, dtype=np.uint32) if return_status else None
        obs = np.zeros((N, self.n_obs_max, _abi.NSTATE)) if return_obs else None
        self._check(
            self._lib.hmx_loglik_batch(self._ctx, _addr(theta), _addr(cid), N, _addr(logL), _addr(obs), _addr(status), _addr(iters)),
            "hmx_loglik_batch",
        )
        out = [logL]
        if return_status:
            out += [status, iters]
        if return_obs:
            out += [obs]
        return out[0] if len(out) == 1 else tuple(out)

    def loglik_batch_device(self, theta, cond_id, logL, status=None, iters=None, obs_state=None):
        """Device-resident variant: all arguments are CUDA torch tensors; only enqueues work."""
        N = theta.shape[0]
        self._check(
            self._lib.hmx_loglik_batch(self._ctx, _addr(theta), _addr(cond_id), N, _addr(logL), _addr(obs_state), _addr(status), _addr(iters)),
            "hmx_loglik_batch",
        )

    def tsm_obs_rows(self, theta, cond_id=None, want_x1=True):
        """BiofilmTSM.solve_tsm at the observation rows (hmx_tsm_obs_rows): returns x0[N,n_obs,12], sigma2[N,n_obs,12],
        x1[N,n_obs,12,20] (or None), logL[N], status[N]."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim != 2 or theta.shape[1] != _abi.NTHETA:
            raise ValueError("theta must have shape (N, 20)")
        N = theta.shape[0]
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        x0 = np.zeros((N, self.n_obs_max, _abi.NSTATE))
        sig2 = np.zeros((N, self.n_obs_max, _abi.NSTATE))
        x1 = np.zeros((N, self.n_obs_max, _abi.NSTATE, _abi.NTHETA)) if want_x1 else None
        logL = np.empty(N)
        status = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_tsm_obs_rows(self._ctx, _addr(theta), _addr(cid), N, _addr(x0), _addr(sig2), _addr(x1), _addr(logL),
                                               _addr(status)), "hmx_tsm_obs_rows")
        return x0, sig2, x1, logL, status

    def solve_tsm_batch(self, theta, cond_id=None, want_x0=True):
        """BiofilmTSM.solve_tsm in full (hmx_solve_tsm_batch): x0[N,T+1,12] (or None), sigma2[N,T+1,12], status[N]."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim != 2 or theta.shape[1] != _abi.NTHETA:
            raise ValueError("theta must have shape (N, 20)")
        N = theta.shape[0]
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        x0 = np.empty((N, self.n_steps + 1, _abi.NSTATE)) if want_x0 else None
        s2 = np.empty((N, self.n_steps + 1, _abi.NSTATE))
        status = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_solve_tsm_batch(self._ctx, _addr(theta), _addr(cid), N, _addr(x0), _addr(s2), _addr(status)), "hmx_solve_tsm_batch")
        return x0, s2, status

    def set_sig2_guard(self, all_rows: bool):
        """EV:657-667 at all T + 1 rows (True: the reference's scan, ~2.5x the cost) or at the observation rows (False, default)."""
        self._check(self._lib.hmx_set_sig2_guard(self._ctx, int(bool(all_rows))), "hmx_set_sig2_guard")

    def loglik_grad_batch(self, theta, cond_id=None, want_dstate=False):
        """(logL[N], grad[N,20][, dstate[N,n_obs,12,20]], status[N]): value and gradient of the log-likelihood
        (hmx_loglik_grad_batch: forward sensitivities with the exact Jacobian, likelihood variance frozen)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        if theta.ndim != 2 or theta.shape[1] != _abi.NTHETA:
            raise ValueError("theta must have shape (N, 20)")
        N = theta.shape[0]
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        logL, grad = np.empty(N), np.empty((N, _abi.NTHETA))
        ds = np.zeros((N, self.n_obs_max, _abi.NSTATE, _abi.NTHETA)) if want_dstate else None
        status = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_loglik_grad_batch(self._ctx, _addr(theta), _addr(cid), N, _addr(logL), _addr(grad), _addr(ds), _addr(status)),
                    "hmx_loglik_grad_batch")
        return (logL, grad, ds, status) if want_dstate else (logL, grad, status)

    def trajectory_batch(self, theta, cond_id=None):
        """g[N, n_steps+1, 12] (hmx_trajectory_batch)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        N = theta.shape[0]
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        g = np.empty((N, self.n_steps + 1, _abi.NSTATE))
        status = np.zeros(N, dtype=np.uint32)
        iters = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_trajectory_batch(self._ctx, _addr(theta), _addr(cid), N, _addr(g), _addr(status), _addr(iters)), "hmx_trajectory_batch")
        return g, status, iters

    def mutate_groups(self, mutations, theta_list, logL_list):
        """Several independent populations in one set of launches (hmx_mutate_groups): mutations[g] applies to
        theta_list[g][N_g, d], logL_list[g][N_g].  Returns a list of (theta_new, logL_new, stats) per group."""
        G = len(mutations)
        if not (1 <= G <= _abi.MUT_MAX_GROUPS and len(theta_list) == G and len(logL_list) == G):
            raise ValueError(f"between 1 and {_abi.MUT_MAX_GROUPS} groups, one (theta, logL) pair each")
        ths = [np.ascontiguousarray(t, dtype=np.float64) for t in theta_list]
        lls = [np.ascontiguousarray(l, dtype=np.float64) for l in logL_list]
        d = mutations[0].d
        for m, t, l in zip(mutations, ths, lls):
            if t.ndim != 2 or t.shape[1] != d or m.d != d or l.shape != (t.shape[0],):
                raise ValueError("every group needs theta (N_g, d) with the common d and logL (N_g,)")
        theta_res = np.concatenate(ths, axis=0)
        logL_res = np.concatenate(lls)
        counts = (C.c_int64 * G)(*[t.shape[0] for t in ths])
        muts = (_abi.HmxMutation * G)(*mutations)
        stats = (_abi.HmxMutationStats * G)()
        theta_out, logL_out = np.empty_like(theta_res), np.empty_like(logL_res)
        self._check(
            self._lib.hmx_mutate_groups(self._ctx, G, muts, counts, _addr(theta_res), _addr(logL_res), _addr(theta_out), _addr(logL_out),
                                        None, None, stats),
            "hmx_mutate_groups",
        )
        out, off = [], 0
        for g, t in enumerate(ths):
            n = t.shape[0]
            st = stats[g]
            out.append((theta_out[off:off + n].copy(), logL_out[off:off + n].copy(),
                        {"n_accepted": st.n_accepted, "n_candidates": st.n_candidates, "n_nonfinite": st.n_nonfinite,
                         "n_maxiter": st.n_maxiter, "n_singular": st.n_singular}))
            off += n
        return out

    def legacy_trajectory_rows(self, legacy_cfg: "_abi.HmxLegacyConfig", theta, rows=None):
        """The legacy 4-species solver (hmx_legacy_trajectory_rows): g[N, len(rows) | T+1, 10], status[N], iters[N]."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        nth = 20 if legacy_cfg.n_species == 5 else 14  # an unsupported n_species is rejected by the library
        if theta.ndim != 2 or theta.shape[1] != nth:
            raise ValueError(f"theta must have shape (N, {nth})")
        N = theta.shape[0]
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
        n_rows = legacy_cfg.n_steps + 1 if r is None else r.size
        g = np.empty((N, n_rows, 2 * max(legacy_cfg.n_species, 4) + 2))
        status = np.zeros(N, dtype=np.uint32)
        iters = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_legacy_trajectory_rows(self._ctx, C.byref(legacy_cfg), _addr(theta), N, _addr(r), 0 if r is None else r.size,
                                                         _addr(g), _addr(status), _addr(iters)), "hmx_legacy_trajectory_rows")
        return g, status, iters

    def trajectory_rows(self, theta, rows, cond_id=None):
        """g[N, len(rows), 12]: the time levels `rows` (strictly ascending) of every trajectory (hmx_trajectory_rows)."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        N = theta.shape[0]
        rows = np.ascontiguousarray(rows, dtype=np.int32).reshape(-1)
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        g = np.empty((N, rows.size, _abi.NSTATE))
        status = np.zeros(N, dtype=np.uint32)
        iters = np.zeros(N, dtype=np.uint32)
        self._check(self._lib.hmx_trajectory_rows(self._ctx, _addr(theta), _addr(cid), N, _addr(rows), rows.size, _addr(g), _addr(status),
                                                  _addr(iters)), "hmx_trajectory_rows")
        return g, status, iters

    # -- device-side mutation -------------------------------------------------------------
    def mutate(self, mutation: "_abi.HmxMutation", theta_res, logL_res, return_proposals: bool = False):
        """K Gaussian MH steps for every resampled particle, entirely on the device (hmx_mutate).

        Returns (theta_new[N,d], logL_new[N], stats dict) and, with `return_proposals`, the N*K proposals and their
        log-likelihoods."""
        theta_res = np.ascontiguousarray(theta_res, dtype=np.float64)
        logL_res = np.ascontiguousarray(logL_res, dtype=np.float64)
        N, d = theta_res.shape
        if d != mutation.d or logL_res.shape != (N,):
            raise ValueError("theta_res must be (N, d) with d == mutation.d and logL_res (N,)")
        theta_out, logL_out = np.empty_like(theta_res), np.empty_like(logL_res)
        props = np.empty((N * mutation.K, d)) if return_proposals else None
        llp = np.empty(N * mutation.K) if return_proposals else None
        st = _abi.HmxMutationStats()
        self._check(
            self._lib.hmx_mutate(self._ctx, C.byref(mutation), _addr(theta_res), _addr(logL_res), N, _addr(theta_out), _addr(logL_out),
                                 _addr(props), _addr(llp), C.byref(st)),
            "hmx_mutate",
        )
        stats = {"n_accepted": st.n_accepted, "n_candidates": st.n_candidates, "n_nonfinite": st.n_nonfinite,
                 "n_maxiter": st.n_maxiter, "n_singular": st.n_singular}
        return (theta_out, logL_out, stats, props, llp) if return_proposals else (theta_out, logL_out, stats)

    # -- stage arithmetic -------------------------------------------------------------------------------
    def beta_step(self, logL, beta, target_ess_ratio=0.5, min_delta_beta=0.02, max_delta_beta=0.5):
        logL = _arg(logL)
        db, ess = C.c_double(0.0), C.c_double(0.0)
        self._check(
            self._lib.hmx_beta_step(self._ctx, _addr(logL), logL.shape[0], beta, target_ess_ratio, min_delta_beta, max_delta_beta,
                                    C.addressof(db), C.addressof(ess)),
            "hmx_beta_step",
        )
        return db.value, ess.value

    def weights(self, logL, delta_beta):
        logL = _arg(logL)
        w = np.empty(logL.shape[0], dtype=np.float64)
        ls = C.c_double(0.0)
        self._check(self._lib.hmx_weights(self._ctx, _addr(logL), logL.shape[0], delta_beta, _addr(w), C.addressof(ls)), "hmx_weights")
        return w, ls.value

    def cumsum(self, w):
        """np.cumsum(w) bit for bit with the last entry forced to 1.0 (hmx_cumsum; the first half of systematic_resample)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        cs = np.empty_like(w)
        self._check(self._lib.hmx_cumsum(self._ctx, _addr(w), w.shape[0], _addr(cs)), "hmx_cumsum")
        return cs

    def resample(self, w, u):
        w = _arg(w)
        idx = np.empty(w.shape[0], dtype=np.int64)
        self._check(self._lib.hmx_resample(self._ctx, _addr(w), w.shape[0], float(u), _addr(idx)), "hmx_resample")
        return idx

    def weighted_cov(self, theta, w):
        theta, w = _arg(theta), _arg(w)
        N, d = theta.shape
        mean, cov = np.empty(d), np.empty((d, d))
        self._check(self._lib.hmx_weighted_cov(self._ctx, _addr(theta), _addr(w), N, d, _addr(mean), _addr(cov)), "hmx_weighted_cov")
        return mean, cov

    def weighted_moments(self, theta, w):
        theta, w = _arg(theta), _arg(w)
        N, d = theta.shape
        m = np.empty(2 + d)
        self._check(self._lib.hmx_weighted_moments(self._ctx, _addr(theta), _addr(w), N, d, _addr(m)), "hmx_weighted_moments")
        return m

    def weighted_scatter(self, theta, w, mean):
        theta, w, mean = _arg(theta), _arg(w), _arg(mean)
        N, d = theta.shape
        s = np.empty((d, d))
        self._check(self._lib.hmx_weighted_scatter(self._ctx, _addr(theta), _addr(w), N, d, _addr(mean), _addr(s)), "hmx_weighted_scatter")
        return s

    # -- device-resident variants (CUDA torch tensors in, CUDA torch tensors out; scalars come back to the host) --
    def weights_device(self, logL, delta_beta, out):
        ls = C.c_double(0.0)
        self._check(self._lib.hmx_weights(self._ctx, _addr(logL), logL.shape[0], delta_beta, _addr(out), C.addressof(ls)), "hmx_weights")
        return ls.value

    def resample_device(self, w, u, out_idx):
        self._check(self._lib.hmx_resample(self._ctx, _addr(w), w.shape[0], float(u), _addr(out_idx)), "hmx_resample")

    def weighted_moments_device(self, theta, w, out):
        N, d = theta.shape
        self._check(self._lib.hmx_weighted_moments(self._ctx, _addr(theta), _addr(w), N, d, _addr(out)), "hmx_weighted_moments")

    def weighted_scatter_device(self, theta, w, mean, out):
        N, d = theta.shape
        self._check(self._lib.hmx_weighted_scatter(self._ctx, _addr(theta), _addr(w), N, d, _addr(mean), _addr(out)), "hmx_weighted_scatter")

    def weighted_cov_device(self, theta, w, out_mean, out_cov):
        N, d = theta.shape
        self._check(self._lib.hmx_weighted_cov(self._ctx, _addr(theta), _addr(w), N, d, _addr(out_mean), _addr(out_cov)), "hmx_weighted_cov")

    def log_likelihood_sparse(self, mu, sig, data, sigma_obs, weights=None) -> float:
        data = np.ascontiguousarray(data, dtype=np.float64)
        n_obs = data.shape[0]
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        sig = np.ascontiguousarray(np.broadcast_to(np.asarray(sig, dtype=np.float64), data.shape))
        so = np.ascontiguousarray(_abi.broadcast_sigma(sigma_obs, n_obs))
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        out = C.c_double(0.0)
        self._check(
            self._lib.hmx_log_likelihood_sparse(self._ctx, _addr(mu), _addr(sig), _addr(data), _addr(so), _addr(w), n_obs, C.addressof(out)),
            "hmx_log_likelihood_sparse",
        )
        return out.value

    # -- measurement ------------------------------------------------------------------------------------
    def set_latency_kernel(self, max_evals: int = -1):
        """Run batches of at most `max_evals` evaluations with K1L, eight lanes per solve (-1: the measured optimum, 0: off).
        Lower launch latency for production-sized stages; results agree with the default kernel to ~1e-12, not bitwise."""
        self._check(self._lib.hmx_set_latency_kernel(self._ctx, int(max_evals)), "hmx_set_latency_kernel")

    def set_timing(self, enabled: bool = True):
        self._check(self._lib.hmx_set_timing(self._ctx, int(enabled)), "hmx_set_timing")

    def counters(self) -> dict:
        c = _abi.HmxCounters()
        self._check(self._lib.hmx_get_counters(self._ctx, C.byref(c)), "hmx_get_counters")
        return {f: getattr(c, f) for f, _ in _abi.HmxCounters._fields_}

    def measure_fp64_peak(self) -> float:
        v = C.c_double(0.0)
        self._check(self._lib.hmx_measure_fp64_peak(self._ctx, C.byref(v)), "hmx_measure_fp64_peak")
        return v.value


class MultiEngine:
    """Several GPUs of one box behind one handle (hmx_multi_*): NumPy arrays in, NumPy arrays out; the flat evaluation
    list is split into contiguous blocks, one per device, solved concurrently.  `stage_engine` is the context of the first
    device as a HamiltonEngine view for the replicated stage arithmetic."""

    def __init__(self, config: _abi.HmxConfig, devices):
        self._lib = _abi.load()
        self._m = C.c_void_p()
        self.config = config
        self.devices = [int(d) for d in devices]
        ids = (C.c_int32 * len(self.devices))(*self.devices)
        rc = self._lib.hmx_multi_create(C.byref(self._m), C.byref(config), ids, len(self.devices))
        if rc != 0:
            msg = self._lib.hmx_multi_last_error(None)
            self._m = C.c_void_p()
            raise HmxError(f"hmx_multi_create failed ({rc}): {msg.decode() if msg else ''}")
        self.n_cond = config.n_cond

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.hmx_multi_last_error(self._m)
            raise HmxError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def stage_engine(self) -> HamiltonEngine:
        """Borrowed view on the first device's context (do not close it)."""
        e = HamiltonEngine.__new__(HamiltonEngine)
        e._lib, e.config, e.device = self._lib, self.config, self.devices[0]
        e._ctx = C.c_void_p(self._lib.hmx_multi_context(self._m, 0))
        e.n_cond, e.n_steps = self.config.n_cond, self.config.n_steps
        e.n_obs_max = max(self.config.cond[k].n_obs for k in range(self.config.n_cond))
        e.close = lambda: None
        return e

    def loglik_batch(self, theta, cond_id=None, return_status=False):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        N = theta.shape[0]
        cid = None if cond_id is None else np.ascontiguousarray(cond_id, dtype=np.int32)
        logL = np.empty(N)
        status = np.zeros(N, dtype=np.uint32) if return_status else None
        iters = np.zeros(N, dtype=np.uint32) if return_status else None
        self._check(self._lib.hmx_multi_loglik_batch(self._m, _addr(theta), _addr(cid), N, _addr(logL), _addr(status), _addr(iters)),
                    "hmx_multi_loglik_batch")
        return (logL, status, iters) if return_status else logL

    def weighted_cov(self, theta, w):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        N, d = theta.shape
        mean, cov = np.empty(d), np.empty((d, d))
        self._check(self._lib.hmx_multi_weighted_cov(self._m, _addr(theta), _addr(w), N, d, _addr(mean), _addr(cov)), "hmx_multi_weighted_cov")
        return mean, cov

    def close(self):
        if self._m:
            self._lib.hmx_multi_destroy(self._m)
            self._m = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
