"""Concurrent TMCMC runs with merged GPU launches -- the B200 answer to the reference's `--batch all`
(MAIN:1100-1267: a ProcessPoolExecutor over the four conditions, each running its chains one after another).

A single 1000-particle chain hands the GPU 1000-5000 evaluations per launch, which occupies ~4 % of a B200
and is bound by the latency of one solve (~90 ms).  Chains and conditions are statistically independent,
so every (condition, chain) pair runs the reference-shaped `run_TMCMC` in its own host thread and a broker
merges the likelihood requests of all threads into ONE hmx_loglik_batch launch per round (conditions are
told apart by cond_id).  Results are bit-identical to running the chains sequentially: chain seeds do not
depend on other chains (TM:2212-2214), likelihood values do not depend on their position in a batch, and
the linearization hand-over between chains (TM:2223-2231) is inert while linearization stays off.
"""

from __future__ import annotations

import threading
import time
from typing import Any, Dict, List, Optional, Sequence

import numpy as np

from . import _abi, tmcmc
from .engine import HamiltonEngine
from .evaluator import LogLikelihoodEvaluator


class MergedLikelihoodBroker:
    """Collects `loglik_batch` requests from client threads; the last client to arrive launches the merged batch."""

    def __init__(self, engine: HamiltonEngine, n_clients: int, engine_lock: Optional[threading.Lock] = None):
        self.engine = engine
        self.engine_lock = engine_lock or threading.Lock()  # the context is not thread-safe: shared with the stage ops
        self.cv = threading.Condition()
        self.active = int(n_clients)
        self.pending: List[dict] = []
        self.launches = 0
        self.evaluations = 0
        self.gpu_seconds = 0.0
        self.error: Optional[BaseException] = None

    def _flush_locked(self):
        reqs, self.pending = self.pending, []
        mut = [r for r in reqs if r["kind"] == "mutate"]
        ev = [r for r in reqs if r["kind"] == "loglik"]
        try:
            if mut:
                self._launch_mutations(mut)
            if ev:
                self._launch_logliks(ev)
        except BaseException as e:  # wake everybody up with the error instead of dead-locking
            self.error = e
            for r in reqs:
                r["done"] = True
            self.cv.notify_all()
            raise
        self.cv.notify_all()

    def _launch_mutations(self, reqs):
        """All chains that reached their mutation step this round: ONE hmx_mutate_groups call (one ODE launch)."""
        t0 = time.perf_counter()
        with self.engine_lock:
            res = self.engine.mutate_groups([r["mutation"] for r in reqs], [r["theta"] for r in reqs], [r["logL"] for r in reqs])
        self.gpu_seconds += time.perf_counter() - t0
        self.launches += 1
        for r, out in zip(reqs, res):
            self.evaluations += r["theta"].shape[0] * r["mutation"].K
            r["result"] = out
            r["done"] = True

    def _launch_logliks(self, reqs):
        theta = np.concatenate([r["theta"] for r in reqs], axis=0)
        cid = np.concatenate([r["cond_id"] for r in reqs])
        t0 = time.perf_counter()
        with self.engine_lock:
            logL, status, iters = self.engine.loglik_batch(theta, cid, return_status=True)
        self.gpu_seconds += time.perf_counter() - t0
        self.launches += 1
        self.evaluations += theta.shape[0]
        off = 0
        for r in reqs:
            n = r["theta"].shape[0]
            r["result"] = (logL[off:off + n].copy(), status[off:off + n].copy(), iters[off:off + n].copy())
            r["done"] = True
            off += n

    def submit(self, theta: np.ndarray, cond_id: np.ndarray):
        return self._submit({"kind": "loglik", "theta": theta, "cond_id": cond_id, "done": False, "result": None})

    def submit_mutation(self, mutation, theta_res: np.ndarray, logL_res: np.ndarray):
        return self._submit({"kind": "mutate", "mutation": mutation, "theta": theta_res, "logL": logL_res, "done": False, "result": None})

    def _submit(self, req):
        with self.cv:
            self.pending.append(req)
            if len(self.pending) >= self.active:
                self._flush_locked()
            else:
                while not req["done"]:
                    self.cv.wait()
            if self.error is not None:
                raise RuntimeError(f"merged likelihood launch failed: {self.error}")
            return req["result"]

    def client_done(self):
        with self.cv:
            self.active -= 1
            if self.pending and len(self.pending) >= self.active:
                self._flush_locked()


class _BrokerEngine:
    """What a LogLikelihoodEvaluator sees instead of its own HamiltonEngine: condition slots (2k, 2k+1) of the shared one."""

    def __init__(self, broker: MergedLikelihoodBroker, slot: int):
        self.broker = broker
        self.slot = int(slot)

    def loglik_batch(self, theta, cond_id=None, return_status=False, return_obs=False):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        local = np.zeros(theta.shape[0], dtype=np.int32) if cond_id is None else np.asarray(cond_id, dtype=np.int32)
        logL, status, iters = self.broker.submit(theta, (2 * self.slot + local).astype(np.int32))
        return (logL, status, iters) if return_status else logL

    def mutate(self, mutation, theta_res, logL_res, return_proposals=False):
        if return_proposals:
            raise NotImplementedError("proposals are not returned through the merged path")
        mutation.cond_default = 2 * self.slot + (mutation.cond_default & 1)
        if mutation.cond_at_lin >= 0:
            mutation.cond_at_lin = 2 * self.slot + (mutation.cond_at_lin & 1)
        return self.broker.submit_mutation(mutation, np.ascontiguousarray(theta_res, dtype=np.float64),
                                           np.ascontiguousarray(logL_res, dtype=np.float64))

    def close(self):
        pass


class _Tally:
    """Launch / evaluation counters of the unmerged mode (what the broker reports in the merged one)."""

    def __init__(self):
        self.lock = threading.Lock()
        self.launches = 0
        self.evaluations = 0
        self.gpu_seconds = 0.0

    def add(self, n, dt):
        with self.lock:
            self.launches += 1
            self.evaluations += int(n)
            self.gpu_seconds += dt

    def client_done(self):
        pass


class _CountingEngine:
    """A chain's own engine in the unmerged mode; forwards everything, counts launches and evaluations."""

    def __init__(self, inner, tally: _Tally):
        self.inner, self._tally = inner, tally

    def __getattr__(self, name):
        return getattr(self.inner, name)

    def loglik_batch(self, theta, *a, **k):
        t0 = time.perf_counter()
        out = self.inner.loglik_batch(theta, *a, **k)
        self._tally.add(np.shape(theta)[0], time.perf_counter() - t0)
        return out

    def mutate(self, mutation, theta_res, *a, **k):
        t0 = time.perf_counter()
        out = self.inner.mutate(mutation, theta_res, *a, **k)
        self._tally.add(np.shape(theta_res)[0] * mutation.K, time.perf_counter() - t0)
        return out


class _LockedStageOps:
    """Stage kernels on one shared context; calls from the chain threads are serialised."""

    def __init__(self, ops, lock):
        self._ops, self._lock = ops, lock

    def _call(self, name, *a):
        with self._lock:
            return getattr(self._ops, name)(*a)

    def beta_step(self, *a):
        return self._call("beta_step", *a)

    def weights(self, *a):
        return self._call("weights", *a)

    def resample(self, *a):
        return self._call("resample", *a)

    def weighted_cov(self, *a):
        return self._call("weighted_cov", *a)


def run_batch_TMCMC(jobs: Sequence[Dict[str, Any]], solver_kwargs: Dict[str, Any], n_particles: int = 1000, n_stages: int = 30,
                    n_chains: int = 2, seed: int = 42, n_mutation_steps: int = 5, min_delta_beta: float = 0.02,
                    max_delta_beta: float = 0.5, device: int = 0, stage_ops=None, engine_factory=None, device_mutation: bool = False,
                    merge_launches: bool = True):
    """Run len(jobs) x n_chains TMCMC chains concurrently on one GPU.

    jobs[k] needs: model_tag, data, idx_sparse, sigma_obs, phi_init, prior_bounds, theta_base, active_indices
    (optional: cov_rel=0.005, weights, use_absolute_volume).  All jobs share `solver_kwargs` (without phi_init).
    Returns one dict per job with the same five entries run_multi_chain_TMCMC returns, plus broker statistics.
    `device_mutation=True`: the chains' mutation steps are merged as well (hmx_mutate_groups: one group per chain).
    `merge_launches=False`: no broker -- every chain owns a context (its own CUDA stream) and launches as soon as it is ready,
    so a chain waiting for the slowest solve of its stage no longer holds the other chains back: the kernels of different
    chains share the GPU.  Same samples (a likelihood does not depend on the launch it travels in); pays off when the
    chains need different numbers of stages or when a merged launch would exceed one resident round of lanes.
    """
    if not 1 <= len(jobs) <= _abi.MAX_COND // 2:
        raise ValueError(f"between 1 and {_abi.MAX_COND // 2} jobs per GPU context")
    kw = {k: v for k, v in solver_kwargs.items() if k != "phi_init"}
    conds = []
    for j in jobs:
        for tsm in (True, False):  # slot 2k: TSM variance on; 2k+1: off (np.allclose shortcut of the evaluator)
            conds.append(_abi.make_cond(j["data"], j["idx_sparse"], j["sigma_obs"], j["phi_init"], cov_rel=j.get("cov_rel", 0.005),
                                        active_theta_indices=[i for i in j["active_indices"] if i < _abi.NTHETA],
                                        weights=j.get("weights"), use_absolute_volume=j.get("use_absolute_volume", False),
                                        tsm_variance=tsm))
    if merge_launches:
        cfg = _abi.make_config(conds, **kw)
        engine = engine_factory(cfg) if engine_factory else HamiltonEngine(cfg, device=device)
        lock = threading.Lock()
        broker = MergedLikelihoodBroker(engine, len(jobs) * n_chains, engine_lock=lock)
        ops = _LockedStageOps(stage_ops if stage_ops is not None else tmcmc.GpuStageOps(engine=engine), lock)
    else:
        engine, broker, ops = None, _Tally(), stage_ops
    results: List[List[Optional[tmcmc.TMCMCResult]]] = [[None] * n_chains for _ in jobs]
    errors: List[BaseException] = []

    def worker(k: int, chain: int):
        j = jobs[k]
        try:
            ev = LogLikelihoodEvaluator(dict(kw, phi_init=j["phi_init"]), [0, 1, 2, 3, 4], list(j["active_indices"]),
                                        np.asarray(j["theta_base"], float), j["data"], j["idx_sparse"], j["sigma_obs"],
                                        cov_rel=j.get("cov_rel", 0.005), theta_linearization=np.asarray(j["theta_base"], float),
                                        weights=j.get("weights"), use_absolute_volume=j.get("use_absolute_volume", False),
                                        engine_factory=(lambda _cfg, k=k: _BrokerEngine(broker, k)) if merge_launches else
                                        ((lambda c: _CountingEngine(engine_factory(c), broker)) if engine_factory else
                                         (lambda c: _CountingEngine(HamiltonEngine(c, device=device), broker))))
            my_ops = ops if ops is not None else tmcmc.GpuStageOps(engine=ev._engine.inner)
            chain_seed = (int(seed or 0) + tmcmc._stable_hash_int(j["model_tag"]) % (2**31) + n_chains * 1000 + chain) % (2**31)
            results[k][chain] = tmcmc.run_TMCMC(
                log_likelihood=ev, prior_bounds=[tuple(b) for b in j["prior_bounds"]], n_particles=n_particles, n_stages=n_stages,
                min_delta_beta=min_delta_beta, max_delta_beta=max_delta_beta, seed=chain_seed,
                model_name=f"{j['model_tag']}_chain{chain + 1}", evaluator=ev, theta_base_full=np.asarray(j["theta_base"], float),
                active_indices=list(j["active_indices"]), n_mutation_steps=n_mutation_steps, stage_ops=my_ops, device_mutation=device_mutation)
            if not merge_launches:
                ev.close()
        except BaseException as e:  # noqa: BLE001 - reported to the caller after all threads stopped
            errors.append(e)
        finally:
            broker.client_done()

    t0 = time.perf_counter()
    threads = [threading.Thread(target=worker, args=(k, c), daemon=True) for k in range(len(jobs)) for c in range(n_chains)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    wall = time.perf_counter() - t0
    if errors:
        raise errors[0]
    out = []
    for k, j in enumerate(jobs):
        res = results[k]
        chains = [r.samples for r in res]
        logL = [r.logL_values for r in res]
        best = max(range(n_chains), key=lambda c: float(np.max(logL[c])))
        out.append({
            "model_tag": j["model_tag"], "chains": chains, "logL": logL, "MAP": chains[best][int(np.argmax(logL[best]))],
            "converged": [r.converged for r in res], "results": res,
        })
    stats = {"wall_s": wall, "merged_launches": broker.launches, "likelihood_evaluations": broker.evaluations,
             "likelihood_wall_s": broker.gpu_seconds}
    if engine is not None and not engine_factory:
        engine.close()
    return out, stats
