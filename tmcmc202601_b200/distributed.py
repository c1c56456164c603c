"""Multi-GPU plumbing: one process per GPU (torchrun), `torch.distributed` over NCCL/NVLink.

The path shards trivially -- evaluations are independent -- so the data path has no
collective inside the ODE kernel.  Per tempering stage the replicas exchange exactly two things:

  * the per-particle log-likelihoods: every rank evaluates a contiguous block of the population / of the
    N*K proposals and one all-gather (8 bytes per evaluation) gives every rank the full vector, after which
    the beta bisection, the weights and the resampling indices are computed redundantly and
    deterministically on every rank (identical inputs -> identical decisions, no further synchronisation);
  * the weighted proposal covariance: shard-local moments (2+d doubles) and centred scatter (d*d doubles)
    are summed with two all-reduces (232 + 400 doubles at d = 20).

All ranks run the same host program with the same seed, so theta never has to be communicated.
With the gloo backend (CPU tests, world_size 2) the same code paths run on host tensors.
"""

from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_bounds(n: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: the first n % G ranks get one extra element."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist

    return dist


def _comm_device(group=None):
    import torch

    dist = _dist()
    backend = dist.get_backend(group)
    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")


def all_gather_blocks(local: np.ndarray, n_total: int, group=None, unit: int = 1) -> np.ndarray:
    """All-gather of contiguous float64 blocks laid out by shard_bounds (blocks differ by at most one row).
    `unit` > 1: rows of `unit` elements (n_total counts elements, the split is by rows)."""
    import torch

    dist = _dist()
    G, r = dist.get_world_size(group), dist.get_rank(group)
    unit = int(unit)
    rows = int(n_total) // unit
    width = (-(-rows // G) if rows else 0) * unit
    dev = _comm_device(group)
    send = torch.zeros(max(width, 1), dtype=torch.float64, device=dev)
    lo, hi = shard_bounds(rows, G, r)
    assert local.shape[0] == (hi - lo) * unit, "local block does not match shard_bounds"
    if hi > lo:
        send[: (hi - lo) * unit] = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64)).to(dev)
    recv = torch.empty(G * max(width, 1), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.cpu().numpy().reshape(G, max(width, 1))
    out = np.empty(rows * unit, dtype=np.float64)
    for k in range(G):
        a, b = shard_bounds(rows, G, k)
        out[a * unit:b * unit] = recv[k, : (b - a) * unit]
    return out


def all_reduce_sum(x: np.ndarray, group=None) -> np.ndarray:
    import torch

    dist = _dist()
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(_comm_device(group))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()


class ShardedEvaluator:
    """Evaluator wrapper: `batch(theta)` evaluates this rank's block and all-gathers the log-likelihoods.

    Everything else (call counters, linearization hooks, `__call__`) is forwarded to the wrapped evaluator,
    so it can be handed to run_TMCMC unchanged.
    """

    def __init__(self, evaluator, group=None):
        self._inner = evaluator
        self._group = group

    def __getattr__(self, name):
        return getattr(self._inner, name)

    def __call__(self, theta_sub):
        return self._inner(theta_sub)

    def batch(self, theta: np.ndarray) -> np.ndarray:
        dist = _dist()
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        n = theta.shape[0]
        G, r = dist.get_world_size(self._group), dist.get_rank(self._group)
        lo, hi = shard_bounds(n, G, r)
        local = self._inner.batch(theta[lo:hi]) if hi > lo else np.empty(0)
        return all_gather_blocks(np.asarray(local, dtype=np.float64), n, self._group)


    def mutate_device(self, theta_res, logL_res, factor, lows, highs, K, beta_next, seed, sequence=0, index_offset=0):
        """Device-side mutation of this rank's block; the Philox counters use the GLOBAL particle index, so the
        gathered population is bitwise what a single GPU would produce."""
        dist = _dist()
        theta_res = np.ascontiguousarray(theta_res, dtype=np.float64)
        logL_res = np.ascontiguousarray(logL_res, dtype=np.float64)
        n, d = theta_res.shape
        G, r = dist.get_world_size(self._group), dist.get_rank(self._group)
        lo, hi = shard_bounds(n, G, r)
        if hi > lo:
            th, ll, st = self._inner.mutate_device(theta_res[lo:hi], logL_res[lo:hi], factor, lows, highs, K, beta_next, seed,
                                                   sequence=sequence, index_offset=index_offset + lo)
        else:
            th, ll, st = np.empty((0, d)), np.empty(0), {}
        keys = ["n_accepted", "n_candidates", "n_nonfinite", "n_maxiter", "n_singular"]
        tot = all_reduce_sum(np.array([float(st.get(k, 0)) for k in keys]), self._group)
        theta = all_gather_blocks(np.ascontiguousarray(th).reshape(-1), n * d, self._group, unit=d).reshape(n, d)
        logL = all_gather_blocks(np.asarray(ll, dtype=np.float64), n, self._group)
        return theta, logL, {k: int(v) for k, v in zip(keys, tot)}


class ShardedStageOps:
    """Stage arithmetic for replicas: bisection / weights / resampling on the full (gathered) vector,
    covariance from shard-local partial sums + two all-reduces."""

    def __init__(self, ops, group=None):
        self._ops = ops
        self._group = group

    def beta_step(self, *a, **k):
        return self._ops.beta_step(*a, **k)

    def weights(self, *a, **k):
        return self._ops.weights(*a, **k)

    def resample(self, *a, **k):
        return self._ops.resample(*a, **k)

    def weighted_cov(self, theta: np.ndarray, w: np.ndarray) -> np.ndarray:
        dist = _dist()
        n, d = theta.shape
        G, r = dist.get_world_size(self._group), dist.get_rank(self._group)
        lo, hi = shard_bounds(n, G, r)
        eng = getattr(self._ops, "engine", self._ops)
        mom = eng.weighted_moments(theta[lo:hi], w[lo:hi]) if hi > lo else np.zeros(2 + d)
        mom = all_reduce_sum(mom, self._group)
        mean = mom[2:] / mom[0]
        sc = eng.weighted_scatter(theta[lo:hi], w[lo:hi], mean) if hi > lo else np.zeros((d, d))
        sc = all_reduce_sum(sc, self._group)
        fact = mom[0] - mom[1] / mom[0]
        return sc * (1.0 / fact)
