"""Multi-process plumbing on the CPU: world_size 2, gloo backend, 127.0.0.1 rendezvous.  Covers the N>1 path
of tmcmc202601_b200.distributed (block sharding, log-likelihood all-gather, covariance all-reduce) and a
sharded run_TMCMC whose replicas must stay bit-identical."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    from helpers import OracleEngine, OracleStageOps, toy_loglik
    from tmcmc202601_b200 import distributed as D
    from tmcmc202601_b200 import tmcmc

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. ragged all-gather (7 elements over 2 ranks, then an empty vector)
        for n in (7, 2, 1, 0):
            lo, hi = D.shard_bounds(n, world, rank)
            full = D.all_gather_blocks(np.arange(lo, hi, dtype=np.float64) * 1.5, n)
            assert np.array_equal(full, np.arange(n) * 1.5), (n, full)

        # 2. sharded covariance == np.cov
        rng = np.random.default_rng(0)
        theta = rng.normal(size=(101, 6))
        w = rng.random(101)
        w /= w.sum()

        class Ops(OracleStageOps):
            engine = OracleEngine(None)

        cov = D.ShardedStageOps(Ops()).weighted_cov(theta, w)
        assert np.allclose(cov, np.cov(theta.T, aweights=w), rtol=1e-11, atol=1e-15)

        # 3. sharded evaluator inside run_TMCMC: each rank evaluates half of every batch
        class Ev:
            n_local = 0

            def __call__(self, t):
                return toy_loglik(t)

            def batch(self, th):
                Ev.n_local += len(th)
                return np.array([toy_loglik(t) for t in th])

        ev = D.ShardedEvaluator(Ev())
        res = tmcmc.run_TMCMC(ev, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=120, n_stages=30, seed=99,
                              min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=2, stage_ops=D.ShardedStageOps(Ops()))
        # 4. sharded device-side mutation: every rank mutates its block with GLOBAL Philox counters
        from tmcmc202601_b200 import workloads
        from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

        cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
        obs_rows = np.array([3, 6, 9, 12, 14])

        def factory(cfg):
            for o, v in enumerate(obs_rows):
                cfg.cond[0].idx_sparse[o] = cfg.cond[1].idx_sparse[o] = int(v)
            return OracleEngine(cfg, n_threads=1)

        theta_base = np.array(cd["bounds"]).mean(axis=1)
        inner = LogLikelihoodEvaluator(dict(workloads.PRODUCTION_SOLVER, maxtimestep=15, phi_init=cd["phi_init"]), [0, 1, 2, 3, 4],
                                       list(range(20)), theta_base, cd["data"], obs_rows, cd["sigma_obs"], cov_rel=0.005,
                                       engine_factory=factory)
        b = np.array(cd["bounds"])
        pop = np.random.default_rng(4).uniform(b[:, 0], b[:, 1], size=(21, 20))
        ll0 = D.ShardedEvaluator(inner).batch(pop)
        fac = np.diag(0.02 * (b[:, 1] - b[:, 0]))
        th_m, ll_m, st_m = D.ShardedEvaluator(inner).mutate_device(pop, ll0, fac, b[:, 0], b[:, 1], 3, 0.3, seed=77, sequence=2)
        np.savez(Path(out_dir) / f"rank{rank}.npz", samples=res.samples, logL=res.logL_values, beta=np.array(res.beta_schedule),
                 n_local=Ev.n_local, mut_theta=th_m, mut_logL=ll_m, mut_stats=np.array([st_m["n_accepted"], st_m["n_candidates"]]),
                 pop=pop, ll0=ll0)
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["samples"], r1["samples"]) and np.array_equal(r0["logL"], r1["logL"])  # replicas in lock-step
    assert np.array_equal(r0["beta"], r1["beta"]) and r0["beta"][-1] == 1.0
    assert abs(int(r0["n_local"]) - int(r1["n_local"])) <= len(r0["beta"]) * 2  # the work really was split

    # and the sharded run equals the single-process run with the same seed
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import OracleStageOps, toy_loglik
    from tmcmc202601_b200 import tmcmc

    solo = tmcmc.run_TMCMC(toy_loglik, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=120, n_stages=30, seed=99,
                           min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=2, stage_ops=OracleStageOps())
    assert np.allclose(solo.beta_schedule, r0["beta"], rtol=1e-12)
    assert np.allclose(solo.samples, r0["samples"], rtol=1e-9, atol=1e-12)

    # device-side mutation: both ranks hold the same gathered result, and it is bitwise what ONE process produces
    assert np.array_equal(r0["mut_theta"], r1["mut_theta"]) and np.array_equal(r0["mut_logL"], r1["mut_logL"])
    assert np.array_equal(r0["mut_stats"], r1["mut_stats"]) and r0["mut_stats"][1] == 21 * 3
    from helpers import OracleEngine
    from tmcmc202601_b200 import workloads
    from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    obs_rows = np.array([3, 6, 9, 12, 14])

    def factory(cfg):
        for o, v in enumerate(obs_rows):
            cfg.cond[0].idx_sparse[o] = cfg.cond[1].idx_sparse[o] = int(v)
        return OracleEngine(cfg)

    b = np.array(cd["bounds"])
    one = LogLikelihoodEvaluator(dict(workloads.PRODUCTION_SOLVER, maxtimestep=15, phi_init=cd["phi_init"]), [0, 1, 2, 3, 4],
                                 list(range(20)), b.mean(axis=1), cd["data"], obs_rows, cd["sigma_obs"], cov_rel=0.005,
                                 engine_factory=factory)
    th1, ll1, st1 = one.mutate_device(r0["pop"], r0["ll0"], np.diag(0.02 * (b[:, 1] - b[:, 0])), b[:, 0], b[:, 1], 3, 0.3, seed=77, sequence=2)
    assert np.array_equal(th1, r0["mut_theta"]) and np.array_equal(ll1, r0["mut_logL"])
    assert [st1["n_accepted"], st1["n_candidates"]] == r0["mut_stats"].tolist()
    assert 0 < st1["n_accepted"] <= 63  # (15 time steps: the likelihood is almost flat)


def test_shard_bounds_partition():
    from tmcmc202601_b200.distributed import shard_bounds

    for n in (0, 1, 7, 8, 1000, 1001):
        for G in (1, 2, 3, 8):
            blocks = [shard_bounds(n, G, r) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
