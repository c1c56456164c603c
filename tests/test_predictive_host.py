"""Batched trajectory consumers, host logic: generate_dataset against vectors of the LIVE reference
generator (tests/golden/training_dataset_kat.npz, made by tests/golden/make_golden_dataset.py) with an oracle-backed
solver stand-in, and its failure handling against a sequential restatement of the reference loop."""
import numpy as np

from helpers import GOLDEN
from oracle import oracle as orc
from tmcmc202601_b200 import predictive


class OracleSolver:
    """run_deterministic_batch(theta, rows) through the CPU oracle (tests only)."""

    def __init__(self, maxtimestep=60, dt=1e-5, fail=None):
        self.maxtimestep, self.dt = maxtimestep, dt
        self.cfg = orc.make_solver_cfg(phi_init=0.2, dt=dt, maxtimestep=maxtimestep, eps=1e-6, Kp1=1e-4, c_const=100.0, alpha_const=100.0,
                                       K_hill=0.05, n_hill=4.0, max_newton_iter=50)
        self.fail = fail or (lambda theta: False)
        self.batches = []

    def time_array(self):
        return np.concatenate([[0.0], (np.arange(self.maxtimestep) + 1) * self.dt])

    def run_one(self, th):
        _, g, _, _, st = orc.run_deterministic(self.cfg, th)
        if self.fail(th):
            g = g.copy()
            g[5, 2] = np.nan
            st |= 1
        return g, st

    def run_deterministic_batch(self, theta, rows=None, return_status=False):
        self.batches.append(len(theta))
        gs, sts = zip(*[self.run_one(t) for t in theta])
        g = np.stack(gs)
        if rows is not None:
            g = g[:, rows]
        return (g, np.array(sts, dtype=np.uint32)) if return_status else g


def test_generate_dataset_equals_the_live_reference_generator():
    k = np.load(GOLDEN / "training_dataset_kat.npz")
    for name, mf, post in (("uniform", 0.0, None), ("map", 0.4, None), ("mix", 0.3, k["posterior"])):
        th, phi, t, b = predictive.generate_dataset(14, k["bounds_in"], seed=11, maxtimestep=60, dt=1e-5, n_time_out=13, theta_map=k["theta_map"],
                                                    map_frac=mf, map_std_frac=0.1, posterior_samples=post, posterior_frac=0.5,
                                                    solver=OracleSolver())
        assert np.array_equal(th, k[f"{name}_theta"]), name          # same random stream, same category order
        assert np.array_equal(b, k[f"{name}_bounds"]) and b[18, 1] > k["bounds_in"][18, 1]
        assert np.allclose(t, k[f"{name}_t"], rtol=1e-15)
        assert np.allclose(phi, k[f"{name}_phi"], rtol=1e-9, atol=1e-14), name


def _reference_loop(n_samples, bounds, seed, theta_map, map_frac, map_std_frac, post, post_frac, solver, rows):
    """Sequential restatement of generate_training_data.py:204-330 (one solve per draw, failures redrawn)."""
    rng = np.random.default_rng(seed)
    bounds = predictive.expand_bounds_to_include(bounds, theta_map)
    tm = theta_map if map_frac > 0 else None
    n_post = int(n_samples * post_frac) if post is not None else 0
    n_map = int((n_samples - n_post) * map_frac) if tm is not None else 0
    predictive.sample_theta(bounds, rng)
    out_t, out_p, pd, md, pi, failed = [], [], 0, 0, 0, 0
    while len(out_t) < n_samples:
        up = post is not None and pd < n_post
        um = (not up) and tm is not None and md < n_map
        if up:
            th = post[pi % len(post)].copy()
            pi += 1
            th = np.clip(th + rng.normal(0, 0.01, size=20), bounds[:, 0], bounds[:, 1])
        else:
            th = predictive.sample_theta(bounds, rng, theta_map=tm if um else None, map_std_frac=map_std_frac)
        g, st = solver.run_one(th)
        phi = g[:, :5]
        if np.any(~np.isfinite(phi)) or np.any(phi < -0.5) or np.any(phi > 1.5):
            failed += 1
            continue
        out_t.append(th)
        out_p.append(phi[rows])
        pd += up
        md += um
    return np.array(out_t), np.array(out_p), failed


def test_failed_solves_are_redrawn_in_stream_order():
    k = np.load(GOLDEN / "training_dataset_kat.npz")
    bad = lambda th: (int(abs(th[0]) * 1e6) % 5) == 0  # a deterministic ~20 % of the draws "fail"
    rows = np.linspace(0, 60, 13, dtype=int)
    for mf, post in ((0.0, None), (0.3, k["posterior"])):
        s1 = OracleSolver(fail=bad)
        th, phi, t, b = predictive.generate_dataset(20, k["bounds_in"], seed=5, maxtimestep=60, n_time_out=13, theta_map=k["theta_map"], map_frac=mf,
                                                    posterior_samples=post, solver=s1)
        want_t, want_p, n_failed = _reference_loop(20, k["bounds_in"], 5, k["theta_map"], mf, 0.1, post, 0.5, OracleSolver(fail=bad), rows)
        assert n_failed >= 2 and len(s1.batches) == n_failed + 1  # one extra batch per failure, never one solve per draw
        assert np.array_equal(th, want_t) and np.array_equal(phi, want_p)
        assert not any(bad(x) for x in th)


def test_sample_theta_and_bound_expansion():
    b = np.array([[0.0, 2.0]] * 20)
    b[3] = [0.0, 0.0]
    rng = np.random.default_rng(0)
    th = predictive.sample_theta(b, rng)
    assert th[3] == 0.0 and np.all((th >= 0) & (th <= 2))
    tm = np.full(20, 1.0)
    tm[5] = 9.0
    th2 = predictive.sample_theta(b, np.random.default_rng(0), theta_map=tm, map_std_frac=0.1)
    assert th2[5] == 2.0 and abs(th2[0] - 1.0) < 1.0
    e = predictive.expand_bounds_to_include(b, tm)
    assert e[5, 1] == 9.0 + 0.2 * 4.5 and e[3].tolist() == [0.0, 0.0] and np.array_equal(e[0], b[0])
