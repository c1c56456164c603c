"""Host control flow of tmcmc202601_b200.tmcmc on the CPU: stage arithmetic injected through the oracle
(OracleStageOps), likelihoods are toy callables.  Mirrors the behaviours the reference pins in
tmcmc/program2602/test_tmcmc.py:33-620 and compares a full run with the LIVE reference's trace."""
import numpy as np
import pytest

from helpers import GOLDEN, OracleStageOps, toy_loglik
from tmcmc202601_b200 import tmcmc

OPS = OracleStageOps()


def _run(ll, bounds, **kw):
    kw.setdefault("stage_ops", OPS)
    kw.setdefault("n_jobs", 1)
    return tmcmc.run_TMCMC(ll, bounds, **kw)


def test_run_tmcmc_reproduces_live_reference_trace():
    """Same seed, same toy likelihood -> same beta schedule, acceptance rates, evidence and final particles
    as the reference's run_TMCMC (tests/golden/stage_kat.npz): the RNG stream is consumed identically."""
    k = np.load(GOLDEN / "stage_kat.npz")
    bounds = [tuple(map(float, b)) for b in k["toy_bounds"]]
    res = _run(toy_loglik, bounds, n_particles=int(k["toy_n_particles"]), n_stages=int(k["toy_n_stages"]),
               seed=int(k["toy_seed"]), min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=int(k["toy_n_mutation_steps"]))
    assert np.allclose(res.beta_schedule, k["toy_beta_schedule"], rtol=1e-12, atol=0)
    assert np.allclose(res.acc_rate_history, k["toy_acc_rate"], rtol=1e-12)
    assert abs(res.log_evidence - float(k["toy_log_evidence"])) < 1e-10
    assert np.allclose(res.samples, k["toy_samples"], rtol=1e-10, atol=1e-12)
    assert np.allclose(res.logL_values, k["toy_logL"], rtol=1e-9, atol=1e-10)
    assert np.allclose(res.theta_MAP, k["toy_theta_MAP"], rtol=1e-10)
    assert np.allclose([r["delta_beta"] for r in res.stage_summary], k["toy_delta_beta"], rtol=1e-12)
    assert res.converged and res.beta_schedule[-1] == 1.0


def test_input_validation_matches_reference_exceptions():
    b = [(-1.0, 1.0)]
    with pytest.raises(TypeError):
        _run("nope", b)
    with pytest.raises(ValueError):
        _run(toy_loglik, [])
    with pytest.raises(ValueError):
        _run(toy_loglik, [(1.0, -1.0)])
    with pytest.raises(TypeError):
        _run(toy_loglik, [("a", 1.0)])
    with pytest.raises(ValueError):
        _run(toy_loglik, b, n_particles=0)
    with pytest.raises(ValueError):
        _run(toy_loglik, b, n_stages=0)
    with pytest.raises(ValueError):
        _run(toy_loglik, b, target_ess_ratio=0.0)
    with pytest.raises(ValueError):
        _run(toy_loglik, b, evaluator=object())
    with pytest.raises(NotImplementedError):
        _run(toy_loglik, b, proposal_type="dreamzs")


def test_gaussian_posterior_recovered_and_seed_reproducible():
    mu, s = np.array([0.5, -0.3]), np.array([0.1, 0.2])
    ll = lambda t: float(-0.5 * np.sum(((t - mu) / s) ** 2))  # noqa: E731
    b = [(-2.0, 2.0), (-2.0, 2.0)]
    r1 = _run(ll, b, n_particles=400, n_stages=40, seed=7, max_delta_beta=0.5, n_mutation_steps=3)
    r2 = _run(ll, b, n_particles=400, n_stages=40, seed=7, max_delta_beta=0.5, n_mutation_steps=3)
    r3 = _run(ll, b, n_particles=400, n_stages=40, seed=8, max_delta_beta=0.5, n_mutation_steps=3)
    assert np.array_equal(r1.samples, r2.samples) and not np.array_equal(r1.samples, r3.samples)
    assert r1.converged and np.all(np.diff(r1.beta_schedule) > 0)
    assert np.all(np.abs(r1.samples.mean(axis=0) - mu) < 0.05)
    assert np.all(np.abs(r1.samples.std(axis=0) - s) < 0.05)
    assert np.all(np.abs(r1.theta_MAP - mu) < 0.5)
    # evidence of a Gaussian likelihood under a uniform prior on [-2,2]^2: log(2 pi s1 s2 / 16)
    assert abs(r1.log_evidence - np.log(2 * np.pi * s[0] * s[1] / 16.0)) < 0.25
    summ = tmcmc.compute_MAP_with_uncertainty(r1.samples, r1.logL_values)
    assert set(summ) == {"mean", "std", "MAP", "ci_lower", "ci_upper"} and np.all(summ["ci_lower"] < summ["ci_upper"])


def test_one_dimensional_and_flat_likelihood_edge_cases():
    r = _run(lambda t: float(-0.5 * ((t[0] - 0.2) / 0.05) ** 2), [(-1.0, 1.0)], n_particles=200, n_stages=30, seed=1, max_delta_beta=0.5)
    assert r.samples.shape == (200, 1) and abs(r.samples.mean() - 0.2) < 0.05
    flat = _run(lambda t: 0.0, [(-1.0, 1.0), (0.0, 2.0)], n_particles=100, n_stages=5, seed=2, max_delta_beta=0.5)
    # flat likelihood: only the jump cap limits the first stage; the 25-step bisection then stops 2^-25 short of
    # the bracket end, exactly like the reference loop (TM:617-650), so one more tiny stage follows
    assert flat.converged and flat.beta_schedule[1] == 0.5 and flat.beta_schedule[-1] == 1.0 and len(flat.beta_schedule) == 4
    assert flat.beta_schedule[2] == pytest.approx(1.0 - 0.5 * 2.0**-25, abs=1e-12)
    assert np.all((flat.samples >= [-1, 0]) & (flat.samples <= [1, 2]))


def test_force_beta_one_and_stage_cap():
    r = _run(toy_loglik, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=100, n_stages=2, seed=3, max_delta_beta=0.05)
    assert not r.converged and len(r.beta_schedule) == 3
    r = _run(toy_loglik, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=100, n_stages=2, seed=3, max_delta_beta=0.05, force_beta_one=True)
    assert r.converged and r.beta_schedule[-1] == 1.0


def test_stuck_mutation_raises_like_reference():
    calls = {"n": 0}

    def ll(t):
        calls["n"] += 1
        return 0.0 if calls["n"] <= 50 else -1e6 * calls["n"]  # every later evaluation is hopelessly worse

    with pytest.raises(RuntimeError, match="mutation stuck"):
        _run(ll, [(-1.0, 1.0)], n_particles=50, n_stages=3, seed=4)


def test_cov_schedule_is_ess_schedule_with_transformed_ratio():
    a = _run(toy_loglik, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=200, n_stages=30, seed=5, max_delta_beta=0.5,
             beta_schedule="cov", target_cov=1.0)
    b = _run(toy_loglik, [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)], n_particles=200, n_stages=30, seed=5, max_delta_beta=0.5,
             target_ess_ratio=0.5)
    assert np.allclose(a.beta_schedule, b.beta_schedule)  # CoV <= 1  <=>  ESS >= N/2


def test_reflect_into_bounds():
    assert tmcmc.reflect_into_bounds(1.2, 0.0, 1.0) == pytest.approx(0.8)
    assert tmcmc.reflect_into_bounds(-0.3, 0.0, 1.0) == pytest.approx(0.3)
    assert tmcmc.reflect_into_bounds(7.3, 0.0, 1.0) == pytest.approx(0.7)
    assert tmcmc.reflect_into_bounds(5.0, 2.0, 2.0) == 2.0  # zero-width (locked) bound clips
    x = np.random.default_rng(0).normal(size=(100, 3)) * 10
    y = tmcmc.reflect_into_bounds(x, np.array([-1.0, 0.0, 2.0])[None, :], np.array([1.0, 3.0, 2.0])[None, :])
    assert np.all(y[:, 0] >= -1) and np.all(y[:, 0] <= 1) and np.all(y[:, 2] == 2.0)
    inside = np.abs(x[:, 0]) <= 1
    assert np.allclose(y[inside, 0], x[inside, 0])


def test_evaluate_particles_parallel_dispatch():
    class Ev:
        def __init__(self):
            self.batches = 0

        def __call__(self, t):
            raise AssertionError("batch() must be used")

        def batch(self, th):
            self.batches += 1
            return th.sum(axis=1)

    ev = Ev()
    th = np.arange(12.0).reshape(4, 3)
    assert np.array_equal(tmcmc.evaluate_particles_parallel(ev, th, n_jobs=12), th.sum(axis=1)) and ev.batches == 1
    out = tmcmc.evaluate_particles_parallel(lambda t: 1 / 0, th)
    assert np.all(out == -1e20)  # worker failure -> sentinel (TM:304-306)
    assert tmcmc.evaluate_particles_parallel(ev, np.zeros((0, 3))).size == 0


def test_multi_chain_seeds_and_outputs():
    seen = []

    class Ev:
        call_count = fom_call_count = 0
        _linearization_enabled = False

        def __init__(self, lin):
            self.lin = lin.copy()
            from tmcmc202601_b200.utils import TimingStats

            self.timing = TimingStats()

        def __call__(self, t):
            return toy_loglik(t)

        def batch(self, th):
            return np.array([toy_loglik(t) for t in th])

        def get_linearization_point(self):
            return self.lin.copy()

        def update_linearization_point(self, t):
            self.lin = t.copy()

        def enable_linearization(self, e):
            self._linearization_enabled = e

        def compute_ROM_error(self, t):
            return float("inf") if self._linearization_enabled else 0.0

        def get_health(self):
            return {"n_calls": 1}

    def make(theta_linearization=None):
        seen.append(theta_linearization.copy())
        return Ev(theta_linearization)

    b = [(-1.0, 1.0), (-1.0, 1.0), (0.0, 2.0)]
    chains, logL, MAP, conv, diag = tmcmc.run_multi_chain_TMCMC(
        "Dysbiotic_HOBIC", make, b, np.zeros(3), [0, 1, 2], n_particles=100, n_stages=30, n_chains=2, seed=42,
        min_delta_beta=0.02, max_delta_beta=0.5, n_mutation_steps=2, stage_ops=OPS)
    assert len(chains) == 2 and chains[0].shape == (100, 3) and all(conv)
    assert not np.array_equal(chains[0], chains[1])
    import zlib

    s0 = (42 + (zlib.crc32(b"Dysbiotic_HOBIC") & 0x7FFFFFFF) % 2**31 + 2000 + 0) % 2**31  # TM:2212-2214
    solo = tmcmc.run_TMCMC(Ev(np.zeros(3)), b, n_particles=100, n_stages=30, seed=s0, min_delta_beta=0.02, max_delta_beta=0.5,
                           n_mutation_steps=2, evaluator=Ev(np.zeros(3)), theta_base_full=np.zeros(3), active_indices=[0, 1, 2], stage_ops=OPS)
    assert np.array_equal(solo.samples, chains[0])
    best = max(range(2), key=lambda c: logL[c].max())
    assert np.array_equal(MAP, chains[best][np.argmax(logL[best])])
    assert diag["total_chains"] == 2 and len(diag["stage_summaries"]) == 2 and diag["Rhat_reference"].shape == (3,)
    assert all(not e for e in [False])  # linearization stays off, like the production runs
    assert diag["stage_summaries"][0][-1]["linearization_enabled"] == 0
