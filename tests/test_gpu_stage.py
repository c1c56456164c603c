"""Per-stage TMCMC kernels through the C ABI against the oracle / NumPy / the live reference's golden
resampling indices.  Bars: resampling indices bit-exact; delta-beta, weights, evidence, covariance <= 1e-12."""
import numpy as np
import pytest

from helpers import GOLDEN
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng(built):
    from tmcmc202601_b200.tmcmc import GpuStageOps

    return GpuStageOps().engine


def test_resample_bit_exact_with_live_reference_golden(eng):
    k = np.load(GOLDEN / "stage_kat.npz")
    for c in range(int(k["resample_n"])):
        idx = eng.resample(k[f"resample_w{c}"], float(k["resample_u"][c]))
        assert idx.dtype == np.int64 and np.array_equal(idx, k[f"resample_idx{c}"])


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 1000, 4097, 250000])
def test_resample_bit_exact_with_numpy(eng, N):
    rng = np.random.default_rng(N)
    w = rng.random(N) ** 6
    w /= w.sum()
    for u in (0.0, rng.random(), 1.0 - 2.0**-53):
        cs = np.cumsum(w)
        cs[-1] = 1.0
        ref = np.clip(np.searchsorted(cs, (u + np.arange(N)) / N), 0, N - 1)
        got = eng.resample(w, u)
        assert np.array_equal(got, ref)
        assert np.all(np.diff(got) >= 0)  # sortedness: a size-independent property of systematic resampling


def test_resample_degenerate_and_counts(eng):
    w = np.zeros(1000)
    w[123] = 1.0
    assert np.all(eng.resample(w, 0.7) == 123)
    N = 2_000_000  # beyond oracle sizes: offspring counts are floor/ceil of N w_i
    rng = np.random.default_rng(5)
    w = rng.random(N)
    w /= w.sum()
    idx = eng.resample(w, 0.25)
    counts = np.bincount(idx, minlength=N)
    assert counts.sum() == N and np.all(np.abs(counts - N * w) < 1.0 + 1e-6) and np.all(np.diff(idx) >= 0)


@pytest.mark.parametrize("N,scale", [(7, 1.0), (1000, 5.0), (1000, 80.0), (100000, 20.0)])
def test_beta_step_and_weights(eng, N, scale):
    rng = np.random.default_rng(N + int(scale))
    logL = rng.normal(size=N) * scale
    logL[rng.integers(0, N)] = -1e20  # a failed solve in the population
    for beta in (0.0, 0.4, 0.99):
        db, ess = eng.beta_step(logL, beta, 0.5, 0.02, 0.5)
        db_o, ess_o = orc.beta_step(logL, beta, 0.5, 0.02, 0.5)
        assert abs(db - db_o) <= 1e-12 * db_o
        assert (np.isnan(ess) and np.isnan(ess_o)) or abs(ess - ess_o) <= 1e-10 * ess_o
        w, ls = eng.weights(logL, db)
        w_o, ls_o, _ = orc.weights(logL, db)
        assert np.allclose(w, w_o, rtol=1e-12, atol=1e-300) and abs(w.sum() - 1.0) < 1e-12
        assert abs(ls - ls_o) <= 1e-12 * max(1.0, abs(ls_o))


def test_weights_uniform_fallback(eng):
    w, ls = eng.weights(np.full(16, -np.inf), 0.3)
    assert np.allclose(w, 1.0 / 16) and np.isnan(ls)


@pytest.mark.parametrize("N,d", [(10, 1), (1000, 20), (1001, 9), (50000, 20), (300, 32)])
def test_weighted_cov_matches_numpy(eng, N, d):
    rng = np.random.default_rng(N * d)
    theta = rng.normal(size=(N, d)) * (1.0 + rng.random(d)) + rng.normal(size=d)
    w = rng.random(N) ** 3
    w /= w.sum()
    mean, cov = eng.weighted_cov(theta, w)
    ref = np.atleast_2d(np.cov(theta.T, aweights=w))
    assert np.allclose(cov, ref, rtol=1e-11, atol=1e-14 * np.abs(ref).max())
    assert np.array_equal(cov, cov.T) and np.allclose(mean, np.average(theta, axis=0, weights=w), rtol=1e-13)
    # the two shard-local halves (multi-GPU path) reproduce it
    half = N // 2
    mom = eng.weighted_moments(theta[:half], w[:half]) + eng.weighted_moments(theta[half:], w[half:]) if half else eng.weighted_moments(theta, w)
    m = mom[2:] / mom[0]
    sc = eng.weighted_scatter(theta[:half], w[:half], m) + eng.weighted_scatter(theta[half:], w[half:], m) if half else eng.weighted_scatter(theta, w, m)
    assert np.allclose(sc / (mom[0] - mom[1] / mom[0]), ref, rtol=1e-11, atol=1e-14 * np.abs(ref).max())


def test_log_likelihood_sparse(eng):
    from tmcmc202601_b200.evaluator import build_likelihood_weights, log_likelihood_sparse

    rng = np.random.default_rng(2)
    mu, data = rng.random((6, 5)), rng.random((6, 5))
    sig = rng.random((6, 5)) * 1e-3
    for sigma in (0.23, np.array([0.1, 0.2, 0.3, 0.1, 0.05]), rng.random((6, 5)) + 0.01):
        for W in (None, build_likelihood_weights(6, 5)):
            got = log_likelihood_sparse(mu, sig, data, sigma, rho=0.0, weights=W)
            ref = orc.log_likelihood_sparse(mu, sig, data, sigma, W)
            assert abs(got - ref) <= 1e-12 * abs(ref)
    h = {}
    v = log_likelihood_sparse(mu, np.full((6, 5), -1.0), data, 0.0, health=h)  # negative variance -> 1e-20 floor
    assert np.isfinite(v) and h["n_var_total_clipped"] == 30 and h["n_var_raw_negative"] == 30


def test_systematic_resample_api(eng):
    from tmcmc202601_b200.tmcmc import systematic_resample

    w = np.random.default_rng(1).random(500)
    w /= w.sum()
    idx = systematic_resample(w, np.random.default_rng(42))
    u = np.random.default_rng(42).random()
    assert np.array_equal(idx, orc.resample(w, u))


def test_cumsum_equals_numpy_chain_bitwise_on_adversarial_weights(eng):
    """hmx_cumsum == np.cumsum bit for bit.  The kernel handles tiles in parallel only where that provably equals the
    sequentially rounded chain; the cases below hit every fall-back: binade crossings, exact rounding ties, weights
    heavier than the running sum, zeros, subnormals, sorted and wildly uneven weights, ragged last tiles."""
    rng = np.random.default_rng(0)
    cases = {}
    w = rng.random(300001)
    cases["uniform"] = w / w.sum()
    w = np.exp(rng.normal(size=262144) * 6)
    cases["lognormal"] = w / w.sum()
    w = np.sort(np.exp(rng.normal(size=100000) * 10))
    cases["ascending"] = w / w.sum()
    cases["descending"] = cases["ascending"][::-1].copy()
    cases["constant"] = np.full(200000, 1.0 / 200000)
    w = rng.random(100000)
    w[rng.random(100000) < 0.3] = 0.0
    cases["zeros"] = w / w.sum()
    cases["ties"] = np.concatenate([[0.6], (rng.integers(1, 1000, 50000) + 0.5) * 2.0 ** -53, rng.random(1000) * 1e-6])
    cases["subnormal"] = np.concatenate([np.full(5000, 5e-324), np.full(5000, 1e-310), rng.random(50000) * 1e-5])
    cases["jumps"] = np.concatenate([rng.random(30000) * 1e-9, [0.7], rng.random(30000) * 1e-9, [0.2999], rng.random(30000) * 1e-7])
    cases["heavy"] = np.concatenate([[1e-3], np.full(5000, 4e-4), rng.random(20000) * 1e-6])
    cases["tiny"] = np.array([0.25, 0.5, 0.25])
    cases["single"] = np.array([1.0])
    w = rng.random(2_000_000)
    cases["two_million"] = w / w.sum()
    # running sum lands exactly on / next to a power of two at tile boundaries: the multi-CTA path's binade guess (made from
    # a differently associated sum) is wrong there and the tile must fall back to the chain
    cases["boundary_exact"] = np.concatenate([np.full(4096, 0.25 / 4096), np.full(4096, 0.25 / 4096), np.full(8192, 0.5 / 8192)])
    cases["boundary_near"] = np.concatenate([np.full(4096, 0.125 / 4096 * (1 - 2.0 ** -50)), np.full(4096 * 7, 0.875 / (4096 * 7))])
    w = np.exp(rng.normal(size=1_500_000) * 12)
    cases["tempered_heavy_tail"] = w / w.sum()
    for name, w in cases.items():
        w = np.ascontiguousarray(w, dtype=np.float64)
        ref = np.cumsum(w)
        ref[-1] = 1.0
        got = eng.cumsum(w)
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), name
    # and the resampling indices that depend on it
    for name in ("uniform", "lognormal", "ties", "two_million", "tempered_heavy_tail"):
        w = cases[name] / cases[name].sum()
        u = float(rng.random())
        cs = np.cumsum(w)
        cs[-1] = 1.0
        want = np.clip(np.searchsorted(cs, (u + np.arange(len(w))) / len(w)), 0, len(w) - 1)
        assert np.array_equal(eng.resample(w, u), want), name


def test_multi_cta_cumsum_equals_single_cta_kernel(built, monkeypatch):
    """The multi-CTA cumulative sum (tiles in parallel, integer prefix of the tile totals inside a binade) against the
    single-CTA kernel it replaces for large N: identical bits on 20 random weight vectors of ragged length."""
    from tmcmc202601_b200 import _abi
    from tmcmc202601_b200.engine import HamiltonEngine

    cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.2, tsm_variance=False)
    multi = HamiltonEngine(_abi.make_config([cond]), device=0)
    monkeypatch.setenv("HMX_CUMSUM_SINGLE_CTA", "1")
    single = HamiltonEngine(_abi.make_config([cond]), device=0)
    try:
        rng = np.random.default_rng(7)
        for k in range(20):
            n = int(rng.integers(4 * 4096, 700_000))
            w = np.exp(rng.normal(size=n) * rng.uniform(0.1, 15.0))
            w /= w.sum()
            a, b = multi.cumsum(w), single.cumsum(w)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (k, n)
    finally:
        multi.close()
        single.close()
