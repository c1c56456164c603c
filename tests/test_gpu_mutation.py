"""hmx_mutate on the B200: the CUDA proposal / accept kernels against their host-compiled twin and the
oracle likelihood, and a TMCMC run with device-side mutation against the reference's stored posterior."""
import ctypes as C
import sys

import numpy as np
import pytest

from helpers import GOLDEN, dptr, load_emu
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, tmcmc, workloads
from tmcmc202601_b200.engine import HamiltonEngine

pytestmark = pytest.mark.gpu


def _setup(N=96, K=3, seed=21, seq=1, off=5):
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    conds = [workloads.cond_struct(cd), workloads.cond_struct(cd, tsm_variance=False)]
    cfg = _abi.make_config(conds, **workloads.PRODUCTION_SOLVER)
    b = np.array(cd["bounds"])
    rng = np.random.default_rng(8)
    theta_res = rng.uniform(b[:, 0], b[:, 1], size=(N, 20))
    a = rng.normal(size=(20, 20)) * 0.02 * (b[:, 1] - b[:, 0]).mean()
    factor = tmcmc._mvn_factor(a @ a.T + 1e-6 * np.eye(20))
    mut = _abi.make_mutation(factor, b[:, 0], b[:, 1], K, 0.6, seed, sequence=seq, index_offset=off, theta_base=b.mean(axis=1),
                             active_map=list(range(20)), theta_lin=b.mean(axis=1), cond_default=0, cond_at_lin=1)
    return cfg, mut, theta_res, b


def test_hmx_mutate_equals_host_twin_and_oracle(built):
    cfg, mut, theta_res, b = _setup()
    N, K = theta_res.shape[0], mut.K
    eng = HamiltonEngine(cfg, 0)
    logL_res = eng.loglik_batch(theta_res, np.zeros(N, dtype=np.int32))
    th, ll, st, props, llp = eng.mutate(mut, theta_res, logL_res, return_proposals=True)
    eng.close()
    # proposals: same Philox counters, same arithmetic (sincospi vs sin/cos differ in the last bits only)
    emu = load_emu()
    p2, full = np.empty((N * K, 20)), np.empty((N * K, 20))
    cond, inside = np.empty(N * K, dtype=np.int32), np.empty(N * K, dtype=np.uint8)
    emu.emu_mutation_propose(C.byref(mut), dptr(theta_res), C.c_longlong(N), dptr(p2), dptr(full), dptr(cond), dptr(inside))
    assert np.allclose(props, p2, rtol=1e-12, atol=1e-13)
    assert np.all((props >= b[:, 0]) & (props <= b[:, 1])) and inside.all()
    # likelihoods of the proposals: the ordinary kernels, 1e-9 against the oracle
    ref = orc.evaluate_hmx(cfg, props, cond)
    assert np.max(np.abs(llp - ref) / np.abs(ref)) < 1e-9
    # accept chain replayed on the host from the GPU's proposals and likelihoods
    out_t, out_l = np.empty_like(theta_res), np.empty_like(logL_res)
    counts = np.zeros(2, dtype=np.int64)
    emu.emu_mutation_accept(C.byref(mut), dptr(props), dptr(llp), dptr(inside), C.c_longlong(N), dptr(theta_res), dptr(logL_res),
                            dptr(out_t), dptr(out_l), dptr(counts))
    assert np.array_equal(th, out_t) and np.array_equal(ll, out_l)
    assert [st["n_accepted"], st["n_candidates"]] == counts.tolist() and st["n_candidates"] == N * K
    assert 0 < st["n_accepted"] < N * K
    # every returned particle is either the resampled state or one of its own proposals, with the matching likelihood
    P = props.reshape(N, K, 20)
    L = llp.reshape(N, K)
    for i in range(N):
        if np.array_equal(th[i], theta_res[i]):
            assert ll[i] == logL_res[i]
        else:
            k = [kk for kk in range(K) if np.array_equal(th[i], P[i, kk])]
            assert k and ll[i] == L[i, k[-1]]


def test_hmx_mutate_shards_reproduce_the_whole(built):
    cfg, mut, theta_res, b = _setup(N=64, off=0)
    eng = HamiltonEngine(cfg, 0)
    logL_res = eng.loglik_batch(theta_res, np.zeros(64, dtype=np.int32))
    whole = eng.mutate(mut, theta_res, logL_res)
    parts = []
    for lo, hi in ((0, 23), (23, 64)):
        mut.index_offset = lo
        parts.append(eng.mutate(mut, theta_res[lo:hi], logL_res[lo:hi]))
    eng.close()
    assert np.array_equal(whole[0], np.concatenate([p[0] for p in parts]))
    assert np.array_equal(whole[1], np.concatenate([p[1] for p in parts]))
    assert whole[2]["n_accepted"] == sum(p[2]["n_accepted"] for p in parts)


def test_hmx_mutate_rejects_bad_arguments(built):
    cfg, mut, theta_res, b = _setup(N=4)
    eng = HamiltonEngine(cfg, 0)
    mut.cond_default = 7
    with pytest.raises(Exception):
        eng.mutate(mut, theta_res, np.zeros(4))
    mut.cond_default = 0
    mut.K = 0
    with pytest.raises(Exception):
        eng.mutate(mut, theta_res, np.zeros(4))
    eng.close()


def test_production_run_with_device_mutation_matches_reference_posterior(built):
    """Same acceptance bar as the host-mutation run (test_gpu_tmcmc): posterior summaries of the stored reference run
    within Monte-Carlo error, now with proposals / accepts drawn on the device."""
    sys.path.insert(0, str(GOLDEN.parents[1] / "scripts"))
    import run_production_tmcmc as rp

    out, samples, ll = rp.run_condition("Dysbiotic_HOBIC_legacy_1k30", 1000, 2, device_mutation=True)
    cmp_ = rp.compare_with_reference(samples, ll)
    print(f"wall {out['wall_s']:.2f} s for {out['likelihood_evaluations']} evaluations; {cmp_}")
    assert all(out["converged"]) and samples.shape == (2000, 20)
    assert cmp_["z_mean_max"] < 4.5
    assert 0.75 < cmp_["std_ratio_min"] and cmp_["std_ratio_max"] < 1.33
    assert cmp_["quantile_shift_over_std_max"] < 0.5
    assert abs(out["logL_max"] - cmp_["logL_max_ref"]) < 1.5 and abs(out["logL_mean"] - cmp_["logL_mean_ref"]) < 0.5


def test_hmx_mutate_and_trajectory_rows_across_chunks(built, monkeypatch):
    """N K proposals above the chunk size go through the ODE kernel in pieces; the result must not depend on it."""
    cfg, mut, theta_res, b = _setup(N=64, off=0)
    eng = HamiltonEngine(cfg, 0)
    logL_res = eng.loglik_batch(theta_res, np.zeros(64, dtype=np.int32))
    whole = eng.mutate(mut, theta_res, logL_res, return_proposals=True)
    rows = np.array([0, 7, 1200, 2500])
    g0, _, _ = eng.trajectory_rows(theta_res, rows)
    eng.close()
    monkeypatch.setenv("HMX_CHUNK_EVALS", "50")
    eng2 = HamiltonEngine(cfg, 0)
    try:
        parts = eng2.mutate(mut, theta_res, logL_res, return_proposals=True)
        g1, _, _ = eng2.trajectory_rows(theta_res, rows)
    finally:
        eng2.close()
    for a, c in zip(whole, parts):
        if isinstance(a, dict):
            assert a == c
        else:
            assert np.array_equal(a, c)
    assert np.array_equal(g0, g1)


def test_mutate_groups_equals_separate_calls(built):
    """hmx_mutate_groups: three populations (different covariance factor, K, beta, seed, condition slot, size) in one
    set of launches == three hmx_mutate calls, bit for bit."""
    cfg, mut, theta_res, b = _setup(N=80, off=0)
    eng = HamiltonEngine(cfg, 0)
    logL_res = eng.loglik_batch(theta_res, np.zeros(80, dtype=np.int32))
    rng = np.random.default_rng(3)
    specs = [(0, 30, 3, 0.6, 21, 0), (30, 31, 5, 0.25, 99, 1), (61, 80, 1, 1.0, 7, 0)]
    muts, ths, lls = [], [], []
    for lo, hi, K, beta, seed, cond in specs:
        a = rng.normal(size=(20, 20)) * 0.03
        m = _abi.make_mutation(tmcmc._mvn_factor(a @ a.T + 1e-6 * np.eye(20)), b[:, 0], b[:, 1], K, beta, seed, sequence=4, index_offset=lo,
                               theta_base=b.mean(axis=1), active_map=list(range(20)), theta_lin=b.mean(axis=1), cond_default=cond, cond_at_lin=-1)
        muts.append(m)
        ths.append(theta_res[lo:hi])
        lls.append(logL_res[lo:hi])
    merged = eng.mutate_groups(muts, ths, lls)
    for m, t, l, got in zip(muts, ths, lls, merged):
        want = eng.mutate(m, t, l)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and got[2] == want[2]
    assert merged[1][0].shape == (1, 20) and merged[2][2]["n_candidates"] == 19
    eng.close()


def test_merged_four_condition_run_with_device_mutation(built):
    """BASELINE.json configs[2] with every stage on the device: four conditions x 2 chains, likelihood AND mutation launches
    merged across the eight chains; each chain equals the same chain run alone with device-side mutation."""
    sys.path.insert(0, str(GOLDEN.parents[1] / "scripts"))
    import run_production_tmcmc as rp

    summary, out = rp.run_merged(workloads.CONDITION_KEYS, 200, 2, device_mutation=True)
    assert all(all(o["converged"]) for o in out) and summary["merged_launches"] < 40
    solo, samples, ll = rp.run_condition("Commensal_HOBIC", 200, 2, device_mutation=True)
    k = workloads.CONDITION_KEYS.index("Commensal_HOBIC")
    assert np.array_equal(np.concatenate(out[k]["chains"], axis=0), samples)
    assert np.array_equal(np.concatenate(out[k]["logL"]), ll)
    print(summary)
