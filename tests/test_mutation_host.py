"""Device-side mutation, host-checkable parts: the kernels' arithmetic (tmcmc202601_b200/csrc/
hmx_mutation.cuh) compiled for the CPU -- Philox known answers, normal / uniform moments, reflection against the
Python mirror of TM:201-226, proposal covariance, theta_sub -> theta_full mapping, the accept chain against a NumPy
restatement of TM:908-962 -- and run_TMCMC(device_mutation=True) end to end on an oracle-backed engine."""
import ctypes as C

import numpy as np
import pytest

from helpers import OracleEngine, OracleStageOps, dptr, load_emu
from tmcmc202601_b200 import _abi, tmcmc, workloads
from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

emu = None


@pytest.fixture(scope="module", autouse=True)
def _emu(built):
    global emu
    emu = load_emu()
    emu.emu_reflect.restype = C.c_double
    emu.emu_reflect.argtypes = [C.c_double] * 3
    emu.emu_accept_uniform.restype = C.c_double
    emu.emu_accept_uniform.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    emu.emu_normals.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
    return emu


def _philox(c, k):
    out = (C.c_uint32 * 4)()
    emu.emu_philox((C.c_uint32 * 4)(*c), (C.c_uint32 * 2)(*k), out)
    return list(out)


def test_struct_layout():
    assert emu.emu_sizeof_mutation() == C.sizeof(_abi.HmxMutation)
    assert emu.emu_sizeof_mutation_stats() == C.sizeof(_abi.HmxMutationStats)


def test_philox4x32_10_known_answers():
    """Random123 kat_vectors for philox4x32-10 (Salmon, Moraes, Dror, Shaw: SC'11)."""
    assert _philox([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert _philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert _philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normal_and_uniform_streams():
    n = 40000
    z = np.empty((n, 6))
    for g in range(n):
        emu.emu_normals(12345, 3, g, 3, dptr(z[g]))
    assert abs(z.mean()) < 4 / np.sqrt(z.size)
    assert abs(z.var() - 1.0) < 0.02
    assert abs((z ** 3).mean()) < 0.03
    assert abs((z ** 4).mean() - 3.0) < 0.1
    c = np.corrcoef(z.T)
    assert np.abs(c - np.eye(6)).max() < 0.03  # components (pairs and blocks) are uncorrelated
    # a different sequence tag / seed gives a different, equally distributed stream
    z2 = np.empty(6)
    emu.emu_normals(12345, 4, 0, 3, dptr(z2))
    assert not np.allclose(z2, z[0])
    u = np.array([emu.emu_accept_uniform(99, 0, g) for g in range(n)])
    assert 0.0 < u.min() and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.003


def test_reflection_equals_python_mirror():
    rng = np.random.default_rng(0)
    lo = rng.uniform(-3, 3, 4000)
    w = rng.uniform(0.01, 5, 4000)
    w[::17] = 0.0       # degenerate width: clip
    hi = lo + w
    x = lo + rng.normal(size=4000) * 4 * np.maximum(w, 1.0)
    x[::13] = lo[::13] + 50.3 * w[::13]   # many widths away
    x[::11] = lo[::11]                    # exactly on the boundary
    got = np.array([emu.emu_reflect(a, b, c) for a, b, c in zip(x, lo, hi)])
    want = np.array([tmcmc.reflect_into_bounds(a, b, c) for a, b, c in zip(x, lo, hi)])
    assert np.array_equal(got, want)
    assert np.all((got >= lo) & (got <= hi))


def _mutation(d=6, K=4, seed=7, seq=0, off=0, lows=None, highs=None, factor=None, beta=0.4, **kw):
    rng = np.random.default_rng(5)
    if factor is None:
        a = rng.normal(size=(d, d)) * 0.05
        factor = tmcmc._mvn_factor(a @ a.T + 1e-4 * np.eye(d))
    lows = np.full(d, -10.0) if lows is None else lows
    highs = np.full(d, 10.0) if highs is None else highs
    return _abi.make_mutation(factor, lows, highs, K, beta, seed, sequence=seq, index_offset=off, **kw), factor


def _propose(mut, theta_res):
    N, d = theta_res.shape
    K = mut.K
    props, full = np.empty((N * K, d)), np.empty((N * K, 20))
    cond, inside = np.empty(N * K, dtype=np.int32), np.empty(N * K, dtype=np.uint8)
    emu.emu_mutation_propose(C.byref(mut), dptr(np.ascontiguousarray(theta_res)), C.c_longlong(N), dptr(props), dptr(full), dptr(cond), dptr(inside))
    return props, full, cond, inside


def test_proposals_have_the_requested_covariance_and_mapping():
    d, K, N = 6, 5, 6000
    base = np.arange(20) * 0.01
    amap = [3, 0, 7, 19, 21, 12]  # 21 is beyond the ODE's 20-vector (viscoelastic slot): dropped
    mut, factor = _mutation(d=d, K=K, theta_base=base, active_map=amap)
    theta_res = np.random.default_rng(1).uniform(-1, 1, (N, d))
    props, full, cond, inside = _propose(mut, theta_res)
    dev = props - np.repeat(theta_res, K, axis=0)
    cov = factor @ factor.T
    emp = dev.T @ dev / dev.shape[0]
    assert np.abs(emp - cov).max() < 0.05 * np.abs(cov).max()
    assert abs(dev.mean()) < 5 * np.sqrt(np.diag(cov).max() / dev.shape[0])
    assert inside.all() and (cond == 0).all()
    want = np.tile(base, (N * K, 1))
    for j, a in enumerate(amap):
        if a < 20:
            want[:, a] = props[:, j]
    assert np.array_equal(full, want)


def test_proposals_are_reflected_and_flag_the_allclose_shortcut():
    d, K = 3, 2
    lows, highs = np.array([0.0, 0.0, 1.0]), np.array([0.1, 0.0, 3.0])  # dim 1 is locked (width 0)
    base = np.zeros(20)
    mut, _ = _mutation(d=d, K=K, lows=lows, highs=highs, factor=np.eye(d) * 0.5, theta_base=base, active_map=[0, 1, 2],
                       theta_lin=np.array([0.05, 0.0, 2.0] + [0.0] * 17), cond_default=0, cond_at_lin=1)
    theta_res = np.tile([0.05, 0.0, 2.0], (500, 1))
    props, full, cond, inside = _propose(mut, theta_res)
    assert inside.all()
    assert np.all((props >= lows) & (props <= highs))
    assert np.all(props[:, 1] == 0.0)
    assert (cond == 0).all()  # sd 0.5 never lands within 1e-5 of the centre
    mut2, _ = _mutation(d=d, K=K, lows=lows, highs=highs, factor=np.zeros((d, d)), theta_base=base, active_map=[0, 1, 2],
                        theta_lin=np.array([0.05, 0.0, 2.0] + [0.0] * 17), cond_default=0, cond_at_lin=1)
    _, _, cond2, _ = _propose(mut2, theta_res)
    assert (cond2 == 1).all()  # zero-width proposal == linearization point -> condition without TSM variance (TSMA:366-401)


def test_random_numbers_depend_on_the_global_index_only():
    d, K = 5, 3
    theta_res = np.random.default_rng(2).uniform(-1, 1, (10, d))
    whole = _propose(_mutation(d=d, K=K)[0], theta_res)[0]
    a = _propose(_mutation(d=d, K=K, off=0)[0], theta_res[:4])[0]
    b = _propose(_mutation(d=d, K=K, off=4)[0], theta_res[4:])[0]
    assert np.array_equal(whole, np.concatenate([a, b]))
    other = _propose(_mutation(d=d, K=K, seq=1)[0], theta_res)[0]
    assert not np.allclose(other, whole)


def test_accept_chain_equals_numpy_restatement():
    """TM:908-962: sequential accepts against the evolving state; proposals outside the prior are no candidates,
    NaN likelihoods draw nothing, the -1e20 sentinel always loses."""
    d, K, N, beta, seed, seq, off = 4, 5, 300, 0.37, 11, 2, 1000
    mut, _ = _mutation(d=d, K=K, seed=seed, seq=seq, off=off, beta=beta)
    rng = np.random.default_rng(3)
    theta_res = rng.normal(size=(N, d))
    logL_res = rng.normal(size=N) * 3
    props = rng.normal(size=(N * K, d))
    ll = rng.normal(size=N * K) * 3
    ll[::7] = np.nan
    ll[::9] = -1e20
    inside = (rng.random(N * K) > 0.1).astype(np.uint8)
    out_t, out_l = np.empty_like(theta_res), np.empty_like(logL_res)
    counts = np.zeros(2, dtype=np.int64)
    emu.emu_mutation_accept(C.byref(mut), dptr(props), dptr(ll), dptr(inside), C.c_longlong(N), dptr(theta_res), dptr(logL_res),
                            dptr(out_t), dptr(out_l), dptr(counts))
    cur_t, cur_l = theta_res.copy(), logL_res.copy()
    n_acc = n_cand = 0
    for i in range(N):
        for k in range(K):
            t = i * K + k
            if not inside[t]:
                continue
            n_cand += 1
            if not np.isfinite(ll[t]):
                continue
            u = emu.emu_accept_uniform(seed, seq, (off + i) * K + k)
            if np.log(u) < beta * ll[t] - beta * cur_l[i]:
                cur_t[i], cur_l[i] = props[t], ll[t]
                n_acc += 1
    assert np.array_equal(out_t, cur_t) and np.array_equal(out_l, cur_l)
    assert counts.tolist() == [n_acc, n_cand]
    assert not np.any(out_l == -1e20)


def _short_evaluator(cd, engines):
    kw = dict(workloads.PRODUCTION_SOLVER, maxtimestep=40)

    def factory(cfg):
        for o, v in enumerate([5, 12, 20, 30, 39]):
            cfg.cond[0].idx_sparse[o] = v
            cfg.cond[1].idx_sparse[o] = v
        e = OracleEngine(cfg)
        engines.append(e)
        return e

    theta_base = np.array(cd["bounds"]).mean(axis=1)
    return LogLikelihoodEvaluator(dict(kw, phi_init=cd["phi_init"]), [0, 1, 2, 3, 4], list(range(20)), theta_base, cd["data"],
                                  np.array([5, 12, 20, 30, 39]), cd["sigma_obs"], cov_rel=0.005, engine_factory=factory)


def test_run_tmcmc_with_device_mutation_on_the_oracle_engine():
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    engines = []
    ev = _short_evaluator(cd, engines)
    bounds = [tuple(b) for b in cd["bounds"]]
    res = tmcmc.run_TMCMC(ev, bounds, n_particles=120, n_stages=30, seed=5, evaluator=ev, theta_base_full=ev.theta_base,
                          active_indices=list(range(20)), n_mutation_steps=3, stage_ops=OracleStageOps(), device_mutation=True)
    assert res.converged and res.beta_schedule[-1] == 1.0
    lows, highs = np.array(bounds).T
    assert np.all((res.samples >= lows) & (res.samples <= highs))
    assert all(0.0 < a <= 1.0 for a in res.acc_rate_history)
    assert all(s["n_proposals"] > 0 and s["n_accepted"] <= s["n_proposals"] for s in res.stage_summary)
    # the returned log-likelihoods belong to the returned particles
    again = ev.batch(res.samples)
    assert np.allclose(again, res.logL_values, rtol=1e-12, atol=0)
    # same seed -> same result (counter-based stream); another seed -> another population
    ev2 = _short_evaluator(cd, engines)
    res2 = tmcmc.run_TMCMC(ev2, bounds, n_particles=120, n_stages=30, seed=5, evaluator=ev2, theta_base_full=ev2.theta_base,
                           active_indices=list(range(20)), n_mutation_steps=3, stage_ops=OracleStageOps(), device_mutation=True)
    assert np.array_equal(res2.samples, res.samples)
    # host path with the same seed: same initial population and beta_1 (the host stream is untouched by the option)
    ev3 = _short_evaluator(cd, engines)
    res3 = tmcmc.run_TMCMC(ev3, bounds, n_particles=120, n_stages=30, seed=5, evaluator=ev3, theta_base_full=ev3.theta_base,
                           active_indices=list(range(20)), n_mutation_steps=3, stage_ops=OracleStageOps())
    assert res3.beta_schedule[1] == res.beta_schedule[1]
    assert abs(res3.log_evidence - res.log_evidence) < 3.0  # same target, different random stream


def test_device_mutation_needs_a_capable_likelihood():
    with pytest.raises(ValueError):
        tmcmc.run_TMCMC(lambda t: 0.0, [(0.0, 1.0)], n_particles=10, n_stages=2, seed=1, stage_ops=OracleStageOps(), device_mutation=True)
