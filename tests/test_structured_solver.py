"""The product's kernel arithmetic (hmx_model / hmx_lane / hmx_obs .cuh) compiled for the CPU and compared
with the dense oracle.  This validates the block-structured direction solve, the per-lane state machine and
the observation-row sensitivities on machines without a GPU; `-m gpu` repeats the comparison on the device."""
import ctypes as C

import numpy as np
import pytest

from helpers import dptr, load_emu
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, workloads

emu = None


@pytest.fixture(scope="module", autouse=True)
def _emu(built):
    global emu
    emu = load_emu()
    emu.emu_solve5x2.restype = C.c_double
    return emu


CONFIGS = [
    dict(K_hill=0.05, n_hill=4.0, alpha_const=0.0, c_const=25.0),
    dict(K_hill=0.05, n_hill=2.0, alpha_const=100.0, c_const=100.0),
    dict(K_hill=0.0, n_hill=2.0, alpha_const=100.0, c_const=100.0),
    dict(K_hill=0.05, n_hill=2.5, alpha_const=0.0, c_const=25.0),
    dict(K_hill=0.05, n_hill=1.0, alpha_const=3.0, c_const=25.0),
    dict(K_hill=0.05, n_hill=3.0, alpha_const=0.0, c_const=25.0),
]


def test_ctypes_struct_layout_matches_header():
    assert emu.emu_sizeof_config() == C.sizeof(_abi.HmxConfig)
    assert emu.emu_sizeof_cond() == C.sizeof(_abi.HmxCond)
    assert emu.emu_sizeof_counters() == C.sizeof(_abi.HmxCounters)


@pytest.mark.parametrize("extra", CONFIGS, ids=lambda d: f"K{d['K_hill']}_n{d['n_hill']}_a{d['alpha_const']}")
def test_direction_equals_dense_lapack_solve(extra):
    """delta from the block-structured solve == np.linalg.solve(K_dense, -Q) with K, Q from the oracle, at
    non-converged points including clipped states and non-unit eta."""
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    kw = dict(dt=1e-4, maxtimestep=2500, Kp1=1e-4, eta_vec=[1.0, 1.3, 0.8, 1.1, 0.9], eta_phi_vec=[1.2, 0.7, 1.0, 1.5, 0.95], **extra)
    cfg = _abi.make_config([workloads.cond_struct(cd)], **kw)
    ocfg = orc.make_solver_cfg(phi_init=cd["phi_init"], **kw)
    b = np.array(cd["bounds"])
    rng = np.random.default_rng(7)
    worst_d = worst_q = 0.0
    for j in range(150):
        th = rng.uniform(b[:, 0], b[:, 1])
        g_old = np.concatenate([rng.uniform(1e-3, 0.3, 5), [0.1], rng.uniform(0.2, 0.999, 5), [rng.normal() * 10]])
        if j % 5 == 0:
            g_old[3] = 1e-9
        g_new = g_old + rng.normal(size=12) * np.array([1e-3] * 11 + [1.0])
        if j % 7 == 0:
            g_new[2] = -1e-3
        if j % 11 == 0:
            g_new[8] = 1.2
        if j % 13 == 0:
            g_new[5] = 1e-16
        Q, d = np.zeros(12), np.zeros(12)
        emu.emu_direction(C.byref(cfg), dptr(th), dptr(g_new.copy()), dptr(g_old.copy()), dptr(Q), dptr(d))
        Ko, Qo = orc.compute_J(ocfg, th, g_new, g_old), orc.compute_Q(ocfg, th, g_new, g_old)
        do = np.linalg.solve(Ko, -Qo)
        worst_d = max(worst_d, np.max(np.abs(d - do) / (np.abs(do) + 1e-300)))
        worst_q = max(worst_q, np.max(np.abs(Q - Qo)) / np.max(np.abs(Qo)))
    assert worst_q < 1e-13, worst_q
    assert worst_d < 1e-9, worst_d  # observed <= 1e-12 with cond(K) ~ 1e8


@pytest.mark.parametrize("extra", CONFIGS[:4], ids=lambda d: f"K{d['K_hill']}_n{d['n_hill']}_a{d['alpha_const']}")
def test_rebased_step_start_residual_equals_fresh_evaluation(extra):
    """The residual a new time step starts with, derived from the accepted trial's residual (rebase_residual),
    equals a fresh evaluation at the same point -- including trials that left the [1e-12, 1-1e-12] box, where
    the new time derivatives (clip(x) - x)/dt are not zero -- up to rounding of the largest term of Q."""
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    kw = dict(dt=1e-4, maxtimestep=2500, Kp1=1e-4, eta_vec=[1.0, 1.3, 0.8, 1.1, 0.9], eta_phi_vec=[1.2, 0.7, 1.0, 1.5, 0.95], **extra)
    cfg = _abi.make_config([workloads.cond_struct(cd)], **kw)
    b = np.array(cd["bounds"])
    rng = np.random.default_rng(21)
    for j in range(120):
        th = rng.uniform(b[:, 0], b[:, 1])
        g_old = np.concatenate([rng.uniform(1e-3, 0.3, 5), [0.1], rng.uniform(0.2, 0.999, 5), [rng.normal() * 10]])
        x = g_old + rng.normal(size=12) * np.array([1e-4] * 11 + [0.1])
        if j % 3 == 0:
            x[rng.integers(0, 5)] = -rng.uniform(0, 1e-5)      # phi below the box
        if j % 4 == 0:
            x[6 + rng.integers(0, 5)] = 1.0 + rng.uniform(0, 1e-5)  # psi above the box
        if j % 5 == 0:
            x[5] = -1e-6                                       # phi_0 is clipped inside Q only
        Qr, Qd, norms, fin = np.zeros(12), np.zeros(12), np.zeros(3), C.c_int()
        emu.emu_rebase(C.byref(cfg), dptr(th), dptr(x), dptr(g_old), dptr(Qr), dptr(Qd), dptr(norms), C.byref(fin))
        assert fin.value == 1
        scale = max(np.abs(Qd).max(), 1.0 / 1e-4)  # terms of size |x|/dt cancel inside Q
        assert np.max(np.abs(Qr - Qd)) <= 1e-12 * scale, (j, np.max(np.abs(Qr - Qd)), scale)
        assert abs(norms[0] - norms[1]) <= 1e-12 * scale
        assert norms[2] == 0.0  # everything else the direction solve reads is bitwise the same
    # a non-finite component must be reported (it doubles as the trajectory's isfinite scan)
    for k in (0, 5, 7):
        x = g_old.copy()
        x[k] = np.inf if k != 7 else np.nan
        emu.emu_rebase(C.byref(cfg), dptr(th), dptr(x), dptr(g_old), dptr(Qr), dptr(Qd), dptr(norms), C.byref(fin))
        assert fin.value == 0, k


def test_solve5x2_natural_and_pivoted_agree_with_numpy():
    rng = np.random.default_rng(11)
    for trial in range(50):
        S = np.eye(5) + 0.1 * rng.normal(size=(5, 5))
        rhs = rng.normal(size=(5, 2))
        M = np.ascontiguousarray(np.hstack([S, rhs]))
        ref = np.linalg.solve(S, rhs)
        for mode in (1, 0, 2):  # pivoted, natural-order LU, 2+3 block elimination (the kernel's fast path)
            X = np.zeros(10)
            health = emu.emu_solve5x2(dptr(M), dptr(X), mode)
            assert np.allclose(X.reshape(2, 5).T, ref, rtol=1e-12, atol=1e-14)
            assert mode == 1 or health > _abi_floor()  # fast paths report a healthy system
    # a system the natural order cannot handle (zero leading pivot): the guard must trip, the pivoted path solves it
    S = np.eye(5)[[1, 0, 2, 3, 4]] + 0.01 * rng.normal(size=(5, 5))
    S[0, 0] = 0.0
    rhs = rng.normal(size=(5, 2))
    M = np.ascontiguousarray(np.hstack([S, rhs]))
    X = np.zeros(10)
    assert not (emu.emu_solve5x2(dptr(M), dptr(X), 0) >= _abi_floor())
    emu.emu_solve5x2(dptr(M), dptr(X), 1)
    assert np.allclose(X.reshape(2, 5).T, np.linalg.solve(S, rhs), rtol=1e-11, atol=1e-13)
    # the 2+3 block path inverts the leading 2x2 block explicitly, so a row swap inside it is harmless ...
    assert emu.emu_solve5x2(dptr(M), dptr(X), 2) >= _abi_floor()
    assert np.allclose(X.reshape(2, 5).T, np.linalg.solve(S, rhs), rtol=1e-11, atol=1e-13)
    # ... but a (near-)singular leading block or Schur complement must trip its guard
    S2 = np.eye(5) + 0.01 * rng.normal(size=(5, 5))
    S2[1, :2] = S2[0, :2] * (1 + 1e-9)
    M2 = np.ascontiguousarray(np.hstack([S2, rhs]))
    assert not (emu.emu_solve5x2(dptr(M2), dptr(X), 2) >= _abi_floor())
    emu.emu_solve5x2(dptr(M2), dptr(X), 1)
    assert np.allclose(X.reshape(2, 5).T, np.linalg.solve(S2, rhs), rtol=1e-9, atol=1e-11)
    S3 = np.eye(5) + 0.01 * rng.normal(size=(5, 5))
    S3[4, 2:] = S3[3, 2:]
    S3[4, :2] = S3[3, :2]
    S3[4, 4] += 1e-3  # rows 3 and 4 almost identical: det T' ~ 1e-3
    M3 = np.ascontiguousarray(np.hstack([S3, rhs]))
    assert not (emu.emu_solve5x2(dptr(M3), dptr(X), 2) >= _abi_floor())


def _abi_floor():
    return 0.0625  # hmx::kPivotFloor


@pytest.mark.parametrize("key", ["Dysbiotic_HOBIC", "Commensal_Static", "Commensal_HOBIC", "Dysbiotic_Static", "Dysbiotic_HOBIC_legacy_1k30"])
def test_lane_state_machine_matches_oracle(key):
    """Full 2500-step solves: trajectory, iteration count and logL (TSM variance ON) of the emulated lane vs the oracle."""
    cd = workloads.load_conditions()[key]
    kw = dict(workloads.PRODUCTION_SOLVER)
    cfg = _abi.make_config([workloads.cond_struct(cd)], **kw)
    (ocfg, occ), = orc.from_hmx_config(cfg)
    rng = np.random.default_rng(5)
    b = np.array(cd["bounds"])
    for j in range(3):
        th = rng.uniform(b[:, 0], b[:, 1])
        g = np.zeros((cfg.n_steps + 1, 12))
        ll, st, it, tr = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert emu.emu_evaluate(C.byref(cfg), 0, dptr(th), dptr(g), C.byref(ll), C.byref(st), C.byref(it), C.byref(tr)) == 0
        _, go, its, trials, sto = orc.run_deterministic(ocfg, th)
        llo = orc.evaluate(ocfg, occ, th)
        rel = np.abs(g - go) / np.maximum(np.abs(go), 1e-300)
        assert rel.max() < 1e-9, rel.max()
        assert it.value == its  # same quasi-Newton iterate sequence
        assert trials < tr.value <= trials + its  # the line-search trials + the few full iteration-start residuals (step starts are rebased)
        assert abs(ll.value - llo) <= 1e-9 * abs(llo)
        assert (st.value & 1) == (sto & 1)


def test_lane_t500_alpha_hill_variants():
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    rng = np.random.default_rng(9)
    b = np.array(cd["bounds"])
    for extra in (dict(), dict(K_hill=0.05, n_hill=2.0), dict(K_hill=0.05, n_hill=2.5, alpha_const=0.0)):
        kw = dict(workloads.T500_SOLVER, **extra)
        cfg = _abi.make_config([workloads.cond_struct(cd, phi_init=0.2)], **kw)
        # observation rows of the 2500-step grid do not exist at T=500: use rows inside the trajectory
        for o, v in enumerate([50, 120, 260, 400, 499]):
            cfg.cond[0].idx_sparse[o] = v
        (ocfg, occ), = orc.from_hmx_config(cfg)
        th = rng.uniform(b[:, 0], b[:, 1])
        g = np.zeros((501, 12))
        ll, st, it, tr = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        assert emu.emu_evaluate(C.byref(cfg), 0, dptr(th), dptr(g), C.byref(ll), C.byref(st), C.byref(it), C.byref(tr)) == 0
        _, go, its, _, _ = orc.run_deterministic(ocfg, th)
        assert (np.abs(g - go) / np.maximum(np.abs(go), 1e-300)).max() < 1e-9
        llo = orc.evaluate(ocfg, occ, th)
        assert abs(ll.value - llo) <= 1e-9 * abs(llo)


def test_observation_row_zero_uses_first_step_sensitivities():
    """idx_sparse = 0 reads g_arr[0] but the sensitivities of row 1 (S4:1139-1141)."""
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    kw = dict(workloads.PRODUCTION_SOLVER, maxtimestep=60)
    hc = _abi.make_cond(cd["data"][:3], [0, 1, 55], cd["sigma_obs"][:3], cd["phi_init"], active_theta_indices=range(20))
    cfg = _abi.make_config([hc], **kw)
    (ocfg, occ), = orc.from_hmx_config(cfg)
    th = np.array(cd["bounds"]).mean(axis=1)
    ll, st, it, tr = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    assert emu.emu_evaluate(C.byref(cfg), 0, dptr(th), None, C.byref(ll), C.byref(st), C.byref(it), C.byref(tr)) == 0
    llo = orc.evaluate(ocfg, occ, th)
    assert abs(ll.value - llo) <= 1e-9 * abs(llo)


def test_config_validation():
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    good = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    assert emu.emu_validate(C.byref(good)) == 0
    bad = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    bad.cond[0].idx_sparse[4] = 2501  # beyond the trajectory: the reference raises IndexError (EV:682-697)
    assert emu.emu_validate(C.byref(bad)) != 0
    bad2 = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    bad2.n_cond = 0
    assert emu.emu_validate(C.byref(bad2)) != 0
    bad3 = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    bad3.abi_version = 99
    assert emu.emu_validate(C.byref(bad3)) != 0


def test_active_species_masks_against_live_reference_trajectories():
    """The lane state machine with inactive species (phi = psi = 0 at t = 0, held at the clip floor) and a vector
    phi_init, against trajectories of the live reference solver (tests/golden/solver_masked_kat.npz)."""
    from helpers import GOLDEN

    k = np.load(GOLDEN / "solver_masked_kat.npz")
    for name in ("four_species", "three_species_vec", "all_vec"):
        hc = _abi.make_cond(np.zeros((1, 5)), [400], 1.0, k[f"{name}_phi_init"], tsm_variance=False)
        cfg = _abi.make_config([hc], dt=1e-4, maxtimestep=400, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0,
                               active_species=list(k[f"{name}_active"]))
        g = np.zeros((401, 12))
        ll, st, it, tr = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        th = np.ascontiguousarray(k[f"{name}_theta"])
        assert emu.emu_evaluate(C.byref(cfg), 0, dptr(th), dptr(g), C.byref(ll), C.byref(st), C.byref(it), C.byref(tr)) == 0
        ref = k[f"{name}_g"]
        assert np.array_equal(g[0], ref[0])
        rel = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-300)
        assert rel.max() < 1e-9, (name, rel.max())
