"""`np.linalg.solve` and its singular-matrix retry `K + 1e-10 I` (S5:369-373).

The production kernel never forms K; when its block-structured solve yields a non-finite direction the evaluation is
flagged and solved again by K1r (csrc/hmx_dense.cuh): the reference's loop nest with the dense 12x12 matrix, pivoted LU and
the ridge retry.  CPU half: that code compiled for the host against the oracle, on ordinary evaluations and on a degenerate
configuration in which K is EXACTLY singular (the oracle's dgesv restatement hits a zero pivot and takes the ridge path).
GPU half (-m gpu): the same through the C ABI, plus K1r against the production kernel on every evaluation of a batch."""
import ctypes as C

import numpy as np
import pytest

from helpers import load_emu
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, workloads

# species blocks with phi = psi, no barrier, no interaction, eta_phi = 0: every 2x2 block of K is exactly singular and so is K
DEGENERATE = dict(dt=1e-4, maxtimestep=6, c_const=0.0, alpha_const=0.0, Kp1=0.0, K_hill=0.0, n_hill=2.0, eta_phi_vec=[0.0] * 5)


def _emu_loop(hcfg, th):
    emu = load_emu()
    g = np.empty((hcfg.n_steps + 1, 12))
    st, it = C.c_uint32(), C.c_uint32()
    assert emu.emu_reference_loop(C.byref(hcfg), 0, np.ascontiguousarray(th).ctypes.data_as(C.c_void_p), g.ctypes.data_as(C.c_void_p),
                                  C.byref(st), C.byref(it)) == 0
    return g, st.value, it.value


def test_reference_loop_with_dense_solve_matches_oracle_on_ordinary_evaluations():
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    b = np.array(cd["bounds"])
    rng = np.random.default_rng(4)
    for kw in (dict(workloads.PRODUCTION_SOLVER, maxtimestep=400), dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=2.5)):
        th = rng.uniform(b[:, 0], b[:, 1])
        hcfg = _abi.make_config([workloads.cond_struct(cd, tsm_variance=False)], **kw)
        g, st, it = _emu_loop(hcfg, th)
        _, go, its, _, sto = orc.run_deterministic(orc.make_solver_cfg(phi_init=cd["phi_init"], **kw), th)
        rel = np.abs(g - go) / np.maximum(np.abs(go), 1e-300)
        assert rel.max() < 1e-9 and st == sto and abs(it - its) <= max(2, its // 200)


def test_ridge_retry_on_an_exactly_singular_matrix_matches_oracle():
    hcfg = _abi.make_config([_abi.make_cond(np.zeros((0, 5)), [], 1.0, 0.999, tsm_variance=False)], **DEGENERATE)
    th = np.ones(20)
    g, st, it = _emu_loop(hcfg, th)
    _, go, its, _, sto = orc.run_deterministic(orc.make_solver_cfg(phi_init=0.999, **DEGENERATE), th)
    assert sto & orc.ST_SINGULAR and st & _abi.ST_SINGULAR  # both took the K + 1e-10 I path
    assert np.isfinite(go).all()
    # the first implicit step (one ridge solve) agrees tightly; later steps amplify rounding of an ill-posed problem
    assert np.abs(g[1] - go[1]).max() <= 1e-9 * np.abs(go[1]).max()
    assert np.abs(g[2] - go[2]).max() <= 1e-6 * np.abs(go[2]).max()
    # the block-structured solve cannot express this: the lane emulation gives up (non-finite direction -> flagged)
    emu = load_emu()
    ll, st2, its2, trips2 = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    emu.emu_evaluate(C.byref(hcfg), 0, th.ctypes.data_as(C.c_void_p), None, C.byref(ll), C.byref(st2), C.byref(its2), C.byref(trips2))
    assert st2.value & _abi.ST_SINGULAR


@pytest.mark.gpu
def test_gpu_flagged_evaluations_are_resolved_by_the_dense_kernel(built):
    from tmcmc202601_b200.solver import BiofilmNewtonSolver5S

    s = BiofilmNewtonSolver5S(phi_init=0.999, **DEGENERATE)
    th = np.ones((3, 20)) * np.array([1.0, 0.5, 2.0])[:, None]
    g, st = s.run_deterministic_batch(th, return_status=True)
    for i in range(3):
        _, go, *_rest, sto = orc.run_deterministic(orc.make_solver_cfg(phi_init=0.999, **DEGENERATE), th[i])
        assert st[i] & _abi.ST_SINGULAR and sto & orc.ST_SINGULAR
        assert np.isfinite(g[i]).all()
        assert np.abs(g[i, 1] - go[1]).max() <= 1e-9 * np.abs(go[1]).max()


@pytest.mark.gpu
def test_gpu_dense_kernel_agrees_with_production_kernel_on_every_evaluation(built, monkeypatch):
    """HMX_FORCE_REFERENCE_LOOP=1 sends every evaluation through K1r: two independent implementations of the same iterate
    sequence (per-lane state machine + block-structured solve vs nested loops + dense LU) on 4 x 64 prior draws."""
    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(64, seed=12)
    eng = HamiltonEngine(wl.config, device=0)
    monkeypatch.setenv("HMX_FORCE_REFERENCE_LOOP", "1")
    eng_ref = HamiltonEngine(wl.config, device=0)
    try:
        ll, st, it = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        ll_r, st_r, it_r = eng_ref.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        rel = np.abs(ll - ll_r) / np.abs(ll_r)
        same = np.mean(it.astype(np.int64) == it_r.astype(np.int64))
        print(f"K1 vs K1r: logL max rel {rel.max():.2e}, iteration counts identical on {100 * same:.1f} %")
        assert rel.max() < 1e-9 and np.array_equal(st & 1, st_r & 1) and same > 0.8
    finally:
        eng.close()
        eng_ref.close()
