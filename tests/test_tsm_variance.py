"""SURVEY 8(a) row a9 pinned DIRECTLY: the TSM sensitivities x1 and the variance sigma2 at the observation rows
against the live reference's `tsm._last_x1`, `_last_var_act` and `sig2` (tests/golden/tsm_kat.npz, written by
make_golden_tsm.py from BiofilmTSM.solve_tsm, S4:1077-1160).

The scalar log-likelihood does not constrain the variance at the 1e-9 bar (it moves logL by 8.6e-10
relative in the legacy configuration), so the variance needs its own assertion.  CPU half: the oracle's
orc_sensitivity_row against the live vectors.  GPU half (-m gpu): hmx_tsm_obs_rows through the C ABI."""
import json

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import oracle as orc
from tmcmc202601_b200 import _abi, workloads

REL = 1e-9


def _cases():
    k = np.load(GOLDEN / "tsm_kat.npz")
    kw = json.loads(str(k["solver_kwargs"]))
    for i, name in enumerate(k["names"]):
        n = int(k["n_obs"][i])
        act = k["active"][i]
        yield dict(name=str(name), cond=str(k["cond"][i]), theta=k["theta_full"][i], act=act[act >= 0], n_obs=n,
                   idx=k["idx_sparse"][i][:n], logL=float(k["logL"][i]), x0=k["x0"][i][:n], sig2=k["sig2"][i][:n],
                   x1=k["x1"][i][:n], var_act=k["var_act"][i], kw=kw, extra_rows=k["extra_rows"], x0_extra=k["x0_extra"][i],
                   sig2_extra=k["sig2_extra"][i], x1_extra=k["x1_extra"][i], sig2_all_max=float(k["sig2_all_max"][i]),
                   sig2_all_finite=bool(k["sig2_all_finite"][i]))


CASES = list(_cases())


def _scaled(a, b):
    """max |a - b| relative to the largest entry of the reference block (x1 has structural zeros)."""
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_oracle_sensitivities_match_live_reference(case):
    conds = workloads.load_conditions()
    cd = conds[case["cond"]]
    cfg = orc.make_solver_cfg(phi_init=cd["phi_init"], **case["kw"])
    _, g, *_ = orc.run_deterministic(cfg, case["theta"])
    mask = sum(1 << int(a) for a in case["act"])
    var = (0.005 * case["theta"]) ** 2
    rows = list(case["idx"]) + list(case["extra_rows"])
    x1_ref = np.concatenate([case["x1"], case["x1_extra"]])
    s2_ref = np.concatenate([case["sig2"], case["sig2_extra"]])
    x0_ref = np.concatenate([case["x0"], case["x0_extra"]])
    for r, x1r, s2r, x0r in zip(rows, x1_ref, s2_ref, x0_ref):
        t = max(int(r), 1)
        assert np.abs(g[r] - x0r).max() <= REL * np.abs(x0r).max()
        x1 = orc.sensitivity_row(cfg, case["theta"], g[t], g[t - 1], mask)
        assert _scaled(x1, x1r) < REL, (case["name"], r)
        s2 = (x1[:, case["act"]] ** 2 * var[case["act"]][None, :]).sum(axis=1) + 1e-12
        assert (np.abs(s2 - s2r) / np.abs(s2r)).max() < REL, (case["name"], r)
    # var_act of the reference (S4:1144-1145)
    assert np.allclose(var[case["act"]], case["var_act"][: len(case["act"])], rtol=1e-15, atol=0)


def test_whole_trajectory_variance_of_the_live_reference_is_far_from_overflow():
    """EV:657-667 rejects a theta whose sig2 is non-finite at ANY of the T+1 rows.  On the live reference's own
    trajectories the variance never exceeds O(1): the guard can only fire together with a non-finite state (which
    the kernel scans at every row) or a numerically singular Jacobian."""
    for c in CASES:
        assert c["sig2_all_finite"] and c["sig2_all_max"] < 1e3


@pytest.mark.gpu
@pytest.mark.parametrize("cond_key", sorted({c["cond"] for c in CASES}))
def test_gpu_sigma2_and_x1_match_live_reference(built, cond_key):
    from tmcmc202601_b200.engine import HamiltonEngine

    cases = [c for c in CASES if c["cond"] == cond_key]
    conds = workloads.load_conditions()
    cd = conds[cond_key]
    sig = cd["sigma_scalar"] if cond_key.endswith("legacy_1k30") else cd["sigma_obs"]
    structs = [workloads.cond_struct(cd, sigma_obs=sig), workloads.cond_struct(cd, sigma_obs=sig, tsm_variance=False)]
    eng = HamiltonEngine(_abi.make_config(structs, **cases[0]["kw"]), device=0)
    try:
        th = np.array([c["theta"] for c in cases])
        x0, s2, x1, ll, st = eng.tsm_obs_rows(th, np.zeros(len(cases), np.int32))
        ll_plain = eng.loglik_batch(th, np.zeros(len(cases), np.int32))
        assert np.array_equal(ll, ll_plain)  # same kernels, same numbers
        for i, c in enumerate(cases):
            n = c["n_obs"]
            assert np.abs(x0[i, :n] - c["x0"]).max() <= REL * np.abs(c["x0"]).max()
            rel_s2 = np.abs(s2[i, :n] - c["sig2"]) / np.abs(c["sig2"])
            print(f"{c['name']}: sigma2 max rel {rel_s2.max():.2e}, x1 scaled {_scaled(x1[i, :n], c['x1']):.2e}")
            assert rel_s2.max() < REL, c["name"]
            for o in range(n):
                assert _scaled(x1[i, o], c["x1"][o]) < REL, (c["name"], o)
            inactive = [k for k in range(20) if k not in set(c["act"].tolist())]
            assert not x1[i, :n][:, :, inactive].any()
            assert abs(ll[i] - c["logL"]) <= REL * abs(c["logL"])
            assert st[i] & 1 == 0
        # the np.allclose short-cut condition returns a zero variance (TSMA:366-401)
        _, s2z, x1z, llz, _ = eng.tsm_obs_rows(th, np.ones(len(cases), np.int32))
        assert not s2z.any() and not x1z.any()
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_sensitivities_match_oracle_on_prior_draws(built):
    """64 prior draws per condition: x1 and sigma2 at every observation row against orc_sensitivity_row."""
    from tmcmc202601_b200.engine import HamiltonEngine

    wl = workloads.four_condition_workload(16, seed=77)
    eng = HamiltonEngine(wl.config, device=0)
    try:
        x0, s2, x1, ll, st = eng.tsm_obs_rows(wl.theta, wl.cond_id)
        pairs = orc.from_hmx_config(wl.config)
        worst_x1 = worst_s2 = 0.0
        for i in range(wl.theta.shape[0]):
            scfg, cc = pairs[int(wl.cond_id[i])]
            _, g, *_ = orc.run_deterministic(scfg, wl.theta[i])
            var = (cc.cov_rel * wl.theta[i]) ** 2
            for o in range(cc.n_obs):
                t = max(int(cc.idx_sparse[o]), 1)
                x1o = orc.sensitivity_row(scfg, wl.theta[i], g[t], g[t - 1], cc.active_theta_mask)
                s2o = (x1o ** 2 * var[None, :]).sum(axis=1) + 1e-12
                worst_x1 = max(worst_x1, _scaled(x1[i, o], x1o))
                worst_s2 = max(worst_s2, float((np.abs(s2[i, o] - s2o) / s2o).max()))
        print(f"x1 scaled max {worst_x1:.2e}, sigma2 rel max {worst_s2:.2e}")
        assert worst_x1 < REL and worst_s2 < REL
    finally:
        eng.close()


@pytest.mark.gpu
def test_bad_cond_id_is_rejected_not_evaluated_as_condition_zero(built):
    import torch

    from tmcmc202601_b200.engine import HamiltonEngine, HmxError

    wl = workloads.four_condition_workload(4, seed=5)
    eng = HamiltonEngine(wl.config, device=0)
    try:
        bad = wl.cond_id.copy()
        bad[3] = 7
        logL = np.empty(bad.size)
        rc = eng._lib.hmx_loglik_batch(eng._ctx, wl.theta.ctypes.data, bad.ctypes.data, bad.size, logL.ctypes.data, None, None, None)
        assert rc == -1 and b"cond_id[3]" in eng._lib.hmx_last_error(eng._ctx)
        # device array: cannot be checked without a sync -> flagged per evaluation
        eng.use_torch_stream()
        th = torch.from_numpy(wl.theta).cuda()
        cid = torch.from_numpy(bad).cuda()
        ll = torch.empty(bad.size, dtype=torch.float64, device="cuda")
        st = torch.zeros(bad.size, dtype=torch.int32, device="cuda")
        eng.loglik_batch_device(th, cid, ll, status=st)
        torch.cuda.synchronize()
        assert ll[3].item() == -1e20 and (st[3].item() & _abi.ST_BAD_COND)
        good = eng.loglik_batch(wl.theta, wl.cond_id)
        keep = np.arange(bad.size) != 3
        assert np.array_equal(ll.cpu().numpy()[keep], good[keep])
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_full_trajectory_variance_and_all_rows_guard(built):
    """hmx_solve_tsm_batch (BiofilmTSM.solve_tsm in full) against the live reference's sig2 at rows off the observation grid,
    and the all-rows finiteness guard of EV:657-667 (hmx_set_sig2_guard): same values, same flags as the default."""
    from tmcmc202601_b200.engine import HamiltonEngine
    from tmcmc202601_b200.solver import BiofilmNewtonSolver5S
    from tmcmc202601_b200.tsm import BiofilmTSM

    conds = workloads.load_conditions()
    for c in CASES[:2] + CASES[-1:]:
        cd = conds[c["cond"]]
        solver = BiofilmNewtonSolver5S(phi_init=cd["phi_init"], **c["kw"])
        tsm = BiofilmTSM(solver, active_theta_indices=c["act"], cov_rel=0.005)
        t_arr, x0, s2 = tsm.solve_tsm(c["theta"])
        assert x0.shape == (2501, 12) and s2.shape == (2501, 12) and t_arr.shape == (2501,)
        rows = np.concatenate([c["idx"], c["extra_rows"]])
        ref_s2 = np.concatenate([c["sig2"], c["sig2_extra"]])
        ref_x0 = np.concatenate([c["x0"], c["x0_extra"]])
        rel = np.abs(s2[rows] - ref_s2) / np.abs(ref_s2)
        print(f"{c['name']}: full-trajectory sigma2 vs live reference at {len(rows)} rows: max rel {rel.max():.2e}")
        assert rel.max() < REL and np.abs(x0[rows] - ref_x0).max() <= REL * np.abs(ref_x0).max()
        assert np.array_equal(s2[0], s2[1])  # x1[0] := x1[1], S4:1139-1141
        assert np.isfinite(s2).all() and abs(np.abs(s2).max() - c["sig2_all_max"]) <= 1e-9 * c["sig2_all_max"]
        # rows interface == the reference's _last_x1 rows
        x0r, s2r, x1r = tsm.solve_tsm_rows(c["theta"], c["extra_rows"])
        assert _scaled(x1r, c["x1_extra"][:, :, c["act"]]) < REL and np.array_equal(s2r, s2[c["extra_rows"]])
        tsm.close()
    # the guard: all rows vs observation rows -- identical values and flags on prior draws (it never fires on finite solves)
    wl = workloads.four_condition_workload(24, seed=99)
    eng = HamiltonEngine(wl.config, device=0)
    try:
        ll0, st0, _ = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        eng.set_sig2_guard(True)
        ll1, st1, _ = eng.loglik_batch(wl.theta, wl.cond_id, return_status=True)
        assert np.array_equal(ll0, ll1) and np.array_equal(st0, st1)
        # ... and it does fire where the variance overflows: cov_rel so large that sig2 = inf at every row
        bad_cfg = workloads.four_condition_workload(2, seed=99).config
        for k in range(4):
            bad_cfg.cond[k].cov_rel = 1e200
        e2 = HamiltonEngine(bad_cfg, device=0)
        e2.set_sig2_guard(True)
        llb, stb, _ = e2.loglik_batch(wl.theta[:4], wl.cond_id[:4], return_status=True)
        assert np.all(llb == -1e20) and np.all(stb & 1)
        e2.close()
    finally:
        eng.close()


def test_factor_once_sensitivities_match_oracle_on_the_host():
    """tsm_row (hmx_obs.cuh: direction_factor once per row + direction_apply per parameter) compiled for the CPU against the
    oracle's dense sensitivity solve, on rows of oracle trajectories for every condition (incl. early, stiff rows)."""
    import ctypes as C

    from helpers import dptr, load_emu

    emu = load_emu()
    wl = workloads.four_condition_workload(3, seed=5)
    pairs = orc.from_hmx_config(wl.config)
    worst_x1 = worst_s2 = 0.0
    for i in range(wl.theta.shape[0]):
        c = int(wl.cond_id[i])
        scfg, cc = pairs[c]
        th = np.ascontiguousarray(wl.theta[i])
        _, g, *_ = orc.run_deterministic(scfg, th)
        var = (cc.cov_rel * th) ** 2
        for t in (1, 2, 40, 339, 1200, 2500):
            x1o = orc.sensitivity_row(scfg, th, g[t], g[t - 1], cc.active_theta_mask)
            s2o = (x1o ** 2 * var[None, :]).sum(axis=1) + 1e-12
            s2, x1 = np.empty(12), np.empty((12, 20))
            gn, go = np.ascontiguousarray(g[t]), np.ascontiguousarray(g[t - 1])
            assert emu.emu_tsm_row(C.byref(wl.config), c, dptr(th), dptr(gn), dptr(go), dptr(s2), dptr(x1)) == 1
            worst_x1 = max(worst_x1, _scaled(x1, x1o))
            worst_s2 = max(worst_s2, float((np.abs(s2 - s2o) / s2o).max()))
    assert worst_x1 < REL and worst_s2 < REL, (worst_x1, worst_s2)
