"""Host logic of the LogLikelihoodEvaluator mirror (theta mapping quirk, np.allclose shortcut, counters,
hooks) with the GPU engine replaced by an oracle-backed stand-in, and its values against the LIVE
reference evaluator (tests/golden/evaluator_kat.npz)."""
import json

import numpy as np
import pytest

from helpers import GOLDEN, OracleEngine
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.evaluator import (LogLikelihoodEvaluator, ViscoelasticPrior, build_likelihood_weights,
                                        build_species_sigma)


def _make(key, **kw):
    cd = workloads.load_conditions()[key]
    theta_base = np.full(20, 1.5)
    theta_base[cd["locked"]] = 0.0
    sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"])
    sig = cd["sigma_scalar"] if key.endswith("legacy_1k30") else cd["sigma_obs"]
    ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], cd["active_indices"], theta_base, cd["data"], cd["idx_sparse"], sig,
                                cov_rel=0.005, theta_linearization=theta_base, engine_factory=OracleEngine, **kw)
    return ev, cd, theta_base


def test_matches_live_reference_evaluator_values():
    k = np.load(GOLDEN / "evaluator_kat.npz")
    for key in ("Commensal_Static", "Dysbiotic_HOBIC_legacy_1k30"):
        ev, cd, base = _make(key)
        sel = [i for i, c in enumerate(k["cond"]) if str(c) == key]
        ll = ev.batch(k["theta_sub"][sel])
        assert np.allclose(ll, k["logL"][sel], rtol=1e-10, atol=0)
        # the engine received the reference's theta_full (incl. the lock/shift quirk for Commensal_Static)
        assert np.array_equal(ev._engine.calls[-1][0], k["theta_full"][sel])
        assert float(ev(k["theta_sub"][sel[0]])) == pytest.approx(k["logL"][sel[0]], rel=1e-10)


def test_theta_mapping_quirk_and_strict_mode():
    ev, cd, base = _make("Commensal_Static")
    active = cd["active_indices"]  # 9 free parameters
    sub = np.arange(1.0, 21.0)
    full = ev._full_theta(sub)[0]
    for i, idx in enumerate(active):
        assert full[idx] == sub[i]  # EV:635-637: positional, NOT by parameter index
    assert np.all(full[cd["locked"]] == 0.0)
    ev2, _, _ = _make("Commensal_Static", strict_mapping=True)
    full2 = ev2._full_theta(sub)[0]
    assert np.all(full2[active] == sub[active]) and np.all(full2[cd["locked"]] == 0.0)


def test_allclose_shortcut_selects_variance_free_condition():
    ev, cd, base = _make("Dysbiotic_HOBIC")
    th = np.vstack([base, base * (1 + 5e-6), base + 0.01])
    ev.batch(th)
    assert list(ev._engine.calls[-1][1]) == [1, 1, 0]  # np.allclose(theta, theta_lin): rtol 1e-5, atol 1e-8
    ev.update_linearization_point(base + 0.01)
    ev.batch(th)
    assert list(ev._engine.calls[-1][1]) == [0, 0, 1]


def test_counters_hooks_and_history():
    ev, cd, base = _make("Dysbiotic_HOBIC")
    b = np.array(cd["bounds"])
    th = np.random.default_rng(0).uniform(b[:, 0], b[:, 1], size=(3, 20))
    ll = ev.batch(th)
    assert ev.call_count == 3 and ev.get_health()["n_calls"] == 3 and ev.timing.get_s("tsm.solve_tsm") > 0
    v = ev(th[0])
    assert v == ll[0] and ev.call_count == 4 and len(ev.logL_history) == 1
    assert ev.get_MAP()[1] == v
    assert ev.compute_ROM_error(base) == 0.0 and ev.fom_call_count == 1
    ev.enable_linearization(True)
    assert ev._linearization_enabled and ev.compute_ROM_error(base) == float("inf")
    ev.enable_linearization(False)
    ev.update_linearization_point(th[1])
    assert np.array_equal(ev.get_linearization_point(), th[1]) and ev.call_count == 0
    nan_th = th[0].copy()
    nan_th[0] = np.nan
    assert ev(nan_th) == -1e20 and ev.get_health()["n_output_nonfinite"] == 1  # never raises (EV:655,667)


def test_constructor_checks():
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"])
    with pytest.raises(IndexError):
        LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], list(range(20)), np.zeros(20), cd["data"], np.array([1, 2, 3, 4, 9999]),
                               0.1, 0.005, engine_factory=OracleEngine)
    with pytest.raises(NotImplementedError):
        LogLikelihoodEvaluator(sk, [0, 1, 2, 3], list(range(14)), np.zeros(14), cd["data"][:, :4], cd["idx_sparse"], 0.1, 0.005,
                               engine_factory=OracleEngine)
    with pytest.raises(NotImplementedError):
        LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], list(range(20)), np.zeros(20), cd["data"], cd["idx_sparse"], 0.1, 0.005, rho=0.3,
                               engine_factory=OracleEngine)


def test_sigma_and_weight_builders_and_ve_prior():
    assert np.array_equal(build_species_sigma(0.1), [0.1, 0.1, 0.2, 0.1, 0.1])
    W = build_likelihood_weights(6, 5)
    assert W[0, 0] == 1 and W[0, 4] == 5 and W[5, 0] == 3 and W[5, 4] == 15 and W[3, 4] == 5
    p = ViscoelasticPrior(list(range(22)))
    t = np.zeros(22)
    t[20], t[21] = 1.5, 2.0
    assert p.log_prior(t) == pytest.approx(-0.5 * (1.0 + 1.0))
    assert ViscoelasticPrior(list(range(20))).log_prior(t) == 0.0
