"""INTEGRATION.md, Level 2: the ctypes stub a reference maintainer would add is EXECUTED here, verbatim (the code block
is extracted from the document), instead of being trusted as prose.

* build container (`not gpu`): the stub's `_lib` is a fake whose two entry points are backed by the CPU oracle, and the
  evaluator it defines is driven by the LIVE reference `core.tmcmc.run_TMCMC` (patched with the one-line `.batch`
  dispatch the document prescribes for `evaluate_particles_parallel`) -- the evaluator surface (`call_count`, `timing`,
  linearization hooks, `get_health`) is exactly what the reference's driver touches.
* GPU box (`-m gpu`): the same stub against the real libhamilton.so, driven by this package's run_TMCMC; its values equal
  the package's own LogLikelihoodEvaluator bit for bit."""
import ctypes
import re
import sys
import types

import numpy as np
import pytest

from helpers import ROOT
from oracle import oracle as orc
from oracle import reference_live as rl
from tmcmc202601_b200 import _abi, workloads

SOLVER_KW = dict(dt=1e-4, maxtimestep=120, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0)


def _stub_source() -> str:
    text = (ROOT / "INTEGRATION.md").read_text()
    level2 = text[text.index("## Level 2"):text.index("## Level 3")]
    blocks = re.findall(r"```python\n(.*?)```", level2, flags=re.S)
    stub = [b for b in blocks if "class HamiltonGPUEvaluator" in b]
    assert len(stub) == 1 and any("hasattr(log_likelihood, \"batch\")" in b for b in blocks)
    return stub[0]


class _FakeLib:
    """libhamilton's two entry points the stub calls, answered by the CPU oracle (tests only)."""

    def __init__(self):
        self.cfg = None
        self.hmx_last_error = lambda ctx: b"fake"

    def hmx_create(self, ctx_ref, cfg_ref, device):
        self.cfg = _abi.HmxConfig.from_buffer_copy(bytes(cfg_ref._obj))  # same layout as include/hmx.h or the stub is wrong
        ctx_ref._obj.value = 1
        return 0

    def hmx_loglik_batch(self, ctx, theta_p, cond_p, n, logL_p, obs_p, st_p, it_p):
        n = n.value
        theta = np.ctypeslib.as_array((ctypes.c_double * (n * 20)).from_address(theta_p.value)).reshape(n, 20)
        out = np.ctypeslib.as_array((ctypes.c_double * n).from_address(logL_p.value))
        out[:] = orc.evaluate_hmx(self.cfg, theta.copy(), None)
        return 0


def _cond():
    cd = dict(workloads.load_conditions()["Dysbiotic_HOBIC"])
    cd["idx_sparse"] = np.array([10, 30, 60, 90, 118])  # the five observation days on a 120-step grid
    return cd


def _make_stub_evaluator(ns, cd):
    base = np.full(20, 1.5)
    return ns["HamiltonGPUEvaluator"](dict(SOLVER_KW, phi_init=cd["phi_init"]), cd["active_indices"], base, cd["data"], cd["idx_sparse"],
                                      cd["sigma_obs"], 0.005)


@pytest.mark.skipif(not rl.available(), reason="needs the live reference tree (build container)")
def test_level2_stub_drives_the_live_reference_run_TMCMC(monkeypatch):
    m = rl.bootstrap()  # puts the reference's `utils` package (TimingStats) on sys.path as the stub expects
    fake = _FakeLib()
    orc.lib()  # load the oracle before ctypes.CDLL is replaced
    real_cdll = ctypes.CDLL
    monkeypatch.setattr(ctypes, "CDLL", lambda name, *a, **k: fake if name == "libhamilton.so" else real_cdll(name, *a, **k))
    ns = {}
    exec(_stub_source(), ns)
    cd = _cond()
    ev = _make_stub_evaluator(ns, cd)
    assert ctypes.sizeof(ns["HmxConfig"]) == ctypes.sizeof(_abi.HmxConfig) and ctypes.sizeof(ns["HmxCond"]) == ctypes.sizeof(_abi.HmxCond)
    # the stub's values are the oracle's for the configuration it built
    b = np.array(cd["bounds"])
    th = np.random.default_rng(0).uniform(b[:, 0], b[:, 1], size=(5, 20))
    cfg = orc.make_solver_cfg(phi_init=cd["phi_init"], **SOLVER_KW)
    cc = orc.make_cond_cfg(cd["data"], cd["idx_sparse"], cd["sigma_obs"], cov_rel=0.005, active_theta_indices=cd["active_indices"])
    assert np.array_equal(ev.batch(th), orc.evaluate_batch(cfg, [cc], th))
    assert ev(th[0]) == ev.batch(th[:1])[0]
    # the live reference's TMCMC driver with the documented one-line dispatch in evaluate_particles_parallel
    tm = m["tmcmc"]
    live_epp = tm.evaluate_particles_parallel

    def patched(log_likelihood, theta, n_jobs=None, use_threads=False):
        if hasattr(log_likelihood, "batch"):
            return np.asarray(log_likelihood.batch(theta))
        return live_epp(log_likelihood, theta, n_jobs, use_threads)

    monkeypatch.setattr(tm, "evaluate_particles_parallel", patched)
    res = tm.run_TMCMC(ev, [tuple(x) for x in cd["bounds"]], n_particles=24, n_stages=6, seed=3, n_jobs=4, min_delta_beta=0.02,
                       max_delta_beta=0.5, n_mutation_steps=2, evaluator=ev, theta_base_full=np.full(20, 1.5),
                       active_indices=cd["active_indices"])
    assert res.samples.shape == (24, 20) and np.isfinite(res.logL_values).all() and res.beta_schedule[-1] <= 1.0
    assert ev.call_count >= 24 and not ev._linearization_enabled
    # every likelihood the reference's driver holds at the end is the oracle's value for that particle
    assert np.allclose(res.logL_values, orc.evaluate_batch(cfg, [cc], res.samples), rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_level2_stub_on_the_gpu_equals_the_package_evaluator(built, monkeypatch):
    from tmcmc202601_b200 import tmcmc
    from tmcmc202601_b200 import utils as pkg_utils
    from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

    real_cdll = ctypes.CDLL
    monkeypatch.setattr(ctypes, "CDLL", lambda name, *a, **k: real_cdll(str(_abi.LIB_PATH)) if name == "libhamilton.so" else real_cdll(name, *a, **k))
    fake_utils = types.ModuleType("utils")  # the reference's data_5species/utils is not on the GPU box
    fake_utils.TimingStats = pkg_utils.TimingStats
    monkeypatch.setitem(sys.modules, "utils", fake_utils)
    ns = {}
    exec(_stub_source(), ns)
    cd = _cond()
    ev = _make_stub_evaluator(ns, cd)
    mine = LogLikelihoodEvaluator(dict(SOLVER_KW, phi_init=cd["phi_init"]), [0, 1, 2, 3, 4], cd["active_indices"], np.full(20, 1.5), cd["data"],
                                  cd["idx_sparse"], cd["sigma_obs"], cov_rel=0.005, theta_linearization=np.full(20, 1e6))
    try:
        b = np.array(cd["bounds"])
        th = np.random.default_rng(1).uniform(b[:, 0], b[:, 1], size=(200, 20))
        assert np.array_equal(ev.batch(th), mine.batch(th))
        res = tmcmc.run_TMCMC(ev, [tuple(x) for x in cd["bounds"]], n_particles=64, n_stages=8, seed=5, min_delta_beta=0.02, max_delta_beta=0.5,
                              n_mutation_steps=2, evaluator=ev, theta_base_full=np.full(20, 1.5), active_indices=cd["active_indices"])
        assert res.samples.shape == (64, 20) and np.array_equal(res.logL_values, mine.batch(res.samples))
    finally:
        mine.close()
