"""Bootstrap of the LIVE reference (keisuke58/Tmcmc202601) for oracle pinning.

TEST INFRASTRUCTURE ONLY, and BUILD-CONTAINER ONLY: /root/reference does not exist on the GPU box, so
nothing under tests -m gpu, smoke() or bench.py may import this module.  It is used by
tests/golden/make_golden.py (which writes the committed fixtures) and by the optional
`tests/test_oracle_vs_live_reference.py` (skipped when the reference tree is absent).

Recipe: SURVEY.md Appendix C (matplotlib/seaborn are not installed -> MagicMock stubs; sys.path order as
data_5species/main/estimate_reduced_nishioka.py:63-74).
"""

from __future__ import annotations

import logging
import os
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

REF = Path(os.environ.get("HMX_REFERENCE_ROOT", "/root/reference"))

CONDITIONS = [
    ("Commensal", "Static"),
    ("Commensal", "HOBIC"),
    ("Dysbiotic", "Static"),
    ("Dysbiotic", "HOBIC"),
]

REP_SPECIES_MAP = {  # MAIN:2041-2049
    "S. oralis": 0,
    "A. naeslundii": 1,
    "V. dispar": 2,
    "V. parvula": 2,
    "F. nucleatum": 3,
    "P. gingivalis_20709": 4,
    "P. gingivalis_W83": 4,
}


def available() -> bool:
    return (REF / "tmcmc" / "program2602" / "improved_5species_jit.py").exists()


_mods = None


def bootstrap():
    """Import the reference modules once; returns a namespace dict."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")
    for m in (
        "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches",
        "matplotlib.colors", "matplotlib.ticker", "matplotlib.lines", "matplotlib.cm", "seaborn",
    ):
        sys.modules.setdefault(m, MagicMock())
    for p in (REF / "data_5species", REF / "tmcmc" / "program2602", REF / "tmcmc", REF):
        sys.path.insert(0, str(p))
    logging.disable(logging.CRITICAL)
    from improved_5species_jit import BiofilmNewtonSolver5S  # noqa: E402
    from core.evaluator import (  # noqa: E402
        LogLikelihoodEvaluator, log_likelihood_sparse, build_replicate_sigma, build_likelihood_weights,
    )
    from core import tmcmc as ref_tmcmc  # noqa: E402
    from core.nishioka_model import get_condition_bounds  # noqa: E402
    # data_5species/main/__init__.py imports stale scripts: load the CLI module by path
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "estimate_reduced_nishioka", str(REF / "data_5species" / "main" / "estimate_reduced_nishioka.py")
    )
    ref_main = importlib.util.module_from_spec(spec)
    sys.modules["estimate_reduced_nishioka"] = ref_main
    spec.loader.exec_module(ref_main)

    _mods = dict(
        BiofilmNewtonSolver5S=BiofilmNewtonSolver5S,
        LogLikelihoodEvaluator=LogLikelihoodEvaluator,
        log_likelihood_sparse=log_likelihood_sparse,
        build_replicate_sigma=build_replicate_sigma,
        build_likelihood_weights=build_likelihood_weights,
        tmcmc=ref_tmcmc,
        get_condition_bounds=get_condition_bounds,
        main=ref_main,
    )
    return _mods


def load_condition(condition: str, cultivation: str, dt=1e-4, maxtimestep=2500):
    """Loader outputs with the CLI defaults (use_exp_init, normalize, replicate sigma): SURVEY Appendix D."""
    m = bootstrap()
    data_dir = REF / "data_5species"
    data, t_days, sigma_scalar, phi_init_exp, meta = m["main"].load_experimental_data(
        data_dir, condition, cultivation, start_from_day=1, normalize=True, use_exp_init=True
    )
    _, idx_sparse = m["main"].convert_days_to_model_time(t_days, dt, maxtimestep, None)
    rep_csv = data_dir / "experiment_data" / "fig3_species_distribution_replicates.csv"
    sigma = m["build_replicate_sigma"](
        replicate_csv=str(rep_csv), condition=condition, cultivation=cultivation, days=meta["days"],
        species_map=REP_SPECIES_MAP, n_species=data.shape[1], min_sigma=0.005,
    )
    bounds, locked = m["get_condition_bounds"](condition, cultivation)
    return dict(
        condition=condition, cultivation=cultivation, data=np.asarray(data, float),
        days=[int(d) for d in meta["days"]], idx_sparse=np.asarray(idx_sparse, int),
        sigma_scalar=float(sigma_scalar), sigma_obs=np.asarray(sigma, float),
        phi_init=np.asarray(phi_init_exp, float), bounds=[(float(a), float(b)) for a, b in bounds],
        locked=[int(i) for i in locked],
    )


def make_evaluator(cond: dict, solver_kwargs: dict, cov_rel=0.005, weights=None, sigma_obs=None,
                   use_absolute_volume=False):
    """The closure at MAIN:2181-2239 for one condition."""
    m = bootstrap()
    locked = cond["locked"]
    active = [i for i in range(20) if i not in locked]
    theta_base = np.full(20, 0.0)  # PRIOR_BOUNDS_DEFAULT mean; locked -> 0.0 (MAIN:1937-1972)
    lo, hi = m["main"].PRIOR_BOUNDS_DEFAULT
    theta_base[:] = (lo + hi) / 2.0
    for i in locked:
        theta_base[i] = 0.0
    ev = m["LogLikelihoodEvaluator"](
        solver_kwargs=solver_kwargs, active_species=[0, 1, 2, 3, 4], active_indices=active,
        theta_base=theta_base, data=cond["data"], idx_sparse=cond["idx_sparse"],
        sigma_obs=cond["sigma_obs"] if sigma_obs is None else sigma_obs, cov_rel=cov_rel, rho=0.0,
        theta_linearization=theta_base, paper_mode=False, use_absolute_volume=use_absolute_volume,
        weights=weights,
    )
    ev.active_indices = list(active)
    ev.theta_base = theta_base.copy()
    return ev, theta_base, active
