"""Observation constants of the four Heine et al. conditions and the synthetic particle batches of
BASELINE.json `configs`.

`data/heine_conditions.json` holds the outputs of the reference loader (load_experimental_data,
convert_days_to_model_time, build_replicate_sigma, get_condition_bounds) at full precision; it was written
by tests/golden/make_golden.py in the build container and is read-only input here.
"""

from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi

_DATA = Path(__file__).resolve().parent / "data" / "heine_conditions.json"

CONDITION_KEYS = ["Commensal_Static", "Commensal_HOBIC", "Dysbiotic_Static", "Dysbiotic_HOBIC"]

# production solver settings: MAIN:2433-2779 defaults + the _runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30 config
PRODUCTION_SOLVER = dict(dt=1e-4, maxtimestep=2500, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0)
# the 500-step configuration of deeponet/surrogate_tmcmc.py:353-365 (also the solver's own defaults, S5:464-481)
T500_SOLVER = dict(dt=1e-5, maxtimestep=500, c_const=100.0, alpha_const=100.0, Kp1=1e-4, K_hill=0.0, n_hill=2.0)


def load_conditions() -> Dict[str, dict]:
    raw = json.loads(_DATA.read_text())["conditions"]
    out = {}
    for k, v in raw.items():
        d = dict(v)
        for f in ("data", "sigma_obs", "phi_init"):
            d[f] = np.asarray(d[f], dtype=np.float64)
        d["idx_sparse"] = np.asarray(d["idx_sparse"], dtype=np.int64)
        d["bounds"] = [tuple(b) for b in d["bounds"]]
        d["active_indices"] = [i for i in range(_abi.NTHETA) if i not in d["locked"]]
        out[k] = d
    return out


def cond_struct(cd: dict, cov_rel=0.005, tsm_variance=True, weights=None, use_absolute_volume=False, phi_init=None,
                sigma_obs=None) -> _abi.HmxCond:
    return _abi.make_cond(
        cd["data"], cd["idx_sparse"], cd["sigma_obs"] if sigma_obs is None else sigma_obs,
        cd["phi_init"] if phi_init is None else phi_init, cov_rel=cov_rel, active_theta_indices=cd["active_indices"],
        weights=weights, use_absolute_volume=use_absolute_volume, tsm_variance=tsm_variance,
    )


@dataclass
class Workload:
    name: str
    config: _abi.HmxConfig
    theta: np.ndarray  # [M,20] full 20-vectors
    cond_id: np.ndarray  # [M] int32
    keys: List[str]
    solver_kwargs: dict


def sample_prior(bounds: Sequence, n: int, rng: np.random.Generator) -> np.ndarray:
    b = np.asarray(bounds, dtype=np.float64)
    return rng.uniform(b[:, 0], b[:, 1], size=(n, b.shape[0]))


def four_condition_workload(n_per_condition: int, seed: int = 20260101, solver: Optional[dict] = None,
                            keys: Optional[Sequence[str]] = None, tsm_variance: bool = True,
                            interleave: bool = False) -> Workload:
    """theta_i ~ U(prior bounds of condition c), rng = default_rng(seed + c), n_per_condition per condition
    (BASELINE.json configs[4] / SURVEY 8(d) C5).  Locked parameters have bounds (0,0) and stay 0."""
    conds = load_conditions()
    keys = list(CONDITION_KEYS if keys is None else keys)
    solver = dict(PRODUCTION_SOLVER if solver is None else solver)
    structs, thetas, cids = [], [], []
    for c, k in enumerate(keys):
        cd = conds[k]
        structs.append(cond_struct(cd, tsm_variance=tsm_variance))
        rng = np.random.default_rng(seed + c)
        thetas.append(sample_prior(cd["bounds"], n_per_condition, rng))
        cids.append(np.full(n_per_condition, c, dtype=np.int32))
    theta = np.concatenate(thetas, axis=0)
    cid = np.concatenate(cids)
    if interleave:
        perm = np.random.default_rng(seed + 1000).permutation(theta.shape[0])
        theta, cid = theta[perm], cid[perm]
    cfg = _abi.make_config(structs, **solver)
    return Workload(f"synthetic {len(keys)}x{n_per_condition}", cfg, np.ascontiguousarray(theta), cid, keys, solver)


def single_condition_workload(key: str, n: int, seed: int = 42, solver: Optional[dict] = None,
                              tsm_variance: bool = True) -> Workload:
    return four_condition_workload(n, seed=seed, solver=solver, keys=[key], tsm_variance=tsm_variance)
