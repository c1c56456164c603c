"""Side-car helpers whose names the reference's callers rely on (data_5species/utils/timing.py:18-115,
data_5species/utils/health.py:13-83).  Host-side bookkeeping only."""

from __future__ import annotations

import time
from collections import defaultdict
from contextlib import contextmanager
from dataclasses import dataclass, field, asdict
from typing import Any, Dict

import numpy as np


@dataclass
class TimingStats:
    """Wall seconds + call counts per label (labels kept: tsm.solve_tsm, fom.run_deterministic; new: gpu.batch)."""

    seconds: Dict[str, float] = field(default_factory=lambda: defaultdict(float))
    counts: Dict[str, int] = field(default_factory=lambda: defaultdict(int))

    def add(self, name: str, dt_s: float) -> None:
        if not name or not np.isfinite(dt_s):
            return
        self.seconds[name] += float(dt_s)
        self.counts[name] += 1

    def get_s(self, name: str) -> float:
        return float(self.seconds.get(name, 0.0))

    def snapshot(self) -> Dict[str, Any]:
        return {"seconds": dict(self.seconds), "counts": dict(self.counts)}


@contextmanager
def timed(stats: TimingStats, name: str):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        stats.add(name, time.perf_counter() - t0)


@dataclass
class LikelihoodHealthCounter:
    n_calls: int = 0
    n_tsm_fail: int = 0
    n_output_nonfinite: int = 0
    n_var_raw_negative: int = 0
    n_var_raw_nonfinite: int = 0
    n_var_total_clipped: int = 0
    # GPU-side additions (status word of hmx_loglik_batch)
    n_max_newton_iter: int = 0
    n_singular_direction: int = 0

    def to_dict(self) -> Dict[str, int]:
        return {k: int(v) for k, v in asdict(self).items()}
