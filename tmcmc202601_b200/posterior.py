"""Posterior summaries of a finished run: host-side, trivial cost.

Same names and return layouts as the reference's helpers in
data_5species/main/estimate_reduced_nishioka.py:378-456 (`compute_hdi`, `compute_credible_intervals`), so the
run-directory writers and plotters downstream can consume GPU results unchanged.
"""

from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def compute_hdi(samples: np.ndarray, credibility: float = 0.95) -> Tuple[np.ndarray, np.ndarray]:
    """Highest-density interval per parameter: the narrowest window of ceil(credibility * n) + 1 sorted
    samples (MAIN:378-423; all samples when the window covers everything)."""
    s = np.asarray(samples, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    n, d = s.shape
    k = int(np.ceil(credibility * n))
    srt = np.sort(s, axis=0)
    if k >= n:
        return srt[0].copy(), srt[-1].copy()
    widths = srt[k:] - srt[:-k]  # (n-k, d)
    j = np.argmin(widths, axis=0)
    cols = np.arange(d)
    return srt[j, cols], srt[j + k, cols]


def compute_credible_intervals(samples: np.ndarray, credibility: float = 0.95) -> Dict[str, np.ndarray]:
    """HDI + equal-tailed interval + median (MAIN:426-456)."""
    alpha = 1 - credibility
    lo, hi = compute_hdi(samples, credibility)
    return {
        "hdi_lower": lo, "hdi_upper": hi,
        "et_lower": np.percentile(samples, 100 * alpha / 2, axis=0),
        "et_upper": np.percentile(samples, 100 * (1 - alpha / 2), axis=0),
        "median": np.median(samples, axis=0), "credibility": credibility,
    }
