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


def compute_gelman_rubin_rhat(chains) -> np.ndarray:
    """Potential scale reduction per parameter (MAIN:207-258): sqrt(((n-1)/n W + B/n) / (W + 1e-10)) with chains
    trimmed to the shortest one; NaN for fewer than two chains."""
    chains = [np.asarray(c, dtype=np.float64) for c in chains]
    if len(chains) < 2:
        return np.full(chains[0].shape[1], np.nan)
    n = min(len(c) for c in chains)
    x = np.stack([c[:n] for c in chains])            # (m, n, d)
    B = n * np.var(x.mean(axis=1), axis=0, ddof=1)
    W = np.var(x, axis=1, ddof=1).mean(axis=0)
    return np.sqrt((((n - 1) / n) * W + B / n) / (W + 1e-10))


def compute_effective_sample_size(samples: np.ndarray, max_lag=None) -> np.ndarray:
    """n / max(1, 1 + 2 sum rho_k) with the autocorrelations summed up to the first negative one and at most
    min(n // 2, 1000) lags (MAIN:261-304); autocorrelation by zero-padded FFT."""
    s = np.asarray(samples, dtype=np.float64)
    n, d = s.shape
    if max_lag is None:
        max_lag = min(n // 2, 1000)
    xc = s - s.mean(axis=0)
    f = np.fft.fft(xc, n=2 * n, axis=0)
    acf = np.fft.ifft(f * np.conj(f), axis=0)[:n].real
    with np.errstate(invalid="ignore", divide="ignore"):  # a constant column (locked parameter) gives NaN, as in the reference
        acf = acf / acf[0]
    ess = np.empty(d)
    top = min(max_lag, n)
    for p in range(d):
        lags = acf[1:top, p]
        neg = np.nonzero(lags < 0)[0]
        stop = neg[0] if neg.size else lags.size
        ess[p] = n / max(1.0 + 2.0 * float(np.sum(lags[:stop])), 1.0)
    return ess


def compute_mcmc_diagnostics(chains, logL, param_names=None) -> Dict[str, object]:
    """R-hat over the chains + ESS of the concatenated samples + the convergence flags the reference writes to
    mcmc_diagnostics.json (MAIN:307-371)."""
    allS = np.concatenate([np.asarray(c, dtype=np.float64) for c in chains], axis=0)
    d = allS.shape[1]
    names = list(param_names) if param_names is not None else [f"theta_{i}" for i in range(d)]
    rhat = compute_gelman_rubin_rhat(chains)
    ess = compute_effective_sample_size(allS)
    means, stds = allS.mean(axis=0), allS.std(axis=0)
    rhat_ok, ess_ok = bool(np.all(rhat < 1.1)), bool(np.all(ess > 100))
    return {
        "rhat": rhat, "ess": ess, "means": means, "stds": stds, "n_samples_total": int(len(allS)), "n_chains": len(chains),
        "rhat_max": float(np.nanmax(rhat)) if np.any(np.isfinite(rhat)) else float("nan"), "ess_min": float(np.nanmin(ess)),
        "converged_rhat": rhat_ok, "converged_ess": ess_ok, "converged": bool(rhat_ok and ess_ok), "param_names": names,
        "per_parameter": [
            {"name": nm, "mean": float(means[i]), "std": float(stds[i]), "rhat": None if np.isnan(rhat[i]) else float(rhat[i]),
             "ess": float(ess[i])} for i, nm in enumerate(names)
        ],
    }


def compute_phibar(x0: np.ndarray, active_species) -> np.ndarray:
    """Observable phi_i psi_i of a trajectory g[T, 12] (visualization/helpers.py:39-67)."""
    g = np.asarray(x0, dtype=np.float64)
    off = (g.shape[1] - 2) // 2 + 1
    return np.stack([g[:, sp] * g[:, off + sp] for sp in active_species], axis=1)


def compute_fit_metrics(t_arr, x0, active_species, data, idx_sparse, weights=None) -> Dict[str, object]:
    """RMSE / MAE between the model observable at the observation rows and the data (visualization/helpers.py:70-127)."""
    resid = compute_phibar(x0, active_species)[np.asarray(idx_sparse)] - np.asarray(data, dtype=np.float64)
    out = {
        "n_obs": int(resid.shape[0]), "n_species": int(resid.shape[1]),
        "rmse_per_species": np.sqrt(np.mean(resid ** 2, axis=0)), "mae_per_species": np.mean(np.abs(resid), axis=0),
        "rmse_total": float(np.sqrt(np.mean(resid ** 2))), "mae_total": float(np.mean(np.abs(resid))),
        "max_abs": float(np.max(np.abs(resid))),
    }
    if weights is not None:
        w = np.asarray(weights, dtype=np.float64)
        out["weighted_rmse_total"] = float(np.sqrt(np.sum(w * resid ** 2) / np.sum(w)))
        if resid.shape[1] > 4:
            out["rmse_pg_last2"] = float(np.sqrt(np.mean(resid[-2:, 4] ** 2)))
    return out
