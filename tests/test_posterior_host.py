"""Posterior-summary helpers against the reference's own outputs for its stored production run
(_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30/parameter_summary.csv, written by MAIN:378-456)."""
import numpy as np
import pytest

from helpers import GOLDEN
from tmcmc202601_b200.posterior import compute_credible_intervals, compute_hdi
from tmcmc202601_b200.tmcmc import compute_MAP_with_uncertainty


def test_credible_intervals_match_reference_run_directory():
    k = np.load(GOLDEN / "posterior_dh_1k30_samples.npz")
    ci = compute_credible_intervals(k["samples"], float(k["credibility"]))
    for name in ("hdi_lower", "hdi_upper", "et_lower", "et_upper", "median"):
        assert np.allclose(ci[name], k[name], rtol=1e-12, atol=1e-14), name
    summ = compute_MAP_with_uncertainty(k["samples"], np.arange(len(k["samples"]), dtype=float))
    assert np.allclose(summ["mean"], k["mean"], rtol=1e-12)
    assert np.array_equal(summ["MAP"], k["samples"][-1])


def test_hdi_edge_cases():
    x = np.arange(10.0)
    lo, hi = compute_hdi(x, 0.5)  # window of ceil(5)+1 = 6 samples: ties -> first
    assert lo[0] == 0.0 and hi[0] == 5.0
    lo, hi = compute_hdi(x, 1.0)
    assert lo[0] == 0.0 and hi[0] == 9.0
    s = np.random.default_rng(0).exponential(size=(5000, 2))
    lo, hi = compute_hdi(s, 0.9)
    assert np.all(lo < 0.05) and np.all(np.abs(hi - 2.3) < 0.25)  # skewed: HDI hugs zero, unlike equal tails


def test_build_replicate_sigma_matches_reference_loader():
    from pathlib import Path

    csv = Path("/root/reference/data_5species/experiment_data/fig3_species_distribution_replicates.csv")
    if not csv.exists():
        pytest.skip("reference data not present (GPU box)")
    from tmcmc202601_b200 import workloads
    from tmcmc202601_b200.evaluator import build_replicate_sigma

    species_map = {"S. oralis": 0, "A. naeslundii": 1, "V. dispar": 2, "V. parvula": 2, "F. nucleatum": 3,
                   "P. gingivalis_20709": 4, "P. gingivalis_W83": 4}
    for key in workloads.CONDITION_KEYS:
        cd = workloads.load_conditions()[key]
        sig = build_replicate_sigma(str(csv), cd["condition"], cd["cultivation"], cd["days"], species_map)
        assert np.array_equal(sig, cd["sigma_obs"])  # heine_conditions.json holds the reference function's output
