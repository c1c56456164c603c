"""Golden vectors for tmcmc202601_b200.predictive.generate_dataset: the LIVE reference generator
(/root/reference/deeponet/generate_training_data.py:144-333, numba solver) run here on three small cases, with its file
loaders replaced by in-memory arrays.  Run in the build container only:

    python tests/golden/make_golden_dataset.py
"""
import importlib.util
import sys
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

OUT = Path(__file__).resolve().parent
for m in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(m, MagicMock())
spec = importlib.util.spec_from_file_location("ref_gen", "/root/reference/deeponet/generate_training_data.py")
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def main():
    bounds = gen.load_prior_bounds("Dysbiotic_HOBIC")
    rng = np.random.default_rng(123)
    theta_map = rng.uniform(bounds[:, 0], bounds[:, 1])
    theta_map[18] = bounds[18, 1] + 0.7  # outside its bound: exercises the expansion rule
    post = rng.uniform(bounds[:, 0], bounds[:, 1], size=(7, 20))
    out = {"bounds_in": bounds, "theta_map": theta_map, "posterior": post}
    cases = {
        "uniform": dict(map_frac=0.0, posterior=None),
        "map": dict(map_frac=0.4, posterior=None),
        "mix": dict(map_frac=0.3, posterior=post),
    }
    for name, c in cases.items():
        gen.load_theta_map = lambda condition: theta_map.copy()
        gen.load_posterior_samples = (lambda p, max_samples=5000: c["posterior"])
        th, phi, t, b = gen.generate_dataset(n_samples=14, condition="Dysbiotic_HOBIC", seed=11, maxtimestep=60, dt=1e-5, n_time_out=13,
                                             map_frac=c["map_frac"], map_std_frac=0.1, expand_bounds=True,
                                             posterior_dir="/nonexistent" if c["posterior"] is not None else None, posterior_frac=0.5)
        out[f"{name}_theta"], out[f"{name}_phi"], out[f"{name}_t"], out[f"{name}_bounds"] = th, phi, t, b
        print(name, th.shape, phi.shape, float(np.abs(phi).max()))
    np.savez_compressed(OUT / "training_dataset_kat.npz", **out)


if __name__ == "__main__":
    main()
