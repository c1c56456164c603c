"""Golden trajectories for solver variants the main fixture set does not cover: `active_species` masks (inactive
species start at phi = psi = 0, S5:563-575) and a vector phi_init, from the LIVE reference solver.
Run in the build container only:  python tests/golden/make_golden_masked.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_live as rl  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    m = rl.bootstrap()
    S = m["BiofilmNewtonSolver5S"]
    rng = np.random.default_rng(31)
    out = {}
    cases = [
        ("four_species", dict(active_species=[0, 1, 2, 3], phi_init=0.05)),
        ("three_species_vec", dict(active_species=[0, 2, 4], phi_init=[0.1, 0.2, 0.05, 0.02, 0.03])),
        ("all_vec", dict(active_species=[0, 1, 2, 3, 4], phi_init=[0.12, 0.08, 0.05, 0.02, 0.01])),
    ]
    for name, kw in cases:
        s = S(dt=1e-4, maxtimestep=400, c_const=25.0, alpha_const=0.0, Kp1=1e-4, K_hill=0.05, n_hill=4.0, **kw)
        th = rng.uniform(0.0, 3.0, 20)
        t, g = s.run_deterministic(th)
        out[f"{name}_theta"], out[f"{name}_g"] = th, np.asarray(g)
        out[f"{name}_active"] = np.array(kw["active_species"])
        out[f"{name}_phi_init"] = np.broadcast_to(np.asarray(kw["phi_init"], dtype=float), 5).copy()
        print(name, np.asarray(g).shape, float(np.abs(g).max()), g[0][:6])
    np.savez_compressed(OUT / "solver_masked_kat.npz", **out)


if __name__ == "__main__":
    main()
