"""Golden data for the run-directory writer, taken from the reference's stored production run
/root/reference/_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30 (written by data_5species/main/estimate_reduced_nishioka.py:2904-3430).
Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_rundir.py
"""
import csv
import json
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent
RUN = Path("/root/reference/_runs/Dysbiotic_HOBIC_K0.05_n4.0_1k30")


def keys_of(obj, depth=0):
    if isinstance(obj, dict) and depth < 2:
        return {k: keys_of(v, depth + 1) for k, v in obj.items()}
    return type(obj).__name__


def main():
    with open(RUN / "parameter_summary.csv") as fh:
        rows = list(csv.DictReader(fh))
    diag = json.loads((RUN / "mcmc_diagnostics.json").read_text())
    tmap = json.loads((RUN / "theta_MAP.json").read_text())
    tmean = json.loads((RUN / "theta_mean.json").read_text())
    fit = json.loads((RUN / "fit_metrics.json").read_text())
    np.savez_compressed(
        OUT / "rundir_dh_1k30.npz",
        logL=np.load(RUN / "logL.npy").astype(np.float64),
        rhat=np.array(diag["rhat"]), ess=np.array(diag["ess"]),
        MAP=np.array([float(r["MAP"]) for r in rows]), mean=np.array([float(r["mean"]) for r in rows]),
        names=np.array([r["name"] for r in rows]), index=np.array([int(r["index"]) for r in rows]),
        theta_MAP_full=np.array(tmap["theta_full"]), theta_mean_full=np.array(tmean["theta_full"]),
        csv_header=np.array(list(rows[0].keys())),
        **{f"fit_{k}_{m}": np.asarray(fit[k][m], dtype=np.float64) for k in ("MAP", "Mean")
           for m in ("rmse_per_species", "mae_per_species", "rmse_total", "mae_total", "max_abs")},
    )
    schema = {}
    for f in sorted(RUN.glob("*.json")):
        schema[f.name] = keys_of(json.loads(f.read_text()))
    schema["_files"] = sorted(p.name for p in RUN.iterdir())
    schema["_npy_shapes"] = {f.name: list(np.load(f).shape) for f in sorted(RUN.glob("*.npy"))}
    (OUT / "rundir_schema.json").write_text(json.dumps(schema, indent=1))
    print(sorted(schema))


if __name__ == "__main__":
    main()
