#!/usr/bin/env python
"""Recipe for oracle/_ref: a runnable snapshot of the reference's OWN CPU implementation of the hot path.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/hamilton_oracle.h).  The reference is Python + numba and lives
at /root/reference, which exists in the build container but not on the GPU box.  `python oracle/make_ref.py`
(also run by __graft_entry__.build() when /root/reference is present) copies -- at BUILD time, into the
git-ignored directory oracle/_ref/, never into the repository's history -- exactly the module files the
reference's `core.evaluator.LogLikelihoodEvaluator` and `core.tmcmc.evaluate_particles_parallel` import
(21 files, 430 KB: data_5species/{core,utils,debug,visualization}, tmcmc/program2602/*.py).  gpurun ships
oracle/_ref/ to the GPU box like a built .so, where `bench.py --impl reference` and the `cpu_baseline` leg
time the reference as shipped (`kind: "reference"`), on the box's host cores.

The file list is not hard-coded: the reference is imported here and every module whose file lies under the
reference root is copied with its relative path.
"""

from __future__ import annotations

import json
import os
import shutil
import sys
from pathlib import Path
from unittest.mock import MagicMock

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("HMX_REFERENCE_ROOT", "/root/reference"))

STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors",
         "matplotlib.ticker", "matplotlib.lines", "matplotlib.cm", "seaborn")
# sys.path entries relative to the reference root, in the order of MAIN:63-74
PATHS = ("data_5species", "tmcmc/program2602", "tmcmc", "")


def import_reference(root: Path):
    """Import the evaluator + batch entry of a reference tree rooted at `root`; returns (LogLikelihoodEvaluator,
    evaluate_particles_parallel)."""
    for m in STUBS:
        sys.modules.setdefault(m, MagicMock())
    for p in PATHS:
        sys.path.insert(0, str(root / p) if p else str(root))
    import logging

    logging.disable(logging.CRITICAL)
    from core.evaluator import LogLikelihoodEvaluator  # noqa: E402
    from core.tmcmc import evaluate_particles_parallel  # noqa: E402

    return LogLikelihoodEvaluator, evaluate_particles_parallel


def available() -> bool:
    return (OUT / "MANIFEST.json").exists()


def build(force: bool = False) -> bool:
    """Returns True when oracle/_ref is in place (freshly copied or already there)."""
    if available() and not force:
        return True
    if not (REF / "tmcmc" / "program2602" / "improved_5species_jit.py").exists():
        return False  # GPU box / no reference tree: use what was shipped, or fall back to the port
    import_reference(REF)
    root = os.path.realpath(REF)
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(root + os.sep)})
    if OUT.exists():
        shutil.rmtree(OUT)
    copied = []
    for f in files:
        rel = os.path.relpath(f, root)
        dst = OUT / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(f, dst)
        copied.append(rel)
    (OUT / "MANIFEST.json").write_text(json.dumps({"source": str(REF), "files": copied}, indent=1))
    return True


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "reference tree not found", "-", len(list(OUT.rglob("*.py"))) if OUT.exists() else 0, "files")
