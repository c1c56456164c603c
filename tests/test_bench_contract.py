"""bench.py contract checks that need no GPU: the reference arm runs on the host cores and prints ONE JSON line
with the keys the driver reads; the GPU arm refuses to run without a B200 instead of falling back."""
import json
import subprocess
import sys

from helpers import ROOT


import os

import pytest


@pytest.mark.parametrize("kind", ["reference", "port"])
def test_reference_arm_prints_contract_line(kind):
    """kind = "reference": the reference as shipped from oracle/_ref (build-time snapshot, oracle/make_ref.py); "port": the
    oracle's C restatement, the fallback when the snapshot is absent."""
    sys.path.insert(0, str(ROOT))
    from oracle import make_ref

    if kind == "reference" and not make_ref.build():
        pytest.skip("no reference tree and no oracle/_ref snapshot on this machine")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, HMX_BENCH_REF_KIND=kind))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["value"] > 0 and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_non_zero_ranks_exit_silently():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_reference_ode_only_baseline_equals_the_oracle():
    """SURVEY 8(d)(ii), the reference without its Python sensitivity loop (its numba solve + its log_likelihood_sparse with
    sig = 0, oracle/ref_runner.py): one more pin of the oracle, and the number bench.py reports as `reference_ode_only`.
    Runs in a subprocess: importing the snapshot puts its `core` package into sys.modules, which other tests import from
    the live reference tree."""
    sys.path.insert(0, str(ROOT))
    from oracle import make_ref

    if not make_ref.build():
        pytest.skip("no reference tree and no oracle/_ref snapshot on this machine")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import oracle as orc, ref_runner\n"
        "from tmcmc202601_b200 import workloads\n"
        "wl = workloads.four_condition_workload(2, seed=11, tsm_variance=False)\n"
        "arm = ref_runner.ReferenceArm(wl.config, n_jobs=2)\n"
        "ll = arm.evaluate_ode_only(wl.theta, wl.cond_id)\n"
        "ref = orc.evaluate_hmx(wl.config, wl.theta, wl.cond_id)\n"
        "print('MAXREL', float(np.max(np.abs(ll - ref) / np.abs(ref))))\n" % str(ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float([l for l in out.stdout.splitlines() if l.startswith("MAXREL")][0].split()[1])
    assert worst < 1e-12
