"""The C-ABI shared library: loads on a CPU-only box, exports every symbol include/hmx.h declares, and
refuses to run without a B200 (no CPU fallback).  No compute calls here."""
import ctypes as C
import re

import pytest

from helpers import ROOT
from tmcmc202601_b200 import _abi, workloads


def _declared_symbols():
    text = (ROOT / "include" / "hmx.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hmx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    lib = _abi.load()
    declared = _declared_symbols()
    assert len(declared) >= 17
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(_abi.EXPORTS) == declared  # the ctypes binding covers exactly the header
    assert lib.hmx_abi_version() == _abi.ABI_VERSION


def test_library_is_sm100a_only(built):
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-lelf", str(_abi.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_create_fails_loudly_without_gpu(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load()
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    cfg = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    ctx = C.c_void_p()
    rc = lib.hmx_create(C.byref(ctx), C.byref(cfg), 0)
    assert rc == -3 and not ctx  # HMX_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.hmx_last_error(None)
    from tmcmc202601_b200.engine import HamiltonEngine, HmxError

    with pytest.raises(HmxError):
        HamiltonEngine(cfg)


def test_create_rejects_invalid_config(built):
    lib = _abi.load()
    cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
    cfg = _abi.make_config([workloads.cond_struct(cd)], **workloads.PRODUCTION_SOLVER)
    cfg.n_cond = 99
    ctx = C.c_void_p()
    assert lib.hmx_create(C.byref(ctx), C.byref(cfg), 0) == -1  # HMX_ERR_INVALID comes before the device probe
    assert b"n_cond" in lib.hmx_last_error(None)
    assert lib.hmx_create(C.byref(ctx), None, 0) == -1


def test_null_context_calls_return_invalid(built):
    lib = _abi.load()
    assert lib.hmx_synchronize(None) == -1
    assert lib.hmx_loglik_batch(None, None, None, 0, None, None, None, None) == -1
    assert lib.hmx_resample(None, None, 0, 0.5, None) == -1
    lib.hmx_destroy(None)  # no-op


def test_make_cond_broadcasting_and_checks():
    import numpy as np

    data = np.full((5, 5), 0.2)
    c = _abi.make_cond(data, [1, 2, 3, 4, 5], 0.1, 0.02)
    assert list(c.phi_init) == [0.02] * 5 and c.n_obs == 5
    assert list(c.sigma_obs[:25]) == [0.1] * 25 and list(c.weights[:25]) == [1.0] * 25
    c = _abi.make_cond(data, [1, 2, 3, 4, 5], [0.1, 0.2, 0.3, 0.4, 0.5], [0.1, 0.2, 0.3, 0.2, 0.1], active_theta_indices=[0, 5, 19])
    assert list(c.sigma_obs[5:10]) == [0.1, 0.2, 0.3, 0.4, 0.5]
    assert c.active_theta_mask == (1 | 1 << 5 | 1 << 19)
    with pytest.raises(ValueError):
        _abi.make_cond(np.zeros((9, 5)), range(9), 0.1, 0.02)
    with pytest.raises(ValueError):
        _abi.make_cond(data, [1, 2, 3], 0.1, 0.02)
    with pytest.raises(ValueError):
        _abi.make_cond(np.zeros((5, 4)), range(5), 0.1, 0.02)
