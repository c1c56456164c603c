"""Replay of measured evaluation costs through schedulers (CPU only; the numbers behind DESIGN.md section 6).

Costs (state-machine trips per evaluation) come from the product's own lane state machine compiled for the CPU
(tests/csrc/host_emu.cpp: emu_evaluate) on prior draws of one condition.  Schedulers: the kernel's (lanes fetch evaluations in
arrival order), longest-first with exact and with noisy cost predictions, and processor sharing (what time slicing approaches).

    python scripts/schedule_replay.py [condition=Dysbiotic_HOBIC] [n_draws=1500]
"""
import ctypes as C
import heapq
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tmcmc202601_b200 import workloads  # noqa: E402

LANES = 148 * 2 * 128      # resident lanes of the ODE kernel on one B200
US_PER_TRIP = 2.72         # measured: one trip of a lane with two warps per scheduler


def costs(condition, n, seed=20260101):
    emu = C.CDLL(str(g.build_host_emulation()))
    wl = workloads.single_condition_workload(condition, n, seed=seed)

    def one(i):
        ll, st, it, tr = C.c_double(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        th = np.ascontiguousarray(wl.theta[i])
        emu.emu_evaluate(C.byref(wl.config), 0, th.ctypes.data_as(C.c_void_p), None, C.byref(ll), C.byref(st), C.byref(it), C.byref(tr))
        return tr.value

    with ThreadPoolExecutor(16) as ex:
        return np.array(list(ex.map(one, range(n))), dtype=float)


def list_schedule(c):
    if len(c) <= LANES:
        return c.max()
    h = list(c[:LANES])
    heapq.heapify(h)
    for x in c[LANES:]:
        heapq.heappush(h, heapq.heappop(h) + x)
    return max(h)


def processor_sharing(c):
    c = np.sort(c)
    n, t, served = len(c), 0.0, 0.0
    for k in range(n):
        rate = min(1.0, LANES / (n - k))
        t += (c[k] - served) / rate
        served = c[k]
    return t


if __name__ == "__main__":
    cond = sys.argv[1] if len(sys.argv) > 1 else "Dysbiotic_HOBIC"
    pool = costs(cond, int(sys.argv[2]) if len(sys.argv) > 2 else 1500)
    print(f"{cond}: trips per evaluation mean {pool.mean():.0f}, min {pool.min():.0f}, max {pool.max():.0f}")
    rng = np.random.default_rng(0)
    ms = US_PER_TRIP * 1e-3
    for n in (40000, 62500, 87500, 137500, 187500, 500000):
        c = rng.choice(pool, size=n)
        row = {"evaluations": n, "work_ms": c.sum() / LANES * ms, "arrival_order_ms": list_schedule(c) * ms,
               "longest_first_ms": list_schedule(np.sort(c)[::-1]) * ms, "processor_sharing_ms": processor_sharing(c) * ms}
        for rho in (0.9, 0.7):
            pred = rho * (c - c.mean()) + np.sqrt(1 - rho**2) * rng.standard_normal(n) * c.std()
            row[f"longest_first_rho{rho}_ms"] = list_schedule(c[np.argsort(-pred)]) * ms
        print({k: (round(v) if isinstance(v, float) else v) for k, v in row.items()})
