"""Per-trip event traces of the ODE kernel's lane state machine (CPU only).

Records, through the host compile of the product headers (tests/csrc/host_emu.cpp: emu_trip_trace), what every trip of an
evaluation did (direction solved? time step ended?) for prior draws of the bench workload and prints the statistics the warp
cost estimates of DESIGN.md section 8 start from: directions and step ends per trip, trips per time step.  The launch-level
replay (arrival order / longest first / processor sharing) is scripts/schedule_replay.py.
Usage: python scripts/trip_sim.py [n_per_condition]
"""
import ctypes as C
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402
from tmcmc202601_b200 import workloads  # noqa: E402


def traces(n_per_condition=64, seed=20260101):
    emu = C.CDLL(str(g.build_host_emulation()))
    wl = workloads.four_condition_workload(n_per_condition, seed=seed)
    cap = 60000

    def one(i):
        ev = np.zeros(cap, dtype=np.uint8)
        n = C.c_longlong(0)
        th = np.ascontiguousarray(wl.theta[i])
        emu.emu_trip_trace(C.byref(wl.config), int(wl.cond_id[i]), th.ctypes.data_as(C.c_void_p), ev.ctypes.data_as(C.c_void_p),
                           C.c_longlong(cap), C.byref(n))
        return ev[: min(n.value, cap)].copy()

    with ThreadPoolExecutor(16) as ex:
        return list(ex.map(one, range(wl.theta.shape[0]))), wl.cond_id


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    tr, cid = traces(n)
    np.savez_compressed("/tmp/sim/traces.npz", cid=cid, **{f"t{i}": t for i, t in enumerate(tr)})
    lens = np.array([len(t) for t in tr])
    print("trips per evaluation: mean %.0f min %d max %d" % (lens.mean(), lens.min(), lens.max()))
    allv = np.concatenate(tr)
    print("direction per trip %.3f, step end per trip %.3f" % ((allv & 1).mean(), ((allv >> 1) & 1).mean()))
    # trips per step histogram
    h = np.zeros(64, dtype=np.int64)
    for t in tr:
        ends = np.flatnonzero(t & 2)
        d = np.diff(np.concatenate([[-1], ends]))
        h += np.bincount(np.minimum(d, 63), minlength=64)
    print("trips per step histogram:", {k: int(v) for k, v in enumerate(h) if v})
