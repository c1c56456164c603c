"""Do the ODE kernels of two contexts (two streams) overlap on one GPU?  (development experiment)"""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch
from tmcmc202601_b200 import workloads
from tmcmc202601_b200.engine import HamiltonEngine

wl = workloads.four_condition_workload(1250, seed=3)   # 5000 evaluations
dev = torch.device("cuda", 0)
engs = [HamiltonEngine(wl.config, device=0) for _ in range(4)]
th = torch.from_numpy(wl.theta).to(dev); cid = torch.from_numpy(wl.cond_id).to(dev)
outs = [torch.empty(th.shape[0], dtype=torch.float64, device=dev) for _ in engs]
torch.cuda.synchronize()

def run_dev(k):
    t0 = time.perf_counter()
    for e, o in zip(engs[:k], outs[:k]):
        e.loglik_batch_device(th, cid, o)
    for e in engs[:k]:
        e.synchronize()
    return time.perf_counter() - t0

for k in (1, 1, 2, 4, 1, 2, 4):
    print("device buffers, one host thread, %d contexts: %.1f ms" % (k, run_dev(k) * 1e3), flush=True)

def run_threads(k, pinned):
    res = [None] * k
    hosts = [torch.empty(th.shape[0], dtype=torch.float64).pin_memory().numpy() if pinned else np.empty(th.shape[0]) for _ in range(k)]
    thh = torch.from_numpy(wl.theta).pin_memory().numpy() if pinned else wl.theta
    def work(i):
        import ctypes as C
        e = engs[i]
        t0 = time.perf_counter()
        e._check(e._lib.hmx_loglik_batch(e._ctx, thh.ctypes.data, wl.cond_id.ctypes.data, th.shape[0], hosts[i].ctypes.data, None, None, None), "x")
        res[i] = time.perf_counter() - t0
    ts = [threading.Thread(target=work, args=(i,)) for i in range(k)]
    t0 = time.perf_counter()
    [t.start() for t in ts]; [t.join() for t in ts]
    return time.perf_counter() - t0, res

for pinned in (False, True):
    for k in (1, 2, 4):
        w, r = run_threads(k, pinned)
        print("host buffers (%s), %d threads: wall %.1f ms, per call %s" % ("pinned" if pinned else "pageable", k, w * 1e3, ["%.0f" % (x * 1e3) for x in r]), flush=True)
