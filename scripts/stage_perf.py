import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from tmcmc202601_b200.tmcmc import GpuStageOps
import torch
eng = GpuStageOps().engine
eng.use_torch_stream()
for N in (65536, 1<<20, 1<<21):
    w = torch.rand(N, dtype=torch.float64, device='cuda'); w /= w.sum()
    idx = torch.empty(N, dtype=torch.int64, device='cuda')
    ll = torch.randn(N, dtype=torch.float64, device='cuda')*20
    for name, fn in (("resample", lambda: eng.resample_device(w, 0.3, idx)), ("beta_step", lambda: eng.beta_step(ll, 0.0)), ("weights", lambda: eng.weights_device(ll, 0.1, w))):
        fn(); torch.cuda.synchronize(); t0=time.perf_counter()
        for _ in range(3): fn()
        torch.cuda.synchronize(); print(N, name, "%.3f ms"%((time.perf_counter()-t0)/3*1e3))
    w = torch.rand(N, dtype=torch.float64, device='cuda'); w /= w.sum()
