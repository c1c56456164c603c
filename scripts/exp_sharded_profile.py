"""Where does a sharded stage spend its time?  (development experiment; torchrun, see run_sharded_tmcmc.py)"""
import json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
from tmcmc202601_b200 import distributed as D, tmcmc, workloads
from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
N = int(sys.argv[1]) if len(sys.argv) > 1 else 25000
cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"], n_hill=4.0)
base = np.full(20, 1.5); base[cd["locked"]] = 0.0
ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], cd["active_indices"], base, cd["data"], cd["idx_sparse"], cd["sigma_obs"], cov_rel=0.005, theta_linearization=base, device=local)
eng = ev._engine
eng.set_timing(True)
log = []
orig_mutate, orig_batch = eng.mutate, eng.loglik_batch
def mutate(*a, **k):
    t0 = time.perf_counter(); out = orig_mutate(*a, **k); dt = time.perf_counter() - t0
    c = eng.counters(); log.append(("mutate", a[1].shape[0] * a[0].K, dt * 1e3, c["last_ode_kernel_ms"])); return out
def batch(*a, **k):
    t0 = time.perf_counter(); out = orig_batch(*a, **k); dt = time.perf_counter() - t0
    c = eng.counters(); log.append(("loglik", a[0].shape[0], dt * 1e3, c["last_ode_kernel_ms"])); return out
eng.mutate, eng.loglik_batch = mutate, batch
like, sops = D.ShardedEvaluator(ev), D.ShardedStageOps(tmcmc.GpuStageOps(engine=eng))
D.all_gather_blocks(np.zeros(D.shard_bounds(world * 4, world, rank)[1] - D.shard_bounds(world * 4, world, rank)[0]), world * 4)
D.all_reduce_sum(np.zeros(8)); dist.barrier()
t0 = time.perf_counter()
res = tmcmc.run_TMCMC(like, [tuple(b) for b in cd["bounds"]], n_particles=N, n_stages=30, seed=4242, min_delta_beta=0.02, max_delta_beta=0.5,
                      evaluator=like, theta_base_full=base, active_indices=cd["active_indices"], n_mutation_steps=5, stage_ops=sops, device_mutation=True)
wall = time.perf_counter() - t0
if rank == 0:
    print(json.dumps({"wall_s": wall, "tsm_s": res.timing_breakdown_s, "stages": len(res.stage_summary),
                      "launches": [(k, n, round(w, 1), round(o, 1)) for k, n, w, o in log]}))
dist.barrier(); dist.destroy_process_group()
