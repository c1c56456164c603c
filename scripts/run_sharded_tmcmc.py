"""One TMCMC run with particles sharded over the GPUs of one box (one process per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        scripts/run_sharded_tmcmc.py [--particles 2000] [--condition Dysbiotic_HOBIC]

Every rank runs the same host program with the same seed; each likelihood batch is split into contiguous blocks,
evaluated on the local GPU and all-gathered; the covariance comes from shard-local partial sums + all-reduce.
Rank 0 checks that all replicas ended bit-identical and (with --check) equal to the single-GPU run.
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", type=int, default=2000)
    ap.add_argument("--condition", default="Dysbiotic_HOBIC")
    ap.add_argument("--mutation-steps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--n-hill", type=float, default=4.0)
    ap.add_argument("--device-mutation", action="store_true", help="proposals / accept-reject on the GPU (Philox stream)")
    ap.add_argument("--tag", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from tmcmc202601_b200 import distributed as D
    from tmcmc202601_b200 import tmcmc, workloads
    from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    cd = workloads.load_conditions()[args.condition]
    sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"], n_hill=args.n_hill)
    base = np.full(20, 1.5)
    base[cd["locked"]] = 0.0
    bounds = [tuple(b) for b in cd["bounds"]]

    def run(sharded: bool):
        ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], cd["active_indices"], base, cd["data"], cd["idx_sparse"], cd["sigma_obs"],
                                    cov_rel=0.005, theta_linearization=base, device=local)
        ops = tmcmc.GpuStageOps(engine=ev._engine)
        like, sops = (D.ShardedEvaluator(ev), D.ShardedStageOps(ops)) if sharded else (ev, ops)
        if sharded:  # communicator set-up (channels, peer connections) belongs to process start-up, not to the run
            D.all_gather_blocks(np.zeros(D.shard_bounds(world * 4, world, rank)[1] - D.shard_bounds(world * 4, world, rank)[0]), world * 4)
            D.all_reduce_sum(np.zeros(8))
            dist.barrier()
        t0 = time.perf_counter()
        res = tmcmc.run_TMCMC(like, bounds, n_particles=args.particles, n_stages=30, seed=4242, min_delta_beta=0.02, max_delta_beta=0.5,
                              evaluator=like, theta_base_full=base, active_indices=cd["active_indices"],
                              n_mutation_steps=args.mutation_steps, stage_ops=sops, device_mutation=args.device_mutation)
        wall = time.perf_counter() - t0
        n_local = ev.health.n_calls
        ev.close()
        return res, wall, n_local

    res, wall, n_local = run(world > 1)
    out = {"world": world, "particles": args.particles, "condition": args.condition, "n_hill": args.n_hill, "device_mutation": bool(args.device_mutation), "wall_s": wall,
           "likelihood_wall_s": res.timing_breakdown_s["tsm_s"] if res.timing_breakdown_s else None,
           "stages": len(res.stage_summary), "beta": res.beta_schedule[-1], "local_evaluations": int(n_local),
           "logL_max": float(res.logL_values.max()), "log_evidence": res.log_evidence,
           "acc": [round(a, 3) for a in res.acc_rate_history]}
    if world > 1:
        # replicas must be bit-identical
        mine = torch.from_numpy(res.samples).cuda()
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([int(torch.equal(mine, ref))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out["replicas_identical"] = bool(same.item())
        counts = torch.tensor([n_local], device="cuda", dtype=torch.int64)
        allc = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(allc, counts)
        out["evaluations_per_rank"] = [int(c.item()) for c in allc]
    if args.check and rank == 0:
        solo, wall1, _ = run(False)
        out["single_gpu_wall_s"] = wall1
        out["equals_single_gpu_beta_schedule"] = bool(np.allclose(solo.beta_schedule, res.beta_schedule, rtol=1e-12))
        out["max_abs_sample_diff_vs_single_gpu"] = float(np.abs(solo.samples - res.samples).max())
    if rank == 0:
        print(json.dumps(out), flush=True)
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / f"sharded_tmcmc_w{world}{args.tag}.json").write_text(json.dumps(out, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
