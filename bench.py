#!/usr/bin/env python
"""bench.py -- particle-ODE likelihood evaluations per second on N B200s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[4], the synthetic likelihood throughput sweep -- per GPU
4 conditions x 262,144 theta ~ U(prior bounds of the condition) (np.random.default_rng(20260101 + c + 97 rank)),
production solver settings (dt=1e-4, 2500 steps, c=25, alpha=0, Kp1=1e-4, K_hill=0.05, n_hill=4), TSM
variance on.  One evaluation = one (theta, condition) pair = one 2500-step implicit solve + its Gaussian
log-likelihood.  Weak scaling: the per-GPU batch is fixed, ranks hold disjoint particles.

A "step" is one pass of the hot path over one batch:
    ODE + likelihood kernels on this rank's particles -> all-gather of the log-likelihoods (N > 1) ->
    per-condition stage arithmetic on the gathered population (delta-beta bisection, weights + evidence,
    systematic resampling, weighted covariance from shard-local partial sums + all-reduce).
`value`  : inputs resident in HBM, timed with CUDA events on the launching stream, barrier + synchronize on
           both sides, max over ranks.
`e2e`    : the same step through the public host-facing API (theta/cond_id H2D from pinned NumPy buffers and
           logL/indices/scalars/covariance D2H inside the timed region; the gathered log-likelihoods and the weights stay
           on the device between the all-gather and the stage kernels).
`roofline`: FP64 FMA pipe (the path is compute bound; HBM traffic is ~1.2 KB per 2e7-flop evaluation).
`cpu_baseline` / `--impl reference`: the reference AS SHIPPED (kind "reference": oracle/_ref, a build-time snapshot of the
           reference's own evaluator modules written by oracle/make_ref.py, driven through its public API
           LogLikelihoodEvaluator + evaluate_particles_parallel on all host cores); when oracle/_ref is absent, the oracle's C
           restatement (kind "port") on all host threads.  The port's number is always reported next to it.
`sweep`, `t500`, `tmcmc_runs`: (N = 1, after the timed regions) the rest of BASELINE.json: the 10^4-10^7 per-condition sweep,
           the 500-step configuration, configs[0]/[1]/[2] wall times and a configs[3] Hill-sweep run.  --quick skips them.
`tmcmc_run`: (N = 1, after the timed regions) the other half of BASELINE.json's metric, "TMCMC wall time per run": configs[1]
           through the reference-shaped API, host-side and device-side mutation, posterior checked against the stored
           reference run (226 865 s there).  --no-tmcmc-run skips it.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_PER_COND = 262144
N_COND = 4
METRIC = "particle_ode_likelihood_evals_per_sec"
UNIT = "evals/s"

# Executed FP64 flops (DFMA = 2, DMUL = DADD = 1) of the ODE kernel per residual evaluation ("trip") and per quasi-Newton
# direction solve, calibrated on the round-2 kernel from the ncu capture profiles/r02d_ode_kernel.json: thread-level SASS
# counters 1.5171e12 DFMA + 1.1205e12 DMUL + 3.4125e11 DADD = 4.496e12 flop for the launch's own counters of 3.9592e9
# residual evaluations and 3.4175e9 directions; attributed per SASS instruction to its source function through the line
# table (residual + rebase + its reciprocals 1.6945e12 flop -> 428 per residual evaluation; direction + 2x2 blocks + 5x5
# solve + its reciprocals 2.8015e12 -> 820 per direction).  DESIGN.md "Roofline accounting" has the derivation.
FLOPS_PER_RESIDUAL = 428.0
FLOPS_PER_DIRECTION = 820.0
# DRAM bytes per evaluation of the same capture: (dram__bytes_read.sum 59.1 MB + dram__bytes_write.sum 213.6 MB) / 262144
DRAM_BYTES_PER_EVAL_NCU = 1040
# the reference-equivalent count of SURVEY.md section 8(d): dense Q + K + dgesv per iteration, Q per extra trial
FP64_NOMINAL_TFLOPS = 37.2  # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
REF_FLOPS_PER_ITER = 2778.0
REF_FLOPS_PER_TRIAL = 384.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(rank: int):
    from tmcmc202601_b200 import workloads

    wl = workloads.four_condition_workload(N_PER_COND, seed=20260101 + 97 * rank)
    return wl


def reference_as_shipped(wl, n_cond_sample: int = 1) -> dict | None:
    """The reference's own evaluator (oracle/_ref) on a bounded sample: `cores` evaluations of the LAST condition(s) of the
    workload through core.tmcmc.evaluate_particles_parallel, after one warm-up call (numba JIT) in this process."""
    from oracle import ref_runner

    if not ref_runner.available():
        return None
    cores = os.cpu_count() or 1
    arm = ref_runner.ReferenceArm(wl.config, n_jobs=cores)
    conds = list(range(N_COND))[-n_cond_sample:]
    sel = np.concatenate([np.nonzero(wl.cond_id == c)[0][:cores] for c in conds])
    warm = np.concatenate([np.nonzero(wl.cond_id == c)[0][cores:cores + 1] for c in conds])
    t_warm = arm.warm_up(wl.theta[warm], wl.cond_id[warm])
    t0 = time.perf_counter()
    arm.evaluate(wl.theta[sel], wl.cond_id[sel])
    dt = time.perf_counter() - t0
    out = {"value": sel.size / dt, "unit": UNIT, "cores": cores, "kind": "reference",
           "sample": f"{sel.size} evaluations (condition(s) {conds}) through the reference's LogLikelihoodEvaluator + "
                     f"evaluate_particles_parallel(n_jobs={cores}) from oracle/_ref, {dt:.1f} s after a {t_warm:.1f} s warm-up (numba JIT)"}
    # the reference without its Python sensitivity loop -- its numba solve + its log_likelihood_sparse with
    # sig = 0, particles over forked workers (not the same quantity as `value`: the TSM variance is off)
    try:
        per = 16 * cores // N_COND
        sel2 = np.concatenate([np.nonzero(wl.cond_id == c)[0][:per] for c in range(N_COND)])
        arm.warm_up_ode_only(wl.theta[sel2], wl.cond_id[sel2])
        t0 = time.perf_counter()
        arm.evaluate_ode_only(wl.theta[sel2], wl.cond_id[sel2])
        dt2 = time.perf_counter() - t0
        out["reference_ode_only"] = {"value": sel2.size / dt2, "unit": UNIT, "cores": cores,
                                     "sample": f"{sel2.size} evaluations ({per} per condition), BiofilmNewtonSolver5S.run_deterministic + "
                                               f"log_likelihood_sparse(sig=0) from oracle/_ref on {cores} forked workers, {dt2:.1f} s"}
    except Exception as e:  # noqa: BLE001
        out["reference_ode_only"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def cpu_baseline(wl, budget_s: float = 12.0, max_evals: int = 16384) -> dict:
    """The oracle port on all host threads, on a bounded prefix-sample of the same workload (equal share per condition)."""
    from oracle import oracle as orc

    cores = os.cpu_count() or 1
    per = max(cores, 8)
    idx_by_cond = [np.nonzero(wl.cond_id == c)[0] for c in range(N_COND)]
    done, t_total, pos = 0, 0.0, 0
    while t_total < budget_s and done < max_evals:
        sel = np.concatenate([ix[pos:pos + per] for ix in idx_by_cond])
        if sel.size == 0:
            break
        t0 = time.perf_counter()
        orc.evaluate_hmx(wl.config, wl.theta[sel], wl.cond_id[sel], n_threads=cores)
        t_total += time.perf_counter() - t0
        done += sel.size
        pos += per
    return {"value": done / t_total if t_total > 0 else 0.0, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {done} evaluations of the workload ({done // N_COND} per condition), oracle/hamilton_oracle.c on {cores} pthreads, {t_total:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  oracle/_ref (the reference as
    shipped, kind "reference") when the build-time snapshot is present, else the oracle's C restatement (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = synth_batch(0)
    from oracle import oracle as orc
    from oracle import ref_runner

    cores = os.cpu_count() or 1
    n_total = args.steps + args.warmup
    if ref_runner.available() and os.environ.get("HMX_BENCH_REF_KIND", "reference") != "port":
        # one step = `cores` evaluations of ONE condition (step s -> condition s mod 4) through evaluate_particles_parallel:
        # ~5.4 s of CPU per evaluation per core, so a step costs ~6 s and all cores are busy
        kind, per_step = "reference", cores
        arm = ref_runner.ReferenceArm(wl.config, n_jobs=cores)
        by_cond = [np.nonzero(wl.cond_id == c)[0] for c in range(N_COND)]
        warm = np.array([ix[-1] for ix in by_cond])
        arm.warm_up(wl.theta[warm], wl.cond_id[warm])  # numba JIT + first call per condition, untimed

        def step_fn(s):
            c = s % N_COND
            sel = by_cond[c][(s // N_COND) * per_step:(s // N_COND + 1) * per_step]
            arm.evaluate(wl.theta[sel], wl.cond_id[sel])

        sample = (f"{per_step} evaluations per step (one condition per step, cycling) x {args.steps} steps through the reference's "
                  f"LogLikelihoodEvaluator + evaluate_particles_parallel(n_jobs={cores}) from oracle/_ref")
    else:
        kind, per_step = "port", max(cores * 64, 256)  # ~2-3 s of CPU work per step
        sel_all = np.concatenate([np.nonzero(wl.cond_id == c)[0][: n_total * per_step // N_COND + per_step] for c in range(N_COND)])
        np.random.default_rng(0).shuffle(sel_all)

        def step_fn(s):
            sel = sel_all[s * per_step:(s + 1) * per_step]
            orc.evaluate_hmx(wl.config, wl.theta[sel], wl.cond_id[sel], n_threads=cores)

        sample = f"{per_step} evaluations per step x {args.steps} steps, oracle/hamilton_oracle.c on {cores} pthreads"
    times = []
    for s in range(n_total):
        t0 = time.perf_counter()
        step_fn(s)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, per_step_note=f"each step = a bounded sample of {per_step} evaluations of the same workload"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int, per_step_note: str = "") -> dict:
    return {
        "workload": "BASELINE.json configs[4]: synthetic likelihood throughput, 4 conditions (Commensal/Dysbiotic x Static/HOBIC) x "
                    f"{N_PER_COND} theta ~ U(prior bounds) per GPU, Hamilton 5-species ODE dt=1e-4 x 2500 steps, K_hill=0.05 n_hill=4, TSM variance on",
        "evals_per_gpu_per_step": N_PER_COND * N_COND, "global_evals_per_step": N_PER_COND * N_COND * n_gpus,
        "n_steps_ode": 2500, "parallelism": f"particles sharded over {n_gpus} GPU(s), logL all-gather + covariance all-reduce per step",
        "l2_policy": "inputs+scratch per step (168 MB theta + 1 GB observation pairs per GPU) exceed the 126 MB L2; no flush needed",
        "note": per_step_note,
    }


def tmcmc_run():
    """Second half of BASELINE.json's metric ("TMCMC wall time per run"), outside the timed region: configs[1], the stored
    production run (Dysbiotic HOBIC, K_hill=0.05, n_hill=4, N=1000 particles x 2 chains, seed 42), through the
    reference-shaped API -- once with the host-side mutation loop (NumPy stream of the reference) and once with the
    device-side mutation (hmx_mutate) -- with the posterior checked against the reference run's own samples."""
    sys.path.insert(0, str(ROOT / "scripts"))
    import run_production_tmcmc as rp

    out = {"config": "BASELINE.json configs[1]: Dysbiotic HOBIC production run, N=1000 x 2 chains, K=0.05 n=4, seed 42",
           "reference_wall_s": 226864.8}
    for key, dev, lat in (("host_mutation", False, False), ("device_mutation", True, False), ("device_mutation_latency_kernel", True, True)):
        try:
            # two runs, the faster one reported: the first pays one-off costs (lazy kernel loading, buffer growth)
            walls = []
            for _ in range(2):
                res, samples, ll = rp.run_condition("Dysbiotic_HOBIC_legacy_1k30", 1000, 2, device_mutation=dev, latency_kernel=lat)
                walls.append(res["wall_s"])
            cmp_ = rp.compare_with_reference(samples, ll)
            out[key] = {"wall_s": min(walls), "wall_s_first_run": walls[0], "likelihood_evaluations": res["likelihood_evaluations"], "stages": res["stages"],
                        "converged": res["converged"], "posterior_z_mean_max": cmp_["z_mean_max"],
                        "posterior_std_ratio": [cmp_["std_ratio_min"], cmp_["std_ratio_max"]], "logL_max": res["logL_max"]}
        except Exception as e:  # the throughput line must not be lost to a failure of the side measurement
            out[key] = {"error": f"{type(e).__name__}: {e}"}
    # the two chains advanced concurrently with merged launches (batch.run_batch_TMCMC: bit-identical to running them one
    # after the other, half the number of launches -- each launch costs the latency of its slowest solve)
    try:
        walls = []
        for _ in range(2):
            summ, res = rp.run_merged(["Dysbiotic_HOBIC_legacy_1k30"], 1000, 2, device_mutation=True)
            walls.append(summ["wall_s"])
        samples = np.concatenate(res[0]["chains"], axis=0)
        cmp_ = rp.compare_with_reference(samples, np.concatenate(res[0]["logL"]))
        out["device_mutation_merged_chains"] = {"wall_s": min(walls), "wall_s_first_run": walls[0], "merged_launches": summ["merged_launches"],
                                                "likelihood_evaluations": summ["likelihood_evaluations"], "posterior_z_mean_max": cmp_["z_mean_max"],
                                                "posterior_std_ratio": [cmp_["std_ratio_min"], cmp_["std_ratio_max"]]}
    except Exception as e:  # noqa: BLE001
        out["device_mutation_merged_chains"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def timed_pass(eng, th_d, cid_d, ll_d, reps: int, local_rank: int) -> dict:
    """`reps` passes of the hot path (ODE + likelihood kernels) over a device-resident batch, CUDA events, clocks sampled."""
    import torch

    eng.loglik_batch_device(th_d, cid_d, ll_d)  # warm-up (buffer growth, cost-ordered work list)
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.loglik_batch_device(th_d, cid_d, ll_d)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = th_d.shape[0]
    c = eng.counters()
    return {"evals": int(n), "reps": reps, "ms_per_pass": ms, "evals_per_s": n / (ms * 1e-3), "newton_iters_per_eval": c["newton_iters"] / n,
            "nonfinite": int((ll_d == -1e20).sum().item()), "clocks": sampler.stop()}


def device_prior_batch(keys, n_per_cond: int, dev, seed: int):
    """theta ~ U(prior bounds of condition c) drawn ON the device (torch Philox): 10^7 x 4 x 20 doubles are 6.4 GB, and host RNG
    output is not the thing measured.  Locked parameters (bounds (0,0)) stay 0, as in workloads.four_condition_workload."""
    import torch

    from tmcmc202601_b200 import workloads

    conds = workloads.load_conditions()
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    th = torch.empty((len(keys) * n_per_cond, 20), dtype=torch.float64, device=dev)
    cid = torch.empty(len(keys) * n_per_cond, dtype=torch.int32, device=dev)
    for c, k in enumerate(keys):
        b = torch.tensor(conds[k]["bounds"], dtype=torch.float64, device=dev)
        blk = th[c * n_per_cond:(c + 1) * n_per_cond]
        blk.uniform_(0.0, 1.0, generator=g)
        blk.mul_(b[:, 1] - b[:, 0]).add_(b[:, 0])
        cid[c * n_per_cond:(c + 1) * n_per_cond] = c
    return th, cid


def throughput_sweep(eng, dev, local_rank: int, sizes=(10**4, 10**5, 10**6, 10**7)) -> list:
    """BASELINE.json configs[4] in full: N in {1e4, 1e5, 1e6, 1e7} theta per condition x 4 conditions on one GPU (the largest
    size runs in ten 4 Mi-evaluation chunks inside hmx_loglik_batch)."""
    import torch

    from tmcmc202601_b200 import workloads

    out = []
    for n in sizes:
        try:
            th, cid = device_prior_batch(workloads.CONDITION_KEYS, n, dev, seed=20260101 + n % 9973)
            ll = torch.empty(th.shape[0], dtype=torch.float64, device=dev)
            r = timed_pass(eng, th, cid, ll, reps=3 if n <= 10**5 else 1, local_rank=local_rank)
            r["n_per_condition"] = n
            out.append(r)
            del th, cid, ll
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001 - a side measurement must not lose the headline line
            out.append({"n_per_condition": n, "error": f"{type(e).__name__}: {e}"})
    return out


def t500_pass(dev, local_rank: int) -> dict:
    """The 500-step configuration of deeponet/surrogate_tmcmc.py:353-365 (dt=1e-5, c=alpha=100, phi_init=0.2, K_hill=0.05,
    n_hill=4) on the same 4 x 262144 prior batch: the configuration for which 10^7 evaluations/s on 8 GPUs is reachable."""
    import torch

    from tmcmc202601_b200 import workloads
    from tmcmc202601_b200.engine import HamiltonEngine

    try:
        solver = dict(workloads.T500_SOLVER, K_hill=0.05, n_hill=4.0)
        wl0 = workloads.four_condition_workload(8, solver=solver)
        for k in range(N_COND):  # observation days rescaled onto the 500-step grid, phi_init of the surrogate set-up
            for o, r in enumerate((68, 136, 226, 339, 475)):
                wl0.config.cond[k].idx_sparse[o] = r
            for i in range(5):
                wl0.config.cond[k].phi_init[i] = 0.2
        eng = HamiltonEngine(wl0.config, device=local_rank)
        eng.use_torch_stream()
        th, cid = device_prior_batch(workloads.CONDITION_KEYS, N_PER_COND, dev, seed=500)
        ll = torch.empty(th.shape[0], dtype=torch.float64, device=dev)
        r = timed_pass(eng, th, cid, ll, reps=3, local_rank=local_rank)
        eng.close()
        r["config"] = "dt=1e-5 x 500 steps, c=alpha=100, phi_init=0.2, K_hill=0.05 n_hill=4, 4 x 262144 prior draws"
        return r
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"}


def tmcmc_runs(local_rank: int) -> dict:
    """Wall time per TMCMC run for BASELINE.json configs[0], [2] and [3] (configs[1] is `tmcmc_run`), through the reference-shaped
    API.  No stored reference posterior exists for these (the reference needs ~60 h per run), so the posterior check is the
    agreement of the two independent RNG paths: host-side mutation (the reference's NumPy stream) vs device-side (Philox)."""
    sys.path.insert(0, str(ROOT / "scripts"))
    import run_production_tmcmc as rp

    from tmcmc202601_b200 import tmcmc, workloads
    from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

    out = {}

    def z_between(a, b, n_eff=200.0):
        return float(np.max(np.abs(a.mean(0) - b.mean(0)) / np.sqrt((a.std(0, ddof=1) ** 2 + b.std(0, ddof=1) ** 2) / n_eff + 1e-300)))

    def clocked(fn):
        sampler = ClockSampler(local_rank)
        sampler.start()
        r = fn()
        return r, sampler.stop()

    # configs[0]: Commensal Static, N = 1000, 2 chains
    try:
        (h, hs, _), ck = clocked(lambda: rp.run_condition("Commensal_Static", 1000, 2, device_mutation=False))
        (d, ds, _), _ = clocked(lambda: rp.run_condition("Commensal_Static", 1000, 2, device_mutation=True))
        act = workloads.load_conditions()["Commensal_Static"]["active_indices"]
        # the two chains advanced together through merged launches (batch.run_batch_TMCMC): same samples as chain after chain
        (mm, mo), _ = clocked(lambda: rp.run_merged(["Commensal_Static"], 1000, 2, device_mutation=True))
        merged_equal = bool(np.array_equal(np.concatenate(mo[0]["chains"]), ds))
        out["configs0_commensal_static"] = {
            "host_mutation_wall_s": h["wall_s"], "device_mutation_wall_s": d["wall_s"],
            "device_mutation_merged_chains_wall_s": mm["wall_s"], "merged_launches": mm["merged_launches"],
            "merged_samples_equal_sequential": merged_equal, "likelihood_evaluations": h["likelihood_evaluations"],
            "stages": h["stages"], "converged": h["converged"], "logL_max": [h["logL_max"], d["logL_max"]],
            "posterior_z_host_vs_device_max": z_between(hs[:, act], ds[:, act]), "clocks": ck}
    except Exception as e:  # noqa: BLE001
        out["configs0_commensal_static"] = {"error": f"{type(e).__name__}: {e}"}

    # configs[2]: the four conditions x 2 chains x 1000 particles, launches merged across the eight chains
    try:
        (hm, ho), ck = clocked(lambda: rp.run_merged(workloads.CONDITION_KEYS, 1000, 2, device_mutation=False))
        (dm, do), _ = clocked(lambda: rp.run_merged(workloads.CONDITION_KEYS, 1000, 2, device_mutation=True))
        conds = workloads.load_conditions()
        zs = []
        for k, a, b in zip(workloads.CONDITION_KEYS, ho, do):
            act = conds[k]["active_indices"]
            zs.append(z_between(np.concatenate(a["chains"])[:, act], np.concatenate(b["chains"])[:, act]))
        out["configs2_four_conditions"] = {
            "host_mutation_wall_s": hm["wall_s"], "device_mutation_wall_s": dm["wall_s"], "merged_launches": dm["merged_launches"],
            "likelihood_evaluations": dm["likelihood_evaluations"], "converged": [p["converged"] for p in dm["per_condition"]],
            "posterior_z_host_vs_device_max": zs, "clocks": ck}
    except Exception as e:  # noqa: BLE001
        out["configs2_four_conditions"] = {"error": f"{type(e).__name__}: {e}"}

    # configs[3]: Hill-gate sweep member K = 0.05, n = 4, N = 10^5 particles per stage, device-side mutation
    try:
        cd = workloads.load_conditions()["Dysbiotic_HOBIC"]
        sk = dict(workloads.PRODUCTION_SOLVER, phi_init=cd["phi_init"], n_hill=4.0)
        base = np.full(20, 1.5)
        base[cd["locked"]] = 0.0

        def run():
            ev = LogLikelihoodEvaluator(sk, [0, 1, 2, 3, 4], cd["active_indices"], base, cd["data"], cd["idx_sparse"], cd["sigma_obs"],
                                        cov_rel=0.005, theta_linearization=base, device=local_rank)
            t0 = time.perf_counter()
            res = tmcmc.run_TMCMC(ev, [tuple(b) for b in cd["bounds"]], n_particles=100000, n_stages=30, seed=4242, min_delta_beta=0.02,
                                  max_delta_beta=0.5, evaluator=ev, theta_base_full=base, active_indices=cd["active_indices"],
                                  n_mutation_steps=5, device_mutation=True)
            wall = time.perf_counter() - t0
            n = ev.health.n_calls
            ev.close()
            return res, wall, n

        (res, wall, n), ck = clocked(run)
        out["configs3_hill_n4_1e5"] = {"wall_s": wall, "stages": len(res.stage_summary), "s_per_stage": wall / max(len(res.stage_summary), 1),
                                       "likelihood_evaluations": int(n), "evals_per_s_in_run": n / wall, "log_evidence": res.log_evidence,
                                       "logL_max": float(res.logL_values.max()), "beta_final": res.beta_schedule[-1], "clocks": ck}
    except Exception as e:  # noqa: BLE001
        out["configs3_hill_n4_1e5"] = {"error": f"{type(e).__name__}: {e}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-tmcmc-run", action="store_true", help="skip the configs[1] TMCMC wall-time side measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline line only: no sweep / t500 / tmcmc side measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = synth_batch(rank)
    M = wl.theta.shape[0]

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the product path has no CPU fallback (use --impl reference for the CPU arm)")

    # CPU baselines first (rank 0, N = 1), before this process creates a CUDA context: the reference forks worker processes
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        port = cpu_baseline(wl)
        try:
            shipped = reference_as_shipped(wl)
        except Exception as e:  # noqa: BLE001
            shipped = None
            port["reference_as_shipped_error"] = f"{type(e).__name__}: {e}"
        cpu = dict(shipped, port=port) if shipped else port

    import torch.distributed as dist

    from tmcmc202601_b200.engine import HamiltonEngine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    eng = HamiltonEngine(wl.config, device=local_rank)
    eng.use_torch_stream()
    fp64_peak = eng.measure_fp64_peak()

    # ---- device-resident buffers ------------------------------------------------------------------------
    th_d = torch.from_numpy(wl.theta).to(dev)
    cid_d = torch.from_numpy(wl.cond_id).to(dev)
    ll_d = torch.empty(M, dtype=torch.float64, device=dev)
    ll_all = torch.empty(world * M, dtype=torch.float64, device=dev) if world > 1 else ll_d
    w_d = torch.empty(world * N_PER_COND, dtype=torch.float64, device=dev)
    idx_d = torch.empty(world * N_PER_COND, dtype=torch.int64, device=dev)
    mom_d = torch.empty(22, dtype=torch.float64, device=dev)
    mean_d = torch.empty(20, dtype=torch.float64, device=dev)
    sc_d = torch.empty(400, dtype=torch.float64, device=dev)
    # the batch is sorted by condition: condition c of rank r sits at [c*N_PER_COND, (c+1)*N_PER_COND) of its block
    gather_idx = torch.cat([torch.arange(c * N_PER_COND, (c + 1) * N_PER_COND, device=dev) + r * M for c in range(N_COND) for r in range(world)]
                           ).reshape(N_COND, world * N_PER_COND)
    u_draw = np.random.default_rng(1234).random(1024)  # identical on every replica

    def stage_arithmetic(step: int):
        """Per-condition stage arithmetic on the gathered population; covariance from shard-local partial sums."""
        for c in range(N_COND):
            pop = ll_all[gather_idx[c]] if world > 1 else ll_d[c * N_PER_COND:(c + 1) * N_PER_COND]
            db, _ = eng.beta_step(pop, 0.0, 0.5, 0.02, 0.5)
            eng.weights_device(pop, db, w_d)
            eng.resample_device(w_d, float(u_draw[(step * N_COND + c) % 1024]), idx_d)
            lo = c * N_PER_COND
            w_loc = w_d[rank * N_PER_COND:(rank + 1) * N_PER_COND]
            eng.weighted_moments_device(th_d[lo:lo + N_PER_COND], w_loc, mom_d)
            if world > 1:
                dist.all_reduce(mom_d)
            torch.div(mom_d[2:], mom_d[0], out=mean_d)
            eng.weighted_scatter_device(th_d[lo:lo + N_PER_COND], w_loc, mean_d, sc_d)
            if world > 1:
                dist.all_reduce(sc_d)

    def device_step(step: int):
        eng.loglik_batch_device(th_d, cid_d, ll_d)
        if world > 1:
            dist.all_gather_into_tensor(ll_all, ll_d)
        stage_arithmetic(step)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for s in range(args.warmup):
        device_step(s)
    barrier()
    launches0 = eng.counters()["kernel_launches"]
    eng.set_timing(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ode_ms, iters_sum, trips_sum = [], 0, 0
    barrier()
    ev0.record()
    for s in range(args.steps):
        device_step(args.warmup + s)
        c = eng.counters()  # synchronises this rank's stream: the step's host scalars are needed anyway
        ode_ms.append(c["last_ode_kernel_ms"])
        iters_sum += c["newton_iters"]
        trips_sum += c["q_evals"]
    ev1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = eng.counters()["kernel_launches"] - launches0
    eng.set_timing(False)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * M * args.steps / (elapsed_ms * 1e-3)

    # ---- end to end through the host-facing API ---------------------------------------------------------
    # The particle matrix is uploaded ONCE per step into a context-owned device array (hmx_population_*) and every call of the
    # step reads it there; everything else (cond_id, log-likelihoods, weights, indices, moments) crosses the bus as NumPy.
    th_h = torch.from_numpy(wl.theta).pin_memory().numpy()
    cid_h = torch.from_numpy(wl.cond_id).pin_memory().numpy()
    th_pop = eng.device_array((M, 20))
    pop_n = world * N_PER_COND
    # results land in pinned host arrays (what a host program reads after the step): log-likelihoods of this rank's block and
    # the resampling indices of the gathered population; the gathered log-likelihoods and the weights never leave the device
    ll_h = torch.empty(M, dtype=torch.float64).pin_memory().numpy()
    idx_h = torch.empty(pop_n, dtype=torch.int64).pin_memory().numpy()
    mom_h, sc_h = np.empty(22), np.empty((20, 20))

    def host_step(step: int):
        th_pop.upload(th_h)                                  # H2D theta (pinned source)
        eng.loglik_batch_device(th_pop, cid_h, ll_d)         # H2D cond_id; logL stays on the device for the all-gather ...
        torch.from_numpy(ll_h).copy_(ll_d)                   # ... and is read back: D2H logL (the step's result)
        if world > 1:
            dist.all_gather_into_tensor(ll_all, ll_d)
        out = None
        for c in range(N_COND):
            pop = ll_all[gather_idx[c]] if world > 1 else ll_d[c * N_PER_COND:(c + 1) * N_PER_COND]
            db, _ = eng.beta_step(pop, 0.0, 0.5, 0.02, 0.5)                   # D2H two scalars
            eng.weights_device(pop, db, w_d)                                  # D2H log S_j
            eng.resample_device(w_d, float(u_draw[(step * N_COND + c) % 1024]), idx_h)  # D2H indices (pinned destination)
            lo = c * N_PER_COND
            w_loc = w_d[rank * N_PER_COND:(rank + 1) * N_PER_COND]
            mom = eng.weighted_moments(th_pop[lo:lo + N_PER_COND], w_loc)     # D2H 22 moments
            if world > 1:
                mt = torch.from_numpy(mom).to(dev)
                dist.all_reduce(mt)
                mom = mt.cpu().numpy()
            sc = eng.weighted_scatter(th_pop[lo:lo + N_PER_COND], w_loc, mom[2:] / mom[0])  # H2D mean, D2H 20 x 20
            if world > 1:
                st = torch.from_numpy(sc).to(dev)
                dist.all_reduce(st)
                sc = st.cpu().numpy()
            out = (idx_h, sc)
        return out

    host_step(0)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_steps = args.steps
    for s in range(e2e_steps):
        host_step(1 + s)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)  # e1 is recorded after the last host-facing call returned (each call synchronises)
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * M * e2e_steps / (float(te.item()) * 1e-3)
    # bytes the host-facing calls copy per step: theta + cond_id in, logL out; per condition the 20-vector mean in (and, on
    # several GPUs, the all-reduced moments / scatter); the resampling indices, 3 scalars, 22 moments and the 20 x 20 scatter out
    h2d = M * (20 * 8 + 4) + N_COND * (20 * 8 + (0 if world == 1 else (22 + 400) * 8))
    d2h = M * 8 + N_COND * (pop_n * 8 + 3 * 8 + (22 + 400) * 8 * (1 if world == 1 else 2))

    # ---- roofline of the dominant kernel (ode_kernel): executed FP64 flops / CUDA-event duration ---------
    ode_avg_ms = float(np.mean(ode_ms))
    flops_per_launch = (FLOPS_PER_RESIDUAL * trips_sum + FLOPS_PER_DIRECTION * iters_sum) / args.steps
    achieved = flops_per_launch / (ode_avg_ms * 1e-3) / 1e12
    # rejected line-search trials = residual evaluations that were not followed by a direction (step starts are rebased, not evaluated)
    ref_equiv = ((REF_FLOPS_PER_ITER * iters_sum + REF_FLOPS_PER_TRIAL * max(trips_sum - iters_sum, 0)) / args.steps
                 / (ode_avg_ms * 1e-3) / 1e12)
    roofline = {
        "bound": "fp64", "kernel": "hmx::ode_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak if fp64_peak > 0 else None,
        "frac_of_nominal_peak": achieved / FP64_NOMINAL_TFLOPS,
        "traffic": DRAM_BYTES_PER_EVAL_NCU * M,  # bytes per launch; ncu measured per evaluation, algorithmic 1144
        "peak_source": "measured live by hmx_measure_fp64_peak (register-resident DFMA loop); MEASURED_PEAKS.json has no FP64 entry "
                       f"(nominal 148 SM x 64 FMA x 2 x 1.965 GHz = {FP64_NOMINAL_TFLOPS}: frac_of_nominal_peak)",
        "kernel_ms_per_launch": ode_avg_ms, "kernel_share_of_step": ode_avg_ms * args.steps / elapsed_ms,
        "flops_per_launch": flops_per_launch, "newton_iters_per_eval": iters_sum / (M * args.steps),
        "residual_evals_per_eval": trips_sum / (M * args.steps),
        "reference_equivalent_tflops": ref_equiv,
        # SURVEY 8(d)'s ALGORITHMIC count (dense Jacobian + LU: 2778 flop per iteration, 384 per extra trial) over the same
        # kernel time: what the reference's formulation would have to sustain for this evaluation rate.  `achieved` / `frac`
        # above count the flops this kernel actually executes (block-structured solve), which is the smaller number.
        "algorithmic_frac_of_peak": ref_equiv / fp64_peak if fp64_peak > 0 else None,
        "hbm_algorithmic_bytes_per_eval": 20 * 8 + 4 + 5 * 24 * 8 + 12 + 8,
    }

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.quick:
            line["sweep"] = throughput_sweep(eng, dev, local_rank)
            line["t500"] = t500_pass(dev, local_rank)
        if world == 1 and not args.no_tmcmc_run and not args.quick:
            line["tmcmc_run"] = tmcmc_run()
            line["tmcmc_runs"] = tmcmc_runs(local_rank)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
