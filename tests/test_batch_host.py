"""run_batch_TMCMC (concurrent chains, merged likelihood launches) on the CPU with oracle-backed stand-ins:
the merged run must be bit-identical to running every chain on its own."""
import numpy as np

from helpers import OracleEngine, OracleStageOps
from tmcmc202601_b200 import batch, tmcmc, workloads
from tmcmc202601_b200.evaluator import LogLikelihoodEvaluator

SOLVER = dict(workloads.PRODUCTION_SOLVER, maxtimestep=80)  # short horizon keeps the CPU oracle fast


def _jobs():
    conds = workloads.load_conditions()
    jobs = []
    for key in ("Commensal_Static", "Dysbiotic_HOBIC"):
        cd = conds[key]
        base = np.full(20, 1.5)
        base[cd["locked"]] = 0.0
        jobs.append(dict(model_tag=key, data=cd["data"], idx_sparse=np.array([10, 25, 40, 60, 79]), sigma_obs=cd["sigma_obs"],
                         phi_init=cd["phi_init"], prior_bounds=cd["bounds"], theta_base=base, active_indices=cd["active_indices"]))
    return jobs


def test_merged_run_equals_independent_chains():
    jobs = _jobs()
    engines = []

    def factory(cfg):
        e = OracleEngine(cfg, n_threads=4)
        engines.append(e)
        return e

    out, stats = batch.run_batch_TMCMC(jobs, SOLVER, n_particles=24, n_stages=4, n_chains=2, seed=7, n_mutation_steps=2,
                                       stage_ops=OracleStageOps(), engine_factory=factory)
    assert len(out) == 2 and stats["merged_launches"] >= 1
    # every merged launch carried the requests of several chains
    sizes = [len(c[0]) for c in engines[0].calls]
    assert max(sizes) >= 4 * 24 and stats["likelihood_evaluations"] == sum(sizes)
    for k, j in enumerate(jobs):
        for chain in range(2):
            ev = LogLikelihoodEvaluator(dict(SOLVER, phi_init=j["phi_init"]), [0, 1, 2, 3, 4], list(j["active_indices"]), j["theta_base"],
                                        j["data"], j["idx_sparse"], j["sigma_obs"], cov_rel=0.005, theta_linearization=j["theta_base"],
                                        engine_factory=lambda cfg: OracleEngine(cfg, n_threads=4))
            s = (7 + tmcmc._stable_hash_int(j["model_tag"]) % (2**31) + 2 * 1000 + chain) % (2**31)
            solo = tmcmc.run_TMCMC(ev, [tuple(b) for b in j["prior_bounds"]], n_particles=24, n_stages=4, seed=s, min_delta_beta=0.02,
                                   max_delta_beta=0.5, evaluator=ev, theta_base_full=j["theta_base"], active_indices=list(j["active_indices"]),
                                   n_mutation_steps=2, stage_ops=OracleStageOps())
            assert np.array_equal(solo.samples, out[k]["chains"][chain])
            assert np.array_equal(solo.logL_values, out[k]["logL"][chain])
            assert solo.beta_schedule == out[k]["results"][chain].beta_schedule


def test_broker_survives_uneven_chain_lengths():
    """Chains that finish early must not dead-lock the others (client_done re-evaluates the barrier)."""
    jobs = _jobs()[:1]
    out, stats = batch.run_batch_TMCMC(jobs, SOLVER, n_particles=16, n_stages=3, n_chains=3, seed=3, n_mutation_steps=1,
                                       stage_ops=OracleStageOps(), engine_factory=lambda cfg: OracleEngine(cfg, n_threads=4))
    assert len(out[0]["chains"]) == 3 and all(c.shape == (16, 20) for c in out[0]["chains"])


def test_merged_device_mutation_equals_independent_chains():
    """device_mutation=True through the broker: the chains' mutation steps are merged into mutate_groups calls, and every
    chain still equals the same chain run on its own (the Philox counters do not depend on the grouping)."""
    jobs = _jobs()
    out, stats = batch.run_batch_TMCMC(jobs, SOLVER, n_particles=24, n_stages=4, n_chains=2, seed=7, n_mutation_steps=2,
                                       stage_ops=OracleStageOps(), engine_factory=lambda cfg: OracleEngine(cfg, n_threads=4),
                                       device_mutation=True)
    assert stats["merged_launches"] >= 2 and stats["likelihood_evaluations"] > 4 * 24
    for k, j in enumerate(jobs):
        for chain in range(2):
            ev = LogLikelihoodEvaluator(dict(SOLVER, phi_init=j["phi_init"]), [0, 1, 2, 3, 4], list(j["active_indices"]), j["theta_base"],
                                        j["data"], j["idx_sparse"], j["sigma_obs"], cov_rel=0.005, theta_linearization=j["theta_base"],
                                        engine_factory=lambda cfg: OracleEngine(cfg, n_threads=4))
            s = (7 + tmcmc._stable_hash_int(j["model_tag"]) % (2**31) + 2 * 1000 + chain) % (2**31)
            solo = tmcmc.run_TMCMC(ev, [tuple(b) for b in j["prior_bounds"]], n_particles=24, n_stages=4, seed=s, min_delta_beta=0.02,
                                   max_delta_beta=0.5, evaluator=ev, theta_base_full=j["theta_base"], active_indices=list(j["active_indices"]),
                                   n_mutation_steps=2, stage_ops=OracleStageOps(), device_mutation=True)
            assert np.array_equal(solo.samples, out[k]["chains"][chain])
            assert np.array_equal(solo.logL_values, out[k]["logL"][chain])
            assert solo.acc_rate_history == out[k]["results"][chain].acc_rate_history


def test_unmerged_streams_mode_equals_merged():
    """merge_launches=False (every chain on its own context / stream) returns the samples of the merged run bit for bit and
    counts one launch per chain and round instead of one per round."""
    jobs = _jobs()
    kw = dict(n_particles=24, n_stages=4, n_chains=2, seed=7, n_mutation_steps=2, stage_ops=OracleStageOps(),
              engine_factory=lambda cfg: OracleEngine(cfg, n_threads=4), device_mutation=True)
    merged, ms = batch.run_batch_TMCMC(jobs, SOLVER, **kw)
    free, fs = batch.run_batch_TMCMC(jobs, SOLVER, merge_launches=False, **kw)
    assert fs["likelihood_evaluations"] == ms["likelihood_evaluations"] and fs["merged_launches"] > ms["merged_launches"]
    for a, b in zip(merged, free):
        for chain in range(2):
            assert np.array_equal(a["chains"][chain], b["chains"][chain])
            assert np.array_equal(a["logL"][chain], b["logL"][chain])
