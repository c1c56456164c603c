"""tmcmc202601_b200 -- B200-native hot path of keisuke58/Tmcmc202601's Stage-1 TMCMC likelihood.

Public surface (names follow the reference):
    solver.BiofilmNewtonSolver5S, solver.BiofilmNewtonSolver (legacy 4-species, alpha_schedule), tsm.BiofilmTSM
    evaluator.LogLikelihoodEvaluator, log_likelihood_sparse, build_species_sigma, build_replicate_sigma,
        build_likelihood_weights, ViscoelasticPrior
    tmcmc.run_TMCMC, run_multi_chain_TMCMC, TMCMCResult, systematic_resample, evaluate_particles_parallel,
        reflect_into_bounds, compute_MAP_with_uncertainty
    nuts.tmcmc_engine (RW / HMC / NUTS TMCMC on the likelihood gradient), evaluator.LogLikelihoodEvaluator.value_and_grad
    posterior.compute_hdi, compute_credible_intervals
    engine.HamiltonEngine (direct handle on the C ABI of include/hmx.h), distributed.ShardedEvaluator / ShardedStageOps
Nothing here imports the CPU oracle; without libhamilton.so and a B200 the compute entry points raise.
"""

__version__ = "0.1.0"
