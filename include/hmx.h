/*
 * hmx.h -- C ABI of libhamilton.so: the B200-native (sm_100a) implementation of the Stage-1 TMCMC
 * likelihood hot path of keisuke58/Tmcmc202601.
 *
 * The reference has no FFI: its "plugin" interface is duck-typed Python.  Each entry
 * point below names the reference interface it replaces (paths relative to the reference root):
 *   S5 = tmcmc/program2602/improved_5species_jit.py      S4 = tmcmc/program2602/improved1207_paper_jit.py
 *   EV = data_5species/core/evaluator.py                  TM = data_5species/core/tmcmc.py
 *   MAIN = data_5species/main/estimate_reduced_nishioka.py
 * The ctypes binding a maintainer would add on the reference side is shown in INTEGRATION.md; this
 * repository's own binding is tmcmc202601_b200/_abi.py.
 *
 * Conventions
 *   - plain C, no exceptions: every call returns 0 on success or a negative hmx_status; the message is
 *     available from hmx_last_error().
 *   - every array argument is caller-owned and may be a HOST pointer (pageable or pinned) or a DEVICE
 *     pointer on the context's GPU (detected with cudaPointerGetAttributes).  Host inputs are copied into
 *     device buffers owned by the context (cudaMemcpyAsync on the context stream; pass pinned memory for
 *     full PCIe rate); a call with any host OUTPUT pointer returns after the
 *     results have landed (stream-synchronised); a call whose outputs are all device pointers only
 *     enqueues work on the context stream (hmx_set_stream / hmx_synchronize).
 *   - calls on one context serialise (not thread-safe, like the reference evaluator whose histories are
 *     mutable).  Several contexts may live on one GPU, one host thread each: every context owns a
 *     non-blocking stream and takes its scratch memory from the stream-ordered allocator
 *     (cudaMallocAsync / cudaFreeAsync), so creating, growing or destroying one context never
 *     synchronises the device under another context's running kernels.
 *   - failed solves follow the reference's error convention: logL = -1e20, never an error return
 *     (EV:655,667,753; TM:306).
 *   - there is NO CPU fallback: hmx_create fails when no sm_100 device is present.
 */
#ifndef HMX_H
#define HMX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMX_ABI_VERSION 1
#define HMX_NSTATE 12  /* g = [phi_1..5, phi_0, psi_1..5, gamma]  (S5:8-13) */
#define HMX_NTHETA 20  /* S5:15-22 */
#define HMX_MAX_OBS 8
#define HMX_MAX_COND 8

typedef enum {
    HMX_OK = 0,
    HMX_ERR_INVALID = -1,   /* bad argument */
    HMX_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    HMX_ERR_NO_DEVICE = -3, /* no sm_100 GPU visible: the product path refuses to run */
    HMX_ERR_OOM = -4
} hmx_status;

/* status word per evaluation (hmx_loglik_batch `status`) */
#define HMX_ST_NONFINITE 1u /* trajectory / likelihood inputs non-finite -> logL = -1e20 (EV:657-667,747-753) */
#define HMX_ST_MAXITER 2u   /* a time step exhausted max_newton_iter (S5:308) */
#define HMX_ST_SINGULAR 4u  /* a Newton direction came out non-finite (the reference's LinAlgError/ridge path, S5:369-373) */
#define HMX_ST_BAD_COND 8u  /* cond_id outside [0, n_cond) in a DEVICE array -> logL = -1e20 (a host array is rejected with
                               HMX_ERR_INVALID before anything runs) */

/* One experimental condition: what LogLikelihoodEvaluator.__init__ receives (EV:381-398) plus the
 * per-condition initial state from the loader (MAIN:1907-1928). */
typedef struct {
    double phi_init[5];                /* solver_kwargs["phi_init"], scalar already broadcast (S5:489-495) */
    int32_t n_obs;                     /* rows of `data` (5 with --use-exp-init, 6 in older run dirs) */
    int32_t idx_sparse[HMX_MAX_OBS];   /* row indices into the trajectory, 0..n_steps (MAIN:1827-1867) */
    double data[HMX_MAX_OBS * 5];      /* row-major [n_obs,5] */
    double sigma_obs[HMX_MAX_OBS * 5]; /* scalar / [5] sigma pre-broadcast to [n_obs,5] (EV:286-290) */
    double weights[HMX_MAX_OBS * 5];   /* 1.0 where the reference passes weights=None (EV:332) */
    double cov_rel;                    /* BiofilmTSM.cov_rel (S4:1067) */
    uint32_t active_theta_mask;        /* bit k set <=> k in BiofilmTSM.active_idx (S4:1075) */
    int32_t use_absolute_volume;       /* EV:703-711: mu = phi*gamma instead of phi*psi */
    int32_t tsm_variance;              /* 1 = EV:700-745 variance (evaluator parity), 0 = deterministic only */
} hmx_cond;

/* mirrors solver_kwargs (MAIN:1919-1928) + BiofilmNewtonSolver5S.__init__ defaults (S5:464-511) */
typedef struct {
    int32_t abi_version; /* must be HMX_ABI_VERSION */
    double dt;
    int32_t n_steps; /* maxtimestep */
    double eps;      /* 1e-6 */
    double Kp1;
    double c_const;
    double alpha_const;
    double K_hill;
    double n_hill;
    int32_t max_newton_iter; /* 50 */
    double eta[5];
    double eta_phi[5];
    int32_t active_species_mask; /* bit i <=> species i active in the initial state (S5:566-575) */
    int32_t n_cond;              /* 1..HMX_MAX_COND */
    hmx_cond cond[HMX_MAX_COND];
} hmx_config;

typedef struct hmx_ctx hmx_ctx;

/* ---- lifetime ------------------------------------------------------------------------------------ */
/* replaces: BiofilmNewtonSolver5S(**solver_kwargs) + LogLikelihoodEvaluator(...) construction
 * (S5:464-511, EV:381-490).  Uploads the constants to __constant__ memory of `device`. */
int hmx_create(hmx_ctx** ctx, const hmx_config* cfg, int32_t device);
void hmx_destroy(hmx_ctx* ctx);
/* message of the last failing call on ctx (ctx may be NULL: last hmx_create failure) */
const char* hmx_last_error(const hmx_ctx* ctx);
/* run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); NULL = the context's own (non-blocking)
 * stream.  The legacy default stream has handle 0: name it with cudaStreamLegacy ((cudaStream_t)0x1) -- work enqueued on
 * the context's own stream is NOT ordered with work on the default stream. */
int hmx_set_stream(hmx_ctx* ctx, void* cuda_stream);
int hmx_synchronize(hmx_ctx* ctx);
int hmx_abi_version(void);

/* ---- device-resident arrays ("populations") --------------------------------------------------------
 * Every array argument of this API may be a device pointer.  A caller without a CUDA binding of its own (the
 * reference is NumPy-only) gets device arrays from the context: upload the particle matrix ONCE per stage and pass
 * the returned pointer (or pointer + offset: plain row-major doubles) to hmx_loglik_batch, hmx_weighted_cov /
 * _moments / _scatter and hmx_mutate instead of re-sending the host copy with every call.  Arrays are freed by
 * hmx_population_destroy or with the context.  upload is asynchronous on the context stream (the source must stay
 * valid until the next synchronising call; pass pinned memory for full PCIe rate), download synchronises. */
int hmx_population_create(hmx_ctx* ctx, int64_t n_rows, int32_t n_cols, double** dev);
int hmx_population_destroy(hmx_ctx* ctx, double* dev);
int hmx_population_upload(hmx_ctx* ctx, double* dev, const double* host, int64_t n_values);
int hmx_population_download(hmx_ctx* ctx, double* host, const double* dev, int64_t n_values);

/* ---- the hot path -------------------------------------------------------------------------------- */
/* replaces: evaluate_particles_parallel(LogLikelihoodEvaluator, theta) (TM:247-312) for FULL 20-vectors:
 * per evaluation BiofilmTSM.solve_tsm -> run_deterministic (S5:561-581, 837-882; S4:1077-1160) and
 * LogLikelihoodEvaluator.__call__ from EV:645 on (guards, mu/var at idx_sparse, log_likelihood_sparse
 * EV:237-371 with rho=0).  The theta_sub -> theta_full map of EV:635-637 stays on the host side.
 *   theta        [N,20] row-major
 *   cond_id      [N] int32 index into cfg.cond, or NULL (= condition 0)
 *   logL         [N]
 *   obs_state    [N,n_obs_max,12] trajectory rows at idx_sparse (n_obs_max = max n_obs over conditions), or NULL
 *   status       [N] HMX_ST_* bits, or NULL
 *   newton_iters [N] quasi-Newton iterations summed over the time steps, or NULL */
int hmx_loglik_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N,
                     double* logL, double* obs_state, uint32_t* status, uint32_t* newton_iters);

/* replaces: BiofilmTSM.solve_tsm (S4:1077-1160) as the evaluator consumes it (EV:645-745), i.e. at the rows idx_sparse:
 *   x0     [N,n_obs_max,12]     = x0[idx_sparse]          (the deterministic trajectory rows), or NULL
 *   sigma2 [N,n_obs_max,12]     = sig2[idx_sparse]        (S4:1143-1149: sum_k x1_k^2 (cov_rel theta_k)^2 + 1e-12), or NULL
 *   x1     [N,n_obs_max,12,20]  = tsm._last_x1[idx_sparse] (S4:1150) with one column per theta_k (zero where k is not in
 *                                 active_idx; the reference stores only the active columns, in active_idx order), or NULL
 *   logL   [N] or NULL, status [N] or NULL as in hmx_loglik_batch.
 * The local sensitivity x1[t,:,k] = solve(J(g_t, g_{t-1}), -dQ/dtheta_k) depends on (g_t, g_{t-1}) only, which is why the
 * observation rows suffice.  idx_sparse = 0 returns the row-1 sensitivities (S4:1139-1141).  Conditions with
 * tsm_variance = 0 (the np.allclose short-cut, TSMA:366-401) return zeros. */
int hmx_tsm_obs_rows(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* x0, double* sigma2,
                     double* x1, double* logL, uint32_t* status);

/* replaces: BiofilmTSM.solve_tsm (S4:1077-1160) in FULL: t_arr is arange(n_steps + 1) * dt,
 *   x0     [N,n_steps+1,12] the deterministic trajectory (may be NULL),
 *   sigma2 [N,n_steps+1,12] the TSM variance at every time level (level 0 repeats level 1, S4:1139-1141; zeros for conditions
 *          with tsm_variance = 0),
 *   status [N] or NULL: HMX_ST_NONFINITE when any entry of x0 or sigma2 is non-finite -- the whole-trajectory test of
 *          EV:657-667.
 * Costs about 2.5 solves per evaluation (one residual + n_active direction solves per time level; the reference spends
 * 98 % of its wall time here) and 480 KB of output per evaluation: for likelihoods use hmx_loglik_batch, which
 * needs the variance at the observation rows only. */
int hmx_solve_tsm_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* x0, double* sigma2,
                        uint32_t* status);
/* EV:657-667 returns -1e20 when sig2 is non-finite at ANY of the T + 1 rows.  hmx_loglik_batch checks the state at every row
 * and sig2 at the observation rows (all_rows = 0, default: a variance can only blow up off the observation rows, with a finite
 * state, where the Jacobian is numerically singular); all_rows = 1 evaluates sig2 at every row like the reference (about 2.5x the
 * cost per evaluation; identical results whenever the guard does not fire). */
int hmx_set_sig2_guard(hmx_ctx* ctx, int32_t all_rows);

/* replaces: jax.value_and_grad(log_likelihood) as the reference's NUTS / HMC engine consumes it
 * (data_5species/main/tmcmc_nuts_engine.py:202; leapfrog :20-34, tempered_vg :247-249): value and gradient of the
 * log-likelihood of hmx_loglik_batch for a batch of FULL 20-vectors.
 *   grad   [N,20]  d logL / d theta_k for all 20 parameters, by forward sensitivities of the implicit-Euler trajectory through
 *                  all n_steps levels (exact Jacobian; csrc/hmx_grad.cuh), with the likelihood variance (sigma_obs^2 + the TSM
 *                  variance) frozen at theta.  Evaluations flagged HMX_ST_NONFINITE get logL = -1e20 and a zero gradient.
 *   dstate [N,n_obs_max,12,20] or NULL: the state sensitivities d g[idx_sparse] / d theta behind it, from which the gradient of
 *                  any other observable-based likelihood (e.g. the normalised-fraction Gaussian of
 *                  estimate_reduced_nishioka_jax.py:115-133) follows by the chain rule.
 * The gradient is that of the converged implicit-Euler map evaluated on the solver's trajectory (which carries the Newton
 * tolerance eps): it equals central differences of a tightly converged solve to <= 1e-6, and differs from them by O(eps)-level
 * relative amounts (1e-4 observed) at the production tolerance. */
int hmx_loglik_grad_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* logL, double* grad,
                          double* dstate, uint32_t* status);

/* replaces: BiofilmNewtonSolver5S.run_deterministic / .solve (S5:561-581, 646) for a batch.
 *   g [N,n_steps+1,12] row-major;  t_arr is (arange(n_steps+1) * dt) and is left to the caller. */
int hmx_trajectory_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N,
                         double* g, uint32_t* status, uint32_t* newton_iters);
/* Same solve, but only the time levels rows[0..n_rows) (strictly ascending, in [0, n_steps]; host or device array) are stored:
 * g[N, n_rows, 12].  For consumers that down-sample in time -- posterior-predictive bands, the surrogate training
 * set of deeponet/generate_training_data.py:283-289 (np.linspace(0, T, n_time_out, dtype=int)). */
int hmx_trajectory_rows(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, const int32_t* rows,
                        int32_t n_rows, double* g, uint32_t* status, uint32_t* newton_iters);

/* ---- legacy 4-species solver -------------------------------------------------------------------
 * replaces: BiofilmNewtonSolver(**kw).run_deterministic(theta) of tmcmc/program2602/improved1207_paper_jit.py
 * (S4:715-1053; numba integrator S4:610-700) for a batch: state g (10) = [phi1..4, phi0, psi1..4, gamma], theta (14) =
 * a11,a12,a22,b1,b2, a33,a34,a44,b3,b4, a13,a14,a23,a24; inactive species are locked to phi = psi = 0 (S4:92-113).
 * alpha_schedule (S4:783-822) arrives resolved to one switch level on the time grid: the step that produces time level n
 * uses alpha_after when n >= alpha_switch_level, alpha_before otherwise (alpha_switch_level < 0: alpha_const throughout,
 * which is what the reference's numba path does with any schedule, S4:1024; the reference honours the schedule only in its
 * pure-Python fallback, S4:895-913).  n_species = 5 runs the same generic code on the 5-species model without the Hill
 * gate (theta [N,20], state 12): a cross-check of this code path against hmx_trajectory_batch.
 *   theta [N, 14 | 20];  rows: n_rows strictly ascending time levels in [0, n_steps] (host array), or NULL = all levels;
 *   g [N, n_rows | n_steps + 1, 10 | 12];  status / newton_iters [N] or NULL.  The solver settings travel with the call;
 *   the context supplies the device, the stream and scratch memory. */
typedef struct {
    int32_t n_species;           /* 4 (S4) or 5 (cross-check) */
    double dt;
    int32_t n_steps;
    double eps, Kp1, c_const, alpha_const;
    int32_t max_newton_iter;     /* the reference's numba loop hard-codes 50 (S4:685) */
    double eta[5], eta_phi[5];
    double phi_init[5];          /* the 4-species solver takes a scalar (S4:761): broadcast it */
    int32_t active_species_mask; /* bit i <=> species i active */
    int32_t alpha_switch_level;
    double alpha_before, alpha_after;
} hmx_legacy_config;
int hmx_legacy_trajectory_rows(hmx_ctx* ctx, const hmx_legacy_config* cfg, const double* theta, int64_t N, const int32_t* rows,
                               int32_t n_rows, double* g, uint32_t* status, uint32_t* newton_iters);

/* ---- per-stage TMCMC arithmetic ------------------------------------------------------------------ */
/* replaces: the 25-step ESS bisection for delta-beta (TM:613-659, beta_schedule="ess"). ess_at_lo is NaN
 * when no bisection step was accepted (the reference's ess_at_delta_low = None). */
int hmx_beta_step(hmx_ctx* ctx, const double* logL, int64_t N, double beta, double target_ess_ratio,
                  double min_delta_beta, double max_delta_beta, double* delta_beta, double* ess_at_lo);
/* replaces: weights + evidence increment (TM:686-701). delta_beta is (beta_next - beta) as the reference
 * forms it.  *log_Sj is NaN and w uniform when the reference would fall back ("Weight sum issue"). */
int hmx_weights(hmx_ctx* ctx, const double* logL, int64_t N, double delta_beta, double* w, double* log_Sj);
/* cs = np.cumsum(w) bit for bit (the sequentially rounded chain of tmcmc.py:195), with cs[N-1] := 1.0 (tmcmc.py:196).
 * Computed tile-parallel where that is provably equal to the chain, as the literal chain elsewhere (csrc/hmx_stage.cuh). */
int hmx_cumsum(hmx_ctx* ctx, const double* w, int64_t N, double* cs);
/* replaces: systematic_resample (TM:186-198) with the uniform draw u = rng.random() supplied by the host.
 * idx [N] int64, bit-exact with NumPy for identical (w, u). */
int hmx_resample(hmx_ctx* ctx, const double* w, int64_t N, double u, int64_t* idx);
/* replaces: np.cov(theta.T, aweights=w) at TM:766 (ddof=1).  mean [d] (may be NULL), cov [d,d]. */
int hmx_weighted_cov(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d,
                     double* mean, double* cov);
/* the two shard-local halves of hmx_weighted_cov for particles sharded over GPUs: the host all-reduces
 * `moments` (2+d doubles: sum w, sum w^2, sum w*theta_j) between them, then all-reduces `scatter` (d*d). */
int hmx_weighted_moments(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d,
                         double* moments);
int hmx_weighted_scatter(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d,
                         const double* mean, double* scatter);

/* replaces: log_likelihood_sparse(mu, sig, data, sigma_obs, rho=0, weights) (EV:237-371), diagonal branch.
 * mu, sig, data, sigma_obs, weights: [n_obs,5] row-major (sigma_obs pre-broadcast; weights may be NULL = 1).
 * *logL receives the scalar; n_obs may exceed HMX_MAX_OBS here. */
int hmx_log_likelihood_sparse(hmx_ctx* ctx, const double* mu, const double* sig, const double* data,
                              const double* sigma_obs, const double* weights, int32_t n_obs, double* logL);

/* ---- several GPUs of one box behind one handle -------------------------------------------------
 * For HOST callers (the reference is a single NumPy process): one context per listed device, the flat evaluation list
 * split into contiguous blocks of ceil(N / n_devices), every block copied, solved and copied back by its own device
 * concurrently (one host thread per device).  The log-likelihoods land in the caller's array in input order -- the
 * "all-gather" -- so the replicated stage arithmetic (hmx_beta_step / hmx_weights / hmx_resample on
 * hmx_multi_context(m, 0)) sees the whole population; hmx_multi_weighted_cov is the "all-reduce": shard-local moments and
 * scatter summed in device order.  Evaluations are independent, so there is no device-to-device traffic on this path.
 * Device-RESIDENT multi-GPU callers use one process and one hmx_ctx per GPU and exchange log-likelihoods / covariance
 * partials with NCCL (tmcmc202601_b200/distributed.py, `torchrun`; INTEGRATION.md shows the recipe).
 * Likelihood values do not depend on the number of devices (bitwise); the covariance agrees to rounding (the partial
 * sums associate by shard). */
typedef struct hmx_multi hmx_multi;
int hmx_multi_create(hmx_multi** m, const hmx_config* cfg, const int32_t* device_ids, int32_t n_devices);
void hmx_multi_destroy(hmx_multi* m);
const char* hmx_multi_last_error(const hmx_multi* m);
int32_t hmx_multi_device_count(const hmx_multi* m);
hmx_ctx* hmx_multi_context(hmx_multi* m, int32_t k);
/* hmx_loglik_batch over all devices; theta / cond_id / logL / status / newton_iters are HOST arrays */
int hmx_multi_loglik_batch(hmx_multi* m, const double* theta, const int32_t* cond_id, int64_t N, double* logL, uint32_t* status,
                           uint32_t* newton_iters);
/* hmx_weighted_cov with the particles sharded over the devices (host arrays) */
int hmx_multi_weighted_cov(hmx_multi* m, const double* theta, const double* w, int64_t N, int32_t d, double* mean, double* cov);

/* ---- measurement support ------------------------------------------------------------------------- */
typedef struct {
    int64_t evals;         /* evaluations processed by the last hmx_loglik_batch / hmx_trajectory_batch */
    int64_t newton_iters;  /* sum of quasi-Newton iterations (direction solves) */
    int64_t q_evals;       /* full residual evaluations = state-machine trips (line-search trials + the iteration starts that
                              cannot reuse a trial's residual; the first residual of a time step is derived from the accepted
                              trial that ended the previous one and is not counted) */
    int64_t kernel_launches; /* CUDA kernels launched by this context since creation */
    double last_ode_kernel_ms; /* CUDA-event duration of the ODE kernel of the last batch (0 if not timed) */
} hmx_counters;
/* synchronises the stream */
int hmx_get_counters(hmx_ctx* ctx, hmx_counters* out);
/* Latency regime (production TMCMC stages: 10^3-10^4 evaluations per launch, which cost the latency of the slowest solve):
 * batches of at most max_evals evaluations run K1L, eight lanes per solve (csrc/hmx_coop.cuh), 20-37 % faster per launch.
 * max_evals < 0 selects the measured optimum (one resident round of groups, 7104 on a B200), 0 switches K1L off (default).
 * K1L agrees with the default kernel to ~1e-12 relative on logL, not bit for bit: with it on, a value may depend (at that
 * level) on the size of the batch it is evaluated in. */
int hmx_set_latency_kernel(hmx_ctx* ctx, int64_t max_evals);
/* enable CUDA-event timing of the ODE kernel inside hmx_loglik_batch (adds two event records) */
int hmx_set_timing(hmx_ctx* ctx, int32_t enabled);
/* register-resident DFMA loop: measured FP64 FMA peak of this GPU in TFLOP/s (roofline denominator) */
int hmx_measure_fp64_peak(hmx_ctx* ctx, double* tflops);

/* ---- device-side MH mutation of one TMCMC stage ---------------------------------------------------------------
 * Replaces the proposal / evaluate / accept block of run_TMCMC's `mutate` (core/tmcmc.py:849-962) for the Gaussian
 * proposal: for each of the N resampled particles, K proposals mean_i + factor z (z ~ N(0, I_d), Philox4x32-10),
 * reflected into [low, high] (tmcmc.py:201-226), mapped to the 20-vector of the ODE (evaluator.py:631-637), solved
 * by the same kernels as hmx_loglik_batch, and accepted sequentially with log u < beta (logL' - logL)
 * (tmcmc.py:908-962; uniform prior inside the box).  Nothing but the outputs crosses the PCIe bus.
 * Parity with the reference is statistical: NumPy's PCG64 stream is not reproduced (the host-side mirror in
 * tmcmc202601_b200/tmcmc.py keeps the bitwise stream).  The random numbers depend on (seed, sequence,
 * index_offset + i, k) only, so a population sharded over several GPUs gets what one GPU would draw. */
#define HMX_MUT_MAX_DIM 32
#define HMX_MUT_MAX_GROUPS 64
typedef struct {
    int32_t d;                          /* dimension of the particles (len(prior_bounds)), 1..HMX_MUT_MAX_DIM */
    int32_t K;                          /* proposals per particle (n_mutation_steps after adaptation) */
    double factor[HMX_MUT_MAX_DIM * HMX_MUT_MAX_DIM]; /* row-major [d,d]; factor factor^T = proposal covariance */
    double low[HMX_MUT_MAX_DIM], high[HMX_MUT_MAX_DIM];
    double theta_base[HMX_NTHETA];      /* theta_full starts from this ... */
    int32_t map[HMX_MUT_MAX_DIM];       /* ... and theta_full[map[j]] = theta_sub[j], j < n_map (entries >= 20 are ignored) */
    int32_t n_map;
    double theta_lin[HMX_NTHETA];       /* np.allclose short-cut (tsm_analytical.py:366-401): a proposal within rtol 1e-5 /  */
    int32_t cond_default, cond_at_lin;  /* atol 1e-8 of theta_lin is evaluated as condition cond_at_lin (< 0: short-cut off) */
    double beta;                        /* beta_next of the stage */
    uint64_t seed, sequence;            /* RNG key and per-call tag (stage / recovery attempt); sequence < 2^30 */
    uint64_t index_offset;              /* global index of particle 0 of this call (sharded populations) */
} hmx_mutation;

typedef struct {
    int64_t n_accepted, n_candidates;   /* accepted proposals / proposals inside the prior (tmcmc.py:952-960) */
    int64_t n_nonfinite, n_maxiter, n_singular; /* status bits over the N K solves */
} hmx_mutation_stats;

/* theta_res[N,d], logL_res[N] -> theta_out[N,d], logL_out[N] (host or device pointers; in-place allowed).
 * Optional outputs (may be NULL): props[N*K,d], logL_props[N*K] for diagnostics.  Synchronises the stream. */
int hmx_mutate(hmx_ctx* ctx, const hmx_mutation* mut, const double* theta_res, const double* logL_res, int64_t N,
               double* theta_out, double* logL_out, double* props, double* logL_props, hmx_mutation_stats* stats);

/* Several independent populations (the chains and conditions of the reference's `--batch all`, estimate_reduced_nishioka.py:
 * 1100-1267) in ONE set of launches: group g owns n_particles[g] consecutive rows of theta_res / logL_res and is mutated with
 * muts[g] (own covariance factor, bounds, K, beta, seed, condition ids); all groups share d.  Equal, bit for bit, to calling
 * hmx_mutate once per group; stats[n_groups].  Optional props / logL_props are group-major [sum_g n_g K_g, d]. */
int hmx_mutate_groups(hmx_ctx* ctx, int32_t n_groups, const hmx_mutation* muts, const int64_t* n_particles, const double* theta_res,
                      const double* logL_res, double* theta_out, double* logL_out, double* props, double* logL_props,
                      hmx_mutation_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* HMX_H */
