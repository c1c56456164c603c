/*
 * hamilton_oracle.h -- CPU restatement of the Tmcmc202601 Stage-1 likelihood hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this
 * library, and only as the checker or as the reported CPU baseline.  The product
 * (tmcmc202601_b200/csrc -> libhamilton.so) never links, imports or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * repository root):
 *   S5 = tmcmc/program2602/improved_5species_jit.py
 *   S4 = tmcmc/program2602/improved1207_paper_jit.py
 *   EV = data_5species/core/evaluator.py
 *   TM = data_5species/core/tmcmc.py
 *
 * Third-party arithmetic on the reference path that is NOT in the reference tree:
 *   - LAPACK dgesv (via numba's np.linalg.solve lowering; scipy 1.18.1 bundled OpenBLAS, unpinned
 *     by the reference).  Restated here as unblocked right-looking LU with partial (row) pivoting,
 *     first-maximum tie rule (idamax), then forward/back substitution -- the published dgetf2/dgetrs
 *     algorithm.
 *   - numba/LLVM fast-math code generation and libm pow/log/exp.
 *   - numpy: np.cumsum (strict left-to-right), np.searchsorted(side='left'), np.cov(aweights).
 * Parity pinning: checked against the LIVE reference (numba 0.65) in the build container by
 * tests/golden/make_golden.py, whose outputs are committed under tests/golden/.
 */
#ifndef HAMILTON_ORACLE_H
#define HAMILTON_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NSTATE 12
#define ORC_NTHETA 20
#define ORC_MAX_OBS 8

/* mirrors solver_kwargs (MAIN:1919-1928) + BiofilmNewtonSolver5S.__init__ (S5:464-511) */
typedef struct {
    double dt;
    int32_t n_steps; /* maxtimestep */
    double eps;      /* Newton residual tolerance base, 1e-6 */
    double Kp1;
    double c_const;
    double alpha_const;
    double K_hill;
    double n_hill;
    int32_t max_newton_iter;
    double eta[5];
    double eta_phi[5];
    double phi_init[5];
    int32_t active_mask; /* bit i set => species i active (S5:566-575) */
} orc_solver_cfg;

/* one experimental condition: observation constants consumed by EV:629-778 */
typedef struct {
    int32_t n_obs;
    int32_t idx_sparse[ORC_MAX_OBS];
    double data[ORC_MAX_OBS * 5];      /* row-major [n_obs,5] */
    double sigma_obs[ORC_MAX_OBS * 5]; /* pre-broadcast to [n_obs,5] */
    double weights[ORC_MAX_OBS * 5];   /* 1.0 where the reference passes weights=None */
    double cov_rel;
    uint32_t active_theta_mask; /* bit k => theta_k is in BiofilmTSM.active_idx (S4:1075) */
    int32_t use_absolute_volume;
    int32_t tsm_variance; /* 1 = EV:700-745 variance, 0 = deterministic only (sig=0) */
} orc_cond_cfg;

/* status bits shared with the product C ABI (include/hmx.h) */
#define ORC_ST_NONFINITE 1u /* -> logL = -1e20 (EV:657-667, 747-753) */
#define ORC_ST_MAXITER 2u   /* some time step exhausted max_newton_iter */
#define ORC_ST_SINGULAR 4u  /* exact zero pivot met; ridge 1e-10 retry used (S5:369-373) */

/* S5:513-559 */
void orc_theta_to_matrices(const double* theta, double* A /*[25]*/, double* b /*[5]*/);

/* S5:43-127.  phi/psi are clipped IN PLACE, phi0 only locally. */
void orc_compute_Q(double* phi, double phi0, double* psi, double gamma, const double* phi_old,
                   double phi0_old, const double* psi_old, const orc_solver_cfg* cfg,
                   const double* A, const double* b, double* Q /*[12]*/);

/* S5:129-289.  Inputs are clipped IN PLACE like the reference does on its copies. */
void orc_compute_J(double* phi, double phi0, double* psi, double gamma, const double* phi_old,
                   const double* psi_old, const orc_solver_cfg* cfg, const double* A,
                   const double* b, double* K /*[144] row-major*/);

/* dgesv restatement: solves K x = rhs in place (K destroyed). returns 0, or k+1 on exact zero pivot */
int orc_dgesv(int n, double* K, int nrhs, double* rhs /* column-major [n,nrhs] */);

/* S5:291-432 one implicit step. returns number of Newton iterations; *n_trials += line-search Q evals */
int orc_newton_step(const double* g_prev, const orc_solver_cfg* cfg, const double* A,
                    const double* b, double eps_tol, double* g_new, int64_t* n_trials,
                    uint32_t* status);

/* S5:561-581 + S5:837-882.  g_arr is [(n_steps+1),12] row-major (may be NULL when only obs rows are
 * wanted: then obs_idx/obs_pairs receive g[idx] and g[max(idx,1)-1]). */
void orc_run_deterministic(const orc_solver_cfg* cfg, const double* theta, double* g_arr,
                           int64_t* n_iters, int64_t* n_trials, uint32_t* status);

/* S4:1095-1150 restricted to one row t>=1: x1[:,k] = solve(J(g_t,g_{t-1}), -dQ/dtheta_k) */
void orc_sensitivity_row(const orc_solver_cfg* cfg, const double* theta, const double* g_new,
                         const double* g_old, uint32_t active_theta_mask,
                         double* x1 /*[12,20] row-major, inactive columns = 0*/, uint32_t* status);

/* EV:237-371, rho = 0 branch */
double orc_log_likelihood_sparse(const double* mu, const double* sig, const double* data,
                                 const double* sigma_obs, const double* weights, int n_obs);

/* EV:629-778 for a FULL 20-vector (theta_sub->theta_full mapping stays in Python) */
double orc_evaluate(const orc_solver_cfg* cfg, const orc_cond_cfg* cond, const double* theta,
                    double* obs_state /*[n_obs,12] or NULL*/, uint32_t* status, int64_t* n_iters,
                    int64_t* n_trials);

/* batch with OpenMP over evaluations; cond_id may be NULL (=0) */
void orc_evaluate_batch(const orc_solver_cfg* cfg, const orc_cond_cfg* conds, int n_cond,
                        const double* theta, const int32_t* cond_id, int64_t N, double* logL,
                        uint32_t* status, int64_t* n_iters, int64_t* n_trials, int n_threads);

/* TM:613-659 (ESS schedule) */
void orc_beta_step(const double* logL, int64_t N, double beta, double target_ess_ratio,
                   double min_db, double max_db, double* delta_beta, double* ess_at_lo);
/* TM:686-701. returns 0, or 1 when the reference falls back to uniform weights */
int orc_weights(const double* logL, int64_t N, double delta_beta, double* w, double* log_Sj);
/* TM:186-198 */
void orc_resample(const double* w, int64_t N, double u, int64_t* idx);
/* np.cov(theta.T, aweights=w) as used at TM:766 */
void orc_weighted_cov(const double* theta, const double* w, int64_t N, int d, double* mean,
                      double* cov);

/* ---- likelihood gradient (hamilton_oracle_grad.c): forward sensitivities with the EXACT Jacobian; parity unpinned against
 * the reference (which has no such code), pinned by finite differences of the pinned oracle ---- */
void orc_exact_jacobians(const orc_solver_cfg* cfg, const double* theta, const double* g_new, const double* g_old,
                         double* Jx /*[144]*/, double* Jp_diag /*[22]*/);
double orc_loglik_grad(const orc_solver_cfg* cfg, const orc_cond_cfg* cond, const double* theta, double* grad /*[20]*/,
                       double* dstate /*[n_obs,12,20] or NULL*/, uint32_t* status);

/* ---- legacy 4-species solver (S4 = improved1207_paper_jit.py, class BiofilmNewtonSolver): hamilton_oracle4.c ---- */
typedef struct {
    double dt;
    int32_t n_steps;
    double eps, Kp1, c_const, alpha_const;
    int32_t max_newton_iter; /* the numba time loop hard-codes 50 (S4:685) */
    double eta[4], eta_phi[4];
    double phi_init;      /* scalar in the 4-species solver (S4:761) */
    int32_t active_mask;  /* bit i => species i active; inactive species are locked to phi = psi = 0 (S4:92-113) */
    int32_t alpha_switch_level; /* < 0: alpha_const throughout; else time level n uses alpha_after when n >= this (S4:783-822) */
    double alpha_before, alpha_after;
} orc4_cfg;
void orc4_theta_to_matrices(const double* theta /*[14]*/, double* A /*[16]*/, double* b /*[4]*/);
void orc4_compute_Q(double* phi, double phi0, double* psi, double gamma, const double* phi_old, double phi0_old,
                    const double* psi_old, const orc4_cfg* cfg, double alpha, const double* A, const double* b, double* Q /*[10]*/);
void orc4_compute_J(double* phi, double phi0, double* psi, const double* phi_old, const double* psi_old, const orc4_cfg* cfg,
                    double alpha, const double* A, const double* b, double* K /*[100]*/);
int orc4_newton_step(const double* g_prev, const orc4_cfg* cfg, double alpha, const double* A, const double* b, double eps_tol,
                     double* g_new, int64_t* n_trials, uint32_t* status);
/* S4:610-700: g_arr [(n_steps+1), 10] row-major */
void orc4_run_deterministic(const orc4_cfg* cfg, const double* theta, double* g_arr, int64_t* n_iters, int64_t* n_trials,
                            uint32_t* status);

#ifdef __cplusplus
}
#endif
#endif
