/*
 * hamilton_oracle_grad.c -- CPU checker for the likelihood gradient (gradients for the NUTS/HMC engine).
 *
 * TEST INFRASTRUCTURE ONLY (see hamilton_oracle.h).
 *
 * PARITY UNPINNED against the reference: the reference has no analytic gradient of this path.  Its NUTS engine
 * (data_5species/main/tmcmc_nuts_engine.py:202) calls jax.value_and_grad on a JAX re-implementation of the model
 * (hamilton_ode_jax.py; six exact-Jacobian Newton steps per time step, a different observable), and JAX is not
 * installed in the build container.  What this file states is the textbook forward-sensitivity recursion of the
 * implicit-Euler scheme the numba solver integrates, written densely and independently of the product's
 * block-structured code; it is pinned by central finite differences of the (reference-pinned) oracle likelihood
 * in tests/test_gradient.py.
 *
 *   Q(g_t; g_{t-1}, theta) = 0  (S5:43-127)   =>   Jx S_t = -dQ/dtheta - Jp S_{t-1},   S_0 = 0,
 *   Jx = dQ/dg_t  : the EXACT derivative of S5:63-127 (NOT the reference's approximate Newton matrix S5:129-289),
 *   Jp = dQ/dg_{t-1}: only the time-derivative terms, block diagonal,
 *   clip() is differentiated as the identity (the reference's own Jacobian does the same, S5:181-194).
 * Likelihood: the evaluator's Gaussian (EV:283-334) with the variance v FROZEN (sigma_obs^2 + the TSM variance at
 * theta): dlogL/dtheta_k = sum_{obs,s} w r / v * d mu_s / dtheta_k, mu_s = phi_s psi_s (or phi_s gamma).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "hamilton_oracle.h"

#define CLIP_EPS 1e-12
static inline double clip01(double v) {
    if (v < CLIP_EPS) return CLIP_EPS;
    if (v > 1.0 - CLIP_EPS) return 1.0 - CLIP_EPS;
    return v;
}
static inline double cube(double x) { return x * x * x; }
static inline double sqr(double x) { return x * x; }
static inline double pow4(double x) { return (x * x) * (x * x); }

static const int TH_P[ORC_NTHETA] = {0, 0, 1, 0, 1, 2, 2, 3, 2, 3, 0, 0, 1, 1, 4, 4, 0, 1, 2, 3};
static const int TH_Q[ORC_NTHETA] = {0, 1, 1, -1, -1, 2, 3, 3, -1, -1, 2, 3, 2, 3, 4, -1, 4, 4, 4, 4};

/* d/dv of Kp1 (2 - 4 v) / ((v - 1)^3 v^3), the phi / phi_0 barrier (S5:99, 110), by the quotient rule.  The reference's
 * Newton matrix uses (-4 + 8 v) where the derivative of the numerator is -4 (S5:185): one of the places it is inexact. */
static double dbarrier_phi(double Kp1, double v) {
    return (Kp1 * (-4.0)) / (cube(v) * cube(v - 1.0)) -
           (Kp1 * (2.0 - 4.0 * v)) * (3.0 / (pow4(v) * cube(v - 1.0)) + 3.0 / (cube(v) * pow4(v - 1.0)));
}
/* d/dv of -2 Kp1 / ((v - 1)^2 v^3) - 2 Kp1 / ((v - 1)^3 v^2), the psi barrier (S5:117-119):
 *   2 Kp1 [ 4/((v-1)^3 v^3) + 3/((v-1)^2 v^4) + 3/((v-1)^4 v^2) ]  =  2 Kp1 (10 v^2 - 10 v + 3) / (v^4 (v-1)^4)
 * (the reference's S5:189 has 4 Kp1 (3 - 5 v + 5 v^2) in the numerator, i.e. a constant 6 where the derivative has 3). */
static double dbarrier_psi(double Kp1, double v) {
    return 2.0 * Kp1 * (4.0 / (cube(v - 1.0) * cube(v)) + 3.0 / (sqr(v - 1.0) * pow4(v)) + 3.0 / (pow4(v - 1.0) * sqr(v)));
}

/* Exact dQ/dg_new (12x12 row-major) and the non-zero entries of dQ/dg_old at (g_new, g_old). */
void orc_exact_jacobians(const orc_solver_cfg* cfg, const double* theta, const double* g_new, const double* g_old, double* Jx,
                         double* Jp_diag /*[11][2]: for phi_i: d/dphi_old, d/dpsi_old; phi0: d/dphi0_old; psi_i: d/dphi_old, d/dpsi_old*/) {
    double A[25], b[5], phi[5], psi[5], p[5], I[5], phd[5], psd[5];
    const double dt = cfg->dt, Kp1 = cfg->Kp1, c = cfg->c_const, alpha = cfg->alpha_const;
    orc_theta_to_matrices(theta, A, b);
    for (int i = 0; i < 5; ++i) {
        phi[i] = clip01(g_new[i]);
        psi[i] = clip01(g_new[6 + i]);
        p[i] = phi[i] * psi[i];
        phd[i] = (phi[i] - g_old[i]) / dt;
        psd[i] = (psi[i] - g_old[6 + i]) / dt;
    }
    const double phi0 = clip01(g_new[5]);
    for (int i = 0; i < 5; ++i) {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s += A[i * 5 + j] * p[j];
        I[i] = s;
    }
    double h = 1.0, dh = 0.0;
    if (cfg->K_hill > 0.0) {
        const double x = p[3];
        if (x > 1e-20) {
            const double Kn = pow(cfg->K_hill, cfg->n_hill), xn = pow(x, cfg->n_hill);
            h = xn / (Kn + xn);
            dh = cfg->n_hill * Kn * pow(x, cfg->n_hill - 1.0) / sqr(Kn + xn);
        } else {
            h = 0.0;
        }
    }
    memset(Jx, 0, 144 * sizeof(double));
    memset(Jp_diag, 0, 22 * sizeof(double));
#define JJ(r, cc) Jx[(r) * 12 + (cc)]
    for (int i = 0; i < 5; ++i) {
        const double eta = cfg->eta[i], etaphi = cfg->eta_phi[i], ce = c / eta;
        const double hi = (i == 4) ? h : 1.0;
        /* Q_i = B(phi_i) + (1/eta)[gamma + (etaphi + eta psi^2) phidot + eta phi psi psidot] - ce psi_i hi I_i */
        JJ(i, i) += dbarrier_phi(Kp1, phi[i]) + (1.0 / eta) * ((etaphi + eta * sqr(psi[i])) / dt + eta * psi[i] * psd[i]);
        JJ(i, 6 + i) += (1.0 / eta) * (2.0 * eta * psi[i] * phd[i] + eta * phi[i] * psd[i] + eta * phi[i] * psi[i] / dt) - ce * hi * I[i];
        JJ(i, 11) = 1.0 / eta;
        /* Q_{6+i} = Bpsi(psi_i) + (b alpha/eta) psi + phi psi phidot + phi^2 psidot - ce phi_i hi I_i */
        JJ(6 + i, i) += psi[i] * phd[i] + phi[i] * psi[i] / dt + 2.0 * phi[i] * psd[i] - ce * hi * I[i];
        JJ(6 + i, 6 + i) += dbarrier_psi(Kp1, psi[i]) + b[i] * alpha / eta + phi[i] * phd[i] + sqr(phi[i]) / dt;
        for (int j = 0; j < 5; ++j) { /* d(hi I_i)/d(phi_j, psi_j) = hi A_ij (psi_j, phi_j) */
            JJ(i, j) += -ce * psi[i] * hi * A[i * 5 + j] * psi[j];
            JJ(i, 6 + j) += -ce * psi[i] * hi * A[i * 5 + j] * phi[j];
            JJ(6 + i, j) += -ce * phi[i] * hi * A[i * 5 + j] * psi[j];
            JJ(6 + i, 6 + j) += -ce * phi[i] * hi * A[i * 5 + j] * phi[j];
        }
        if (i == 4 && cfg->K_hill > 0.0) { /* dh/dx with x = phi_3 psi_3 */
            JJ(4, 3) += -ce * psi[4] * I[4] * dh * psi[3];
            JJ(4, 9) += -ce * psi[4] * I[4] * dh * phi[3];
            JJ(10, 3) += -ce * phi[4] * I[4] * dh * psi[3];
            JJ(10, 9) += -ce * phi[4] * I[4] * dh * phi[3];
        }
        Jp_diag[i * 2 + 0] = -(1.0 / eta) * (etaphi + eta * sqr(psi[i])) / dt; /* dQ_i/dphi_i,old */
        Jp_diag[i * 2 + 1] = -phi[i] * psi[i] / dt;                             /* dQ_i/dpsi_i,old */
        Jp_diag[(6 + i) * 2 + 0] = -phi[i] * psi[i] / dt;                       /* dQ_{6+i}/dphi_i,old */
        Jp_diag[(6 + i) * 2 + 1] = -sqr(phi[i]) / dt;                           /* dQ_{6+i}/dpsi_i,old */
    }
    JJ(5, 5) = dbarrier_phi(Kp1, phi0) + 1.0 / dt;
    JJ(5, 11) = 1.0;
    Jp_diag[5 * 2 + 0] = -1.0 / dt;
    for (int j = 0; j < 6; ++j) JJ(11, j) = 1.0;
#undef JJ
}

/* dQ/dtheta_k at (g_new): Q is linear in A and b (same closed form as orc_sensitivity_row) */
static void dQ_dtheta(const orc_solver_cfg* cfg, const double* g_new, int k, double* dG) {
    double phi[5], psi[5];
    for (int i = 0; i < 5; ++i) {
        phi[i] = clip01(g_new[i]);
        psi[i] = clip01(g_new[6 + i]);
    }
    double hg = 1.0;
    if (cfg->K_hill > 0.0) {
        const double x = phi[3] * psi[3];
        hg = (x > 1e-20) ? pow(x, cfg->n_hill) / (pow(cfg->K_hill, cfg->n_hill) + pow(x, cfg->n_hill)) : 0.0;
    }
    memset(dG, 0, 12 * sizeof(double));
    const int p = TH_P[k], q = TH_Q[k];
    if (q < 0) {
        dG[6 + p] = (cfg->alpha_const / cfg->eta[p]) * psi[p];
        return;
    }
    const double hp = (p == 4) ? hg : 1.0, hq = (q == 4) ? hg : 1.0;
    const double cep = cfg->c_const / cfg->eta[p], ceq = cfg->c_const / cfg->eta[q];
    dG[p] += -cep * psi[p] * (phi[q] * psi[q] * hp);
    dG[6 + p] += -cep * phi[p] * (phi[q] * psi[q] * hp);
    if (p != q) {
        dG[q] += -ceq * psi[q] * (phi[p] * psi[p] * hq);
        dG[6 + q] += -ceq * phi[q] * (phi[p] * psi[p] * hq);
    }
}

/* logL (as orc_evaluate) and its gradient with the variance frozen; dstate [n_obs,12,20] (may be NULL) receives dg[idx]/dtheta */
double orc_loglik_grad(const orc_solver_cfg* cfg, const orc_cond_cfg* cond, const double* theta, double* grad /*[20]*/,
                       double* dstate, uint32_t* status) {
    const int T = cfg->n_steps;
    double* g_arr = (double*)malloc(sizeof(double) * 12 * (size_t)(T + 1));
    double* S = (double*)calloc((size_t)12 * ORC_NTHETA, sizeof(double)); /* [k][12] */
    uint32_t st = 0;
    orc_run_deterministic(cfg, theta, g_arr, NULL, NULL, &st);
    /* the likelihood value and the pulls w r / v at the observation rows, with the evaluator's own variance */
    double obs_state[ORC_MAX_OBS * 12];
    int64_t its = 0, tr = 0;
    const double logL = orc_evaluate(cfg, cond, theta, obs_state, &st, &its, &tr);
    memset(grad, 0, ORC_NTHETA * sizeof(double));
    if (dstate) memset(dstate, 0, sizeof(double) * (size_t)cond->n_obs * 12 * ORC_NTHETA);
    /* frozen variance per observation entry: recompute as orc_evaluate does */
    double vfro[ORC_MAX_OBS * 5];
    for (int o = 0; o < cond->n_obs; ++o) {
        const int idx = cond->idx_sparse[o];
        const double* g = g_arr + (size_t)idx * 12;
        double x1[12 * ORC_NTHETA], sig2[12], var_th[ORC_NTHETA];
        memset(x1, 0, sizeof(x1));
        for (int s = 0; s < 12; ++s) sig2[s] = 0.0;
        for (int k = 0; k < ORC_NTHETA; ++k) var_th[k] = sqr(cond->cov_rel * theta[k]);
        if (cond->tsm_variance) {
            const int t = idx < 1 ? 1 : idx;
            orc_sensitivity_row(cfg, theta, g_arr + (size_t)t * 12, g_arr + (size_t)(t - 1) * 12, cond->active_theta_mask, x1, &st);
            for (int s = 0; s < 12; ++s) {
                double acc = 0.0;
                for (int k = 0; k < ORC_NTHETA; ++k)
                    if ((cond->active_theta_mask >> k) & 1u) acc += sqr(x1[s * ORC_NTHETA + k]) * var_th[k];
                sig2[s] = acc + 1e-12;
            }
        }
        for (int sp = 0; sp < 5; ++sp) {
            const int oi = cond->use_absolute_volume ? 11 : 6 + sp;
            double var = sqr(g[sp]) * sig2[oi] + sqr(g[oi]) * sig2[sp];
            if (cond->tsm_variance) {
                double cv = 0.0;
                for (int k = 0; k < ORC_NTHETA; ++k)
                    if ((cond->active_theta_mask >> k) & 1u) cv += x1[sp * ORC_NTHETA + k] * x1[oi * ORC_NTHETA + k] * var_th[k];
                var += 2.0 * g[sp] * g[oi] * cv;
            }
            const double sg = cond->sigma_obs[o * 5 + sp];
            const double vr = var + sg * sg;
            vfro[o * 5 + sp] = (!isfinite(vr) || vr <= 1e-20) ? 1e-20 : vr;
        }
    }
    double Jx[144], Jp[22], Jc[144], rhs[12 * ORC_NTHETA], dG[12];
    int next_o = 0;
    while (next_o < cond->n_obs && cond->idx_sparse[next_o] == 0) ++next_o; /* dg_0/dtheta = 0 */
    for (int t = 1; t <= T && next_o < cond->n_obs; ++t) {
        const double* gn = g_arr + (size_t)t * 12;
        const double* go = g_arr + (size_t)(t - 1) * 12;
        orc_exact_jacobians(cfg, theta, gn, go, Jx, Jp);
        for (int k = 0; k < ORC_NTHETA; ++k) {
            double* r = rhs + (size_t)k * 12;
            const double* Sk = S + (size_t)k * 12;
            dQ_dtheta(cfg, gn, k, dG);
            for (int i = 0; i < 5; ++i) {
                r[i] = -dG[i] - (Jp[i * 2] * Sk[i] + Jp[i * 2 + 1] * Sk[6 + i]);
                r[6 + i] = -dG[6 + i] - (Jp[(6 + i) * 2] * Sk[i] + Jp[(6 + i) * 2 + 1] * Sk[6 + i]);
            }
            r[5] = -dG[5] - Jp[5 * 2] * Sk[5];
            r[11] = -dG[11];
        }
        memcpy(Jc, Jx, sizeof(Jc));
        if (orc_dgesv(12, Jc, ORC_NTHETA, rhs) != 0) st |= ORC_ST_SINGULAR;
        memcpy(S, rhs, sizeof(double) * 12 * ORC_NTHETA);
        while (next_o < cond->n_obs && cond->idx_sparse[next_o] == t) {
            const int o = next_o++;
            for (int k = 0; k < ORC_NTHETA; ++k) {
                const double* Sk = S + (size_t)k * 12;
                if (dstate)
                    for (int i = 0; i < 12; ++i) dstate[((size_t)o * 12 + i) * ORC_NTHETA + k] = Sk[i];
                for (int sp = 0; sp < 5; ++sp) {
                    const int oi = cond->use_absolute_volume ? 11 : 6 + sp;
                    const double mu = gn[sp] * gn[oi];
                    const double dmu = gn[oi] * Sk[sp] + gn[sp] * Sk[oi];
                    const double r = cond->data[o * 5 + sp] - mu;
                    grad[k] += cond->weights[o * 5 + sp] * r / vfro[o * 5 + sp] * dmu;
                }
            }
        }
    }
    free(g_arr);
    free(S);
    if (status) *status = st;
    return logL;
}
