/*
 * hamilton_oracle4.c -- CPU restatement of the reference's legacy 4-species solver
 * (S4 = tmcmc/program2602/improved1207_paper_jit.py, class BiofilmNewtonSolver).
 *
 * TEST INFRASTRUCTURE ONLY (see hamilton_oracle.h): checker for the product's hmx_legacy_trajectory_batch.
 *
 * State g (10): [phi1..phi4, phi0, psi1..psi4, gamma]; theta (14): a11,a12,a22,b1,b2, a33,a34,a44,b3,b4,
 * a13,a14,a23,a24 (S4:8-16).  Follows the numba path line by line:
 *   residual      S4:73-166    (inactive species are LOCKED: phi = psi = 0, dummy equations Q = phi, Q = psi)
 *   Jacobian      S4:168-340   (identity rows for locked species)
 *   Newton step   S4:440-608   (locks re-imposed on every trial and on the full-step fallback)
 *   time loop     S4:610-700   (50 iterations hard-coded at S4:685, tolerance eps (1 + 10 t))
 *
 * alpha_schedule (S4:783-822): the reference evaluates `alpha(t)` only in its pure-Python fallback
 * (`use_numba=False`: S4:895-913 inside compute_Q_vector); the numba integrator receives `alpha_const`
 * (S4:1024) and ignores the schedule.  Here the schedule is honoured INSIDE the numba algorithm: the step that
 * produces time level n (t = n dt) uses alpha_after when n >= switch_level, alpha_before otherwise -- what
 * `alpha(t)` returns for that step's t (S4:799-820 reduce to one switch level on the time grid).
 * switch_level < 0 disables it (alpha_const throughout, identical to the numba path).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "hamilton_oracle.h"

#define N4 4
#define NG 10
#define CLIP_EPS 1e-12

static inline double clip01(double v) {
    if (v < CLIP_EPS) return CLIP_EPS;
    if (v > 1.0 - CLIP_EPS) return 1.0 - CLIP_EPS;
    return v;
}
static inline double cube(double x) { return x * x * x; }
static inline double sqr(double x) { return x * x; }
static inline double pow4(double x) { return (x * x) * (x * x); }

/* S4:824-855 */
void orc4_theta_to_matrices(const double* th, double* A, double* b) {
    memset(A, 0, 16 * sizeof(double));
    memset(b, 0, 4 * sizeof(double));
#define AS(i, j, v)             \
    do {                        \
        A[(i) * 4 + (j)] = (v); \
        A[(j) * 4 + (i)] = (v); \
    } while (0)
    A[0] = th[0];
    AS(0, 1, th[1]);
    A[5] = th[2];
    b[0] = th[3];
    b[1] = th[4];
    A[10] = th[5];
    AS(2, 3, th[6]);
    A[15] = th[7];
    b[2] = th[8];
    b[3] = th[9];
    AS(0, 2, th[10]);
    AS(0, 3, th[11]);
    AS(1, 2, th[12]);
    AS(1, 3, th[13]);
#undef AS
}

static void clip_lock(double* phi, double* psi, double* phi0, int mask) {
    /* S4:92-113: active species clipped, inactive species locked to 0 */
    for (int i = 0; i < N4; ++i) {
        if ((mask >> i) & 1) {
            phi[i] = clip01(phi[i]);
            psi[i] = clip01(psi[i]);
        } else {
            phi[i] = 0.0;
            psi[i] = 0.0;
        }
    }
    *phi0 = clip01(*phi0);
}

/* S4:73-166 */
void orc4_compute_Q(double* phi, double phi0, double* psi, double gamma, const double* phi_old, double phi0_old,
                    const double* psi_old, const orc4_cfg* cfg, double alpha, const double* A, const double* b, double* Q) {
    const double dt = cfg->dt, Kp1 = cfg->Kp1, c = cfg->c_const;
    const int mask = cfg->active_mask;
    double phidot[N4], psidot[N4], inter[N4];
    clip_lock(phi, psi, &phi0, mask);
    for (int i = 0; i < N4; ++i) {
        phidot[i] = (phi[i] - phi_old[i]) / dt;
        psidot[i] = (psi[i] - psi_old[i]) / dt;
    }
    const double phi0dot = (phi0 - phi0_old) / dt;
    for (int i = 0; i < N4; ++i) {
        double s = 0.0;
        for (int j = 0; j < N4; ++j) s += A[i * 4 + j] * (phi[j] * psi[j]);
        inter[i] = s;
    }
    for (int i = 0; i < N4; ++i) {
        if ((mask >> i) & 1) {
            const double eta = cfg->eta[i], etaphi = cfg->eta_phi[i];
            double term1 = (Kp1 * (2.0 - 4.0 * phi[i])) / (cube(phi[i] - 1.0) * cube(phi[i]));
            double term2 = (1.0 / eta) * (gamma + (etaphi + eta * sqr(psi[i])) * phidot[i] + eta * phi[i] * psi[i] * psidot[i]);
            double term3 = (c / eta) * psi[i] * inter[i];
            Q[i] = term1 + term2 - term3;
        } else {
            Q[i] = phi[i];
        }
    }
    Q[4] = gamma + (Kp1 * (2.0 - 4.0 * phi0)) / (cube(phi0 - 1.0) * cube(phi0)) + phi0dot;
    for (int i = 0; i < N4; ++i) {
        if ((mask >> i) & 1) {
            const double eta = cfg->eta[i];
            double term1 = (-2.0 * Kp1) / (sqr(psi[i] - 1.0) * cube(psi[i])) - (2.0 * Kp1) / (cube(psi[i] - 1.0) * sqr(psi[i]));
            double term2 = (b[i] * alpha / eta) * psi[i];
            double term3 = phi[i] * psi[i] * phidot[i] + sqr(phi[i]) * psidot[i];
            double term4 = (c / eta) * phi[i] * inter[i];
            Q[5 + i] = term1 + term2 + term3 - term4;
        } else {
            Q[5 + i] = psi[i];
        }
    }
    double s = 0.0;
    for (int i = 0; i < N4; ++i) s += phi[i];
    Q[9] = s + phi0 - 1.0;
}

/* S4:168-340 */
void orc4_compute_J(double* phi, double phi0, double* psi, const double* phi_old, const double* psi_old, const orc4_cfg* cfg,
                    double alpha, const double* A, const double* b, double* K) {
    const double dt = cfg->dt, Kp1 = cfg->Kp1, c = cfg->c_const;
    const int mask = cfg->active_mask;
    double phidot[N4], psidot[N4], inter[N4], phi_p[N4], psi_p[N4];
    memset(K, 0, NG * NG * sizeof(double));
#define KK(r, cc) K[(r) * NG + (cc)]
    clip_lock(phi, psi, &phi0, mask);
    for (int i = 0; i < N4; ++i) {
        phidot[i] = (phi[i] - phi_old[i]) / dt;
        psidot[i] = (psi[i] - psi_old[i]) / dt;
    }
    for (int i = 0; i < N4; ++i) {
        double s = 0.0;
        for (int j = 0; j < N4; ++j) s += A[i * 4 + j] * (phi[j] * psi[j]);
        inter[i] = s;
    }
    for (int i = 0; i < N4; ++i) {
        phi_p[i] = psi_p[i] = 0.0;
        if ((mask >> i) & 1) {
            double v = phi[i];
            phi_p[i] = (Kp1 * (-4.0 + 8.0 * v)) / (cube(v) * cube(v - 1.0)) -
                       (Kp1 * (2.0 - 4.0 * v)) * (3.0 / (pow4(v) * cube(v - 1.0)) + 3.0 / (cube(v) * pow4(v - 1.0)));
            v = psi[i];
            psi_p[i] = (4.0 * Kp1 * (3.0 - 5.0 * v + 5.0 * sqr(v))) / (pow4(v) * pow4(v - 1.0));
        }
    }
    const double v0 = phi0;
    const double phi0_p = (Kp1 * (-4.0 + 8.0 * v0)) / (cube(v0) * cube(v0 - 1.0)) -
                          (Kp1 * (2.0 - 4.0 * v0)) * (3.0 / (pow4(v0) * cube(v0 - 1.0)) + 3.0 / (cube(v0) * pow4(v0 - 1.0)));
    for (int i = 0; i < N4; ++i) {
        if ((mask >> i) & 1) {
            const double eta = cfg->eta[i], etaphi = cfg->eta_phi[i];
            for (int j = 0; j < N4; ++j) KK(i, j) = (c / eta) * psi[i] * (-A[i * 4 + j] * psi[j]);
            KK(i, i) += phi_p[i] + (1.0 / eta) * ((etaphi + eta * sqr(psi[i])) / dt + eta * psi[i] * psidot[i]) -
                        (c / eta) * (psi[i] * (inter[i] + A[i * 4 + i] * psi[i]));
            for (int j = 0; j < N4; ++j) KK(i, j + 5) = (c / eta) * psi[i] * (-A[i * 4 + j] * phi[j]);
            KK(i, i + 5) += (1.0 / eta) * (2.0 * eta * psi[i] * phidot[i] + eta * phi[i] * psidot[i] + eta * phi[i] * psi[i] / dt) -
                            (c / eta) * ((inter[i] + A[i * 4 + i] * phi[i] * psi[i]) + psi[i] * (A[i * 4 + i] * phi[i]));
            KK(i, 9) = 1.0 / eta;
        } else {
            KK(i, i) = 1.0;
        }
    }
    KK(4, 4) = phi0_p + 1.0 / dt;
    KK(4, 9) = 1.0;
    for (int i = 0; i < N4; ++i) {
        const int k = i + 5;
        if ((mask >> i) & 1) {
            const double eta = cfg->eta[i];
            for (int j = 0; j < N4; ++j)
                KK(k, j) = -(c / eta) * (A[i * 4 + j] * psi[j] * phi[i] + inter[i] * (i == j ? 1.0 : 0.0));
            KK(k, i) += (psi[i] * phidot[i] + psi[i] * phi[i] / dt + 2.0 * phi[i] * psidot[i]) -
                        (c / eta) * (A[i * 4 + i] * psi[i] * phi[i] + inter[i] + phi[i] * A[i * 4 + i] * psi[i]);
            for (int j = 0; j < N4; ++j) KK(k, j + 5) = -(c / eta) * phi[i] * A[i * 4 + j] * phi[j];
            KK(k, i + 5) += psi_p[i] + (b[i] * alpha / eta) + (phi[i] * phidot[i] + sqr(phi[i]) / dt) -
                            (c / eta) * phi[i] * A[i * 4 + i] * phi[i];
        } else {
            KK(k, k) = 1.0;
        }
    }
    for (int j = 0; j < 5; ++j) KK(9, j) = 1.0;
#undef KK
}

static int solve_with_ridge10(double* K, double* rhs, uint32_t* status) {
    /* S4:518-522 */
    double Kc[NG * NG], rc[NG];
    memcpy(Kc, K, sizeof(Kc));
    memcpy(rc, rhs, sizeof(rc));
    int info = orc_dgesv(NG, Kc, 1, rc);
    if (info != 0) {
        if (status) *status |= ORC_ST_SINGULAR;
        memcpy(Kc, K, sizeof(Kc));
        memcpy(rc, rhs, sizeof(rc));
        for (int i = 0; i < NG; ++i) Kc[i * (NG + 1)] += 1e-10;
        info = orc_dgesv(NG, Kc, 1, rc);
    }
    memcpy(rhs, rc, sizeof(rc));
    return info;
}

static inline int any_nan(const double* q) {
    for (int i = 0; i < NG; ++i)
        if (isnan(q[i])) return 1;
    return 0;
}
static inline double norm_inf(const double* q) {
    double m = 0.0;
    for (int i = 0; i < NG; ++i) {
        double v = fabs(q[i]);
        if (v > m) m = v;
    }
    return m;
}
static inline void force_lock(double* g, int mask) {
    for (int i = 0; i < N4; ++i)
        if (!((mask >> i) & 1)) {
            g[i] = 0.0;
            g[i + 5] = 0.0;
        }
}

/* S4:440-608 */
int orc4_newton_step(const double* g_prev, const orc4_cfg* cfg, double alpha, const double* A, const double* b, double eps_tol,
                     double* g_new, int64_t* n_trials, uint32_t* status) {
    double Q[NG], K[NG * NG], delta[NG], phi[N4], psi[N4], g_trial[NG], Qt[NG];
    const double* phi_old = g_prev;
    const double phi0_old = g_prev[4];
    const double* psi_old = g_prev + 5;
    const int mask = cfg->active_mask;
    int it;
    memcpy(g_new, g_prev, NG * sizeof(double));
    force_lock(g_new, mask);
    for (it = 0; it < cfg->max_newton_iter; ++it) {
        memcpy(phi, g_new, sizeof(phi));
        memcpy(psi, g_new + 5, sizeof(psi));
        const double phi0 = g_new[4], gamma = g_new[9];
        orc4_compute_Q(phi, phi0, psi, gamma, phi_old, phi0_old, psi_old, cfg, alpha, A, b, Q);
        memcpy(g_new, phi, sizeof(phi)); /* clipped / locked phi, psi written back; phi0 (passed by value) is not */
        memcpy(g_new + 5, psi, sizeof(psi));
        if (any_nan(Q)) return it;
        double phic[N4], psic[N4];
        memcpy(phic, phi, sizeof(phic));
        memcpy(psic, psi, sizeof(psic));
        orc4_compute_J(phic, phi0, psic, phi_old, psi_old, cfg, alpha, A, b, K);
        for (int i = 0; i < NG; ++i) delta[i] = -Q[i];
        solve_with_ridge10(K, delta, status);
        double norm_Q = norm_inf(Q);
        double step = 1.0;
        int improved = 0;
        while (step > 1e-4) {
            for (int i = 0; i < NG; ++i) g_trial[i] = g_new[i] + step * delta[i];
            force_lock(g_trial, mask);
            double phit[N4], psit[N4];
            memcpy(phit, g_trial, sizeof(phit));
            memcpy(psit, g_trial + 5, sizeof(psit));
            orc4_compute_Q(phit, g_trial[4], psit, g_trial[9], phi_old, phi0_old, psi_old, cfg, alpha, A, b, Qt);
            if (n_trials) ++*n_trials;
            if (!any_nan(Qt)) {
                const double norm_trial = norm_inf(Qt);
                if (norm_trial < norm_Q) {
                    memcpy(g_new, g_trial, sizeof(g_trial));
                    norm_Q = norm_trial;
                    improved = 1;
                    break;
                }
            }
            step *= 0.5;
        }
        if (!improved) {
            for (int i = 0; i < NG; ++i) g_new[i] = g_new[i] + delta[i];
            force_lock(g_new, mask);
        }
        if (norm_Q < eps_tol) return it + 1;
    }
    if (status) *status |= ORC_ST_MAXITER;
    return it;
}

/* S4:610-700 (+ the alpha schedule, see the header comment) */
void orc4_run_deterministic(const orc4_cfg* cfg, const double* theta, double* g_arr, int64_t* n_iters, int64_t* n_trials,
                            uint32_t* status) {
    double A[16], b[4], g_prev[NG], g_new[NG];
    orc4_theta_to_matrices(theta, A, b);
    double s = 0.0;
    for (int i = 0; i < N4; ++i) {
        g_prev[i] = ((cfg->active_mask >> i) & 1) ? cfg->phi_init : 0.0;
        s += g_prev[i];
    }
    g_prev[4] = 1.0 - s;
    for (int i = 0; i < N4; ++i) g_prev[5 + i] = ((cfg->active_mask >> i) & 1) ? 0.999 : 0.0;
    g_prev[9] = 0.0;
    memcpy(g_arr, g_prev, sizeof(g_prev));
    for (int step = 0; step < cfg->n_steps; ++step) {
        const double tt = (step + 1) * cfg->dt;
        const double tol_t = cfg->eps * (1.0 + 10.0 * tt);
        double alpha = cfg->alpha_const;
        if (cfg->alpha_switch_level >= 0) alpha = (step + 1 >= cfg->alpha_switch_level) ? cfg->alpha_after : cfg->alpha_before;
        int it = orc4_newton_step(g_prev, cfg, alpha, A, b, tol_t, g_new, n_trials, status);
        if (n_iters) *n_iters += it;
        memcpy(g_prev, g_new, sizeof(g_new));
        memcpy(g_arr + (size_t)(step + 1) * NG, g_prev, sizeof(g_prev));
    }
}
