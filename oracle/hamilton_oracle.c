/*
 * hamilton_oracle.c -- CPU restatement of the reference hot path (see hamilton_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: checker + reported CPU baseline.  Not linked into the product.
 *
 * The code follows the reference line by line on purpose (dense 12x12 Jacobian, LAPACK-style
 * LU, same clip / write-back / line-search semantics).  It is deliberately NOT optimised and NOT
 * restructured: the product's CUDA kernels use a different (block-structured) linear solve and
 * are checked against this file.
 */
#include "hamilton_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define CLIP_EPS 1e-12

static inline double clip01(double v) {
    /* S5:64-79 -- NaN compares false twice and stays NaN, as in the reference */
    if (v < CLIP_EPS) return CLIP_EPS;
    if (v > 1.0 - CLIP_EPS) return 1.0 - CLIP_EPS;
    return v;
}

static inline double cube(double x) { return x * x * x; }
static inline double sqr(double x) { return x * x; }
static inline double pow4(double x) { return (x * x) * (x * x); }

/* S5:513-559 */
void orc_theta_to_matrices(const double* th, double* A, double* b) {
    memset(A, 0, 25 * sizeof(double));
    memset(b, 0, 5 * sizeof(double));
#define AS(i, j, v) \
    do {            \
        A[(i) * 5 + (j)] = (v); \
        A[(j) * 5 + (i)] = (v); \
    } while (0)
    A[0] = th[0];
    AS(0, 1, th[1]);
    A[6] = th[2];
    b[0] = th[3];
    b[1] = th[4];
    A[12] = th[5];
    AS(2, 3, th[6]);
    A[18] = th[7];
    b[2] = th[8];
    b[3] = th[9];
    AS(0, 2, th[10]);
    AS(0, 3, th[11]);
    AS(1, 2, th[12]);
    AS(1, 3, th[13]);
    A[24] = th[14];
    b[4] = th[15];
    AS(0, 4, th[16]);
    AS(1, 4, th[17]);
    AS(2, 4, th[18]);
    AS(3, 4, th[19]);
#undef AS
}

/* Hill gate H(x) = x^n/(K^n+x^n), S5:89-95 */
static inline double hill_gate(double fn_conc, double K_hill, double n_hill) {
    if (fn_conc > 1e-20) {
        double xn = pow(fn_conc, n_hill);
        return xn / (pow(K_hill, n_hill) + xn);
    }
    return 0.0;
}

/* S5:43-127 */
void orc_compute_Q(double* phi, double phi0, double* psi, double gamma, const double* phi_old,
                   double phi0_old, const double* psi_old, const orc_solver_cfg* cfg,
                   const double* A, const double* b, double* Q) {
    const double dt = cfg->dt, Kp1 = cfg->Kp1, c = cfg->c_const, alpha = cfg->alpha_const;
    double phidot[5], psidot[5], inter[5];
    for (int i = 0; i < 5; ++i) {
        phi[i] = clip01(phi[i]);
        psi[i] = clip01(psi[i]);
    }
    phi0 = clip01(phi0);
    for (int i = 0; i < 5; ++i) {
        phidot[i] = (phi[i] - phi_old[i]) / dt;
        psidot[i] = (psi[i] - psi_old[i]) / dt;
    }
    double phi0dot = (phi0 - phi0_old) / dt;
    for (int i = 0; i < 5; ++i) {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s += A[i * 5 + j] * (phi[j] * psi[j]);
        inter[i] = s;
    }
    if (cfg->K_hill > 0.0) inter[4] *= hill_gate(phi[3] * psi[3], cfg->K_hill, cfg->n_hill);

    for (int i = 0; i < 5; ++i) {
        const double eta = cfg->eta[i], etaphi = cfg->eta_phi[i];
        double term1 = (Kp1 * (2.0 - 4.0 * phi[i])) / (cube(phi[i] - 1.0) * cube(phi[i]));
        double term2 = (1.0 / eta) * (gamma + (etaphi + eta * sqr(psi[i])) * phidot[i] +
                                      eta * phi[i] * psi[i] * psidot[i]);
        double term3 = (c / eta) * psi[i] * inter[i];
        Q[i] = term1 + term2 - term3;
    }
    Q[5] = gamma + (Kp1 * (2.0 - 4.0 * phi0)) / (cube(phi0 - 1.0) * cube(phi0)) + phi0dot;
    for (int i = 0; i < 5; ++i) {
        const double eta = cfg->eta[i];
        double term1 = (-2.0 * Kp1) / (sqr(psi[i] - 1.0) * cube(psi[i])) -
                       (2.0 * Kp1) / (cube(psi[i] - 1.0) * sqr(psi[i]));
        double term2 = (b[i] * alpha / eta) * psi[i];
        double term3 = phi[i] * psi[i] * phidot[i] + sqr(phi[i]) * psidot[i];
        double term4 = (c / eta) * phi[i] * inter[i];
        Q[6 + i] = term1 + term2 + term3 - term4;
    }
    double s = 0.0;
    for (int i = 0; i < 5; ++i) s += phi[i];
    Q[11] = s + phi0 - 1.0;
}

/* S5:129-289 */
void orc_compute_J(double* phi, double phi0, double* psi, double gamma, const double* phi_old,
                   const double* psi_old, const orc_solver_cfg* cfg, const double* A,
                   const double* b, double* K) {
    (void)gamma;
    const double dt = cfg->dt, Kp1 = cfg->Kp1, c = cfg->c_const, alpha = cfg->alpha_const;
    double phidot[5], psidot[5], inter[5], phi_p[5], psi_p[5];
    memset(K, 0, 144 * sizeof(double));
#define KK(r, cc) K[(r) * 12 + (cc)]
    for (int i = 0; i < 5; ++i) {
        phi[i] = clip01(phi[i]);
        psi[i] = clip01(psi[i]);
    }
    phi0 = clip01(phi0);
    for (int i = 0; i < 5; ++i) {
        phidot[i] = (phi[i] - phi_old[i]) / dt;
        psidot[i] = (psi[i] - psi_old[i]) / dt;
    }
    for (int i = 0; i < 5; ++i) {
        double s = 0.0;
        for (int j = 0; j < 5; ++j) s += A[i * 5 + j] * (phi[j] * psi[j]);
        inter[i] = s;
    }
    double hg = 1.0;
    const double I_raw_4 = inter[4];
    if (cfg->K_hill > 0.0) {
        hg = hill_gate(phi[3] * psi[3], cfg->K_hill, cfg->n_hill);
        inter[4] *= hg;
    }
    for (int i = 0; i < 5; ++i) {
        double v = phi[i];
        phi_p[i] = (Kp1 * (-4.0 + 8.0 * v)) / (cube(v) * cube(v - 1.0)) -
                   (Kp1 * (2.0 - 4.0 * v)) *
                       (3.0 / (pow4(v) * cube(v - 1.0)) + 3.0 / (cube(v) * pow4(v - 1.0)));
        v = psi[i];
        psi_p[i] = (4.0 * Kp1 * (3.0 - 5.0 * v + 5.0 * sqr(v))) / (pow4(v) * pow4(v - 1.0));
    }
    double v0 = phi0;
    double phi0_p = (Kp1 * (-4.0 + 8.0 * v0)) / (cube(v0) * cube(v0 - 1.0)) -
                    (Kp1 * (2.0 - 4.0 * v0)) *
                        (3.0 / (pow4(v0) * cube(v0 - 1.0)) + 3.0 / (cube(v0) * pow4(v0 - 1.0)));

    for (int i = 0; i < 5; ++i) {
        const double eta = cfg->eta[i], etaphi = cfg->eta_phi[i];
        for (int j = 0; j < 5; ++j) KK(i, j) = (c / eta) * psi[i] * (-A[i * 5 + j] * psi[j]);
        KK(i, i) += phi_p[i] +
                    (1.0 / eta) * ((etaphi + eta * sqr(psi[i])) / dt + eta * psi[i] * psidot[i]) -
                    (c / eta) * (psi[i] * (inter[i] + A[i * 5 + i] * psi[i]));
        for (int j = 0; j < 5; ++j) KK(i, j + 6) = (c / eta) * psi[i] * (-A[i * 5 + j] * phi[j]);
        KK(i, i + 6) += (1.0 / eta) * (2.0 * eta * psi[i] * phidot[i] + eta * phi[i] * psidot[i] +
                                       eta * phi[i] * psi[i] / dt) -
                        (c / eta) * ((inter[i] + A[i * 5 + i] * phi[i] * psi[i]) +
                                     psi[i] * (A[i * 5 + i] * phi[i]));
        KK(i, 11) = 1.0 / eta;
    }
    KK(5, 5) = phi0_p + 1.0 / dt;
    KK(5, 11) = 1.0;
    for (int i = 0; i < 5; ++i) {
        const double eta = cfg->eta[i];
        const int k = i + 6;
        for (int j = 0; j < 5; ++j)
            KK(k, j) = -(c / eta) * (A[i * 5 + j] * psi[j] * phi[i] + inter[i] * (i == j ? 1.0 : 0.0));
        KK(k, i) += (psi[i] * phidot[i] + psi[i] * phi[i] / dt + 2.0 * phi[i] * psidot[i]) -
                    (c / eta) * (A[i * 5 + i] * psi[i] * phi[i] + inter[i] +
                                 phi[i] * A[i * 5 + i] * psi[i]);
        for (int j = 0; j < 5; ++j) KK(k, j + 6) = -(c / eta) * phi[i] * A[i * 5 + j] * phi[j];
        KK(k, i + 6) += psi_p[i] + (b[i] * alpha / eta) + (phi[i] * phidot[i] + sqr(phi[i]) / dt) -
                        (c / eta) * phi[i] * A[i * 5 + i] * phi[i];
    }
    for (int j = 0; j < 6; ++j) KK(11, j) = 1.0;

    /* Hill-gate corrections, S5:249-287 */
    if (cfg->K_hill > 0.0 && hg < 1.0) {
        const double ce4 = c / cfg->eta[4];
        const double hm1 = hg - 1.0;
        for (int j = 0; j < 5; ++j) {
            KK(4, j) += hm1 * (-(ce4)*psi[4] * A[20 + j] * psi[j]);
            KK(4, j + 6) += hm1 * (-(ce4)*psi[4] * A[20 + j] * phi[j]);
        }
        KK(4, 4) += hm1 * (-(ce4)*psi[4] * A[24] * psi[4]);
        KK(4, 10) += hm1 * (-(ce4) * (A[24] * phi[4] * psi[4] + psi[4] * A[24] * phi[4]));
        for (int j = 0; j < 5; ++j) {
            KK(10, j) += hm1 * (-(ce4)*A[20 + j] * psi[j] * phi[4]);
            KK(10, j + 6) += hm1 * (-(ce4)*phi[4] * A[20 + j] * phi[j]);
        }
        KK(10, 4) += hm1 * (-(ce4) * (A[24] * psi[4] * phi[4] + phi[4] * A[24] * psi[4]));
        KK(10, 10) += hm1 * (-(ce4)*phi[4] * A[24] * phi[4]);

        const double fn = phi[3] * psi[3];
        double dh_dx = 0.0;
        if (fn > 1e-20) {
            double Kn = pow(cfg->K_hill, cfg->n_hill);
            double xn = pow(fn, cfg->n_hill);
            dh_dx = cfg->n_hill * Kn * pow(fn, cfg->n_hill - 1.0) / sqr(Kn + xn);
        }
        const double dh_dphi3 = dh_dx * psi[3];
        const double dh_dpsi3 = dh_dx * phi[3];
        KK(4, 3) += -(ce4)*psi[4] * I_raw_4 * dh_dphi3;
        KK(4, 9) += -(ce4)*psi[4] * I_raw_4 * dh_dpsi3;
        KK(10, 3) += -(ce4)*phi[4] * I_raw_4 * dh_dphi3;
        KK(10, 9) += -(ce4)*phi[4] * I_raw_4 * dh_dpsi3;
    }
#undef KK
}

/* LAPACK dgesv restated: dgetf2 (row-major storage here, same arithmetic) + dgetrs.
 * rhs is column-major [n, nrhs].  Returns 0 or (k+1) for an exact zero pivot, like INFO>0 which
 * numpy turns into LinAlgError("Singular matrix") (call sites S5:370-373, S4:1123-1126). */
int orc_dgesv(int n, double* K, int nrhs, double* rhs) {
    int piv[ORC_NSTATE];
    if (n > ORC_NSTATE) return -1;
    for (int k = 0; k < n; ++k) {
        int p = k;
        double amax = fabs(K[k * n + k]);
        for (int r = k + 1; r < n; ++r) {
            double a = fabs(K[r * n + k]);
            if (a > amax) { /* idamax: first maximum */
                amax = a;
                p = r;
            }
        }
        piv[k] = p;
        if (K[p * n + k] == 0.0) return k + 1;
        if (p != k) {
            for (int cc = 0; cc < n; ++cc) {
                double t = K[k * n + cc];
                K[k * n + cc] = K[p * n + cc];
                K[p * n + cc] = t;
            }
        }
        const double rp = 1.0 / K[k * n + k]; /* dgetf2 scales by the reciprocal when |pivot|>=sfmin */
        for (int r = k + 1; r < n; ++r) K[r * n + k] *= rp;
        for (int r = k + 1; r < n; ++r) {
            const double l = K[r * n + k];
            for (int cc = k + 1; cc < n; ++cc) K[r * n + cc] -= l * K[k * n + cc];
        }
    }
    for (int q = 0; q < nrhs; ++q) {
        double* x = rhs + (size_t)q * n;
        for (int k = 0; k < n; ++k) {
            if (piv[k] != k) {
                double t = x[k];
                x[k] = x[piv[k]];
                x[piv[k]] = t;
            }
        }
        for (int k = 0; k < n; ++k) /* L y = P b, unit lower */
            for (int r = k + 1; r < n; ++r) x[r] -= K[r * n + k] * x[k];
        for (int k = n - 1; k >= 0; --k) { /* U x = y */
            x[k] /= K[k * n + k];
            for (int r = 0; r < k; ++r) x[r] -= K[r * n + k] * x[k];
        }
    }
    return 0;
}

static int solve_with_ridge(double* K, int nrhs, double* rhs, uint32_t* status) {
    /* S5:369-373 / S4:1122-1126: try plain solve, on LinAlgError retry with K + 1e-10 I */
    double Kc[144], rc[12 * ORC_NTHETA];
    memcpy(Kc, K, sizeof(Kc));
    memcpy(rc, rhs, sizeof(double) * 12 * (size_t)nrhs);
    int info = orc_dgesv(12, Kc, nrhs, rc);
    if (info != 0) {
        if (status) *status |= ORC_ST_SINGULAR;
        memcpy(Kc, K, sizeof(Kc));
        memcpy(rc, rhs, sizeof(double) * 12 * (size_t)nrhs);
        for (int i = 0; i < 12; ++i) Kc[i * 13] += 1e-10;
        info = orc_dgesv(12, Kc, nrhs, rc);
    }
    memcpy(rhs, rc, sizeof(double) * 12 * (size_t)nrhs);
    return info;
}

static inline int any_nan12(const double* q) {
    for (int i = 0; i < 12; ++i)
        if (isnan(q[i])) return 1;
    return 0;
}
static inline double norm_inf12(const double* q) {
    double m = 0.0;
    for (int i = 0; i < 12; ++i) {
        double v = fabs(q[i]);
        if (v > m) m = v;
    }
    return m;
}

/* S5:291-432 */
int orc_newton_step(const double* g_prev, const orc_solver_cfg* cfg, const double* A,
                    const double* b, double eps_tol, double* g_new, int64_t* n_trials,
                    uint32_t* status) {
    double Q[12], K[144], delta[12], phi[5], psi[5], g_trial[12], Qt[12];
    const double* phi_old = g_prev;
    const double phi0_old = g_prev[5];
    const double* psi_old = g_prev + 6;
    int it = 0;
    memcpy(g_new, g_prev, 12 * sizeof(double));
    for (it = 0; it < cfg->max_newton_iter; ++it) {
        memcpy(phi, g_new, 5 * sizeof(double));
        memcpy(psi, g_new + 6, 5 * sizeof(double));
        const double phi0 = g_new[5], gamma = g_new[11];
        orc_compute_Q(phi, phi0, psi, gamma, phi_old, phi0_old, psi_old, cfg, A, b, Q);
        memcpy(g_new, phi, 5 * sizeof(double)); /* clipped phi/psi written back, phi0 is not */
        memcpy(g_new + 6, psi, 5 * sizeof(double));
        if (any_nan12(Q)) return it;

        double phic[5], psic[5];
        memcpy(phic, phi, sizeof(phic));
        memcpy(psic, psi, sizeof(psic));
        orc_compute_J(phic, phi0, psic, gamma, phi_old, psi_old, cfg, A, b, K);
        for (int i = 0; i < 12; ++i) delta[i] = -Q[i];
        solve_with_ridge(K, 1, delta, status);

        double norm_Q = norm_inf12(Q);
        double step = 1.0;
        int improved = 0;
        while (step > 1e-4) {
            for (int i = 0; i < 12; ++i) g_trial[i] = g_new[i] + step * delta[i];
            double phit[5], psit[5];
            memcpy(phit, g_trial, sizeof(phit));
            memcpy(psit, g_trial + 6, sizeof(psit));
            orc_compute_Q(phit, g_trial[5], psit, g_trial[11], phi_old, phi0_old, psi_old, cfg, A, b,
                          Qt);
            if (n_trials) ++*n_trials;
            if (!any_nan12(Qt)) {
                double norm_trial = norm_inf12(Qt);
                if (norm_trial < norm_Q) {
                    memcpy(g_new, g_trial, sizeof(g_trial)); /* accepted trial is NOT clipped */
                    norm_Q = norm_trial;
                    improved = 1;
                    break;
                }
            }
            step *= 0.5;
        }
        if (!improved)
            for (int i = 0; i < 12; ++i) g_new[i] = g_new[i] + delta[i];
        if (norm_Q < eps_tol) return it + 1;
    }
    if (status) *status |= ORC_ST_MAXITER;
    return it;
}

static void initial_state(const orc_solver_cfg* cfg, double* g) {
    /* S5:564-576 */
    double s = 0.0;
    for (int i = 0; i < 5; ++i) {
        g[i] = ((cfg->active_mask >> i) & 1) ? cfg->phi_init[i] : 0.0;
        s += g[i];
    }
    g[5] = 1.0 - s;
    for (int i = 0; i < 5; ++i) g[6 + i] = ((cfg->active_mask >> i) & 1) ? 0.999 : 0.0;
    g[11] = 0.0;
}

/* S5:561-581, S5:837-882 */
void orc_run_deterministic(const orc_solver_cfg* cfg, const double* theta, double* g_arr,
                           int64_t* n_iters, int64_t* n_trials, uint32_t* status) {
    double A[25], b[5], g_prev[12], g_new[12];
    orc_theta_to_matrices(theta, A, b);
    initial_state(cfg, g_prev);
    memcpy(g_arr, g_prev, sizeof(g_prev));
    for (int step = 0; step < cfg->n_steps; ++step) {
        const double tt = (step + 1) * cfg->dt;
        const double tol_t = cfg->eps * (1.0 + 10.0 * tt);
        int it = orc_newton_step(g_prev, cfg, A, b, tol_t, g_new, n_trials, status);
        if (n_iters) *n_iters += it;
        memcpy(g_prev, g_new, sizeof(g_new));
        memcpy(g_arr + (size_t)(step + 1) * 12, g_prev, sizeof(g_prev));
    }
}

/* Which (p,q) entry of A, or which b_p, a theta index fills: S5:519-557. q<0 => b_p */
static const int TH_P[ORC_NTHETA] = {0, 0, 1, 0, 1, 2, 2, 3, 2, 3, 0, 0, 1, 1, 4, 4, 0, 1, 2, 3};
static const int TH_Q[ORC_NTHETA] = {0, 1, 1, -1, -1, 2, 3, 3, -1, -1, 2, 3, 2, 3, 4, -1, 4, 4, 4, 4};

/* S4:1095-1129 for one row.  dG_k = Im Q(theta + i h e_k)/h is exact because Q is linear in A,b
 * (S5:761-832 evaluated on np.where-clipped phi, psi). */
void orc_sensitivity_row(const orc_solver_cfg* cfg, const double* theta, const double* g_new,
                         const double* g_old, uint32_t active_theta_mask, double* x1,
                         uint32_t* status) {
    double A[25], b[5], K[144], phi[5], psi[5], rhs[12 * ORC_NTHETA];
    int cols[ORC_NTHETA], ncol = 0;
    orc_theta_to_matrices(theta, A, b);
    memcpy(phi, g_new, sizeof(phi));
    memcpy(psi, g_new + 6, sizeof(psi));
    /* compute_Jacobian_matrix (S5:720-758): copies are clipped inside the numba kernel */
    orc_compute_J(phi, g_new[5], psi, g_new[11], g_old, g_old + 6, cfg, A, b, K);
    /* phi, psi are now the clipped values the complex-step Q also sees */
    double hg = 1.0;
    if (cfg->K_hill > 0.0) hg = hill_gate(phi[3] * psi[3], cfg->K_hill, cfg->n_hill);
    memset(x1, 0, sizeof(double) * 12 * ORC_NTHETA);
    for (int k = 0; k < ORC_NTHETA; ++k) {
        if (!((active_theta_mask >> k) & 1u)) continue;
        double* r = rhs + (size_t)ncol * 12;
        double dG[12];
        memset(dG, 0, sizeof(dG));
        const int p = TH_P[k], q = TH_Q[k];
        if (q < 0) {
            dG[6 + p] = (cfg->alpha_const / cfg->eta[p]) * psi[p];
        } else {
            const double hp = (p == 4) ? hg : 1.0, hq = (q == 4) ? hg : 1.0;
            const double cep = cfg->c_const / cfg->eta[p], ceq = cfg->c_const / cfg->eta[q];
            double dIp = phi[q] * psi[q] * hp; /* dI_p/dA_pq (gated when p is species 4) */
            dG[p] += -cep * psi[p] * dIp;
            dG[6 + p] += -cep * phi[p] * dIp;
            if (p != q) {
                double dIq = phi[p] * psi[p] * hq;
                dG[q] += -ceq * psi[q] * dIq;
                dG[6 + q] += -ceq * phi[q] * dIq;
            }
        }
        for (int i = 0; i < 12; ++i) r[i] = -dG[i];
        cols[ncol++] = k;
    }
    if (ncol == 0) return;
    solve_with_ridge(K, ncol, rhs, status);
    for (int cidx = 0; cidx < ncol; ++cidx)
        for (int i = 0; i < 12; ++i) x1[i * ORC_NTHETA + cols[cidx]] = rhs[(size_t)cidx * 12 + i];
}

/* EV:237-371 with rho == 0 (hard-coded by the caller, MAIN:2221) */
double orc_log_likelihood_sparse(const double* mu, const double* sig, const double* data,
                                 const double* sigma_obs, const double* weights, int n_obs) {
    double logL = 0.0;
    for (int i = 0; i < n_obs; ++i) {
        double v[5];
        for (int j = 0; j < 5; ++j) {
            const double s = sigma_obs[i * 5 + j];
            const double var_raw = sig[i * 5 + j] + s * s;
            v[j] = (!isfinite(var_raw) || var_raw <= 1e-20) ? 1e-20 : var_raw;
        }
        for (int j = 0; j < 5; ++j) {
            const double r = data[i * 5 + j] - mu[i * 5 + j];
            const double w = weights ? weights[i * 5 + j] : 1.0;
            logL -= w * 0.5 * log(2 * M_PI * v[j]);
            logL -= w * 0.5 * (r * r) / v[j];
        }
    }
    return logL;
}

/* EV:629-778 on a full 20-vector */
double orc_evaluate(const orc_solver_cfg* cfg, const orc_cond_cfg* cond, const double* theta,
                    double* obs_state, uint32_t* status, int64_t* n_iters, int64_t* n_trials) {
    const int T = cfg->n_steps;
    double* g_arr = (double*)malloc(sizeof(double) * 12 * (size_t)(T + 1));
    uint32_t st = 0;
    int64_t its = 0, trials = 0;
    double mu[ORC_MAX_OBS * 5], sig[ORC_MAX_OBS * 5];
    double logL;
    orc_run_deterministic(cfg, theta, g_arr, &its, &trials, &st);

    /* EV:657-667: every entry of x0 (and sig2) must be finite */
    int bad = 0;
    for (size_t i = 0; i < (size_t)12 * (T + 1); ++i)
        if (!isfinite(g_arr[i])) {
            bad = 1;
            break;
        }
    for (int o = 0; o < cond->n_obs && !bad; ++o) {
        const int idx = cond->idx_sparse[o];
        const double* g = g_arr + (size_t)idx * 12;
        if (obs_state) memcpy(obs_state + (size_t)o * 12, g, 12 * sizeof(double));
        double x1[12 * ORC_NTHETA];
        double sig2[12];
        double var_th[ORC_NTHETA];
        const int other0 = cond->use_absolute_volume ? 11 : 6;
        if (cond->tsm_variance) {
            /* x1[0] := x1[1] (S4:1139-1141): row 0 uses the first step's sensitivities */
            const int t = idx < 1 ? 1 : idx;
            orc_sensitivity_row(cfg, theta, g_arr + (size_t)t * 12, g_arr + (size_t)(t - 1) * 12,
                                cond->active_theta_mask, x1, &st);
            for (int k = 0; k < ORC_NTHETA; ++k) var_th[k] = sqr(cond->cov_rel * theta[k]);
            for (int s = 0; s < 12; ++s) { /* S4:324-332, 1143-1150 */
                double acc = 0.0;
                for (int k = 0; k < ORC_NTHETA; ++k)
                    if ((cond->active_theta_mask >> k) & 1u) acc += sqr(x1[s * ORC_NTHETA + k]) * var_th[k];
                sig2[s] = acc + 1e-12;
                if (!isfinite(sig2[s])) bad = 1;
            }
        } else {
            for (int s = 0; s < 12; ++s) sig2[s] = 0.0;
            memset(x1, 0, sizeof(x1));
            memset(var_th, 0, sizeof(var_th));
        }
        for (int sp = 0; sp < 5; ++sp) { /* EV:679-745 */
            const double phi = g[sp];
            const int oi = cond->use_absolute_volume ? 11 : other0 + sp;
            const double other = g[oi];
            double var = sqr(phi) * sig2[oi] + sqr(other) * sig2[sp];
            if (cond->tsm_variance) {
                double covpo = 0.0;
                for (int k = 0; k < ORC_NTHETA; ++k)
                    if ((cond->active_theta_mask >> k) & 1u)
                        covpo += x1[sp * ORC_NTHETA + k] * x1[oi * ORC_NTHETA + k] * var_th[k];
                var = var + 2.0 * phi * other * covpo;
            }
            mu[o * 5 + sp] = phi * other;
            sig[o * 5 + sp] = var;
            if (!isfinite(mu[o * 5 + sp]) || !isfinite(var)) bad = 1;
        }
    }
    if (bad) {
        st |= ORC_ST_NONFINITE;
        logL = -1e20;
    } else {
        logL = orc_log_likelihood_sparse(mu, sig, cond->data, cond->sigma_obs, cond->weights,
                                         cond->n_obs);
    }
    free(g_arr);
    if (status) *status = st;
    if (n_iters) *n_iters = its;
    if (n_trials) *n_trials = trials;
    return logL;
}

typedef struct {
    const orc_solver_cfg* cfg;
    const orc_cond_cfg* conds;
    const double* theta;
    const int32_t* cond_id;
    int64_t N;
    double* logL;
    uint32_t* status;
    int64_t* n_iters;
    int64_t* n_trials;
    int64_t next; /* shared work counter, advanced with an atomic fetch-add */
} orc_batch_job;

static void* batch_worker(void* arg) {
    orc_batch_job* job = (orc_batch_job*)arg;
    for (;;) {
        const int64_t i = __atomic_fetch_add(&job->next, 1, __ATOMIC_RELAXED);
        if (i >= job->N) break;
        uint32_t st = 0;
        int64_t its = 0, tr = 0;
        const orc_cond_cfg* cond = job->conds + (job->cond_id ? job->cond_id[i] : 0);
        job->logL[i] = orc_evaluate(job->cfg, cond, job->theta + i * ORC_NTHETA, NULL, &st, &its, &tr);
        if (job->status) job->status[i] = st;
        if (job->n_iters) job->n_iters[i] = its;
        if (job->n_trials) job->n_trials[i] = tr;
    }
    return NULL;
}

/* the reference fans evaluations out over a process pool (TM:247-312); here: a pthread pool */
void orc_evaluate_batch(const orc_solver_cfg* cfg, const orc_cond_cfg* conds, int n_cond,
                        const double* theta, const int32_t* cond_id, int64_t N, double* logL,
                        uint32_t* status, int64_t* n_iters, int64_t* n_trials, int n_threads) {
    (void)n_cond;
    orc_batch_job job = {cfg, conds, theta, cond_id, N, logL, status, n_iters, n_trials, 0};
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((int64_t)n_threads > N) n_threads = (int)(N > 0 ? N : 1);
    pthread_t tid[256];
    int started = 0;
    for (int t = 1; t < n_threads; ++t)
        if (pthread_create(&tid[started], NULL, batch_worker, &job) == 0) ++started;
    batch_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(tid[t], NULL);
}

/* ---------------------------------------------------------------- stage arithmetic ------- */

/* numpy's pairwise summation for contiguous float64 (numpy/_core/src/umath/loops_utils.h.src,
 * DOUBLE_pairwise_sum): what np.sum does at TM:621,627,689 */
static double np_pairwise_sum(const double* a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        int64_t i;
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
    }
}

static double max_of(const double* a, int64_t n) {
    double m = a[0];
    for (int64_t i = 1; i < n; ++i)
        if (a[i] > m) m = a[i];
    return m;
}

/* TM:613-659, beta_schedule == "ess" */
void orc_beta_step(const double* logL, int64_t N, double beta, double target_ess_ratio,
                   double min_db, double max_db, double* delta_beta, double* ess_at_lo) {
    double lo = 0.0, hi = 1.0 - beta;
    double ess_lo = NAN;
    const double lmax = max_of(logL, N);
    double* w = (double*)malloc(sizeof(double) * (size_t)N);
    for (int it = 0; it < 25; ++it) {
        const double mid = 0.5 * (lo + hi);
        for (int64_t i = 0; i < N; ++i) w[i] = exp(mid * (logL[i] - lmax));
        const double sum_w = np_pairwise_sum(w, N);
        double ess = 0.0;
        if (sum_w > 0) {
            for (int64_t i = 0; i < N; ++i) {
                double wn = w[i] / sum_w;
                w[i] = wn * wn;
            }
            ess = 1.0 / np_pairwise_sum(w, N);
        }
        if (ess >= target_ess_ratio * (double)N) {
            lo = mid;
            ess_lo = ess;
        } else {
            hi = mid;
        }
    }
    free(w);
    double raw = lo > min_db ? lo : min_db;
    double db = raw;
    if (max_db < db) db = max_db;
    if (1.0 - beta < db) db = 1.0 - beta;
    *delta_beta = db;
    if (ess_at_lo) *ess_at_lo = ess_lo;
}

/* TM:686-701 */
int orc_weights(const double* logL, int64_t N, double delta_beta, double* w, double* log_Sj) {
    double shift = delta_beta * logL[0];
    for (int64_t i = 0; i < N; ++i) {
        w[i] = delta_beta * logL[i];
        if (w[i] > shift) shift = w[i];
    }
    for (int64_t i = 0; i < N; ++i) w[i] = exp(w[i] - shift);
    const double w_sum = np_pairwise_sum(w, N);
    if (w_sum <= 0 || !isfinite(w_sum)) {
        for (int64_t i = 0; i < N; ++i) w[i] = 1.0 / (double)N;
        if (log_Sj) *log_Sj = NAN;
        return 1;
    }
    if (log_Sj) *log_Sj = log(w_sum) - log((double)N) + shift;
    for (int64_t i = 0; i < N; ++i) w[i] /= w_sum;
    return 0;
}

/* TM:186-198 */
void orc_resample(const double* w, int64_t N, double u, int64_t* idx) {
    double* cs = (double*)malloc(sizeof(double) * (size_t)N);
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) { /* np.cumsum: strict left-to-right */
        s += w[i];
        cs[i] = s;
    }
    cs[N - 1] = 1.0;
    for (int64_t i = 0; i < N; ++i) {
        const double pos = (u + (double)i) / (double)N;
        int64_t lo = 0, hi = N; /* searchsorted side='left': first j with cs[j] >= pos */
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            if (cs[mid] < pos)
                lo = mid + 1;
            else
                hi = mid;
        }
        if (lo > N - 1) lo = N - 1;
        idx[i] = lo;
    }
    free(cs);
}

/* np.cov(m, aweights=w) with m = theta.T, ddof = 1 (numpy/lib/_function_base_impl.py cov) */
void orc_weighted_cov(const double* theta, const double* w, int64_t N, int d, double* mean,
                      double* cov) {
    double w_sum = 0.0, w2 = 0.0;
    for (int64_t i = 0; i < N; ++i) {
        w_sum += w[i];
        w2 += w[i] * w[i];
    }
    for (int j = 0; j < d; ++j) {
        double s = 0.0;
        for (int64_t i = 0; i < N; ++i) s += theta[i * d + j] * w[i];
        mean[j] = s / w_sum;
    }
    const double fact = w_sum - 1.0 * w2 / w_sum;
    for (int a = 0; a < d; ++a)
        for (int bb = a; bb < d; ++bb) {
            double s = 0.0;
            for (int64_t i = 0; i < N; ++i)
                s += (theta[i * d + a] - mean[a]) * ((theta[i * d + bb] - mean[bb]) * w[i]);
            s *= 1.0 / fact;
            cov[a * d + bb] = s;
            cov[bb * d + a] = s;
        }
}
