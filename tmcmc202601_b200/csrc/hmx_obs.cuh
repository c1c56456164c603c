// hmx_obs.cuh -- observation-row sensitivities, TSM variance and Gaussian log-likelihood of one evaluation.
//
// Reference path: BiofilmTSM.solve_tsm (S4:1095-1150) computes x1[t,:,k] = solve(K(g_t, g_{t-1}), -dQ/dtheta_k)
// for ALL 2500 rows t by complex-step differentiation in a Python loop (98 % of the reference's wall time),
// but LogLikelihoodEvaluator.__call__ (EV:669-765) reads only the rows idx_sparse.  x1 at row t depends on
// (g_t, g_{t-1}) alone, so only those rows are computed here.
//
// dQ/dtheta_k is exact in closed form because Q is linear in A and b (S5:761-832); for theta_k = A_pq it is
// the vector  -c/eta_p h_p (phi_q psi_q) [psi_p; phi_p]  on the rows (p, 6+p)  (and symmetrically for q), i.e. it
// lies in the column space of the same rank-one factors the direction solve uses, so the sensitivity solve
// is hmx::direction() with -Q replaced by -dQ/dtheta_k.
#pragma once

#include "hmx_consts.h"
#include "hmx_lane.cuh"

namespace hmx {

// which entry of A (p,q) or of b (q = -1) theta_k fills: S5:519-557
struct ThetaMap {
    int8_t p[HMX_NTHETA], q[HMX_NTHETA];
};
#define HMX_THETA_P {0, 0, 1, 0, 1, 2, 2, 3, 2, 3, 0, 0, 1, 1, 4, 4, 0, 1, 2, 3}
#define HMX_THETA_Q {0, 1, 1, -1, -1, 2, 3, 3, -1, -1, 2, 3, 2, 3, 4, -1, 4, 4, 4, 4}

// pair = [g_t (12), phi_old(5), phi0_old, psi_old(5), pad] for t = max(idx,1); 24 doubles per observation row
constexpr int kPairStride = 24;

// TSM variance of ONE trajectory row (S4:1095-1150): x = g_t, gp = g_{t-1} (first 11 entries).  sig2[12] receives
// sum_k x1_k^2 (cov_rel theta_k)^2 + 1e-12, covp[5] the phi/psi (or phi/gamma) covariance sums of EV:723-740, x1_out
// (optional, [12,20] row-major, pre-zeroed by the caller) the sensitivities.  Returns false when a variance is non-finite.
template <bool ALPHA, int HILL>
HMX_HD bool tsm_row(const SolverConsts& K, const CondConsts& C, const ThetaMap& TM, const double* theta, const double (&A)[15],
                    const double (&b)[5], const double (&x)[12], const double (&gp)[11], double (&sig2)[12], double (&covp)[5],
                    double* x1_out) {
    bool ok = true;
    Eval e;
    residual<ALPHA, HILL>(K, x, gp, A, b, e);  // clipped copies, as compute_Jacobian_matrix does (S5:734-758)
    // every sensitivity system of this row has the same matrix: factor it once (direction_factor), then one cheap solve per
    // active parameter; a row whose elimination meets a small pivot keeps the full (pivoting) direction() per parameter
    DirFactor F;
    direction_factor<ALPHA, HILL>(K, e, A, b, F);
    for (int k = 0; k < HMX_NTHETA; ++k) {
        if (!((C.active_theta_mask >> k) & 1u)) continue;
        const int p = TM.p[k], q = TM.q[k];
        const double tk = C.cov_rel * theta[k];
        const double var_k = tk * tk;  // S4:1144
#pragma unroll
        for (int i = 0; i < 12; ++i) e.Q[i] = 0.0;
        if (q < 0) {
#pragma unroll
            for (int i = 0; i < 5; ++i)
                if (i == p) e.Q[6 + i] = K.alpha_eta[i] * e.ps[i];
        } else {
            double pq = 0.0, pp = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                pq = (i == q) ? e.p[i] : pq;
                pp = (i == p) ? e.p[i] : pp;
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const double hi = (HILL != 0 && i == 4) ? e.h : 1.0;
                double rho = 0.0;
                if (i == p) rho += K.c_eta[i] * hi * pq;
                if (i == q && p != q) rho += K.c_eta[i] * hi * pp;
                e.Q[i] = -rho * e.ps[i];
                e.Q[6 + i] = -rho * e.ph[i];
            }
        }
        double x1[12];
        if (F.healthy) direction_apply<HILL>(e, A, F, e.Q, x1);
        else direction<ALPHA, HILL>(K, e, A, b, x1);
        if (x1_out) {
#pragma unroll
            for (int i = 0; i < 12; ++i) x1_out[(size_t)i * HMX_NTHETA + k] = x1[i];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) sig2[i] = fma(x1[i] * x1[i], var_k, sig2[i]);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double xo = C.use_absolute_volume ? x1[11] : x1[6 + i];
            covp[i] = fma(x1[i] * xo, var_k, covp[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        sig2[i] += 1e-12;  // S4:324-332
        if (!(fabs(sig2[i]) <= 1.79e308)) ok = false;
    }
    return ok;
}

template <bool ALPHA, int HILL>
HMX_HD double obs_loglik(const SolverConsts& K, const CondConsts& C, const ThetaMap& TM, const double* theta,
                         const double* pairs, uint32_t solve_status, double* obs_state, uint32_t* status_out,
                         double* sig2_out = nullptr, double* x1_out = nullptr, double* pull_out = nullptr) {
    double A[15], b[5];
    theta_to_Ab(theta, A, b);
    bool bad = (solve_status & kStNonfinite) != 0u;
    double logL = 0.0;
    uint32_t st = solve_status;

    for (int o = 0; o < C.n_obs; ++o) {
        double x[12], gp[11];
        const double* pr = pairs + (size_t)o * kPairStride;
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = pr[i];
#pragma unroll
        for (int i = 0; i < 11; ++i) gp[i] = pr[12 + i];
        double state[12];  // g_arr[idx]; differs from the pair's g_t only for idx == 0 (S4:1139-1141)
#pragma unroll
        for (int i = 0; i < 12; ++i) state[i] = (C.idx[o] == 0) ? C.g0[i] : x[i];
        if (obs_state) {
#pragma unroll
            for (int i = 0; i < 12; ++i) obs_state[(size_t)o * 12 + i] = state[i];
        }

        double sig2[12], covp[5];
#pragma unroll
        for (int i = 0; i < 12; ++i) sig2[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 5; ++i) covp[i] = 0.0;

        if (x1_out) {  // _last_x1[idx] (S4:1150): column k = theta_k, zero where theta_k is not active
            for (int i = 0; i < 12 * HMX_NTHETA; ++i) x1_out[(size_t)o * 12 * HMX_NTHETA + i] = 0.0;
        }
        if (C.tsm_variance) {
            if (!tsm_row<ALPHA, HILL>(K, C, TM, theta, A, b, x, gp, sig2, covp, x1_out ? x1_out + (size_t)o * 12 * HMX_NTHETA : nullptr)) bad = true;
        }
        if (sig2_out) {  // sigma2[idx] of solve_tsm (S4:1146-1149); zeros on the np.allclose short-cut (TSMA:366-401)
#pragma unroll
            for (int i = 0; i < 12; ++i) sig2_out[(size_t)o * 12 + i] = sig2[i];
        }

        // EV:679-745 then EV:305-334 for this row
        for (int sp = 0; sp < 5; ++sp) {
            double phi = 0.0, other = 0.0, s2p = 0.0, s2o = 0.0, cv = 0.0;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i == sp) {
                    phi = state[i];
                    other = C.use_absolute_volume ? state[11] : state[6 + i];
                    s2p = sig2[i];
                    s2o = C.use_absolute_volume ? sig2[11] : sig2[6 + i];
                    cv = covp[i];
                }
            }
            const double mu = phi * other;
            double var = phi * phi * s2o + other * other * s2p;
            if (C.tsm_variance) var = var + 2.0 * phi * other * cv;
            if (!(fabs(mu) <= 1.79e308) || !(fabs(var) <= 1.79e308)) bad = true;
            const double var_raw = var + C.sigma2[o * 5 + sp];
            const double v = (!(fabs(var_raw) <= 1.79e308) || var_raw <= 1e-20) ? 1e-20 : var_raw;
            const double r = C.data[o * 5 + sp] - mu;
            const double w = C.weights[o * 5 + sp];
            logL -= w * 0.5 * log(2.0 * 3.14159265358979323846 * v);
            logL -= w * 0.5 * (r * r) / v;
            if (pull_out) pull_out[o * 5 + sp] = w * r / v;  // d logL / d mu at frozen variance (hmx_grad.cuh)
        }
    }
    if (bad) {
        st |= kStNonfinite;
        logL = -1e20;
    }
    if (status_out) *status_out = st;
    return logL;
}

}  // namespace hmx
