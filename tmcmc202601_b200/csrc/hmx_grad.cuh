// hmx_grad.cuh -- likelihood gradient d logL / d theta by forward sensitivities (what the reference's
// NUTS / HMC engine obtains from jax.value_and_grad, data_5species/main/tmcmc_nuts_engine.py:202, leapfrog :20-34).
//
// Two kinds of "sensitivity" exist on this path and they are NOT the same thing:
//   * the TSM's x1 (S4:1095-1129, hmx_obs.cuh) is LOCAL: solve(K(g_t, g_{t-1}), -dQ/dtheta_k) with the reference's
//     approximate Newton matrix K and no memory of earlier steps -- a variance proxy;
//   * the gradient needs the derivative of the trajectory itself.  For the implicit-Euler scheme Q(g_t; g_{t-1}, theta) = 0:
//         Jx S_t = -dQ/dtheta - Jp S_{t-1},   S_0 = 0,    S_t = dg_t/dtheta,
//     Jx = dQ/dg_t EXACT (direction_exact, hmx_model.cuh), Jp = dQ/dg_{t-1} = the time-derivative terms only
//     (block diagonal).  clip() is differentiated as the identity, like the reference's own matrix does.
// The recursion runs over all n_steps levels on the stored trajectory (K1b).  One thread per (evaluation, parameter):
// every thread rebuilds the step's block factorisation (cheap next to keeping 20 right-hand sides in one thread's
// registers) and carries its own 12-vector S; the 20 threads of an evaluation read the same trajectory rows (L1 hits), all
// lanes run the same T-step loop: no divergence.  ~2.3 MFLOP per (evaluation, parameter).
//
// Likelihood: the evaluator's Gaussian with the variance v FROZEN at theta (sigma_obs^2 + the TSM variance):
//     d logL / d theta_k = sum_{rows, species} (w r / v) (other_s dphi_s + phi_s dother_s),   other = psi_s (or gamma).
// The reference has no analytic gradient of this path (its JAX model is a different solver and observable), so parity
// is "unpinned" in the tier's sense: the checker is oracle/hamilton_oracle_grad.c (dense, independent), itself pinned by
// central differences of the reference-pinned oracle likelihood (tests/test_gradient.py).
#pragma once

#include "hmx_obs.cuh"

namespace hmx {

struct GradParams {
    SolverConsts K;
    ThetaMap TM;
    const CondConsts* cond;  // device array [n_cond]
    const double* theta;     // [N,20]
    const int32_t* cond_id;  // [N] or null
    long long N;
    const double* traj;      // [N, n_steps + 1, 12]
    const double* pull;      // [N, n_obs_max, 5]: w r / v (obs_kernel)
    const uint32_t* status;  // [N]: non-finite solves get a zero gradient
    double* grad;            // [N,20]
    double* dstate;          // [N, n_obs_max, 12, 20] or null: dg[idx_sparse]/dtheta
    int n_obs_max;
    int n_cond;
};

// dQ/dtheta_k at the point of `e` (Q is linear in A and b, S5:761-832): same closed form as the TSM's right-hand side
template <int HILL>
HMX_HD void dQ_dtheta(const SolverConsts& K, const Eval& e, int p, int q, double (&dG)[12]) {
#pragma unroll
    for (int i = 0; i < 12; ++i) dG[i] = 0.0;
    if (q < 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i)
            if (i == p) dG[6 + i] = K.alpha_eta[i] * e.ps[i];
        return;
    }
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
        dG[i] = -rho * e.ps[i];
        dG[6 + i] = -rho * e.ph[i];
    }
}

// One forward-sensitivity step for parameter k: S <- solve(Jx, -dQ/dtheta_k - Jp S) at (x = g_t, gp = g_{t-1}).
template <bool ALPHA, int HILL>
HMX_HD void sensitivity_step(const SolverConsts& K, const double (&x)[12], const double (&gp)[11], const double (&A)[15],
                             const double (&b)[5], int p, int q, double (&S)[12]) {
    Eval e;
    residual<ALPHA, HILL>(K, x, gp, A, b, e);
    double dG[12];
    dQ_dtheta<HILL>(K, e, p, q, dG);
    // direction_exact solves Jx delta = -e.Q: load e.Q with dQ/dtheta + Jp S  (Jp: S5:98-123, the (.. - old)/dt terms)
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = e.ph[i], u = e.ps[i];
        e.Q[i] = dG[i] - K.inv_dt * fma(fma(u, u, K.etaphi_eta[i]), S[i], e.p[i] * S[6 + i]);
        e.Q[6 + i] = dG[6 + i] - K.inv_dt * fma(e.p[i], S[i], v * v * S[6 + i]);
    }
    e.Q[5] = dG[5] - K.inv_dt * S[5];
    e.Q[11] = dG[11];
    double Sn[12];
    direction_exact<ALPHA, HILL>(K, e, A, b, Sn, clip01(x[5]));
#pragma unroll
    for (int i = 0; i < 12; ++i) S[i] = Sn[i];
}

#if defined(__CUDACC__)
template <bool ALPHA, int HILL>
__global__ void __launch_bounds__(128) grad_kernel(const __grid_constant__ GradParams P) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = tid / HMX_NTHETA;
    const int k = (int)(tid - w * HMX_NTHETA);
    if (w >= P.N) return;
    const int c = cond_index(P.cond_id, w, P.n_cond);
    const CondConsts& C = P.cond[c];
    double th[HMX_NTHETA], A[15], b[5];
#pragma unroll
    for (int j = 0; j < HMX_NTHETA; ++j) th[j] = __ldg(P.theta + w * HMX_NTHETA + j);
    theta_to_Ab(th, A, b);
    const int p = P.TM.p[k], q = P.TM.q[k];
    const int T = P.K.n_steps;
    const double* tr = P.traj + w * (long long)(T + 1) * 12;
    const double* pull = P.pull + w * (long long)P.n_obs_max * 5;
    double* ds = P.dstate ? P.dstate + w * (long long)P.n_obs_max * 12 * HMX_NTHETA : nullptr;
    double S[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) S[i] = 0.0;
    double acc = 0.0;
    int next_o = 0;
    while (next_o < C.n_obs && C.idx[next_o] == 0) {  // dg_0/dtheta = 0
        if (ds)
            for (int i = 0; i < 12; ++i) ds[((size_t)next_o * 12 + i) * HMX_NTHETA + k] = 0.0;
        ++next_o;
    }
    const bool ok = (P.status[w] & kStNonfinite) == 0u;
    double gp[11], x[12];
#pragma unroll
    for (int i = 0; i < 11; ++i) gp[i] = tr[i];
    for (int t = 1; t <= T && next_o < C.n_obs && ok; ++t) {
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = tr[(size_t)t * 12 + i];
        sensitivity_step<ALPHA, HILL>(P.K, x, gp, A, b, p, q, S);
#pragma unroll
        for (int i = 0; i < 11; ++i) gp[i] = x[i];
        while (next_o < C.n_obs && C.idx[next_o] == t) {
            const int o = next_o++;
            if (ds)
                for (int i = 0; i < 12; ++i) ds[((size_t)o * 12 + i) * HMX_NTHETA + k] = S[i];
#pragma unroll
            for (int sp = 0; sp < 5; ++sp) {
                const double phi = x[sp], other = C.use_absolute_volume ? x[11] : x[6 + sp];
                const double dphi = S[sp], dother = C.use_absolute_volume ? S[11] : S[6 + sp];
                acc = fma(pull[o * 5 + sp], fma(other, dphi, phi * dother), acc);
            }
        }
    }
    P.grad[w * HMX_NTHETA + k] = ok ? acc : 0.0;
}
#endif

}  // namespace hmx
