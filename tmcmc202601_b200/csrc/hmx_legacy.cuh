// hmx_legacy.cuh -- the reference's LEGACY 4-species solver (S4 = tmcmc/program2602/improved1207_paper_jit.py,
// class BiofilmNewtonSolver, S4:715-1053) on the GPU: SURVEY 8(f)4.
//
// Same model family as the 5-species solver without the Hill gate, written for a compile-time species count NS:
//   state g (2 NS + 2) = [phi_1..NS, phi_0, psi_1..NS, gamma]          (S4:8-13)
//   residual   S4:73-166,  Jacobian S4:168-340 -- never formed: the block structure of hmx_model.cuh
//              (blockdiag(D_i) + P C Q^T, phi_0 row and constraint row / gamma column eliminated analytically)
//   Newton     S4:440-608  (damped quasi-Newton, backtracking on ||Q||_inf; same semantics as S5:291-432)
//   time loop  S4:610-700  (tolerance eps (1 + 10 t_{n+1}), 50 iterations)
// What the 4-species solver adds:
//   * LOCKED species (S4:92-113, 233-340, 458-462, 566-570, 597-603): an inactive species is held at phi = psi = 0, its two
//     equations are the dummies Q = phi, Q = psi with identity Jacobian rows, and its columns vanish because every
//     A-derived entry carries a factor phi_j or psi_j = 0.  In the block structure this is D_i = I, zero rank-one
//     factors and zero right-hand sides: the species drops out of the NS x NS system.
//   * alpha_schedule (S4:783-822): piecewise-constant antibiotic strength.  The reference honours it only in its
//     pure-Python fallback (compute_Q_vector calls self.alpha(t), S4:895-913); its numba integrator passes alpha_const
//     (S4:1024).  Here it is a per-step lookup inside the numba algorithm: the step producing time level n uses
//     alpha_after when n >= switch_level (the host resolves switch_time / switch_frac / switch_step to that level).
//
// The legacy pipeline is served for completeness, not speed: one thread per solve with the reference's nested
// loops (time step -> Newton iteration -> line-search trial); the per-lane state machine of hmx_lane.cuh is what the
// production 5-species path uses.  NS = 5 is instantiated as well (no Hill gate) so the tests can check this
// generic arithmetic against the production kernel.
#pragma once

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#endif

#include "hmx_lane.cuh"

namespace hmx {

constexpr int kLegacyMaxNS = 5;

struct LegacyConsts {
    double dt, inv_dt, eps, Kp1, c;
    double alpha_const, alpha_before, alpha_after;
    double eta[kLegacyMaxNS], eta_phi[kLegacyMaxNS], inv_eta[kLegacyMaxNS], c_eta[kLegacyMaxNS], etaphi_eta[kLegacyMaxNS];
    double g0[2 * kLegacyMaxNS + 2];
    int32_t n_steps, max_iter, switch_level;  // switch_level < 0: alpha_const throughout
    uint32_t active_mask;
};

template <int NS>
HMX_HD constexpr int lpk(int i, int j) {  // packed symmetric NS x NS
    return (i <= j) ? (i * NS - (i * (i - 1)) / 2 + (j - i)) : (j * NS - (j * (j - 1)) / 2 + (i - j));
}

// theta -> packed A, b.  NS = 4: S4:824-855 (14 parameters); NS = 5: S5:513-559 (20 parameters)
template <int NS>
HMX_HD void legacy_theta_to_Ab(const double* th, double (&A)[NS * (NS + 1) / 2], double (&b)[NS]) {
#pragma unroll
    for (int k = 0; k < NS * (NS + 1) / 2; ++k) A[k] = 0.0;
    A[lpk<NS>(0, 0)] = th[0];
    A[lpk<NS>(0, 1)] = th[1];
    A[lpk<NS>(1, 1)] = th[2];
    b[0] = th[3];
    b[1] = th[4];
    A[lpk<NS>(2, 2)] = th[5];
    A[lpk<NS>(2, 3)] = th[6];
    A[lpk<NS>(3, 3)] = th[7];
    b[2] = th[8];
    b[3] = th[9];
    A[lpk<NS>(0, 2)] = th[10];
    A[lpk<NS>(0, 3)] = th[11];
    A[lpk<NS>(1, 2)] = th[12];
    A[lpk<NS>(1, 3)] = th[13];
    if (NS == 5) {
        A[lpk<NS>(4, 4)] = th[14];
        b[NS - 1] = th[15];
        A[lpk<NS>(0, 4)] = th[16];
        A[lpk<NS>(1, 4)] = th[17];
        A[lpk<NS>(2, 4)] = th[18];
        A[lpk<NS>(3, 4)] = th[19];
    }
}

template <int NS>
struct LegacyEval {
    double ph[NS], ps[NS], p[NS], I[NS], phd[NS], psd[NS], phip[NS], psip[NS];
    double K55;
    double Q[2 * NS + 2];
    double norm;
    bool nan;
};

// Residual at x (S4:73-166).  gp = [phi_old(NS), phi0_old, psi_old(NS)].
template <int NS>
HMX_HD void legacy_residual(const LegacyConsts& K, double alpha, const double (&x)[2 * NS + 2], const double (&gp)[2 * NS + 1],
                            const double (&A)[NS * (NS + 1) / 2], const double (&b)[NS], LegacyEval<NS>& e) {
    const double gam = x[2 * NS + 1];
    const double ph0 = clip01(x[NS]);
    const double m2K = -2.0 * K.Kp1;
    double fph[NS], fps[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const bool act = (K.active_mask >> i) & 1u;
        const double v = act ? clip01(x[i]) : 0.0, u = act ? clip01(x[NS + 1 + i]) : 0.0;  // S4:92-113
        e.ph[i] = v;
        e.ps[i] = u;
        e.p[i] = v * u;
        e.phd[i] = (v - gp[i]) * K.inv_dt;
        e.psd[i] = (u - gp[NS + 1 + i]) * K.inv_dt;
        if (act) {
            const double wv = fma(v, v, -v), wu = fma(u, u, -u);
            const double rv = rcp(wv), ru = rcp(wu);
            const double sv = fma(2.0, v, -1.0), su = fma(2.0, u, -1.0);
            fph[i] = (m2K * sv) * (rv * rv * rv);
            fps[i] = (m2K * su) * (ru * ru * ru);
            e.phip[i] = -fph[i] * fma(3.0 * sv, rv, 2.0);
            const double ru2 = ru * ru;
            e.psip[i] = (4.0 * K.Kp1) * fma(5.0, wu, 3.0) * (ru2 * ru2);
        } else {
            fph[i] = fps[i] = e.phip[i] = e.psip[i] = 0.0;
        }
    }
    double f0;
    {
        const double r0 = rcp(fma(ph0, ph0, -ph0));
        const double s0 = fma(2.0, ph0, -1.0);
        f0 = (m2K * s0) * (r0 * r0 * r0);
        e.K55 = -f0 * fma(3.0 * s0, r0, 2.0) + K.inv_dt;
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double s = A[lpk<NS>(i, 0)] * e.p[0];
#pragma unroll
        for (int j = 1; j < NS; ++j) s = fma(A[lpk<NS>(i, j)], e.p[j], s);
        e.I[i] = s;
    }
    double nrm = 0.0, asum = 0.0, sphi = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const bool act = (K.active_mask >> i) & 1u;
        const double v = e.ph[i], u = e.ps[i];
        double qi, qs;
        if (act) {
            const double t2 = fma(K.inv_eta[i], gam, fma(fma(u, u, K.etaphi_eta[i]), e.phd[i], e.p[i] * e.psd[i]));
            qi = fph[i] + t2 - K.c_eta[i] * u * e.I[i];
            qs = fps[i] + fma(e.p[i], e.phd[i], v * v * e.psd[i]) - K.c_eta[i] * v * e.I[i] + (b[i] * alpha * K.inv_eta[i]) * u;
        } else {
            qi = v;  // dummy equations of a locked species: Q = phi, Q = psi (S4:141-143, 160-162)
            qs = u;
        }
        e.Q[i] = qi;
        e.Q[NS + 1 + i] = qs;
        nrm = max_abs(nrm, max_abs(qi, qs));
        asum += fabs(qi) + fabs(qs);
        sphi = (i == 0) ? v : sphi + v;
    }
    e.Q[NS] = gam + f0 + (ph0 - gp[NS]) * K.inv_dt;
    e.Q[2 * NS + 1] = sphi + ph0 - 1.0;
    nrm = max_abs(nrm, max_abs(e.Q[NS], e.Q[2 * NS + 1]));
    asum += fabs(e.Q[NS]) + fabs(e.Q[2 * NS + 1]);
    e.norm = nrm;
    e.nan = (asum != asum);
}

// N x N solve with two right-hand sides, partial pivoting (runtime loops: the legacy path is not tuned)
template <int N>
HMX_HD void legacy_solve_pivoted(double (&S)[N][N + 2]) {
    for (int k = 0; k < N; ++k) {
        int p = k;
        double best = fabs(S[k][k]);
        for (int r = k + 1; r < N; ++r) {
            const double a = fabs(S[r][k]);
            if (a > best) {
                best = a;
                p = r;
            }
        }
        if (p != k)
            for (int cc = 0; cc < N + 2; ++cc) {
                const double t = S[k][cc];
                S[k][cc] = S[p][cc];
                S[p][cc] = t;
            }
        const double rpk = 1.0 / S[k][k];
        for (int r = k + 1; r < N; ++r) {
            const double l = S[r][k] * rpk;
            for (int cc = k + 1; cc < N + 2; ++cc) S[r][cc] = fma(-l, S[k][cc], S[r][cc]);
        }
    }
    for (int k = N - 1; k >= 0; --k) {
        double s0 = S[k][N], s1 = S[k][N + 1];
        for (int cc = k + 1; cc < N; ++cc) {
            s0 = fma(-S[k][cc], S[cc][N], s0);
            s1 = fma(-S[k][cc], S[cc][N + 1], s1);
        }
        S[k][N] = s0 / S[k][k];
        S[k][N + 1] = s1 / S[k][k];
    }
}

// delta = solve(K, -Q) with K the reference's Jacobian (S4:168-340) through the block structure of hmx_model.cuh
template <int NS>
HMX_HD void legacy_direction(const LegacyConsts& K, double alpha, const LegacyEval<NS>& e, const double (&A)[NS * (NS + 1) / 2],
                             const double (&b)[NS], double (&delta)[2 * NS + 2]) {
    double S[NS][NS + 2];
    double a0[NS], a1[NS], b0[NS], b1[NS], dd0[NS], dd1[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        const bool act = (K.active_mask >> i) & 1u;
        if (act) {
            const double v = e.ph[i], u = e.ps[i], pd = e.phd[i], sd = e.psd[i];
            const double ce = K.c_eta[i];
            const double aii = A[lpk<NS>(i, i)];
            const double ap = aii * e.p[i];
            const double d00 = e.phip[i] + fma(fma(u, u, K.etaphi_eta[i]), K.inv_dt, u * sd) - ce * u * fma(2.0 * aii, u, e.I[i]);
            const double d01 = fma(2.0 * u, pd, fma(v, sd, e.p[i] * K.inv_dt)) - ce * fma(3.0, ap, e.I[i]);
            const double d10 = fma(u, pd, fma(e.p[i], K.inv_dt, 2.0 * v * sd)) - ce * fma(3.0, ap, 2.0 * e.I[i]);
            const double d11 = e.psip[i] + fma(v, pd, v * v * K.inv_dt) - 2.0 * ce * aii * v * v + b[i] * alpha * K.inv_eta[i];
            const double t = d01 * d10;
            const double det = fma(d00, d11, -t) + fma(-d01, d10, t);
            const double rdet = rcp(det);
            const double r0 = -e.Q[i], r1 = -e.Q[NS + 1 + i];
            a0[i] = fma(d11, r0, -d01 * r1) * rdet;
            a1[i] = fma(d00, r1, -d10 * r0) * rdet;
            const double rdi = rdet * K.inv_eta[i];
            b0[i] = d11 * rdi;
            b1[i] = -d10 * rdi;
            dd0[i] = fma(d11, u, -d01 * v) * rdet;
            dd1[i] = fma(d00, v, -d10 * u) * rdet;
            const double gi = fma(u, dd0[i], v * dd1[i]) * (-ce);
#pragma unroll
            for (int j = 0; j < NS; ++j) S[i][j] = (i == j) ? 1.0 : gi * A[lpk<NS>(i, j)];
            S[i][NS] = fma(u, a0[i], v * a1[i]);
            S[i][NS + 1] = fma(u, b0[i], v * b1[i]);
        } else {
            // identity rows (S4:274-276, 336-338), no gamma column, zero rank-one factors
            a0[i] = -e.Q[i];
            a1[i] = -e.Q[NS + 1 + i];
            b0[i] = b1[i] = dd0[i] = dd1[i] = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) S[i][j] = (i == j) ? 1.0 : 0.0;
            S[i][NS] = 0.0;
            S[i][NS + 1] = 0.0;
        }
    }
    legacy_solve_pivoted<NS>(S);
    const double rK55 = rcp(e.K55);
    double y10[NS], y11[NS], y20[NS], y21[NS];
    double sumY1 = 0.0, sumY2 = 0.0;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double t1 = 0.0, t2 = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (j != i) {
                t1 = fma(A[lpk<NS>(i, j)], S[j][NS], t1);
                t2 = fma(A[lpk<NS>(i, j)], S[j][NS + 1], t2);
            }
        }
        t1 *= -K.c_eta[i];
        t2 *= -K.c_eta[i];
        y10[i] = fma(-dd0[i], t1, a0[i]);
        y11[i] = fma(-dd1[i], t1, a1[i]);
        y20[i] = fma(-dd0[i], t2, b0[i]);
        y21[i] = fma(-dd1[i], t2, b1[i]);
        sumY1 += y10[i];
        sumY2 += y20[i];
    }
    const double srhs = fma(e.Q[NS], rK55, -e.Q[2 * NS + 1]);
    const double dgam = (sumY1 - srhs) * rcp(sumY2 + rK55);
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        delta[i] = fma(-dgam, y20[i], y10[i]);
        delta[NS + 1 + i] = fma(-dgam, y21[i], y11[i]);
    }
    delta[NS] = (-e.Q[NS] - dgam) * rK55;
    delta[2 * NS + 1] = dgam;
}

struct LegacyParams {
    LegacyConsts K;
    const double* theta;  // [N, NTH]
    long long N;
    double* traj;         // [N, n_rows, 2 NS + 2]
    const int32_t* row_map;  // [n_steps + 1] -> output row or -1; null = every level
    int n_rows;
    uint32_t* status;
    uint32_t* iters;
};

template <int NS>
HMX_HD void legacy_lock(const LegacyConsts& K, double (&g)[2 * NS + 2]) {
#pragma unroll
    for (int i = 0; i < NS; ++i)
        if (!((K.active_mask >> i) & 1u)) {
            g[i] = 0.0;
            g[NS + 1 + i] = 0.0;
        }
}

// One solve, the reference's loop nest (S4:440-700).  store(level, g) receives every time level.
template <int NS, class Store>
HMX_HD void legacy_solve(const LegacyConsts& K, const double* theta, uint32_t& status, uint32_t& iters, Store&& store) {
    constexpr int NG = 2 * NS + 2;
    double A[NS * (NS + 1) / 2], b[NS];
    legacy_theta_to_Ab<NS>(theta, A, b);
    double g[NG], gp[NG - 1], d[NG], x[NG];
#pragma unroll
    for (int i = 0; i < NG; ++i) g[i] = K.g0[i];
    store(0, g);
    status = 0u;
    iters = 0u;
    for (int step = 0; step < K.n_steps; ++step) {
        const double tol = K.eps * (1.0 + 10.0 * ((double)(step + 1) * K.dt));
        const double alpha = K.switch_level < 0 ? K.alpha_const : ((step + 1 >= K.switch_level) ? K.alpha_after : K.alpha_before);
#pragma unroll
        for (int i = 0; i < NG - 1; ++i) gp[i] = g[i];
        legacy_lock<NS>(K, g);  // S4:458-462
        LegacyEval<NS> e;
        bool have_e = false, converged = false, nan_break = false;
        double normQ = 0.0;
        for (int it = 0; it < K.max_iter; ++it) {
            if (!have_e) legacy_residual<NS>(K, alpha, g, gp, A, b, e);
#pragma unroll
            for (int i = 0; i < NS; ++i) {  // clipped / locked phi, psi written back (S4:489-493)
                g[i] = e.ph[i];
                g[NS + 1 + i] = e.ps[i];
            }
            have_e = false;
            if (e.nan) {
                nan_break = true;
                break;
            }
            legacy_direction<NS>(K, alpha, e, A, b, d);
            ++iters;
            normQ = e.norm;
            double sl = 1.0;
            bool improved = false;
            while (sl > 1e-4) {
#pragma unroll
                for (int i = 0; i < NG; ++i) x[i] = g[i] + sl * d[i];
                legacy_lock<NS>(K, x);
                LegacyEval<NS> et;
                legacy_residual<NS>(K, alpha, x, gp, A, b, et);
                if (!et.nan && et.norm < normQ) {
#pragma unroll
                    for (int i = 0; i < NG; ++i) g[i] = x[i];  // the accepted trial is NOT clipped
                    normQ = et.norm;
                    e = et;  // the next iteration would recompute exactly this residual
                    have_e = true;
                    improved = true;
                    break;
                }
                sl *= 0.5;
            }
            if (!improved) {
                double chk = 0.0;
#pragma unroll
                for (int i = 0; i < NG; ++i) {
                    chk = fma(d[i], 0.0, chk);
                    g[i] = g[i] + d[i];
                }
                if (chk != 0.0) status |= kStSingular;
                legacy_lock<NS>(K, g);
            }
            if (normQ < tol) {
                converged = true;
                break;
            }
        }
        if (!converged && !nan_break) status |= kStMaxIter;
        double chk = 0.0;
#pragma unroll
        for (int i = 0; i < NG; ++i) chk = fma(g[i], 0.0, chk);
        if (chk != 0.0) status |= kStNonfinite;
        store(step + 1, g);
    }
}

#if defined(__CUDACC__)
template <int NS>
__global__ void __launch_bounds__(128) legacy_ode_kernel(const __grid_constant__ LegacyParams P) {
    constexpr int NG = 2 * NS + 2;
    constexpr int NTH = (NS == 4) ? 14 : 20;
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.N) return;
    double th[NTH];
#pragma unroll
    for (int k = 0; k < NTH; ++k) th[k] = __ldg(P.theta + w * NTH + k);
    uint32_t st = 0u, its = 0u;
    auto store = [&](int level, const double (&g)[NG]) {
        const int r = P.row_map ? P.row_map[level] : level;
        if (r >= 0) {
            double* q = P.traj + (w * (long long)P.n_rows + r) * NG;
#pragma unroll
            for (int i = 0; i < NG; ++i) q[i] = g[i];
        }
    };
    legacy_solve<NS>(P.K, th, st, its, store);
    if (P.status) P.status[w] = st;
    if (P.iters) P.iters[w] = its;
}
#endif

}  // namespace hmx
