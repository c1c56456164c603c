// hmx_dense.cuh -- the reference's loop nest with a DENSE linear solve, for the evaluations the production kernel could
// not handle (K1r, "resolve" kernel).
//
// The production kernel (hmx_lane.cuh) solves K delta = -Q through the block structure of K (hmx_model.cuh), which needs
// every 2x2 species block D_i, the 5x5 Schur system and two scalar closures to be invertible.  K itself can be perfectly
// regular when one of those is not, and where K IS singular the reference has a defined behaviour that the structured solve
// cannot express:   try: delta = solve(K, -Q)   except LinAlgError: delta = solve(K + 1e-10 I, -Q)        (S5:369-373).
// An evaluation whose structured solve ever produced a non-finite direction is flagged HMX_ST_SINGULAR by K1 and is then
// solved AGAIN, from its initial state, by this kernel: the reference's nested loops (S5:837-882 -> S5:291-432) with the
// dense 12x12 Newton matrix, LAPACK-style LU with partial pivoting (first maximum), and the ridge retry on an exact zero
// pivot.  One thread per flagged evaluation, everything in local memory: correctness path, not a throughput path -- on the
// production workloads no evaluation is ever flagged (0 of 4 x 10^7 in the throughput sweep).
#pragma once

#include "hmx_obs.cuh"

namespace hmx {

// The reference's Newton matrix K (S5:129-289) assembled densely from the quantities of one residual evaluation.
template <bool ALPHA, int HILL, class AT>
HMX_HD void dense_newton_matrix(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], double* Kd /*[144] row-major*/) {
    for (int i = 0; i < 144; ++i) Kd[i] = 0.0;
    double c43x = 0.0;
    if (HILL != 0) {
        if (e.h < 1.0) c43x = -K.c_eta[4] * e.Iraw4 * e.dh_dx;  // S5:273-287
    }
    for (int i = 0; i < 5; ++i) {
        const double hi = (HILL != 0 && i == 4) ? e.h : 1.0;
        const double srow = -K.c_eta[i] * hi;
        for (int j = 0; j < 5; ++j) {
            if (j == i) continue;
            double cij = srow * A[apk(i, j)];
            if (HILL != 0 && i == 4 && j == 3) cij += c43x;
            Kd[i * 12 + j] = cij * e.ps[i] * e.ps[j];
            Kd[i * 12 + 6 + j] = cij * e.ps[i] * e.ph[j];
            Kd[(6 + i) * 12 + j] = cij * e.ph[i] * e.ps[j];
            Kd[(6 + i) * 12 + 6 + j] = cij * e.ph[i] * e.ph[j];
        }
        double d00, d01, d10, d11;
        block2x2<ALPHA, HILL, false>(K, e, A, b, i, d00, d01, d10, d11);
        Kd[i * 12 + i] = d00;
        Kd[i * 12 + 6 + i] = d01;
        Kd[(6 + i) * 12 + i] = d10;
        Kd[(6 + i) * 12 + 6 + i] = d11;
        Kd[i * 12 + 11] = K.inv_eta[i];  // S5:220
    }
    Kd[5 * 12 + 5] = e.K55;  // S5:222-223
    Kd[5 * 12 + 11] = 1.0;
    for (int j = 0; j < 6; ++j) Kd[11 * 12 + j] = 1.0;  // S5:247
}

// dgesv for one right-hand side: right-looking LU with partial (row) pivoting, first maximum; returns 0, or k + 1 when the
// pivot of column k is exactly zero (LAPACK INFO > 0 -> numpy raises LinAlgError).  Kd is destroyed.
HMX_HD int dense_solve12(double* Kd, double* x) {
    for (int k = 0; k < 12; ++k) {
        int p = k;
        double amax = fabs(Kd[k * 12 + k]);
        for (int r = k + 1; r < 12; ++r) {
            const double a = fabs(Kd[r * 12 + k]);
            if (a > amax) {
                amax = a;
                p = r;
            }
        }
        if (Kd[p * 12 + k] == 0.0) return k + 1;
        if (p != k) {
            for (int cc = 0; cc < 12; ++cc) {
                const double t = Kd[k * 12 + cc];
                Kd[k * 12 + cc] = Kd[p * 12 + cc];
                Kd[p * 12 + cc] = t;
            }
            const double t = x[k];
            x[k] = x[p];
            x[p] = t;
        }
        const double rp = 1.0 / Kd[k * 12 + k];
        for (int r = k + 1; r < 12; ++r) {
            const double l = Kd[r * 12 + k] * rp;
            for (int cc = k + 1; cc < 12; ++cc) Kd[r * 12 + cc] -= l * Kd[k * 12 + cc];
            x[r] -= l * x[k];
        }
    }
    for (int k = 11; k >= 0; --k) {
        double s = x[k];
        for (int cc = k + 1; cc < 12; ++cc) s -= Kd[k * 12 + cc] * x[cc];
        x[k] = s / Kd[k * 12 + k];
    }
    return 0;
}

// delta = solve(K, -Q), retried with K + 1e-10 I on an exact zero pivot (S5:369-373).  Returns true when the ridge was used.
template <bool ALPHA, int HILL, class AT>
HMX_HD bool dense_direction(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], double (&d)[12]) {
    double Kd[144];
    dense_newton_matrix<ALPHA, HILL>(K, e, A, b, Kd);
    for (int i = 0; i < 12; ++i) d[i] = -e.Q[i];
    if (dense_solve12(Kd, d) == 0) return false;
    dense_newton_matrix<ALPHA, HILL>(K, e, A, b, Kd);
    for (int i = 0; i < 12; ++i) {
        Kd[i * 13] += 1e-10;
        d[i] = -e.Q[i];
    }
    dense_solve12(Kd, d);
    return true;
}

// One evaluation with the reference's loop nest.  step_end(g, gp, step_idx): g = g_arr[step_idx + 1], gp = g_arr[step_idx].
template <bool ALPHA, int HILL, class StepEnd>
HMX_HD void reference_loop_solve(const SolverConsts& K, const double* theta, const double* g0, uint32_t& status, uint32_t& iters,
                                 uint32_t& trips, StepEnd&& step_end) {
    double A[15], b[5], g[12], d[12], x[12];
    RegArr<11> gp;
    theta_to_Ab(theta, A, b);
    for (int i = 0; i < 12; ++i) g[i] = g0[i];
    status = 0u;
    iters = 0u;
    trips = 0u;
    for (int step = 0; step < K.n_steps; ++step) {
        const double tol = step_tol(K, step);
        for (int i = 0; i < 11; ++i) gp[i] = g[i];
        bool converged = false, nan_break = false;
        for (int it = 0; it < K.max_iter; ++it) {
            Eval e;
            residual<ALPHA, HILL>(K, g, gp, A, b, e);
            ++trips;
            for (int i = 0; i < 5; ++i) {  // clipped phi / psi written back, phi_0 and gamma are not (S5:337-340)
                g[i] = e.ph[i];
                g[6 + i] = e.ps[i];
            }
            if (e.nan) {  // S5:342-348
                nan_break = true;
                break;
            }
            if (dense_direction<ALPHA, HILL>(K, e, A, b, d)) status |= kStSingular;
            ++iters;
            double normQ = e.norm, sl = 1.0;
            bool improved = false;
            while (sl > 1e-4) {  // S5:381-425
                for (int i = 0; i < 12; ++i) x[i] = g[i] + sl * d[i];
                Eval et;
                residual<ALPHA, HILL>(K, x, gp, A, b, et);
                ++trips;
                if (!et.nan && et.norm < normQ) {
                    for (int i = 0; i < 12; ++i) g[i] = x[i];  // the accepted trial is NOT clipped
                    normQ = et.norm;
                    improved = true;
                    break;
                }
                sl *= 0.5;
            }
            if (!improved)
                for (int i = 0; i < 12; ++i) g[i] = g[i] + d[i];  // S5:427-428
            if (normQ < tol) {
                converged = true;
                break;
            }
        }
        if (!converged && !nan_break) status |= kStMaxIter;
        double chk = 0.0;
        for (int i = 0; i < 12; ++i) chk = fma(g[i], 0.0, chk);
        if (chk != 0.0) status |= kStNonfinite;
        step_end(g, gp, step);
    }
}

}  // namespace hmx
