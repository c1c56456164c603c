// hmx_coop.cuh -- K1L, the latency-regime twin of the ODE kernel: EIGHT lanes cooperate on one solve.
//
// A production TMCMC stage hands the GPU 10^3-10^4 evaluations ("latency-bound in the production
// regime"); with one thread per solve (K1, hmx_kernels.cuh) such a launch costs the latency of its slowest
// solve, ~35 000 dependent trips of ~1500 instructions each, while 95 % of the machine idles.  Here the twelve
// unknowns of a solve are spread over a group of eight lanes (four groups per warp):
//
//     lanes 0..4   species i: phi_i, psi_i, their residual rows Q_i, Q_{6+i}, the 2x2 diagonal block D_i and
//                  row i of the 5x5 system S = I + diag(g) C
//     lane  5      phi_0 and gamma: rows Q_5 (phi_0) and Q_11 (constraint)
//     lanes 6, 7   idle (they take part in the shuffles and contribute zeros to the reductions)
//
// the per-species work is embarrassingly parallel, the interaction sums / norms / closure sums are 8-lane shuffle
// reductions, and the 5x5 elimination runs row-distributed (the pivot row is broadcast, every lane updates its own
// row).  The arithmetic follows K1 (hmx_model.cuh / hmx_lane.cuh) operation by operation -- same clipping, same rebase
// of the step-start residual, same reduction trees, same pivot guard and pivoted fall-back -- except that the
// elimination is division-free and the back substitution column-oriented (shorter dependency chains, same solution up
// to rounding).  The two kernels agree to ~1e-12 on log-likelihoods with identical iteration counts on the test
// batches, and both meet the 1e-9 bar against the oracle (tests/test_gpu_coop.py).
//
// Measured (B200): a 1000-evaluation launch takes 45 ms instead of 61 ms, a 7000-evaluation launch 55 instead of 66;
// beyond ~10^4 evaluations K1 is faster (its lanes are all busy), so the C ABI switches at one resident round of
// groups (7104).  The gain is modest because the dependency chain of a trip is long even when spread out (~2500
// cycles, a third of it shuffle latency) and the decisions-as-data style costs moves and selects.
//
// The control flow is uniform per warp: the four groups of a warp are at different (time step, iteration, trial)
// positions, so decisions are data (predicates and selects), and blocks that contain shuffles are entered by the
// whole warp or not at all (__any_sync).  The state machine itself is lane_trip's (hmx_lane.cuh), replicated in
// every lane of a group; the reductions make all lanes of a group take identical decisions.
#pragma once

#include <cuda_runtime.h>

#include "hmx_kernels.cuh"

namespace hmx {

constexpr int kCoopLanes = 8;           // lanes per solve
constexpr int kCoopThreads = 128;       // 16 solves per CTA
constexpr unsigned kFullMask = 0xffffffffu;

__device__ __forceinline__ double grp_bcast(double x, int src_lane) { return __shfl_sync(kFullMask, x, src_lane); }

// butterfly over the 8 lanes of a group: ((0+1)+(2+3)) + ((4+5)+(6+7)) -- K1's tree when lanes 6, 7 hold zero
__device__ __forceinline__ double grp_sum(double x) {
    x += __shfl_xor_sync(kFullMask, x, 1);
    x += __shfl_xor_sync(kFullMask, x, 2);
    x += __shfl_xor_sync(kFullMask, x, 4);
    return x;
}
__device__ __forceinline__ double grp_max_abs(double x) {  // x >= 0 on entry
    x = max_abs(x, __shfl_xor_sync(kFullMask, x, 1));
    x = max_abs(x, __shfl_xor_sync(kFullMask, x, 2));
    x = max_abs(x, __shfl_xor_sync(kFullMask, x, 4));
    return x;
}
__device__ __forceinline__ int grp_or(int x) {
    x |= __shfl_xor_sync(kFullMask, x, 1);
    x |= __shfl_xor_sync(kFullMask, x, 2);
    x |= __shfl_xor_sync(kFullMask, x, 4);
    return x;
}

// What one lane keeps of a residual evaluation (the distributed Eval of hmx_model.cuh).
struct CoopEval {
    double v, u;        // clipped phi_i, psi_i (lane 5: clipped phi_0, dummy)
    double p, I;        // phi psi, gated interaction
    double phd, psd;    // time derivatives
    double phip, psip;  // barrier derivatives (lane 5: phip = phi0_p)
    double Q0, Q1;      // Q_i, Q_{6+i} (lane 5: Q_5, Q_11; lanes 6, 7: 0)
    double h, dh_dx, Iraw4;  // Hill gate state (uniform in the group; Iraw4 meaningful on lane 4)
    double norm;        // ||Q||_inf of the group
    bool nan;
};

struct CoopConsts {  // per-lane copies of the per-species constants
    double inv_eta, etaphi_eta, c_eta, alpha_eta, b;
    double a_row[5];  // A[role][0..4] (zero on lanes >= 5)
    double a_diag;
};

template <bool ALPHA, int HILL>
__device__ __forceinline__ void coop_residual(const SolverConsts& K, const CoopConsts& C, int role, int gbase, double x_ph, double x_ps,
                                              double gp_ph, double gp_ps, CoopEval& e) {
    const bool sp = role < 5;
    const double gam = grp_bcast(x_ps, gbase + 5);  // gamma is never clipped
    const double v = clip01(x_ph);                   // lane 5: phi_0, clipped inside Q only
    const double u = sp ? clip01(x_ps) : 0.5;
    e.v = v;
    e.u = u;
    const double m2K = -2.0 * K.Kp1;
    const double wv = fma(v, v, -v), wu = fma(u, u, -u);
    const double rv = rcp(wv), ru = rcp(wu);
    const double sv = fma(2.0, v, -1.0), su = fma(2.0, u, -1.0);
    const double fph = (m2K * sv) * (rv * rv * rv);
    const double fps = (m2K * su) * (ru * ru * ru);
    e.phip = -fph * fma(3.0 * sv, rv, 2.0);
    const double ru2 = ru * ru;
    e.psip = (4.0 * K.Kp1) * fma(5.0, wu, 3.0) * (ru2 * ru2);
    e.p = v * u;
    e.phd = (v - gp_ph) * K.inv_dt;
    e.psd = (u - gp_ps) * K.inv_dt;
    double pj[5], vj[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        pj[j] = grp_bcast(e.p, gbase + j);
        vj[j] = grp_bcast(v, gbase + j);
    }
    double s = C.a_row[0] * pj[0];
#pragma unroll
    for (int j = 1; j < 5; ++j) s = fma(C.a_row[j], pj[j], s);
    e.h = 1.0;
    e.dh_dx = 0.0;
    e.Iraw4 = s;
    if (HILL != 0) {
        const double fn = pj[3];
        if (fn > 1e-20) {
            double xn, xn1;
            hill_powers<HILL>(K, fn, xn, xn1);
            const double rden = rcp(K.Kn + xn);
            e.h = xn * rden;
            e.dh_dx = K.n_hill * K.Kn * xn1 * (rden * rden);
        } else {
            e.h = 0.0;
        }
        if (role == 4) s *= e.h;
    }
    e.I = s;
    // species rows (S5:98-123)
    const double t2 = fma(C.inv_eta, gam, fma(fma(u, u, C.etaphi_eta), e.phd, e.p * e.psd));
    const double qi = fph + t2 - C.c_eta * u * e.I;
    double qs = fps + fma(e.p, e.phd, v * v * e.psd) - C.c_eta * v * e.I;
    if (ALPHA) qs += (C.b * C.alpha_eta) * u;
    // phi_0 row and constraint row on lane 5 (S5:109-113, 126): same barrier form as a species' phi
    const double q5 = gam + fph + e.phd;
    const double q11 = ((((vj[0] + vj[1]) + vj[2]) + vj[3]) + vj[4]) + v - 1.0;
    e.Q0 = sp ? qi : (role == 5 ? q5 : 0.0);
    e.Q1 = sp ? qs : (role == 5 ? q11 : 0.0);
    e.norm = grp_max_abs(max_abs(e.Q0, e.Q1));
    const double asum = grp_sum(fabs(e.Q0) + fabs(e.Q1));
    e.nan = (asum != asum);
}

// First residual of the next time step from the accepted trial's (rebase_residual of hmx_lane.cuh, distributed).
// gp_* already hold the new previous level (the raw accepted trial); gp_ph_old is lane 5's previous phi_0.
__device__ __forceinline__ bool coop_rebase(const SolverConsts& K, const CoopConsts& C, int role, double gp_ph, double gp_ps,
                                            double gp_ph_old, CoopEval& e) {
    const bool sp = role < 5;
    const double phd = (e.v - gp_ph) * K.inv_dt, psd = (e.u - gp_ps) * K.inv_dt;
    const double dph = phd - e.phd, dps = psd - e.psd;
    const double qi = e.Q0 + fma(fma(e.u, e.u, C.etaphi_eta), dph, e.p * dps);
    const double qs = e.Q1 + fma(e.p, dph, e.v * e.v * dps);
    const double q5 = e.Q0 + ((e.v - gp_ph) - (e.v - gp_ph_old)) * K.inv_dt;
    if (sp) {
        e.phd = phd;
        e.psd = psd;
        e.Q0 = qi;
        e.Q1 = qs;
    } else if (role == 5) {
        e.phd = phd;
        e.Q0 = q5;
    }
    e.norm = grp_max_abs(max_abs(e.Q0, e.Q1));
    const double asum = grp_sum(fabs(e.Q0) + fabs(e.Q1));
    e.nan = (asum != asum);
    return asum <= 1.7976931348623157e308;
}

// Quasi-Newton direction (direction() of hmx_model.cuh, distributed).  Returns this lane's two components.
template <bool ALPHA, int HILL>
__device__ __forceinline__ void coop_direction(const SolverConsts& K, const CoopConsts& C, int role, int gbase, const CoopEval& e,
                                               double& delta_ph, double& delta_ps) {
    const bool sp = role < 5;
    const double v = e.v, u = e.u, pd = e.phd, sd = e.psd;
    const double ce = C.c_eta;
    double srow = -ce;
    if (HILL != 0 && role == 4) srow *= e.h;
    double c43x = 0.0;
    if (HILL != 0) {
        if (role == 4 && e.h < 1.0) c43x = -ce * e.Iraw4 * e.dh_dx;
    }
    const double hi = (HILL != 0 && role == 4) ? e.h : 1.0;
    const double aii = hi * C.a_diag;
    const double ap = aii * e.p;
    const double d00 = e.phip + fma(fma(u, u, C.etaphi_eta), K.inv_dt, u * sd) - ce * u * fma(2.0 * aii, u, e.I);
    const double d01 = fma(2.0 * u, pd, fma(v, sd, e.p * K.inv_dt)) - ce * fma(3.0, ap, e.I);
    const double d10 = fma(u, pd, fma(e.p, K.inv_dt, 2.0 * v * sd)) - ce * fma(3.0, ap, 2.0 * e.I);
    double d11 = e.psip + fma(v, pd, v * v * K.inv_dt) - 2.0 * ce * aii * v * v;
    if (ALPHA) d11 += C.b * C.alpha_eta;
    const double t = d01 * d10;
    const double det = fma(d00, d11, -t) + fma(-d01, d10, t);
    const double rdet = rcp(det);
    const double r0 = -e.Q0, r1 = -e.Q1;
    const double a0 = fma(d11, r0, -d01 * r1) * rdet;
    const double a1 = fma(d00, r1, -d10 * r0) * rdet;
    const double rdi = rdet * C.inv_eta;
    const double b0 = d11 * rdi;
    const double b1 = -d10 * rdi;
    const double dd0 = fma(d11, u, -d01 * v) * rdet;
    const double dd1 = fma(d00, v, -d10 * u) * rdet;
    const double gq = fma(u, dd0, v * dd1);
    const double gi = gq * srow;
    // own row of S = I + diag(g) C and its two right-hand sides (lanes >= 5: an identity row nobody reads)
    double S[7], S0[7];
#pragma unroll
    for (int j = 0; j < 5; ++j) S[j] = (j == role) ? 1.0 : (sp ? gi * C.a_row[j] : 0.0);
    if (HILL != 0) {
        if (role == 4) S[3] = fma(gq, c43x, S[3]);
    }
    S[5] = sp ? fma(u, a0, v * a1) : 0.0;
    S[6] = sp ? fma(u, b0, v * b1) : 0.0;
#pragma unroll
    for (int c = 0; c < 7; ++c) S0[c] = S[c];
    // Natural-order elimination, row-distributed: lane k broadcasts the pivot row, lanes k+1..4 update theirs.
    // Latency is what this kernel is about, so the elimination is DIVISION-FREE (row_i <- p_k row_i - S_ik row_k):
    // no reciprocal sits between two broadcasts, the five pivot reciprocals are computed side by side afterwards.
    // Rows get scaled by the pivots (all ~1: S = I + diag(g) C), which the back substitution undoes; the solution
    // is the same up to rounding.  `scale` tracks the scaling so that the pivot guard keeps K1's meaning.
    double piv[5];
    double scale = 1.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double rowk[7];
#pragma unroll
        for (int c = k; c < 7; ++c) rowk[c] = grp_bcast(S[c], gbase + k);
        piv[k] = rowk[k];
        ok = ok && (fabs(rowk[k]) >= kPivotFloor * fabs(scale));
        scale *= rowk[k];
        if (sp && role > k) {
            const double sik = S[k];
#pragma unroll
            for (int c = k + 1; c < 7; ++c) S[c] = fma(rowk[k], S[c], -sik * rowk[c]);
        }
    }
    double rp[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) rp[k] = rcp(piv[k]);
    // back substitution, column-oriented: z_k is broadcast, every row above subtracts its share at once
    double z0[5], z1[5];
    double s0 = S[5], s1 = S[6];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        z0[k] = grp_bcast(s0 * rp[k], gbase + k);
        z1[k] = grp_bcast(s1 * rp[k], gbase + k);
        s0 = fma(-S[k], z0[k], s0);  // (rows <= k only matter; the others are never read again)
        s1 = fma(-S[k], z1[k], s1);
    }
    if (__any_sync(kFullMask, !ok)) {
        // rare: some group of the warp met a small pivot -> that group redoes the solve with row pivoting
        double T[5][7], X[10];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int c = 0; c < 7; ++c) T[i][c] = grp_bcast(S0[c], gbase + i);
        if (!ok) {
            solve5x2_pivoted(&T[0][0], X);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                z0[j] = X[j];
                z1[j] = X[5 + j];
            }
        }
    }
    // closure: sum_j C_ij z_j on the own row, the phi_0 row and the constraint row (S5:222-223, 247)
    double t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const double aj = (j == role) ? 0.0 : C.a_row[j];
        t1 = fma(aj, z0[j], t1);
        t2 = fma(aj, z1[j], t2);
    }
    t1 *= srow;
    t2 *= srow;
    if (HILL != 0) {
        if (role == 4) {
            t1 = fma(c43x, z0[3], t1);
            t2 = fma(c43x, z1[3], t2);
        }
    }
    const double y10 = sp ? fma(-dd0, t1, a0) : 0.0;
    const double y11 = sp ? fma(-dd1, t1, a1) : 0.0;
    const double y20 = sp ? fma(-dd0, t2, b0) : 0.0;
    const double y21 = sp ? fma(-dd1, t2, b1) : 0.0;
    const double sumY1 = grp_sum(y10);
    const double sumY2 = grp_sum(y20);
    // lane 5 holds K55 = phi0_p + 1/dt, Q_5, Q_11
    const double rK55_own = rcp(e.phip + K.inv_dt);
    const double srhs_own = fma(e.Q0, rK55_own, -e.Q1);
    const double rK55 = grp_bcast(rK55_own, gbase + 5);
    const double srhs = grp_bcast(srhs_own, gbase + 5);
    const double dgam = (sumY1 - srhs) * rcp(sumY2 + rK55);
    if (sp) {
        delta_ph = fma(-dgam, y20, y10);
        delta_ps = fma(-dgam, y21, y11);
    } else if (role == 5) {
        delta_ph = (-e.Q0 - dgam) * rK55;
        delta_ps = dgam;
    } else {
        delta_ph = 0.0;
        delta_ps = 0.0;
    }
}

// One solve per group of eight lanes; groups fetch work from the same global counter as K1's lanes.
template <bool ALPHA, int HILL>
__global__ void __launch_bounds__(kCoopThreads) coop_ode_kernel(const __grid_constant__ OdeParams P) {
    const SolverConsts& K = P.K;
    const int lane = threadIdx.x & 31;
    const int role = lane & (kCoopLanes - 1);
    const int gbase = lane & ~(kCoopLanes - 1);
    const bool sp = role < 5;
    const long long n_groups = ((long long)gridDim.x * blockDim.x) / kCoopLanes;
    long long pos = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kCoopLanes;

    CoopConsts C;
    // groups that never receive work still run the shuffles with the warp: give them a benign system (S = I)
    C.inv_eta = C.etaphi_eta = C.c_eta = C.alpha_eta = C.b = C.a_diag = 0.0;
#pragma unroll
    for (int j = 0; j < 5; ++j) C.a_row[j] = 0.0;
    double g_ph = 0.0, g_ps = 0.0, gp_ph = 0.0, gp_ps = 0.0, d_ph = 0.0, d_ps = 0.0;
    double normQ = 0.0, step = 1.0, tol = 0.0;
    int step_idx = 0, it = 0, mode = kModeEval;
    uint32_t status = 0u, iters = 0u, trips = 0u;
    long long w = 0;
    int c = 0, next_obs = 0, next_t = 0x7fffffff;
    bool have = false, exhausted = false;

    for (;;) {
        if (!have && !exhausted) {
            if (pos >= P.N) {
                exhausted = true;
            } else {
                w = P.order ? (long long)P.order[pos] : pos;
                c = cond_index(P.cond_id, w, P.n_cond);
                const double* tp = P.theta + w * HMX_NTHETA;
                double th[HMX_NTHETA], A[15], b[5];
#pragma unroll
                for (int k = 0; k < HMX_NTHETA; ++k) th[k] = __ldg(tp + k);
                theta_to_Ab(th, A, b);
                const int r5 = sp ? role : 0;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    double a = 0.0;
#pragma unroll
                    for (int i = 0; i < 5; ++i) a = (i == r5) ? A[apk(i, j)] : a;
                    C.a_row[j] = sp ? a : 0.0;
                }
                double ad = 0.0, bb = 0.0, ie = 1.0, ee = 1.0, cc = 0.0, ae = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    if (i == r5) {
                        ad = A[apk(i, i)];
                        bb = b[i];
                        ie = K.inv_eta[i];
                        ee = K.etaphi_eta[i];
                        cc = K.c_eta[i];
                        ae = K.alpha_eta[i];
                    }
                }
                C.a_diag = sp ? ad : 0.0;
                C.b = sp ? bb : 0.0;
                C.inv_eta = sp ? ie : 0.0;     // lane 5: gamma enters Q_5 directly, not through 1/eta
                C.etaphi_eta = sp ? ee : 0.0;
                C.c_eta = sp ? cc : 0.0;
                C.alpha_eta = sp ? ae : 0.0;
                const double* g0 = P.cond[c].g0;
                double gph = 0.0, gps = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    if (i == r5) {
                        gph = g0[i];
                        gps = g0[6 + i];
                    }
                }
                g_ph = sp ? gph : (role == 5 ? g0[5] : 0.0);
                g_ps = sp ? gps : (role == 5 ? g0[11] : 0.0);
                gp_ph = g_ph;
                gp_ps = sp ? g_ps : 0.0;
                d_ph = 0.0;
                d_ps = 0.0;
                normQ = 0.0;
                step = 1.0;
                step_idx = 0;
                it = 0;
                mode = kModeEval;
                status = 0u;
                iters = 0u;
                trips = 0u;
                tol = step_tol(K, 0);
                if (P.traj) {
                    const int r0 = P.traj_row_map ? P.traj_row_map[0] : 0;
                    if (r0 >= 0) {
                        double* tr = P.traj + (w * (long long)P.traj_n_rows + r0) * 12;
                        if (sp) {
                            tr[role] = g_ph;
                            tr[6 + role] = g_ps;
                        } else if (role == 5) {
                            tr[5] = g_ph;
                            tr[11] = g_ps;
                        }
                    }
                }
                next_obs = 0;
                next_t = (P.cond[c].n_obs > 0) ? max(P.cond[c].idx[0], 1) : 0x7fffffff;
                have = true;
            }
        }
        if (__all_sync(kFullMask, !have)) break;  // every group of the warp is out of work

        // ---- candidate and residual (all lanes, also of idle groups: the shuffles need the whole warp) ----
        const bool trial = (mode == kModeTrial);
        const double sl = trial ? step : 0.0;
        const double x_ph = fma(sl, d_ph, g_ph);
        const double x_ps = fma(sl, d_ps, g_ps);
        CoopEval e;
        coop_residual<ALPHA, HILL>(K, C, role, gbase, x_ph, x_ps, gp_ph, gp_ps, e);
        if (have) ++trips;

        // ---- accept / reject: lane_trip of hmx_lane.cuh, decisions as data ----
        bool need_dir = false, finished = false, take_clipped = false, take_raw = false, nan_break = false, full_step = false;
        if (have) {
            if (!trial) {
                take_clipped = true;
                if (e.nan) {
                    finished = true;
                    nan_break = true;
                } else {
                    normQ = e.norm;
                    need_dir = true;
                }
            } else {
                const bool accept = (!e.nan) && (e.norm < normQ);
                if (accept) {
                    normQ = e.norm;
                    if (normQ < tol || it >= K.max_iter) {
                        take_raw = true;
                        finished = true;
                    } else {
                        take_clipped = true;
                        need_dir = true;
                    }
                } else {
                    step *= 0.5;
                    if (!(step > 1e-4)) full_step = true;
                }
            }
        }
        if (__any_sync(kFullMask, full_step)) {
            // no improving step length: full step, old norm kept; a non-finite direction is detected here
            const int bad = grp_or((fma(d_ph, 0.0, 0.0) != 0.0 || fma(d_ps, 0.0, 0.0) != 0.0) ? 1 : 0);
            if (full_step) {
                g_ph = g_ph + d_ph;
                g_ps = g_ps + d_ps;
                if (bad) {
                    status |= kStSingular;
                    d_ph = 0.0;
                    d_ps = 0.0;
                }
                if (normQ < tol || it >= K.max_iter) finished = true;
                else mode = kModeEval;
            }
        }
        if (take_clipped) {
            // species: clipped write-back; lane 5: phi_0 and gamma are never clipped in the iterate
            g_ph = sp ? e.v : x_ph;
            g_ps = sp ? e.u : x_ps;
        }
        if (take_raw) {
            g_ph = x_ph;
            g_ps = x_ps;
        }

        // ---- end of a time step ----
        bool done = false;
        if (__any_sync(kFullMask, finished)) {
            if (finished) {
                if (!nan_break && !(normQ < tol)) status |= kStMaxIter;
                // g_* = g_arr[step_idx + 1], gp_* = g_arr[step_idx]
                if (step_idx + 1 == next_t) {
                    const OdeCond& oc = P.cond[c];
                    double* pr = P.pairs + w * P.pair_stride;
                    do {
                        double* q = pr + next_obs * kPairStride;
                        if (sp) {
                            q[role] = g_ph;
                            q[6 + role] = g_ps;
                            q[12 + role] = gp_ph;
                            q[18 + role] = gp_ps;
                        } else if (role == 5) {
                            q[5] = g_ph;
                            q[11] = g_ps;
                            q[17] = gp_ph;
                        } else if (role == 6) {
                            q[23] = 0.0;
                        }
                        ++next_obs;
                        next_t = (next_obs < oc.n_obs) ? max(oc.idx[next_obs], 1) : 0x7fffffff;
                    } while (step_idx + 1 == next_t);
                }
                if (P.traj) {
                    const int r = P.traj_row_map ? P.traj_row_map[step_idx + 1] : step_idx + 1;
                    if (r >= 0) {
                        double* q = P.traj + (w * (long long)P.traj_n_rows + r) * 12;
                        if (sp) {
                            q[role] = g_ph;
                            q[6 + role] = g_ps;
                        } else if (role == 5) {
                            q[5] = g_ph;
                            q[11] = g_ps;
                        }
                    }
                }
            }
            // isfinite scan of the new row on the ordinary way out (the rebase covers it otherwise)
            const int nonfin = grp_or((fma(g_ph, 0.0, 0.0) != 0.0 || fma(g_ps, 0.0, 0.0) != 0.0) ? 1 : 0);
            const double gp_ph_old = gp_ph;
            if (finished) {
                if (!take_raw && nonfin) status |= kStNonfinite;
                gp_ph = g_ph;
                gp_ps = sp ? g_ps : 0.0;
                ++step_idx;
                it = 0;
                mode = kModeEval;
                tol = step_tol(K, step_idx);
                done = step_idx >= K.n_steps;
            }
            // rebase for the groups that left the step through an accepted trial (all lanes run the shuffles)
            CoopEval r = e;
            const bool finite = coop_rebase(K, C, role, gp_ph, gp_ps, gp_ph_old, r);
            if (finished && take_raw) {
                if (!finite) status |= kStNonfinite;
                if (finite && !done) {
                    e = r;
                    normQ = e.norm;
                    if (sp) {
                        g_ph = e.v;
                        g_ps = e.u;
                    }
                    need_dir = true;
                }
            }
        }

        // ---- direction solve ----
        if (__any_sync(kFullMask, need_dir)) {
            double nd_ph, nd_ps;
            coop_direction<ALPHA, HILL>(K, C, role, gbase, e, nd_ph, nd_ps);
            if (need_dir) {
                d_ph = nd_ph;
                d_ps = nd_ps;
                ++it;
                ++iters;
                step = 1.0;
                mode = kModeTrial;
            }
        }

        // ---- evaluation finished: results, next work item ----
        if (__any_sync(kFullMask, done)) {
            long long nxt = 0;
            if (done && role == 0) {
                P.status[w] = status;
                P.iters[w] = iters;
                P.trips[w] = trips;
                atomicAdd(P.totals + 0, (unsigned long long)iters);
                atomicAdd(P.totals + 1, (unsigned long long)trips);
                atomicAdd(P.cond_trips + c, (unsigned long long)trips);
                atomicAdd(P.cond_evals + c, 1ULL);
                nxt = (long long)atomicAdd(P.work_counter, 1ULL) + n_groups;
            }
            nxt = __shfl_sync(kFullMask, nxt, gbase);
            if (done) {
                pos = nxt;
                have = false;
            }
        }
    }
}

}  // namespace hmx
