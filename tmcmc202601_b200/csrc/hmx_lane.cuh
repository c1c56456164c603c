// hmx_lane.cuh -- per-lane state machine of the batched implicit integrator.
//
// The reference advances one particle at a time: for each of n_steps time steps run <= max_newton_iter
// damped quasi-Newton iterations, each with a backtracking line search (S5:291-432, 837-882).  Iteration
// and trial counts are data dependent (2-50 iterations per step, 1-14 trials per iteration), so running
// 32 particles of a warp in lock-step over (time step, iteration, trial) would idle most lanes.
//
// Here every lane carries its own (time step, iteration, line-search) position and all lanes of a warp
// execute the same "trip":   candidate -> residual -> accept/reject decision -> [direction solve].
// A lane that converges simply starts its next time step on the next trip; a lane that finishes its
// particle fetches the next one from a global work counter.  Divergence is limited to the predicated
// direction solve and the (rare) step-advance bookkeeping.
//
// Iterate semantics reproduced from the reference ("same iterate sequence"):
//   * phi, psi are clipped to [1e-12, 1-1e-12] and WRITTEN BACK to the iterate at the start of every
//     iteration; phi_0 is clipped only inside Q/K; gamma is never clipped            (S5:66-79, 337-340)
//   * NaN residual at an iteration start ends the Newton loop with the current iterate (S5:342-348)
//   * the first step length in {1, 1/2, ...} (> 1e-4) with ||Q_trial||_inf < ||Q||_inf is accepted and
//     the accepted trial is NOT clipped                                               (S5:381-425)
//   * if no trial improves, the full step is taken and the old norm is kept           (S5:427-428)
//   * the loop stops when ||Q||_inf < eps (1 + 10 t_{n+1}) or after max_newton_iter   (S5:430-431, 861)
#pragma once

#include "hmx_model.cuh"

#if defined(__CUDA_ARCH__)
#define HMX_WARP_CONVERGE() __syncwarp()
#else
#define HMX_WARP_CONVERGE() ((void)0)
#endif

namespace hmx {

enum : int32_t { kModeEval = 0, kModeTrial = 1 };

constexpr uint32_t kStNonfinite = 1u, kStMaxIter = 2u, kStSingular = 4u;  // == HMX_ST_* in include/hmx.h

#if defined(HMX_A_IN_SMEM) && defined(__CUDACC__)
using LaneA = SmemCol<HMX_A_IN_SMEM>;  // interaction matrix parked in shared memory (30 registers freed)
#else
struct LaneA {
    double v[15];
    HMX_HD double& operator[](int k) { return v[k]; }
    HMX_HD const double& operator[](int k) const { return v[k]; }
};
#endif

#if defined(HMX_STATE_IN_SMEM) && defined(__CUDACC__)
// previous time level and Newton direction parked in shared memory too: 23 more doubles out of the register
// file, which is what lets a third warp per scheduler become resident (168 registers / thread)
using LaneGP = SmemCol<HMX_STATE_IN_SMEM>;
using LaneD = SmemCol<HMX_STATE_IN_SMEM>;
#else
template <int N>
struct RegArr {
    double v[N];
    HMX_HD double& operator[](int k) { return v[k]; }
    HMX_HD const double& operator[](int k) const { return v[k]; }
};
using LaneGP = RegArr<11>;
#if defined(HMX_D_IN_SMEM) && defined(__CUDACC__)
using LaneD = SmemCol<HMX_D_IN_SMEM>;  // the Newton direction alone (written once, read once per trip): 24 registers freed
#else
using LaneD = RegArr<12>;
#endif
#endif

#if defined(HMX_STASH_IN_SMEM) && defined(__CUDACC__)
using LaneStash = SmemCol<HMX_STASH_IN_SMEM>;  // direction-solve scratch (30 doubles) parked in shared memory
#else
using LaneStash = RegStash;
#endif

struct Lane {
    LaneGP gp;       // previous time level: phi(5), phi_0, psi(5)
    double g[12];    // current Newton iterate
    LaneD d;         // current Newton direction
    LaneA A;         // packed symmetric interaction matrix
    LaneStash stash; // scratch of the direction solve
    double b[5];     // decay vector (only read when alpha != 0)
    double normQ;    // ||Q||_inf at the iterate
    double step;     // current line-search step length
    double tol;      // eps (1 + 10 t_{n+1})
    int32_t step_idx;  // time level of gp
    int32_t it;        // directions computed in this time step
    int32_t mode;
    uint32_t status;
    uint32_t iters;  // total directions (quasi-Newton iterations)
    uint32_t trips;  // total residual evaluations
    uint32_t rebased; // time steps whose first residual was derived from the previous step's last one
};

HMX_HD double step_tol(const SolverConsts& K, int32_t step_idx) {
    const double tt = (double)(step_idx + 1) * K.dt;  // S5:860-861
    return K.eps * (1.0 + 10.0 * tt);
}

// g0 = [phi_init, 1 - sum(phi_init), 0.999 x5, 0]  (S5:564-576)
HMX_HD void lane_init(const SolverConsts& K, Lane& s, const double* theta, const double* g0) {
    theta_to_Ab(theta, s.A, s.b);
#pragma unroll
    for (int i = 0; i < 12; ++i) s.g[i] = g0[i];
#pragma unroll
    for (int i = 0; i < 11; ++i) s.gp[i] = g0[i];
#pragma unroll
    for (int i = 0; i < 12; ++i) s.d[i] = 0.0;
    s.normQ = 0.0;
    s.step = 1.0;
    s.step_idx = 0;
    s.it = 0;
    s.mode = kModeEval;
    s.status = 0u;
    s.iters = 0u;
    s.trips = 0u;
    s.rebased = 0u;
    s.tol = step_tol(K, 0);
}

// Bookkeeping after a finished time step.  Returns true when all n_steps levels are done.
HMX_HD bool lane_next_level(const SolverConsts& K, Lane& s) {
    ++s.step_idx;
    s.it = 0;
    s.mode = kModeEval;
    s.tol = step_tol(K, s.step_idx);
    return s.step_idx >= K.n_steps;
}

// ... including the hand-over s.g -> previous level and the isfinite scan of the new trajectory row.
HMX_HD bool lane_advance(const SolverConsts& K, Lane& s) {
    double chk = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) chk = fma(s.g[i], 0.0, chk);
    if (chk != 0.0) s.status |= kStNonfinite;  // EV:657-667 scans the whole trajectory
#pragma unroll
    for (int i = 0; i < 11; ++i) s.gp[i] = s.g[i];
    return lane_next_level(K, s);
}

// First residual of the NEXT time step, derived from the residual of the accepted trial that just ended the
// current one.  Both are evaluated at the same clipped point; only the previous time level differs
// (gp: g_arr[n] -> the raw accepted trial), so only the time-derivative terms of Q change (S5:98-123):
//   Q_i     += (psi^2 + eta_phi/eta) (phidot' - phidot) + phi psi (psidot' - psidot)
//   Q_{6+i} += phi psi (phidot' - phidot) + phi^2 (psidot' - psidot)
//   Q_5     += (phi0dot' - phi0dot)
// with phidot' = (clip(x) - x)/dt (zero unless the accepted trial left the box).  This replaces one full
// residual evaluation per time step (2500 of ~17500 per solve) by ~60 flops, and it turns the trip that detects
// convergence into a trip that also computes the next step's first direction.
// x: the raw accepted trial (== s.g == new s.gp); gp5_old: phi_0 of the previous level before the advance.
HMX_HD bool rebase_residual(const SolverConsts& K, const Lane& s, const double (&x)[12], double gp5_old, Eval& e) {
#if defined(HMX_REBASE_FAST)
    // Common case: the accepted trial lies strictly inside the clip box, so clip(x) = x = the new previous level and the new
    // time derivatives are exactly zero: the update is "subtract the old time-derivative terms" (30 FP64 instructions fewer).
    if (e.inbox && strictly_inside_hi(x[5])) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double v = e.ph[i], u = e.ps[i];
            const double qi = e.Q[i] - fma(fma(u, u, K.etaphi_eta[i]), e.phd[i], e.p[i] * e.psd[i]);
            const double qs = e.Q[6 + i] - fma(e.p[i], e.phd[i], v * v * e.psd[i]);
#if !defined(HMX_TD_FROM_GP)
            e.phd[i] = 0.0;
            e.psd[i] = 0.0;
#endif
            e.Q[i] = qi;
            e.Q[6 + i] = qs;
        }
        e.Q[5] = e.Q[5] - (x[5] - gp5_old) * K.inv_dt;
    } else
#endif
    {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double v = e.ph[i], u = e.ps[i];
            const double phd = (v - s.gp[i]) * K.inv_dt, psd = (u - s.gp[6 + i]) * K.inv_dt;
            const double dph = phd - e.phd[i], dps = psd - e.psd[i];
#if !defined(HMX_TD_FROM_GP)
            e.phd[i] = phd;
            e.psd[i] = psd;
#endif
            const double qi = e.Q[i] + fma(fma(u, u, K.etaphi_eta[i]), dph, e.p[i] * dps);
            const double qs = e.Q[6 + i] + fma(e.p[i], dph, v * v * dps);
            e.Q[i] = qi;
            e.Q[6 + i] = qs;
        }
        const double ph0 = clip01(x[5]);
        e.Q[5] = e.Q[5] + ((ph0 - s.gp[5]) - (ph0 - gp5_old)) * K.inv_dt;
    }
    bool finite;
    norm12(e.Q, e.norm, e.nan, finite);
    // A non-finite component of x makes (clip(x) - x)/dt, hence a residual component, non-finite (gamma entered the
    // accepted residual itself, which was finite): this test doubles as the trajectory's isfinite scan (EV:657-667).
    // (On the in-box path every component of x is finite by the box test itself.)
    return finite;
}

// One trip.  `step_end(g, gp, step_idx)` is called whenever the Newton loop of the current time step has ended,
// with g = g_arr[step_idx + 1] and gp = g_arr[step_idx]; the lane then advances to the next time level by itself.  Returns true when all n_steps levels are done.
template <bool ALPHA, int HILL, class StepEnd>
HMX_HD bool lane_trip(const SolverConsts& K, Lane& s, StepEnd&& step_end) {
    double x[12];
    const bool trial = (s.mode == kModeTrial);
    // candidate = g + step d in a trial, g itself at an iteration start: one fma with step := 0 instead of a
    // select per component (0 * d + g == g exactly because d is finite whenever the mode is "iteration start":
    // it is zeroed at lane_init and after a non-finite direction, see below)
    const double sl = trial ? s.step : 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = fma(sl, s.d[i], s.g[i]);

    Eval e;
    residual<ALPHA, HILL>(K, x, s.gp, s.A, s.b, e);
    ++s.trips;

    bool need_dir = false, finished = false, take_clipped = false, take_raw = false, nan_break = false;
    if (!trial) {
        take_clipped = true;  // write-back of the clipped phi/psi
        if (e.nan) {
            finished = true;
            nan_break = true;
        } else {
            s.normQ = e.norm;
            need_dir = true;
        }
    } else {
        const bool accept = (!e.nan) && (e.norm < s.normQ);
        if (accept) {
            s.normQ = e.norm;
            if (s.normQ < s.tol || s.it >= K.max_iter) {
                take_raw = true;  // the accepted trial leaves the Newton loop unclipped
                finished = true;
            } else {
                take_clipped = true;  // next iteration: same residual, clipped write-back
                need_dir = true;
            }
        } else {
            s.step *= 0.5;
            if (!(s.step > 1e-4)) {
                // no improving step length: full step, keep the old norm.  This is also where a non-finite
                // direction ends up (every trial along it has a non-finite residual and is rejected), so the
                // finiteness test of the direction lives here, off the hot path.
                double chk = 0.0;
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                    chk = fma(s.d[i], 0.0, chk);
                    s.g[i] = s.g[i] + s.d[i];
                }
                if (chk != 0.0) {
                    s.status |= kStSingular;
#pragma unroll
                    for (int i = 0; i < 12; ++i) s.d[i] = 0.0;
                }
                if (s.normQ < s.tol || s.it >= K.max_iter) {
                    finished = true;
                } else {
                    s.mode = kModeEval;
                }
            }
        }
    }
    if (take_clipped) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            s.g[i] = e.ph[i];
            s.g[6 + i] = e.ps[i];
        }
        s.g[5] = x[5];
        s.g[11] = x[11];
    }
    bool done = false;
    if (finished) {
        if (!nan_break && !(s.normQ < s.tol)) s.status |= kStMaxIter;  // range(max_newton_iter) exhausted
#if !defined(HMX_NO_REBASE)
        if (take_raw) {
            // the common way out of a time step: an accepted trial x, which is g_arr[step_idx + 1] unclipped
            step_end(x, s.gp, s.step_idx);
            const double gp5_old = s.gp[5];
#pragma unroll
            for (int i = 0; i < 11; ++i) s.gp[i] = x[i];
            done = lane_next_level(K, s);
            const bool finite = rebase_residual(K, s, x, gp5_old, e);
            if (!finite) s.status |= kStNonfinite;
            if (finite && !done) {
                s.normQ = e.norm;
#pragma unroll
                for (int i = 0; i < 5; ++i) {  // clipped write-back at the start of the new step's first iteration
                    s.g[i] = e.ph[i];
                    s.g[6 + i] = e.ps[i];
                }
                s.g[5] = x[5];
                s.g[11] = x[11];
                need_dir = true;
                ++s.rebased;
            } else {  // ordinary path: full residual at x on the next trip (S5:342-348 handles a NaN there)
#pragma unroll
                for (int i = 0; i < 12; ++i) s.g[i] = x[i];
            }
        } else
#endif
        {
            // NaN residual at an iteration start, no improving step length at the tolerance / iteration limit
            if (take_raw) {
#pragma unroll
                for (int i = 0; i < 12; ++i) s.g[i] = x[i];
            }
            step_end(s.g, s.gp, s.step_idx);
            done = lane_advance(K, s);
        }
    }
    // All lanes of the warp meet here before the direction solve.  Without the barrier the compiler lets the
    // lanes that skipped the step-end block run ahead, and the warp executes the (expensive) solve twice:
    // once for them and once for the lanes that just started a new time step (measured: -27 % throughput).
    HMX_WARP_CONVERGE();
    if (need_dir) {
#if defined(HMX_TD_FROM_GP)
        // time derivatives at the point of `e` against the CURRENT previous level, recomputed here (20 FP64 instructions)
        // instead of being carried -- and patched by the rebase -- across the accept / step-end join (20 registers)
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            e.phd[i] = (e.ph[i] - s.gp[i]) * K.inv_dt;
            e.psd[i] = (e.ps[i] - s.gp[6 + i]) * K.inv_dt;
        }
#endif
        direction<ALPHA, HILL>(K, e, s.A, s.b, s.d, s.stash);
        ++s.it;
        ++s.iters;
        s.step = 1.0;
        s.mode = kModeTrial;
    }
    return done;
}

}  // namespace hmx
