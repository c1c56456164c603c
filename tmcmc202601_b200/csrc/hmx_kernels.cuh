// hmx_kernels.cuh -- sm_100a kernels of the likelihood hot path.
//
//   K1  ode_kernel        one thread per (particle, condition) solve; persistent lanes with a global work
//                         counter; per-lane state machine (hmx_lane.cuh); all solver state in registers;
//                         solver constants + per-condition initial state / observation rows arrive as
//                         __grid_constant__ kernel parameters (constant bank, no loads on the hot loop).
//                         Emits the (g_t, g_{t-1}) pairs at the observation rows, the status word and the
//                         iteration / residual-evaluation counters.  Optionally the whole trajectory (K1b).
//   K1c obs_kernel        one thread per evaluation: sensitivities at the observation rows, TSM variance,
//                         Gaussian log-likelihood, the reference's -1e20 guards (hmx_obs.cuh).
//   fp64_peak_kernel      register-resident DFMA chains: measured FP64 roofline denominator.
//
// Bound: FP64 FMA pipe (64 lanes/SM/clk on B200 -> a warp-wide DFMA/DMUL/DADD issues every 2 clk per SM
// sub-partition).  HBM traffic is 160 B in + ~1 KB out per evaluation against ~2e7 flop: irrelevant.
// Tensor cores are not used: the largest dense contraction on the path is 5x5.
#pragma once

#include <cuda_runtime.h>

#include "hmx_obs.cuh"
#include "hmx_dense.cuh"

namespace hmx {

#ifndef HMX_ODE_THREADS
#define HMX_ODE_THREADS 128
#endif
#ifndef HMX_ODE_MINBLOCKS
#define HMX_ODE_MINBLOCKS 2
#endif
#ifndef HMX_OBS_MINBLOCKS
#define HMX_OBS_MINBLOCKS 1  // resident 128-thread CTAs per SM the observation / TSM-row kernels are compiled for
#endif
constexpr int kOdeThreads = HMX_ODE_THREADS;      // 4 warps: one per SM sub-partition
constexpr int kOdeMinBlocks = HMX_ODE_MINBLOCKS;  // 2 -> 255 registers / thread available, 8 warps / SM

// A host cond_id array is validated by the API (HMX_ERR_INVALID).  A device array cannot be checked without a
// synchronisation, so the kernels do it: the solve runs as condition 0 (the lane needs some valid work) and the
// observation kernel then returns logL = -1e20 with HMX_ST_BAD_COND set for that evaluation.
__device__ __forceinline__ int cond_index(const int32_t* cond_id, long long i, int n_cond) {
    const int c = cond_id ? cond_id[i] : 0;
    return ((unsigned)c < (unsigned)n_cond) ? c : 0;
}
__device__ __forceinline__ bool cond_valid(const int32_t* cond_id, long long i, int n_cond) {
    return !cond_id || ((unsigned)cond_id[i] < (unsigned)n_cond);
}
constexpr uint32_t kStBadCond = 8u;  // == HMX_ST_BAD_COND

struct OdeCond {
    double g0[12];
    int32_t idx[HMX_MAX_OBS];
    int32_t n_obs;
    int32_t pad_;
};

struct OdeParams {
    SolverConsts K;
    OdeCond cond[HMX_MAX_COND];
    const double* theta;     // [N,20]
    const int32_t* cond_id;  // [N] or null
    long long N;
    double* pairs;           // [N, n_obs_max, 24]
    long long pair_stride;   // n_obs_max * 24
    double* traj;            // [N, n_steps+1, 12] (or [N, traj_n_rows, 12] with a row map) or null
    const int32_t* traj_row_map;  // [n_steps+1]: output row of time level t, -1 = not stored; null = every level
    int traj_n_rows;
    uint32_t* status;        // [N]
    uint32_t* iters;         // [N]
    uint32_t* trips;         // [N]
    unsigned long long* work_counter;  // zeroed before launch
    unsigned long long* totals;        // [2] += sum iters, sum trips
    unsigned long long* cond_trips;    // [HMX_MAX_COND] += residual evaluations per condition (cost model of the scheduler)
    unsigned long long* cond_evals;    // [HMX_MAX_COND] += evaluations per condition
    const int* order;                  // [N] work position -> evaluation index (most expensive condition first), or null
    int n_cond;
};

// ONE flat loop per lane.  Every pass is one state-machine trip; fetching the next particle, the per-step
// bookkeeping and the final stores are (rare) predicated blocks INSIDE that loop.  With a nested
// "for each particle { while (!done) trip }" the lanes of a warp that finish a particle early wait at the
// inner loop's reconvergence point for the slowest lane (SIMT), which measured as 18.5 of 32 active threads
// per instruction; the flat loop keeps all lanes of a warp in the same trip until the work list is empty.
#if defined(HMX_ODE_MAXNREG)
#define HMX_ODE_BOUNDS __maxnreg__(HMX_ODE_MAXNREG)
#else
#define HMX_ODE_BOUNDS __launch_bounds__(kOdeThreads, kOdeMinBlocks)
#endif
template <bool ALPHA, int HILL>
__global__ void HMX_ODE_BOUNDS ode_kernel(const __grid_constant__ OdeParams P) {
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    Lane s;
#if defined(HMX_A_IN_SMEM)
    __shared__ double smA[15 * kOdeThreads];
    static_assert(HMX_A_IN_SMEM == kOdeThreads, "HMX_A_IN_SMEM must equal the ODE block size");
    s.A.p = smA + threadIdx.x;
#endif
#if defined(HMX_STASH_IN_SMEM)
    __shared__ double smStash[30 * kOdeThreads];
    static_assert(HMX_STASH_IN_SMEM == kOdeThreads, "HMX_STASH_IN_SMEM must equal the ODE block size");
    s.stash.p = smStash + threadIdx.x;
#endif
#if defined(HMX_D_IN_SMEM) && !defined(HMX_STATE_IN_SMEM)
    __shared__ double smD[12 * kOdeThreads];
    static_assert(HMX_D_IN_SMEM == kOdeThreads, "HMX_D_IN_SMEM must equal the ODE block size");
    s.d.p = smD + threadIdx.x;
#endif
#if defined(HMX_STATE_IN_SMEM)
    __shared__ double smGP[11 * kOdeThreads];
    __shared__ double smD[12 * kOdeThreads];
    static_assert(HMX_STATE_IN_SMEM == kOdeThreads, "HMX_STATE_IN_SMEM must equal the ODE block size");
    s.gp.p = smGP + threadIdx.x;
    s.d.p = smD + threadIdx.x;
#endif
    long long w = 0;       // evaluation this lane is working on
    int c = 0;             // its condition
    int next_obs = 0, next_t = 0x7fffffff;
    bool have = false;     // lane holds an unfinished evaluation

    for (;;) {
        if (!have) {
            if (pos >= P.N) break;  // work list empty: this lane is finished for good
            w = P.order ? (long long)P.order[pos] : pos;
            c = cond_index(P.cond_id, w, P.n_cond);
            double th[HMX_NTHETA];
            const double* tp = P.theta + w * HMX_NTHETA;
#pragma unroll
            for (int k = 0; k < HMX_NTHETA; ++k) th[k] = __ldg(tp + k);
            lane_init(P.K, s, th, P.cond[c].g0);
            if (P.traj) {
                const int r0 = P.traj_row_map ? P.traj_row_map[0] : 0;
                if (r0 >= 0) {
                    double* tr = P.traj + (w * (long long)P.traj_n_rows + r0) * 12;
#pragma unroll
                    for (int i = 0; i < 12; ++i) tr[i] = s.g[i];
                }
            }
            next_obs = 0;
            next_t = (P.cond[c].n_obs > 0) ? max(P.cond[c].idx[0], 1) : 0x7fffffff;
            have = true;
        }
        // end of a time step (called from inside the trip): g = g_arr[step_idx + 1], gp = g_arr[step_idx]
        auto step_end = [&](const double (&g)[12], const LaneGP& gp, int step_idx) {
            if (step_idx + 1 == next_t) {
                // observation row reached: keep (g_t, g_{t-1}) for the sensitivity kernel
                const OdeCond& oc = P.cond[c];
                // one contiguous 192-byte block per (evaluation, row): six full sectors per store burst.  (An entry-major
                // layout makes the observation kernel's reads coalesce -- K1c 3.89 -> 3.57 ms per 1.05 M evaluations -- but
                // turns these stores into 24 partial-sector writes: DRAM traffic of this kernel 1040 -> 2200 B per evaluation.)
                double* pr = P.pairs + w * P.pair_stride;
                do {
                    double* q = pr + next_obs * kPairStride;
#pragma unroll
                    for (int i = 0; i < 12; ++i) q[i] = g[i];
#pragma unroll
                    for (int i = 0; i < 11; ++i) q[12 + i] = gp[i];
                    q[23] = 0.0;
                    ++next_obs;
                    next_t = (next_obs < oc.n_obs) ? max(oc.idx[next_obs], 1) : 0x7fffffff;
                } while (step_idx + 1 == next_t);
            }
            if (P.traj) {
                const int r = P.traj_row_map ? P.traj_row_map[step_idx + 1] : step_idx + 1;
                if (r >= 0) {
                    double* q = P.traj + (w * (long long)P.traj_n_rows + r) * 12;
#pragma unroll
                    for (int i = 0; i < 12; ++i) q[i] = g[i];
                }
            }
        };
        if (lane_trip<ALPHA, HILL>(P.K, s, step_end)) {
            P.status[w] = s.status;
            P.iters[w] = s.iters;
            P.trips[w] = s.trips;
            atomicAdd(P.totals + 0, (unsigned long long)s.iters);
            atomicAdd(P.totals + 1, (unsigned long long)s.trips);
            atomicAdd(P.cond_trips + c, (unsigned long long)s.trips);
            atomicAdd(P.cond_evals + c, 1ULL);
            pos = (long long)atomicAdd(P.work_counter, 1ULL) + nthreads;
            have = false;
        }
    }
}

// K1r: evaluations flagged HMX_ST_SINGULAR by the production kernel are solved again from their initial state with the
// reference's loop nest and a dense, pivoted linear solve incl. the 1e-10 ridge retry (hmx_dense.cuh).  One thread per
// evaluation; threads of unflagged evaluations return at once (force != 0: every evaluation, a test / cross-check mode).
template <bool ALPHA, int HILL>
__global__ void __launch_bounds__(128) resolve_kernel(const __grid_constant__ OdeParams P, int force) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.N) return;
    if (!force && !(P.status[w] & kStSingular)) return;
    const int c = cond_index(P.cond_id, w, P.n_cond);
    const OdeCond& oc = P.cond[c];
    double th[HMX_NTHETA];
    for (int k = 0; k < HMX_NTHETA; ++k) th[k] = P.theta[w * HMX_NTHETA + k];
    if (P.traj) {
        const int r0 = P.traj_row_map ? P.traj_row_map[0] : 0;
        if (r0 >= 0)
            for (int i = 0; i < 12; ++i) P.traj[(w * (long long)P.traj_n_rows + r0) * 12 + i] = oc.g0[i];
    }
    int next_obs = 0;
    auto step_end = [&](const double (&g)[12], const RegArr<11>& gp, int step_idx) {
        while (next_obs < oc.n_obs && max(oc.idx[next_obs], 1) == step_idx + 1) {
            double* q = P.pairs + w * P.pair_stride + next_obs * kPairStride;
            for (int i = 0; i < 12; ++i) q[i] = g[i];
            for (int i = 0; i < 11; ++i) q[12 + i] = gp[i];
            q[23] = 0.0;
            ++next_obs;
        }
        if (P.traj) {
            const int r = P.traj_row_map ? P.traj_row_map[step_idx + 1] : step_idx + 1;
            if (r >= 0)
                for (int i = 0; i < 12; ++i) P.traj[(w * (long long)P.traj_n_rows + r) * 12 + i] = g[i];
        }
    };
    uint32_t st = 0u, its = 0u, trips = 0u;
    reference_loop_solve<ALPHA, HILL>(P.K, th, oc.g0, st, its, trips, step_end);
    P.status[w] = st;
    P.iters[w] = its;
    P.trips[w] = trips;
}

// ---- cost-ordered work list ---------------------------------------------------------------------------
// Evaluation cost is data dependent (9k-34k residual evaluations) and differs systematically between
// conditions (means 13k-23k on the synthetic prior workload).  Lanes fetch work dynamically, so the kernel's
// tail is about one evaluation long; handing out the expensive condition first shortens it.  The expected
// cost per condition is an exponential moving average of the previous launches' own counters, kept on the
// device (no host synchronisation); the first launch of a context uses the input order.
struct SchedState {
    double ema_trips[HMX_MAX_COND];      // expected residual evaluations per evaluation, 0 = unknown
    unsigned int count[HMX_MAX_COND];    // histogram of the current batch
    unsigned int cursor[HMX_MAX_COND];   // scatter cursors
    unsigned int offset[HMX_MAX_COND];
    int use_order;                       // 0 -> identity order (no cost information yet, or a single condition)
};

__global__ void sched_hist_kernel(const int32_t* __restrict__ cond_id, long long N, SchedState* st, int n_cond) {
    __shared__ unsigned int h[HMX_MAX_COND];
    if (threadIdx.x < HMX_MAX_COND) h[threadIdx.x] = 0u;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(&h[cond_index(cond_id, i, n_cond)], 1u);
    __syncthreads();
    if (threadIdx.x < HMX_MAX_COND && h[threadIdx.x]) atomicAdd(&st->count[threadIdx.x], h[threadIdx.x]);
}

__global__ void sched_plan_kernel(SchedState* st, int n_cond) {
    int rank[HMX_MAX_COND];
    int known = 0, used = 0;
    for (int c = 0; c < n_cond; ++c) {
        rank[c] = c;
        if (st->count[c]) {
            ++used;
            if (st->ema_trips[c] > 0.0) ++known;
        }
    }
    st->use_order = (used > 1 && known == used) ? 1 : 0;
    for (int a = 1; a < n_cond; ++a)  // insertion sort by expected cost, descending
        for (int b = a; b > 0 && st->ema_trips[rank[b]] > st->ema_trips[rank[b - 1]]; --b) {
            const int t = rank[b];
            rank[b] = rank[b - 1];
            rank[b - 1] = t;
        }
    unsigned int off = 0;
    for (int k = 0; k < n_cond; ++k) {
        st->offset[rank[k]] = off;
        off += st->count[rank[k]];
        st->cursor[rank[k]] = 0u;
    }
}

// Counting-sort scatter.  The cursors are bumped once per (warp, condition) -- the lanes of a warp that hold the same condition
// elect a leader (__match_any_sync) -- instead of once per evaluation: batches arrive sorted or in long runs per condition, and
// 10^6 atomics on four addresses cost 0.5 ms per launch.  The order inside a condition is arbitrary either way (it only
// decides which lane solves which evaluation).
__global__ void sched_scatter_kernel(const int32_t* __restrict__ cond_id, long long N, SchedState* st, int* __restrict__ order, int n_cond) {
    const bool use = st->use_order != 0;
    const unsigned lane = threadIdx.x & 31u;
    for (long long base = (long long)blockIdx.x * blockDim.x; base < N; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        const bool act = i < N;
        if (!use) {
            if (act) order[i] = (int)i;
            continue;
        }
        const int c = act ? cond_index(cond_id, i, n_cond) : -1;
        const unsigned peers = __match_any_sync(0xffffffffu, c);
        if (act) {
            const int leader = __ffs(peers) - 1;
            unsigned first = 0u;
            if ((int)lane == leader) first = atomicAdd(&st->cursor[c], (unsigned)__popc(peers));
            first = __shfl_sync(peers, first, leader);
            order[st->offset[c] + first + __popc(peers & ((1u << lane) - 1u))] = (int)i;
        }
    }
}

// after the ODE kernel: fold this launch's per-condition cost into the moving average, re-arm the accumulators
__global__ void sched_update_kernel(SchedState* st, unsigned long long* cond_trips, unsigned long long* cond_evals, int n_cond) {
    for (int c = 0; c < n_cond; ++c) {
        if (cond_evals[c]) {
            const double m = (double)cond_trips[c] / (double)cond_evals[c];
            st->ema_trips[c] = st->ema_trips[c] > 0.0 ? 0.5 * (st->ema_trips[c] + m) : m;
        }
        cond_trips[c] = 0ull;
        cond_evals[c] = 0ull;
        st->count[c] = 0u;
    }
}

struct ObsParams {
    SolverConsts K;
    ThetaMap TM;
    const CondConsts* cond;  // device array [n_cond]
    const double* theta;
    const int32_t* cond_id;
    long long N;
    const double* pairs;
    long long pair_stride;
    const uint32_t* status_in;
    uint32_t* status_out;  // may alias status_in, may be null
    double* logL;
    double* obs_state;     // [N, n_obs_max, 12] or null
    long long obs_stride;  // n_obs_max * 12
    double* sigma2;        // [N, n_obs_max, 12] or null: TSM variance at the observation rows (S4:1146-1149)
    double* x1;            // [N, n_obs_max, 12, 20] or null: local sensitivities at the observation rows (S4:1150)
    double* pull;          // [N, n_obs_max, 5] or null: w r / v per observation entry (input of the gradient kernel)
    int n_cond;
};

template <bool ALPHA, int HILL>
__global__ void __launch_bounds__(128, HMX_OBS_MINBLOCKS) obs_kernel(const __grid_constant__ ObsParams P) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= P.N) return;
    const int c = cond_index(P.cond_id, w, P.n_cond);
    const CondConsts& C = P.cond[c];
    double th[HMX_NTHETA];
#pragma unroll
    for (int k = 0; k < HMX_NTHETA; ++k) th[k] = __ldg(P.theta + w * HMX_NTHETA + k);
    uint32_t st = 0u;
    const double ll = obs_loglik<ALPHA, HILL>(P.K, C, P.TM, th, P.pairs + w * P.pair_stride, P.status_in[w],
                                              P.obs_state ? P.obs_state + w * P.obs_stride : nullptr, &st,
                                              P.sigma2 ? P.sigma2 + w * P.obs_stride : nullptr,
                                              P.x1 ? P.x1 + w * P.obs_stride * HMX_NTHETA : nullptr,
                                              P.pull ? P.pull + w * (P.obs_stride / 12) * 5 : nullptr);
    const bool okc = cond_valid(P.cond_id, w, P.n_cond);
    P.logL[w] = okc ? ll : -1e20;
    if (P.status_out) P.status_out[w] = okc ? st : (st | kStBadCond);
}

// ---- TSM variance at EVERY time level (BiofilmTSM.solve_tsm in full, S4:1095-1160) -------------------------------------
// One thread per (evaluation, time level t >= 1) on a stored trajectory: sigma2[t] from (g_t, g_{t-1}); level 0 repeats
// level 1 (S4:1139-1141).  Used by hmx_solve_tsm_batch and by the all-rows finiteness guard (EV:657-667 scans sig2 at all
// T + 1 rows, hmx_set_sig2_guard).  ~(1 residual + n_active directions) per level: ~1.5x the cost of the solve itself.
struct TsmRowsParams {
    SolverConsts K;
    ThetaMap TM;
    const CondConsts* cond;
    const double* theta;
    const int32_t* cond_id;
    long long N;
    const double* traj;  // [N, n_steps + 1, 12]
    double* sigma2;      // [N, n_steps + 1, 12] or null (guard only)
    uint32_t* status;    // [N]: kStNonfinite is OR-ed in when a variance (or a state) is non-finite
    int n_cond;
};

template <bool ALPHA, int HILL>
__global__ void __launch_bounds__(128, HMX_OBS_MINBLOCKS) tsm_rows_kernel(const __grid_constant__ TsmRowsParams P) {
    const int T = P.K.n_steps;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long w = tid / T;
    const int t = (int)(tid - w * T) + 1;
    if (w >= P.N) return;
    const CondConsts& C = P.cond[cond_index(P.cond_id, w, P.n_cond)];
    double th[HMX_NTHETA], A[15], b[5], x[12], gp[11], sig2[12], covp[5];
#pragma unroll
    for (int k = 0; k < HMX_NTHETA; ++k) th[k] = __ldg(P.theta + w * HMX_NTHETA + k);
    theta_to_Ab(th, A, b);
    const double* tr = P.traj + (w * (long long)(T + 1) + t) * 12;
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = tr[i];
#pragma unroll
    for (int i = 0; i < 11; ++i) gp[i] = tr[i - 12];
#pragma unroll
    for (int i = 0; i < 12; ++i) sig2[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) covp[i] = 0.0;
    bool ok = true;
    if (C.tsm_variance) ok = tsm_row<ALPHA, HILL>(P.K, C, P.TM, th, A, b, x, gp, sig2, covp, nullptr);
    double chk = 0.0;
#pragma unroll
    for (int i = 0; i < 12; ++i) chk = fma(x[i], 0.0, chk);
    if (!ok || chk != 0.0) atomicOr(P.status + w, kStNonfinite);
    if (P.sigma2) {
        double* out = P.sigma2 + (w * (long long)(T + 1) + t) * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) out[i] = sig2[i];
        if (t == 1) {
#pragma unroll
            for (int i = 0; i < 12; ++i) out[i - 12] = sig2[i];
        }
    }
}

// ---- FP64 peak: 8 independent DFMA chains per thread, operands in registers ---------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // keeps the chains alive; practically never taken
}
constexpr int kPeakFmaPerIter = 64;  // 8 chains x 8 unrolled

}  // namespace hmx
