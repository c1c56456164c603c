// hmx_api.cu -- C ABI (include/hmx.h) over the sm_100a kernels.  Build: see __graft_entry__.build().
//
// Memory model: every array argument may be a host or a device pointer.  Host inputs are copied to
// context-owned device buffers with cudaMemcpyAsync on the context stream; host outputs are copied back
// and the call synchronises the stream before returning.  Device pointers are used in place and the call
// only enqueues work.  Scratch (observation pairs, status words, counters, reduction partials) lives in
// growable device buffers owned by the context, sized for 180 GB HBM3e: the per-evaluation footprint is
// 20*8 (theta) + n_obs*24*8 (pairs) + 3*4 (status/iters/trips) + 8 (logL) bytes ~= 1.1 KB, so one 4 Mi-evaluation
// chunk needs ~4.6 GB.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hmx.h"
#include "hmx_consts.h"
#include "hmx_kernels.cuh"
#include "hmx_stage.cuh"
#include "hmx_mutation.cuh"
#include "hmx_coop.cuh"
#include "hmx_legacy.cuh"
#include "hmx_grad.cuh"

using namespace hmx;

namespace {

constexpr long long kDefaultChunkEvals = 1ll << 22;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

thread_local std::string g_create_error;

}  // namespace

struct hmx_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    hmx_config cfg;
    SolverConsts K;
    CondConsts C[HMX_MAX_COND];
    OdeCond oc[HMX_MAX_COND];
    int n_obs_max = 0;
    bool alpha = false;
    int hill = 0;
    int ode_blocks_per_sm = 1;
    int coop_blocks_per_sm = 1;
    long long coop_max_evals = 0;   // batches up to this size run the 8-lanes-per-solve kernel (HMX_COOP_MAX_EVALS overrides)
    bool last_kernel_coop = false;
    bool force_reference_loop = false;  // HMX_FORCE_REFERENCE_LOOP=1: every evaluation goes through K1r (cross-check / tests)
    bool sig2_guard_all = false;  // EV:657-667 on all T + 1 rows instead of the observation rows (hmx_set_sig2_guard)
    std::string err;

    // device scratch
    DevBuf theta, cond_id, pairs, status, iters, trips, logL, obs_state, traj, sig2, x1;
    DevBuf st_in0, st_in1, st_in2, st_out0, st_out1, partial, cs, order, logL_scratch;
    DevBuf g_traj, g_pull, g_grad, g_dstate;  // hmx_loglik_grad_batch
    DevBuf cs_info;                            // multi-CTA cumulative sum: per-tile records
    bool cumsum_single_cta = false;            // HMX_CUMSUM_SINGLE_CTA=1: the single-CTA kernel for every N (tests compare the two)
    DevBuf m_props, m_theta, m_cond, m_inside, m_ll, m_status, m_in_theta, m_in_logL, m_out_theta, m_out_logL, m_groups;  // hmx_mutate
    unsigned long long* d_mut_counts = nullptr;  // [8] accepted, candidates, nonfinite, maxiter, singular
    SchedState* d_sched = nullptr;
    CondConsts* d_cond = nullptr;
    unsigned long long* d_counters = nullptr;  // [0] work counter, [1] sum iters, [2] sum trips, [4..12) trips per condition, [12..20) evaluations per condition
    StageState* d_stage = nullptr;
    unsigned int* d_tickets = nullptr;  // [64]
    double* d_small = nullptr;          // [2048] moments / mean / scatter / small outputs

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false;
    bool timing_pending = false;
    double last_ode_ms = 0.0;
    double ode_ms_acc = 0.0;  // ODE kernel time of the chunks of the batch in flight that have already been read
    long long last_evals = 0;
    long long chunk_evals = kDefaultChunkEvals;  // HMX_CHUNK_EVALS overrides (tests exercise the multi-chunk path with it)
    long long launches = 0;
    // row selection of the trajectory call in flight (hmx_trajectory_rows); null = all n_steps + 1 levels
    const int32_t* cur_row_map = nullptr;
    int cur_n_rows = 0;
    DevBuf row_map;
    std::vector<void*> populations;  // hmx_population_create
};

namespace {

int fail(hmx_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail(ctx, HMX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));       \
    } while (0)

// Scratch buffers come from the stream-ordered allocator: cudaMalloc / cudaFree synchronise the whole DEVICE, so a context
// growing (or releasing) a buffer would wait for -- and, holding the driver's lock, stall -- the kernels of every other
// context on the GPU (chains on their own streams, run_batch_TMCMC(merge_launches=False)).  cudaMallocAsync / cudaFreeAsync
// are ordered on the context's stream only.
cudaError_t dev_alloc(hmx_ctx* ctx, void** p, size_t bytes) {
    cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaErrorNotSupported) {
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    return e;
}

void dev_free(hmx_ctx* ctx, void* p) {
    if (!p) return;
    if (cudaFreeAsync(p, ctx->stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
    }
}

int ensure(hmx_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    dev_free(ctx, b.p);  // ordered after the work already queued on the stream that may still read it
    b.p = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = dev_alloc(ctx, &b.p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.p = nullptr;
        return fail(ctx, HMX_ERR_OOM, std::string("cudaMallocAsync(") + std::to_string(want) + "): " + cudaGetErrorString(e));
    }
    b.cap = want;
    return 0;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// input: returns a device pointer valid on ctx->stream (the original one, or a staged copy)
int stage_in(hmx_ctx* ctx, const void* p, size_t bytes, DevBuf& buf, const void** out) {
    if (!p) {
        *out = nullptr;
        return 0;
    }
    if (is_device_ptr(p)) {
        *out = p;
        return 0;
    }
    int rc = ensure(ctx, buf, bytes);
    if (rc) return rc;
    CU(cudaMemcpyAsync(buf.p, p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = buf.p;
    return 0;
}

// output: device pointer to write to; *host_dst set when a copy-back is needed afterwards
int stage_out(hmx_ctx* ctx, void* p, size_t bytes, DevBuf& buf, void** dev, void** host_dst) {
    *host_dst = nullptr;
    if (!p) {
        *dev = nullptr;
        return 0;
    }
    if (is_device_ptr(p)) {
        *dev = p;
        return 0;
    }
    int rc = ensure(ctx, buf, bytes);
    if (rc) return rc;
    *dev = buf.p;
    *host_dst = p;
    return 0;
}

int red_blocks(long long N) {
    long long b = (N + (long long)kRedThreads * 4 - 1) / ((long long)kRedThreads * 4);
    return (int)std::max(1ll, std::min<long long>(b, kRedMaxBlocks));
}

template <bool ALPHA, int HILL>
void launch_ode(hmx_ctx* ctx, const OdeParams& P, int blocks) {
    // the production exponent n_hill = 4 has its own instance of the hot kernel (HILL = 4: no exponent switch per trip);
    // the arithmetic is the same as the generic integer-exponent instance, bit for bit
    if (HILL == 1 && ctx->K.hill_mode == 4) ode_kernel<ALPHA, 4><<<blocks, kOdeThreads, 0, ctx->stream>>>(P);
    else ode_kernel<ALPHA, HILL><<<blocks, kOdeThreads, 0, ctx->stream>>>(P);
}
template <bool ALPHA, int HILL>
void launch_resolve(hmx_ctx* ctx, const OdeParams& P, int blocks, int force) {
    resolve_kernel<ALPHA, HILL><<<blocks, 128, 0, ctx->stream>>>(P, force);
}
template <bool ALPHA, int HILL>
void launch_coop(hmx_ctx* ctx, const OdeParams& P, int blocks) {
    coop_ode_kernel<ALPHA, HILL><<<blocks, kCoopThreads, 0, ctx->stream>>>(P);
}
template <bool ALPHA, int HILL>
int coop_occupancy() {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, coop_ode_kernel<ALPHA, HILL>, kCoopThreads, 0);
    return nb;
}
template <bool ALPHA, int HILL>
void launch_obs(hmx_ctx* ctx, const ObsParams& P, int blocks) {
    obs_kernel<ALPHA, HILL><<<blocks, 128, 0, ctx->stream>>>(P);
}
template <bool ALPHA, int HILL>
void launch_tsm_rows(hmx_ctx* ctx, const TsmRowsParams& P, long long blocks) {
    tsm_rows_kernel<ALPHA, HILL><<<(unsigned)blocks, 128, 0, ctx->stream>>>(P);
}
template <bool ALPHA, int HILL>
void launch_grad(hmx_ctx* ctx, const GradParams& P, int blocks) {
    grad_kernel<ALPHA, HILL><<<blocks, 128, 0, ctx->stream>>>(P);
}
template <bool ALPHA, int HILL>
int ode_occupancy() {
    int nb = 0, nb4 = 1 << 30;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, ode_kernel<ALPHA, HILL>, kOdeThreads, 0);
    if (HILL == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb4, ode_kernel<ALPHA, 4>, kOdeThreads, 0);
    return std::min(nb, nb4);
}

#define HMX_DISPATCH(FN, ...)                                  \
    do {                                                       \
        if (ctx->alpha) {                                      \
            if (ctx->hill == 0) FN<true, 0>(__VA_ARGS__);      \
            else if (ctx->hill == 1) FN<true, 1>(__VA_ARGS__); \
            else FN<true, 2>(__VA_ARGS__);                     \
        } else {                                               \
            if (ctx->hill == 0) FN<false, 0>(__VA_ARGS__);     \
            else if (ctx->hill == 1) FN<false, 1>(__VA_ARGS__);\
            else FN<false, 2>(__VA_ARGS__);                    \
        }                                                      \
    } while (0)

int query_occupancy(hmx_ctx* ctx) {
    int nb = 0;
    if (ctx->alpha) {
        nb = ctx->hill == 0 ? ode_occupancy<true, 0>() : ctx->hill == 1 ? ode_occupancy<true, 1>() : ode_occupancy<true, 2>();
    } else {
        nb = ctx->hill == 0 ? ode_occupancy<false, 0>() : ctx->hill == 1 ? ode_occupancy<false, 1>() : ode_occupancy<false, 2>();
    }
    return std::max(nb, 1);
}

int query_coop_occupancy(hmx_ctx* ctx) {
    int nb = 0;
    if (ctx->alpha) {
        nb = ctx->hill == 0 ? coop_occupancy<true, 0>() : ctx->hill == 1 ? coop_occupancy<true, 1>() : coop_occupancy<true, 2>();
    } else {
        nb = ctx->hill == 0 ? coop_occupancy<false, 0>() : ctx->hill == 1 ? coop_occupancy<false, 1>() : coop_occupancy<false, 2>();
    }
    return std::max(nb, 1);
}

// the ODE solve + observation likelihood for one chunk of evaluations already resident on the device
int run_chunk(hmx_ctx* ctx, const double* d_theta, const int32_t* d_cond, long long N, double* d_logL,
              double* d_obs_state, uint32_t* d_status, uint32_t* d_iters, double* d_traj, bool want_logL,
              double* d_sigma2 = nullptr, double* d_x1 = nullptr, double* d_pull = nullptr, double* d_sigma2_full = nullptr) {
    int rc;
    // TSM variance at every time level (hmx_solve_tsm_batch) or the all-rows guard (hmx_set_sig2_guard): needs the trajectory
    const bool tsm_rows = d_sigma2_full || (ctx->sig2_guard_all && want_logL);
    if (tsm_rows && !d_traj) {
        if ((rc = ensure(ctx, ctx->g_traj, (size_t)N * (ctx->cfg.n_steps + 1) * 12 * sizeof(double)))) return rc;
        d_traj = (double*)ctx->g_traj.p;
    }
    const long long pair_stride = (long long)std::max(ctx->n_obs_max, 1) * kPairStride;
    if ((rc = ensure(ctx, ctx->pairs, (size_t)N * pair_stride * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->trips, (size_t)N * sizeof(uint32_t)))) return rc;
    if (!d_status) {
        if ((rc = ensure(ctx, ctx->status, (size_t)N * sizeof(uint32_t)))) return rc;
        d_status = (uint32_t*)ctx->status.p;
    }
    if (!d_iters) {
        if ((rc = ensure(ctx, ctx->iters, (size_t)N * sizeof(uint32_t)))) return rc;
        d_iters = (uint32_t*)ctx->iters.p;
    }
    CU(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned long long), ctx->stream));  // work counter only

    OdeParams P;
    P.K = ctx->K;
    for (int k = 0; k < HMX_MAX_COND; ++k) P.cond[k] = ctx->oc[k];
    P.theta = d_theta;
    P.cond_id = d_cond;
    P.N = N;
    P.pairs = (double*)ctx->pairs.p;
    P.pair_stride = pair_stride;
    P.traj = d_traj;
    P.traj_row_map = ctx->cur_row_map;
    P.traj_n_rows = ctx->cur_row_map ? ctx->cur_n_rows : ctx->cfg.n_steps + 1;
    P.status = d_status;
    P.iters = d_iters;
    P.trips = (uint32_t*)ctx->trips.p;
    P.work_counter = ctx->d_counters;
    P.totals = ctx->d_counters + 1;
    P.cond_trips = ctx->d_counters + 4;
    P.cond_evals = ctx->d_counters + 12;
    P.order = nullptr;
    P.n_cond = ctx->cfg.n_cond;
    if (d_cond && ctx->cfg.n_cond > 1 && N < (1ll << 31)) {
        // expensive conditions first (see SchedState); three tiny launches, no host synchronisation
        if ((rc = ensure(ctx, ctx->order, (size_t)N * sizeof(int)))) return rc;
        const int sb = red_blocks(N);
        sched_hist_kernel<<<sb, kRedThreads, 0, ctx->stream>>>(d_cond, N, ctx->d_sched, ctx->cfg.n_cond);
        sched_plan_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_sched, ctx->cfg.n_cond);
        sched_scatter_kernel<<<sb, kRedThreads, 0, ctx->stream>>>(d_cond, N, ctx->d_sched, (int*)ctx->order.p, ctx->cfg.n_cond);
        CU(cudaGetLastError());
        ctx->launches += 3;
        P.order = (const int*)ctx->order.p;
    }

    if (ctx->timing) {
        // one event pair per context: fold the previous chunk's duration into the running sum before re-recording
        if (ctx->timing_pending) {
            float ms = 0.f;
            CU(cudaEventSynchronize(ctx->ev1));
            if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->ode_ms_acc += ms;
            else cudaGetLastError();
            ctx->timing_pending = false;
        }
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
    }
    if (N <= ctx->coop_max_evals) {
        // latency regime: eight lanes per solve (K1L, hmx_coop.cuh)
        const int per_cta = kCoopThreads / kCoopLanes;
        const long long want = (N + per_cta - 1) / per_cta;
        const int blocks = (int)std::min<long long>(want, (long long)ctx->sm_count * ctx->coop_blocks_per_sm);
        HMX_DISPATCH(launch_coop, ctx, P, blocks);
        ctx->last_kernel_coop = true;
    } else {
        const long long want_blocks = (N + kOdeThreads - 1) / kOdeThreads;
        const int blocks = (int)std::min<long long>(want_blocks, (long long)ctx->sm_count * ctx->ode_blocks_per_sm);
        HMX_DISPATCH(launch_ode, ctx, P, blocks);
        ctx->last_kernel_coop = false;
    }
    CU(cudaGetLastError());
    ++ctx->launches;
    if (ctx->timing) {
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        ctx->timing_pending = true;
    }
    sched_update_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_sched, ctx->d_counters + 4, ctx->d_counters + 12, ctx->cfg.n_cond);
    CU(cudaGetLastError());
    ++ctx->launches;
    // K1r: evaluations flagged HMX_ST_SINGULAR are solved again with the dense, pivoted solve + ridge retry (hmx_dense.cuh)
    HMX_DISPATCH(launch_resolve, ctx, P, (int)((N + 127) / 128), ctx->force_reference_loop ? 1 : 0);
    CU(cudaGetLastError());
    ++ctx->launches;

    if (tsm_rows) {
        if (ctx->cur_row_map) return fail(ctx, HMX_ERR_INVALID, "the all-rows TSM variance needs the full trajectory (no row selection)");
        TsmRowsParams R;
        R.K = ctx->K;
        ThetaMap tmr = {HMX_THETA_P, HMX_THETA_Q};
        R.TM = tmr;
        R.cond = ctx->d_cond;
        R.theta = d_theta;
        R.cond_id = d_cond;
        R.N = N;
        R.traj = d_traj;
        R.sigma2 = d_sigma2_full;
        R.status = d_status;
        R.n_cond = ctx->cfg.n_cond;
        const long long rb = (N * (long long)ctx->cfg.n_steps + 127) / 128;
        HMX_DISPATCH(launch_tsm_rows, ctx, R, rb);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    if (want_logL) {
        ObsParams O;
        O.K = ctx->K;
        ThetaMap tm = {HMX_THETA_P, HMX_THETA_Q};
        O.TM = tm;
        O.cond = ctx->d_cond;
        O.theta = d_theta;
        O.cond_id = d_cond;
        O.N = N;
        O.pairs = (const double*)ctx->pairs.p;
        O.pair_stride = pair_stride;
        O.status_in = d_status;
        O.status_out = d_status;
        O.logL = d_logL;
        O.obs_state = d_obs_state;
        O.obs_stride = (long long)std::max(ctx->n_obs_max, 1) * 12;
        O.sigma2 = d_sigma2;
        O.x1 = d_x1;
        O.pull = d_pull;
        O.n_cond = ctx->cfg.n_cond;
        const int ob = (int)((N + 127) / 128);
        HMX_DISPATCH(launch_obs, ctx, O, ob);
        CU(cudaGetLastError());
        ++ctx->launches;
    }
    return 0;
}

int batch_impl(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, long long N, double* logL,
               double* obs_state, uint32_t* status, uint32_t* newton_iters, double* traj, double* sigma2 = nullptr,
               double* x1 = nullptr) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N < 0 || (N > 0 && !theta)) return fail(ctx, HMX_ERR_INVALID, "theta is NULL or N < 0");
    if (!traj && !logL && N > 0) return fail(ctx, HMX_ERR_INVALID, "logL is NULL");
    CU(cudaSetDevice(ctx->device));
    if (cond_id && N > 0 && !is_device_ptr(cond_id)) {
        // a host array is checked here; a device array is checked by the kernels (HMX_ST_BAD_COND, logL = -1e20)
        for (long long i = 0; i < N; ++i)
            if (cond_id[i] < 0 || cond_id[i] >= ctx->cfg.n_cond)
                return fail(ctx, HMX_ERR_INVALID, "cond_id[" + std::to_string(i) + "] = " + std::to_string(cond_id[i]) +
                                                      " outside [0, n_cond)");
    }
    if (ctx->timing_pending) {  // an unread timing of an earlier batch is dropped, not added to this one
        ctx->timing_pending = false;
    }
    ctx->ode_ms_acc = 0.0;
    ctx->last_evals = N;
    CU(cudaMemsetAsync(ctx->d_counters + 1, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (N == 0) return 0;
    const long long obs_stride = (long long)std::max(ctx->n_obs_max, 1) * 12;
    const long long traj_stride = (long long)(ctx->cur_row_map ? ctx->cur_n_rows : ctx->cfg.n_steps + 1) * 12;
    // trajectories are 100-240 KB per evaluation: keep a chunk's output below ~8 GiB
    long long chunk = traj ? std::max(1ll, std::min(ctx->chunk_evals, (8ll << 30) / (traj_stride * 8))) : ctx->chunk_evals;
    if (ctx->sig2_guard_all && !traj)  // the guard keeps a scratch trajectory per chunk
        chunk = std::max(1ll, std::min(chunk, (4ll << 30) / ((long long)(ctx->cfg.n_steps + 1) * 12 * 8)));
    bool need_sync = false;
    for (long long off = 0; off < N; off += chunk) {
        const long long n = std::min(chunk, N - off);
        int rc;
        const void *d_theta, *d_cond;
        void *d_logL, *d_obs, *d_status, *d_iters, *d_traj, *d_sig2, *d_x1;
        void *h_logL, *h_obs, *h_status, *h_iters, *h_traj, *h_sig2, *h_x1;
        if ((rc = stage_in(ctx, theta + off * HMX_NTHETA, (size_t)n * HMX_NTHETA * sizeof(double), ctx->theta, &d_theta))) return rc;
        if ((rc = stage_in(ctx, cond_id ? cond_id + off : nullptr, (size_t)n * sizeof(int32_t), ctx->cond_id, &d_cond))) return rc;
        if ((rc = stage_out(ctx, logL ? logL + off : nullptr, (size_t)n * sizeof(double), ctx->logL, &d_logL, &h_logL))) return rc;
        if ((rc = stage_out(ctx, obs_state ? obs_state + off * obs_stride : nullptr, (size_t)n * obs_stride * sizeof(double),
                            ctx->obs_state, &d_obs, &h_obs))) return rc;
        if ((rc = stage_out(ctx, status ? status + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out0, &d_status, &h_status))) return rc;
        if ((rc = stage_out(ctx, newton_iters ? newton_iters + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out1, &d_iters, &h_iters))) return rc;
        if ((rc = stage_out(ctx, traj ? traj + off * traj_stride : nullptr, (size_t)n * traj_stride * sizeof(double), ctx->traj,
                            &d_traj, &h_traj))) return rc;
        if ((rc = stage_out(ctx, sigma2 ? sigma2 + off * obs_stride : nullptr, (size_t)n * obs_stride * sizeof(double),
                            ctx->sig2, &d_sig2, &h_sig2))) return rc;
        if ((rc = stage_out(ctx, x1 ? x1 + off * obs_stride * HMX_NTHETA : nullptr,
                            (size_t)n * obs_stride * HMX_NTHETA * sizeof(double), ctx->x1, &d_x1, &h_x1))) return rc;
        if (!traj && !d_logL) return fail(ctx, HMX_ERR_INVALID, "logL is NULL");
        if ((rc = run_chunk(ctx, (const double*)d_theta, (const int32_t*)d_cond, n, (double*)d_logL, (double*)d_obs,
                            (uint32_t*)d_status, (uint32_t*)d_iters, (double*)d_traj, d_logL != nullptr, (double*)d_sig2,
                            (double*)d_x1))) return rc;
        if (h_sig2) CU(cudaMemcpyAsync(h_sig2, d_sig2, (size_t)n * obs_stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_x1) CU(cudaMemcpyAsync(h_x1, d_x1, (size_t)n * obs_stride * HMX_NTHETA * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_logL) CU(cudaMemcpyAsync(h_logL, d_logL, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_obs) CU(cudaMemcpyAsync(h_obs, d_obs, (size_t)n * obs_stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_status) CU(cudaMemcpyAsync(h_status, d_status, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_iters) CU(cudaMemcpyAsync(h_iters, d_iters, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_traj) CU(cudaMemcpyAsync(h_traj, d_traj, (size_t)n * traj_stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        const bool any_host = h_logL || h_obs || h_status || h_iters || h_traj || h_sig2 || h_x1;
        need_sync = need_sync || any_host;
        // staging buffers are reused by the next chunk: drain before overwriting them
        if (off + chunk < N && (any_host || !is_device_ptr(theta))) CU(cudaStreamSynchronize(ctx->stream));
    }
    if (need_sync) CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}

}  // namespace

struct hmx_multi {
    std::vector<hmx_ctx*> ctx;
    std::string err;
};

static thread_local std::string g_multi_create_error;

// contiguous block of rank k of G over N items (the partition of SURVEY 8(e); identical to distributed.shard_bounds)
static void multi_bounds(long long N, int G, int k, long long* lo, long long* hi) {
    const long long per = (N + G - 1) / G;
    *lo = std::min(N, per * k);
    *hi = std::min(N, per * (k + 1));
}

// run fn(k, lo, hi) for every device on its own host thread; first error wins
template <class F>
static int multi_run(hmx_multi* m, long long N, F&& fn) {
    const int G = (int)m->ctx.size();
    std::vector<int> rc(G, HMX_OK);
    std::vector<std::thread> th;
    for (int k = 0; k < G; ++k) {
        long long lo, hi;
        multi_bounds(N, G, k, &lo, &hi);
        th.emplace_back([&, k, lo, hi] { rc[k] = (hi > lo) ? fn(k, lo, hi) : HMX_OK; });
    }
    for (auto& t : th) t.join();
    for (int k = 0; k < G; ++k)
        if (rc[k] != HMX_OK) {
            m->err = std::string("device slot ") + std::to_string(k) + ": " + hmx_last_error(m->ctx[k]);
            return rc[k];
        }
    return HMX_OK;
}

extern "C" {

int hmx_abi_version(void) { return HMX_ABI_VERSION; }

const char* hmx_last_error(const hmx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int hmx_create(hmx_ctx** out, const hmx_config* cfg, int32_t device) {
    hmx_ctx* ctx = nullptr;  // errors before allocation go to the thread-local message
    if (!out) return fail(nullptr, HMX_ERR_INVALID, "ctx out-pointer is NULL");
    *out = nullptr;
    const char* why = "";
    if (validate_config(cfg, &why) != 0) return fail(nullptr, HMX_ERR_INVALID, std::string("invalid config: ") + why);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(nullptr, HMX_ERR_NO_DEVICE, "no CUDA device visible: libhamilton has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, HMX_ERR_INVALID, "device index out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, HMX_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, HMX_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                                     "; this library is built for sm_100a (B200) only");
    ctx = new (std::nothrow) hmx_ctx();
    if (!ctx) return fail(nullptr, HMX_ERR_OOM, "out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cfg = *cfg;
    make_solver_consts(*cfg, ctx->K);
    ctx->alpha = (cfg->alpha_const != 0.0);
    ctx->hill = hill_variant(ctx->K);
    for (int k = 0; k < cfg->n_cond; ++k) {
        make_cond_consts(*cfg, k, ctx->C[k]);
        memset(&ctx->oc[k], 0, sizeof(OdeCond));
        for (int i = 0; i < 12; ++i) ctx->oc[k].g0[i] = ctx->C[k].g0[i];
        for (int o = 0; o < HMX_MAX_OBS; ++o) ctx->oc[k].idx[o] = ctx->C[k].idx[o];
        ctx->oc[k].n_obs = ctx->C[k].n_obs;
        ctx->n_obs_max = std::max(ctx->n_obs_max, (int)ctx->C[k].n_obs);
    }
    for (int k = cfg->n_cond; k < HMX_MAX_COND; ++k) memset(&ctx->oc[k], 0, sizeof(OdeCond));

#define CU_CREATE(call)                                                                                    \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_);                           \
            hmx_destroy(ctx);                                                                              \
            return HMX_ERR_CUDA;                                                                           \
        }                                                                                                  \
    } while (0)
    CU_CREATE(cudaSetDevice(device));
    CU_CREATE(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    // small per-context state, stream-ordered like the scratch buffers (no device-wide synchronisation: see dev_alloc)
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_cond, sizeof(CondConsts) * HMX_MAX_COND));
    CU_CREATE(cudaMemcpyAsync(ctx->d_cond, ctx->C, sizeof(CondConsts) * HMX_MAX_COND, cudaMemcpyHostToDevice, ctx->stream));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_counters, 20 * sizeof(unsigned long long)));
    CU_CREATE(cudaMemsetAsync(ctx->d_counters, 0, 20 * sizeof(unsigned long long), ctx->stream));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_sched, sizeof(SchedState)));
    CU_CREATE(cudaMemsetAsync(ctx->d_sched, 0, sizeof(SchedState), ctx->stream));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_stage, sizeof(StageState)));
    CU_CREATE(cudaMemsetAsync(ctx->d_stage, 0, sizeof(StageState), ctx->stream));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_tickets, 64 * sizeof(unsigned int)));
    CU_CREATE(cudaMemsetAsync(ctx->d_tickets, 0, 64 * sizeof(unsigned int), ctx->stream));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_mut_counts, 8 * kMutMaxGroups * sizeof(unsigned long long)));
    CU_CREATE(dev_alloc(ctx, (void**)&ctx->d_small, 4096 * sizeof(double)));
    CU_CREATE(cudaStreamSynchronize(ctx->stream));
    CU_CREATE(cudaEventCreate(&ctx->ev0));
    CU_CREATE(cudaEventCreate(&ctx->ev1));
    ctx->ode_blocks_per_sm = query_occupancy(ctx);
    ctx->coop_blocks_per_sm = query_coop_occupancy(ctx);
    // K1L is OPT-IN (hmx_set_latency_kernel / HMX_COOP_MAX_EVALS): it agrees with K1 to ~1e-12, not bit for bit, so a
    // size-dependent switch would make a likelihood depend on the size of the batch it travels in -- and with it the
    // guarantees "merged launches == separate launches" and "sharded likelihoods == single-GPU likelihoods" (both bitwise) would go.
    ctx->coop_max_evals = 0;
    if (const char* cm = getenv("HMX_COOP_MAX_EVALS")) ctx->coop_max_evals = atoll(cm);
    if (const char* sc = getenv("HMX_CUMSUM_SINGLE_CTA")) ctx->cumsum_single_cta = atoi(sc) != 0;
    if (const char* fr = getenv("HMX_FORCE_REFERENCE_LOOP")) ctx->force_reference_loop = atoi(fr) != 0;
    if (const char* ce = getenv("HMX_CHUNK_EVALS")) {
        const long long v = atoll(ce);
        if (v > 0) ctx->chunk_evals = v;
    }
    *out = ctx;
    return HMX_OK;
}

void hmx_destroy(hmx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf* bufs[] = {&ctx->theta, &ctx->cond_id, &ctx->pairs, &ctx->status, &ctx->iters, &ctx->trips, &ctx->logL,
                      &ctx->obs_state, &ctx->traj, &ctx->sig2, &ctx->x1, &ctx->st_in0, &ctx->st_in1, &ctx->st_in2, &ctx->st_out0, &ctx->st_out1,
                      &ctx->partial, &ctx->cs, &ctx->order, &ctx->m_props, &ctx->m_theta, &ctx->m_cond, &ctx->m_inside,
                      &ctx->m_ll, &ctx->m_status, &ctx->m_in_theta, &ctx->m_in_logL, &ctx->m_out_theta, &ctx->m_out_logL, &ctx->row_map, &ctx->m_groups, &ctx->logL_scratch,
                      &ctx->g_traj, &ctx->g_pull, &ctx->g_grad, &ctx->g_dstate, &ctx->cs_info};
    for (DevBuf* b : bufs) dev_free(ctx, b->p);
    void* small[] = {ctx->d_cond, ctx->d_counters, ctx->d_sched, ctx->d_stage, ctx->d_tickets, ctx->d_mut_counts, ctx->d_small};
    for (void* q : small) dev_free(ctx, q);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->populations) cudaFree(p);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

// ---- device-resident arrays owned by the context ---------------------------------------------------------
int hmx_population_create(hmx_ctx* ctx, int64_t n_rows, int32_t n_cols, double** dev) {
    if (!ctx) return HMX_ERR_INVALID;
    if (!dev || n_rows < 0 || n_cols < 1) return fail(ctx, HMX_ERR_INVALID, "hmx_population_create: bad arguments");
    *dev = nullptr;
    CU(cudaSetDevice(ctx->device));
    void* p = nullptr;
    const size_t bytes = std::max<size_t>((size_t)n_rows * (size_t)n_cols * sizeof(double), 256);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, HMX_ERR_OOM, std::string("hmx_population_create: cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    ctx->populations.push_back(p);
    *dev = (double*)p;
    return HMX_OK;
}

int hmx_population_destroy(hmx_ctx* ctx, double* dev) {
    if (!ctx) return HMX_ERR_INVALID;
    auto it = std::find(ctx->populations.begin(), ctx->populations.end(), (void*)dev);
    if (it == ctx->populations.end()) return fail(ctx, HMX_ERR_INVALID, "hmx_population_destroy: not an array of this context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(*it);
    ctx->populations.erase(it);
    return HMX_OK;
}

int hmx_population_upload(hmx_ctx* ctx, double* dev, const double* host, int64_t n_values) {
    if (!ctx) return HMX_ERR_INVALID;
    if (n_values < 0 || (n_values > 0 && (!dev || !host))) return fail(ctx, HMX_ERR_INVALID, "hmx_population_upload: bad arguments");
    CU(cudaSetDevice(ctx->device));
    // asynchronous for pinned sources; a pageable source is staged by the runtime before the call returns
    if (n_values) CU(cudaMemcpyAsync(dev, host, (size_t)n_values * sizeof(double), cudaMemcpyDefault, ctx->stream));
    return HMX_OK;
}

int hmx_population_download(hmx_ctx* ctx, double* host, const double* dev, int64_t n_values) {
    if (!ctx) return HMX_ERR_INVALID;
    if (n_values < 0 || (n_values > 0 && (!dev || !host))) return fail(ctx, HMX_ERR_INVALID, "hmx_population_download: bad arguments");
    CU(cudaSetDevice(ctx->device));
    if (n_values) CU(cudaMemcpyAsync(host, dev, (size_t)n_values * sizeof(double), cudaMemcpyDefault, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_set_stream(hmx_ctx* ctx, void* cuda_stream) {
    if (!ctx) return HMX_ERR_INVALID;
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (next != ctx->stream) {
        // scratch buffers are allocated (and released) in stream order: drain the old stream so that everything it allocated
        // is complete memory for the new one
        CU(cudaSetDevice(ctx->device));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stream = next;
    return HMX_OK;
}

int hmx_synchronize(hmx_ctx* ctx) {
    if (!ctx) return HMX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_loglik_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* logL,
                     double* obs_state, uint32_t* status, uint32_t* newton_iters) {
    return batch_impl(ctx, theta, cond_id, N, logL, obs_state, status, newton_iters, nullptr);
}

int hmx_tsm_obs_rows(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* x0, double* sigma2,
                     double* x1, double* logL, uint32_t* status) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N > 0 && !sigma2 && !x1 && !x0) return fail(ctx, HMX_ERR_INVALID, "hmx_tsm_obs_rows: no output requested");
    CU(cudaSetDevice(ctx->device));
    double* ll = logL;
    if (!ll && N > 0) {  // the observation kernel always writes the log-likelihood: give it context scratch
        int rc = ensure(ctx, ctx->logL_scratch, (size_t)N * sizeof(double));
        if (rc) return rc;
        ll = (double*)ctx->logL_scratch.p;
    }
    return batch_impl(ctx, theta, cond_id, N, ll, x0, status, nullptr, nullptr, sigma2, x1);
}

int hmx_set_sig2_guard(hmx_ctx* ctx, int32_t all_rows) {
    if (!ctx) return HMX_ERR_INVALID;
    ctx->sig2_guard_all = all_rows != 0;
    return HMX_OK;
}

int hmx_solve_tsm_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* x0, double* sigma2,
                        uint32_t* status) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N < 0 || (N > 0 && (!theta || !sigma2))) return fail(ctx, HMX_ERR_INVALID, "hmx_solve_tsm_batch: NULL argument or N < 0");
    CU(cudaSetDevice(ctx->device));
    if (cond_id && N > 0 && !is_device_ptr(cond_id))
        for (long long i = 0; i < N; ++i)
            if (cond_id[i] < 0 || cond_id[i] >= ctx->cfg.n_cond) return fail(ctx, HMX_ERR_INVALID, "hmx_solve_tsm_batch: cond_id outside [0, n_cond)");
    ctx->last_evals = N;
    ctx->ode_ms_acc = 0.0;
    ctx->timing_pending = false;
    CU(cudaMemsetAsync(ctx->d_counters + 1, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (N == 0) return HMX_OK;
    const long long stride = (long long)(ctx->cfg.n_steps + 1) * 12;
    const long long chunk = std::max(1ll, std::min(ctx->chunk_evals, (4ll << 30) / (stride * 8)));
    int rc;
    bool need_sync = false;
    for (long long off = 0; off < N; off += chunk) {
        const long long n = std::min(chunk, N - off);
        const void *d_theta, *d_cond;
        void *d_x0, *h_x0, *d_s2, *h_s2, *d_st, *h_st;
        if ((rc = stage_in(ctx, theta + off * HMX_NTHETA, (size_t)n * HMX_NTHETA * sizeof(double), ctx->theta, &d_theta))) return rc;
        if ((rc = stage_in(ctx, cond_id ? cond_id + off : nullptr, (size_t)n * sizeof(int32_t), ctx->cond_id, &d_cond))) return rc;
        if ((rc = stage_out(ctx, x0 ? x0 + off * stride : nullptr, (size_t)n * stride * sizeof(double), ctx->traj, &d_x0, &h_x0))) return rc;
        if ((rc = stage_out(ctx, sigma2 + off * stride, (size_t)n * stride * sizeof(double), ctx->g_dstate, &d_s2, &h_s2))) return rc;
        if ((rc = stage_out(ctx, status ? status + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out0, &d_st, &h_st))) return rc;
        if ((rc = run_chunk(ctx, (const double*)d_theta, (const int32_t*)d_cond, n, nullptr, nullptr, (uint32_t*)d_st, nullptr, (double*)d_x0,
                            false, nullptr, nullptr, nullptr, (double*)d_s2))) return rc;
        if (h_x0) CU(cudaMemcpyAsync(h_x0, d_x0, (size_t)n * stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_s2) CU(cudaMemcpyAsync(h_s2, d_s2, (size_t)n * stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_st) CU(cudaMemcpyAsync(h_st, d_st, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        need_sync = need_sync || h_x0 || h_s2 || h_st;
        if (off + chunk < N) CU(cudaStreamSynchronize(ctx->stream));
    }
    if (need_sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_loglik_grad_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* logL, double* grad,
                          double* dstate, uint32_t* status) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N < 0 || (N > 0 && (!theta || !logL || !grad))) return fail(ctx, HMX_ERR_INVALID, "hmx_loglik_grad_batch: NULL argument or N < 0");
    CU(cudaSetDevice(ctx->device));
    if (cond_id && N > 0 && !is_device_ptr(cond_id))
        for (long long i = 0; i < N; ++i)
            if (cond_id[i] < 0 || cond_id[i] >= ctx->cfg.n_cond) return fail(ctx, HMX_ERR_INVALID, "hmx_loglik_grad_batch: cond_id outside [0, n_cond)");
    ctx->last_evals = N;
    ctx->ode_ms_acc = 0.0;
    ctx->timing_pending = false;
    CU(cudaMemsetAsync(ctx->d_counters + 1, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (N == 0) return HMX_OK;
    const int nobs = std::max(ctx->n_obs_max, 1);
    const long long traj_stride = (long long)(ctx->cfg.n_steps + 1) * 12;
    const long long ds_stride = (long long)nobs * 12 * HMX_NTHETA;
    // the recursion reads the whole trajectory (240 KB per evaluation at 2500 steps): chunks of <= 4 GiB of it
    const long long chunk = std::max(1ll, std::min(ctx->chunk_evals, (4ll << 30) / (traj_stride * 8)));
    int rc;
    bool need_sync = false;
    for (long long off = 0; off < N; off += chunk) {
        const long long n = std::min(chunk, N - off);
        const void *d_theta, *d_cond;
        void *d_logL, *h_logL, *d_grad, *h_grad, *d_ds, *h_ds, *d_st, *h_st;
        if ((rc = stage_in(ctx, theta + off * HMX_NTHETA, (size_t)n * HMX_NTHETA * sizeof(double), ctx->theta, &d_theta))) return rc;
        if ((rc = stage_in(ctx, cond_id ? cond_id + off : nullptr, (size_t)n * sizeof(int32_t), ctx->cond_id, &d_cond))) return rc;
        if ((rc = stage_out(ctx, logL + off, (size_t)n * sizeof(double), ctx->logL, &d_logL, &h_logL))) return rc;
        if ((rc = stage_out(ctx, grad + off * HMX_NTHETA, (size_t)n * HMX_NTHETA * sizeof(double), ctx->g_grad, &d_grad, &h_grad))) return rc;
        if ((rc = stage_out(ctx, dstate ? dstate + off * ds_stride : nullptr, (size_t)n * ds_stride * sizeof(double), ctx->g_dstate, &d_ds, &h_ds))) return rc;
        if ((rc = stage_out(ctx, status ? status + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out0, &d_st, &h_st))) return rc;
        if (!d_st) {
            if ((rc = ensure(ctx, ctx->status, (size_t)n * sizeof(uint32_t)))) return rc;
            d_st = ctx->status.p;
        }
        if ((rc = ensure(ctx, ctx->g_traj, (size_t)n * traj_stride * sizeof(double)))) return rc;
        if ((rc = ensure(ctx, ctx->g_pull, (size_t)n * nobs * 5 * sizeof(double)))) return rc;
        // K1b (trajectory + observation pairs) and K1c (logL, pulls w r / v), then the sensitivity recursion
        if ((rc = run_chunk(ctx, (const double*)d_theta, (const int32_t*)d_cond, n, (double*)d_logL, nullptr, (uint32_t*)d_st, nullptr,
                            (double*)ctx->g_traj.p, true, nullptr, nullptr, (double*)ctx->g_pull.p))) return rc;
        GradParams G;
        G.K = ctx->K;
        ThetaMap tm = {HMX_THETA_P, HMX_THETA_Q};
        G.TM = tm;
        G.cond = ctx->d_cond;
        G.theta = (const double*)d_theta;
        G.cond_id = (const int32_t*)d_cond;
        G.N = n;
        G.traj = (const double*)ctx->g_traj.p;
        G.pull = (const double*)ctx->g_pull.p;
        G.status = (const uint32_t*)d_st;
        G.grad = (double*)d_grad;
        G.dstate = (double*)d_ds;
        G.n_obs_max = nobs;
        G.n_cond = ctx->cfg.n_cond;
        const int blocks = (int)((n * HMX_NTHETA + 127) / 128);
        HMX_DISPATCH(launch_grad, ctx, G, blocks);
        CU(cudaGetLastError());
        ++ctx->launches;
        if (h_logL) CU(cudaMemcpyAsync(h_logL, d_logL, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_grad) CU(cudaMemcpyAsync(h_grad, d_grad, (size_t)n * HMX_NTHETA * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_ds) CU(cudaMemcpyAsync(h_ds, d_ds, (size_t)n * ds_stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_st) CU(cudaMemcpyAsync(h_st, d_st, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        const bool any_host = h_logL || h_grad || h_ds || h_st;
        need_sync = need_sync || any_host;
        if (off + chunk < N) CU(cudaStreamSynchronize(ctx->stream));  // trajectory / staging scratch is reused by the next chunk
    }
    if (need_sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_trajectory_batch(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, double* g,
                         uint32_t* status, uint32_t* newton_iters) {
    if (!ctx) return HMX_ERR_INVALID;
    if (!g && N > 0) return fail(ctx, HMX_ERR_INVALID, "g is NULL");
    return batch_impl(ctx, theta, cond_id, N, nullptr, nullptr, status, newton_iters, g);
}

int hmx_trajectory_rows(hmx_ctx* ctx, const double* theta, const int32_t* cond_id, int64_t N, const int32_t* rows, int32_t n_rows,
                        double* g, uint32_t* status, uint32_t* newton_iters) {
    if (!ctx) return HMX_ERR_INVALID;
    if (!g && N > 0) return fail(ctx, HMX_ERR_INVALID, "g is NULL");
    const int T = ctx->cfg.n_steps;
    if (!rows || n_rows < 1 || n_rows > T + 1) return fail(ctx, HMX_ERR_INVALID, "hmx_trajectory_rows: need 1 <= n_rows <= n_steps + 1");
    CU(cudaSetDevice(ctx->device));
    std::vector<int32_t> rows_host;
    if (is_device_ptr(rows)) {  // a device row list is fetched (it is a handful of integers): the map is built on the host
        rows_host.resize((size_t)n_rows);
        CU(cudaMemcpyAsync(rows_host.data(), rows, (size_t)n_rows * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        rows = rows_host.data();
    }
    std::string map((size_t)(T + 1) * sizeof(int32_t), '\0');
    int32_t* m = (int32_t*)&map[0];
    for (int t = 0; t <= T; ++t) m[t] = -1;
    for (int r = 0; r < n_rows; ++r) {
        if (rows[r] < 0 || rows[r] > T || (r > 0 && rows[r] <= rows[r - 1]))
            return fail(ctx, HMX_ERR_INVALID, "hmx_trajectory_rows: rows must be strictly ascending in [0, n_steps]");
        m[rows[r]] = r;
    }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure(ctx, ctx->row_map, map.size());
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->row_map.p, m, map.size(), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // `map` lives on this stack frame
    ctx->cur_row_map = (const int32_t*)ctx->row_map.p;
    ctx->cur_n_rows = n_rows;
    rc = batch_impl(ctx, theta, cond_id, N, nullptr, nullptr, status, newton_iters, g);
    ctx->cur_row_map = nullptr;
    ctx->cur_n_rows = 0;
    return rc;
}

int hmx_legacy_trajectory_rows(hmx_ctx* ctx, const hmx_legacy_config* cfg, const double* theta, int64_t N, const int32_t* rows,
                               int32_t n_rows, double* g, uint32_t* status, uint32_t* newton_iters) {
    if (!ctx) return HMX_ERR_INVALID;
    if (!cfg || N < 0 || (N > 0 && (!theta || !g))) return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: NULL argument or N < 0");
    if (cfg->n_species != 4 && cfg->n_species != 5) return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: n_species must be 4 or 5");
    if (!(cfg->dt > 0.0) || cfg->n_steps < 1 || cfg->max_newton_iter < 1)
        return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: dt > 0, n_steps >= 1, max_newton_iter >= 1 required");
    const int NS = cfg->n_species, NG = 2 * NS + 2, NTH = NS == 4 ? 14 : 20, T = cfg->n_steps;
    for (int i = 0; i < NS; ++i)
        if (!(cfg->eta[i] != 0.0)) return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: eta must be non-zero");
    if (rows && (n_rows < 1 || n_rows > T + 1)) return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: need 1 <= n_rows <= n_steps + 1");
    if (rows && is_device_ptr(rows)) return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: rows must be a host array");
    CU(cudaSetDevice(ctx->device));
    if (N == 0) return HMX_OK;
    LegacyParams P;
    make_legacy_consts(*cfg, P.K);
    const int out_rows = rows ? n_rows : T + 1;
    int rc;
    P.row_map = nullptr;
    if (rows) {
        std::vector<int32_t> m((size_t)T + 1, -1);
        for (int r = 0; r < n_rows; ++r) {
            if (rows[r] < 0 || rows[r] > T || (r > 0 && rows[r] <= rows[r - 1]))
                return fail(ctx, HMX_ERR_INVALID, "hmx_legacy_trajectory_rows: rows must be strictly ascending in [0, n_steps]");
            m[rows[r]] = r;
        }
        if ((rc = ensure(ctx, ctx->row_map, m.size() * sizeof(int32_t)))) return rc;
        CU(cudaMemcpyAsync(ctx->row_map.p, m.data(), m.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));  // `m` dies with this scope
        P.row_map = (const int32_t*)ctx->row_map.p;
    }
    const long long stride = (long long)out_rows * NG;
    const long long chunk = std::max(1ll, std::min(ctx->chunk_evals, (8ll << 30) / (stride * 8)));
    bool need_sync = false;
    for (long long off = 0; off < N; off += chunk) {
        const long long n = std::min(chunk, N - off);
        const void* d_theta;
        void *d_g, *h_g, *d_st, *h_st, *d_it, *h_it;
        if ((rc = stage_in(ctx, theta + off * NTH, (size_t)n * NTH * sizeof(double), ctx->theta, &d_theta))) return rc;
        if ((rc = stage_out(ctx, g + off * stride, (size_t)n * stride * sizeof(double), ctx->traj, &d_g, &h_g))) return rc;
        if ((rc = stage_out(ctx, status ? status + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out0, &d_st, &h_st))) return rc;
        if ((rc = stage_out(ctx, newton_iters ? newton_iters + off : nullptr, (size_t)n * sizeof(uint32_t), ctx->st_out1, &d_it, &h_it))) return rc;
        P.theta = (const double*)d_theta;
        P.N = n;
        P.traj = (double*)d_g;
        P.n_rows = out_rows;
        P.status = (uint32_t*)d_st;
        P.iters = (uint32_t*)d_it;
        const int blocks = (int)((n + 127) / 128);
        if (NS == 4) legacy_ode_kernel<4><<<blocks, 128, 0, ctx->stream>>>(P);
        else legacy_ode_kernel<5><<<blocks, 128, 0, ctx->stream>>>(P);
        CU(cudaGetLastError());
        ++ctx->launches;
        if (h_g) CU(cudaMemcpyAsync(h_g, d_g, (size_t)n * stride * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_st) CU(cudaMemcpyAsync(h_st, d_st, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        if (h_it) CU(cudaMemcpyAsync(h_it, d_it, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        const bool any_host = h_g || h_st || h_it;
        need_sync = need_sync || any_host;
        if (off + chunk < N && (any_host || !is_device_ptr(theta))) CU(cudaStreamSynchronize(ctx->stream));
    }
    if (need_sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_beta_step(hmx_ctx* ctx, const double* logL, int64_t N, double beta, double target_ess_ratio,
                  double min_delta_beta, double max_delta_beta, double* delta_beta, double* ess_at_lo) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !logL || !delta_beta) return fail(ctx, HMX_ERR_INVALID, "hmx_beta_step: bad arguments");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void* d_l;
    if ((rc = stage_in(ctx, logL, (size_t)N * sizeof(double), ctx->st_in0, &d_l))) return rc;
    const int nb = red_blocks(N);
    if ((rc = ensure(ctx, ctx->partial, (size_t)2 * kRedMaxBlocks * sizeof(double)))) return rc;
    double* part = (double*)ctx->partial.p;
    const bool out_dev = is_device_ptr(delta_beta);
    double* d_db = out_dev ? delta_beta : ctx->d_small;
    double* d_ess = ess_at_lo ? (is_device_ptr(ess_at_lo) ? ess_at_lo : ctx->d_small + 1) : nullptr;
    max_kernel<<<nb, kRedThreads, 0, ctx->stream>>>((const double*)d_l, N, 1.0, part, ctx->d_stage);
    bisect_init_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_stage, beta);
    const double target = target_ess_ratio * (double)N;
    for (int it = 0; it < 25; ++it)
        bisect_step_kernel<<<nb, kRedThreads, 0, ctx->stream>>>((const double*)d_l, N, target, part, ctx->d_stage);
    bisect_final_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_stage, beta, min_delta_beta, max_delta_beta, d_db, d_ess);
    CU(cudaGetLastError());
    ctx->launches += 28;
    bool sync = false;
    if (!out_dev) {
        CU(cudaMemcpyAsync(delta_beta, d_db, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (ess_at_lo && !is_device_ptr(ess_at_lo)) {
        CU(cudaMemcpyAsync(ess_at_lo, d_ess, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_weights(hmx_ctx* ctx, const double* logL, int64_t N, double delta_beta, double* w, double* log_Sj) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !logL || !w) return fail(ctx, HMX_ERR_INVALID, "hmx_weights: bad arguments");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void* d_l;
    void *d_w, *h_w;
    if ((rc = stage_in(ctx, logL, (size_t)N * sizeof(double), ctx->st_in0, &d_l))) return rc;
    if ((rc = stage_out(ctx, w, (size_t)N * sizeof(double), ctx->st_out0, &d_w, &h_w))) return rc;
    if ((rc = ensure(ctx, ctx->partial, (size_t)2 * kRedMaxBlocks * sizeof(double)))) return rc;
    double* part = (double*)ctx->partial.p;
    const int nb = red_blocks(N);
    double* d_ls = log_Sj ? (is_device_ptr(log_Sj) ? log_Sj : ctx->d_small + 2) : nullptr;
    max_kernel<<<nb, kRedThreads, 0, ctx->stream>>>((const double*)d_l, N, delta_beta, part, ctx->d_stage);
    weights_exp_kernel<<<nb, kRedThreads, 0, ctx->stream>>>((const double*)d_l, N, delta_beta, (double*)d_w, part, ctx->d_stage);
    weights_norm_kernel<<<nb, kRedThreads, 0, ctx->stream>>>((double*)d_w, N, ctx->d_stage, d_ls);
    CU(cudaGetLastError());
    ctx->launches += 3;
    bool sync = false;
    if (h_w) {
        CU(cudaMemcpyAsync(h_w, d_w, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (log_Sj && !is_device_ptr(log_Sj)) {
        CU(cudaMemcpyAsync(log_Sj, d_ls, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

// cs = np.cumsum(w) exactly, on the whole GPU (hmx_stage.cuh, K4a multi-CTA); the single-CTA kernel serves small N
static int cumsum_impl(hmx_ctx* ctx, const double* d_w, long long N, double* d_cs) {
    const long long n_tiles = (N + kCumsumTile - 1) / kCumsumTile;
    if (n_tiles < 4 || ctx->cumsum_single_cta) {
        cumsum_serial_kernel<<<1, kRedThreads, 0, ctx->stream>>>(d_w, N, d_cs);
        CU(cudaGetLastError());
        ++ctx->launches;
        return 0;
    }
    int rc;
    if ((rc = ensure(ctx, ctx->cs_info, (size_t)n_tiles * (sizeof(CumsumTileInfo) + sizeof(double))))) return rc;
    CumsumTileInfo* info = (CumsumTileInfo*)ctx->cs_info.p;
    double* tile_sum = (double*)((char*)ctx->cs_info.p + (size_t)n_tiles * sizeof(CumsumTileInfo));
    cumsum_tile_sum_kernel<<<(unsigned)n_tiles, kRedThreads, 0, ctx->stream>>>(d_w, N, tile_sum);
    cumsum_guess_kernel<<<1, kRedThreads, 0, ctx->stream>>>(tile_sum, (int)n_tiles, info);
    cumsum_tile_round_kernel<<<(unsigned)n_tiles, kRedThreads, 0, ctx->stream>>>(d_w, N, info);
    cumsum_carry_kernel<<<1, kRedThreads, 0, ctx->stream>>>(d_w, N, (int)n_tiles, info, d_cs);
    cumsum_tile_write_kernel<<<(unsigned)n_tiles, kRedThreads, 0, ctx->stream>>>(d_w, N, info, d_cs);
    CU(cudaGetLastError());
    ctx->launches += 5;
    return 0;
}

int hmx_cumsum(hmx_ctx* ctx, const double* w, int64_t N, double* cs) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !w || !cs) return fail(ctx, HMX_ERR_INVALID, "hmx_cumsum: bad arguments");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void* d_w;
    void *d_cs, *h_cs;
    if ((rc = stage_in(ctx, w, (size_t)N * sizeof(double), ctx->st_in0, &d_w))) return rc;
    if ((rc = stage_out(ctx, cs, (size_t)N * sizeof(double), ctx->cs, &d_cs, &h_cs))) return rc;
    if ((rc = cumsum_impl(ctx, (const double*)d_w, N, (double*)d_cs))) return rc;
    if (h_cs) {
        CU(cudaMemcpyAsync(h_cs, d_cs, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return HMX_OK;
}

int hmx_resample(hmx_ctx* ctx, const double* w, int64_t N, double u, int64_t* idx) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !w || !idx) return fail(ctx, HMX_ERR_INVALID, "hmx_resample: bad arguments");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void* d_w;
    void *d_idx, *h_idx;
    if ((rc = stage_in(ctx, w, (size_t)N * sizeof(double), ctx->st_in0, &d_w))) return rc;
    if ((rc = stage_out(ctx, idx, (size_t)N * sizeof(int64_t), ctx->st_out0, &d_idx, &h_idx))) return rc;
    if ((rc = ensure(ctx, ctx->cs, (size_t)N * sizeof(double)))) return rc;
    if ((rc = cumsum_impl(ctx, (const double*)d_w, N, (double*)ctx->cs.p))) return rc;
    searchsorted_kernel<<<red_blocks(N), kRedThreads, 0, ctx->stream>>>((const double*)ctx->cs.p, N, u, (long long*)d_idx);
    CU(cudaGetLastError());
    ctx->launches += 1;
    if (h_idx) {
        CU(cudaMemcpyAsync(h_idx, d_idx, (size_t)N * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return HMX_OK;
}

static int moments_impl(hmx_ctx* ctx, const double* d_theta, const double* d_w, long long N, int d, double* d_mom) {
    int rc;
    const int nb = red_blocks(N);
    if ((rc = ensure(ctx, ctx->partial, std::max((size_t)2 * kRedMaxBlocks, (size_t)(2 + d) * nb) * sizeof(double)))) return rc;
    dim3 grid(nb, 2 + d);
    wmoments_kernel<<<grid, kRedThreads, 0, ctx->stream>>>(d_theta, d_w, N, d, (double*)ctx->partial.p, d_mom, ctx->d_tickets);
    CU(cudaGetLastError());
    ++ctx->launches;
    return 0;
}

static int scatter_impl(hmx_ctx* ctx, const double* d_theta, const double* d_w, long long N, int d, const double* d_mean,
                        double* d_scatter) {
    int rc;
    const int nb = (int)std::max(1ll, std::min<long long>((N + kScatterChunk - 1) / kScatterChunk, 4ll * ctx->sm_count));
    if ((rc = ensure(ctx, ctx->partial, std::max((size_t)2 * kRedMaxBlocks, (size_t)nb * d * d) * sizeof(double)))) return rc;
    const size_t smem = (size_t)kScatterChunk * (d + 1) * sizeof(double);
    wscatter_kernel<<<nb, kRedThreads, smem, ctx->stream>>>(d_theta, d_w, N, d, d_mean, (double*)ctx->partial.p, d_scatter,
                                                            ctx->d_tickets + 40);
    CU(cudaGetLastError());
    ++ctx->launches;
    return 0;
}

int hmx_weighted_moments(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d, double* moments) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !theta || !w || !moments || d < 1 || d > 32) return fail(ctx, HMX_ERR_INVALID, "hmx_weighted_moments: bad arguments (1 <= d <= 32)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void *d_t, *d_w;
    if ((rc = stage_in(ctx, theta, (size_t)N * d * sizeof(double), ctx->st_in1, &d_t))) return rc;
    if ((rc = stage_in(ctx, w, (size_t)N * sizeof(double), ctx->st_in0, &d_w))) return rc;
    const bool dev = is_device_ptr(moments);
    double* d_m = dev ? moments : ctx->d_small + 8;
    if ((rc = moments_impl(ctx, (const double*)d_t, (const double*)d_w, N, d, d_m))) return rc;
    if (!dev) {
        CU(cudaMemcpyAsync(moments, d_m, (size_t)(2 + d) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return HMX_OK;
}

int hmx_weighted_scatter(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d, const double* mean,
                         double* scatter) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !theta || !w || !mean || !scatter || d < 1 || d > 32)
        return fail(ctx, HMX_ERR_INVALID, "hmx_weighted_scatter: bad arguments (1 <= d <= 32)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void *d_t, *d_w, *d_mean;
    if ((rc = stage_in(ctx, theta, (size_t)N * d * sizeof(double), ctx->st_in1, &d_t))) return rc;
    if ((rc = stage_in(ctx, w, (size_t)N * sizeof(double), ctx->st_in0, &d_w))) return rc;
    if ((rc = stage_in(ctx, mean, (size_t)d * sizeof(double), ctx->st_in2, &d_mean))) return rc;
    const bool dev = is_device_ptr(scatter);
    double* d_s = dev ? scatter : ctx->d_small + 64;
    if ((rc = scatter_impl(ctx, (const double*)d_t, (const double*)d_w, N, d, (const double*)d_mean, d_s))) return rc;
    if (!dev) {
        CU(cudaMemcpyAsync(scatter, d_s, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return HMX_OK;
}

int hmx_weighted_cov(hmx_ctx* ctx, const double* theta, const double* w, int64_t N, int32_t d, double* mean, double* cov) {
    if (!ctx) return HMX_ERR_INVALID;
    if (N <= 0 || !theta || !w || !cov || d < 1 || d > 32) return fail(ctx, HMX_ERR_INVALID, "hmx_weighted_cov: bad arguments (1 <= d <= 32)");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const void *d_t, *d_w;
    if ((rc = stage_in(ctx, theta, (size_t)N * d * sizeof(double), ctx->st_in1, &d_t))) return rc;
    if ((rc = stage_in(ctx, w, (size_t)N * sizeof(double), ctx->st_in0, &d_w))) return rc;
    double* d_mom = ctx->d_small + 8;      // [2+d]
    double* d_mean = ctx->d_small + 48;    // [d]
    double* d_scat = ctx->d_small + 1100;  // [d*d]
    const bool cov_dev = is_device_ptr(cov);
    double* d_cov = cov_dev ? cov : ctx->d_small + 2200;
    const bool mean_dev = mean && is_device_ptr(mean);
    double* d_mean_out = mean ? (mean_dev ? mean : ctx->d_small + 3300) : nullptr;
    if ((rc = moments_impl(ctx, (const double*)d_t, (const double*)d_w, N, d, d_mom))) return rc;
    wmean_kernel<<<1, 32, 0, ctx->stream>>>(d_mom, d, d_mean);
    if ((rc = scatter_impl(ctx, (const double*)d_t, (const double*)d_w, N, d, d_mean, d_scat))) return rc;
    wcov_final_kernel<<<1, 256, 0, ctx->stream>>>(d_mom, d_scat, d, d_cov, d_mean_out, d_mean);
    CU(cudaGetLastError());
    ctx->launches += 2;
    bool sync = false;
    if (!cov_dev) {
        CU(cudaMemcpyAsync(cov, d_cov, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (mean && !mean_dev) {
        CU(cudaMemcpyAsync(mean, d_mean_out, (size_t)d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        sync = true;
    }
    if (sync) CU(cudaStreamSynchronize(ctx->stream));
    return HMX_OK;
}

int hmx_log_likelihood_sparse(hmx_ctx* ctx, const double* mu, const double* sig, const double* data,
                              const double* sigma_obs, const double* weights, int32_t n_obs, double* logL) {
    if (!ctx) return HMX_ERR_INVALID;
    if (n_obs < 0 || !mu || !sig || !data || !sigma_obs || !logL) return fail(ctx, HMX_ERR_INVALID, "hmx_log_likelihood_sparse: bad arguments");
    CU(cudaSetDevice(ctx->device));
    int rc;
    const size_t bytes = (size_t)n_obs * 5 * sizeof(double);
    // five small inputs share one staging buffer
    if ((rc = ensure(ctx, ctx->st_in2, 5 * bytes + 64))) return rc;
    const double* src[5] = {mu, sig, data, sigma_obs, weights};
    const double* dev[5];
    for (int k = 0; k < 5; ++k) {
        if (!src[k]) {
            dev[k] = nullptr;
        } else if (is_device_ptr(src[k])) {
            dev[k] = src[k];
        } else {
            double* dst = (double*)((char*)ctx->st_in2.p + k * bytes);
            if (bytes) CU(cudaMemcpyAsync(dst, src[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
            dev[k] = dst;
        }
    }
    const bool out_dev = is_device_ptr(logL);
    double* d_out = out_dev ? logL : ctx->d_small + 3;
    loglik_sparse_kernel<<<1, 1, 0, ctx->stream>>>(dev[0], dev[1], dev[2], dev[3], dev[4], n_obs, d_out);
    CU(cudaGetLastError());
    ++ctx->launches;
    if (!out_dev) {
        CU(cudaMemcpyAsync(logL, d_out, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return HMX_OK;
}

int hmx_mutate_groups(hmx_ctx* ctx, int32_t n_groups, const hmx_mutation* muts, const int64_t* n_particles, const double* theta_res,
                      const double* logL_res, double* theta_out, double* logL_out, double* props, double* logL_props,
                      hmx_mutation_stats* stats) {
    if (!ctx) return HMX_ERR_INVALID;
    if (n_groups < 1 || n_groups > kMutMaxGroups || !muts || !n_particles || !theta_res || !logL_res || !theta_out || !logL_out)
        return fail(ctx, HMX_ERR_INVALID, "hmx_mutate_groups: NULL argument or n_groups outside [1, 64]");
    static_assert(HMX_MUT_MAX_DIM == kMutMaxDim && HMX_MUT_MAX_GROUPS == kMutMaxGroups, "header and kernel disagree");
    const int d = muts[0].d;
    long long N = 0, NK = 0;
    std::vector<MutationGroup> host(n_groups);
    for (int g = 0; g < n_groups; ++g) {
        const hmx_mutation& m = muts[g];
        if (m.d < 1 || m.d > HMX_MUT_MAX_DIM || m.K < 1 || m.n_map < 0 || m.n_map > m.d || n_particles[g] < 0)
            return fail(ctx, HMX_ERR_INVALID, "hmx_mutate: need 1 <= d <= 32, K >= 1, 0 <= n_map <= d, N >= 0");
        if (m.d != d) return fail(ctx, HMX_ERR_INVALID, "hmx_mutate_groups: all groups must share the particle dimension d");
        if (m.cond_default < 0 || m.cond_default >= ctx->cfg.n_cond || m.cond_at_lin >= ctx->cfg.n_cond)
            return fail(ctx, HMX_ERR_INVALID, "hmx_mutate: condition id outside [0, n_cond)");
        if (m.sequence >= (1ull << 30)) return fail(ctx, HMX_ERR_INVALID, "hmx_mutate: sequence must be < 2^30");
        make_mutation_params(m, host[g].M);
        host[g].beta = m.beta;
        host[g].part0 = N;
        host[g].n_part = n_particles[g];
        host[g].prop0 = NK;
        N += n_particles[g];
        NK += (long long)n_particles[g] * m.K;
    }
    CU(cudaSetDevice(ctx->device));
    if (stats) memset(stats, 0, sizeof(*stats) * n_groups);
    if (N == 0) return HMX_OK;

    int rc;
    const void *d_res, *d_lres;
    if ((rc = stage_in(ctx, theta_res, (size_t)N * d * sizeof(double), ctx->m_in_theta, &d_res))) return rc;
    if ((rc = stage_in(ctx, logL_res, (size_t)N * sizeof(double), ctx->m_in_logL, &d_lres))) return rc;
    if ((rc = ensure(ctx, ctx->m_groups, host.size() * sizeof(MutationGroup)))) return rc;
    if ((rc = ensure(ctx, ctx->m_props, (size_t)NK * d * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->m_theta, (size_t)NK * HMX_NTHETA * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->m_cond, (size_t)NK * sizeof(int32_t)))) return rc;
    if ((rc = ensure(ctx, ctx->m_inside, (size_t)NK))) return rc;
    if ((rc = ensure(ctx, ctx->m_ll, (size_t)NK * sizeof(double)))) return rc;
    if ((rc = ensure(ctx, ctx->m_status, (size_t)NK * sizeof(uint32_t)))) return rc;
    void *d_tout, *d_lout, *h_tout, *h_lout;
    if ((rc = stage_out(ctx, theta_out, (size_t)N * d * sizeof(double), ctx->m_out_theta, &d_tout, &h_tout))) return rc;
    if ((rc = stage_out(ctx, logL_out, (size_t)N * sizeof(double), ctx->m_out_logL, &d_lout, &h_lout))) return rc;
    CU(cudaMemcpyAsync(ctx->m_groups.p, host.data(), host.size() * sizeof(MutationGroup), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));  // `host` is pageable and dies with this frame
    const MutationGroup* d_groups = (const MutationGroup*)ctx->m_groups.p;

    const int threads = 128;
    const int grid_p = (int)std::min<long long>((NK + threads - 1) / threads, (long long)ctx->sm_count * 16);
    propose_kernel<<<grid_p, threads, 0, ctx->stream>>>(d_groups, n_groups, (const double*)d_res, NK, (double*)ctx->m_props.p,
                                                        (double*)ctx->m_theta.p, (int32_t*)ctx->m_cond.p, (uint8_t*)ctx->m_inside.p);
    CU(cudaGetLastError());
    ctx->launches += 1;
    // K1 + K1c on all proposals of all groups in ONE launch, device-resident (same code path as hmx_loglik_batch)
    if ((rc = batch_impl(ctx, (const double*)ctx->m_theta.p, (const int32_t*)ctx->m_cond.p, NK, (double*)ctx->m_ll.p, nullptr,
                         (uint32_t*)ctx->m_status.p, nullptr, nullptr))) return rc;
    CU(cudaMemsetAsync(ctx->d_mut_counts, 0, 8 * kMutMaxGroups * sizeof(unsigned long long), ctx->stream));
    const int grid_a = (int)std::min<long long>((N + threads - 1) / threads, (long long)ctx->sm_count * 16);
    accept_kernel<<<grid_a, threads, 0, ctx->stream>>>(d_groups, n_groups, (const double*)ctx->m_props.p, (const double*)ctx->m_ll.p,
                                                       (const uint8_t*)ctx->m_inside.p, N, (const double*)d_res, (const double*)d_lres,
                                                       (double*)d_tout, (double*)d_lout, ctx->d_mut_counts);
    status_count_kernel<<<grid_p, threads, 0, ctx->stream>>>(d_groups, n_groups, (const uint32_t*)ctx->m_status.p, NK, ctx->d_mut_counts);
    CU(cudaGetLastError());
    ctx->launches += 2;
    if (h_tout) CU(cudaMemcpyAsync(h_tout, d_tout, (size_t)N * d * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (h_lout) CU(cudaMemcpyAsync(h_lout, d_lout, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (props)
        CU(cudaMemcpyAsync(props, ctx->m_props.p, (size_t)NK * d * sizeof(double),
                           is_device_ptr(props) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    if (logL_props)
        CU(cudaMemcpyAsync(logL_props, ctx->m_ll.p, (size_t)NK * sizeof(double),
                           is_device_ptr(logL_props) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<unsigned long long> h(8 * (size_t)n_groups, 0ull);
    CU(cudaMemcpyAsync(h.data(), ctx->d_mut_counts, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        for (int g = 0; g < n_groups; ++g) {
            stats[g].n_accepted = (int64_t)h[8 * g + 0];
            stats[g].n_candidates = (int64_t)h[8 * g + 1];
            stats[g].n_nonfinite = (int64_t)h[8 * g + 2];
            stats[g].n_maxiter = (int64_t)h[8 * g + 3];
            stats[g].n_singular = (int64_t)h[8 * g + 4];
        }
    }
    return HMX_OK;
}

int hmx_mutate(hmx_ctx* ctx, const hmx_mutation* mut, const double* theta_res, const double* logL_res, int64_t N,
               double* theta_out, double* logL_out, double* props, double* logL_props, hmx_mutation_stats* stats) {
    if (!ctx) return HMX_ERR_INVALID;
    if (!mut || N < 0) return fail(ctx, HMX_ERR_INVALID, "hmx_mutate: NULL argument or N < 0");
    const int64_t n = N;
    return hmx_mutate_groups(ctx, 1, mut, &n, theta_res, logL_res, theta_out, logL_out, props, logL_props, stats);
}

// ---- several GPUs behind ONE handle ------------------------------------------------------------------
int hmx_multi_create(hmx_multi** out, const hmx_config* cfg, const int32_t* device_ids, int32_t n_devices) {
    if (!out) return HMX_ERR_INVALID;
    *out = nullptr;
    if (!cfg || !device_ids || n_devices < 1 || n_devices > 64) {
        g_multi_create_error = "hmx_multi_create: need cfg, device_ids and 1 <= n_devices <= 64";
        return HMX_ERR_INVALID;
    }
    hmx_multi* m = new (std::nothrow) hmx_multi();
    if (!m) return HMX_ERR_OOM;
    for (int k = 0; k < n_devices; ++k) {
        hmx_ctx* c = nullptr;
        const int rc = hmx_create(&c, cfg, device_ids[k]);
        if (rc != HMX_OK) {
            g_multi_create_error = std::string("device ") + std::to_string(device_ids[k]) + ": " + hmx_last_error(nullptr);
            for (hmx_ctx* q : m->ctx) hmx_destroy(q);
            delete m;
            return rc;
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return HMX_OK;
}

void hmx_multi_destroy(hmx_multi* m) {
    if (!m) return;
    for (hmx_ctx* c : m->ctx) hmx_destroy(c);
    delete m;
}

const char* hmx_multi_last_error(const hmx_multi* m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }
int32_t hmx_multi_device_count(const hmx_multi* m) { return m ? (int32_t)m->ctx.size() : 0; }
hmx_ctx* hmx_multi_context(hmx_multi* m, int32_t k) { return (m && k >= 0 && k < (int32_t)m->ctx.size()) ? m->ctx[k] : nullptr; }

int hmx_multi_loglik_batch(hmx_multi* m, const double* theta, const int32_t* cond_id, int64_t N, double* logL, uint32_t* status,
                           uint32_t* newton_iters) {
    if (!m) return HMX_ERR_INVALID;
    if (N < 0 || (N > 0 && (!theta || !logL))) {
        m->err = "hmx_multi_loglik_batch: NULL argument or N < 0";
        return HMX_ERR_INVALID;
    }
    if (is_device_ptr(theta) || is_device_ptr(logL)) {
        m->err = "hmx_multi_loglik_batch shards HOST arrays; device-resident callers run one context (one process) per GPU";
        return HMX_ERR_INVALID;
    }
    return multi_run(m, N, [&](int k, long long lo, long long hi) {
        return hmx_loglik_batch(m->ctx[k], theta + lo * HMX_NTHETA, cond_id ? cond_id + lo : nullptr, hi - lo, logL + lo,
                                nullptr, status ? status + lo : nullptr, newton_iters ? newton_iters + lo : nullptr);
    });
}

int hmx_multi_weighted_cov(hmx_multi* m, const double* theta, const double* w, int64_t N, int32_t d, double* mean, double* cov) {
    if (!m) return HMX_ERR_INVALID;
    if (N <= 0 || !theta || !w || !cov || d < 1 || d > 32) {
        m->err = "hmx_multi_weighted_cov: bad arguments (1 <= d <= 32)";
        return HMX_ERR_INVALID;
    }
    const int G = (int)m->ctx.size();
    std::vector<double> mom((size_t)G * (2 + d), 0.0), sc((size_t)G * d * d, 0.0);
    int rc = multi_run(m, N, [&](int k, long long lo, long long hi) {
        return hmx_weighted_moments(m->ctx[k], theta + lo * d, w + lo, hi - lo, d, mom.data() + (size_t)k * (2 + d));
    });
    if (rc) return rc;
    // "all-reduce" of the shard moments in device order (deterministic), then the centred scatter per shard
    std::vector<double> tot(2 + d, 0.0), mu(d, 0.0);
    for (int k = 0; k < G; ++k)
        for (int j = 0; j < 2 + d; ++j) tot[j] += mom[(size_t)k * (2 + d) + j];
    for (int j = 0; j < d; ++j) mu[j] = tot[2 + j] / tot[0];
    rc = multi_run(m, N, [&](int k, long long lo, long long hi) {
        return hmx_weighted_scatter(m->ctx[k], theta + lo * d, w + lo, hi - lo, d, mu.data(), sc.data() + (size_t)k * d * d);
    });
    if (rc) return rc;
    const double fact = tot[0] - tot[1] / tot[0];
    for (int e = 0; e < d * d; ++e) {
        double s = 0.0;
        for (int k = 0; k < G; ++k) s += sc[(size_t)k * d * d + e];
        cov[e] = s * (1.0 / fact);
    }
    if (mean)
        for (int j = 0; j < d; ++j) mean[j] = mu[j];
    return HMX_OK;
}

int hmx_set_latency_kernel(hmx_ctx* ctx, int64_t max_evals) {
    if (!ctx) return HMX_ERR_INVALID;
    // < 0: the measured sweet spot, one resident round of 8-lane groups (7104 evaluations on a B200)
    ctx->coop_max_evals = max_evals < 0 ? 1ll * ctx->sm_count * ctx->coop_blocks_per_sm * (kCoopThreads / kCoopLanes) : max_evals;
    return HMX_OK;
}

int hmx_set_timing(hmx_ctx* ctx, int32_t enabled) {
    if (!ctx) return HMX_ERR_INVALID;
    ctx->timing = enabled != 0;
    return HMX_OK;
}

int hmx_get_counters(hmx_ctx* ctx, hmx_counters* out) {
    if (!ctx || !out) return HMX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    unsigned long long h[4] = {0, 0, 0, 0};
    CU(cudaMemcpy(h, ctx->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    if (ctx->timing_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) ctx->last_ode_ms = ctx->ode_ms_acc + ms;
        else cudaGetLastError();
        ctx->ode_ms_acc = 0.0;
        ctx->timing_pending = false;
    }
    out->evals = ctx->last_evals;
    out->newton_iters = (int64_t)h[1];
    out->q_evals = (int64_t)h[2];
    out->kernel_launches = ctx->launches;
    out->last_ode_kernel_ms = ctx->last_ode_ms;
    return HMX_OK;
}

int hmx_measure_fp64_peak(hmx_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return HMX_ERR_INVALID;
    CU(cudaSetDevice(ctx->device));
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 4096;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        CU(cudaEventRecord(a, ctx->stream));
        fp64_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(ctx->d_small, iters, 0.999999, 1e-6);
        CU(cudaEventRecord(b, ctx->stream));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * (double)blocks * threads * (double)iters * kPeakFmaPerIter;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    ctx->launches += 4;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return HMX_OK;
}

}  // extern "C"
