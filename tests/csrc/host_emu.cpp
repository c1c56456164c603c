// host_emu.cpp -- CPU emulation of ONE GPU lane, for tests only.
//
// Compiles the product's __host__ __device__ arithmetic (tmcmc202601_b200/csrc/hmx_{model,lane,obs}.cuh)
// with g++ so the structured direction solve and the lane state machine can be compared with the dense
// oracle on machines without a GPU (`pytest -m "not gpu"`).  It is NOT part of libhamilton.so and is never
// loaded by the product package: there is no CPU fallback on the product path.
#include <stdint.h>
#include <string.h>

#include "../../tmcmc202601_b200/csrc/hmx_obs.cuh"
#include "../../tmcmc202601_b200/csrc/hmx_mutation.cuh"
#include "../../tmcmc202601_b200/csrc/hmx_grad.cuh"
#include "../../tmcmc202601_b200/csrc/hmx_dense.cuh"

using namespace hmx;

template <bool ALPHA, int HILL>
static void run_lane(const SolverConsts& K, const CondConsts& C, const ThetaMap& TM, const double* theta,
                     double* g_arr, double* logL, uint32_t* status, uint32_t* iters, uint32_t* trips) {
    Lane s;
    lane_init(K, s, theta, C.g0);
    if (g_arr) memcpy(g_arr, s.g, 12 * sizeof(double));
    double pairs[HMX_MAX_OBS * kPairStride];
    memset(pairs, 0, sizeof(pairs));
    auto step_end = [&](const double (&g)[12], const LaneGP& gp, int step_idx) {  // g = g_arr[step_idx + 1], gp = g_arr[step_idx]
        const int row = step_idx + 1;
        if (g_arr) memcpy(g_arr + (size_t)row * 12, g, 12 * sizeof(double));
        for (int o = 0; o < C.n_obs; ++o) {
            const int t = C.idx[o] < 1 ? 1 : C.idx[o];
            if (t == row) {
                memcpy(pairs + o * kPairStride, g, 12 * sizeof(double));
                for (int i = 0; i < 11; ++i) pairs[o * kPairStride + 12 + i] = gp[i];
            }
        }
    };
    while (!lane_trip<ALPHA, HILL>(K, s, step_end)) {
    }
    uint32_t st = s.status;
    if (logL) *logL = obs_loglik<ALPHA, HILL>(K, C, TM, theta, pairs, s.status, nullptr, &st);
    if (status) *status = st;
    if (iters) *iters = s.iters;
    if (trips) *trips = s.trips;
}

template <bool ALPHA, int HILL>
static void dir_once(const SolverConsts& K, const double* theta, const double* g_new, const double* g_old,
                     double* Q, double* delta) {
    double A[15], b[5], x[12], gp[11];
    theta_to_Ab(theta, A, b);
    for (int i = 0; i < 12; ++i) x[i] = g_new[i];
    for (int i = 0; i < 11; ++i) gp[i] = g_old[i];
    Eval e;
    residual<ALPHA, HILL>(K, x, gp, A, b, e);
    for (int i = 0; i < 12; ++i) Q[i] = e.Q[i];
    double d[12];
    direction<ALPHA, HILL>(K, e, A, b, d);
    for (int i = 0; i < 12; ++i) delta[i] = d[i];
}

// rebase_residual() against a fresh residual: x is the accepted trial that ended a step whose previous level
// was g_old; the next step's first residual is Q(clip(x); previous level = x).
template <bool ALPHA, int HILL>
static void rebase_once(const SolverConsts& K, const double* theta, const double* x_in, const double* g_old,
                        double* Q_rebased, double* Q_direct, double* norms, int* finite) {
    Lane s;
    double x[12], xc[12];
    theta_to_Ab(theta, s.A, s.b);
    for (int i = 0; i < 12; ++i) x[i] = x_in[i];
    for (int i = 0; i < 11; ++i) s.gp[i] = g_old[i];
    Eval e;
    residual<ALPHA, HILL>(K, x, s.gp, s.A, s.b, e);
    const double gp5_old = s.gp[5];
    for (int i = 0; i < 11; ++i) s.gp[i] = x[i];
    *finite = rebase_residual(K, s, x, gp5_old, e) ? 1 : 0;
    for (int i = 0; i < 12; ++i) Q_rebased[i] = e.Q[i];
    norms[0] = e.norm;
    // what the ordinary path evaluates on the next trip: iterate = x, previous level = x
    for (int i = 0; i < 12; ++i) xc[i] = x[i];
    Eval f;
    residual<ALPHA, HILL>(K, xc, s.gp, s.A, s.b, f);
    for (int i = 0; i < 12; ++i) Q_direct[i] = f.Q[i];
    norms[1] = f.norm;
    // the direction solve must see identical inputs apart from Q
    double dev = 0.0;
    for (int i = 0; i < 5; ++i) {
        dev = fmax(dev, fabs(e.phd[i] - f.phd[i]));
        dev = fmax(dev, fabs(e.psd[i] - f.psd[i]));
        dev = fmax(dev, fabs(e.ph[i] - f.ph[i]) + fabs(e.ps[i] - f.ps[i]) + fabs(e.I[i] - f.I[i]));
        dev = fmax(dev, fabs(e.phip[i] - f.phip[i]) + fabs(e.psip[i] - f.psip[i]));
    }
    norms[2] = dev + fabs(e.K55 - f.K55) + fabs(e.h - f.h);
}

#define DISPATCH(FN, ...)                                   \
    do {                                                    \
        const bool al = (K.alpha != 0.0);                   \
        const int hv = hill_variant(K);                     \
        if (al) {                                           \
            if (hv == 0) FN<true, 0>(__VA_ARGS__);          \
            else if (hv == 1) FN<true, 1>(__VA_ARGS__);     \
            else FN<true, 2>(__VA_ARGS__);                  \
        } else {                                            \
            if (hv == 0) FN<false, 0>(__VA_ARGS__);         \
            else if (hv == 1) FN<false, 1>(__VA_ARGS__);    \
            else FN<false, 2>(__VA_ARGS__);                 \
        }                                                   \
    } while (0)

// forward-sensitivity recursion of hmx_grad.cuh over a stored trajectory g_arr [(n_steps + 1), 12]: S_out [n_rows,12] at `rows`
template <bool ALPHA, int HILL>
static void emu_sens(const SolverConsts& K, const double* theta, const double* g_arr, int k, const int32_t* rows, int n_rows, double* S_out) {
    ThetaMap tm = {HMX_THETA_P, HMX_THETA_Q};
    double A[15], b[5], S[12], x[12], gp[11];
    theta_to_Ab(theta, A, b);
    for (int i = 0; i < 12; ++i) S[i] = 0.0;
    int nr = 0;
    for (int t = 1; t <= K.n_steps && nr < n_rows; ++t) {
        for (int i = 0; i < 12; ++i) x[i] = g_arr[(size_t)t * 12 + i];
        for (int i = 0; i < 11; ++i) gp[i] = g_arr[(size_t)(t - 1) * 12 + i];
        sensitivity_step<ALPHA, HILL>(K, x, gp, A, b, tm.p[k], tm.q[k], S);
        while (nr < n_rows && rows[nr] == t) {
            memcpy(S_out + (size_t)nr * 12, S, sizeof(S));
            ++nr;
        }
    }
}

template <bool ALPHA, int HILL>
static void run_reference_loop(const SolverConsts& K, const CondConsts& C, const double* theta, double* g_arr, uint32_t* status, uint32_t* iters) {
    uint32_t st = 0, it = 0, tr = 0;
    memcpy(g_arr, C.g0, 12 * sizeof(double));
    reference_loop_solve<ALPHA, HILL>(K, theta, C.g0, st, it, tr, [&](const double (&g)[12], const RegArr<11>&, int step) {
        memcpy(g_arr + (size_t)(step + 1) * 12, g, 12 * sizeof(double));
    });
    *status = st;
    *iters = it;
}

// per-trip event trace of one lane (scripts/trip_sim.py): bit0 = a direction was solved, bit1 = the time step ended
template <bool ALPHA, int HILL>
static void trace_lane(const SolverConsts& K, const CondConsts& C, const double* theta, uint8_t* ev, long long cap, long long* n_out) {
    Lane s;
    lane_init(K, s, theta, C.g0);
    auto step_end = [&](const double (&)[12], const LaneGP&, int) {};
    long long n = 0;
    bool done = false;
    while (!done) {
        const uint32_t it0 = s.iters;
        const int32_t st0 = s.step_idx;
        done = lane_trip<ALPHA, HILL>(K, s, step_end);
        if (n < cap) ev[n] = (uint8_t)((s.iters != it0 ? 1 : 0) | (s.step_idx != st0 ? 2 : 0));
        ++n;
    }
    *n_out = n;
}

// tsm_row of hmx_obs.cuh (factor once, one solve per active parameter) on one trajectory row
template <bool ALPHA, int HILL>
static void tsm_row_once(const SolverConsts& K, const CondConsts& C, const double* theta, const double* g_new, const double* g_old,
                         double* sig2_out, double* x1_out, int* ok) {
    ThetaMap tm = {HMX_THETA_P, HMX_THETA_Q};
    double A[15], b[5], x[12], gp[11], sig2[12], covp[5];
    theta_to_Ab(theta, A, b);
    for (int i = 0; i < 12; ++i) x[i] = g_new[i];
    for (int i = 0; i < 11; ++i) gp[i] = g_old[i];
    for (int i = 0; i < 12; ++i) sig2[i] = 0.0;
    for (int i = 0; i < 5; ++i) covp[i] = 0.0;
    for (int i = 0; i < 12 * HMX_NTHETA; ++i) x1_out[i] = 0.0;
    *ok = tsm_row<ALPHA, HILL>(K, C, tm, theta, A, b, x, gp, sig2, covp, x1_out) ? 1 : 0;
    for (int i = 0; i < 12; ++i) sig2_out[i] = sig2[i];
}

extern "C" {

// sigma2[12] and x1[12,20] of one trajectory row through the product's tsm_row (hmx_obs.cuh)
int emu_tsm_row(const hmx_config* cfg, int cond, const double* theta, const double* g_new, const double* g_old, double* sig2,
                double* x1) {
    SolverConsts K;
    CondConsts C;
    make_solver_consts(*cfg, K);
    make_cond_consts(*cfg, cond, C);
    int ok = 0;
    DISPATCH(tsm_row_once, K, C, theta, g_new, g_old, sig2, x1, &ok);
    return ok;
}


// hmx_dense.cuh on the CPU: the reference's loop nest with the dense pivoted solve + ridge retry; g_arr [(n_steps + 1), 12]
int emu_reference_loop(const hmx_config* cfg, int cond, const double* theta, double* g_arr, uint32_t* status, uint32_t* iters) {
    SolverConsts K;
    CondConsts C;
    make_solver_consts(*cfg, K);
    make_cond_consts(*cfg, cond, C);
    const bool alpha = cfg->alpha_const != 0.0;
    const int hill = hill_variant(K);
    if (alpha) {
        if (hill == 0) run_reference_loop<true, 0>(K, C, theta, g_arr, status, iters);
        else if (hill == 1) run_reference_loop<true, 1>(K, C, theta, g_arr, status, iters);
        else run_reference_loop<true, 2>(K, C, theta, g_arr, status, iters);
    } else {
        if (hill == 0) run_reference_loop<false, 0>(K, C, theta, g_arr, status, iters);
        else if (hill == 1) run_reference_loop<false, 1>(K, C, theta, g_arr, status, iters);
        else run_reference_loop<false, 2>(K, C, theta, g_arr, status, iters);
    }
    return 0;
}


// the generic NS-species solver (hmx_legacy.cuh) on the CPU: g_arr [(n_steps + 1), 2 NS + 2]
int emu_legacy_solve(const hmx_legacy_config* cfg, const double* theta, double* g_arr, uint32_t* status, uint32_t* iters) {
    LegacyConsts K;
    make_legacy_consts(*cfg, K);
    uint32_t st = 0, it = 0;
    if (cfg->n_species == 4) {
        legacy_solve<4>(K, theta, st, it, [&](int level, const double (&g)[10]) { memcpy(g_arr + (size_t)level * 10, g, sizeof(g)); });
    } else if (cfg->n_species == 5) {
        legacy_solve<5>(K, theta, st, it, [&](int level, const double (&g)[12]) { memcpy(g_arr + (size_t)level * 12, g, sizeof(g)); });
    } else {
        return -1;
    }
    *status = st;
    *iters = it;
    return 0;
}
int emu_sensitivities(const hmx_config* cfg, const double* theta, const double* g_arr, int k, const int32_t* rows, int n_rows, double* S_out) {
    SolverConsts K;
    make_solver_consts(*cfg, K);
    const bool alpha = cfg->alpha_const != 0.0;
    const int hill = hill_variant(K);
    if (alpha) {
        if (hill == 0) emu_sens<true, 0>(K, theta, g_arr, k, rows, n_rows, S_out);
        else if (hill == 1) emu_sens<true, 1>(K, theta, g_arr, k, rows, n_rows, S_out);
        else emu_sens<true, 2>(K, theta, g_arr, k, rows, n_rows, S_out);
    } else {
        if (hill == 0) emu_sens<false, 0>(K, theta, g_arr, k, rows, n_rows, S_out);
        else if (hill == 1) emu_sens<false, 1>(K, theta, g_arr, k, rows, n_rows, S_out);
        else emu_sens<false, 2>(K, theta, g_arr, k, rows, n_rows, S_out);
    }
    return 0;
}
int emu_sizeof_legacy_config(void) { return (int)sizeof(hmx_legacy_config); }


int emu_trip_trace(const hmx_config* cfg, int cond, const double* theta, uint8_t* ev, long long cap, long long* n_out) {
    SolverConsts K;
    CondConsts C;
    make_solver_consts(*cfg, K);
    make_cond_consts(*cfg, cond, C);
    DISPATCH(trace_lane, K, C, theta, ev, cap, n_out);
    return 0;
}

// full solve of one (theta, condition): trajectory (optional), logL, counters
int emu_evaluate(const hmx_config* cfg, int cond, const double* theta, double* g_arr, double* logL,
                 uint32_t* status, uint32_t* iters, uint32_t* trips) {
    const char* why;
    if (validate_config(cfg, &why) != 0) return -1;
    SolverConsts K;
    CondConsts C;
    make_solver_consts(*cfg, K);
    make_cond_consts(*cfg, cond, C);
    ThetaMap TM = {HMX_THETA_P, HMX_THETA_Q};
    DISPATCH(run_lane, K, C, TM, theta, g_arr, logL, status, iters, trips);
    return 0;
}

int emu_rebase(const hmx_config* cfg, const double* theta, const double* x, const double* g_old,
               double* Q_rebased, double* Q_direct, double* norms, int* finite) {
    SolverConsts K;
    make_solver_consts(*cfg, K);
    DISPATCH(rebase_once, K, theta, x, g_old, Q_rebased, Q_direct, norms, finite);
    return 0;
}

// ---- device-side mutation (hmx_mutation.cuh) run serially on the host ------------------------------------------
void emu_philox(const uint32_t* c, const uint32_t* k, uint32_t* out) {
    const Philox4 p = philox4x32_10(c[0], c[1], c[2], c[3], k[0], k[1]);
    for (int i = 0; i < 4; ++i) out[i] = p.v[i];
}
double emu_reflect(double x, double lo, double hi) { return reflect_into_bounds(x, lo, hi); }
void emu_normals(uint64_t seed, uint32_t seq, uint64_t gidx, int n_pairs, double* z) {
    MutationRng R = {(uint32_t)seed, (uint32_t)(seed >> 32), seq};
    for (int j = 0; j < n_pairs; ++j) normal_pair(R, gidx, (uint32_t)j, z[2 * j], z[2 * j + 1]);
}
double emu_accept_uniform(uint64_t seed, uint32_t seq, uint64_t gidx) {
    MutationRng R = {(uint32_t)seed, (uint32_t)(seed >> 32), seq};
    return accept_uniform(R, gidx);
}
// phase 1 of hmx_mutate: proposals, full 20-vectors, condition ids, inside flags
int emu_mutation_propose(const hmx_mutation* mut, const double* theta_res, long long N, double* props, double* theta_full,
                         int32_t* cond_id, uint8_t* inside) {
    MutationParams M;
    make_mutation_params(*mut, M);
    for (long long t = 0; t < N * M.K; ++t) {
        const long long i = t / M.K;
        double prop[kMutMaxDim], full[HMX_NTHETA];
        const uint64_t gidx = (uint64_t)(M.index_offset * (unsigned long long)M.K) + (uint64_t)t;
        const bool in = propose_one(M, theta_res + i * M.d, gidx, prop);
        for (int j = 0; j < M.d; ++j) props[t * M.d + j] = prop[j];
        cond_id[t] = expand_theta(M, prop, full);
        for (int k = 0; k < HMX_NTHETA; ++k) theta_full[t * HMX_NTHETA + k] = full[k];
        inside[t] = in ? 1 : 0;
    }
    return 0;
}
// phase 3: sequential accept / reject; counts[0] += accepted, counts[1] += candidates
int emu_mutation_accept(const hmx_mutation* mut, const double* props, const double* ll, const uint8_t* inside, long long N,
                        const double* theta_res, const double* logL_res, double* theta_out, double* logL_out, long long* counts) {
    MutationParams M;
    make_mutation_params(*mut, M);
    for (long long i = 0; i < N; ++i) {
        double cur[kMutMaxDim];
        for (int j = 0; j < M.d; ++j) cur[j] = theta_res[i * M.d + j];
        double cl = logL_res[i];
        int acc = 0, cand = 0;
        accept_chain(M, mut->beta, (uint64_t)(M.index_offset + (unsigned long long)i), props + (size_t)i * M.K * M.d, ll + i * M.K,
                     inside + i * M.K, cur, cl, acc, cand);
        for (int j = 0; j < M.d; ++j) theta_out[i * M.d + j] = cur[j];
        logL_out[i] = cl;
        counts[0] += acc;
        counts[1] += cand;
    }
    return 0;
}
int emu_sizeof_mutation(void) { return (int)sizeof(hmx_mutation); }
int emu_sizeof_mutation_stats(void) { return (int)sizeof(hmx_mutation_stats); }

// struct layout checks for the ctypes mirror (tmcmc202601_b200/_abi.py)
int emu_sizeof_config(void) { return (int)sizeof(hmx_config); }
int emu_sizeof_cond(void) { return (int)sizeof(hmx_cond); }
int emu_sizeof_counters(void) { return (int)sizeof(hmx_counters); }
int emu_validate(const hmx_config* cfg) {
    const char* why;
    return validate_config(cfg, &why);
}

// the two 5x5 solvers of hmx_model.cuh on a caller-supplied system M[5][7]; returns the smallest pivot
double emu_solve5x2(const double* M, double* X, int pivoted) {
    if (pivoted == 1) {
        solve5x2_pivoted(M, X);
        return 0.0;
    }
    double S[5][7];
    for (int r = 0; r < 5; ++r)
        for (int c = 0; c < 7; ++c) S[r][c] = M[r * 7 + c];
    const double pmin = (pivoted == 2) ? solve5x2_block23(S) : (solve5x2_natural(S) ? 1.0 : 0.0);
    for (int r = 0; r < 5; ++r) {
        X[r] = S[r][5];
        X[5 + r] = S[r][6];
    }
    return pmin;
}

// residual and structured quasi-Newton direction at one point
int emu_direction(const hmx_config* cfg, const double* theta, const double* g_new, const double* g_old,
                  double* Q, double* delta) {
    SolverConsts K;
    make_solver_consts(*cfg, K);
    DISPATCH(dir_once, K, theta, g_new, g_old, Q, delta);
    return 0;
}
}
