// hmx_mutation.cuh -- device-side MH mutation of a TMCMC stage (reference TM:849-962).
//
// The reference generates, for every resampled particle, K Gaussian proposals around the RESAMPLED state
// (intermediate accepts are ignored, TM:871-906), reflects them into the prior box (TM:201-226), evaluates the
// likelihood of the ones inside the prior, and then accepts them sequentially against the evolving current
// state with log u < (lp' + beta logL') - (lp + beta logL) (TM:908-962).  Here the three phases are
//
//   K6 propose_kernel   one thread per (particle, step): Philox normals -> mean + F z -> reflection -> prior
//                       check -> the FULL 20-vector the ODE kernel consumes (theta_sub -> theta_full mapping of
//                       EV:635-637 and the np.allclose short-cut of TSMA:366-401 as a condition id)
//   K1  ode_kernel + K1c obs_kernel on the N K proposals (unchanged)
//   K7 accept_kernel    one thread per particle: K sequential accept/reject decisions, acceptance counters
//
// Random numbers: Philox4x32-10 (Salmon et al., SC'11), counter = (proposal index lo, hi, block, tag),
// key = seed.  The proposal index is GLOBAL (index_offset + i) K + k, so a population sharded over several
// GPUs draws exactly the numbers a single GPU would: results do not depend on the world size.  Parity with
// the reference is statistical (NumPy's PCG64 stream is not reproduced); the host path in tmcmc.py remains
// the bitwise-RNG mirror.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "hmx_model.cuh"

namespace hmx {

constexpr int kMutMaxDim = 32;

struct Philox4 {
    uint32_t v[4];
};

HMX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.v[0] = c0;
    o.v[1] = c1;
    o.v[2] = c2;
    o.v[3] = c3;
    return o;
}

// 52 random bits -> (0, 1), both ends excluded, every value exactly representable
HMX_HD double uniform52(uint32_t a, uint32_t b) {
    const uint64_t x = ((uint64_t)(a >> 6) << 26) | (uint64_t)(b >> 6);
    return ((double)x + 0.5) * 2.220446049250313e-16;  // 2^-52
}

enum : uint32_t { kMutTagNormal = 0u, kMutTagAccept = 1u };

struct MutationRng {
    uint32_t k0, k1;   // seed
    uint32_t seq;      // stage / attempt number chosen by the caller
};

// two standard normals for (proposal gidx, block j)
HMX_HD void normal_pair(const MutationRng& R, uint64_t gidx, uint32_t block, double& z0, double& z1) {
    const Philox4 p = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), block, (R.seq << 2) | kMutTagNormal, R.k0, R.k1);
    const double u1 = uniform52(p.v[0], p.v[1]);
    const double u2 = uniform52(p.v[2], p.v[3]);
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
#if defined(__CUDA_ARCH__)
    sincospi(2.0 * u2, &s, &c);
#else
    s = sin(6.283185307179586 * u2);
    c = cos(6.283185307179586 * u2);
#endif
    z0 = r * c;
    z1 = r * s;
}

HMX_HD double accept_uniform(const MutationRng& R, uint64_t gidx) {
    const Philox4 p = philox4x32_10((uint32_t)gidx, (uint32_t)(gidx >> 32), 0u, (R.seq << 2) | kMutTagAccept, R.k0, R.k1);
    return uniform52(p.v[0], p.v[1]);
}

// TM:201-226 with NumPy's modulo (result takes the sign of the divisor); width <= 0 clips
HMX_HD double reflect_into_bounds(double x, double low, double high) {
    const double width = high - low;
    if (!(width > 0.0)) return (x < low) ? low : ((x > high) ? high : x);
    double y = x - low;
    const double two_w = 2.0 * width;
    if (y < 0.0 || y >= two_w) {
        double r = fmod(y, two_w);
        if (r != 0.0 && r < 0.0) r += two_w;
        y = r;
    }
    if (y > width) y = two_w - y;
    return low + y;
}

// Everything the proposal kernel needs besides the particles (passed by value as a kernel parameter).
struct MutationParams {
    double factor[kMutMaxDim * kMutMaxDim];  // row-major [d,d]: proposal = mean + factor z
    double low[kMutMaxDim], high[kMutMaxDim];
    double theta_base[HMX_NTHETA];           // full 20-vector the sub-vector is scattered into (EV:631-637)
    double theta_lin[HMX_NTHETA];            // linearization point of the np.allclose short-cut
    int32_t map[kMutMaxDim];                 // theta_full[map[j]] = theta_sub[j] for j < n_map; entries >= 20 are dropped
    int32_t n_map;
    int32_t d, K;
    int32_t cond_default, cond_at_lin;       // condition ids (cond_at_lin < 0: short-cut off)
    MutationRng rng;
    unsigned long long index_offset;         // global index of local particle 0
};

// One proposal.  Writes prop[d]; returns true when it lies inside the prior box (log prior finite).
HMX_HD bool propose_one(const MutationParams& M, const double* mean, uint64_t gidx, double* prop) {
    double z[kMutMaxDim];
    const int d = M.d;
    for (int j = 0; j < d; j += 2) {
        double a, b;
        normal_pair(M.rng, gidx, (uint32_t)(j >> 1), a, b);
        z[j] = a;
        if (j + 1 < d) z[j + 1] = b;
    }
    bool inside = true;
    for (int r = 0; r < d; ++r) {
        double s = 0.0;
        for (int c = 0; c < d; ++c) s = fma(M.factor[r * d + c], z[c], s);
        const double v = reflect_into_bounds(mean[r] + s, M.low[r], M.high[r]);
        prop[r] = v;
        inside = inside && (v >= M.low[r]) && (v <= M.high[r]);
    }
    return inside;
}

// theta_sub -> the 20-vector of the ODE kernel + the condition id of the np.allclose short-cut
HMX_HD int32_t expand_theta(const MutationParams& M, const double* prop, double* full) {
    for (int k = 0; k < HMX_NTHETA; ++k) full[k] = M.theta_base[k];
    for (int j = 0; j < M.n_map; ++j) {
        const int t = M.map[j];
        if (t >= 0 && t < HMX_NTHETA) full[t] = prop[j];
    }
    if (M.cond_at_lin < 0) return M.cond_default;
    bool at = true;
    for (int k = 0; k < HMX_NTHETA; ++k) at = at && (fabs(full[k] - M.theta_lin[k]) <= 1e-8 + 1e-5 * fabs(M.theta_lin[k]));
    return at ? M.cond_at_lin : M.cond_default;
}

// K sequential accept/reject decisions of one particle.  cur[d], cur_logL are updated in place.
// Proposals outside the prior were never candidates (TM:915-929); a NaN likelihood draws no uniform and is
// rejected (TM:931-939); the -1e20 sentinel is finite and loses against any real value.
HMX_HD void accept_chain(const MutationParams& M, double beta, uint64_t particle, const double* props, const double* ll,
                         const uint8_t* inside, double* cur, double& cur_logL, int& n_acc, int& n_cand) {
    const int d = M.d, K = M.K;
    for (int k = 0; k < K; ++k) {
        if (!inside[k]) continue;
        ++n_cand;
        const double lp = ll[k];
        if (!(lp == lp) || fabs(lp) > 1.7976931348623157e308) continue;
        const double u = accept_uniform(M.rng, particle * (uint64_t)K + (uint64_t)k);
        const double log_ratio = beta * lp - beta * cur_logL;  // uniform prior: lp' = lp = 0 inside the box
        if (log(u) < log_ratio) {
            for (int j = 0; j < d; ++j) cur[j] = props[(size_t)k * d + j];
            cur_logL = lp;
            ++n_acc;
        }
    }
}

// C-ABI struct (include/hmx.h) -> kernel parameters
inline void make_mutation_params(const hmx_mutation& m, MutationParams& M) {
    memset(&M, 0, sizeof(M));
    const int d = m.d;
    memcpy(M.factor, m.factor, sizeof(double) * (size_t)d * d);
    memcpy(M.low, m.low, sizeof(double) * d);
    memcpy(M.high, m.high, sizeof(double) * d);
    memcpy(M.theta_base, m.theta_base, sizeof(M.theta_base));
    memcpy(M.theta_lin, m.theta_lin, sizeof(M.theta_lin));
    memcpy(M.map, m.map, sizeof(int32_t) * d);
    M.n_map = m.n_map;
    M.d = d;
    M.K = m.K;
    M.cond_default = m.cond_default;
    M.cond_at_lin = m.cond_at_lin;
    M.rng.k0 = (uint32_t)m.seed;
    M.rng.k1 = (uint32_t)(m.seed >> 32);
    M.rng.seq = (uint32_t)m.sequence;
    M.index_offset = m.index_offset;
}

// One entry per independent population (chain / condition) handled by a launch: its parameters and where its
// particles and proposals sit in the concatenated arrays.
constexpr int kMutMaxGroups = 64;
struct MutationGroup {
    MutationParams M;
    double beta;
    long long part0, n_part;  // particles [part0, part0 + n_part)
    long long prop0;          // proposals [prop0, prop0 + n_part * M.K)
};

// group of a flat index given the ascending start offsets (n_groups <= 64: a linear scan over shared memory)
HMX_HD int group_of(const long long* start, int n_groups, long long idx) {
    int g = 0;
    while (g + 1 < n_groups && idx >= start[g + 1]) ++g;
    return g;
}

#if defined(__CUDACC__)
__global__ void propose_kernel(const MutationGroup* __restrict__ groups, int n_groups, const double* __restrict__ theta_res,
                               long long total, double* __restrict__ props, double* __restrict__ theta_full,
                               int32_t* __restrict__ cond_id, uint8_t* __restrict__ inside) {
    __shared__ long long start[kMutMaxGroups];
    if (threadIdx.x < n_groups) start[threadIdx.x] = groups[threadIdx.x].prop0;
    __syncthreads();
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const MutationGroup& G = groups[group_of(start, n_groups, t)];
        const MutationParams& M = G.M;
        const long long tl = t - G.prop0;
        const long long i = G.part0 + tl / M.K;
        double mean[kMutMaxDim], prop[kMutMaxDim], full[HMX_NTHETA];
        for (int j = 0; j < M.d; ++j) mean[j] = theta_res[i * M.d + j];
        const uint64_t gidx = (uint64_t)(M.index_offset * (unsigned long long)M.K) + (uint64_t)tl;
        const bool in = propose_one(M, mean, gidx, prop);
        for (int j = 0; j < M.d; ++j) props[t * M.d + j] = prop[j];
        const int32_t c = expand_theta(M, prop, full);
        for (int k = 0; k < HMX_NTHETA; ++k) theta_full[t * HMX_NTHETA + k] = full[k];
        cond_id[t] = c;
        inside[t] = in ? 1 : 0;
    }
}

// counts: [n_groups][8] = accepted, candidates, nonfinite, maxiter, singular
__global__ void accept_kernel(const MutationGroup* __restrict__ groups, int n_groups, const double* __restrict__ props,
                              const double* __restrict__ ll, const uint8_t* __restrict__ inside, long long N,
                              const double* __restrict__ theta_res, const double* __restrict__ logL_res,
                              double* __restrict__ theta_out, double* __restrict__ logL_out, unsigned long long* counts) {
    __shared__ long long start[kMutMaxGroups];
    if (threadIdx.x < n_groups) start[threadIdx.x] = groups[threadIdx.x].part0;
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const int g = group_of(start, n_groups, i);
        const MutationGroup& G = groups[g];
        const MutationParams& M = G.M;
        const long long il = i - G.part0;
        const long long p0 = G.prop0 + il * M.K;
        double cur[kMutMaxDim];
        for (int j = 0; j < M.d; ++j) cur[j] = theta_res[i * M.d + j];
        double cl = logL_res[i];
        int acc = 0, cand = 0;
        accept_chain(M, G.beta, (uint64_t)(M.index_offset + (unsigned long long)il), props + (size_t)p0 * M.d, ll + p0, inside + p0, cur, cl,
                     acc, cand);
        for (int j = 0; j < M.d; ++j) theta_out[i * M.d + j] = cur[j];
        logL_out[i] = cl;
        if (acc) atomicAdd(counts + 8 * g + 0, (unsigned long long)acc);
        if (cand) atomicAdd(counts + 8 * g + 1, (unsigned long long)cand);
    }
}

// status words of the proposal solves -> health counters per group (EV: LikelihoodHealthCounter)
__global__ void status_count_kernel(const MutationGroup* __restrict__ groups, int n_groups, const uint32_t* __restrict__ status,
                                    long long total, unsigned long long* counts) {
    __shared__ long long start[kMutMaxGroups];
    if (threadIdx.x < n_groups) start[threadIdx.x] = groups[threadIdx.x].prop0;
    __syncthreads();
    int cur = -1, cnt[3] = {0, 0, 0};
    auto flush = [&]() {
        // one atomic per warp when every lane holds counts of the same group (the common case), else one per lane
        const bool uniform = __all_sync(0xffffffffu, cur == __shfl_sync(0xffffffffu, cur, 0));
        for (int bit = 0; bit < 3; ++bit) {
            int v = cnt[bit];
            if (uniform) {
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                if ((threadIdx.x & 31) == 0 && cur >= 0 && v) atomicAdd(counts + 8 * cur + 2 + bit, (unsigned long long)v);
            } else if (cur >= 0 && v) {
                atomicAdd(counts + 8 * cur + 2 + bit, (unsigned long long)v);
            }
            cnt[bit] = 0;
        }
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long rounds = (total + stride - 1) / stride;  // every lane runs the same number of rounds (warp collectives inside)
    for (long long r = 0; r < rounds; ++r) {
        const long long t = r * stride + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        const int g = (t < total) ? group_of(start, n_groups, t) : cur;
        if (__any_sync(0xffffffffu, g != cur && cur >= 0)) flush();
        cur = g;
        if (t < total) {
            const uint32_t s = status[t];
            cnt[0] += (s & 1u) != 0;
            cnt[1] += (s & 2u) != 0;
            cnt[2] += (s & 4u) != 0;
        }
    }
    flush();
}
#endif

}  // namespace hmx
