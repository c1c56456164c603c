// hmx_stage.cuh -- per-stage TMCMC arithmetic on the device (K2-K5 of SURVEY.md section 2.3).
//
// All reductions are deterministic: every block writes its partial to a fixed slot and the block that
// draws the last ticket adds the slots in index order, so replicas on different GPUs that received the
// same log-likelihood vector (NCCL all-gather) take bit-identical decisions.
//
//   K2 beta_bisect_*   25-step bisection on ESS(delta) >= ratio*N                         (TM:613-659)
//   K3 weights_*       w = exp(d*logL - max), sum, normalise, log S_j                      (TM:686-701)
//   K4 cumsum_serial   np.cumsum semantics (strict left-to-right adds) + searchsorted      (TM:186-198)
//   K5 wcov_*          np.cov(theta.T, aweights=w), as moments + centred scatter           (TM:766)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace hmx {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 1024;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum / max of one value per thread (blockDim.x == kRedThreads); result valid in thread 0
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double* sm /*[8]*/) {
    v = IS_MAX ? warp_max(v) : warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[wid] = v;
    __syncthreads();
    if (wid == 0) {
        v = (lane < (kRedThreads >> 5)) ? sm[lane] : (IS_MAX ? -INFINITY : 0.0);
        v = IS_MAX ? warp_max(v) : warp_sum(v);
    }
    return v;
}

// returns true in thread 0 of the block that finished last
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
    __shared__ bool is_last;
    __threadfence();   // publish this thread's partial(s)
    __syncthreads();   // ... for every thread of the block, before the ticket is drawn
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0u;  // re-arm for the next launch on this stream
    }
    __syncthreads();
    return is_last;
}

// ---- state shared by the stage kernels (device memory, owned by the context) ----------------------------
struct StageState {
    double vmax;          // max(logL) or max(delta*logL)
    double lo, hi;        // bisection bracket
    double ess_lo;        // ESS at the last accepted mid (NaN if none)
    double sum_w, sum_w2;
    double delta_beta;    // result of K2
    double log_Sj;        // result of K3
    int32_t uniform_fallback;
    unsigned int ticket;
};

// max over x[i]*scale (scale = 1 for K2, delta for K3: the reference maxes the scaled vector, TM:619,687)
__global__ void __launch_bounds__(kRedThreads) max_kernel(const double* __restrict__ x, long long N, double scale,
                                                          double* partial, StageState* st) {
    __shared__ double sm[8];
    double m = -INFINITY;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        m = fmax(m, x[i] * scale);
    m = block_reduce<true>(m, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = m;
    if (last_block(&st->ticket)) {
        double v = -INFINITY;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) v = fmax(v, __ldcg(partial + b));
        v = block_reduce<true>(v, sm);
        if (threadIdx.x == 0) st->vmax = v;
    }
}

__global__ void bisect_init_kernel(StageState* st, double beta) {
    st->lo = 0.0;
    st->hi = 1.0 - beta;
    st->ess_lo = nan("");
}

// one bisection step: ESS(mid) = (sum w)^2 / sum w^2 with w = exp(mid (logL - max))   (TM:618-650)
__global__ void __launch_bounds__(kRedThreads) bisect_step_kernel(const double* __restrict__ logL, long long N,
                                                                   double target, double* partial, StageState* st) {
    __shared__ double sm[8];
    const double mid = 0.5 * (st->lo + st->hi);
    const double vmax = st->vmax;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double w = exp(mid * (logL[i] - vmax));
        s1 += w;
        s2 = fma(w, w, s2);
    }
    s1 = block_reduce<false>(s1, sm);
    s2 = block_reduce<false>(s2, sm);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = s1;
        partial[2 * blockIdx.x + 1] = s2;
    }
    if (last_block(&st->ticket)) {
        double a = 0.0, b = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) {
            a += __ldcg(partial + 2 * k);
            b += __ldcg(partial + 2 * k + 1);
        }
        a = block_reduce<false>(a, sm);
        b = block_reduce<false>(b, sm);
        if (threadIdx.x == 0) {
            const double ess = (a > 0.0) ? (a * a) / b : 0.0;  // sum_w <= 0 -> ess = 0 (TM:622-623)
            if (ess >= target) {
                st->lo = mid;
                st->ess_lo = ess;
            } else {
                st->hi = mid;
            }
        }
    }
}

__global__ void bisect_final_kernel(StageState* st, double beta, double min_db, double max_db, double* out_db,
                                    double* out_ess) {
    // TM:656-657
    const double raw = fmax(st->lo, min_db);
    const double db = fmin(fmin(raw, max_db), 1.0 - beta);
    st->delta_beta = db;
    if (out_db) *out_db = db;
    if (out_ess) *out_ess = st->ess_lo;
}

// K3: w_i = exp(delta logL_i - max);  partial sums
__global__ void __launch_bounds__(kRedThreads) weights_exp_kernel(const double* __restrict__ logL, long long N, double delta,
                                                                   double* __restrict__ w, double* partial, StageState* st) {
    __shared__ double sm[8];
    const double vmax = st->vmax;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double v = exp(__dsub_rn(__dmul_rn(delta, logL[i]), vmax));  // product rounded first, as NumPy does
        w[i] = v;
        s += v;
    }
    s = block_reduce<false>(s, sm);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
    if (last_block(&st->ticket)) {
        double a = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) a += __ldcg(partial + k);
        a = block_reduce<false>(a, sm);
        if (threadIdx.x == 0) {
            st->sum_w = a;
            const bool bad = !(a > 0.0) || !(fabs(a) <= 1.79e308);  // TM:691
            st->uniform_fallback = bad ? 1 : 0;
            st->log_Sj = bad ? nan("") : (log(a) - log((double)N) + vmax);  // TM:697-698
        }
    }
}

__global__ void __launch_bounds__(kRedThreads) weights_norm_kernel(double* __restrict__ w, long long N, const StageState* st,
                                                                    double* out_logS) {
    const double s = st->sum_w;
    const bool uni = st->uniform_fallback != 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
        w[i] = uni ? 1.0 / (double)N : w[i] / s;
    if (out_logS && blockIdx.x == 0 && threadIdx.x == 0) *out_logS = st->log_Sj;
}

// K4a: cs = np.cumsum(w) with the exact left-to-right association, cs[N-1] := 1.0  (TM:195-197).
// np.cumsum is one chain of N dependent, individually rounded FP64 adds; re-associating it could move a resampling
// index, so the result must equal that chain bit for bit.  It does NOT have to be computed as a chain, though:
// while the running sum s = S u stays inside one binade [2^e, 2^(e+1)) (u = 2^(e-52) its ulp, S an integer), adding a
// non-negative w that is not an exact tie gives RN(s + w) = (S + q) u with q = RN(w / u) -- the rounding of each
// addend does not depend on the running sum.  So a tile of 4096 weights is handled in parallel: every thread rounds
// its 16 weights to integers q (bit operations on the mantissa), an integer block scan gives the prefixes, and the
// doubles are assembled from (e, S + prefix) directly.  A tile that contains a tie (w / u has fractional part exactly
// 1/2: the result would depend on the parity of S), a binade crossing, or anything unusual (negative, non-finite,
// subnormal running sum) is walked by thread 0 as the literal dependent chain, ~11 clk per element.  For N = 2^21
// normalised weights ~3 % of the tiles take the chain (the binade crossings sit in the first 1/2^k of the array):
// 1.3 ms instead of 11.7 ms.  Equality with np.cumsum is tested bitwise on adversarial inputs (ties, crossings,
// sorted / zero / subnormal weights) in tests/test_gpu_stage.py; scripts in tests/tools reproduce it in NumPy.
constexpr int kCumsumTile = 4096;
constexpr int kCumsumPerThread = kCumsumTile / kRedThreads;  // 16 contiguous weights per thread
static_assert(kCumsumPerThread * kRedThreads == kCumsumTile, "tile must be a multiple of the block size");

// q = RN(w / 2^(e-1075)) for a non-negative finite w; flags an exact tie or a weight above the binade of s
__device__ __forceinline__ unsigned long long cumsum_round_weight(double w, int e, bool& bad) {
    const unsigned long long wb = (unsigned long long)__double_as_longlong(w);
    const int ew = (int)((wb >> 52) & 0x7ffull);
    if ((wb >> 63) || ew == 0x7ff) {
        bad = true;
        return 0ull;
    }
    const unsigned long long mw = (wb & ((1ull << 52) - 1)) | (ew > 0 ? (1ull << 52) : 0ull);
    const int sh = e - (ew > 0 ? ew : 1);  // w = mw 2^(ew-1075); u = 2^(e-1075)
    if (sh < 2) {
        bad = true;  // at least a quarter of the running sum: (almost) certainly a binade crossing; also keeps 4096 q < 2^64
        return 0ull;
    }
    if (sh >= 64) return 0ull;  // mw < 2^53 <= 2^(sh-1): rounds to zero, cannot tie
    const unsigned long long fl = mw >> sh;
    const unsigned long long rem = mw - (fl << sh);
    const unsigned long long half = 1ull << (sh - 1);
    if (rem == half) bad = true;  // exact tie: the rounding direction depends on the parity of the running sum
    return fl + (rem > half ? 1ull : 0ull);
}

__global__ void __launch_bounds__(kRedThreads) cumsum_serial_kernel(const double* __restrict__ w, long long N, double* __restrict__ cs) {
    __shared__ double tile[kCumsumTile];
    __shared__ unsigned long long warp_tot[kRedThreads / 32];
    __shared__ double carry_sh;
    if (threadIdx.x == 0) carry_sh = 0.0;
    __syncthreads();
    for (long long base = 0; base < N; base += kCumsumTile) {
        const int n = (int)((N - base < kCumsumTile) ? (N - base) : kCumsumTile);
        for (int t = threadIdx.x; t < kCumsumTile; t += blockDim.x) tile[t] = (t < n) ? w[base + t] : 0.0;
        __syncthreads();
        const double carry = carry_sh;
        // ---- parallel attempt ----
        const unsigned long long sb = (unsigned long long)__double_as_longlong(carry);
        const int e = (int)((sb >> 52) & 0x7ffull);
        bool bad = (sb >> 63) || e == 0 || e == 0x7ff;  // zero / subnormal / negative / non-finite running sum
        const unsigned long long S = (sb & ((1ull << 52) - 1)) | (1ull << 52);
        unsigned long long pre[kCumsumPerThread];
        unsigned long long run = 0ull;
        const int t0 = threadIdx.x * kCumsumPerThread;
#pragma unroll
        for (int j = 0; j < kCumsumPerThread; ++j) {
            run += cumsum_round_weight(tile[t0 + j], e > 0 ? e : 1, bad);
            pre[j] = run;
        }
        // exclusive scan of the per-thread totals over the block
        unsigned long long incl = run;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((threadIdx.x & 31) >= o) incl += v;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
        const int any_bad = __syncthreads_or(bad ? 1 : 0);
        unsigned long long offset = incl - run, total = 0ull;
        for (int k = 0; k < kRedThreads / 32; ++k) {
            const unsigned long long wt = warp_tot[k];
            if (k < (int)(threadIdx.x >> 5)) offset += wt;
            total += wt;
        }
        const bool fast = !any_bad && (total < (1ull << 53)) && (S + total < (1ull << 53));
        if (fast) {
            const unsigned long long ebits = (unsigned long long)e << 52;
#pragma unroll
            for (int j = 0; j < kCumsumPerThread; ++j) {
                const unsigned long long tot = S + offset + pre[j];  // in [2^52, 2^53): the mantissa with its hidden bit
                tile[t0 + j] = __longlong_as_double((long long)(ebits | (tot & ((1ull << 52) - 1))));
            }
        }
        __syncthreads();
        if (!fast && threadIdx.x == 0) {
            // the literal chain (operands come from shared memory, off the dependency chain)
            double c = carry;
            int t = 0;
            for (; t + 8 <= n; t += 8) {
                const double x0 = tile[t], x1 = tile[t + 1], x2 = tile[t + 2], x3 = tile[t + 3];
                const double x4 = tile[t + 4], x5 = tile[t + 5], x6 = tile[t + 6], x7 = tile[t + 7];
                c = __dadd_rn(c, x0); tile[t] = c;
                c = __dadd_rn(c, x1); tile[t + 1] = c;
                c = __dadd_rn(c, x2); tile[t + 2] = c;
                c = __dadd_rn(c, x3); tile[t + 3] = c;
                c = __dadd_rn(c, x4); tile[t + 4] = c;
                c = __dadd_rn(c, x5); tile[t + 5] = c;
                c = __dadd_rn(c, x6); tile[t + 6] = c;
                c = __dadd_rn(c, x7); tile[t + 7] = c;
            }
            for (; t < n; ++t) {
                c = __dadd_rn(c, tile[t]);
                tile[t] = c;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = tile[n - 1];
        for (int t = threadIdx.x; t < n; t += blockDim.x) cs[base + t] = (base + t == N - 1) ? 1.0 : tile[t];
        __syncthreads();
    }
}


// ---- K4a, multi-CTA: the same exact cumulative sum, tiles processed by the whole GPU ------------------------------------
// The dependency between tiles is only the running sum that enters a tile.  Inside a run of consecutive tiles whose
// running sum stays in ONE binade and that contain no tie / oversized / irregular weight, that running sum is
// S_entry(t) = S_start + sum of the tiles' integer totals -- an INTEGER prefix sum, exact and associative.  So:
//   A  cumsum_tile_sum_kernel      (all CTAs)  plain floating-point sum of every tile (only used to GUESS the binade)
//   B  cumsum_guess_kernel         (1 CTA)     scan of the tile sums -> guessed exponent of the running sum entering each tile
//   C  cumsum_tile_round_kernel    (all CTAs)  per tile, with the guessed exponent: q_i = RN(w_i / ulp), total = sum q_i, tie / irregular flag
//   D  cumsum_carry_kernel         (1 thread)  walks the TILES in order with the exact running sum: a tile whose guess was right,
//                                              is regular and does not cross its binade advances the running sum by one integer
//                                              add; any other tile is walked as the literal chain of rounded adds (11 clk/element)
//                                              and written at once.  Records the entering mantissa of every fast tile.
//   E  cumsum_tile_write_kernel    (all CTAs)  fast tiles: integer block scan from the recorded mantissa, doubles assembled from bits
// The guess can be wrong (a tile sum within rounding of a power of two): D detects it by comparing exponents and falls back
// to the chain for that tile, so the result never depends on the guess -- it is the sequentially rounded np.cumsum chain
// bit for bit, like the single-CTA kernel above (kept as the N < 2 tiles path and as the reference in the tests).
struct CumsumTileInfo {
    unsigned long long total;   // sum of the rounded weights in units of the guessed ulp
    unsigned long long S_entry; // mantissa (with hidden bit) of the running sum entering the tile (written by D)
    int e_guess;                // guessed biased exponent of the running sum entering the tile
    int mode;                   // C: 0 regular, 1 irregular;  D: 0 fast (E writes it), 1 done by the chain
};

__global__ void __launch_bounds__(kRedThreads) cumsum_tile_sum_kernel(const double* __restrict__ w, long long N, double* __restrict__ tile_sum) {
    __shared__ double sm[8];
    const long long base = (long long)blockIdx.x * kCumsumTile;
    double s = 0.0;
    for (int t = threadIdx.x; t < kCumsumTile; t += blockDim.x)
        if (base + t < N) s += w[base + t];
    s = block_reduce<false>(s, sm);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s;
}

__global__ void __launch_bounds__(kRedThreads) cumsum_guess_kernel(const double* __restrict__ tile_sum, int n_tiles, CumsumTileInfo* info) {
    // serial over the tiles in one thread: n_tiles = N / 4096 (512 for 2 M particles), a few microseconds
    if (threadIdx.x == 0) {
        double c = 0.0;
        for (int t = 0; t < n_tiles; ++t) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(c);
            info[t].e_guess = (int)((b >> 52) & 0x7ffull);
            c += tile_sum[t];
        }
    }
}

__global__ void __launch_bounds__(kRedThreads) cumsum_tile_round_kernel(const double* __restrict__ w, long long N, CumsumTileInfo* info) {
    __shared__ unsigned long long sm_tot[kRedThreads / 32];
    const long long base = (long long)blockIdx.x * kCumsumTile;
    const int e = info[blockIdx.x].e_guess;
    bool bad = (e == 0 || e == 0x7ff);
    unsigned long long run = 0ull;
    const int t0 = threadIdx.x * kCumsumPerThread;
#pragma unroll
    for (int j = 0; j < kCumsumPerThread; ++j) {
        const long long i = base + t0 + j;
        const double x = (i < N) ? w[i] : 0.0;
        run += cumsum_round_weight(x, e > 0 ? e : 1, bad);
    }
    for (int o = 16; o > 0; o >>= 1) run += __shfl_down_sync(0xffffffffu, run, o);
    if ((threadIdx.x & 31) == 0) sm_tot[threadIdx.x >> 5] = run;
    const int any_bad = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        unsigned long long tot = 0ull;
        for (int k = 0; k < kRedThreads / 32; ++k) tot += sm_tot[k];
        info[blockIdx.x].total = tot;
        info[blockIdx.x].mode = (any_bad || tot >= (1ull << 53)) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(kRedThreads) cumsum_carry_kernel(const double* __restrict__ w, long long N, int n_tiles, CumsumTileInfo* info,
                                                                   double* __restrict__ cs) {
    __shared__ double tile[kCumsumTile];
    __shared__ double carry_sh;
    __shared__ int chain_sh;
    if (threadIdx.x == 0) carry_sh = 0.0;
    __syncthreads();
    for (int t = 0; t < n_tiles; ++t) {
        if (threadIdx.x == 0) {
            const double c = carry_sh;
            const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
            const int e = (int)((cb >> 52) & 0x7ffull);
            const unsigned long long S = (cb & ((1ull << 52) - 1)) | (1ull << 52);
            const bool regular = !(cb >> 63) && e != 0 && e != 0x7ff && e == info[t].e_guess && info[t].mode == 0;
            if (regular && S + info[t].total < (1ull << 53)) {
                info[t].S_entry = S;
                const unsigned long long tot = S + info[t].total;
                carry_sh = __longlong_as_double((long long)(((unsigned long long)e << 52) | (tot & ((1ull << 52) - 1))));
                chain_sh = 0;
            } else {
                info[t].mode = 1;  // guess wrong, irregular weights or a binade crossing inside the tile
                chain_sh = 1;
            }
        }
        __syncthreads();
        if (chain_sh) {  // the literal chain of rounded adds, operands staged in shared memory (11 clk per element)
            const long long base = (long long)t * kCumsumTile;
            const int n = (int)((N - base < kCumsumTile) ? (N - base) : kCumsumTile);
            for (int k = threadIdx.x; k < n; k += blockDim.x) tile[k] = w[base + k];
            __syncthreads();
            if (threadIdx.x == 0) {
                double c = carry_sh;
                int k = 0;
                for (; k + 8 <= n; k += 8) {
                    const double x0 = tile[k], x1 = tile[k + 1], x2 = tile[k + 2], x3 = tile[k + 3];
                    const double x4 = tile[k + 4], x5 = tile[k + 5], x6 = tile[k + 6], x7 = tile[k + 7];
                    c = __dadd_rn(c, x0); tile[k] = c;
                    c = __dadd_rn(c, x1); tile[k + 1] = c;
                    c = __dadd_rn(c, x2); tile[k + 2] = c;
                    c = __dadd_rn(c, x3); tile[k + 3] = c;
                    c = __dadd_rn(c, x4); tile[k + 4] = c;
                    c = __dadd_rn(c, x5); tile[k + 5] = c;
                    c = __dadd_rn(c, x6); tile[k + 6] = c;
                    c = __dadd_rn(c, x7); tile[k + 7] = c;
                }
                for (; k < n; ++k) {
                    c = __dadd_rn(c, tile[k]);
                    tile[k] = c;
                }
                carry_sh = c;
            }
            __syncthreads();
            for (int k = threadIdx.x; k < n; k += blockDim.x) cs[base + k] = tile[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cs[N - 1] = 1.0;  // TM:196 (the write kernel skips this element)
}

__global__ void __launch_bounds__(kRedThreads) cumsum_tile_write_kernel(const double* __restrict__ w, long long N, const CumsumTileInfo* info,
                                                                        double* __restrict__ cs) {
    __shared__ unsigned long long warp_tot[kRedThreads / 32];
    const CumsumTileInfo ti = info[blockIdx.x];
    if (ti.mode != 0) return;  // written by the chain
    const long long base = (long long)blockIdx.x * kCumsumTile;
    const int e = ti.e_guess;
    bool bad = false;
    unsigned long long pre[kCumsumPerThread];
    unsigned long long run = 0ull;
    const int t0 = threadIdx.x * kCumsumPerThread;
#pragma unroll
    for (int j = 0; j < kCumsumPerThread; ++j) {
        const long long i = base + t0 + j;
        run += cumsum_round_weight((i < N) ? w[i] : 0.0, e, bad);
        pre[j] = run;
    }
    unsigned long long incl = run;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long offset = incl - run;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) offset += warp_tot[k];
    const unsigned long long ebits = (unsigned long long)e << 52;
#pragma unroll
    for (int j = 0; j < kCumsumPerThread; ++j) {
        const long long i = base + t0 + j;
        if (i < N - 1) {  // element N - 1 is forced to 1.0 by the carry kernel
            const unsigned long long tot = ti.S_entry + offset + pre[j];
            cs[i] = __longlong_as_double((long long)(ebits | (tot & ((1ull << 52) - 1))));
        }
    }
}

// K4b: idx_i = clip(searchsorted(cs, (u+i)/N, side='left'), 0, N-1)   (TM:194,198)
__global__ void __launch_bounds__(kRedThreads) searchsorted_kernel(const double* __restrict__ cs, long long N, double u,
                                                                    long long* __restrict__ idx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double pos = __ddiv_rn(__dadd_rn(u, (double)i), (double)N);
        long long lo = 0, hi = N;
        while (lo < hi) {
            const long long mid = lo + ((hi - lo) >> 1);
            if (cs[mid] < pos)
                lo = mid + 1;
            else
                hi = mid;
        }
        idx[i] = (lo > N - 1) ? N - 1 : lo;
    }
}

// K5a: moments = [sum w, sum w^2, sum w theta_j (j<d)]; one block-column per moment keeps it simple:
// grid.y = 2 + d, grid.x blocks over particles.
__global__ void __launch_bounds__(kRedThreads) wmoments_kernel(const double* __restrict__ theta, const double* __restrict__ w,
                                                                long long N, int d, double* partial /*[gridDim.y][gridDim.x]*/,
                                                                double* out /*[2+d]*/, unsigned int* tickets /*[gridDim.y]*/) {
    __shared__ double sm[8];
    const int m = blockIdx.y;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        const double wi = w[i];
        s += (m == 0) ? wi : (m == 1) ? wi * wi : theta[i * d + (m - 2)] * wi;
    }
    s = block_reduce<false>(s, sm);
    if (threadIdx.x == 0) partial[(size_t)m * gridDim.x + blockIdx.x] = s;
    if (last_block(tickets + m)) {
        double a = 0.0;
        for (int k = threadIdx.x; k < (int)gridDim.x; k += blockDim.x) a += __ldcg(partial + (size_t)m * gridDim.x + k);
        a = block_reduce<false>(a, sm);
        if (threadIdx.x == 0) out[m] = a;
    }
}

__global__ void wmean_kernel(const double* moments, int d, double* mean) {
    const int j = threadIdx.x;
    if (j < d) mean[j] = moments[2 + j] / moments[0];  // np.average(..., weights=w)
}

// K5b: scatter[a,b] = sum_i (theta_ia - mean_a) ((theta_ib - mean_b) w_i); thread (a,b), particles staged in smem
constexpr int kScatterChunk = 64;
__global__ void __launch_bounds__(kRedThreads) wscatter_kernel(const double* __restrict__ theta, const double* __restrict__ w,
                                                                long long N, int d, const double* __restrict__ mean,
                                                                double* partial /*[gridDim.x][d*d]*/, double* out /*[d*d]*/,
                                                                unsigned int* ticket) {
    extern __shared__ double smem[];  // [kScatterChunk][d] centred thetas, then [kScatterChunk] weights
    double* xs = smem;
    double* ws = smem + (size_t)kScatterChunk * d;
    const int dd = d * d;
    // each thread owns up to ceil(dd / blockDim) output entries
    const int per = (dd + blockDim.x - 1) / blockDim.x;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};  // d <= 32 -> dd <= 1024 -> per <= 4
    for (long long base = (long long)blockIdx.x * kScatterChunk; base < N; base += (long long)gridDim.x * kScatterChunk) {
        const int n = (int)((N - base < kScatterChunk) ? (N - base) : kScatterChunk);
        __syncthreads();
        for (int t = threadIdx.x; t < n * d; t += blockDim.x) xs[t] = theta[base * d + t] - mean[t % d];
        for (int t = threadIdx.x; t < n; t += blockDim.x) ws[t] = w[base + t];
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = threadIdx.x + q * blockDim.x;
            if (q < per && e < dd) {
                // canonical operand order so that entry (a,b) and its mirror (b,a) round identically
                const int r0 = e / d, c0 = e % d;
                const int a = r0 < c0 ? r0 : c0, b = r0 < c0 ? c0 : r0;
                double s = acc[q];
                for (int i = 0; i < n; ++i) s = fma(xs[i * d + a], xs[i * d + b] * ws[i], s);
                acc[q] = s;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = threadIdx.x + q * blockDim.x;
        if (q < per && e < dd) partial[(size_t)blockIdx.x * dd + e] = acc[q];
    }
    if (last_block(ticket)) {
        for (int e = threadIdx.x; e < dd; e += blockDim.x) {
            double s = 0.0;
            for (int k = 0; k < (int)gridDim.x; ++k) s += __ldcg(partial + (size_t)k * dd + e);
            out[e] = s;
        }
    }
}

// cov = scatter * (1/fact), fact = sum w - sum w^2 / sum w  (ddof = 1 with aweights)
__global__ void wcov_final_kernel(const double* moments, const double* scatter, int d, double* cov, double* mean_out,
                                  const double* mean) {
    const double fact = moments[0] - moments[1] / moments[0];
    const double rf = 1.0 / fact;
    for (int e = threadIdx.x; e < d * d; e += blockDim.x) cov[e] = scatter[e] * rf;
    if (mean_out)
        for (int j = threadIdx.x; j < d; j += blockDim.x) mean_out[j] = mean[j];
}

// EV:305-334, rho = 0: strictly sequential accumulation in (row, species) order like the reference loop
__global__ void loglik_sparse_kernel(const double* mu, const double* sig, const double* data, const double* sigma_obs,
                                     const double* weights, int n_obs, double* out) {
    double logL = 0.0;
    for (int i = 0; i < n_obs; ++i) {
        for (int j = 0; j < 5; ++j) {
            const int e = i * 5 + j;
            const double var_raw = sig[e] + sigma_obs[e] * sigma_obs[e];
            const double v = (!(fabs(var_raw) <= 1.79e308) || var_raw <= 1e-20) ? 1e-20 : var_raw;
            const double r = data[e] - mu[e];
            const double w = weights ? weights[e] : 1.0;
            logL -= w * 0.5 * log(2.0 * 3.14159265358979323846 * v);
            logL -= w * 0.5 * (r * r) / v;
        }
    }
    *out = logL;
}

}  // namespace hmx
