// hmx_model.cuh -- arithmetic of one implicit time step of the 5-species Hamilton ODE, written for
// one CUDA thread per (particle, condition) solve with every live quantity in registers.
//
// What the reference computes per quasi-Newton iteration (S5 = tmcmc/program2602/improved_5species_jit.py):
//   residual Q(g_new; g_prev)            S5:43-127
//   approximate Jacobian K (12x12)       S5:129-289   (NOT the exact derivative -- reproduced as is)
//   delta = solve(K, -Q)  (LAPACK dgesv) S5:369-373
//   backtracking on ||Q||_inf            S5:381-431
//
// B200-first restructuring (same mathematics, different algorithm; parity is checked against the
// dense oracle in tests/):
//   * K is never formed.  With unknowns ordered per species (dphi_i, dpsi_i), the 10x10 (phi,psi) part is
//       M = blockdiag(D_i) + P C Q^T,   P = blockdiag([psi_i; phi_i]),  Q = blockdiag([psi_i, phi_i]),
//     because every A-derived off-diagonal 2x2 block of S5:197-245 is the rank-one outer product
//     -c/eta_i * h_i * A_ij * [psi_i; phi_i] [psi_j, phi_j]   (Hill corrections S5:249-287 included: they
//     scale row-species 4 by h and add a multiple of the same outer product at block (4,3)).
//     The phi_0 row (S5:222-223) and the constraint row / gamma column (S5:220,247) are eliminated
//     analytically.  What remains is five 2x2 inverses, ONE 5x5 pivoted solve with two right-hand sides
//     and a scalar closure for d-gamma: ~0.9 kflop instead of ~2.4 kflop for K + dgesv, and ~60 live
//     doubles instead of 144+, which is what lets the whole step stay in registers.
//   * all barrier terms of one variable v share ONE reciprocal r = 1/(v(v-1)):
//       Kp1(2-4v)/((v-1)^3 v^3) = -2Kp1(2v-1) r^3      (phi, phi_0 rows AND, after combining its two
//                                                         fractions, the psi rows S5:117-119)
//       phi_p (S5:185-187)      = 2Kp1(2v-1) r^3 (2 + 3(2v-1) r)
//       psi_p (S5:189)          = 4Kp1(3 + 5 v(v-1)) r^4
//   * the residual of an accepted line-search trial IS the residual the next iteration would recompute
//     (same clipped inputs), so each state-machine trip evaluates Q exactly once.
//
// Everything here is __host__ __device__ so tests can run the identical arithmetic on the CPU
// (tests/csrc/host_emulation.cpp); the shipped library only instantiates the device side.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HMX_HD __host__ __device__ __forceinline__
#else
#define HMX_HD inline
#endif

namespace hmx {

constexpr double kClipLo = 1e-12;         // S5:64
constexpr double kClipHi = 1.0 - 1e-12;   // S5:68

// Solver constants shared by all evaluations of a context (lives in __constant__ memory on the device).
struct SolverConsts {
    double dt, inv_dt, eps, Kp1, c, alpha, K_hill, n_hill, Kn;  // Kn = K_hill^n_hill
    double eta[5], eta_phi[5], inv_eta[5], c_eta[5], alpha_eta[5];  // c/eta_i, alpha/eta_i
    double etaphi_eta[5];  // eta_phi_i / eta_i: lets (1/eta)(eta_phi + eta psi^2) be one fma
    int32_t n_steps, max_iter;
    int32_t hill_mode;  // 0 off, 1..4 integer exponent, 5 generic pow
    int32_t pad_;
};

// Reciprocal without the slow-path branch of IEEE division: MUFU.RCP64H seed (2^-23) + two Newton steps.
// Every divisor on the hot path is a normal, non-zero number (v(v-1) >= 1e-12, 2x2 determinants, pivots of
// I + diag(g) C, phi0_p + 1/dt); 0 / inf / NaN inputs still give inf / 0 / NaN, only the last-bit rounding
// differs from 1.0 / x (<= 1.5 ulp), which is far inside the 1e-9 parity bar.  It removes ~24 branches and
// ~100 fix-up instructions per trip from a latency-bound loop.
HMX_HD double rcp(double x) {
#if defined(__CUDA_ARCH__) && !defined(HMX_IEEE_DIV)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if defined(HMX_RCP_CUBIC)
    // one third-order step: y (1 + e + e^2) with e = 1 - x y; the seed is good to 2^-23, so the error becomes e^3 <= 2^-69
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
#else
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
#endif
#else
    return 1.0 / x;
#endif
}

HMX_HD double clip01(double v) {
    // two ordered compares: NaN stays NaN exactly like S5:66-69
    return (v < kClipLo) ? kClipLo : ((v > kClipHi) ? kClipHi : v);
}

// Strictly-inside-the-box test on the HIGH 32-bit word only (one integer subtract + one unsigned compare on the ALU
// instead of two DSETP on the FP64 pipe): bits(1e-12) = 0x3D719799_812DEA11 and bits(1 - 1e-12) = 0x3FEFFFFF_FFFFDCD0, so
// 0x3D71979A <= hi32(v) < 0x3FEFFFFF implies 1e-12 < v < 1 - 1e-12.  Negative values, NaN, inf and the two thin bands next to
// the bounds fail the test and take the exact clip (clip01) -- the test is conservative, never wrong.
HMX_HD bool strictly_inside_hi(double v) {
#if defined(__CUDA_ARCH__)
    const uint32_t h = (uint32_t)__double2hiint(v);
#else
    unsigned long long u;
    memcpy(&u, &v, sizeof(u));
    const uint32_t h = (uint32_t)(u >> 32);
#endif
    return (h - 0x3D71979Au) < (0x3FEFFFFFu - 0x3D71979Au);
}

// max(|a|, |b|) as one ordered compare and a select.  fmax() costs twice the instructions because of its NaN
// rules; the callers detect NaN separately (sum of |Q|) and never use the maximum when one is present.
HMX_HD double max_abs(double a, double b) {
    const double fa = fabs(a), fb = fabs(b);
    return (fa > fb) ? fa : fb;
}

// |a| as its IEEE-754 bit pattern.  For non-negative doubles the order of the patterns (as unsigned integers) IS the
// order of the values, +inf is 0x7ff0..0 and every NaN pattern lies above it.  The infinity norm of the residual, its
// NaN test (S5:342-348) and the finiteness scan are therefore ONE unsigned-integer max tree on the ALU instead of
// DSETP / DADD work on the FP64 pipe, which is the pipe this kernel is bound by (ncu r02a: the compare-select maximum and
// the |Q| sum were 63 of ~1030 FP64-pipe instructions per trip).  The resulting double is bit-identical.
constexpr unsigned long long kInfBits = 0x7ff0000000000000ull;
HMX_HD unsigned long long abs_bits(double a) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffull;
#else
    unsigned long long u;
    memcpy(&u, &a, sizeof(u));
    return u & 0x7fffffffffffffffull;
#endif
}
HMX_HD double bits_double(unsigned long long u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double a;
    memcpy(&a, &u, sizeof(a));
    return a;
#endif
}
HMX_HD unsigned long long umax64(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// The same maximum with fewer issue slots (HMX_NORM_FP): the magnitude comparison is ONE DSETP with |.| operand modifiers and
// the select picks the RAW value of larger magnitude (two SEL), so no absolute value is ever materialised; |.| is applied once
// at the root.  Non-finite entries are detected by a 32-bit integer max tree over the high words (exponent all ones);
// only then (never on the hot path) is the NaN test done entry by entry.  Values identical to the 64-bit integer tree.
HMX_HD double larger_mag(double a, double b) { return (fabs(a) > fabs(b)) ? a : b; }
HMX_HD uint32_t abs_hi(double a) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2hiint(a) & 0x7fffffffu;
#else
    unsigned long long u;
    memcpy(&u, &a, sizeof(u));
    return (uint32_t)(u >> 32) & 0x7fffffffu;
#endif
}
HMX_HD uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
// ||Q||_inf, any(isnan(Q)) and all(isfinite(Q)) of the twelve residual entries
HMX_HD void norm12(const double (&Q)[12], double& norm, bool& nan, bool& finite) {
#if defined(HMX_NORM_FP)
    const double m0 = larger_mag(larger_mag(Q[0], Q[6]), larger_mag(Q[1], Q[7]));
    const double m1 = larger_mag(larger_mag(Q[2], Q[8]), larger_mag(Q[3], Q[9]));
    const double m2 = larger_mag(larger_mag(Q[4], Q[10]), larger_mag(Q[5], Q[11]));
    norm = fabs(larger_mag(larger_mag(m0, m1), m2));
    const uint32_t h0 = umax32(umax32(abs_hi(Q[0]), abs_hi(Q[6])), umax32(abs_hi(Q[1]), abs_hi(Q[7])));
    const uint32_t h1 = umax32(umax32(abs_hi(Q[2]), abs_hi(Q[8])), umax32(abs_hi(Q[3]), abs_hi(Q[9])));
    const uint32_t h2 = umax32(umax32(abs_hi(Q[4]), abs_hi(Q[10])), umax32(abs_hi(Q[5]), abs_hi(Q[11])));
    finite = umax32(umax32(h0, h1), h2) < 0x7ff00000u;
    nan = false;
    if (!finite) {  // rare: an inf or a NaN somewhere -- redo it with the exact semantics (a NaN makes the norm irrelevant)
        double mx = 0.0;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            nan = nan || (Q[i] != Q[i]);
            mx = (fabs(Q[i]) > mx) ? fabs(Q[i]) : mx;
        }
        norm = mx;
    }
#else
    unsigned long long nb[6];
#pragma unroll
    for (int i = 0; i < 5; ++i) nb[i] = umax64(abs_bits(Q[i]), abs_bits(Q[6 + i]));
    nb[5] = umax64(abs_bits(Q[5]), abs_bits(Q[11]));
    const unsigned long long m = umax64(umax64(umax64(nb[0], nb[1]), umax64(nb[2], nb[3])), umax64(nb[4], nb[5]));
    norm = bits_double(m);  // a NaN pattern when a NaN is present: every comparison with it is false
    nan = m > kInfBits;     // true for NaN only, +inf entries are not NaN (S5:342-348)
    finite = m < kInfBits;
#endif
}

// packed symmetric A: index of (i,j)
HMX_HD constexpr int apk(int i, int j) {
    return (i <= j) ? (i * 5 - (i * (i - 1)) / 2 + (j - i)) : (j * 5 - (j * (j - 1)) / 2 + (i - j));
}

// Column of a [k][thread] array in shared memory: element k of this thread lives at p[k * STRIDE]
// (consecutive threads -> consecutive 8-byte words: conflict-free).  Used to park per-lane data that is
// read a few times per trip (the interaction matrix) outside the register file.
template <int STRIDE>
struct SmemCol {
    double* p;
    HMX_HD double& operator[](int k) const { return p[k * STRIDE]; }
};

// theta (20) -> packed symmetric A (15) and b (5): S5:513-559
template <class AT>
HMX_HD void theta_to_Ab(const double* th, AT& A, double (&b)[5]) {
#pragma unroll
    for (int k = 0; k < 15; ++k) A[k] = 0.0;
    A[apk(0, 0)] = th[0];
    A[apk(0, 1)] = th[1];
    A[apk(1, 1)] = th[2];
    b[0] = th[3];
    b[1] = th[4];
    A[apk(2, 2)] = th[5];
    A[apk(2, 3)] = th[6];
    A[apk(3, 3)] = th[7];
    b[2] = th[8];
    b[3] = th[9];
    A[apk(0, 2)] = th[10];
    A[apk(0, 3)] = th[11];
    A[apk(1, 2)] = th[12];
    A[apk(1, 3)] = th[13];
    A[apk(4, 4)] = th[14];
    b[4] = th[15];
    A[apk(0, 4)] = th[16];
    A[apk(1, 4)] = th[17];
    A[apk(2, 4)] = th[18];
    A[apk(3, 4)] = th[19];
}

// x^n and x^(n-1) for the Hill gate (S5:92, 278); integer exponents avoid pow()
template <int HILL>
HMX_HD void hill_powers(const SolverConsts& K, double x, double& xn, double& xn1) {
    if (HILL == 2) {
        xn = pow(x, K.n_hill);
        xn1 = pow(x, K.n_hill - 1.0);
    } else if (HILL == 4) {  // n_hill == 4 at compile time (the production setting): no switch on the hot path
        const double x2 = x * x;
        xn = x2 * x2;
        xn1 = x2 * x;
    } else {
        const double x2 = x * x;
        switch (K.hill_mode) {
            case 1: xn = x; xn1 = 1.0; break;
            case 2: xn = x2; xn1 = x; break;
            case 3: xn = x2 * x; xn1 = x2; break;
            default: xn = x2 * x2; xn1 = x2 * x; break;
        }
    }
}

// Everything one residual evaluation produces that the direction solve reuses.  Kept small on purpose: the
// kernel is register-bound, so the barrier derivatives are finished inside the residual (while 1/(v(v-1)) is
// still at hand) instead of carrying the reciprocals across the accept/reject decision.
struct Eval {
    double ph[5], ps[5];           // clipped values the residual was evaluated at
    double p[5];                   // phi_i psi_i
    double I[5];                   // gated interaction (A p)_i * h_i
    double phd[5], psd[5];         // time derivatives
    double phip[5], psip[5];       // barrier derivatives phi_p, psi_p of S5:185-189
    double K55;                    // phi0_p + 1/dt  (S5:191-194, 222)
    double Iraw4, h, dh_dx;        // Hill gate state (h = 1 when the gate is off)
    double Q[12];
    double norm;                   // ||Q||_inf
    bool nan;                      // any(isnan(Q))
    bool inbox;                    // every phi_i, psi_i of the point was strictly inside the clip box (nothing was clipped)
};

// Residual Q at the candidate x (S5:43-127).  gp = [phi_old(5), phi0_old, psi_old(5)].
template <bool ALPHA, int HILL, class GPT, class AT>
HMX_HD void residual(const SolverConsts& K, const double (&x)[12], const GPT& gp,
                     const AT& A, const double (&b)[5], Eval& e) {
    const double gam = x[11];
    // phi, psi inside the box is the overwhelmingly common case: one ordered compare pair per value and a
    // branch around the selects (NaN fails the test and takes the clip path, which leaves it NaN like S5:66-69)
#if defined(HMX_CLIP_BRANCH) || defined(HMX_CLIP_HIWORD)
    bool inbox = true;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#if defined(HMX_CLIP_HIWORD)
        inbox = inbox && strictly_inside_hi(x[i]) && strictly_inside_hi(x[6 + i]);
#else
        inbox = inbox && (x[i] >= kClipLo) && (x[i] <= kClipHi) && (x[6 + i] >= kClipLo) && (x[6 + i] <= kClipHi);
#endif
    e.inbox = inbox;
    if (inbox) {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            e.ph[i] = x[i];
            e.ps[i] = x[6 + i];
        }
    } else
#else
    e.inbox = false;
#endif
    {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            e.ph[i] = clip01(x[i]);
            e.ps[i] = clip01(x[6 + i]);
        }
    }
    const double ph0 = clip01(x[5]);
    const double m2K = -2.0 * K.Kp1;
    double fph[5], fps[5];  // barrier values -2Kp1(2v-1) r^3
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = e.ph[i], u = e.ps[i];
        const double wv = fma(v, v, -v), wu = fma(u, u, -u);
#if defined(HMX_RCP_PAIR)
        const double rvu = rcp(wv * wu);  // one reciprocal for the pair: 1/wv = wu / (wv wu), 1/wu = wv / (wv wu)
        const double rv = rvu * wu, ru = rvu * wv;
#else
        const double rv = rcp(wv), ru = rcp(wu);
#endif
        const double sv = fma(2.0, v, -1.0), su = fma(2.0, u, -1.0);
#if defined(HMX_FOLD_CONSTS)
        // -2 Kp1 (2v - 1) as ONE fma, 4 Kp1 (5 w + 3) as ONE fma (the constants are folded on the host side of the bank)
        fph[i] = fma(2.0 * m2K, v, -m2K) * (rv * rv * rv);
        fps[i] = fma(2.0 * m2K, u, -m2K) * (ru * ru * ru);
        e.phip[i] = -fph[i] * fma(3.0 * sv, rv, 2.0);                        // S5:185-187
        const double ru2 = ru * ru;
        e.psip[i] = fma(20.0 * K.Kp1, wu, 12.0 * K.Kp1) * (ru2 * ru2);        // S5:189
#else
        fph[i] = (m2K * sv) * (rv * rv * rv);
        fps[i] = (m2K * su) * (ru * ru * ru);
        e.phip[i] = -fph[i] * fma(3.0 * sv, rv, 2.0);                        // S5:185-187
        const double ru2 = ru * ru;
        e.psip[i] = (4.0 * K.Kp1) * fma(5.0, wu, 3.0) * (ru2 * ru2);          // S5:189
#endif
        e.p[i] = v * u;
        e.phd[i] = (v - gp[i]) * K.inv_dt;
        e.psd[i] = (u - gp[6 + i]) * K.inv_dt;
    }
    double f0;
    {
        const double v = ph0;
        const double r0 = rcp(fma(v, v, -v));
        const double s0 = fma(2.0, v, -1.0);
        f0 = (m2K * s0) * (r0 * r0 * r0);
        e.K55 = -f0 * fma(3.0 * s0, r0, 2.0) + K.inv_dt;                      // S5:191-194, 222
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double s = A[apk(i, 0)] * e.p[0];
#pragma unroll
        for (int j = 1; j < 5; ++j) s = fma(A[apk(i, j)], e.p[j], s);
        e.I[i] = s;
    }
    e.h = 1.0;
    e.dh_dx = 0.0;
    e.Iraw4 = e.I[4];
    if (HILL != 0) {
        const double fn = e.p[3];
        if (fn > 1e-20) {
            double xn, xn1;
            hill_powers<HILL>(K, fn, xn, xn1);
            const double rden = rcp(K.Kn + xn);
            e.h = xn * rden;
            e.dh_dx = K.n_hill * K.Kn * xn1 * (rden * rden);
        } else {
            e.h = 0.0;
        }
        e.I[4] *= e.h;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = e.ph[i], u = e.ps[i];
        // Q_i, S5:98-106: (1/eta)[gamma + (eta_phi + eta psi^2) phidot + eta phi psi psidot], 1/eta distributed
        const double t2 = fma(K.inv_eta[i], gam, fma(fma(u, u, K.etaphi_eta[i]), e.phd[i], e.p[i] * e.psd[i]));
        const double qi = fph[i] + t2 - K.c_eta[i] * u * e.I[i];
        // Q_{6+i}, S5:116-123
        double qs = fps[i] + fma(e.p[i], e.phd[i], v * v * e.psd[i]) - K.c_eta[i] * v * e.I[i];
        if (ALPHA) qs += (b[i] * K.alpha_eta[i]) * u;
        e.Q[i] = qi;
        e.Q[6 + i] = qs;
    }
    e.Q[5] = gam + f0 + (ph0 - gp[5]) * K.inv_dt;                                  // S5:109-113
    e.Q[11] = ((((e.ph[0] + e.ph[1]) + e.ph[2]) + e.ph[3]) + e.ph[4]) + ph0 - 1.0;    // S5:126
    bool finite;
    norm12(e.Q, e.norm, e.nan, finite);  // ||Q||_inf; any(isnan(Q)): true for NaN only, +inf entries are not NaN (S5:342-348)
}

// 5x5 solve with two right-hand sides.  S[r][0..4] matrix, S[r][5], S[r][6] the two right-hand sides ->
// overwritten with the solutions.
//
// The matrix is S = I + diag(g) C with g_i = q_i^T D_i^{-1} p_i ~ dt and |C_ij| = c/eta |A_ij|, i.e. a small
// perturbation of the identity for every physically meaningful input (|g C| ~ 0.07 at the production MAP).
// Fast path: elimination in natural order, fully unrolled, registers only.  It reports whether every pivot
// magnitude stayed above kPivotFloor; when not, the caller redoes the solve with partial pivoting
// (solve5x2_pivoted, local memory, dynamic indexing -- off the hot path).
constexpr double kPivotFloor = 0.0625;

HMX_HD bool solve5x2_natural(double (&S)[5][7]) {
    double rp[5];
    bool ok = true;  // every pivot at least kPivotFloor in magnitude (a NaN pivot fails the test)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        ok = ok && (fabs(S[k][k]) >= kPivotFloor);
        rp[k] = rcp(S[k][k]);
#pragma unroll
        for (int r = k + 1; r < 5; ++r) {
            const double l = S[r][k] * rp[k];
#pragma unroll
            for (int cc = k + 1; cc < 7; ++cc) S[r][cc] = fma(-l, S[k][cc], S[r][cc]);
        }
    }
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        double s0 = S[k][5], s1 = S[k][6];
#pragma unroll
        for (int cc = k + 1; cc < 5; ++cc) {
            s0 = fma(-S[k][cc], S[cc][5], s0);
            s1 = fma(-S[k][cc], S[cc][6], s1);
        }
        S[k][5] = s0 * rp[k];
        S[k][6] = s1 * rp[k];
    }
    return ok;
}

// Division-free variant of the natural-order elimination: row_r <- S_kk row_r - S_rk row_k.  No reciprocal sits on the
// elimination chain (the five divisions of the back substitution are independent of each other); the pivots of
// I + diag(g) C are ~1, so the scaled rows stay O(1).  Healthy when every scaled pivot keeps at least kPivotFloor of the
// previous one.
HMX_HD bool solve5x2_divfree(double (&S)[5][7]) {
    bool ok = fabs(S[0][0]) >= kPivotFloor;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double pk = S[k][k];
#pragma unroll
        for (int r = k + 1; r < 5; ++r) {
            const double l = S[r][k];
#pragma unroll
            for (int cc = k + 1; cc < 7; ++cc) S[r][cc] = fma(pk, S[r][cc], -l * S[k][cc]);
        }
        ok = ok && (fabs(S[k + 1][k + 1]) >= kPivotFloor * fabs(pk));
    }
    double rp[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) rp[k] = rcp(S[k][k]);
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        double s0 = S[k][5], s1 = S[k][6];
#pragma unroll
        for (int cc = k + 1; cc < 5; ++cc) {
            s0 = fma(-S[k][cc], S[cc][5], s0);
            s1 = fma(-S[k][cc], S[cc][6], s1);
        }
        S[k][5] = s0 * rp[k];
        S[k][6] = s1 * rp[k];
    }
    return ok;
}

// Same system by 2+3 block elimination: explicit 2x2 inverse, Schur complement, 3x3 adjugate.  Only two
// reciprocals sit on the dependency chain (five in the natural-order LU) and every other level is a batch of
// independent FMAs, which is what a latency-bound FP64 kernel with two warps per scheduler needs.
// Returns min(|det P|, |det T'|): both are 1 + O(|g C|) for a healthy system.
HMX_HD double solve5x2_block23(double (&S)[5][7]) {
    const double tp = S[0][1] * S[1][0];
    const double detP = fma(S[0][0], S[1][1], -tp) + fma(-S[0][1], S[1][0], tp);
    const double rP = rcp(detP);
    const double i00 = S[1][1] * rP, i01 = -S[0][1] * rP, i10 = -S[1][0] * rP, i11 = S[0][0] * rP;
    double X[2][5];  // P^{-1} [Q | r_top]: columns 0..2 = Q, 3..4 = right-hand sides
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double q0 = S[0][2 + c], q1 = S[1][2 + c];
        X[0][c] = fma(i00, q0, i01 * q1);
        X[1][c] = fma(i10, q0, i11 * q1);
    }
    double T[3][5];  // Schur complement and reduced right-hand sides
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) T[r][c] = fma(-S[2 + r][1], X[1][c], fma(-S[2 + r][0], X[0][c], S[2 + r][2 + c]));
    // adjugate of the 3x3 block
    const double c00 = fma(T[1][1], T[2][2], -T[1][2] * T[2][1]);
    const double c01 = fma(T[1][2], T[2][0], -T[1][0] * T[2][2]);
    const double c02 = fma(T[1][0], T[2][1], -T[1][1] * T[2][0]);
    const double c10 = fma(T[0][2], T[2][1], -T[0][1] * T[2][2]);
    const double c11 = fma(T[0][0], T[2][2], -T[0][2] * T[2][0]);
    const double c12 = fma(T[0][1], T[2][0], -T[0][0] * T[2][1]);
    const double c20 = fma(T[0][1], T[1][2], -T[0][2] * T[1][1]);
    const double c21 = fma(T[0][2], T[1][0], -T[0][0] * T[1][2]);
    const double c22 = fma(T[0][0], T[1][1], -T[0][1] * T[1][0]);
    const double det3 = fma(T[0][0], c00, fma(T[0][1], c01, T[0][2] * c02));
    const double r3 = rcp(det3);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const double b0 = T[0][3 + q], b1 = T[1][3 + q], b2 = T[2][3 + q];
        const double z2 = fma(c00, b0, fma(c10, b1, c20 * b2)) * r3;
        const double z3 = fma(c01, b0, fma(c11, b1, c21 * b2)) * r3;
        const double z4 = fma(c02, b0, fma(c12, b1, c22 * b2)) * r3;
        S[2][5 + q] = z2;
        S[3][5 + q] = z3;
        S[4][5 + q] = z4;
        S[0][5 + q] = fma(-X[0][2], z4, fma(-X[0][1], z3, fma(-X[0][0], z2, X[0][3 + q])));
        S[1][5 + q] = fma(-X[1][2], z4, fma(-X[1][1], z3, fma(-X[1][0], z2, X[1][3 + q])));
    }
    return fmin(fabs(detP), fabs(det3));
}

// Partial (row) pivoting variant for the rare ill-scaled case.  M is the untouched copy of the system.
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#endif
static void solve5x2_pivoted(const double* M /*[35]*/, double* X /*[10]: x0[5], x1[5]*/) {
    double T[5][7];
    for (int r = 0; r < 5; ++r)
        for (int cc = 0; cc < 7; ++cc) T[r][cc] = M[r * 7 + cc];
    for (int k = 0; k < 5; ++k) {
        int p = k;
        double best = fabs(T[k][k]);
        for (int r = k + 1; r < 5; ++r) {
            const double a = fabs(T[r][k]);
            if (a > best) {
                best = a;
                p = r;
            }
        }
        if (p != k)
            for (int cc = 0; cc < 7; ++cc) {
                const double t = T[k][cc];
                T[k][cc] = T[p][cc];
                T[p][cc] = t;
            }
        const double rpk = 1.0 / T[k][k];
        for (int r = k + 1; r < 5; ++r) {
            const double l = T[r][k] * rpk;
            for (int cc = k + 1; cc < 7; ++cc) T[r][cc] = fma(-l, T[k][cc], T[r][cc]);
        }
    }
    for (int k = 4; k >= 0; --k) {
        double s0 = T[k][5], s1 = T[k][6];
        for (int cc = k + 1; cc < 5; ++cc) {
            s0 = fma(-T[k][cc], X[cc], s0);
            s1 = fma(-T[k][cc], X[5 + cc], s1);
        }
        X[k] = s0 / T[k][k];
        X[5 + k] = s1 / T[k][k];
    }
}

// Scratch for the per-species vectors D_i^{-1} r_i, D_i^{-1} u_i, D_i^{-1} [psi_i; phi_i] (30 doubles) that are
// produced before and consumed after the 5x5 solve.  In the kernel they are parked in shared memory so they
// do not occupy registers during the elimination; on the host / in the observation kernel a plain array.
struct RegStash {
    double v[30];
    HMX_HD double& operator[](int k) { return v[k]; }
};

// 2x2 diagonal block of species i (rows i, 6+i; columns i, 6+i) of the reference's Newton matrix, S5:199-245 (+ S5:254-271),
// or -- EXACT -- of the true derivative dQ/dg_new.
template <bool ALPHA, int HILL, bool EXACT, class AT>
HMX_HD void block2x2(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], int i, double& d00, double& d01,
                     double& d10, double& d11) {
        const double v = e.ph[i], u = e.ps[i], pd = e.phd[i], sd = e.psd[i];
        const double ce = K.c_eta[i];
        const double hi = (HILL != 0 && i == 4) ? e.h : 1.0;
        const double aii = hi * A[apk(i, i)];
        const double ap = aii * e.p[i];
        if (!EXACT) {
            d00 = e.phip[i] + fma(fma(u, u, K.etaphi_eta[i]), K.inv_dt, u * sd) - ce * u * fma(2.0 * aii, u, e.I[i]);
            d01 = fma(2.0 * u, pd, fma(v, sd, e.p[i] * K.inv_dt)) - ce * fma(3.0, ap, e.I[i]);  // (1/eta) eta = 1
            d10 = fma(u, pd, fma(e.p[i], K.inv_dt, 2.0 * v * sd)) - ce * fma(3.0, ap, 2.0 * e.I[i]);
            d11 = e.psip[i] + fma(v, pd, v * v * K.inv_dt) - 2.0 * ce * aii * v * v;
        } else {
            // barrier derivatives: d/dv[Kp1(2-4v)/((v-1)^3 v^3)] = 2 Kp1 r^3 (3 s^2 r - 2),
            //                      d/du[-2Kp1/((u-1)^2 u^3) - 2Kp1/((u-1)^3 u^2)] = 2 Kp1 (10 w + 3) r^4,  w = u(u-1), r = 1/w, s = 2v-1
            const double wv = fma(v, v, -v), wu = fma(u, u, -u);
            const double rv = rcp(wv), ru = rcp(wu);
            const double sv = fma(2.0, v, -1.0);
            const double phip = (2.0 * K.Kp1) * (rv * rv * rv) * fma(3.0 * sv * sv, rv, -2.0);
            const double ru2 = ru * ru;
            const double psip = (2.0 * K.Kp1) * fma(10.0, wu, 3.0) * (ru2 * ru2);
            // e.I[i] is the gated interaction h_i (A p)_i; the A_ii part of the rank-one term is folded into the block
            d00 = phip + fma(fma(u, u, K.etaphi_eta[i]), K.inv_dt, u * sd) - ce * aii * u * u;
            d01 = fma(2.0 * u, pd, fma(v, sd, e.p[i] * K.inv_dt)) - ce * (e.I[i] + ap);
            d10 = fma(u, pd, fma(e.p[i], K.inv_dt, 2.0 * v * sd)) - ce * (e.I[i] + ap);
            d11 = psip + fma(v, pd, v * v * K.inv_dt) - ce * aii * v * v;
        }
        if (ALPHA) d11 += b[i] * K.alpha_eta[i];
}

// Quasi-Newton direction delta = solve(K, -Q) with K the reference's approximate Jacobian (S5:129-289)
// evaluated at the point of `e`, computed through the block structure described at the top.
//
// EXACT = true solves with the TRUE derivative dQ/dg_new instead (same block structure, different 2x2 blocks): what the
// forward-sensitivity recursion of the likelihood gradient needs (hmx_grad.cuh).  The reference's matrix differs from
// the derivative in four places: the numerator derivative of the phi barrier (-4 + 8v for -4, S5:185), the constant of
// the psi barrier derivative (6 for 3, S5:189), an extra -c/eta psi I term in dQ_phi/dphi and a doubled A_ii term on the
// block diagonals (S5:199-245).  ph0 = clip(phi_0) is only read when EXACT.
template <bool ALPHA, int HILL, bool EXACT, class AT, class DT, class ST>
HMX_HD void direction_impl(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], DT& delta, ST& stash, double ph0) {
    double S[5][7];
    double srow[5];  // -c/eta_i h_i: row scaling of the A-derived coefficients C_ij = srow_i A_ij
#pragma unroll
    for (int i = 0; i < 5; ++i) srow[i] = -K.c_eta[i];
    if (HILL != 0) srow[4] *= e.h;  // h == 1 reproduces "no correction" (S5:250)
    // the dh/dx outer product at block (4,3), S5:273-287
    double c43x = 0.0;
    if (HILL != 0) {
        if (e.h < 1.0) c43x = -K.c_eta[4] * e.Iraw4 * e.dh_dx;
    }

#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = e.ph[i], u = e.ps[i], pd = e.phd[i], sd = e.psd[i];
        const double ce = K.c_eta[i];
        const double hi = (HILL != 0 && i == 4) ? e.h : 1.0;
        const double aii = hi * A[apk(i, i)];
        const double ap = aii * e.p[i];
        // 2x2 diagonal block of species i (rows i, 6+i; columns i, 6+i), S5:199-245 (+ S5:254-271)
        double d00, d01, d10, d11;
        block2x2<ALPHA, HILL, EXACT>(K, e, A, b, i, d00, d01, d10, d11);
        // inverse by Cramer with an fma-compensated determinant
        const double t = d01 * d10;
        const double det = fma(d00, d11, -t) + fma(-d01, d10, t);
        const double rdet = rcp(det);
        const double r0 = -e.Q[i], r1 = -e.Q[6 + i];
        const double a0 = fma(d11, r0, -d01 * r1) * rdet;
        const double a1 = fma(d00, r1, -d10 * r0) * rdet;
        const double rdi = rdet * K.inv_eta[i];  // u_i = (1/eta_i, 0): S5:220
        const double b0 = d11 * rdi;
        const double b1 = -d10 * rdi;
        const double dd0 = fma(d11, u, -d01 * v) * rdet;
        const double dd1 = fma(d00, v, -d10 * u) * rdet;
        stash[6 * i + 0] = a0;
        stash[6 * i + 1] = a1;
        stash[6 * i + 2] = b0;
        stash[6 * i + 3] = b1;
        stash[6 * i + 4] = dd0;
        stash[6 * i + 5] = dd1;
        // row i of S = I + diag(g) C and its two right-hand sides
        const double gi = fma(u, dd0, v * dd1) * srow[i];  // q_i^T D_i^{-1} p_i times the row scaling
#pragma unroll
        for (int j = 0; j < 5; ++j) S[i][j] = (i == j) ? 1.0 : gi * A[apk(i, j)];
        if (HILL != 0 && i == 4) S[4][3] = fma(fma(u, dd0, v * dd1), c43x, S[4][3]);
        S[i][5] = fma(u, a0, v * a1);
        S[i][6] = fma(u, b0, v * b1);
    }
#if defined(HMX_SOLVE_DIVFREE)
    const bool healthy = solve5x2_divfree(S);
#elif defined(HMX_SOLVE_BLOCK23)
    const bool healthy = solve5x2_block23(S) >= kPivotFloor;  // shorter dependency chain, a few more live values; measured 2 % slower
#else
    const bool healthy = solve5x2_natural(S);
#endif
    if (!healthy) {  // rare: rebuild the system and redo it with row pivoting
        double T[5][7], X[10];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const double gq = fma(e.ps[i], stash[6 * i + 4], e.ph[i] * stash[6 * i + 5]);
#pragma unroll
            for (int j = 0; j < 5; ++j) T[i][j] = (i == j) ? 1.0 : gq * srow[i] * A[apk(i, j)];
            if (HILL != 0 && i == 4) T[4][3] = fma(gq, c43x, T[4][3]);
            T[i][5] = fma(e.ps[i], stash[6 * i + 0], e.ph[i] * stash[6 * i + 1]);
            T[i][6] = fma(e.ps[i], stash[6 * i + 2], e.ph[i] * stash[6 * i + 3]);
        }
        solve5x2_pivoted(&T[0][0], X);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            S[i][5] = X[i];
            S[i][6] = X[5 + i];
        }
    }
    // phi_0 row and the constraint row: S5:222-223, 247
    double K55 = e.K55;
    if (EXACT) {
        const double r0 = rcp(fma(ph0, ph0, -ph0)), s0 = fma(2.0, ph0, -1.0);
        K55 = (2.0 * K.Kp1) * (r0 * r0 * r0) * fma(3.0 * s0 * s0, r0, -2.0) + K.inv_dt;
    }
    const double rK55 = rcp(K55);
    double sumY1 = 0.0, sumY2 = 0.0;
    double y10[5], y11[5], y20[5], y21[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double t1 = 0.0, t2 = 0.0;  // sum_j C_ij z_j = srow_i sum_j A_ij z_j (+ the (4,3) extra)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (j != i) {
                t1 = fma(A[apk(i, j)], S[j][5], t1);
                t2 = fma(A[apk(i, j)], S[j][6], t2);
            }
        }
        t1 *= srow[i];
        t2 *= srow[i];
        if (HILL != 0 && i == 4) {
            t1 = fma(c43x, S[3][5], t1);
            t2 = fma(c43x, S[3][6], t2);
        }
        const double dd0 = stash[6 * i + 4], dd1 = stash[6 * i + 5];
        y10[i] = fma(-dd0, t1, stash[6 * i + 0]);
        y11[i] = fma(-dd1, t1, stash[6 * i + 1]);
        y20[i] = fma(-dd0, t2, stash[6 * i + 2]);
        y21[i] = fma(-dd1, t2, stash[6 * i + 3]);
    }
    sumY1 = ((y10[0] + y10[1]) + (y10[2] + y10[3])) + y10[4];
    sumY2 = ((y20[0] + y20[1]) + (y20[2] + y20[3])) + y20[4];
    const double srhs = fma(e.Q[5], rK55, -e.Q[11]);
    const double dgam = (sumY1 - srhs) * rcp(sumY2 + rK55);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        delta[i] = fma(-dgam, y20[i], y10[i]);
        delta[6 + i] = fma(-dgam, y21[i], y11[i]);
    }
    delta[5] = (-e.Q[5] - dgam) * rK55;
    delta[11] = dgam;
}

// ---- factor once, solve many --------------------------------------------------------------------------------------
// The <= 20 sensitivity systems of one trajectory row (hmx_obs.cuh) share the Newton matrix and differ in the right-hand
// side only.  direction_factor() does everything of direction_impl() that does not depend on -Q (2x2 blocks and their
// inverses, S = I + diag(g) C, its elimination with the multipliers kept, the whole second right-hand side -- the gamma
// column --, the closure denominator); direction_apply() runs the right-hand-side half with the SAME operations in the same
// order, so its result equals direction()'s.  ~130 instead of ~650 FP64 instructions per additional right-hand side.
struct DirFactor {
    double d00[5], d01[5], d10[5], d11[5], rdet[5];
    double dd0[5], dd1[5];
    double srow[5], c43x;
    double LU[5][5];  // upper triangle (incl. the unit-scaled rows' entries) and, below the diagonal, the multipliers l_rk
    double rp[5];     // reciprocal pivots
    double y20[5], y21[5];
    double rK55, rden;  // 1/K55, 1/(sum y20 + 1/K55)
    bool healthy;       // every pivot above kPivotFloor; when false the caller falls back to direction()
};

template <bool ALPHA, int HILL, class AT>
HMX_HD void direction_factor(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], DirFactor& F) {
#pragma unroll
    for (int i = 0; i < 5; ++i) F.srow[i] = -K.c_eta[i];
    if (HILL != 0) F.srow[4] *= e.h;
    F.c43x = 0.0;
    if (HILL != 0) {
        if (e.h < 1.0) F.c43x = -K.c_eta[4] * e.Iraw4 * e.dh_dx;
    }
    double S[5][5], r2[5], b0[5], b1[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double v = e.ph[i], u = e.ps[i];
        double d00, d01, d10, d11;
        block2x2<ALPHA, HILL, false>(K, e, A, b, i, d00, d01, d10, d11);
        const double t = d01 * d10;
        const double det = fma(d00, d11, -t) + fma(-d01, d10, t);
        const double rdet = rcp(det);
        const double rdi = rdet * K.inv_eta[i];
        b0[i] = d11 * rdi;
        b1[i] = -d10 * rdi;
        const double dd0 = fma(d11, u, -d01 * v) * rdet;
        const double dd1 = fma(d00, v, -d10 * u) * rdet;
        F.d00[i] = d00;
        F.d01[i] = d01;
        F.d10[i] = d10;
        F.d11[i] = d11;
        F.rdet[i] = rdet;
        F.dd0[i] = dd0;
        F.dd1[i] = dd1;
        const double gi = fma(u, dd0, v * dd1) * F.srow[i];
#pragma unroll
        for (int j = 0; j < 5; ++j) S[i][j] = (i == j) ? 1.0 : gi * A[apk(i, j)];
        if (HILL != 0 && i == 4) S[4][3] = fma(fma(u, dd0, v * dd1), F.c43x, S[4][3]);
        r2[i] = fma(u, b0[i], v * b1[i]);
    }
    // solve5x2_natural with the multipliers kept and only the second right-hand side carried along
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        ok = ok && (fabs(S[k][k]) >= kPivotFloor);
        F.rp[k] = rcp(S[k][k]);
#pragma unroll
        for (int r = k + 1; r < 5; ++r) {
            const double l = S[r][k] * F.rp[k];
#pragma unroll
            for (int cc = k + 1; cc < 5; ++cc) S[r][cc] = fma(-l, S[k][cc], S[r][cc]);
            r2[r] = fma(-l, r2[k], r2[r]);
            S[r][k] = l;
        }
    }
    F.healthy = ok;
    double z2[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        double s1 = r2[k];
#pragma unroll
        for (int cc = k + 1; cc < 5; ++cc) s1 = fma(-S[k][cc], z2[cc], s1);
        z2[k] = s1 * F.rp[k];
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) F.LU[i][j] = S[i][j];
    F.rK55 = rcp(e.K55);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double t2 = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (j != i) t2 = fma(A[apk(i, j)], z2[j], t2);
        t2 *= F.srow[i];
        if (HILL != 0 && i == 4) t2 = fma(F.c43x, z2[3], t2);
        F.y20[i] = fma(-F.dd0[i], t2, b0[i]);
        F.y21[i] = fma(-F.dd1[i], t2, b1[i]);
    }
    const double sumY2 = ((F.y20[0] + F.y20[1]) + (F.y20[2] + F.y20[3])) + F.y20[4];
    F.rden = rcp(sumY2 + F.rK55);
}

// delta = solve(K, -Q) for one more right-hand side Q (12 entries) with the factorisation of the point of `e`
template <int HILL, class AT, class DT>
HMX_HD void direction_apply(const Eval& e, const AT& A, const DirFactor& F, const double (&Q)[12], DT& delta) {
    double a0[5], a1[5], c[5], z[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const double r0 = -Q[i], r1 = -Q[6 + i];
        a0[i] = fma(F.d11[i], r0, -F.d01[i] * r1) * F.rdet[i];
        a1[i] = fma(F.d00[i], r1, -F.d10[i] * r0) * F.rdet[i];
        c[i] = fma(e.ps[i], a0[i], e.ph[i] * a1[i]);
    }
#pragma unroll
    for (int k = 0; k < 5; ++k)
#pragma unroll
        for (int r = k + 1; r < 5; ++r) c[r] = fma(-F.LU[r][k], c[k], c[r]);
#pragma unroll
    for (int k = 4; k >= 0; --k) {
        double s0 = c[k];
#pragma unroll
        for (int cc = k + 1; cc < 5; ++cc) s0 = fma(-F.LU[k][cc], z[cc], s0);
        z[k] = s0 * F.rp[k];
    }
    double y10[5], y11[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        double t1 = 0.0;
#pragma unroll
        for (int j = 0; j < 5; ++j)
            if (j != i) t1 = fma(A[apk(i, j)], z[j], t1);
        t1 *= F.srow[i];
        if (HILL != 0 && i == 4) t1 = fma(F.c43x, z[3], t1);
        y10[i] = fma(-F.dd0[i], t1, a0[i]);
        y11[i] = fma(-F.dd1[i], t1, a1[i]);
    }
    const double sumY1 = ((y10[0] + y10[1]) + (y10[2] + y10[3])) + y10[4];
    const double srhs = fma(Q[5], F.rK55, -Q[11]);
    const double dgam = (sumY1 - srhs) * F.rden;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        delta[i] = fma(-dgam, F.y20[i], y10[i]);
        delta[6 + i] = fma(-dgam, F.y21[i], y11[i]);
    }
    delta[5] = (-Q[5] - dgam) * F.rK55;
    delta[11] = dgam;
}

template <bool ALPHA, int HILL, class AT, class DT, class ST>
HMX_HD void direction(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], DT& delta, ST& stash) {
    direction_impl<ALPHA, HILL, false>(K, e, A, b, delta, stash, 0.0);
}

// convenience overload with a register-array stash (host emulation, observation kernel)
template <bool ALPHA, int HILL, class AT, class DT>
HMX_HD void direction(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], DT& delta) {
    RegStash st;
    direction<ALPHA, HILL>(K, e, A, b, delta, st);
}

// delta = solve(dQ/dg_new, -e.Q) with the exact derivative (ph0 = clip(phi_0) of the point `e` was evaluated at)
template <bool ALPHA, int HILL, class AT, class DT>
HMX_HD void direction_exact(const SolverConsts& K, const Eval& e, const AT& A, const double (&b)[5], DT& delta, double ph0) {
    RegStash st;
    direction_impl<ALPHA, HILL, true>(K, e, A, b, delta, st, ph0);
}

}  // namespace hmx
