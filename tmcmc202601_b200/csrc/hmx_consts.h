// hmx_consts.h -- hmx_config (C ABI) -> kernel constant blocks.  Shared by the CUDA library and by the
// CPU emulation used in tests; plain C++.
#pragma once

#include <math.h>
#include <string.h>

#include "../../include/hmx.h"
#include "hmx_model.cuh"

namespace hmx {

// per-condition constants consumed by the kernels
struct CondConsts {
    double g0[12];  // initial state (S5:564-576)
    double data[HMX_MAX_OBS * 5], sigma2[HMX_MAX_OBS * 5], weights[HMX_MAX_OBS * 5];  // sigma2 = sigma_obs^2
    double cov_rel;
    int32_t idx[HMX_MAX_OBS];
    int32_t n_obs;
    uint32_t active_theta_mask;
    int32_t use_absolute_volume;
    int32_t tsm_variance;
};

inline int validate_config(const hmx_config* c, const char** why) {
    *why = "";
    if (!c) { *why = "cfg is NULL"; return -1; }
    if (c->abi_version != HMX_ABI_VERSION) { *why = "abi_version mismatch"; return -1; }
    if (!(c->dt > 0.0) || c->n_steps < 1) { *why = "dt must be > 0 and n_steps >= 1"; return -1; }
    if (c->max_newton_iter < 1) { *why = "max_newton_iter must be >= 1"; return -1; }
    if (c->n_cond < 1 || c->n_cond > HMX_MAX_COND) { *why = "n_cond out of range"; return -1; }
    for (int i = 0; i < 5; ++i)
        if (!(c->eta[i] != 0.0)) { *why = "eta must be non-zero"; return -1; }
    for (int k = 0; k < c->n_cond; ++k) {
        const hmx_cond& q = c->cond[k];
        if (q.n_obs < 0 || q.n_obs > HMX_MAX_OBS) { *why = "n_obs out of range"; return -1; }
        for (int o = 0; o < q.n_obs; ++o) {
            // the reference raises IndexError for rows outside the trajectory (EV:682-697)
            if (q.idx_sparse[o] < 0 || q.idx_sparse[o] > c->n_steps) { *why = "idx_sparse outside [0, n_steps]"; return -1; }
            if (o > 0 && q.idx_sparse[o] <= q.idx_sparse[o - 1]) { *why = "idx_sparse must be strictly increasing"; return -1; }
        }
    }
    return 0;
}

inline void make_solver_consts(const hmx_config& c, SolverConsts& K) {
    memset(&K, 0, sizeof(K));
    K.dt = c.dt;
    K.inv_dt = 1.0 / c.dt;
    K.eps = c.eps;
    K.Kp1 = c.Kp1;
    K.c = c.c_const;
    K.alpha = c.alpha_const;
    K.K_hill = c.K_hill;
    K.n_hill = c.n_hill;
    K.Kn = (c.K_hill > 0.0) ? pow(c.K_hill, c.n_hill) : 0.0;
    for (int i = 0; i < 5; ++i) {
        K.eta[i] = c.eta[i];
        K.eta_phi[i] = c.eta_phi[i];
        K.inv_eta[i] = 1.0 / c.eta[i];
        K.c_eta[i] = c.c_const / c.eta[i];
        K.alpha_eta[i] = c.alpha_const / c.eta[i];
        K.etaphi_eta[i] = c.eta_phi[i] / c.eta[i];
    }
    K.n_steps = c.n_steps;
    K.max_iter = c.max_newton_iter;
    K.hill_mode = 0;
    if (c.K_hill > 0.0) {
        const double n = c.n_hill;
        K.hill_mode = (n == 1.0) ? 1 : (n == 2.0) ? 2 : (n == 3.0) ? 3 : (n == 4.0) ? 4 : 5;
    }
}

// template selector: 0 = gate off, 1 = integer exponent, 2 = generic pow
inline int hill_variant(const SolverConsts& K) { return K.hill_mode == 0 ? 0 : (K.hill_mode == 5 ? 2 : 1); }

inline void make_cond_consts(const hmx_config& c, int k, CondConsts& q) {
    const hmx_cond& s = c.cond[k];
    memset(&q, 0, sizeof(q));
    double sum = 0.0;
    for (int i = 0; i < 5; ++i) {
        const bool act = (c.active_species_mask >> i) & 1;
        q.g0[i] = act ? s.phi_init[i] : 0.0;
        sum += q.g0[i];
        q.g0[6 + i] = act ? 0.999 : 0.0;
    }
    q.g0[5] = 1.0 - sum;
    q.g0[11] = 0.0;
    for (int i = 0; i < HMX_MAX_OBS * 5; ++i) {
        q.data[i] = s.data[i];
        q.sigma2[i] = s.sigma_obs[i] * s.sigma_obs[i];  // s_ij**2, EV:310
        q.weights[i] = s.weights[i];
    }
    for (int o = 0; o < HMX_MAX_OBS; ++o) q.idx[o] = s.idx_sparse[o];
    q.n_obs = s.n_obs;
    q.cov_rel = s.cov_rel;
    q.active_theta_mask = s.active_theta_mask & ((1u << HMX_NTHETA) - 1u);
    q.use_absolute_volume = s.use_absolute_volume;
    q.tsm_variance = s.tsm_variance;
}

}  // namespace hmx

#include "hmx_legacy.cuh"

namespace hmx {

// hmx_legacy_config -> kernel constants of the generic NS-species solver (initial state: S4:857-865)
inline void make_legacy_consts(const hmx_legacy_config& c, LegacyConsts& K) {
    memset(&K, 0, sizeof(K));
    const int NS = c.n_species;
    K.dt = c.dt;
    K.inv_dt = 1.0 / c.dt;
    K.eps = c.eps;
    K.Kp1 = c.Kp1;
    K.c = c.c_const;
    K.alpha_const = c.alpha_const;
    K.alpha_before = c.alpha_before;
    K.alpha_after = c.alpha_after;
    K.switch_level = c.alpha_switch_level;
    K.n_steps = c.n_steps;
    K.max_iter = c.max_newton_iter;
    K.active_mask = (uint32_t)c.active_species_mask & ((1u << NS) - 1u);
    double sum = 0.0;
    for (int i = 0; i < NS; ++i) {
        K.eta[i] = c.eta[i];
        K.eta_phi[i] = c.eta_phi[i];
        K.inv_eta[i] = 1.0 / c.eta[i];
        K.c_eta[i] = c.c_const / c.eta[i];
        K.etaphi_eta[i] = c.eta_phi[i] / c.eta[i];
        const bool act = (K.active_mask >> i) & 1u;
        K.g0[i] = act ? c.phi_init[i] : 0.0;
        sum += K.g0[i];
        K.g0[NS + 1 + i] = act ? 0.999 : 0.0;
    }
    K.g0[NS] = 1.0 - sum;
    K.g0[2 * NS + 1] = 0.0;
}

}  // namespace hmx
