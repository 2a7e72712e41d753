// Per-element math of the count likelihoods (Poisson / NB / zero-inflated variants, DU:206-305,
// ZI:180-199) shared by the streaming likelihood kernel and the fused heads kernel.
#pragma once

#include "common.cuh"

namespace scvae {

template <int KIND>
struct Lik;
template <>
struct Lik<SCVAE_LIK_POISSON> { static constexpr int P = 1; static constexpr bool NB = false; static constexpr bool ZI = false; };
template <>
struct Lik<SCVAE_LIK_NB> { static constexpr int P = 2; static constexpr bool NB = true; static constexpr bool ZI = false; };
template <>
struct Lik<SCVAE_LIK_ZIP> { static constexpr int P = 2; static constexpr bool NB = false; static constexpr bool ZI = true; };
template <>
struct Lik<SCVAE_LIK_ZINB> { static constexpr int P = 3; static constexpr bool NB = true; static constexpr bool ZI = true; };

// tf.clip_by_value gradient mask: passes inside [lo, hi] (boundaries included).
__device__ __forceinline__ float clip_mask(float a, float lo, float hi) {
    return (a < lo || a > hi) ? 0.f : 1.f;
}

// Everything of one (cell, gene) term except the x>0-only special functions:
//   lp  : log p without [lgamma(r+x) - lgamma(r)] and without lgamma(1+x)
//   g[] : d lp / d a[] without the digamma part of log_r
//   r, cr : total_count and the coefficient of [digamma(r+x) - digamma(r)] in g[log_r]
template <int KIND, bool BWD>
__device__ __forceinline__ void lik_elem(float x, const float (&a)[3], float &lp, float (&g)[3],
                                         float &r, float &cr) {
    using T = Lik<KIND>;
    constexpr int iD = T::ZI ? 1 : 0;  // first head of the wrapped count distribution
    const bool pos = x > 0.f;

    // wrapped distribution: lp_d (x-linear part), its gradients, log p_d(0)
    float lp_d, l0, gd0 = 0.f, gd1 = 0.f;
    if (T::NB) {
        const float ap_raw = a[iD], ar_raw = a[iD + 1];
        const float ap = fmaxf(ap_raw, kLogitFloor);
        const float lr = fminf(fmaxf(ar_raw, -10.f), 10.f);
        r = fast_ex2(lr * kLog2e);
        float e, u;
        const float sp = softplus_eu(ap, e, u);  // -log(1-p)
        l0 = -r * sp;
        // x == 1 (about half of all non-zero counts): lgamma(r+1) - lgamma(r) = log r exactly
        const bool one = (x == 1.f);
        lp_d = fmaf(x, ap - sp, l0) + (one ? lr : 0.f);
        if (BWD) {
            const float p = sigmoid_from_eu(ap, e, u);
            const float mp = ap_raw < kLogitFloor ? 0.f : 1.f;
            const float mr = fabsf(ar_raw) > 10.f ? 0.f : 1.f;
            gd0 = fmaf(-(x + r), p, x) * mp;
            gd1 = (l0 + (one ? 1.f : 0.f)) * mr;  // r (digamma(r+1) - digamma(r)) = 1
            cr = r * mr;
        }
    } else {
        const float al = a[iD];
        const float ll = fminf(fmaxf(al, -10.f), 10.f);
        const float lam = fast_ex2(ll * kLog2e);
        r = 0.f;
        cr = 0.f;
        l0 = -lam;
        lp_d = fmaf(x, ll, -lam);
        if (BWD) gd0 = (x - lam) * (fabsf(al) > 10.f ? 0.f : 1.f);
    }

    if (!T::ZI) {
        lp = lp_d;
        if (BWD) {
            g[0] = gd0;
            g[1] = gd1;
        }
        return;
    }
    // zero inflation (ZI:194-199) with pi = sigmoid(a_pi):
    //   x > 0 : log(1-pi) + lp_d                = -softplus(a_pi) + lp_d
    //   x <= 0: log(pi + (1-pi) exp(l0))        = softplus(l0 - a_pi) - softplus(-a_pi)
    const float api_raw = a[0];
    const float api = fmaxf(api_raw, kLogitFloor);
    float e_pi, u_pi, e_u, u_u;
    const float sp_pi = softplus_eu(api, e_pi, u_pi);
    const float u = l0 - api;
    const float sp_u = softplus_eu(u, e_u, u_u);
    lp = pos ? (lp_d - sp_pi) : (sp_u - (sp_pi - api));
    if (BWD) {
        const float pi = sigmoid_from_eu(api, e_pi, u_pi);
        const float wz = sigmoid_from_eu(u, e_u, u_u);  // (1-pi) e^{l0} / (pi + (1-pi) e^{l0})
        const float mpi = api_raw < kLogitFloor ? 0.f : 1.f;
        g[0] = (pos ? -pi : (1.f - pi) - wz) * mpi;
        const float w = pos ? 1.f : wz;
        g[1] = gd0 * w;
        g[2] = gd1 * w;  // NB only; unused for ZIP
        cr = pos ? cr : 0.f;
    }
}


// One group of W (cell, gene) terms of a single cell: accumulates log p into `acc`, writes
// d log p / d a (unscaled) into gv.  The branch-free part of all W elements comes first (W
// independent dependency chains for the scheduler); counts >= 2 (rare) are then fixed up quad
// by quad with a per-thread loop over the quad's non-zeros, so that a warp loops max-popcount
// times instead of once per element (x == 1 is handled branch-free in lik_elem; lgamma(1 + 1)
// = 0 needs no fix-up either).  W is 1 or a multiple of 4.
template <int KIND, bool BWD, int W>
__device__ __forceinline__ void lik_group(const float (&x)[W], const float (&av)[3][W], bool has_const,
                                          float &acc, float (&gv)[3][W]) {
    using T = Lik<KIND>;
    constexpr int P = T::P;
    constexpr int Q = W >= 4 ? 4 : W;     // fix-up granularity
    float rv[W], crv[W];
    unsigned nz = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const float aj[3] = {av[0][j], P > 1 ? av[1][j] : 0.f, P > 2 ? av[2][j] : 0.f};
        float lp, g[3] = {0.f, 0.f, 0.f};
        lik_elem<KIND, BWD>(x[j], aj, lp, g, rv[j], crv[j]);
        acc += lp;
        if (BWD) {
            gv[0][j] = g[0];
            gv[1][j] = g[1];
            gv[2][j] = g[2];
        }
        nz |= ((x[j] > 0.f && x[j] != 1.f) ? 1u : 0u) << j;
    }
    if (T::NB || !has_const) {
#pragma unroll
        for (int b = 0; b < W; b += Q) {
            unsigned m = (nz >> b) & ((1u << Q) - 1u);
            while (m) {
                const int j = __ffs(m) - 1;
                m &= m - 1;
                float xj = x[b], rj = rv[b], cj = crv[b];
#pragma unroll
                for (int q = 1; q < Q; ++q) {
                    xj = j == q ? x[b + q] : xj;
                    rj = j == q ? rv[b + q] : rj;
                    cj = j == q ? crv[b + q] : cj;
                }
                float extra = 0.f;
                if (T::NB) {
                    float D, Pd;
                    lgamma_diff(rj, xj, D, Pd);
                    extra = D;
                    if (BWD) {
                        const float add = cj * Pd;
                        constexpr int ir = P - 1;  // log_r is the last head
#pragma unroll
                        for (int q = 0; q < Q; ++q) gv[ir][b + q] += (j == q) ? add : 0.f;
                    }
                }
                if (!has_const) extra -= lgammaf(1.f + xj);
                acc += extra;
            }
        }
    }
}

}  // namespace scvae
