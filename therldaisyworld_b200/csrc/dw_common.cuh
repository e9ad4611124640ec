// Shared device-side definitions for the B200-native RLDaisyWorld step.
//
// "Literal" arithmetic = the canonical IEEE-binary64 operation order of the reference's step
// (daisy/daisy_world_rl.py:340-461 with daisy/nn/functional.py:12-49 restated as a 3x3 toroidal
// stencil).  The whole library is compiled with -fmad=false, so a*b+c below is two roundings exactly
// like NumPy; fused multiply-adds appear only where __fma_rn is written out (fast lattice path).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/daisyworld_b200.h"

// Physics constants as the kernels see them (copied from dw_config on every launch, so attribute
// mutation between steps behaves like the reference).
struct DevParams {
    int B, N, n_agents, pad_;
    double p, g, S, sigma, gamma, q, q2, temp_optimal, dt, agent_gamma;
    double ab, al, ad;
    double w[9];      // daisy-spread taps, row-major
    double adj[9];    // adjacent-albedo taps, row-major (centre tap skipped when it is 0, see nb_sum)
    double mask[9];   // observation mask
};

__device__ __forceinline__ double dw_root4(double x) { return sqrt(sqrt(x)); }
__device__ __forceinline__ double dw_pow4(double x) { double x2 = x * x; return x2 * x2; }
__device__ __forceinline__ double dw_div1000(double k);
// np.round(x, 3) = rint(x*1000)/1000 (true division); the division-free form is exact for |rint(x*1000)| <= 2e6
__device__ __forceinline__ double dw_round3(double x) {
    const double k = rint(x * 1000.0);
    return fabs(k) <= 2.0e6 ? dw_div1000(k) : k / 1000.0;
}
__device__ __forceinline__ double dw_clip01(double x) {
    x = x < 0.0 ? 0.0 : x;
    return x > 1.0 ? 1.0 : x;
}

struct LitCell {
    double nb, nl, nd;             // unrounded new covers (clipped l', d'; b' = (p-l')-d')
    double T, Tl, Td, Te;          // unrounded temperatures
    double beta, beta_l, beta_d;   // growth rates
    double dl, dd;                 // growth
    double b0;                     // (p-l)-d of the centre cell (written back in place by forward, :381)
};

// One cell of RLDaisyWorld.forward in literal order. l9/d9: 3x3 neighbourhood (row-major, wrap applied
// by the caller) of the light / dark covers; SL = S*L.
__device__ __forceinline__ LitCell dw_literal_cell(const DevParams &P, double SL, const double (&l9)[9],
                                                   const double (&d9)[9]) {
    LitCell o;
    double nb_b = 0.0, nb_l = 0.0, nb_d = 0.0, rho_l = 0.0, rho_d = 0.0;
    bool first = true;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        double tl = P.w[k] * l9[k], td = P.w[k] * d9[k];
        if (k == 0) { rho_l = tl; rho_d = td; }
        else { rho_l = rho_l + tl; rho_d = rho_d + td; }
        if (k == 4 && P.adj[4] == 0.0) continue;      // zero centre tap is skipped (oracle/daisy_numpy.py)
        double bk = (P.p - l9[k]) - d9[k];
        double vb = P.adj[k] * bk, vl = P.adj[k] * l9[k], vd = P.adj[k] * d9[k];
        if (first) { nb_b = vb; nb_l = vl; nb_d = vd; first = false; }
        else { nb_b = nb_b + vb; nb_l = nb_l + vl; nb_d = nb_d + vd; }
    }
    const double l = l9[4], d = d9[4];
    o.b0 = (P.p - l) - d;
    const double Al = (P.ab * o.b0 + P.al * l) + P.ad * d;
    const double A = (P.ab * nb_b + P.al * nb_l) + P.ad * nb_d;
    o.Te = dw_root4((SL * (1 - A)) / P.sigma);
    o.T = dw_root4(P.q * (A - Al) + dw_pow4(o.Te));
    const double T4 = dw_pow4(o.T);
    o.Tl = dw_root4(P.q2 * (Al - P.al) + T4);
    o.Td = dw_root4(P.q2 * (Al - P.ad) + T4);
    const double dT = P.temp_optimal - o.T, dTl = P.temp_optimal - o.Tl, dTd = P.temp_optimal - o.Td;
    o.beta = 1 - P.g * (dT * dT);
    o.beta_l = 1 - P.g * (dTl * dTl);
    o.beta_d = 1 - P.g * (dTd * dTd);
    const double rb = (P.p - rho_l) - rho_d;
    o.dl = rho_l * (rb * o.beta_l - P.gamma);
    o.dd = rho_d * (rb * o.beta_d - P.gamma);
    o.nl = dw_clip01(l + P.dt * o.dl);
    o.nd = dw_clip01(d + P.dt * o.dd);
    o.nb = (P.p - o.nl) - o.nd;
    return o;
}

// Packed lattice cell: light milli-cover in bits 0..15, dark in bits 16..31 (both 0..1000).
__device__ __forceinline__ uint32_t dw_pack(int kl, int kd) { return (uint32_t)kl | ((uint32_t)kd << 16); }
// k/1000 correctly rounded (== the value np.round(x,3) stores) without a division: Markstein's sequence
// q = k*(1/1000), r = k - 1000q (exact in an FMA), q + r*(1/1000).  Verified exhaustively against k/1000.0 for every
// integer 0 <= k <= 2e6 (tests/test_oracle_c.py::test_markstein_division_exhaustive and the GPU twin).
__device__ __forceinline__ double dw_div1000(double k) {
    const double q = k * 0.001;
    const double r = __fma_rn(-1000.0, q, k);
    return __fma_rn(r, 0.001, q);
}
__device__ __forceinline__ double dw_milli(uint32_t k) {
    return dw_div1000(__hiloint2double(0x43300000, (int)k) - 4503599627370496.0);
}

#define DW_CUDA_TRY(h, expr)                                                           \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) return dw_fail((h), DW_E_CUDA, #expr, cudaGetErrorString(e__)); \
    } while (0)
