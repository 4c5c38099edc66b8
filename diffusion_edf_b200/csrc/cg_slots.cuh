// Depthwise ('uvu', mul2 = 1) Clebsch-Gordan tensor product of an irreps triple
// (M0 x0e + M1 x1e + M2 x2e) with the l<=2 spherical harmonics: per-channel
// micro-kernels and the lane <-> channel "slot" mapping shared by the fused edge
// kernels.
//
// Restates o3.TensorProduct as built by DepthwiseTensorProduct
// (/root/reference/diffusion_edf/equiformer/tensor_product_rescale.py:352-382):
// 15 paths in creation order k, each
//     out_k[u, :] = sqrt(2 lo + 1) * w_k[u] * sum_ij C_ijk x[u, i] sh[j].
//
// Every irreps set used by the reference's configs satisfies M0 = 2 G, M1 = G,
// M2 = G / 2 (G = 32: 64x0e+32x1e+16x2e, G = 16: 32x0e+16x1e+8x2e).  A warp works
// on a PACK of P = 64 / G edges at once so that all 32 lanes run the same
// instruction stream: 4 "l0 slots", 2 "l1 slots" and 1 "l2 slot" per lane.
#pragma once
#include "common.cuh"
#include "cg_paths.cuh"

namespace dedf {

template <int G>
struct Dtp {
    static constexpr int M0 = 2 * G, M1 = G, M2 = G / 2;
    static constexpr int F = M0 + 3 * M1 + 5 * M2;          // input feature dim (240 / 120)
    static constexpr int P = 32 / M2;                        // edges per pack (2 / 4)
    static constexpr int NUMEL = 3 * M0 + 6 * M1 + 6 * M2;   // weights per edge (480 / 240)
    static constexpr int D0 = M0 + M1 + M2;                  // 0e channels of the output (112 / 56)
    static constexpr int D1 = M0 + 3 * M1 + 2 * M2;          // 1e channels (192 / 96)
    static constexpr int D2 = M0 + 2 * M1 + 3 * M2;          // 2e channels (176 / 88)
    static constexpr int FOUT = D0 + 3 * D1 + 5 * D2;        // 1568 / 784
    // weight offsets of path k (creation order)
    static constexpr int W_K0 = 0, W_K1 = M0, W_K2 = 2 * M0, W_K3 = 3 * M0;
    static constexpr int W_K9 = 3 * M0 + 6 * M1;
    // column of path k inside its l_out block (i_out order)
    static constexpr int C0_K0 = 0, C0_K4 = M0, C0_K12 = M0 + M1;
    static constexpr int C1_K1 = 0, C1_K3 = M0, C1_K5 = M0 + M1, C1_K7 = M0 + 2 * M1, C1_K10 = M0 + 3 * M1,
                         C1_K13 = M0 + 3 * M1 + M2;
    static constexpr int C2_K2 = 0, C2_K6 = M0, C2_K8 = M0 + M1, C2_K9 = M0 + 2 * M1, C2_K11 = M0 + 2 * M1 + M2,
                         C2_K14 = M0 + 2 * M1 + 2 * M2;
    // slot -> (edge in pack, channel)
    __device__ static __forceinline__ void slot0(int lane, int s, int& e, int& ch) { int j = lane + 32 * s; e = j / M0; ch = j % M0; }
    __device__ static __forceinline__ void slot1(int lane, int s, int& e, int& ch) { int j = lane + 32 * s; e = j / M1; ch = j % M1; }
    __device__ static __forceinline__ void slot2(int lane, int& e, int& ch) { e = lane / M2; ch = lane % M2; }
};

// l1 = 0 channel: x scalar, weights (k0,k1,k2).  o[0] -> lo=0 ; o[1..3] -> lo=1 ; o[4..8] -> lo=2
__device__ __forceinline__ void dtp_l0(float x, float w0, float w1, float w2, const float* __restrict__ sh, float* __restrict__ o) {
    const float a0 = x * w0, a1 = x * w1, a2 = x * w2;
    o[0] = a0 * sh[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[1 + k] = a1 * sh[1 + k];
#pragma unroll
    for (int k = 0; k < 5; ++k) o[4 + k] = a2 * sh[4 + k];
}

// l1 = 1 channel: x[3], weights w[0..5] = (k3..k8).
// o[0..2]=k3 (lo1) ; o[3]=k4 (lo0) ; o[4..6]=k5 (lo1) ; o[7..11]=k6 (lo2) ; o[12..14]=k7 (lo1) ; o[15..19]=k8 (lo2)
__device__ __forceinline__ void dtp_l1(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ sh, float* __restrict__ o) {
    float t[5];
    const float a = w[0] * sh[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = a * x[k];
    cg_110(x, sh + 1, t); o[3] = w[1] * t[0];
    cg_111(x, sh + 1, t);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[4 + k] = w[2] * t[k];
    cg_112(x, sh + 1, t);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[7 + k] = w[3] * t[k];
    cg_121(x, sh + 4, t);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[12 + k] = w[4] * t[k];
    cg_122(x, sh + 4, t);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[15 + k] = w[5] * t[k];
}

// l1 = 2 channel: x[5], weights w[0..5] = (k9..k14).
// o[0..4]=k9 (lo2) ; o[5..7]=k10 (lo1) ; o[8..12]=k11 (lo2) ; o[13]=k12 (lo0) ; o[14..16]=k13 (lo1) ; o[17..21]=k14 (lo2)
__device__ __forceinline__ void dtp_l2(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ sh, float* __restrict__ o) {
    float t[5];
    const float a = w[0] * sh[0];
#pragma unroll
    for (int k = 0; k < 5; ++k) o[k] = a * x[k];
    cg_211(x, sh + 1, t);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[5 + k] = w[1] * t[k];
    cg_212(x, sh + 1, t);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[8 + k] = w[2] * t[k];
    cg_220(x, sh + 4, t); o[13] = w[3] * t[0];
    cg_221(x, sh + 4, t);
#pragma unroll
    for (int k = 0; k < 3; ++k) o[14 + k] = w[4] * t[k];
    cg_222(x, sh + 4, t);
#pragma unroll
    for (int k = 0; k < 5; ++k) o[17 + k] = w[5] * t[k];
}

}  // namespace dedf
