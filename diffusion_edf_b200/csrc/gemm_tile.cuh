// SIMT fp32 tile-GEMM micro-kernel shared by the edge and node kernels.
//
// A lives in shared memory, row-major [rows][lda] (K contiguous, lda % 4 == 0 and
// lda % 32 == 4 or 20 so that consecutive row-groups hit distinct banks); W lives
// in global memory, row-major [K][N] ("in x out"), read through L1 (every CTA
// reads the same few weight matrices), or in shared memory (kWShared) when the
// caller staged the whole matrix there to take L2 latency off the K loop.  One work item = 4 rows x 4 columns; the 4
// rows of item (rg, cg) are rg, rg+n_rg, rg+2 n_rg, rg+3 n_rg.
//
// fp32 FMA is used on purpose: the reference is an fp32 network checked to 1e-4
// relative, which TF32 tensor-core MMA (10-bit mantissa) cannot meet without a
// 3xTF32 split; see DESIGN.md "precision".
#pragma once
#include "common.cuh"

namespace dedf {

// acc[i][j] += sum_k A[row_i][k] * W[k][col0 + j]     (K % 4 == 0, N % 4 == 0)
//
// kPrefetch > 0 (global W only): every iteration also issues one `prefetch.global.L1` for the W row this thread's
// column group needs kPrefetch k-steps ahead (row k + kPrefetch + (rg & 3): the 4 threads that share a column group
// cover the 4 rows of a step between them).  The K loop of the edge kernels is bound by the L2 latency of the first
// touch of each weight row (profiles/r1_tp_lin_act_ncu_full_summary.txt); the prefetch turns it into an L1 hit.
template <bool kVecW, bool kWShared = false, int kPrefetch = 0>
__device__ __forceinline__ void gemm_item_4x4(const float* __restrict__ A, int lda, int n_rg, int rg,
                                              const float* __restrict__ W, int N, int col0, int K,
                                              float acc[4][4]) {
    const float* a0 = A + (size_t)rg * lda;
    const float* a1 = a0 + (size_t)n_rg * lda;
    const float* a2 = a1 + (size_t)n_rg * lda;
    const float* a3 = a2 + (size_t)n_rg * lda;
    const int K4 = K & ~3;
#pragma unroll 2
    for (int k = 0; k < K4; k += 4) {
        if (kPrefetch > 0 && !kWShared) {
            const int kp = k + kPrefetch + (rg & 3);
            if (kp < K) asm volatile("prefetch.global.L1 [%0];" ::"l"(W + (size_t)kp * N + col0));
        }
        const float4 x0 = *reinterpret_cast<const float4*>(a0 + k);
        const float4 x1 = *reinterpret_cast<const float4*>(a1 + k);
        const float4 x2 = *reinterpret_cast<const float4*>(a2 + k);
        const float4 x3 = *reinterpret_cast<const float4*>(a3 + k);
        float w[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            if (kVecW) {
                const float4 t = kWShared ? *reinterpret_cast<const float4*>(W + (size_t)(k + kk) * N + col0)
                                          : __ldg(reinterpret_cast<const float4*>(W + (size_t)(k + kk) * N + col0));
                w[kk][0] = t.x; w[kk][1] = t.y; w[kk][2] = t.z; w[kk][3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) w[kk][j] = (col0 + j < N) ? (kWShared ? W[(size_t)(k + kk) * N + col0 + j] : __ldg(W + (size_t)(k + kk) * N + col0 + j)) : 0.f;
            }
        }
        const float xa[4][4] = {{x0.x, x0.y, x0.z, x0.w}, {x1.x, x1.y, x1.z, x1.w}, {x2.x, x2.y, x2.z, x2.w}, {x3.x, x3.y, x3.z, x3.w}};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i][kk], w[kk][j], acc[i][j]);
    }
    for (int k = K4; k < K; ++k) {   // K tail (e.g. the 3x0e input embedding)
        const float xa[4] = {a0[k], a1[k], a2[k], a3[k]};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float wv = (col0 + j < N) ? (kWShared ? W[(size_t)k * N + col0 + j] : __ldg(W + (size_t)k * N + col0 + j)) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(xa[i], wv, acc[i][j]);
        }
    }
}

__host__ __device__ inline int pad_lda(int K) {
    // smallest lda >= K + 4 with lda % 4 == 0 and (lda % 32) in {4, 12, 20, 28}
    int lda = (K + 3) & ~3;
    lda += 4;
    while (((lda & 31) & 7) != 4) lda += 4;
    return lda;
}

}  // namespace dedf
