// Tensor-core (tcgen05, kind::tf32, 3xTF32 split) kernels: self-test GEMM and the per-edge MLP.
#include "common.cuh"
#include "tc.cuh"
#include "../../include/dedf.h"

namespace dedf {

// ---------------------------------------------------------------------------------------------------------------
// Self-test: D[128, N] = A[128, K] . B[N, K]^T through the same descriptor / TMEM / 3xTF32 path the MLP kernel uses.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int K,
                                                            int n_split, float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sA_hi = smem_raw;
    unsigned char* sA_lo = sA_hi + 128 * K * 4;
    unsigned char* sB_hi = sA_lo + 128 * K * 4;
    unsigned char* sB_lo = sB_hi + N * K * 4;
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        const float v = A[i], hi = tc::tf32_hi(v);
        *reinterpret_cast<float*>(sA_hi + tc::cm_off(128, r, k)) = hi;
        *reinterpret_cast<float*>(sA_lo + tc::cm_off(128, r, k)) = (n_split > 1) ? v - hi : 0.f;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float v = B[i], hi = tc::tf32_hi(v);
        *reinterpret_cast<float*>(sB_hi + tc::cm_off(N, r, k)) = hi;
        *reinterpret_cast<float*>(sB_lo + tc::cm_off(N, r, k)) = (n_split > 1) ? v - hi : 0.f;
    }
    uint32_t cols = 32;
    while ((int)cols < N) cols <<= 1;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, cols);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t a_hi = tc::smem_desc(smem_u32(sA_hi) + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t a_lo = tc::smem_desc(smem_u32(sA_lo) + ks * 2 * 128 * 16, 128 * 16, 128);
            const uint64_t b_hi = tc::smem_desc(smem_u32(sB_hi) + ks * 2 * N * 16, N * 16, 128);
            const uint64_t b_lo = tc::smem_desc(smem_u32(sB_lo) + ks * 2 * N * 16, N * 16, 128);
            tc::mma_tf32(tmem_base, a_hi, b_hi, idesc, ks > 0);
            if (n_split > 1) {
                tc::mma_tf32(tmem_base, a_lo, b_hi, idesc, 1);
                tc::mma_tf32(tmem_base, a_hi, b_lo, idesc, 1);
            }
        }
        tc::commit(&bar);
    }
    tc::mbar_wait_bounded(&bar, 0);
    tc::fence_after();
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tc::tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)tid * N + c0 + j] = v[j];
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, cols);
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_tc_selftest(const float* A, const float* B, int N, int K, int n_split, float* D, cudaStream_t stream) {
    if (!A || !B || !D || N < 16 || N > 256 || (N % 16) || K < 8 || (K % 8)) return DEDF_ERR_ARG;
    const size_t smem = (size_t)2 * (128 + N) * K * 4;
    if (smem > 200 * 1024) return DEDF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    tc_selftest_kernel<<<1, 128, smem, stream>>>(A, B, N, K, n_split, D);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
