// Micro-benchmark of tcgen05.mma throughput for the operand forms the per-edge kernels use (not part of the public header:
// a debug hook like dedf_tc_set_debug).  One CTA issues `reps` MMAs of shape M = 128 x N x (K = 8 tf32 | 16 bf16) back to back
// into one accumulator, with the A operand either in shared memory (descriptor, the no-swizzle chunk-major layout of tc.cuh)
// or in tensor memory, and reports the cycles from the first issue to the completion of the last one.
#include "common.cuh"
#include "tc.cuh"

namespace dedf {

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(int kind, int N, int reps, int a_tmem, int n_ksteps, int n_acc, int commit_every, long long* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char* sA = smem_raw;                                   // [n_ksteps][2][128][16 B]
    unsigned char* sB = sA + (size_t)n_ksteps * 2 * 128 * 16;       // [n_ksteps][2][N][16 B]
    for (int i = tid; i < (n_ksteps * 2 * (128 + N) * 16) / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        // instruction descriptor: D = F32; A/B format 2 = TF32, 1 = BF16; K-major both; N >> 3; M >> 4
        const uint32_t fmt = kind == 0 ? 2u : 1u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        // descriptors of the 8 K steps precomputed; the issue loop is unrolled and does nothing else (the MLP kernel's loop costs
        // ~200 cycles of scalar instructions per MMA, which would hide what the tensor pipe itself needs)
        uint64_t da[8], db[8];
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            da[ks] = tc::smem_desc(a0 + (ks % n_ksteps) * 2 * 128 * 16, 128 * 16, 128);
            db[ks] = tc::smem_desc(b0 + (ks % n_ksteps) * 2 * N * 16, (uint32_t)N * 16, 128);
        }
        const uint32_t step = (n_acc > 1) ? (uint32_t)N : 0u;
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // commit_every = c > 0: a tcgen05.commit (to a barrier nobody waits on) after every c MMAs, like the per-chunk
                // "stage free" commits of the MLP kernel's ring
                if (commit_every > 0 && u > 0 && (u % commit_every) == 0) tc::commit(&bar2);
                const uint32_t dacc = tmem_base + (uint32_t)(u % 4 < n_acc ? u % 4 : 0) * step;
                const uint32_t accf = (uint32_t)(r > 0 || u >= 4);
                if (a_tmem) {
                    const uint32_t ta = tmem_base + 448 + u * 8;
                    if (kind == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                                     ::"r"(dacc), "r"(ta), "l"(db[u]), "r"(idesc), "r"(accf) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                                     ::"r"(dacc), "r"(ta), "l"(db[u]), "r"(idesc), "r"(accf) : "memory");
                } else {
                    if (kind == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(dacc), "l"(da[u]), "l"(db[u]), "r"(idesc), "r"(accf) : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                                     ::"r"(dacc), "l"(da[u]), "l"(db[u]), "r"(idesc), "r"(accf) : "memory");
                }
            }
        }
        const long long t1 = clock64();
        tc::commit(&bar);
        tc::mbar_wait_bounded(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

__device__ __forceinline__ uint32_t probe_elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(pred));
    return pred;
}

// The MLP kernel's issue loop in isolation: 16 K chunks x 3 MMAs (hi.hi, lo.hi, hi.lo) + one commit per `cps` chunks, operands
// resident in shared memory.  mode 0: one thread runs the loop (if lane == 0, as the round-1 kernels do); mode 1: the whole warp
// runs the loop and an elected lane issues (addresses stay warp-uniform: no per-MMA R2UR traffic).
template <int MODE>
__global__ void __launch_bounds__(128, 1) tc_probe_loop_kernel(int N, int nkc, int cps, int reps, long long* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (2 * 128 * 128 * 4 + 2 * 8 * N * 32) / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1 << 20); mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 1 && (MODE >= 1 || lane == 0)) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t a_hi0 = smem_u32(smem_raw), a_lo0 = a_hi0 + 128 * 128 * 4, b0 = a_lo0 + 128 * 128 * 4;
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            uint64_t da_hi = tc::smem_desc(a_hi0, 128 * 16, 128), da_lo = tc::smem_desc(a_lo0, 128 * 16, 128);
            for (int kc = 0; kc < nkc; ++kc) {
                const uint32_t bs = b0 + (uint32_t)(kc & 7) * (uint32_t)N * 64u;
                const uint64_t db_hi = tc::smem_desc(bs, (uint32_t)N * 16, 128), db_lo = tc::smem_desc(bs + (uint32_t)N * 32, (uint32_t)N * 16, 128);
                if (MODE == 0 || probe_elect_one()) {
                    // MODE 2: the three terms of the split go to three accumulators (no MMA depends on the previous one)
                    const uint32_t d1 = MODE == 2 ? tmem_base + (uint32_t)N : tmem_base, d2 = MODE == 2 ? tmem_base + 2u * (uint32_t)N : tmem_base;
                    tc::mma_tf32(tmem_base, da_hi, db_hi, idesc, (r | kc) > 0);
                    tc::mma_tf32(d1, da_lo, db_hi, idesc, MODE == 2 ? (uint32_t)((r | kc) > 0) : 1u);
                    tc::mma_tf32(d2, da_hi, db_lo, idesc, MODE == 2 ? (uint32_t)((r | kc) > 0) : 1u);
                    if ((kc + 1) % cps == 0) tc::commit(&bar2);
                }
                if (MODE >= 1) __syncwarp();
                da_hi += (2u * 128 * 16u) >> 4; da_lo += (2u * 128 * 16u) >> 4;
            }
        }
        const long long t1 = clock64();
        if (MODE == 0 || probe_elect_one()) {
            tc::commit(&bar);
            tc::mbar_wait_bounded(&bar, 0);
            out[0] = t1 - t0; out[1] = clock64() - t0;
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace dedf

extern "C" int dedf_tc_probe_loop(int mode, int N, int nkc, int cps, int reps, long long* out, cudaStream_t stream) {
    if (!out || N < 16 || N > 256 || (N % 16) || nkc < 1 || nkc > 16 || cps < 1 || reps < 1) return DEDF_ERR_ARG;
    const size_t smem = (size_t)2 * 128 * 128 * 4 + (size_t)2 * 8 * N * 32;
    if (smem > 220 * 1024) return DEDF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(dedf::tc_probe_loop_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(dedf::tc_probe_loop_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(dedf::tc_probe_loop_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (mode == 0) dedf::tc_probe_loop_kernel<0><<<1, 128, smem, stream>>>(N, nkc, cps, reps, out);
    else if (mode == 1) dedf::tc_probe_loop_kernel<1><<<1, 128, smem, stream>>>(N, nkc, cps, reps, out);
    else dedf::tc_probe_loop_kernel<2><<<1, 128, smem, stream>>>(N, nkc, cps, reps, out);
    return cudaGetLastError() == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}

/* debug hook: out[0] = cycles to ISSUE `reps` MMAs, out[1] = cycles until the last one has completed */
extern "C" int dedf_tc_probe(int kind, int N, int reps, int a_tmem, int n_ksteps, int n_acc, int commit_every, long long* out, cudaStream_t stream) {
    if (!out || N < 16 || N > 256 || (N % 16) || reps < 1 || n_ksteps < 1 || n_acc < 1 || (n_acc > 4 ? 4 : n_acc) * N > 448 || (reps % 8)) return DEDF_ERR_ARG;
    const size_t smem = (size_t)n_ksteps * 2 * (128 + N) * 16;
    if (smem > 200 * 1024) return DEDF_ERR_UNSUPPORTED;
    cudaFuncSetAttribute(dedf::tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dedf::tc_probe_kernel<<<1, 128, smem, stream>>>(kind, N, reps, a_tmem, n_ksteps, n_acc, commit_every, out);
    return cudaGetLastError() == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}

// Swizzle probe: the same back-to-back issue loop with the operands described as K-major SWIZZLE_128B tiles ([rows][32 floats], 128-byte
// rows, 8-row groups 1024 bytes apart; the K = 8 step advances the start address by 32 bytes inside the row) instead of the no-swizzle
// core-matrix layout.  Timing only (the operands are zeros): does the tensor pipe fetch a swizzled operand faster?
namespace dedf {
__global__ void __launch_bounds__(128, 1) tc_probe_swz_kernel(int N, int reps, int layout, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    // align the tiles to 1024 bytes (swizzle atom)
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = base;                     // [128][128 B]
    unsigned char* sB = sA + 128 * 128;           // [N][128 B]
    for (int i = tid; i < (128 + N) * 32; i += 128) reinterpret_cast<float*>(base)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        uint64_t da[4], db[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (layout == 0) {      // reference: no swizzle, chunk-major (the product's layout), K step = two 16-byte chunks rows*16 apart
                da[ks] = tc::smem_desc(a0 + ks * 2 * 128 * 16, 128 * 16, 128);
                db[ks] = tc::smem_desc(b0 + ks * 2 * N * 16, (uint32_t)N * 16, 128);
            } else {                // swizzled K-major: LBO unused (1), SBO = 8 rows * row bytes, layout type in bits 61..63
                const uint32_t row_bytes = layout == 2 ? 128u : layout == 4 ? 64u : 32u;
                da[ks] = tc::smem_desc(a0 + (ks * 32) % row_bytes, 16, 8 * row_bytes) | ((uint64_t)layout << 61);
                db[ks] = tc::smem_desc(b0 + (ks * 32) % row_bytes, 16, 8 * row_bytes) | ((uint64_t)layout << 61);
            }
        }
        const long long t0 = clock64();
        for (int r = 0; r < reps; r += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) tc::mma_tf32(tmem_base, da[u], db[u], idesc, (uint32_t)(r > 0 || u > 0));
        }
        const long long t1 = clock64();
        tc::commit(&bar);
        tc::mbar_wait_bounded(&bar, 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}
}  // namespace dedf

/* debug hook: layout 0 = no swizzle (chunk-major), 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B */
extern "C" int dedf_tc_probe_swz(int N, int reps, int layout, long long* out, cudaStream_t stream) {
    if (!out || N < 8 || N > 256 || (N % 8) || reps < 4 || (reps % 4) || !(layout == 0 || layout == 2 || layout == 4 || layout == 6)) return DEDF_ERR_ARG;
    const size_t smem = (size_t)(128 + N) * 128 + 2048;
    cudaFuncSetAttribute(dedf::tc_probe_swz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dedf::tc_probe_swz_kernel<<<1, 128, smem, stream>>>(N, reps, layout, out);
    return cudaGetLastError() == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}

// Issue-loop probe: every MMA gets a DIFFERENT descriptor pair (as in the real kernels).  variant 1: one thread runs the loop and
// advances the descriptors in its own registers; variant 2: the whole warp runs the loop on warp-uniform values (kernel parameters,
// loop counters: the descriptors can live in uniform registers) and a lane elected ONCE issues under a predicate; variant 3: like 1
// but the descriptors are rebuilt from scratch per MMA (tc::smem_desc), as the product kernels do.
namespace dedf {
template <int VARIANT>
__global__ void __launch_bounds__(128, 1) tc_probe_issue_kernel(int N, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sA = smem_raw;                 // 16 K-steps x [2][128][16 B]
    unsigned char* sB = sA + 16 * 2 * 128 * 16;   // 16 K-steps x [2][N][16 B]
    for (int i = tid; i < (16 * 2 * (128 + N) * 16) / 4; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (warp == 1) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        const uint64_t da0 = tc::smem_desc(a0, 128 * 16, 128), db0 = tc::smem_desc(b0, (uint32_t)N * 16, 128);
        const uint64_t a_step = (2u * 128u * 16u) >> 4, b_step = (2u * (uint32_t)N * 16u) >> 4;
        if (VARIANT == 2) {
            uint32_t leader = 0;
            asm volatile("{\n.reg .pred px;\nelect.sync _|px, 0xffffffff;\nselp.u32 %0, 1, 0, px;\n}" : "=r"(leader));
            const long long t0 = clock64();
            for (int r = 0; r < reps; r += 16) {
                uint64_t da = da0, db = db0;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 p, %4, 0;\nsetp.ne.b32 q, %5, 0;\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                                 ::"r"(tmem_base), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(r > 0 || u > 0)), "r"(leader) : "memory");
                    da += a_step; db += b_step;
                }
            }
            const long long t1 = clock64();
            if (leader) {
                tc::commit(&bar);
                tc::mbar_wait_bounded(&bar, 0);
                out[0] = t1 - t0; out[1] = clock64() - t0;
            }
        } else if (lane == 0) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; r += 16) {
                uint64_t da = da0, db = db0;
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    if (VARIANT == 3) {
                        da = tc::smem_desc(a0 + u * 2 * 128 * 16, 128 * 16, 128);
                        db = tc::smem_desc(b0 + u * 2 * N * 16, (uint32_t)N * 16, 128);
                    }
                    tc::mma_tf32(tmem_base, da, db, idesc, (uint32_t)(r > 0 || u > 0));
                    da += a_step; db += b_step;
                }
            }
            const long long t1 = clock64();
            tc::commit(&bar);
            tc::mbar_wait_bounded(&bar, 0);
            out[0] = t1 - t0; out[1] = clock64() - t0;
        }
        __syncwarp();
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_base, 512);
}
}  // namespace dedf

extern "C" int dedf_tc_probe_issue(int variant, int N, int reps, long long* out, cudaStream_t stream) {
    if (!out || N < 8 || N > 256 || (N % 8) || reps < 16 || (reps % 16) || variant < 1 || variant > 3) return DEDF_ERR_ARG;
    const size_t smem = (size_t)16 * 2 * (128 + N) * 16;
    cudaFuncSetAttribute(dedf::tc_probe_issue_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dedf::tc_probe_issue_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(dedf::tc_probe_issue_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (variant == 1) dedf::tc_probe_issue_kernel<1><<<1, 128, smem, stream>>>(N, reps, out);
    else if (variant == 2) dedf::tc_probe_issue_kernel<2><<<1, 128, smem, stream>>>(N, reps, out);
    else dedf::tc_probe_issue_kernel<3><<<1, 128, smem, stream>>>(N, reps, out);
    return cudaGetLastError() == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}
