// Shared device helpers for the Diffusion-EDF score-network kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include "../../include/dedf.h"

namespace dedf {

// e3nn.math.normalize2mom constants (Monte-Carlo values of e3nn==0.4.4's recipe:
// manual_seed(0), randn(1_000_000, float64)); used by fast_activation.py:69 in the
// reference.  tests/test_host.py::test_kernel_constants_match_oracle checks them against the oracle's recomputation.
constexpr float kCSilu = 1.6791767923989418f;
constexpr float kCSigmoid = 1.8467055342154763f;
constexpr float kCSlrelu = 1.531320475574866f;

// SM count of the current device (B200: 148), queried once per process: grids are sized in multiples of it.  Host only.
inline int num_sms() {
    static int n = 0;
    if (n <= 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
#define kNumSMs (::dedf::num_sms())

// return codes: DEDF_OK / DEDF_ERR_* of include/dedf.h

// 1 / (1 + e^-x) with the hardware reciprocal (__fdividef: MUFU.RCP + one multiply, <= 2 ulp).  The IEEE division this replaced
// compiled to a range check + BRANCH (+ a subroutine call on the slow path) per element: ~35 instructions and a reconvergence point
// for every SiLU / gate, which made "activate + split + store" the longest phase of the tensor-core MLP's epilogues (5.7-6.2 k cycles
// for 32 elements per thread; CTA-0 phase stamps, profiles/r2_s9_mlp_tc_timeline_128.txt).
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_(x); }
// SmoothLeakyReLU(0.2): 0.6 x + 0.4 x (2 sigmoid(x) - 1)   (fast_activation.py:14-23)
__device__ __forceinline__ float slreluf_(float x) { return 0.6f * x + 0.4f * x * (2.0f * sigmoidf_(x) - 1.0f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// soft_step(x, n=3) = 4x^3 - 3x^4 on (0,1), 0 below, 1 above   (radial_func.py:16-17)
__device__ __forceinline__ float soft_step3(float x) {
    if (x <= 0.0f) return 0.0f;
    if (x >= 1.0f) return 1.0f;
    float x3 = x * x * x;
    return 4.0f * x3 - 3.0f * x3 * x;
}

// real spherical harmonics l<=2 of a UNIT vector, e3nn 'component' normalisation
// (o3.SphericalHarmonics(normalize=True)).  sh[0..8].
__device__ __forceinline__ void sph_harm_l2(float x, float y, float z, float* sh) {
    const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f, s15 = 3.872983346207417f;
    sh[0] = 1.0f;
    sh[1] = s3 * x;
    sh[2] = s3 * y;
    sh[3] = s3 * z;
    sh[4] = s15 * x * z;
    sh[5] = s15 * x * y;
    sh[6] = s5 * (y * y - 0.5f * (x * x + z * z));
    sh[7] = s15 * y * z;
    sh[8] = 0.5f * s15 * (z * z - x * x);
}

// ---- edge geometry with EXPLICIT rounding -------------------------------------------------------------------------------
// Every multiply / add is an _rn intrinsic, so no compiler context can contract a pair into an FMA: the un-fused
// edge_geom_kernel and the fused head_front_kernel produce the same bits for the same edge.
// (graph_parser.py:146-224: length, o3.SphericalHarmonics(normalize=True, 'component'), non-scalar min-cut, edge logits)
__device__ __forceinline__ float soft_step3_rn(float x) {
    if (x <= 0.0f) return 0.0f;
    if (x >= 1.0f) return 1.0f;
    const float x3 = __fmul_rn(__fmul_rn(x, x), x);
    return __fsub_rn(__fmul_rn(4.0f, x3), __fmul_rn(__fmul_rn(3.0f, x3), x));
}
// vec = x_src - x_dst; ns_hi <= 0: no non-scalar cut; r < 0: all-pairs scale (logit 0); lg may be null
__device__ __forceinline__ void edge_geometry_rn(float vx, float vy, float vz, float ns_lo, float ns_hi, float r, float* len_out,
                                                 float* sh, float* lg) {
    const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), __fmul_rn(vz, vz)));
    const float inv = 1.0f / fmaxf(len, 1e-12f);          // F.normalize
    const float x = __fmul_rn(vx, inv), y = __fmul_rn(vy, inv), z = __fmul_rn(vz, inv);
    const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f, s15 = 3.872983346207417f;
    sh[0] = 1.0f;
    sh[1] = __fmul_rn(s3, x); sh[2] = __fmul_rn(s3, y); sh[3] = __fmul_rn(s3, z);
    sh[4] = __fmul_rn(__fmul_rn(s15, x), z);
    sh[5] = __fmul_rn(__fmul_rn(s15, x), y);
    sh[6] = __fmul_rn(s5, __fsub_rn(__fmul_rn(y, y), __fmul_rn(0.5f, __fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)))));
    sh[7] = __fmul_rn(__fmul_rn(s15, y), z);
    sh[8] = __fmul_rn(0.5f * s15, __fsub_rn(__fmul_rn(z, z), __fmul_rn(x, x)));
    if (ns_hi > 0.f) {                                     // graph_parser.py:174-177, 199-204
        const float c = soft_step3_rn(__fsub_rn(len, ns_lo) / __fsub_rn(ns_hi, ns_lo));
#pragma unroll
        for (int j = 1; j < 9; ++j) sh[j] = __fmul_rn(sh[j], c);
    }
    *len_out = len;
    if (lg) {
        float v = 0.f;
        if (r >= 0.f) {                                    // graph_parser.py:170-173, 206-215
            const float r8 = __fmul_rn(0.8f, r);
            const float cut = __fsub_rn(1.0f, soft_step3_rn(__fsub_rn(len, r8) / __fsub_rn(r, r8)));
            v = logf(fmaxf(cut, 1e-12f));
        }
        *lg = v;
    }
}

// Feature layout of an irreps triple (m0 x0e + m1 x1e + m2 x2e), e3nn mul_ir order.
struct Irr {
    int m0, m1, m2;
    __host__ __device__ int dim() const { return m0 + 3 * m1 + 5 * m2; }
    __host__ __device__ int off1() const { return m0; }
    __host__ __device__ int off2() const { return m0 + 3 * m1; }
    __host__ __device__ int nirr() const { return m0 + m1 + m2; }
};

// ---- TMA (1-D bulk async copy, SASS UBLKCP) + mbarrier helpers ---------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// make an mbarrier initialised by this thread visible to the async proxy before a bulk copy signals it
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte aligned) in chunks the tx-count can hold
__device__ __forceinline__ void bulk_g2s_chunked(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    constexpr uint32_t kChunk = 32768;
#pragma unroll 1       // one thread issues a handful of copies: unrolled, these loops were a quarter of node_chain_kernel's code
    for (uint32_t o = 0; o < bytes; o += kChunk)
        bulk_g2s(static_cast<char*>(dst) + o, static_cast<const char*>(src) + o, min(kChunk, bytes - o), bar);
}


inline int grid_for(long long work_items, int per_block, int max_blocks) {
    long long b = (work_items + per_block - 1) / per_block;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Kernels launched through launch_pdl() may start while the previous kernel of the stream is still running: everything a
// kernel does BEFORE pdl_wait() (barrier init, TMEM allocation, TMA prefetch of weights, parameter tables) overlaps with
// the tail of its predecessor; pdl_wait() returns once the predecessor has completed and its writes are visible.  Every
// kernel launched this way calls pdl_wait() unconditionally in every thread before touching anything a previous kernel
// produced (so completion stays transitive along the stream) and then pdl_launch() to let its own successor start.
// DEDF_PDL=0 in the environment turns the attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DEDF_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define DEDF_CHECK_LAUNCH()                                         \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return DEDF_ERR_LAUNCH;             \
    } while (0)

}  // namespace dedf
