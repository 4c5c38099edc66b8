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


// ---------------------------------------------------------------------------------------------------------------
// Per-edge MLP on the tensor cores: input embedding -> [Linear (+bias / per-pose row bias) (+LayerNorm) (+SiLU)]* -> out
//   UNet blocks:   GaussianRadialBasisLayerFiniteCutoff(length) -> RadialProfile          (radial_func.py:231-278,
//                                                                                          equiformer/radial_func.py:56-59)
//   tensor field:  length encoder -> edge_scalars_pre_linears[scale] (+ time rows) -> SiLU -> RadialProfile, ONE launch
//                                                                     (graph_parser.py:180-183, multiscale_tensor_field.py:225-234)
// Tile = 128 edges = the M of one tcgen05.mma (cta_group::1); accumulator row m lives in TMEM lane m, so the thread that
// owns edge m reads its whole output row with tcgen05.ld and does bias / LayerNorm / SiLU without any cross-thread traffic,
// then writes the next layer's A operand (hi / lo tf32 split, chunk-major) straight back into shared memory.
// Warp roles: warps 0-3 = one edge per thread (input embedding, epilogues, output rows); warp 4 lane 0 = weight producer
// (TMA 1-D bulk copies of pre-packed hi/lo weight chunks into a 4-stage ring, mbarrier complete_tx); warp 5 lane 0 = MMA issuer
// (3 tcgen05.mma per K=8 step: Ahi.Bhi + Alo.Bhi + Ahi.Blo; tcgen05.commit frees ring stages and publishes the accumulator).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTcM = 128;
constexpr int kTcStages = 4;
constexpr int kTcThreads = 192;
constexpr int kTcMaxHidden = 128;      // widest hidden layer (one row lives in 128 registers during its epilogue)

struct MlpTcArgs {
    int mode;                          // DEDF_MLP_IN_RBF / DEDF_MLP_IN_FIELD
    const int* n_edges;
    const float* length;
    const float* rbf_mean; const float* rbf_std_logit; const float* rbf_weight_logit;
    float rbf_cutoff, rbf_offset;
    int n_scales, n_dst;
    const int* row_ptr; const int* edge_dst;
    const float* enc_mean[DEDF_MAX_SCALES]; const float* enc_std_logit[DEDF_MAX_SCALES];
    const float* enc_weight_logit[DEDF_MAX_SCALES];
    float enc_r[DEDF_MAX_SCALES];
    float enc_max_r, enc_n;
    const float* enc_freq;
    const float* row_bias; int n_rb, rb_div;
    int n_layers;
    int K[DEDF_MLP_MAX_LAYERS + 1];
    const float* Wp[DEDF_MLP_MAX_LAYERS];   // packed hi/lo weight chunks; FIELD layer 0: n_scales blocks of 2 K0 K1 floats
    const float* b[DEDF_MLP_MAX_LAYERS];
    const float* ln_g[DEDF_MLP_MAX_LAYERS]; const float* ln_b[DEDF_MLP_MAX_LAYERS];
    int flags[DEDF_MLP_MAX_LAYERS];
    const float* out_offset;
    float* out;
    int a_bytes;                       // bytes of one A operand (hi or lo) = 128 * max K * 4
    int stage_bytes;                   // ring stage size (>= widest N block * 64)
    int tmem_cols;
};

__device__ __forceinline__ int tc_nblocks(int N) { return (N + 255) / 256; }

__global__ void __launch_bounds__(kTcThreads, 1) edge_mlp_tc_kernel(MlpTcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA_hi = smem;
    unsigned char* sA_lo = sA_hi + a.a_bytes;
    unsigned char* sB = sA_lo + a.a_bytes;
    float* s_tab = reinterpret_cast<float*>(sB + kTcStages * a.stage_bytes);     // [n_scales][K0][3] = mean, std, weight
    __shared__ __align__(8) uint64_t full_bar[kTcStages], empty_bar[kTcStages], a_ready, acc_ready;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K0 = a.K[0];
    const bool field = (a.mode == DEDF_MLP_IN_FIELD);
    const int E = *a.n_edges;

    // ---- tiles: FIELD mode keeps a tile inside one scale (the first layer's weights differ per scale) ----
    int n_tiles, tile_base[DEDF_MAX_SCALES + 1];
    if (field) {
        tile_base[0] = 0;
        for (int s = 0; s < a.n_scales; ++s) {
            const int es = a.row_ptr[(size_t)(s + 1) * a.n_dst] - a.row_ptr[(size_t)s * a.n_dst];
            tile_base[s + 1] = tile_base[s] + (es + kTcM - 1) / kTcM;
        }
        n_tiles = tile_base[a.n_scales];
    } else {
        n_tiles = (E + kTcM - 1) / kTcM;
    }
    auto tile_range = [&](int tile, int& e0, int& e1, int& scale) {
        scale = 0;
        if (field) {
            while (tile >= tile_base[scale + 1]) ++scale;
            const int sbeg = a.row_ptr[(size_t)scale * a.n_dst], send = a.row_ptr[(size_t)(scale + 1) * a.n_dst];
            e0 = sbeg + (tile - tile_base[scale]) * kTcM;
            e1 = min(e0 + kTcM, send);
        } else {
            e0 = tile * kTcM;
            e1 = min(e0 + kTcM, E);
        }
    };

    // ---- one-time setup ----
    if (tid == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&a_ready, kTcM);
        mbar_init(&acc_ready, 1);
        mbar_init_fence();
    }
    if (warp == 4) tc::tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid < kTcM) {          // input-encoder tables
        const int ns = field ? a.n_scales : 1;
        for (int i = tid; i < ns * K0; i += kTcM) {
            const int s = i / K0, k = i % K0;
            float mean = 0.f, sd = 1.f, wg = 0.f;
            if (field) {
                if (a.enc_r[s] >= 0.f) {
                    mean = a.enc_mean[s][k];
                    const float sl = a.enc_std_logit[s][k];
                    sd = ((sl > 20.f) ? sl : log1pf(expf(sl))) + 1e-5f;
                    wg = sigmoidf_(a.enc_weight_logit[s][k]) * (4.0f * sqrtf((float)K0));
                }
            } else {
                mean = a.rbf_mean[k];
                const float sl = a.rbf_std_logit[k];
                sd = ((sl > 20.f) ? sl : log1pf(expf(sl))) + 1e-5f;
                wg = sigmoidf_(a.rbf_weight_logit[k]) * 4.0f;
            }
            s_tab[i * 3] = mean; s_tab[i * 3 + 1] = sd; s_tab[i * 3 + 2] = wg;
        }
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 4) {
        // =========================== weight producer ===========================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int e0, e1, scale; tile_range(tile, e0, e1, scale);
                for (int L = 0; L < a.n_layers; ++L) {
                    const int K = a.K[L], N = a.K[L + 1];
                    const int NB = tc_nblocks(N), Nb = N / NB, nkc = K / 8;
                    const uint32_t bytes = (uint32_t)Nb * 64u;
                    const float* Wl = a.Wp[L] + ((field && L == 0) ? (size_t)scale * 2 * K * N : 0);
                    for (int i = 0; i < NB * nkc; ++i) {
                        tc::mbar_wait_bounded(&empty_bar[st], ph ^ 1u);
                        mbar_expect_tx(&full_bar[st], bytes);
                        bulk_g2s(sB + (size_t)st * a.stage_bytes, Wl + (size_t)i * Nb * 16, bytes, &full_bar[st]);
                        if (++st == kTcStages) { st = 0; ph ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            uint32_t st = 0, ph = 0, pa = 0;
            const uint32_t a_hi0 = smem_u32(sA_hi), a_lo0 = smem_u32(sA_lo), b0 = smem_u32(sB);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int L = 0; L < a.n_layers; ++L) {
                    const int K = a.K[L], N = a.K[L + 1];
                    const int NB = tc_nblocks(N), Nb = N / NB, nkc = K / 8;
                    const uint32_t idesc = tc::idesc_tf32(kTcM, Nb);
                    tc::mbar_wait_bounded(&a_ready, pa); pa ^= 1u;      // A operand written, accumulator drained
                    tc::fence_after();
                    for (int nb = 0; nb < NB; ++nb) {
                        for (int kc = 0; kc < nkc; ++kc) {
                            tc::mbar_wait_bounded(&full_bar[st], ph);
                            tc::fence_after();
                            const uint32_t koff = (uint32_t)kc * 2u * kTcM * 16u;
                            const uint64_t da_hi = tc::smem_desc(a_hi0 + koff, kTcM * 16, 128);
                            const uint64_t da_lo = tc::smem_desc(a_lo0 + koff, kTcM * 16, 128);
                            const uint32_t bs = b0 + st * (uint32_t)a.stage_bytes;
                            const uint64_t db_hi = tc::smem_desc(bs, (uint32_t)Nb * 16, 128);
                            const uint64_t db_lo = tc::smem_desc(bs + (uint32_t)Nb * 32, (uint32_t)Nb * 16, 128);
                            const uint32_t d = tmem_base + (uint32_t)(nb * Nb);
                            tc::mma_tf32(d, da_hi, db_hi, idesc, kc > 0);
                            tc::mma_tf32(d, da_lo, db_hi, idesc, 1);
                            tc::mma_tf32(d, da_hi, db_lo, idesc, 1);
                            tc::commit(&empty_bar[st]);                // stage reusable once these MMAs have read it
                            if (++st == kTcStages) { st = 0; ph ^= 1u; }
                        }
                    }
                    tc::commit(&acc_ready);
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== one edge per thread ===========================
        const int m = tid;
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t pc = 0;
        // input embedding of edge e of `scale` -> A operand (hi / lo)
        auto gen_input = [&](int e, bool valid, int scale) {
            const float len = valid ? a.length[e] : 0.f;
            const float* tab = s_tab + (size_t)scale * K0 * 3;
            float xs = 0.f, dd = 0.f, cut = 1.f, nrm = 1.f;
            bool sinus = false;
            if (field) {
                const float r_s = a.enc_r[scale];
                if (r_s >= 0.f) dd = len / r_s;
                else { sinus = true; xs = len / a.enc_max_r * a.enc_n; }
            } else {
                dd = (len - a.rbf_offset) * (1.0f / (a.rbf_cutoff - a.rbf_offset));
                cut = (dd > 0.5f) ? 1.0f : (1.0f - soft_step3(((1.0f - dd) - 0.8f) / (1.0f - 0.8f)));
                nrm = sqrtf((float)K0);
            }
            const int half = K0 / 2;
            for (int k4 = 0; k4 < K0; k4 += 4) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int k = k4 + j;
                    float t;
                    if (sinus) {
                        const int kk = (k < half) ? k : k - half;
                        const float arg = __fmul_rn(xs, a.enc_freq[kk]);
                        t = (k < half) ? sinf(arg) : cosf(arg);
                    } else {
                        const float z = (dd - tab[k * 3]) / tab[k * 3 + 1];
                        t = expf(-0.5f * z * z) * tab[k * 3 + 2];
                        if (!field) t = t * cut * nrm;
                    }
                    v[j] = valid ? t : 0.f;
                }
                float4 hi, lo;
                hi.x = tc::tf32_hi(v[0]); hi.y = tc::tf32_hi(v[1]); hi.z = tc::tf32_hi(v[2]); hi.w = tc::tf32_hi(v[3]);
                lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
                const uint32_t off = (uint32_t)((k4 >> 2) * kTcM + m) * 16u;
                *reinterpret_cast<float4*>(sA_hi + off) = hi;
                *reinterpret_cast<float4*>(sA_lo + off) = lo;
            }
        };
        int tile = blockIdx.x;
        int e0 = 0, e1 = 0, scale = 0;
        if (tile < n_tiles) {
            tile_range(tile, e0, e1, scale);
            gen_input(e0 + m, e0 + m < e1, scale);
            tc::fence_async_smem();
            tc::mbar_arrive(&a_ready);
        }
        for (; tile < n_tiles; tile += gridDim.x) {
            const int e = e0 + m;
            const bool valid = e < e1;
            for (int L = 0; L < a.n_layers; ++L) {
                const int N = a.K[L + 1];
                const bool last = (L == a.n_layers - 1);
                tc::mbar_wait_bounded(&acc_ready, pc); pc ^= 1u;
                tc::fence_after();
                if (!last) {
                    float v[kTcMaxHidden];
#pragma unroll
                    for (int q = 0; q < kTcMaxHidden / 16; ++q) {
                        if (q * 16 < N) {
                            float t[16];
                            tc::tmem_ld16(t_lane + q * 16, t);
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[q * 16 + j] = t[j];
                        }
                    }
                    tc::fence_before();
                    const float* rb = nullptr;
                    if (field && L == 0 && a.row_bias && valid)
                        rb = a.row_bias + ((size_t)scale * a.n_rb + min(a.edge_dst[e] / a.rb_div, a.n_rb - 1)) * N;
                    const float* bL = a.b[L];
                    float s = 0.f;
#pragma unroll
                    for (int c = 0; c < kTcMaxHidden; ++c) {
                        if (c < N) {
                            float t = v[c];
                            if (bL) t += __ldg(bL + c);
                            if (rb) t += __ldg(rb + c);
                            v[c] = t; s += t;
                        }
                    }
                    if (a.flags[L] & 1) {
                        const float mean = s / (float)N;
                        float ss = 0.f;
#pragma unroll
                        for (int c = 0; c < kTcMaxHidden; ++c) if (c < N) { const float t = v[c] - mean; ss += t * t; }
                        const float rstd = rsqrtf(ss / (float)N + 1e-5f);
                        const float* g = a.ln_g[L]; const float* bb = a.ln_b[L];
#pragma unroll
                        for (int c = 0; c < kTcMaxHidden; ++c) if (c < N) v[c] = (v[c] - mean) * rstd * __ldg(g + c) + __ldg(bb + c);
                    }
                    if (a.flags[L] & 2) {
#pragma unroll
                        for (int c = 0; c < kTcMaxHidden; ++c) if (c < N) v[c] = siluf_(v[c]);
                    }
#pragma unroll
                    for (int c4 = 0; c4 < kTcMaxHidden; c4 += 4) {
                        if (c4 < N) {
                            float4 hi, lo;
                            hi.x = tc::tf32_hi(v[c4]); hi.y = tc::tf32_hi(v[c4 + 1]); hi.z = tc::tf32_hi(v[c4 + 2]); hi.w = tc::tf32_hi(v[c4 + 3]);
                            lo.x = v[c4] - hi.x; lo.y = v[c4 + 1] - hi.y; lo.z = v[c4 + 2] - hi.z; lo.w = v[c4 + 3] - hi.w;
                            const uint32_t off = (uint32_t)((c4 >> 2) * kTcM + m) * 16u;
                            *reinterpret_cast<float4*>(sA_hi + off) = hi;
                            *reinterpret_cast<float4*>(sA_lo + off) = lo;
                        }
                    }
                    tc::fence_async_smem();
                    tc::mbar_arrive(&a_ready);
                } else {
                    float* orow = a.out + (size_t)e * N;
                    for (int c0 = 0; c0 < N; c0 += 16) {
                        float t[16];
                        tc::tmem_ld16(t_lane + c0, t);
                        if (valid) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                float4 o = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
                                if (a.b[L]) { o.x += __ldg(a.b[L] + c0 + j); o.y += __ldg(a.b[L] + c0 + j + 1); o.z += __ldg(a.b[L] + c0 + j + 2); o.w += __ldg(a.b[L] + c0 + j + 3); }
                                if (a.out_offset) {
                                    const float4 of = __ldg(reinterpret_cast<const float4*>(a.out_offset + c0 + j));
                                    o.x += of.x; o.y += of.y; o.z += of.z; o.w += of.w;
                                }
                                *reinterpret_cast<float4*>(orow + c0 + j) = o;
                            }
                        }
                    }
                    tc::fence_before();
                    // the accumulator is drained and the A buffers are free: stage the next tile's input right away
                    const int nt = tile + gridDim.x;
                    if (nt < n_tiles) {
                        tile_range(nt, e0, e1, scale);
                        gen_input(e0 + m, e0 + m < e1, scale);
                        tc::fence_async_smem();
                        tc::mbar_arrive(&a_ready);
                    }
                }
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 4) tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
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

extern "C" int dedf_edge_mlp_tc(const dedf_mlp_desc* d, int max_edges, cudaStream_t stream) {
    if (max_edges <= 0) return DEDF_OK;
    if (!d || !d->n_edges_dev || !d->out || !d->length) return DEDF_ERR_ARG;
    if (d->n_layers < 1 || d->n_layers > DEDF_MLP_MAX_LAYERS) return DEDF_ERR_ARG;
    if (d->mode != DEDF_MLP_IN_RBF && d->mode != DEDF_MLP_IN_FIELD) return DEDF_ERR_UNSUPPORTED;
    MlpTcArgs a{};
    a.mode = d->mode; a.n_edges = d->n_edges_dev; a.length = d->length;
    a.rbf_mean = d->rbf_mean; a.rbf_std_logit = d->rbf_std_logit; a.rbf_weight_logit = d->rbf_weight_logit;
    a.rbf_cutoff = d->rbf_cutoff; a.rbf_offset = d->rbf_offset;
    a.n_scales = d->n_scales; a.n_dst = d->n_dst; a.row_ptr = d->row_ptr; a.edge_dst = d->edge_dst;
    a.enc_max_r = d->enc_max_r; a.enc_n = d->enc_n; a.enc_freq = d->enc_freq;
    a.row_bias = d->row_bias; a.n_rb = d->n_rb; a.rb_div = d->rb_div;
    a.n_layers = d->n_layers;
    int max_k = 0, max_nb = 0, max_n = 0;
    for (int i = 0; i <= d->n_layers; ++i) a.K[i] = d->dims[i];
    for (int i = 0; i < d->n_layers; ++i) {
        const int K = a.K[i], N = a.K[i + 1];
        const int NB = (N + 255) / 256;
        if (K < 8 || (K % 8) || K > 128 || N < 16 || (N % NB) || ((N / NB) % 16)) return DEDF_ERR_UNSUPPORTED;
        if (i < d->n_layers - 1 && N > kTcMaxHidden) return DEDF_ERR_UNSUPPORTED;
        if (NB * (N / NB) > 512) return DEDF_ERR_UNSUPPORTED;
        max_k = K > max_k ? K : max_k; max_nb = (N / NB) > max_nb ? (N / NB) : max_nb; max_n = N > max_n ? N : max_n;
        a.Wp[i] = (d->mode == DEDF_MLP_IN_FIELD && i == 0) ? d->pre_w_tc : d->W_tc[i];
        if (!a.Wp[i] || (reinterpret_cast<uintptr_t>(a.Wp[i]) & 15)) return DEDF_ERR_ARG;
        a.b[i] = d->b[i]; a.ln_g[i] = d->ln_g[i]; a.ln_b[i] = d->ln_b[i]; a.flags[i] = d->flags[i];
        if ((a.flags[i] & 1) && (!a.ln_g[i] || !a.ln_b[i])) return DEDF_ERR_ARG;
    }
    if (a.K[0] % 4 || a.K[0] > 128) return DEDF_ERR_UNSUPPORTED;
    if (d->mode == DEDF_MLP_IN_RBF) { if (!a.rbf_mean || !a.rbf_std_logit || !a.rbf_weight_logit) return DEDF_ERR_ARG; }
    else {
        if (!a.row_ptr || !a.edge_dst || a.n_scales < 1 || a.n_scales > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
        if (a.row_bias && (a.rb_div < 1 || a.n_rb < 1)) return DEDF_ERR_ARG;
        for (int s = 0; s < a.n_scales; ++s) {
            a.enc_mean[s] = d->enc_mean[s]; a.enc_std_logit[s] = d->enc_std_logit[s]; a.enc_weight_logit[s] = d->enc_weight_logit[s];
            a.enc_r[s] = d->enc_r[s];
            if (a.enc_r[s] >= 0.f && (!a.enc_mean[s] || !a.enc_std_logit[s] || !a.enc_weight_logit[s])) return DEDF_ERR_ARG;
            if (a.enc_r[s] < 0.f && !a.enc_freq) return DEDF_ERR_ARG;
        }
    }
    a.out_offset = d->out_offset; a.out = d->out;
    a.a_bytes = kTcM * max_k * 4;
    a.stage_bytes = ((max_nb * 64) + 1023) / 1024 * 1024;
    a.tmem_cols = 32;
    while (a.tmem_cols < max_n) a.tmem_cols <<= 1;
    const int ns = (d->mode == DEDF_MLP_IN_FIELD) ? a.n_scales : 1;
    const size_t smem = (size_t)2 * a.a_bytes + (size_t)kTcStages * a.stage_bytes + (size_t)ns * a.K[0] * 3 * sizeof(float);
    if (smem > 220 * 1024) return DEDF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(edge_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr_done = true; }
    const int n_tiles = (max_edges + kTcM - 1) / kTcM + DEDF_MAX_SCALES;
    edge_mlp_tc_kernel<<<grid_for(n_tiles, 1, kNumSMs), kTcThreads, smem, stream>>>(a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
