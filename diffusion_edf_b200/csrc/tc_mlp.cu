// Tensor-core (tcgen05) kernels: self-test GEMM (kind::tf32, 3xTF32 split) and the per-edge MLP (kind::f16 with an fp16 hi / lo
// operand split by default, kind::tf32 3xTF32 as the F16 = false variant: TcGroup below).
#include <cuda_fp16.h>
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
// Warp roles: warps 0-15 = 4 warpgroups; a warp may only touch TMEM lanes 32*(warp%4)..+31, so the four threads
// (warp%4, lane) of the 4 warpgroups share edge row 32*(warp%4)+lane and split its columns (LayerNorm statistics are exchanged
// through shared memory); warp 16 lane 0 = weight producer (TMA 1-D bulk copies of pre-packed hi/lo weight chunks into a
// 4-stage ring, mbarrier complete_tx); warp 17 lane 0 = MMA issuer (3 tcgen05.mma per K=8 step: Ahi.Bhi + Alo.Bhi + Ahi.Blo;
// tcgen05.commit frees ring stages and publishes the accumulator).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTcM = 128;
constexpr int kTcStages = 4;            // ring capacity in units of the widest stage (stage_bytes): ring bytes = kTcStages * stage_bytes
constexpr int kTcMaxStages = 16;        // narrower layers cut the same ring into more, smaller stages (see the producer)
constexpr int kTcEpiWG = 4;            // epilogue warpgroups: the 4 threads with the same (warp % 4, lane) share one edge row
constexpr int kTcEpiThreads = kTcEpiWG * kTcM;          // 512
constexpr int kTcProdWarp = kTcEpiThreads / 32;         // warp 16: weight producer (+ TMEM alloc)
constexpr int kTcMmaWarp = kTcProdWarp + 1;             // warp 17: MMA issuer
constexpr int kTcThreads = kTcEpiThreads + 64;          // 576
constexpr int kTcMaxHidden = 128;      // widest hidden layer

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
    int out_cw;                        // output staging chunk width (columns)
    long long* dbg;                    // optional (host debug): per-phase clock64() stamps of CTA 0
};

__device__ __forceinline__ int tc_nblocks(int N) { return (N + 255) / 256; }
// K chunks per ring stage: as many as fit in HALF the ring (two stages in flight at least), a divisor of the layer's chunk count
__device__ __forceinline__ int tc_chunks_per_stage(uint32_t ring_bytes, uint32_t chunk_bytes, int nkc) {
    int cps = (int)((ring_bytes / 2) / chunk_bytes);
    cps = max(1, min(cps, nkc));
    while (nkc % cps) --cps;
    return cps;
}

// hi / lo operand group of one edge row: KG consecutive K columns = one 16-byte core-matrix row.
//   tf32 split (F16 = false): 4 floats, hi = 13 low mantissa bits cleared, lo = x - hi, K = 8 per MMA
//   fp16 split (F16 = true):  8 halves, hi = fp16(x), lo = fp16(x - hi), K = 16 per MMA: half the MMAs, half the weight bytes
//                             streamed per tile and half the operand bytes written / fetched for the same products (see the
//                             note on edge_tp_act_tc_kernel; the activations here are O(1) post-LayerNorm values, the inputs radial
//                             basis values <= 4 sqrt(K0): far inside fp16 range)
template <bool F16> struct TcGroup;
template <> struct TcGroup<false> {
    static constexpr int KG = 4;
    static __device__ __forceinline__ void store(unsigned char* hi_p, unsigned char* lo_p, const float* v) {
        float4 hi, lo;
        hi.x = tc::tf32_hi(v[0]); hi.y = tc::tf32_hi(v[1]); hi.z = tc::tf32_hi(v[2]); hi.w = tc::tf32_hi(v[3]);
        lo.x = v[0] - hi.x; lo.y = v[1] - hi.y; lo.z = v[2] - hi.z; lo.w = v[3] - hi.w;
        *reinterpret_cast<float4*>(hi_p) = hi;
        *reinterpret_cast<float4*>(lo_p) = lo;
    }
};
template <> struct TcGroup<true> {
    static constexpr int KG = 8;
    static __device__ __forceinline__ void store(unsigned char* hi_p, unsigned char* lo_p, const float* v) {
        uint32_t H[4], Lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            const float2 f = __half22float2(h);
            const __half2 l = __floats2half2_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
            H[i] = *reinterpret_cast<const uint32_t*>(&h); Lo[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        *reinterpret_cast<uint4*>(hi_p) = make_uint4(H[0], H[1], H[2], H[3]);
        *reinterpret_cast<uint4*>(lo_p) = make_uint4(Lo[0], Lo[1], Lo[2], Lo[3]);
    }
};

template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1) edge_mlp_tc_kernel(MlpTcArgs a) {
    using GR = TcGroup<F16>;
    constexpr int KG = GR::KG;               // K columns per 16-byte operand group
    constexpr int KM = 2 * KG;               // K per MMA (two groups)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA_hi = smem;
    unsigned char* sA_lo = sA_hi + a.a_bytes;
    unsigned char* sB = sA_lo + a.a_bytes;
    float* s_tab = reinterpret_cast<float*>(sB + kTcStages * a.stage_bytes);     // [n_scales][K0][4] = mean, 1/std, weight, -
    float* s_freq = s_tab + (size_t)(a.mode == DEDF_MLP_IN_FIELD ? a.n_scales : 1) * a.K[0] * 4;   // [K0/2] sinusoidal frequencies
    float* s_par = s_freq + ((a.K[0] / 2 + 3) & ~3);                               // hidden layer L: [b | ln_g | ln_b] x kTcMaxHidden
    float* s_last = s_par + (DEDF_MLP_MAX_LAYERS - 1) * 3 * kTcMaxHidden;          // last layer: bias + offset, [N_last]
    __shared__ __align__(8) uint64_t full_bar[kTcMaxStages], empty_bar[kTcMaxStages], a_ready, acc_ready, layer_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K0 = a.K[0];
    const bool field = (a.mode == DEDF_MLP_IN_FIELD);
    int dbg_i = 0;
#define TC_STAMP(tidx) do { if (a.dbg && blockIdx.x == 0 && tid == (tidx) && dbg_i < 64) a.dbg[(tidx == 0 ? 0 : (tidx == kTcProdWarp * 32 ? 64 : 128)) + dbg_i++] = clock64(); } while (0)
    TC_STAMP(0); TC_STAMP(kTcProdWarp * 32); TC_STAMP(kTcMmaWarp * 32);

    // ---- one-time setup ----
    if (tid == 0) {
        for (int s = 0; s < kTcMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&layer_done, 1);
        mbar_init(&a_ready, kTcEpiThreads);
        mbar_init(&acc_ready, 1);
        mbar_init_fence();
    }
    if (warp == kTcProdWarp) tc::tmem_alloc(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid < kTcM) {          // input-encoder / parameter tables
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
            s_tab[i * 4] = mean; s_tab[i * 4 + 1] = 1.0f / sd; s_tab[i * 4 + 2] = wg; s_tab[i * 4 + 3] = 0.f;
        }
        if (field && a.enc_freq) for (int i = tid; i < K0 / 2; i += kTcM) s_freq[i] = a.enc_freq[i];
        for (int L = 0; L + 1 < a.n_layers; ++L) {
            const int N = a.K[L + 1];
            for (int c = tid; c < N; c += kTcM) {
                s_par[(L * 3 + 0) * kTcMaxHidden + c] = a.b[L] ? a.b[L][c] : 0.f;
                s_par[(L * 3 + 1) * kTcMaxHidden + c] = (a.flags[L] & 1) ? a.ln_g[L][c] : 1.f;
                s_par[(L * 3 + 2) * kTcMaxHidden + c] = (a.flags[L] & 1) ? a.ln_b[L][c] : 0.f;
            }
        }
        {
            const int L = a.n_layers - 1, N = a.K[L + 1];
            for (int c = tid; c < N; c += kTcM) s_last[c] = (a.b[L] ? a.b[L][c] : 0.f) + (a.out_offset ? a.out_offset[c] : 0.f);
        }
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait(); pdl_launch();     // PDL: everything above (barriers, TMEM, parameter tables) overlapped the previous kernel
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
    TC_STAMP(0); TC_STAMP(kTcProdWarp * 32); TC_STAMP(kTcMmaWarp * 32);

    if (warp == kTcProdWarp) {
        // =========================== weight producer ===========================
        if (lane == 0) {
            // The ring (kTcStages * stage_bytes) is cut into stages of ONE K chunk of ONE N block of the CURRENT layer: a 128-wide
            // layer gets 8 stages of 8 KB, a 64-wide one 16 of 4 KB, the 240-wide blocks of the last layer 4 of 15 KB.  A chunk's
            // MMAs are bound by the round trip "MMAs done -> commit -> refill (L2, every CTA asks for the same lines at the same
            // time) -> full": ~2400 cycles measured (profiles/r2_s1_mlp_tc_timeline_128.txt: 600 cycles per chunk at 4 stages, whatever
            // the chunk's size), so the stage COUNT is what hides it.  Stage geometry changes at a layer boundary: the producer
            // waits for the previous layer's last MMA (layer_done) before it writes with the new geometry; it still runs ahead of
            // the issuer by the whole ring during the epilogue between two layers.
            uint32_t st = 0, pmask = 0, pl = 0;            // pmask bit s: parity the next wait on empty_bar[s] uses
            bool first_layer = true;
            const uint32_t ring_bytes = (uint32_t)kTcStages * (uint32_t)a.stage_bytes;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                int e0, e1, scale; tile_range(tile, e0, e1, scale);
                for (int L = 0; L < a.n_layers; ++L) {
                    const int K = a.K[L], N = a.K[L + 1];
                    const int NB = tc_nblocks(N), Nb = N / NB, nkc = K / KM;
                    const uint32_t cbytes = (uint32_t)Nb * 64u;                       // one K chunk of one N block (hi + lo)
                    const int cps = tc_chunks_per_stage(ring_bytes, cbytes, nkc);      // K chunks per stage (divides nkc)
                    const uint32_t bytes = cbytes * (uint32_t)cps;
                    const uint32_t ns = min((uint32_t)kTcMaxStages, ring_bytes / bytes);
                    const float* Wl = a.Wp[L] + ((field && L == 0) ? (size_t)scale * (F16 ? 1 : 2) * K * N : 0);
                    TC_STAMP(kTcProdWarp * 32);
                    if (!first_layer) { tc::mbar_wait_bounded(&layer_done, pl); pl ^= 1u; }
                    first_layer = false;
                    st = 0;
                    for (int i = 0; i < NB * nkc; i += cps) {                          // (the chunks of a stage are contiguous in Wp)
                        tc::mbar_wait_bounded(&empty_bar[st], ((pmask >> st) & 1u) ^ 1u);
                        pmask ^= 1u << st;
                        mbar_expect_tx(&full_bar[st], bytes);
                        bulk_g2s_chunked(sB + (size_t)st * bytes, Wl + (size_t)i * Nb * 16, bytes, &full_bar[st]);
                        if (++st == ns) st = 0;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == kTcMmaWarp) {
        // =========================== MMA issuer ===========================
        // warp-uniform issue loop (tc.cuh: mma_tf32_if): all 32 lanes run it on uniform values (descriptors in uniform registers,
        // no per-MMA R2UR chain), the lane elected here issues
        {
            const uint32_t leader = tc::elect_one();
            uint32_t st = 0, cmask = 0, pa = 0;            // cmask bit s: parity the next wait on full_bar[s] uses
            const uint32_t ring_bytes = (uint32_t)kTcStages * (uint32_t)a.stage_bytes;
            const uint32_t a_hi0 = smem_u32(sA_hi), a_lo0 = smem_u32(sA_lo), b0 = smem_u32(sB);
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int L = 0; L < a.n_layers; ++L) {
                    const int K = a.K[L], N = a.K[L + 1];
                    const int NB = tc_nblocks(N), Nb = N / NB, nkc = K / KM;
                    const uint32_t idesc = F16 ? tc::idesc_f16(kTcM, Nb) : tc::idesc_tf32(kTcM, Nb);
                    TC_STAMP(kTcMmaWarp * 32);
                    tc::mbar_wait_bounded(&a_ready, pa); pa ^= 1u;      // A operand written, accumulator drained
                    tc::fence_after();
                    TC_STAMP(kTcMmaWarp * 32);
                    const uint64_t da_hi_l = tc::smem_desc(a_hi0, kTcM * 16, 128), da_lo_l = tc::smem_desc(a_lo0, kTcM * 16, 128);
                    constexpr uint64_t kAStep = (2u * kTcM * 16u) >> 4;          // address-field increment per K chunk
                    const uint32_t cbytes = (uint32_t)Nb * 64u;
                    const int cps = tc_chunks_per_stage(ring_bytes, cbytes, nkc);
                    const uint32_t sbytes = cbytes * (uint32_t)cps;
                    const uint32_t ns = min((uint32_t)kTcMaxStages, ring_bytes / sbytes);     // this layer's stage count (see the producer)
                    st = 0;
                    for (int nb = 0; nb < NB; ++nb) {
                        const uint32_t d = tmem_base + (uint32_t)(nb * Nb);
                        uint64_t da_hi = da_hi_l, da_lo = da_lo_l;
                        for (int kc = 0; kc < nkc; kc += cps) {
                            // one wait, one fence and one commit per STAGE of cps chunks (3 cps MMAs): their fixed cost is what a
                            // one-chunk stage spent most of its ~600 cycles on
                            tc::mbar_wait_bounded(&full_bar[st], (cmask >> st) & 1u);
                            cmask ^= 1u << st;
                            tc::fence_after();
                            uint32_t bs = b0 + st * sbytes;
                            for (int c = 0; c < cps; ++c) {
                                const uint64_t db_hi = tc::smem_desc(bs, (uint32_t)Nb * 16, 128);
                                const uint64_t db_lo = tc::smem_desc(bs + (uint32_t)Nb * 32, (uint32_t)Nb * 16, 128);
                                if constexpr (F16) {
                                    tc::mma_f16_if(leader, d, da_hi, db_hi, idesc, (kc + c) > 0);
                                    tc::mma_f16_if(leader, d, da_lo, db_hi, idesc, 1);
                                    tc::mma_f16_if(leader, d, da_hi, db_lo, idesc, 1);
                                } else {
                                    tc::mma_tf32_if(leader, d, da_hi, db_hi, idesc, (kc + c) > 0);
                                    tc::mma_tf32_if(leader, d, da_lo, db_hi, idesc, 1);
                                    tc::mma_tf32_if(leader, d, da_hi, db_lo, idesc, 1);
                                }
                                da_hi += kAStep; da_lo += kAStep;
                                bs += cbytes;
                            }
                            tc::commit_if(leader, &empty_bar[st]);     // stage reusable once these MMAs have read it
                            if (++st == ns) st = 0;
                        }
                    }
                    tc::commit_if(leader, &layer_done);                // the producer may re-cut the ring for the next layer
                    tc::commit_if(leader, &acc_ready);
                    TC_STAMP(kTcMmaWarp * 32);
                }
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue warpgroups: 4 threads per edge row ===========================
        const int wg = warp >> 2;                                   // column slice of this thread
        const int m = ((warp & 3) << 5) | lane;                     // edge row = TMEM lane
        const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        float* s_red = s_last + 512;                                // [2][kTcEpiWG][kTcM] LayerNorm partial sums
        uint32_t pc = 0;
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); };
        // input embedding of edge e of `scale` -> this thread's K0/4 columns of the A operand (hi / lo)
        auto gen_input = [&](int e, bool valid, int scale) {
            const float len = valid ? a.length[e] : 0.f;
            const float* tab = s_tab + (size_t)scale * K0 * 4;
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
            for (int k4 = wg * KG; k4 < K0; k4 += KG * kTcEpiWG) {
                float v[KG];
#pragma unroll
                for (int j = 0; j < KG; ++j) {
                    const int k = k4 + j;
                    float t;
                    if (sinus) {
                        const int kk = (k < half) ? k : k - half;
                        const float arg = __fmul_rn(xs, s_freq[kk]);
                        t = (k < half) ? sinf(arg) : cosf(arg);
                    } else {
                        const float4 p = *reinterpret_cast<const float4*>(tab + k * 4);
                        const float z = (dd - p.x) * p.y;
                        t = expf(-0.5f * z * z) * p.z;
                        if (!field) t = t * cut * nrm;
                    }
                    v[j] = valid ? t : 0.f;
                }
                const uint32_t off = (uint32_t)((k4 / KG) * kTcM + m) * 16u;
                GR::store(sA_hi + off, sA_lo + off, v);
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
                TC_STAMP(0);
                tc::mbar_wait_bounded(&acc_ready, pc); pc ^= 1u;
                tc::fence_after();
                TC_STAMP(0);
                if (!last) {
                    // this thread's columns [c_lo, c_lo + 16 n16): 32 per thread at N = 128, 16 at N = 64, fewer warpgroups below
                    const int n16 = (N >= 128) ? 2 : 1;
                    const int c_lo = wg * 16 * n16;
                    const bool act_cols = c_lo < N;
                    float v[32];
                    const float* rb = nullptr;
                    if (field && L == 0 && a.row_bias && valid)
                        rb = a.row_bias + ((size_t)scale * a.n_rb + min(a.edge_dst[e] / a.rb_div, a.n_rb - 1)) * N;
                    const float* pb = s_par + (L * 3 + 0) * kTcMaxHidden;
                    const float* pg = s_par + (L * 3 + 1) * kTcMaxHidden;
                    const float* pbb = s_par + (L * 3 + 2) * kTcMaxHidden;
                    float s = 0.f;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (q < n16 && act_cols) {
                            float t[16];
                            tc::tmem_ld16(t_lane + c_lo + q * 16, t);
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const int c = c_lo + q * 16 + j;
                                const float4 b4 = *reinterpret_cast<const float4*>(pb + c);
                                float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (rb) r4 = __ldg(reinterpret_cast<const float4*>(rb + c));
                                v[q * 16 + j] = t[j] + b4.x + r4.x; v[q * 16 + j + 1] = t[j + 1] + b4.y + r4.y;
                                v[q * 16 + j + 2] = t[j + 2] + b4.z + r4.z; v[q * 16 + j + 3] = t[j + 3] + b4.w + r4.w;
                                s += (v[q * 16 + j] + v[q * 16 + j + 1]) + (v[q * 16 + j + 2] + v[q * 16 + j + 3]);
                            }
                        }
                    }
                    tc::fence_before();
                    float mean = 0.f, rstd = 1.f;
                    const bool ln = (a.flags[L] & 1) != 0, act = (a.flags[L] & 2) != 0;
                    if (ln) {      // two-pass statistics over the row, partial sums exchanged between the 4 threads of the row
                        s_red[wg * kTcM + m] = act_cols ? s : 0.f;
                        epi_sync();
                        mean = ((s_red[m] + s_red[kTcM + m]) + (s_red[2 * kTcM + m] + s_red[3 * kTcM + m])) / (float)N;
                        float ss = 0.f;
#pragma unroll
                        for (int c = 0; c < 32; ++c) if (c < 16 * n16 && act_cols) { const float t = v[c] - mean; ss += t * t; }
                        s_red[(kTcEpiWG + wg) * kTcM + m] = ss;
                        epi_sync();
                        const float* r2 = s_red + kTcEpiWG * kTcM;
                        rstd = rsqrtf(((r2[m] + r2[kTcM + m]) + (r2[2 * kTcM + m] + r2[3 * kTcM + m])) / (float)N + 1e-5f);
                    }
#pragma unroll
                    for (int cg = 0; cg < 32; cg += KG) {
                        if (cg < 16 * n16 && act_cols) {
                            float o[KG];
#pragma unroll
                            for (int c4 = 0; c4 < KG; c4 += 4) {
                                const int c = c_lo + cg + c4;
                                o[c4] = v[cg + c4]; o[c4 + 1] = v[cg + c4 + 1]; o[c4 + 2] = v[cg + c4 + 2]; o[c4 + 3] = v[cg + c4 + 3];
                                if (ln) {
                                    const float4 g4 = *reinterpret_cast<const float4*>(pg + c), b4 = *reinterpret_cast<const float4*>(pbb + c);
                                    o[c4] = (o[c4] - mean) * rstd * g4.x + b4.x; o[c4 + 1] = (o[c4 + 1] - mean) * rstd * g4.y + b4.y;
                                    o[c4 + 2] = (o[c4 + 2] - mean) * rstd * g4.z + b4.z; o[c4 + 3] = (o[c4 + 3] - mean) * rstd * g4.w + b4.w;
                                }
                                if (act) { o[c4] = siluf_(o[c4]); o[c4 + 1] = siluf_(o[c4 + 1]); o[c4 + 2] = siluf_(o[c4 + 2]); o[c4 + 3] = siluf_(o[c4 + 3]); }
                            }
                            const uint32_t off = (uint32_t)(((c_lo + cg) / KG) * kTcM + m) * 16u;
                            GR::store(sA_hi + off, sA_lo + off, o);
                        }
                    }
                    tc::fence_async_smem();
                    tc::mbar_arrive(&a_ready);
                } else {
                    // output rows: TMEM -> registers (+ bias + offset) -> the row's slot of the staging tile (the free A buffers;
                    // row stride CW + 4 floats: conflict-free float4 stores), 16-column chunks dealt round-robin to the 4
                    // threads of the row -> ONE TMA bulk store per row and staging pass, issued by warpgroup 0
                    float* s_out = reinterpret_cast<float*>(sA_hi);
                    const int CW = a.out_cw, ldo = CW + 4;
                    for (int cb = 0; cb < N; cb += CW) {
                        const int cw = min(CW, N - cb);
                        float* srow = s_out + (size_t)m * ldo;
                        for (int c0 = wg * 16; c0 < cw; c0 += 16 * kTcEpiWG) {
                            float t[16];
                            tc::tmem_ld16(t_lane + cb + c0, t);
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const float4 of = *reinterpret_cast<const float4*>(s_last + cb + c0 + j);
                                *reinterpret_cast<float4*>(srow + c0 + j) = make_float4(t[j] + of.x, t[j + 1] + of.y, t[j + 2] + of.z, t[j + 3] + of.w);
                            }
                        }
                        tc::fence_async_smem();
                        epi_sync();                                   // the whole row is staged
                        if (wg == 0) {
                            if (valid)
                                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                             ::"l"(a.out + (size_t)e * N + cb), "r"(smem_u32(srow)), "r"((uint32_t)cw * 4u) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        }
                        epi_sync();                                   // staging tile reusable (next pass / next tile's input)
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
    TC_STAMP(0); TC_STAMP(kTcProdWarp * 32); TC_STAMP(kTcMmaWarp * 32);
    tc::fence_before();
    __syncthreads();
    if (warp == kTcProdWarp) tc::tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    TC_STAMP(0);
#undef TC_STAMP
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

static long long* g_tc_dbg = nullptr;
/* debug hook (not part of the public header): 192 x int64 device buffer receiving clock64() stamps of CTA 0 */
extern "C" int dedf_tc_set_debug(long long* dbg) { g_tc_dbg = dbg; return DEDF_OK; }

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
        if (i < d->n_layers - 1 && !(N == 128 || N <= 64)) return DEDF_ERR_UNSUPPORTED;     // epilogue column split
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
    const bool f16 = d->tc_f16 != 0;
    if (f16) for (int i = 0; i < d->n_layers; ++i) if (a.K[i] % 16) return DEDF_ERR_UNSUPPORTED;      // K = 16 per kind::f16 MMA
    a.out_offset = d->out_offset; a.out = d->out;
    a.dbg = g_tc_dbg;
    a.a_bytes = kTcM * max_k * 4;
    a.stage_bytes = ((max_nb * 64) + 1023) / 1024 * 1024;
    a.tmem_cols = 32;
    while (a.tmem_cols < max_n) a.tmem_cols <<= 1;
    const int ns = (d->mode == DEDF_MLP_IN_FIELD) ? a.n_scales : 1;
    {   // widest staging chunk (multiple of 16 columns, <= one N block) whose 128 padded rows fit in the two A buffers
        const int n_last = a.K[a.n_layers];
        int cw = ((2 * a.a_bytes) / (kTcM * 4) - 4) / 16 * 16;
        if (cw > n_last) cw = n_last;
        if (cw < 16) return DEDF_ERR_UNSUPPORTED;
        a.out_cw = cw;
        if (n_last > 512 || (n_last % 4)) return DEDF_ERR_UNSUPPORTED;
    }
    const size_t smem = (size_t)2 * a.a_bytes + (size_t)kTcStages * a.stage_bytes +
                        ((size_t)ns * a.K[0] * 4 + ((a.K[0] / 2 + 3) & ~3) + (DEDF_MLP_MAX_LAYERS - 1) * 3 * kTcMaxHidden + 512 + 2 * kTcEpiWG * kTcM) * sizeof(float);
    if (smem > 220 * 1024) return DEDF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(edge_mlp_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        cudaFuncSetAttribute(edge_mlp_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        attr_done = true;
    }
    const int n_tiles = (max_edges + kTcM - 1) / kTcM + DEDF_MAX_SCALES;
    if (f16) launch_pdl(edge_mlp_tc_kernel<true>, dim3(grid_for(n_tiles, 1, kNumSMs)), dim3(kTcThreads), smem, stream, a);
    else launch_pdl(edge_mlp_tc_kernel<false>, dim3(grid_for(n_tiles, 1, kNumSMs)), dim3(kTcThreads), smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
