// Per-edge kernels of the equivariant graph attention
// (GraphAttentionMLP / GraphAttentionMLP2, /root/reference/diffusion_edf/graph_attention.py:84-122, :218-273):
//
//   edge_geom       vec / length / spherical harmonics / soft cut-offs          (graph_parser.py:146-224,
//                                                                                unet_feature_extractor.py:284-288)
//   edge_mlp        length (+time) embedding -> RadialProfile MLP -> per-edge TP weights
//                                                                               (equiformer/radial_func.py:56-59,
//                                                                                multiscale_tensor_field.py:225-234)
//   edge_tp_lin     gather -> depthwise CG tensor product -> block-diagonal linear -> (alpha logits + gate | bias)
//                   with the (E,1568) tensor-product output kept in shared memory  (graph_attention.py:231-246)
//   segment_softmax_reduce   per-destination softmax over incoming edges + weighted sum (graph_attention.py:254-265)
//   edge_tp_reduce  "K1": gather -> depthwise CG TP with per-edge weights -> x alpha_h -> segment reduce
#include "common.cuh"
#include "cg_slots.cuh"
#include "gemm_tile.cuh"
#include "../../include/dedf.h"

namespace dedf {

// ===========================================================================
// edge geometry
// ===========================================================================
struct GeomArgs {
    const float* x_src; const float* x_dst;
    const int* edge_src; const int* edge_dst;
    const int* n_edges;                 // device scalar
    float* length; float* sh; float* logit;   // logit may be null
    float ns_lo, ns_hi;                 // non-scalar SH min-cut range; ns_hi <= 0: none
    int n_scales;
    int src_off[DEDF_MAX_SCALES + 1];
    float r[DEDF_MAX_SCALES];           // < 0: infinite scale (logit 0)
};

__global__ void __launch_bounds__(256) edge_geom_kernel(GeomArgs a) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    const int E = *a.n_edges;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int s = a.edge_src[e], d = a.edge_dst[e];
        const float vx = a.x_src[3 * s] - a.x_dst[3 * d];
        const float vy = a.x_src[3 * s + 1] - a.x_dst[3 * d + 1];
        const float vz = a.x_src[3 * s + 2] - a.x_dst[3 * d + 2];
        int sc = 0;
        if (a.logit) while (sc + 1 < a.n_scales && s >= a.src_off[sc + 1]) ++sc;
        float len, sh[9], lg;
        edge_geometry_rn(vx, vy, vz, a.ns_lo, a.ns_hi, a.logit ? a.r[sc] : -1.f, &len, sh, &lg);
        a.length[e] = len;
#pragma unroll
        for (int j = 0; j < 9; ++j) a.sh[(size_t)e * 9 + j] = sh[j];
        if (a.logit) a.logit[e] = lg;
    }
}

// ===========================================================================
// edge MLP
// ===========================================================================
constexpr int kMlpTE = 64;          // edges per tile
constexpr int kMlpThreads = 256;
constexpr int kMlpMaxW = 192;       // widest layer input / hidden layer kept in shared memory (sapien configs: 64 + 128 edge scalars)

struct MlpArgs {
    int mode;                       // DEDF_MLP_IN_ROWS / _RBF / _FIELD
    const int* n_edges;             // device scalar
    // input
    const float* x_in;              // ROWS: (E, K[0])
    const float* length;            // RBF / FIELD
    // RBF (GaussianRadialBasisLayerFiniteCutoff, radial_func.py:231-278)
    const float* rbf_mean; const float* rbf_std_logit; const float* rbf_weight_logit;
    float rbf_cutoff, rbf_offset;
    // FIELD (graph_parser length encoders + edge_scalars_pre_linears, per scale)
    int n_scales; int n_dst;        // edges of (scale s, dst d) = [row_ptr[s*n_dst+d], row_ptr[s*n_dst+d+1])
    const int* row_ptr;
    const int* edge_dst;
    const float* enc_mean[DEDF_MAX_SCALES]; const float* enc_std_logit[DEDF_MAX_SCALES];
    const float* enc_weight_logit[DEDF_MAX_SCALES];
    float enc_r[DEDF_MAX_SCALES];   // < 0: sinusoidal encoder with max_val = enc_max_r, n = enc_n
    float enc_max_r, enc_n;
    const float* enc_freq;          // (K0/2) sinusoidal frequency table
    const float* pre_w[DEDF_MAX_SCALES];   // (len_dim, K1) = W_s[:, :len_dim]^T
    const float* row_bias;          // (n_scales, n_rb, K1): W_s[:, len_dim:] t_emb + b_s
    int n_rb; int rb_div;           // row of row_bias = edge_dst / rb_div  (clamped to n_rb - 1)
    // layers
    int n_layers;
    int K[DEDF_MLP_MAX_LAYERS + 1];
    const float* W[DEDF_MLP_MAX_LAYERS];   // (K[i], K[i+1])
    const float* b[DEDF_MLP_MAX_LAYERS];   // may be null
    const float* ln_g[DEDF_MLP_MAX_LAYERS]; const float* ln_b[DEDF_MLP_MAX_LAYERS];
    int flags[DEDF_MLP_MAX_LAYERS];        // 1: LayerNorm, 2: SiLU
    const float* out_offset;        // added to the last layer (RadialProfile.offset), may be null
    float* out;                     // (E, K[n_layers])
    int w_smem;                     // 1: all weight matrices are staged in shared memory by TMA bulk copies at kernel start
};

__device__ __forceinline__ float softplusf_(float x) { return (x > 20.f) ? x : log1pf(expf(x)); }

__global__ void __launch_bounds__(kMlpThreads) edge_mlp_kernel(MlpArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int lda = pad_lda(kMlpMaxW);
    float* buf0 = smem;
    float* buf1 = smem + kMlpTE * lda;
    const int tid = threadIdx.x;
    // ---- weights -> shared memory (TMA bulk copies, issued before anything else; waited for before the first GEMM) ----
    __shared__ __align__(8) uint64_t wbar;
    float* sW = buf1 + kMlpTE * lda;
    int w_off[DEDF_MLP_MAX_LAYERS + 1];
    bool w_pending = false;
    if (a.w_smem) {
        const int n_first = (a.mode == DEDF_MLP_IN_FIELD) ? a.n_scales : 1;      // FIELD: one first-layer matrix per scale
        w_off[0] = 0;
        w_off[1] = n_first * a.K[0] * a.K[1];
        for (int L = 1; L < a.n_layers; ++L) w_off[L + 1] = w_off[L] + a.K[L] * a.K[L + 1];
        if (tid == 0) {
            mbar_init(&wbar, 1);
            mbar_init_fence();
            mbar_expect_tx(&wbar, (uint32_t)w_off[a.n_layers] * 4u);
            for (int s = 0; s < n_first; ++s)
                bulk_g2s_chunked(sW + s * a.K[0] * a.K[1], (a.mode == DEDF_MLP_IN_FIELD) ? a.pre_w[s] : a.W[0], (uint32_t)(a.K[0] * a.K[1]) * 4u, &wbar);
            for (int L = 1; L < a.n_layers; ++L) bulk_g2s_chunked(sW + w_off[L], a.W[L], (uint32_t)(a.K[L] * a.K[L + 1]) * 4u, &wbar);
        }
        w_pending = true;
    }
    pdl_wait(); pdl_launch();     // PDL: the weight copies above overlap the previous kernel's tail
    const int E = *a.n_edges;

    // tiles: FIELD mode keeps tiles inside one scale (first layer weights differ per scale)
    int n_tiles = 0;
    int tile_base[DEDF_MAX_SCALES + 1];
    if (a.mode == DEDF_MLP_IN_FIELD) {
        tile_base[0] = 0;
        for (int s = 0; s < a.n_scales; ++s) {
            const int es = a.row_ptr[(size_t)(s + 1) * a.n_dst] - a.row_ptr[(size_t)s * a.n_dst];
            tile_base[s + 1] = tile_base[s] + (es + kMlpTE - 1) / kMlpTE;
        }
        n_tiles = tile_base[a.n_scales];
    } else {
        n_tiles = (E + kMlpTE - 1) / kMlpTE;
    }

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int e0, e1, scale = 0;
        if (a.mode == DEDF_MLP_IN_FIELD) {
            while (tile >= tile_base[scale + 1]) ++scale;
            const int sbeg = a.row_ptr[(size_t)scale * a.n_dst], send = a.row_ptr[(size_t)(scale + 1) * a.n_dst];
            e0 = sbeg + (tile - tile_base[scale]) * kMlpTE;
            e1 = min(e0 + kMlpTE, send);
        } else {
            e0 = tile * kMlpTE;
            e1 = min(e0 + kMlpTE, E);
        }
        const int rows = e1 - e0;
        const int K0 = a.K[0];
        __syncthreads();   // previous tile's readers are done with buf0/buf1
        // ---- stage the input tile into buf0 ---------------------------------
        if (a.mode == DEDF_MLP_IN_ROWS) {
            for (int i = tid; i < kMlpTE * K0; i += kMlpThreads) {
                const int r = i / K0, k = i % K0;
                buf0[r * lda + k] = (r < rows) ? a.x_in[(size_t)(e0 + r) * K0 + k] : 0.f;
            }
        } else if (a.mode == DEDF_MLP_IN_RBF) {
            const float inv_span = 1.0f / (a.rbf_cutoff - a.rbf_offset);
            const float normalizer = sqrtf((float)K0);
            for (int i = tid; i < kMlpTE * K0; i += kMlpThreads) {
                const int r = i / K0, k = i % K0;
                float v = 0.f;
                if (r < rows) {
                    const float d = (a.length[e0 + r] - a.rbf_offset) * inv_span;
                    const float std = softplusf_(a.rbf_std_logit[k]) + 1e-5f;
                    const float z = (d - a.rbf_mean[k]) / std;
                    v = expf(-0.5f * z * z) * (sigmoidf_(a.rbf_weight_logit[k]) * 4.0f);
                    // soft_square_cutoff(d, thr=0.8, infinite=False): only the inner (d -> 0) edge is cut
                    const float cut = (d > 0.5f) ? 1.0f : (1.0f - soft_step3(((1.0f - d) - 0.8f) / (1.0f - 0.8f)));
                    v = v * cut * normalizer;
                }
                buf0[r * lda + k] = v;
            }
        } else {   // FIELD: length embedding of this scale
            const float r_s = a.enc_r[scale];
            for (int i = tid; i < kMlpTE * K0; i += kMlpThreads) {
                const int r = i / K0, k = i % K0;
                float v = 0.f;
                if (r < rows) {
                    const float len = a.length[e0 + r];
                    if (r_s >= 0.f) {   // GaussianRadialBasis (radial_func.py:208-227)
                        const float d = len / r_s;
                        const float std = softplusf_(a.enc_std_logit[scale][k]) + 1e-5f;
                        const float z = (d - a.enc_mean[scale][k]) / std;
                        v = expf(-0.5f * z * z) * (sigmoidf_(a.enc_weight_logit[scale][k]) * (4.0f * sqrtf((float)K0)));
                    } else {            // SinusoidalPositionEmbeddings (radial_func.py:291-316)
                        const int half = K0 / 2;
                        const int kk = (k < half) ? k : k - half;
                        const float x = len / a.enc_max_r * a.enc_n;
                        const float arg = __fmul_rn(x, a.enc_freq[kk]);
                        v = (k < half) ? sinf(arg) : cosf(arg);
                    }
                }
                buf0[r * lda + k] = v;
            }
        }
        __syncthreads();
        if (w_pending) { mbar_wait(&wbar, 0); w_pending = false; }

        float* in = buf0;
        float* outb = buf1;
        for (int L = 0; L < a.n_layers; ++L) {
            const int K = a.K[L], N = a.K[L + 1];
            const bool last = (L == a.n_layers - 1);
            const float* W = (a.mode == DEDF_MLP_IN_FIELD && L == 0) ? a.pre_w[scale] : a.W[L];
            if (a.w_smem) W = sW + w_off[L] + ((a.mode == DEDF_MLP_IN_FIELD && L == 0) ? scale * K * N : 0);
            const int n_rg = kMlpTE / 4, n_cg = N / 4;
            for (int item = tid; item < n_rg * n_cg; item += kMlpThreads) {
                const int cg = item % n_cg, rg = item / n_cg;
                float acc[4][4] = {};
                if (a.w_smem) gemm_item_4x4<true, true>(in, lda, n_rg, rg, W, N, 4 * cg, K, acc);
                else gemm_item_4x4<true, false>(in, lda, n_rg, rg, W, N, 4 * cg, K, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = rg + i * n_rg;
                    if (r >= rows && last) continue;
                    float v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j;
                        float t = acc[i][j];
                        if (a.b[L]) t += a.b[L][c];
                        if (a.mode == DEDF_MLP_IN_FIELD && L == 0 && r < rows && a.row_bias) {
                            const int rb = min(a.edge_dst[e0 + r] / a.rb_div, a.n_rb - 1);
                            t += a.row_bias[((size_t)scale * a.n_rb + rb) * N + c];
                        }
                        if (!(a.flags[L] & 1) && (a.flags[L] & 2)) t = siluf_(t);
                        if (last && a.out_offset) t += a.out_offset[c];
                        v[j] = t;
                    }
                    if (last) {
                        *reinterpret_cast<float4*>(a.out + (size_t)(e0 + r) * N + 4 * cg) = make_float4(v[0], v[1], v[2], v[3]);
                    } else {
                        *reinterpret_cast<float4*>(outb + r * lda + 4 * cg) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
            __syncthreads();
            if (!last && (a.flags[L] & 1)) {   // LayerNorm (+SiLU) in place, 4 threads per row
                const int r = tid >> 2, q = tid & 3;
                float s = 0.f, ss = 0.f;
                for (int c = q; c < N; c += 4) { const float t = outb[r * lda + c]; s += t; }
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2);
                const float mean = s / (float)N;
                for (int c = q; c < N; c += 4) { const float t = outb[r * lda + c] - mean; ss += t * t; }
                ss += __shfl_xor_sync(0xffffffffu, ss, 1); ss += __shfl_xor_sync(0xffffffffu, ss, 2);
                const float rstd = rsqrtf(ss / (float)N + 1e-5f);
                for (int c = q; c < N; c += 4) {
                    float t = (outb[r * lda + c] - mean) * rstd * a.ln_g[L][c] + a.ln_b[L][c];
                    if (a.flags[L] & 2) t = siluf_(t);
                    outb[r * lda + c] = t;
                }
                __syncthreads();
            }
            float* t = in; in = outb; outb = t;
        }
    }
}

// ===========================================================================
// edge_tp_lin: gather -> depthwise TP -> block-diagonal linear -> epilogue
// ===========================================================================
struct TpLinArgs {
    const float* x_src;      // (N_src, F) messages of the source nodes
    const float* x_dst;      // (N_dst, F) or null: added to the gathered source message
    const int* edge_src; const int* edge_dst;
    const int* n_edges;      // device scalar
    const float* sh;         // (E, 9)
    const float* w; long long w_stride;   // per-edge TP weights (E, NUMEL), or shared (NUMEL,) with stride 0
    const float* W0; const float* W1; const float* W2;   // (D0,N0) (D1,N1) (D2,N2)
    const float* bias0;      // (N0) added to the 0e outputs (may be null)
    // EPI_ACT
    const float* alpha_dot;  // (H * Ma/H)
    const float* edge_logit; // (E) or null
    float* logits;           // (E, 4)
    float* out;              // (E, F_out)
    int per_edge_x;          // 1: x_src is (E, F) indexed by the edge itself (no gather)
    long long* dbg;          // optional (host debug): clock64() stamps of CTA 0
};

enum { EPI_ACT = 0, EPI_LIN = 1 };

#ifndef DEDF_TPLIN_TE32
#define DEDF_TPLIN_TE32 16
#endif
constexpr int kTpLinTE32 = DEDF_TPLIN_TE32;
#ifndef DEDF_TPLIN_PREFETCH
#define DEDF_TPLIN_PREFETCH 0
#endif
constexpr int kTpLinPrefetch = DEDF_TPLIN_PREFETCH;   // k-steps of weight rows prefetched into L1 ahead of the K loop

template <int G, int EPI>
struct TpLinCfg {
    using D = Dtp<G>;
    // edges per tile (16 / 32).  Measured alternative (profiles/run_tp_lin.py): G = 32 with 8-edge tiles and an L1 large
    // enough for the whole weight set halves the latency of a K step but doubles the steps per edge: 953 us vs 683 us.
    static constexpr int TE = (G == 32) ? kTpLinTE32 : 32;
    static constexpr int MA = (EPI == EPI_ACT) ? D::M0 : 0;  // alpha channels (= mul of 0e in the heads irreps)
    static constexpr int N0 = (EPI == EPI_ACT) ? (MA + D::M0 + D::M1 + D::M2) : D::M0;
    static constexpr int N1 = D::M1, N2 = D::M2;
    static constexpr int FO = D::M0 + 3 * D::M1 + 5 * D::M2;  // output feature dim (= F)
    static constexpr int NITEMS = (TE / 4) * (N0 / 4) + (3 * TE / 4) * (N1 / 4) + (5 * TE / 4) * (N2 / 4);
    static constexpr int THREADS = ((NITEMS + 31) / 32) * 32;
};

template <int G, int EPI>
__global__ void __launch_bounds__(TpLinCfg<G, EPI>::THREADS, 2)
edge_tp_lin_kernel(TpLinArgs a, int lda0, int lda1, int lda2) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    using C = TpLinCfg<G, EPI>;
    using D = Dtp<G>;
    constexpr int TE = C::TE, P = D::P;
    extern __shared__ __align__(16) float smem[];
    float* A0 = smem;                            // [TE][lda0]
    float* A1 = A0 + TE * lda0;                  // [3 TE][lda1]   row = k * TE + e
    float* A2 = A1 + 3 * TE * lda1;              // [5 TE][lda2]
    float* s_sh = A2 + 5 * TE * lda2;            // [TE][9]
    int* s_src = reinterpret_cast<int*>(s_sh + TE * 9);
    int* s_dst = s_src + TE;
    float* O0 = smem;                            // outputs alias the A region after the GEMM
    float* O1 = O0 + TE * C::N0;
    float* O2 = O1 + 3 * TE * C::N1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NWARPS = C::THREADS / 32;
    const int E = *a.n_edges;
    const int n_tiles = (E + TE - 1) / TE;

    int dbg_i = 0;
#define TPL_STAMP() do { if (a.dbg && blockIdx.x == 0 && tid == 0 && dbg_i < 60) a.dbg[dbg_i++] = clock64(); } while (0)
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int e0 = tile * TE;
        const int rows = min(TE, E - e0);
        __syncthreads();
        TPL_STAMP();
        for (int i = tid; i < TE; i += C::THREADS) {
            const bool ok = i < rows;
            s_src[i] = ok ? (a.per_edge_x ? (e0 + i) : a.edge_src[e0 + i]) : 0;
            s_dst[i] = ok ? a.edge_dst[e0 + i] : 0;
        }
        for (int i = tid; i < TE * 9; i += C::THREADS) s_sh[i] = (i / 9 < rows) ? a.sh[(size_t)e0 * 9 + i] : 0.f;
        __syncthreads();
        TPL_STAMP();

        // ---------------- CG phase: one pack of P edges per warp iteration ----------------
        for (int pack = warp; pack < TE / P; pack += NWARPS) {
            const int pe0 = pack * P;
            // l = 0 slots
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                int ei, ch; D::slot0(lane, s, ei, ch);
                const int e = pe0 + ei;
                const bool ok = e < rows;
                float o[9];
                float x = 0.f, w0 = 0.f, w1 = 0.f, w2 = 0.f;
                if (ok) {
                    x = a.x_src[(size_t)s_src[e] * D::F + ch];
                    if (a.x_dst) x += a.x_dst[(size_t)s_dst[e] * D::F + ch];
                    const float* w = a.w + (size_t)(e0 + e) * a.w_stride;
                    w0 = w[D::W_K0 + ch]; w1 = w[D::W_K1 + ch]; w2 = w[D::W_K2 + ch];
                }
                dtp_l0(x, w0, w1, w2, s_sh + e * 9, o);
                A0[e * lda0 + D::C0_K0 + ch] = o[0];
#pragma unroll
                for (int k = 0; k < 3; ++k) A1[(k * TE + e) * lda1 + D::C1_K1 + ch] = o[1 + k];
#pragma unroll
                for (int k = 0; k < 5; ++k) A2[(k * TE + e) * lda2 + D::C2_K2 + ch] = o[4 + k];
            }
            // l = 1 slots
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                int ei, ch; D::slot1(lane, s, ei, ch);
                const int e = pe0 + ei;
                const bool ok = e < rows;
                float x[3] = {0.f, 0.f, 0.f}, w[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[20];
                if (ok) {
                    const float* xs = a.x_src + (size_t)s_src[e] * D::F + D::M0 + 3 * ch;
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[i] = xs[i];
                    if (a.x_dst) {
                        const float* xd = a.x_dst + (size_t)s_dst[e] * D::F + D::M0 + 3 * ch;
#pragma unroll
                        for (int i = 0; i < 3; ++i) x[i] += xd[i];
                    }
                    const float* wp = a.w + (size_t)(e0 + e) * a.w_stride + D::W_K3 + ch;
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M1];
                }
                dtp_l1(x, w, s_sh + e * 9, o);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float* row = A1 + (k * TE + e) * lda1 + ch;
                    row[D::C1_K3] = o[k]; row[D::C1_K5] = o[4 + k]; row[D::C1_K7] = o[12 + k];
                }
                A0[e * lda0 + D::C0_K4 + ch] = o[3];
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    float* row = A2 + (k * TE + e) * lda2 + ch;
                    row[D::C2_K6] = o[7 + k]; row[D::C2_K8] = o[15 + k];
                }
            }
            // l = 2 slot
            {
                int ei, ch; D::slot2(lane, ei, ch);
                const int e = pe0 + ei;
                const bool ok = e < rows;
                float x[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, w[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, o[22];
                if (ok) {
                    const float* xs = a.x_src + (size_t)s_src[e] * D::F + D::M0 + 3 * D::M1 + 5 * ch;
#pragma unroll
                    for (int i = 0; i < 5; ++i) x[i] = xs[i];
                    if (a.x_dst) {
                        const float* xd = a.x_dst + (size_t)s_dst[e] * D::F + D::M0 + 3 * D::M1 + 5 * ch;
#pragma unroll
                        for (int i = 0; i < 5; ++i) x[i] += xd[i];
                    }
                    const float* wp = a.w + (size_t)(e0 + e) * a.w_stride + D::W_K9 + ch;
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M2];
                }
                dtp_l2(x, w, s_sh + e * 9, o);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    float* row = A2 + (k * TE + e) * lda2 + ch;
                    row[D::C2_K9] = o[k]; row[D::C2_K11] = o[8 + k]; row[D::C2_K14] = o[17 + k];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float* row = A1 + (k * TE + e) * lda1 + ch;
                    row[D::C1_K10] = o[5 + k]; row[D::C1_K13] = o[14 + k];
                }
                A0[e * lda0 + D::C0_K12 + ch] = o[13];
            }
        }
        __syncthreads();
        TPL_STAMP();

        // ---------------- GEMM phase: one 4x4 item per thread ----------------
        constexpr int I0 = (TE / 4) * (C::N0 / 4), I1 = (3 * TE / 4) * (C::N1 / 4), I2 = (5 * TE / 4) * (C::N2 / 4);
        float acc[4][4] = {};
        int which = -1, rg = 0, cg = 0;
        if (tid < I0) { which = 0; cg = tid % (C::N0 / 4); rg = tid / (C::N0 / 4); }
        else if (tid < I0 + I1) { which = 1; const int t = tid - I0; cg = t % (C::N1 / 4); rg = t / (C::N1 / 4); }
        else if (tid < I0 + I1 + I2) { which = 2; const int t = tid - I0 - I1; cg = t % (C::N2 / 4); rg = t / (C::N2 / 4); }
        if (which == 0) gemm_item_4x4<true, false, kTpLinPrefetch>(A0, lda0, TE / 4, rg, a.W0, C::N0, 4 * cg, D::D0, acc);
        else if (which == 1) gemm_item_4x4<true, false, kTpLinPrefetch>(A1, lda1, 3 * TE / 4, rg, a.W1, C::N1, 4 * cg, D::D1, acc);
        else if (which == 2) gemm_item_4x4<true, false, kTpLinPrefetch>(A2, lda2, 5 * TE / 4, rg, a.W2, C::N2, 4 * cg, D::D2, acc);
        TPL_STAMP();
        __syncthreads();   // all A reads done -> the region may be overwritten with the outputs
        if (which == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(O0 + (rg + i * (TE / 4)) * C::N0 + 4 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else if (which == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(O1 + (rg + i * (3 * TE / 4)) * C::N1 + 4 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else if (which == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                *reinterpret_cast<float4*>(O2 + (rg + i * (5 * TE / 4)) * C::N2 + 4 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        TPL_STAMP();

        // ---------------- epilogue: one warp per edge ----------------
        for (int e = warp; e < rows; e += NWARPS) {
            const size_t eg = (size_t)(e0 + e);
            if constexpr (EPI == EPI_ACT) {
                // attention logits: sum_k c * SLReLU(pre[h, k]) * alpha_dot[h, k] (+ edge logit)   (graph_attention.py:241-246)
                constexpr int MA = C::MA, HD = MA / 4;          // channels per head
                float part[MA / 32];
#pragma unroll
                for (int j = 0; j < MA / 32; ++j) {
                    const int c = lane + 32 * j;
                    const float pre = O0[e * C::N0 + c] + (a.bias0 ? a.bias0[c] : 0.f);
                    part[j] = kCSlrelu * slreluf_(pre) * a.alpha_dot[c];
                }
#pragma unroll
                for (int j = 0; j < MA / 32; ++j) {
#pragma unroll
                    for (int o = HD / 2; o > 0; o >>= 1) part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
                }
                const float el = a.edge_logit ? a.edge_logit[eg] : 0.f;
#pragma unroll
                for (int j = 0; j < MA / 32; ++j) {
                    if ((lane % HD) == 0) a.logits[eg * 4 + (lane + 32 * j) / HD] = part[j] + el;
                }
                // gate (fast_activation.py:210-224): scalars | gates | gated
                const float* b = a.bias0 ? a.bias0 + MA : nullptr;
                for (int c = lane; c < C::FO; c += 32) {
                    float v;
                    if (c < D::M0) {
                        v = kCSilu * siluf_(O0[e * C::N0 + MA + c] + (b ? b[c] : 0.f));
                    } else if (c < D::M0 + 3 * D::M1) {
                        const int u = (c - D::M0) / 3, k = (c - D::M0) % 3;
                        const float g = kCSigmoid * sigmoidf_(O0[e * C::N0 + MA + D::M0 + u] + (b ? b[D::M0 + u] : 0.f));
                        v = O1[(k * TE + e) * C::N1 + u] * g;
                    } else {
                        const int u = (c - D::M0 - 3 * D::M1) / 5, k = (c - D::M0 - 3 * D::M1) % 5;
                        const float g = kCSigmoid * sigmoidf_(O0[e * C::N0 + MA + D::M0 + D::M1 + u] + (b ? b[D::M0 + D::M1 + u] : 0.f));
                        v = O2[(k * TE + e) * C::N2 + u] * g;
                    }
                    a.out[eg * C::FO + c] = v;
                }
            } else {
                for (int c = lane; c < C::FO; c += 32) {
                    float v;
                    if (c < C::N0) v = O0[e * C::N0 + c] + (a.bias0 ? a.bias0[c] : 0.f);
                    else if (c < C::N0 + 3 * C::N1) { const int u = (c - C::N0) / 3, k = (c - C::N0) % 3; v = O1[(k * TE + e) * C::N1 + u]; }
                    else { const int u = (c - C::N0 - 3 * C::N1) / 5, k = (c - C::N0 - 3 * C::N1) % 5; v = O2[(k * TE + e) * C::N2 + u]; }
                    a.out[eg * C::FO + c] = v;
                }
            }
        }
    }
}

// ===========================================================================
// per-destination softmax over incoming edges + weighted sum of the values
// ===========================================================================
struct SoftmaxArgs {
    const int* row_ptr; int n_dst; int n_seg;   // edges of dst d in segment s: [row_ptr[s*n_dst+d], row_ptr[s*n_dst+d+1])
    const float* logits;   // (E, 4)
    const float* val;      // (E, F)
    float* out;            // (n_dst, F)
    int m0, m1, m2;        // value irreps; heads split every mul in 4
};

__global__ void __launch_bounds__(128) segment_softmax_reduce_kernel(SoftmaxArgs a) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    // One CTA (4 warps) per destination: the statistics passes stride the destination's edges with all 128 threads, the
    // weighted sum gives every warp a quarter of the edges (interleaved) and folds the four partial rows through shared
    // memory in a fixed order -- deterministic, no atomics.  (A warp per destination left the small graphs of the coarse
    // scales and of the score head -- a few hundred destinations -- on a fraction of the SMs with one long dependent chain.)
    __shared__ float s_red[4][4];
    __shared__ float s_acc[4][256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int F = a.m0 + 3 * a.m1 + 5 * a.m2;
    constexpr int MAXC = 8;   // channels per lane (F <= 256)
    // head of each of this lane's channels
    int hd[MAXC];
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
        const int c = lane + 32 * j;
        int h = 0;
        if (c < a.m0) h = c / (a.m0 / 4);
        else if (c < a.m0 + 3 * a.m1) h = ((c - a.m0) / 3) / (a.m1 / 4);
        else if (c < F) h = ((c - a.m0 - 3 * a.m1) / 5) / (a.m2 / 4);
        hd[j] = h;
    }
    for (int d = blockIdx.x; d < a.n_dst; d += gridDim.x) {
        // pass 1: per-head max over all edges of the destination
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int deg = 0;
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            deg += e - b;
            for (int i = b + tid; i < e; i += 128) {
                const float4 l = *reinterpret_cast<const float4*>(a.logits + (size_t)i * 4);
                mx[0] = fmaxf(mx[0], l.x); mx[1] = fmaxf(mx[1], l.y); mx[2] = fmaxf(mx[2], l.z); mx[3] = fmaxf(mx[3], l.w);
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) mx[h] = warp_max(mx[h]);
        __syncthreads();                       // previous destination's readers of s_red / s_acc are done
        if (lane == 0) { s_red[warp][0] = mx[0]; s_red[warp][1] = mx[1]; s_red[warp][2] = mx[2]; s_red[warp][3] = mx[3]; }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < 4; ++h) mx[h] = fmaxf(fmaxf(s_red[0][h], s_red[1][h]), fmaxf(s_red[2][h], s_red[3][h]));
        // pass 2: sum of exp
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            for (int i = b + tid; i < e; i += 128) {
                const float4 l = *reinterpret_cast<const float4*>(a.logits + (size_t)i * 4);
                sm[0] += __expf(l.x - mx[0]); sm[1] += __expf(l.y - mx[1]); sm[2] += __expf(l.z - mx[2]); sm[3] += __expf(l.w - mx[3]);
            }
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) sm[h] = warp_sum(sm[h]);
        __syncthreads();
        if (lane == 0) { s_red[warp][0] = sm[0]; s_red[warp][1] = sm[1]; s_red[warp][2] = sm[2]; s_red[warp][3] = sm[3]; }
        __syncthreads();
        float logZ[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const float t = (s_red[0][h] + s_red[1][h]) + (s_red[2][h] + s_red[3][h]);
            logZ[h] = (deg > 0) ? (logf(t + 1e-12f) + mx[h]) : 0.f;
        }
        // pass 3: weighted sum; warp w takes the edges w, w+4, ... of every segment, two edges in flight
        float acc[MAXC];
#pragma unroll
        for (int j = 0; j < MAXC; ++j) acc[j] = 0.f;
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            for (int i0 = b + warp; i0 < e; i0 += 8) {
                float4 l[2];
                float v[2][MAXC];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int i = min(i0 + 4 * u, e - 1);
                    l[u] = *reinterpret_cast<const float4*>(a.logits + (size_t)i * 4);
                    const float* vp = a.val + (size_t)i * F;
#pragma unroll
                    for (int j = 0; j < MAXC; ++j) { const int c = lane + 32 * j; v[u][j] = (c < F) ? vp[c] : 0.f; }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const bool on = (i0 + 4 * u) < e;
                    const float al0 = on ? __expf(l[u].x - logZ[0]) : 0.f, al1 = on ? __expf(l[u].y - logZ[1]) : 0.f;
                    const float al2 = on ? __expf(l[u].z - logZ[2]) : 0.f, al3 = on ? __expf(l[u].w - logZ[3]) : 0.f;
#pragma unroll
                    for (int j = 0; j < MAXC; ++j) {
                        const float al = (hd[j] == 0) ? al0 : (hd[j] == 1) ? al1 : (hd[j] == 2) ? al2 : al3;
                        acc[j] = fmaf(al, v[u][j], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MAXC; ++j) s_acc[warp][lane + 32 * j] = acc[j];
        __syncthreads();
        for (int c = tid; c < F; c += 128) a.out[(size_t)d * F + c] = (s_acc[0][c] + s_acc[1][c]) + (s_acc[2][c] + s_acc[3][c]);
    }
}

// ===========================================================================
// value path, reassociated: out[d] = lin( sum_e alpha_{e,h} dtp(v_e, sh_e, w_shared) ) + bias * sum_e alpha_{e,h}
// ===========================================================================
// graph_attention.py:237-239 + :254-266 computes value_e = lin(dtp(v_e, sh_e)) per EDGE (79 kMAC each) and then the
// alpha-weighted sum per destination.  sep_value.lin is linear with nothing after it and the heads partition its OUTPUT
// channels, so per head h   sum_e alpha_{e,h} lin_h(d_e) = lin_h(sum_e alpha_{e,h} d_e) + b_h sum_e alpha_{e,h}:
// the (E, 49 G) tensor-product outputs are reduced over the incoming edges first (4 head copies, in registers) and the
// linear layer runs once per destination instead of once per edge -- the K1 pattern (gather -> depthwise CG -> x alpha ->
// segment reduce) inside the real attention block, with the softmax statistics and the linear epilogue fused in.
//
// One CTA per destination.  Warp (cw, hp): channel group cw (16 l=0, 8 l=1, 4 l=2 channels) x head pair hp (heads 2hp, 2hp+1).
// Edges are staged 64 at a time in shared memory (value rows and logits by TMA bulk copies); a warp walks them in packs of 8
// so that all 32 lanes run the same code: 4 x (2 edges x 16 l=0 channels), 2 x (4 edges x 8 l=1 channels), 1 x (8 edges x 4 l=2
// channels); lanes that share a channel are folded with shuffles at the end.  Deterministic, no atomics.
struct ValueReduceArgs {
    const int* row_ptr; int n_dst; int n_seg;
    const float* v;          // (E, F) gated values (ACT epilogue output)
    const float* sh;         // (E, 9)
    const float* logits;     // (E, 4)
    const float* post;       // (E) optional factor applied to alpha after the softmax (source-point attention) or null
    const float* wv;         // (NUMEL) shared depthwise-TP weights (sep_value.dtp.tp.weight)
    const float* V0; const float* V1; const float* V2;   // sep_value.lin blocks (D_l, M_l) row-major
    const float* vb;         // (M0) bias or null
    float* out;              // (n_dst, F)
};

constexpr int kVrChunk = 64;

// -DDEDF_VR_TRACE (profiles/run_vr_trace.py builds its own copy of the library): thread 0 of CTA 0 stamps clock64() at the phase
// boundaries of its FIRST destination into a device array that dedf_vr_trace() copies out.
#ifdef DEDF_VR_TRACE
__device__ long long g_vr_trace[24];
#define VR_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0 && d == (int)blockIdx.x) g_vr_trace[i] = clock64(); } while (0)
#define VR_STAMP0(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_vr_trace[i] = clock64(); } while (0)
#else
#define VR_STAMP(i) do { } while (0)
#define VR_STAMP0(i) do { } while (0)
#endif

template <int G>
__global__ void __launch_bounds__(G * 8, G == 16 ? 3 : 1) value_reduce_kernel(ValueReduceArgs a) {
    using D = Dtp<G>;
    constexpr int NCW = G / 8, NW = 2 * NCW, NT = NW * 32, CH = kVrChunk;
    constexpr int B1 = D::D0, B2 = D::D0 + 3 * D::D1;             // block offsets inside one head copy of the reduced TP output
    constexpr int NV0 = D::D0 * D::M0, NV1 = D::D1 * D::M1, NV2 = D::D2 * D::M2;
    extern __shared__ __align__(16) float smem[];
    float* s_v = smem;                          // [CH][F]
    float* s_lg = s_v + CH * D::F;              // [CH][4] logits -> alpha (in place)
    float* s_sh = s_lg + CH * 4;                // [CH][12]
    float* s_D = s_sh + CH * 12;                // [4 heads][FOUT]
    float* s_V = s_D + 4 * D::FOUT;             // sep_value.lin weights [V0 | V1 | V2], staged once per CTA
    __shared__ float s_red[NW][4], s_sal[4];
    __shared__ int s_cum[DEDF_MAX_SCALES + 1], s_beg[DEDF_MAX_SCALES];
    __shared__ __align__(8) uint64_t bar, vbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    VR_STAMP0(0);
    const int cw = warp % NCW, hp = warp / NCW;
    const int ch0 = cw * 16 + (lane & 15), ch1 = cw * 8 + (lane & 7), ch2 = cw * 4 + (lane & 3);
    float w0[3], w1[6], w2[6];
    w0[0] = a.wv[D::W_K0 + ch0]; w0[1] = a.wv[D::W_K1 + ch0]; w0[2] = a.wv[D::W_K2 + ch0];
#pragma unroll
    for (int i = 0; i < 6; ++i) { w1[i] = a.wv[D::W_K3 + ch1 + i * D::M1]; w2[i] = a.wv[D::W_K9 + ch2 + i * D::M2]; }
    // The tensor-product weights of the value path are SHARED (sep_value.dtp.tp.weight, one per path and channel, the same for every
    // edge), so  sum_e alpha_e w_p cg_p(x_e, sh_e) = w_p sum_e alpha_e cg_p(x_e, sh_e):  the edge loop runs with unit weights and the
    // path weights are applied ONCE per destination when the reduced outputs are written (one multiply per output instead of one per
    // output and edge: ~15 % of the loop's FP instructions, and 15 registers that are no longer live across it).
    const float kOnes[6] = {1.0f, 1.0f, 1.0f, 1.0f, 1.0f, 1.0f};
    if (tid == 0) {
        mbar_init(&bar, 1); mbar_init(&vbar, 1);
        mbar_init_fence();
        // the linear layer's weights do not depend on the previous kernel: their copy overlaps its tail (PDL)
        mbar_expect_tx(&vbar, (uint32_t)(NV0 + NV1 + NV2) * 4u);
        bulk_g2s_chunked(s_V, a.V0, NV0 * 4u, &vbar);
        bulk_g2s_chunked(s_V + NV0, a.V1, NV1 * 4u, &vbar);
        bulk_g2s_chunked(s_V + NV0 + NV1, a.V2, NV2 * 4u, &vbar);
    }
    VR_STAMP0(1);
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    VR_STAMP0(2);
    __syncthreads();
    uint32_t ph = 0;
    bool v_pending = true;

    // stage edges [f0, f0 + n) of destination d's FLAT edge list (all segments back to back) into the chunk buffers
    auto issue_chunk = [&](int f0, int n) {       // thread 0 only; s_cum / s_beg are valid
        mbar_expect_tx(&bar, (uint32_t)n * (D::F + 4) * 4u);
        for (int s = 0; s < a.n_seg; ++s) {
            const int lo = max(f0, s_cum[s]), hi = min(f0 + n, s_cum[s + 1]);
            if (lo < hi) {
                const size_t e = (size_t)s_beg[s] + (lo - s_cum[s]);
                bulk_g2s_chunked(s_v + (size_t)(lo - f0) * D::F, a.v + e * D::F, (uint32_t)(hi - lo) * D::F * 4u, &bar);
                bulk_g2s(s_lg + (lo - f0) * 4, a.logits + e * 4, (uint32_t)(hi - lo) * 16u, &bar);
            }
        }
    };

    for (int d = blockIdx.x; d < a.n_dst; d += gridDim.x) {
        __syncthreads();                                           // previous destination done with every shared buffer
        if (warp == 0) {
            // the segments' row pointers in parallel (lane = segment; one L2 round trip instead of a serial loop of 2 n_seg loads on
            // one thread: 2.3 k of the ~20 k cycles a destination costs after the PDL wait, profiles/run_vr_trace.py), then a shuffle scan
            int b = 0, e = 0;
            if (lane < a.n_seg) { b = a.row_ptr[(size_t)lane * a.n_dst + d]; e = a.row_ptr[(size_t)lane * a.n_dst + d + 1]; }
            const int len = e - b;
            int cum = len;
#pragma unroll
            for (int o = 1; o < DEDF_MAX_SCALES; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += t; }
            if (lane < a.n_seg) { s_cum[lane] = cum - len; s_beg[lane] = b; }
            if (lane == a.n_seg - 1) s_cum[a.n_seg] = cum;
            __syncwarp();
            if (lane == 0) {
                const int c = s_cum[a.n_seg];
                if (c > 0) issue_chunk(0, min(CH, c));             // the first chunk flies while the statistics are computed
            }
        }
        if (tid < 4) s_sal[tid] = 0.f;
        __syncthreads();
        VR_STAMP(3);
        const int deg = s_cum[a.n_seg];
        // ---- softmax statistics over all incoming edges: per-head max, log Z (logits straight from global / L2) ----
        // thread t owns the flat edges t, t + NT, ... of the destination (all segments back to back); the first one -- the only one
        // unless the destination has more than NT edges -- stays in registers for the second pass
        auto flat_edge = [&](int f) { int s = 0; while (f >= s_cum[s + 1]) ++s; return (size_t)s_beg[s] + (f - s_cum[s]); };
        // the harmonics of the FIRST chunk are fetched here, together with the logits: their L2 round trip overlaps the statistics
        // instead of following them (9-float rows are not 16-byte aligned: plain loads)
        constexpr int SHN = (CH * 9 + NT - 1) / NT;
        const int n_first = min(CH, deg);
        float shv[SHN];
#pragma unroll
        for (int k = 0; k < SHN; ++k) {
            const int i = tid + k * NT;
            shv[k] = (i < n_first * 9) ? a.sh[flat_edge(i / 9) * 9 + (i % 9)] : 0.f;
        }
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        float4 l_first = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid < deg) l_first = *reinterpret_cast<const float4*>(a.logits + flat_edge(tid) * 4);
#pragma unroll
        for (int k = 0; k < SHN; ++k) {
            const int i = tid + k * NT;
            if (i < n_first * 9) s_sh[(i / 9) * 12 + (i % 9)] = shv[k];
        }
        if (tid < deg) { mx[0] = l_first.x; mx[1] = l_first.y; mx[2] = l_first.z; mx[3] = l_first.w; }
        for (int f = tid + NT; f < deg; f += NT) {
            const float4 l = *reinterpret_cast<const float4*>(a.logits + flat_edge(f) * 4);
            mx[0] = fmaxf(mx[0], l.x); mx[1] = fmaxf(mx[1], l.y); mx[2] = fmaxf(mx[2], l.z); mx[3] = fmaxf(mx[3], l.w);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) mx[h] = warp_max(mx[h]);
        if (lane == 0) { s_red[warp][0] = mx[0]; s_red[warp][1] = mx[1]; s_red[warp][2] = mx[2]; s_red[warp][3] = mx[3]; }
        __syncthreads();
        VR_STAMP(4);
#pragma unroll
        for (int h = 0; h < 4; ++h) { float m = s_red[0][h]; for (int w = 1; w < NW; ++w) m = fmaxf(m, s_red[w][h]); mx[h] = m; }
        float sm[4] = {0.f, 0.f, 0.f, 0.f};
        if (tid < deg) {
            sm[0] = __expf(l_first.x - mx[0]); sm[1] = __expf(l_first.y - mx[1]); sm[2] = __expf(l_first.z - mx[2]); sm[3] = __expf(l_first.w - mx[3]);
        }
        for (int f = tid + NT; f < deg; f += NT) {
            const float4 l = *reinterpret_cast<const float4*>(a.logits + flat_edge(f) * 4);
            sm[0] += __expf(l.x - mx[0]); sm[1] += __expf(l.y - mx[1]); sm[2] += __expf(l.z - mx[2]); sm[3] += __expf(l.w - mx[3]);
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) sm[h] = warp_sum(sm[h]);
        __syncthreads();
        if (lane == 0) { s_red[warp][0] = sm[0]; s_red[warp][1] = sm[1]; s_red[warp][2] = sm[2]; s_red[warp][3] = sm[3]; }
        __syncthreads();
        float logZ[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float t = 0.f;
            for (int w = 0; w < NW; ++w) t += s_red[w][h];
            logZ[h] = (deg > 0) ? (logf(t + 1e-12f) + mx[h]) : 0.f;
        }
        VR_STAMP(5);

        // ---- edge loop: accumulate alpha_h * dtp(v_e, sh_e, w) for this warp's channels and its two heads ----
        float acc0[2][9], acc1[2][20], acc2[2][22];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc0[h][k] = 0.f;
#pragma unroll
            for (int k = 0; k < 20; ++k) acc1[h][k] = 0.f;
#pragma unroll
            for (int k = 0; k < 22; ++k) acc2[h][k] = 0.f;
        }
        for (int f0 = 0; f0 < deg; f0 += CH) {
            const int n = min(CH, deg - f0);
            if (f0 > 0) {
                __syncthreads();                                   // previous chunk fully consumed
                if (tid == 0) issue_chunk(f0, n);
            }
            // harmonics of the later chunks, per segment piece (the first chunk's were staged with the statistics)
            for (int i = tid; f0 > 0 && i < n * 9; i += NT) {
                const int r = i / 9, f = f0 + r;
                int s = 0;
                while (f >= s_cum[s + 1]) ++s;
                s_sh[r * 12 + (i % 9)] = a.sh[((size_t)s_beg[s] + (f - s_cum[s])) * 9 + (i % 9)];
            }
            VR_STAMP(6);
            mbar_wait(&bar, ph); ph ^= 1u;
            __syncthreads();
            VR_STAMP(7);
            // logits -> alpha (x optional post factor), in place; per-head sums for the bias term (fixed order)
            if (tid < CH) {
                float4 al = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tid < n) {
                    const float4 l = *reinterpret_cast<const float4*>(s_lg + tid * 4);
                    float pf = 1.0f;
                    if (a.post) {
                        const int f = f0 + tid;
                        int s = 0;
                        while (f >= s_cum[s + 1]) ++s;
                        pf = a.post[(size_t)s_beg[s] + (f - s_cum[s])];
                    }
                    al = make_float4(__expf(l.x - logZ[0]) * pf, __expf(l.y - logZ[1]) * pf, __expf(l.z - logZ[2]) * pf, __expf(l.w - logZ[3]) * pf);
                }
                *reinterpret_cast<float4*>(s_lg + tid * 4) = al;
                float t0 = warp_sum(al.x), t1 = warp_sum(al.y), t2 = warp_sum(al.z), t3 = warp_sum(al.w);
                if (lane == 0) { s_red[warp][0] = t0; s_red[warp][1] = t1; s_red[warp][2] = t2; s_red[warp][3] = t3; }
            }
            __syncthreads();
            if (tid < 4) s_sal[tid] += s_red[0][tid] + s_red[1][tid];      // CH = 64 = the first two warps
            VR_STAMP(8);
            // packs of 8 edges
            for (int p0 = 0; p0 < n; p0 += 8) {
#pragma unroll
                for (int it = 0; it < 4; ++it) {                           // l = 0: 2 edges x 16 channels
                    const int e = p0 + 2 * it + (lane >> 4);
                    const bool ok = e < n;
                    const float x = ok ? s_v[e * D::F + ch0] : 0.f;
                    const float al0 = ok ? s_lg[e * 4 + 2 * hp] : 0.f, al1 = ok ? s_lg[e * 4 + 2 * hp + 1] : 0.f;
                    float o[9];
                    dtp_l0(x, 1.0f, 1.0f, 1.0f, s_sh + (ok ? e : 0) * 12, o);
#pragma unroll
                    for (int k = 0; k < 9; ++k) { acc0[0][k] = fmaf(al0, o[k], acc0[0][k]); acc0[1][k] = fmaf(al1, o[k], acc0[1][k]); }
                }
#pragma unroll
                for (int it = 0; it < 2; ++it) {                           // l = 1: 4 edges x 8 channels
                    const int e = p0 + 4 * it + (lane >> 3);
                    const bool ok = e < n;
                    const float* xs = s_v + (ok ? e : 0) * D::F + D::M0 + 3 * ch1;
                    const float xv[3] = {ok ? xs[0] : 0.f, ok ? xs[1] : 0.f, ok ? xs[2] : 0.f};
                    const float al0 = ok ? s_lg[e * 4 + 2 * hp] : 0.f, al1 = ok ? s_lg[e * 4 + 2 * hp + 1] : 0.f;
                    float o[20];
                    dtp_l1(xv, kOnes, s_sh + (ok ? e : 0) * 12, o);
#pragma unroll
                    for (int k = 0; k < 20; ++k) { acc1[0][k] = fmaf(al0, o[k], acc1[0][k]); acc1[1][k] = fmaf(al1, o[k], acc1[1][k]); }
                }
                {                                                          // l = 2: 8 edges x 4 channels
                    const int e = p0 + (lane >> 2);
                    const bool ok = e < n;
                    const float* xs = s_v + (ok ? e : 0) * D::F + D::M0 + 3 * D::M1 + 5 * ch2;
                    float xv[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) xv[i] = ok ? xs[i] : 0.f;
                    const float al0 = ok ? s_lg[e * 4 + 2 * hp] : 0.f, al1 = ok ? s_lg[e * 4 + 2 * hp + 1] : 0.f;
                    float o[22];
                    dtp_l2(xv, kOnes, s_sh + (ok ? e : 0) * 12, o);
#pragma unroll
                    for (int k = 0; k < 22; ++k) { acc2[0][k] = fmaf(al0, o[k], acc2[0][k]); acc2[1][k] = fmaf(al1, o[k], acc2[1][k]); }
                }
            }
        }
        VR_STAMP(9);
        // ---- fold the lanes that share a channel, write the 4 head copies of the reduced TP output to shared memory ----
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc0[h][k] += __shfl_xor_sync(0xffffffffu, acc0[h][k], 16);
#pragma unroll
            for (int k = 0; k < 20; ++k) { acc1[h][k] += __shfl_xor_sync(0xffffffffu, acc1[h][k], 8); acc1[h][k] += __shfl_xor_sync(0xffffffffu, acc1[h][k], 16); }
#pragma unroll
            for (int k = 0; k < 22; ++k) {
                acc2[h][k] += __shfl_xor_sync(0xffffffffu, acc2[h][k], 4); acc2[h][k] += __shfl_xor_sync(0xffffffffu, acc2[h][k], 8);
                acc2[h][k] += __shfl_xor_sync(0xffffffffu, acc2[h][k], 16);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float* Dh = s_D + (size_t)(2 * hp + h) * D::FOUT;
            if (lane < 16) {
                Dh[D::C0_K0 + ch0] = acc0[h][0] * w0[0];
#pragma unroll
                for (int k = 0; k < 3; ++k) Dh[B1 + (D::C1_K1 + ch0) * 3 + k] = acc0[h][1 + k] * w0[1];
#pragma unroll
                for (int k = 0; k < 5; ++k) Dh[B2 + (D::C2_K2 + ch0) * 5 + k] = acc0[h][4 + k] * w0[2];
            }
            if (lane < 8) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    Dh[B1 + (D::C1_K3 + ch1) * 3 + k] = acc1[h][k] * w1[0]; Dh[B1 + (D::C1_K5 + ch1) * 3 + k] = acc1[h][4 + k] * w1[2];
                    Dh[B1 + (D::C1_K7 + ch1) * 3 + k] = acc1[h][12 + k] * w1[4];
                }
                Dh[D::C0_K4 + ch1] = acc1[h][3] * w1[1];
#pragma unroll
                for (int k = 0; k < 5; ++k) { Dh[B2 + (D::C2_K6 + ch1) * 5 + k] = acc1[h][7 + k] * w1[3]; Dh[B2 + (D::C2_K8 + ch1) * 5 + k] = acc1[h][15 + k] * w1[5]; }
            }
            if (lane < 4) {
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    Dh[B2 + (D::C2_K9 + ch2) * 5 + k] = acc2[h][k] * w2[0]; Dh[B2 + (D::C2_K11 + ch2) * 5 + k] = acc2[h][8 + k] * w2[2];
                    Dh[B2 + (D::C2_K14 + ch2) * 5 + k] = acc2[h][17 + k] * w2[5];
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) { Dh[B1 + (D::C1_K10 + ch2) * 3 + k] = acc2[h][5 + k] * w2[1]; Dh[B1 + (D::C1_K13 + ch2) * 3 + k] = acc2[h][14 + k] * w2[4]; }
                Dh[D::C0_K12 + ch2] = acc2[h][13] * w2[3];
            }
        }
        VR_STAMP(10);
        if (v_pending) { mbar_wait(&vbar, 0); v_pending = false; }
        __syncthreads();
        VR_STAMP(11);
        // ---- linear layer, once per destination: out[c] = sum_k V_l[k, u] D_{head(u)}[k, m] (+ bias * sum alpha) ----
        // One thread per (l, m, channel PAIR u, u + 1): the two outputs read the same reduced TP value D[k, m] (same head: the heads'
        // channel ranges are even) and adjacent weights V[k, u], V[k, u + 1] -- one LDS.64 + one LDS.32 per two FMAs instead of four
        // LDS.32 (the phase is bound by shared-memory wavefronts).  Same products in the same order per output as before.
        constexpr int P0 = D::M0 / 2, P1 = 3 * (D::M1 / 2), P2 = 5 * (D::M2 / 2);
        static_assert((D::M0 / 4) % 2 == 0 && (D::M1 / 4) % 2 == 0 && (D::M2 / 4) % 2 == 0, "channel pairs inside one head");
        for (int t = tid; t < P0 + P1 + P2; t += NT) {
            int l, u, m;
            if (t < P0) { l = 0; u = 2 * t; m = 0; }
            else if (t < P0 + P1) { l = 1; u = 2 * ((t - P0) / 3); m = (t - P0) % 3; }
            else { l = 2; u = 2 * ((t - P0 - P1) / 5); m = (t - P0 - P1) % 5; }
            const int ML = (l == 0) ? D::M0 : (l == 1) ? D::M1 : D::M2;
            const int KL = (l == 0) ? D::D0 : (l == 1) ? D::D1 : D::D2;
            const int dd = 2 * l + 1;
            const int h = u / (ML / 4);
            const float* V = s_V + ((l == 0) ? 0 : (l == 1) ? NV0 : NV0 + NV1) + u;
            const float* Dl = s_D + (size_t)h * D::FOUT + ((l == 0) ? 0 : (l == 1) ? B1 : B2) + m;
            float acc0 = 0.f, accb0 = 0.f, acc1 = 0.f, accb1 = 0.f;
            for (int k = 0; k < KL; k += 4) {                      // every D_l is a multiple of 4
                const float2 v0 = *reinterpret_cast<const float2*>(V + k * ML), v1 = *reinterpret_cast<const float2*>(V + (k + 1) * ML);
                const float2 v2 = *reinterpret_cast<const float2*>(V + (k + 2) * ML), v3 = *reinterpret_cast<const float2*>(V + (k + 3) * ML);
                const float d0 = Dl[k * dd], d1 = Dl[(k + 1) * dd], d2 = Dl[(k + 2) * dd], d3 = Dl[(k + 3) * dd];
                acc0 = fmaf(v0.x, d0, acc0); accb0 = fmaf(v1.x, d1, accb0); acc0 = fmaf(v2.x, d2, acc0); accb0 = fmaf(v3.x, d3, accb0);
                acc1 = fmaf(v0.y, d0, acc1); accb1 = fmaf(v1.y, d1, accb1); acc1 = fmaf(v2.y, d2, acc1); accb1 = fmaf(v3.y, d3, accb1);
            }
            acc0 += accb0; acc1 += accb1;
            if (l == 0 && a.vb) { acc0 = fmaf(a.vb[u], s_sal[h], acc0); acc1 = fmaf(a.vb[u + 1], s_sal[h], acc1); }
            const int c = ((l == 0) ? 0 : (l == 1) ? D::M0 : D::M0 + 3 * D::M1) + u * dd + m;
            a.out[(size_t)d * D::F + c] = acc0;
            a.out[(size_t)d * D::F + c + dd] = acc1;
        }
        VR_STAMP(12);
    }
}
#ifdef DEDF_VR_TRACE
}  // namespace dedf
extern "C" int dedf_vr_trace(long long* host_out24) {
    return cudaMemcpyFromSymbol(host_out24, dedf::g_vr_trace, sizeof(long long) * 24) == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}
namespace dedf {
#endif

// ---------------------------------------------------------------------------


// ===========================================================================
// K1: gather -> depthwise CG TP (per-edge weights) -> x alpha_h -> segment reduce
// ===========================================================================
struct TpReduceArgs {
    const float* x;        // (N_src, F)
    const int* row_ptr;    // (N_dst + 1) CSR by destination
    const int* edge_src;   // (E)
    const float* sh;       // (E, 9)
    const float* w;        // (E, NUMEL)
    const float* alpha;    // (E, 4)
    float* out;            // (N_dst, FOUT)
    int n_dst;
    int task_d;            // destinations per warp task (TMA kernel), 1..16
};

template <int G>
__global__ void __launch_bounds__(256) edge_tp_reduce_kernel(TpReduceArgs a) {
    using D = Dtp<G>;
    constexpr int P = D::P;
    constexpr int S0 = 4 / P;   // distinct l0 channels per lane (2 / 1)
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    float* stage = smem + (size_t)warp * D::FOUT;
    for (int d = blockIdx.x * wpb + warp; d < a.n_dst; d += gridDim.x * wpb) {
        const int eb = a.row_ptr[d], ee = a.row_ptr[d + 1];
        // accumulators, indexed by slot: slots that map to the same channel (different edge of
        // the pack) are summed when written out.
        float acc0[4][9], acc1[2][20], acc2[22];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int k = 0; k < 9; ++k) acc0[s][k] = 0.f;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int k = 0; k < 20; ++k) acc1[s][k] = 0.f;
#pragma unroll
        for (int k = 0; k < 22; ++k) acc2[k] = 0.f;

        for (int pe = eb; pe < ee; pe += P) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                int ei, ch; D::slot0(lane, s, ei, ch);
                const int e = pe + ei;
                if (e < ee) {
                    const float al = a.alpha[(size_t)e * 4 + ch / (D::M0 / 4)];
                    const float x = a.x[(size_t)a.edge_src[e] * D::F + ch];
                    const float* w = a.w + (size_t)e * D::NUMEL;
                    float o[9];
                    dtp_l0(x * al, w[D::W_K0 + ch], w[D::W_K1 + ch], w[D::W_K2 + ch], a.sh + (size_t)e * 9, o);
#pragma unroll
                    for (int k = 0; k < 9; ++k) acc0[s][k] += o[k];
                }
            }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                int ei, ch; D::slot1(lane, s, ei, ch);
                const int e = pe + ei;
                if (e < ee) {
                    const float al = a.alpha[(size_t)e * 4 + ch / (D::M1 / 4)];
                    const float* xs = a.x + (size_t)a.edge_src[e] * D::F + D::M0 + 3 * ch;
                    float x[3] = {xs[0] * al, xs[1] * al, xs[2] * al}, w[6], o[20];
                    const float* wp = a.w + (size_t)e * D::NUMEL + D::W_K3 + ch;
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M1];
                    dtp_l1(x, w, a.sh + (size_t)e * 9, o);
#pragma unroll
                    for (int k = 0; k < 20; ++k) acc1[s][k] += o[k];
                }
            }
            {
                int ei, ch; D::slot2(lane, ei, ch);
                const int e = pe + ei;
                if (e < ee) {
                    const float al = a.alpha[(size_t)e * 4 + ch / (D::M2 / 4)];
                    const float* xs = a.x + (size_t)a.edge_src[e] * D::F + D::M0 + 3 * D::M1 + 5 * ch;
                    float x[5], w[6], o[22];
#pragma unroll
                    for (int i = 0; i < 5; ++i) x[i] = xs[i] * al;
                    const float* wp = a.w + (size_t)e * D::NUMEL + D::W_K9 + ch;
#pragma unroll
                    for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M2];
                    dtp_l2(x, w, a.sh + (size_t)e * 9, o);
#pragma unroll
                    for (int k = 0; k < 22; ++k) acc2[k] += o[k];
                }
            }
        }
        // ---- fold the pack dimension ------------------------------------------------
        // l0: slots s and s' with the same channel: G=32: (0,2),(1,3); G=16: all four slots -> channel = lane
        // l1: G=32: slots (0,1) same channel; G=16: slot s, lane>>4 selects the edge -> fold across lane^16 too
        // l2: fold across the P lane groups of width M2
        if (G == 32) {
#pragma unroll
            for (int k = 0; k < 9; ++k) { acc0[0][k] += acc0[2][k]; acc0[1][k] += acc0[3][k]; }
#pragma unroll
            for (int k = 0; k < 20; ++k) acc1[0][k] += acc1[1][k];
#pragma unroll
            for (int k = 0; k < 22; ++k) acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 16);
        } else {
#pragma unroll
            for (int k = 0; k < 9; ++k) acc0[0][k] += acc0[1][k] + acc0[2][k] + acc0[3][k];
#pragma unroll
            for (int k = 0; k < 20; ++k) { acc1[0][k] += acc1[1][k]; acc1[0][k] += __shfl_xor_sync(0xffffffffu, acc1[0][k], 16); }
#pragma unroll
            for (int k = 0; k < 22; ++k) { acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 8); acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 16); }
        }
        // ---- stage the output row in shared memory in the sorted-irreps layout ---------
        constexpr int B1 = D::D0, B2 = D::D0 + 3 * D::D1;      // block offsets of lo=1 / lo=2
        __syncwarp();
#pragma unroll
        for (int s = 0; s < S0; ++s) {
            const int ch = lane + 32 * s;                       // G=32: channels lane, lane+32 ; G=16: lane
            stage[D::C0_K0 + ch] = acc0[s][0];
#pragma unroll
            for (int k = 0; k < 3; ++k) stage[B1 + (D::C1_K1 + ch) * 3 + k] = acc0[s][1 + k];
#pragma unroll
            for (int k = 0; k < 5; ++k) stage[B2 + (D::C2_K2 + ch) * 5 + k] = acc0[s][4 + k];
        }
        if (lane < D::M1) {
            const int ch = lane;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                stage[B1 + (D::C1_K3 + ch) * 3 + k] = acc1[0][k];
                stage[B1 + (D::C1_K5 + ch) * 3 + k] = acc1[0][4 + k];
                stage[B1 + (D::C1_K7 + ch) * 3 + k] = acc1[0][12 + k];
            }
            stage[D::C0_K4 + ch] = acc1[0][3];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                stage[B2 + (D::C2_K6 + ch) * 5 + k] = acc1[0][7 + k];
                stage[B2 + (D::C2_K8 + ch) * 5 + k] = acc1[0][15 + k];
            }
        }
        if (lane < D::M2) {
            const int ch = lane;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                stage[B2 + (D::C2_K9 + ch) * 5 + k] = acc2[k];
                stage[B2 + (D::C2_K11 + ch) * 5 + k] = acc2[8 + k];
                stage[B2 + (D::C2_K14 + ch) * 5 + k] = acc2[17 + k];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                stage[B1 + (D::C1_K10 + ch) * 3 + k] = acc2[5 + k];
                stage[B1 + (D::C1_K13 + ch) * 3 + k] = acc2[14 + k];
            }
            stage[D::C0_K12 + ch] = acc2[13];
        }
        __syncwarp();
        float4* dst = reinterpret_cast<float4*>(a.out + (size_t)d * D::FOUT);
        const float4* srcv = reinterpret_cast<const float4*>(stage);
        for (int i = lane; i < D::FOUT / 4; i += 32) dst[i] = srcv[i];
    }
}


// ---------------------------------------------------------------------------
// K1, TMA variant: every warp owns a private multi-stage ring in shared memory that is filled by 1-D bulk
// async copies (cp.async.bulk global -> shared, SASS UBLKCP) signalled through mbarriers: one stage = one pack
// of P edges = [P weight rows | P gathered source-feature rows | P spherical harmonics (padded to 12) | P alphas].
// The CG math reads only shared memory; the output row is staged and written with coalesced float4 stores.
// ---------------------------------------------------------------------------
constexpr int kK1Warps = 12;
constexpr int kK1Stages = 3;
constexpr int kK1TaskD = 16;      // destinations per warp task (row pointers live in lanes 0..16)

// CG math of one pack read from a shared-memory stage.  FULL = all P edges valid: no per-slot predicates.
template <int G, bool FULL>
__device__ __forceinline__ void k1_pack_math(int lane, int nv, const float* __restrict__ sw, const float* __restrict__ sx,
                                             const float* __restrict__ ssh, const float* __restrict__ sal,
                                             float (&acc0)[4][9], float (&acc1)[2][20], float (&acc2)[22]) {
    using D = Dtp<G>;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        int ei, ch; D::slot0(lane, s, ei, ch);
        if (FULL || ei < nv) {
            const float al = sal[ei * 4 + ch / (D::M0 / 4)];
            const float* w = sw + ei * D::NUMEL;
            float o[9];
            dtp_l0(sx[ei * D::F + ch] * al, w[D::W_K0 + ch], w[D::W_K1 + ch], w[D::W_K2 + ch], ssh + ei * 12, o);
#pragma unroll
            for (int k = 0; k < 9; ++k) acc0[s][k] += o[k];
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        int ei, ch; D::slot1(lane, s, ei, ch);
        if (FULL || ei < nv) {
            const float al = sal[ei * 4 + ch / (D::M1 / 4)];
            const float* xs = sx + ei * D::F + D::M0 + 3 * ch;
            float x[3] = {xs[0] * al, xs[1] * al, xs[2] * al}, w[6], o[20];
            const float* wp = sw + ei * D::NUMEL + D::W_K3 + ch;
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M1];
            dtp_l1(x, w, ssh + ei * 12, o);
#pragma unroll
            for (int k = 0; k < 20; ++k) acc1[s][k] += o[k];
        }
    }
    {
        int ei, ch; D::slot2(lane, ei, ch);
        if (FULL || ei < nv) {
            const float al = sal[ei * 4 + ch / (D::M2 / 4)];
            const float* xs = sx + ei * D::F + D::M0 + 3 * D::M1 + 5 * ch;
            float x[5], w[6], o[22];
#pragma unroll
            for (int i = 0; i < 5; ++i) x[i] = xs[i] * al;
            const float* wp = sw + ei * D::NUMEL + D::W_K9 + ch;
#pragma unroll
            for (int i = 0; i < 6; ++i) w[i] = wp[i * D::M2];
            dtp_l2(x, w, ssh + ei * 12, o);
#pragma unroll
            for (int k = 0; k < 22; ++k) acc2[k] += o[k];
        }
    }
}

template <int G>
struct K1Stage {
    using D = Dtp<G>;
    static constexpr int P = D::P;
    static constexpr int W_OFF = 0, X_OFF = P * D::NUMEL, SH_OFF = X_OFF + P * D::F, AL_OFF = SH_OFF + P * 12;
    static constexpr int FLOATS = AL_OFF + P * 4;
    static constexpr int EDGE_BYTES = (D::NUMEL + D::F + 12 + 4) * 4;
    static constexpr int WARP_FLOATS = kK1Stages * FLOATS;
};

template <int G>
__global__ void __launch_bounds__(kK1Warps * 32, 1) edge_tp_reduce_tma_kernel(TpReduceArgs a) {
    using D = Dtp<G>;
    using ST = K1Stage<G>;
    constexpr int P = D::P, S = kK1Stages;
    constexpr int S0 = 4 / P;
    extern __shared__ __align__(128) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem) + warp * S;                 // [warps][S]
    float* base = smem + (kK1Warps * S * 2 + 31) / 32 * 32 + (size_t)warp * ST::WARP_FLOATS;
    // streamed operands (weights / harmonics / alphas) are read once: evict-first, so that the gathered feature
    // table (evict-last) stays resident in the 126 MB L2 across the launch
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    uint32_t phase = 0;                         // bit s = parity to wait for on stage s
    const int task_d = a.task_d;
    const int n_tasks = (a.n_dst + task_d - 1) / task_d;
    const int gw = blockIdx.x * kK1Warps + warp, nw = gridDim.x * kK1Warps;

    for (int task = gw; task < n_tasks; task += nw) {
        const int d0 = task * task_d;
        const int nd = min(task_d, a.n_dst - d0);
        const int rp_l = a.row_ptr[d0 + min(lane, nd)];
        const int e_end = __shfl_sync(0xffffffffu, rp_l, nd);
        // source-index windows (current / next 32 edges of this task's contiguous edge range)
        int wbase = __shfl_sync(0xffffffffu, rp_l, 0);
        int srcw = (wbase + lane < e_end) ? a.edge_src[wbase + lane] : 0;
        int srcn = (wbase + 32 + lane < e_end) ? a.edge_src[wbase + 32 + lane] : 0;
        // producer iterator over the task's packs
        int pd = 0, ppe = 0, pee = 0;
        while (pd < nd) {
            ppe = __shfl_sync(0xffffffffu, rp_l, pd); pee = __shfl_sync(0xffffffffu, rp_l, pd + 1);
            if (ppe < pee) break;
            ++pd;
        }
        int pstage = 0;
        auto issue = [&](int stage) {
            if (pd >= nd) return;
            while (ppe >= wbase + 32) {
                srcw = srcn; wbase += 32;
                srcn = (wbase + 32 + lane < e_end) ? a.edge_src[wbase + 32 + lane] : 0;
            }
            const int nv = min(P, pee - ppe);
            int srcs[P];
#pragma unroll
            for (int i = 0; i < P; ++i) {
                const int idx = ppe + i - wbase;
                const int v0 = __shfl_sync(0xffffffffu, srcw, idx & 31), v1 = __shfl_sync(0xffffffffu, srcn, idx & 31);
                srcs[i] = (idx < 32) ? v0 : v1;
            }
            if (lane == 0) {
                float* st = base + stage * ST::FLOATS;
                uint64_t* bar = bars + stage;
                mbar_expect_tx(bar, (uint32_t)(nv * ST::EDGE_BYTES));
                bulk_g2s_hint(st + ST::W_OFF, a.w + (size_t)ppe * D::NUMEL, (uint32_t)(nv * D::NUMEL * 4), bar, pol_stream);
#pragma unroll
                for (int i = 0; i < P; ++i)
                    if (i < nv) bulk_g2s_hint(st + ST::X_OFF + i * D::F, a.x + (size_t)srcs[i] * D::F, (uint32_t)(D::F * 4), bar, pol_keep);
                bulk_g2s_hint(st + ST::SH_OFF, a.sh + (size_t)ppe * 12, (uint32_t)(nv * 48), bar, pol_stream);
                bulk_g2s_hint(st + ST::AL_OFF, a.alpha + (size_t)ppe * 4, (uint32_t)(nv * 16), bar, pol_stream);
            }
            ppe += P;
            if (ppe >= pee) {
                ++pd;
                while (pd < nd) {
                    ppe = __shfl_sync(0xffffffffu, rp_l, pd); pee = __shfl_sync(0xffffffffu, rp_l, pd + 1);
                    if (ppe < pee) break;
                    ++pd;
                }
            }
        };
#pragma unroll
        for (int s = 0; s < S; ++s) { issue(pstage); pstage = (pstage + 1) % S; }

        int cstage = 0;
        for (int cd = 0; cd < nd; ++cd) {
            const int ceb = __shfl_sync(0xffffffffu, rp_l, cd), cee = __shfl_sync(0xffffffffu, rp_l, cd + 1);
            float acc0[4][9], acc1[2][20], acc2[22];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int k = 0; k < 9; ++k) acc0[s][k] = 0.f;
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int k = 0; k < 20; ++k) acc1[s][k] = 0.f;
#pragma unroll
            for (int k = 0; k < 22; ++k) acc2[k] = 0.f;

            for (int pe = ceb; pe < cee; pe += P) {
                const int nv = min(P, cee - pe);
                mbar_wait(bars + cstage, (phase >> cstage) & 1u);
                phase ^= (1u << cstage);
                const float* st = base + cstage * ST::FLOATS;
                const float* sw = st + ST::W_OFF;
                const float* sx = st + ST::X_OFF;
                const float* ssh = st + ST::SH_OFF;
                const float* sal = st + ST::AL_OFF;
                if (nv == P) k1_pack_math<G, true>(lane, nv, sw, sx, ssh, sal, acc0, acc1, acc2);
                else k1_pack_math<G, false>(lane, nv, sw, sx, ssh, sal, acc0, acc1, acc2);
                __syncwarp();                    // every lane is done reading this stage -> refill it
                issue(cstage);
                cstage = (cstage + 1) % S;
            }
            // ---- fold the pack dimension, stage the row, coalesced store (same layout as the LDG variant) ----
            if (G == 32) {
#pragma unroll
                for (int k = 0; k < 9; ++k) { acc0[0][k] += acc0[2][k]; acc0[1][k] += acc0[3][k]; }
#pragma unroll
                for (int k = 0; k < 20; ++k) acc1[0][k] += acc1[1][k];
#pragma unroll
                for (int k = 0; k < 22; ++k) acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 16);
            } else {
#pragma unroll
                for (int k = 0; k < 9; ++k) acc0[0][k] += acc0[1][k] + acc0[2][k] + acc0[3][k];
#pragma unroll
                for (int k = 0; k < 20; ++k) { acc1[0][k] += acc1[1][k]; acc1[0][k] += __shfl_xor_sync(0xffffffffu, acc1[0][k], 16); }
#pragma unroll
                for (int k = 0; k < 22; ++k) { acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 8); acc2[k] += __shfl_xor_sync(0xffffffffu, acc2[k], 16); }
            }
            constexpr int B1 = D::D0, B2 = D::D0 + 3 * D::D1;
            // every lane owns whole (channel, m) groups of the row: 4-byte streaming stores, merged into full sectors in L2
            float* outst = a.out + (size_t)(d0 + cd) * D::FOUT;
#define OUTST(i, v) __stcs(outst + (i), (v))
#pragma unroll
            for (int s = 0; s < S0; ++s) {
                const int ch = lane + 32 * s;
                OUTST(D::C0_K0 + ch, acc0[s][0]);
#pragma unroll
                for (int k = 0; k < 3; ++k) OUTST(B1 + (D::C1_K1 + ch) * 3 + k, acc0[s][1 + k]);
#pragma unroll
                for (int k = 0; k < 5; ++k) OUTST(B2 + (D::C2_K2 + ch) * 5 + k, acc0[s][4 + k]);
            }
            if (lane < D::M1) {
                const int ch = lane;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    OUTST(B1 + (D::C1_K3 + ch) * 3 + k, acc1[0][k]);
                    OUTST(B1 + (D::C1_K5 + ch) * 3 + k, acc1[0][4 + k]);
                    OUTST(B1 + (D::C1_K7 + ch) * 3 + k, acc1[0][12 + k]);
                }
                OUTST(D::C0_K4 + ch, acc1[0][3]);
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    OUTST(B2 + (D::C2_K6 + ch) * 5 + k, acc1[0][7 + k]);
                    OUTST(B2 + (D::C2_K8 + ch) * 5 + k, acc1[0][15 + k]);
                }
            }
            if (lane < D::M2) {
                const int ch = lane;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    OUTST(B2 + (D::C2_K9 + ch) * 5 + k, acc2[k]);
                    OUTST(B2 + (D::C2_K11 + ch) * 5 + k, acc2[8 + k]);
                    OUTST(B2 + (D::C2_K14 + ch) * 5 + k, acc2[17 + k]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    OUTST(B1 + (D::C1_K10 + ch) * 3 + k, acc2[5 + k]);
                    OUTST(B1 + (D::C1_K13 + ch) * 3 + k, acc2[14 + k]);
                }
                OUTST(D::C0_K12 + ch, acc2[13]);
            }
#undef OUTST
        }
    }
}

}  // namespace dedf

using namespace dedf;

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" int dedf_edge_geom(const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst,
                              const int* n_edges_dev, int max_edges, int n_scales, const int* src_off, const float* r,
                              float ns_lo, float ns_hi, float* length, float* sh, float* logit, cudaStream_t stream) {
    if (max_edges <= 0) return DEDF_OK;
    if (!x_src || !x_dst || !edge_src || !edge_dst || !n_edges_dev || !length || !sh) return DEDF_ERR_ARG;
    if (n_scales < 1 || n_scales > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
    if (logit && (!src_off || !r)) return DEDF_ERR_ARG;
    if (max_edges <= 0) return DEDF_OK;
    GeomArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.edge_src = edge_src; a.edge_dst = edge_dst; a.n_edges = n_edges_dev;
    a.length = length; a.sh = sh; a.logit = logit; a.ns_lo = ns_lo; a.ns_hi = ns_hi; a.n_scales = n_scales;
    for (int s = 0; s < n_scales; ++s) { a.src_off[s] = src_off ? src_off[s] : 0; a.r[s] = r ? r[s] : -1.f; }
    a.src_off[n_scales] = src_off ? src_off[n_scales] : 0x7fffffff;
    launch_pdl(edge_geom_kernel, dim3(grid_for(max_edges, 256, kNumSMs * 8)), dim3(256), 0, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_edge_mlp(const dedf_mlp_desc* d, int max_edges, cudaStream_t stream) {
    if (max_edges <= 0) return DEDF_OK;
    if (!d || !d->n_edges_dev || !d->out) return DEDF_ERR_ARG;
    if (d->n_layers < 1 || d->n_layers > DEDF_MLP_MAX_LAYERS) return DEDF_ERR_ARG;
    for (int i = 0; i <= d->n_layers; ++i) {
        if (d->dims[i] < 4 || d->dims[i] % 4) return DEDF_ERR_ARG;
        if (i < d->n_layers && d->dims[i] > kMlpMaxW) return DEDF_ERR_ARG;      // hidden widths staged in smem
    }
    if (max_edges <= 0) return DEDF_OK;
    MlpArgs a{};
    a.mode = d->mode; a.n_edges = d->n_edges_dev; a.x_in = d->x_in; a.length = d->length;
    a.rbf_mean = d->rbf_mean; a.rbf_std_logit = d->rbf_std_logit; a.rbf_weight_logit = d->rbf_weight_logit;
    a.rbf_cutoff = d->rbf_cutoff; a.rbf_offset = d->rbf_offset;
    a.n_scales = d->n_scales; a.n_dst = d->n_dst; a.row_ptr = d->row_ptr; a.edge_dst = d->edge_dst;
    a.enc_max_r = d->enc_max_r; a.enc_n = d->enc_n; a.enc_freq = d->enc_freq; a.row_bias = d->row_bias; a.n_rb = d->n_rb; a.rb_div = d->rb_div;
    if (a.mode == DEDF_MLP_IN_ROWS) { if (!a.x_in) return DEDF_ERR_ARG; }
    else if (a.mode == DEDF_MLP_IN_RBF) { if (!a.length || !a.rbf_mean || !a.rbf_std_logit || !a.rbf_weight_logit) return DEDF_ERR_ARG; }
    else if (a.mode == DEDF_MLP_IN_FIELD) {
        if (!a.length || !a.row_ptr || !a.edge_dst || a.n_scales < 1 || a.n_scales > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
        if (a.row_bias && (a.rb_div < 1 || a.n_rb < 1)) return DEDF_ERR_ARG;
        for (int s = 0; s < a.n_scales; ++s) {
            a.enc_mean[s] = d->enc_mean[s]; a.enc_std_logit[s] = d->enc_std_logit[s]; a.enc_weight_logit[s] = d->enc_weight_logit[s];
            a.enc_r[s] = d->enc_r[s]; a.pre_w[s] = d->pre_w[s];
            if (!a.pre_w[s]) return DEDF_ERR_ARG;
            if (a.enc_r[s] >= 0.f && (!a.enc_mean[s] || !a.enc_std_logit[s] || !a.enc_weight_logit[s])) return DEDF_ERR_ARG;
            if (a.enc_r[s] < 0.f && !a.enc_freq) return DEDF_ERR_ARG;
        }
    } else return DEDF_ERR_ARG;
    for (int i = 0; i <= d->n_layers; ++i) a.K[i] = d->dims[i];
    a.n_layers = d->n_layers;
    for (int i = 0; i < d->n_layers; ++i) {
        a.W[i] = d->W[i]; a.b[i] = d->b[i]; a.ln_g[i] = d->ln_g[i]; a.ln_b[i] = d->ln_b[i]; a.flags[i] = d->flags[i];
        if (!(a.mode == DEDF_MLP_IN_FIELD && i == 0) && !a.W[i]) return DEDF_ERR_ARG;
        if ((a.flags[i] & 1) && (!a.ln_g[i] || !a.ln_b[i])) return DEDF_ERR_ARG;
    }
    a.out_offset = d->out_offset; a.out = d->out;
    const size_t act_bytes = (size_t)2 * kMlpTE * pad_lda(kMlpMaxW) * sizeof(float);
    // stage the weights in shared memory when they fit beside the activation buffers (all but the field's RadialProfile)
    size_t w_floats = (size_t)((a.mode == DEDF_MLP_IN_FIELD) ? a.n_scales : 1) * a.K[0] * a.K[1];
    for (int i = 1; i < a.n_layers; ++i) w_floats += (size_t)a.K[i] * a.K[i + 1];
    constexpr size_t kMaxSmem = 220 * 1024;
    bool w_ok = act_bytes + w_floats * sizeof(float) <= kMaxSmem;
    for (int i = 0; i < a.n_layers && w_ok; ++i) {
        if (a.mode == DEDF_MLP_IN_FIELD && i == 0) { for (int s = 0; s < a.n_scales; ++s) w_ok = w_ok && (reinterpret_cast<uintptr_t>(a.pre_w[s]) & 15) == 0; }
        else w_ok = (reinterpret_cast<uintptr_t>(a.W[i]) & 15) == 0;
    }
    a.w_smem = w_ok ? 1 : 0;
    const size_t smem = act_bytes + (w_ok ? w_floats * sizeof(float) : 0);
    static bool attr_done = false;
    if (!attr_done) { cudaFuncSetAttribute(edge_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem); attr_done = true; }
    const int n_tiles = (max_edges + kMlpTE - 1) / kMlpTE + DEDF_MAX_SCALES;
    launch_pdl(edge_mlp_kernel, dim3(grid_for(n_tiles, 1, kNumSMs * 3)), dim3(kMlpThreads), smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

template <int G, int EPI>
static int launch_tp_lin(const TpLinArgs& a, int max_edges, cudaStream_t stream) {
    using C = TpLinCfg<G, EPI>;
    using D = Dtp<G>;
    const int lda0 = pad_lda(D::D0), lda1 = pad_lda(D::D1), lda2 = pad_lda(D::D2);
    const size_t smem = ((size_t)C::TE * lda0 + 3 * C::TE * lda1 + 5 * C::TE * lda2 + C::TE * 9 + 2 * C::TE) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(edge_tp_lin_kernel<G, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        // ask for just enough shared memory for the resident CTAs: the rest of the 256 KB stays L1 and caches the weights
        const int ctas = (smem * 4 <= 120 * 1024) ? 4 : 2;
        int pct = (int)((smem * ctas + 2048 * ctas) * 100 / (228 * 1024)) + 1;
        if (pct > 100) pct = 100;
        cudaFuncSetAttribute(edge_tp_lin_kernel<G, EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
        attr_done = true;
    }
    const int n_tiles = (max_edges + C::TE - 1) / C::TE;
    launch_pdl((edge_tp_lin_kernel<G, EPI>), dim3(grid_for(n_tiles, 1, kNumSMs * 2)), dim3(C::THREADS), smem, stream, a, lda0, lda1, lda2);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

static long long* g_tpl_dbg = nullptr;
/* debug hook (not in the public header): 64 x int64 device buffer receiving clock64() stamps of CTA 0 of dedf_edge_tp_lin */
extern "C" int dedf_tp_lin_set_debug(long long* dbg) { g_tpl_dbg = dbg; return DEDF_OK; }

extern "C" int dedf_edge_tp_lin(int mul1, int epilogue, const float* x_src, const float* x_dst, int per_edge_x,
                                const int* edge_src, const int* edge_dst, const int* n_edges_dev, int max_edges,
                                const float* sh, const float* w, long long w_stride, const float* W0, const float* W1,
                                const float* W2, const float* bias0, const float* alpha_dot, const float* edge_logit,
                                float* logits, float* out, cudaStream_t stream) {
    if (max_edges <= 0) return DEDF_OK;
    if (!x_src || !edge_dst || !n_edges_dev || !sh || !w || !W0 || !W1 || !W2 || !out) return DEDF_ERR_ARG;
    if (!per_edge_x && !edge_src) return DEDF_ERR_ARG;
    if (epilogue == DEDF_EPI_ACT && (!alpha_dot || !logits)) return DEDF_ERR_ARG;
    if (max_edges <= 0) return DEDF_OK;
    TpLinArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.edge_src = edge_src; a.edge_dst = edge_dst; a.n_edges = n_edges_dev; a.sh = sh;
    a.w = w; a.w_stride = w_stride; a.W0 = W0; a.W1 = W1; a.W2 = W2; a.bias0 = bias0; a.alpha_dot = alpha_dot;
    a.edge_logit = edge_logit; a.logits = logits; a.out = out; a.per_edge_x = per_edge_x;
    a.dbg = g_tpl_dbg;
    if (mul1 == 32 && epilogue == DEDF_EPI_ACT) return launch_tp_lin<32, EPI_ACT>(a, max_edges, stream);
    if (mul1 == 32 && epilogue == DEDF_EPI_LIN) return launch_tp_lin<32, EPI_LIN>(a, max_edges, stream);
    if (mul1 == 16 && epilogue == DEDF_EPI_ACT) return launch_tp_lin<16, EPI_ACT>(a, max_edges, stream);
    if (mul1 == 16 && epilogue == DEDF_EPI_LIN) return launch_tp_lin<16, EPI_LIN>(a, max_edges, stream);
    return DEDF_ERR_UNSUPPORTED;
}

extern "C" int dedf_segment_softmax_reduce(const int* row_ptr, int n_dst, int n_seg, const float* logits,
                                           const float* val, int m0, int m1, int m2, float* out, cudaStream_t stream) {
    if (!row_ptr || !logits || !val || !out || n_seg < 1) return DEDF_ERR_ARG;
    if (m0 % 4 || m1 % 4 || m2 % 4 || m0 + 3 * m1 + 5 * m2 > 256) return DEDF_ERR_UNSUPPORTED;
    if (n_dst <= 0) return DEDF_OK;
    SoftmaxArgs a{row_ptr, n_dst, n_seg, logits, val, out, m0, m1, m2};
    launch_pdl(segment_softmax_reduce_kernel, dim3(grid_for(n_dst, 1, kNumSMs * 16)), dim3(128), 0, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

template <int G>
static int launch_k1_tma(const TpReduceArgs& a, cudaStream_t stream) {
    using ST = K1Stage<G>;
    const size_t smem = ((size_t)(kK1Warps * kK1Stages * 2 + 31) / 32 * 32 + (size_t)kK1Warps * ST::WARP_FLOATS) * sizeof(float);
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(edge_tp_reduce_tma_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
    // small graphs: shrink the per-warp task so that every warp of the 148 CTAs has work (a task is sequential)
    TpReduceArgs b = a;
    b.task_d = max(1, min(kK1TaskD, a.n_dst / (kNumSMs * kK1Warps * 2)));
    const int n_tasks = (a.n_dst + b.task_d - 1) / b.task_d;
    edge_tp_reduce_tma_kernel<G><<<grid_for(n_tasks, kK1Warps, kNumSMs), kK1Warps * 32, smem, stream>>>(b);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_edge_tp_reduce(int mul1, const float* x, const int* row_ptr, const int* edge_src, const float* sh,
                                   int sh_stride, const float* w, const float* alpha, int n_dst, float* out,
                                   cudaStream_t stream) {
    if (!x || !row_ptr || !edge_src || !sh || !w || !alpha || !out) return DEDF_ERR_ARG;
    if (sh_stride != 9 && sh_stride != 12) return DEDF_ERR_ARG;
    if (n_dst <= 0) return DEDF_OK;
    TpReduceArgs a{x, row_ptr, edge_src, sh, w, alpha, out, n_dst, kK1TaskD};
    if (sh_stride == 12) {      // bulk-copy (TMA) pipeline: needs 16-byte rows everywhere
        if (mul1 == 32) return launch_k1_tma<32>(a, stream);
        if (mul1 == 16) return launch_k1_tma<16>(a, stream);
        return DEDF_ERR_UNSUPPORTED;
    }
    if (mul1 == 32) {
        const size_t smem = (size_t)8 * Dtp<32>::FOUT * sizeof(float);
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(edge_tp_reduce_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
        edge_tp_reduce_kernel<32><<<grid_for(n_dst, 8, kNumSMs * 4), 256, smem, stream>>>(a);
    } else if (mul1 == 16) {
        const size_t smem = (size_t)8 * Dtp<16>::FOUT * sizeof(float);
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(edge_tp_reduce_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
        edge_tp_reduce_kernel<16><<<grid_for(n_dst, 8, kNumSMs * 4), 256, smem, stream>>>(a);
    } else return DEDF_ERR_UNSUPPORTED;
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

template <int G>
static int launch_value_reduce(const ValueReduceArgs& a, cudaStream_t stream) {
    using D = Dtp<G>;
    const size_t smem = ((size_t)kVrChunk * (D::F + 4 + 12) + 4 * (size_t)D::FOUT + (size_t)D::D0 * D::M0 + (size_t)D::D1 * D::M1 + (size_t)D::D2 * D::M2) * sizeof(float);
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(value_reduce_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); done = true; }
    // Up to two destinations per resident CTA slot (one CTA per SM at G = 32, three at G = 16): persistent CTAs that walk their
    // destinations -- the prologue (tensor-product weights into registers, the linear layer's weights into shared memory) is paid
    // once per CTA instead of once per destination: 39.0 -> 35.5 us for the 256 destinations of a 128-pose denoise step.  More
    // destinations: one CTA each, handed out by the hardware as SMs free up -- the static striding of persistent CTAs loses more
    // to the uneven degrees than the prologues cost (2048 destinations: 206 us persistent against 181 us).
    const int resident = kNumSMs * (G == 16 ? 3 : 1);
    const int max_ctas = (a.n_dst <= 2 * resident) ? resident : kNumSMs * 8;
    launch_pdl(value_reduce_kernel<G>, dim3(grid_for(a.n_dst, 1, max_ctas)), dim3(G * 8), smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_value_reduce(int mul1, const int* row_ptr, int n_dst, int n_seg, const float* v, const float* sh, const float* logits,
                                 const float* post, const float* wv, const float* V0, const float* V1, const float* V2, const float* vb,
                                 float* out, cudaStream_t stream) {
    if (!row_ptr || !v || !sh || !logits || !wv || !V0 || !V1 || !V2 || !out || n_seg < 1) return DEDF_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(v) & 15) || (reinterpret_cast<uintptr_t>(logits) & 15)) return DEDF_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(V0) & 15) || (reinterpret_cast<uintptr_t>(V1) & 15) || (reinterpret_cast<uintptr_t>(V2) & 15)) return DEDF_ERR_ARG;
    if (n_seg > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
    if (n_dst <= 0) return DEDF_OK;
    ValueReduceArgs a{row_ptr, n_dst, n_seg, v, sh, logits, post, wv, V0, V1, V2, vb, out};
    if (mul1 == 32) return launch_value_reduce<32>(a, stream);
    if (mul1 == 16) return launch_value_reduce<16>(a, stream);
    return DEDF_ERR_UNSUPPORTED;
}
