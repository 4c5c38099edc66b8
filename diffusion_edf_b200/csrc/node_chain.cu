// The per-node tail of an Equiformer block in ONE launch:
//
//     y1 = proj(x) + b_p (+ res1)                         GraphAttention.proj (+ the block's first residual)
//     y  = y1 + fctp_2(Gate(fctp_1(LN(y1))))              norm_2 / post_norm -> FeedForwardNetwork -> second residual
//
// Replaces, per block, three dedf_node_linear launches (proj -> [LN, fctp_1, gate] -> [fctp_2, +res]) and the two global
// round trips between them.  Reference: /root/reference/diffusion_edf/graph_attention.py:118-121,268-272 (proj),
// gnn_block.py:51-57,207-216 and block.py:51-57,165-173 (FeedForwardNetwork + residuals),
// equiformer/layer_norm.py:91-156 (EquivariantLayerNormV2), equiformer/fast_activation.py:210-224 (Gate),
// equiformer/tensor_product_rescale.py:176-185 (LinearRS semantics: y[w,m] = sum_u W_l[u,w] x[u,m], bias on 0e only).
//
// One CTA per tile of TN nodes (8 for the 240-dim irreps, 16 for the 120-dim ones).  All three weight sets are staged in
// shared memory by TMA bulk copies issued BEFORE the PDL wait (they are parameters): for `64x0e+32x1e+16x2e` that is
// 21.5 + 101 + 64.5 KB, so the proj weights share their region with the FFN intermediates and are re-fetched (from L2) per
// tile; everything else stays resident for the CTA's lifetime.  The arithmetic and its summation order are those of
// node_linear_tma_kernel (same gemm_item_4x4 micro-kernel, K ascending), so results are bit-identical to the un-fused path
// for the 8- and 16-node tiles; the 4-node tiles of small launches split K over four lanes (chain_gemm_splitk below).
#include "common.cuh"
#include "gemm_tile.cuh"
#include "../../include/dedf.h"

namespace dedf {

constexpr int kChainThreads = 256;

// -DDEDF_CHAIN_TRACE (profiles/run_chain_trace.py builds its own copy of the library with it): thread 0 of CTA 0 stamps clock64()
// at every phase boundary into a device array that dedf_chain_trace() copies out.  Not compiled into the shipped library.
#ifdef DEDF_CHAIN_TRACE
__device__ long long g_chain_trace[32];
#define CHAIN_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_chain_trace[i] = clock64(); } while (0)
#else
#define CHAIN_STAMP(i) do { } while (0)
#endif

struct ChainArgs {
    const float* x; int n;
    Irr emb, pre, mid;                    // mid = gate(pre): m0 = pre.m0 - pre.m1 - pre.m2
    const float *P0, *P1, *P2, *pb;       // proj    emb -> emb
    const float* res1;
    const float *ln_w, *ln_b; float ln_eps;
    const float *A0, *A1, *A2, *ab;       // fctp_1  emb -> pre
    const float *B0, *B1, *B2, *bb;       // fctp_2  mid -> emb
    float* y;
    // shared-memory plan (float offsets), computed by the host
    int o_w1, o_w2, o_y1, o_a, o_u, o_stat;
    int o_wp, o_x, o_res;                 // inside the U region (proj phase)
    int o_o0, o_m;                        // inside the U region (FFN phases)
    int wp_transient;                     // the proj weights are overwritten by the FFN phase: re-fetch them per tile
    int o_par;                            // biases + layer-norm affine parameters: [pb m0][ab pre.m0][bb m0][ln_w nirr][ln_b m0]
};

// block-diagonal GEMM over the three l blocks; epi(l, r, k, c, v): node row r, harmonic index k, output channel c.
// The l block of an item only selects pointers and sizes: ONE inlined copy of the micro-kernel and of the epilogue per GEMM (three
// copies, one per l, made the kernel 10.7 k instructions that each CTA runs once -- instruction fetch of straight-line code).
template <int TN, typename Epi>
__device__ __forceinline__ void chain_gemm(const float* A0, int lda0, const float* A1, int lda1, const float* A2, int lda2,
                                           const float* W0, const float* W1, const float* W2, Irr in, Irr out, Epi epi) {
    const int cg0 = out.m0 >> 2, cg1 = out.m1 >> 2, cg2 = out.m2 >> 2;
    const int I0 = (TN / 4) * cg0, I1 = (3 * TN / 4) * cg1, I2 = (5 * TN / 4) * cg2;
    for (int item = threadIdx.x; item < I0 + I1 + I2; item += kChainThreads) {
        float acc[4][4] = {};
        int l, cg, rg, lda, nrg, N, K; const float* A; const float* W;
        if (item < I0) { l = 0; cg = item % cg0; rg = item / cg0; A = A0; lda = lda0; nrg = TN / 4; W = W0; N = out.m0; K = in.m0; }
        else if (item < I0 + I1) { const int t = item - I0; l = 1; cg = t % cg1; rg = t / cg1; A = A1; lda = lda1; nrg = 3 * TN / 4; W = W1; N = out.m1; K = in.m1; }
        else { const int t = item - I0 - I1; l = 2; cg = t % cg2; rg = t / cg2; A = A2; lda = lda2; nrg = 5 * TN / 4; W = W2; N = out.m2; K = in.m2; }
        gemm_item_4x4<true, true>(A, lda, nrg, rg, W, N, 4 * cg, K, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = rg + i * nrg, r = row % TN, k = row / TN;
#pragma unroll
            for (int j = 0; j < 4; ++j) epi(l, r, k, 4 * cg + j, acc[i][j]);
        }
    }
}

// Few-node tiles (TN = 4: at most 2 x SMs nodes in the whole launch, e.g. the coarse UNet scales and a 128-pose denoise step): the
// launch is ONE round of latency, and with one 4 x 4 item per thread only 60 of 256 threads had work in the emb -> emb GEMMs while
// one warp walked K = 192 alone (warm-cache ncu, profiles/r2_s11_node_chain_*: 30 % of the kernel waiting for that warp).  Here KS
// adjacent lanes split an item's K range (k4-steps partitioned evenly, each slice K ascending) and add their partial sums with two
// butterfly shuffles -- (s0 + s1) + (s2 + s3): not the summation order of node_linear_tma_kernel any more (1 ulp-level differences).
template <int TN, int KS, typename Epi>
__device__ __forceinline__ void chain_gemm_splitk(const float* A0, int lda0, const float* A1, int lda1, const float* A2, int lda2,
                                                  const float* W0, const float* W1, const float* W2, Irr in, Irr out, Epi epi) {
    static_assert(KS == 4, "two butterfly rounds");
    const int cg0 = out.m0 >> 2, cg1 = out.m1 >> 2, cg2 = out.m2 >> 2;
    const int I0 = (TN / 4) * cg0, I1 = (3 * TN / 4) * cg1, I2 = (5 * TN / 4) * cg2;
    const int total = I0 + I1 + I2;
    const int ks = threadIdx.x & (KS - 1);
    constexpr int kItemsPerRound = kChainThreads / KS;
    const int rounds = (total + kItemsPerRound - 1) / kItemsPerRound;
    for (int rd = 0; rd < rounds; ++rd) {                       // warp-uniform trip count: every lane takes part in the shuffles
        const int item = rd * kItemsPerRound + (threadIdx.x >> 2);
        float acc[4][4] = {};
        int l = -1, cg = 0, rg = 0;
        if (item < total) {
            const float* A; const float* W; int lda, nrg, N, K;
            if (item < I0) { l = 0; cg = item % cg0; rg = item / cg0; A = A0; lda = lda0; nrg = TN / 4; W = W0; N = out.m0; K = in.m0; }
            else if (item < I0 + I1) { const int t = item - I0; l = 1; cg = t % cg1; rg = t / cg1; A = A1; lda = lda1; nrg = 3 * TN / 4; W = W1; N = out.m1; K = in.m1; }
            else { const int t = item - I0 - I1; l = 2; cg = t % cg2; rg = t / cg2; A = A2; lda = lda2; nrg = 5 * TN / 4; W = W2; N = out.m2; K = in.m2; }
            const int k4 = K >> 2, lo = 4 * ((k4 * ks) / KS), hi = 4 * ((k4 * (ks + 1)) / KS);
            if (hi > lo) gemm_item_4x4<true, true>(A + lo, lda, nrg, rg, W + (size_t)lo * N, N, 4 * cg, hi - lo, acc);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = acc[i][j];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                acc[i][j] = v;
            }
        if (l < 0) continue;
        // every lane of the quad holds the sums: lane ks writes row-group member ks (4 of the 16 outputs each)
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ks == 0 ? acc[0][j] : ks == 1 ? acc[1][j] : ks == 2 ? acc[2][j] : acc[3][j];
        const int nrg = l == 0 ? TN / 4 : l == 1 ? 3 * TN / 4 : 5 * TN / 4;
        const int row = rg + ks * nrg;
#pragma unroll
        for (int j = 0; j < 4; ++j) epi(l, row % TN, row / TN, 4 * cg + j, o[j]);
    }
}

// SPLIT: the emb -> emb and mid -> emb GEMMs (60 items for the 240-dim irreps, K up to 192).  emb -> pre has 216 items of K <= 64:
// splitting it costs four rounds of shuffles + epilogues instead of one K loop (clock64 phase trace, profiles/run_chain_trace.py:
// 12.3 k cycles split against ~6 k whole).
template <int TN, bool SPLIT, typename Epi>
__device__ __forceinline__ void chain_gemm_any(const float* A0, int lda0, const float* A1, int lda1, const float* A2, int lda2,
                                               const float* W0, const float* W1, const float* W2, Irr in, Irr out, Epi epi) {
    if constexpr (TN == 4 && SPLIT) chain_gemm_splitk<TN, 4>(A0, lda0, A1, lda1, A2, lda2, W0, W1, W2, in, out, epi);
    else chain_gemm<TN>(A0, lda0, A1, lda1, A2, lda2, W0, W1, W2, in, out, epi);
}

template <int TN>
__global__ void __launch_bounds__(kChainThreads, 1) node_chain_kernel(ChainArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t w1bar, w2bar, wpbar, xbar, rbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    CHAIN_STAMP(0);
    const Irr emb = a.emb, pre = a.pre, mid = a.mid;
    const int F = emb.dim();
    const int lda0 = pad_lda(emb.m0), lda1 = pad_lda(emb.m1), lda2 = pad_lda(emb.m2);
    const int ldm0 = pad_lda(mid.m0), ldm1 = pad_lda(mid.m1), ldm2 = pad_lda(mid.m2);
    const int ldo0 = pre.m0 + 4;
    float* sW1 = smem + a.o_w1; float* sW2 = smem + a.o_w2; float* sWp = smem + a.o_wp;
    float* Y1 = smem + a.o_y1;                                   // [TN][F]
    float* X = smem + a.o_x; float* R = smem + a.o_res;          // [TN][F] each
    float* A0 = smem + a.o_a; float* A1 = A0 + TN * lda0; float* A2 = A1 + 3 * TN * lda1;
    float* O0 = smem + a.o_o0;                                   // [TN][ldo0]   fctp_1's 0e block (scalars | gates)
    float* M0 = smem + a.o_m; float* M1 = M0 + TN * ldm0; float* M2 = M1 + 3 * TN * ldm1;
    float* s_mean = smem + a.o_stat; float* s_scale = s_mean + TN;
    const int nP0 = emb.m0 * emb.m0, nP1 = emb.m1 * emb.m1, nP2 = emb.m2 * emb.m2;
    const int nA0 = emb.m0 * pre.m0, nA1 = emb.m1 * pre.m1, nA2 = emb.m2 * pre.m2;
    const int nB0 = mid.m0 * emb.m0, nB1 = mid.m1 * emb.m1, nB2 = mid.m2 * emb.m2;
    const int n_tiles = (a.n + TN - 1) / TN;
    const bool has_tile = (int)blockIdx.x < n_tiles;
    auto load_wp = [&]() {
        mbar_expect_tx(&wpbar, (uint32_t)(nP0 + nP1 + nP2) * 4u);
        bulk_g2s_chunked(sWp, a.P0, (uint32_t)nP0 * 4u, &wpbar);
        bulk_g2s_chunked(sWp + nP0, a.P1, (uint32_t)nP1 * 4u, &wpbar);
        bulk_g2s_chunked(sWp + nP0 + nP1, a.P2, (uint32_t)nP2 * 4u, &wpbar);
    };
    if (tid == 0) {
        mbar_init(&w1bar, 1); mbar_init(&w2bar, 1); mbar_init(&wpbar, 1); mbar_init(&xbar, 1); mbar_init(&rbar, 1);
        mbar_init_fence();
        if (has_tile) {
            load_wp();
            mbar_expect_tx(&w1bar, (uint32_t)(nA0 + nA1 + nA2) * 4u);
            bulk_g2s_chunked(sW1, a.A0, (uint32_t)nA0 * 4u, &w1bar);
            bulk_g2s_chunked(sW1 + nA0, a.A1, (uint32_t)nA1 * 4u, &w1bar);
            bulk_g2s_chunked(sW1 + nA0 + nA1, a.A2, (uint32_t)nA2 * 4u, &w1bar);
            mbar_expect_tx(&w2bar, (uint32_t)(nB0 + nB1 + nB2) * 4u);
            bulk_g2s_chunked(sW2, a.B0, (uint32_t)nB0 * 4u, &w2bar);
            bulk_g2s_chunked(sW2 + nB0, a.B1, (uint32_t)nB1 * 4u, &w2bar);
            bulk_g2s_chunked(sW2 + nB0 + nB1, a.B2, (uint32_t)nB2 * 4u, &w2bar);
        }
    }
    // biases and layer-norm parameters -> shared memory (parameters: read before the PDL wait; the epilogues and the LN pass then
    // take them at shared-memory latency instead of one L2 round trip per phase on the critical path of a one-tile CTA)
    float* s_pb = smem + a.o_par; float* s_ab = s_pb + emb.m0; float* s_bb = s_ab + pre.m0;
    float* s_lnw = s_bb + emb.m0; float* s_lnb = s_lnw + emb.nirr();
    if (has_tile && tid >= 32) {          // (warp 0 is busy issuing the bulk copies)
        constexpr int NT = kChainThreads - 32;
        for (int c = tid - 32; c < emb.m0; c += NT) { s_pb[c] = a.pb ? a.pb[c] : 0.f; s_bb[c] = a.bb ? a.bb[c] : 0.f; s_lnb[c] = a.ln_b[c]; }
        for (int c = tid - 32; c < pre.m0; c += NT) s_ab[c] = a.ab ? a.ab[c] : 0.f;
        for (int c = tid - 32; c < emb.nirr(); c += NT) s_lnw[c] = a.ln_w[c];
    }
    CHAIN_STAMP(1);
    pdl_wait(); pdl_launch();     // PDL: barrier init and the weight copies above overlap the previous kernel's tail
    CHAIN_STAMP(2);
    const int m0s = mid.m0;       // scalars that survive the gate
    uint32_t ph = 0, php = 0;
    bool first = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = tile * TN;
        const int rows = min(TN, a.n - n0);
        __syncthreads();          // barriers initialised / previous tile's output store has drained Y1, U region is free
        if (tid == 0) {
            if (!first && a.wp_transient) load_wp();
            mbar_expect_tx(&xbar, (uint32_t)(rows * F) * 4u);
            bulk_g2s_chunked(X, a.x + (size_t)n0 * F, (uint32_t)(rows * F) * 4u, &xbar);
            if (a.res1) {
                mbar_expect_tx(&rbar, (uint32_t)(rows * F) * 4u);
                bulk_g2s_chunked(R, a.res1 + (size_t)n0 * F, (uint32_t)(rows * F) * 4u, &rbar);
            }
        }
        mbar_wait(&xbar, ph);
        CHAIN_STAMP(3);
        // ---- A tiles of the attention output (l-split, K contiguous) ----
        for (int r = warp; r < TN; r += kChainThreads / 32) {
            const float* xr = X + r * F;
            const bool ok = r < rows;
            for (int c = lane; c < emb.m0; c += 32) A0[r * lda0 + c] = ok ? xr[c] : 0.f;
            for (int c = lane; c < 3 * emb.m1; c += 32) { const int u = c / 3, k = c - 3 * u; A1[(k * TN + r) * lda1 + u] = ok ? xr[emb.off1() + c] : 0.f; }
            for (int c = lane; c < 5 * emb.m2; c += 32) { const int u = c / 5, k = c - 5 * u; A2[(k * TN + r) * lda2 + u] = ok ? xr[emb.off2() + c] : 0.f; }
        }
        __syncthreads();
        CHAIN_STAMP(4);
        if (first || a.wp_transient) { mbar_wait(&wpbar, php); php ^= 1u; }
        if (a.res1) mbar_wait(&rbar, ph);
        CHAIN_STAMP(5);
        // ---- y1 = proj(x) + b (+ res1) ----
        {
            const bool has_res = a.res1 != nullptr;
            const int off1 = emb.off1(), off2 = emb.off2();
            chain_gemm_any<TN, true>(A0, lda0, A1, lda1, A2, lda2, sWp, sWp + nP0, sWp + nP0 + nP1, emb, emb,
                           [&](int l, int r, int k, int c, float v) {
                               const int col = (l == 0) ? c : (l == 1) ? off1 + 3 * c + k : off2 + 5 * c + k;
                               if (l == 0) v += s_pb[c];
                               if (has_res) v += R[r * F + col];
                               Y1[r * F + col] = v;
                           });
        }
        __syncthreads();
        CHAIN_STAMP(6);
        // ---- layer-norm statistics of y1: one warp per node ----
        for (int r = warp; r < TN; r += kChainThreads / 32) {
            float mean = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
            if (r < rows) {
                const float* xr = Y1 + r * F;
                float s = 0.f;
                for (int c = lane; c < emb.m0; c += 32) s += xr[c];
                s = warp_sum(s);
                mean = s / (float)emb.m0;
                float q0 = 0.f, q1 = 0.f, q2 = 0.f;
                for (int c = lane; c < emb.m0; c += 32) { const float t = xr[c] - mean; q0 += t * t; }
                for (int c = lane; c < 3 * emb.m1; c += 32) { const float t = xr[emb.off1() + c]; q1 += t * t; }
                for (int c = lane; c < 5 * emb.m2; c += 32) { const float t = xr[emb.off2() + c]; q2 += t * t; }
                q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
                sc0 = rsqrtf(q0 / (float)emb.m0 + a.ln_eps);
                sc1 = rsqrtf(q1 / (float)(3 * emb.m1) + a.ln_eps);
                sc2 = rsqrtf(q2 / (float)(5 * emb.m2) + a.ln_eps);
            }
            if (lane == 0) { s_mean[r] = mean; s_scale[r * 3] = sc0; s_scale[r * 3 + 1] = sc1; s_scale[r * 3 + 2] = sc2; }
        }
        __syncthreads();
        CHAIN_STAMP(7);
        // ---- A tiles of LN(y1) ----
        for (int r = warp; r < TN; r += kChainThreads / 32) {
            const float* xr = Y1 + r * F;
            const bool ok = r < rows;
            const float mean = s_mean[r], sc0 = s_scale[r * 3], sc1 = s_scale[r * 3 + 1], sc2 = s_scale[r * 3 + 2];
            for (int c = lane; c < emb.m0; c += 32) A0[r * lda0 + c] = ok ? (xr[c] - mean) * sc0 * s_lnw[c] + s_lnb[c] : 0.f;
            for (int c = lane; c < 3 * emb.m1; c += 32) {
                const int u = c / 3, k = c - 3 * u;
                A1[(k * TN + r) * lda1 + u] = ok ? xr[emb.off1() + c] * sc1 * s_lnw[emb.m0 + u] : 0.f;
            }
            for (int c = lane; c < 5 * emb.m2; c += 32) {
                const int u = c / 5, k = c - 5 * u;
                A2[(k * TN + r) * lda2 + u] = ok ? xr[emb.off2() + c] * sc2 * s_lnw[emb.m0 + emb.m1 + u] : 0.f;
            }
        }
        __syncthreads();          // (also: the proj GEMM is done with sWp / X / R -- the U region now belongs to the FFN)
        CHAIN_STAMP(8);
        if (first) mbar_wait(&w1bar, 0);
        CHAIN_STAMP(9);
        // ---- fctp_1: 0e block raw (+bias) into O0, l > 0 blocks raw into the next GEMM's A tiles ----
        {
            chain_gemm_any<TN, false>(A0, lda0, A1, lda1, A2, lda2, sW1, sW1 + nA0, sW1 + nA0 + nA1, emb, pre,
                           [&](int l, int r, int k, int c, float v) {
                               if (l == 0) O0[r * ldo0 + c] = v + s_ab[c];
                               else if (l == 1) M1[(k * TN + r) * ldm1 + c] = v;
                               else M2[(k * TN + r) * ldm2 + c] = v;
                           });
        }
        __syncthreads();
        CHAIN_STAMP(10);
        // ---- Gate: SiLU on the scalars, sigmoid gates multiplied onto the l > 0 channels (in place) ----
        constexpr int NW = kChainThreads / 32, SUB = TN < NW ? NW / TN : 1;       // few-node tiles: SUB warps share a node's columns
        for (int r = warp % (TN < NW ? TN : NW); r < TN; r += NW) {
            const float* o = O0 + r * ldo0;
            const int c0 = lane + 32 * (TN < NW ? warp / TN : 0);
            for (int c = c0; c < m0s; c += 32 * SUB) M0[r * ldm0 + c] = kCSilu * siluf_(o[c]);
            for (int c = c0; c < mid.m1; c += 32 * SUB) {
                const float g = kCSigmoid * sigmoidf_(o[m0s + c]);
#pragma unroll
                for (int k = 0; k < 3; ++k) M1[(k * TN + r) * ldm1 + c] *= g;
            }
            for (int c = c0; c < mid.m2; c += 32 * SUB) {
                const float g = kCSigmoid * sigmoidf_(o[m0s + mid.m1 + c]);
#pragma unroll
                for (int k = 0; k < 5; ++k) M2[(k * TN + r) * ldm2 + c] *= g;
            }
        }
        __syncthreads();
        CHAIN_STAMP(11);
        if (first) mbar_wait(&w2bar, 0);
        CHAIN_STAMP(12);
        // ---- y = y1 + fctp_2(h) + b ----
        {
            const int off1 = emb.off1(), off2 = emb.off2();
            chain_gemm_any<TN, true>(M0, ldm0, M1, ldm1, M2, ldm2, sW2, sW2 + nB0, sW2 + nB0 + nB1, mid, emb,
                           [&](int l, int r, int k, int c, float v) {
                               const int col = (l == 0) ? c : (l == 1) ? off1 + 3 * c + k : off2 + 5 * c + k;
                               if (l == 0) v += s_bb[c];
                               Y1[r * F + col] = v + Y1[r * F + col];
                           });
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        CHAIN_STAMP(13);
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(rows * F) * 4u;
            constexpr uint32_t kChunk = 32768;
            for (uint32_t off = 0; off < bytes; off += kChunk)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(reinterpret_cast<char*>(a.y + (size_t)n0 * F) + off), "r"(smem_u32(Y1) + off), "r"(min(kChunk, bytes - off)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        CHAIN_STAMP(14);
        ph ^= 1u;
        first = false;
    }
    // a CTA without a tile issued no copies; one with tiles has waited for all of them (w1 / w2 / wp on its first tile)
}

template <int TN>
static int launch_chain(ChainArgs a, cudaStream_t stream) {
    const Irr emb = a.emb, pre = a.pre, mid = a.mid;
    const int F = emb.dim();
    auto r4 = [](int v) { return (v + 3) & ~3; };
    const int nP = emb.m0 * emb.m0 + emb.m1 * emb.m1 + emb.m2 * emb.m2;
    const int nA = emb.m0 * pre.m0 + emb.m1 * pre.m1 + emb.m2 * pre.m2;
    const int nB = mid.m0 * emb.m0 + mid.m1 * emb.m1 + mid.m2 * emb.m2;
    const int a_tiles = TN * pad_lda(emb.m0) + 3 * TN * pad_lda(emb.m1) + 5 * TN * pad_lda(emb.m2);
    const int m_tiles = TN * pad_lda(mid.m0) + 3 * TN * pad_lda(mid.m1) + 5 * TN * pad_lda(mid.m2);
    const int o0 = TN * (pre.m0 + 4);
    const int proj_u = r4(nP) + 2 * TN * F, ffn_u = o0 + m_tiles;
    constexpr int kMax = 226 * 1024 / 4;
    int off = 0;
    a.o_w1 = off; off += r4(nA);
    a.o_w2 = off; off += r4(nB);
    a.o_y1 = off; off += TN * F;
    a.o_a = off; off += a_tiles;
    a.o_stat = off; off += 4 * TN;
    a.o_par = off; off += r4(2 * emb.m0 + pre.m0 + emb.nirr() + emb.m0);
    a.o_u = off;
    if (off + proj_u + ffn_u <= kMax) {            // everything resident
        a.wp_transient = 0;
        a.o_wp = off; a.o_x = off + r4(nP); a.o_res = a.o_x + TN * F;
        a.o_o0 = off + proj_u; a.o_m = a.o_o0 + o0;
        off += proj_u + ffn_u;
    } else {                                       // proj phase and FFN phases share the U region
        a.wp_transient = 1;
        a.o_wp = off; a.o_x = off + r4(nP); a.o_res = a.o_x + TN * F;
        a.o_o0 = off; a.o_m = off + o0;
        off += proj_u > ffn_u ? proj_u : ffn_u;
    }
    if (off > kMax) return DEDF_ERR_UNSUPPORTED;
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(node_chain_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); done = true; }
    const int n_tiles = (a.n + TN - 1) / TN;
    launch_pdl(node_chain_kernel<TN>, dim3(grid_for(n_tiles, 1, kNumSMs)), dim3(kChainThreads), (size_t)off * sizeof(float), stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_node_chain(const dedf_node_chain_desc* d, cudaStream_t stream) {
    if (!d || !d->x || !d->y) return DEDF_ERR_ARG;
    ChainArgs a{};
    a.x = d->x; a.n = d->n;
    a.emb = Irr{d->irr_emb[0], d->irr_emb[1], d->irr_emb[2]};
    a.pre = Irr{d->irr_pre[0], d->irr_pre[1], d->irr_pre[2]};
    a.mid = Irr{a.pre.m0 - a.pre.m1 - a.pre.m2, a.pre.m1, a.pre.m2};
    a.P0 = d->P0; a.P1 = d->P1; a.P2 = d->P2; a.pb = d->pb; a.res1 = d->res1;
    a.ln_w = d->ln_w; a.ln_b = d->ln_b; a.ln_eps = d->ln_eps;
    a.A0 = d->A0; a.A1 = d->A1; a.A2 = d->A2; a.ab = d->ab;
    a.B0 = d->B0; a.B1 = d->B1; a.B2 = d->B2; a.bb = d->bb;
    a.y = d->y;
    if (!a.P0 || !a.P1 || !a.P2 || !a.A0 || !a.A1 || !a.A2 || !a.B0 || !a.B1 || !a.B2 || !a.ln_w || !a.ln_b) return DEDF_ERR_ARG;
    if (a.mid.m0 <= 0) return DEDF_ERR_ARG;
    const int ms[9] = {a.emb.m0, a.emb.m1, a.emb.m2, a.pre.m0, a.pre.m1, a.pre.m2, a.mid.m0, a.mid.m1, a.mid.m2};
    for (int i = 0; i < 9; ++i) if (ms[i] <= 0 || (ms[i] & 3)) return DEDF_ERR_UNSUPPORTED;       // float4 weight rows, K % 4 == 0
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (!al16(a.x) || !al16(a.y) || !al16(a.res1) || !al16(a.P0) || !al16(a.P1) || !al16(a.P2) || !al16(a.A0) || !al16(a.A1) || !al16(a.A2) ||
        !al16(a.B0) || !al16(a.B1) || !al16(a.B2)) return DEDF_ERR_UNSUPPORTED;
    if (d->n <= 0) return DEDF_OK;
    // Few nodes (the denoise step at 128 poses has 256): 4-node tiles put them on twice as many SMs and each GEMM phase is one
    // round of work items instead of two -- the launch is latency, not throughput.  Otherwise 16-node tiles when the three
    // weight sets and the tiles fit, 8-node tiles for the wide irreps.
    int rc = DEDF_ERR_UNSUPPORTED;
    if (d->n <= 2 * kNumSMs && !getenv("DEDF_CHAIN_NO_TN4")) rc = launch_chain<4>(a, stream);
    if (rc == DEDF_ERR_UNSUPPORTED) rc = launch_chain<16>(a, stream);
    if (rc == DEDF_ERR_UNSUPPORTED) rc = launch_chain<8>(a, stream);
    return rc;
}

#ifdef DEDF_CHAIN_TRACE
extern "C" int dedf_chain_trace(long long* host_out32) {
    return cudaMemcpyFromSymbol(host_out32, dedf::g_chain_trace, sizeof(long long) * 32) == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}
#endif

// Two independent dedf_node_linear problems without layer norm / gate / residual in ONE launch (blockIdx.y selects): the
// linear_src / linear_dst pair at the head of every UNet block (block.py:149-153).
namespace dedf {
struct DualLinArgs {
    const float* x[2]; int n[2]; Irr in[2], out;
    const float* W0[2]; const float* W1[2]; const float* W2[2]; const float* b0[2];
    float* y[2];
};

template <int TN>
__global__ void __launch_bounds__(kChainThreads) dual_linear_kernel(DualLinArgs a) {
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t wbar, xbar;
    const int p = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const Irr in = a.in[p], out = a.out;
    const int Fin = in.dim(), Fout = out.dim();
    const int lda0 = pad_lda(in.m0), lda1 = pad_lda(in.m1), lda2 = pad_lda(in.m2);
    const int n0w = in.m0 * out.m0, n1w = in.m1 * out.m1, n2w = in.m2 * out.m2;
    float* A0 = smem; float* A1 = A0 + TN * lda0; float* A2 = A1 + 3 * TN * lda1;
    float* IO = A2 + 5 * TN * lda2;                  // [TN][max(Fin, Fout)]
    float* sW = IO + TN * (Fin > Fout ? Fin : Fout);
    const int n_tiles = (a.n[p] + TN - 1) / TN;
    if (tid == 0) {
        mbar_init(&wbar, 1); mbar_init(&xbar, 1);
        mbar_init_fence();
        if ((int)blockIdx.x < n_tiles) {
            mbar_expect_tx(&wbar, (uint32_t)(n0w + n1w + n2w) * 4u);
            if (n0w) bulk_g2s_chunked(sW, a.W0[p], (uint32_t)n0w * 4u, &wbar);
            if (n1w) bulk_g2s_chunked(sW + n0w, a.W1[p], (uint32_t)n1w * 4u, &wbar);
            if (n2w) bulk_g2s_chunked(sW + n0w + n1w, a.W2[p], (uint32_t)n2w * 4u, &wbar);
        }
    }
    pdl_wait(); pdl_launch();
    uint32_t ph = 0; bool first = true;
    const float* b0 = a.b0[p];
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = tile * TN, rows = min(TN, a.n[p] - n0);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&xbar, (uint32_t)(rows * Fin) * 4u);
            bulk_g2s_chunked(IO, a.x[p] + (size_t)n0 * Fin, (uint32_t)(rows * Fin) * 4u, &xbar);
        }
        mbar_wait(&xbar, ph);
        for (int r = warp; r < TN; r += kChainThreads / 32) {
            const float* xr = IO + r * Fin;
            const bool ok = r < rows;
            for (int c = lane; c < in.m0; c += 32) A0[r * lda0 + c] = ok ? xr[c] : 0.f;
            for (int c = lane; c < 3 * in.m1; c += 32) { const int u = c / 3, k = c - 3 * u; A1[(k * TN + r) * lda1 + u] = ok ? xr[in.off1() + c] : 0.f; }
            for (int c = lane; c < 5 * in.m2; c += 32) { const int u = c / 5, k = c - 5 * u; A2[(k * TN + r) * lda2 + u] = ok ? xr[in.off2() + c] : 0.f; }
        }
        __syncthreads();
        if (first) { mbar_wait(&wbar, 0); first = false; }
        const int off1 = out.off1(), off2 = out.off2();
        chain_gemm<TN>(A0, lda0, A1, lda1, A2, lda2, sW, sW + n0w, sW + n0w + n1w, in, out,
                       [&](int l, int r, int k, int c, float v) {
                           const int col = (l == 0) ? c : (l == 1) ? off1 + 3 * c + k : off2 + 5 * c + k;
                           if (l == 0 && b0) v += b0[c];
                           IO[r * Fout + col] = v;
                       });
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(rows * Fout) * 4u;
            constexpr uint32_t kChunk = 32768;
            for (uint32_t off = 0; off < bytes; off += kChunk)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(reinterpret_cast<char*>(a.y[p] + (size_t)n0 * Fout) + off), "r"(smem_u32(IO) + off), "r"(min(kChunk, bytes - off)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        ph ^= 1u;
    }
}
}  // namespace dedf

extern "C" int dedf_node_linear_pair(const float* x_a, int n_a, const int* irr_in_a, const float* const* W_a, const float* bias_a, float* y_a,
                                     const float* x_b, int n_b, const int* irr_in_b, const float* const* W_b, const float* bias_b, float* y_b,
                                     const int* irr_out, cudaStream_t stream) {
    if (!x_a || !x_b || !y_a || !y_b || !irr_in_a || !irr_in_b || !irr_out || !W_a || !W_b) return DEDF_ERR_ARG;
    DualLinArgs a{};
    a.x[0] = x_a; a.x[1] = x_b; a.n[0] = n_a; a.n[1] = n_b; a.y[0] = y_a; a.y[1] = y_b;
    a.in[0] = Irr{irr_in_a[0], irr_in_a[1], irr_in_a[2]}; a.in[1] = Irr{irr_in_b[0], irr_in_b[1], irr_in_b[2]};
    a.out = Irr{irr_out[0], irr_out[1], irr_out[2]};
    a.b0[0] = bias_a; a.b0[1] = bias_b;
    auto al16 = [](const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    size_t w_floats = 0, tile_floats = 0;
    constexpr int TN = 16;
    for (int p = 0; p < 2; ++p) {
        const float* const* W = p ? W_b : W_a;
        a.W0[p] = W[0]; a.W1[p] = W[1]; a.W2[p] = W[2];
        const Irr in = a.in[p];
        const int ms[6] = {in.m0, in.m1, in.m2, a.out.m0, a.out.m1, a.out.m2};
        for (int i = 0; i < 6; ++i) if (ms[i] <= 0 || (ms[i] & 3)) return DEDF_ERR_UNSUPPORTED;
        if (!W[0] || !W[1] || !W[2]) return DEDF_ERR_ARG;
        if (!al16(W[0]) || !al16(W[1]) || !al16(W[2])) return DEDF_ERR_UNSUPPORTED;
        const size_t wf = (size_t)in.m0 * a.out.m0 + (size_t)in.m1 * a.out.m1 + (size_t)in.m2 * a.out.m2;
        const int Fin = in.dim(), Fout = a.out.dim();
        const size_t tf = (size_t)TN * pad_lda(in.m0) + 3 * TN * pad_lda(in.m1) + 5 * TN * pad_lda(in.m2) + (size_t)TN * (Fin > Fout ? Fin : Fout);
        w_floats = wf > w_floats ? wf : w_floats; tile_floats = tf > tile_floats ? tf : tile_floats;
    }
    if (!al16(x_a) || !al16(x_b) || !al16(y_a) || !al16(y_b)) return DEDF_ERR_UNSUPPORTED;
    if (n_a <= 0 && n_b <= 0) return DEDF_OK;
    const size_t smem = (w_floats + tile_floats) * sizeof(float);
    if (smem > 226 * 1024) return DEDF_ERR_UNSUPPORTED;
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(dual_linear_kernel<TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); done = true; }
    const int nt = ((n_a > n_b ? n_a : n_b) + TN - 1) / TN;
    launch_pdl(dual_linear_kernel<TN>, dim3(grid_for(nt, 1, kNumSMs), 2), dim3(kChainThreads), smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
