// Per-node kernels: equivariant layer norm -> block-diagonal linear (LinearRS /
// FullyConnectedTensorProductRescale with a 1x0e second operand) -> gate / residual.
//
// Replaces, on the reference's hot path (/root/reference/diffusion_edf):
//   equiformer/layer_norm.py:91-156        EquivariantLayerNormV2 ('component', affine)
//   equiformer/tensor_product_rescale.py:176-185, :155-173, :241-268   LinearRS / FCTP (+SwishGate)
//   equiformer/fast_activation.py:210-224  Gate
//   skip.py:13-35                          ProjectIfMismatch
//   gnn_block.py:51-57, block.py:51-57     FeedForwardNetwork
// LinearRS semantics: per l,  y[w, m] = sum_u W_l[u, w] x[u, m];
// bias only on the 0e block; nothing is rescaled at run time.
#include "common.cuh"
#include "gemm_tile.cuh"
#include "../../include/dedf.h"

namespace dedf {

constexpr int kNodeTN = 16;        // nodes per tile
constexpr int kNodeThreads = 256;

struct NodeLinArgs {
    const float* x; int n;                 // (n, Fin)
    Irr in, out;                           // irreps of x and of the linear output (pre-gate)
    const float* W0; const float* W1; const float* W2;   // (in.m_l, out.m_l) row-major; null if either mul is 0
    const float* bias0;                    // (out.m0) or null
    // optional layer-norm prologue
    const float* ln_w; const float* ln_b; float ln_eps; int ln;
    // epilogue
    int gate;                              // 1: out.m0 = scalars + gates, gates = out.m1 + out.m2
    const float* res; float res_scale;     // y = (y + res) * res_scale   (res may be null)
    float* y;                              // (n, Fy)
    int bulk_w;                            // weights are 16-byte aligned with sizes % 16 == 0: stage them with TMA bulk copies
};

template <bool kVecW, bool kWShared>
__global__ void __launch_bounds__(kNodeThreads) node_linear_kernel(NodeLinArgs a, int lda0, int lda1, int lda2, int ldo) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TN = kNodeTN;
    const int Fin = a.in.dim(), Fout = a.out.dim();
    float* A0 = smem;                       // [TN][lda0]
    float* A1 = A0 + TN * lda0;             // [3 TN][lda1]  row = k * TN + n
    float* A2 = A1 + 3 * TN * lda1;         // [5 TN][lda2]
    float* O = A2 + 5 * TN * lda2;          // [TN][ldo]  linear output in e3nn layout
    float* s_scale = O + TN * ldo;          // [TN][3]    layer-norm scale per l
    float* s_mean = s_scale + TN * 3;       // [TN]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Stage the weight matrices in shared memory once per CTA (they are reused by every tile): the K loop then runs at
    // shared-memory latency instead of L2 latency, which is what bounds the small (few-tile) launches of the coarse scales.
    const float* W0 = a.W0; const float* W1 = a.W1; const float* W2 = a.W2;
    __shared__ __align__(8) uint64_t wbar;
    bool w_pending = false;
    if (kWShared) {
        float* sW = s_mean + ((TN + 3) & ~3);
        const int n0w = a.W0 ? a.in.m0 * a.out.m0 : 0, n1w = a.W1 ? a.in.m1 * a.out.m1 : 0, n2w = a.W2 ? a.in.m2 * a.out.m2 : 0;
        const int o1 = (n0w + 3) & ~3, o2 = o1 + ((n1w + 3) & ~3);
        if (a.bulk_w) {
            // TMA: one thread issues 1-D bulk copies of the three matrices; they land while the layer-norm statistics
            // and the A tiles of the first tile are being prepared (waited for right before the GEMM)
            if (tid == 0) {
                mbar_init(&wbar, 1);
                mbar_init_fence();
                mbar_expect_tx(&wbar, (uint32_t)(n0w + n1w + n2w) * 4u);
                if (n0w) bulk_g2s_chunked(sW, a.W0, (uint32_t)n0w * 4u, &wbar);
                if (n1w) bulk_g2s_chunked(sW + o1, a.W1, (uint32_t)n1w * 4u, &wbar);
                if (n2w) bulk_g2s_chunked(sW + o2, a.W2, (uint32_t)n2w * 4u, &wbar);
            }
            w_pending = true;
        } else {
            for (int i = tid; i < n0w; i += kNodeThreads) sW[i] = __ldg(a.W0 + i);
            for (int i = tid; i < n1w; i += kNodeThreads) sW[o1 + i] = __ldg(a.W1 + i);
            for (int i = tid; i < n2w; i += kNodeThreads) sW[o2 + i] = __ldg(a.W2 + i);
        }
        if (a.W0) W0 = sW;
        if (a.W1) W1 = sW + o1;
        if (a.W2) W2 = sW + o2;
    }
    pdl_wait(); pdl_launch();     // PDL: the weight copies above overlap the previous kernel's tail
    const int n_tiles = (a.n + TN - 1) / TN;
    const int Fy = a.gate ? (Fout - a.out.m1 - a.out.m2) : Fout;
    const int m0s = a.gate ? (a.out.m0 - a.out.m1 - a.out.m2) : a.out.m0;   // scalars that survive the gate

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = tile * TN;
        const int rows = min(TN, a.n - n0);
        __syncthreads();
        // ---- layer-norm statistics: one warp per node --------------------------------
        if (a.ln) {
            for (int r = warp; r < TN; r += kNodeThreads / 32) {
                float mean = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
                if (r < rows) {
                    const float* xr = a.x + (size_t)(n0 + r) * Fin;
                    float s = 0.f;
                    for (int c = lane; c < a.in.m0; c += 32) s += xr[c];
                    s = warp_sum(s);
                    mean = (a.in.m0 > 0) ? s / (float)a.in.m0 : 0.f;
                    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
                    for (int c = lane; c < a.in.m0; c += 32) { const float t = xr[c] - mean; q0 += t * t; }
                    for (int c = lane; c < 3 * a.in.m1; c += 32) { const float t = xr[a.in.off1() + c]; q1 += t * t; }
                    for (int c = lane; c < 5 * a.in.m2; c += 32) { const float t = xr[a.in.off2() + c]; q2 += t * t; }
                    q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
                    // field.pow(2).mean(-1) then mean over mul  == sum / (mul * d)
                    sc0 = (a.in.m0 > 0) ? rsqrtf(q0 / (float)a.in.m0 + a.ln_eps) : 0.f;
                    sc1 = (a.in.m1 > 0) ? rsqrtf(q1 / (float)(3 * a.in.m1) + a.ln_eps) : 0.f;
                    sc2 = (a.in.m2 > 0) ? rsqrtf(q2 / (float)(5 * a.in.m2) + a.ln_eps) : 0.f;
                }
                if (lane == 0) { s_mean[r] = mean; s_scale[r * 3] = sc0; s_scale[r * 3 + 1] = sc1; s_scale[r * 3 + 2] = sc2; }
            }
            __syncthreads();
        }
        // ---- stage A tiles (normalised on the fly) -----------------------------------
        for (int i = tid; i < TN * Fin; i += kNodeThreads) {
            const int r = i / Fin, c = i % Fin;
            float v = (r < rows) ? a.x[(size_t)(n0 + r) * Fin + c] : 0.f;
            if (c < a.in.m0) {
                if (a.ln) v = (v - s_mean[r]) * s_scale[r * 3] * a.ln_w[c] + a.ln_b[c];
                A0[r * lda0 + c] = v;
            } else if (c < a.in.off2()) {
                const int u = (c - a.in.m0) / 3, k = (c - a.in.m0) % 3;
                if (a.ln) v = v * s_scale[r * 3 + 1] * a.ln_w[a.in.m0 + u];
                A1[(k * TN + r) * lda1 + u] = v;
            } else {
                const int u = (c - a.in.off2()) / 5, k = (c - a.in.off2()) % 5;
                if (a.ln) v = v * s_scale[r * 3 + 2] * a.ln_w[a.in.m0 + a.in.m1 + u];
                A2[(k * TN + r) * lda2 + u] = v;
            }
        }
        __syncthreads();
        if (w_pending) { mbar_wait(&wbar, 0); w_pending = false; }     // weights have landed (first tile only)
        // ---- block-diagonal GEMM -----------------------------------------------------
        const int cg0 = (a.out.m0 + 3) / 4, cg1 = (a.out.m1 + 3) / 4, cg2 = (a.out.m2 + 3) / 4;
        const int I0 = (a.W0 ? (TN / 4) * cg0 : 0), I1 = (a.W1 ? (3 * TN / 4) * cg1 : 0), I2 = (a.W2 ? (5 * TN / 4) * cg2 : 0);
        // outputs without a path are zero (e.g. l>0 of the 3x0e input embedding)
        for (int i = tid; i < TN * Fout; i += kNodeThreads) {
            const int c = i % Fout;
            const bool has = (c < a.out.m0) ? (a.W0 != nullptr) : (c < a.out.off2()) ? (a.W1 != nullptr) : (a.W2 != nullptr);
            if (!has) O[(i / Fout) * ldo + c] = 0.f;
        }
        for (int item = tid; item < I0 + I1 + I2; item += kNodeThreads) {
            float acc[4][4] = {};
            if (item < I0) {
                const int cg = item % cg0, rg = item / cg0;
                gemm_item_4x4<kVecW, kWShared>(A0, lda0, TN / 4, rg, W0, a.out.m0, 4 * cg, a.in.m0, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, r = rg + i * (TN / 4);
                        if (c < a.out.m0) O[r * ldo + c] = acc[i][j] + (a.bias0 ? a.bias0[c] : 0.f);
                    }
            } else if (item < I0 + I1) {
                const int t = item - I0, cg = t % cg1, rg = t / cg1;
                gemm_item_4x4<kVecW, kWShared>(A1, lda1, 3 * TN / 4, rg, W1, a.out.m1, 4 * cg, a.in.m1, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, row = rg + i * (3 * TN / 4), k = row / TN, r = row % TN;
                        if (c < a.out.m1) O[r * ldo + a.out.off1() + c * 3 + k] = acc[i][j];
                    }
            } else {
                const int t = item - I0 - I1, cg = t % cg2, rg = t / cg2;
                gemm_item_4x4<kVecW, kWShared>(A2, lda2, 5 * TN / 4, rg, W2, a.out.m2, 4 * cg, a.in.m2, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, row = rg + i * (5 * TN / 4), k = row / TN, r = row % TN;
                        if (c < a.out.m2) O[r * ldo + a.out.off2() + c * 5 + k] = acc[i][j];
                    }
            }
        }
        __syncthreads();
        // ---- epilogue: gate / residual, coalesced store --------------------------------
        for (int i = tid; i < rows * Fy; i += kNodeThreads) {
            const int r = i / Fy, c = i % Fy;
            const float* o = O + r * ldo;
            float v;
            if (!a.gate) {
                v = o[c];
            } else if (c < m0s) {
                v = kCSilu * siluf_(o[c]);
            } else if (c < m0s + 3 * a.out.m1) {
                const int u = (c - m0s) / 3;
                v = o[a.out.off1() + (c - m0s)] * (kCSigmoid * sigmoidf_(o[m0s + u]));
            } else {
                const int u = (c - m0s - 3 * a.out.m1) / 5;
                v = o[a.out.off2() + (c - m0s - 3 * a.out.m1)] * (kCSigmoid * sigmoidf_(o[m0s + a.out.m1 + u]));
            }
            if (a.res) v = (v + a.res[(size_t)(n0 + r) * Fy + c]) * a.res_scale;
            a.y[(size_t)(n0 + r) * Fy + c] = v;
        }
    }
}

// ---------------------------------------------------------------------------
// TMA variant (the default): every global access is a 1-D bulk async copy.  The weights (before the PDL wait: they do not
// depend on the previous kernel), the x tile (16 rows are one contiguous range), the residual tile and the output tile move
// as cp.async.bulk transfers signalled through mbarriers; the arithmetic only touches shared memory.  Same math and
// summation order as node_linear_kernel above (which stays as the fallback for unaligned / odd-width operands).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kNodeThreads) node_linear_tma_kernel(NodeLinArgs a, int lda0, int lda1, int lda2, int ldo, int io_floats) {
    extern __shared__ __align__(16) float smem[];
    constexpr int TN = kNodeTN;
    const int Fin = a.in.dim(), Fout = a.out.dim();
    float* A0 = smem;
    float* A1 = A0 + TN * lda0;
    float* A2 = A1 + 3 * TN * lda1;
    float* O = A2 + 5 * TN * lda2;
    float* s_scale = O + TN * ldo;
    float* s_mean = s_scale + TN * 3;
    float* io = s_mean + ((TN + 3) & ~3);          // [io_floats]: x tile, then residual tile, then the output tile
    float* sW = io + io_floats;
    __shared__ __align__(8) uint64_t wbar, xbar, rbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n0w = a.W0 ? a.in.m0 * a.out.m0 : 0, n1w = a.W1 ? a.in.m1 * a.out.m1 : 0, n2w = a.W2 ? a.in.m2 * a.out.m2 : 0;
    const int o1 = (n0w + 3) & ~3, o2 = o1 + ((n1w + 3) & ~3);
    if (tid == 0) {
        mbar_init(&wbar, 1); mbar_init(&xbar, 1); mbar_init(&rbar, 1);
        mbar_init_fence();
        mbar_expect_tx(&wbar, (uint32_t)(n0w + n1w + n2w) * 4u);
        if (n0w) bulk_g2s_chunked(sW, a.W0, (uint32_t)n0w * 4u, &wbar);
        if (n1w) bulk_g2s_chunked(sW + o1, a.W1, (uint32_t)n1w * 4u, &wbar);
        if (n2w) bulk_g2s_chunked(sW + o2, a.W2, (uint32_t)n2w * 4u, &wbar);
    }
    const float* W0 = a.W0 ? sW : nullptr;
    const float* W1 = a.W1 ? sW + o1 : nullptr;
    const float* W2 = a.W2 ? sW + o2 : nullptr;
    pdl_wait(); pdl_launch();     // PDL: barrier init and the weight copies above overlap the previous kernel's tail
    const int n_tiles = (a.n + TN - 1) / TN;
    const int Fy = a.gate ? (Fout - a.out.m1 - a.out.m2) : Fout;
    const int m0s = a.gate ? (a.out.m0 - a.out.m1 - a.out.m2) : a.out.m0;
    uint32_t ph = 0;
    bool w_pending = true;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = tile * TN;
        const int rows = min(TN, a.n - n0);
        __syncthreads();          // barriers initialised / previous tile's output store has drained the io tile
        if (tid == 0) {
            mbar_expect_tx(&xbar, (uint32_t)(rows * Fin) * 4u);
            bulk_g2s_chunked(io, a.x + (size_t)n0 * Fin, (uint32_t)(rows * Fin) * 4u, &xbar);
        }
        mbar_wait(&xbar, ph);
        // ---- layer-norm statistics: one warp per node (from shared memory) ----
        if (a.ln) {
            for (int r = warp; r < TN; r += kNodeThreads / 32) {
                float mean = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
                if (r < rows) {
                    const float* xr = io + r * Fin;
                    float s = 0.f;
                    for (int c = lane; c < a.in.m0; c += 32) s += xr[c];
                    s = warp_sum(s);
                    mean = (a.in.m0 > 0) ? s / (float)a.in.m0 : 0.f;
                    float q0 = 0.f, q1 = 0.f, q2 = 0.f;
                    for (int c = lane; c < a.in.m0; c += 32) { const float t = xr[c] - mean; q0 += t * t; }
                    for (int c = lane; c < 3 * a.in.m1; c += 32) { const float t = xr[a.in.off1() + c]; q1 += t * t; }
                    for (int c = lane; c < 5 * a.in.m2; c += 32) { const float t = xr[a.in.off2() + c]; q2 += t * t; }
                    q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
                    sc0 = (a.in.m0 > 0) ? rsqrtf(q0 / (float)a.in.m0 + a.ln_eps) : 0.f;
                    sc1 = (a.in.m1 > 0) ? rsqrtf(q1 / (float)(3 * a.in.m1) + a.ln_eps) : 0.f;
                    sc2 = (a.in.m2 > 0) ? rsqrtf(q2 / (float)(5 * a.in.m2) + a.ln_eps) : 0.f;
                }
                if (lane == 0) { s_mean[r] = mean; s_scale[r * 3] = sc0; s_scale[r * 3 + 1] = sc1; s_scale[r * 3 + 2] = sc2; }
            }
            __syncthreads();
        }
        // ---- stage A tiles (normalised on the fly): warp per row, lanes over columns ----
        for (int r = warp; r < TN; r += kNodeThreads / 32) {
            const float* xr = io + r * Fin;
            const bool ok = r < rows;
            const float mean = a.ln ? s_mean[r] : 0.f;
            const float sc0 = a.ln ? s_scale[r * 3] : 1.f, sc1 = a.ln ? s_scale[r * 3 + 1] : 1.f, sc2 = a.ln ? s_scale[r * 3 + 2] : 1.f;
            for (int c = lane; c < a.in.m0; c += 32) {
                float v = ok ? xr[c] : 0.f;
                if (a.ln) v = (v - mean) * sc0 * a.ln_w[c] + a.ln_b[c];
                A0[r * lda0 + c] = v;
            }
            for (int c = lane; c < 3 * a.in.m1; c += 32) {
                const int u = c / 3, k = c - 3 * u;
                float v = ok ? xr[a.in.off1() + c] : 0.f;
                if (a.ln) v = v * sc1 * a.ln_w[a.in.m0 + u];
                A1[(k * TN + r) * lda1 + u] = v;
            }
            for (int c = lane; c < 5 * a.in.m2; c += 32) {
                const int u = c / 5, k = c - 5 * u;
                float v = ok ? xr[a.in.off2() + c] : 0.f;
                if (a.ln) v = v * sc2 * a.ln_w[a.in.m0 + a.in.m1 + u];
                A2[(k * TN + r) * lda2 + u] = v;
            }
        }
        __syncthreads();          // A staged: the io tile is free for the residual rows
        if (a.res && tid == 0) {
            mbar_expect_tx(&rbar, (uint32_t)(rows * Fy) * 4u);
            bulk_g2s_chunked(io, a.res + (size_t)n0 * Fy, (uint32_t)(rows * Fy) * 4u, &rbar);
        }
        if (w_pending) { mbar_wait(&wbar, 0); w_pending = false; }
        // ---- block-diagonal GEMM (identical to node_linear_kernel) ----
        const int cg0 = (a.out.m0 + 3) / 4, cg1 = (a.out.m1 + 3) / 4, cg2 = (a.out.m2 + 3) / 4;
        const int I0 = (a.W0 ? (TN / 4) * cg0 : 0), I1 = (a.W1 ? (3 * TN / 4) * cg1 : 0), I2 = (a.W2 ? (5 * TN / 4) * cg2 : 0);
        if (!a.W0 || !a.W1 || !a.W2) {
            for (int r = warp; r < TN; r += kNodeThreads / 32)
                for (int c = lane; c < Fout; c += 32) {
                    const bool has = (c < a.out.m0) ? (a.W0 != nullptr) : (c < a.out.off2()) ? (a.W1 != nullptr) : (a.W2 != nullptr);
                    if (!has) O[r * ldo + c] = 0.f;
                }
        }
        for (int item = tid; item < I0 + I1 + I2; item += kNodeThreads) {
            float acc[4][4] = {};
            if (item < I0) {
                const int cg = item % cg0, rg = item / cg0;
                gemm_item_4x4<true, true>(A0, lda0, TN / 4, rg, W0, a.out.m0, 4 * cg, a.in.m0, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, r = rg + i * (TN / 4);
                        if (c < a.out.m0) O[r * ldo + c] = acc[i][j] + (a.bias0 ? a.bias0[c] : 0.f);
                    }
            } else if (item < I0 + I1) {
                const int t = item - I0, cg = t % cg1, rg = t / cg1;
                gemm_item_4x4<true, true>(A1, lda1, 3 * TN / 4, rg, W1, a.out.m1, 4 * cg, a.in.m1, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, row = rg + i * (3 * TN / 4), k = row / TN, r = row % TN;
                        if (c < a.out.m1) O[r * ldo + a.out.off1() + c * 3 + k] = acc[i][j];
                    }
            } else {
                const int t = item - I0 - I1, cg = t % cg2, rg = t / cg2;
                gemm_item_4x4<true, true>(A2, lda2, 5 * TN / 4, rg, W2, a.out.m2, 4 * cg, a.in.m2, acc);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = 4 * cg + j, row = rg + i * (5 * TN / 4), k = row / TN, r = row % TN;
                        if (c < a.out.m2) O[r * ldo + a.out.off2() + c * 5 + k] = acc[i][j];
                    }
            }
        }
        __syncthreads();
        if (a.res) mbar_wait(&rbar, ph);
        // ---- epilogue: gate / residual, in place into the io tile; one bulk store of the tile ----
        for (int r = warp; r < rows; r += kNodeThreads / 32) {
            const float* o = O + r * ldo;
            float* yr = io + r * Fy;
            for (int c = lane; c < Fy; c += 32) {
                float v;
                if (!a.gate) {
                    v = o[c];
                } else if (c < m0s) {
                    v = kCSilu * siluf_(o[c]);
                } else if (c < m0s + 3 * a.out.m1) {
                    const int u = (c - m0s) / 3;
                    v = o[a.out.off1() + (c - m0s)] * (kCSigmoid * sigmoidf_(o[m0s + u]));
                } else {
                    const int u = (c - m0s - 3 * a.out.m1) / 5;
                    v = o[a.out.off2() + (c - m0s - 3 * a.out.m1)] * (kCSigmoid * sigmoidf_(o[m0s + a.out.m1 + u]));
                }
                if (a.res) v = (v + yr[c]) * a.res_scale;
                yr[c] = v;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(rows * Fy) * 4u;
            constexpr uint32_t kChunk = 32768;
            for (uint32_t off = 0; off < bytes; off += kChunk)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(reinterpret_cast<char*>(a.y + (size_t)n0 * Fy) + off), "r"(smem_u32(io) + off), "r"(min(kChunk, bytes - off)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        ph ^= 1u;
    }
}

// y[i, :] = x[idx[i], :]
__global__ void gather_rows_kernel(const float* __restrict__ x, const long long* __restrict__ idx, int n, int F,
                                   float* __restrict__ y) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * F; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / F), c = (int)(i % F);
        y[i] = x[(size_t)idx[r] * F + c];
    }
}

// y = (a + b) * s
__global__ void add_scale_kernel(const float* __restrict__ a, const float* __restrict__ b, float s, long long n,
                                 float* __restrict__ y) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (a[i] + b[i]) * s;
}

// KeypointExtractor.weight_post (keypoint_extractor.py:129-134): LayerNorm -> SiLU -> Linear(D, 1) -> sigmoid (optional),
// times an optional scalar multiplier softplus(weight_mult_logit).  One warp per node.
__global__ void weight_post_kernel(const float* __restrict__ x, int n, int D, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                   const float* __restrict__ w, const float* __restrict__ b, int use_sigmoid,
                                   const float* __restrict__ mult_logit, float* __restrict__ y) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const float* xr = x + (size_t)i * D;
        float s = 0.f;
        for (int c = lane; c < D; c += 32) s += xr[c];
        const float mean = warp_sum(s) / (float)D;
        float q = 0.f;
        for (int c = lane; c < D; c += 32) { const float t = xr[c] - mean; q += t * t; }
        const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
        float acc = 0.f;
        for (int c = lane; c < D; c += 32) acc = fmaf(siluf_((xr[c] - mean) * rstd * ln_g[c] + ln_b[c]), w[c], acc);
        acc = warp_sum(acc) + b[0];
        if (use_sigmoid) acc = sigmoidf_(acc);
        if (mult_logit) { const float m = mult_logit[0]; acc *= (m > 20.f) ? m : log1pf(expf(m)); }
        if (lane == 0) y[i] = acc;
    }
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_node_linear(const float* x, int n, const int* irr_in, const int* irr_out, const float* W0,
                                const float* W1, const float* W2, const float* bias0, const float* ln_w,
                                const float* ln_b, float ln_eps, int gate, const float* res, float res_scale,
                                float* y, cudaStream_t stream) {
    if (!x || !irr_in || !irr_out || !y) return DEDF_ERR_ARG;
    NodeLinArgs a{};
    a.x = x; a.n = n;
    a.in = Irr{irr_in[0], irr_in[1], irr_in[2]};
    a.out = Irr{irr_out[0], irr_out[1], irr_out[2]};
    // a block needs weights iff both sides have that l; a missing path leaves zeros
    a.W0 = (a.in.m0 && a.out.m0) ? W0 : nullptr;
    a.W1 = (a.in.m1 && a.out.m1) ? W1 : nullptr;
    a.W2 = (a.in.m2 && a.out.m2) ? W2 : nullptr;
    if ((a.in.m0 && a.out.m0 && !W0) || (a.in.m1 && a.out.m1 && !W1) || (a.in.m2 && a.out.m2 && !W2)) return DEDF_ERR_ARG;
    a.bias0 = bias0;
    a.ln = (ln_w != nullptr); a.ln_w = ln_w; a.ln_b = ln_b; a.ln_eps = ln_eps;
    if (a.ln && a.in.m0 && !ln_b) return DEDF_ERR_ARG;
    a.gate = gate; a.res = res; a.res_scale = res_scale; a.y = y;
    if (gate && a.out.m0 < a.out.m1 + a.out.m2) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const int lda0 = pad_lda(a.in.m0), lda1 = pad_lda(a.in.m1), lda2 = pad_lda(a.in.m2);
    const int ldo = a.out.dim() + 1;
    const size_t base_floats = (size_t)kNodeTN * lda0 + 3 * kNodeTN * lda1 + 5 * kNodeTN * lda2 + (size_t)kNodeTN * ldo + kNodeTN * 3 + ((kNodeTN + 3) & ~3);
    const size_t w_floats = (size_t)((a.W0 ? a.in.m0 * a.out.m0 : 0) + 3) / 4 * 4 + (size_t)((a.W1 ? a.in.m1 * a.out.m1 : 0) + 3) / 4 * 4 +
                            (size_t)((a.W2 ? a.in.m2 * a.out.m2 : 0) + 3) / 4 * 4;
    constexpr size_t kMaxSmem = 220 * 1024;
    const bool wshared = (base_floats + w_floats) * sizeof(float) <= kMaxSmem;
    auto ok16 = [](const float* p, int n) { return !p || ((reinterpret_cast<uintptr_t>(p) & 15) == 0 && (n & 3) == 0); };
    a.bulk_w = wshared && ok16(a.W0, a.in.m0 * a.out.m0) && ok16(a.W1, a.in.m1 * a.out.m1) && ok16(a.W2, a.in.m2 * a.out.m2);
    const size_t smem = (base_floats + (wshared ? w_floats : 0)) * sizeof(float);
    if (smem > kMaxSmem) return DEDF_ERR_UNSUPPORTED;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(node_linear_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        cudaFuncSetAttribute(node_linear_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        cudaFuncSetAttribute(node_linear_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        cudaFuncSetAttribute(node_linear_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        attr_done = true;
    }
    const int n_tiles = (n + kNodeTN - 1) / kNodeTN;
    // float4 weight loads need every output multiplicity to be a multiple of 4 (true for all feature irreps)
    const bool vec = (a.out.m0 % 4 == 0) && (a.out.m1 % 4 == 0) && (a.out.m2 % 4 == 0);
    const int grid = grid_for(n_tiles, 1, kNumSMs * 2);
    // TMA variant: all operands 16-byte aligned, row widths multiples of 4 floats, everything fits in shared memory
    {
        const int Fin = a.in.dim(), Fy = gate ? (a.out.dim() - a.out.m1 - a.out.m2) : a.out.dim();
        const int io_floats = kNodeTN * (Fin > Fy ? Fin : Fy);
        const size_t smem_tma = (base_floats + (size_t)io_floats + w_floats) * sizeof(float);
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        const bool tma_ok = vec && a.bulk_w && (Fin % 4 == 0) && (Fy % 4 == 0) && al16(x) && al16(y) && (!res || al16(res)) &&
                            smem_tma <= 226 * 1024 && !getenv("DEDF_NO_TMA_NODE");
        if (tma_ok) {
            static bool done = false;
            if (!done) { cudaFuncSetAttribute(node_linear_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); done = true; }
            launch_pdl(node_linear_tma_kernel, dim3(grid), dim3(kNodeThreads), smem_tma, stream, a, lda0, lda1, lda2, ldo, io_floats);
            DEDF_CHECK_LAUNCH();
            return DEDF_OK;
        }
    }
    if (vec && wshared) launch_pdl((node_linear_kernel<true, true>), dim3(grid), dim3(kNodeThreads), smem, stream, a, lda0, lda1, lda2, ldo);
    else if (vec) launch_pdl((node_linear_kernel<true, false>), dim3(grid), dim3(kNodeThreads), smem, stream, a, lda0, lda1, lda2, ldo);
    else if (wshared) launch_pdl((node_linear_kernel<false, true>), dim3(grid), dim3(kNodeThreads), smem, stream, a, lda0, lda1, lda2, ldo);
    else launch_pdl((node_linear_kernel<false, false>), dim3(grid), dim3(kNodeThreads), smem, stream, a, lda0, lda1, lda2, ldo);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_gather_rows(const float* x, const long long* idx, int n, int F, float* y, cudaStream_t stream) {
    if (!x || !idx || !y || F <= 0) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    launch_pdl(gather_rows_kernel, dim3(grid_for((long long)n * F, 256, kNumSMs * 8)), dim3(256), 0, stream, x, idx, n, F, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_add_scale(const float* a, const float* b, float s, long long n, float* y, cudaStream_t stream) {
    if (!a || !b || !y) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    launch_pdl(add_scale_kernel, dim3(grid_for(n, 256, kNumSMs * 8)), dim3(256), 0, stream, a, b, s, n, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_weight_post(const float* x, int n, int dim, const float* ln_g, const float* ln_b, const float* w,
                                const float* b, int use_sigmoid, const float* mult_logit, float* y, cudaStream_t stream) {
    if (!x || !ln_g || !ln_b || !w || !b || !y || dim <= 0) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    weight_post_kernel<<<grid_for(n, 8, kNumSMs * 8), 256, 0, stream>>>(x, n, dim, ln_g, ln_b, w, b, use_sigmoid, mult_logit, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
