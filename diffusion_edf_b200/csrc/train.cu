// Training path: un-fused forward primitives that keep their intermediates, and the backward kernels of every
// primitive of the score network (get_train_loss -> loss.backward(), /root/reference/diffusion_edf/score_model_base.py:41-107,
// trainer.py:308-346).  The inference kernels (edge.cu, node.cu, head.cu, tc_mlp.cu) fuse these steps and never materialise
// the (E, 49 G) tensor-product outputs; a training step needs them for the weight gradients, so the autograd functions in
// diffusion_edf_b200/autograd_ops.py compose the model from the primitives below instead.  Correctness first: one thread
// per output element, fp32, parameter gradients accumulated with atomics (summation order is not deterministic).
//
//   lin_wgrad             dW / dbias of a block-diagonal linear (LinearRS, FCTP with 1x0e, nn.Linear)
//   ln_fwd / ln_bwd       EquivariantLayerNormV2 ('component', affine)   (equiformer/layer_norm.py:91-156); nn.LayerNorm is the
//                         (N, 0, 0) special case
//   gate_fwd / gate_bwd   Gate (fast_activation.py:210-224) ; act_fwd / act_bwd: plain SiLU
//   dtp_fwd / dtp_bwd     DepthwiseTensorProduct 'uvu' with the l<=2 harmonics (tensor_product_rescale.py:352-382)
//   gather_rows_i32 / scatter_add_rows
//   alpha_fwd / alpha_bwd attention logits: sum_k c SLReLU(pre[h,k]) alpha_dot[h,k] + edge_logit   (graph_attention.py:241-246)
//   softmax_reduce_bwd    backward of scatter_logsumexp + exp + scatter(sum)                     (graph_attention.py:254-265)
//   rbf_fwd / rbf_bwd     Gaussian radial bases with learnable mean / std / weight               (radial_func.py:168-278)
//   sinusoid              SinusoidalPositionEmbeddings (no parameters)                           (radial_func.py:291-316)
//   score_tp_fwd / _bwd   'uvu' tensor product of two feature vectors with shared weights, l_out <= 1 (score_head.py:123-139)
//   query_tf_bwd          adjoint of f' = D(q) f w.r.t. f                                        (wigner.py:257-283)
//   assemble_fwd / _bwd   mean over the vectors, rotate by q^-1, orbital term, weighted sum      (score_head.py:196-209)
#include "common.cuh"
#include "cg_slots.cuh"
#include "so3.cuh"
#include <curand_kernel.h>
#include "../../include/dedf.h"

namespace dedf {

__device__ __forceinline__ float dsiluf_(float x) { const float s = sigmoidf_(x); return s * (1.0f + x * (1.0f - s)); }
__device__ __forceinline__ float dslreluf_(float x) { const float s = sigmoidf_(x); return 0.2f + 0.8f * (s + x * s * (1.0f - s)); }

// ---------------------------------------------------------------------------------------------------------------
// block-diagonal linear: weight / bias gradients
// ---------------------------------------------------------------------------------------------------------------
// dW_l[u, w] += sum_n sum_m x[n, off_in_l + u d + m] dy[n, off_out_l + w d + m],  db[w] += sum_n dy[n, w]
// CTA = 16 x 16 weight tile of one l, over one chunk of rows; partial sums are added atomically.
constexpr int kWgRows = 16;      // rows staged per iteration
constexpr int kWgChunk = 512;    // rows per CTA

__global__ void __launch_bounds__(256) lin_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int n, Irr in, Irr out,
                                                       float* __restrict__ dW0, float* __restrict__ dW1, float* __restrict__ dW2,
                                                       float* __restrict__ db) {
    __shared__ float xs[kWgRows][16][5], ds[kWgRows][16][5];
    const int l = blockIdx.z;
    const int mi = (l == 0) ? in.m0 : (l == 1) ? in.m1 : in.m2;
    const int mo = (l == 0) ? out.m0 : (l == 1) ? out.m1 : out.m2;
    float* dW = (l == 0) ? dW0 : (l == 1) ? dW1 : dW2;
    if (mi == 0 || mo == 0 || dW == nullptr) return;
    const int d = 2 * l + 1;
    const int off_i = (l == 0) ? 0 : (l == 1) ? in.off1() : in.off2();
    const int off_o = (l == 0) ? 0 : (l == 1) ? out.off1() : out.off2();
    const int tiles_w = (mo + 15) / 16, tiles_u = (mi + 15) / 16;
    if ((int)blockIdx.x >= tiles_u * tiles_w) return;
    const int u0 = (blockIdx.x / tiles_w) * 16, w0 = (blockIdx.x % tiles_w) * 16;
    const int tu = threadIdx.x / 16, tw = threadIdx.x % 16;
    const int Fi = in.dim(), Fo = out.dim();
    const int r_beg = blockIdx.y * kWgChunk, r_end = min(n, r_beg + kWgChunk);
    float acc = 0.f, accb = 0.f;
    for (int r0 = r_beg; r0 < r_end; r0 += kWgRows) {
        __syncthreads();
        for (int i = threadIdx.x; i < kWgRows * 16 * d; i += 256) {
            const int r = i / (16 * d), c = i % (16 * d), uu = c / d, m = c % d;
            const int row = r0 + r;
            xs[r][uu][m] = (row < r_end && u0 + uu < mi) ? x[(size_t)row * Fi + off_i + (u0 + uu) * d + m] : 0.f;
            ds[r][uu][m] = (row < r_end && w0 + uu < mo) ? dy[(size_t)row * Fo + off_o + (w0 + uu) * d + m] : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < kWgRows; ++r) {
            for (int m = 0; m < d; ++m) acc = fmaf(xs[r][tu][m], ds[r][tw][m], acc);
            if (l == 0 && tu == 0) accb += ds[r][tw][0];
        }
    }
    if (u0 + tu < mi && w0 + tw < mo) atomicAdd(dW + (size_t)(u0 + tu) * mo + w0 + tw, acc);
    if (l == 0 && db && u0 == 0 && tu == 0 && w0 + tw < mo) atomicAdd(db + w0 + tw, accb);
}

// ---------------------------------------------------------------------------------------------------------------
// equivariant layer norm (standalone) forward / backward: one warp per node
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, int n, Irr irr, const float* __restrict__ w,
                                                    const float* __restrict__ b, float eps, float* __restrict__ y) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, F = irr.dim();
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const float* xr = x + (size_t)i * F;
        float* yr = y + (size_t)i * F;
        float s = 0.f;
        for (int c = lane; c < irr.m0; c += 32) s += xr[c];
        const float mean = (irr.m0 > 0) ? warp_sum(s) / (float)irr.m0 : 0.f;
        float q0 = 0.f, q1 = 0.f, q2 = 0.f;
        for (int c = lane; c < irr.m0; c += 32) { const float t = xr[c] - mean; q0 += t * t; }
        for (int c = lane; c < 3 * irr.m1; c += 32) { const float t = xr[irr.off1() + c]; q1 += t * t; }
        for (int c = lane; c < 5 * irr.m2; c += 32) { const float t = xr[irr.off2() + c]; q2 += t * t; }
        q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2);
        const float s0 = (irr.m0 > 0) ? rsqrtf(q0 / (float)irr.m0 + eps) : 0.f;
        const float s1 = (irr.m1 > 0) ? rsqrtf(q1 / (float)(3 * irr.m1) + eps) : 0.f;
        const float s2 = (irr.m2 > 0) ? rsqrtf(q2 / (float)(5 * irr.m2) + eps) : 0.f;
        for (int c = lane; c < irr.m0; c += 32) yr[c] = (xr[c] - mean) * s0 * w[c] + b[c];
        for (int c = lane; c < 3 * irr.m1; c += 32) yr[irr.off1() + c] = xr[irr.off1() + c] * s1 * w[irr.m0 + c / 3];
        for (int c = lane; c < 5 * irr.m2; c += 32) yr[irr.off2() + c] = xr[irr.off2() + c] * s2 * w[irr.m0 + irr.m1 + c / 5];
    }
}

__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, int n, Irr irr,
                                                    const float* __restrict__ w, float eps, float* __restrict__ dx,
                                                    float* __restrict__ dw, float* __restrict__ db) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, F = irr.dim();
    for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
        const float* xr = x + (size_t)i * F;
        const float* gr = g + (size_t)i * F;
        float* dr = dx + (size_t)i * F;
        float s = 0.f;
        for (int c = lane; c < irr.m0; c += 32) s += xr[c];
        const float mean = (irr.m0 > 0) ? warp_sum(s) / (float)irr.m0 : 0.f;
        float q0 = 0.f, q1 = 0.f, q2 = 0.f, p0 = 0.f, p1 = 0.f, p2 = 0.f;       // q: sum x^2 ; p: sum g w x
        for (int c = lane; c < irr.m0; c += 32) { const float t = xr[c] - mean; q0 += t * t; p0 += gr[c] * w[c] * t; }
        for (int c = lane; c < 3 * irr.m1; c += 32) { const float t = xr[irr.off1() + c]; q1 += t * t; p1 += gr[irr.off1() + c] * w[irr.m0 + c / 3] * t; }
        for (int c = lane; c < 5 * irr.m2; c += 32) { const float t = xr[irr.off2() + c]; q2 += t * t; p2 += gr[irr.off2() + c] * w[irr.m0 + irr.m1 + c / 5] * t; }
        q0 = warp_sum(q0); q1 = warp_sum(q1); q2 = warp_sum(q2); p0 = warp_sum(p0); p1 = warp_sum(p1); p2 = warp_sum(p2);
        const float n0 = (float)max(irr.m0, 1), n1 = (float)max(3 * irr.m1, 1), n2 = (float)max(5 * irr.m2, 1);
        const float s0 = (irr.m0 > 0) ? rsqrtf(q0 / n0 + eps) : 0.f, s1 = (irr.m1 > 0) ? rsqrtf(q1 / n1 + eps) : 0.f,
                    s2 = (irr.m2 > 0) ? rsqrtf(q2 / n2 + eps) : 0.f;
        // l = 0: dxc = g w s - xc s^3 p / n ; dx = dxc - mean(dxc)
        float dsum = 0.f;
        for (int c = lane; c < irr.m0; c += 32) {
            const float t = xr[c] - mean;
            const float dxc = gr[c] * w[c] * s0 - t * s0 * s0 * s0 * p0 / n0;
            dr[c] = dxc; dsum += dxc;
            atomicAdd(dw + c, gr[c] * t * s0);
            atomicAdd(db + c, gr[c]);
        }
        dsum = warp_sum(dsum) / n0;
        for (int c = lane; c < irr.m0; c += 32) dr[c] -= dsum;
        for (int c = lane; c < 3 * irr.m1; c += 32) {
            const int o = irr.off1() + c, u = irr.m0 + c / 3;
            const float t = xr[o];
            dr[o] = gr[o] * w[u] * s1 - t * s1 * s1 * s1 * p1 / n1;
            atomicAdd(dw + u, gr[o] * t * s1);
        }
        for (int c = lane; c < 5 * irr.m2; c += 32) {
            const int o = irr.off2() + c, u = irr.m0 + irr.m1 + c / 5;
            const float t = xr[o];
            dr[o] = gr[o] * w[u] * s2 - t * s2 * s2 * s2 * p2 / n2;
            atomicAdd(dw + u, gr[o] * t * s2);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// gate / activation
// ---------------------------------------------------------------------------------------------------------------
// pre: (n, ms + m1 + m2 | 3 m1 | 5 m2)  ->  y: (n, ms | 3 m1 | 5 m2);  `out` holds the PRE-gate irreps (m0 = ms + m1 + m2)
__global__ void gate_fwd_kernel(const float* __restrict__ pre, int n, Irr out, float* __restrict__ y) {
    const int ms = out.m0 - out.m1 - out.m2, Fp = out.dim(), Fy = Fp - out.m1 - out.m2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * Fy; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / Fy), c = (int)(i % Fy);
        const float* o = pre + (size_t)r * Fp;
        float v;
        if (c < ms) v = kCSilu * siluf_(o[c]);
        else if (c < ms + 3 * out.m1) v = o[out.off1() + (c - ms)] * (kCSigmoid * sigmoidf_(o[ms + (c - ms) / 3]));
        else v = o[out.off2() + (c - ms - 3 * out.m1)] * (kCSigmoid * sigmoidf_(o[ms + out.m1 + (c - ms - 3 * out.m1) / 5]));
        y[i] = v;
    }
}

// one thread per (row, pre-gate scalar channel): scalars, and gates together with their gated irrep
__global__ void gate_bwd_kernel(const float* __restrict__ pre, const float* __restrict__ g, int n, Irr out, float* __restrict__ dpre) {
    const int ms = out.m0 - out.m1 - out.m2, Fp = out.dim(), Fy = Fp - out.m1 - out.m2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * out.m0; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / out.m0), c = (int)(i % out.m0);
        const float* o = pre + (size_t)r * Fp;
        const float* gr = g + (size_t)r * Fy;
        float* d = dpre + (size_t)r * Fp;
        if (c < ms) { d[c] = gr[c] * kCSilu * dsiluf_(o[c]); continue; }
        const bool is1 = c < ms + out.m1;
        const int u = is1 ? c - ms : c - ms - out.m1, dd = is1 ? 3 : 5;
        const int po = (is1 ? out.off1() : out.off2()) + u * dd, yo = (is1 ? ms : ms + 3 * out.m1) + u * dd;
        const float sg = sigmoidf_(o[c]);
        float dot = 0.f;
        for (int k = 0; k < dd; ++k) { dot += gr[yo + k] * o[po + k]; d[po + k] = gr[yo + k] * kCSigmoid * sg; }
        d[c] = dot * kCSigmoid * sg * (1.0f - sg);
    }
}

// mode 0: SiLU ; mode 1: sigmoid
__global__ void act_fwd_kernel(const float* __restrict__ x, long long n, int mode, float* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = mode ? sigmoidf_(x[i]) : siluf_(x[i]);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, long long n, int mode, float* __restrict__ dx) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float s = sigmoidf_(x[i]);
        dx[i] = g[i] * (mode ? s * (1.0f - s) : dsiluf_(x[i]));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise tensor product, un-fused: out (E, 49 G) in the sorted-irreps layout the following LinearRS expects
// ---------------------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) dtp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ sh, const float* __restrict__ w,
                                                     long long w_stride, int E, float* __restrict__ out) {
    using D = Dtp<G>;
    constexpr int NCH = D::M0 + D::M1 + D::M2, B1 = D::D0, B2 = D::D0 + 3 * D::D1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * NCH; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / NCH), c = (int)(i % NCH);
        const float* xe = x + (size_t)e * D::F;
        const float* we = w + (size_t)e * w_stride;
        const float* s = sh + (size_t)e * 9;
        float* o = out + (size_t)e * D::FOUT;
        if (c < D::M0) {
            const int ch = c;
            float r[9];
            dtp_l0(xe[ch], we[D::W_K0 + ch], we[D::W_K1 + ch], we[D::W_K2 + ch], s, r);
            o[D::C0_K0 + ch] = r[0];
            for (int k = 0; k < 3; ++k) o[B1 + (D::C1_K1 + ch) * 3 + k] = r[1 + k];
            for (int k = 0; k < 5; ++k) o[B2 + (D::C2_K2 + ch) * 5 + k] = r[4 + k];
        } else if (c < D::M0 + D::M1) {
            const int ch = c - D::M0;
            float xv[3], wv[6], r[20];
            for (int k = 0; k < 3; ++k) xv[k] = xe[D::M0 + 3 * ch + k];
            for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K3 + ch + k * D::M1];
            dtp_l1(xv, wv, s, r);
            for (int k = 0; k < 3; ++k) {
                o[B1 + (D::C1_K3 + ch) * 3 + k] = r[k]; o[B1 + (D::C1_K5 + ch) * 3 + k] = r[4 + k]; o[B1 + (D::C1_K7 + ch) * 3 + k] = r[12 + k];
            }
            o[D::C0_K4 + ch] = r[3];
            for (int k = 0; k < 5; ++k) { o[B2 + (D::C2_K6 + ch) * 5 + k] = r[7 + k]; o[B2 + (D::C2_K8 + ch) * 5 + k] = r[15 + k]; }
        } else {
            const int ch = c - D::M0 - D::M1;
            float xv[5], wv[6], r[22];
            for (int k = 0; k < 5; ++k) xv[k] = xe[D::M0 + 3 * D::M1 + 5 * ch + k];
            for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K9 + ch + k * D::M2];
            dtp_l2(xv, wv, s, r);
            for (int k = 0; k < 5; ++k) {
                o[B2 + (D::C2_K9 + ch) * 5 + k] = r[k]; o[B2 + (D::C2_K11 + ch) * 5 + k] = r[8 + k]; o[B2 + (D::C2_K14 + ch) * 5 + k] = r[17 + k];
            }
            for (int k = 0; k < 3; ++k) { o[B1 + (D::C1_K10 + ch) * 3 + k] = r[5 + k]; o[B1 + (D::C1_K13 + ch) * 3 + k] = r[14 + k]; }
            o[D::C0_K12 + ch] = r[13];
        }
    }
}

// dx (E, F) and dw: per edge (E, NUMEL) written, or shared (NUMEL) accumulated with atomics (w_stride == 0)
template <int G>
__global__ void __launch_bounds__(256) dtp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ sh, const float* __restrict__ w,
                                                     long long w_stride, const float* __restrict__ g, int E, float* __restrict__ dx,
                                                     float* __restrict__ dw) {
    using D = Dtp<G>;
    constexpr int NCH = D::M0 + D::M1 + D::M2, B1 = D::D0, B2 = D::D0 + 3 * D::D1;
    const float c5 = 2.23606797749979f, c3 = 1.7320508075688772f;
    (void)c5; (void)c3;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * NCH; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / NCH), c = (int)(i % NCH);
        const float* xe = x + (size_t)e * D::F;
        const float* we = w + (size_t)e * w_stride;
        const float* s = sh + (size_t)e * 9;
        const float* go = g + (size_t)e * D::FOUT;
        float* dxe = dx + (size_t)e * D::F;
        float* dwe = dw + (size_t)e * w_stride;
        auto put_w = [&](int idx, float v) { if (w_stride) dwe[idx] = v; else atomicAdd(dw + idx, v); };
        if (c < D::M0) {
            const int ch = c;
            const float xv = xe[ch];
            // o0 = x w0 sh0 ; o1[k] = x w1 sh[1+k] ; o2[k] = x w2 sh[4+k]
            const float t0 = s[0] * go[D::C0_K0 + ch];
            float t1 = 0.f, t2 = 0.f;
            for (int k = 0; k < 3; ++k) t1 += s[1 + k] * go[B1 + (D::C1_K1 + ch) * 3 + k];
            for (int k = 0; k < 5; ++k) t2 += s[4 + k] * go[B2 + (D::C2_K2 + ch) * 5 + k];
            dxe[ch] = we[D::W_K0 + ch] * t0 + we[D::W_K1 + ch] * t1 + we[D::W_K2 + ch] * t2;
            put_w(D::W_K0 + ch, xv * t0); put_w(D::W_K1 + ch, xv * t1); put_w(D::W_K2 + ch, xv * t2);
        } else if (c < D::M0 + D::M1) {
            const int ch = c - D::M0;
            float xv[3], wv[6], d[3] = {0.f, 0.f, 0.f}, t[5], a[3];
            for (int k = 0; k < 3; ++k) xv[k] = xe[D::M0 + 3 * ch + k];
            for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K3 + ch + k * D::M1];
            const float* g3 = go + B1 + (D::C1_K3 + ch) * 3; const float g4 = go[D::C0_K4 + ch];
            const float* g5 = go + B1 + (D::C1_K5 + ch) * 3; const float* g6 = go + B2 + (D::C2_K6 + ch) * 5;
            const float* g7 = go + B1 + (D::C1_K7 + ch) * 3; const float* g8 = go + B2 + (D::C2_K8 + ch) * 5;
            // k3: o = w sh0 x
            { float dot = 0.f; for (int k = 0; k < 3; ++k) { dot += xv[k] * g3[k]; d[k] += wv[0] * s[0] * g3[k]; } put_w(D::W_K3 + ch, dot * s[0]); }
            // k4: o = w cg_110(x, sh1)
            { cg_110(xv, s + 1, t); put_w(D::W_K3 + ch + 1 * D::M1, t[0] * g4); cg_110_dx(s + 1, &g4, a); for (int k = 0; k < 3; ++k) d[k] += wv[1] * a[k]; }
            // k5: cg_111
            { cg_111(xv, s + 1, t); float dot = 0.f; for (int k = 0; k < 3; ++k) dot += t[k] * g5[k]; put_w(D::W_K3 + ch + 2 * D::M1, dot);
              cg_111_dx(s + 1, g5, a); for (int k = 0; k < 3; ++k) d[k] += wv[2] * a[k]; }
            // k6: cg_112
            { cg_112(xv, s + 1, t); float dot = 0.f; for (int k = 0; k < 5; ++k) dot += t[k] * g6[k]; put_w(D::W_K3 + ch + 3 * D::M1, dot);
              cg_112_dx(s + 1, g6, a); for (int k = 0; k < 3; ++k) d[k] += wv[3] * a[k]; }
            // k7: cg_121 (with the l=2 harmonics)
            { cg_121(xv, s + 4, t); float dot = 0.f; for (int k = 0; k < 3; ++k) dot += t[k] * g7[k]; put_w(D::W_K3 + ch + 4 * D::M1, dot);
              cg_121_dx(s + 4, g7, a); for (int k = 0; k < 3; ++k) d[k] += wv[4] * a[k]; }
            // k8: cg_122
            { cg_122(xv, s + 4, t); float dot = 0.f; for (int k = 0; k < 5; ++k) dot += t[k] * g8[k]; put_w(D::W_K3 + ch + 5 * D::M1, dot);
              cg_122_dx(s + 4, g8, a); for (int k = 0; k < 3; ++k) d[k] += wv[5] * a[k]; }
            for (int k = 0; k < 3; ++k) dxe[D::M0 + 3 * ch + k] = d[k];
        } else {
            const int ch = c - D::M0 - D::M1;
            float xv[5], wv[6], d[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, t[5], a[5];
            for (int k = 0; k < 5; ++k) xv[k] = xe[D::M0 + 3 * D::M1 + 5 * ch + k];
            for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K9 + ch + k * D::M2];
            const float* g9 = go + B2 + (D::C2_K9 + ch) * 5; const float* g10 = go + B1 + (D::C1_K10 + ch) * 3;
            const float* g11 = go + B2 + (D::C2_K11 + ch) * 5; const float g12 = go[D::C0_K12 + ch];
            const float* g13 = go + B1 + (D::C1_K13 + ch) * 3; const float* g14 = go + B2 + (D::C2_K14 + ch) * 5;
            // k9: o = w sh0 x
            { float dot = 0.f; for (int k = 0; k < 5; ++k) { dot += xv[k] * g9[k]; d[k] += wv[0] * s[0] * g9[k]; } put_w(D::W_K9 + ch, dot * s[0]); }
            // k10: cg_211
            { cg_211(xv, s + 1, t); float dot = 0.f; for (int k = 0; k < 3; ++k) dot += t[k] * g10[k]; put_w(D::W_K9 + ch + 1 * D::M2, dot);
              cg_211_dx(s + 1, g10, a); for (int k = 0; k < 5; ++k) d[k] += wv[1] * a[k]; }
            // k11: cg_212
            { cg_212(xv, s + 1, t); float dot = 0.f; for (int k = 0; k < 5; ++k) dot += t[k] * g11[k]; put_w(D::W_K9 + ch + 2 * D::M2, dot);
              cg_212_dx(s + 1, g11, a); for (int k = 0; k < 5; ++k) d[k] += wv[2] * a[k]; }
            // k12: cg_220
            { cg_220(xv, s + 4, t); put_w(D::W_K9 + ch + 3 * D::M2, t[0] * g12); cg_220_dx(s + 4, &g12, a); for (int k = 0; k < 5; ++k) d[k] += wv[3] * a[k]; }
            // k13: cg_221
            { cg_221(xv, s + 4, t); float dot = 0.f; for (int k = 0; k < 3; ++k) dot += t[k] * g13[k]; put_w(D::W_K9 + ch + 4 * D::M2, dot);
              cg_221_dx(s + 4, g13, a); for (int k = 0; k < 5; ++k) d[k] += wv[4] * a[k]; }
            // k14: cg_222
            { cg_222(xv, s + 4, t); float dot = 0.f; for (int k = 0; k < 5; ++k) dot += t[k] * g14[k]; put_w(D::W_K9 + ch + 5 * D::M2, dot);
              cg_222_dx(s + 4, g14, a); for (int k = 0; k < 5; ++k) d[k] += wv[5] * a[k]; }
            for (int k = 0; k < 5; ++k) dxe[D::M0 + 3 * D::M1 + 5 * ch + k] = d[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// depthwise tensor product for ANY even-parity l <= 2 irreps and harmonics degree, driven by a path table
// (irreps.dtp_paths: creation order of tensor_product_rescale.py:352-382).  The un-fused tensor field of irreps outside the
// fused kernels' family runs on it (BASELINE config C1: 16x0e+8x1e with the l <= 1 harmonics).  One thread per (edge, weight):
// 'uvu' paths own disjoint output channels, so there are no atomics.  Forward only.
// ---------------------------------------------------------------------------------------------------------------
struct DtpPathTable {
    int n_paths;
    int l1[DEDF_DTP_MAX_PATHS], l2[DEDF_DTP_MAX_PATHS], lo[DEDF_DTP_MAX_PATHS], mul[DEDF_DTP_MAX_PATHS], w_off[DEDF_DTP_MAX_PATHS],
        ch_off[DEDF_DTP_MAX_PATHS];
};

__device__ __forceinline__ void cg_any(int l1, int l2, int lo, const float* x, const float* y, float* o) {
    switch (l1 * 9 + l2 * 3 + lo) {
        case 0: cg_000(x, y, o); break;   case 4: cg_011(x, y, o); break;   case 8: cg_022(x, y, o); break;
        case 10: cg_101(x, y, o); break;  case 12: cg_110(x, y, o); break;  case 13: cg_111(x, y, o); break;
        case 14: cg_112(x, y, o); break;  case 16: cg_121(x, y, o); break;  case 17: cg_122(x, y, o); break;
        case 20: cg_202(x, y, o); break;  case 22: cg_211(x, y, o); break;  case 23: cg_212(x, y, o); break;
        case 24: cg_220(x, y, o); break;  case 25: cg_221(x, y, o); break;  case 26: cg_222(x, y, o); break;
        default: for (int k = 0; k < 5; ++k) o[k] = 0.f;
    }
}

__global__ void __launch_bounds__(256) dtp_generic_fwd_kernel(const float* __restrict__ x, Irr in, const float* __restrict__ sh,
                                                             const float* __restrict__ w, long long w_stride, DtpPathTable t, Irr out,
                                                             int numel, int E, float* __restrict__ y) {
    const int Fin = in.dim(), Fout = out.dim();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * numel; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / numel), wi = (int)(i % numel);
        int p = 0;
        while (p + 1 < t.n_paths && wi >= t.w_off[p + 1]) ++p;
        const int u = wi - t.w_off[p], l1 = t.l1[p], l2 = t.l2[p], lo = t.lo[p];
        const float* xe = x + (size_t)e * Fin + (l1 == 0 ? 0 : l1 == 1 ? in.off1() : in.off2()) + u * (2 * l1 + 1);
        const float* se = sh + (size_t)e * 9 + (l2 == 0 ? 0 : l2 == 1 ? 1 : 4);
        float xv[5], sv[5], o[5];
        for (int k = 0; k < 2 * l1 + 1; ++k) xv[k] = xe[k];
        for (int k = 0; k < 2 * l2 + 1; ++k) sv[k] = se[k];
        cg_any(l1, l2, lo, xv, sv, o);
        const float wv = w[(size_t)e * w_stride + wi];
        float* ye = y + (size_t)e * Fout + (lo == 0 ? 0 : lo == 1 ? out.off1() : out.off2()) + (t.ch_off[p] + u) * (2 * lo + 1);
        for (int k = 0; k < 2 * lo + 1; ++k) ye[k] = wv * o[k];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// gather / scatter-add
// ---------------------------------------------------------------------------------------------------------------
__global__ void gather_rows_i32_kernel(const float* __restrict__ x, const int* __restrict__ idx, int n, int F, float* __restrict__ y) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * F; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[(size_t)idx[i / F] * F + (i % F)];
}
template <typename IdxT>
__global__ void scatter_add_rows_kernel(const float* __restrict__ g, const IdxT* __restrict__ idx, int n, int F, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * F; i += (long long)gridDim.x * blockDim.x)
        atomicAdd(out + (size_t)idx[i / F] * F + (i % F), g[i]);
}

// ---------------------------------------------------------------------------------------------------------------
// attention logits
// ---------------------------------------------------------------------------------------------------------------
__global__ void alpha_fwd_kernel(const float* __restrict__ pre, int E, int MA, const float* __restrict__ alpha_dot,
                                 const float* __restrict__ edge_logit, float* __restrict__ logits) {
    const int HD = MA / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * 4; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / 4), h = (int)(i % 4);
        float s = 0.f;
        for (int k = 0; k < HD; ++k) s += kCSlrelu * slreluf_(pre[(size_t)e * MA + h * HD + k]) * alpha_dot[h * HD + k];
        logits[i] = s + (edge_logit ? edge_logit[e] : 0.f);
    }
}
__global__ void alpha_bwd_kernel(const float* __restrict__ pre, int E, int MA, const float* __restrict__ alpha_dot,
                                 const float* __restrict__ g, float* __restrict__ dpre, float* __restrict__ dalpha) {
    const int HD = MA / 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * MA; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / MA), c = (int)(i % MA);
        const float gl = g[(size_t)e * 4 + c / HD], p = pre[i];
        dpre[i] = gl * kCSlrelu * dslreluf_(p) * alpha_dot[c];
        atomicAdd(dalpha + c, gl * kCSlrelu * slreluf_(p));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of the per-destination softmax + weighted sum: warp per destination
// ---------------------------------------------------------------------------------------------------------------
struct SmBwdArgs {
    const int* row_ptr; int n_dst; int n_seg;
    const float* logits; const float* val; const float* gout;     // (E,4), (E,F), (n_dst,F)
    float* dlogits; float* dval;                                    // (E,4), (E,F)
    int m0, m1, m2;
};

__global__ void __launch_bounds__(128) softmax_reduce_bwd_kernel(SmBwdArgs a) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int F = a.m0 + 3 * a.m1 + 5 * a.m2;
    constexpr int MAXC = 8;
    for (int d = blockIdx.x * wpb + (threadIdx.x >> 5); d < a.n_dst; d += gridDim.x * wpb) {
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int deg = 0;
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            deg += e - b;
            for (int i = b + lane; i < e; i += 32)
                for (int h = 0; h < 4; ++h) mx[h] = fmaxf(mx[h], a.logits[(size_t)i * 4 + h]);
        }
        float sm[4] = {0.f, 0.f, 0.f, 0.f}, logZ[4];
        for (int h = 0; h < 4; ++h) mx[h] = warp_max(mx[h]);
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            for (int i = b + lane; i < e; i += 32)
                for (int h = 0; h < 4; ++h) sm[h] += __expf(a.logits[(size_t)i * 4 + h] - mx[h]);
        }
        for (int h = 0; h < 4; ++h) { sm[h] = warp_sum(sm[h]); logZ[h] = (deg > 0) ? (logf(sm[h] + 1e-12f) + mx[h]) : 0.f; }
        int hd[MAXC];
        float go[MAXC];
        for (int j = 0; j < MAXC; ++j) {
            const int c = lane + 32 * j;
            int h = 0;
            if (c < a.m0) h = c / (a.m0 / 4);
            else if (c < a.m0 + 3 * a.m1) h = ((c - a.m0) / 3) / (a.m1 / 4);
            else if (c < F) h = ((c - a.m0 - 3 * a.m1) / 5) / (a.m2 / 4);
            hd[j] = h;
            go[j] = (c < F) ? a.gout[(size_t)d * F + c] : 0.f;
        }
        // pass A: gl[e,h] = sum_{c in h} val[e,c] gout[d,c] (stored in dlogits), S[h] = sum_e alpha gl ; dval = alpha gout
        float S[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            for (int i = b; i < e; ++i) {
                float al[4], gl[4] = {0.f, 0.f, 0.f, 0.f};
                for (int h = 0; h < 4; ++h) al[h] = __expf(a.logits[(size_t)i * 4 + h] - logZ[h]);
                for (int j = 0; j < MAXC; ++j) {
                    const int c = lane + 32 * j;
                    if (c < F) {
                        const float v = a.val[(size_t)i * F + c];
                        gl[hd[j]] += v * go[j];
                        a.dval[(size_t)i * F + c] = al[hd[j]] * go[j];
                    }
                }
                for (int h = 0; h < 4; ++h) { gl[h] = warp_sum(gl[h]); S[h] += al[h] * gl[h]; }
                if (lane < 4) a.dlogits[(size_t)i * 4 + lane] = gl[lane];
            }
        }
        __syncwarp();
        // pass B: dlogit = alpha (gl - S)
        for (int s = 0; s < a.n_seg; ++s) {
            const int b = a.row_ptr[(size_t)s * a.n_dst + d], e = a.row_ptr[(size_t)s * a.n_dst + d + 1];
            for (int i = b + lane / 4; i < e; i += 8) {
                const int h = lane & 3;
                const float al = __expf(a.logits[(size_t)i * 4 + h] - logZ[h]);
                a.dlogits[(size_t)i * 4 + h] = al * (a.dlogits[(size_t)i * 4 + h] - S[h]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Gaussian radial bases:  v[e,k] = exp(-z^2/2) sigmoid(wl_k) amp cut(d),  z = (d - mean_k) / (softplus(sl_k) + 1e-5),
// d = (len - offset) * inv_span.   mode 0: GaussianRadialBasis (amp = 4 sqrt(K), no cut);
// mode 1: GaussianRadialBasisLayerFiniteCutoff (amp = 4 sqrt(K), cut = inner soft cut-off)
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rbf_cut(float d) { return (d > 0.5f) ? 1.0f : (1.0f - soft_step3(((1.0f - d) - 0.8f) / (1.0f - 0.8f))); }
__device__ __forceinline__ float softplus_(float x) { return (x > 20.f) ? x : log1pf(expf(x)); }

__global__ void rbf_fwd_kernel(const float* __restrict__ len, int E, int K, const float* __restrict__ mean, const float* __restrict__ sl,
                               const float* __restrict__ wl, float offset, float inv_span, int mode, float* __restrict__ out) {
    const float amp = 4.0f * sqrtf((float)K);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * K; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / K), k = (int)(i % K);
        const float d = (len[e] - offset) * inv_span;
        const float z = (d - mean[k]) / (softplus_(sl[k]) + 1e-5f);
        out[i] = expf(-0.5f * z * z) * sigmoidf_(wl[k]) * amp * (mode ? rbf_cut(d) : 1.0f);
    }
}
__global__ void rbf_bwd_kernel(const float* __restrict__ len, int E, int K, const float* __restrict__ mean, const float* __restrict__ sl,
                               const float* __restrict__ wl, float offset, float inv_span, int mode, const float* __restrict__ g,
                               float* __restrict__ dmean, float* __restrict__ dsl, float* __restrict__ dwl) {
    const float amp = 4.0f * sqrtf((float)K);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)E * K; i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / K), k = (int)(i % K);
        const float d = (len[e] - offset) * inv_span;
        const float sd = softplus_(sl[k]) + 1e-5f, sw = sigmoidf_(wl[k]);
        const float z = (d - mean[k]) / sd;
        const float v = expf(-0.5f * z * z) * sw * amp * (mode ? rbf_cut(d) : 1.0f);
        const float gv = g[i] * v;
        atomicAdd(dmean + k, gv * z / sd);
        atomicAdd(dsl + k, gv * z * z / sd * sigmoidf_(sl[k]));      // d softplus = sigmoid
        atomicAdd(dwl + k, gv * (1.0f - sw));
    }
}

__global__ void sinusoid_kernel(const float* __restrict__ x, int n, int dim, const float* __restrict__ freq, float scale, float* __restrict__ out) {
    const int half = dim / 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * dim; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / dim), k = (int)(i % dim);
        const float arg = __fmul_rn(x[r] * scale, freq[k < half ? k : k - half]);
        out[i] = (k < half) ? sinf(arg) : cosf(arg);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// position gradients (EbmScoreModelHead.forward, score_head_ebm.py:192-222: score = d(-energy)/d(pose)).  The reference gets
// them from torch autograd through graph_parser._encode_edges; here every adjoint is its own kernel:
//   dtp_bwd_sh        adjoint of the depthwise tensor product w.r.t. the harmonics          (one warp per edge)
//   rbf_bwd_len       adjoint of the Gaussian radial bases w.r.t. the edge length
//   sinusoid_bwd      adjoint of the sinusoidal length embedding w.r.t. the edge length
//   edge_geom_bwd     (d length, d harmonics, d edge logit) -> d x_dst                      (graph_parser.py:146-224)
//   ebm_energy_bwd    adjoint of the energy tail w.r.t. the field and the rotated query features
//   ebm_pose_grad     (d x', d f') -> body-frame angular / linear score of every pose
// ---------------------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(256) dtp_bwd_sh_kernel(const float* __restrict__ x, const float* __restrict__ w, long long w_stride,
                                                        const float* __restrict__ g, int E, float* __restrict__ dsh) {
    using D = Dtp<G>;
    constexpr int NCH = D::M0 + D::M1 + D::M2, B1 = D::D0, B2 = D::D0 + 3 * D::D1;
    const int lane = threadIdx.x & 31;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < E; e += n_warps) {
        const float* xe = x + (size_t)e * D::F;
        const float* we = w + (size_t)e * w_stride;
        const float* go = g + (size_t)e * D::FOUT;
        float d[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) d[j] = 0.f;
        for (int c = lane; c < NCH; c += 32) {
            if (c < D::M0) {
                const int ch = c;
                const float xv = xe[ch];
                d[0] = fmaf(xv * we[D::W_K0 + ch], go[D::C0_K0 + ch], d[0]);
                const float a1 = xv * we[D::W_K1 + ch], a2 = xv * we[D::W_K2 + ch];
                for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(a1, go[B1 + (D::C1_K1 + ch) * 3 + k], d[1 + k]);
                for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(a2, go[B2 + (D::C2_K2 + ch) * 5 + k], d[4 + k]);
            } else if (c < D::M0 + D::M1) {
                const int ch = c - D::M0;
                float xv[3], wv[6], a[5];
                for (int k = 0; k < 3; ++k) xv[k] = xe[D::M0 + 3 * ch + k];
                for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K3 + ch + k * D::M1];
                const float* g3 = go + B1 + (D::C1_K3 + ch) * 3; const float g4 = go[D::C0_K4 + ch];
                const float* g5 = go + B1 + (D::C1_K5 + ch) * 3; const float* g6 = go + B2 + (D::C2_K6 + ch) * 5;
                const float* g7 = go + B1 + (D::C1_K7 + ch) * 3; const float* g8 = go + B2 + (D::C2_K8 + ch) * 5;
                d[0] = fmaf(wv[0], xv[0] * g3[0] + xv[1] * g3[1] + xv[2] * g3[2], d[0]);
                cg_110_dy(xv, &g4, a); for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(wv[1], a[k], d[1 + k]);
                cg_111_dy(xv, g5, a);  for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(wv[2], a[k], d[1 + k]);
                cg_112_dy(xv, g6, a);  for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(wv[3], a[k], d[1 + k]);
                cg_121_dy(xv, g7, a);  for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(wv[4], a[k], d[4 + k]);
                cg_122_dy(xv, g8, a);  for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(wv[5], a[k], d[4 + k]);
            } else {
                const int ch = c - D::M0 - D::M1;
                float xv[5], wv[6], a[5];
                for (int k = 0; k < 5; ++k) xv[k] = xe[D::M0 + 3 * D::M1 + 5 * ch + k];
                for (int k = 0; k < 6; ++k) wv[k] = we[D::W_K9 + ch + k * D::M2];
                const float* g9 = go + B2 + (D::C2_K9 + ch) * 5; const float* g10 = go + B1 + (D::C1_K10 + ch) * 3;
                const float* g11 = go + B2 + (D::C2_K11 + ch) * 5; const float g12 = go[D::C0_K12 + ch];
                const float* g13 = go + B1 + (D::C1_K13 + ch) * 3; const float* g14 = go + B2 + (D::C2_K14 + ch) * 5;
                float dot = 0.f;
                for (int k = 0; k < 5; ++k) dot = fmaf(xv[k], g9[k], dot);
                d[0] = fmaf(wv[0], dot, d[0]);
                cg_211_dy(xv, g10, a);  for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(wv[1], a[k], d[1 + k]);
                cg_212_dy(xv, g11, a);  for (int k = 0; k < 3; ++k) d[1 + k] = fmaf(wv[2], a[k], d[1 + k]);
                cg_220_dy(xv, &g12, a); for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(wv[3], a[k], d[4 + k]);
                cg_221_dy(xv, g13, a);  for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(wv[4], a[k], d[4 + k]);
                cg_222_dy(xv, g14, a);  for (int k = 0; k < 5; ++k) d[4 + k] = fmaf(wv[5], a[k], d[4 + k]);
            }
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) d[j] = warp_sum(d[j]);
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 9; ++j) dsh[(size_t)e * 9 + j] = d[j];
        }
    }
}

__device__ __forceinline__ float dsoft_step3(float x) { return (x <= 0.0f || x >= 1.0f) ? 0.0f : 12.0f * x * x * (1.0f - x); }

// dlen[e] = sum_k g[e, k] d out[e, k] / d len[e]
__global__ void rbf_bwd_len_kernel(const float* __restrict__ len, int E, int K, const float* __restrict__ mean, const float* __restrict__ sl,
                                   const float* __restrict__ wl, float offset, float inv_span, int mode, const float* __restrict__ g,
                                   float* __restrict__ dlen) {
    const float amp = 4.0f * sqrtf((float)K);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const float d = (len[e] - offset) * inv_span;
        const float cut = mode ? rbf_cut(d) : 1.0f;
        // rbf_cut(d) = 1 - soft_step3((0.2 - d) / 0.2) for d <= 0.5  ->  d cut / d d = soft_step3'((0.2 - d) / 0.2) / 0.2
        const float dcut = (mode && d <= 0.5f) ? dsoft_step3(((1.0f - d) - 0.8f) / (1.0f - 0.8f)) / (1.0f - 0.8f) : 0.0f;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float sd = softplus_(sl[k]) + 1e-5f;
            const float z = (d - mean[k]) / sd;
            const float base = expf(-0.5f * z * z) * sigmoidf_(wl[k]) * amp;
            acc = fmaf(g[(size_t)e * K + k], base * (dcut - cut * z / sd), acc);
        }
        dlen[e] = acc * inv_span;
    }
}

__global__ void sinusoid_bwd_kernel(const float* __restrict__ x, int n, int dim, const float* __restrict__ freq, float scale,
                                    const float* __restrict__ g, float* __restrict__ dx) {
    const int half = dim / 2;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < half; ++k) {
            const float arg = __fmul_rn(x[r] * scale, freq[k]);
            float sn, cs; sincosf(arg, &sn, &cs);
            acc = fmaf(scale * freq[k], g[(size_t)r * dim + k] * cs - g[(size_t)r * dim + half + k] * sn, acc);
        }
        dx[r] = acc;
    }
}

struct GeomBwdArgs {
    const float* x_src; const float* x_dst;
    const int* edge_src; const int* edge_dst;
    int n_edges;
    const float* g_len; const float* g_sh; const float* g_logit;     // g_logit may be null
    float* dx_dst;                                                    // (n_dst, 3), accumulated
    float ns_lo, ns_hi;
    int n_scales;
    int src_off[DEDF_MAX_SCALES + 1];
    float r[DEDF_MAX_SCALES];
};

__global__ void __launch_bounds__(256) edge_geom_bwd_kernel(GeomBwdArgs a) {
    const float s3 = 1.7320508075688772f, s5 = 2.23606797749979f, s15 = 3.872983346207417f;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < a.n_edges; e += gridDim.x * blockDim.x) {
        const int s = a.edge_src[e], d = a.edge_dst[e];
        const float vx = a.x_src[3 * s] - a.x_dst[3 * d], vy = a.x_src[3 * s + 1] - a.x_dst[3 * d + 1], vz = a.x_src[3 * s + 2] - a.x_dst[3 * d + 2];
        const float len = sqrtf(vx * vx + vy * vy + vz * vz);
        const float inv = 1.0f / fmaxf(len, 1e-12f);
        const float x = vx * inv, y = vy * inv, z = vz * inv;
        const float* gs = a.g_sh + (size_t)e * 9;
        float c = 1.0f, dc = 0.0f;
        if (a.ns_hi > 0.f) {
            const float t = (len - a.ns_lo) / (a.ns_hi - a.ns_lo);
            c = soft_step3(t); dc = dsoft_step3(t) / (a.ns_hi - a.ns_lo);
        }
        // un-cut harmonics (for the derivative of the cut) and their gradient w.r.t. the unit vector
        float Y[9]; sph_harm_l2(x, y, z, Y);
        float g_len = a.g_len[e];
        if (dc != 0.0f) {
            float acc = 0.f;
#pragma unroll
            for (int j = 1; j < 9; ++j) acc = fmaf(gs[j], Y[j], acc);
            g_len = fmaf(acc, dc, g_len);
        }
        if (a.g_logit) {
            int sc = 0;
            while (sc + 1 < a.n_scales && s >= a.src_off[sc + 1]) ++sc;
            const float r = a.r[sc];
            if (r >= 0.f) {
                const float r8 = 0.8f * r, t = (len - r8) / (r - r8);
                const float cut = 1.0f - soft_step3(t);
                if (cut > 1e-12f) g_len = fmaf(a.g_logit[e], -dsoft_step3(t) / ((r - r8) * cut), g_len);
            }
        }
        float gx = c * (s3 * gs[1] + s15 * (z * gs[4] + y * gs[5]) - s5 * x * gs[6] - s15 * x * gs[8]);
        float gy = c * (s3 * gs[2] + s15 * (x * gs[5] + z * gs[7]) + 2.0f * s5 * y * gs[6]);
        float gz = c * (s3 * gs[3] + s15 * (x * gs[4] + y * gs[7]) - s5 * z * gs[6] + s15 * z * gs[8]);
        const float gu = gx * x + gy * y + gz * z;
        gx = (gx - gu * x) * inv + g_len * x;      // d/d vec = (I - u u^T) / len . g_u + g_len u
        gy = (gy - gu * y) * inv + g_len * y;
        gz = (gz - gu * z) * inv + g_len * z;
        atomicAdd(a.dx_dst + 3 * d, -gx); atomicAdd(a.dx_dst + 3 * d + 1, -gy); atomicAdd(a.dx_dst + 3 * d + 2, -gz);    // vec = x_src - x_dst
    }
}

__global__ void ebm_energy_bwd_kernel(const float* __restrict__ key_f, const float* __restrict__ query_f, const float* __restrict__ qw,
                                      const float* __restrict__ g_energy, int n_t, int n_q, int F, float scale,
                                      float* __restrict__ dkey, float* __restrict__ dquery) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n_t * n_q * F; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / F;
        const int t = (int)(row / n_q), q = (int)(row % n_q);
        const float v = 2.0f * scale * qw[q] * g_energy[t] * (key_f[i] - query_f[i]);
        dkey[i] = v; dquery[i] = -v;
    }
}

// one CTA per pose:  lin = lin_mult R^T sum_q g_x ;  ang = ang_mult sum_q [ x_q x (R^T g_x) + sum_u f_u x (R^T g_u) (l = 1)
//                                                                          + sum_u (X_a f_u) . (D2^T g_u) (l = 2) ]
// g_x (n_t n_q, 3), g_f (n_t n_q, F): gradients of log P w.r.t. the transformed coordinates / rotated features
__global__ void __launch_bounds__(128) ebm_pose_grad_kernel(const float* __restrict__ Ts, int n_t, int n_q, Irr irr, const float* __restrict__ qx,
                                                           const float* __restrict__ qf, const float* __restrict__ g_x,
                                                           const float* __restrict__ g_f, float ang_mult, float lin_mult,
                                                           float* __restrict__ ang, float* __restrict__ lin) {
    __shared__ float sR[9], sD2[25], red[4][6];
    const int t = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        const float* T = Ts + (size_t)t * 7;
        const float nrm = sqrtf(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
        float qn[4] = {T[0] / nrm, T[1] / nrm, T[2] / nrm, T[3] / nrm};
        float R[9]; quat_to_matrix<float>(qn, R);
        for (int i = 0; i < 9; ++i) sR[i] = R[i];
        float D[25]; wigner_d2_from_R(R, D);
        for (int i = 0; i < 25; ++i) sD2[i] = D[i];
    }
    __syncthreads();
    const int F = irr.dim();
    const int per_q = 1 + irr.m1 + irr.m2;         // items of a query point: its coordinates, every l = 1 / l = 2 channel
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // ang xyz | lin xyz (world frame sums; rotated at the end)
    for (int i = tid; i < n_q * per_q; i += blockDim.x) {
        const int q = i / per_q, it = i % per_q;
        const size_t row = (size_t)t * n_q + q;
        if (it == 0 || it <= irr.m1) {
            const float* gv = (it == 0) ? g_x + row * 3 : g_f + row * F + irr.m0 + 3 * (it - 1);
            const float* v = (it == 0) ? qx + (size_t)q * 3 : qf + (size_t)q * F + irr.m0 + 3 * (it - 1);
            float b[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) b[m] = sR[0 * 3 + m] * gv[0] + sR[1 * 3 + m] * gv[1] + sR[2 * 3 + m] * gv[2];       // R^T g
            acc[0] += v[1] * b[2] - v[2] * b[1]; acc[1] += v[2] * b[0] - v[0] * b[2]; acc[2] += v[0] * b[1] - v[1] * b[0];
            if (it == 0) { acc[3] += b[0]; acc[4] += b[1]; acc[5] += b[2]; }
        } else {
            const int u = it - 1 - irr.m1;
            const float* gv = g_f + row * F + irr.off2() + 5 * u;
            const float* v = qf + (size_t)q * F + irr.off2() + 5 * u;
            float b[5];
#pragma unroll
            for (int m = 0; m < 5; ++m) { float s = 0.f; for (int j = 0; j < 5; ++j) s = fmaf(sD2[j * 5 + m], gv[j], s); b[m] = s; }   // D2^T g
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float s = 0.f;
#pragma unroll
                for (int i2 = 0; i2 < 5; ++i2)
#pragma unroll
                    for (int j = 0; j < 5; ++j) if (kGenL2[a][i2][j] != 0.0f) s = fmaf(kGenL2[a][i2][j] * v[j], b[i2], s);
                acc[a] += s;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] = warp_sum(acc[k]);
    if ((tid & 31) == 0) for (int k = 0; k < 6; ++k) red[tid >> 5][k] = acc[k];
    __syncthreads();
    if (tid < 3) {
        ang[(size_t)t * 3 + tid] = ang_mult * (red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]);
        lin[(size_t)t * 3 + tid] = lin_mult * (red[0][3 + tid] + red[1][3 + tid] + red[2][3 + tid] + red[3][3 + tid]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// score tensor product ('uvu', shared weights, mul2 > 1, l_out <= 1), un-fused
//   paths:  0: 0x0->0 | 1: 0x1->1 | 2: 1x0->1 | 3: 1x1->0 | 4: 1x1->1 | 5: 1x2->1 | 6: 2x1->1 | 7: 2x2->0 | 8: 2x2->1
//   out: [lo0: p0 (M0) | p3 (M1) | p7 (M2)] then [lo1: (p1 M0 | p2 M1 | p4 M1 | p5 M1 | p6 M2 | p8 M2) x 3]
// ---------------------------------------------------------------------------------------------------------------
struct StpCfg {
    int M0, M1, M2, F, D0, D1;
    int m1s[9], m2s[9], l1s[9], l2s[9], los[9], woff[10], uoff[10], ocol[9];
};
__host__ __device__ inline StpCfg stp_cfg(Irr irr) {
    StpCfg c;
    c.M0 = irr.m0; c.M1 = irr.m1; c.M2 = irr.m2; c.F = irr.dim();
    c.D0 = c.M0 + c.M1 + c.M2; c.D1 = c.M0 + 3 * c.M1 + 2 * c.M2;
    const int m1[9] = {c.M0, c.M0, c.M1, c.M1, c.M1, c.M1, c.M2, c.M2, c.M2};
    const int m2[9] = {c.M0, c.M1, c.M0, c.M1, c.M1, c.M2, c.M1, c.M2, c.M2};
    const int l1[9] = {0, 0, 1, 1, 1, 1, 2, 2, 2}, l2[9] = {0, 1, 0, 1, 1, 2, 1, 2, 2}, lo[9] = {0, 1, 1, 0, 1, 1, 1, 0, 1};
    // column (channel index inside the l_out block) of each path
    const int oc[9] = {0, 0, c.M0, c.M0, c.M0 + c.M1, c.M0 + 2 * c.M1, c.M0 + 3 * c.M1, c.M0 + c.M1, c.M0 + 3 * c.M1 + c.M2};
    c.woff[0] = 0; c.uoff[0] = 0;
    for (int p = 0; p < 9; ++p) {
        c.m1s[p] = m1[p]; c.m2s[p] = m2[p]; c.l1s[p] = l1[p]; c.l2s[p] = l2[p]; c.los[p] = lo[p]; c.ocol[p] = oc[p];
        c.woff[p + 1] = c.woff[p] + m1[p] * m2[p]; c.uoff[p + 1] = c.uoff[p] + m1[p];
    }
    return c;
}
__device__ __forceinline__ int stp_loff(const StpCfg& c, int l) { return (l == 0) ? 0 : (l == 1) ? c.M0 : c.M0 + 3 * c.M1; }

// o = cg(a, t) for path p (o has 2 lo + 1 entries)
__device__ __forceinline__ void stp_cg(int p, const float* a, const float* t, float* o) {
    switch (p) {
        case 0: o[0] = a[0] * t[0]; break;
        case 1: o[0] = a[0] * t[0]; o[1] = a[0] * t[1]; o[2] = a[0] * t[2]; break;
        case 2: o[0] = a[0] * t[0]; o[1] = a[1] * t[0]; o[2] = a[2] * t[0]; break;
        case 3: cg_110(a, t, o); break;
        case 4: cg_111(a, t, o); break;
        case 5: cg_121(a, t, o); break;
        case 6: cg_211(a, t, o); break;
        case 7: cg_220(a, t, o); break;
        default: cg_221(a, t, o); break;
    }
}
// da = d o / d a . g ; dt = d o / d t . g
__device__ __forceinline__ void stp_cg_bwd(int p, const float* a, const float* t, const float* g, float* da, float* dt) {
    switch (p) {
        case 0: da[0] = t[0] * g[0]; dt[0] = a[0] * g[0]; break;
        case 1: da[0] = t[0] * g[0] + t[1] * g[1] + t[2] * g[2]; dt[0] = a[0] * g[0]; dt[1] = a[0] * g[1]; dt[2] = a[0] * g[2]; break;
        case 2: da[0] = t[0] * g[0]; da[1] = t[0] * g[1]; da[2] = t[0] * g[2]; dt[0] = a[0] * g[0] + a[1] * g[1] + a[2] * g[2]; break;
        case 3: cg_110_dx(t, g, da); cg_110_dy(a, g, dt); break;
        case 4: cg_111_dx(t, g, da); cg_111_dy(a, g, dt); break;
        case 5: cg_121_dx(t, g, da); cg_121_dy(a, g, dt); break;
        case 6: cg_211_dx(t, g, da); cg_211_dy(a, g, dt); break;
        case 7: cg_220_dx(t, g, da); cg_220_dy(a, g, dt); break;
        default: cg_221_dx(t, g, da); cg_221_dy(a, g, dt); break;
    }
}

// one thread per (node, path, u)
__global__ void score_tp_fwd_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ W, int n, Irr irr,
                                    float* __restrict__ out) {
    const StpCfg c = stp_cfg(irr);
    const int NU = c.uoff[9], FO = c.D0 + 3 * c.D1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * NU; i += (long long)gridDim.x * blockDim.x) {
        const int node = (int)(i / NU), r = (int)(i % NU);
        int p = 0;
        while (r >= c.uoff[p + 1]) ++p;
        const int u = r - c.uoff[p], d1 = 2 * c.l1s[p] + 1, d2 = 2 * c.l2s[p] + 1, dout = 2 * c.los[p] + 1;
        const float* a = A + (size_t)node * c.F + stp_loff(c, c.l1s[p]) + u * d1;
        const float* b = B + (size_t)node * c.F + stp_loff(c, c.l2s[p]);
        const float* w = W + c.woff[p] + (size_t)u * c.m2s[p];
        float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, o[3];
        for (int v = 0; v < c.m2s[p]; ++v)
            for (int j = 0; j < d2; ++j) t[j] = fmaf(w[v], b[v * d2 + j], t[j]);
        stp_cg(p, a, t, o);
        float* orow = out + (size_t)node * FO;
        if (c.los[p] == 0) orow[c.ocol[p] + u] = o[0];
        else for (int k = 0; k < dout; ++k) orow[c.D0 + (c.ocol[p] + u) * 3 + k] = o[k];
    }
}

__global__ void score_tp_bwd_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ W, int n, Irr irr,
                                    const float* __restrict__ g, float* __restrict__ dA, float* __restrict__ dB, float* __restrict__ dW) {
    const StpCfg c = stp_cfg(irr);
    const int NU = c.uoff[9], FO = c.D0 + 3 * c.D1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * NU; i += (long long)gridDim.x * blockDim.x) {
        const int node = (int)(i / NU), r = (int)(i % NU);
        int p = 0;
        while (r >= c.uoff[p + 1]) ++p;
        const int u = r - c.uoff[p], d1 = 2 * c.l1s[p] + 1, d2 = 2 * c.l2s[p] + 1;
        const float* a = A + (size_t)node * c.F + stp_loff(c, c.l1s[p]) + u * d1;
        const float* b = B + (size_t)node * c.F + stp_loff(c, c.l2s[p]);
        const float* w = W + c.woff[p] + (size_t)u * c.m2s[p];
        float t[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, go[3] = {0.f, 0.f, 0.f}, da[5], dt[5];
        for (int v = 0; v < c.m2s[p]; ++v)
            for (int j = 0; j < d2; ++j) t[j] = fmaf(w[v], b[v * d2 + j], t[j]);
        const float* grow = g + (size_t)node * FO;
        if (c.los[p] == 0) go[0] = grow[c.ocol[p] + u];
        else for (int k = 0; k < 3; ++k) go[k] = grow[c.D0 + (c.ocol[p] + u) * 3 + k];
        stp_cg_bwd(p, a, t, go, da, dt);
        float* dar = dA + (size_t)node * c.F + stp_loff(c, c.l1s[p]) + u * d1;
        for (int k = 0; k < d1; ++k) atomicAdd(dar + k, da[k]);         // several paths share an input channel
        float* dbr = dB + (size_t)node * c.F + stp_loff(c, c.l2s[p]);
        float* dw = dW + c.woff[p] + (size_t)u * c.m2s[p];
        for (int v = 0; v < c.m2s[p]; ++v) {
            float s = 0.f;
            for (int j = 0; j < d2; ++j) { s = fmaf(dt[j], b[v * d2 + j], s); atomicAdd(dbr + v * d2 + j, w[v] * dt[j]); }
            atomicAdd(dw + v, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// adjoint of the query transform w.r.t. the query features: dqf[q] += D(q_t)^T g[t, q]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) query_tf_bwd_kernel(const float* __restrict__ Ts, int n_t, int n_q, Irr irr, const float* __restrict__ g,
                                                          float* __restrict__ dqf) {
    __shared__ float sR[9], sD2[25];
    const int t = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        const float* T = Ts + (size_t)t * 7;
        const float nrm = sqrtf(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
        float qn[4] = {T[0] / nrm, T[1] / nrm, T[2] / nrm, T[3] / nrm};
        float R[9]; quat_to_matrix<float>(qn, R);
        for (int i = 0; i < 9; ++i) sR[i] = R[i];
        float D[25]; wigner_d2_from_R(R, D);
        for (int i = 0; i < 25; ++i) sD2[i] = D[i];
    }
    __syncthreads();
    const int F = irr.dim();
    for (int i = tid; i < n_q * F; i += blockDim.x) {
        const int q = i / F, c = i % F;
        const float* gr = g + ((size_t)t * n_q + q) * F;
        float v;
        if (c < irr.m0) v = gr[c];
        else if (c < irr.off2()) {
            const int u = (c - irr.m0) / 3, m = (c - irr.m0) % 3;
            const float* gu = gr + irr.m0 + 3 * u;
            v = sR[0 * 3 + m] * gu[0] + sR[1 * 3 + m] * gu[1] + sR[2 * 3 + m] * gu[2];       // R^T
        } else {
            const int u = (c - irr.off2()) / 5, m = (c - irr.off2()) % 5;
            const float* gu = gr + irr.off2() + 5 * u;
            v = 0.f;
            for (int j = 0; j < 5; ++j) v = fmaf(sD2[j * 5 + m], gu[j], v);                   // D2^T
        }
        atomicAdd(dqf + i, v);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// final assembly (score_head.py:196-209): y (n_t n_q, 1 + 3 NV) gated outputs of the lin / ang products
// ---------------------------------------------------------------------------------------------------------------
__global__ void assemble_fwd_kernel(const float* __restrict__ Ts, int n_t, int n_q, int NV, const float* __restrict__ ylin,
                                    const float* __restrict__ yang, const float* __restrict__ qx, const float* __restrict__ qw,
                                    float lin_mult, float* __restrict__ ang, float* __restrict__ lin) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_t) return;
    const float* T = Ts + (size_t)t * 7;
    const float qinv[4] = {T[0], -T[1], -T[2], -T[3]};
    float la[3] = {0.f, 0.f, 0.f}, aa[3] = {0.f, 0.f, 0.f};
    const int FY = 1 + 3 * NV;
    for (int q = 0; q < n_q; ++q) {
        const float* yl = ylin + ((size_t)t * n_q + q) * FY + 1;
        const float* ya = yang + ((size_t)t * n_q + q) * FY + 1;
        float ml[3] = {0.f, 0.f, 0.f}, ma[3] = {0.f, 0.f, 0.f}, l[3], s[3];
        for (int c = 0; c < NV; ++c) for (int k = 0; k < 3; ++k) { ml[k] += yl[3 * c + k]; ma[k] += ya[3 * c + k]; }
        for (int k = 0; k < 3; ++k) { ml[k] /= (float)NV; ma[k] /= (float)NV; }
        quat_apply<float>(qinv, ml, l);
        quat_apply<float>(qinv, ma, s);
        const float px = qx[3 * q] / lin_mult, py = qx[3 * q + 1] / lin_mult, pz = qx[3 * q + 2] / lin_mult;
        const float ox = py * l[2] - pz * l[1], oy = pz * l[0] - px * l[2], oz = px * l[1] - py * l[0];
        const float w = qw[q];
        la[0] += w * l[0]; la[1] += w * l[1]; la[2] += w * l[2];
        aa[0] += w * (ox + s[0]); aa[1] += w * (oy + s[1]); aa[2] += w * (oz + s[2]);
    }
    for (int k = 0; k < 3; ++k) { lin[(size_t)t * 3 + k] = la[k]; ang[(size_t)t * 3 + k] = aa[k]; }
}

__global__ void assemble_bwd_kernel(const float* __restrict__ Ts, int n_t, int n_q, int NV, const float* __restrict__ ylin,
                                    const float* __restrict__ yang, const float* __restrict__ qx, const float* __restrict__ qw,
                                    float lin_mult, const float* __restrict__ gang, const float* __restrict__ glin,
                                    float* __restrict__ dylin, float* __restrict__ dyang, float* __restrict__ dqw) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_t) return;
    const float* T = Ts + (size_t)t * 7;
    const float qinv[4] = {T[0], -T[1], -T[2], -T[3]}, qf[4] = {T[0], T[1], T[2], T[3]};
    const float ga[3] = {gang[(size_t)t * 3], gang[(size_t)t * 3 + 1], gang[(size_t)t * 3 + 2]};
    const float gl[3] = {glin[(size_t)t * 3], glin[(size_t)t * 3 + 1], glin[(size_t)t * 3 + 2]};
    const int FY = 1 + 3 * NV;
    for (int q = 0; q < n_q; ++q) {
        const size_t row = ((size_t)t * n_q + q) * FY;
        const float* yl = ylin + row + 1;
        const float* ya = yang + row + 1;
        float ml[3] = {0.f, 0.f, 0.f}, ma[3] = {0.f, 0.f, 0.f}, l[3], s[3];
        for (int c = 0; c < NV; ++c) for (int k = 0; k < 3; ++k) { ml[k] += yl[3 * c + k]; ma[k] += ya[3 * c + k]; }
        for (int k = 0; k < 3; ++k) { ml[k] /= (float)NV; ma[k] /= (float)NV; }
        quat_apply<float>(qinv, ml, l);
        quat_apply<float>(qinv, ma, s);
        const float p[3] = {qx[3 * q] / lin_mult, qx[3 * q + 1] / lin_mult, qx[3 * q + 2] / lin_mult};
        const float o[3] = {p[1] * l[2] - p[2] * l[1], p[2] * l[0] - p[0] * l[2], p[0] * l[1] - p[1] * l[0]};
        const float w = qw[q];
        float dwq = 0.f;
        for (int k = 0; k < 3; ++k) dwq += gl[k] * l[k] + ga[k] * (o[k] + s[k]);
        atomicAdd(dqw + q, dwq);
        // d l = w (g_lin + g_ang x p) ; d s = w g_ang   (orbital = p x l)
        const float dl[3] = {w * (gl[0] + ga[1] * p[2] - ga[2] * p[1]), w * (gl[1] + ga[2] * p[0] - ga[0] * p[2]), w * (gl[2] + ga[0] * p[1] - ga[1] * p[0])};
        const float ds[3] = {w * ga[0], w * ga[1], w * ga[2]};
        float dml[3], dma[3];
        quat_apply<float>(qf, dl, dml);      // adjoint of v -> q^-1 v q is v -> q v q^-1 (same |q|^2 scale)
        quat_apply<float>(qf, ds, dma);
        dylin[row] = 0.f; dyang[row] = 0.f;    // the dummy scalar is dropped (score_head.py:196)
        for (int c = 0; c < NV; ++c) for (int k = 0; k < 3; ++k) { dylin[row + 1 + 3 * c + k] = dml[k] / (float)NV; dyang[row + 1 + 3 * c + k] = dma[k] / (float)NV; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// train-mode dropout (graph_attention.py:111-112 alpha_dropout = nn.Dropout on the attention weights; :119-120 proj_drop =
// EquivariantDropout: one Bernoulli per (node, irrep channel), equiformer/drop.py:76-96)
// ---------------------------------------------------------------------------------------------------------------
// out[i] = 0 with probability p, else 1 / (1 - p)   (Philox, counter = offset + i)
__global__ void dropout_mask_kernel(unsigned long long seed, unsigned long long offset, long long n, float p, float* __restrict__ out) {
    const float keep = 1.0f / (1.0f - p);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i + offset, 0, &st);
        out[i] = (curand_uniform(&st) <= p) ? 0.f : keep;
    }
}
// y[r, c] = x[r, c] * mask[r, g(c)]: mode 0: g = attention head of channel c (mask (n,4)); mode 1: g = irrep channel (mask (n, m0+m1+m2));
// mode 2: one factor per row (mask (n,1): the source-point weight of an edge, gnn_block.py:190-193 -> graph_attention.py:258-259)
__global__ void group_scale_kernel(const float* __restrict__ x, const float* __restrict__ mask, int n, Irr irr, int mode, float* __restrict__ y) {
    const int F = irr.dim(), NG = (mode == 1) ? irr.nirr() : (mode == 2) ? 1 : 4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (long long)n * F; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / F), c = (int)(i % F);
        int u;
        if (c < irr.m0) u = c;
        else if (c < irr.off2()) u = irr.m0 + (c - irr.m0) / 3;
        else u = irr.m0 + irr.m1 + (c - irr.off2()) / 5;
        int gidx = u;
        if (mode == 2) gidx = 0;
        else if (!mode) gidx = (u < irr.m0) ? u / (irr.m0 / 4) : (u < irr.m0 + irr.m1) ? (u - irr.m0) / (irr.m1 / 4) : (u - irr.m0 - irr.m1) / (irr.m2 / 4);
        y[i] = x[i] * mask[(size_t)r * NG + gidx];
    }
}

// out[e] = w[edge_src[e]] for e < *n_edges, 0 beyond (edge buffers may be sized by a capacity)
__global__ void edge_gather_scalar_kernel(const float* __restrict__ w, const int* __restrict__ edge_src, const int* __restrict__ n_edges,
                                          int max_edges, float* __restrict__ out) {
    const int E = *n_edges;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < max_edges; e += gridDim.x * blockDim.x) out[e] = (e < E) ? w[edge_src[e]] : 0.f;
}
// out[r] = sum_c a[r, c] b[r, c]   (gradient of a per-row scale factor)
__global__ void rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int F, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < n; r += gridDim.x * wpb) {
        float s = 0.f;
        for (int c = lane; c < F; c += 32) s = fmaf(a[(size_t)r * F + c], b[(size_t)r * F + c], s);
        s = warp_sum(s);
        if (lane == 0) out[r] = s;
    }
}

}  // namespace dedf

using namespace dedf;

// ---------------------------------------------------------------------------------------------------------------
// C ABI (training path)
// ---------------------------------------------------------------------------------------------------------------
#define DEDF_GRID(n) grid_for((long long)(n), 256, kNumSMs * 8)

extern "C" int dedf_lin_wgrad(const float* x, const float* dy, int n, const int* irr_in, const int* irr_out, float* dW0, float* dW1,
                              float* dW2, float* db, cudaStream_t stream) {
    if (!x || !dy || !irr_in || !irr_out) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const Irr in{irr_in[0], irr_in[1], irr_in[2]}, out{irr_out[0], irr_out[1], irr_out[2]};
    int tiles = 1;
    for (int l = 0; l < 3; ++l) {
        const int mi = (l == 0) ? in.m0 : (l == 1) ? in.m1 : in.m2, mo = (l == 0) ? out.m0 : (l == 1) ? out.m1 : out.m2;
        const int t = ((mi + 15) / 16) * ((mo + 15) / 16);
        tiles = t > tiles ? t : tiles;
    }
    dim3 grid(tiles, (n + kWgChunk - 1) / kWgChunk, 3);
    lin_wgrad_kernel<<<grid, 256, 0, stream>>>(x, dy, n, in, out, dW0, dW1, dW2, db);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_ln_fwd(const float* x, int n, const int* irr, const float* w, const float* b, float eps, float* y, cudaStream_t stream) {
    if (!x || !irr || !w || !y || (irr[0] > 0 && !b)) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    ln_fwd_kernel<<<grid_for(n, 8, kNumSMs * 8), 256, 0, stream>>>(x, n, Irr{irr[0], irr[1], irr[2]}, w, b, eps, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_ln_bwd(const float* x, const float* g, int n, const int* irr, const float* w, float eps, float* dx, float* dw,
                           float* db, cudaStream_t stream) {
    if (!x || !g || !irr || !w || !dx || !dw || (irr[0] > 0 && !db)) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    ln_bwd_kernel<<<grid_for(n, 8, kNumSMs * 8), 256, 0, stream>>>(x, g, n, Irr{irr[0], irr[1], irr[2]}, w, eps, dx, dw, db);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_gate_fwd(const float* pre, int n, const int* irr_pre, float* y, cudaStream_t stream) {
    if (!pre || !irr_pre || !y || irr_pre[0] < irr_pre[1] + irr_pre[2]) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const Irr o{irr_pre[0], irr_pre[1], irr_pre[2]};
    gate_fwd_kernel<<<DEDF_GRID((long long)n * o.dim()), 256, 0, stream>>>(pre, n, o, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_gate_bwd(const float* pre, const float* g, int n, const int* irr_pre, float* dpre, cudaStream_t stream) {
    if (!pre || !g || !irr_pre || !dpre || irr_pre[0] < irr_pre[1] + irr_pre[2]) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const Irr o{irr_pre[0], irr_pre[1], irr_pre[2]};
    gate_bwd_kernel<<<DEDF_GRID((long long)n * o.m0), 256, 0, stream>>>(pre, g, n, o, dpre);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_act_fwd(const float* x, long long n, int mode, float* y, cudaStream_t stream) {
    if (!x || !y || mode < 0 || mode > 1) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    act_fwd_kernel<<<DEDF_GRID(n), 256, 0, stream>>>(x, n, mode, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_act_bwd(const float* x, const float* g, long long n, int mode, float* dx, cudaStream_t stream) {
    if (!x || !g || !dx || mode < 0 || mode > 1) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    act_bwd_kernel<<<DEDF_GRID(n), 256, 0, stream>>>(x, g, n, mode, dx);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_dtp_fwd(int mul1, const float* x, const float* sh, const float* w, long long w_stride, int n_edges, float* out,
                            cudaStream_t stream) {
    if (!x || !sh || !w || !out) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    if (mul1 == 32) dtp_fwd_kernel<32><<<DEDF_GRID((long long)n_edges * 112), 256, 0, stream>>>(x, sh, w, w_stride, n_edges, out);
    else if (mul1 == 16) dtp_fwd_kernel<16><<<DEDF_GRID((long long)n_edges * 56), 256, 0, stream>>>(x, sh, w, w_stride, n_edges, out);
    else return DEDF_ERR_UNSUPPORTED;
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_dtp_bwd(int mul1, const float* x, const float* sh, const float* w, long long w_stride, const float* g, int n_edges,
                            float* dx, float* dw, cudaStream_t stream) {
    if (!x || !sh || !w || !g || !dx || !dw) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    if (mul1 == 32) dtp_bwd_kernel<32><<<DEDF_GRID((long long)n_edges * 112), 256, 0, stream>>>(x, sh, w, w_stride, g, n_edges, dx, dw);
    else if (mul1 == 16) dtp_bwd_kernel<16><<<DEDF_GRID((long long)n_edges * 56), 256, 0, stream>>>(x, sh, w, w_stride, g, n_edges, dx, dw);
    else return DEDF_ERR_UNSUPPORTED;
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_gather_rows_i32(const float* x, const int* idx, int n, int F, float* y, cudaStream_t stream) {
    if (!x || !idx || !y || F <= 0) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    gather_rows_i32_kernel<<<DEDF_GRID((long long)n * F), 256, 0, stream>>>(x, idx, n, F, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_scatter_add_rows(const float* g, const void* idx, int idx_is_i64, int n, int F, float* out, cudaStream_t stream) {
    if (!g || !idx || !out || F <= 0) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    if (idx_is_i64) scatter_add_rows_kernel<long long><<<DEDF_GRID((long long)n * F), 256, 0, stream>>>(g, static_cast<const long long*>(idx), n, F, out);
    else scatter_add_rows_kernel<int><<<DEDF_GRID((long long)n * F), 256, 0, stream>>>(g, static_cast<const int*>(idx), n, F, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_alpha_fwd(const float* pre, int n_edges, int ma, const float* alpha_dot, const float* edge_logit, float* logits,
                              cudaStream_t stream) {
    if (!pre || !alpha_dot || !logits || ma % 4) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    alpha_fwd_kernel<<<DEDF_GRID((long long)n_edges * 4), 256, 0, stream>>>(pre, n_edges, ma, alpha_dot, edge_logit, logits);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_alpha_bwd(const float* pre, int n_edges, int ma, const float* alpha_dot, const float* g, float* dpre, float* dalpha,
                              cudaStream_t stream) {
    if (!pre || !alpha_dot || !g || !dpre || !dalpha || ma % 4) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    alpha_bwd_kernel<<<DEDF_GRID((long long)n_edges * ma), 256, 0, stream>>>(pre, n_edges, ma, alpha_dot, g, dpre, dalpha);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_softmax_reduce_bwd(const int* row_ptr, int n_dst, int n_seg, const float* logits, const float* val, const float* gout,
                                       int m0, int m1, int m2, float* dlogits, float* dval, cudaStream_t stream) {
    if (!row_ptr || !logits || !val || !gout || !dlogits || !dval || n_seg < 1) return DEDF_ERR_ARG;
    if (m0 % 4 || m1 % 4 || m2 % 4 || m0 + 3 * m1 + 5 * m2 > 256) return DEDF_ERR_UNSUPPORTED;
    if (n_dst <= 0) return DEDF_OK;
    SmBwdArgs a{row_ptr, n_dst, n_seg, logits, val, gout, dlogits, dval, m0, m1, m2};
    softmax_reduce_bwd_kernel<<<grid_for(n_dst, 4, kNumSMs * 16), 128, 0, stream>>>(a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_rbf_fwd(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                            float offset, float inv_span, int mode, float* out, cudaStream_t stream) {
    if (!len || !mean || !std_logit || !weight_logit || !out || k <= 0) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    rbf_fwd_kernel<<<DEDF_GRID((long long)n_edges * k), 256, 0, stream>>>(len, n_edges, k, mean, std_logit, weight_logit, offset, inv_span, mode, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_rbf_bwd(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                            float offset, float inv_span, int mode, const float* g, float* dmean, float* dstd_logit, float* dweight_logit,
                            cudaStream_t stream) {
    if (!len || !mean || !std_logit || !weight_logit || !g || !dmean || !dstd_logit || !dweight_logit || k <= 0) return DEDF_ERR_ARG;
    if (n_edges <= 0) return DEDF_OK;
    rbf_bwd_kernel<<<DEDF_GRID((long long)n_edges * k), 256, 0, stream>>>(len, n_edges, k, mean, std_logit, weight_logit, offset, inv_span, mode, g,
                                                                        dmean, dstd_logit, dweight_logit);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_sinusoid(const float* x, int n, int dim, const float* freq, float scale, float* out, cudaStream_t stream) {
    if (!x || !freq || !out || dim < 2 || dim % 2) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    sinusoid_kernel<<<DEDF_GRID((long long)n * dim), 256, 0, stream>>>(x, n, dim, freq, scale, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_score_tp_fwd(const float* a, const float* b, const float* w, int n, const int* irr, float* out, cudaStream_t stream) {
    if (!a || !b || !w || !irr || !out) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const Irr ir{irr[0], irr[1], irr[2]};
    score_tp_fwd_kernel<<<DEDF_GRID((long long)n * (2 * ir.m0 + 4 * ir.m1 + 3 * ir.m2)), 256, 0, stream>>>(a, b, w, n, ir, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_score_tp_bwd(const float* a, const float* b, const float* w, int n, const int* irr, const float* g, float* da, float* db,
                                 float* dw, cudaStream_t stream) {
    if (!a || !b || !w || !irr || !g || !da || !db || !dw) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    const Irr ir{irr[0], irr[1], irr[2]};
    score_tp_bwd_kernel<<<DEDF_GRID((long long)n * (2 * ir.m0 + 4 * ir.m1 + 3 * ir.m2)), 256, 0, stream>>>(a, b, w, n, ir, g, da, db, dw);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_query_transform_bwd(const float* Ts, int n_t, int n_q, const int* irr, const float* g, float* dqf, cudaStream_t stream) {
    if (!Ts || !irr || !g || !dqf) return DEDF_ERR_ARG;
    if (n_t <= 0 || n_q <= 0) return DEDF_OK;
    query_tf_bwd_kernel<<<n_t, 128, 0, stream>>>(Ts, n_t, n_q, Irr{irr[0], irr[1], irr[2]}, g, dqf);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_assemble_fwd(const float* Ts, int n_t, int n_q, int n_vec, const float* ylin, const float* yang, const float* qx,
                                 const float* qw, float lin_mult, float* ang, float* lin, cudaStream_t stream) {
    if (!Ts || !ylin || !yang || !qx || !qw || !ang || !lin) return DEDF_ERR_ARG;
    if (n_t <= 0) return DEDF_OK;
    assemble_fwd_kernel<<<(n_t + 127) / 128, 128, 0, stream>>>(Ts, n_t, n_q, n_vec, ylin, yang, qx, qw, lin_mult, ang, lin);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_assemble_bwd(const float* Ts, int n_t, int n_q, int n_vec, const float* ylin, const float* yang, const float* qx,
                                 const float* qw, float lin_mult, const float* gang, const float* glin, float* dylin, float* dyang,
                                 float* dqw, cudaStream_t stream) {
    if (!Ts || !ylin || !yang || !qx || !qw || !gang || !glin || !dylin || !dyang || !dqw) return DEDF_ERR_ARG;
    if (n_t <= 0) return DEDF_OK;
    assemble_bwd_kernel<<<(n_t + 127) / 128, 128, 0, stream>>>(Ts, n_t, n_q, n_vec, ylin, yang, qx, qw, lin_mult, gang, glin, dylin, dyang, dqw);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_dropout_mask(unsigned long long seed, unsigned long long offset, long long n, float p, float* out, cudaStream_t stream) {
    if (!out || p < 0.f || p >= 1.f) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    dropout_mask_kernel<<<DEDF_GRID(n), 256, 0, stream>>>(seed, offset, n, p, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_group_scale(const float* x, const float* mask, int n, const int* irr, int mode, float* y, cudaStream_t stream) {
    if (!x || !mask || !irr || !y) return DEDF_ERR_ARG;
    if (mode == 0 && (irr[0] % 4 || irr[1] % 4 || irr[2] % 4)) return DEDF_ERR_UNSUPPORTED;
    if (n <= 0) return DEDF_OK;
    const Irr ir{irr[0], irr[1], irr[2]};
    group_scale_kernel<<<DEDF_GRID((long long)n * ir.dim()), 256, 0, stream>>>(x, mask, n, ir, mode, y);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_edge_gather_scalar(const float* w, const int* edge_src, const int* n_edges_dev, int max_edges, float* out,
                                       cudaStream_t stream) {
    if (!w || !edge_src || !n_edges_dev || !out) return DEDF_ERR_ARG;
    if (max_edges <= 0) return DEDF_OK;
    edge_gather_scalar_kernel<<<DEDF_GRID(max_edges), 256, 0, stream>>>(w, edge_src, n_edges_dev, max_edges, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
extern "C" int dedf_rowdot(const float* a, const float* b, int n, int F, float* out, cudaStream_t stream) {
    if (!a || !b || !out || F <= 0) return DEDF_ERR_ARG;
    if (n <= 0) return DEDF_OK;
    rowdot_kernel<<<grid_for(n, 8, kNumSMs * 8), 256, 0, stream>>>(a, b, n, F, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_dtp_bwd_sh(int mul1, const float* x, const float* w, long long w_stride, const float* g, int n_edges, float* dsh,
                               cudaStream_t stream) {
    if (n_edges <= 0) return DEDF_OK;
    if (!x || !w || !g || !dsh) return DEDF_ERR_ARG;
    const int grid = grid_for((long long)n_edges * 32, 256, kNumSMs * 8);
    if (mul1 == 32) dtp_bwd_sh_kernel<32><<<grid, 256, 0, stream>>>(x, w, w_stride, g, n_edges, dsh);
    else if (mul1 == 16) dtp_bwd_sh_kernel<16><<<grid, 256, 0, stream>>>(x, w, w_stride, g, n_edges, dsh);
    else return DEDF_ERR_UNSUPPORTED;
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_rbf_bwd_len(const float* len, int n_edges, int k, const float* mean, const float* std_logit, const float* weight_logit,
                                float offset, float inv_span, int mode, const float* g, float* dlen, cudaStream_t stream) {
    if (n_edges <= 0) return DEDF_OK;
    if (!len || !mean || !std_logit || !weight_logit || !g || !dlen || k <= 0) return DEDF_ERR_ARG;
    rbf_bwd_len_kernel<<<grid_for(n_edges, 128, kNumSMs * 8), 128, 0, stream>>>(len, n_edges, k, mean, std_logit, weight_logit, offset,
                                                                                 inv_span, mode, g, dlen);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_sinusoid_bwd(const float* x, int n, int dim, const float* freq, float scale, const float* g, float* dx,
                                 cudaStream_t stream) {
    if (n <= 0) return DEDF_OK;
    if (!x || !freq || !g || !dx || dim < 2 || (dim & 1)) return DEDF_ERR_ARG;
    sinusoid_bwd_kernel<<<grid_for(n, 128, kNumSMs * 8), 128, 0, stream>>>(x, n, dim, freq, scale, g, dx);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_edge_geom_bwd(const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst, int n_edges,
                                  int n_scales, const int* src_off, const float* r, float ns_lo, float ns_hi, const float* g_len,
                                  const float* g_sh, const float* g_logit, float* dx_dst, cudaStream_t stream) {
    if (n_edges <= 0) return DEDF_OK;
    if (!x_src || !x_dst || !edge_src || !edge_dst || !g_len || !g_sh || !dx_dst) return DEDF_ERR_ARG;
    if (n_scales < 1 || n_scales > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
    if (g_logit && (!src_off || !r)) return DEDF_ERR_ARG;
    GeomBwdArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.edge_src = edge_src; a.edge_dst = edge_dst; a.n_edges = n_edges;
    a.g_len = g_len; a.g_sh = g_sh; a.g_logit = g_logit; a.dx_dst = dx_dst; a.ns_lo = ns_lo; a.ns_hi = ns_hi; a.n_scales = n_scales;
    for (int s = 0; s < n_scales; ++s) { a.src_off[s] = src_off ? src_off[s] : 0; a.r[s] = r ? r[s] : -1.f; }
    a.src_off[n_scales] = src_off ? src_off[n_scales] : 0x7fffffff;
    edge_geom_bwd_kernel<<<grid_for(n_edges, 256, kNumSMs * 8), 256, 0, stream>>>(a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_ebm_energy_bwd(const float* key_f, const float* query_f, const float* qw, const float* g_energy, int n_t, int n_q,
                                   int F, float scale, float* dkey, float* dquery, cudaStream_t stream) {
    if (n_t <= 0 || n_q <= 0) return DEDF_OK;
    if (!key_f || !query_f || !qw || !g_energy || !dkey || !dquery || F <= 0) return DEDF_ERR_ARG;
    ebm_energy_bwd_kernel<<<grid_for((long long)n_t * n_q * F, 256, kNumSMs * 8), 256, 0, stream>>>(key_f, query_f, qw, g_energy, n_t, n_q,
                                                                                                     F, scale, dkey, dquery);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_ebm_pose_grad(const float* Ts, int n_t, int n_q, const int* irr, const float* qx, const float* qf, const float* g_x,
                                  const float* g_f, float ang_mult, float lin_mult, float* ang, float* lin, cudaStream_t stream) {
    if (n_t <= 0) return DEDF_OK;
    if (!Ts || !irr || !qx || !qf || !g_x || !g_f || !ang || !lin) return DEDF_ERR_ARG;
    ebm_pose_grad_kernel<<<n_t, 128, 0, stream>>>(Ts, n_t, n_q, Irr{irr[0], irr[1], irr[2]}, qx, qf, g_x, g_f, ang_mult, lin_mult, ang, lin);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_dtp_generic_fwd(const float* x, const int* irr_in, const float* sh, const float* w, long long w_stride, int n_paths,
                                    const int* paths, const int* irr_out, int n_edges, float* out, cudaStream_t stream) {
    if (n_edges <= 0) return DEDF_OK;
    if (!x || !irr_in || !sh || !w || !paths || !irr_out || !out || n_paths < 1 || n_paths > DEDF_DTP_MAX_PATHS) return DEDF_ERR_ARG;
    DtpPathTable t{};
    t.n_paths = n_paths;
    const Irr in{irr_in[0], irr_in[1], irr_in[2]}, o{irr_out[0], irr_out[1], irr_out[2]};
    int numel = 0;
    for (int p = 0; p < n_paths; ++p) {
        const int* r = paths + 6 * p;
        t.l1[p] = r[0]; t.l2[p] = r[1]; t.lo[p] = r[2]; t.mul[p] = r[3]; t.w_off[p] = r[4]; t.ch_off[p] = r[5];
        if (r[0] < 0 || r[0] > 2 || r[1] < 0 || r[1] > 2 || r[2] < abs(r[0] - r[1]) || r[2] > 2 || r[2] > r[0] + r[1]) return DEDF_ERR_ARG;
        const int m_in = r[0] == 0 ? in.m0 : r[0] == 1 ? in.m1 : in.m2, m_out = r[2] == 0 ? o.m0 : r[2] == 1 ? o.m1 : o.m2;
        if (r[3] != m_in || r[4] != numel || r[5] < 0 || r[5] + r[3] > m_out) return DEDF_ERR_ARG;
        numel += r[3];
    }
    dtp_generic_fwd_kernel<<<grid_for((long long)n_edges * numel, 256, kNumSMs * 8), 256, 0, stream>>>(x, in, sh, w, w_stride, t, o, numel,
                                                                                                        n_edges, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
