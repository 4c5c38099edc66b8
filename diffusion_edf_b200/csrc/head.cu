// Score-head kernels that run once per (pose, diffusion step)
// (/root/reference/diffusion_edf/score_head.py:142-211, score_model_base.py:146-199):
//
//   time_embed        sinusoidal time encoding -> per-scale MLP -> time half of the edge pre-linear
//                     (score_head.py:53-63,160-164 ; multiscale_tensor_field.py:225-234, reassociated:
//                      W [len_emb | t_emb] + b = W_len len_emb + (W_t t_emb + b))
//   query_transform   x' = R(q) x + t,  f' = D(q) f with D^1 = R and D^2 built from R directly
//                     (gnn_data.py:88-100, wigner.py:257-283; no Euler angles / J matrices)
//   score_tp          the two 240 x 240 'uvu' tensor products (lin / ang), linear, gate, mean over the 32
//                     vectors, rotate back by q^-1, orbital term, weighted sum over the query points
//   pose_update       one annealed-Langevin step on SE(3) in float64 (score_model_base.py:178-193)
#include "common.cuh"
#include "cg_paths.cuh"
#include "so3.cuh"
#include <curand_kernel.h>
#include "../../include/dedf.h"

namespace dedf {

// ---------------------------------------------------------------------------
// time embedding
// ---------------------------------------------------------------------------
struct TimeArgs {
    const float* time; int n_t;
    float max_time, enc_n; int enc_dim;        // SinusoidalPositionEmbeddings(dim, max_val, n)
    const float* enc_freq;                     // (enc_dim/2) frequency table
    int h_dim, e_dim, out_dim;                 // MLP enc_dim -> h_dim -> e_dim ; pre-linear e_dim -> out_dim
    int n_scales;
    const float* W1[DEDF_MAX_SCALES]; const float* b1[DEDF_MAX_SCALES];   // (enc_dim, h_dim) transposed
    const float* W2[DEDF_MAX_SCALES]; const float* b2[DEDF_MAX_SCALES];   // (h_dim, e_dim)
    const float* Wp[DEDF_MAX_SCALES]; const float* bp[DEDF_MAX_SCALES];   // (e_dim, out_dim) = W_s[:, len_dim:]^T, b_s
    float* out;                                // (n_scales, n_t, out_dim)
};

__global__ void __launch_bounds__(128) time_embed_kernel(TimeArgs a) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    extern __shared__ float sm[];
    float* enc = sm;                 // enc_dim
    float* h = enc + a.enc_dim;      // h_dim
    float* e = h + a.h_dim;          // e_dim
    const int t = blockIdx.x, s = blockIdx.y, tid = threadIdx.x;
    const float x = a.time[t] / a.max_time * a.enc_n;
    const int half = a.enc_dim / 2;
    for (int i = tid; i < a.enc_dim; i += blockDim.x) {
        const int k = (i < half) ? i : i - half;
        const float arg = __fmul_rn(x, a.enc_freq[k]);
        enc[i] = (i < half) ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    // GEMVs: the weight loads of a column are independent of the accumulation chain; 16 of them are kept in flight
    // (the launch is a handful of CTAs whose time is pure load latency otherwise)
    auto gemv = [&](const float* __restrict__ W, const float* __restrict__ bias, const float* x, int K, int N, int o) {
        float acc0 = bias[o], acc1 = 0.f;
        int i = 0;
        for (; i + 16 <= K; i += 16) {      // 16 in flight: the three layers are 32 + 16 + 8 dependent round trips at 8
            float w[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = __ldg(W + (size_t)(i + j) * N + o);
#pragma unroll
            for (int j = 0; j < 16; j += 2) { acc0 = fmaf(x[i + j], w[j], acc0); acc1 = fmaf(x[i + j + 1], w[j + 1], acc1); }
        }
        for (; i + 8 <= K; i += 8) {
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = __ldg(W + (size_t)(i + j) * N + o);
#pragma unroll
            for (int j = 0; j < 8; j += 2) { acc0 = fmaf(x[i + j], w[j], acc0); acc1 = fmaf(x[i + j + 1], w[j + 1], acc1); }
        }
        for (; i < K; ++i) acc0 = fmaf(x[i], __ldg(W + (size_t)i * N + o), acc0);
        return acc0 + acc1;
    };
    for (int o = tid; o < a.h_dim; o += blockDim.x) h[o] = siluf_(gemv(a.W1[s], a.b1[s], enc, a.enc_dim, a.h_dim, o));
    __syncthreads();
    for (int o = tid; o < a.e_dim; o += blockDim.x) e[o] = gemv(a.W2[s], a.b2[s], h, a.h_dim, a.e_dim, o);
    __syncthreads();
    for (int o = tid; o < a.out_dim; o += blockDim.x)
        a.out[((size_t)s * a.n_t + t) * a.out_dim + o] = gemv(a.Wp[s], a.bp[s], e, a.e_dim, a.out_dim, o);
}

struct QueryArgs {
    const float* Ts; int n_t;       // (n_t, 7)
    const float* qx; const float* qf; int n_q;   // (n_q,3), (n_q,F)
    Irr irr;
    float* x_out; float* f_out;     // (n_t*n_q, 3), (n_t*n_q, F)
};

__global__ void __launch_bounds__(128) query_transform_kernel(QueryArgs a) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    __shared__ float sR[9], sD2[25], sq[4], st[3];
    const int t = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        const float* T = a.Ts + (size_t)t * 7;
        // features: q / |q|, standardised (the sign does not change R); points: raw q like transform_points
        const float nrm = sqrtf(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
        float qn[4] = {T[0] / nrm, T[1] / nrm, T[2] / nrm, T[3] / nrm};
        float R[9]; quat_to_matrix<float>(qn, R);
        for (int i = 0; i < 9; ++i) sR[i] = R[i];
        float D[25]; wigner_d2_from_R(R, D);
        for (int i = 0; i < 25; ++i) sD2[i] = D[i];
        for (int i = 0; i < 4; ++i) sq[i] = T[i];
        for (int i = 0; i < 3; ++i) st[i] = T[4 + i];
    }
    __syncthreads();
    const int F = a.irr.dim();
    for (int q = tid; q < a.n_q; q += blockDim.x) {
        float p[3] = {a.qx[3 * q], a.qx[3 * q + 1], a.qx[3 * q + 2]}, o[3];
        const float T7[7] = {sq[0], sq[1], sq[2], sq[3], st[0], st[1], st[2]};
        transform_point_f(T7, p, o);
        float* xo = a.x_out + ((size_t)t * a.n_q + q) * 3;
        xo[0] = o[0]; xo[1] = o[1]; xo[2] = o[2];
    }
    for (int i = tid; i < a.n_q * F; i += blockDim.x) {
        const int q = i / F, c = i % F;
        const float* f = a.qf + (size_t)q * F;
        float v;
        if (c < a.irr.m0) v = f[c];
        else if (c < a.irr.off2()) {
            const int u = (c - a.irr.m0) / 3, m = (c - a.irr.m0) % 3;
            const float* fu = f + a.irr.m0 + 3 * u;
            v = sR[m * 3] * fu[0] + sR[m * 3 + 1] * fu[1] + sR[m * 3 + 2] * fu[2];
        } else {
            const int u = (c - a.irr.off2()) / 5, m = (c - a.irr.off2()) % 5;
            const float* fu = f + a.irr.off2() + 5 * u;
            v = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) v = fmaf(sD2[m * 5 + j], fu[j], v);
        }
        a.f_out[((size_t)t * a.n_q) * F + i] = v;
    }
}

// ---------------------------------------------------------------------------
// score tensor products + final reduction
// ---------------------------------------------------------------------------
// uvu tensor product of two (M0,M1,M2) irreps with shared weights, outputs l<=1 only
// (score_head.py:123-139; path/weight layout SURVEY.md App. E.2):
//   k : l1 x l2 -> lo   weight block (mul1 x mul2)
//   0 : 0 x 0 -> 0 | 1 : 0 x 1 -> 1 | 2 : 1 x 0 -> 1 | 3 : 1 x 1 -> 0 | 4 : 1 x 1 -> 1
//   5 : 1 x 2 -> 1 | 6 : 2 x 1 -> 1 | 7 : 2 x 2 -> 0 | 8 : 2 x 2 -> 1
struct ScoreArgs {
    const float* Ts; int n_t;          // (n_t,7)
    const float* qf_rot;               // (n_t*n_q, F): D(q) psi  (first TP operand, 'node_input')
    const float* key_f;                // (n_t*n_q, F): field output (second operand, 'edge_attr')
    const float* qx; const float* qw; int n_q;   // query coords (n_q,3), weights (n_q)
    Irr irr;
    const float* Wd[2];                // dtp weights (lin, ang), flat in path order, each path block TRANSPOSED to [mul2][mul1]
    const float* Wl0[2]; const float* Wl1[2]; const float* bl[2];   // lin: (D0, 1+NV), (D1, NV), bias (1+NV)
    int n_vec;                         // NV = 32
    float lin_mult;
    float* ang_out; float* lin_out;    // (n_t,3)
    const float* qf;                   // (n_q, F) UN-rotated query features: used when qf_rot is null -- D(q) is applied here
                                       // (the arithmetic of query_transform_kernel), so a denoise step needs no separate launch
    // optional fused Langevin step (denoise loop): the pose update of pose_update_kernel + cast_pose_kernel for the CTA's poses
    double* T64; const double* sched; int n_steps; int* counter; const double* noise; unsigned long long seed;
    const unsigned long long* seed_dev;    // optional device copy of the seed (a cached graph serves any seed)
    double ang_mult_d, lin_mult_d; double* traj; float* T32; unsigned* ticket;
};

// One annealed-Langevin step of pose i in float64 (score_model_base.py:178-193), shared by pose_update_kernel and the fused
// tail of score_tp_kernel.  z: 6 standard normals or null -> Philox (seed, subsequence = pose, window = step).
constexpr unsigned long long kPhiloxPerStep = 16;   // cuRAND's Philox offset counts 32-bit outputs and one step draws 3 x
                                                    // curand_normal2_double = 12 of them: steps are 16 outputs apart
// The step in three parts so that the fused tail of score_tp_kernel can compute the two that do not depend on the score -- the
// noise (Philox + Box-Muller in fp64) and the per-step coefficients (square roots) -- BEFORE it waits for its inputs:
struct LangevinCoef { double den_a, den_l, c_a, c_l, n_a, n_l; };
__device__ __forceinline__ LangevinCoef langevin_coef(double t, double ang_mult, double lin_mult, double alpha_ang, double alpha_lin,
                                                      double temperature) {
    const double sq_t = sqrt(t);
    LangevinCoef c;
    c.den_a = ang_mult * sq_t; c.den_l = lin_mult * sq_t;
    c.c_a = alpha_ang / 2; c.c_l = alpha_lin / 2;
    c.n_a = sqrt(temperature * alpha_ang); c.n_l = sqrt(temperature * alpha_lin);
    return c;
}
__device__ __forceinline__ void langevin_noise(const double* z_in, unsigned long long seed, unsigned long long pose,
                                               unsigned long long step, double* z) {
    if (z_in) {
        for (int k = 0; k < 6; ++k) z[k] = z_in[k];
    } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, pose, step * kPhiloxPerStep, &st);
        for (int k = 0; k < 6; k += 2) { double2 n = curand_normal2_double(&st); z[k] = n.x; z[k + 1] = n.y; }
    }
}
__device__ __forceinline__ void langevin_apply(double* T, const float* ang, const float* lin, const double* z, const LangevinCoef& c) {
    double ang_disp[3], lin_disp[3];
    for (int k = 0; k < 3; ++k) {
        const double as = (double)ang[k] / c.den_a;
        const double ls = (double)lin[k] / c.den_l;
        ang_disp[k] = c.c_a * as + c.n_a * z[k];
        lin_disp[k] = c.c_l * ls + c.n_l * z[3 + k];
    }
    const double q[4] = {T[0], T[1], T[2], T[3]};
    // L = T[q_indices] * q_factor   (score_model_base.py:31-32,188)
    const int qi[4][3] = {{1, 2, 3}, {0, 3, 2}, {3, 0, 1}, {2, 1, 0}};
    const double qf[4][3] = {{-0.5, -0.5, -0.5}, {0.5, -0.5, 0.5}, {0.5, 0.5, -0.5}, {-0.5, 0.5, 0.5}};
    double qn[4], nrm = 0;
    for (int r = 0; r < 4; ++r) {
        double dq = 0;
        for (int c2 = 0; c2 < 3; ++c2) dq += q[qi[r][c2]] * qf[r][c2] * ang_disp[c2];
        qn[r] = q[r] + dq;
        nrm += qn[r] * qn[r];
    }
    nrm = sqrt(nrm);
    double dx[3];
    quat_apply<double>(q, lin_disp, dx);
    for (int r = 0; r < 4; ++r) T[r] = qn[r] / nrm;
    for (int k = 0; k < 3; ++k) T[4 + k] += dx[k];
}
__device__ __forceinline__ void langevin_step(double* T, const float* ang, const float* lin, const double* z_in,
                                              unsigned long long seed, unsigned long long pose, unsigned long long step,
                                              double t, double ang_mult, double lin_mult, double alpha_ang, double alpha_lin,
                                              double temperature) {
    double z[6];
    langevin_noise(z_in, seed, pose, step, z);
    langevin_apply(T, ang, lin, z, langevin_coef(t, ang_mult, lin_mult, alpha_ang, alpha_lin, temperature));
}

constexpr int kScoreQB = 2;     // (pose, query node) rows processed together: every weight load serves all of them, for both tensor
                                // products.  Large batches take 2 poses per CTA and 4 rows at a time (the kernel is then bound by the
                                // L2 stream of the 218 KB of weights per CTA pass).

// -DDEDF_SCORE_TRACE (profiles/run_score_trace.py builds its own copy of the library): thread 0 of CTA 0 stamps clock64() at every
// phase boundary of its first (pose, query-row) pass into a device array that dedf_score_trace() copies out.
#ifdef DEDF_SCORE_TRACE
__device__ long long g_score_trace[16];
#define SCORE_STAMP(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_score_trace[i] = clock64(); } while (0)
#else
#define SCORE_STAMP(i) do { } while (0)
#endif

constexpr int kScoreThreads = 384;   // 2 (NU) = 704 tasks in step 1 -> 2 rounds; 2 NY = 258 tasks in step 3 -> 1 round (256 threads: 3 and 2)

// step 1 of score_tp for one (path, u): t[qq][j] += sum_v w[v] b[qq][v][j], KB consecutive v per call.  D2 = 2 l2 + 1 is a
// compile-time constant (no predicated-off FMAs) and the KB x D2 operand values of a row are contiguous in shared memory: they
// are fetched as float4s (one LDS per four FMAs instead of one per FMA).  Same summation order as a plain loop over v.
// Weight accessor of score_tp_kernel.  The weights live in shared memory (resident variant) or in global memory; a plain pointer
// that may be either compiles to GENERIC loads with 64-bit address arithmetic (four integer instructions per weight: step 3 spent
// 560 cycles per batch of 8 weights that way, profiles/run_score_trace.py).  WPtr<true> is a 32-bit shared-memory address read with
// ld.shared; WPtr<false> a global pointer.
template <bool SM> struct WPtr;
template <> struct WPtr<true> {
    uint32_t a;
    __device__ __forceinline__ float operator[](int i) const {
        float v; asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a + 4u * (uint32_t)i)); return v;
    }
    __device__ __forceinline__ WPtr operator+(int o) const { return WPtr{a + 4u * (uint32_t)o}; }
};
template <> struct WPtr<false> {
    const float* p;
    __device__ __forceinline__ float operator[](int i) const { return p[i]; }
    __device__ __forceinline__ WPtr operator+(int o) const { return WPtr{p + o}; }
};

template <int D2, int QB, int KB, typename WP>
__device__ __forceinline__ void score_s1_batch(WP w, int m1, const float* __restrict__ b, int F, float (&acc)[QB][D2]) {
    float wv[KB];
#pragma unroll
    for (int k = 0; k < KB; ++k) wv[k] = w[k * m1];
#pragma unroll
    for (int qq = 0; qq < QB; ++qq) {
        float bb[KB * D2];
        const float4* bp = reinterpret_cast<const float4*>(b + qq * F);
#pragma unroll
        for (int x = 0; x < KB * D2 / 4; ++x) {
            const float4 t = bp[x];
            bb[4 * x] = t.x; bb[4 * x + 1] = t.y; bb[4 * x + 2] = t.z; bb[4 * x + 3] = t.w;
        }
#pragma unroll
        for (int k = 0; k < KB; ++k)
#pragma unroll
            for (int j = 0; j < D2; ++j) acc[qq][j] = fmaf(wv[k], bb[k * D2 + j], acc[qq][j]);
    }
}
template <int D2, int QB, typename WP>
__device__ __forceinline__ void score_s1_path(WP w, int m1, int m2, const float* __restrict__ b, int F,
                                              float* __restrict__ tout, int TT) {
    float acc[QB][D2];
#pragma unroll
    for (int qq = 0; qq < QB; ++qq)
#pragma unroll
        for (int j = 0; j < D2; ++j) acc[qq][j] = 0.f;
    // the chain of dependent weight loads is what a small batch pays for: 8 in flight (4 for a tail; every mul is a multiple of 4)
    int v = 0;
    for (; v + 8 <= m2; v += 8) score_s1_batch<D2, QB, 8>(w + v * m1, m1, b + v * D2, F, acc);
    for (; v < m2; v += 4) score_s1_batch<D2, QB, 4>(w + v * m1, m1, b + v * D2, F, acc);
#pragma unroll
    for (int qq = 0; qq < QB; ++qq)
#pragma unroll
        for (int j = 0; j < D2; ++j) tout[qq * TT + j] = acc[qq][j];
}

// Persistent: min(n_t / pb, 148) CTAs; each stages the tensor-product weights of both products (2 x 70 KB) and the 1e linear
// layers (2 x 24 KB) in shared memory ONCE -- by TMA bulk copies issued before the PDL wait, they are parameters -- and then
// walks its poses: the per-pose chain of dependent L2 round trips (45 us at one pose per CTA) becomes shared-memory reads.
template <int QB>
__global__ void __launch_bounds__(kScoreThreads, QB >= 8 ? 2 : 1) score_tp_kernel(ScoreArgs a, int pb, int w_smem) {
    extern __shared__ __align__(16) float sm[];
    __shared__ __align__(8) uint64_t wbar;
    SCORE_STAMP(0);
    const int M0 = a.irr.m0, M1 = a.irr.m1, M2 = a.irr.m2, F = a.irr.dim();
    const int D0 = M0 + M1 + M2, D1 = M0 + 3 * M1 + 2 * M2;   // 112, 192 channels
    const int NV = a.n_vec, NY = 1 + 4 * NV;
    // t buffers: T[p][u][j]
    const int tsz[9] = {M0, M0 * 3, M1, M1 * 3, M1 * 3, M1 * 5, M2 * 3, M2 * 5, M2 * 5};
    int toff[10]; toff[0] = 0;
    for (int p = 0; p < 9; ++p) toff[p + 1] = toff[p] + tsz[p];
    const int TT = toff[9];
    float* sa = sm;                     // [QB][F]
    float* sb = sa + QB * F;            // [QB][F]
    float* st = sb + QB * F;            // [2][QB][TT]
    float* sd0 = st + 2 * QB * TT;      // [2][QB][D0]
    float* sd1 = sd0 + 2 * QB * D0;     // [2][QB][3][D1]  planar in the vector component: step 3 reads 4 consecutive channels per LDS
    float* sy = st;                     // [2][QB][NY]     aliases the t buffers (dead after step 2; NY <= TT checked by the launcher)
    float* sres = sd1 + 2 * QB * 3 * D1;  // [pb n_q][2][3]
    float* srd = sres + (((6 * pb * a.n_q) + 3) & ~3);   // [pb][36]: R (9) and D^2 (25) of the pass's poses (on-the-fly rotation)
    const int tid = threadIdx.x;
    // weight offsets (per path: [mul2][mul1], i.e. transposed blocks, see ScoreArgs)
    const int m1s[9] = {M0, M0, M1, M1, M1, M1, M2, M2, M2};
    const int m2s[9] = {M0, M1, M0, M1, M1, M2, M1, M2, M2};
    const int l2s[9] = {0, 1, 0, 1, 1, 2, 1, 2, 2};
    int woff[10], uoff[10]; woff[0] = 0; uoff[0] = 0;
    for (int p = 0; p < 9; ++p) { woff[p + 1] = woff[p] + m1s[p] * m2s[p]; uoff[p + 1] = uoff[p] + m1s[p]; }
    const int NU = uoff[9];             // (path, u) pairs
    const int boff[3] = {0, M0, M0 + 3 * M1};   // offsets of l blocks in a feature vector
    // resident weights
    const int nWd = woff[9], nWl1 = D1 * NV, nWl0 = D0 * (1 + NV);
    float* sw = srd + 36 * pb;
    const float* Wd[2] = {a.Wd[0], a.Wd[1]};
    const float* Wl1[2] = {a.Wl1[0], a.Wl1[1]};
    const float* Wl0[2] = {a.Wl0[0], a.Wl0[1]};
    // w_smem = 2: the scalar linear layer's weights are resident too (step 3 read them from L2 in dependent batches of 8 rows:
    // 9.6 k cycles of the 38 k a 128-pose denoise step spends after its PDL wait, profiles/run_score_trace.py)
    if (w_smem) {
        if (tid == 0) {
            mbar_init(&wbar, 1);
            mbar_init_fence();
            mbar_expect_tx(&wbar, (uint32_t)(2 * (nWd + nWl1 + (w_smem > 1 ? nWl0 : 0))) * 4u);
            for (int i = 0; i < 2; ++i) {
                bulk_g2s_chunked(sw + i * nWd, a.Wd[i], (uint32_t)nWd * 4u, &wbar);
                bulk_g2s_chunked(sw + 2 * nWd + i * nWl1, a.Wl1[i], (uint32_t)nWl1 * 4u, &wbar);
                if (w_smem > 1) bulk_g2s_chunked(sw + 2 * (nWd + nWl1) + i * nWl0, a.Wl0[i], (uint32_t)nWl0 * 4u, &wbar);
            }
        }
        for (int i = 0; i < 2; ++i) {
            Wd[i] = sw + i * nWd; Wl1[i] = sw + 2 * nWd + i * nWl1;
            if (w_smem > 1) Wl0[i] = sw + 2 * (nWd + nWl1) + i * nWl0;
        }
    }
    // Fused Langevin step: the noise of this CTA's first poses and the step's coefficients depend on (seed, pose, step) and the
    // schedule only, so they are computed HERE, in the shadow of the previous kernels (Philox + Box-Muller + four square roots in
    // fp64 on one thread were 5 k cycles of the kernel's serial tail).  Reading the step counter before the PDL wait is safe: it was
    // written by the previous step's launch of this kernel, five launches back -- a kernel starts only after its predecessor
    // passed ITS wait, i.e. after everything two or more launches back has completed.
    __shared__ double s_z[8][6];
    __shared__ LangevinCoef s_coef;
    auto stage_noise = [&](int t0_, int step_) {      // threads 32 .. 32 + pb - 1: one pose each; thread 64: the coefficients
        if (tid >= 32 && tid < 32 + pb && pb <= 8) {
            const int t = t0_ + (tid - 32);
            if (t < a.n_t && step_ < a.n_steps)
                langevin_noise(a.noise ? a.noise + ((size_t)step_ * a.n_t + t) * 6 : nullptr, a.seed_dev ? *a.seed_dev : a.seed,
                               (unsigned long long)t, (unsigned long long)step_, s_z[tid - 32]);
        }
        if (tid == 64 && step_ < a.n_steps) {
            const double* row = a.sched + (size_t)step_ * 4;
            s_coef = langevin_coef(row[0], a.ang_mult_d, a.lin_mult_d, row[1], row[2], row[3]);
        }
    };
    // R(q / |q|) and D^2 from R of the poses of a pass, exactly as query_transform_kernel
    auto compute_srd = [&](int t0_, int np_) {
        if (tid < np_) {
            const float* T = a.Ts + (size_t)(t0_ + tid) * 7;
            const float nrm = sqrtf(T[0] * T[0] + T[1] * T[1] + T[2] * T[2] + T[3] * T[3]);
            float qn[4] = {T[0] / nrm, T[1] / nrm, T[2] / nrm, T[3] / nrm};
            float R[9]; quat_to_matrix<float>(qn, R);
            float D[25]; wigner_d2_from_R(R, D);
            float* o = srd + tid * 36;
            for (int i = 0; i < 9; ++i) o[i] = R[i];
            for (int i = 0; i < 25; ++i) o[9 + i] = D[i];
        }
    };
    // first operand of the rows [q0, q0 + QB) of a pass: D(q) psi (rotated on the fly, or the caller's rotated copy)
    auto stage_a = [&](int t0_, int q0_, int nq_) {
        for (int i = tid; i < QB * F; i += blockDim.x) {
            const int qq = i / F, c = i % F;
            const int row = q0_ + min(qq, nq_ - 1);
            float v;
            if (a.qf_rot) {
                v = a.qf_rot[((size_t)t0_ * a.n_q + row) * F + c];
            } else {
                const float* f = a.qf + (size_t)(row % a.n_q) * F;
                const float* sR = srd + (row / a.n_q) * 36;
                const float* sD2 = sR + 9;
                if (c < M0) v = f[c];
                else if (c < M0 + 3 * M1) {
                    const int u = (c - M0) / 3, m = (c - M0) % 3;
                    const float* fu = f + M0 + 3 * u;
                    v = sR[m * 3] * fu[0] + sR[m * 3 + 1] * fu[1] + sR[m * 3 + 2] * fu[2];
                } else {
                    const int u = (c - M0 - 3 * M1) / 5, m = (c - M0 - 3 * M1) % 5;
                    const float* fu = f + M0 + 3 * M1 + 5 * u;
                    v = 0.f;
#pragma unroll
                    for (int j = 0; j < 5; ++j) v = fmaf(sD2[m * 5 + j], fu[j], v);
                }
            }
            sa[i] = v;
        }
    };
    // Denoise step: the poses (written by the previous step's launch of this kernel) and the static query features are final two
    // or more launches back, so the Wigner matrices and the rotated query features of the CTA's first rows are staged before the
    // PDL wait as well; after it only the field output (the previous kernel's) has to be fetched.
    const int t0_first = (int)blockIdx.x * pb;
    const bool early = a.T64 != nullptr && a.qf_rot == nullptr && t0_first < a.n_t;
    if (a.T64) stage_noise(t0_first, *a.counter);
    if (early) {
        const int np_f = min(pb, a.n_t - t0_first);
        compute_srd(t0_first, np_f);
        __syncthreads();
        stage_a(t0_first, 0, min(QB, np_f * a.n_q));
    }
    SCORE_STAMP(1);
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    SCORE_STAMP(2);
    const int step_now = a.T64 ? *a.counter : 0;
    if (w_smem) { __syncthreads(); mbar_wait(&wbar, 0); }
    SCORE_STAMP(3);

    for (int t0 = blockIdx.x * pb; t0 < a.n_t; t0 += gridDim.x * pb) {
    const int np = min(pb, a.n_t - t0);
    if (a.T64 && t0 != (int)blockIdx.x * pb) {          // later passes of a persistent CTA: their noise, off the tail as well
        __syncthreads();                                // (the previous pass's tail has read s_z)
        stage_noise(t0, step_now);
    }
    const int rows_total = np * a.n_q;  // (pose, query node) rows of this pass: consecutive nodes of qf_rot / key_f
    const bool pre = early && t0 == t0_first;       // this pass's Wigner matrices and first rows were staged before the wait
    if (!a.qf_rot && !pre) {
        __syncthreads();                // (the previous pass is done with srd)
        compute_srd(t0, np);
    }
    for (int q0 = 0; q0 < rows_total; q0 += QB) {
        const int nq = min(QB, rows_total - q0);
        __syncthreads();
        if (!(pre && q0 == 0)) stage_a(t0, q0, nq);
        for (int i = tid; i < QB * F; i += blockDim.x) {
            const int qq = i / F, c = i % F;
            const int row = q0 + min(qq, nq - 1);
            sb[i] = a.key_f[((size_t)t0 * a.n_q + row) * F + c];
        }
        __syncthreads();
        SCORE_STAMP(4);
        // step 1: t_p[u][j] = sum_v W_p[u][v] b_{l2}[v][j] for both tensor products and both query nodes: one thread per
        // (which, path, u); consecutive threads take consecutive u (coalesced weight rows, 4 loads in flight), b is broadcast
        auto step1 = [&](auto wd0, auto wd1) {
            for (int i = tid; i < 2 * NU; i += blockDim.x) {
                const int which = i / NU, r = i % NU;
                int p = 0;
                while (r >= uoff[p + 1]) ++p;
                const int u = r - uoff[p], l2 = l2s[p];
                const auto w = (which ? wd1 : wd0) + (woff[p] + u);
                const float* b = sb + boff[l2];
                float* tout = st + (size_t)(which * QB) * TT + toff[p] + u * (2 * l2 + 1);
                if (l2 == 0) score_s1_path<1, QB>(w, m1s[p], m2s[p], b, F, tout, TT);
                else if (l2 == 1) score_s1_path<3, QB>(w, m1s[p], m2s[p], b, F, tout, TT);
                else score_s1_path<5, QB>(w, m1s[p], m2s[p], b, F, tout, TT);
            }
        };
        if (w_smem) step1(WPtr<true>{smem_u32(sw)}, WPtr<true>{smem_u32(sw + nWd)});
        else step1(WPtr<false>{a.Wd[0]}, WPtr<false>{a.Wd[1]});
        __syncthreads();
        SCORE_STAMP(5);
        // step 2: d[p][u][:] = cg(a[u], t_p[u])  ->  sd0 [112] , sd1 [192][3] in i_out order
        //   lo=0 block: [p0 (M0) | p3 (M1) | p7 (M2)] ; lo=1 block: [p1 (M0) | p2 (M1) | p4 (M1) | p5 (M1) | p6 (M2) | p8 (M2)]
        const int ntask = 2 * M0 + 4 * M1 + 3 * M2;
        // one thread per (path, u): the task is decoded once and applied to the 2 QB (product, row) copies
        for (int r = tid; r < ntask; r += blockDim.x) {
            int p, u;
            if (r < M0) { p = 0; u = r; }
            else if (r < 2 * M0) { p = 1; u = r - M0; }
            else if (r < 2 * M0 + 4 * M1) { p = 2 + (r - 2 * M0) / M1; u = (r - 2 * M0) % M1; }
            else { p = 6 + (r - 2 * M0 - 4 * M1) / M2; u = (r - 2 * M0 - 4 * M1) % M2; }
#pragma unroll 2
            for (int wq = 0; wq < 2 * QB; ++wq) {             // wq = which * QB + qq
            const float* xa = sa + (wq % QB) * F;
            const float* tt = st + wq * TT;
            float* d0 = sd0 + wq * D0;
            float* d1 = sd1 + wq * 3 * D1;
            float o[3];
            switch (p) {
                case 0: d0[u] = xa[u] * tt[toff[0] + u]; break;
                case 1: { const float x = xa[u]; const float* y = tt + toff[1] + 3 * u;
                          d1[u] = x * y[0]; d1[D1 + u] = x * y[1]; d1[2 * D1 + u] = x * y[2]; } break;
                case 2: { const float* x = xa + boff[1] + 3 * u; const float y = tt[toff[2] + u];
                          float* d = d1 + M0 + u; d[0] = x[0] * y; d[D1] = x[1] * y; d[2 * D1] = x[2] * y; } break;
                case 3: cg_110(xa + boff[1] + 3 * u, tt + toff[3] + 3 * u, o); d0[M0 + u] = o[0]; break;
                case 4: { cg_111(xa + boff[1] + 3 * u, tt + toff[4] + 3 * u, o);
                          float* d = d1 + M0 + M1 + u; d[0] = o[0]; d[D1] = o[1]; d[2 * D1] = o[2]; } break;
                case 5: { cg_121(xa + boff[1] + 3 * u, tt + toff[5] + 5 * u, o);
                          float* d = d1 + M0 + 2 * M1 + u; d[0] = o[0]; d[D1] = o[1]; d[2 * D1] = o[2]; } break;
                case 6: { cg_211(xa + boff[2] + 5 * u, tt + toff[6] + 3 * u, o);
                          float* d = d1 + M0 + 3 * M1 + u; d[0] = o[0]; d[D1] = o[1]; d[2 * D1] = o[2]; } break;
                case 7: cg_220(xa + boff[2] + 5 * u, tt + toff[7] + 5 * u, o); d0[M0 + M1 + u] = o[0]; break;
                default: { cg_221(xa + boff[2] + 5 * u, tt + toff[8] + 5 * u, o);
                           float* d = d1 + M0 + 3 * M1 + M2 + u; d[0] = o[0]; d[D1] = o[1]; d[2 * D1] = o[2]; } break;
            }
            }
        }
        __syncthreads();
        SCORE_STAMP(6);
        // step 3: linear  (D0 -> 1+NV scalars with bias ; D1 -> NV vectors).  One thread per (which, scalar output) and one per
        // (which, vector CHANNEL): the three components of a vector output share their weight column, so the thread that owns the
        // channel loads each weight once and every activation float4 is a warp-wide broadcast -- with one thread per component the
        // phase was bound by shared-memory wavefronts (20 per warp and batch of 8 rows, three-way bank conflicts between the
        // component planes: ~10 k cycles at QB = 2).  Scalar tasks sit in threads 0 .. 2 NS - 1, vector tasks start at a warp
        // boundary (no warp runs both bodies).  Same summation order per output as before (K ascending).
        auto step3 = [&](auto wl0_0, auto wl0_1, auto wl1_0, auto wl1_1) {
        const int NS = 1 + NV, v0 = (2 * NS + 31) & ~31;
        for (int i = tid; i < v0 + 2 * NV; i += blockDim.x) {
            if (i < 2 * NS) {
                const int which = i / NS, o = i % NS;
                float acc[QB];
                const auto W = (which ? wl0_1 : wl0_0) + o;
                const float* S = sd0 + (size_t)(which * QB) * D0;
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) acc[qq] = a.bl[which][o];
                int r = 0;
                SCORE_STAMP(11);
                for (; r + 8 <= D0; r += 8) {
                    float wv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) wv[k] = W[(r + k) * NS];
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq) {
                        const float4 s0 = *reinterpret_cast<const float4*>(S + qq * D0 + r);
                        const float4 s1 = *reinterpret_cast<const float4*>(S + qq * D0 + r + 4);
                        float t = acc[qq];
                        t = fmaf(s0.x, wv[0], t); t = fmaf(s0.y, wv[1], t); t = fmaf(s0.z, wv[2], t); t = fmaf(s0.w, wv[3], t);
                        t = fmaf(s1.x, wv[4], t); t = fmaf(s1.y, wv[5], t); t = fmaf(s1.z, wv[6], t); t = fmaf(s1.w, wv[7], t);
                        acc[qq] = t;
                    }
                }
                for (; r < D0; r += 4) {
                    float wv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) wv[k] = W[(r + k) * NS];
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq) {
                        const float4 s0 = *reinterpret_cast<const float4*>(S + qq * D0 + r);
                        float t = acc[qq];
                        t = fmaf(s0.x, wv[0], t); t = fmaf(s0.y, wv[1], t); t = fmaf(s0.z, wv[2], t); t = fmaf(s0.w, wv[3], t);
                        acc[qq] = t;
                    }
                }
                SCORE_STAMP(12);
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) sy[(which * QB + qq) * NY + o] = acc[qq];
            } else if (i >= v0) {
                const int which = (i - v0) / NV, c = (i - v0) % NV;
                float acc[QB][3];
#pragma unroll
                for (int qq = 0; qq < QB; ++qq) acc[qq][0] = acc[qq][1] = acc[qq][2] = 0.f;
                const auto W = (which ? wl1_1 : wl1_0) + c;
                const float* S = sd1 + (size_t)(which * QB) * 3 * D1;          // [qq][k3][D1]
                int r = 0;
                for (; r + 8 <= D1; r += 8) {
                    float wv[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) wv[k] = W[(r + k) * NV];
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq)
#pragma unroll
                        for (int k3 = 0; k3 < 3; ++k3) {
                            const float4 s0 = *reinterpret_cast<const float4*>(S + (qq * 3 + k3) * D1 + r);
                            const float4 s1 = *reinterpret_cast<const float4*>(S + (qq * 3 + k3) * D1 + r + 4);
                            float t = acc[qq][k3];
                            t = fmaf(s0.x, wv[0], t); t = fmaf(s0.y, wv[1], t); t = fmaf(s0.z, wv[2], t); t = fmaf(s0.w, wv[3], t);
                            t = fmaf(s1.x, wv[4], t); t = fmaf(s1.y, wv[5], t); t = fmaf(s1.z, wv[6], t); t = fmaf(s1.w, wv[7], t);
                            acc[qq][k3] = t;
                        }
                }
                for (; r < D1; r += 4) {
                    float wv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) wv[k] = W[(r + k) * NV];
#pragma unroll
                    for (int qq = 0; qq < QB; ++qq)
#pragma unroll
                        for (int k3 = 0; k3 < 3; ++k3) {
                            const float4 s0 = *reinterpret_cast<const float4*>(S + (qq * 3 + k3) * D1 + r);
                            float t = acc[qq][k3];
                            t = fmaf(s0.x, wv[0], t); t = fmaf(s0.y, wv[1], t); t = fmaf(s0.z, wv[2], t); t = fmaf(s0.w, wv[3], t);
                            acc[qq][k3] = t;
                        }
                }
#pragma unroll
                for (int qq = 0; qq < QB; ++qq)
#pragma unroll
                    for (int k3 = 0; k3 < 3; ++k3) sy[(which * QB + qq) * NY + NS + 3 * c + k3] = acc[qq][k3];
            }
        }
        };
        if (w_smem > 1) step3(WPtr<true>{smem_u32(Wl0[0])}, WPtr<true>{smem_u32(Wl0[1])}, WPtr<true>{smem_u32(Wl1[0])}, WPtr<true>{smem_u32(Wl1[1])});
        else step3(WPtr<false>{Wl0[0]}, WPtr<false>{Wl0[1]}, WPtr<false>{Wl1[0]}, WPtr<false>{Wl1[1]});
        SCORE_STAMP(10);
        __syncthreads();
        SCORE_STAMP(7);
        // step 4: gate, mean over the NV vectors (drop the scalar)
        // 8 lanes per (product, row, component), each NV / 8 vectors apart, folded with shuffles (2 QB 3 8 is a multiple of 32:
        // whole warps enter or skip the loop)
        for (int i = tid; i < 2 * QB * 3 * 8; i += blockDim.x) {
            const int part = i & 7, o3 = i >> 3, wq = o3 / 3, k3 = o3 % 3, which = wq / QB, qq = wq % QB;
            const float* y = sy + wq * NY;
            float s = 0.f;
            for (int c = part; c < NV; c += 8) s += y[1 + NV + 3 * c + k3] * (kCSigmoid * sigmoidf_(y[1 + c]));
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (part == 0 && qq < nq) sres[((q0 + qq) * 2 + which) * 3 + k3] = s / (float)NV;
        }
    }
    __syncthreads();
    SCORE_STAMP(8);
    if (tid < np) {
        const int t = t0 + tid;
        const float* T = a.Ts + (size_t)t * 7;
        const float qinv[4] = {T[0], -T[1], -T[2], -T[3]};
        float lin[3] = {0.f, 0.f, 0.f}, ang[3] = {0.f, 0.f, 0.f};
        for (int q = 0; q < a.n_q; ++q) {
            float l[3], s[3];
            quat_apply<float>(qinv, sres + ((tid * a.n_q + q) * 2 + 0) * 3, l);
            quat_apply<float>(qinv, sres + ((tid * a.n_q + q) * 2 + 1) * 3, s);
            const float px = a.qx[3 * q] / a.lin_mult, py = a.qx[3 * q + 1] / a.lin_mult, pz = a.qx[3 * q + 2] / a.lin_mult;
            const float ox = py * l[2] - pz * l[1], oy = pz * l[0] - px * l[2], oz = px * l[1] - py * l[0];
            const float w = a.qw[q];
            lin[0] += w * l[0]; lin[1] += w * l[1]; lin[2] += w * l[2];
            ang[0] += w * (ox + s[0]); ang[1] += w * (oy + s[1]); ang[2] += w * (oz + s[2]);
        }
        for (int i = 0; i < 3; ++i) { a.lin_out[(size_t)t * 3 + i] = lin[i]; a.ang_out[(size_t)t * 3 + i] = ang[i]; }
        if (a.T64 && step_now < a.n_steps) {       // (a replay past the end of the schedule is a no-op, not an out-of-bounds row)
            // fused Langevin step of pose t (pose_update_kernel + cast_pose_kernel); every CTA read `step_now` before any
            // CTA can advance the counter (the last ticket holder does, below)
            double* Td = a.T64 + (size_t)t * 7;
            if (pb <= 8) {
                langevin_apply(Td, ang, lin, s_z[tid], s_coef);
            } else {
                const double* row = a.sched + (size_t)step_now * 4;
                langevin_step(Td, ang, lin, a.noise ? a.noise + ((size_t)step_now * a.n_t + t) * 6 : nullptr,
                              a.seed_dev ? *a.seed_dev : a.seed, (unsigned long long)t,
                              (unsigned long long)step_now, row[0], a.ang_mult_d, a.lin_mult_d, row[1], row[2], row[3]);
            }
            if (a.traj) for (int k = 0; k < 7; ++k) a.traj[((size_t)(step_now + 1) * a.n_t + t) * 7 + k] = Td[k];
            for (int k = 0; k < 7; ++k) a.T32[(size_t)t * 7 + k] = (float)Td[k];
        }
    }
    }   // poses of this CTA
    SCORE_STAMP(9);
    if (a.T64) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) { *a.ticket = 0; __threadfence(); *a.counter = step_now + 1; }
        }
    }
}

#ifdef DEDF_SCORE_TRACE
}  // namespace dedf
extern "C" int dedf_score_trace(long long* host_out16) {
    return cudaMemcpyFromSymbol(host_out16, dedf::g_score_trace, sizeof(long long) * 16) == cudaSuccess ? DEDF_OK : DEDF_ERR_LAUNCH;
}
namespace dedf {
#endif

// ---------------------------------------------------------------------------
// pose update (float64)
// ---------------------------------------------------------------------------
// Device-resident schedule of a denoise loop: one row per step, advanced by sample_advance_kernel so that a whole step
// (score head + pose update) can be replayed as a CUDA graph without host-side parameters.
//   row = [t, alpha_ang, alpha_lin, temperature]
struct SampleState {
    const double* sched; int n_steps;   // (n_steps, 4)
    int* counter;                       // current step (device)
    float* time_out;                    // (1) fp32 time of the current step, read by time_embed
    double* cur;                        // (4) current row, read by pose_update
    // optional: time rows of EVERY step, precomputed in one dedf_time_embed launch before the loop (the schedule is known up front)
    const float* rows_all;              // (n_scales, n_steps, K)
    float* rows_cur;                    // (n_scales, 1, K): this step's rows, read by the edge MLP as its per-pose bias
    int n_scales, K;
};

__global__ void sample_advance_kernel(SampleState s) {
    const int i = min(*s.counter, s.n_steps - 1);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) s.cur[k] = s.sched[(size_t)i * 4 + k];
        s.time_out[0] = (float)s.sched[(size_t)i * 4];
    }
    if (s.rows_all)
        for (int j = threadIdx.x; j < s.n_scales * s.K; j += blockDim.x)
            s.rows_cur[j] = s.rows_all[((size_t)(j / s.K) * s.n_steps + i) * s.K + (j % s.K)];
}

struct PoseArgs {
    double* T; int n_t;                 // (n_t,7) in place
    const float* ang; const float* lin; // dimensionless scores (n_t,3)
    const double* noise;                // (n_t,6) standard normals (ang, lin) or null -> Philox
    unsigned long long seed, offset;    // Philox: (seed, subsequence = pose); `offset` = STEP index, each step owns a disjoint
                                        // window of kPhiloxPerStep 32-bit outputs of the pose's stream (see pose_update_kernel)
    double t, ang_mult, lin_mult, alpha_ang, alpha_lin, temperature;
    double* traj_out;                   // optional (n_t,7) copy of the new pose
    const double* dev_row;              // optional device row [t, alpha_ang, alpha_lin, temperature] overriding the host values
    int* dev_counter;                   // optional device step counter: Philox offset, trajectory row (counter+1), incremented here
};

__global__ void pose_update_kernel(PoseArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_t) return;
    if (a.dev_row) { a.t = a.dev_row[0]; a.alpha_ang = a.dev_row[1]; a.alpha_lin = a.dev_row[2]; a.temperature = a.dev_row[3]; }
    if (a.dev_counter) {
        const int step = *a.dev_counter;
        a.offset = (unsigned long long)step;
        if (a.noise) a.noise += (size_t)step * a.n_t * 6;
        if (a.traj_out) a.traj_out += (size_t)(step + 1) * a.n_t * 7;
    }
    double* T = a.T + (size_t)i * 7;
    langevin_step(T, a.ang + (size_t)i * 3, a.lin + (size_t)i * 3, a.noise ? a.noise + (size_t)i * 6 : nullptr, a.seed,
                  (unsigned long long)i, a.offset, a.t, a.ang_mult, a.lin_mult, a.alpha_ang, a.alpha_lin, a.temperature);
    if (a.traj_out) for (int k = 0; k < 7; ++k) a.traj_out[(size_t)i * 7 + k] = T[k];
}

__global__ void cast_pose_kernel(const double* __restrict__ T, int n, float* __restrict__ out, int* counter) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)T[i];
    if (counter && i == 0) *counter += 1;      // runs after pose_update_kernel on the same stream
}

// EbmScoreModelHead.compute_energy tail (score_head_ebm.py:171-172): one CTA per pose
__global__ void __launch_bounds__(128) ebm_energy_kernel(const float* __restrict__ key_f, const float* __restrict__ query_f,
                                                        const float* __restrict__ qw, int n_q, int F, float scale, float* __restrict__ out) {
    pdl_wait(); pdl_launch();
    __shared__ float red[4];
    const int t = blockIdx.x, tid = threadIdx.x;
    float acc = 0.f;
    for (int q = 0; q < n_q; ++q) {
        const size_t row = ((size_t)t * n_q + q) * F;
        float s = 0.f;
        for (int c = tid; c < F; c += blockDim.x) { const float d = key_f[row + c] - query_f[row + c]; s = fmaf(d, d, s); }
        acc = fmaf(qw[q], s, acc);
    }
    acc = warp_sum(acc);
    if ((tid & 31) == 0) red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) out[t] = (red[0] + red[1] + red[2] + red[3]) * scale;
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_time_embed(const dedf_time_desc* d, const float* time, int n_t, float* out, cudaStream_t stream) {
    if (!d || !time || !out || d->n_scales < 1 || d->n_scales > DEDF_MAX_SCALES) return DEDF_ERR_ARG;
    if (n_t <= 0) return DEDF_OK;
    TimeArgs a{};
    a.time = time; a.n_t = n_t; a.max_time = d->max_time; a.enc_n = d->enc_n; a.enc_dim = d->enc_dim; a.enc_freq = d->enc_freq;
    if (!a.enc_freq) return DEDF_ERR_ARG;
    a.h_dim = d->h_dim; a.e_dim = d->e_dim; a.out_dim = d->out_dim; a.n_scales = d->n_scales; a.out = out;
    for (int s = 0; s < d->n_scales; ++s) {
        a.W1[s] = d->W1[s]; a.b1[s] = d->b1[s]; a.W2[s] = d->W2[s]; a.b2[s] = d->b2[s]; a.Wp[s] = d->Wp[s]; a.bp[s] = d->bp[s];
        if (!a.W1[s] || !a.b1[s] || !a.W2[s] || !a.b2[s] || !a.Wp[s] || !a.bp[s]) return DEDF_ERR_ARG;
    }
    const size_t smem = (size_t)(a.enc_dim + a.h_dim + a.e_dim) * sizeof(float);
    launch_pdl(time_embed_kernel, dim3(dim3(n_t, d->n_scales)), dim3(128), smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_query_transform(const float* Ts, int n_t, const float* qx, const float* qf, int n_q,
                                    const int* irr, float* x_out, float* f_out, cudaStream_t stream) {
    if (!Ts || !qx || !qf || !irr || !x_out || !f_out) return DEDF_ERR_ARG;
    if (n_t <= 0 || n_q <= 0) return DEDF_OK;
    QueryArgs a{Ts, n_t, qx, qf, n_q, Irr{irr[0], irr[1], irr[2]}, x_out, f_out};
    launch_pdl(query_transform_kernel, dim3(n_t), dim3(128), 0, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

static int launch_score_tp(ScoreArgs a, cudaStream_t stream) {
    const int n_t = a.n_t, n_q = a.n_q, n_vec = a.n_vec;
    const int M0 = a.irr.m0, M1 = a.irr.m1, M2 = a.irr.m2, F = a.irr.dim();
    const int tt = M0 + 3 * M0 + M1 + 3 * M1 + 3 * M1 + 5 * M1 + 3 * M2 + 5 * M2 + 5 * M2;
    const int D0 = M0 + M1 + M2, D1 = M0 + 3 * M1 + 2 * M2;
    if ((M0 % 4) || (M1 % 4) || (M2 % 4) || (D0 % 4) || (D1 % 4)) return DEDF_ERR_UNSUPPORTED;
    const int pb = 1, qb = kScoreQB;
    if (1 + 4 * n_vec > tt) return DEDF_ERR_UNSUPPORTED;      // the linear outputs reuse the t buffers
    const size_t act = (size_t)(qb * (2 * F + 2 * (tt + D0 + 3 * D1)) + ((6 * pb * n_q + 3) & ~3) + 36 * pb) * sizeof(float);
    const int nWd = M0 * M0 + M0 * M1 + M1 * M0 + 2 * M1 * M1 + M1 * M2 + M2 * M1 + 2 * M2 * M2;
    const size_t wbytes = (size_t)2 * (nWd + D1 * n_vec) * sizeof(float);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    // Up to two poses per SM: persistent CTAs with the weights resident in shared memory (37 us vs 70 us at 128 poses).
    // Larger batches: one CTA per pose, weights through L1/L2, three CTAs per SM hide each other's latency (139 us at 1024
    // poses; the resident variant runs one 12-warp CTA per SM and takes 181 us).
    const bool w_smem = n_t <= 2 * kNumSMs && act + wbytes <= 226 * 1024 && al16(a.Wd[0]) && al16(a.Wd[1]) && al16(a.Wl1[0]) && al16(a.Wl1[1]) &&
                        (nWd % 4 == 0) && ((D1 * n_vec) % 4 == 0) && !getenv("DEDF_NO_SMEM_SCORE");
    // More than two poses per SM: 8 (pose, query node) rows per weight pass.  Every weight load then serves 8 rows of both tensor
    // products, so the L2 stream of the 218 KB of weights -- what bounds the one-pose-per-CTA form -- shrinks 4x.
    if (n_t > 2 * kNumSMs && !getenv("DEDF_SCORE_QB2")) {
        constexpr int QB8 = 8;
        const int pb8 = QB8 / n_q > 1 ? QB8 / n_q : 1;
        const size_t act8 = (size_t)(QB8 * (2 * F + 2 * (tt + D0 + 3 * D1)) + ((6 * pb8 * n_q + 3) & ~3) + 36 * pb8) * sizeof(float);
        if (act8 <= 226 * 1024) {
            static bool done8 = false;
            if (!done8) { cudaFuncSetAttribute(score_tp_kernel<QB8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); done8 = true; }
            launch_pdl(score_tp_kernel<QB8>, dim3((n_t + pb8 - 1) / pb8), dim3(kScoreThreads), act8, stream, a, pb8, 0);
            DEDF_CHECK_LAUNCH();
            return DEDF_OK;
        }
    }
    const size_t wl0_bytes = (size_t)2 * D0 * (1 + n_vec) * sizeof(float);
    const bool wl0_smem = w_smem && act + wbytes + wl0_bytes <= 226 * 1024 && al16(a.Wl0[0]) && al16(a.Wl0[1]) && ((D0 * (1 + n_vec)) % 4 == 0);
    const size_t smem = act + (w_smem ? wbytes : 0) + (wl0_smem ? wl0_bytes : 0);
    if (smem > 226 * 1024) return DEDF_ERR_UNSUPPORTED;
    static bool done = false;
    if (!done) { cudaFuncSetAttribute(score_tp_kernel<kScoreQB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024); done = true; }
    if (w_smem) launch_pdl(score_tp_kernel<kScoreQB>, dim3(grid_for(n_t, pb, kNumSMs)), dim3(kScoreThreads), smem, stream, a, pb, wl0_smem ? 2 : 1);
    else launch_pdl(score_tp_kernel<kScoreQB>, dim3((n_t + pb - 1) / pb), dim3(256), smem, stream, a, pb, 0);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_score_tp(const float* Ts, int n_t, const float* qf_rot, const float* key_f, const float* qx,
                             const float* qw, int n_q, const int* irr, const float* const* Wd, const float* const* Wl0,
                             const float* const* Wl1, const float* const* bl, int n_vec, float lin_mult,
                             float* ang_out, float* lin_out, cudaStream_t stream) {
    if (!Ts || !qf_rot || !key_f || !qx || !qw || !irr || !Wd || !Wl0 || !Wl1 || !bl || !ang_out || !lin_out) return DEDF_ERR_ARG;
    if (n_t <= 0) return DEDF_OK;
    ScoreArgs a{};
    a.Ts = Ts; a.n_t = n_t; a.qf_rot = qf_rot; a.key_f = key_f; a.qx = qx; a.qw = qw; a.n_q = n_q;
    a.irr = Irr{irr[0], irr[1], irr[2]};
    for (int i = 0; i < 2; ++i) { a.Wd[i] = Wd[i]; a.Wl0[i] = Wl0[i]; a.Wl1[i] = Wl1[i]; a.bl[i] = bl[i];
                                  if (!Wd[i] || !Wl0[i] || !Wl1[i] || !bl[i]) return DEDF_ERR_ARG; }
    a.n_vec = n_vec; a.lin_mult = lin_mult; a.ang_out = ang_out; a.lin_out = lin_out;
    return launch_score_tp(a, stream);
}

extern "C" int dedf_score_tp_step(const dedf_score_step_desc* d, cudaStream_t stream) {
    if (!d || !d->Ts || !d->qf || !d->key_f || !d->qx || !d->qw || !d->ang_out || !d->lin_out) return DEDF_ERR_ARG;
    if (d->n_t <= 0) return DEDF_OK;
    ScoreArgs a{};
    a.Ts = d->Ts; a.n_t = d->n_t; a.qf_rot = nullptr; a.qf = d->qf; a.key_f = d->key_f; a.qx = d->qx; a.qw = d->qw; a.n_q = d->n_q;
    a.irr = Irr{d->irr[0], d->irr[1], d->irr[2]};
    for (int i = 0; i < 2; ++i) { a.Wd[i] = d->Wd[i]; a.Wl0[i] = d->Wl0[i]; a.Wl1[i] = d->Wl1[i]; a.bl[i] = d->bl[i];
                                  if (!a.Wd[i] || !a.Wl0[i] || !a.Wl1[i] || !a.bl[i]) return DEDF_ERR_ARG; }
    a.n_vec = d->n_vec; a.lin_mult = d->lin_mult; a.ang_out = d->ang_out; a.lin_out = d->lin_out;
    if (d->T64) {           // fused Langevin step
        if (!d->sched || d->n_steps < 1 || !d->counter || !d->T32 || !d->ticket) return DEDF_ERR_ARG;
        if (d->T32 != d->Ts) return DEDF_ERR_ARG;       // the fp32 copy the network reads IS the one the step refreshes
        a.T64 = d->T64; a.sched = d->sched; a.n_steps = d->n_steps; a.counter = d->counter; a.noise = d->noise; a.seed = d->seed;
        a.seed_dev = d->seed_dev;
        a.ang_mult_d = d->ang_mult; a.lin_mult_d = d->lin_mult_d; a.traj = d->traj; a.T32 = d->T32; a.ticket = d->ticket;
    }
    return launch_score_tp(a, stream);
}

extern "C" int dedf_pose_update(double* T, int n_t, const float* ang, const float* lin, const double* noise,
                                unsigned long long seed, unsigned long long offset, double t, double ang_mult,
                                double lin_mult, double alpha_ang, double alpha_lin, double temperature,
                                double* traj_out, float* T_f32_out, const double* dev_row, int* dev_counter,
                                cudaStream_t stream) {
    if (!T || !ang || !lin) return DEDF_ERR_ARG;
    if (dev_counter && !T_f32_out) return DEDF_ERR_ARG;     // the counter is advanced by the cast kernel
    if (n_t <= 0) return DEDF_OK;
    PoseArgs a{T, n_t, ang, lin, noise, seed, offset, t, ang_mult, lin_mult, alpha_ang, alpha_lin, temperature, traj_out,
               dev_row, dev_counter};
    pose_update_kernel<<<(n_t + 127) / 128, 128, 0, stream>>>(a);
    DEDF_CHECK_LAUNCH();
    if (T_f32_out) {
        cast_pose_kernel<<<(n_t * 7 + 255) / 256, 256, 0, stream>>>(T, n_t * 7, T_f32_out, dev_counter);
        DEDF_CHECK_LAUNCH();
    }
    return DEDF_OK;
}

extern "C" int dedf_sample_advance(const double* sched, int n_steps, int* counter, float* time_out, double* cur_row,
                                   const float* rows_all, float* rows_cur, int n_scales, int k, cudaStream_t stream) {
    if (!sched || !counter || !time_out || !cur_row || n_steps <= 0) return DEDF_ERR_ARG;
    if (rows_all && (!rows_cur || n_scales <= 0 || k <= 0)) return DEDF_ERR_ARG;
    SampleState s{sched, n_steps, counter, time_out, cur_row, rows_all, rows_cur, n_scales, k};
    sample_advance_kernel<<<1, 256, 0, stream>>>(s);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_ebm_energy(const float* key_f, const float* query_f, const float* qw, int n_t, int n_q, int F, float scale,
                               float* out, cudaStream_t stream) {
    if (!key_f || !query_f || !qw || !out || F <= 0) return DEDF_ERR_ARG;
    if (n_t <= 0) return DEDF_OK;
    launch_pdl(ebm_energy_kernel, dim3(n_t), dim3(128), 0, stream, key_f, query_f, qw, n_q, F, scale, out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
