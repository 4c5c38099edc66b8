// edge_tp_act_tc: the attention block's per-edge hot op on the tensor cores.
//
//   message = x_src[src] (+ x_dst[dst]) -> depthwise CG tensor product with the 9 harmonics and the per-edge radial
//   weights (o3.TensorProduct 'uvu', equiformer/tensor_product_rescale.py:352-382) -> block-diagonal LinearRS
//   (sep_alpha | sep_act.lin) -> SmoothLeakyReLU . alpha_dot (+ edge logit) = attention logits, Gate = value rows
//   (graph_attention.py:231-246, fast_activation.py:210-224)
//
// Same arithmetic and outputs as edge_tp_lin_kernel<G, EPI_ACT> (edge.cu); the block-diagonal linear layer (104 kMAC per
// edge, 90 % of the FLOPs) moves from the fp32 FMA pipe to tcgen05.mma with a hi / lo operand split (tc.cuh: kind::f16 with fp16
// halves by default, kind::tf32 with the 3xTF32 split as the F16 = false variant -- see TaStore below), fp32 accumulators in TMEM.
// What makes it fit (sizes below are those of the tf32 variant; the fp16 one packs two chunks into the same bytes):
//
//   * K is consumed in CHUNKS of input channels: chunk j = l=0 channels [8j, 8j+8), l=1 channels [4j, 4j+4), l=2 channels
//     [2j, 2j+2) (8 chunks for 64x0e+32x1e+16x2e, 4 for 32x0e+16x1e+8x2e).  One chunk contributes 14 / 24 / 22 columns
//     to the l_out = 0 / 1 / 2 GEMMs (padded to 16 / 24 / 24); the host packs the weight rows in the same order
//     (layers.pack_tp_act_tc).  The (E, 1568) tensor-product output therefore never exists, not even per tile: a chunk
//     of it (32 edges) is 26 KB per hi / lo part, double-buffered.
//   * the tile's rows are the M side: 32 rows (l_out = 0), 96 rows (m, e) (l_out = 1), 160 rows (l_out = 2, two MMAs);
//     M is always 128 and the rows past the stored ones alias the following shared memory (finite or not, they only
//     reach accumulator rows nobody reads).
//   * the 2e output's m = 4 rows ride in rows 96..127 of the 1e operand and the 1e / 2e weights sit side by side in one
//     B operand (N = 48 / 32): one MMA serves both, the cross terms land in accumulator cells nobody reads.  The 0e GEMM
//     runs with the operands swapped (M = output channels, N = 32 edges) so that its epilogue work spreads over all four
//     TMEM lane quarters instead of one.
//   * warp roles: 6 producer warps (lane = edge; warps 0-3 take one l=1 channel + two l=0 channels of the chunk, warps
//     4-5 one l=2 channel: balanced to 7 %), each thread writing whole float4 K-groups of the chunk-major operand;
//     8 epilogue warps (TMEM -> bias, activations, gates -> staged value rows -> one bulk store per edge row), one TMA
//     weight-ring warp, one MMA issuer (30 tcgen05.mma per chunk); double-buffered accumulators (2 x 128 TMEM columns)
//     so the epilogue of tile t overlaps the production of tile t + 1.
//   * measured (profiles/run_tp_lin.py, 86 k edges, G = 32): 248 us vs 686 us for the fp32 FMA kernel.  The MMA issuer's
//     cycle accounting (dedf_tp_act_tc_set_debug) shows the tensor pipe waiting on shared-memory bandwidth: every
//     M = 128 x K = 8 tf32 MMA fetches 4 KB for the M side whatever N is, the producers store 53 KB per chunk next to it;
//     a second producer group (12 warps) made both slower (275 us).
#include <cuda_fp16.h>
#include "common.cuh"
#include "tc.cuh"
#include "cg_slots.cuh"
#include "../../include/dedf.h"

namespace dedf {

// -DDEDF_TA_TRACE (profiles/run_tp_act_trace.py builds its own copy of the library): producer warp 0 of CTA 0 accumulates the
// cycles of each segment of its chunk loop into dbg[8..15].  Not compiled into the shipped library.
#ifdef DEDF_TA_TRACE
#define TA_T0() long long ta_c = clock64()
#define TA_SEG(i) do { const long long ta_n = clock64(); ta_seg[i] += ta_n - ta_c; ta_c = ta_n; } while (0)
#else
#define TA_T0() do { } while (0)
#define TA_SEG(i) do { } while (0)
#endif

constexpr int kTaTE = 32;
constexpr int kTaProdWarps = 6;
constexpr int kTaTmaWarp = 6, kTaMmaWarp = 7, kTaEpiWarp0 = 8;
constexpr int kTaEpiWarps = 8;                      // two warpgroups: each takes half of the tile's edges / output columns
constexpr int kTaThreads = (kTaEpiWarp0 + kTaEpiWarps) * 32;
constexpr int kTaAccCols = 128;
constexpr int kTaXLd = 36;                          // staged message slice: 8 (0e) + 12 (1e) + 12 (2e, 10 used) floats + 4 pad                     // TMEM columns of one accumulator buffer
constexpr int kTaStages = 2;                        // A chunk stages and weight ring stages
constexpr int kKC0 = 16, kKC1 = 24, kKC2 = 24;      // K columns per chunk of the three GEMMs
constexpr int kR0 = kTaTE, kR1 = 4 * kTaTE, kR2 = 4 * kTaTE;   // stored rows: 0e (e) | 1e (m, e) m < 3 + 2e m = 4 | 2e (m, e) m < 4
constexpr int kA0Off = 0;
constexpr int kA1Off = kA0Off + (kKC0 / 4) * kR0 * 16;          //  2048
constexpr int kA2Off = kA1Off + (kKC1 / 4) * kR1 * 16;          // 11264
constexpr int kAPart = kA2Off + (kKC2 / 4) * kR2 * 16;          // 26624 (hi or lo)
constexpr int kAStage = 2 * kAPart;

template <int G>
struct TaCfg {
    using D = Dtp<G>;
    static constexpr int NCH = D::M0 / 8;                                  // chunks
    static constexpr int MA = D::M0;                                       // alpha channels
    static constexpr int N0 = MA + D::M0 + D::M1 + D::M2;                  // 176 / 88
    static constexpr int N0P = (N0 + 15) / 16 * 16;                        // 176 / 96
    static constexpr int N1P = (D::M1 + 15) / 16 * 16;                     // 32 / 16
    static constexpr int N2P = (D::M2 + 15) / 16 * 16;                     // 16 / 16
    static constexpr int W0Off = 0;
    static constexpr int N1C = N1P + N2P;                                  // combined 1e | 2e (m = 4) GEMM: 48 / 32 columns
    static constexpr int W1Off = W0Off + (kKC0 / 4) * N0P * 16;
    static constexpr int W2Off = W1Off + (kKC1 / 4) * N1C * 16;
    static constexpr int WPart = W2Off + (kKC2 / 4) * N2P * 16;            // 15872 / 9216
    static constexpr int WStage = 2 * WPart;
    static constexpr int MT0 = (N0P + 127) / 128;                          // M tiles of the (swapped) l_out = 0 GEMM
    static constexpr int T0 = 0, T1 = 2 * kTaTE, T2B = T1 + N1P, T2A = T1 + N1C;   // TMEM columns of the accumulators
    static_assert(T2A + N2P <= kTaAccCols && kKC1 == kKC2, "accumulator buffer / combined GEMM");
    static constexpr int LDO = D::F + 4;                                   // staged value row stride (floats)
    static constexpr int NG = D::M1 + D::M2;                               // gates
    // dynamic shared memory
    static constexpr int OffA = 0;
    static constexpr int OffW = OffA + kTaStages * kAStage;
    static constexpr int OffOut = OffW + kTaStages * WStage;
    static constexpr int OffGate = OffOut + kTaTE * LDO * 4;
    static constexpr int NGP = NG + 1;                                     // gate table row stride (odd: conflict-free both ways)
    static constexpr int OffRed = OffGate + kTaTE * NGP * 4;               // [MA][33] attention-logit terms
    static constexpr int OffX = OffRed + MA * 33 * 4;                      // [2][32][kTaXLd] staged message slices
    static constexpr int OffPar = OffX + 2 * kTaTE * kTaXLd * 4;
    static constexpr int Smem = OffPar + (N0P + MA) * 4;
};

struct TpActArgs {
    const float* x_src; const float* x_dst;
    const int* edge_src; const int* edge_dst; const int* n_edges;
    const float* sh; const float* w; long long w_stride; int w_perm;
    const float* Wp;          // packed weight chunks (layers.pack_tp_act_tc)
    const float* bias0; const float* alpha_dot; const float* edge_logit;
    float* logits; float* out;
    long long* dbg;
};

// fp16 operand split (F16 = true): x = h + l with h = fp16(x), l = fp16(x - h) (22 significand bits between them, like the 21 of
// the tf32 split; the lo part of a value below ~0.1 is an fp16 subnormal, i.e. an ABSOLUTE error of <= 3e-8 per operand element),
// products on kind::f16 with K = 16 per MMA: two channel chunks share one operand stage (the 16-byte K-group of a row holds four
// halves of the even chunk and four of the odd one), so per chunk the producers store, and the tensor core fetches, HALF the
// shared-memory bytes of the tf32 split -- and shared-memory bandwidth is what bounds this kernel (clock64 trace of the producers
// and the issuer, profiles/r2_s11_tp_act_trace.txt: 150 KB of operand fetch + ~1000 LSU wavefronts per 2750-cycle chunk).
// Limit: |operand| must stay below 65504 (fp16); larger values come out as NaN, never silently wrong.  DEDF_TPACT_F16=0 on the
// host side selects the tf32 split.
template <bool F16> struct TaStore;
template <> struct TaStore<false> {
    static __device__ __forceinline__ void put4(unsigned char* hi, unsigned char* lo, int goff, int, float v0, float v1, float v2, float v3) {
        const float4 h = make_float4(tc::tf32_hi(v0), tc::tf32_hi(v1), tc::tf32_hi(v2), tc::tf32_hi(v3));
        *reinterpret_cast<float4*>(hi + goff) = h;
        *reinterpret_cast<float4*>(lo + goff) = make_float4(v0 - h.x, v1 - h.y, v2 - h.z, v3 - h.w);
    }
    static __device__ __forceinline__ void put2(unsigned char* hi, unsigned char* lo, int goff, int, int col, float v0, float v1) {
        const float2 h = make_float2(tc::tf32_hi(v0), tc::tf32_hi(v1));
        *reinterpret_cast<float2*>(hi + goff + 4 * col) = h;
        *reinterpret_cast<float2*>(lo + goff + 4 * col) = make_float2(v0 - h.x, v1 - h.y);
    }
    static __device__ __forceinline__ void put1(unsigned char* hi, unsigned char* lo, int goff, int, int col, float v) {
        const float h = tc::tf32_hi(v);
        *reinterpret_cast<float*>(hi + goff + 4 * col) = h;
        *reinterpret_cast<float*>(lo + goff + 4 * col) = v - h;
    }
};
template <> struct TaStore<true> {      // odd: 0 / 1 = which half of the 16-byte K-group this chunk fills
    static __device__ __forceinline__ void split2(float v0, float v1, __half2& h, __half2& l) {
        h = __floats2half2_rn(v0, v1);
        const float2 f = __half22float2(h);
        l = __floats2half2_rn(v0 - f.x, v1 - f.y);
    }
    static __device__ __forceinline__ void put4(unsigned char* hi, unsigned char* lo, int goff, int odd, float v0, float v1, float v2, float v3) {
        __half2 h0, l0, h1, l1;
        split2(v0, v1, h0, l0); split2(v2, v3, h1, l1);
        uint2 H, Lo;
        H.x = *reinterpret_cast<uint32_t*>(&h0); H.y = *reinterpret_cast<uint32_t*>(&h1);
        Lo.x = *reinterpret_cast<uint32_t*>(&l0); Lo.y = *reinterpret_cast<uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(hi + goff + 8 * odd) = H;
        *reinterpret_cast<uint2*>(lo + goff + 8 * odd) = Lo;
    }
    static __device__ __forceinline__ void put2(unsigned char* hi, unsigned char* lo, int goff, int odd, int col, float v0, float v1) {
        __half2 h, l;
        split2(v0, v1, h, l);
        *reinterpret_cast<__half2*>(hi + goff + 8 * odd + 2 * col) = h;
        *reinterpret_cast<__half2*>(lo + goff + 8 * odd + 2 * col) = l;
    }
    static __device__ __forceinline__ void put1(unsigned char* hi, unsigned char* lo, int goff, int odd, int col, float v) {
        const __half h = __float2half_rn(v);
        *reinterpret_cast<__half*>(hi + goff + 8 * odd + 2 * col) = h;
        *reinterpret_cast<__half*>(lo + goff + 8 * odd + 2 * col) = __float2half_rn(v - __half2float(h));
    }
};

// One tile's worth of a producer warp's work.  pr = 0..5: the warp's role (0-3: two l=0 channels + one l=1 channel of the chunk,
// 4-5: one l=2 channel); grp / NGRP / MASK: the warp's producer group handles the chunks whose bit is set in MASK.
//   NGRP = 1: one group of 6 warps takes every chunk, the x-slice staging buffer is double-buffered by chunk parity, one named
//             barrier per chunk; an fp16 stage is complete after the odd chunk of its pair (6 arrivals).
//   NGRP = 2 (fp16 split only): group 0 = warps 0-5, group 1 = the epilogue warps 8-13, which produce their share of a tile before
//             they run its epilogue.  A chunk fills one HALF of an operand stage (even / odd chunk of the pair), whichever group it
//             belongs to: every half ends with a fence + 6 arrivals, a stage is complete at 12.  The stage and its mbarrier parity
//             follow from the running pair count (pair_base + j / 2).  Each group owns ONE staging buffer (index grp), so a second
//             barrier separates the reads of a chunk's slice from staging the next.
// Every load below has lane = edge, i.e. 32 different rows per instruction: what bounds the producers is the NUMBER of gather
// instructions.  The chunk's slice of the message row (x_src[src] + x_dst[dst]: 8 + 12 + 10 floats) is therefore fetched as
// 8 float4 slices spread over the 6 warps, staged in shared memory and read back by the warp that needs it; the per-edge radial
// weights come as 3 float4 / 3 float2 per lane when the caller permuted their columns chunk-major (w_perm).
template <int G, bool F16, int NGRP, uint32_t MASK>
__device__ __forceinline__ void ta_produce_tile(const TpActArgs& a, int E, int tile, int pr, int grp, int lane, unsigned char* sA,
                                                float* s_x, uint64_t* fullA, uint64_t* emptyA, uint32_t& st, uint32_t& ph, uint32_t pair_base) {
    using C = TaCfg<G>;
    using D = Dtp<G>;
    using S = TaStore<F16>;
    constexpr int NCH = C::NCH;
    static_assert(NGRP == 1 || F16, "two producer groups fill the two halves of an fp16 stage");
    const int e = lane;
    auto prod_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(2 + grp), "n"(kTaProdWarps * 32) : "memory"); };
    auto xbuf = [&](int j) { return NGRP == 1 ? (j & 1) : grp; };
    static_assert(MASK != 0 && (MASK >> NCH) == 0, "chunk mask");
    const int j_first = __ffs((int)MASK) - 1;
    auto j_next = [&](int j) { const uint32_t rest = MASK >> (j + 1); return rest ? j + __ffs((int)rest) : NCH; };
    static_assert(kTaStages == 2, "stage = pair count & 1");
    if (NGRP == 2) { st = 0; ph = 0; }      // (unused: the stage of a chunk follows from the pair count)
    auto stage_of = [&](int j) { return NGRP == 2 ? ((pair_base + (uint32_t)(j >> 1)) & 1u) : st; };
    auto wait_par = [&](int j) { return NGRP == 2 ? ((((pair_base + (uint32_t)(j >> 1)) >> 1) & 1u) ^ 1u) : (ph ^ 1u); };
    const int sliceA = pr, sliceB = (pr < 2) ? 6 + pr : -1;
    auto slice_off = [&](int j, int sl) {
        return (sl < 2) ? 8 * j + 4 * sl : (sl < 5) ? D::M0 + 12 * j + 4 * (sl - 2) : D::M0 + 3 * D::M1 + 10 * j - 2 * (j & 1) + 4 * (sl - 5);
    };
    const int e0 = tile * kTaTE;
    const bool ok = e0 + e < E;
    const int eg = ok ? e0 + e : E - 1;
    const int src = a.edge_src[eg], dst = a.edge_dst[eg];
    float sh[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) sh[i] = ok ? a.sh[(size_t)eg * 9 + i] : 0.f;
    const float* xs = a.x_src + (size_t)src * D::F;
    const float* xd = a.x_dst ? a.x_dst + (size_t)dst * D::F : nullptr;
    const float* wr = a.w + (size_t)eg * a.w_stride;
    const float okf = ok ? 1.f : 0.f;
    // the source and destination halves of a message slice stay in separate registers until stx() adds them: the adds (the
    // first USE of the gathered values) then sit a whole chunk of math behind the loads instead of right after them
    // (ncu source view: the producers' long-scoreboard samples were on exactly those adds)
    float4 xa, xb = make_float4(0.f, 0.f, 0.f, 0.f), xad = make_float4(0.f, 0.f, 0.f, 0.f), xbd = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ldx = [&](int j) {
        const int oa = slice_off(j, sliceA);
        xa = *reinterpret_cast<const float4*>(xs + oa);
        if (xd) xad = *reinterpret_cast<const float4*>(xd + oa);
        if (sliceB >= 0) {
            const int ob = slice_off(j, sliceB);
            xb = *reinterpret_cast<const float4*>(xs + ob);
            if (xd) xbd = *reinterpret_cast<const float4*>(xd + ob);
        }
    };
    auto stx = [&](int buf) {
        float* row = s_x + (buf * kTaTE + e) * kTaXLd;
        *reinterpret_cast<float4*>(row + 4 * sliceA) = make_float4((xa.x + xad.x) * okf, (xa.y + xad.y) * okf, (xa.z + xad.z) * okf, (xa.w + xad.w) * okf);
        if (sliceB >= 0) *reinterpret_cast<float4*>(row + 4 * sliceB) = make_float4((xb.x + xbd.x) * okf, (xb.y + xbd.y) * okf, (xb.z + xbd.z) * okf, (xb.w + xbd.w) * okf);
    };
    // stage bookkeeping: with one group an fp16 stage is complete after the odd chunk of a pair; with two groups every chunk of a
    // group is its half of a new stage
    auto stage_open = [&](int j) { return !F16 || NGRP == 2 || !(j & 1); };
    auto stage_close = [&](int j) { return !F16 || NGRP == 2 || (j & 1); };
    if (pr < 4) {
        const int i = pr;
#ifdef DEDF_TA_TRACE
        long long ta_seg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#endif
        float2 w0, w1, w2; float w6[6];
        auto ldw = [&](int j) {
            if (a.w_perm) {
                const float4* q = reinterpret_cast<const float4*>(wr + 60 * j + 12 * i);
                const float4 q0 = q[0], q1 = q[1], q2 = q[2];
                w0 = make_float2(q0.x, q0.y); w1 = make_float2(q0.z, q0.w); w2 = make_float2(q1.x, q1.y);
                w6[0] = q1.z; w6[1] = q1.w; w6[2] = q2.x; w6[3] = q2.y; w6[4] = q2.z; w6[5] = q2.w;
            } else {
                const int ch = 8 * j + 2 * i, p = 4 * j + i;
                w0 = *reinterpret_cast<const float2*>(wr + D::W_K0 + ch);
                w1 = *reinterpret_cast<const float2*>(wr + D::W_K1 + ch);
                w2 = *reinterpret_cast<const float2*>(wr + D::W_K2 + ch);
#pragma unroll
                for (int k = 0; k < 6; ++k) w6[k] = wr[D::W_K3 + p + k * D::M1];
            }
        };
        TA_T0();
        ldx(j_first); ldw(j_first); stx(xbuf(j_first));
        prod_sync();
        TA_SEG(0);
#pragma unroll 1
        for (int j = j_first; j < NCH; j = j_next(j)) {
            const int jn = j_next(j);
            const float* xrow = s_x + (xbuf(j) * kTaTE + e) * kTaXLd;
            const float2 xab = *reinterpret_cast<const float2*>(xrow + 2 * i);
            float x1[3] = {xrow[8 + 3 * i], xrow[9 + 3 * i], xrow[10 + 3 * i]};
            if (NGRP == 2) prod_sync();                                 // every warp of the group has read the (single) staging buffer
            if (jn < NCH) ldx(jn);                                      // next chunk's gathers fly during the math and the stores
            float oa[9], ob[9], o[20];
            dtp_l0(xab.x, w0.x, w1.x, w2.x, sh, oa);
            dtp_l0(xab.y, w0.y, w1.y, w2.y, sh, ob);
            dtp_l1(x1, w6, sh, o);
            TA_SEG(1);
            if (jn < NCH) ldw(jn);
            const int odd = F16 ? (j & 1) : 0;
            const uint32_t sj = stage_of(j);
            if (stage_open(j)) tc::mbar_wait_bounded(&emptyA[sj], wait_par(j));
            TA_SEG(2);
            unsigned char* hi = sA + sj * kAStage;
            unsigned char* lo = hi + kAPart;
            // l_out = 0, group i: [k0 a, k0 b, k4 p, (k12: the l=2 warps)]
            {
                const int off = kA0Off + (i * kR0 + e) * 16;
                S::put2(hi, lo, off, odd, 0, oa[0], ob[0]); S::put1(hi, lo, off, odd, 2, o[3]);
            }
#pragma unroll
            for (int m = 0; m < 3; ++m) {      // l_out = 1, group i: [k3, k5, k7, k1 a]; group 4 column i: k1 b
                S::put4(hi, lo, kA1Off + (i * kR1 + m * kTaTE + e) * 16, odd, o[m], o[4 + m], o[12 + m], oa[1 + m]);
                S::put1(hi, lo, kA1Off + (4 * kR1 + m * kTaTE + e) * 16, odd, i, ob[1 + m]);
            }
#pragma unroll
            for (int m = 0; m < 5; ++m)        // l_out = 2, group i: [k2 a, k2 b, k6, k8]; m = 4 rides in rows 96.. of the 1e operand
                S::put4(hi, lo, m < 4 ? kA2Off + (i * kR2 + m * kTaTE + e) * 16 : kA1Off + (i * kR1 + 3 * kTaTE + e) * 16, odd,
                        oa[4 + m], ob[4 + m], o[7 + m], o[15 + m]);
            TA_SEG(3);
            if (stage_close(j)) {
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&fullA[sj]);
                if (NGRP == 1 && ++st == kTaStages) { st = 0; ph ^= 1u; }
            }
            TA_SEG(4);
            if (jn < NCH) stx(xbuf(jn));
            TA_SEG(5);
            prod_sync();
            TA_SEG(6);
        }
#ifdef DEDF_TA_TRACE
        if (a.dbg && blockIdx.x == 0 && pr == 0 && grp == 0 && lane == 0)
            for (int k = 0; k < 8; ++k) a.dbg[8 + k] += ta_seg[k];
#endif
    } else {
        const int t = pr - 4;
        float w6[6];
        auto ldw = [&](int j) {
            if (a.w_perm) {
                const float2* q = reinterpret_cast<const float2*>(wr + 60 * j + 48 + 6 * t);
                const float2 q0 = q[0], q1 = q[1], q2 = q[2];
                w6[0] = q0.x; w6[1] = q0.y; w6[2] = q1.x; w6[3] = q1.y; w6[4] = q2.x; w6[5] = q2.y;
            } else {
                const int q = 2 * j + t;
#pragma unroll
                for (int k = 0; k < 6; ++k) w6[k] = wr[D::W_K9 + q + k * D::M2];
            }
        };
        ldx(j_first); ldw(j_first); stx(xbuf(j_first));
        prod_sync();
#pragma unroll 1
        for (int j = j_first; j < NCH; j = j_next(j)) {
            const int jn = j_next(j);
            const float* xrow = s_x + (xbuf(j) * kTaTE + e) * kTaXLd + 20 + 2 * (j & 1) + 5 * t;
            float x2[5] = {xrow[0], xrow[1], xrow[2], xrow[3], xrow[4]};
            if (NGRP == 2) prod_sync();
            if (jn < NCH) ldx(jn);
            float o[22];
            dtp_l2(x2, w6, sh, o);
            if (jn < NCH) ldw(jn);
            const int odd = F16 ? (j & 1) : 0;
            const uint32_t sj = stage_of(j);
            if (stage_open(j)) tc::mbar_wait_bounded(&emptyA[sj], wait_par(j));
            unsigned char* hi = sA + sj * kAStage;
            unsigned char* lo = hi + kAPart;
            S::put1(hi, lo, kA0Off + (t * kR0 + e) * 16, odd, 3, o[13]);       // l_out = 0, group t, column 3: k12
#pragma unroll
            for (int m = 0; m < 3; ++m)        // l_out = 1, group 5, columns 2t, 2t+1: [k10, k13]
                S::put2(hi, lo, kA1Off + (5 * kR1 + m * kTaTE + e) * 16, odd, 2 * t, o[5 + m], o[14 + m]);
#pragma unroll
            for (int m = 0; m < 5; ++m)        // l_out = 2, group 4 + t: [k9, k11, k14, 0]
                S::put4(hi, lo, m < 4 ? kA2Off + ((4 + t) * kR2 + m * kTaTE + e) * 16 : kA1Off + ((4 + t) * kR1 + 3 * kTaTE + e) * 16, odd,
                        o[m], o[8 + m], o[17 + m], 0.f);
            if (stage_close(j)) {
                tc::fence_async_smem();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(&fullA[sj]);
                if (NGRP == 1 && ++st == kTaStages) { st = 0; ph ^= 1u; }
            }
            if (jn < NCH) stx(xbuf(jn));
            prod_sync();
        }
    }
}

template <int G, bool F16>
__global__ void __launch_bounds__(kTaThreads, 1) edge_tp_act_tc_kernel(TpActArgs a) {
    using C = TaCfg<G>;
    using D = Dtp<G>;
    using S = TaStore<F16>;
    constexpr int NCH = C::NCH;
    constexpr int NST = F16 ? NCH / 2 : NCH;            // operand stages per tile: one per chunk (tf32) / per chunk pair (fp16)
    // producer groups (ta_produce_tile).  -DDEDF_TA_NGRP=2 (fp16 split only, NOT the default): the epilogue warps 8-13 produce the odd
    // chunk of every second pair (chunks 1 and 5 of 8; chunk 1 of 4) before they run the tile's epilogue, warps 0-5 the rest.
    // Measured at 86 k edges, G = 32: one group 230 us; two groups, even split 243 us; two groups, 6 + 2 split 242 us (G = 16,
    // 60 k edges: 84 -> 98 us).  With both groups running every segment of a producer's chunk gets slower (stores 819 -> 1045
    // cycles, the second group needs 3.8 k cycles per chunk instead of 2.6 k): the SM's LSU / issue slots are the bound, not the
    // latency of one warp's chain, so more producer warps buy nothing.
#ifndef DEDF_TA_NGRP
#define DEDF_TA_NGRP 1
#endif
    constexpr int NGRP = F16 ? DEDF_TA_NGRP : 1;
    constexpr uint32_t kAll = (1u << NCH) - 1u;
    constexpr uint32_t MASK1 = NGRP == 2 ? (NCH == 8 ? 0x22u : 0x2u) : 0u;
    constexpr uint32_t MASK0 = kAll & ~MASK1;
    static_assert(NCH % 2 == 0, "chunk pairs");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem + C::OffA;
    unsigned char* sW = smem + C::OffW;
    float* s_out = reinterpret_cast<float*>(smem + C::OffOut);
    float* s_gate = reinterpret_cast<float*>(smem + C::OffGate);      // [32][NGP]
    float* s_red = reinterpret_cast<float*>(smem + C::OffRed);        // [MA][33]
    float* s_x = reinterpret_cast<float*>(smem + C::OffX);
    float* s_b0 = reinterpret_cast<float*>(smem + C::OffPar);         // [N0P] bias
    float* s_adot = s_b0 + C::N0P;                                    // [MA]
    __shared__ __align__(8) uint64_t fullA[kTaStages], emptyA[kTaStages], fullW[kTaStages], emptyW[kTaStages], accFull[2], accEmpty[2];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup (overlaps the previous kernel under PDL: parameters only) ----
    if (tid == 0) {
        for (int s = 0; s < kTaStages; ++s) {
            mbar_init(&fullA[s], kTaProdWarps * NGRP); mbar_init(&emptyA[s], 1);      // NGRP = 2: 6 arrivals per HALF stage
            mbar_init(&fullW[s], 1); mbar_init(&emptyW[s], 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(&accFull[b], 1); mbar_init(&accEmpty[b], kTaEpiWarps); }
        mbar_init_fence();
    }
    if (warp == kTaTmaWarp) tc::tmem_alloc(&tmem_base_s, 2 * kTaAccCols);
    for (int i = tid; i < kTaStages * kAStage / 16; i += kTaThreads)            // pad columns must be finite (x zero weights)
        reinterpret_cast<float4*>(sA)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < C::N0P; i += kTaThreads) s_b0[i] = (i < C::N0 && a.bias0) ? a.bias0[i] : 0.f;
    for (int i = tid; i < C::MA; i += kTaThreads) s_adot[i] = a.alpha_dot[i];
    tc::fence_async_smem();
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = tmem_base_s;
    pdl_wait(); pdl_launch();
    const int E = *a.n_edges;
    const int n_tiles = (E + kTaTE - 1) / kTaTE;

    if (warp < kTaProdWarps) {
        // =========================== producers (group 0): CG chunk -> hi / lo A operand ===========================
        uint32_t st = 0, ph = 0, pair_base = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, pair_base += NST)
            ta_produce_tile<G, F16, NGRP, MASK0>(a, E, tile, warp, 0, lane, sA, s_x, fullA, emptyA, st, ph, pair_base);
    } else if (warp == kTaTmaWarp) {
        // =========================== weight ring ===========================
        if (lane == 0) {
            uint32_t st = 0, ph = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int j = 0; j < NST; ++j) {
                    tc::mbar_wait_bounded(&emptyW[st], ph ^ 1u);
                    mbar_expect_tx(&fullW[st], (uint32_t)C::WStage);
                    bulk_g2s_chunked(sW + st * C::WStage, reinterpret_cast<const unsigned char*>(a.Wp) + (size_t)j * C::WStage,
                                     (uint32_t)C::WStage, &fullW[st]);
                    if (++st == kTaStages) { st = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == kTaMmaWarp) {
        // =========================== MMA issuer ===========================
        // warp-uniform issue loop (tc.cuh: mma_tf32_if): all 32 lanes run it, the lane elected here issues
        {
            const uint32_t leader = tc::elect_one();
            uint32_t st = 0, ph = 0;
            const uint32_t a_base = smem_u32(sA), w_base = smem_u32(sW);
            const uint32_t id0 = F16 ? tc::idesc_f16(128, kTaTE) : tc::idesc_tf32(128, kTaTE);
            const uint32_t id1 = F16 ? tc::idesc_f16(128, C::N1C) : tc::idesc_tf32(128, C::N1C);
            const uint32_t id2 = F16 ? tc::idesc_f16(128, C::N2P) : tc::idesc_tf32(128, C::N2P);
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
                if constexpr (F16) tc::mma_f16_if(leader, d, da, db, id, acc);
                else tc::mma_tf32_if(leader, d, da, db, id, acc);
            };
            int it = 0;
            long long wA = 0, wW = 0, wE = 0, tIssue = 0;
            const long long tStart = clock64();
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = (uint32_t)it & 1u;
                { const long long c = clock64();
                tc::mbar_wait_bounded(&accEmpty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);     // epilogue drained this buffer
                wE += clock64() - c; }
                tc::fence_after();
                const uint32_t d0 = tmem_base + buf * (uint32_t)kTaAccCols;
                for (int j = 0; j < NST; ++j) {
                    { const long long c = clock64();
                    tc::mbar_wait_bounded(&fullA[st], ph);
                    const long long c2 = clock64();
                    tc::mbar_wait_bounded(&fullW[st], ph);
                    const long long c3 = clock64(); wA += c2 - c; wW += c3 - c2; }
                    tc::fence_after();
                    const long long ci = clock64();
                    const uint32_t ah = a_base + st * kAStage, al = ah + kAPart;
                    const uint32_t wh = w_base + st * C::WStage, wl = wh + C::WPart;
                    // One MMA = (GEMM g, K step ks, pass p of the 3xTF32 split); the four GEMMs of the chunk are issued round-robin
                    // so that consecutive instructions target different TMEM columns (measured: ~64 cycles per MMA either way --
                    // what an M = 128 x K = 8 tf32 MMA costs with N <= 48 is the fetch of its 4 KB M-side operand).
                    //   g = 0, 1: l_out = 0 with the operands swapped, D0^T[n][e] = W0^T[n][k] . A0[e][k] (M = channels, N = 32 edges)
                    //   g = 2:    rows 0..95 = 1e rows (m, e) against [W1 | .], rows 96..127 = the 2e m = 4 rows against [. | W2]:
                    //             one MMA with N = N1P + N2P, the cross terms land in accumulator cells nobody reads
                    //   g = 3:    2e rows (m, e), m = 0..3
                    auto issue = [&](int g, int i) {
                        const int ks = i / 3, p = i % 3;
                        const uint32_t acc = (j > 0 || i > 0) ? 1u : 0u;
                        if (g < 2) {
                            const uint32_t wo = C::W0Off + ks * 2 * (C::N0P * 16) + g * 128 * 16, ao = kA0Off + ks * 2 * (kR0 * 16);
                            const uint64_t dw = tc::smem_desc((p == 2 ? wl : wh) + wo, C::N0P * 16, 128);
                            const uint64_t da = tc::smem_desc((p == 1 ? al : ah) + ao, kR0 * 16, 128);
                            mma(d0 + C::T0 + g * kTaTE, dw, da, id0, acc);
                        } else {
                            const uint32_t a_off = (g == 2) ? kA1Off : kA2Off, lbo_a = (g == 2 ? kR1 : kR2) * 16;
                            const uint32_t w_off = (g == 2) ? C::W1Off : C::W2Off, lbo_w = (g == 2 ? C::N1C : C::N2P) * 16;
                            const uint64_t da = tc::smem_desc((p == 1 ? al : ah) + a_off + ks * 2 * lbo_a, lbo_a, 128);
                            const uint64_t dw = tc::smem_desc((p == 2 ? wl : wh) + w_off + ks * 2 * lbo_w, lbo_w, 128);
                            mma(d0 + (g == 2 ? C::T1 : C::T2A), da, dw, g == 2 ? id1 : id2, acc);
                        }
                    };
                    constexpr int n0 = 3 * (kKC0 / 8), n12 = 3 * (kKC1 / 8);
#pragma unroll
                    for (int i = 0; i < n12; ++i) {
                        if (i < n0) { issue(0, i); if (C::MT0 > 1) issue(1, i); }
                        issue(2, i);
                        issue(3, i);
                    }
                    tc::commit_if(leader, &emptyA[st]);
                    tc::commit_if(leader, &emptyW[st]);
                    tIssue += clock64() - ci;
                    if (++st == kTaStages) { st = 0; ph ^= 1u; }
                }
                tc::commit_if(leader, &accFull[buf]);
            }
            if (leader && a.dbg && blockIdx.x == 0) {      // host debug: where the MMA issuer of CTA 0 spent its cycles
                a.dbg[0] = clock64() - tStart; a.dbg[1] = wA; a.dbg[2] = wW; a.dbg[3] = wE; a.dbg[4] = tIssue; a.dbg[5] = it;
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                            // TMEM lane quarter
        const int wg = (warp - kTaEpiWarp0) >> 2;          // warpgroup: edges [16 wg, 16 wg + 16) of the l_out = 0 tile, half of the 1e columns
        auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kTaEpiWarps * 32) : "memory"); };
        constexpr int HD = C::MA / 4;
        int it = 0;
        uint32_t pst = 0, pph = 0;                          // producer-side stage / phase (second producer group)
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint32_t buf = (uint32_t)it & 1u;
            const int e0 = tile * kTaTE;
            if constexpr (NGRP == 2) {
                if (warp < kTaEpiWarp0 + kTaProdWarps)                  // this tile's share of the production first, then its epilogue
                    ta_produce_tile<G, F16, NGRP, MASK1>(a, E, tile, warp - kTaEpiWarp0, 1, lane, sA, s_x, fullA, emptyA, pst, pph,
                                                         (uint32_t)it * NST);
            }
#ifdef DEDF_TA_TRACE
            const long long ep_c0 = clock64();
#endif
            tc::mbar_wait_bounded(&accFull[buf], ((uint32_t)it >> 1) & 1u);
            tc::fence_after();
#ifdef DEDF_TA_TRACE
            const long long ep_c1 = clock64();
#endif
            // the previous tile's value rows have left the staging tile (their bulk stores were issued a whole tile ago: the wait is
            // free here, at the end of that tile's epilogue it cost the drain time of 30 KB)
            if (warp == kTaEpiWarp0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            epi_sync();
            const uint32_t tq = tmem_base + buf * (uint32_t)kTaAccCols + ((uint32_t)(q * 32) << 16);
            // ---- phase 1: the 0e GEMM (lane = output channel n, column = edge): logit terms | scalars | gates ----
#pragma unroll
            for (int mt = 0; mt < C::MT0; ++mt) {
                const int n = mt * 128 + q * 32 + lane;
                if (mt * 128 + q * 32 < C::N0) {           // warp-uniform
                    float t[16];
                    tc::tmem_ld16(tq + C::T0 + mt * kTaTE + 16 * wg, t);
                    if (n < C::N0) {
                        const float b = s_b0[n];
                        if (n < C::MA) {                   // graph_attention.py:241-246
                            const float ad = kCSlrelu * s_adot[n];
#pragma unroll
                            for (int k = 0; k < 16; ++k) s_red[n * 33 + 16 * wg + k] = slreluf_(t[k] + b) * ad;
                        } else if (n < C::MA + D::M0) {    // fast_activation.py:210-224: scalars | gates | gated
#pragma unroll
                            for (int k = 0; k < 16; ++k) s_out[(16 * wg + k) * C::LDO + (n - C::MA)] = kCSilu * siluf_(t[k] + b);
                        } else {
#pragma unroll
                            for (int k = 0; k < 16; ++k) s_gate[(16 * wg + k) * C::NGP + (n - C::MA - D::M0)] = kCSigmoid * sigmoidf_(t[k] + b);
                        }
                    }
                }
            }
            epi_sync();
            // ---- phase 2: logits (thread = (edge, head, half of the head's channels)) ----
            {
                const int e = 16 * wg + (lane & 15), kh = lane >> 4, h = q;
                float sum = 0.f;
#pragma unroll
                for (int k = 0; k < HD / 2; ++k) sum += s_red[(h * HD + kh * (HD / 2) + k) * 33 + e];
                sum += __shfl_xor_sync(0xffffffffu, sum, 16);
                if (kh == 0 && e0 + e < E) a.logits[(size_t)(e0 + e) * 4 + h] = sum + (a.edge_logit ? a.edge_logit[e0 + e] : 0.f);
            }
            // ---- 1e / 2e GEMMs (lane = edge, quarter = component m, column = channel): x gate ----
            {
                const int e = lane;
                float* orow = s_out + e * C::LDO;
                const float* grow = s_gate + e * C::NGP;
                if (q < 3) {
                    constexpr int W1 = (D::M1 >= 32) ? 16 : D::M1;          // columns per warpgroup (16 | all 16 in warpgroup 0)
                    if (wg * W1 < D::M1) {
                        float t[16];
                        tc::tmem_ld16(tq + C::T1 + wg * W1, t);
#pragma unroll
                        for (int k = 0; k < 16; ++k)
                            if (k < W1) orow[D::M0 + 3 * (wg * W1 + k) + q] = t[k] * grow[wg * W1 + k];
                    }
                }
                if (wg == 0) {
                    float t[16];
                    tc::tmem_ld16(tq + C::T2A, t);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < D::M2) orow[D::M0 + 3 * D::M1 + 5 * k + q] = t[k] * grow[D::M1 + k];
                } else if (q == 3) {
                    float t[16];
                    tc::tmem_ld16(tq + C::T2B, t);
#pragma unroll
                    for (int k = 0; k < 16; ++k)
                        if (k < D::M2) orow[D::M0 + 3 * D::M1 + 5 * k + 4] = t[k] * grow[D::M1 + k];
                }
            }
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&accEmpty[buf]);
            tc::fence_async_smem();
            epi_sync();                                    // the 32 value rows are staged
            if (warp == kTaEpiWarp0) {
                if (e0 + lane < E)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(a.out + (size_t)(e0 + lane) * D::F), "r"(smem_u32(s_out + lane * C::LDO)), "r"((uint32_t)D::F * 4u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // (gate and logit tables: every read of them precedes the barrier above; the staging tile is released at the top of the
            //  next tile's epilogue)
#ifdef DEDF_TA_TRACE
            if (a.dbg && blockIdx.x == 0 && warp == kTaEpiWarp0 && lane == 0) { a.dbg[16] += ep_c1 - ep_c0; a.dbg[17] += clock64() - ep_c1; }
#endif
        }
    }
    if (warp == kTaEpiWarp0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // shared memory stays valid until the last rows left
    tc::fence_before();
    __syncthreads();
    if (warp == kTaTmaWarp) tc::tmem_dealloc(tmem_base, 2 * kTaAccCols);
}

template <int G, bool F16>
static int launch_tp_act_tc(const TpActArgs& a, int max_edges, cudaStream_t stream) {
    using C = TaCfg<G>;
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(edge_tp_act_tc_kernel<G, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::Smem);
        attr_done = true;
    }
    const int n_tiles = (max_edges + kTaTE - 1) / kTaTE;
    launch_pdl((edge_tp_act_tc_kernel<G, F16>), dim3(grid_for(n_tiles, 1, kNumSMs)), dim3(kTaThreads), (size_t)C::Smem, stream, a);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

}  // namespace dedf

using namespace dedf;

static long long* g_ta_dbg = nullptr;
/* debug hook (not in the public header): 8 x int64 device buffer receiving the MMA issuer's cycle accounting of CTA 0 */
extern "C" int dedf_tp_act_tc_set_debug(long long* dbg) { g_ta_dbg = dbg; return DEDF_OK; }

extern "C" int dedf_edge_tp_act_tc(int mul1, const float* x_src, const float* x_dst, const int* edge_src, const int* edge_dst,
                                   const int* n_edges_dev, int max_edges, const float* sh, const float* w, long long w_stride,
                                   int w_perm, const float* W_tc, const float* bias0, const float* alpha_dot, const float* edge_logit,
                                   float* logits, float* out, cudaStream_t stream) {
    if (max_edges <= 0) return DEDF_OK;
    if (!x_src || !edge_src || !edge_dst || !n_edges_dev || !sh || !w || !W_tc || !alpha_dot || !logits || !out) return DEDF_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(W_tc) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(logits) & 15) ||
        (reinterpret_cast<uintptr_t>(w) & 15) || (reinterpret_cast<uintptr_t>(x_src) & 15) || (x_dst && (reinterpret_cast<uintptr_t>(x_dst) & 15)) ||
        (w_stride & 3))
        return DEDF_ERR_ARG;
    TpActArgs a{};
    a.x_src = x_src; a.x_dst = x_dst; a.edge_src = edge_src; a.edge_dst = edge_dst; a.n_edges = n_edges_dev; a.sh = sh;
    a.w = w; a.w_stride = w_stride; a.w_perm = (w_perm & DEDF_TPACT_W_PERM) ? 1 : 0; a.Wp = W_tc; a.bias0 = bias0; a.alpha_dot = alpha_dot; a.edge_logit = edge_logit;
    a.logits = logits; a.out = out; a.dbg = g_ta_dbg;
    const bool f16 = (w_perm & DEDF_TPACT_F16) != 0;
    if (mul1 == 32) return f16 ? launch_tp_act_tc<32, true>(a, max_edges, stream) : launch_tp_act_tc<32, false>(a, max_edges, stream);
    if (mul1 == 16) return f16 ? launch_tp_act_tc<16, true>(a, max_edges, stream) : launch_tp_act_tc<16, false>(a, max_edges, stream);
    return DEDF_ERR_UNSUPPORTED;
}
