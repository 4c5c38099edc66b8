// Graph construction kernels: farthest-point sampling, radius search (CSR by
// destination, sources ascending -- the order torch_cluster.radius returns), scan.
//
// Replaces, for the reference's hot path:
//   torch_cluster.fps            (connectivity.py:62)
//   torch_cluster.radius         (graph_parser.py:339, connectivity.py:42)
//   torch_cluster.radius_graph   (connectivity.py:22)
//   the all-pairs meshgrid of InfiniteBipartite (graph_parser.py:272-286)
// Index results are bit-exact against oracle/graph.py: squared distances are
// evaluated as (dx*dx + dy*dy) + dz*dz with no FMA contraction.
#include "common.cuh"
#include <cooperative_groups.h>
#include "../../include/dedf.h"

namespace dedf {

__device__ __forceinline__ float sqdist_exact(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------
// FPS: one persistent CTA per call; points and running min-distances live in
// registers (PT per thread), one barrier per iteration (double-buffered slots).
// ---------------------------------------------------------------------------
constexpr int kFpsThreads = 1024;

__device__ __forceinline__ void argmax_pair(float& v, int& i, float ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}

// Per iteration: every thread updates its PT points (10 FP instructions each: the exact, unfused squared distance, a min and
// a max), the block arg-max is two REDUX pairs (max of the distance bits, then min index among the holders of the max) around
// ONE barrier, and the winner's coordinates are read back from a shared-memory copy of the cloud (no global load, no second
// barrier on the critical path).
template <int PT>
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel(const float* __restrict__ x, int n, int m, int start, const long long* __restrict__ start_dev, int idx_base,
           long long* __restrict__ out_idx) {
    extern __shared__ float s_pts[];            // [3 n] copy of the cloud for the winner broadcast
    __shared__ unsigned s_val[2][32];
    __shared__ int s_idx[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 3 * n; i += kFpsThreads) s_pts[i] = x[i];
    float px[PT], py[PT], pz[PT], dist[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int i = tid + j * kFpsThreads;
        if (i < n) { px[j] = x[3 * i]; py[j] = x[3 * i + 1]; pz[j] = x[3 * i + 2]; dist[j] = __int_as_float(0x7f800000); }
        else       { px[j] = py[j] = pz[j] = 0.f; dist[j] = 0.f; }      // padding: never beats a real point (ties -> lowest index)
    }
    __syncthreads();
    int cur = start_dev ? (int)min(max(start_dev[0], 0ll), (long long)(n - 1)) : start;
    for (int it = 0; it < m; ++it) {
        if (tid == 0) out_idx[it] = (long long)(cur + idx_base);
        const float cx = s_pts[3 * cur], cy = s_pts[3 * cur + 1], cz = s_pts[3 * cur + 2];
        float bv = 0.f;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            const float d = sqdist_exact(px[j], py[j], pz[j], cx, cy, cz);      // padding: dist = 0 stays 0 (see fps_cluster_kernel)
            dist[j] = fminf(dist[j], d);
            bv = fmaxf(bv, dist[j]);
        }
        // distances are >= 0, so their bit patterns order like unsigned integers
        int mine = 0x7fffffff;
#pragma unroll
        for (int j = PT - 1; j >= 0; --j)
            if (dist[j] == bv) mine = tid + j * kFpsThreads;                        // lowest index among this thread's ties
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(bv));
        const int cand = (__float_as_uint(bv) == wm) ? mine : 0x7fffffff;
        const int wi = __reduce_min_sync(0xffffffffu, cand);
        const int buf = it & 1;
        if (lane == 0) { s_val[buf][warp] = wm; s_idx[buf][warp] = wi; }
        __syncthreads();
        const unsigned v = s_val[buf][lane];
        const unsigned gm = __reduce_max_sync(0xffffffffu, v);
        cur = __reduce_min_sync(0xffffffffu, (v == gm) ? s_idx[buf][lane] : 0x7fffffff);
    }
}

// Cluster variant for the large scales: 8 CTAs (8 SMs, one thread-block cluster) split the points, so the per-iteration
// distance update shrinks 8x; every CTA keeps a full shared-memory copy of the cloud (winner coordinates need no
// communication).  Candidate exchange: every WARP publishes one 64-bit key [distance bits | ~index | iteration tag] straight
// into the slot tables of all 8 CTAs with plain distributed-shared-memory stores; every warp then polls its own CTA's 64
// slots until all carry this iteration's tag and reduces them.  One one-way DSMEM flight per iteration: no __syncthreads,
// no mbarrier, no cluster barrier (a cluster.sync costs ~380 cycles and flushes L1).  Slots are double-buffered by iteration
// parity: a peer can only write the slots of iteration it+2 after it has received this CTA's keys of iteration it+1, which
// are sent after the slots of iteration it were read.  Same arithmetic and tie-breaking (largest distance, then lowest
// index) as the single-CTA kernel.
constexpr int kFpsClusterSize = 8;
constexpr int kFpsClusterThreads = 256;
constexpr int kFpsTagBits = 18;                 // iteration tag; 14 bits of (complemented) index; 32 bits of distance
constexpr int kFpsClusterMaxN = 1 << 14;

__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t fps_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}

template <int PT>
__global__ void __launch_bounds__(kFpsClusterThreads, 1)
fps_cluster_kernel(const float* __restrict__ x, int n, int m, int start, const long long* __restrict__ start_dev, int idx_base,
                   long long* __restrict__ out_idx) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int NW = kFpsClusterThreads / 32, STRIDE = kFpsClusterSize * kFpsClusterThreads;
    constexpr int NSLOT = kFpsClusterSize * NW;                     // 64 keys per iteration
    constexpr int SPL = NSLOT / 32;                                 // slots per lane
    constexpr unsigned long long TAG_MASK = (1ull << kFpsTagBits) - 1;
    extern __shared__ float s_pts[];
    __shared__ __align__(8) unsigned long long c_slot[2][NSLOT];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rank = (int)cluster.block_rank();
    const int gtid = rank * kFpsClusterThreads + tid;
    for (int i = tid; i < 2 * NSLOT; i += kFpsClusterThreads) (&c_slot[0][0])[i] = TAG_MASK;          // a tag no iteration uses (m < 2^18 - 1)
    for (int i = tid; i < 3 * n; i += kFpsClusterThreads) s_pts[i] = x[i];
    float px[PT], py[PT], pz[PT], dist[PT];
#pragma unroll
    for (int j = 0; j < PT; ++j) {
        const int i = gtid + j * STRIDE;
        if (i < n) { px[j] = x[3 * i]; py[j] = x[3 * i + 1]; pz[j] = x[3 * i + 2]; dist[j] = __int_as_float(0x7f800000); }
        else       { px[j] = py[j] = pz[j] = 0.f; dist[j] = 0.f; }
    }
    // remote address of this warp's slot in CTA `lane` (lanes 0..7 publish)
    const uint32_t peer = (uint32_t)(lane & (kFpsClusterSize - 1));
    const uint32_t r_slot0 = fps_mapa(fps_smem_u32(&c_slot[0][rank * NW + warp]), peer);
    const uint32_t r_slot1 = fps_mapa(fps_smem_u32(&c_slot[1][rank * NW + warp]), peer);
    const uint32_t l_slot = fps_smem_u32(&c_slot[0][lane]);
    cluster.sync();     // slot tables initialised and clouds staged everywhere before the first remote store
    int cur = start_dev ? (int)min(max(start_dev[0], 0ll), (long long)(n - 1)) : start;
    for (int it = 0; it < m; ++it) {
        const int buf = it & 1;
        if (gtid == 0) out_idx[it] = (long long)(cur + idx_base);
        const float cx = s_pts[3 * cur], cy = s_pts[3 * cur + 1], cz = s_pts[3 * cur + 2];
        float bv = 0.f;
#pragma unroll
        for (int j = 0; j < PT; ++j) {
            // no bounds test: a padding slot has dist = 0 and fminf(0, d) = 0 for every d >= 0 -- the per-point branch put each of
            // the PT updates into its own reconvergence region, i.e. serialised five dependent FADD/FMUL chains per iteration
            const float d = sqdist_exact(px[j], py[j], pz[j], cx, cy, cz);
            dist[j] = fminf(dist[j], d);
            bv = fmaxf(bv, dist[j]);
        }
        // distances are >= 0, so their bit patterns order like unsigned integers.  The thread's own candidate (lowest index among
        // ITS maxima) does not depend on the warp maximum: it is computed in the shadow of the first REDUX, and only one
        // compare + select sits between the two reductions.
        int mine = kFpsClusterMaxN - 1;
#pragma unroll
        for (int j = PT - 1; j >= 0; --j)
            if (dist[j] == bv) mine = gtid + j * STRIDE;
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(bv));
        const int cand = (__float_as_uint(bv) == wm) ? mine : kFpsClusterMaxN - 1;
        const int wi = __reduce_min_sync(0xffffffffu, cand);
        // key: larger distance first, then LOWER index (stored complemented so that one max picks both), then the tag
        const unsigned long long tag = (unsigned long long)it & TAG_MASK;
        const unsigned long long key = ((unsigned long long)wm << 32) |
                                       ((unsigned long long)((kFpsClusterMaxN - 1) - wi) << kFpsTagBits) | tag;
        if (lane < kFpsClusterSize)
            asm volatile("st.relaxed.cluster.shared::cluster.b64 [%0], %1;" ::"r"(buf ? r_slot1 : r_slot0), "l"(key) : "memory");
        // poll this CTA's slots (SPL per lane) for the keys of this iteration (bounded: a lost store traps, not hangs)
        unsigned long long kmax;
        {
            const uint32_t a0 = l_slot + (uint32_t)buf * NSLOT * 8;
            int spins = 0;
            bool ok;
            do {
                ok = true; kmax = 0;
#pragma unroll
                for (int q = 0; q < SPL; ++q) {
                    unsigned long long k;
                    asm volatile("ld.relaxed.cluster.shared::cta.b64 %0, [%1];" : "=l"(k) : "r"(a0 + q * 32 * 8) : "memory");
                    ok = ok && ((k & TAG_MASK) == tag);
                    kmax = k > kmax ? k : kmax;
                }
                if (++spins > (1 << 24)) __trap();
            } while (!__all_sync(0xffffffffu, ok));
        }
        const unsigned long long kk = kmax >> kFpsTagBits;     // [distance | ~index], 46 bits
        const unsigned hi = (unsigned)(kk >> 14);
        const unsigned gm = __reduce_max_sync(0xffffffffu, hi);
        const unsigned gl = __reduce_max_sync(0xffffffffu, (hi == gm) ? (unsigned)(kk & (kFpsClusterMaxN - 1)) : 0u);
        cur = (kFpsClusterMaxN - 1) - (int)gl;
    }
    cluster.sync();     // no CTA may exit while a peer could still write into its shared memory
}

// fallback for very large clouds: running distances in global scratch
__global__ void __launch_bounds__(kFpsThreads, 1)
fps_kernel_large(const float* __restrict__ x, int n, int m, int start, const long long* __restrict__ start_dev, int idx_base,
                 long long* __restrict__ out_idx, float* __restrict__ dist) {
    __shared__ float s_val[2][32];
    __shared__ int s_idx[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < n; i += kFpsThreads) dist[i] = __int_as_float(0x7f800000);
    __syncthreads();
    int cur = start_dev ? (int)min(max(start_dev[0], 0ll), (long long)(n - 1)) : start;
    for (int it = 0; it < m; ++it) {
        if (tid == 0) out_idx[it] = (long long)(cur + idx_base);
        const float cx = __ldg(x + 3 * cur), cy = __ldg(x + 3 * cur + 1), cz = __ldg(x + 3 * cur + 2);
        float bv = -2.0f; int bi = 0x7fffffff;
        for (int i = tid; i < n; i += kFpsThreads) {
            float d = sqdist_exact(x[3 * i], x[3 * i + 1], x[3 * i + 2], cx, cy, cz);
            float nd = fminf(dist[i], d);
            dist[i] = nd;
            argmax_pair(bv, bi, nd, i);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmax_pair(bv, bi, ov, oi);
        }
        const int buf = it & 1;
        if (lane == 0) { s_val[buf][warp] = bv; s_idx[buf][warp] = bi; }
        __syncthreads();
        bv = s_val[buf][lane]; bi = s_idx[buf][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            argmax_pair(bv, bi, ov, oi);
        }
        cur = bi;
    }
}

// ---------------------------------------------------------------------------
// Radius search, brute force, one warp per (destination, scale): the warp walks
// the sources 32 at a time in index order and compacts hits with a ballot, so
// every destination's neighbour list comes out ascending and the
// ``max_num_neighbors`` truncation keeps the first ones, like torch_cluster.
// ---------------------------------------------------------------------------
struct RadiusArgs {
    const float* x_src;            // (sum n_src, 3)
    const float* x_dst;            // (n_dst, 3)
    const long long* b_src;        // optional batch ids
    const long long* b_dst;
    const long long* excl;         // exclusion table (mode 1: per dst, mode 3: per src)
    int n_dst, n_scales, max_nb, excl_mode;
    int src_off[DEDF_MAX_SCALES + 1];
    float r[DEDF_MAX_SCALES];      // < 0: all pairs
};

// SMEM: the concatenated source clouds (and their batch ids) are staged in shared memory once per CTA -- the score head's
// query points meet the same <= 2.5k scene points at every scale, and a warp's walk is a chain of dependent iterations, so
// the latency of one iteration (global: ~600 cycles, shared: ~30) is what the kernel costs.  Four 32-source chunks are
// evaluated per iteration before their ordered compaction; chunks past the max_nb cap keep nothing, as before.
template <bool FILL, bool SMEM>
__global__ void __launch_bounds__(256)
radius_kernel(RadiusArgs a, int* __restrict__ counts, const int* __restrict__ row_ptr,
              int* __restrict__ edge_src, int* __restrict__ edge_dst) {
    extern __shared__ __align__(16) float s_radius[];
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const long long n_items = (long long)a.n_dst * a.n_scales;
    const int n_total = a.src_off[a.n_scales];
    const float* xs = a.x_src;
    const int* s_b = nullptr;
    if (SMEM) {
        // Batches of independent loads, then the stores: written as one load + one store per iteration the loop is a chain of
        // ~30 serialised L2 round trips per thread (the stores cannot be hoisted over possibly aliasing loads), which was two
        // thirds of this kernel's 31 us at 128 poses (ncu: long-scoreboard stalls on the STS, profiles/r1_s5_radius_head128_ncu.txt).
        constexpr int UX = 16, UB = 8;
        const int n3 = 3 * n_total, bd = blockDim.x;
        for (int i0 = threadIdx.x; i0 < n3; i0 += bd * UX) {
            float v[UX];
#pragma unroll
            for (int k = 0; k < UX; ++k) { const int i = i0 + k * bd; v[k] = (i < n3) ? __ldg(a.x_src + i) : 0.f; }
#pragma unroll
            for (int k = 0; k < UX; ++k) { const int i = i0 + k * bd; if (i < n3) s_radius[i] = v[k]; }
        }
        if (a.b_src) {
            int* sb = reinterpret_cast<int*>(s_radius + n3);
            for (int i0 = threadIdx.x; i0 < n_total; i0 += bd * UB) {
                long long v[UB];
#pragma unroll
                for (int k = 0; k < UB; ++k) { const int i = i0 + k * bd; v[k] = (i < n_total) ? __ldg(a.b_src + i) : 0ll; }
#pragma unroll
                for (int k = 0; k < UB; ++k) { const int i = i0 + k * bd; if (i < n_total) sb[i] = (int)v[k]; }
            }
            s_b = sb;
        }
        __syncthreads();
        xs = s_radius;
    }
    for (long long item = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); item < n_items;
         item += (long long)gridDim.x * warps_per_block) {
        const int s = (int)(item / a.n_dst), d = (int)(item % a.n_dst);
        const int s0 = a.src_off[s], s1 = a.src_off[s + 1];
        const float r = a.r[s];
        const bool all = r < 0.f;
        const float r2 = r * r;
        const float qx = a.x_dst[3 * d], qy = a.x_dst[3 * d + 1], qz = a.x_dst[3 * d + 2];
        const long long qb = a.b_dst ? a.b_dst[d] : 0;
        const long long ex = (a.excl_mode == 1) ? a.excl[d] : -1;
        // ``max_nb`` caps the hits INCLUDING excluded ones (torch_cluster truncates first, the
        // reference filters self pairs afterwards: connectivity.py:68-70, radius_graph's +1).
        // ... and only for radius scales: InfiniteBipartite builds the full meshgrid, its max_neighbors is a placeholder
        // (graph_parser.py:274-278)
        const int cap_nb = all ? 0x7fffffff : a.max_nb;
        int cnt_all = 0, cnt_keep = 0;
        int base = 0, seg_end = 0x7fffffff;
        if (FILL) { base = row_ptr[item]; seg_end = row_ptr[item + 1]; }
        constexpr int U = 4;
        for (int c = s0; c < s1 && cnt_all < cap_nb; c += 32 * U) {
            bool hit[U], excluded[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = c + 32 * u + lane;
                hit[u] = false; excluded[u] = false;
                if (i < s1) {
                    hit[u] = all || sqdist_exact(xs[3 * i], xs[3 * i + 1], xs[3 * i + 2], qx, qy, qz) < r2;
                    if (a.b_src && !all)      // the all-pairs scale ignores batches (graph_parser.py:276-278)
                        hit[u] = hit[u] && (SMEM ? ((long long)s_b[i] == qb) : (a.b_src[i] == qb));
                    const int li = i - s0;   // index local to this scale's cloud
                    if (a.excl_mode == 1) excluded[u] = ((long long)li == ex);
                    else if (a.excl_mode == 2) excluded[u] = (li == d);
                    else if (a.excl_mode == 3) excluded[u] = (a.excl[li] == (long long)d);
                }
            }
            const unsigned lt = (1u << lane) - 1u;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = c + 32 * u + lane;
                const unsigned bal_all = __ballot_sync(0xffffffffu, hit[u]);
                const bool keep = hit[u] && !excluded[u] && (cnt_all + __popc(bal_all & lt) < cap_nb);
                const unsigned bal_keep = __ballot_sync(0xffffffffu, keep);
                if (FILL && keep) {
                    const int pos = base + cnt_keep + __popc(bal_keep & lt);
                    if (pos < seg_end) {
                        edge_src[pos] = i;              // flat index into the concatenated clouds
                        edge_dst[pos] = d;
                    }
                }
                cnt_all += __popc(bal_all);
                cnt_keep += __popc(bal_keep);
            }
        }
        if (!FILL && lane == 0) counts[item] = cnt_keep;
    }
}

// launch the brute-force radius kernel, staging the sources in shared memory when they fit
template <bool FILL>
static void launch_radius(const RadiusArgs& a, long long items, int* counts, const int* row_ptr, int* edge_src, int* edge_dst,
                          cudaStream_t stream) {
    const int n_total = a.src_off[a.n_scales];
    const size_t smem = (size_t)n_total * (12 + (a.b_src ? 4 : 0));
    if (smem > 0 && smem <= 96 * 1024 && !getenv("DEDF_NO_SMEM_RADIUS")) {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(radius_kernel<FILL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); done = true; }
        launch_pdl((radius_kernel<FILL, true>), dim3(grid_for(items, 8, kNumSMs * 2)), dim3(256), smem, stream, a, counts, row_ptr, edge_src, edge_dst);
    } else {
        launch_pdl((radius_kernel<FILL, false>), dim3(grid_for(items, 8, kNumSMs * 8)), dim3(256), 0, stream, a, counts, row_ptr, edge_src, edge_dst);
    }
}

// single-CTA exclusive scan: row_ptr[0..n] from counts[0..n-1]
// STAGED: the counts are first copied into shared memory with coalesced, independent loads (n <= kScanStageMax); every thread then
// scans its contiguous chunk there.  The direct variant walked its chunk with one dependent global load per element: 59 us for the
// 32 768 buckets of the 10k-point hash grid, all of it on the critical path in front of the first UNet block
// (profiles/r2_s8_timeline_128_before.txt).
constexpr int kScanStageMax = 48 * 1024;           // ints staged in shared memory (+ 1/32 padding): 198 KB
__device__ __forceinline__ int scan_pad(int i) { return i + (i >> 5); }     // chunk starts of consecutive threads fall into different banks

template <bool STAGED>
__global__ void __launch_bounds__(1024, 1)
exclusive_scan_kernel(const int* __restrict__ counts, int n, int* __restrict__ row_ptr, int capacity,
                      int* __restrict__ n_edges_out, int* __restrict__ overflow) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    extern __shared__ int s_all[];
    __shared__ int s_part[1024];
    const int tid = threadIdx.x;
    const int chunk = (n + 1023) / 1024;
    const int lo = min(tid * chunk, n), hi = min(lo + chunk, n);
    if (STAGED) {
        for (int i = tid; i < n; i += 1024) s_all[scan_pad(i)] = counts[i];
        __syncthreads();
    }
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += STAGED ? s_all[scan_pad(i)] : counts[i];
    s_part[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
        int v = (tid >= o) ? s_part[tid - o] : 0;
        __syncthreads();
        s_part[tid] += v;
        __syncthreads();
    }
    // capacity > 0: the edge buffers were sized ahead of time (CUDA-graph replay); clamp the CSR so that no consumer can
    // index past them and raise the overflow flag -- the host then re-plans with larger buffers.
    const int cap = capacity > 0 ? capacity : 0x7fffffff;
    int run = s_part[tid] - sum;
    if (STAGED) {
        for (int i = lo; i < hi; ++i) { const int c = s_all[scan_pad(i)]; s_all[scan_pad(i)] = min(run, cap); run += c; }
        __syncthreads();
        for (int i = tid; i < n; i += 1024) row_ptr[i] = s_all[scan_pad(i)];        // coalesced
    } else {
        for (int i = lo; i < hi; ++i) { row_ptr[i] = min(run, cap); run += counts[i]; }
    }
    if (tid == 1023) {
        const int total = s_part[1023];
        row_ptr[n] = min(total, cap);
        if (n_edges_out) *n_edges_out = min(total, cap);
        if (overflow && total > cap) atomicOr(overflow, 1);
    }
}

static cudaError_t launch_scan(const int* counts, int n, int* row_ptr, int capacity, int* n_edges_out, int* overflow, cudaStream_t stream) {
    if (n <= kScanStageMax && !getenv("DEDF_SCAN_DIRECT")) {
        static bool done = false;
        if (!done) { cudaFuncSetAttribute(exclusive_scan_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (kScanStageMax + kScanStageMax / 32 + 32) * 4); done = true; }
        return launch_pdl(exclusive_scan_kernel<true>, dim3(1), dim3(1024), (size_t)(n + n / 32 + 32) * sizeof(int), stream, counts, n, row_ptr,
                          capacity, n_edges_out, overflow);
    }
    return launch_pdl(exclusive_scan_kernel<false>, dim3(1), dim3(1024), 0, stream, counts, n, row_ptr, capacity, n_edges_out, overflow);
}


// ---------------------------------------------------------------------------
// Grid-hash radius search (one source cloud, finite radius): sources are counting-sorted into the buckets of a hashed
// uniform grid with cell edge slightly larger than r (so every neighbour lies in the 27 surrounding cells); a warp per
// destination scans those cells, collects the hits in shared memory, sorts them ascending (bitonic) and applies
// torch_cluster's truncation + the reference's exclusion filter on the ordered list.  Results are identical, element for
// element, to the brute-force kernel above (same exact squared distance); destinations with more hits than the shared
// buffer holds fall back to the ordered brute-force scan.
// ---------------------------------------------------------------------------
constexpr int kGridMaxHits = 1024;          // per-warp hit buffer (ints)
constexpr int kGridWarps = 8;

__device__ __forceinline__ int3 grid_cell(float x, float y, float z, float inv_cell) {
    return make_int3((int)floorf(x * inv_cell), (int)floorf(y * inv_cell), (int)floorf(z * inv_cell));
}
__device__ __forceinline__ unsigned grid_hash(int3 c, unsigned mask) {
    return ((unsigned)c.x * 73856093u ^ (unsigned)c.y * 19349663u ^ (unsigned)c.z * 83492791u) & mask;
}

__global__ void grid_count_kernel(const float* __restrict__ x, int n, float inv_cell, unsigned mask, int* __restrict__ bucket_cnt) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(bucket_cnt + grid_hash(grid_cell(x[3 * i], x[3 * i + 1], x[3 * i + 2], inv_cell), mask), 1);
}

__global__ void grid_fill_kernel(const float* __restrict__ x, int n, float inv_cell, unsigned mask, const int* __restrict__ bucket_start,
                                 int* __restrict__ bucket_fill, int* __restrict__ sorted_idx) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned b = grid_hash(grid_cell(x[3 * i], x[3 * i + 1], x[3 * i + 2], inv_cell), mask);
        sorted_idx[bucket_start[b] + atomicAdd(bucket_fill + b, 1)] = i;
    }
}

// Lay the coordinates out in bucket order so that the query kernel streams them.  The order INSIDE a bucket is whatever the fill
// atomics produced: the query kernel sorts every destination's hit list by source index anyway, so the CSR does not depend on it.
// (Until the end of round 2 a thread per bucket insertion-sorted its bucket in global memory first: with cell = r a bucket of the
// coarser scales holds dozens of points, and that sort was 40-76 us per grid -- 124 us per forward -- of purely redundant work.)
__global__ void grid_gather_kernel(const float* __restrict__ x, int n, const int* __restrict__ sorted_idx, float* __restrict__ sorted_xyz) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int v = sorted_idx[i];
        sorted_xyz[3 * i] = x[3 * v]; sorted_xyz[3 * i + 1] = x[3 * v + 1]; sorted_xyz[3 * i + 2] = x[3 * v + 2];
    }
}

struct GridArgs {
    const float* x_src; const float* x_dst;      // x_src only for the brute-force fallback
    const long long* b_src; const long long* b_dst; const long long* excl;
    const int* bucket_start; const int* sorted_idx; const float* sorted_xyz;
    int n_src, n_dst, max_nb, excl_mode;
    float r, inv_cell; unsigned mask;
};

template <bool FILL>
__global__ void __launch_bounds__(kGridWarps * 32)
radius_grid_kernel(GridArgs a, int* __restrict__ counts, const int* __restrict__ row_ptr, int* __restrict__ edge_src,
                   int* __restrict__ edge_dst) {
    pdl_wait(); pdl_launch();     // PDL: see common.cuh
    __shared__ int s_hits[kGridWarps][kGridMaxHits];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int* hits = s_hits[warp];
    const float r2 = a.r * a.r;
    for (int d = blockIdx.x * kGridWarps + warp; d < a.n_dst; d += gridDim.x * kGridWarps) {
        const float qx = a.x_dst[3 * d], qy = a.x_dst[3 * d + 1], qz = a.x_dst[3 * d + 2];
        const long long qb = a.b_dst ? a.b_dst[d] : 0;
        const int3 qc = grid_cell(qx, qy, qz, a.inv_cell);
        int n_hits = 0;
        __syncwarp();
        // Lane nb < 27 owns ONE of the 27 neighbouring cells: the 27 bucket ranges are fetched in one round trip and each lane walks
        // its own bucket, so the dependent-load chain is 1 + (largest bucket) round trips instead of 27 x 2 (the cells used to be
        // visited one after the other by the whole warp: 32 us for 2 000 destinations, all of it load latency).  The hits are
        // collected in arbitrary order; the bitonic sort below restores ascending source order.
        const int nb = lane < 27 ? lane : 0;
        const int3 c = make_int3(qc.x + nb % 3 - 1, qc.y + (nb / 3) % 3 - 1, qc.z + nb / 9 - 1);
        int s = 0, e = 0;
        if (lane < 27) { const unsigned b = grid_hash(c, a.mask); s = a.bucket_start[b]; e = a.bucket_start[b + 1]; }
        const int longest = __reduce_max_sync(0xffffffffu, e - s);
        for (int k = 0; k < longest; ++k) {
            const int i = s + k;
            bool hit = false;
            int src = 0;
            if (i < e) {
                const float px = a.sorted_xyz[3 * i], py = a.sorted_xyz[3 * i + 1], pz = a.sorted_xyz[3 * i + 2];
                const int3 pc = grid_cell(px, py, pz, a.inv_cell);
                // bucket collisions: only accept points that really live in the cell being visited (also prevents
                // duplicates when two of the 27 cells share a bucket)
                hit = (pc.x == c.x && pc.y == c.y && pc.z == c.z) && (sqdist_exact(px, py, pz, qx, qy, qz) < r2);
                src = a.sorted_idx[i];
                if (hit && a.b_src) hit = (a.b_src[src] == qb);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int pos = n_hits + __popc(bal & ((1u << lane) - 1u));
                if (pos < kGridMaxHits) hits[pos] = src;
            }
            n_hits += __popc(bal);
        }
        __syncwarp();
        int base = 0, seg_end = 0x7fffffff;
        if (FILL) { base = row_ptr[d]; seg_end = row_ptr[d + 1]; }
        const long long ex = (a.excl_mode == 1) ? a.excl[d] : -1;
        int cnt_keep = 0;
        if (n_hits > kGridMaxHits) {
            // rare, very dense neighbourhood: ordered brute-force scan (same code path as radius_kernel)
            int cnt_all = 0;
            for (int c0 = 0; c0 < a.n_src && cnt_all < a.max_nb; c0 += 32) {
                const int i = c0 + lane;
                bool hit = false, excluded = false;
                if (i < a.n_src) {
                    hit = sqdist_exact(a.x_src[3 * i], a.x_src[3 * i + 1], a.x_src[3 * i + 2], qx, qy, qz) < r2;
                    if (a.b_src) hit = hit && (a.b_src[i] == qb);
                    if (a.excl_mode == 1) excluded = ((long long)i == ex);
                    else if (a.excl_mode == 2) excluded = (i == d);
                    else if (a.excl_mode == 3) excluded = (a.excl[i] == (long long)d);
                }
                const unsigned lt = (1u << lane) - 1u;
                const unsigned bal_all = __ballot_sync(0xffffffffu, hit);
                const bool keep = hit && !excluded && (cnt_all + __popc(bal_all & lt) < a.max_nb);
                const unsigned bal_keep = __ballot_sync(0xffffffffu, keep);
                if (FILL && keep) {
                    const int pos = base + cnt_keep + __popc(bal_keep & lt);
                    if (pos < seg_end) { edge_src[pos] = i; edge_dst[pos] = d; }
                }
                cnt_all += __popc(bal_all);
                cnt_keep += __popc(bal_keep);
            }
        } else {
            // bitonic sort of the hit list (padded to a power of two with INT_MAX), ascending source index
            int np2 = 32;
            while (np2 < n_hits) np2 <<= 1;
            for (int i = n_hits + lane; i < np2; i += 32) hits[i] = 0x7fffffff;
            __syncwarp();
            for (int k = 2; k <= np2; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = lane; i < np2; i += 32) {
                        const int p = i ^ j;
                        if (p > i) {
                            const int vi = hits[i], vp = hits[p];
                            const bool up = ((i & k) == 0);
                            if ((vi > vp) == up) { hits[i] = vp; hits[p] = vi; }
                        }
                    }
                    __syncwarp();
                }
            }
            const int n_all = min(n_hits, a.max_nb);             // torch_cluster keeps the first max_nb in index order
            for (int i0 = 0; i0 < n_all; i0 += 32) {
                const int i = i0 + lane;
                bool keep = false;
                int src = 0;
                if (i < n_all) {
                    src = hits[i];
                    keep = true;
                    if (a.excl_mode == 1) keep = ((long long)src != ex);
                    else if (a.excl_mode == 2) keep = (src != d);
                    else if (a.excl_mode == 3) keep = (a.excl[src] != (long long)d);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, keep);
                if (FILL && keep) {
                    const int pos = base + cnt_keep + __popc(bal & ((1u << lane) - 1u));
                    if (pos < seg_end) { edge_src[pos] = src; edge_dst[pos] = d; }
                }
                cnt_keep += __popc(bal);
            }
        }
        if (!FILL && lane == 0) counts[d] = cnt_keep;
    }
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_fps(const float* x, int n, int m, int start, const long long* start_dev, int idx_base, long long* out_idx,
                        float* scratch_dist, cudaStream_t stream) {
    if (!x || !out_idx || n <= 0 || m <= 0 || m > n || start < 0 || start >= n) return DEDF_ERR_ARG;
    const size_t smem = (size_t)3 * n * sizeof(float);
    if (n >= 4096 && n <= 16384) {
        // cluster launch: 8 CTAs x 256 threads
        const int ptc = (n + kFpsClusterSize * kFpsClusterThreads - 1) / (kFpsClusterSize * kFpsClusterThreads);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(kFpsClusterSize); cfg.blockDim = dim3(kFpsClusterThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kFpsClusterSize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        cudaError_t err;
#define DEDF_FPSC_CASE(PT)                                                                                             \
        {                                                                                                              \
            static bool done = false;                                                                                  \
            if (!done) { cudaFuncSetAttribute(fps_cluster_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done = true; } \
            err = cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PT>, x, n, m, start, start_dev, idx_base, out_idx);      \
        }
        if (ptc <= 2) DEDF_FPSC_CASE(2)
        else if (ptc <= 3) DEDF_FPSC_CASE(3)
        else if (ptc <= 4) DEDF_FPSC_CASE(4)
        else if (ptc <= 5) DEDF_FPSC_CASE(5)
        else if (ptc <= 6) DEDF_FPSC_CASE(6)
        else DEDF_FPSC_CASE(8)
#undef DEDF_FPSC_CASE
        if (err != cudaSuccess) return DEDF_ERR_LAUNCH;
        DEDF_CHECK_LAUNCH();
        return DEDF_OK;
    }
    const int pt = (n + kFpsThreads - 1) / kFpsThreads;
#define DEDF_FPS_CASE(PT)                                                                                          \
    {                                                                                                              \
        static bool done = false;                                                                                  \
        if (!done) { cudaFuncSetAttribute(fps_kernel<PT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); done = true; } \
        fps_kernel<PT><<<1, kFpsThreads, smem, stream>>>(x, n, m, start, start_dev, idx_base, out_idx);                       \
    }
    if (pt <= 1) DEDF_FPS_CASE(1)
    else if (pt <= 2) DEDF_FPS_CASE(2)
    else if (pt <= 4) DEDF_FPS_CASE(4)
    else if (pt <= 6) DEDF_FPS_CASE(6)
    else if (pt <= 8) DEDF_FPS_CASE(8)
    else if (pt <= 10) DEDF_FPS_CASE(10)
    else if (pt <= 12) DEDF_FPS_CASE(12)
    else if (pt <= 16) DEDF_FPS_CASE(16)
#undef DEDF_FPS_CASE
    else {
        if (!scratch_dist) return DEDF_ERR_ARG;
        fps_kernel_large<<<1, kFpsThreads, 0, stream>>>(x, n, m, start, start_dev, idx_base, out_idx, scratch_dist);
    }
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

static int fill_radius_args(RadiusArgs& a, const float* x_src, const float* x_dst, int n_dst, int n_scales,
                            const int* src_off, const float* r, const long long* b_src, const long long* b_dst,
                            int excl_mode, const long long* excl, int max_nb) {
    if (!x_src || !x_dst || !src_off || !r || n_dst < 0 || n_scales < 1 || n_scales > DEDF_MAX_SCALES || max_nb < 1)
        return DEDF_ERR_ARG;
    if ((excl_mode == 1 || excl_mode == 3) && !excl) return DEDF_ERR_ARG;
    if (excl_mode < 0 || excl_mode > 3) return DEDF_ERR_ARG;
    if ((b_src == nullptr) != (b_dst == nullptr)) return DEDF_ERR_ARG;
    a.x_src = x_src; a.x_dst = x_dst; a.b_src = b_src; a.b_dst = b_dst; a.excl = excl;
    a.n_dst = n_dst; a.n_scales = n_scales; a.max_nb = max_nb; a.excl_mode = excl_mode;
    for (int s = 0; s <= n_scales; ++s) a.src_off[s] = src_off[s];
    for (int s = 0; s < n_scales; ++s) a.r[s] = r[s];
    return DEDF_OK;
}

extern "C" int dedf_radius_count(const float* x_src, const float* x_dst, int n_dst, int n_scales, const int* src_off,
                                 const float* r, const long long* b_src, const long long* b_dst, int excl_mode,
                                 const long long* excl, int max_nb, int* counts, int* row_ptr, int capacity, int* n_edges_out,
                                 int* overflow, cudaStream_t stream) {
    RadiusArgs a;
    int rc = fill_radius_args(a, x_src, x_dst, n_dst, n_scales, src_off, r, b_src, b_dst, excl_mode, excl, max_nb);
    if (rc) return rc;
    if (!counts || !row_ptr) return DEDF_ERR_ARG;
    const long long items = (long long)n_dst * n_scales;
    if (items > 0) {
        launch_radius<false>(a, items, counts, nullptr, nullptr, nullptr, stream);
        DEDF_CHECK_LAUNCH();
    }
    launch_scan(counts, (int)items, row_ptr, capacity, n_edges_out, overflow, stream);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_radius_fill(const float* x_src, const float* x_dst, int n_dst, int n_scales, const int* src_off,
                                const float* r, const long long* b_src, const long long* b_dst, int excl_mode,
                                const long long* excl, int max_nb, const int* row_ptr, int* edge_src, int* edge_dst,
                                cudaStream_t stream) {
    RadiusArgs a;
    int rc = fill_radius_args(a, x_src, x_dst, n_dst, n_scales, src_off, r, b_src, b_dst, excl_mode, excl, max_nb);
    if (rc) return rc;
    if (!row_ptr || !edge_src || !edge_dst) return DEDF_ERR_ARG;
    const long long items = (long long)n_dst * n_scales;
    if (items == 0) return DEDF_OK;
    launch_radius<true>(a, items, nullptr, row_ptr, edge_src, edge_dst, stream);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

// ---- grid-hash variant --------------------------------------------------------------------------------------------
static int fill_grid_args(GridArgs& a, const float* x_src, int n_src, const float* x_dst, int n_dst, float r, int n_buckets,
                          const int* bucket_start, const int* sorted_idx, const float* sorted_xyz, const long long* b_src,
                          const long long* b_dst, int excl_mode, const long long* excl, int max_nb) {
    if (!x_src || !x_dst || !bucket_start || !sorted_idx || !sorted_xyz || n_src < 0 || n_dst < 0 || r <= 0.f || max_nb < 1) return DEDF_ERR_ARG;
    if (n_buckets < 32 || (n_buckets & (n_buckets - 1))) return DEDF_ERR_ARG;
    if ((excl_mode == 1 || excl_mode == 3) && !excl) return DEDF_ERR_ARG;
    if (excl_mode < 0 || excl_mode > 3) return DEDF_ERR_ARG;
    if ((b_src == nullptr) != (b_dst == nullptr)) return DEDF_ERR_ARG;
    a.x_src = x_src; a.x_dst = x_dst; a.b_src = b_src; a.b_dst = b_dst; a.excl = excl;
    a.bucket_start = bucket_start; a.sorted_idx = sorted_idx; a.sorted_xyz = sorted_xyz;
    a.n_src = n_src; a.n_dst = n_dst; a.max_nb = max_nb; a.excl_mode = excl_mode;
    a.r = r; a.inv_cell = 1.0f / (r * 1.001f); a.mask = (unsigned)(n_buckets - 1);
    return DEDF_OK;
}

extern "C" int dedf_grid_build(const float* x_src, int n_src, float r, int n_buckets, int* bucket_cnt, int* bucket_start,
                               int* sorted_idx, float* sorted_xyz, cudaStream_t stream) {
    if (!x_src || !bucket_cnt || !bucket_start || !sorted_idx || !sorted_xyz || n_src < 0 || r <= 0.f) return DEDF_ERR_ARG;
    if (n_buckets < 32 || (n_buckets & (n_buckets - 1))) return DEDF_ERR_ARG;
    const float inv_cell = 1.0f / (r * 1.001f);
    const unsigned mask = (unsigned)(n_buckets - 1);
    cudaMemsetAsync(bucket_cnt, 0, sizeof(int) * n_buckets, stream);
    if (n_src > 0) { launch_pdl(grid_count_kernel, dim3(grid_for(n_src, 256, kNumSMs * 8)), dim3(256), 0, stream, x_src, n_src, inv_cell, mask, bucket_cnt); DEDF_CHECK_LAUNCH(); }
    launch_scan(bucket_cnt, n_buckets, bucket_start, 0, nullptr, nullptr, stream);
    DEDF_CHECK_LAUNCH();
    cudaMemsetAsync(bucket_cnt, 0, sizeof(int) * n_buckets, stream);
    if (n_src > 0) { launch_pdl(grid_fill_kernel, dim3(grid_for(n_src, 256, kNumSMs * 8)), dim3(256), 0, stream, x_src, n_src, inv_cell, mask, bucket_start, bucket_cnt, sorted_idx); DEDF_CHECK_LAUNCH(); }
    if (n_src > 0) launch_pdl(grid_gather_kernel, dim3(grid_for(n_src, 256, kNumSMs * 8)), dim3(256), 0, stream, x_src, n_src, sorted_idx, sorted_xyz);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_radius_grid_count(const float* x_src, int n_src, const float* x_dst, int n_dst, float r, int n_buckets,
                                      const int* bucket_start, const int* sorted_idx, const float* sorted_xyz,
                                      const long long* b_src, const long long* b_dst, int excl_mode, const long long* excl,
                                      int max_nb, int* counts, int* row_ptr, int capacity, int* n_edges_out, int* overflow,
                                      cudaStream_t stream) {
    GridArgs a;
    int rc = fill_grid_args(a, x_src, n_src, x_dst, n_dst, r, n_buckets, bucket_start, sorted_idx, sorted_xyz, b_src, b_dst, excl_mode, excl, max_nb);
    if (rc) return rc;
    if (!counts || !row_ptr) return DEDF_ERR_ARG;
    if (n_dst > 0) { launch_pdl((radius_grid_kernel<false>), dim3(grid_for(n_dst, kGridWarps, kNumSMs * 4)), dim3(kGridWarps * 32), 0, stream, a, counts, nullptr, nullptr, nullptr); DEDF_CHECK_LAUNCH(); }
    launch_scan(counts, n_dst, row_ptr, capacity, n_edges_out, overflow, stream);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_radius_grid_fill(const float* x_src, int n_src, const float* x_dst, int n_dst, float r, int n_buckets,
                                     const int* bucket_start, const int* sorted_idx, const float* sorted_xyz,
                                     const long long* b_src, const long long* b_dst, int excl_mode, const long long* excl,
                                     int max_nb, const int* row_ptr, int* edge_src, int* edge_dst, cudaStream_t stream) {
    GridArgs a;
    int rc = fill_grid_args(a, x_src, n_src, x_dst, n_dst, r, n_buckets, bucket_start, sorted_idx, sorted_xyz, b_src, b_dst, excl_mode, excl, max_nb);
    if (rc) return rc;
    if (!row_ptr || !edge_src || !edge_dst) return DEDF_ERR_ARG;
    if (n_dst == 0) return DEDF_OK;
    launch_pdl((radius_grid_kernel<true>), dim3(grid_for(n_dst, kGridWarps, kNumSMs * 4)), dim3(kGridWarps * 32), 0, stream, a, nullptr, row_ptr, edge_src, edge_dst);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
