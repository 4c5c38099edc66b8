// Point-cloud pre-processing in front of the score network:
//   voxel_filter  (edf_interface/edf_interface/data/pcd_utils.py:123-152; called by preprocess.downsample, preprocess.py:69-80,
//                  from agent.py:122-124 on every request: 50-60 k raw points -> ~5 k voxels)
// The reference ravels the voxel indices on the HOST (np.ravel_multi_index, pcd_utils.py:133) and scatter-adds into a dense
// grid with torch_scatter atomics.  Here everything stays on the device: dense voxel counts -> two scans (occupied-voxel
// rank, point offsets) -> counting sort of the points -> one thread per occupied voxel sums its points in ASCENDING POINT
// INDEX order, so the result is deterministic and bit-identical to a sequential CPU scatter (oracle/graph.py voxel_filter).
// Output order = ascending ravelled voxel index (C order: x slowest), exactly the reference's `nonzero()` order.
#include "common.cuh"
#include "../../include/dedf.h"

namespace dedf {

// per-axis min / max of the cloud: one CTA
__global__ void __launch_bounds__(1024) bbox_kernel(const float* __restrict__ p, int n, float* __restrict__ mins, float* __restrict__ maxs) {
    __shared__ float smin[3][32], smax[3][32];
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int d = 0; d < 3; ++d) { const float v = p[3 * (size_t)i + d]; lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) { lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o)); hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o)); }
        if (lane == 0) { smin[d][warp] = lo[d]; smax[d][warp] = hi[d]; }
    }
    __syncthreads();
    if (warp == 0) {
        for (int d = 0; d < 3; ++d) {
            float a = (lane < (int)(blockDim.x >> 5)) ? smin[d][lane] : INFINITY, b = (lane < (int)(blockDim.x >> 5)) ? smax[d][lane] : -INFINITY;
            for (int o = 16; o > 0; o >>= 1) { a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o)); b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o)); }
            if (lane == 0) { mins[d] = a; maxs[d] = b; }
        }
    }
}

// trunc((p - min) / voxel) per axis, exactly as torch.div(points - mins, voxel_size, rounding_mode='trunc')
__device__ __forceinline__ int vox_coord(float p, float mn, float vs) { return (int)truncf(__fdiv_rn(__fsub_rn(p, mn), vs)); }

__global__ void voxel_count_kernel(const float* __restrict__ p, int n, const float* __restrict__ mins, float vs, int sy, int sz,
                                   int* __restrict__ key, int* __restrict__ cnt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ix = vox_coord(p[3 * (size_t)i], mins[0], vs), iy = vox_coord(p[3 * (size_t)i + 1], mins[1], vs), iz = vox_coord(p[3 * (size_t)i + 2], mins[2], vs);
        const int k = (ix * sy + iy) * sz + iz;
        key[i] = k;
        atomicAdd(cnt + k, 1);
    }
}

// single-CTA exclusive scans over the dense grid: off[k] = sum_{j<k} cnt[j] ; rank[k] = #{j < k : cnt[j] > 0}; totals -> n_occ
__global__ void __launch_bounds__(1024) voxel_scan_kernel(const int* __restrict__ cnt, int S, int* __restrict__ off, int* __restrict__ rank,
                                                         int* __restrict__ n_occ) {
    __shared__ int s_a[1024], s_b[1024];
    const int tid = threadIdx.x, per = (S + 1023) / 1024;
    const int b = tid * per, e = min(S, b + per);
    int a0 = 0, b0 = 0;
    for (int i = b; i < e; ++i) { a0 += cnt[i]; b0 += cnt[i] > 0; }
    s_a[tid] = a0; s_b[tid] = b0;
    __syncthreads();
    if (tid == 0) {
        int ra = 0, rb = 0;
        for (int i = 0; i < 1024; ++i) { const int ta = s_a[i], tb = s_b[i]; s_a[i] = ra; s_b[i] = rb; ra += ta; rb += tb; }
        *n_occ = rb;
    }
    __syncthreads();
    int ra = s_a[tid], rb = s_b[tid];
    for (int i = b; i < e; ++i) { off[i] = ra; rank[i] = rb; ra += cnt[i]; rb += cnt[i] > 0; }
}

__global__ void voxel_fill_kernel(const int* __restrict__ key, int n, const int* __restrict__ off, int* __restrict__ cursor, int* __restrict__ sorted) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = key[i];
        sorted[off[k] + atomicAdd(cursor + k, 1)] = i;
    }
}

// one thread per grid cell; occupied cells reduce their points in ascending point-index order (selection by repeated minimum)
__global__ void voxel_reduce_kernel(const float* __restrict__ p, const float* __restrict__ f, int F, const int* __restrict__ cnt,
                                    const int* __restrict__ off, const int* __restrict__ rank, const int* __restrict__ sorted, int S,
                                    int sy, int sz, const float* __restrict__ mins, float vs, int center,
                                    float* __restrict__ out_p, float* __restrict__ out_f) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < S; k += gridDim.x * blockDim.x) {
        const int c = cnt[k];
        if (c == 0) continue;
        const int* idx = sorted + off[k];
        float sp[3] = {0.f, 0.f, 0.f};
        float* of = out_f + (size_t)rank[k] * F;
        for (int j = 0; j < F; ++j) of[j] = 0.f;
        int last = -1;
        for (int t = 0; t < c; ++t) {
            int cur = 0x7fffffff;
            for (int u = 0; u < c; ++u) { const int v = idx[u]; if (v > last && v < cur) cur = v; }
            last = cur;
            for (int d = 0; d < 3; ++d) sp[d] = __fadd_rn(sp[d], p[3 * (size_t)cur + d]);
            for (int j = 0; j < F; ++j) of[j] = __fadd_rn(of[j], f[(size_t)cur * F + j]);
        }
        const float fc = (float)c;
        for (int j = 0; j < F; ++j) of[j] = __fdiv_rn(of[j], fc);
        float* op = out_p + (size_t)rank[k] * 3;
        if (center) {
            const int iz = k % sz, iy = (k / sz) % sy, ix = k / (sz * sy);
            const int ii[3] = {ix, iy, iz};
            for (int d = 0; d < 3; ++d) op[d] = __fadd_rn(__fadd_rn(__fmul_rn((float)ii[d], vs), mins[d]), vs * 0.5f);
        } else {
            for (int d = 0; d < 3; ++d) op[d] = __fdiv_rn(sp[d], fc);
        }
    }
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_bbox(const float* points, int n, float* mins, float* maxs, cudaStream_t stream) {
    if (!points || !mins || !maxs || n <= 0) return DEDF_ERR_ARG;
    bbox_kernel<<<1, 1024, 0, stream>>>(points, n, mins, maxs);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_voxel_count(const float* points, int n, const float* mins, float voxel_size, int sx, int sy, int sz, int* key,
                                int* cnt_zeroed, int* off, int* rank, int* n_occupied, cudaStream_t stream) {
    if (!points || !mins || !key || !cnt_zeroed || !off || !rank || !n_occupied || n <= 0 || voxel_size <= 0.f) return DEDF_ERR_ARG;
    if (sx <= 0 || sy <= 0 || sz <= 0 || (long long)sx * sy * sz > (1ll << 27)) return DEDF_ERR_UNSUPPORTED;
    const int S = sx * sy * sz;
    voxel_count_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, stream>>>(points, n, mins, voxel_size, sy, sz, key, cnt_zeroed);
    DEDF_CHECK_LAUNCH();
    voxel_scan_kernel<<<1, 1024, 0, stream>>>(cnt_zeroed, S, off, rank, n_occupied);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_voxel_reduce(const float* points, const float* feats, int n, int F, const float* mins, float voxel_size, int sx,
                                 int sy, int sz, const int* key, const int* cnt, const int* off, const int* rank, int* cursor_zeroed,
                                 int* sorted, int center, float* out_points, float* out_feats, cudaStream_t stream) {
    if (!points || !feats || !mins || !key || !cnt || !off || !rank || !cursor_zeroed || !sorted || !out_points || !out_feats) return DEDF_ERR_ARG;
    if (n <= 0 || F <= 0 || sx <= 0 || sy <= 0 || sz <= 0) return DEDF_ERR_ARG;
    const int S = sx * sy * sz;
    voxel_fill_kernel<<<grid_for(n, 256, kNumSMs * 8), 256, 0, stream>>>(key, n, off, cursor_zeroed, sorted);
    DEDF_CHECK_LAUNCH();
    voxel_reduce_kernel<<<grid_for(S, 256, kNumSMs * 16), 256, 0, stream>>>(points, feats, F, cnt, off, rank, sorted, S, sy, sz, mins, voxel_size,
                                                                           center, out_points, out_feats);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
