// Collision-aware pre-place trajectory optimisation: the post-processing step that consumes the sampled place
// poses.  Replaces, for /root/reference/edf_interface/edf_interface/utils/collision_utils.py:
//   _check_pcd_collision            (:18-34)    torch_cluster.radius + scatter_sum              -> collision_check_kernel
//   _pcd_energy                     (:40-110)   torch_cluster.knn / radius, L1 energy, autograd  -> collision_energy_kernel
//   _se3_adjoint_lie_grad, _optimize_pcd_collision_once (:116-196): adjoint, scaling, se3 exp map, pose product -> collision_step_kernel
// The reference differentiates the energy with autograd through three infinitesimal cross products; the gradient at zero is closed
// form (d y / d trans = I, d y / d rot_a = e_a x y), so energy and gradient come out of ONE pass over the neighbours.
// One warp per (pose, grasp point); the scene cloud is walked 32 points at a time (brute force: a few 10^3 scene points, ~10^4
// queries per step).  fp32 throughout; per-pose sums by atomics (the reference's scatter_sum / index_add are unordered too).
#include "common.cuh"
#include "so3.cuh"
#include "../../include/dedf.h"

namespace dedf {

struct ColQuery { float x, y, z; };

// transformed grasp point of (pose, j): y is (n_y, 3) shared by all poses (y_pose_stride = 0) or (n_pose, n_y, 3); Ts NULL: y is already
// in the scene frame
__device__ __forceinline__ ColQuery col_query(const float* __restrict__ y, long long y_pose_stride, const float* __restrict__ Ts, int pose,
                                              int j) {
    const float* p = y + (size_t)pose * y_pose_stride + (size_t)j * 3;
    float o[3] = {p[0], p[1], p[2]};
    if (Ts) transform_point_f(Ts + (size_t)pose * 7, p, o);
    return ColQuery{o[0], o[1], o[2]};
}

__global__ void __launch_bounds__(256) collision_check_kernel(const float* __restrict__ x, int n_x, const float* __restrict__ y,
                                                             long long y_pose_stride, const float* __restrict__ Ts, int n_pose, int n_y,
                                                             float r2, int* __restrict__ hit) {
    const int lane = threadIdx.x & 31;
    const long long n_q = (long long)n_pose * n_y, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_q; w += n_warps) {
        const int pose = (int)(w / n_y), j = (int)(w % n_y);
        if (hit[pose]) continue;                         // another point of this pose already collides (benign race: monotone flag)
        const ColQuery q = col_query(y, y_pose_stride, Ts, pose, j);
        bool any = false;
        for (int i0 = 0; i0 < n_x && !any; i0 += 32) {
            const int i = i0 + lane;
            bool in = false;
            if (i < n_x) {
                const float dx = x[3 * i] - q.x, dy = x[3 * i + 1] - q.y, dz = x[3 * i + 2] - q.z;
                in = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < r2;      // torch_cluster.radius: d^2 < r^2
            }
            any = __any_sync(0xffffffffu, in);
        }
        if (any && lane == 0) atomicOr(hit + pose, 1);
    }
}

// one neighbour's contribution: e = c / (r1 + eps c), r1 = |x - y|_1;  d e / d y = c / (r1 + eps c)^2 * sign(x - y)
struct ColAcc { float e, gx, gy, gz; };
__device__ __forceinline__ void col_add(ColAcc& a, float dx, float dy, float dz, float c, float eps_c) {
    const float r1 = fabsf(dx) + fabsf(dy) + fabsf(dz);
    const float inv = 1.0f / (r1 + eps_c);
    const float e = c * inv, g = e * inv;
    a.e += e;
    a.gx += g * ((dx > 0.f) - (dx < 0.f)); a.gy += g * ((dy > 0.f) - (dy < 0.f)); a.gz += g * ((dz > 0.f) - (dz < 0.f));
}

// method 0: 'knn' (the k nearest scene points by squared distance, then those with L1 distance <= cutoff); 1: 'radius' (the first k
// scene points in index order with d^2 < cutoff^2, no L1 filter)
__global__ void __launch_bounds__(256) collision_energy_kernel(const float* __restrict__ x, int n_x, const float* __restrict__ y,
                                                              long long y_pose_stride, const float* __restrict__ Ts, int n_pose, int n_y,
                                                              float c, int k, float eps, int method, float* __restrict__ energy,
                                                              float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const float c2 = c * c, eps_c = eps * c;
    const long long n_q = (long long)n_pose * n_y, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n_q; w += n_warps) {
        const int pose = (int)(w / n_y), j = (int)(w % n_y);
        const ColQuery q = col_query(y, y_pose_stride, Ts, pose, j);
        ColAcc a{0.f, 0.f, 0.f, 0.f};
        if (method == 1) {
            int taken = 0;
            for (int i0 = 0; i0 < n_x && taken < k; i0 += 32) {
                const int i = i0 + lane;
                float dx = 0.f, dy = 0.f, dz = 0.f;
                bool in = false;
                if (i < n_x) {
                    dx = x[3 * i] - q.x; dy = x[3 * i + 1] - q.y; dz = x[3 * i + 2] - q.z;
                    in = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < c2;
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in && taken + __popc(m & ((1u << lane) - 1u)) < k) col_add(a, dx, dy, dz, c, eps_c);
                taken += __popc(m);
            }
        } else {
            // pass 1: every scene point with L1 <= c (hence d^2 <= c^2) contributes, provided it is among the k nearest -- which it is
            // whenever at most k points lie within d^2 <= c^2 (or the cloud has at most k points)
            int n_c = 0;
            for (int i0 = 0; i0 < n_x; i0 += 32) {
                const int i = i0 + lane;
                if (i < n_x) {
                    const float dx = x[3 * i] - q.x, dy = x[3 * i + 1] - q.y, dz = x[3 * i + 2] - q.z;
                    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (d2 <= c2) {
                        ++n_c;
                        if (fabsf(dx) + fabsf(dy) + fabsf(dz) <= c) col_add(a, dx, dy, dz, c, eps_c);
                    }
                }
            }
            n_c = __reduce_add_sync(0xffffffffu, n_c);
            if (n_c > k && n_x > k) {
                // dense neighbourhood: the k-th smallest squared distance by a bit-wise radix select over the float bit patterns
                // (non-negative floats order like unsigned integers), then the contributions of the points up to it
                unsigned cur = 0;
                for (int bit = 30; bit >= 0; --bit) {
                    const unsigned T = cur | (1u << bit);
                    int cnt = 0;
                    for (int i0 = 0; i0 < n_x; i0 += 32) {
                        const int i = i0 + lane;
                        if (i < n_x) {
                            const float dx = x[3 * i] - q.x, dy = x[3 * i + 1] - q.y, dz = x[3 * i + 2] - q.z;
                            const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                            cnt += (__float_as_uint(d2) < T);
                        }
                    }
                    if (__reduce_add_sync(0xffffffffu, cnt) < k) cur = T;      // the k-th smallest is >= T
                }
                a = ColAcc{0.f, 0.f, 0.f, 0.f};
                for (int i0 = 0; i0 < n_x; i0 += 32) {
                    const int i = i0 + lane;
                    if (i < n_x) {
                        const float dx = x[3 * i] - q.x, dy = x[3 * i + 1] - q.y, dz = x[3 * i + 2] - q.z;
                        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                        if (__float_as_uint(d2) <= cur && fabsf(dx) + fabsf(dy) + fabsf(dz) <= c) col_add(a, dx, dy, dz, c, eps_c);
                    }
                }
            }
        }
        a.e = warp_sum(a.e); a.gx = warp_sum(a.gx); a.gy = warp_sum(a.gy); a.gz = warp_sum(a.gz);
        if (lane == 0 && a.e != 0.f) {
            atomicAdd(energy + pose, a.e);
            if (grad) {
                float* gp = grad + (size_t)pose * 6;     // (rot xyz | trans xyz): d y / d rot_a = e_a x y  ->  rot = y x g
                atomicAdd(gp + 0, q.y * a.gz - q.z * a.gy); atomicAdd(gp + 1, q.z * a.gx - q.x * a.gz); atomicAdd(gp + 2, q.x * a.gy - q.y * a.gx);
                atomicAdd(gp + 3, a.gx); atomicAdd(gp + 4, a.gy); atomicAdd(gp + 5, a.gz);
            }
        }
    }
}

// one thread per pose: adjoint of the world-frame gradient (collision_utils.py:140-145), scaling and sign (:185-187), se3 exponential map
// (pytorch3d, transforms.py:425-561, squared rotation norm clamped at 1e-4), matrix_to_quaternion (:23-80), T_new = T * exp(disp)
// (se3.py:13-23).  grad NULL: no step, only the quaternion normalisation of the product with the identity is skipped too (copy).
__global__ void collision_step_kernel(const float* __restrict__ Ts, const float* __restrict__ grad, int n_pose, float dt, float c,
                                      float* __restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pose) return;
    const float* T = Ts + (size_t)t * 7;
    const float* g = grad + (size_t)t * 6;
    const float q[4] = {T[0], T[1], T[2], T[3]}, qi[4] = {T[0], -T[1], -T[2], -T[3]};
    const float p[3] = {T[4], T[5], T[6]};
    // adjoint: rot' = q^-1 (g_rot - p x g_trans), trans' = q^-1 g_trans
    const float gr[3] = {g[0] - (p[1] * g[5] - p[2] * g[4]), g[1] - (p[2] * g[3] - p[0] * g[5]), g[2] - (p[0] * g[4] - p[1] * g[3])};
    const float gt[3] = {g[3], g[4], g[5]};
    float ar[3], at[3];
    quat_apply<float>(qi, gr, ar);
    quat_apply<float>(qi, gt, at);
    const float s = -dt * c;
    const float w[3] = {ar[0] * s, ar[1] * s, ar[2] * s};                    // log rotation
    const float v[3] = {at[0] * c * s, at[1] * c * s, at[2] * c * s};        // log translation (scaled by cutoff_r once more)
    const float n2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    const float ang = sqrtf(fmaxf(n2, 1e-4f));
    const float sn = sinf(ang), cs = cosf(ang);
    const float f1 = sn / ang, f2 = (1.0f - cs) / (ang * ang), f3 = (ang - sn) / (ang * ang * ang);
    // K = hat(w), K2 = K K
    const float K[9] = {0.f, -w[2], w[1], w[2], 0.f, -w[0], -w[1], w[0], 0.f};
    float K2[9], R[9], V[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) K2[i * 3 + j] = K[i * 3] * K[j] + K[i * 3 + 1] * K[3 + j] + K[i * 3 + 2] * K[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const float id = (i % 4 == 0) ? 1.0f : 0.0f;
        R[i] = f1 * K[i] + f2 * K2[i] + id;
        V[i] = id + K[i] * f2 + K2[i] * f3;
    }
    const float tr[3] = {V[0] * v[0] + V[1] * v[1] + V[2] * v[2], V[3] * v[0] + V[4] * v[1] + V[5] * v[2], V[6] * v[0] + V[7] * v[1] + V[8] * v[2]};
    // matrix_to_quaternion: best-conditioned of four candidates
    float qa[4] = {1.0f + R[0] + R[4] + R[8], 1.0f + R[0] - R[4] - R[8], 1.0f - R[0] + R[4] - R[8], 1.0f - R[0] - R[4] + R[8]};
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
#pragma unroll
    for (int i = 1; i < 4; ++i) if (qa[i] > qa[best]) best = i;
    float cq[4];
    if (best == 0)      { cq[0] = qa[0] * qa[0]; cq[1] = R[7] - R[5]; cq[2] = R[2] - R[6]; cq[3] = R[3] - R[1]; }
    else if (best == 1) { cq[0] = R[7] - R[5]; cq[1] = qa[1] * qa[1]; cq[2] = R[3] + R[1]; cq[3] = R[2] + R[6]; }
    else if (best == 2) { cq[0] = R[2] - R[6]; cq[1] = R[3] + R[1]; cq[2] = qa[2] * qa[2]; cq[3] = R[5] + R[7]; }
    else                { cq[0] = R[3] - R[1]; cq[1] = R[6] + R[2]; cq[2] = R[7] + R[5]; cq[3] = qa[3] * qa[3]; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
    float dq[4] = {cq[0] / den, cq[1] / den, cq[2] / den, cq[3] / den};
    // _multiply(T, exp): normalise dq, x = q tr + p, q' = normalise(q dq)
    float nrm = sqrtf(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
#pragma unroll
    for (int i = 0; i < 4; ++i) dq[i] /= nrm;
    float xr[3];
    quat_apply<float>(q, tr, xr);
    float qq[4] = {q[0] * dq[0] - q[1] * dq[1] - q[2] * dq[2] - q[3] * dq[3],
                   q[0] * dq[1] + q[1] * dq[0] + q[2] * dq[3] - q[3] * dq[2],
                   q[0] * dq[2] - q[1] * dq[3] + q[2] * dq[0] + q[3] * dq[1],
                   q[0] * dq[3] + q[1] * dq[2] - q[2] * dq[1] + q[3] * dq[0]};
    nrm = sqrtf(qq[0] * qq[0] + qq[1] * qq[1] + qq[2] * qq[2] + qq[3] * qq[3]);
    float* o = out + (size_t)t * 7;
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = qq[i] / nrm;
    o[4] = xr[0] + p[0]; o[5] = xr[1] + p[1]; o[6] = xr[2] + p[2];
}

}  // namespace dedf

using namespace dedf;

extern "C" int dedf_collision_check(const float* x, int n_x, const float* y, long long y_pose_stride, const float* Ts, int n_pose, int n_y,
                                    float r, int* hit, cudaStream_t stream) {
    if (n_pose <= 0) return DEDF_OK;
    if ((!x && n_x > 0) || (!y && n_y > 0) || !hit || n_x < 0 || n_y < 0) return DEDF_ERR_ARG;
    cudaMemsetAsync(hit, 0, sizeof(int) * (size_t)n_pose, stream);
    if (n_x == 0 || n_y == 0) return DEDF_OK;
    const float r2 = r * r;
    collision_check_kernel<<<grid_for((long long)n_pose * n_y * 32, 256, kNumSMs * 8), 256, 0, stream>>>(x, n_x, y, y_pose_stride, Ts, n_pose,
                                                                                                        n_y, r2, hit);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_collision_energy(const float* x, int n_x, const float* y, long long y_pose_stride, const float* Ts, int n_pose, int n_y,
                                     float cutoff_r, int max_num_neighbors, float eps, int method, float* energy, float* grad,
                                     cudaStream_t stream) {
    if (n_pose <= 0) return DEDF_OK;
    if ((!x && n_x > 0) || (!y && n_y > 0) || !energy || n_x < 0 || n_y < 0 || max_num_neighbors < 1 || cutoff_r <= 0.f) return DEDF_ERR_ARG;
    if (method != 0 && method != 1) return DEDF_ERR_UNSUPPORTED;
    cudaMemsetAsync(energy, 0, sizeof(float) * (size_t)n_pose, stream);
    if (grad) cudaMemsetAsync(grad, 0, sizeof(float) * 6 * (size_t)n_pose, stream);
    if (n_x == 0 || n_y == 0) return DEDF_OK;
    collision_energy_kernel<<<grid_for((long long)n_pose * n_y * 32, 256, kNumSMs * 8), 256, 0, stream>>>(
        x, n_x, y, y_pose_stride, Ts, n_pose, n_y, cutoff_r, max_num_neighbors, eps, method, energy, grad);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}

extern "C" int dedf_collision_step(const float* Ts, const float* grad, int n_pose, float dt, float cutoff_r, float* Ts_out,
                                   cudaStream_t stream) {
    if (n_pose <= 0) return DEDF_OK;
    if (!Ts || !grad || !Ts_out) return DEDF_ERR_ARG;
    collision_step_kernel<<<(n_pose + 127) / 128, 128, 0, stream>>>(Ts, grad, n_pose, dt, cutoff_r, Ts_out);
    DEDF_CHECK_LAUNCH();
    return DEDF_OK;
}
