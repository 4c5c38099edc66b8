// Quaternion / SO(3) device helpers shared by the score-head kernels (head.cu) and the training path (train.cu).
#pragma once
#include "common.cuh"

namespace dedf {

// ---------------------------------------------------------------------------
// quaternion helpers (transforms.py:83-110, :113-163)
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void quat_to_matrix(const T* q, T* R) {   // R row-major 3x3
    const T r = q[0], i = q[1], j = q[2], k = q[3];
    const T two_s = T(2) / (r * r + i * i + j * j + k * k);
    R[0] = T(1) - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r); R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r); R[4] = T(1) - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r); R[7] = two_s * (j * k + i * r); R[8] = T(1) - two_s * (i * i + j * j);
}

// quaternion_apply(q, p) = vec(q * (0,p) * conj(q))   -- not normalised, like the reference
template <typename T>
__device__ __forceinline__ void quat_apply(const T* q, const T* p, T* o) {
    const T aw = q[0], ax = q[1], ay = q[2], az = q[3];
    // t = q * (0, p)
    const T tw = -ax * p[0] - ay * p[1] - az * p[2];
    const T tx = aw * p[0] + ay * p[2] - az * p[1];
    const T ty = aw * p[1] - ax * p[2] + az * p[0];
    const T tz = aw * p[2] + ax * p[1] - ay * p[0];
    // o = t * conj(q)
    o[0] = -tw * ax + tx * aw - ty * az + tz * ay;
    o[1] = -tw * ay + tx * az + ty * aw - tz * ax;
    o[2] = -tw * az - tx * ay + ty * ax + tz * aw;
}

// transform_points (edf_interface/.../pcd_utils.py:55-81): x' = quaternion_apply(q, x) + t with the RAW (un-normalised) q.
// The final add is an explicit __fadd_rn so that no caller's context can contract it into the rotation's last FMA: every kernel
// that transforms a query point (query_transform_kernel, head_front_kernel) produces the same bits.
__device__ __forceinline__ void transform_point_f(const float* T7, const float* p, float* out) {
    const float aw = T7[0], ax = T7[1], ay = T7[2], az = T7[3];
#define M_(a, b) __fmul_rn(a, b)
    // quat_apply<float>, every operation rounded on its own (left to right as written there)
    const float tw = __fsub_rn(__fsub_rn(M_(-ax, p[0]), M_(ay, p[1])), M_(az, p[2]));
    const float tx = __fsub_rn(__fadd_rn(M_(aw, p[0]), M_(ay, p[2])), M_(az, p[1]));
    const float ty = __fadd_rn(__fsub_rn(M_(aw, p[1]), M_(ax, p[2])), M_(az, p[0]));
    const float tz = __fsub_rn(__fadd_rn(M_(aw, p[2]), M_(ax, p[1])), M_(ay, p[0]));
    const float o0 = __fadd_rn(__fsub_rn(__fadd_rn(M_(-tw, ax), M_(tx, aw)), M_(ty, az)), M_(tz, ay));
    const float o1 = __fsub_rn(__fadd_rn(__fadd_rn(M_(-tw, ay), M_(tx, az)), M_(ty, aw)), M_(tz, ax));
    const float o2 = __fadd_rn(__fadd_rn(__fsub_rn(M_(-tw, az), M_(tx, ay)), M_(ty, ax)), M_(tz, aw));
#undef M_
    out[0] = __fadd_rn(o0, T7[4]); out[1] = __fadd_rn(o1, T7[5]); out[2] = __fadd_rn(o2, T7[6]);
}

// D^2(R): Y2_a(R x) = sum_b D_ab Y2_b(x), via the symmetric traceless matrices M_a of the l=2 harmonics
// (|M_a|_F^2 = 7.5 for every a):  D_ab = <R^T M_a R, M_b>_F / 7.5
__device__ __forceinline__ void wigner_d2_from_R(const float* R, float* D) {
    const float s15 = 3.872983346207417f, s5 = 2.23606797749979f;
    // M_a as (xx, yy, zz, xy, xz, yz)
    const float M[5][6] = {
        {0.f, 0.f, 0.f, 0.f, 0.5f * s15, 0.f},          // sqrt15 x z
        {0.f, 0.f, 0.f, 0.5f * s15, 0.f, 0.f},          // sqrt15 x y
        {-0.5f * s5, s5, -0.5f * s5, 0.f, 0.f, 0.f},    // sqrt5 (y^2 - (x^2+z^2)/2)
        {0.f, 0.f, 0.f, 0.f, 0.f, 0.5f * s15},          // sqrt15 y z
        {-0.5f * s15, 0.f, 0.5f * s15, 0.f, 0.f, 0.f},  // sqrt15/2 (z^2 - x^2)
    };
#pragma unroll
    for (int a = 0; a < 5; ++a) {
        // full symmetric M
        const float m[3][3] = {{M[a][0], M[a][3], M[a][4]}, {M[a][3], M[a][1], M[a][5]}, {M[a][4], M[a][5], M[a][2]}};
        float t[3][3], n[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) t[i][j] = m[i][0] * R[0 * 3 + j] + m[i][1] * R[1 * 3 + j] + m[i][2] * R[2 * 3 + j];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) n[i][j] = R[0 * 3 + i] * t[0][j] + R[1 * 3 + i] * t[1][j] + R[2 * 3 + i] * t[2][j];
#pragma unroll
        for (int b = 0; b < 5; ++b) {
            const float ip = M[b][0] * n[0][0] + M[b][1] * n[1][1] + M[b][2] * n[2][2] +
                             2.f * (M[b][3] * n[0][1] + M[b][4] * n[0][2] + M[b][5] * n[1][2]);
            D[a * 5 + b] = ip * (1.0f / 7.5f);
        }
    }
}


}  // namespace dedf
