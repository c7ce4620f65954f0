// Device math shared by the apply / grid kernels.  The whole library is built
// with -fmad=false so that float expressions round exactly like the reference's
// host code (x86-64 baseline, no FMA contraction); fused multiply-adds are
// written explicitly (fma()) where fusion is wanted in double hot loops.
#pragma once
#include "common.cuh"

namespace arapgs {

struct Quat { float w, x, y, z; };

// Eigen::Quaternionf(Matrix3f) (Eigen/src/Geometry/Quaternion.h); m row-major.
// Restated from SURVEY Appendix B.4 — the sign convention the reference's
// `Eigen::Quaternionf q(dest_rot_o)` (GaussianView.cpp:3113) relies on.
__device__ __forceinline__ Quat quat_from_matrix(const float (&m)[3][3]) {
  Quat q;
  float t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0f) {
    t = sqrtf(t + 1.0f);
    q.w = 0.5f * t; t = 0.5f / t;
    q.x = (m[2][1] - m[1][2]) * t; q.y = (m[0][2] - m[2][0]) * t; q.z = (m[1][0] - m[0][1]) * t;
  } else if (m[0][0] >= m[1][1] && m[0][0] >= m[2][2]) {  // i = 0
    t = sqrtf(m[0][0] - m[1][1] - m[2][2] + 1.0f);
    q.x = 0.5f * t; t = 0.5f / t;
    q.w = (m[2][1] - m[1][2]) * t; q.y = (m[1][0] + m[0][1]) * t; q.z = (m[2][0] + m[0][2]) * t;
  } else if (m[1][1] > m[0][0] && m[1][1] >= m[2][2]) {   // i = 1
    t = sqrtf(m[1][1] - m[2][2] - m[0][0] + 1.0f);
    q.y = 0.5f * t; t = 0.5f / t;
    q.w = (m[0][2] - m[2][0]) * t; q.z = (m[2][1] + m[1][2]) * t; q.x = (m[0][1] + m[1][0]) * t;
  } else {                                                 // i = 2
    t = sqrtf(m[2][2] - m[0][0] - m[1][1] + 1.0f);
    q.z = 0.5f * t; t = 0.5f / t;
    q.w = (m[1][0] - m[0][1]) * t; q.x = (m[0][2] + m[2][0]) * t; q.y = (m[1][2] + m[2][1]) * t;
  }
  return q;
}
// squaredNorm over (x,y,z,w): (x^2+z^2)+(y^2+w^2) — same order as the oracle
__device__ __forceinline__ float quat_n2(const Quat& q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
__device__ __forceinline__ Quat quat_normalized(const Quat& q) {
  float n = sqrtf(quat_n2(q));
  return Quat{q.w / n, q.x / n, q.y / n, q.z / n};
}
__device__ __forceinline__ Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
__device__ __forceinline__ Quat quat_inverse(const Quat& q) {
  float n2 = quat_n2(q);
  if (n2 > 0.0f) return Quat{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return Quat{0.f, 0.f, 0.f, 0.f};
}
__device__ __forceinline__ void quat_to_matrix(const Quat& q, float (&m)[3][3]) {
  float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0][0] = 1.0f - (tyy + tzz); m[0][1] = txy - twz; m[0][2] = txz + twy;
  m[1][0] = txy + twz; m[1][1] = 1.0f - (txx + tzz); m[1][2] = tyz - twx;
  m[2][0] = txz - twy; m[2][1] = tyz + twx; m[2][2] = 1.0f - (txx + tyy);
}

// ---------------------------------------------------------------------------
// SH rotation (reference helper.cpp:938-1075 / cudakdtree.cu:11-199): the
// Ivanic-Ruedenberg recurrence, fully unrolled at compile time.  Coefficient
// tables (u, v, w per (m,n)) live in constant memory, filled by the host with
// std::sqrt so they equal the reference's `sqrt(a/b)` doubles bit for bit.
// ---------------------------------------------------------------------------
struct ShCoef { double u2[25], v2[25], w2[25], u3[49], v3[49], w3[49]; };
static __constant__ ShCoef c_sh;  // per translation unit (no -rdc); filled by the TU that launches SH kernels

template <int L, int I, int A, int B, typename Prev>
__device__ __forceinline__ float shP(const float (&r1)[3][3], const Prev& prev) {
  constexpr int o = L - 1;
  const float ri1 = r1[I + 1][2], rim1 = r1[I + 1][0], ri0 = r1[I + 1][1];
  if constexpr (B == L) return ri1 * prev[A + o][L - 1 + o] - rim1 * prev[A + o][-L + 1 + o];
  else if constexpr (B == -L) return ri1 * prev[A + o][-L + 1 + o] + rim1 * prev[A + o][L - 1 + o];
  else return ri0 * prev[A + o][B + o];
}

template <int L, int M, int N, typename Prev>
__device__ __forceinline__ float sh_entry(const float (&r1)[3][3], const Prev& prev) {
  constexpr int AM = M < 0 ? -M : M;
  constexpr int idx = (M + L) * (2 * L + 1) + (N + L);
  const double cu = (L == 2) ? c_sh.u2[idx] : c_sh.u3[idx];
  const double cv = (L == 2) ? c_sh.v2[idx] : c_sh.v3[idx];
  const double cw = (L == 2) ? c_sh.w2[idx] : c_sh.w3[idx];
  double acc = 0.0;
  if constexpr (AM != L) acc += cu * (double)shP<L, 0, M, N>(r1, prev);
  {
    double V;
    if constexpr (M == 0) V = (double)(shP<L, 1, 1, N>(r1, prev) + shP<L, -1, -1, N>(r1, prev));
    else if constexpr (M == 1) V = 1.4142135623730951 * (double)shP<L, 1, 0, N>(r1, prev);
    else if constexpr (M > 1) V = (double)(shP<L, 1, M - 1, N>(r1, prev) - shP<L, -1, -M + 1, N>(r1, prev));
    else if constexpr (M == -1) V = 1.4142135623730951 * (double)shP<L, -1, 0, N>(r1, prev);
    else V = (double)(shP<L, 1, M + 1, N>(r1, prev) + shP<L, -1, -M - 1, N>(r1, prev));
    acc += cv * V;
  }
  if constexpr (M != 0 && AM < L - 1) {
    double W;
    if constexpr (M > 0) W = (double)(shP<L, 1, M + 1, N>(r1, prev) + shP<L, -1, -M - 1, N>(r1, prev));
    else W = (double)(shP<L, 1, M - 1, N>(r1, prev) - shP<L, -1, -M + 1, N>(r1, prev));
    acc += cw * W;
  }
  return (float)acc;
}

// Rotate one Gaussian's / sample's 16x3 interleaved SH block in place, with the
// reference's odd-index sign flips (GaussianView.cpp:3138-3154, cudakdtree.cu:160-196).
// `sh` may point to shared or global memory; element i of channel c is sh[(i*3+c)*stride].
__device__ __forceinline__ void sh_rotate_flipped(const float (&R)[3][3], float* sh) {
  float b1[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) b1[i][j] = R[(i + 1) % 3][(j + 1) % 3];

  // band 1 (coefficients 1..3; flips: index 1 and 3 are odd)
  {
    float in[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { float v = sh[(1 + i) * 3 + c]; in[i][c] = ((1 + i) & 1) ? -v : v; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k < 3; k++) a += b1[i][k] * in[k][c];
        sh[(1 + i) * 3 + c] = ((1 + i) & 1) ? -a : a;
      }
  }
  float b2[5][5];
  static_for<5>([&](auto mi) {
    static_for<5>([&](auto ni) {
      constexpr int m = decltype(mi)::value, n = decltype(ni)::value;
      b2[m][n] = sh_entry<2, m - 2, n - 2>(b1, b1);
    });
  });
  {
    float in[5][3];
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { float v = sh[(4 + i) * 3 + c]; in[i][c] = ((4 + i) & 1) ? -v : v; }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k < 5; k++) a += b2[i][k] * in[k][c];
        sh[(4 + i) * 3 + c] = ((4 + i) & 1) ? -a : a;
      }
  }
  {
    float in[7][3];
#pragma unroll
    for (int i = 0; i < 7; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { float v = sh[(9 + i) * 3 + c]; in[i][c] = ((9 + i) & 1) ? -v : v; }
    // band-3 rows are produced one at a time and consumed immediately
    static_for<7>([&](auto mi) {
      constexpr int m = decltype(mi)::value;
      float row[7];
      static_for<7>([&](auto ni) {
        constexpr int n = decltype(ni)::value;
        row[n] = sh_entry<3, m - 3, n - 3>(b1, b2);
      });
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float a = 0.0f;
#pragma unroll
        for (int k = 0; k < 7; k++) a += row[k] * in[k][c];
        sh[(9 + m) * 3 + c] = ((9 + m) & 1) ? -a : a;
      }
    });
  }
}

}  // namespace arapgs
