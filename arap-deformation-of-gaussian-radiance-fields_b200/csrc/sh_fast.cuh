// SH rotation, float-only production version.
//
// The reference builds the degree-2/3 rotation matrices with *double* coefficient products
// (`kSqrt03_04 * (float expr)`, helper.cpp:30-45, 990-1075) only because its constants are
// double macros; device_math.cuh reproduces that rounding bit for bit (used by the parity test
// kernel).  The per-step kernels use this float version instead: coefficients rounded to float,
// fused multiply-adds allowed.  Matrix entries differ from the reference's by <= 2 float ulp, i.e.
// SH coefficients by ~1e-7 relative per step — three orders below the 1e-5 parity tolerance and
// invisible in a render (> 120 dB) — and it removes ~500 FP64/conversion instructions per item,
// which is what made the sample-SH pass compute-bound.
#pragma once
#include "device_math.cuh"

namespace arapgs {

struct ShCoefF { float u2[25], v2[25], w2[25], u3[49], v3[49], w3[49]; };  // v already times sqrt(2) where |m| == 1
static __constant__ ShCoefF c_shf;

template <int L, int I, int A, int B, typename Prev>
__device__ __forceinline__ float shPf(const float (&r1)[3][3], const Prev& prev) {
  constexpr int o = L - 1;
  const float ri1 = r1[I + 1][2], rim1 = r1[I + 1][0], ri0 = r1[I + 1][1];
  if constexpr (B == L) return fmaf(ri1, prev[A + o][L - 1 + o], -(rim1 * prev[A + o][-L + 1 + o]));
  else if constexpr (B == -L) return fmaf(ri1, prev[A + o][-L + 1 + o], rim1 * prev[A + o][L - 1 + o]);
  else return ri0 * prev[A + o][B + o];
}

template <int L, int M, int N, typename Prev>
__device__ __forceinline__ float sh_entry_f(const float (&r1)[3][3], const Prev& prev) {
  constexpr int AM = M < 0 ? -M : M;
  constexpr int idx = (M + L) * (2 * L + 1) + (N + L);
  const float cu = (L == 2) ? c_shf.u2[idx] : c_shf.u3[idx];
  const float cv = (L == 2) ? c_shf.v2[idx] : c_shf.v3[idx];
  const float cw = (L == 2) ? c_shf.w2[idx] : c_shf.w3[idx];
  float V;
  if constexpr (M == 0) V = shPf<L, 1, 1, N>(r1, prev) + shPf<L, -1, -1, N>(r1, prev);
  else if constexpr (M == 1) V = shPf<L, 1, 0, N>(r1, prev);
  else if constexpr (M > 1) V = shPf<L, 1, M - 1, N>(r1, prev) - shPf<L, -1, -M + 1, N>(r1, prev);
  else if constexpr (M == -1) V = shPf<L, -1, 0, N>(r1, prev);
  else V = shPf<L, 1, M + 1, N>(r1, prev) + shPf<L, -1, -M - 1, N>(r1, prev);
  float acc = cv * V;
  if constexpr (AM != L) acc = fmaf(cu, shPf<L, 0, M, N>(r1, prev), acc);
  if constexpr (M != 0 && AM < L - 1) {
    float W;
    if constexpr (M > 0) W = shPf<L, 1, M + 1, N>(r1, prev) + shPf<L, -1, -M - 1, N>(r1, prev);
    else W = shPf<L, 1, M - 1, N>(r1, prev) - shPf<L, -1, -M + 1, N>(r1, prev);
    acc = fmaf(cw, W, acc);
  }
  return acc;
}

// In-place rotation of a 16 x 3 interleaved SH block, with the reference's odd-index sign flips
// (GaussianView.cpp:3138-3154, cudakdtree.cu:160-196).
__device__ __forceinline__ void sh_rotate_flipped_fast(const float (&R)[3][3], float* sh) {
  float b1[3][3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) b1[i][j] = R[(i + 1) % 3][(j + 1) % 3];
  {
    float in[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { const float v = sh[(1 + i) * 3 + c]; in[i][c] = ((1 + i) & 1) ? -v : v; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float a = fmaf(b1[i][2], in[2][c], fmaf(b1[i][1], in[1][c], b1[i][0] * in[0][c]));
        sh[(1 + i) * 3 + c] = ((1 + i) & 1) ? -a : a;
      }
  }
  float b2[5][5];
  static_for<5>([&](auto mi) {
    static_for<5>([&](auto ni) {
      constexpr int m = decltype(mi)::value, n = decltype(ni)::value;
      b2[m][n] = sh_entry_f<2, m - 2, n - 2>(b1, b1);
    });
  });
  {
    float in[5][3];
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { const float v = sh[(4 + i) * 3 + c]; in[i][c] = ((4 + i) & 1) ? -v : v; }
#pragma unroll
    for (int i = 0; i < 5; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float a = b2[i][0] * in[0][c];
#pragma unroll
        for (int k = 1; k < 5; k++) a = fmaf(b2[i][k], in[k][c], a);
        sh[(4 + i) * 3 + c] = ((4 + i) & 1) ? -a : a;
      }
  }
  {
    float in[7][3];
#pragma unroll
    for (int i = 0; i < 7; i++)
#pragma unroll
      for (int c = 0; c < 3; c++) { const float v = sh[(9 + i) * 3 + c]; in[i][c] = ((9 + i) & 1) ? -v : v; }
    static_for<7>([&](auto mi) {
      constexpr int m = decltype(mi)::value;
      float row[7];
      static_for<7>([&](auto ni) {
        constexpr int n = decltype(ni)::value;
        row[n] = sh_entry_f<3, m - 3, n - 3>(b1, b2);
      });
#pragma unroll
      for (int c = 0; c < 3; c++) {
        float a = row[0] * in[0][c];
#pragma unroll
        for (int k = 1; k < 7; k++) a = fmaf(row[k], in[k][c], a);
        sh[(9 + m) * 3 + c] = ((9 + m) & 1) ? -a : a;
      }
    });
  }
}

}  // namespace arapgs
