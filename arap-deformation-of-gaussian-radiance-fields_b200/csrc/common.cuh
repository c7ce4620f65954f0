// Shared device/host helpers for libarapgs (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <utility>

namespace arapgs {

constexpr int KNN_MAX = 12;       // reference helper.hpp:48
constexpr int SAMPLES_PER_CELL = 64;  // SAMPLES_PER_GRID^3, helper.hpp:56
constexpr int SH_FLOATS = 48;     // 16 coeffs x RGB, helper.hpp:81-85

void set_error(const std::string& msg);

#define ARAP_CUDA_TRY(expr)                                                             \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      ::arapgs::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                          __FILE__ + ":" + std::to_string(__LINE__));                   \
      return ARAP_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define ARAP_KERNEL_CHECK() ARAP_CUDA_TRY(cudaGetLastError())

// compile-time loop: f(std::integral_constant<int,0>{}) ... f(<N-1>)
template <int... Is, class F>
__host__ __device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F>
__host__ __device__ __forceinline__ void static_for(F&& f) {
  static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F&&>(f));
}

// Per-node transform record consumed by the LBS kernels (built once per step
// from the solver's x and the current node positions).  112 B, 16-B aligned:
// seven LDG.128 per (point, neighbour) pair.  The pitch is deliberately NOT a power
// of two: with a 128-B pitch every lane of a gather hits the same L1 data banks
// (measured: 2x slower LBS); 112 B spreads neighbouring node ids across banks.
//   A : column-major 3x3 (reference DeformGraph::rot layout, Deform.hpp:29-36)
//   c : trans + (double)pos        g : node position (float, for the float
//                                      subtraction `cur - position`, Deform.hpp:240)
struct __align__(16) NodeXf {
  double A[9];
  double c[3];
  float g[3];
  float pad;
};
static_assert(sizeof(NodeXf) == 112, "NodeXf layout");

// Float record of the tolerance-mode skinning kernels (arap_params.lbs_mode = 3): A - I row-major, t, g.  64 B.
struct __align__(16) NodeXf32 {
  float dA[9];
  float t[3];
  float g[3];
  float pad;
};
static_assert(sizeof(NodeXf32) == 64, "NodeXf32 layout");

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming 128-bit load/store that bypass L1 allocation (data touched once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace arapgs

// status codes (mirrored in include/arapgs.h)
#ifndef ARAP_OK
#define ARAP_OK 0
#define ARAP_ERR_INVALID 1
#define ARAP_ERR_CUDA 2
#define ARAP_ERR_STATE 3
#define ARAP_ERR_IO 4
#define ARAP_ERR_KNN_TIES 5
#define ARAP_ERR_UNSUPPORTED 6
#define ARAP_ERR_NUMERIC 7
#endif
