// Stage (a): density grid — cell assignment, Gaussian re-ordering, per-Gaussian
// cutoff boxes, per-cell Gaussian lists, sample emit, adaptive LPF, evaluation.
//
//   cell_assign       GetGsGrid + CUB scan + host re-order   (cudakdtree.cu:403-423, 508-528; GaussianView.cpp:3896-3953)
//   gs_aabbs          per-Gaussian cutoff AABB               (GaussianView.cpp:3961-4019)
//   footprint_count   GetBoxesGsGrid + scan                  (cudakdtree.cu:425-451, 533-553)
//   footprint_fill    serial host fill, replaced             (GaussianView.cpp:4077-4100)
//   valid_cells / emit_samples                               (GaussianView.cpp:4040-4053, 4111-4133)
//   ada_lpf           GetAdaLpfRatio                         (GaussianView.cpp:4670-4751)
//   grid_eval         Rasterizer::forward3d_grid (EXTERNAL, source absent; call site GaussianView.cpp:4159-4186)
//
// Cell indices and list contents are integer work and must equal the
// reference's bit for bit: cells use the same float expression
// floor((p - min) / step) (IEEE division, -fmad=false), and the lists are
// produced by an atomic fill followed by an in-place per-cell sort, which gives
// the serial host loop's order (ascending Gaussian index within a cell).
#include <cfloat>
#include "device_math.cuh"
#include "kernels.h"

namespace arapgs {

// ------------------------------------------------------------------ scan (int32, inclusive)
constexpr int SCAN_BLOCK = 1024;
constexpr int SCAN_ITEMS = 4;  // per thread -> 4096 per block

__device__ __forceinline__ int block_scan_incl(int v, int* s_warp) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
  if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    const int wv = s_warp[threadIdx.x];
    int y = wv;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, y, o); if (threadIdx.x >= o) y += t; }
    s_warp[threadIdx.x] = y - wv;
  }
  __syncthreads();
  const int r = x + s_warp[threadIdx.x >> 5];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block_sums(const int* __restrict__ in, long long n, int* __restrict__ sums) {
  __shared__ int s_warp[32];
  const long long base = (long long)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (long long)threadIdx.x * SCAN_ITEMS;
  int t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) if (base + i < n) t += in[base + i];
  const int inc = block_scan_incl(t, s_warp);
  if (threadIdx.x == SCAN_BLOCK - 1) sums[blockIdx.x] = inc;
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int* __restrict__ sums, int nb) {  // exclusive, single block
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int b = 0; b < nb; b += SCAN_BLOCK) {
    const int i = b + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    const int inc = block_scan_incl(v, s_warp);
    const int carry = s_carry;
    if (i < nb) sums[i] = carry + inc - v;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) s_carry = carry + inc;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_final(const int* __restrict__ in, long long n, const int* __restrict__ sums,
                                                           int* __restrict__ out) {
  __shared__ int s_warp[32];
  const long long base = (long long)blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + (long long)threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS]; int t = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { v[i] = (base + i < n) ? in[base + i] : 0; t += v[i]; }
  const int inc = block_scan_incl(t, s_warp);
  int run = sums[blockIdx.x] + inc - t;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; i++) { run += v[i]; if (base + i < n) out[base + i] = run; }
}

static int scan_inclusive(const int* in, int* out, long long n, int* sums_scratch, cudaStream_t st) {
  if (n <= 0) return ARAP_OK;
  const int per = SCAN_BLOCK * SCAN_ITEMS;
  const int nb = (int)((n + per - 1) / per);
  k_scan_block_sums<<<nb, SCAN_BLOCK, 0, st>>>(in, n, sums_scratch);
  k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(sums_scratch, nb);
  k_scan_final<<<nb, SCAN_BLOCK, 0, st>>>(in, n, sums_scratch, out);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

}  // namespace arapgs
// exported for the table builders of apply.cu; sums_scratch >= ceil(n / 4096) ints
extern "C" int arapk_scan_inclusive_i32(const int* in, int* out, long long n, int* sums_scratch, cudaStream_t st) {
  return arapgs::scan_inclusive(in, out, n, sums_scratch, st);
}
namespace arapgs {

// ------------------------------------------------------------------ segmented in-place sort (ascending int)
// Normalised bitonic network (all comparators point up) with virtual +inf
// padding, so arbitrary segment lengths sort in place.
constexpr int SEG_WARP_CAP = 1024;

__device__ __forceinline__ void bitonic_steps(int* a, int n, int tid, int nthreads, bool block_sync) {
  int np = 1; while (np < n) np <<= 1;
  for (int k = 2; k <= np; k <<= 1) {
    for (int i = tid; i < np; i += nthreads) {  // first sub-step: mirror within the k-block
      const int p = i ^ (k - 1);
      if (p > i && p < n) { const int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
    }
    if (block_sync) __syncthreads(); else __syncwarp();
    for (int j = k >> 2; j > 0; j >>= 1) {
      for (int i = tid; i < np; i += nthreads) {
        const int p = i ^ j;
        if (p > i && p < n) { const int x = a[i], y = a[p]; if (x > y) { a[i] = y; a[p] = x; } }
      }
      if (block_sync) __syncthreads(); else __syncwarp();
    }
  }
}

// one warp per cell; segments longer than SEG_WARP_CAP go to the block kernel
constexpr int SEG_MED_CAP = 8192;      // "medium" cells: 32 KB of shared memory per CTA, seven CTAs per SM
__global__ void __launch_bounds__(256)
k_segsort_warp(long long ncell, const int* __restrict__ prefix, int* __restrict__ data, int* __restrict__ big_count,
               int* __restrict__ big_list) {   // big_count[0] / big_list ascending: medium cells; big_count[1] / big_list descending from its end: large
  __shared__ int s_buf[8][SEG_WARP_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long c = (long long)blockIdx.x * 8 + warp;
  if (c >= ncell) return;
  const int b = c ? prefix[c - 1] : 0, e = prefix[c];
  const int n = e - b;
  if (n <= 1) return;
  if (n > SEG_WARP_CAP) {
    if (lane == 0) {
      if (n <= SEG_MED_CAP) big_list[atomicAdd(big_count, 1)] = (int)c;
      else big_list[ncell - 1 - atomicAdd(big_count + 1, 1)] = (int)c;
    }
    return;
  }
  int* a = s_buf[warp];
  for (int i = lane; i < n; i += 32) a[i] = data[b + i];
  __syncwarp();
  if (n <= 32) {
    // rank sort: values are distinct Gaussian indices
    const int v = lane < n ? a[lane] : 0x7fffffff;
    int r = 0;
    for (int i = 0; i < n; i++) r += (a[i] < v);
    if (lane < n) data[b + r] = v;
    return;
  }
  bitonic_steps(a, n, lane, 32, false);
  for (int i = lane; i < n; i += 32) data[b + i] = a[i];
}

// cells longer than SEG_WARP_CAP: one CTA per cell, sorted in shared memory when the list fits (SEG_BLOCK_CAP entries),
// else in place in global memory
constexpr int SEG_BLOCK_CAP = 49152;   // 192 KB of dynamic shared memory
__global__ void __launch_bounds__(1024)
k_segsort_block(const int* __restrict__ big_list, int list_stride, const int* __restrict__ prefix, int* __restrict__ data) {
  extern __shared__ int s_seg[];
  const int c = big_list[(long long)blockIdx.x * list_stride];
  const int b = c ? prefix[c - 1] : 0, e = prefix[c];
  const int n = e - b;
  if (n > SEG_BLOCK_CAP) { bitonic_steps(data + b, n, threadIdx.x, blockDim.x, true); return; }   // only reachable in the "large" launch
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_seg[i] = data[b + i];
  __syncthreads();
  bitonic_steps(s_seg, n, threadIdx.x, blockDim.x, true);
  for (int i = threadIdx.x; i < n; i += blockDim.x) data[b + i] = s_seg[i];
}

// ------------------------------------------------------------------ cell assignment
__device__ __forceinline__ int cell_coord(float p, float mn, float step) { return (int)floorf((p - mn) / step); }

__global__ void k_cell_hist(long long N, const float* __restrict__ pos, float3 mn, float step, int G, int* __restrict__ cell_out,
                            int* __restrict__ cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int x = cell_coord(pos[3 * i], mn.x, step), y = cell_coord(pos[3 * i + 1], mn.y, step), z = cell_coord(pos[3 * i + 2], mn.z, step);
  const int c = x * G * G + y * G + z;
  cell_out[i] = c;
  if (c >= 0 && c < G * G * G) atomicAdd(&cnt[c], 1);
}
__global__ void k_cell_fill(long long N, int G, const int* __restrict__ cell, const int* __restrict__ prefix, int* __restrict__ fill,
                            int* __restrict__ members) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int c = cell[i];
  if (c < 0 || c >= G * G * G) return;  // outside the scene box (the reference would index out of bounds)
  const int start = c ? prefix[c - 1] : 0;
  members[start + atomicAdd(&fill[c], 1)] = (int)i;
}
__global__ void k_invert_perm(long long N, const int* __restrict__ members, int* __restrict__ new_idx) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= N) return;
  new_idx[members[t]] = (int)t;
}

// out[new_idx[i]] = in[i] for all five attribute arrays; one thread per (gaussian, 16-byte chunk of SH)
__global__ void k_permute(long long N, const int* __restrict__ new_idx, const float* __restrict__ pos, const float* __restrict__ rot,
                          const float* __restrict__ scale, const float* __restrict__ opacity, const float* __restrict__ shs,
                          float* __restrict__ pos_o, float* __restrict__ rot_o, float* __restrict__ scale_o,
                          float* __restrict__ opacity_o, float* __restrict__ shs_o) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t / 12; const int c = (int)(t - i * 12);
  if (i >= N) return;
  const long long d = new_idx[i];
  reinterpret_cast<float4*>(shs_o + d * 48)[c] = __ldg(reinterpret_cast<const float4*>(shs + i * 48) + c);
  if (c == 0) {
    pos_o[3 * d] = pos[3 * i]; pos_o[3 * d + 1] = pos[3 * i + 1]; pos_o[3 * d + 2] = pos[3 * i + 2];
    scale_o[3 * d] = scale[3 * i]; scale_o[3 * d + 1] = scale[3 * i + 1]; scale_o[3 * d + 2] = scale[3 * i + 2];
    reinterpret_cast<float4*>(rot_o)[d] = __ldg(reinterpret_cast<const float4*>(rot) + i);
    opacity_o[d] = opacity[i];
  }
}

// ------------------------------------------------------------------ Gaussian cutoff boxes
__global__ void k_gs_aabbs(long long N, const float* __restrict__ pos, const float* __restrict__ rot, const float* __restrict__ scale,
                           const float* __restrict__ opacity, float* __restrict__ aabb, float* __restrict__ clip,
                           float* __restrict__ smax) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const float4 r4 = ldg4(rot + 4 * g);
  const Quat q = quat_normalized(Quat{r4.x, r4.y, r4.z, r4.w});
  float R[3][3]; quat_to_matrix(q, R);
  const float op = opacity[g];
  float s3[3];
  if (op <= 1.0f / 255.0f) { s3[0] = s3[1] = s3[2] = 0.0f; }
  else {
    const float rad = sqrtf(-2.0f * logf(1.0f / 255.0f / op));  // CUTOFF_ALPHA (helper.hpp:62)
#pragma unroll
    for (int i = 0; i < 3; i++) s3[i] = rad * (scale[3 * g + i] + 0.0f);
  }
  if (clip) { clip[3 * g] = s3[0]; clip[3 * g + 1] = s3[1]; clip[3 * g + 2] = s3[2]; }
  if (smax) { float m = s3[0]; if (s3[1] > m) m = s3[1]; if (s3[2] > m) m = s3[2]; smax[g] = m; }
  const float p[3] = {pos[3 * g], pos[3 * g + 1], pos[3 * g + 2]};
  float mn[3] = {p[0], p[1], p[2]}, mx[3] = {p[0], p[1], p[2]};
#pragma unroll
  for (int d = 0; d < 3; d++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float v = R[c][d] * s3[d];
      const float l = p[c] + v, r = p[c] + (-v);
      mn[c] = fminf(mn[c], l); mx[c] = fmaxf(mx[c], l);
      mn[c] = fminf(mn[c], r); mx[c] = fmaxf(mx[c], r);
    }
#pragma unroll
  for (int c = 0; c < 3; c++) { aabb[6 * g + c] = mn[c]; aabb[6 * g + 3 + c] = mx[c]; }
}

__device__ __forceinline__ void cell_range(const float* a, float3 mn, float step, int G, int padding, int (&lo)[3], int (&hi)[3]) {
  const float m[3] = {mn.x, mn.y, mn.z};
#pragma unroll
  for (int c = 0; c < 3; c++) {
    lo[c] = max((int)floorf((a[c] - m[c]) / step) - padding, 0);
    hi[c] = min((int)floorf((a[3 + c] - m[c]) / step) + padding, G - 1);
  }
}

// [xlo, xhi): the x-slab of cells this launch bins into (the whole grid: 0, G) — multi-GPU grid sharding, SURVEY 8(e)
template <bool FILL>
__global__ void k_footprint(long long N, const float* __restrict__ aabb, float3 mn, float step, int G, int padding,
                            int* __restrict__ cnt_or_fill, const int* __restrict__ prefix, int* __restrict__ lists, int xlo, int xhi) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  float a[6];
#pragma unroll
  for (int c = 0; c < 6; c++) a[c] = aabb[6 * g + c];
  int lo[3], hi[3]; cell_range(a, mn, step, G, padding, lo, hi);
  lo[0] = max(lo[0], xlo); hi[0] = min(hi[0], xhi - 1);
  for (int x = lo[0]; x <= hi[0]; x++) for (int y = lo[1]; y <= hi[1]; y++) for (int z = lo[2]; z <= hi[2]; z++) {
    const int c = x * G * G + y * G + z;
    if (FILL) {
      const int start = c ? prefix[c - 1] : 0;
      lists[start + atomicAdd(&cnt_or_fill[c], 1)] = (int)g;
    } else atomicAdd(&cnt_or_fill[c], 1);
  }
}

// Gaussians per x-layer of cells (balanced slab cuts for the multi-GPU grid): same cell expression as k_cell_hist
__global__ void k_xlayer_hist(long long N, const float* __restrict__ pos, float mnx, float step, int G, int* __restrict__ hist) {
  __shared__ int s_h[256];
  for (int t = threadIdx.x; t < 256; t += blockDim.x) s_h[t] = 0;
  __syncthreads();
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < N; g += (long long)gridDim.x * blockDim.x) {
    int x = (int)floorf((pos[3 * g] - mnx) / step);
    x = min(max(x, 0), G - 1);
    atomicAdd(&s_h[x], 1);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < G; t += blockDim.x) if (s_h[t]) atomicAdd(&hist[t], s_h[t]);
}

// total of the per-cell pair counts in 64 bits: the prefix sums and list offsets are int32
__global__ void k_count_total(long long ncell, const int* __restrict__ cnt, unsigned long long* __restrict__ total) {
  unsigned long long a = 0;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (long long)gridDim.x * blockDim.x) a += (unsigned)cnt[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0 && a) atomicAdd(total, a);
}

// ------------------------------------------------------------------ valid cells + samples
__global__ void k_valid_flags(long long ncell, const int* __restrict__ prefix, int* __restrict__ flags) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  flags[c] = (prefix[c] - (c ? prefix[c - 1] : 0)) != 0;
}
__global__ void k_valid_compact(long long ncell, const int* __restrict__ flags, const int* __restrict__ incl, int* __restrict__ out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  if (flags[c]) out[incl[c] - 1] = (int)c;
}

__global__ void k_emit_samples(int V, const int* __restrict__ valid, float3 mn, float step, int G, float* __restrict__ out) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = t >> 6; const int s = (int)(t & 63);
  if (i >= V) return;
  const int c = valid[i];
  const int xi = c / (G * G), yi = (c - xi * G * G) / G, zi = c % G;
  const float interval = step / 4;  // SAMPLES_PER_GRID
  // float + int*float, then + 0.5*interval evaluated in double, rounded to float (GaussianView.cpp:4119-4121)
  const float gx = (float)((double)(mn.x + xi * step) + 0.5 * (double)interval);
  const float gy = (float)((double)(mn.y + yi * step) + 0.5 * (double)interval);
  const float gz = (float)((double)(mn.z + zi * step) + 0.5 * (double)interval);
  const int sx = s / 16, sy = (s - sx * 16) / 4, sz = s % 4;
  float* o = out + t * 3;
  o[0] = gx + sx * interval; o[1] = gy + sy * interval; o[2] = gz + sz * interval;
}

// adaptive LPF (closed form of the 24x9 least squares: M = P Q^T (Q Q^T)^-1,
// (Q Q^T)^-1 = 0.5 I - 0.125 ones)
__global__ void k_ada_lpf(int V, const float* __restrict__ samples, const int* __restrict__ valid, float lpf, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float* base = samples + (size_t)i * 64 * 3;
  float PQt[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
#pragma unroll
  for (int col = 0; col < 8; col++) {
    int idx = 0; float q[3] = {0.f, 0.f, 0.f};
    if (col % 2 == 1) { q[0] = 1.0f; idx += 48; }
    if ((col / 2) % 2 == 1) { q[1] = 1.0f; idx += 12; }
    if ((col / 4) % 2 == 1) { q[2] = 1.0f; idx += 3; }
    float p[3];
#pragma unroll
    for (int r = 0; r < 3; r++) p[r] = (base[idx * 3 + r] - base[r]) / 3.0f;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) PQt[r][c] += p[r] * q[c];
  }
  float Mx[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const float rs = PQt[r][0] + PQt[r][1] + PQt[r][2];
#pragma unroll
    for (int c = 0; c < 3; c++) Mx[r][c] = 0.5f * PQt[r][c] - 0.125f * rs;
  }
  float* o = out + (size_t)valid[i] * 9;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 3; t++) s += Mx[r][t] * Mx[c][t];
      o[3 * r + c] = s * lpf;
    }
}

// ------------------------------------------------------------------ empty cells
// JudgeEmptyGrid (GaussianView.cpp:4272-4318).  Pass 1: a valid cell whose 64 aim opacities (summed in sample order, float)
// exceed 1e-6 clears the "bad" mark of itself and of its six face neighbours (clamped at the grid border, like the
// reference).  Pass 2: empty_grid[i] = bad[valid[i]].  The marks only ever go 1 -> 0, so the plain stores race benignly.
__global__ void k_judge_mark(int V, const int* __restrict__ valid, const float* __restrict__ aim_opacity, int G,
                             uint8_t* __restrict__ bad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  const float* o = aim_opacity + (size_t)i * 64;
  float total = 0.0f;
  for (int j = 0; j < 64; j++) total += o[j];
  if (!(total > 1e-6)) return;   // float compared with the double literal, as in the reference
  const int idx = valid[i];
  const int z = idx % G, y = (idx / G) % G, x = idx / (G * G);
  bad[idx] = 0;
  bad[max(x - 1, 0) * (G * G) + y * G + z] = 0;
  bad[min(x + 1, G - 1) * (G * G) + y * G + z] = 0;
  bad[x * (G * G) + max(y - 1, 0) * G + z] = 0;
  bad[x * (G * G) + min(y + 1, G - 1) * G + z] = 0;
  bad[x * (G * G) + y * G + max(z - 1, 0)] = 0;
  bad[x * (G * G) + y * G + min(z + 1, G - 1)] = 0;
}
__global__ void k_judge_collect(int V, const int* __restrict__ valid, const uint8_t* __restrict__ bad, int* __restrict__ empty) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= V) return;
  empty[i] = bad[valid[i]] ? 1 : 0;
}

// ------------------------------------------------------------------ field evaluation
// forward3d_grid is EXTERNAL to the reference tree (XinhaoT/CudaRasterizer @
// 96ea96c, source absent) — PARITY UNPINNED.  Implemented from the call-site
// parameter list and the field definition:
//   w_g(x) = alpha_g exp(-1/2 d^T (Sigma_g + LPF_cell)^-1 d),  feature = sum w_g SH_g,  opacity = sum w_g
// One CTA per valid cell, one thread per sample; the cell's list is staged in
// shared memory in chunks (inverse covariance, centre, alpha, cutoff, 48 SH
// floats per Gaussian) and consumed in list order.
constexpr int EV_CHUNK = 64;   // list entries examined per pass: one per thread
struct EvGauss { float px, py, pz, a, i00, i01, i02, i11, i12, i22, cut, pad; };

// One CTA per valid cell, one thread per sample.  A cell's list holds every Gaussian whose cut-off box, grown by `padding`
// cells, meets the cell (GV:4030-4100) — with padding 1 at least 27 cells per Gaussian, most of which it cannot reach.  Each
// pass therefore first decides PER (cell, Gaussian) pair, one thread per list entry, whether the LPF-widened Gaussian can
// reach any of the cell's 64 samples at all: the set where the evaluator's own cut-off test  -1/2 d^T (Sigma + LPF)^-1 d >= ln(1/255 / alpha)
// holds has the axis-aligned half extent sqrt(r^2 (Sigma + LPF)_aa), r^2 = -2 ln(1/255 / alpha), and is tested against the
// bounding box of the cell's sample positions (measured, so deformed samples are handled too).  Survivors are compacted in
// list order into shared memory and only THEIR 192-byte SH rows are fetched; the per-sample loop then runs over the
// survivors exactly as before, so the sums (and their order) are unchanged.
__global__ void __launch_bounds__(64)
k_grid_eval(const int* __restrict__ valid, const int* __restrict__ prefix, const int* __restrict__ lists,
            const float* __restrict__ samples, const float* __restrict__ pos, const float* __restrict__ rot,
            const float* __restrict__ scale, const float* __restrict__ opacity, const float* __restrict__ shs,
            const float* __restrict__ ada_lpf, float* __restrict__ out_feature, float* __restrict__ out_opacity) {
  __shared__ EvGauss s_g[EV_CHUNK];
  __shared__ int s_gi[EV_CHUNK];
  __shared__ float s_sh[EV_CHUNK][48];
  __shared__ float s_box[2][6];
  __shared__ int s_wcnt[2];
  const int i = blockIdx.x;
  const int c = valid[i];
  const int beg = c ? prefix[c - 1] : 0, end = prefix[c];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* lp = ada_lpf + (size_t)c * 9;
  const float* xs = samples + ((size_t)i * 64 + tid) * 3;
  const float x0 = xs[0], x1 = xs[1], x2 = xs[2];
  // bounding box of the cell's samples
  float bmn[3] = {x0, x1, x2}, bmx[3] = {x0, x1, x2};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int a = 0; a < 3; a++) { bmn[a] = fminf(bmn[a], __shfl_xor_sync(0xffffffffu, bmn[a], o)); bmx[a] = fmaxf(bmx[a], __shfl_xor_sync(0xffffffffu, bmx[a], o)); }
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 3; a++) { s_box[warp][a] = bmn[a]; s_box[warp][3 + a] = bmx[a]; }
  __syncthreads();
  float bc[3], bh[3];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float mn = fminf(s_box[0][a], s_box[1][a]), mx = fmaxf(s_box[0][3 + a], s_box[1][3 + a]);
    bc[a] = 0.5f * (mn + mx); bh[a] = 0.5f * (mx - mn);
  }
  float acc[48];
#pragma unroll
  for (int u = 0; u < 48; u++) acc[u] = 0.f;
  float opa = 0.f;
  for (int base = beg; base < end; base += EV_CHUNK) {
    __syncthreads();
    // ---- one list entry per thread: evaluator constants + reach test
    bool keep = false; EvGauss e; int g = 0;
    if (base + tid < end) {
      g = lists[base + tid];
      const float a = opacity[g];
      e.px = pos[3 * g]; e.py = pos[3 * g + 1]; e.pz = pos[3 * g + 2]; e.a = a; e.pad = 0.f;
      if (a > 1.0f / 255.0f) {
        const float4 r4 = ldg4(rot + 4 * g);
        const Quat q = quat_normalized(Quat{r4.x, r4.y, r4.z, r4.w});
        float R[3][3]; quat_to_matrix(q, R);
        float Sg[3][3];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            float sum = 0.f;
#pragma unroll
            for (int d = 0; d < 3; d++) sum += R[r][d] * (scale[3 * g + d] * scale[3 * g + d]) * R[cc][d];
            Sg[r][cc] = sum + lp[3 * r + cc];
          }
        const float c00 = Sg[1][1] * Sg[2][2] - Sg[1][2] * Sg[2][1];
        const float c01 = Sg[1][2] * Sg[2][0] - Sg[1][0] * Sg[2][2];
        const float c02 = Sg[1][0] * Sg[2][1] - Sg[1][1] * Sg[2][0];
        const float det = Sg[0][0] * c00 + Sg[0][1] * c01 + Sg[0][2] * c02;
        if (det > 0.0f) {
          const float id = 1.0f / det;
          e.i00 = c00 * id; e.i01 = c01 * id; e.i02 = c02 * id;
          e.i11 = (Sg[0][0] * Sg[2][2] - Sg[0][2] * Sg[2][0]) * id;
          e.i12 = (Sg[0][2] * Sg[1][0] - Sg[0][0] * Sg[1][2]) * id;
          e.i22 = (Sg[0][0] * Sg[1][1] - Sg[0][1] * Sg[1][0]) * id;
          e.cut = logf(1.0f / 255.0f / a);
          const float r2 = -2.0f * e.cut * 1.001f;    // margin for the rounded arithmetic of the per-sample test
          keep = fabsf(e.px - bc[0]) <= bh[0] + sqrtf(r2 * fmaxf(Sg[0][0], 0.f)) * 1.0001f + 1e-7f &&
                 fabsf(e.py - bc[1]) <= bh[1] + sqrtf(r2 * fmaxf(Sg[1][1], 0.f)) * 1.0001f + 1e-7f &&
                 fabsf(e.pz - bc[2]) <= bh[2] + sqrtf(r2 * fmaxf(Sg[2][2], 0.f)) * 1.0001f + 1e-7f;
        }
      }
    }
    // ---- compaction in list order
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_wcnt[warp] = __popc(m);
    __syncthreads();
    const int slot = (warp ? s_wcnt[0] : 0) + __popc(m & ((1u << lane) - 1u));
    const int nk = s_wcnt[0] + s_wcnt[1];
    if (keep) { s_g[slot] = e; s_gi[slot] = g; }
    __syncthreads();
    for (int v = tid; v < nk * 12; v += 64) {
      const int r = v / 12, c4 = v - r * 12;
      const float4 x = __ldg(reinterpret_cast<const float4*>(shs + (size_t)s_gi[r] * 48) + c4);
      float* d = &s_sh[r][c4 * 4];
      d[0] = x.x; d[1] = x.y; d[2] = x.z; d[3] = x.w;
    }
    __syncthreads();
    for (int t = 0; t < nk; t++) {
      const EvGauss ev = s_g[t];
      const float d0 = x0 - ev.px, d1 = x1 - ev.py, d2 = x2 - ev.pz;
      const float pw = -0.5f * (ev.i00 * d0 * d0 + ev.i11 * d1 * d1 + ev.i22 * d2 * d2) - (ev.i01 * d0 * d1 + ev.i02 * d0 * d2 + ev.i12 * d1 * d2);
      if (pw > 0.0f || pw < ev.cut) continue;
      const float w = ev.a * expf(pw);
      opa += w;
#pragma unroll
      for (int u = 0; u < 48; u++) acc[u] += w * s_sh[t][u];
    }
  }
  float* of = out_feature + ((size_t)i * 64 + tid) * 48;
#pragma unroll
  for (int u = 0; u < 48; u += 4) *reinterpret_cast<float4*>(of + u) = make_float4(acc[u], acc[u + 1], acc[u + 2], acc[u + 3]);
  out_opacity[(size_t)i * 64 + tid] = opa;
}

}  // namespace arapgs

using namespace arapgs;

// scratch layout helper: [cnt G^3][fill G^3][scan sums][big list G^3 + 1][members N]
extern "C" size_t arapk_grid_scratch_bytes(long long N, int G) {
  const size_t gc = (size_t)G * G * G;
  return (gc * 4 + 8192 + (size_t)N) * sizeof(int) + 1024;
}

namespace {
struct GridScratch { int *cnt, *fill, *sums, *big, *members; };
static GridScratch carve(void* scratch, int G, long long N) {
  const size_t gc = (size_t)G * G * G;
  GridScratch s; int* p = (int*)scratch;
  s.cnt = p; p += gc; s.fill = p; p += gc; s.sums = p; p += 4096; s.big = p; p += gc + 4096; s.members = p;
  (void)N; return s;
}
static int sort_segments(long long ncell, const int* prefix, int* data, int* big, cudaStream_t st) {
  // big[0], big[1]: counts of medium / large cells; big + 2: the list (ncell entries: medium from the front, large from the back)
  ARAP_CUDA_TRY(cudaMemsetAsync(big, 0, 2 * sizeof(int), st));
  k_segsort_warp<<<(unsigned)((ncell + 7) / 8), 256, 0, st>>>(ncell, prefix, data, big, big + 2);
  ARAP_KERNEL_CHECK();
  int nbig[2] = {0, 0};
  ARAP_CUDA_TRY(cudaMemcpyAsync(nbig, big, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  ARAP_CUDA_TRY(cudaStreamSynchronize(st));
  static bool attr_set = false;
  if (!attr_set) { ARAP_CUDA_TRY(cudaFuncSetAttribute(k_segsort_block, cudaFuncAttributeMaxDynamicSharedMemorySize, SEG_BLOCK_CAP * (int)sizeof(int))); attr_set = true; }
  if (nbig[0] > 0) { k_segsort_block<<<nbig[0], 1024, SEG_MED_CAP * sizeof(int), st>>>(big + 2, 1, prefix, data); ARAP_KERNEL_CHECK(); }
  if (nbig[1] > 0) { k_segsort_block<<<nbig[1], 1024, SEG_BLOCK_CAP * sizeof(int), st>>>(big + 2 + ncell - 1, -1, prefix, data); ARAP_KERNEL_CHECK(); }
  return ARAP_OK;
}
}  // namespace

extern "C" int arapk_cell_assign(const float* pos, long long N, const float* min3, float step, int G, int* cell_out,
                                 int* prefix_out, int* new_idx_out, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (scratch_bytes < arapk_grid_scratch_bytes(N, G)) { set_error("cell_assign: scratch too small"); return ARAP_ERR_INVALID; }
  const long long gc = (long long)G * G * G;
  GridScratch s = carve(scratch, G, N);
  ARAP_CUDA_TRY(cudaMemsetAsync(s.cnt, 0, sizeof(int) * gc * 2, st));  // cnt + fill
  const float3 mn = make_float3(min3[0], min3[1], min3[2]);
  if (N > 0) { k_cell_hist<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, pos, mn, step, G, cell_out, s.cnt); ARAP_KERNEL_CHECK(); }
  int rc = scan_inclusive(s.cnt, prefix_out, gc, s.sums, st); if (rc) return rc;
  if (new_idx_out && N > 0) {
    k_cell_fill<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, G, cell_out, prefix_out, s.fill, s.members); ARAP_KERNEL_CHECK();
    rc = sort_segments(gc, prefix_out, s.members, s.big, st); if (rc) return rc;
    k_invert_perm<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, s.members, new_idx_out); ARAP_KERNEL_CHECK();
  }
  return ARAP_OK;
}

extern "C" int arapk_permute_gaussians(long long N, const int* new_idx, const float* pos, const float* rot, const float* scale,
                                       const float* opacity, const float* shs, float* pos_o, float* rot_o, float* scale_o,
                                       float* opacity_o, float* shs_o, cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  const long long T = N * 12;
  k_permute<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(N, new_idx, pos, rot, scale, opacity, shs, pos_o, rot_o, scale_o, opacity_o, shs_o);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_gs_aabbs(long long N, const float* pos, const float* rot, const float* scale, const float* opacity,
                              float* aabb, float* clip, float* smax, cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  k_gs_aabbs<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, pos, rot, scale, opacity, aabb, clip, smax);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_footprint_count(long long N, const float* aabb, const float* min3, float step, int G, int padding,
                                     int* prefix_out, long long* total_host, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  return arapk_footprint_count_slab(N, aabb, min3, step, G, padding, 0, G, prefix_out, total_host, scratch, scratch_bytes, st);
}
extern "C" int arapk_footprint_count_slab(long long N, const float* aabb, const float* min3, float step, int G, int padding, int xlo, int xhi,
                                          int* prefix_out, long long* total_host, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (xlo < 0 || xhi > G || xlo > xhi) { set_error("footprint_count: bad slab"); return ARAP_ERR_INVALID; }
  if (scratch_bytes < arapk_grid_scratch_bytes(0, G)) { set_error("footprint_count: scratch too small"); return ARAP_ERR_INVALID; }
  const long long gc = (long long)G * G * G;
  GridScratch s = carve(scratch, G, N);
  ARAP_CUDA_TRY(cudaMemsetAsync(s.cnt, 0, sizeof(int) * gc, st));
  const float3 mn = make_float3(min3[0], min3[1], min3[2]);
  if (N > 0) { k_footprint<false><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, aabb, mn, step, G, padding, s.cnt, nullptr, nullptr, xlo, xhi); ARAP_KERNEL_CHECK(); }
  if (total_host) {   // 64-bit total first: more than 2^31 - 1 pairs do not fit the int32 offsets (shard the grid: arap_comm_grid_build)
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(s.big);   // 8-byte aligned scratch, rewritten by the fill pass
    ARAP_CUDA_TRY(cudaMemsetAsync(tot, 0, sizeof(unsigned long long), st));
    k_count_total<<<148 * 4, 256, 0, st>>>(gc, s.cnt, tot); ARAP_KERNEL_CHECK();
    unsigned long long h = 0;
    ARAP_CUDA_TRY(cudaMemcpyAsync(&h, tot, sizeof(h), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaStreamSynchronize(st));
    if (h > 2147483647ULL) { set_error("footprint_count: " + std::to_string(h) + " (cell, Gaussian) pairs exceed the int32 list offsets of one GPU: shard the grid over more ranks (arap_comm_grid_build)"); return ARAP_ERR_INVALID; }
    *total_host = (long long)h;
  }
  return scan_inclusive(s.cnt, prefix_out, gc, s.sums, st);
}

extern "C" int arapk_footprint_fill(long long N, const float* aabb, const float* min3, float step, int G, int padding,
                                    const int* prefix, int* lists_out, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  return arapk_footprint_fill_slab(N, aabb, min3, step, G, padding, 0, G, prefix, lists_out, scratch, scratch_bytes, st);
}
extern "C" int arapk_footprint_fill_slab(long long N, const float* aabb, const float* min3, float step, int G, int padding, int xlo, int xhi,
                                         const int* prefix, int* lists_out, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  if (xlo < 0 || xhi > G || xlo > xhi) { set_error("footprint_fill: bad slab"); return ARAP_ERR_INVALID; }
  if (scratch_bytes < arapk_grid_scratch_bytes(0, G)) { set_error("footprint_fill: scratch too small"); return ARAP_ERR_INVALID; }
  const long long gc = (long long)G * G * G;
  GridScratch s = carve(scratch, G, N);
  ARAP_CUDA_TRY(cudaMemsetAsync(s.fill, 0, sizeof(int) * gc, st));
  const float3 mn = make_float3(min3[0], min3[1], min3[2]);
  if (N > 0) { k_footprint<true><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, aabb, mn, step, G, padding, s.fill, prefix, lists_out, xlo, xhi); ARAP_KERNEL_CHECK(); }
  return sort_segments(gc, prefix, lists_out, s.big, st);
}

extern "C" int arapk_xlayer_hist(const float* pos, long long N, const float* min3, float step, int G, int* hist_dev, cudaStream_t st) {
  if (G < 1 || G > 256) { set_error("xlayer_hist: grid_num out of range"); return ARAP_ERR_INVALID; }
  ARAP_CUDA_TRY(cudaMemsetAsync(hist_dev, 0, sizeof(int) * G, st));
  if (N > 0) { k_xlayer_hist<<<(unsigned)std::min<long long>((N + 255) / 256, 148 * 8), 256, 0, st>>>(N, pos, min3[0], step, G, hist_dev); ARAP_KERNEL_CHECK(); }
  return ARAP_OK;
}

extern "C" int arapk_valid_cells(const int* prefix, int G, int* valid_out, int* count_host, void* scratch, size_t scratch_bytes,
                                 cudaStream_t st) {
  if (scratch_bytes < arapk_grid_scratch_bytes(0, G)) { set_error("valid_cells: scratch too small"); return ARAP_ERR_INVALID; }
  const long long gc = (long long)G * G * G;
  GridScratch s = carve(scratch, G, 0);
  k_valid_flags<<<(unsigned)((gc + 255) / 256), 256, 0, st>>>(gc, prefix, s.cnt); ARAP_KERNEL_CHECK();
  int rc = scan_inclusive(s.cnt, s.fill, gc, s.sums, st); if (rc) return rc;
  int V = 0;
  ARAP_CUDA_TRY(cudaMemcpyAsync(&V, s.fill + gc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  ARAP_CUDA_TRY(cudaStreamSynchronize(st));
  if (count_host) *count_host = V;
  if (valid_out && V > 0) { k_valid_compact<<<(unsigned)((gc + 255) / 256), 256, 0, st>>>(gc, s.cnt, s.fill, valid_out); ARAP_KERNEL_CHECK(); }
  return ARAP_OK;
}

extern "C" int arapk_emit_samples(const int* valid, int V, const float* min3, float step, int G, float* out, cudaStream_t st) {
  if (V <= 0) return ARAP_OK;
  const long long T = (long long)V * 64;
  k_emit_samples<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(V, valid, make_float3(min3[0], min3[1], min3[2]), step, G, out);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_ada_lpf(const float* samples, const int* valid, int V, float lpf_parameter, float* out, cudaStream_t st) {
  if (V <= 0) return ARAP_OK;
  k_ada_lpf<<<(V + 127) / 128, 128, 0, st>>>(V, samples, valid, lpf_parameter, out);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

// scratch: G^3 bytes (device)
extern "C" int arapk_judge_empty_grid(const int* valid, int V, const float* aim_opacity, int G, int* empty_out, void* scratch,
                                      size_t scratch_bytes, cudaStream_t st) {
  if (V <= 0) return ARAP_OK;
  const size_t gc = (size_t)G * G * G;
  if (scratch_bytes < gc) { set_error("judge_empty_grid: scratch too small"); return ARAP_ERR_INVALID; }
  uint8_t* bad = (uint8_t*)scratch;
  ARAP_CUDA_TRY(cudaMemsetAsync(bad, 1, gc, st));
  k_judge_mark<<<(V + 127) / 128, 128, 0, st>>>(V, valid, aim_opacity, G, bad);
  k_judge_collect<<<(V + 127) / 128, 128, 0, st>>>(V, valid, bad, empty_out);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_grid_eval(const int* valid, int V, const int* prefix, const int* lists, const float* samples, const float* pos,
                               const float* rot, const float* scale, const float* opacity, const float* shs, const float* ada_lpf,
                               float* out_feature, float* out_opacity, cudaStream_t st) {
  if (V <= 0) return ARAP_OK;
  k_grid_eval<<<V, 64, 0, st>>>(valid, prefix, lists, samples, pos, rot, scale, opacity, shs, ada_lpf, out_feature, out_opacity);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}
