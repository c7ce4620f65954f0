// Stage (b): farthest-point node sampling, grid-bucketed exact kNN and
// skinning weights.
//
//   fps          farthest_control_points_sampling + FetchFirstNodeIdx (helper.cpp:139-195)
//   knn_weights  DeformGraph::findNearestNodes + computeWeights       (Deform.hpp:153-208)
//
// The reference's kNN is a brute-force float distance + partial selection sort
// whose tie-break depends on the swap history.  Here nodes are
// bucketed in a uniform grid; each query expands Chebyshev rings until the
// (k+1)-th distance is provably final.  Distances use the reference's exact
// float arithmetic (library built with -fmad=false): sqrtf((dx*dx+dy*dy)+dz*dz).
// When no two candidates tie, the selection sort's output is simply ascending
// distance.  Queries that see a tie are re-done by a warp that replays the
// literal selection sort on the compacted candidate set (exact, see
// k_knn_slow).
#include <cooperative_groups.h>
#include <cfloat>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace arapgs {

// ------------------------------------------------------------------ FPS
struct FpsBest { float v; int i; };

__device__ __forceinline__ FpsBest fps_better(FpsBest a, FpsBest b) {
  // std::max_element: first (lowest index) maximum
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// Persistent cooperative kernel: one pass over all points per selected node.
// pts_distance (helper.cpp:60-63) evaluates in double (std::pow(float,int)
// promotes), rounds the sqrt to float.
__global__ void __launch_bounds__(512)
k_fps(const float* __restrict__ pos, long long N, int m, int first, float* __restrict__ dist,
      FpsBest* __restrict__ block_best /* 2 x gridDim */, int* __restrict__ out) {
  cg::grid_group grid = cg::this_grid();
  __shared__ FpsBest s_best[16];
  __shared__ int s_last;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  for (long long j = tid; j < N; j += nth) dist[j] = FLT_MAX;
  if (tid == 0) out[0] = first;
  int last = first;
  for (int c = 1; c < m; c++) {
    const float lx = pos[3LL * last], ly = pos[3LL * last + 1], lz = pos[3LL * last + 2];
    FpsBest best{-1.0f, 0x7fffffff};
    for (long long j = tid; j < N; j += nth) {
      const double dx = (double)(lx - pos[3 * j]), dy = (double)(ly - pos[3 * j + 1]), dz = (double)(lz - pos[3 * j + 2]);
      const float dd = (float)sqrt(dx * dx + dy * dy + dz * dz);
      float cur = dist[j];
      if (dd < cur) { cur = dd; dist[j] = dd; }
      best = fps_better(best, FpsBest{cur, (int)j});
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      FpsBest b{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
      best = fps_better(best, b);
    }
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
      FpsBest b = (threadIdx.x < (blockDim.x >> 5)) ? s_best[threadIdx.x] : FpsBest{-1.0f, 0x7fffffff};
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        FpsBest t{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
        b = fps_better(b, t);
      }
      if (threadIdx.x == 0) block_best[(c & 1) * gridDim.x + blockIdx.x] = b;
    }
    grid.sync();
    if (threadIdx.x < 32) {
      FpsBest b{-1.0f, 0x7fffffff};
      for (int t = threadIdx.x; t < (int)gridDim.x; t += 32) b = fps_better(b, block_best[(c & 1) * gridDim.x + t]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        FpsBest t{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)};
        b = fps_better(b, t);
      }
      if (threadIdx.x == 0) { s_last = b.i; if (blockIdx.x == 0) out[c] = b.i; }
    }
    __syncthreads();
    last = s_last;
  }
}

// FetchFirstNodeIdx: argmax of (x+y)+z, first maximum, start value -100000.
__global__ void k_fps_first(const float* __restrict__ pos, long long N, FpsBest* __restrict__ block_best) {
  __shared__ FpsBest s_best[32];
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  FpsBest best{-100000.0f, 0};  // ties with the start value never win (strict >), index 0 default
  bool any = false;
  for (long long j = tid; j < N; j += nth) {
    float c = 0.0f; c += pos[3 * j]; c += pos[3 * j + 1]; c += pos[3 * j + 2];
    if (c > best.v || (any && c == best.v && (int)j < best.i)) { best.v = c; best.i = (int)j; any = true; }
  }
  if (!any) best.i = 0x7fffffff;
  auto comb = [](FpsBest a, FpsBest b) {
    if (b.i == 0x7fffffff) return a;
    if (a.i == 0x7fffffff) return b;
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    FpsBest b{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)};
    best = comb(best, b);
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    FpsBest b = s_best[0];
    for (int t = 1; t < (int)(blockDim.x >> 5); t++) b = comb(b, s_best[t]);
    block_best[blockIdx.x] = b;
  }
}

// ------------------------------------------------------------------ node buckets
struct BucketGrid {
  float min[3];
  float h, inv_h;
  int dim[3];
};

__device__ __forceinline__ int bucket_coord(float p, float mn, float inv_h, int dim) {
  int c = (int)floorf((p - mn) * inv_h);
  return max(0, min(dim - 1, c));
}

__global__ void k_bucket_count(int M, const float* __restrict__ nodes, BucketGrid bg, int* __restrict__ cell_cnt,
                               int* __restrict__ node_cell) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int cx = bucket_coord(nodes[3 * i], bg.min[0], bg.inv_h, bg.dim[0]);
  const int cy = bucket_coord(nodes[3 * i + 1], bg.min[1], bg.inv_h, bg.dim[1]);
  const int cz = bucket_coord(nodes[3 * i + 2], bg.min[2], bg.inv_h, bg.dim[2]);
  const int c = (cx * bg.dim[1] + cy) * bg.dim[2] + cz;
  node_cell[i] = c;
  atomicAdd(&cell_cnt[c], 1);
}

// single-block exclusive scan (cells <= 2^21), writes start[ncell+1]
__global__ void k_bucket_scan(int ncell, const int* __restrict__ cnt, int* __restrict__ start) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < ncell; base += blockDim.x) {
    const int i = base + threadIdx.x;
    int v = (i < ncell) ? cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int wv = (threadIdx.x < (blockDim.x >> 5)) ? s_warp[threadIdx.x] : 0;
      int y = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, y, o); if (threadIdx.x >= o) y += t; }
      s_warp[threadIdx.x] = y - wv;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[threadIdx.x >> 5] + x - v;
    if (i < ncell) start[i] = excl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) start[ncell] = s_carry;
}

__global__ void k_bucket_fill(int M, const float* __restrict__ nodes, const int* __restrict__ node_cell,
                              const int* __restrict__ start, int* __restrict__ fill, float4* __restrict__ sorted) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  const int c = node_cell[i];
  const int slot = start[c] + atomicAdd(&fill[c], 1);
  sorted[slot] = make_float4(nodes[3 * i], nodes[3 * i + 1], nodes[3 * i + 2], __int_as_float(i));
}

// ------------------------------------------------------------------ kNN fast path
__device__ __forceinline__ float knn_dist(float qx, float qy, float qz, float nx, float ny, float nz) {
  const float t0 = qx - nx, t1 = qy - ny, t2 = qz - nz;
  return sqrtf(t0 * t0 + t1 * t1 + t2 * t2);
}

// weights from the k+1 sorted distances (computeWeights, Deform.hpp:187-208)
template <int KQ>
__device__ __forceinline__ void knn_weights(const float (&d)[KQ], int k, double (&w)[KQ]) {
  const double dmax = (double)d[k];
  double sum = 0.0;
#pragma unroll
  for (int j = 0; j < KQ - 1; j++) {
    if (j < k) {
      const double u = 1.0 - (double)d[j] / dmax;
      w[j] = u * u;
      sum += w[j];
    }
  }
  if (k == 1) w[0] = 1.0;
  else {
#pragma unroll
    for (int j = 0; j < KQ - 1; j++) if (j < k) w[j] = w[j] / sum;
  }
}

struct KnnOut {
  // plain layout (tests / generic API): idx Q x k (uint32), w Q x k (double), either may be null
  uint32_t* idx_plain; double* w_plain;
  // blocked layout (internal tables): idx uint16, w double and/or float, any may be null
  uint16_t* idx_blk; double* w_blk; float* wf_blk;
  // optional full k+1 index output (node graph edges use idx[1..k])
  uint32_t* idx_kq;
};

template <int KQ>
__device__ __forceinline__ void knn_store(const KnnOut& o, long long q, int k, const float (&d)[KQ], const int (&id)[KQ]) {
  double w[KQ];
  knn_weights<KQ>(d, k, w);
  const long long base = (q >> 5) * (long long)(k * 32) + (q & 31);
#pragma unroll
  for (int j = 0; j < KQ; j++) {
    if (j < k) {
      if (o.idx_plain) o.idx_plain[q * k + j] = (uint32_t)id[j];
      if (o.w_plain) o.w_plain[q * k + j] = w[j];
      if (o.idx_blk) o.idx_blk[base + j * 32] = (uint16_t)id[j];
      if (o.w_blk) o.w_blk[base + j * 32] = w[j];
      if (o.wf_blk) o.wf_blk[base + j * 32] = (float)w[j];
    }
    if (j <= k && o.idx_kq) o.idx_kq[q * (k + 1) + j] = (uint32_t)id[j];
  }
}

// KQ = compile-time capacity (k+1 <= KQ)
template <int KQ>
__global__ void __launch_bounds__(128)
k_knn_fast(const float* __restrict__ queries, long long Q, int k, BucketGrid bg, const int* __restrict__ cell_start,
           const float4* __restrict__ sorted, KnnOut out, int* __restrict__ slow_count, long long* __restrict__ slow_list,
           float* __restrict__ slow_thr) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int kq = k + 1;
  const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
  float d[KQ]; int id[KQ];
#pragma unroll
  for (int j = 0; j < KQ; j++) { d[j] = FLT_MAX; id[j] = -1; }
  bool tie = false;
  int found = 0;
  const int cx = bucket_coord(qx, bg.min[0], bg.inv_h, bg.dim[0]);
  const int cy = bucket_coord(qy, bg.min[1], bg.inv_h, bg.dim[1]);
  const int cz = bucket_coord(qz, bg.min[2], bg.inv_h, bg.dim[2]);
  const int rmax = max(max(max(cx, bg.dim[0] - 1 - cx), max(cy, bg.dim[1] - 1 - cy)), max(cz, bg.dim[2] - 1 - cz));
  for (int r = 0; r <= rmax; r++) {
    const int x0 = max(cx - r, 0), x1 = min(cx + r, bg.dim[0] - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, bg.dim[1] - 1);
    const int z0 = max(cz - r, 0), z1 = min(cz + r, bg.dim[2] - 1);
    for (int x = x0; x <= x1; x++) {
      const bool xs = (x == cx - r) || (x == cx + r);
      for (int y = y0; y <= y1; y++) {
        const bool ys = xs || (y == cy - r) || (y == cy + r);
        // shell cells only: if not on an x/y face, visit just the two z faces
        const int zstep = ys ? 1 : max(2 * r, 1);
        for (int z = ys ? z0 : cz - r; z <= z1; z += zstep) {
          if (z < z0) continue;
          const int c = (x * bg.dim[1] + y) * bg.dim[2] + z;
          const int b = cell_start[c], e = cell_start[c + 1];
          for (int t = b; t < e; t++) {
            const float4 n = __ldg(sorted + t);
            const float dd = knn_dist(qx, qy, qz, n.x, n.y, n.z);
            found++;
            const float worst = d[KQ - 1 < kq - 1 ? KQ - 1 : kq - 1];
            (void)worst;
            float dw = FLT_MAX;
#pragma unroll
            for (int j = 0; j < KQ; j++) if (j == kq - 1) dw = d[j];
            if (dd > dw) continue;
            if (dd == dw) { tie = true; continue; }
            // insert, keeping ascending order among the first kq slots
            float cd = dd; int ci = __float_as_int(n.w);
#pragma unroll
            for (int j = 0; j < KQ; j++) {
              if (j < kq) {
                if (cd == d[j] && cd != FLT_MAX) tie = true;  // empty slots hold FLT_MAX and must not count as ties
                if (cd < d[j]) { const float td = d[j]; const int ti = id[j]; d[j] = cd; id[j] = ci; cd = td; ci = ti; }
              }
            }
          }
        }
      }
    }
    if (found >= kq) {
      // distance from the query to the unexplored region (cells beyond ring r);
      // faces on the bucket-grid boundary have nothing behind them.
      float bound = FLT_MAX;
      if (cx - r > 0) bound = fminf(bound, qx - (bg.min[0] + (cx - r) * bg.h));
      if (cx + r < bg.dim[0] - 1) bound = fminf(bound, (bg.min[0] + (cx + r + 1) * bg.h) - qx);
      if (cy - r > 0) bound = fminf(bound, qy - (bg.min[1] + (cy - r) * bg.h));
      if (cy + r < bg.dim[1] - 1) bound = fminf(bound, (bg.min[1] + (cy + r + 1) * bg.h) - qy);
      if (cz - r > 0) bound = fminf(bound, qz - (bg.min[2] + (cz - r) * bg.h));
      if (cz + r < bg.dim[2] - 1) bound = fminf(bound, (bg.min[2] + (cz + r + 1) * bg.h) - qz);
      float dw = FLT_MAX;
#pragma unroll
      for (int j = 0; j < KQ; j++) if (j == kq - 1) dw = d[j];
      // conservative margin: cell assignment and bound use rounded float arithmetic
      if (dw < bound - 1e-5f * (fabsf(bound) + bg.h)) break;
    }
  }
  if (tie) {
    float dw = FLT_MAX;
#pragma unroll
    for (int j = 0; j < KQ; j++) if (j == kq - 1) dw = d[j];
    const int slot = atomicAdd(slow_count, 1);
    slow_list[slot] = q;
    slow_thr[slot] = dw;
    return;
  }
  knn_store<KQ>(out, q, k, d, id);
}

// ------------------------------------------------------------------ kNN exact slow path
// One warp per tied query.  The candidate set U = {nodes 0..k} U {d <= thr} is
// compacted in index order; positions of U's elements are invariant under the
// reference's swaps, so replaying the literal selection sort (Deform.hpp:167-184)
// on the compacted array gives the reference's output exactly.
constexpr int SLOW_CAP = 192;
template <int KQ>
__global__ void __launch_bounds__(128)
k_knn_slow(const float* __restrict__ queries, int nslow, const long long* __restrict__ slow_list,
           const float* __restrict__ slow_thr, int k, int M, const float* __restrict__ nodes, KnnOut out,
           int* __restrict__ err_flag) {
  __shared__ float s_d[4][SLOW_CAP];
  __shared__ int s_i[4][SLOW_CAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x * 4 + warp;
  if (item >= nslow) return;
  const long long q = slow_list[item];
  const float thr = slow_thr[item];
  const int kq = k + 1;
  const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
  int cnt = 0;
  for (int base = 0; base < M; base += 32) {
    const int j = base + lane;
    float dd = FLT_MAX; bool keep = false;
    if (j < M) {
      dd = knn_dist(qx, qy, qz, nodes[3 * j], nodes[3 * j + 1], nodes[3 * j + 2]);
      keep = (dd <= thr) || (j < kq);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int p = cnt + __popc(m & ((1u << lane) - 1));
      if (p < SLOW_CAP) { s_d[warp][p] = dd; s_i[warp][p] = j; }
    }
    cnt += __popc(m);
  }
  __syncwarp();
  if (cnt > SLOW_CAP) { if (lane == 0) atomicExch(err_flag, 1); return; }
  if (lane == 0) {
    float d[KQ]; int id[KQ];
    float* sd = s_d[warp]; int* si = s_i[warp];
    for (int i = 0; i < kq; i++) {
      int m = i;
      for (int j = cnt - 1; j > i; j--) if (sd[j] < sd[m]) m = j;
      const float dm = sd[m]; const int im = si[m];
      sd[m] = sd[i]; si[m] = si[i]; sd[i] = dm; si[i] = im;
    }
#pragma unroll
    for (int j = 0; j < KQ; j++) { d[j] = (j < kq) ? sd[j] : FLT_MAX; id[j] = (j < kq) ? si[j] : -1; }
    knn_store<KQ>(out, q, k, d, id);
  }
}

// min/max reduction of a point set (bucket grid bounds, scene AABB)
__global__ void k_minmax(const float* __restrict__ p, long long N, float* __restrict__ out6 /* init +inf/-inf */) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (long long i = tid; i < N; i += nth)
#pragma unroll
    for (int c = 0; c < 3; c++) { const float v = p[3 * i + c]; mn[c] = fminf(mn[c], v); mx[c] = fmaxf(mx[c], v); }
#pragma unroll
  for (int c = 0; c < 3; c++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      // float atomic min/max via ordered-int trick
      int* a = reinterpret_cast<int*>(out6 + c);
      int* b = reinterpret_cast<int*>(out6 + 3 + c);
      if (mn[c] >= 0) atomicMin(a, __float_as_int(mn[c])); else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(mn[c]));
      if (mx[c] >= 0) atomicMax(b, __float_as_int(mx[c])); else atomicMin(reinterpret_cast<unsigned*>(b), __float_as_uint(mx[c]));
    }
  }
}

}  // namespace arapgs

using namespace arapgs;

// ===========================================================================
// launchers
// ===========================================================================
extern "C" int arapk_minmax(const float* pts, long long N, float* out6_dev, cudaStream_t st) {
  const float init[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
  ARAP_CUDA_TRY(cudaMemcpyAsync(out6_dev, init, sizeof(init), cudaMemcpyHostToDevice, st));
  if (N <= 0) return ARAP_OK;
  const int grid = (int)std::min<long long>((N + 255) / 256, 148 * 8);
  k_minmax<<<grid, 256, 0, st>>>(pts, N, out6_dev);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

// out_idx: m ints (device).  scratch: N floats + 2*grid FpsBest (device), provided by caller:
//   scratch_bytes >= N*4 + 64*1024
extern "C" int arapk_fps(const float* pos, long long N, int node_num, int* out_idx_dev, void* scratch, size_t scratch_bytes,
                         int* out_count_host, cudaStream_t st) {
  const int m = (int)std::min<long long>(node_num, N);
  if (out_count_host) *out_count_host = m;
  if (m <= 0) return ARAP_OK;
  if (scratch_bytes < (size_t)N * 4 + 65536) { set_error("fps: scratch too small"); return ARAP_ERR_INVALID; }
  float* dist = (float*)scratch;
  FpsBest* bb = (FpsBest*)((char*)scratch + (((size_t)N * 4 + 255) / 256) * 256);
  int dev = 0, sms = 0, per_sm = 0;
  ARAP_CUDA_TRY(cudaGetDevice(&dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // first node
  const int g1 = (int)std::min<long long>((N + 255) / 256, 1024);
  k_fps_first<<<g1, 256, 0, st>>>(pos, N, bb);
  ARAP_KERNEL_CHECK();
  std::vector<FpsBest> hb(g1);
  ARAP_CUDA_TRY(cudaMemcpyAsync(hb.data(), bb, sizeof(FpsBest) * g1, cudaMemcpyDeviceToHost, st));
  ARAP_CUDA_TRY(cudaStreamSynchronize(st));
  FpsBest best{-100000.0f, 0x7fffffff};
  for (auto& b : hb) {
    if (b.i == 0x7fffffff) continue;
    if (best.i == 0x7fffffff || b.v > best.v || (b.v == best.v && b.i < best.i)) best = b;
  }
  int first = best.i == 0x7fffffff ? 0 : best.i;
  ARAP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fps, 512, 0));
  if (per_sm < 1) { set_error("fps: kernel cannot be co-resident"); return ARAP_ERR_CUDA; }
  int grid = sms * std::min(per_sm, 2);
  grid = (int)std::min<long long>(grid, std::max<long long>(1, (N + 511) / 512));
  if ((size_t)grid * 2 * sizeof(FpsBest) > 65536) grid = 65536 / (2 * sizeof(FpsBest));
  long long Nn = N; int mm = m;
  void* args[] = {(void*)&pos, (void*)&Nn, (void*)&mm, (void*)&first, (void*)&dist, (void*)&bb, (void*)&out_idx_dev};
  ARAP_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_fps, dim3(grid), dim3(512), args, 0, st));
  return ARAP_OK;
}

namespace {
struct BucketState {
  BucketGrid bg; int ncell;
};
}

// Bucket workspace layout (device, caller-owned, >= arapk_knn_workspace_bytes(M)):
//   [cell_cnt ncell][cell_start ncell+1][fill ncell][node_cell M][sorted M float4][minmax 6 floats][slow_count][err]
extern "C" size_t arapk_knn_workspace_bytes(int M) {
  const size_t max_cells = 128 * 128 * 128;
  return (max_cells * 3 + 16) * sizeof(int) + (size_t)M * sizeof(int) + (size_t)M * sizeof(float4) + 4096;
}

struct ArapKnnIndex {
  BucketGrid bg; int ncell; int M;
  int *cell_cnt, *cell_start, *fill, *node_cell; float4* sorted; float* minmax; int* counters;
  const float* nodes;
};

extern "C" size_t arapk_knn_index_struct_bytes() { return sizeof(ArapKnnIndex); }

extern "C" int arapk_knn_build(const float* nodes_dev, int M, void* workspace, size_t workspace_bytes, void* index_out,
                               cudaStream_t st) {
  if (M < 1) { set_error("knn_build: no nodes"); return ARAP_ERR_INVALID; }
  if (workspace_bytes < arapk_knn_workspace_bytes(M)) { set_error("knn_build: workspace too small"); return ARAP_ERR_INVALID; }
  ArapKnnIndex* ix = (ArapKnnIndex*)index_out;
  char* p = (char*)workspace;
  const size_t max_cells = 128 * 128 * 128;
  ix->cell_cnt = (int*)p; p += max_cells * sizeof(int);
  ix->cell_start = (int*)p; p += (max_cells + 16) * sizeof(int);
  ix->fill = (int*)p; p += max_cells * sizeof(int);
  ix->node_cell = (int*)p; p += (((size_t)M * sizeof(int) + 255) / 256) * 256;
  ix->sorted = (float4*)p; p += (size_t)M * sizeof(float4);
  p = (char*)((((uintptr_t)p + 255) / 256) * 256);
  ix->minmax = (float*)p; p += 256;
  ix->counters = (int*)p;
  ix->nodes = nodes_dev; ix->M = M;
  int rc = arapk_minmax(nodes_dev, M, ix->minmax, st); if (rc) return rc;
  float mm[6];
  ARAP_CUDA_TRY(cudaMemcpyAsync(mm, ix->minmax, sizeof(mm), cudaMemcpyDeviceToHost, st));
  ARAP_CUDA_TRY(cudaStreamSynchronize(st));
  const float ext = std::max(std::max(mm[3] - mm[0], mm[4] - mm[1]), std::max(mm[5] - mm[2], 1e-20f));
  // ~cbrt(M) buckets along the longest axis: a few nodes per occupied bucket for
  // surface-like node sets, ring 1-2 usually closes the search
  int B = (int)std::lround(std::cbrt((double)M) * 1.25);
  B = std::max(1, std::min(128, B));
  BucketGrid bg;
  bg.h = ext / (float)B * 1.0001f;
  bg.inv_h = 1.0f / bg.h;
  for (int c = 0; c < 3; c++) {
    bg.min[c] = mm[c];
    bg.dim[c] = std::max(1, std::min(128, (int)std::floor((mm[3 + c] - mm[c]) * bg.inv_h) + 1));
  }
  ix->bg = bg; ix->ncell = bg.dim[0] * bg.dim[1] * bg.dim[2];
  ARAP_CUDA_TRY(cudaMemsetAsync(ix->cell_cnt, 0, sizeof(int) * ix->ncell, st));
  ARAP_CUDA_TRY(cudaMemsetAsync(ix->fill, 0, sizeof(int) * ix->ncell, st));
  k_bucket_count<<<(M + 255) / 256, 256, 0, st>>>(M, nodes_dev, bg, ix->cell_cnt, ix->node_cell);
  ARAP_KERNEL_CHECK();
  k_bucket_scan<<<1, 1024, 0, st>>>(ix->ncell, ix->cell_cnt, ix->cell_start);
  ARAP_KERNEL_CHECK();
  k_bucket_fill<<<(M + 255) / 256, 256, 0, st>>>(M, nodes_dev, ix->node_cell, ix->cell_start, ix->fill, ix->sorted);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

// slow_scratch: Q * 12 bytes (list + thr) device
extern "C" int arapk_knn_query(const void* index, const float* queries_dev, long long Q, int k, uint32_t* idx_plain,
                               double* w_plain, uint16_t* idx_blk, double* w_blk, float* wf_blk, uint32_t* idx_kq,
                               void* slow_scratch, size_t slow_scratch_bytes, int* n_slow_host, cudaStream_t st) {
  const ArapKnnIndex* ix = (const ArapKnnIndex*)index;
  if (Q <= 0) return ARAP_OK;
  if (k < 1 || k > KNN_MAX) { set_error("knn_query: k must be in [1,12]"); return ARAP_ERR_INVALID; }
  if (ix->M < k + 1) { set_error("knn_query: need at least k+1 nodes"); return ARAP_ERR_INVALID; }
  if (idx_blk && ix->M > 65536) { set_error("knn_query: blocked uint16 tables need <= 65536 nodes"); return ARAP_ERR_UNSUPPORTED; }
  if (slow_scratch_bytes < (size_t)Q * 12 + 256) { set_error("knn_query: slow scratch too small"); return ARAP_ERR_INVALID; }
  long long* slow_list = (long long*)slow_scratch;
  float* slow_thr = (float*)((char*)slow_scratch + (((size_t)Q * 8 + 255) / 256) * 256);
  if ((char*)(slow_thr + Q) > (char*)slow_scratch + slow_scratch_bytes) { set_error("knn_query: slow scratch too small"); return ARAP_ERR_INVALID; }
  KnnOut out{idx_plain, w_plain, idx_blk, w_blk, wf_blk, idx_kq};
  ARAP_CUDA_TRY(cudaMemsetAsync(ix->counters, 0, 2 * sizeof(int), st));
  const unsigned grid = (unsigned)((Q + 127) / 128);
  k_knn_fast<KNN_MAX + 1><<<grid, 128, 0, st>>>(queries_dev, Q, k, ix->bg, ix->cell_start, ix->sorted, out, ix->counters,
                                                slow_list, slow_thr);
  ARAP_KERNEL_CHECK();
  int h[2] = {0, 0};
  ARAP_CUDA_TRY(cudaMemcpyAsync(h, ix->counters, sizeof(h), cudaMemcpyDeviceToHost, st));
  ARAP_CUDA_TRY(cudaStreamSynchronize(st));
  if (n_slow_host) *n_slow_host = h[0];
  if (h[0] > 0) {
    k_knn_slow<KNN_MAX + 1><<<(h[0] + 3) / 4, 128, 0, st>>>(queries_dev, h[0], slow_list, slow_thr, k, ix->M, ix->nodes, out,
                                                           ix->counters + 1);
    ARAP_KERNEL_CHECK();
    ARAP_CUDA_TRY(cudaMemcpyAsync(h, ix->counters, sizeof(h), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaStreamSynchronize(st));
    if (h[1]) { set_error("knn_query: more than 192 tied candidates for one query"); return ARAP_ERR_KNN_TIES; }
  }
  return ARAP_OK;
}
