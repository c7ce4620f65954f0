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
#include <cstdlib>
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

// ------------------------------------------------------------------ FPS, pruned by the density grid
// After the first passes the update of a selection step is local: a point's minimum distance can only drop if the new node
// is closer to it than its current value, and every current value is <= the distance at which the new node was picked.
// The candidate points are the cell-ordered Gaussians (cell c = [prefix[c-1], prefix[c])), so per-cell maxima
// (value, first index) and per-supercell maxima (SG^3 cells) are kept next to the per-point distances; a step visits only
// the supercells / cells whose box is closer to the new node than their maximum, updates their points with the exact
// reference expression and repairs the two maxima.  The arg-max is the reduction of the supercell maxima with the
// reference's first-maximum tie-break (value, then lowest index), so the selected sequence is the reference's, bit for bit.
// One CTA runs the whole loop (a step is ~2 us of dependent loads; a grid-wide barrier per step would cost more than that).
struct FpsGrid { float min[3]; float step; int G, SG, ns; };
constexpr int FPS_LIST = 3072;   // marked cells of one selection step (shared-memory list); more -> the per-supercell walk   // ns = supercells per axis, SG = cells per supercell axis

__device__ __forceinline__ FpsBest fps_better_or_empty(FpsBest a, FpsBest b) {   // v < 0 marks "no point"
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}

// conservative lower bound of the distance from p to the box [lo, hi]^3 of cells (rounded float arithmetic on both sides)
__device__ __forceinline__ float fps_box_mind(const float* p, const FpsGrid& g, int x0, int y0, int z0, int x1, int y1, int z1) {
  const int lo[3] = {x0, y0, z0}, hi[3] = {x1, y1, z1};
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const float bl = g.min[a] + lo[a] * g.step, bh = g.min[a] + (hi[a] + 1) * g.step;
    const float gap = fmaxf(0.f, fmaxf(bl - p[a], p[a] - bh));
    s += gap * gap;
  }
  return fmaxf(0.f, sqrtf(s) - 1e-4f * g.step) * 0.99999f;
}

__global__ void k_fps_cellmax(long long ncell, const int* __restrict__ prefix, const float* __restrict__ dist, FpsBest* __restrict__ cm) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int b = c ? prefix[c - 1] : 0, e = prefix[c];
  FpsBest best{-1.0f, 0x7fffffff};
  for (int j = b; j < e; j++) best = fps_better_or_empty(best, FpsBest{dist[j], j});
  cm[c] = best;
}

__global__ void __launch_bounds__(1024)
k_fps_pruned(const float* __restrict__ pos, FpsGrid g, const int* __restrict__ prefix, float* __restrict__ dist, FpsBest* __restrict__ cm,
             int m0, int m, int* __restrict__ out) {
  __shared__ FpsBest s_sm[4096];       // supercell maxima
  __shared__ FpsBest s_w[32];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ns = g.ns, nsup = ns * ns * ns, G = g.G, SG = g.SG, cps = SG * SG * SG;
  const int sgs = 31 - __clz(SG);    // SG is a power of two
  // supercell maxima from the cell maxima
  for (int s = warp; s < nsup; s += 32) {
    const int sx = s / (ns * ns), sy = (s / ns) % ns, sz = s % ns;
    FpsBest best{-1.0f, 0x7fffffff};
    for (int t = lane; t < cps; t += 32) {
      const int x = sx * SG + t / (SG * SG), y = sy * SG + (t / SG) % SG, z = sz * SG + t % SG;
      if (x < G && y < G && z < G) best = fps_better_or_empty(best, cm[((long long)x * G + y) * G + z]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fps_better_or_empty(best, FpsBest{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)});
    if (lane == 0) s_sm[s] = best;
  }
  __shared__ float s_rad;
  __shared__ int s_list[FPS_LIST];
  __shared__ int s_nlist;
  __syncthreads();
  {   // radius of the first pruned step: the current maximum of the minimum distances (the value out[m0 - 1] was picked at)
    float v = -1.0f;
    for (int s = tid; s < nsup; s += 1024) v = fmaxf(v, s_sm[s].v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) s_w[warp].v = v;
    __syncthreads();
    if (tid == 0) {
      float r = -1.0f;
      for (int w = 0; w < 32; w++) r = fmaxf(r, s_w[w].v);
      s_rad = r; s_last = out[m0 - 1];
    }
  }
  __syncthreads();
  for (int c = m0; c < m; c++) {
    const int last = s_last;
    const float p[3] = {pos[3LL * last], pos[3LL * last + 1], pos[3LL * last + 2]};
    // every current minimum distance is <= the value at which `last` was picked: only supercells meeting that ball matter
    const float R = s_rad;
    int slo[3], shi[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float ext = SG * g.step;
      slo[a] = R > 1e30f ? 0 : max(0, min(ns - 1, (int)floorf((p[a] - R - g.min[a]) / ext) - 1));
      shi[a] = R > 1e30f ? ns - 1 : max(0, min(ns - 1, (int)floorf((p[a] + R - g.min[a]) / ext) + 1));
    }
    const int by = shi[1] - slo[1] + 1, bz = shi[2] - slo[2] + 1, nbox = (shi[0] - slo[0] + 1) * by * bz;
    // (1) mark: every cell of the supercells in the box whose own box is closer to the new node than its maximum
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    // one warp per supercell of the box; a supercell farther from the new node than its own maximum is skipped as a whole
    // (the selection loop runs on ONE SM: it is bound by instructions issued, so cells are only looked at where they can matter)
    for (int t = warp; t < nbox; t += 32) {
      const int sx = slo[0] + t / (by * bz), sy = slo[1] + (t / bz) % by, sz = slo[2] + t % bz;
      const FpsBest sb = s_sm[(sx * ns + sy) * ns + sz];
      if (sb.v < 0.f) continue;
      const int x0 = sx << sgs, y0 = sy << sgs, z0 = sz << sgs;
      if (fps_box_mind(p, g, x0, y0, z0, min(x0 + SG, G) - 1, min(y0 + SG, G) - 1, min(z0 + SG, G) - 1) >= sb.v) continue;
      // four cells per lane and pass: the (dependent, L2-latency) cell-maximum loads of a pass are in flight together —
      // the loop runs on one SM and is bound by exposed latency, not by bandwidth
      for (int u0 = lane; u0 < cps; u0 += 128) {
        int cell[4]; float md[4]; FpsBest cb[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int u = u0 + 32 * i;
          const int x = x0 + (u >> (2 * sgs)), y = y0 + ((u >> sgs) & (SG - 1)), z = z0 + (u & (SG - 1));
          cell[i] = -1; md[i] = 0.f; cb[i] = FpsBest{-1.0f, 0x7fffffff};
          if (u < cps && x < G && y < G && z < G) {
            md[i] = fps_box_mind(p, g, x, y, z, x, y, z);
            if (md[i] < sb.v) cell[i] = (x * G + y) * G + z;
          }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) if (cell[i] >= 0) cb[i] = cm[cell[i]];
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (cell[i] >= 0 && cb[i].v >= 0.f && md[i] < cb[i].v) {
            const int slot = atomicAdd(&s_nlist, 1);
            if (slot < FPS_LIST) s_list[slot] = cell[i];
          }
      }
    }
    __syncthreads();
    const int nlist = s_nlist;
    if (nlist <= FPS_LIST) {
      // (2) update: eight lanes per marked cell
      const int sub = tid & 7;
      for (int base = 0; base < nlist; base += 128) {     // warp-uniform trip count: the group reductions below use full-warp shuffles
        const int ci = base + (tid >> 3);
        const bool have = ci < nlist;
        const long long cc = have ? s_list[ci] : 0;
        const int b = have ? (cc ? prefix[cc - 1] : 0) : 0, e = have ? prefix[cc] : 0;
        FpsBest best{-1.0f, 0x7fffffff};
        for (int j0 = b + sub; j0 < e; j0 += 32) {     // four points per lane and pass, loads first
          float px[4], py[4], pz[4], cur[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int j = j0 + 8 * i;
            if (j < e) { px[i] = pos[3LL * j]; py[i] = pos[3LL * j + 1]; pz[i] = pos[3LL * j + 2]; cur[i] = dist[j]; }
          }
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int j = j0 + 8 * i;
            if (j < e) {
              const double dx = (double)(p[0] - px[i]), dy = (double)(p[1] - py[i]), dz = (double)(p[2] - pz[i]);
              const float dd = (float)sqrt(dx * dx + dy * dy + dz * dz);
              float cv = cur[i];
              if (dd < cv) { cv = dd; dist[j] = dd; }
              best = fps_better_or_empty(best, FpsBest{cv, j});
            }
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) best = fps_better_or_empty(best, FpsBest{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)});
        if (sub == 0 && have) cm[cc] = best;
      }
      __syncthreads();
      // (3) repair the maxima of the supercells in the box
      for (int t = warp; t < nbox; t += 32) {
        const int sx = slo[0] + t / (by * bz), sy = slo[1] + (t / bz) % by, sz = slo[2] + t % bz;
        const int s = (sx * ns + sy) * ns + sz;
        const float sv = s_sm[s].v;
        if (sv < 0.f) continue;
        const int x0 = sx << sgs, y0 = sy << sgs, z0 = sz << sgs;
        if (fps_box_mind(p, g, x0, y0, z0, min(x0 + SG, G) - 1, min(y0 + SG, G) - 1, min(z0 + SG, G) - 1) >= sv) continue;   // untouched in (1)
        FpsBest sup{-1.0f, 0x7fffffff};
        for (int u0 = lane; u0 < cps; u0 += 128) {
          FpsBest cb[4];
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int u = u0 + 32 * i;
            const int x = x0 + (u >> (2 * sgs)), y = y0 + ((u >> sgs) & (SG - 1)), z = z0 + (u & (SG - 1));
            cb[i] = FpsBest{-1.0f, 0x7fffffff};
            if (u < cps && x < G && y < G && z < G) cb[i] = cm[((long long)x * G + y) * G + z];
          }
#pragma unroll
          for (int i = 0; i < 4; i++) sup = fps_better_or_empty(sup, cb[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sup = fps_better_or_empty(sup, FpsBest{__shfl_xor_sync(0xffffffffu, sup.v, o), __shfl_xor_sync(0xffffffffu, sup.i, o)});
        __syncwarp();   // every lane has read s_sm[s] above (the shuffles already order it; this states it for racecheck)
        if (lane == 0) s_sm[s] = sup;
      }
    } else
    // the marked cells do not fit the list (early steps, large radius): one warp per supercell walks its cells
    for (int t = warp; t < nbox; t += 32) {
      const int sx = slo[0] + t / (by * bz), sy = slo[1] + (t / bz) % by, sz = slo[2] + t % bz;
      const int s = (sx * ns + sy) * ns + sz;
      const FpsBest sb = s_sm[s];
      if (sb.v < 0.f) continue;
      const int x0 = sx * SG, y0 = sy * SG, z0 = sz * SG;
      if (fps_box_mind(p, g, x0, y0, z0, min(x0 + SG, G) - 1, min(y0 + SG, G) - 1, min(z0 + SG, G) - 1) >= sb.v) continue;   // warp-uniform
      FpsBest sup{-1.0f, 0x7fffffff};
      for (int tb = 0; tb < cps; tb += 32) {
        const int t2 = tb + lane;
        const int x = x0 + t2 / (SG * SG), y = y0 + (t2 / SG) % SG, z = z0 + t2 % SG;
        const bool in = t2 < cps && x < G && y < G && z < G;
        const long long cell = ((long long)x * G + y) * G + z;
        FpsBest cb{-1.0f, 0x7fffffff};
        if (in) cb = cm[cell];
        const bool need = in && cb.v >= 0.f && fps_box_mind(p, g, x, y, z, x, y, z) < cb.v;
        unsigned todo = __ballot_sync(0xffffffffu, need);
        while (todo) {     // the warp updates the points of one marked cell at a time
          const int src = __ffs(todo) - 1; todo &= todo - 1;
          const long long cc = __shfl_sync(0xffffffffu, cell, src);
          const int b = cc ? prefix[cc - 1] : 0, e = prefix[cc];
          FpsBest best{-1.0f, 0x7fffffff};
          for (int j = b + lane; j < e; j += 32) {
            const double dx = (double)(p[0] - pos[3LL * j]), dy = (double)(p[1] - pos[3LL * j + 1]), dz = (double)(p[2] - pos[3LL * j + 2]);
            const float dd = (float)sqrt(dx * dx + dy * dy + dz * dz);
            float cur = dist[j];
            if (dd < cur) { cur = dd; dist[j] = dd; }
            best = fps_better_or_empty(best, FpsBest{cur, j});
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) best = fps_better_or_empty(best, FpsBest{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)});
          if (lane == src) { cb = best; cm[cc] = best; }
        }
        sup = fps_better_or_empty(sup, cb);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sup = fps_better_or_empty(sup, FpsBest{__shfl_xor_sync(0xffffffffu, sup.v, o), __shfl_xor_sync(0xffffffffu, sup.i, o)});
      if (lane == 0) s_sm[s] = sup;
    }
    __syncthreads();
    FpsBest best{-1.0f, 0x7fffffff};
    for (int s = tid; s < nsup; s += 1024) best = fps_better_or_empty(best, s_sm[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fps_better_or_empty(best, FpsBest{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.i, o)});
    if (lane == 0) s_w[warp] = best;
    __syncthreads();
    if (warp == 0) {
      FpsBest b = s_w[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = fps_better_or_empty(b, FpsBest{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.i, o)});
      if (lane == 0) { s_last = b.i; s_rad = b.v; out[c] = b.i; }
    }
    __syncthreads();
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

// Ring search of one query over the node buckets: expands Chebyshev rings until the (k+1)-th distance is provably final.
// d / id: ascending (k+1) best; tie: two candidates at the same distance among / at the rim of the best (exact replay needed).
template <int KQ>
__device__ __forceinline__ void knn_ring_search(float qx, float qy, float qz, int kq, const BucketGrid& bg, const int* __restrict__ cell_start,
                                                const float4* __restrict__ sorted, float (&d)[KQ], int (&id)[KQ], bool& tie) {
#pragma unroll
  for (int j = 0; j < KQ; j++) { d[j] = FLT_MAX; id[j] = -1; }
  tie = false;
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
}

// first version: one thread per query, every query walks the buckets on its own (kept as the fallback of k_knn_tile and
// for the few-query calls: node graph, bucket bounds)
template <int KQ>
__global__ void __launch_bounds__(128)
k_knn_fast(const float* __restrict__ queries, long long Q, int k, BucketGrid bg, const int* __restrict__ cell_start,
           const float4* __restrict__ sorted, KnnOut out, int* __restrict__ slow_count, long long* __restrict__ slow_list,
           float* __restrict__ slow_thr) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int kq = k + 1;
  const float qx = queries[3 * q], qy = queries[3 * q + 1], qz = queries[3 * q + 2];
  float d[KQ]; int id[KQ]; bool tie;
  knn_ring_search<KQ>(qx, qy, qz, kq, bg, cell_start, sorted, d, id, tie);
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

// Upper bound of the (KNN_MAX + 1)-th neighbour distance at every bucket centre (index build): for ANY point q,
//   d_(k+1)(q) <= |q - centre_b| + D_b     (triangle inequality, k <= KNN_MAX),
// which is what lets a tile of queries agree on one candidate set before looking at a single node.
__global__ void __launch_bounds__(128)
k_bucket_bounds(int ncell, int kq, BucketGrid bg, const int* __restrict__ cell_start, const float4* __restrict__ sorted,
                float* __restrict__ bound) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int z = c % bg.dim[2], y = (c / bg.dim[2]) % bg.dim[1], x = c / (bg.dim[2] * bg.dim[1]);
  const float qx = bg.min[0] + (x + 0.5f) * bg.h, qy = bg.min[1] + (y + 0.5f) * bg.h, qz = bg.min[2] + (z + 0.5f) * bg.h;
  float d[KNN_MAX + 1]; int id[KNN_MAX + 1]; bool tie;
  knn_ring_search<KNN_MAX + 1>(qx, qy, qz, kq, bg, cell_start, sorted, d, id, tie);
  float dw = FLT_MAX;
#pragma unroll
  for (int j = 0; j < KNN_MAX + 1; j++) if (j == kq - 1) dw = d[j];
  bound[c] = dw;
}

// ------------------------------------------------------------------ kNN, one candidate set per tile of queries
// Query families arrive in cell order (end points of cell-ordered Gaussians, the 64 samples of a cell), so a tile of 128
// consecutive queries is a compact cloud.  Each query bounds its own (k+1)-th distance by B(q) = |q - centre_b| + D_b;
// every node that can enter ANY of the tile's results lies in the tile's bounding box grown by max B.  The CTA collects
// the nodes of the buckets meeting that box once into shared memory (a few dozen), and every thread then selects its
// k+1 nearest from the staged list: broadcast shared-memory reads and no per-query bucket walk.  Arithmetic and tie
// handling are those of the per-query kernel (exact float distance expression, ties replayed by k_knn_slow); a tile
// whose candidate list does not fit falls back to the per-query walk.
constexpr int KT_TILE = 128;
constexpr int KT_CAP = 1024;   // staged candidates (power of two: sorted in place by a bitonic network)

// KQ = k + 1 at compile time (9, 11, 13 for k = 8, 10, 12; other k run with the next larger KQ and a run-time kq).
template <int KQ, bool EXACT_KQ>
__global__ void __launch_bounds__(KT_TILE)
k_knn_tile(const float* __restrict__ queries, long long Q, int k, BucketGrid bg, const int* __restrict__ cell_start,
           const float4* __restrict__ sorted, const float* __restrict__ bucket_bound, KnnOut out, int* __restrict__ slow_count,
           long long* __restrict__ slow_list, float* __restrict__ slow_thr) {
  __shared__ float4 s_cand[KT_CAP];
  __shared__ float s_key[KT_CAP];       // distance of a candidate to the tile centre (sort key)
  __shared__ float s_red[7][KT_TILE / 32];
  __shared__ int s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long q = (long long)blockIdx.x * KT_TILE + tid;
  const bool live = q < Q;
  const int kq = EXACT_KQ ? KQ : k + 1;
  float qx = 0.f, qy = 0.f, qz = 0.f, B = 0.f;
  float r[7] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f};   // bbox min, bbox max, max B
  if (live) {
    qx = queries[3 * q]; qy = queries[3 * q + 1]; qz = queries[3 * q + 2];
    const int cx = bucket_coord(qx, bg.min[0], bg.inv_h, bg.dim[0]);
    const int cy = bucket_coord(qy, bg.min[1], bg.inv_h, bg.dim[1]);
    const int cz = bucket_coord(qz, bg.min[2], bg.inv_h, bg.dim[2]);
    const float ux = qx - (bg.min[0] + (cx + 0.5f) * bg.h), uy = qy - (bg.min[1] + (cy + 0.5f) * bg.h), uz = qz - (bg.min[2] + (cz + 0.5f) * bg.h);
    const float Db = bucket_bound[(cx * bg.dim[1] + cy) * bg.dim[2] + cz];
    B = (sqrtf(ux * ux + uy * uy + uz * uz) + Db) * 1.00001f + 1e-6f * bg.h;   // margin: rounded float arithmetic on both sides
    r[0] = qx; r[1] = qy; r[2] = qz; r[3] = qx; r[4] = qy; r[5] = qz; r[6] = B;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int c = 0; c < 3; c++) { r[c] = fminf(r[c], __shfl_xor_sync(0xffffffffu, r[c], o)); r[3 + c] = fmaxf(r[3 + c], __shfl_xor_sync(0xffffffffu, r[3 + c], o)); }
    r[6] = fmaxf(r[6], __shfl_xor_sync(0xffffffffu, r[6], o));
  }
  if (lane == 0)
#pragma unroll
    for (int c = 0; c < 7; c++) s_red[c][warp] = r[c];
  if (tid == 0) s_cnt = 0;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < KT_TILE / 32; w++) {
#pragma unroll
    for (int c = 0; c < 3; c++) { r[c] = fminf(r[c], s_red[c][w]); r[3 + c] = fmaxf(r[3 + c], s_red[3 + c][w]); }
    r[6] = fmaxf(r[6], s_red[6][w]);
  }
  const float Bmax = r[6];
  bool fits = Bmax < 1e30f;
  int lo[3], hi[3];
#pragma unroll
  for (int c = 0; c < 3; c++) {
    lo[c] = bucket_coord(r[c] - Bmax, bg.min[c], bg.inv_h, bg.dim[c]);
    hi[c] = bucket_coord(r[3 + c] + Bmax, bg.min[c], bg.inv_h, bg.dim[c]);
  }
  const int ny = hi[1] - lo[1] + 1, nz = hi[2] - lo[2] + 1;
  const int nb = (hi[0] - lo[0] + 1) * ny * nz;
  if (nb > 8192) fits = false;
  const float mx = 0.5f * (r[0] + r[3]), my = 0.5f * (r[1] + r[4]), mz = 0.5f * (r[2] + r[5]);   // tile centre
  if (fits) {
    for (int t = tid; t < nb; t += KT_TILE) {
      const int x = lo[0] + t / (ny * nz), y = lo[1] + (t / nz) % ny, z = lo[2] + t % nz;
      const int c = (x * bg.dim[1] + y) * bg.dim[2] + z;
      const int b = cell_start[c], e = cell_start[c + 1];
      if (e > b) {
        const int p0 = atomicAdd(&s_cnt, e - b);
        if (p0 + (e - b) <= KT_CAP)
          for (int u = b; u < e; u++) {
            const float4 n = __ldg(sorted + u);
            const float ax = n.x - mx, ay = n.y - my, az = n.z - mz;
            s_cand[p0 + u - b] = n; s_key[p0 + u - b] = ax * ax + ay * ay + az * az;
          }
      }
    }
  }
  __syncthreads();
  const int cnt = s_cnt;
  const bool staged = fits && cnt <= KT_CAP;
  if (staged && cnt > 1) {
    // candidates in ascending distance from the tile centre: a query's first k+1 candidates are then (nearly) its nearest,
    // the running (k+1)-th distance is tight at once and almost every later candidate fails the first comparison.
    // The order is only a speed-up: without ties the result is the ascending-distance list whatever the scan order.
    int np = 2; while (np < cnt) np <<= 1;
    for (int t = cnt + tid; t < np; t += KT_TILE) s_key[t] = FLT_MAX;
    __syncthreads();
    for (int kk = 2; kk <= np; kk <<= 1)
      for (int j = kk >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np; i += KT_TILE) {
          const int p = i ^ j;
          if (p > i) {
            const bool up = (i & kk) == 0;
            const float a = s_key[i], b = s_key[p];
            if ((a > b) == up) { s_key[i] = b; s_key[p] = a; const float4 t4 = s_cand[i]; s_cand[i] = s_cand[p]; s_cand[p] = t4; }
          }
        }
        __syncthreads();
      }
  }
  if (!live) return;
  float d[KQ]; int id[KQ]; bool tie = false;
  if (!staged) {
    knn_ring_search<KQ>(qx, qy, qz, kq, bg, cell_start, sorted, d, id, tie);
  } else {
#pragma unroll
    for (int j = 0; j < KQ; j++) { d[j] = FLT_MAX; id[j] = -1; }
    float dw = B;    // B >= this query's (k+1)-th distance: farther nodes can neither enter nor tie; then the running (k+1)-th distance
    bool full = false;
    for (int t = 0; t < cnt; t++) {
      const float4 n = s_cand[t];
      const float dd = knn_dist(qx, qy, qz, n.x, n.y, n.z);
      if (dd > dw) continue;
      if (full && dd == dw) { tie = true; continue; }
      float cd = dd; int ci = __float_as_int(n.w);
#pragma unroll
      for (int j = 0; j < KQ; j++) {
        if (EXACT_KQ || j < kq) {
          if (cd == d[j] && cd != FLT_MAX) tie = true;   // empty slots hold FLT_MAX and must not count as ties
          if (cd < d[j]) { const float td = d[j]; const int ti = id[j]; d[j] = cd; id[j] = ci; cd = td; ci = ti; }
        }
      }
      float last = d[KQ - 1];
      if (!EXACT_KQ) {
#pragma unroll
        for (int j = 0; j < KQ; j++) if (j == kq - 1) last = d[j];
      }
      if (last != FLT_MAX) { dw = last; full = true; }
    }
  }
  if (tie) {
    float dw = d[KQ - 1];
    if (!EXACT_KQ) {
#pragma unroll
      for (int j = 0; j < KQ; j++) if (j == kq - 1) dw = d[j];
    }
    const int slot = atomicAdd(slow_count, 1);
    slow_list[slot] = q;
    slow_thr[slot] = dw;
    return;
  }
  if (EXACT_KQ && KQ < KNN_MAX + 1) {   // knn_store works on the full-width arrays
    float df[KNN_MAX + 1]; int idf[KNN_MAX + 1];
#pragma unroll
    for (int j = 0; j < KNN_MAX + 1; j++) { df[j] = j < KQ ? d[j < KQ ? j : 0] : FLT_MAX; idf[j] = j < KQ ? id[j < KQ ? j : 0] : -1; }
    knn_store<KNN_MAX + 1>(out, q, k, df, idf);
  } else {
    knn_store<KQ>(out, q, k, d, id);
  }
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

// FPS over cell-ordered points with the density grid as the pruning structure (see k_fps_pruned).  prefix = inclusive
// per-cell point counts (G^3), min3 / step / G = the grid of arap_grid_build.  scratch: N floats + 64 KB + G^3 * 8 bytes.
extern "C" size_t arapk_fps_grid_scratch_bytes(long long N, int G) {
  return (((size_t)N * 4 + 255) / 256) * 256 + 65536 + (size_t)G * G * G * sizeof(FpsBest) + 256;
}
extern "C" int arapk_fps_grid(const float* pos, long long N, int node_num, const int* cell_prefix, const float* min3_host, float step,
                              int G, int* out_idx_dev, void* scratch, size_t scratch_bytes, int* out_count_host, cudaStream_t st) {
  const int m = (int)std::min<long long>(node_num, N);
  const int M0 = 128;   // full passes before the pruned loop takes over (the first steps touch most of the cloud)
  if (scratch_bytes < arapk_fps_grid_scratch_bytes(N, G)) { set_error("fps_grid: scratch too small"); return ARAP_ERR_INVALID; }
  int cnt = 0;
  int rc = arapk_fps(pos, N, std::min(m, M0), out_idx_dev, scratch, (((size_t)N * 4 + 255) / 256) * 256 + 65536, &cnt, st);
  if (rc) return rc;
  if (out_count_host) *out_count_host = m;
  if (m <= M0) return ARAP_OK;
  float* dist = (float*)scratch;
  FpsBest* cm = (FpsBest*)((char*)scratch + (((size_t)N * 4 + 255) / 256) * 256 + 65536);
  const long long ncell = (long long)G * G * G;
  k_fps_cellmax<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(ncell, cell_prefix, dist, cm);
  ARAP_KERNEL_CHECK();
  FpsGrid g; g.min[0] = min3_host[0]; g.min[1] = min3_host[1]; g.min[2] = min3_host[2]; g.step = step; g.G = G;
  g.SG = 1; while (g.SG * 16 < G) g.SG <<= 1;     // power of two with at most 16 supercells per axis (4096 in shared memory)
  g.ns = (G + g.SG - 1) / g.SG;
  k_fps_pruned<<<1, 1024, 0, st>>>(pos, g, cell_prefix, dist, cm, M0, m, out_idx_dev);
  ARAP_KERNEL_CHECK();
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
  return (max_cells * 4 + 16) * sizeof(int) + (size_t)M * sizeof(int) + (size_t)M * sizeof(float4) + 4096;
}

struct ArapKnnIndex {
  BucketGrid bg; int ncell; int M;
  int *cell_cnt, *cell_start, *fill, *node_cell; float4* sorted; float* minmax; int* counters; float* bucket_bound;
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
  ix->bucket_bound = (float*)p; p += max_cells * sizeof(float);
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
  k_bucket_bounds<<<(ix->ncell + 127) / 128, 128, 0, st>>>(ix->ncell, std::min(M, KNN_MAX + 1), bg, ix->cell_start, ix->sorted, ix->bucket_bound);
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
  static int mode = -1;   // ARAP_KNN_TILE=0: the per-query bucket walk of the first version for every call (measurement aid)
  if (mode < 0) { const char* ev = getenv("ARAP_KNN_TILE"); mode = ev ? atoi(ev) : 1; }
  if (mode && ix->M >= KNN_MAX + 1) {
#define ARAP_KNN_TILE_ARGS queries_dev, Q, k, ix->bg, ix->cell_start, ix->sorted, ix->bucket_bound, out, ix->counters, slow_list, slow_thr
    if (k == 8) k_knn_tile<9, true><<<grid, KT_TILE, 0, st>>>(ARAP_KNN_TILE_ARGS);
    else if (k == 10) k_knn_tile<11, true><<<grid, KT_TILE, 0, st>>>(ARAP_KNN_TILE_ARGS);
    else if (k == 12) k_knn_tile<13, true><<<grid, KT_TILE, 0, st>>>(ARAP_KNN_TILE_ARGS);
    else k_knn_tile<KNN_MAX + 1, false><<<grid, KT_TILE, 0, st>>>(ARAP_KNN_TILE_ARGS);
#undef ARAP_KNN_TILE_ARGS
  }
  else
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
