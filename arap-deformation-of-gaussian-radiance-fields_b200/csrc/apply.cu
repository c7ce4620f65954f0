// Stage (d) apply pass + the per-step sample passes (a9, a10).
//
//   lbs_points      DeformGraph::predict_mesh / predict_samples  (Deform.hpp:230-268)
//   end_points      GaussianView::GetEndPoints                   (GaussianView.cpp:4643-4668)
//   fit_gaussians   GaussianView::UpdateAsSixPointsWithdrawBad   (GaussianView.cpp:3081-3166)
//   node_quats      GaussianView::FastUpdateSamplesSH host part  (GaussianView.cpp:3169-3186)
//   rotate_sample_shs  RotateSHs kernel                          (cudakdtree.cu:201-222)
//   static_flags    CheckStaticSamples / CheckMovedGaussians     (GaussianView.cpp:2024-2109)
//   lbs_tiles / lbs_build_tiles   the same LBS through per-tile staged node records (ours)
//   replay_shs      the SH update of fit_gaussians repeated by a multi-GPU receiver (ours)
//
// HBM layout.  The rasteriser-facing SoA keeps the reference layout
// (pos N x 3, rot N x 4 wxyz, scale N x 3, opacity N, shs N x 48 — the
// Rasterizer::forward argument layout, GaussianView.cpp:1106-1112).  Internal
// tables are ours: skinning rows are stored in blocks of 32 rows,
//   idx[(blk*K + j)*32 + lane]  (uint16)     w[(blk*K + j)*32 + lane]  (double)
// so a warp reading neighbour j of 32 consecutive rows issues one 64 B and one
// 256 B fully-coalesced request; the staged LBS kernel reads one-byte slot numbers
//   slots[(blk*3 + m)*32 + lane]  (uint32 = 4 slots)
// into a per-tile list of distinct nodes instead of the uint16 ids.
#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include "device_math.cuh"
#include "sh_fast.cuh"
#include "kernels.h"

namespace arapgs {

// ------------------------------------------------------------------ node xf
__global__ void k_node_xf(int M, const double* __restrict__ rot, const double* __restrict__ trans,
                          const float* __restrict__ node_pos, NodeXf* __restrict__ out, NodeXf32* __restrict__ out32) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  NodeXf x;
#pragma unroll
  for (int t = 0; t < 9; t++) x.A[t] = rot[9 * i + t];
#pragma unroll
  for (int t = 0; t < 3; t++) {
    float g = node_pos[3 * i + t];
    x.g[t] = g;
    x.c[t] = trans[3 * i + t] + (double)g;
  }
  x.pad = 0.f;
  out[i] = x;
  if (out32) {   // tolerance-mode record: A - I (row-major), t, g as floats
    NodeXf32 y;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) y.dA[3 * r + c] = (float)(x.A[c * 3 + r] - (r == c ? 1.0 : 0.0));
#pragma unroll
    for (int t = 0; t < 3; t++) { y.t[t] = (float)trans[3 * i + t]; y.g[t] = x.g[t]; }
    y.pad = 0.f;
    out32[i] = y;
  }
}

// ------------------------------------------------------------------ LBS
// One thread per point.  Double products, float accumulator rounded after each
// neighbour — exactly the reference's `Pos output += double_expr` sequence.
// `in` and `out` may alias (the session skins end points / samples / mesh points in place: every element is read and
// written by its own thread only), so neither is __restrict__.
//
// round_to_float: RN-even rounding of a double to float precision WITHOUT leaving the FP64 pipe.
// M = 1.5 * 2^(e+29) (e = exponent of s) puts the unit in the last place of s + M at 2^(e-23), the float
// ulp of s; the hardware add rounds to nearest-even there and the subtraction is exact.  Identical to
// (double)(float)s for every s in the normal float range and for 0 (float denormals, |s| < 2^-126, excluded);
// it replaces two XU-pipe conversions (16 lanes/clk/SM) by two DADD (64 lanes/clk/SM) and two integer ops.
__device__ __forceinline__ double round_to_float(double s) {
  const int hi = __double2hiint(s);
  const double M = __hiloint2double((hi & 0x7ff00000) + 0x01d80000, 0);
  return (s + M) - M;
}

struct LbsAcc {  // the float accumulator, kept either as float or as a float-valued double
  double d0, d1, d2;
};

template <bool MAGIC>
__device__ __forceinline__ void lbs_neighbour(float c0, float c1, float c2, double w, const double2 a01, const double2 a23,
                                              const double2 a45, const double2 a67, const double2 a8c0, const double2 c12,
                                              const float4 g, LbsAcc& o) {
  const double t0 = (double)(c0 - g.x), t1 = (double)(c1 - g.y), t2 = (double)(c2 - g.z);
  // A column-major: row r = (A[r], A[r+3], A[r+6])
  const double e0 = fma(a67.x, t2, fma(a23.y, t1, fma(a01.x, t0, a8c0.y)));
  const double e1 = fma(a67.y, t2, fma(a45.x, t1, fma(a01.y, t0, c12.x)));
  const double e2 = fma(a8c0.x, t2, fma(a45.y, t1, fma(a23.x, t0, c12.y)));
  if (MAGIC) {
    o.d0 = round_to_float(fma(w, e0, o.d0));
    o.d1 = round_to_float(fma(w, e1, o.d1));
    o.d2 = round_to_float(fma(w, e2, o.d2));
  } else {
    o.d0 = (double)(float)fma(w, e0, o.d0);
    o.d1 = (double)(float)fma(w, e1, o.d1);
    o.d2 = (double)(float)fma(w, e2, o.d2);
  }
}

template <int K>
__device__ __forceinline__ void lbs_row_global(const float* in, float* out, long long i, int k,
                                               const uint16_t* __restrict__ ridx, const double* __restrict__ rw,
                                               const NodeXf* __restrict__ nodes) {
  const float c0 = in[3 * i], c1 = in[3 * i + 1], c2 = in[3 * i + 2];
  LbsAcc o{0.0, 0.0, 0.0};
  const long long base = (i >> 5) * (long long)(k * 32) + (i & 31);
#pragma unroll
  for (int j = 0; j < (K > 0 ? K : KNN_MAX); j++) {
    if (K == 0 && j >= k) break;
    const uint16_t nd = ridx[base + j * 32];
    const double w = rw[base + j * 32];
    const double2* n = reinterpret_cast<const double2*>(nodes + nd);
    const double2 a01 = __ldg(n + 0), a23 = __ldg(n + 1), a45 = __ldg(n + 2), a67 = __ldg(n + 3);
    const double2 a8c0 = __ldg(n + 4), c12 = __ldg(n + 5);
    const float4 g = __ldg(reinterpret_cast<const float4*>(n + 6));
    lbs_neighbour<false>(c0, c1, c2, w, a01, a23, a45, a67, a8c0, c12, g, o);
  }
  out[3 * i] = (float)o.d0; out[3 * i + 1] = (float)o.d1; out[3 * i + 2] = (float)o.d2;
}

template <int K>
__global__ void __launch_bounds__(256)
k_lbs_points(const float* in, float* out, long long P, int k_rt,
             const uint16_t* __restrict__ ridx, const double* __restrict__ rw,
             const NodeXf* __restrict__ nodes, const uint8_t* __restrict__ skip, int group) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  if (skip && skip[i / group]) return;
  lbs_row_global<K>(in, out, i, K > 0 ? K : k_rt, ridx, rw, nodes);
}

// ------------------------------------------------------------------ LBS, node records staged per tile
// The skinning tables never change after set-up, so the set of DISTINCT nodes a tile of LT_ROWS consecutive rows
// touches is precomputed (k_build_tiles): rows are in cell order, hence a tile sees a few dozen nodes, not
// LT_ROWS * k.  Per step a CTA copies those records once into shared memory (coalesced 16-byte loads, one L2 read
// per record per tile) and every (row, neighbour) pair reads its record from there through a one-byte slot number.
// That replaces ten 112-byte L1 gathers per row (the global-memory kernel sits at 83-95 % l1tex throughput with
// DRAM at 25-34 %) by shared-memory reads that mostly broadcast, and the id table shrinks from 2 to 1.2 bytes per pair.
// Staged records sit at a 112-byte pitch: bank group of 16-byte part p of slot s is (p - s) mod 8.
// A tile with more than `cap` distinct nodes has cnt = 0 and takes the global-memory path (same arithmetic).
constexpr int LT_ROWS = 128;
constexpr int LT_CAP = 96;
constexpr int LT_PITCH = 7;   // 16-byte units per staged record (6 used): bank group of part p of slot s = (p - s) mod 8
constexpr int LT_WORDS = 3;   // slot words per row: 12 one-byte slots

template <int K, bool MAGIC>
__global__ void __launch_bounds__(LT_ROWS, 8)   // 64 registers: 8 CTAs per SM (72 registers / 7 CTAs measured 3 % slower on the samples)
k_lbs_tiles(const float* in, float* out, long long P, int k_rt, const uint32_t* __restrict__ slots,
            const double* __restrict__ rw, const uint16_t* __restrict__ ridx, const uint16_t* __restrict__ tile_cnt,
            const uint16_t* __restrict__ tile_nodes, const NodeXf* __restrict__ nodes, const uint8_t* __restrict__ skip,
            int group) {
  __shared__ double2 s_rec[LT_CAP * LT_PITCH];   // A (9 doubles) and c (3 doubles) of every staged node
  __shared__ float s_g[3 * LT_CAP];               // node positions, one array per component
  const int tid = threadIdx.x;
  const long long tile = blockIdx.x;
  const long long i = tile * LT_ROWS + tid;
  const int k = K > 0 ? K : k_rt;
  const int cnt = tile_cnt[tile];
  const bool live = i < P && !(skip && skip[i / group]);
  if (cnt == 0) {  // more distinct nodes than the staging area holds (uniform per CTA)
    if (live) lbs_row_global<K>(in, out, i, k, ridx, rw, nodes);
    return;
  }
  const uint16_t* tn = tile_nodes + tile * LT_CAP;
  for (int v = tid; v < cnt * 7; v += LT_ROWS) {
    const int r = v / 7, part = v - r * 7;
    const double2 x = __ldg(reinterpret_cast<const double2*>(nodes + tn[r]) + part);
    if (part < 6) s_rec[r * LT_PITCH + part] = x;
    else {
      const float4 g = *reinterpret_cast<const float4*>(&x);
      s_g[r] = g.x; s_g[LT_CAP + r] = g.y; s_g[2 * LT_CAP + r] = g.z;
    }
  }
  // this row's streamed operands, all in flight before the barrier
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;
  uint32_t sw[LT_WORDS] = {0u, 0u, 0u};
  double w[K > 0 ? K : KNN_MAX];
  if (live) {
    c0 = in[3 * i]; c1 = in[3 * i + 1]; c2 = in[3 * i + 2];
    const long long sb = (i >> 5) * (long long)(LT_WORDS * 32) + (i & 31);
#pragma unroll
    for (int m = 0; m < LT_WORDS; m++) sw[m] = slots[sb + m * 32];
    const long long base = (i >> 5) * (long long)(k * 32) + (i & 31);
#pragma unroll
    for (int j = 0; j < (K > 0 ? K : KNN_MAX); j++) w[j] = (K > 0 || j < k) ? rw[base + j * 32] : 0.0;
  }
  __syncthreads();
  if (!live) return;
  LbsAcc o{0.0, 0.0, 0.0};
#pragma unroll
  for (int j = 0; j < (K > 0 ? K : KNN_MAX); j++) {
    if (K == 0 && j >= k) break;
    const unsigned slot = (sw[j >> 2] >> ((j & 3) * 8)) & 0xffu;
    const double2* n = s_rec + slot * LT_PITCH;
    const double2 a01 = n[0], a23 = n[1], a45 = n[2], a67 = n[3], a8c0 = n[4], c12 = n[5];
    // node position from three float arrays: 3 one-wavefront LDS.32 (conflict-free: bank = slot) instead of a fourth
    // wavefront for 4 bytes of padding in an LDS.128 — the kernel is bound by exactly these wavefronts
    lbs_neighbour<MAGIC>(c0, c1, c2, w[j], a01, a23, a45, a67, a8c0, c12, make_float4(s_g[slot], s_g[LT_CAP + slot], s_g[2 * LT_CAP + slot], 0.f), o);
  }
  out[3 * i] = (float)o.d0; out[3 * i + 1] = (float)o.d1; out[3 * i + 2] = (float)o.d2;
}

// Set-up: distinct node list + one-byte slots of every tile.  One CTA per tile: a 64 Kbit bitmap of the node ids
// the tile touches, ranked by a popcount prefix.
// Slot numbers are then chosen to avoid shared-memory bank conflicts in k_lbs_tiles: an LDS.128 is served one
// quarter-warp (8 consecutive rows) at a time, and two records conflict there iff their slots are congruent mod 8
// (112-byte pitch).  Nodes that are read by the same quarter-warp for the same neighbour position j are joined in a
// graph weighted by how often that happens, every node takes the colour (of 8) that costs the fewest conflicts with the
// nodes placed before it, and slot = colour + 8 * (index within the colour).  Measured on the 6M-Gaussian workload: 5.8 -> ~4 wavefronts per LDS.128 for the end-point rows.
// tile_cnt = number of slots to stage (highest slot + 1; unused slots repeat a node of the tile), 0 = too many nodes.
__global__ void __launch_bounds__(LT_ROWS)
k_build_tiles(long long rows, int k, const uint16_t* __restrict__ ridx, uint32_t* __restrict__ slots,
              uint16_t* __restrict__ tile_cnt, uint16_t* __restrict__ tile_nodes) {
  __shared__ uint32_t bm[2048];
  __shared__ uint16_t pre[2048];
  __shared__ int wsum[LT_ROWS / 32];
  __shared__ uint16_t cow[LT_CAP][LT_CAP];   // co-occurrence counts: how many (quarter-warp, j) reads see both nodes
  __shared__ uint16_t node_of[LT_CAP];
  __shared__ uint8_t slot_of[LT_CAP];
  __shared__ int s_nslots;
  const int tid = threadIdx.x;
  const long long tile = blockIdx.x;
  const long long i = tile * LT_ROWS + tid;
  for (int v = tid; v < 2048; v += LT_ROWS) bm[v] = 0u;
  for (int v = tid; v < LT_CAP * LT_CAP / 2; v += LT_ROWS) reinterpret_cast<uint32_t*>(&cow[0][0])[v] = 0u;
  __syncthreads();
  const long long base = (i >> 5) * (long long)(k * 32) + (i & 31);
  if (i < rows)
    for (int j = 0; j < k; j++) { const unsigned n = ridx[base + j * 32]; atomicOr(&bm[n >> 5], 1u << (n & 31)); }
  __syncthreads();
  constexpr int WPT = 2048 / LT_ROWS;  // bitmap words per thread
  int c = 0;
#pragma unroll
  for (int v = 0; v < WPT; v++) c += __popc(bm[tid * WPT + v]);
  int inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if ((tid & 31) >= d) inc += t; }
  if ((tid & 31) == 31) wsum[tid >> 5] = inc;
  __syncthreads();
  int off = inc - c, total = 0;
#pragma unroll
  for (int v = 0; v < LT_ROWS / 32; v++) { if (v < (tid >> 5)) off += wsum[v]; total += wsum[v]; }
  const bool ok = total <= LT_CAP;
  int run = off;
#pragma unroll
  for (int v = 0; v < WPT; v++) {
    const int wd = tid * WPT + v;
    pre[wd] = (uint16_t)run;
    uint32_t b = bm[wd];
    while (b) {
      const int bit = __ffs(b) - 1; b &= b - 1;
      if (ok) node_of[run] = (uint16_t)(wd * 32 + bit);
      run++;
    }
  }
  __syncthreads();
  if (!ok) {   // uniform per CTA
    if (tid == 0) tile_cnt[tile] = 0;
    if (i < rows) {
      const long long sb = (i >> 5) * (long long)(LT_WORDS * 32) + (i & 31);
      for (int m = 0; m < LT_WORDS; m++) slots[sb + m * 32] = 0u;
    }
    return;
  }
  // ranks of this row's neighbours (ascending node id), and the conflict graph of the tile
  int rk[KNN_MAX];
  for (int j = 0; j < KNN_MAX; j++) {
    rk[j] = -1;
    if (i < rows && j < k) { const unsigned n = ridx[base + j * 32]; rk[j] = pre[n >> 5] + __popc(bm[n >> 5] & ((1u << (n & 31)) - 1u)); }
  }
  for (int j = 0; j < k; j++)
    for (int d = 1; d < 8; d++) {
      const int other = __shfl_xor_sync(0xffffffffu, rk[j], d);
      // 16-bit counters updated through their 32-bit word (a tile has 128 * 12 * 7 < 65536 increments per counter)
      if (rk[j] >= 0 && other >= 0 && other != rk[j])
        atomicAdd(reinterpret_cast<uint32_t*>(&cow[rk[j]][other & ~1]), (other & 1) ? 0x10000u : 1u);
    }
  __syncthreads();
  if (tid == 0) {
    int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int maxload = min(LT_CAP / 8, (total + 7) / 8 + 1);   // keeps the slot range (= records staged per step) tight
    uint8_t colour[LT_CAP];
    int hi = 0;
    for (int v = 0; v < total; v++) {
      unsigned cost[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // conflicts this node would have with the nodes already placed
      for (int u = 0; u < v; u++) cost[colour[u]] += cow[v][u];
      int best = -1;
      for (int cc = 0; cc < 8; cc++)
        if (load[cc] < maxload && (best < 0 || cost[cc] < cost[best] || (cost[cc] == cost[best] && load[cc] < load[best]))) best = cc;
      colour[v] = (uint8_t)best;
      load[best]++;
    }
    // local search: move a node to a cheaper colour while that lowers its conflict count (three sweeps)
    for (int sweep = 0; sweep < 3; sweep++) {
      int moved = 0;
      for (int v = 0; v < total; v++) {
        unsigned cost[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int u = 0; u < total; u++) cost[colour[u]] += cow[v][u];   // cow[v][v] = 0
        int best = colour[v];
        for (int cc = 0; cc < 8; cc++)
          if (cc != colour[v] && load[cc] < maxload && cost[cc] < cost[best]) best = cc;
        if (best != colour[v]) { load[colour[v]]--; load[best]++; colour[v] = (uint8_t)best; moved++; }
      }
      if (!moved) break;
    }
    int used[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int v = 0; v < total; v++) {
      const int sl = colour[v] + 8 * used[colour[v]]++;
      slot_of[v] = (uint8_t)sl;
      hi = sl > hi ? sl : hi;
    }
    s_nslots = hi + 1;
    tile_cnt[tile] = (uint16_t)(hi + 1);
  }
  __syncthreads();
  const int nslots = s_nslots;
  for (int v = tid; v < nslots; v += LT_ROWS) tile_nodes[tile * LT_CAP + v] = node_of[0];   // filler: any node of the tile
  __syncthreads();
  for (int v = tid; v < total; v += LT_ROWS) tile_nodes[tile * LT_CAP + slot_of[v]] = node_of[v];
  if (i < rows) {
    uint32_t sw[LT_WORDS] = {0u, 0u, 0u};
    for (int j = 0; j < k; j++) sw[j >> 2] |= (uint32_t)slot_of[rk[j]] << ((j & 3) * 8);
    const long long sb = (i >> 5) * (long long)(LT_WORDS * 32) + (i & 31);
    for (int m = 0; m < LT_WORDS; m++) slots[sb + m * 32] = sw[m];
  }
}

// ------------------------------------------------------------------ end points
__global__ void k_end_points(long long N, const float* __restrict__ pos, const float* __restrict__ rot,
                             const float* __restrict__ scale, float* __restrict__ ends) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= N) return;
  const float4 r4 = ldg4(rot + 4 * g);
  Quat q = quat_normalized(Quat{r4.x, r4.y, r4.z, r4.w});
  float R[3][3]; quat_to_matrix(q, R);
  const float p0 = pos[3 * g], p1 = pos[3 * g + 1], p2 = pos[3 * g + 2];
  float* e = ends + g * 18;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float len = (scale[3 * g + i] + 1e-3f) * 2.0f;  // axis_padding, end_coeff (GaussianView.hpp:116-117)
    const float v0 = R[0][i] * len, v1 = R[1][i] * len, v2 = R[2][i] * len;
    e[6 * i + 0] = p0 + v0; e[6 * i + 1] = p1 + v1; e[6 * i + 2] = p2 + v2;
    e[6 * i + 3] = p0 - v0; e[6 * i + 4] = p1 - v1; e[6 * i + 5] = p2 - v2;
  }
}

// ------------------------------------------------------------------ fit
// Polar factor of a 3x3 (double) by determinant-scaled Newton iteration
// X <- (mu X + X^-T / mu)/2.  The polar decomposition is unique, so this
// reproduces the reference's double JacobiSVD U V^T (helper.cpp:429-440) to
// ~1e-15.  K = diag(R^T M) = diag(V Sigma V^T).
__device__ __forceinline__ void polar_newton(const double (&M)[3][3], double (&X)[3][3]) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) X[r][c] = M[r][c];
  for (int it = 0; it < 14; it++) {
    double C[3][3];  // cofactors: X^-T = C / det
    C[0][0] = fma(X[1][1], X[2][2], -(X[1][2] * X[2][1]));
    C[0][1] = fma(X[1][2], X[2][0], -(X[1][0] * X[2][2]));
    C[0][2] = fma(X[1][0], X[2][1], -(X[1][1] * X[2][0]));
    C[1][0] = fma(X[0][2], X[2][1], -(X[0][1] * X[2][2]));
    C[1][1] = fma(X[0][0], X[2][2], -(X[0][2] * X[2][0]));
    C[1][2] = fma(X[0][1], X[2][0], -(X[0][0] * X[2][1]));
    C[2][0] = fma(X[0][1], X[1][2], -(X[0][2] * X[1][1]));
    C[2][1] = fma(X[0][2], X[1][0], -(X[0][0] * X[1][2]));
    C[2][2] = fma(X[0][0], X[1][1], -(X[0][1] * X[1][0]));
    const double det = fma(X[0][0], C[0][0], fma(X[0][1], C[0][1], X[0][2] * C[0][2]));
    // scaling mu = |det|^(-1/3) while far from orthogonal, 1 near convergence
    const float adet = fabsf((float)det);
    const double mu = (adet > 1.25f || adet < 0.8f) ? (double)rcbrtf(adet) : 1.0;
    const double a = 0.5 * mu, b = 0.5 / (mu * det);
    // After any step every singular value is (a + 1/a)/2 >= 1, so | |det| - 1 | bounds max(sigma - 1); the error
    // squares per step ((sigma - 1)^2 / 2 sigma): below 2e-8 this step is the last one (result error ~2e-16).
    // Replaces a max |X_new - X| test that cost more FP64 instructions than the update itself and one extra iteration.
    const bool last = it > 0 && fabs(fabs(det) - 1.0) < 2e-8;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) X[r][c] = fma(a, X[r][c], b * C[r][c]);
    if (last) break;
  }
}

constexpr int FIT_TILE = 128;
constexpr int FIT_PITCH4 = 13;  // float4 chunks per staged SH row
constexpr int END_PITCH = 19;   // odd pitch: conflict-free per-thread rows


__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 128-Gaussian tile per CTA.  The SH tile (24.5 KB) is fetched with 16-byte cp.async into rows at a 13-chunk pitch
// (a thread's 12 LDS.128 / STS.128 on its own row are conflict-free, as in k_rotate_sample_shs) and lands while the pose
// (centre, polar factor, quaternion, scale) is computed from the end points; waiting for it in registers before doing
// anything else was 36 % of the kernel's stall samples.  The rotation itself runs on a register copy of the row.
// (A first version copied the tile with 4-byte cp.async into odd-pitch rows for scalar access: that copy loop alone was
// 25 % of the kernel's instructions.)
__global__ void __launch_bounds__(FIT_TILE, 4)   // 128 registers, no spills (5 / 6 CTAs per SM spill and measured no faster)
k_fit_gaussians(long long N, const float* __restrict__ ends, const float* __restrict__ scale_backup,
                const uint8_t* __restrict__ is_static, float* __restrict__ pos, float* __restrict__ rot,
                float* __restrict__ scale, float* __restrict__ shs) {
  extern __shared__ float4 s_sh4[];                                        // FIT_TILE x FIT_PITCH4 float4
  float* s_end = reinterpret_cast<float*>(s_sh4 + FIT_TILE * FIT_PITCH4);  // FIT_TILE x END_PITCH
  __shared__ uint8_t s_static[FIT_TILE];

  const long long g0 = (long long)blockIdx.x * FIT_TILE;
  const int tid = threadIdx.x;
  const int rows = (int)min((long long)FIT_TILE, N - g0);
  s_static[tid] = (tid < rows) ? (is_static ? is_static[g0 + tid] : 0) : 1;
  // endpoint tile (18 floats per row, 8-byte granules) through registers into odd-pitch rows
  const float* gend = ends + g0 * 18;
  const int nend = rows * 18;
  for (int v = tid * 2; v < nend; v += FIT_TILE * 2) {
    const int r = v / 18, c = v - r * 18;
    const float2 x = __ldg(reinterpret_cast<const float2*>(gend + v));
    s_end[r * END_PITCH + c] = x.x; s_end[r * END_PITCH + c + 1] = x.y;
  }
  // per-Gaussian operands of the pose, requested BEFORE the bulk SH copy: memory responses return roughly in issue order
  // per SM, so a load queued behind the 24.5 KB tile waits for all of it (these two were 22 % of the stall samples)
  const long long g = g0 + tid;
  float4 o4 = make_float4(1.f, 0.f, 0.f, 0.f);
  float sb0 = 1.f, sb1 = 1.f, sb2 = 1.f;
  if (tid < rows) { o4 = ldg4(rot + 4 * g); sb0 = scale_backup[3 * g]; sb1 = scale_backup[3 * g + 1]; sb2 = scale_backup[3 * g + 2]; }
  __syncthreads();
  const float4* gsh = reinterpret_cast<const float4*>(shs + g0 * SH_FLOATS);
  const int cp_r0 = tid / 12, cp_c4 = tid - 12 * cp_r0;   // 120 threads x 13 passes of 10 rows: loop-invariant row / chunk
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) cp_async16(s_sh4 + r * FIT_PITCH4 + cp_c4, gsh + tid + 120 * t);
    }
  }
  cp_async_commit();

  const bool act = tid < rows && !s_static[tid];
  float Rs[3][3];
  if (act) {
    const float* e = s_end + tid * END_PITCH;
    const Quat oq{o4.x, o4.y, o4.z, o4.w};
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; r++)  // Eigen redux order: (p0+(p1+p2)) + (p3+(p4+p5))
      c[r] = ((e[r] + (e[3 + r] + e[6 + r])) + (e[9 + r] + (e[12 + r] + e[15 + r]))) / 6.0f;
    double Md[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float a = e[6 * i + r] - c[r], b = e[6 * i + 3 + r] - c[r];
        Md[r][i] = (double)(0.5f * a + (-0.5f) * b);
      }
    double Rd[3][3];
    polar_newton(Md, Rd);
    float Rf[3][3], K[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) Rf[i][j] = (float)Rd[i][j];
      K[i] = (float)fma(Rd[0][i], Md[0][i], fma(Rd[1][i], Md[1][i], Rd[2][i] * Md[2][i]));
    }
    const Quat q = quat_normalized(quat_from_matrix(Rf));
    *reinterpret_cast<float4*>(rot + 4 * g) = make_float4(q.w, q.x, q.y, q.z);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float s0 = i == 0 ? sb0 : i == 1 ? sb1 : sb2;
      scale[3 * g + i] = K[i] / ((s0 + 1e-3f) * 2.0f) * s0;
      pos[3 * g + i] = c[i];
    }
    const Quat rq = quat_normalized(quat_mul(q, quat_inverse(oq)));
    quat_to_matrix(rq, Rs);
  }
  cp_async_wait<0>();
  __syncthreads();
  if (act) {
    float4* row = s_sh4 + tid * FIT_PITCH4;
    float v[SH_FLOATS];
#pragma unroll
    for (int c = 0; c < 12; c++) { const float4 x = row[c]; v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w; }
    sh_rotate_flipped_fast(Rs, v);
#pragma unroll
    for (int c = 1; c < 12; c++) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    row[0].w = v[3];   // DC term (floats 0-2) is rotation invariant
  }
  __syncthreads();
  float* osh = shs + g0 * SH_FLOATS;
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) st_stream4(osh + (size_t)(tid + 120 * t) * 4, s_sh4[r * FIT_PITCH4 + cp_c4]);
    }
  }
}

// ------------------------------------------------------------------ tolerance-mode skinning (lbs_mode = 3)
// The bit-faithful kernels above reproduce the reference's `float += double` chain and are bound by the shared-memory
// return path: every (point, neighbour) pair needs a 108-byte fp64 record in registers (27 wavefronts per warp and
// neighbour).  north_star's bar for means / covariances is 1e-5 relative, not bit equality, so this mode evaluates the
// same skinning  p' = sum_j w_j (A_j (p - g_j) + g_j + t_j)  as
//     p' = p + sum_j w_j ((A_j - I)(p - c) + t_j - (A_j - I)(g_j - c))                   (sum_j w_j = 1)
// with c a reference point next to p: every float operand is a small displacement (|A - I| ~ 1e-2, |p - c| ~ 1e-1,
// |t| ~ 1e-3), so float rounding of the products is ~1e-10 absolute — 2-3 orders below the ulp of p — and the result is
// the correctly rounded exact skinning in all but a few per cent of the coordinates.  It differs from the reference's
// chain by the chain's own rounding noise (~1 ulp of p).
//
// The order of the neighbours no longer matters, which is what makes the kernels fast: neighbour sets of nearby points
// overlap almost completely (measured on the 6M / 16k-node workload: the 60 (end point, neighbour) pairs of a Gaussian
// touch 12.6 distinct nodes on average, max 20; the 320 pairs of 32 consecutive samples about as many), so the tables are
// re-organised at set-up into UNIONS with dense float weights (zero where a point does not have the node):
//   * samples: one union per 32-row block; every lane of the warp reads the same 48-byte record (shared-memory broadcast,
//     one wavefront per LDS.128) and its own weight — ~14 x 3 wavefronts per warp instead of 10 x 27;
//   * end points: one union per Gaussian, a thread skins its six end points against each record it loads.
// Per-step float node records NodeXf32 = [A - I | t | g] are written next to the fp64 ones by k_node_xf.
__device__ __forceinline__ void stage32(const NodeXf32* __restrict__ nd, float cx, float cy, float cz, float4& r0, float4& r1, float4& r2) {
  const float4* n = reinterpret_cast<const float4*>(nd);
  const float4 v0 = __ldg(n), v1 = __ldg(n + 1), v2 = __ldg(n + 2), v3 = __ldg(n + 3);
  const float g0 = v3.x - cx, g1 = v3.y - cy, g2 = v3.z - cz;
  r0 = make_float4(v0.x, v0.y, v0.z, v2.y - fmaf(v0.z, g2, fmaf(v0.y, g1, v0.x * g0)));
  r1 = make_float4(v0.w, v1.x, v1.y, v2.z - fmaf(v1.y, g2, fmaf(v1.x, g1, v0.w * g0)));
  r2 = make_float4(v1.z, v1.w, v2.x, v2.w - fmaf(v2.x, g2, fmaf(v1.w, g1, v1.z * g0)));
}

// ---- samples (and any other row family): one union per 32-row block ----------------------------------------------------
// Tables: boff[nblk + 1] (rows of the block's union), blist[row] (uint16 node id), bw[row * 32 + lane] (float weight of the
// row's node for the block's lane-th point, 0 if it is not among the point's k neighbours).
constexpr int SU_WARPS = 4;
constexpr int SU_CHUNK = 16;

__global__ void __launch_bounds__(SU_WARPS * 32, 6)
k_lbs_union32(const float* in, float* out, long long P, const int* __restrict__ boff, const uint16_t* __restrict__ blist,
              const float* __restrict__ bw, const NodeXf32* __restrict__ nodes, const uint8_t* __restrict__ skip, int group) {
  __shared__ float4 s_rec[SU_WARPS][SU_CHUNK * 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long blk = (long long)blockIdx.x * SU_WARPS + warp;
  const long long i = blk * 32 + lane;
  if (blk * 32 >= P) return;
  const bool live = i < P && !(skip && skip[i / group]);
  if (!__any_sync(0xffffffffu, live)) return;
  const int off = boff[blk], cnt = boff[blk + 1] - off;
  const float cx = in[3 * blk * 32], cy = in[3 * blk * 32 + 1], cz = in[3 * blk * 32 + 2];   // the block's reference point
  float p0 = 0.f, p1 = 0.f, p2 = 0.f;
  if (live) { p0 = in[3 * i]; p1 = in[3 * i + 1]; p2 = in[3 * i + 2]; }
  float4 T0 = make_float4(0.f, 0.f, 0.f, 0.f), T1 = T0, T2 = T0;
  float4* rec = s_rec[warp];
  const float* wrow = bw + (long long)off * 32 + lane;
  // Chunks of SU_CHUNK union rows (one chunk covers almost every block).  All weights of a chunk are requested at once, the
  // next chunk's before the current one is consumed: the kernel is a stream of independent 128-byte row reads and lives on
  // how many of them are in flight, not on occupancy.
  float wc[SU_CHUNK];
#pragma unroll
  for (int t = 0; t < SU_CHUNK; t++) wc[t] = t < cnt ? __ldg(wrow + (long long)t * 32) : 0.f;
  for (int base = 0; base < cnt; base += SU_CHUNK) {
    const int n = min(SU_CHUNK, cnt - base);
    __syncwarp();
    if (lane < n) {
      float4 r0, r1, r2;
      stage32(nodes + blist[off + base + lane], cx, cy, cz, r0, r1, r2);
      rec[3 * lane] = r0; rec[3 * lane + 1] = r1; rec[3 * lane + 2] = r2;
    }
    float wn[SU_CHUNK];
#pragma unroll
    for (int t = 0; t < SU_CHUNK; t++) wn[t] = base + SU_CHUNK + t < cnt ? __ldg(wrow + (long long)(base + SU_CHUNK + t) * 32) : 0.f;
    __syncwarp();
#pragma unroll
    for (int t = 0; t < SU_CHUNK; t++) {
      if (t < n) {
        const float w = wc[t];
        const float4 a = rec[3 * t], b = rec[3 * t + 1], c = rec[3 * t + 2];
        T0.x = fmaf(w, a.x, T0.x); T0.y = fmaf(w, a.y, T0.y); T0.z = fmaf(w, a.z, T0.z); T0.w = fmaf(w, a.w, T0.w);
        T1.x = fmaf(w, b.x, T1.x); T1.y = fmaf(w, b.y, T1.y); T1.z = fmaf(w, b.z, T1.z); T1.w = fmaf(w, b.w, T1.w);
        T2.x = fmaf(w, c.x, T2.x); T2.y = fmaf(w, c.y, T2.y); T2.z = fmaf(w, c.z, T2.z); T2.w = fmaf(w, c.w, T2.w);
      }
    }
#pragma unroll
    for (int t = 0; t < SU_CHUNK; t++) wc[t] = wn[t];
  }
  if (!live) return;
  const float q0 = p0 - cx, q1 = p1 - cy, q2 = p2 - cz;
  out[3 * i] = p0 + fmaf(T0.z, q2, fmaf(T0.y, q1, fmaf(T0.x, q0, T0.w)));
  out[3 * i + 1] = p1 + fmaf(T1.z, q2, fmaf(T1.y, q1, fmaf(T1.x, q0, T1.w)));
  out[3 * i + 2] = p2 + fmaf(T2.z, q2, fmaf(T2.y, q1, fmaf(T2.x, q0, T2.w)));
}

// Set-up of the block unions.  One warp per 32-row block; the distinct node ids of the block's 32 k neighbours are
// extracted in ascending order by repeated warp-minimum.  FILL = false: counts only (boff is their exclusive scan).
template <bool FILL>
__global__ void __launch_bounds__(128)
k_sunion_build(long long rows, int k, const uint16_t* __restrict__ ridx, const float* __restrict__ wf, const double* __restrict__ wd,
               int* __restrict__ bcnt, const int* __restrict__ boff, uint16_t* __restrict__ blist, float* __restrict__ bw) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long blk = (long long)blockIdx.x * 4 + warp;
  if (blk * 32 >= rows) return;
  const long long i = blk * 32 + lane;
  const long long base = blk * (long long)(k * 32) + lane;
  int id[KNN_MAX]; float w[KNN_MAX];
#pragma unroll
  for (int j = 0; j < KNN_MAX; j++) {
    const bool ok = j < k && i < rows;
    id[j] = ok ? (int)ridx[base + j * 32] : 0x7fffffff;
    w[j] = (FILL && ok) ? (wf ? wf[base + j * 32] : (float)wd[base + j * 32]) : 0.f;
  }
  int last = -1, cnt = 0;
  const long long off = FILL ? boff[blk] : 0;
  for (;;) {
    int cand = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < KNN_MAX; j++) if (id[j] > last && id[j] < cand) cand = id[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if (cand == 0x7fffffff) break;
    if (FILL) {
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < KNN_MAX; j++) if (id[j] == cand) mine += w[j];
      if (lane == 0) blist[off + cnt] = (uint16_t)cand;
      bw[(off + cnt) * 32 + lane] = mine;
    }
    last = cand; cnt++;
  }
  if (!FILL && lane == 0) bcnt[blk] = cnt;
}

// ---- end points: one union per Gaussian, fused with the six-point fit and the SH rotation ------------------------------
// Tables, blocked by 32 Gaussians (gblk = g / 32; rows = positions t of the union, padded to the block's longest union):
//   uoff[ngblk + 1]                      first row of a block
//   woff[ngblk + 1]                      first slot word of a block (ceil(rows / 4) words per block)
//   usw[(woff + t / 4) * 32 + lane]      four one-byte slots per word: index into the tile's staged node list
//   unode[(uoff + t) * 32 + lane]        the node id itself (read only by tiles whose node list overflows the staging area)
//   uw[((uoff + t) * 6 + e) * 32 + lane] float weight of the row's node for end point e (0 if absent; padding rows are all 0)
//   gtile_nodes[tile * GT_CAP + slot], gtile_cnt[tile]   distinct nodes of a 128-Gaussian tile (0 = more than GT_CAP)
constexpr int GT_CAP = 256;
constexpr int GU_MAX = 60;   // a Gaussian has 6 k <= 72 pairs; unions above GU_MAX rows fail the build (never seen: max 20 on the bench scene)

// distinct nodes of the Gaussian's 6 k pairs, ascending, with the six weights of each; returns the count
__device__ __forceinline__ int gaussian_union(long long g, int k, const uint16_t* __restrict__ ridx, const double* __restrict__ rw,
                                              uint16_t* un, float (*uwt)[6]) {
  int cnt = 0;
  for (int e = 0; e < 6; e++) {
    const long long row = g * 6 + e;
    const long long base = (row >> 5) * (long long)(k * 32) + (row & 31);
    for (int j = 0; j < k; j++) {
      const uint16_t n = ridx[base + j * 32];
      const float w = (float)rw[base + j * 32];
      int pos = 0;
      while (pos < cnt && un[pos] < n) pos++;
      if (pos == cnt || un[pos] != n) {
        if (cnt >= GU_MAX) return -1;
        for (int m = cnt; m > pos; m--) { un[m] = un[m - 1]; for (int c = 0; c < 6; c++) uwt[m][c] = uwt[m - 1][c]; }
        un[pos] = n; for (int c = 0; c < 6; c++) uwt[pos][c] = 0.f;
        cnt++;
      }
      uwt[pos][e] += w;
    }
  }
  return cnt;
}

template <bool FILL>
__global__ void __launch_bounds__(128)
k_gunion_build(long long N, int k, const uint16_t* __restrict__ ridx, const double* __restrict__ rw, int* __restrict__ ucnt,
               int* __restrict__ err, const int* __restrict__ uoff, const int* __restrict__ woff, uint32_t* __restrict__ usw,
               uint16_t* __restrict__ unode, float* __restrict__ uw, uint16_t* __restrict__ gtile_cnt, uint16_t* __restrict__ gtile_nodes) {
  __shared__ uint32_t bm[2048];
  __shared__ uint16_t pre[2048];
  __shared__ int wsum[4];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long tile = blockIdx.x;
  const long long g = tile * 128 + tid;
  const long long gblk = g >> 5;
  uint16_t un[GU_MAX]; float uwt[GU_MAX][6];
  int cnt = 0;
  if (g < N) cnt = gaussian_union(g, k, ridx, rw, un, uwt);
  if (cnt < 0) { atomicExch(err, 1); cnt = 0; }
  int mx = cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (!FILL) { if (lane == 0 && gblk * 32 < N) ucnt[gblk] = mx; return; }
  // tile slot list: bitmap of the node ids, ranked by a popcount prefix (as k_build_tiles)
  for (int v = tid; v < 2048; v += 128) bm[v] = 0u;
  __syncthreads();
  for (int t = 0; t < cnt; t++) atomicOr(&bm[un[t] >> 5], 1u << (un[t] & 31));
  __syncthreads();
  constexpr int WPT = 2048 / 128;
  int c = 0;
#pragma unroll
  for (int v = 0; v < WPT; v++) c += __popc(bm[tid * WPT + v]);
  int inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += t; }
  if (lane == 31) wsum[tid >> 5] = inc;
  __syncthreads();
  int off = inc - c, total = 0;
#pragma unroll
  for (int v = 0; v < 4; v++) { if (v < (tid >> 5)) off += wsum[v]; total += wsum[v]; }
  const bool ok = total <= GT_CAP;
  int run = off;
#pragma unroll
  for (int v = 0; v < WPT; v++) {
    const int wd = tid * WPT + v;
    pre[wd] = (uint16_t)run;
    uint32_t b = bm[wd];
    while (b) {
      const int bit = __ffs(b) - 1; b &= b - 1;
      if (ok) gtile_nodes[tile * GT_CAP + run] = (uint16_t)(wd * 32 + bit);
      run++;
    }
  }
  if (tid == 0) gtile_cnt[tile] = ok ? (uint16_t)total : (uint16_t)0;
  __syncthreads();
  if (gblk * 32 >= N) return;
  const long long r0 = uoff[gblk], w0 = woff[gblk];
  uint32_t word = 0u;
  for (int t = 0; t < mx; t++) {
    const bool have = t < cnt;
    const uint16_t n = have ? un[t] : (cnt ? un[0] : 0);
    const int slot = (ok && (have || cnt)) ? pre[n >> 5] + __popc(bm[n >> 5] & ((1u << (n & 31)) - 1u)) : 0;
    word |= (uint32_t)slot << ((t & 3) * 8);
    if ((t & 3) == 3 || t == mx - 1) { usw[(w0 + (t >> 2)) * 32 + lane] = word; word = 0u; }
    unode[(r0 + t) * 32 + lane] = n;
    for (int e = 0; e < 6; e++) uw[((r0 + t) * 6 + e) * 32 + lane] = have ? uwt[t][e] : 0.f;
  }
}

__global__ void k_words_of_rows(long long n, const int* __restrict__ rows, int* __restrict__ words) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) words[i] = (rows[i] + 3) >> 2;
}

// PUSH: the fused multi-GPU epilogue: when the tile's pose is final the CTA copies it — 40 bytes per Gaussian, as
// coalesced 16-byte stores — straight into every peer's gathered arrays over NVLink (peer pointers from cudaIpc, see
// arap_comm_set_mode), so the exchange of the deformed Gaussians rides inside the apply pass instead of following it as a
// collective.  Static Gaussians are copied too (their values have not changed: same bits on both sides).
// PUSH = 2: one store per value into the NVSwitch multicast mapping of the gathered arrays (multimem.st: the switch replicates it
// into every rank's copy, this rank's included) — NVLink egress of 40 bytes per Gaussian whatever the number of ranks.
__device__ __forceinline__ void multimem_st(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
template <int PUSH>
__global__ void __launch_bounds__(FIT_TILE, 4)
k_apply_union(long long N, const NodeXf32* __restrict__ nodes, const uint16_t* __restrict__ gtile_cnt,
              const uint16_t* __restrict__ gtile_nodes, const int* __restrict__ uoff, const int* __restrict__ woff,
              const uint32_t* __restrict__ usw, const uint16_t* __restrict__ unode, const float* __restrict__ uw, float* ends,
              const float* __restrict__ scale_backup, const uint8_t* __restrict__ is_static, float* __restrict__ pos,
              float* __restrict__ rot, float* __restrict__ scale, float* __restrict__ shs, ArapPeerPush pp) {
  extern __shared__ float4 s_sh4[];                                        // FIT_TILE x FIT_PITCH4 float4
  float* s_end = reinterpret_cast<float*>(s_sh4 + FIT_TILE * FIT_PITCH4);  // FIT_TILE x END_PITCH
  float4* s_rec = reinterpret_cast<float4*>(s_end + FIT_TILE * END_PITCH); // GT_CAP x 3 float4
  __shared__ uint8_t s_static[FIT_TILE];

  const long long tile = blockIdx.x;
  const long long g0 = tile * FIT_TILE;
  const int tid = threadIdx.x, lane = tid & 31;
  const int rows = (int)min((long long)FIT_TILE, N - g0);
  const long long g = g0 + tid;
  s_static[tid] = (tid < rows) ? (is_static ? is_static[g] : 0) : 1;
  const int tcnt = gtile_cnt[tile];
  float* gend = ends + g0 * 18;
  const float cx = gend[0], cy = gend[1], cz = gend[2];   // tile reference point: its first end point
  for (int v = tid; v < tcnt; v += FIT_TILE) {
    float4 r0, r1, r2;
    stage32(nodes + gtile_nodes[tile * GT_CAP + v], cx, cy, cz, r0, r1, r2);
    s_rec[3 * v] = r0; s_rec[3 * v + 1] = r1; s_rec[3 * v + 2] = r2;
  }
  const int nend = rows * 18;
  for (int v = tid * 2; v < nend; v += FIT_TILE * 2) {
    const int r = v / 18, c = v - r * 18;
    const float2 x = *reinterpret_cast<const float2*>(gend + v);
    s_end[r * END_PITCH + c] = x.x; s_end[r * END_PITCH + c + 1] = x.y;
  }
  float4 o4 = make_float4(1.f, 0.f, 0.f, 0.f);
  float sb0 = 1.f, sb1 = 1.f, sb2 = 1.f;
  const long long gblk = g >> 5;
  const bool act = tid < rows && !s_static[tid];   // own flag: written by this thread above
  int ur0 = 0, ucnt = 0, uw0 = 0;
  if (tid < rows) { ur0 = uoff[gblk]; ucnt = uoff[gblk + 1] - ur0; uw0 = woff[gblk]; }
  if (act) { o4 = ldg4(rot + 4 * g); sb0 = scale_backup[3 * g]; sb1 = scale_backup[3 * g + 1]; sb2 = scale_backup[3 * g + 2]; }
  __syncthreads();
  const float4* gsh = reinterpret_cast<const float4*>(shs + g0 * SH_FLOATS);
  const int cp_r0 = tid / 12, cp_c4 = tid - 12 * cp_r0;
  // The union rows are consumed in groups of four (one slot word); a group's 24 weights are requested one group ahead, the
  // first group before the bulk SH copy (responses return roughly in issue order per SM).
  const float* wrow = uw + (long long)ur0 * 6 * 32 + lane;
  const uint32_t* srow = usw + (long long)uw0 * 32 + lane;
  float wn[4][6];
  uint32_t wordn = 0u;
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int ep = 0; ep < 6; ep++) wn[r][ep] = 0.f;
  if (act && ucnt > 0) {
    wordn = __ldg(srow);
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (r < ucnt)
#pragma unroll
        for (int ep = 0; ep < 6; ep++) wn[r][ep] = __ldg(wrow + (r * 6 + ep) * 32);
  }
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) cp_async16(s_sh4 + r * FIT_PITCH4 + cp_c4, gsh + tid + 120 * t);
    }
  }
  cp_async_commit();

  float Rs[3][3];
  if (act) {
    float* e = s_end + tid * END_PITCH;
    // ---- end-point skinning (tolerance mode): p' = p + sum over the Gaussian's union of w (rec . [p - c; 1])
    float q[18], sd[18];
#pragma unroll
    for (int v = 0; v < 18; v++) { q[v] = e[v] - (v % 3 == 0 ? cx : v % 3 == 1 ? cy : cz); sd[v] = 0.f; }
    for (int tb = 0; tb < ucnt; tb += 4) {
      float w[4][6];
      const uint32_t word = wordn;
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int ep = 0; ep < 6; ep++) w[r][ep] = wn[r][ep];
      if (tb + 4 < ucnt) {   // next group
        wordn = __ldg(srow + (long long)((tb + 4) >> 2) * 32);
#pragma unroll
        for (int r = 0; r < 4; r++)
          if (tb + 4 + r < ucnt)
#pragma unroll
            for (int ep = 0; ep < 6; ep++) wn[r][ep] = __ldg(wrow + ((long long)(tb + 4 + r) * 6 + ep) * 32);
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (tb + r < ucnt) {
          const unsigned slot = (word >> (r * 8)) & 0xffu;
          float4 a, b, c;
          if (tcnt) { a = s_rec[3 * slot]; b = s_rec[3 * slot + 1]; c = s_rec[3 * slot + 2]; }
          else stage32(nodes + unode[((long long)ur0 + tb + r) * 32 + lane], cx, cy, cz, a, b, c);   // tile's node list overflowed (uniform per CTA)
#pragma unroll
          for (int ep = 0; ep < 6; ep++) {
            const float d0 = fmaf(a.z, q[3 * ep + 2], fmaf(a.y, q[3 * ep + 1], fmaf(a.x, q[3 * ep], a.w)));
            const float d1 = fmaf(b.z, q[3 * ep + 2], fmaf(b.y, q[3 * ep + 1], fmaf(b.x, q[3 * ep], b.w)));
            const float d2 = fmaf(c.z, q[3 * ep + 2], fmaf(c.y, q[3 * ep + 1], fmaf(c.x, q[3 * ep], c.w)));
            sd[3 * ep] = fmaf(w[r][ep], d0, sd[3 * ep]); sd[3 * ep + 1] = fmaf(w[r][ep], d1, sd[3 * ep + 1]); sd[3 * ep + 2] = fmaf(w[r][ep], d2, sd[3 * ep + 2]);
          }
        }
      }
    }
#pragma unroll
    for (int v = 0; v < 18; v++) e[v] = e[v] + sd[v];
    // ---- six-point fit (UpdateAsSixPointsWithdrawBad, GV:3081-3166), as k_fit_gaussians
    const Quat oq{o4.x, o4.y, o4.z, o4.w};
    float c[3];
#pragma unroll
    for (int r = 0; r < 3; r++)
      c[r] = ((e[r] + (e[3 + r] + e[6 + r])) + (e[9 + r] + (e[12 + r] + e[15 + r]))) / 6.0f;
    double Md[3][3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        const float a = e[6 * i + r] - c[r], b = e[6 * i + 3 + r] - c[r];
        Md[r][i] = (double)(0.5f * a + (-0.5f) * b);
      }
    double Rd[3][3];
    polar_newton(Md, Rd);
    float Rf[3][3], Kd[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) Rf[i][j] = (float)Rd[i][j];
      Kd[i] = (float)fma(Rd[0][i], Md[0][i], fma(Rd[1][i], Md[1][i], Rd[2][i] * Md[2][i]));
    }
    const Quat qn = quat_normalized(quat_from_matrix(Rf));
    *reinterpret_cast<float4*>(rot + 4 * g) = make_float4(qn.w, qn.x, qn.y, qn.z);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const float s0 = i == 0 ? sb0 : i == 1 ? sb1 : sb2;
      scale[3 * g + i] = Kd[i] / ((s0 + 1e-3f) * 2.0f) * s0;
      pos[3 * g + i] = c[i];
    }
    const Quat rq = quat_normalized(quat_mul(qn, quat_inverse(oq)));
    quat_to_matrix(rq, Rs);
  }
  cp_async_wait<0>();
  __syncthreads();
  // deformed end points back to global (the next step's input), coalesced; rows of static Gaussians are unchanged and skipped
  for (int v = tid * 2; v < nend; v += FIT_TILE * 2) {
    const int r = v / 18, c = v - r * 18;
    if (!s_static[r]) *reinterpret_cast<float2*>(gend + v) = make_float2(s_end[r * END_PITCH + c], s_end[r * END_PITCH + c + 1]);
  }
  if (act) {
    float4* row = s_sh4 + tid * FIT_PITCH4;
    float v[SH_FLOATS];
#pragma unroll
    for (int c = 0; c < 12; c++) { const float4 x = row[c]; v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w; }
    sh_rotate_flipped_fast(Rs, v);
#pragma unroll
    for (int c = 1; c < 12; c++) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    row[0].w = v[3];
  }
  __syncthreads();
  float* osh = shs + g0 * SH_FLOATS;
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) st_stream4(osh + (size_t)(tid + 120 * t) * 4, s_sh4[r * FIT_PITCH4 + cp_c4]);
    }
  }
  if (PUSH) {   // every thread's pos / rot / scale stores precede the __syncthreads above: the tile's pose is complete in L2
    auto push = [&](const float* own, float* const* peer, int w) {
      const long long o = g0 * w;
      const int nf = rows * w;
      const int n4 = (reinterpret_cast<uintptr_t>(own) & 15) == 0 ? nf >> 2 : 0;   // g0 * w * 4 bytes is a multiple of 16
      for (int c = tid; c < n4; c += FIT_TILE) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(own + o) + c);
#pragma unroll
        for (int q = 0; q < ARAP_MAX_PEERS; q++)
          if (q < pp.n) { if (PUSH == 2) multimem_st(peer[q] + o + 4 * c, v); else reinterpret_cast<float4*>(peer[q] + o)[c] = v; }
      }
      for (int c = n4 * 4 + tid; c < nf; c += FIT_TILE) {
        const float v = __ldcg(own + o + c);
#pragma unroll
        for (int q = 0; q < ARAP_MAX_PEERS; q++)
          if (q < pp.n) { if (PUSH == 2) multimem_st(peer[q] + o + c, v); else peer[q][o + c] = v; }
      }
    };
    push(pos, pp.pos, 3); push(rot, pp.rot, 4); push(scale, pp.scale, 3);
  }
}

// ------------------------------------------------------------------ SH replay (multi-GPU receivers)
// A receiver of another rank's deformed Gaussians does not need their SH rows (192 of the 232 bytes per Gaussian): it holds
// last step's rows and rotations, receives the new rotations, and repeats the owner's SH update
//   R_sh = (q_new * q_old^-1).normalized -> sh_rotate_flipped_fast                      (k_fit_gaussians, GV:3138-3154)
// with the same inline arithmetic on the same float inputs, so its copy stays bit-identical to the owner's.
__global__ void __launch_bounds__(FIT_TILE, 4)
k_replay_shs(long long N, const float* __restrict__ rot_old, const float* __restrict__ rot_new,
             const uint8_t* __restrict__ is_static, float* __restrict__ shs) {
  extern __shared__ float4 s_sh4[];   // FIT_TILE x FIT_PITCH4 float4
  __shared__ uint8_t s_static[FIT_TILE];
  const long long g0 = (long long)blockIdx.x * FIT_TILE;
  const int tid = threadIdx.x;
  const int rows = (int)min((long long)FIT_TILE, N - g0);
  const long long g = g0 + tid;
  s_static[tid] = (tid < rows) ? (is_static ? is_static[g] : 0) : 1;
  float4 o4 = make_float4(1.f, 0.f, 0.f, 0.f), n4 = o4;
  if (tid < rows) { if (rot_old) o4 = ldg4(rot_old + 4 * g); n4 = ldg4(rot_new + 4 * g); }
  __syncthreads();
  const float4* gsh = reinterpret_cast<const float4*>(shs + g0 * SH_FLOATS);
  const int cp_r0 = tid / 12, cp_c4 = tid - 12 * cp_r0;
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) cp_async16(s_sh4 + r * FIT_PITCH4 + cp_c4, gsh + tid + 120 * t);
    }
  }
  cp_async_commit();
  const bool act = tid < rows && !s_static[tid];
  float Rs[3][3];
  if (act) {
    const Quat oq{o4.x, o4.y, o4.z, o4.w}, q{n4.x, n4.y, n4.z, n4.w};
    const Quat rq = quat_normalized(quat_mul(q, quat_inverse(oq)));
    quat_to_matrix(rq, Rs);
  }
  cp_async_wait<0>();
  __syncthreads();
  if (act) {
    float4* row = s_sh4 + tid * FIT_PITCH4;
    float v[SH_FLOATS];
#pragma unroll
    for (int c = 0; c < 12; c++) { const float4 x = row[c]; v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w; }
    sh_rotate_flipped_fast(Rs, v);
#pragma unroll
    for (int c = 1; c < 12; c++) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
    row[0].w = v[3];
  }
  __syncthreads();
  float* osh = shs + g0 * SH_FLOATS;
  if (tid < 120) {
#pragma unroll
    for (int t = 0; t < 13; t++) {
      const int r = cp_r0 + 10 * t;
      if (r < rows && !s_static[r]) st_stream4(osh + (size_t)(tid + 120 * t) * 4, s_sh4[r * FIT_PITCH4 + cp_c4]);
    }
  }
}

// ------------------------------------------------------------------ node quaternions
// FastgetOthogonalMatrix (helper.cpp:506-517): float Newton polar iteration,
// returns the iterate BEFORE the one that met the 1e-6 max-abs test.
__global__ void k_node_quats(int M, const double* __restrict__ rot, float4* __restrict__ q_xyzw) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= M) return;
  float A[3][3];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) A[r][c] = (float)rot[9 * n + c * 3 + r];
  for (int it = 0; it < 100; it++) {
    float T[3][3];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) T[r][c] = A[c][r];
    float cof[3][3];
    cof[0][0] = T[1][1] * T[2][2] - T[1][2] * T[2][1];
    cof[0][1] = T[1][2] * T[2][0] - T[1][0] * T[2][2];
    cof[0][2] = T[1][0] * T[2][1] - T[1][1] * T[2][0];
    cof[1][0] = T[0][2] * T[2][1] - T[0][1] * T[2][2];
    cof[1][1] = T[0][0] * T[2][2] - T[0][2] * T[2][0];
    cof[1][2] = T[0][1] * T[2][0] - T[0][0] * T[2][1];
    cof[2][0] = T[0][1] * T[1][2] - T[0][2] * T[1][1];
    cof[2][1] = T[0][2] * T[1][0] - T[0][0] * T[1][2];
    cof[2][2] = T[0][0] * T[1][1] - T[0][1] * T[1][0];
    const float det = T[0][0] * cof[0][0] + T[0][1] * cof[0][1] + T[0][2] * cof[0][2];
    const float invdet = 1.0f / det;
    float nx[3][3]; float md = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const float inv = cof[c][r] * invdet;
        nx[r][c] = 0.5f * (A[r][c] + inv);
        md = fmaxf(md, fabsf(A[r][c] - nx[r][c]));
      }
    if (md < 1e-6f) break;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) A[r][c] = nx[r][c];
  }
  const Quat q = quat_normalized(quat_from_matrix(A));
  q_xyzw[n] = make_float4(q.x, q.y, q.z, q.w);
}

// ------------------------------------------------------------------ sample SH rotation
// Persistent CTAs, 128-sample tiles, two shared-memory stages filled with 16-byte cp.async: the next tile's 24.5 KB
// of SH rows is in flight while the current tile's quaternions are blended and its rows rotated, so DRAM stays busy
// at 4 CTAs / SM without needing occupancy to hide the load -> compute -> store phases of a tile.
// Rows sit at a 13-chunk (208 B) pitch: a thread's 12 LDS.128 / STS.128 on its own row are conflict-free
// (13 r + c mod 8 is a permutation over 8 consecutive rows).  Static samples are copied through unchanged.
// Skinning weights (float) and node ids use the same 32-row blocked layout as the LBS tables.
constexpr int RS_TILE = 128;
constexpr int RS_PITCH4 = 13;                       // float4 chunks per staged row
constexpr int RS_STAGE4 = RS_TILE * RS_PITCH4;      // float4 per stage

// the rare sin branch of Q_SlerpCUDA, kept out of line so the blend loop stays small
__device__ __noinline__ void slerp_ratios_slow(float cf, float t, float& rA, float& rB) {
  const double cosa = (double)cf, td = (double)t;
  const double sina = sqrt(1.0 - cosa * cosa);
  const double ang = atan2(sina, cosa);
  rA = (float)(sin((1.0 - td) * ang) / sina);
  rB = (float)(sin(td * ang) / sina);
}

template <int K, bool PP>
__global__ void __launch_bounds__(RS_TILE, 3)
k_rotate_sample_shs(long long S, long long ntiles, int k, const float* __restrict__ w, const uint16_t* __restrict__ idx,
                    const float4* __restrict__ q_xyzw, const uint8_t* __restrict__ is_static,
                    float* __restrict__ feature, int order, ArapPosePush ps) {
  // stage = [SH rows: RS_STAGE4 float4][weights: 128 K float][node ids: 128 K u16], all blocked like the global tables
  // K = compile-time bound of the neighbour count k (register arrays); k sets the table layout
  const int STAGE16 = RS_STAGE4 + k * 32 + k * 16;   // 16-byte units
  extern __shared__ float4 s_tile[];
  const int tid = threadIdx.x;
  const int cp_r0 = tid / 12, cp_c4 = tid - 12 * cp_r0;   // this thread's row (mod 10) and chunk in the tile copy loops
  auto issue = [&](long long tile, int stage) {
    const long long s0 = tile * RS_TILE;
    const int rows = (int)min((long long)RS_TILE, S - s0);
    const float4* g = reinterpret_cast<const float4*>(feature + s0 * SH_FLOATS);
    float4* d = s_tile + stage * STAGE16;
    // 120 threads copy 10 rows x 12 chunks per pass: row and chunk of a thread are loop-invariant, every address is
    // base + compile-time constant (the v / 12 form cost ~8 instructions per copy, this one 2)
    if (tid < 120) {
      const float4* gs = g + tid;
      float4* ds = d + cp_r0 * RS_PITCH4 + cp_c4;
#pragma unroll
      for (int i = 0; i < 13; i++)
        if (cp_r0 + 10 * i < rows) cp_async16(ds + i * 10 * RS_PITCH4, gs + i * 120);
    }
    const int nblk = (rows + 31) >> 5;
    const float4* gw = reinterpret_cast<const float4*>(w + (s0 >> 5) * (long long)(k * 32));
    const float4* gi = reinterpret_cast<const float4*>(idx + (s0 >> 5) * (long long)(k * 32));
    for (int c = tid; c < nblk * k * 8; c += RS_TILE) cp_async16(d + RS_STAGE4 + c, gw + c);
    for (int c = tid; c < nblk * k * 4; c += RS_TILE) cp_async16(d + RS_STAGE4 + k * 32 + c, gi + c);
  };
  // multi-GPU (PP): the first ps.copy_ctas CTAs of the launch do nothing but the exchange — they stream this rank's final pose arrays
  // into the other ranks' copies (unicast peer stores or one multicast store) while the remaining CTAs rotate the SH rows; the
  // transfer is spread under the longest kernel of the step without stalling its tile pipeline on NVLink back-pressure
  const int ncopy = PP ? ps.copy_ctas : 0;
  if (PP && (int)blockIdx.x < ncopy) {
    const long long p4 = ps.n * 3 / 4, total4 = 2 * p4 + ps.n;
    for (long long i = (long long)blockIdx.x * RS_TILE + tid; i < total4; i += (long long)ncopy * RS_TILE) {
      if (i < p4) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(ps.pos) + i);
#pragma unroll
        for (int q = 0; q < ARAP_MAX_PEERS; q++)
          if (q < ps.peers.n) { if (ps.peers.multicast) multimem_st(ps.peers.pos[q] + 4 * i, v); else reinterpret_cast<float4*>(ps.peers.pos[q])[i] = v; }
      } else if (i < p4 + ps.n) {
        const long long o = i - p4;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(ps.rot) + o);
#pragma unroll
        for (int q = 0; q < ARAP_MAX_PEERS; q++)
          if (q < ps.peers.n) { if (ps.peers.multicast) multimem_st(ps.peers.rot[q] + 4 * o, v); else reinterpret_cast<float4*>(ps.peers.rot[q])[o] = v; }
      } else {
        const long long o = i - p4 - ps.n;
        const float4 v = __ldcg(reinterpret_cast<const float4*>(ps.scale) + o);
#pragma unroll
        for (int q = 0; q < ARAP_MAX_PEERS; q++)
          if (q < ps.peers.n) { if (ps.peers.multicast) multimem_st(ps.peers.scale[q] + 4 * o, v); else reinterpret_cast<float4*>(ps.peers.scale[q])[o] = v; }
      }
    }
    return;
  }
  const long long tstride = (long long)gridDim.x - ncopy;
  long long tile = (long long)blockIdx.x - ncopy;
  if (tile < ntiles) issue(tile, 0);
  cp_async_commit();
  for (int it = 0; tile < ntiles; tile += tstride, it++) {
    const int stage = it & 1;
    const long long s0 = tile * RS_TILE;
    const int rows = (int)min((long long)RS_TILE, S - s0);
    const bool stat = (is_static && tid < rows) ? is_static[s0 + tid] != 0 : false;
    const long long next = tile + tstride;
    if (!order) {
      if (next < ntiles) issue(next, stage ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    float4* st4 = s_tile + stage * STAGE16;
    // The quaternion gathers go out BEFORE the next tile's 32 KB of bulk loads (order = 1): memory responses come back
    // roughly in issue order per SM, so a gather queued behind the bulk loads waits for all of them (16 % of the stall
    // samples sat on the first use of e[0]).
    float4 e[K]; float cw[K];
    if (tid < rows) {
      const float* sw = reinterpret_cast<const float*>(st4 + RS_STAGE4) + (tid >> 5) * (k * 32) + (tid & 31);
      const uint16_t* si = reinterpret_cast<const uint16_t*>(st4 + RS_STAGE4 + k * 32) + (tid >> 5) * (k * 32) + (tid & 31);
#pragma unroll
      for (int j = 0; j < K; j++) if (j < k) { e[j] = __ldg(q_xyzw + si[j * 32]); cw[j] = sw[j * 32]; }
    }
    if (order) {
      if (next < ntiles) issue(next, stage ^ 1);
      cp_async_commit();
    }
    const bool live = tid < rows && !stat;
    if (live) {
      Quat wq{1.0f, 0.0f, 0.0f, 0.0f};
      float last = 0.0f;
#pragma unroll
      for (int j = 0; j < K; j++) {
        if (j >= k) break;
        // Q_SlerpCUDA (cudakdtree.cu:113-148).  The reference mixes double ratios with float quaternions; here the
        // common lerp branch (incremental steps: cos > 0.99995) runs in float, the sin branch keeps double ratios.
        // Quaternions agree with the reference's to ~1e-7, far inside the sample-SH tolerance.
        const float t = __fdividef(cw[j], cw[j] + last);
        Quat eq{e[j].w, e[j].x, e[j].y, e[j].z};
        float cf = wq.x * eq.x + wq.y * eq.y + wq.z * eq.z + wq.w * eq.w;
        if (cf < 0.0f) { eq.x = -eq.x; eq.y = -eq.y; eq.z = -eq.z; eq.w = -eq.w; cf = -cf; }
        float rA, rB;
        if (cf > 0.99995f) { rA = 1.0f - t; rB = t; }
        else slerp_ratios_slow(cf, t, rA, rB);
        Quat l;
        l.x = fmaf(rA, wq.x, rB * eq.x); l.y = fmaf(rA, wq.y, rB * eq.y);
        l.z = fmaf(rA, wq.z, rB * eq.z); l.w = fmaf(rA, wq.w, rB * eq.w);
        const float inv = rsqrtf(quat_n2(l));
        wq = Quat{l.w * inv, l.x * inv, l.y * inv, l.z * inv};
        last += cw[j];
      }
      float R[3][3];
      quat_to_matrix(quat_normalized(wq), R);
      float4* row = st4 + tid * RS_PITCH4;
      float v[SH_FLOATS];
#pragma unroll
      for (int c = 0; c < 12; c++) { const float4 x = row[c]; v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w; }
      sh_rotate_flipped_fast(R, v);
#pragma unroll
      for (int c = 1; c < 12; c++) row[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      row[0].w = v[3];   // DC term (floats 0-2) is rotation invariant
    }
    __syncthreads();
    float4* o = reinterpret_cast<float4*>(feature + s0 * SH_FLOATS);
    if (tid < 120) {
      float4* os = o + tid;
      const float4* ss = st4 + cp_r0 * RS_PITCH4 + cp_c4;
#pragma unroll
      for (int i = 0; i < 13; i++)
        if (cp_r0 + 10 * i < rows) st_stream4(reinterpret_cast<float*>(os + i * 120), ss[i * 10 * RS_PITCH4]);
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ sample SH rotation, deferred (lazy_sample_sh = 1)
// The aim features are only read at a stroke end (UpdateFeatures, GV:1578-1617) or by the stage-II optimiser, never between
// two drag steps, and SH rotation is a group representation: D(R_T) ... D(R_1) f = D(R_T ... R_1) f.  So a drag step only has
// to compose the step's blended sample quaternion (same Q_SlerpCUDA chain as k_rotate_sample_shs) onto a per-sample
// accumulator (16 B read + written instead of 384 B); the row is rotated once, by the accumulated quaternion, when a
// consumer asks for it (arapk_replay_shs with rot_old = identity).  Off by default: the per-step body is then exactly the
// reference's (FastUpdateSamplesSH every step, GV:1519).
template <int K>
__global__ void __launch_bounds__(256)
k_accumulate_sample_quats(long long S, int k, const float* __restrict__ w, const uint16_t* __restrict__ idx, const float4* __restrict__ q_xyzw,
                          const uint8_t* __restrict__ is_static, float4* __restrict__ qacc_wxyz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S || (is_static && is_static[i])) return;
  const long long base = (i >> 5) * (long long)(k * 32) + (i & 31);
  float4 e[K]; float cw[K];
#pragma unroll
  for (int j = 0; j < K; j++) if (j < k) { e[j] = __ldg(q_xyzw + idx[base + j * 32]); cw[j] = w[base + j * 32]; }
  const float4 a4 = qacc_wxyz[i];
  Quat wq{1.0f, 0.0f, 0.0f, 0.0f};
  float last = 0.0f;
#pragma unroll
  for (int j = 0; j < K; j++) {
    if (j >= k) break;
    const float t = __fdividef(cw[j], cw[j] + last);
    Quat eq{e[j].w, e[j].x, e[j].y, e[j].z};
    float cf = wq.x * eq.x + wq.y * eq.y + wq.z * eq.z + wq.w * eq.w;
    if (cf < 0.0f) { eq.x = -eq.x; eq.y = -eq.y; eq.z = -eq.z; eq.w = -eq.w; cf = -cf; }
    float rA, rB;
    if (cf > 0.99995f) { rA = 1.0f - t; rB = t; }
    else slerp_ratios_slow(cf, t, rA, rB);
    Quat l;
    l.x = fmaf(rA, wq.x, rB * eq.x); l.y = fmaf(rA, wq.y, rB * eq.y);
    l.z = fmaf(rA, wq.z, rB * eq.z); l.w = fmaf(rA, wq.w, rB * eq.w);
    const float inv = rsqrtf(quat_n2(l));
    wq = Quat{l.w * inv, l.x * inv, l.y * inv, l.z * inv};
    last += cw[j];
  }
  const Quat acc = quat_normalized(quat_mul(quat_normalized(wq), Quat{a4.x, a4.y, a4.z, a4.w}));
  qacc_wxyz[i] = make_float4(acc.w, acc.x, acc.y, acc.z);
}
__global__ void k_fill_identity_quats(long long S, float4* __restrict__ q) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) q[i] = make_float4(1.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ sample SH rotation, TMA version
// Same arithmetic as k_rotate_sample_shs; what changes is how a tile moves.  The feature array is described once by a 2-D
// tensor map ([S rows] x [48 floats], box 128 rows x 16 floats, 64-byte swizzle): a tile arrives as three
// cp.async.bulk.tensor loads issued by ONE thread and completes on an mbarrier, the tile's weights and node ids as two flat
// bulk copies on the same barrier, and the rotated tile leaves as three bulk tensor stores — no per-thread copy loops, no
// per-thread global address arithmetic (the first version spent ~15 % of its instructions and its top stall there).
// With the 64-byte swizzle a thread's four 16-byte chunks of a 64-byte row segment sit at chunk ^ ((row >> 1) & 3): the
// eight rows of a quarter-warp hit eight different 16-byte bank groups, so the per-thread LDS.128 / STS.128 on its own row
// are conflict-free without the 13-chunk padding.
constexpr int RT_FEAT = 3 * RS_TILE * 64;          // three column blocks of 16 floats, 128 rows x 64 B each
constexpr int RT_STAGE = 34816;                    // RT_FEAT + 128 * 12 * 4 (weights) + 128 * 12 * 2 (ids), rounded up to 1024

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "LAB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE_%=;\n"
      "bra LAB_WAIT_%=;\n"
      "LAB_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__device__ __noinline__ void rot_issue_tile(const CUtensorMap* map, long long tile, long long S, int k, const float* __restrict__ w,
                                           const uint16_t* __restrict__ idx, unsigned char* st, uint64_t* bar) {
  const long long s0 = tile * RS_TILE;
  const int rows = (int)min((long long)RS_TILE, S - s0);
  const int nblk = (rows + 31) >> 5;
  const unsigned wbytes = (unsigned)(nblk * k * 32 * 4), ibytes = (unsigned)(nblk * k * 32 * 2);
  mbar_expect_tx(bar, (unsigned)RT_FEAT + wbytes + ibytes);
#pragma unroll
  for (int c = 0; c < 3; c++) tma_load_2d(st + c * (RS_TILE * 64), map, 16 * c, (int)s0, bar);
  bulk_load_1d(st + RT_FEAT, w + (s0 >> 5) * (long long)(k * 32), wbytes, bar);
  bulk_load_1d(st + RT_FEAT + RS_TILE * 12 * 4, idx + (s0 >> 5) * (long long)(k * 32), ibytes, bar);
}

template <int K>
__global__ void __launch_bounds__(RS_TILE, 3)
k_rotate_sample_shs_tma(const __grid_constant__ CUtensorMap tmap, long long S, long long ntiles, int k, const float* __restrict__ w,
                        const uint16_t* __restrict__ idx, const float4* __restrict__ q_xyzw, const uint8_t* __restrict__ is_static) {
  extern __shared__ unsigned char s_raw[];
  __shared__ uint64_t s_bar[2];
  unsigned char* base = s_raw + ((1024u - (smem_u32(s_raw) & 1023u)) & 1023u);   // swizzled TMA boxes want 1024-byte alignment
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  auto issue = [&](long long tile, int stage) {   // one thread; out of line so that its address arithmetic does not cost every thread registers
    rot_issue_tile(&tmap, tile, S, k, w, idx, base + stage * RT_STAGE, &s_bar[stage]);
  };
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) issue(tile, 0);
  for (int it = 0; tile < ntiles; tile += gridDim.x, it++) {
    const int stage = it & 1;
    const long long s0 = tile * RS_TILE;
    const int rows = (int)min((long long)RS_TILE, S - s0);
    const bool stat = (is_static && tid < rows) ? is_static[s0 + tid] != 0 : false;
    unsigned char* st = base + stage * RT_STAGE;
    mbar_wait(&s_bar[stage], (unsigned)((it >> 1) & 1));
    // quaternion gathers of this tile first, then the next tile's bulk loads (responses return roughly in issue order per SM)
    float4 e[K]; float cw[K];
    if (tid < rows) {
      const float* sw = reinterpret_cast<const float*>(st + RT_FEAT) + (tid >> 5) * (k * 32) + (tid & 31);
      const uint16_t* si = reinterpret_cast<const uint16_t*>(st + RT_FEAT + RS_TILE * 12 * 4) + (tid >> 5) * (k * 32) + (tid & 31);
#pragma unroll
      for (int j = 0; j < K; j++) if (j < k) { e[j] = __ldg(q_xyzw + si[j * 32]); cw[j] = sw[j * 32]; }
    }
    const long long next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) {
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store that last read the other stage has drained it
      issue(next, stage ^ 1);
    }
    const bool live = tid < rows && !stat;
    if (live) {
      Quat wq{1.0f, 0.0f, 0.0f, 0.0f};
      float last = 0.0f;
#pragma unroll
      for (int j = 0; j < K; j++) {
        if (j >= k) break;
        const float t = __fdividef(cw[j], cw[j] + last);
        Quat eq{e[j].w, e[j].x, e[j].y, e[j].z};
        float cf = wq.x * eq.x + wq.y * eq.y + wq.z * eq.z + wq.w * eq.w;
        if (cf < 0.0f) { eq.x = -eq.x; eq.y = -eq.y; eq.z = -eq.z; eq.w = -eq.w; cf = -cf; }
        float rA, rB;
        if (cf > 0.99995f) { rA = 1.0f - t; rB = t; }
        else slerp_ratios_slow(cf, t, rA, rB);
        Quat l;
        l.x = fmaf(rA, wq.x, rB * eq.x); l.y = fmaf(rA, wq.y, rB * eq.y);
        l.z = fmaf(rA, wq.z, rB * eq.z); l.w = fmaf(rA, wq.w, rB * eq.w);
        const float inv = rsqrtf(quat_n2(l));
        wq = Quat{l.w * inv, l.x * inv, l.y * inv, l.z * inv};
        last += cw[j];
      }
      float R[3][3];
      quat_to_matrix(quat_normalized(wq), R);
      // this thread's four swizzled chunk positions inside a 64-byte row segment (the three column blocks are 8 KB apart)
      const unsigned sx = (unsigned)((tid >> 1) & 3);
      unsigned char* rb0 = st + tid * 64 + ((0u ^ sx) << 4);
      unsigned char* rb1 = st + tid * 64 + ((1u ^ sx) << 4);
      unsigned char* rb2 = st + tid * 64 + ((2u ^ sx) << 4);
      unsigned char* rb3 = st + tid * 64 + ((3u ^ sx) << 4);
      float v[SH_FLOATS];
#pragma unroll
      for (int c = 0; c < 12; c++) {
        unsigned char* a = ((c & 3) == 0 ? rb0 : (c & 3) == 1 ? rb1 : (c & 3) == 2 ? rb2 : rb3) + (c >> 2) * (RS_TILE * 64);
        const float4 x = *reinterpret_cast<const float4*>(a);
        v[4 * c] = x.x; v[4 * c + 1] = x.y; v[4 * c + 2] = x.z; v[4 * c + 3] = x.w;
      }
      sh_rotate_flipped_fast(R, v);
#pragma unroll
      for (int c = 1; c < 12; c++) {
        unsigned char* a = ((c & 3) == 0 ? rb0 : (c & 3) == 1 ? rb1 : (c & 3) == 2 ? rb2 : rb3) + (c >> 2) * (RS_TILE * 64);
        *reinterpret_cast<float4*>(a) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      }
      reinterpret_cast<float*>(rb0)[3] = v[3];   // DC term (floats 0-2) is rotation invariant
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the tile -> visible to the bulk store
    __syncthreads();
    if (tid == 0) {
#pragma unroll
      for (int c = 0; c < 3; c++) tma_store_2d(&tmap, 16 * c, (int)s0, st + c * (RS_TILE * 64));
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ static flags
// out[g] = 1 iff every neighbour of every row in group g is an excluded node.
__global__ void k_static_flags(long long G, int group, int k, const uint16_t* __restrict__ idx,
                               const uint8_t* __restrict__ node_static, uint8_t* __restrict__ out) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  uint8_t st = 1;
  for (long long r = g * group; r < (g + 1) * group; r++) {
    const long long base = (r >> 5) * (long long)(k * 32) + (r & 31);
    for (int j = 0; j < k; j++) st &= node_static[idx[base + j * 32]];
  }
  out[g] = st;
}

}  // namespace arapgs

// ===========================================================================
// launchers (kernel-level C ABI, see include/arapgs.h)
// ===========================================================================
using namespace arapgs;

static bool g_sh_ready = false;
static int ensure_sh_tables() {
  if (g_sh_ready) return ARAP_OK;
  ShCoef h;
  auto fill = [](int L, double* u, double* v, double* w) {
    for (int m = -L; m <= L; m++) for (int n = -L; n <= L; n++) {
      const int d = (m == 0), am = m < 0 ? -m : m, an = n < 0 ? -n : n;
      const double denom = (an == L) ? double(2 * L * (2 * L - 1)) : double((L + n) * (L - n));
      const int i = (m + L) * (2 * L + 1) + (n + L);
      u[i] = std::sqrt(double((L + m) * (L - m)) / denom);
      v[i] = 0.5 * std::sqrt(double((1 + d) * (L + am - 1) * (L + am)) / denom) * (1 - 2 * d);
      w[i] = -0.5 * std::sqrt(double((L - am - 1) * (L - am)) / denom) * (1 - d);
    }
  };
  fill(2, h.u2, h.v2, h.w2);
  fill(3, h.u3, h.v3, h.w3);
  ARAP_CUDA_TRY(cudaMemcpyToSymbol(c_sh, &h, sizeof(h)));
  ShCoefF f;
  for (int i = 0; i < 25; i++) { const int m = i / 5 - 2; f.u2[i] = (float)h.u2[i]; f.v2[i] = (float)(h.v2[i] * ((m == 1 || m == -1) ? 1.4142135623730951 : 1.0)); f.w2[i] = (float)h.w2[i]; }
  for (int i = 0; i < 49; i++) { const int m = i / 7 - 3; f.u3[i] = (float)h.u3[i]; f.v3[i] = (float)(h.v3[i] * ((m == 1 || m == -1) ? 1.4142135623730951 : 1.0)); f.w3[i] = (float)h.w3[i]; }
  ARAP_CUDA_TRY(cudaMemcpyToSymbol(c_shf, &f, sizeof(f)));
  g_sh_ready = true;
  return ARAP_OK;
}

extern "C" int arapk_node_xf(int M, const double* rot, const double* trans, const float* node_pos, void* node_xf,
                             void* node_xf32, cudaStream_t st) {
  if (M <= 0) return ARAP_OK;
  k_node_xf<<<(M + 127) / 128, 128, 0, st>>>(M, rot, trans, node_pos, (NodeXf*)node_xf, (NodeXf32*)node_xf32);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_lbs_points(const float* in, float* out, long long P, int k, const uint16_t* ridx, const double* rw,
                                const void* node_xf, const uint8_t* skip, int group, cudaStream_t st) {
  if (P <= 0) return ARAP_OK;
  if (k < 1 || k > KNN_MAX) { set_error("lbs_points: k out of range"); return ARAP_ERR_INVALID; }
  const int bs = 256;
  const unsigned grid = (unsigned)((P + bs - 1) / bs);
  const NodeXf* nx = (const NodeXf*)node_xf;
  if (group < 1) group = 1;
  switch (k) {
    case 8: k_lbs_points<8><<<grid, bs, 0, st>>>(in, out, P, k, ridx, rw, nx, skip, group); break;
    case 10: k_lbs_points<10><<<grid, bs, 0, st>>>(in, out, P, k, ridx, rw, nx, skip, group); break;
    case 12: k_lbs_points<12><<<grid, bs, 0, st>>>(in, out, P, k, ridx, rw, nx, skip, group); break;
    default: k_lbs_points<0><<<grid, bs, 0, st>>>(in, out, P, k, ridx, rw, nx, skip, group); break;
  }
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" long long arapk_lbs_tile_count(long long rows) { return (rows + LT_ROWS - 1) / LT_ROWS; }
extern "C" int arapk_lbs_tile_cap(void) { return LT_CAP; }

// slots: ceil(rows/32)*32*3 words; tile_cnt: tile_count entries; tile_nodes: tile_count * cap entries
extern "C" int arapk_lbs_build_tiles(long long rows, int k, const uint16_t* ridx, uint32_t* slots, uint16_t* tile_cnt,
                                     uint16_t* tile_nodes, cudaStream_t st) {
  if (rows <= 0) return ARAP_OK;
  if (k < 1 || k > KNN_MAX) { set_error("lbs_build_tiles: k out of range"); return ARAP_ERR_INVALID; }
  k_build_tiles<<<(unsigned)arapk_lbs_tile_count(rows), LT_ROWS, 0, st>>>(rows, k, ridx, slots, tile_cnt, tile_nodes);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

template <bool MAGIC>
static void launch_lbs_tiles(const float* in, float* out, long long P, int k, const uint32_t* slots, const double* rw,
                             const uint16_t* ridx, const uint16_t* tile_cnt, const uint16_t* tile_nodes, const NodeXf* nx,
                             const uint8_t* skip, int group, cudaStream_t st) {
  const unsigned grid = (unsigned)arapk_lbs_tile_count(P);
  switch (k) {
    case 8: k_lbs_tiles<8, MAGIC><<<grid, LT_ROWS, 0, st>>>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group); break;
    case 10: k_lbs_tiles<10, MAGIC><<<grid, LT_ROWS, 0, st>>>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group); break;
    case 12: k_lbs_tiles<12, MAGIC><<<grid, LT_ROWS, 0, st>>>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group); break;
    default: k_lbs_tiles<0, MAGIC><<<grid, LT_ROWS, 0, st>>>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group); break;
  }
}

// LBS through per-tile staged node records.  magic = 1: float rounding of the accumulator on the FP64 pipe
// (round_to_float), 0: by conversion instructions.  Results are bit-identical to arapk_lbs_points.
extern "C" int arapk_lbs_tiles(const float* in, float* out, long long P, int k, const uint32_t* slots, const double* rw,
                               const uint16_t* ridx, const uint16_t* tile_cnt, const uint16_t* tile_nodes,
                               const void* node_xf, const uint8_t* skip, int group, int magic, cudaStream_t st) {
  if (P <= 0) return ARAP_OK;
  if (k < 1 || k > KNN_MAX) { set_error("lbs_tiles: k out of range"); return ARAP_ERR_INVALID; }
  if (group < 1) group = 1;
  const NodeXf* nx = (const NodeXf*)node_xf;
  if (magic) launch_lbs_tiles<true>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group, st);
  else launch_lbs_tiles<false>(in, out, P, k, slots, rw, ridx, tile_cnt, tile_nodes, nx, skip, group, st);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_end_points(long long N, const float* pos, const float* rot, const float* scale, float* ends,
                                cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  k_end_points<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(N, pos, rot, scale, ends);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_fit_gaussians(long long N, const float* ends, const float* scale_backup, const uint8_t* is_static,
                                   float* pos, float* rot, float* scale, float* shs, cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  int rc = ensure_sh_tables(); if (rc) return rc;
  const size_t smem = sizeof(float4) * FIT_TILE * FIT_PITCH4 + sizeof(float) * FIT_TILE * END_PITCH;
  static bool attr_set = false;
  if (!attr_set) {
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_fit_gaussians, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_fit_gaussians<<<(unsigned)((N + FIT_TILE - 1) / FIT_TILE), FIT_TILE, smem, st>>>(N, ends, scale_backup, is_static, pos, rot, scale, shs);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

// ---- tolerance-mode (lbs_mode = 3) launchers
extern "C" int arapk_scan_inclusive_i32(const int* in, int* out, long long n, int* sums_scratch, cudaStream_t st);   // grid.cu

// Block unions of a row family.  Two passes around a scan; *rows_out = total union rows.  boff: ceil(rows/32) + 1 ints
// (device).  blist / bw are allocated by the caller between the passes: call with blist == NULL first (counts + scan,
// returns the total), then with the arrays.
extern "C" int arapk_sunion_build(long long rows, int k, const uint16_t* ridx, const float* wf, const double* wd, int* boff,
                                  uint16_t* blist, float* bw, long long* rows_out, int* scratch /* nblk + 8192 ints */,
                                  cudaStream_t st) {
  if (rows <= 0) { if (rows_out) *rows_out = 0; return ARAP_OK; }
  if (k < 1 || k > KNN_MAX) { set_error("sunion_build: k out of range"); return ARAP_ERR_INVALID; }
  const long long nblk = (rows + 31) / 32;
  const unsigned grid = (unsigned)((nblk + 3) / 4);
  if (!blist) {
    int* cnt = scratch;
    k_sunion_build<false><<<grid, 128, 0, st>>>(rows, k, ridx, nullptr, nullptr, cnt, nullptr, nullptr, nullptr);
    ARAP_KERNEL_CHECK();
    ARAP_CUDA_TRY(cudaMemsetAsync(boff, 0, sizeof(int), st));
    int rc = arapk_scan_inclusive_i32(cnt, boff + 1, nblk, scratch + nblk, st); if (rc) return rc;
    int total = 0;
    ARAP_CUDA_TRY(cudaMemcpyAsync(&total, boff + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaStreamSynchronize(st));
    if (rows_out) *rows_out = total;
    return ARAP_OK;
  }
  k_sunion_build<true><<<grid, 128, 0, st>>>(rows, k, ridx, wf, wd, nullptr, boff, blist, bw);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_lbs_union32(const float* in, float* out, long long P, const int* boff, const uint16_t* blist, const float* bw,
                                 const void* node_xf32, const uint8_t* skip, int group, cudaStream_t st) {
  if (P <= 0) return ARAP_OK;
  if (group < 1) group = 1;
  const long long nblk = (P + 31) / 32;
  // one warp per 32-row block.  (A persistent grid-stride version with a two-deep software pipeline — next block's weights,
  // points and offsets in flight while the current one is skinned — measured 1.91 ms against 1.77 ms: the kernel is bound
  // by l1tex wavefronts, not by exposed latency.)
  k_lbs_union32<<<(unsigned)((nblk + SU_WARPS - 1) / SU_WARPS), SU_WARPS * 32, 0, st>>>(in, out, P, boff, blist, bw, (const NodeXf32*)node_xf32, skip, group);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" long long arapk_gtile_count(long long n_gaussians) { return (n_gaussians + FIT_TILE - 1) / FIT_TILE; }
extern "C" int arapk_gtile_cap(void) { return GT_CAP; }

// Per-Gaussian unions of the end-point rows.  Same two-pass protocol as arapk_sunion_build: usw == NULL -> counts, scans
// (uoff, woff: ceil(N/32) + 1 ints each) and totals (*rows_out union rows, *words_out slot words); then the fill.
extern "C" int arapk_gunion_build(long long N, int k, const uint16_t* end_ridx, const double* end_rw, int* uoff, int* woff,
                                  uint32_t* usw, uint16_t* unode, float* uw, uint16_t* gtile_cnt, uint16_t* gtile_nodes,
                                  long long* rows_out, long long* words_out, int* scratch /* 2 * nblk + 8192 + 2 ints */, cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  if (k < 1 || k > KNN_MAX) { set_error("gunion_build: k out of range"); return ARAP_ERR_INVALID; }
  const long long nblk = (N + 31) / 32;
  const unsigned grid = (unsigned)arapk_gtile_count(N);
  int* cnt = scratch; int* wcnt = scratch + nblk; int* sums = scratch + 2 * nblk; int* err = sums + 8192;
  if (!usw) {
    ARAP_CUDA_TRY(cudaMemsetAsync(err, 0, sizeof(int), st));
    k_gunion_build<false><<<grid, 128, 0, st>>>(N, k, end_ridx, end_rw, cnt, err, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    k_words_of_rows<<<(unsigned)((nblk + 255) / 256), 256, 0, st>>>(nblk, cnt, wcnt);
    ARAP_KERNEL_CHECK();
    ARAP_CUDA_TRY(cudaMemsetAsync(uoff, 0, sizeof(int), st)); ARAP_CUDA_TRY(cudaMemsetAsync(woff, 0, sizeof(int), st));
    int rc = arapk_scan_inclusive_i32(cnt, uoff + 1, nblk, sums, st); if (rc) return rc;
    rc = arapk_scan_inclusive_i32(wcnt, woff + 1, nblk, sums, st); if (rc) return rc;
    int h[3] = {0, 0, 0};
    ARAP_CUDA_TRY(cudaMemcpyAsync(&h[0], uoff + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaMemcpyAsync(&h[1], woff + nblk, sizeof(int), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaMemcpyAsync(&h[2], err, sizeof(int), cudaMemcpyDeviceToHost, st));
    ARAP_CUDA_TRY(cudaStreamSynchronize(st));
    if (h[2]) { set_error("gunion_build: a Gaussian touches more than 60 distinct nodes"); return ARAP_ERR_UNSUPPORTED; }
    if (rows_out) *rows_out = h[0];
    if (words_out) *words_out = h[1];
    return ARAP_OK;
  }
  k_gunion_build<true><<<grid, 128, 0, st>>>(N, k, end_ridx, end_rw, nullptr, err, uoff, woff, usw, unode, uw, gtile_cnt, gtile_nodes);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_apply_union(long long N, const void* node_xf32, const uint16_t* gtile_cnt, const uint16_t* gtile_nodes,
                                 const int* uoff, const int* woff, const uint32_t* usw, const uint16_t* unode, const float* uw,
                                 float* ends, const float* scale_backup, const uint8_t* is_static, float* pos, float* rot,
                                 float* scale, float* shs, cudaStream_t st) {
  return arapk_apply_union_push(N, node_xf32, gtile_cnt, gtile_nodes, uoff, woff, usw, unode, uw, ends, scale_backup, is_static, pos, rot, scale, shs, nullptr, st);
}
extern "C" int arapk_apply_union_push(long long N, const void* node_xf32, const uint16_t* gtile_cnt, const uint16_t* gtile_nodes,
                                      const int* uoff, const int* woff, const uint32_t* usw, const uint16_t* unode, const float* uw,
                                      float* ends, const float* scale_backup, const uint8_t* is_static, float* pos, float* rot,
                                      float* scale, float* shs, const ArapPeerPush* peers, cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  if (peers && (peers->n < 0 || peers->n > ARAP_MAX_PEERS)) { set_error("apply_union: bad peer count"); return ARAP_ERR_INVALID; }
  int rc = ensure_sh_tables(); if (rc) return rc;
  const size_t smem = sizeof(float4) * FIT_TILE * FIT_PITCH4 + sizeof(float) * FIT_TILE * END_PITCH + sizeof(float4) * GT_CAP * 3;
  static bool attr_set = false;
  if (!attr_set) {
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_apply_union<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_apply_union<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_apply_union<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  ArapPeerPush pp{};
  if (peers && peers->n > 0 && peers->multicast) {
    pp = *peers;
    k_apply_union<2><<<(unsigned)arapk_gtile_count(N), FIT_TILE, smem, st>>>(N, (const NodeXf32*)node_xf32, gtile_cnt, gtile_nodes, uoff, woff, usw,
                                                                           unode, uw, ends, scale_backup, is_static, pos, rot, scale, shs, pp);
  } else if (peers && peers->n > 0) {
    pp = *peers;
    k_apply_union<1><<<(unsigned)arapk_gtile_count(N), FIT_TILE, smem, st>>>(N, (const NodeXf32*)node_xf32, gtile_cnt, gtile_nodes, uoff, woff, usw,
                                                                           unode, uw, ends, scale_backup, is_static, pos, rot, scale, shs, pp);
  } else
    k_apply_union<0><<<(unsigned)arapk_gtile_count(N), FIT_TILE, smem, st>>>(N, (const NodeXf32*)node_xf32, gtile_cnt, gtile_nodes, uoff, woff, usw,
                                                                               unode, uw, ends, scale_backup, is_static, pos, rot, scale, shs, pp);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_replay_shs(long long N, const float* rot_old, const float* rot_new, const uint8_t* is_static, float* shs,
                                cudaStream_t st) {
  if (N <= 0) return ARAP_OK;
  int rc = ensure_sh_tables(); if (rc) return rc;
  const size_t smem = sizeof(float4) * FIT_TILE * FIT_PITCH4;
  static bool attr_set = false;
  if (!attr_set) {
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_replay_shs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k_replay_shs<<<(unsigned)((N + FIT_TILE - 1) / FIT_TILE), FIT_TILE, smem, st>>>(N, rot_old, rot_new, is_static, shs);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_node_quats(int M, const double* rot, float* q_xyzw, cudaStream_t st) {
  if (M <= 0) return ARAP_OK;
  k_node_quats<<<(M + 127) / 128, 128, 0, st>>>(M, rot, (float4*)q_xyzw);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr; cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

extern "C" int arapk_rotate_sample_shs(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                                       const uint8_t* is_static, float* feature, cudaStream_t st) {
  return arapk_rotate_sample_shs_push(S, k, w, idx, q_xyzw, is_static, feature, nullptr, st);
}
extern "C" int arapk_rotate_sample_shs_push(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                                            const uint8_t* is_static, float* feature, const ArapPosePush* push, cudaStream_t st) {
  if (S <= 0) return ARAP_OK;
  if (push && (push->peers.n <= 0 || push->peers.n > ARAP_MAX_PEERS || (push->n & 3) || (((uintptr_t)push->pos | (uintptr_t)push->rot | (uintptr_t)push->scale) & 15))) {
    set_error("rotate_sample_shs: pose push needs 1..7 peers, a Gaussian count divisible by 4 and 16-byte aligned arrays"); return ARAP_ERR_INVALID;
  }
  int rc = ensure_sh_tables(); if (rc) return rc;
  if (k < 1 || k > 12) { set_error("rotate_sample_shs: k out of range"); return ARAP_ERR_INVALID; }
  const long long ntiles = (S + RS_TILE - 1) / RS_TILE;
  static int sms = 0;
  if (!sms) {
    int dev = 0; ARAP_CUDA_TRY(cudaGetDevice(&dev)); ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int mx = (int)(2 * (sizeof(float4) * RS_STAGE4 + (size_t)RS_TILE * 12 * 6));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<10, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<10, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs<12, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
    const int mt = 2 * RT_STAGE + 1024;
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs_tma<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, mt));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs_tma<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, mt));
    ARAP_CUDA_TRY(cudaFuncSetAttribute(k_rotate_sample_shs_tma<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, mt));
  }
  const unsigned nb = (unsigned)std::min<long long>(ntiles, (long long)sms * 3);
  const float4* q4 = (const float4*)q_xyzw;
  // ARAP_ROT_TMA=1 selects the TMA version.  Measured on the 58.8M-sample workload: 5.89 ms against 5.25 ms for the cp.async version —
  // the kernel is bound by instruction issue on the SH recurrences (~2500 instructions per sample = 4.1 ms at 4 IPC per SM), the
  // tensor-map pipeline costs it registers (spills at 168 registers, 3 CTAs per SM), so off-loading the copies does not pay; kept
  // selectable, default off.
  static int use_tma = -1;
  if (use_tma < 0) { const char* ev = getenv("ARAP_ROT_TMA"); use_tma = ev ? atoi(ev) : 0; }
  EncodeTiledFn enc = (use_tma && !push) ? tensor_map_encoder() : nullptr;
  if (enc && ((uintptr_t)feature & 15) == 0 && S < (1LL << 31)) {
    // one tensor map per (pointer, row count): cached for the session's aim-feature array
    static CUtensorMap tmap; static const float* m_ptr = nullptr; static long long m_S = 0;
    if (m_ptr != feature || m_S != S) {
      const cuuint64_t dims[2] = {48, (cuuint64_t)S};
      const cuuint64_t strides[1] = {192};
      const cuuint32_t box[2] = {16, RS_TILE};
      const cuuint32_t estr[2] = {1, 1};
      const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)feature, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { set_error("rotate_sample_shs: cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"); return ARAP_ERR_CUDA; }
      m_ptr = feature; m_S = S;
    }
    const size_t smem = 2 * RT_STAGE + 1024;
    if (k <= 8) k_rotate_sample_shs_tma<8><<<nb, RS_TILE, smem, st>>>(tmap, S, ntiles, k, w, idx, q4, is_static);
    else if (k <= 10) k_rotate_sample_shs_tma<10><<<nb, RS_TILE, smem, st>>>(tmap, S, ntiles, k, w, idx, q4, is_static);
    else k_rotate_sample_shs_tma<12><<<nb, RS_TILE, smem, st>>>(tmap, S, ntiles, k, w, idx, q4, is_static);
    ARAP_KERNEL_CHECK();
    return ARAP_OK;
  }
  const size_t smem = 2 * (sizeof(float4) * RS_STAGE4 + (size_t)RS_TILE * k * 6);
  static int order = -1;   // ARAP_ROT_ORDER=0: bulk loads of the next tile issued before the gathers (first version)
  if (order < 0) { const char* ev = getenv("ARAP_ROT_ORDER"); order = ev ? atoi(ev) : 1; }
  ArapPosePush ps{};
  if (push) {
    ps = *push;
    static int copy_ctas = -1;   // ARAP_PUSH_CTAS: CTAs of the launch that carry the exchange (default 16)
    if (copy_ctas < 0) { const char* ev = getenv("ARAP_PUSH_CTAS"); copy_ctas = ev ? atoi(ev) : 16; }
    ps.copy_ctas = std::max(1, std::min(copy_ctas, (int)nb - 1));
    if (k <= 8) k_rotate_sample_shs<8, true><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
    else if (k <= 10) k_rotate_sample_shs<10, true><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
    else k_rotate_sample_shs<12, true><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
  } else if (k <= 8) k_rotate_sample_shs<8, false><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
  else if (k <= 10) k_rotate_sample_shs<10, false><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
  else k_rotate_sample_shs<12, false><<<nb, RS_TILE, smem, st>>>(S, ntiles, k, w, idx, q4, is_static, feature, order, ps);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_accumulate_sample_quats(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                                             const uint8_t* is_static, float* qacc_wxyz, cudaStream_t st) {
  if (S <= 0) return ARAP_OK;
  if (k < 1 || k > 12) { set_error("accumulate_sample_quats: k out of range"); return ARAP_ERR_INVALID; }
  const unsigned grid = (unsigned)((S + 255) / 256);
  const float4* q4 = (const float4*)q_xyzw; float4* qa = (float4*)qacc_wxyz;
  if (k <= 8) k_accumulate_sample_quats<8><<<grid, 256, 0, st>>>(S, k, w, idx, q4, is_static, qa);
  else if (k <= 10) k_accumulate_sample_quats<10><<<grid, 256, 0, st>>>(S, k, w, idx, q4, is_static, qa);
  else k_accumulate_sample_quats<12><<<grid, 256, 0, st>>>(S, k, w, idx, q4, is_static, qa);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}
extern "C" int arapk_fill_identity_quats(long long S, float* q_wxyz, cudaStream_t st) {
  if (S <= 0) return ARAP_OK;
  k_fill_identity_quats<<<(unsigned)((S + 255) / 256), 256, 0, st>>>(S, (float4*)q_wxyz);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

extern "C" int arapk_static_flags(long long G, int group, int k, const uint16_t* idx, const uint8_t* node_static,
                                  uint8_t* out, cudaStream_t st) {
  if (G <= 0) return ARAP_OK;
  k_static_flags<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(G, group, k, idx, node_static, out);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}

namespace arapgs {
__global__ void k_sh_rotate_test(const float* R9, float* shs, int fast) {
  float R[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = R9[3 * i + j];
  if (fast) sh_rotate_flipped_fast(R, shs); else sh_rotate_flipped(R, shs);
}
}  // namespace arapgs
// fast = 0: the reference's exact rounding (double coefficient products); fast = 1: the float version the per-step kernels use
extern "C" int arapk_sh_rotate_test(const float* R9, float* shs48_dev, int fast, cudaStream_t st) {
  int rc = ensure_sh_tables(); if (rc) return rc;
  k_sh_rotate_test<<<1, 1, 0, st>>>(R9, shs48_dev, fast);
  ARAP_KERNEL_CHECK();
  return ARAP_OK;
}
