// Stage (c): embedded-deformation Gauss-Newton solve, entirely on device.
//
// Replaces  Deform::real_time_deform / optimize / FastCalcJacobiMat /
// CalcEnergyFunc  (reference Deform.cpp:77-581) and putFreeInputs
// (Deform.hpp:140-151).  Same energy, same Jacobian, same Gauss-Newton loop
// (start from identity, <= 30 iterations, step halving, |h| < (|x|+1e-6)1e-6
// stop) — the sparse Cholesky of J^T J is replaced by a matrix-free,
// block-Jacobi (12x12) preconditioned conjugate gradient in double, run to a
// residual that makes every iterate agree with the direct solve.
//
// One persistent cooperative kernel runs the whole solve; work items are graph
// nodes (then constraint groups), one 16-lane tile per item.  J is never
// stored: rows are regenerated from node positions, and J^T J p is applied in
// two phases (u = J p by row owner; y = J^T u gathered by column owner) with
// one grid barrier each; the barrier also carries the CG dot products.
//
// Unknown layout per node (Deform.hpp:29-36, 140-151): x[0..8] = A column-major,
// x[9..11] = t.  Non-free (excluded) nodes keep identity and carry no unknowns.
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace arapgs {

constexpr int TILE = 16;
constexpr int SOLVE_THREADS = 512;
constexpr int TILES_PER_BLOCK = SOLVE_THREADS / TILE;
constexpr int NRED = 4;  // scalars reduced per barrier

struct SolveDev {
  // graph (static per graph build / block change)
  int M, k, n_groups;
  const float* node_pos;      // M x 3
  const int* nbr;             // M x k
  const int* in_off;          // M + 1
  const int* in_src;          // E
  const int* in_slot;         // E
  const int* anc_idx;         // M x k
  const double* anc_w;        // M x k
  const uint8_t* node_free;   // M
  const int* static_in_cnt;   // M
  const int* grp_off;         // n_groups + 1
  const int* grp_member;      // members (node ids)
  const float* grp_aim;       // n_groups x 3
  const int* cin_off;         // M + 1
  const int* cin_grp;         // entries sorted by group within a node
  const int* cin_member;
  const int* cin_slot;
  // weights (already square-rooted, Deform.hpp:452-454)
  double w_rot, w_reg, w_con;
  int max_gn, max_cg;
  double cg_tol;
  // work vectors (double)
  double *x, *h, *r, *z, *p0, *p1, *dinv, *u_rot, *u_reg, *u_con;
  double* partial;  // 2 x gridDim x NRED
  // outputs
  double *rot_out, *trans_out;
  double* stats;    // [0] gn iters [1] energy [2] halvings [3] |h| [4] total cg iters [5] last rel residual [6] flag
};

struct Tile {
  cg::thread_block_tile<TILE> t;
  int lane;
};

// ---- grid-wide sum of NRED scalars; doubles as the phase barrier -----------
__device__ __forceinline__ void grid_reduce(cg::grid_group& grid, const SolveDev& S, int& phase, double (&v)[NRED]) {
  __shared__ double s_part[SOLVE_THREADS / 32][NRED];
  __shared__ double s_tot[NRED];
#pragma unroll
  for (int q = 0; q < NRED; q++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NRED; q++) s_part[warp][q] = v[q];
  __syncthreads();
  double* buf = S.partial + (size_t)(phase & 1) * gridDim.x * NRED;
  if (threadIdx.x < NRED) {
    double a = 0.0;
    for (int w = 0; w < SOLVE_THREADS / 32; w++) a += s_part[w][threadIdx.x];
    buf[(size_t)blockIdx.x * NRED + threadIdx.x] = a;
  }
  grid.sync();
  if (warp == 0) {
    double a[NRED];
#pragma unroll
    for (int q = 0; q < NRED; q++) a[q] = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32)
#pragma unroll
      for (int q = 0; q < NRED; q++) a[q] += buf[(size_t)b * NRED + q];
#pragma unroll
    for (int q = 0; q < NRED; q++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
      if (lane == 0) s_tot[q] = a[q];
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NRED; q++) v[q] = s_tot[q];
  __syncthreads();
  phase++;
}

// value of vector `a + s*b` component (used for x+h and z+beta*p on the fly)
__device__ __forceinline__ double comb(const double* a, const double* b, double s, size_t i) {
  return b ? fma(s, b[i], a[i]) : a[i];
}

// Jrot entry (row r in 0..5, column c in 0..8) for current A (a[9], column-major), times w.
__device__ __forceinline__ double jrot(const double* a, int r, int c, double w) {
  const int col = c / 3, t = c - 3 * col;
  if (r < 3) {
    const int ca = (r == 2) ? 1 : 0, cb = (r == 0) ? 1 : 2;  // pairs (0,1),(0,2),(1,2)
    if (col == ca) return a[3 * cb + t] * w;
    if (col == cb) return a[3 * ca + t] * w;
    return 0.0;
  }
  return (col == r - 3) ? 2.0 * a[3 * col + t] * w : 0.0;
}

// ---------------------------------------------------------------------------
// Row phase.  MODE 0: nonlinear residual f(xa + sc*xb)   (CalcEnergyFunc)
//             MODE 1: linear u = J v, v = va + sc*vb, Jrot taken at S.x
// Results go to u_rot / u_reg / u_con; returns this lane's sum of squares.
// In MODE 1 the tile also stores its node's v to `vstore` (the new p).
// ---------------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ double row_phase(const SolveDev& S, const Tile& T, int gtile, int ntiles, const double* va,
                                            const double* vb, double sc, double* vstore) {
  const int M = S.M, k = S.k;
  double sq = 0.0;
  for (int item = gtile; item < M + S.n_groups; item += ntiles) {
    if (item < M) {
      const int i = item;
      if (!S.node_free[i]) continue;
      double v[12];
#pragma unroll
      for (int c = 0; c < 12; c++) v[c] = comb(va, vb, sc, (size_t)i * 12 + c);
      if (MODE == 1 && vstore && T.lane < 12) {
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < 12; c++) if (c == T.lane) mine = v[c];
        vstore[(size_t)i * 12 + T.lane] = mine;
      }
      const float gi0 = S.node_pos[3 * i], gi1 = S.node_pos[3 * i + 1], gi2 = S.node_pos[3 * i + 2];
      for (int l = T.lane; l < 3 * k; l += TILE) {
        const int s = l / 3, j = l - 3 * s;
        const int q = S.nbr[i * k + s];
        const float gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
        double tq = 0.0;
        if (S.node_free[q]) tq = comb(va, vb, sc, (size_t)q * 12 + 9 + j);
        double aj0 = 0, aj1 = 0, aj2 = 0, tj = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) if (c == j) { aj0 = v[c]; aj1 = v[c + 3]; aj2 = v[c + 6]; tj = v[9 + c]; }
        double val;
        if (MODE == 0) {
          const double d0 = (double)gq0 - (double)gi0, d1 = (double)gq1 - (double)gi1, d2 = (double)gq2 - (double)gi2;
          const double gij = (j == 0) ? (double)gi0 : (j == 1) ? (double)gi1 : (double)gi2;
          const double gqj = (j == 0) ? (double)gq0 : (j == 1) ? (double)gq1 : (double)gq2;
          val = S.w_reg * ((((fma(aj2, d2, fma(aj1, d1, aj0 * d0)) + gij) + tj) - gqj) - tq);
        } else {
          const double e0 = (double)(gq0 - gi0), e1 = (double)(gq1 - gi1), e2 = (double)(gq2 - gi2);  // float differences (Deform.cpp:254-256)
          val = S.w_reg * ((fma(aj2, e2, fma(aj1, e1, aj0 * e0)) + tj) - tq);
        }
        S.u_reg[((size_t)i * k + s) * 3 + j] = val;
        sq = fma(val, val, sq);
      }
      if (T.lane < 6) {
        const int rr = T.lane;
        double val;
        if (MODE == 0) {
          const double* a = v;
          const int ca = (rr < 3) ? ((rr == 2) ? 1 : 0) : rr - 3, cb = (rr < 3) ? ((rr == 0) ? 1 : 2) : rr - 3;
          double dt = 0.0;
#pragma unroll
          for (int t = 0; t < 3; t++) {
            double xa = 0, xb = 0;
#pragma unroll
            for (int c = 0; c < 9; c++) { if (c == 3 * ca + t) xa = a[c]; if (c == 3 * cb + t) xb = a[c]; }
            dt = fma(xa, xb, dt);
          }
          val = S.w_rot * (rr < 3 ? dt : dt - 1.0);
        } else {
          const double* a = S.x + (size_t)i * 12;
          val = 0.0;
#pragma unroll
          for (int c = 0; c < 9; c++) val = fma(jrot(a, rr, c, S.w_rot), v[c], val);
        }
        S.u_rot[(size_t)i * 6 + rr] = val;
        sq = fma(val, val, sq);
      }
      if (T.lane < 3) {  // static-side rows: one per (excluded node, slot) pointing here (Deform.cpp:268-297, 458-482)
        const int cnt = S.static_in_cnt[i];
        double tj = 0.0;
#pragma unroll
        for (int c = 0; c < 3; c++) if (c == T.lane) tj = v[9 + c];
        const double val = S.w_reg * tj;
        sq = fma((double)cnt * val, val, sq);
      }
    } else {
      // constraint group: rows = sum over members of skin(member) (- aim)
      const int g = item - M;
      const int mb = S.grp_off[g], me = S.grp_off[g + 1];
      double acc[3] = {0.0, 0.0, 0.0};
      const int total = (me - mb) * k;
      for (int l = T.lane; l < total; l += TILE) {
        const int m = l / k, s = l - m * k;
        const int c = S.grp_member[mb + m];
        const int q = S.anc_idx[c * k + s];
        const double wei = S.anc_w[c * k + s];
        const float vc0 = S.node_pos[3 * c], vc1 = S.node_pos[3 * c + 1], vc2 = S.node_pos[3 * c + 2];
        if (!S.node_free[q]) {
          if (MODE == 0) { acc[0] = fma(wei, (double)vc0, acc[0]); acc[1] = fma(wei, (double)vc1, acc[1]); acc[2] = fma(wei, (double)vc2, acc[2]); }
          continue;
        }
        const float gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
        double xv[12];
#pragma unroll
        for (int t = 0; t < 12; t++) xv[t] = comb(va, vb, sc, (size_t)q * 12 + t);
        if (MODE == 0) {
          const double d0 = (double)vc0 - (double)gq0, d1 = (double)vc1 - (double)gq1, d2 = (double)vc2 - (double)gq2;
#pragma unroll
          for (int j = 0; j < 3; j++) {
            const double gqj = (j == 0) ? (double)gq0 : (j == 1) ? (double)gq1 : (double)gq2;
            acc[j] = fma(wei, (fma(xv[j + 6], d2, fma(xv[j + 3], d1, xv[j] * d0)) + gqj) + xv[9 + j], acc[j]);
          }
        } else {
          const double e0 = (double)(vc0 - gq0), e1 = (double)(vc1 - gq1), e2 = (double)(vc2 - gq2);  // Deform.cpp:325-327
#pragma unroll
          for (int j = 0; j < 3; j++) acc[j] = fma(wei, fma(xv[j + 6], e2, fma(xv[j + 3], e1, xv[j] * e0)) + xv[9 + j], acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int o = TILE / 2; o > 0; o >>= 1) acc[j] += T.t.shfl_xor(acc[j], o);
      if (T.lane < 3) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 3; j++) if (j == T.lane) a = acc[j];
        double val;
        if (MODE == 0) val = S.w_con * (a - (double)(me - mb) * (double)S.grp_aim[3 * g + T.lane]);
        else val = S.w_con * a;
        S.u_con[(size_t)g * 3 + T.lane] = val;
        sq = fma(val, val, sq);
      }
    }
  }
  return sq;
}

// y_c = (J^T u)_c for node i, lane c < 12 (others return 0).  vt = the node's own
// translation part of the vector J was applied to (for the static-side rows).
__device__ __forceinline__ double gather_jt(const SolveDev& S, int i, int c, double vt_c) {
  const int k = S.k;
  const int col = c / 3, row = c - 3 * col;  // for c < 9: A(row, col) = x[row + 3 col]
  const bool is_t = c >= 9;
  const int j = is_t ? c - 9 : row;
  double y = 0.0;
  const float gi = is_t ? 0.f : S.node_pos[3 * i + col];
  if (!is_t) {
    const double* a = S.x + (size_t)i * 12;
#pragma unroll
    for (int r = 0; r < 6; r++) y = fma(jrot(a, r, c, S.w_rot), S.u_rot[(size_t)i * 6 + r], y);
  }
  for (int s = 0; s < k; s++) {
    const double u = S.u_reg[((size_t)i * k + s) * 3 + j];
    if (is_t) y = fma(S.w_reg, u, y);
    else {
      const int q = S.nbr[i * k + s];
      const double e = (double)(S.node_pos[3 * q + col] - gi);
      y = fma(S.w_reg * e, u, y);
    }
  }
  if (is_t) {
    for (int t = S.in_off[i]; t < S.in_off[i + 1]; t++) {
      const int src = S.in_src[t];
      if (S.node_free[src]) y = fma(-S.w_reg, S.u_reg[((size_t)src * k + S.in_slot[t]) * 3 + j], y);
    }
    y = fma((double)S.static_in_cnt[i] * S.w_reg * S.w_reg, vt_c, y);
  }
  for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {
    const int g = S.cin_grp[t], m = S.cin_member[t], s = S.cin_slot[t];
    const double wei = S.anc_w[m * k + s];
    const double u = S.u_con[(size_t)g * 3 + j];
    if (is_t) y = fma(S.w_con * wei, u, y);
    else {
      const double e = (double)(S.node_pos[3 * m + col] - gi);
      y = fma(S.w_con * wei * e, u, y);
    }
  }
  return y;
}

// Build the 12x12 diagonal block of J^T J for node i in shared memory (row-major
// D[12][12]), invert it via Cholesky, store the inverse (symmetric) to S.dinv.
__device__ __forceinline__ void build_dinv(const SolveDev& S, const Tile& T, int i, double* D /* 144 */, double* Jr /* 54 */) {
  const int k = S.k;
  const int lane = T.lane;
  const double* a = S.x + (size_t)i * 12;
  for (int t = lane; t < 144; t += TILE) D[t] = 0.0;
  for (int t = lane; t < 54; t += TILE) Jr[t] = jrot(a, t / 9, t % 9, S.w_rot);
  T.t.sync();
  // rot part: D[c'][c] += sum_r Jr[r][c'] Jr[r][c], lane = c
  if (lane < 9) {
    for (int cp = 0; cp < 9; cp++) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) s = fma(Jr[r * 9 + cp], Jr[r * 9 + lane], s);
      D[cp * 12 + lane] += s;
    }
  }
  T.t.sync();
  // 4x4 pattern shared by the three components: index map (a, j) -> j + 3a (a<3) or 9 + j
  {
    const int pa = lane >> 2, pb = lane & 3;
    const float gia = pa < 3 ? S.node_pos[3 * i + pa] : 0.f, gib = pb < 3 ? S.node_pos[3 * i + pb] : 0.f;
    double val = 0.0;
    for (int s = 0; s < k; s++) {
      const int q = S.nbr[i * k + s];
      const double ba = pa < 3 ? S.w_reg * (double)(S.node_pos[3 * q + pa] - gia) : S.w_reg;
      const double bb = pb < 3 ? S.w_reg * (double)(S.node_pos[3 * q + pb] - gib) : S.w_reg;
      val = fma(ba, bb, val);
    }
    // in-edges (free source: entry -w in its row; excluded source: static-side row -w): w^2 each on t_j
    if (pa == 3 && pb == 3) val = fma((double)(S.in_off[i + 1] - S.in_off[i]) * S.w_reg, S.w_reg, val);
    // constraints: per group, the node's aggregated entry b = sum_(member,slot) w wei (d,1); block = b b^T
    int cur = -1; double sa = 0.0, sb = 0.0;
    for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {
      const int g = S.cin_grp[t], m = S.cin_member[t], s = S.cin_slot[t];
      if (g != cur) { val = fma(sa, sb, val); sa = 0.0; sb = 0.0; cur = g; }
      const double wv = S.w_con * S.anc_w[m * k + s];
      sa += pa < 3 ? wv * (double)(S.node_pos[3 * m + pa] - gia) : wv;
      sb += pb < 3 ? wv * (double)(S.node_pos[3 * m + pb] - gib) : wv;
    }
    val = fma(sa, sb, val);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int ia = pa < 3 ? j + 3 * pa : 9 + j, ib = pb < 3 ? j + 3 * pb : 9 + j;
      D[ia * 12 + ib] += val;
    }
  }
  T.t.sync();
  // Cholesky (lower, in place)
  for (int j = 0; j < 12; j++) {
    if (lane == j) {
      double s = D[j * 12 + j];
      for (int t = 0; t < j; t++) s = fma(-D[j * 12 + t], D[j * 12 + t], s);
      D[j * 12 + j] = sqrt(s);
    }
    T.t.sync();
    if (lane > j && lane < 12) {
      double s = D[lane * 12 + j];
      for (int t = 0; t < j; t++) s = fma(-D[lane * 12 + t], D[j * 12 + t], s);
      D[lane * 12 + j] = s / D[j * 12 + j];
    }
    T.t.sync();
  }
  // inverse column `lane`: L y = e, L^T x = y
  if (lane < 12) {
    double y[12];
#pragma unroll
    for (int r = 0; r < 12; r++) {
      double s = (r == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int t = 0; t < 12; t++) if (t < r) s = fma(-D[r * 12 + t], y[t], s);
      y[r] = s / D[r * 12 + r];
    }
#pragma unroll
    for (int r = 11; r >= 0; r--) {
      double s = y[r];
#pragma unroll
      for (int t = 0; t < 12; t++) if (t > r) s = fma(-D[t * 12 + r], y[t], s);
      y[r] = s / D[r * 12 + r];
    }
#pragma unroll
    for (int r = 0; r < 12; r++) S.dinv[(size_t)i * 144 + r * 12 + lane] = y[r];
  }
  T.t.sync();
}

// z_c = sum_c' Dinv[c][c'] r_c'  with r distributed one per lane
__device__ __forceinline__ double apply_dinv(const SolveDev& S, const Tile& T, int i, double r_c) {
  double z = 0.0;
  const double* Di = S.dinv + (size_t)i * 144 + (T.lane < 12 ? T.lane : 0) * 12;
#pragma unroll
  for (int c = 0; c < 12; c++) {
    const double rc = T.t.shfl(r_c, c);
    z = fma(Di[c], rc, z);
  }
  return T.lane < 12 ? z : 0.0;
}

__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(SolveDev S) {
  cg::grid_group grid = cg::this_grid();
  cg::thread_block block = cg::this_thread_block();
  Tile T{cg::tiled_partition<TILE>(block), 0};
  T.lane = T.t.thread_rank();
  extern __shared__ double s_dyn[];  // TILES_PER_BLOCK x (144 + 54) doubles
  const int tile_in_block = threadIdx.x / TILE;
  double* const s_Dt = s_dyn + (size_t)tile_in_block * 198;
  double* const s_Jt = s_Dt + 144;
  const int gtile = blockIdx.x * TILES_PER_BLOCK + tile_in_block;
  const int ntiles = gridDim.x * TILES_PER_BLOCK;
  const int M = S.M;
  int phase = 0;
  double red[NRED];

  // x = identity (setIdentityRots, Deform.cpp:83-93); h = p = 0
  for (int i = gtile; i < M; i += ntiles)
    if (T.lane < 12) {
      const size_t o = (size_t)i * 12 + T.lane;
      S.x[o] = (T.lane == 0 || T.lane == 4 || T.lane == 8) ? 1.0 : 0.0;
      S.h[o] = 0.0; S.p0[o] = 0.0; S.p1[o] = 0.0; S.z[o] = 0.0; S.r[o] = 0.0;
    }
  grid.sync();

  int gn_iters = 0, halvings = 0, total_cg = 0;
  double energy = 0.0, normh = 0.0, last_rel = 0.0, abs_target = -1.0;
  bool have_f = false; double E0 = 0.0;
  int flag = 0;

  for (int gn = 0; gn < S.max_gn; gn++) {
    gn_iters = gn + 1;
    if (!have_f) {
      red[0] = row_phase<0>(S, T, gtile, ntiles, S.x, nullptr, 0.0, nullptr);
      red[1] = red[2] = red[3] = 0.0;
      grid_reduce(grid, S, phase, red);
      E0 = red[0];
    }
    energy = E0;
    // gradient g = -J^T f, preconditioner, first search direction
    double rz_l = 0.0, gg_l = 0.0, xx_l = 0.0;
    for (int i = gtile; i < M; i += ntiles) {
      if (!S.node_free[i]) continue;
      build_dinv(S, T, i, s_Dt, s_Jt);
      double g_c = 0.0, x_c = 0.0;
      if (T.lane < 12) {
        x_c = S.x[(size_t)i * 12 + T.lane];
        const double vt = T.lane >= 9 ? x_c : 0.0;
        g_c = -gather_jt(S, i, T.lane, vt);
      }
      const double z_c = apply_dinv(S, T, i, g_c);
      if (T.lane < 12) {
        const size_t o = (size_t)i * 12 + T.lane;
        S.r[o] = g_c; S.z[o] = z_c; S.h[o] = 0.0;
        rz_l = fma(g_c, z_c, rz_l); gg_l = fma(g_c, g_c, gg_l); xx_l = fma(x_c, x_c, xx_l);
      }
    }
    red[0] = rz_l; red[1] = gg_l; red[2] = xx_l; red[3] = 0.0;
    grid_reduce(grid, S, phase, red);
    double rz = red[0]; const double gg = red[1]; const double normv = sqrt(red[2]);
    if (abs_target < 0.0) abs_target = S.cg_tol * S.cg_tol * gg;  // absolute residual^2 target set by the first linear system

    // ---- PCG on (J^T J) h = g ------------------------------------------------
    double beta = 0.0;
    int cur = 0;
    if (gg > 0.0) {
      for (int it = 0; it < S.max_cg; it++) {
        double* pnew = cur ? S.p1 : S.p0;
        const double* pold = cur ? S.p0 : S.p1;
        // p = z + beta p_old (own and, on the fly, neighbours'); u = J p
        red[0] = row_phase<1>(S, T, gtile, ntiles, S.z, pold, beta, pnew);
        red[1] = red[2] = red[3] = 0.0;
        grid_reduce(grid, S, phase, red);
        const double pHp = red[0];
        const double alpha = rz / pHp;
        double rzn_l = 0.0, rr_l = 0.0;
        for (int i = gtile; i < M; i += ntiles) {
          if (!S.node_free[i]) continue;
          double r_c = 0.0;
          if (T.lane < 12) {
            const size_t o = (size_t)i * 12 + T.lane;
            const double p_c = pnew[o];
            const double y_c = gather_jt(S, i, T.lane, T.lane >= 9 ? p_c : 0.0);
            S.h[o] = fma(alpha, p_c, S.h[o]);
            r_c = fma(-alpha, y_c, S.r[o]);
            S.r[o] = r_c;
          }
          const double z_c = apply_dinv(S, T, i, r_c);
          if (T.lane < 12) {
            S.z[(size_t)i * 12 + T.lane] = z_c;
            rzn_l = fma(r_c, z_c, rzn_l); rr_l = fma(r_c, r_c, rr_l);
          }
        }
        red[0] = rzn_l; red[1] = rr_l; red[2] = red[3] = 0.0;
        grid_reduce(grid, S, phase, red);
        total_cg++;
        const double rzn = red[0], rr = red[1];
        last_rel = sqrt(rr / gg);
        cur ^= 1;
        if (!(pHp > 0.0) || !(rr == rr)) { flag = 1; break; }
        if (rr <= abs_target || rr <= 1e-30 * gg) break;
        beta = rzn / rz; rz = rzn;
        if (it == S.max_cg - 1) flag |= 2;
      }
    }

    // ---- step halving (Deform.cpp:144-156) ----------------------------------
    bool accepted = false;
    for (double alpha_ls = 1.0; alpha_ls > 1e-15; alpha_ls *= 0.5) {
      red[0] = row_phase<0>(S, T, gtile, ntiles, S.x, S.h, 1.0, nullptr);
      double hh_l = 0.0;
      for (int i = gtile; i < M; i += ntiles)
        if (S.node_free[i] && T.lane < 12) { const double hv = S.h[(size_t)i * 12 + T.lane]; hh_l = fma(hv, hv, hh_l); }
      red[1] = hh_l; red[2] = red[3] = 0.0;
      grid_reduce(grid, S, phase, red);
      const double E1 = red[0];
      if (E1 > E0) {
        for (int i = gtile; i < M; i += ntiles)
          if (S.node_free[i] && T.lane < 12) S.h[(size_t)i * 12 + T.lane] *= 0.5;
        halvings++;
        normh = 0.5 * sqrt(red[1]);
        grid.sync();
      } else {
        for (int i = gtile; i < M; i += ntiles)
          if (S.node_free[i] && T.lane < 12) { const size_t o = (size_t)i * 12 + T.lane; S.x[o] += S.h[o]; }
        normh = sqrt(red[1]);
        E0 = E1; have_f = true; accepted = true;
        grid.sync();
        break;
      }
    }
    if (!accepted) have_f = false;  // x unchanged, but the row buffers now hold f(x + h): recompute f(x)
    if (normh < (normv + 1e-6) * 1e-6) break;
  }

  // putFreeInputs (Deform.hpp:140-151); excluded nodes keep identity
  for (int i = gtile; i < M; i += ntiles)
    if (T.lane < 12) {
      const bool fr = S.node_free[i];
      const double v = fr ? S.x[(size_t)i * 12 + T.lane] : ((T.lane == 0 || T.lane == 4 || T.lane == 8) ? 1.0 : 0.0);
      if (T.lane < 9) S.rot_out[(size_t)i * 9 + T.lane] = v; else S.trans_out[(size_t)i * 3 + T.lane - 9] = v;
    }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    S.stats[0] = gn_iters; S.stats[1] = energy; S.stats[2] = halvings; S.stats[3] = normh;
    S.stats[4] = total_cg; S.stats[5] = last_rel; S.stats[6] = flag;
  }
}

}  // namespace arapgs

using namespace arapgs;

extern "C" size_t arapk_solve_workspace_bytes(int M, int k, int n_groups) {
  size_t d = (size_t)M * 12 * 6 + (size_t)M * 144 + (size_t)M * 6 + (size_t)M * k * 3 + (size_t)(n_groups + 1) * 3 + 2 * 1024 * NRED + 64;
  return d * sizeof(double);
}

extern "C" int arapk_solve(const ArapSolveGraph* G, const ArapSolveParams* P, void* workspace, size_t workspace_bytes,
                           double* rot_out, double* trans_out, double* stats_dev, cudaStream_t st) {
  if (G->M < 1 || G->k < 1 || G->k > KNN_MAX) { set_error("solve: bad graph"); return ARAP_ERR_INVALID; }
  if (workspace_bytes < arapk_solve_workspace_bytes(G->M, G->k, G->n_groups)) { set_error("solve: workspace too small"); return ARAP_ERR_INVALID; }
  SolveDev S;
  S.M = G->M; S.k = G->k; S.n_groups = G->n_groups;
  S.node_pos = G->node_pos; S.nbr = G->nbr; S.in_off = G->in_off; S.in_src = G->in_src; S.in_slot = G->in_slot;
  S.anc_idx = G->anc_idx; S.anc_w = G->anc_w; S.node_free = G->node_free; S.static_in_cnt = G->static_in_cnt;
  S.grp_off = G->grp_off; S.grp_member = G->grp_member; S.grp_aim = G->grp_aim;
  S.cin_off = G->cin_off; S.cin_grp = G->cin_grp; S.cin_member = G->cin_member; S.cin_slot = G->cin_slot;
  S.w_rot = std::sqrt(P->w_rot); S.w_reg = std::sqrt(P->w_reg); S.w_con = std::sqrt(P->w_con);
  S.max_gn = P->max_gn_iters > 0 ? P->max_gn_iters : 30;
  S.max_cg = P->max_cg_iters > 0 ? P->max_cg_iters : 4000;
  S.cg_tol = P->cg_tol > 0 ? P->cg_tol : 1e-10;
  double* w = (double*)workspace;
  const size_t v12 = (size_t)G->M * 12;
  S.x = w; w += v12; S.h = w; w += v12; S.r = w; w += v12; S.z = w; w += v12; S.p0 = w; w += v12; S.p1 = w; w += v12;
  S.dinv = w; w += (size_t)G->M * 144;
  S.u_rot = w; w += (size_t)G->M * 6;
  S.u_reg = w; w += (size_t)G->M * G->k * 3;
  S.u_con = w; w += (size_t)(G->n_groups + 1) * 3;
  S.partial = w;
  S.rot_out = rot_out; S.trans_out = trans_out; S.stats = stats_dev;
  int dev = 0, sms = 0;
  ARAP_CUDA_TRY(cudaGetDevice(&dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int per_sm = 0;
  const size_t smem = sizeof(double) * TILES_PER_BLOCK * 198;
  ARAP_CUDA_TRY(cudaFuncSetAttribute(k_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ARAP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_solve, SOLVE_THREADS, smem));
  if (per_sm < 1) { set_error("solve: kernel cannot be co-resident"); return ARAP_ERR_CUDA; }
  const int items = G->M + G->n_groups;
  int grid = std::min(sms, (items + TILES_PER_BLOCK - 1) / TILES_PER_BLOCK);
  grid = std::max(1, std::min(grid, 1024));
  void* args[] = {(void*)&S};
  ARAP_CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_solve, dim3(grid), dim3(SOLVE_THREADS), args, smem, st));
  return ARAP_OK;
}
