// Stage (c): embedded-deformation Gauss-Newton solve, entirely on device.
//
// Replaces  Deform::real_time_deform / optimize / FastCalcJacobiMat /
// CalcEnergyFunc  (reference Deform.cpp:77-581) and putFreeInputs
// (Deform.hpp:140-151).  Same energy, same Jacobian, same Gauss-Newton loop
// (start from identity, <= 30 iterations, step halving, |h| < (|x|+1e-6)1e-6
// stop) — the sparse Cholesky of J^T J is replaced by a matrix-free, Jacobi
// preconditioned conjugate gradient in double, run to a residual that makes
// every iterate agree with the direct solve.
//
// One persistent cooperative kernel runs the whole solve.  The work per PCG
// iteration is tiny (a few MFLOP), so the design goal is latency: every phase
// is "one thread per row" or "one thread per unknown" with short, independent
// load chains, and there are exactly two grid barriers per iteration, which
// also carry the dot products:
//   phase A  u = J p     one thread per E_reg row (node, slot, component), per node (E_rot), per constraint row
//   phase B  y = J^T u   one thread per unknown; + x/r/z updates (Jacobi: z = r / diag)
// J is never stored.  E_reg / E_con rows act identically on the three components j of a node's unknowns, each
// on the 4-vector (A[j,0], A[j,1], A[j,2], t_j); vectors are stored [node][component][4].  Row values of an edge
// are written twice: at the source (u_reg, its own gather) and into the destination's in-edge slot (u_in), so
// both gathers are contiguous.  Per-edge constants (float position differences, Deform.cpp:254-256) and
// constraint coefficients are computed once per solve.
//
// Reference unknown layout (Deform.hpp:29-36): x[0..8] = A column-major, x[9..11] = t, i.e. A[j,c] = x[j + 3c].
// Non-free (excluded) nodes keep identity and carry no unknowns.
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"
#include "solve_dev.h"

namespace cg = cooperative_groups;

namespace arapgs {

constexpr int SOLVE_THREADS = 1024;
constexpr int NRED = 3;

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// ---- grid-wide deterministic sum of NRED scalars; doubles as the phase barrier
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
#pragma unroll
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

struct D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ld4(const double* p) {
  const double2 lo = *reinterpret_cast<const double2*>(p), hi = *reinterpret_cast<const double2*>(p + 2);
  return D4{lo.x, lo.y, hi.x, hi.y};
}
__device__ __forceinline__ void st4(double* p, const D4& v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.a, v.b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v.c, v.d);
}
// va + s*vb (vb may be null)
__device__ __forceinline__ D4 ldcomb(const double* va, const double* vb, double s, size_t o) {
  D4 v = ld4(va + o);
  if (vb) { const D4 w = ld4(vb + o); v.a = fma(s, w.a, v.a); v.b = fma(s, w.b, v.b); v.c = fma(s, w.c, v.c); v.d = fma(s, w.d, v.d); }
  return v;
}
__device__ __forceinline__ double ldcomb1(const double* va, const double* vb, double s, size_t o) {
  return vb ? fma(s, vb[o], va[o]) : va[o];
}

// E_rot rows for a node with rows A0,A1,A2 (current x) applied to direction rows P0,P1,P2 (Deform.cpp:186-220):
// u0..2 = w (c_a . pc_b + c_b . pc_a) for column pairs (0,1),(0,2),(1,2); u3+c = 2 w (c_c . pc_c).  D4 fields a,b,c = columns.
__device__ __forceinline__ void rot_rows_lin(const D4& A0, const D4& A1, const D4& A2, const D4& P0, const D4& P1, const D4& P2,
                                             double w, double (&u)[6]) {
  const double a0p1 = fma(A0.a, P0.b, fma(A1.a, P1.b, A2.a * P2.b)), a1p0 = fma(A0.b, P0.a, fma(A1.b, P1.a, A2.b * P2.a));
  const double a0p2 = fma(A0.a, P0.c, fma(A1.a, P1.c, A2.a * P2.c)), a2p0 = fma(A0.c, P0.a, fma(A1.c, P1.a, A2.c * P2.a));
  const double a1p2 = fma(A0.b, P0.c, fma(A1.b, P1.c, A2.b * P2.c)), a2p1 = fma(A0.c, P0.b, fma(A1.c, P1.b, A2.c * P2.b));
  u[0] = w * (a0p1 + a1p0); u[1] = w * (a0p2 + a2p0); u[2] = w * (a1p2 + a2p1);
  u[3] = 2.0 * w * fma(A0.a, P0.a, fma(A1.a, P1.a, A2.a * P2.a));
  u[4] = 2.0 * w * fma(A0.b, P0.b, fma(A1.b, P1.b, A2.b * P2.b));
  u[5] = 2.0 * w * fma(A0.c, P0.c, fma(A1.c, P1.c, A2.c * P2.c));
}
// nonlinear E_rot residual (Deform.cpp:384-404)
__device__ __forceinline__ void rot_rows_res(const D4& A0, const D4& A1, const D4& A2, double w, double (&f)[6]) {
  f[0] = w * fma(A0.a, A0.b, fma(A1.a, A1.b, A2.a * A2.b));
  f[1] = w * fma(A0.a, A0.c, fma(A1.a, A1.c, A2.a * A2.c));
  f[2] = w * fma(A0.b, A0.c, fma(A1.b, A1.c, A2.b * A2.c));
  f[3] = w * (fma(A0.a, A0.a, fma(A1.a, A1.a, A2.a * A2.a)) - 1.0);
  f[4] = w * (fma(A0.b, A0.b, fma(A1.b, A1.b, A2.b * A2.b)) - 1.0);
  f[5] = w * (fma(A0.c, A0.c, fma(A1.c, A1.c, A2.c * A2.c)) - 1.0);
}
// entry c (< 3) of (Jrot^T u) on row j of A, whose own row is Aj
__device__ __forceinline__ double rot_rows_t(const D4& Aj, double w, const double (&u)[6], int c) {
  if (c == 0) return w * fma(u[0], Aj.b, fma(u[1], Aj.c, 2.0 * u[3] * Aj.a));
  if (c == 1) return w * fma(u[0], Aj.a, fma(u[2], Aj.c, 2.0 * u[4] * Aj.b));
  return w * fma(u[1], Aj.a, fma(u[2], Aj.b, 2.0 * u[5] * Aj.c));
}

// ---------------------------------------------------------------------------
// Row phase.  MODE 0: nonlinear residual f(va + sc*vb)  (CalcEnergyFunc, Deform.cpp:378-581)
//             MODE 1: u = J v, v = va + sc*vb, Jrot at S.x; v is also stored to vstore (the new search direction)
// Returns this thread's sum of squares.
// ---------------------------------------------------------------------------
template <int K, int MODE>
__device__ __forceinline__ double row_phase(const SolveDev& S, int gthread, int nthreads, const double* va, const double* vb,
                                            double sc, double* vstore) {
  const int M = S.M, k = K > 0 ? K : S.k;
  const int k3 = 3 * k;
  double sq = 0.0;
  // --- E_reg rows: (i, s, j)
  for (int t = gthread; t < M * k3; t += nthreads) {
    const int i = t / k3, rem = t - i * k3, s = rem / 3, j = rem - 3 * s;
    if (!S.node_free[i]) continue;
    const int e = i * k + s;
    const int q = S.nbr[e];
    const int slot = S.out_to_in[e];
    const size_t o = ((size_t)i * 3 + j) * 4;
    const D4 v = ldcomb(va, vb, sc, o);
    double tq = 0.0;
    if (S.node_free[q]) tq = ldcomb1(va, vb, sc, ((size_t)q * 3 + j) * 4 + 3);
    double val;
    if (MODE == 1) {
      const D4 b = ld4(S.bedge + (size_t)e * 4);
      val = S.w_reg * ((fma(v.c, b.c, fma(v.b, b.b, v.a * b.a)) + v.d) - tq);
      if (s == 0 && vstore) st4(vstore + o, v);
    } else {
      // mat*(gk-gj) + gj + tj - gk - tk with double differences (Deform.cpp:444-448)
      const double gi0 = S.node_pos[3 * i], gi1 = S.node_pos[3 * i + 1], gi2 = S.node_pos[3 * i + 2];
      const double gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
      const double gij = j == 0 ? gi0 : j == 1 ? gi1 : gi2, gqj = j == 0 ? gq0 : j == 1 ? gq1 : gq2;
      val = S.w_reg * ((((fma(v.c, gq2 - gi2, fma(v.b, gq1 - gi1, v.a * (gq0 - gi0))) + gij) + v.d) - gqj) - tq);
    }
    S.u_reg[(size_t)e * 3 + j] = val;
    if (slot >= 0) S.u_in[(size_t)slot * 3 + j] = val;
    sq = fma(val, val, sq);
    if (s == 0) {  // static-side rows: one per (excluded node, slot) pointing here (Deform.cpp:268-297, 458-482)
      const double sv = S.w_reg * v.d;
      sq = fma((double)S.static_in_cnt[i] * sv, sv, sq);
    }
  }
  // --- E_rot rows: one thread per node.  The short extra loops are dealt from the far end of the grid so they do
  // not pile onto the blocks that also own the constraint rows (barrier wait = slowest block).
  for (int i = nthreads - 1 - gthread; i < M; i += nthreads) {
    if (!S.node_free[i]) continue;
    const size_t ob = (size_t)i * 12;
    double u[6];
    if (MODE == 1) {
      const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
      const D4 P0 = ldcomb(va, vb, sc, ob), P1 = ldcomb(va, vb, sc, ob + 4), P2 = ldcomb(va, vb, sc, ob + 8);
      rot_rows_lin(A0, A1, A2, P0, P1, P2, S.w_rot, u);
    } else {
      const D4 A0 = ldcomb(va, vb, sc, ob), A1 = ldcomb(va, vb, sc, ob + 4), A2 = ldcomb(va, vb, sc, ob + 8);
      rot_rows_res(A0, A1, A2, S.w_rot, u);
    }
#pragma unroll
    for (int t = 0; t < 6; t++) sq = fma(u[t], u[t], sq);
  }
  // --- constraint rows: one 16-lane team per (group, component); lane = neighbour slot of the member's anchor row
  {
    const int l16 = threadIdx.x & 15;
    const unsigned tmask = 0xFFFFu << (threadIdx.x & 16);
    const int nteams = nthreads >> 4;
    int team = (gthread >> 4) + (nteams >> 1);   // start in the middle of the grid (see E_rot note)
    if (team >= nteams) team -= nteams;
    for (int t = team; t < 3 * S.n_groups; t += nteams) {
      const int g = t / 3, j = t - 3 * g;
      const int mb = S.grp_off[g], me = S.grp_off[g + 1];
      double acc = 0.0;
      for (int m = mb; m < me; m++) {
        const int c = S.grp_member[m];
        if (l16 < k) {
          const int q = S.anc_idx[c * k + l16];
          const double wei = S.anc_w[c * k + l16];
          const float vc0 = S.node_pos[3 * c], vc1 = S.node_pos[3 * c + 1], vc2 = S.node_pos[3 * c + 2];
          if (!S.node_free[q]) {
            if (MODE == 0) acc = fma(wei, j == 0 ? (double)vc0 : j == 1 ? (double)vc1 : (double)vc2, acc);
          } else {
            const float gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
            const D4 v = ldcomb(va, vb, sc, ((size_t)q * 3 + j) * 4);
            if (MODE == 1) {
              const double e0 = (double)(vc0 - gq0), e1 = (double)(vc1 - gq1), e2 = (double)(vc2 - gq2);  // float differences, Deform.cpp:325-327
              acc = fma(S.w_con * wei, fma(v.c, e2, fma(v.b, e1, v.a * e0)) + v.d, acc);
            } else {
              const double gqj = j == 0 ? (double)gq0 : j == 1 ? (double)gq1 : (double)gq2;
              acc = fma(wei, (fma(v.c, (double)vc2 - (double)gq2, fma(v.b, (double)vc1 - (double)gq1, v.a * ((double)vc0 - (double)gq0))) + gqj) + v.d, acc);
            }
          }
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(tmask, acc, o, 16);
      if (l16 == 0) {
        const double val = MODE == 1 ? acc : S.w_con * (acc - (double)(me - mb) * (double)S.grp_aim[3 * g + j]);
        S.u_con[t] = val;
        sq = fma(val, val, sq);
      }
    }
  }
  return sq;
}

// (J^T u) for unknown (i, j, c).  urot: the node's six E_rot row values; vt: this unknown's own value in the vector J
// was applied to (only used for c == 3: static-side rows).
template <int K>
__device__ __forceinline__ double gather_jt(const SolveDev& S, int i, int j, int c, const D4& Aj, const double (&urot)[6], double vt) {
  constexpr int KK = K > 0 ? K : KNN_MAX;
  const int k = K > 0 ? K : S.k;
  const double* ur = S.u_reg + (size_t)i * k * 3 + j;
  double y = 0.0;
  if (c < 3) {
    y = rot_rows_t(Aj, S.w_rot, urot, c);
    const double* be = S.bedge + (size_t)i * k * 4 + c;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < KK; s++) if (s < k) acc = fma(be[4 * s], ur[3 * s], acc);
    y = fma(S.w_reg, acc, y);
  } else {
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < KK; s++) if (s < k) acc += ur[3 * s];
    const double* ui = S.u_in + j;
    const int ib = S.in_off[i], ie = S.in_off[i + 1];
    double acc2 = 0.0;
    for (int t = ib; t < ie; t++) acc2 += ui[(size_t)t * 3];
    y = S.w_reg * (acc - acc2);
    y = fma((double)S.static_in_cnt[i] * S.w_reg * S.w_reg, vt, y);
  }
  const int cb = S.cin_off[i], ce = S.cin_off[i + 1];
  for (int t0 = cb; t0 < ce; t0 += 4) {   // 4 at a time so the dependent u_con gathers overlap
    double cc[4], uu[4]; int gg[4];
#pragma unroll
    for (int t = 0; t < 4; t++) { const bool ok = t0 + t < ce; gg[t] = ok ? S.cin_grp[t0 + t] : -1; cc[t] = ok ? S.ccoef[(size_t)(t0 + t) * 4 + c] : 0.0; }
#pragma unroll
    for (int t = 0; t < 4; t++) uu[t] = gg[t] >= 0 ? S.u_con[(size_t)gg[t] * 3 + j] : 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++) y = fma(cc[t], uu[t], y);
  }
  return y;
}

// diagonal of J^T J for unknown (i, j, c)
template <int K>
__device__ __forceinline__ double diag_jtj(const SolveDev& S, int i, int j, int c, const D4& A0, const D4& A1, const D4& A2) {
  constexpr int KK = K > 0 ? K : KNN_MAX;
  const int k = K > 0 ? K : S.k;
  double d = 0.0;
  if (c < 3) {
    // E_rot column of A[j][c]: two pair rows (entries A[j][other columns]) and the norm row of column c (2 A[j][c])
    const D4 Aj = j == 0 ? A0 : j == 1 ? A1 : A2;
    const double w2 = S.w_rot * S.w_rot;
    const double o1 = c == 0 ? Aj.b : Aj.a, o2 = c == 2 ? Aj.b : Aj.c, own = c == 0 ? Aj.a : c == 1 ? Aj.b : Aj.c;
    d = w2 * (o1 * o1 + o2 * o2 + 4.0 * own * own);
    const double* be = S.bedge + (size_t)i * k * 4 + c;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < KK; s++) if (s < k) acc = fma(be[4 * s], be[4 * s], acc);
    d = fma(S.w_reg * S.w_reg, acc, d);
  } else {
    // own rows (w each), in-edges from free sources (-w each), static-side rows (-w each)
    d = S.w_reg * S.w_reg * (double)(k + (S.in_off[i + 1] - S.in_off[i]) + S.static_in_cnt[i]);
  }
  int cur = -1; double sa = 0.0;
  for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {   // per group the node's aggregated entry (setFromTriplets sums duplicates)
    const int g = S.cin_grp[t];
    if (g != cur) { d = fma(sa, sa, d); sa = 0.0; cur = g; }
    sa += S.ccoef[(size_t)t * 4 + c];
  }
  d = fma(sa, sa, d);
  return d;
}

template <int K>
__global__ void __launch_bounds__(SOLVE_THREADS, 1) k_solve(SolveDev S) {
  cg::grid_group grid = cg::this_grid();
  const int gthread = blockIdx.x * SOLVE_THREADS + threadIdx.x, nthreads = gridDim.x * SOLVE_THREADS;
  const int M = S.M, k = K > 0 ? K : S.k;
  const int NU = M * 12;
  int phase = 0;
  double red[NRED];

  // ---- per-solve constants + x = identity (setIdentityRots, Deform.cpp:83-93)
  for (int t = gthread; t < NU; t += nthreads) {
    const int c = t & 3, jj = (t >> 2) % 3;
    S.x[t] = (c == jj) ? 1.0 : 0.0;
    S.h[t] = 0.0; S.p0[t] = 0.0; S.p1[t] = 0.0; S.z[t] = 0.0; S.r[t] = 0.0; S.dinv[t] = 0.0;
  }
  for (int t = gthread; t < M * k; t += nthreads) {  // float differences g_q - g_i (Deform.cpp:254-256)
    const int i = t / k, q = S.nbr[t];
    double* b = S.bedge + (size_t)t * 4;
    b[0] = (double)(S.node_pos[3 * q] - S.node_pos[3 * i]);
    b[1] = (double)(S.node_pos[3 * q + 1] - S.node_pos[3 * i + 1]);
    b[2] = (double)(S.node_pos[3 * q + 2] - S.node_pos[3 * i + 2]);
    b[3] = 1.0;
  }
  for (int i = gthread; i < M; i += nthreads)   // constraint coefficients w_con wei (v_c - g_q, 1) (Deform.cpp:325-328)
    for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {
      const int m = S.cin_member[t];
      const double wv = S.w_con * S.anc_w[m * k + S.cin_slot[t]];
      double* c = S.ccoef + (size_t)t * 4;
      c[0] = wv * (double)(S.node_pos[3 * m] - S.node_pos[3 * i]);
      c[1] = wv * (double)(S.node_pos[3 * m + 1] - S.node_pos[3 * i + 1]);
      c[2] = wv * (double)(S.node_pos[3 * m + 2] - S.node_pos[3 * i + 2]);
      c[3] = wv;
    }
  grid.sync();

  int gn_iters = 0, halvings = 0, total_cg = 0, flag = 0;
  double energy = 0.0, normh = 0.0, last_rel = 0.0, abs_target = -1.0, E0 = 0.0;
  bool have_f = false;
  double tphase[4] = {0.0, 0.0, 0.0, 0.0};  // ns spent in: rows, barrier 1, gather/update, barrier 2 (this thread's view)

  for (int gn = 0; gn < S.max_gn; gn++) {
    gn_iters = gn + 1;
    if (!have_f) {
      red[0] = row_phase<K, 0>(S, gthread, nthreads, S.x, nullptr, 0.0, nullptr);
      red[1] = red[2] = 0.0;
      grid_reduce(grid, S, phase, red);
      E0 = red[0];
    }
    energy = E0;
    // ---- gradient g = -J^T f, Jacobi preconditioner, z = g / diag
    double rz_l = 0.0, gg_l = 0.0, xx_l = 0.0;
    for (int t = gthread; t < NU; t += nthreads) {
      const int i = t / 12, qi = t - 12 * i, j = qi >> 2, c = qi & 3;
      if (!S.node_free[i]) continue;
      const size_t ob = (size_t)i * 12;
      const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
      double f[6]; rot_rows_res(A0, A1, A2, S.w_rot, f);
      const D4 Aj = j == 0 ? A0 : j == 1 ? A1 : A2;
      const double xv = c == 0 ? Aj.a : c == 1 ? Aj.b : c == 2 ? Aj.c : Aj.d;
      const double g = -gather_jt<K>(S, i, j, c, Aj, f, xv);
      const double di = 1.0 / diag_jtj<K>(S, i, j, c, A0, A1, A2);
      const double zv = g * di;
      S.dinv[t] = di; S.r[t] = g; S.z[t] = zv; S.h[t] = 0.0;
      rz_l = fma(g, zv, rz_l); gg_l = fma(g, g, gg_l); xx_l = fma(xv, xv, xx_l);
    }
    red[0] = rz_l; red[1] = gg_l; red[2] = xx_l;
    grid_reduce(grid, S, phase, red);
    double rz = red[0]; const double gg = red[1]; const double normv = sqrt(red[2]);
    if (abs_target < 0.0) abs_target = S.cg_tol * S.cg_tol * gg;  // absolute residual^2 target set by the first linear system

    // ---- PCG on (J^T J) h = g
    double beta = 0.0; int cur = 0;
    if (gg > 0.0) {
      for (int it = 0; it < S.max_cg; it++) {
        double* pnew = cur ? S.p1 : S.p0;
        const double* pold = cur ? S.p0 : S.p1;
        const unsigned long long t0 = gtime();
        red[0] = row_phase<K, 1>(S, gthread, nthreads, S.z, pold, beta, pnew);   // p = z + beta p_old; u = J p
        red[1] = red[2] = 0.0;
        const unsigned long long t1 = gtime();
        grid_reduce(grid, S, phase, red);
        const unsigned long long t2 = gtime();
        const double pHp = red[0];
        const double alpha = rz / pHp;
        double rzn_l = 0.0, rr_l = 0.0;
        for (int t = gthread; t < NU; t += nthreads) {
          const int i = t / 12, qi = t - 12 * i, j = qi >> 2, c = qi & 3;
          if (!S.node_free[i]) continue;
          const size_t ob = (size_t)i * 12;
          const D4 Aj = ld4(S.x + ob + 4 * j);
          double u[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
          if (c < 3) {
            const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
            const D4 P0 = ld4(pnew + ob), P1 = ld4(pnew + ob + 4), P2 = ld4(pnew + ob + 8);
            rot_rows_lin(A0, A1, A2, P0, P1, P2, S.w_rot, u);
          }
          const double pv = pnew[t];
          const double y = gather_jt<K>(S, i, j, c, Aj, u, pv);
          S.h[t] = fma(alpha, pv, S.h[t]);
          const double rv = fma(-alpha, y, S.r[t]);
          S.r[t] = rv;
          const double zv = rv * S.dinv[t];
          S.z[t] = zv;
          rzn_l = fma(rv, zv, rzn_l); rr_l = fma(rv, rv, rr_l);
        }
        red[0] = rzn_l; red[1] = rr_l; red[2] = 0.0;
        const unsigned long long t3 = gtime();
        grid_reduce(grid, S, phase, red);
        const unsigned long long t4 = gtime();
        tphase[0] += (double)(t1 - t0); tphase[1] += (double)(t2 - t1); tphase[2] += (double)(t3 - t2); tphase[3] += (double)(t4 - t3);
        total_cg++;
        const double rzn = red[0], rr = red[1];
        last_rel = sqrt(rr / gg);
        cur ^= 1;
        if (!(pHp > 0.0) || !(rr == rr)) { flag |= 1; break; }
        if (rr <= abs_target || rr <= 1e-30 * gg) break;
        beta = rzn / rz; rz = rzn;
        if (it == S.max_cg - 1) flag |= 2;
      }
    }

    if (flag & 1) {   // numeric breakdown (uniform over the grid): take a zero step, see solve_smem.cu
      for (int t = gthread; t < NU; t += nthreads) S.h[t] = 0.0;
      grid.sync();
    }
    // ---- step halving (Deform.cpp:144-156)
    bool accepted = false;
    for (double alpha_ls = 1.0; alpha_ls > 1e-15; alpha_ls *= 0.5) {
      red[0] = row_phase<K, 0>(S, gthread, nthreads, S.x, S.h, 1.0, nullptr);
      double hh_l = 0.0;
      for (int t = gthread; t < NU; t += nthreads) { const double hv = S.h[t]; hh_l = fma(hv, hv, hh_l); }
      red[1] = hh_l; red[2] = 0.0;
      grid_reduce(grid, S, phase, red);
      const double E1 = red[0];
      if (!(E1 <= E0)) {   // also rejects a non-finite energy
        for (int t = gthread; t < NU; t += nthreads) S.h[t] *= 0.5;
        halvings++;
        normh = 0.5 * sqrt(red[1]);
        grid.sync();
      } else {
        for (int t = gthread; t < NU; t += nthreads) S.x[t] += S.h[t];
        normh = sqrt(red[1]);
        E0 = E1; have_f = true; accepted = true;
        grid.sync();
        break;
      }
    }
    if (!accepted) have_f = false;  // x unchanged, but the row buffers now hold f(x + h): recompute f(x)
    if (normh < (normv + 1e-6) * 1e-6) break;
  }

  // putFreeInputs (Deform.hpp:140-151): rot column-major; excluded nodes keep identity (their x never changes)
  for (int t = gthread; t < NU; t += nthreads) {
    const int i = t / 12, qi = t - 12 * i, jj = qi >> 2, c = qi & 3;
    const double v = S.x[t];
    if (c < 3) S.rot_out[(size_t)i * 9 + jj + 3 * c] = v; else S.trans_out[(size_t)i * 3 + jj] = v;
  }
  if (gthread == 0) {
    S.stats[0] = gn_iters; S.stats[1] = energy; S.stats[2] = halvings; S.stats[3] = normh;
    S.stats[4] = total_cg; S.stats[5] = last_rel; S.stats[6] = flag;
    for (int t = 13; t < 24; t++) S.stats[t] = 0.0;
    S.stats[8] = tphase[0]; S.stats[9] = tphase[1]; S.stats[10] = tphase[2]; S.stats[11] = tphase[3]; S.stats[12] = gridDim.x;
  }
}

}  // namespace arapgs

using namespace arapgs;

extern "C" size_t arapk_solve_workspace_bytes(int M, int k, int n_groups) {
  // 7 vectors + edge constants + row buffers (source + destination copies) + constraint coefficients
  // (<= groups * 20 * k entries) + partials
  size_t d = (size_t)M * 12 * 7 + (size_t)M * k * 4 + (size_t)M * k * 3 * 2 + (size_t)(n_groups + 1) * 3 +
             (size_t)(n_groups + 1) * 20 * k * 4 * 2 + (size_t)(n_groups + 1) * 20 * k / 2 + 2 * 2048 * NRED + 64 + 32;
  d += solve_pipe_extra_doubles(M, k, n_groups) + 2;   // one-barrier kernel (solve_pipe.cu): double-buffered partials
  return d * sizeof(double);
}

extern "C" size_t arapk_solve_warm_doubles(int M) { return 8 + (size_t)SOLVE_WARM_MAX * (size_t)M * 12; }

extern "C" int arapk_solve(const ArapSolveGraph* G, const ArapSolveParams* P, void* workspace, size_t workspace_bytes,
                           double* rot_out, double* trans_out, double* stats_dev, cudaStream_t st) {
  if (G->M < 1 || G->k < 1 || G->k > KNN_MAX) { set_error("solve: bad graph"); return ARAP_ERR_INVALID; }
  if (workspace_bytes < arapk_solve_workspace_bytes(G->M, G->k, G->n_groups)) { set_error("solve: workspace too small"); return ARAP_ERR_INVALID; }
  if (G->n_cin_entries > (long long)(G->n_groups + 1) * 20 * G->k) { set_error("solve: constraint entry count exceeds the workspace bound"); return ARAP_ERR_INVALID; }
  SolveDev S;
  S.M = G->M; S.k = G->k; S.n_groups = G->n_groups;
  S.node_pos = G->node_pos; S.nbr = G->nbr; S.in_off = G->in_off; S.out_to_in = G->out_to_in;
  S.anc_idx = G->anc_idx; S.anc_w = G->anc_w; S.node_free = G->node_free; S.static_in_cnt = G->static_in_cnt;
  S.grp_off = G->grp_off; S.grp_member = G->grp_member; S.grp_aim = G->grp_aim;
  S.cin_off = G->cin_off; S.cin_grp = G->cin_grp; S.cin_member = G->cin_member; S.cin_slot = G->cin_slot;
  S.w_rot = std::sqrt(P->w_rot); S.w_reg = std::sqrt(P->w_reg); S.w_con = std::sqrt(P->w_con);
  S.max_gn = P->max_gn_iters > 0 ? P->max_gn_iters : 30;
  S.max_cg = P->max_cg_iters > 0 ? P->max_cg_iters : 4000;
  S.cg_tol = P->cg_tol > 0 ? P->cg_tol : 1e-10;
  S.eta0 = P->newton_eta0 > 0 ? P->newton_eta0 : 0.0;
  double* w = (double*)workspace;
  const size_t v12 = (size_t)G->M * 12;
  S.x = w; w += v12; S.h = w; w += v12; S.r = w; w += v12; S.z = w; w += v12; S.p0 = w; w += v12; S.p1 = w; w += v12; S.dinv = w; w += v12;
  S.bedge = w; w += (size_t)G->M * G->k * 4;
  S.ccoef = w; w += (size_t)(G->n_groups + 1) * 20 * G->k * 4;  // 16-byte aligned arrays (double2 loads) first
  S.gent_c = w; w += (size_t)(G->n_groups + 1) * 20 * G->k * 4;
  S.gent_q = reinterpret_cast<int*>(w); w += ((size_t)(G->n_groups + 1) * 20 * G->k + 1) / 2 + 1;
  if ((w - (double*)workspace) & 1) w += 1;
  S.u_reg = w; w += (size_t)G->M * G->k * 3;
  S.u_in = w; w += (size_t)G->M * G->k * 3;
  S.u_con = w; w += (size_t)(G->n_groups + 1) * 3;
  S.partial = w; w += 2 * 2048 * NRED;
  unsigned* counter = reinterpret_cast<unsigned*>(w);
  w += 64 + 32;
  if ((w - (double*)workspace) & 1) w += 1;
  double* pipe_extra = w;
  S.rot_out = rot_out; S.trans_out = trans_out; S.stats = stats_dev;
  S.warm = P->warm_buf;
  S.warm_systems = P->warm_systems > 0 ? std::min(P->warm_systems, SOLVE_WARM_MAX) : SOLVE_WARM_MAX;
  if (!P->force_global_kernel && P->pipelined) {   // fastest path: one grid barrier per PCG iteration (solve_pipe.cu)
    const int rc = launch_solve_pipe(S, reinterpret_cast<unsigned*>(S.partial), pipe_extra, st, P->max_ctas);
    if (rc >= 0) return rc;
  }
  if (!P->force_global_kernel) {   // fast path: per-node state resident in shared memory (solve_smem.cu)
    const int rc = launch_solve_smem(S, reinterpret_cast<unsigned*>(S.partial), st, P->max_ctas);   // barrier slots live in the partials area (2 x 148 x 64 B)
    if (rc >= 0) return rc;
  }
  int dev = 0, sms = 0, per_sm = 0;
  ARAP_CUDA_TRY(cudaGetDevice(&dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void* kern = G->k == 8 ? (void*)k_solve<8> : G->k == 10 ? (void*)k_solve<10> : G->k == 12 ? (void*)k_solve<12> : (void*)k_solve<0>;
  ARAP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, SOLVE_THREADS, 0));
  if (per_sm < 1) { set_error("solve: kernel cannot be co-resident"); return ARAP_ERR_CUDA; }
  // one thread per E_reg row is the widest phase; small graphs get a small grid (cheaper barriers)
  const long long rows = (long long)G->M * G->k * 3;
  int grid = (int)std::min<long long>((long long)sms * std::min(per_sm, 1), (rows + SOLVE_THREADS - 1) / SOLVE_THREADS);
  grid = std::max(1, std::min(grid, 2048));
  void* args[] = {(void*)&S};
  ARAP_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(SOLVE_THREADS), args, 0, st));
  return ARAP_OK;
}
