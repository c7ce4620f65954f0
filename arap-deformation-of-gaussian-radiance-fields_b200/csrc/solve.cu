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
// One persistent cooperative kernel runs the whole solve.  J is never stored:
//   * E_reg and E_con rows act identically on the three components j of a node's
//     unknowns, each on the 4-vector q_ij = (A[j,0], A[j,1], A[j,2], t_j); only
//     E_rot couples the components.  Vectors are therefore stored as
//     [node][component][4] and one QUAD of lanes (3 active) owns a node.
//   * J^T J p runs in two phases with one grid barrier each: u = J p by the row
//     owner (rows of an edge live with its source node), y = J^T u gathered by the
//     column owner.  The barriers also carry the CG dot products (p.Hp = |Jp|^2).
//   * per-edge constants (float position differences, Deform.cpp:254-256) and
//     constraint coefficients are computed once per solve.
//
// Reference unknown layout (Deform.hpp:29-36): x[0..8] = A column-major, x[9..11] = t,
// i.e. A[j,c] = x[j + 3c].  Non-free (excluded) nodes keep identity.
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace arapgs {

constexpr int SOLVE_THREADS = 256;
constexpr int QUADS_PER_BLOCK = SOLVE_THREADS / 4;
constexpr int TILE = 16;                              // D-block build / inversion
constexpr int TILES_PER_BLOCK = SOLVE_THREADS / TILE;
constexpr int NRED = 3;

struct SolveDev {
  int M, k, n_groups, n_entries;
  const float* node_pos;      // M x 3
  const int* nbr;             // M x k
  const int* in_off;          // M + 1
  const int* in_src;
  const int* in_slot;
  const int* anc_idx;         // M x k
  const double* anc_w;        // M x k
  const uint8_t* node_free;   // M
  const int* static_in_cnt;   // M
  const int* grp_off;         // n_groups + 1 -> members
  const int* grp_member;
  const float* grp_aim;       // n_groups x 3
  const int* cin_off;         // M + 1 -> (group, member, slot) entries touching the node, sorted by group
  const int* cin_grp;
  const int* cin_member;
  const int* cin_slot;
  double w_rot, w_reg, w_con;  // square-rooted (Deform.hpp:452-454)
  int max_gn, max_cg;
  double cg_tol;
  // work (double).  Vectors: [M][3][4]
  double *x, *h, *r, *z, *p0, *p1, *dinv /* M x 144, quad ordering */, *bedge /* M x k x 4 */, *ccoef /* cin entries x 4 */;
  double *u_reg /* M x k x 3 */, *u_con /* groups x 3 */;
  double* partial;             // 2 x gridDim x NRED
  double *rot_out, *trans_out, *stats;
};

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

// E_rot rows for a node with rows A0,A1,A2 (current x) applied to direction rows P0,P1,P2:
// u0..2 = w (c_a . pc_b + c_b . pc_a) for column pairs (0,1),(0,2),(1,2); u3+c = 2 w (c_c . pc_c)   (Deform.cpp:186-220)
__device__ __forceinline__ void rot_rows_lin(const D4& A0, const D4& A1, const D4& A2, const D4& P0, const D4& P1, const D4& P2,
                                             double w, double (&u)[6]) {
  // column c of A = (A0.c, A1.c, A2.c); D4 fields a,b,c = columns 0,1,2
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
// (Jrot^T u) restricted to row j of A: needs only the node's own row Aj
__device__ __forceinline__ void rot_rows_t(const D4& Aj, double w, const double (&u)[6], D4& y) {
  y.a = fma(w, fma(u[0], Aj.b, fma(u[1], Aj.c, 2.0 * u[3] * Aj.a)), y.a);
  y.b = fma(w, fma(u[0], Aj.a, fma(u[2], Aj.c, 2.0 * u[4] * Aj.b)), y.b);
  y.c = fma(w, fma(u[1], Aj.a, fma(u[2], Aj.b, 2.0 * u[5] * Aj.c)), y.c);
}

// ---------------------------------------------------------------------------
// Row phases (one quad per item; lanes 0..2 = component j)
// ---------------------------------------------------------------------------
// LIN: u = J v with v = va + sc*vb (Jrot at S.x); own v stored to vstore.  Returns this lane's sum of squares.
// K > 0: compile-time neighbour count, so the k neighbour gathers are issued as one batch (latency, not bandwidth,
// bounds this kernel).
template <int K>
__device__ __forceinline__ double rows_lin(const SolveDev& S, int gquad, int nquads, int j, const double* va, const double* vb,
                                           double sc, double* vstore) {
  constexpr int KK = K > 0 ? K : KNN_MAX;
  const int M = S.M, k = K > 0 ? K : S.k;
  double sq = 0.0;
  for (int item = gquad; item < M + S.n_groups; item += nquads) {
    if (j > 2) continue;
    if (item < M) {
      const int i = item;
      if (!S.node_free[i]) continue;
      const size_t o = ((size_t)i * 3 + j) * 4;
      const D4 v = ldcomb(va, vb, sc, o);
      if (vstore) st4(vstore + o, v);
      const double* be = S.bedge + (size_t)i * k * 4;
      double* ur = S.u_reg + (size_t)i * k * 3 + j;
      int q[KK]; double tqa[KK], tqb[KK]; uint8_t fr[KK];
#pragma unroll
      for (int s = 0; s < KK; s++) q[s] = (s < k) ? S.nbr[i * k + s] : i;
#pragma unroll
      for (int s = 0; s < KK; s++) {
        const size_t oq = ((size_t)q[s] * 3 + j) * 4 + 3;
        fr[s] = S.node_free[q[s]];
        tqa[s] = va[oq];
        tqb[s] = vb ? vb[oq] : 0.0;
      }
#pragma unroll
      for (int s = 0; s < KK; s++) {
        if (s < k) {
          const D4 b = ld4(be + 4 * s);
          const double tq = fr[s] ? fma(sc, tqb[s], tqa[s]) : 0.0;
          const double val = S.w_reg * ((fma(v.c, b.c, fma(v.b, b.b, v.a * b.a)) + v.d) - tq);
          ur[3 * s] = val;
          sq = fma(val, val, sq);
        }
      }
      {  // static-side rows: one per (excluded node, slot) pointing here (Deform.cpp:268-297)
        const double val = S.w_reg * v.d;
        sq = fma((double)S.static_in_cnt[i] * val, val, sq);
      }
      if (j == 0) {
        const size_t ob = (size_t)i * 12;
        const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
        const D4 P0 = v, P1 = ldcomb(va, vb, sc, ob + 4), P2 = ldcomb(va, vb, sc, ob + 8);
        double u[6]; rot_rows_lin(A0, A1, A2, P0, P1, P2, S.w_rot, u);
#pragma unroll
        for (int t = 0; t < 6; t++) sq = fma(u[t], u[t], sq);
      }
    } else {
      const int g = item - M;
      double acc = 0.0;
      for (int m = S.grp_off[g]; m < S.grp_off[g + 1]; m++) {
        const int c = S.grp_member[m];
        const float vc0 = S.node_pos[3 * c], vc1 = S.node_pos[3 * c + 1], vc2 = S.node_pos[3 * c + 2];
        int q[KK]; double wv[KK];
#pragma unroll
        for (int s = 0; s < KK; s++) { q[s] = (s < k) ? S.anc_idx[c * k + s] : c; wv[s] = (s < k) ? S.w_con * S.anc_w[c * k + s] : 0.0; }
#pragma unroll
        for (int s = 0; s < KK; s++) {
          if (s < k && S.node_free[q[s]]) {
            const double e0 = (double)(vc0 - S.node_pos[3 * q[s]]), e1 = (double)(vc1 - S.node_pos[3 * q[s] + 1]), e2 = (double)(vc2 - S.node_pos[3 * q[s] + 2]);  // Deform.cpp:325-327
            const D4 v = ldcomb(va, vb, sc, ((size_t)q[s] * 3 + j) * 4);
            acc = fma(wv[s], fma(v.c, e2, fma(v.b, e1, v.a * e0)) + v.d, acc);
          }
        }
      }
      S.u_con[(size_t)g * 3 + j] = acc;
      sq = fma(acc, acc, sq);
    }
  }
  return sq;
}

// RES: nonlinear residual f(xa + sc*xb) (CalcEnergyFunc, Deform.cpp:378-581) into the row buffers.
__device__ __forceinline__ double rows_res(const SolveDev& S, int gquad, int nquads, int j, const double* va, const double* vb, double sc) {
  const int M = S.M, k = S.k;
  double sq = 0.0;
  for (int item = gquad; item < M + S.n_groups; item += nquads) {
    if (j > 2) continue;
    if (item < M) {
      const int i = item;
      if (!S.node_free[i]) continue;
      const size_t o = ((size_t)i * 3 + j) * 4;
      const D4 v = ldcomb(va, vb, sc, o);
      const double gi0 = S.node_pos[3 * i], gi1 = S.node_pos[3 * i + 1], gi2 = S.node_pos[3 * i + 2];
      const double gij = j == 0 ? gi0 : j == 1 ? gi1 : gi2;
      double* ur = S.u_reg + (size_t)i * k * 3 + j;
      for (int s = 0; s < k; s++) {
        const int q = S.nbr[i * k + s];
        const double gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
        const double gqj = j == 0 ? gq0 : j == 1 ? gq1 : gq2;
        double tq = 0.0;
        if (S.node_free[q]) tq = ldcomb1(va, vb, sc, ((size_t)q * 3 + j) * 4 + 3);
        // mat*(gk-gj) + gj + tj - gk - tk with double differences (Deform.cpp:444-448)
        const double val = S.w_reg * ((((fma(v.c, gq2 - gi2, fma(v.b, gq1 - gi1, v.a * (gq0 - gi0))) + gij) + v.d) - gqj) - tq);
        ur[3 * s] = val;
        sq = fma(val, val, sq);
      }
      {
        const double val = S.w_reg * v.d;
        sq = fma((double)S.static_in_cnt[i] * val, val, sq);
      }
      if (j == 0) {
        const size_t ob = (size_t)i * 12;
        const D4 A1 = ldcomb(va, vb, sc, ob + 4), A2 = ldcomb(va, vb, sc, ob + 8);
        double f[6]; rot_rows_res(v, A1, A2, S.w_rot, f);
#pragma unroll
        for (int t = 0; t < 6; t++) sq = fma(f[t], f[t], sq);
      }
    } else {
      const int g = item - M;
      double acc = 0.0;
      const int mb = S.grp_off[g], me = S.grp_off[g + 1];
      for (int m = mb; m < me; m++) {
        const int c = S.grp_member[m];
        const double vc0 = S.node_pos[3 * c], vc1 = S.node_pos[3 * c + 1], vc2 = S.node_pos[3 * c + 2];
        const double vcj = j == 0 ? vc0 : j == 1 ? vc1 : vc2;
        for (int s = 0; s < k; s++) {
          const int q = S.anc_idx[c * k + s];
          const double wei = S.anc_w[c * k + s];
          if (!S.node_free[q]) { acc = fma(wei, vcj, acc); continue; }
          const double gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
          const double gqj = j == 0 ? gq0 : j == 1 ? gq1 : gq2;
          const D4 v = ldcomb(va, vb, sc, ((size_t)q * 3 + j) * 4);
          acc = fma(wei, (fma(v.c, vc2 - gq2, fma(v.b, vc1 - gq1, v.a * (vc0 - gq0))) + gqj) + v.d, acc);
        }
      }
      const double val = S.w_con * (acc - (double)(me - mb) * (double)S.grp_aim[3 * g + j]);
      S.u_con[(size_t)g * 3 + j] = val;
      sq = fma(val, val, sq);
    }
  }
  return sq;
}

// y = (J^T u) for (node i, component j).  urot: the node's six E_rot row values; vt: t_j of the vector J was applied to
// (static-side rows).
template <int K>
__device__ __forceinline__ D4 gather_jt(const SolveDev& S, int i, int j, const double (&urot)[6], double vt) {
  constexpr int KK = K > 0 ? K : KNN_MAX;
  const int k = K > 0 ? K : S.k;
  D4 y{0.0, 0.0, 0.0, 0.0};
  const D4 Aj = ld4(S.x + ((size_t)i * 3 + j) * 4);
  rot_rows_t(Aj, S.w_rot, urot, y);
  const double* be = S.bedge + (size_t)i * k * 4;
  const double* ur = S.u_reg + (size_t)i * k * 3 + j;
  const int ib = S.in_off[i], ie = S.in_off[i + 1];
#pragma unroll
  for (int s = 0; s < KK; s++) {
    if (s < k) {
      const double wu = S.w_reg * ur[3 * s];
      const D4 b = ld4(be + 4 * s);
      y.a = fma(b.a, wu, y.a); y.b = fma(b.b, wu, y.b); y.c = fma(b.c, wu, y.c); y.d += wu;
    }
  }
  for (int t0 = ib; t0 < ie; t0 += 8) {   // in-edges, 8 at a time so the dependent gathers overlap
    int src[8], sl[8]; double uu[8];
#pragma unroll
    for (int t = 0; t < 8; t++) { const bool ok = t0 + t < ie; src[t] = ok ? S.in_src[t0 + t] : -1; sl[t] = ok ? S.in_slot[t0 + t] : 0; }
#pragma unroll
    for (int t = 0; t < 8; t++) uu[t] = (src[t] >= 0 && S.node_free[src[t]]) ? S.u_reg[((size_t)src[t] * k + sl[t]) * 3 + j] : 0.0;
#pragma unroll
    for (int t = 0; t < 8; t++) y.d = fma(-S.w_reg, uu[t], y.d);
  }
  y.d = fma((double)S.static_in_cnt[i] * S.w_reg * S.w_reg, vt, y.d);
  for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {
    const D4 c = ld4(S.ccoef + (size_t)t * 4);
    const double u = S.u_con[(size_t)S.cin_grp[t] * 3 + j];
    y.a = fma(c.a, u, y.a); y.b = fma(c.b, u, y.b); y.c = fma(c.c, u, y.c); y.d = fma(c.d, u, y.d);
  }
  return y;
}

// z = Dinv r for one node; each lane holds its component's 4 residual entries.
__device__ __forceinline__ D4 apply_dinv(const SolveDev& S, int i, int j, unsigned qmask, const D4& r) {
  D4 z{0.0, 0.0, 0.0, 0.0};
  const double* Di = S.dinv + (size_t)i * 144 + (size_t)(j < 3 ? j : 0) * 48;  // rows 4j..4j+3
#pragma unroll
  for (int jj = 0; jj < 3; jj++) {
    const double ra = __shfl_sync(qmask, r.a, jj, 4), rb = __shfl_sync(qmask, r.b, jj, 4);
    const double rc = __shfl_sync(qmask, r.c, jj, 4), rd = __shfl_sync(qmask, r.d, jj, 4);
    const D4 d0 = ld4(Di + 0 * 12 + 4 * jj), d1 = ld4(Di + 1 * 12 + 4 * jj), d2 = ld4(Di + 2 * 12 + 4 * jj), d3 = ld4(Di + 3 * 12 + 4 * jj);
    z.a = fma(d0.a, ra, fma(d0.b, rb, fma(d0.c, rc, fma(d0.d, rd, z.a))));
    z.b = fma(d1.a, ra, fma(d1.b, rb, fma(d1.c, rc, fma(d1.d, rd, z.b))));
    z.c = fma(d2.a, ra, fma(d2.b, rb, fma(d2.c, rc, fma(d2.d, rd, z.c))));
    z.d = fma(d3.a, ra, fma(d3.b, rb, fma(d3.c, rc, fma(d3.d, rd, z.d))));
  }
  return z;
}

// ---- 12x12 diagonal block of J^T J for node i (reference index order in shared memory), Cholesky inverse stored in
// quad order: index (j,c) -> 4j + c, reference index of (j,c) = (c < 3) ? j + 3c : 9 + j.
__device__ __forceinline__ int ref_index(int qi) { const int j = qi >> 2, c = qi & 3; return c < 3 ? j + 3 * c : 9 + j; }

__device__ __forceinline__ double jrot_ref(const double* x12 /* [3][4] quad layout */, int r, int c, double w) {
  // reference column index c = 3*col + t  (entry A[t][col]); a(col,t) = x12[t*4 + col]
  const int col = c / 3, t = c - 3 * col;
  if (r < 3) {
    const int ca = (r == 2) ? 1 : 0, cb = (r == 0) ? 1 : 2;
    if (col == ca) return x12[t * 4 + cb] * w;
    if (col == cb) return x12[t * 4 + ca] * w;
    return 0.0;
  }
  return (col == r - 3) ? 2.0 * x12[t * 4 + col] * w : 0.0;
}

__device__ __forceinline__ void build_dinv(const SolveDev& S, cg::thread_block_tile<TILE>& T, int i, double* D /*144*/, double* Jr /*54*/) {
  const int k = S.k, lane = T.thread_rank();
  const double* a = S.x + (size_t)i * 12;
  for (int t = lane; t < 144; t += TILE) D[t] = 0.0;
  for (int t = lane; t < 54; t += TILE) Jr[t] = jrot_ref(a, t / 9, t % 9, S.w_rot);
  T.sync();
  if (lane < 9) {
    for (int cp = 0; cp < 9; cp++) {
      double s = 0.0;
#pragma unroll
      for (int r = 0; r < 6; r++) s = fma(Jr[r * 9 + cp], Jr[r * 9 + lane], s);
      D[cp * 12 + lane] += s;
    }
  }
  T.sync();
  {
    const int pa = lane >> 2, pb = lane & 3;
    double val = 0.0;
    const double* be = S.bedge + (size_t)i * k * 4;
    for (int s = 0; s < k; s++) {
      const double ba = pa < 3 ? S.w_reg * be[4 * s + pa] : S.w_reg;
      const double bb = pb < 3 ? S.w_reg * be[4 * s + pb] : S.w_reg;
      val = fma(ba, bb, val);
    }
    // in-edges: a free source has -w in its row, an excluded source contributes a static-side row -w: w^2 each on t_j
    if (pa == 3 && pb == 3) val = fma((double)(S.in_off[i + 1] - S.in_off[i]) * S.w_reg, S.w_reg, val);
    // constraints: per group the node's aggregated entry (setFromTriplets sums duplicates, Deform.cpp:374)
    int cur = -1; double sa = 0.0, sb = 0.0;
    for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {
      const int g = S.cin_grp[t];
      if (g != cur) { val = fma(sa, sb, val); sa = 0.0; sb = 0.0; cur = g; }
      sa += S.ccoef[(size_t)t * 4 + pa]; sb += S.ccoef[(size_t)t * 4 + pb];
    }
    val = fma(sa, sb, val);
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int ia = pa < 3 ? j + 3 * pa : 9 + j, ib = pb < 3 ? j + 3 * pb : 9 + j;
      D[ia * 12 + ib] += val;
    }
  }
  T.sync();
  for (int j = 0; j < 12; j++) {  // Cholesky, lower, in place
    if (lane == j) {
      double s = D[j * 12 + j];
      for (int t = 0; t < j; t++) s = fma(-D[j * 12 + t], D[j * 12 + t], s);
      D[j * 12 + j] = sqrt(s);
    }
    T.sync();
    if (lane > j && lane < 12) {
      double s = D[lane * 12 + j];
      for (int t = 0; t < j; t++) s = fma(-D[lane * 12 + t], D[j * 12 + t], s);
      D[lane * 12 + j] = s / D[j * 12 + j];
    }
    T.sync();
  }
  if (lane < 12) {  // inverse column for quad index `lane`
    const int rc = ref_index(lane);
    double y[12];
#pragma unroll
    for (int r = 0; r < 12; r++) {
      double s = (r == rc) ? 1.0 : 0.0;
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
    // y is indexed by reference index; store row-wise in quad order (matrix is symmetric)
#pragma unroll
    for (int qi = 0; qi < 12; qi++) {
      double v = 0.0;
      const int rr = ref_index(qi);
#pragma unroll
      for (int t = 0; t < 12; t++) if (t == rr) v = y[t];
      S.dinv[(size_t)i * 144 + (size_t)qi * 12 + lane] = v;
    }
  }
  T.sync();
}

template <int K>
__global__ void __launch_bounds__(SOLVE_THREADS, 3) k_solve(SolveDev S) {
  cg::grid_group grid = cg::this_grid();
  cg::thread_block block = cg::this_thread_block();
  cg::thread_block_tile<TILE> T = cg::tiled_partition<TILE>(block);
  __shared__ double s_D[TILES_PER_BLOCK][144 + 54];
  const int gquad = blockIdx.x * QUADS_PER_BLOCK + (threadIdx.x >> 2), nquads = gridDim.x * QUADS_PER_BLOCK;
  const int j = threadIdx.x & 3;
  const unsigned qmask = 0xFu << (threadIdx.x & 28);
  const int gtile = blockIdx.x * TILES_PER_BLOCK + threadIdx.x / TILE, ntiles = gridDim.x * TILES_PER_BLOCK;
  const int gthread = blockIdx.x * SOLVE_THREADS + threadIdx.x, nthreads = gridDim.x * SOLVE_THREADS;
  const int M = S.M, k = S.k;
  int phase = 0;
  double red[NRED];

  // ---- per-solve constants + x = identity (setIdentityRots, Deform.cpp:83-93)
  for (int t = gthread; t < M * 12; t += nthreads) {
    const int c = t & 3;
    const int jj = (t >> 2) % 3;
    S.x[t] = (c == jj) ? 1.0 : 0.0;   // A[j][c] = delta, t_j = 0
    S.h[t] = 0.0; S.p0[t] = 0.0; S.p1[t] = 0.0; S.z[t] = 0.0; S.r[t] = 0.0;
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

  for (int gn = 0; gn < S.max_gn; gn++) {
    gn_iters = gn + 1;
    if (!have_f) {
      red[0] = rows_res(S, gquad, nquads, j, S.x, nullptr, 0.0);
      red[1] = red[2] = 0.0;
      grid_reduce(grid, S, phase, red);
      E0 = red[0];
    }
    energy = E0;
    // ---- block-Jacobi preconditioner
    for (int i = gtile; i < M; i += ntiles)
      if (S.node_free[i]) build_dinv(S, T, i, s_D[threadIdx.x / TILE], s_D[threadIdx.x / TILE] + 144);
    __syncthreads();
    // ---- gradient g = -J^T f, z = Dinv g
    double rz_l = 0.0, gg_l = 0.0, xx_l = 0.0;
    for (int i = gquad; i < M; i += nquads) {
      if (!S.node_free[i]) continue;
      D4 g4{0.0, 0.0, 0.0, 0.0}, x4{0.0, 0.0, 0.0, 0.0};
      if (j < 3) {
        const size_t ob = (size_t)i * 12;
        const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
        double f[6]; rot_rows_res(A0, A1, A2, S.w_rot, f);
        x4 = j == 0 ? A0 : j == 1 ? A1 : A2;
        const D4 y = gather_jt<K>(S, i, j, f, x4.d);
        g4 = D4{-y.a, -y.b, -y.c, -y.d};
      }
      const D4 z4 = apply_dinv(S, i, j, qmask, g4);
      if (j < 3) {
        const size_t o = ((size_t)i * 3 + j) * 4;
        st4(S.r + o, g4); st4(S.z + o, z4); st4(S.h + o, D4{0.0, 0.0, 0.0, 0.0});
        rz_l += g4.a * z4.a + g4.b * z4.b + g4.c * z4.c + g4.d * z4.d;
        gg_l += g4.a * g4.a + g4.b * g4.b + g4.c * g4.c + g4.d * g4.d;
        xx_l += x4.a * x4.a + x4.b * x4.b + x4.c * x4.c + x4.d * x4.d;
      }
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
        red[0] = rows_lin<K>(S, gquad, nquads, j, S.z, pold, beta, pnew);   // p = z + beta p_old; u = J p
        red[1] = red[2] = 0.0;
        grid_reduce(grid, S, phase, red);
        const double pHp = red[0];
        const double alpha = rz / pHp;
        double rzn_l = 0.0, rr_l = 0.0;
        for (int i = gquad; i < M; i += nquads) {
          if (!S.node_free[i]) continue;
          D4 r4{0.0, 0.0, 0.0, 0.0};
          const size_t o = ((size_t)i * 3 + (j < 3 ? j : 0)) * 4;
          if (j < 3) {
            const size_t ob = (size_t)i * 12;
            const D4 A0 = ld4(S.x + ob), A1 = ld4(S.x + ob + 4), A2 = ld4(S.x + ob + 8);
            const D4 P0 = ld4(pnew + ob), P1 = ld4(pnew + ob + 4), P2 = ld4(pnew + ob + 8);
            double u[6]; rot_rows_lin(A0, A1, A2, P0, P1, P2, S.w_rot, u);
            const D4 p4 = j == 0 ? P0 : j == 1 ? P1 : P2;
            const D4 y = gather_jt<K>(S, i, j, u, p4.d);
            D4 h4 = ld4(S.h + o); r4 = ld4(S.r + o);
            h4.a = fma(alpha, p4.a, h4.a); h4.b = fma(alpha, p4.b, h4.b); h4.c = fma(alpha, p4.c, h4.c); h4.d = fma(alpha, p4.d, h4.d);
            r4.a = fma(-alpha, y.a, r4.a); r4.b = fma(-alpha, y.b, r4.b); r4.c = fma(-alpha, y.c, r4.c); r4.d = fma(-alpha, y.d, r4.d);
            st4(S.h + o, h4); st4(S.r + o, r4);
          }
          const D4 z4 = apply_dinv(S, i, j, qmask, r4);
          if (j < 3) {
            st4(S.z + o, z4);
            rzn_l += r4.a * z4.a + r4.b * z4.b + r4.c * z4.c + r4.d * z4.d;
            rr_l += r4.a * r4.a + r4.b * r4.b + r4.c * r4.c + r4.d * r4.d;
          }
        }
        red[0] = rzn_l; red[1] = rr_l; red[2] = 0.0;
        grid_reduce(grid, S, phase, red);
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

    // ---- step halving (Deform.cpp:144-156)
    bool accepted = false;
    for (double alpha_ls = 1.0; alpha_ls > 1e-15; alpha_ls *= 0.5) {
      red[0] = rows_res(S, gquad, nquads, j, S.x, S.h, 1.0);
      double hh_l = 0.0;
      for (int t = gthread; t < M * 12; t += nthreads) { const double hv = S.h[t]; hh_l = fma(hv, hv, hh_l); }
      red[1] = hh_l; red[2] = 0.0;
      grid_reduce(grid, S, phase, red);
      const double E1 = red[0];
      if (E1 > E0) {
        for (int t = gthread; t < M * 12; t += nthreads) S.h[t] *= 0.5;
        halvings++;
        normh = 0.5 * sqrt(red[1]);
        grid.sync();
      } else {
        for (int t = gthread; t < M * 12; t += nthreads) S.x[t] += S.h[t];
        normh = sqrt(red[1]);
        E0 = E1; have_f = true; accepted = true;
        grid.sync();
        break;
      }
    }
    if (!accepted) have_f = false;  // x unchanged, but the row buffers now hold f(x + h): recompute f(x)
    if (normh < (normv + 1e-6) * 1e-6) break;
  }

  // putFreeInputs (Deform.hpp:140-151): rot column-major, excluded nodes keep identity (h, x of excluded nodes never change)
  for (int t = gthread; t < M * 12; t += nthreads) {
    const int i = t / 12, qi = t - 12 * i, jj = qi >> 2, c = qi & 3;
    const double v = S.x[t];
    if (c < 3) S.rot_out[(size_t)i * 9 + jj + 3 * c] = v; else S.trans_out[(size_t)i * 3 + jj] = v;
  }
  if (gthread == 0) {
    S.stats[0] = gn_iters; S.stats[1] = energy; S.stats[2] = halvings; S.stats[3] = normh;
    S.stats[4] = total_cg; S.stats[5] = last_rel; S.stats[6] = flag;
  }
}

}  // namespace arapgs

using namespace arapgs;

extern "C" size_t arapk_solve_workspace_bytes(int M, int k, int n_groups) {
  // 6 vectors + dinv + edge constants + row buffers + constraint coefficients (<= groups*20*k entries) + partials
  size_t d = (size_t)M * 12 * 6 + (size_t)M * 144 + (size_t)M * k * 4 + (size_t)M * k * 3 + (size_t)(n_groups + 1) * 3 +
             (size_t)(n_groups + 1) * 20 * k * 4 + 2 * 2048 * NRED + 64;
  return d * sizeof(double);
}

extern "C" int arapk_solve(const ArapSolveGraph* G, const ArapSolveParams* P, void* workspace, size_t workspace_bytes,
                           double* rot_out, double* trans_out, double* stats_dev, cudaStream_t st) {
  if (G->M < 1 || G->k < 1 || G->k > KNN_MAX) { set_error("solve: bad graph"); return ARAP_ERR_INVALID; }
  if (workspace_bytes < arapk_solve_workspace_bytes(G->M, G->k, G->n_groups)) { set_error("solve: workspace too small"); return ARAP_ERR_INVALID; }
  if (G->n_cin_entries > (long long)(G->n_groups + 1) * 20 * G->k) { set_error("solve: constraint entry count exceeds the workspace bound"); return ARAP_ERR_INVALID; }
  SolveDev S;
  S.M = G->M; S.k = G->k; S.n_groups = G->n_groups; S.n_entries = (int)G->n_cin_entries;
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
  S.bedge = w; w += (size_t)G->M * G->k * 4;
  S.ccoef = w; w += (size_t)(G->n_groups + 1) * 20 * G->k * 4;  // 16-byte aligned arrays (double2 loads) first
  S.u_reg = w; w += (size_t)G->M * G->k * 3;
  S.u_con = w; w += (size_t)(G->n_groups + 1) * 3;
  S.partial = w;
  S.rot_out = rot_out; S.trans_out = trans_out; S.stats = stats_dev;
  int dev = 0, sms = 0, per_sm = 0;
  ARAP_CUDA_TRY(cudaGetDevice(&dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void* kern = G->k == 8 ? (void*)k_solve<8> : G->k == 10 ? (void*)k_solve<10> : G->k == 12 ? (void*)k_solve<12> : (void*)k_solve<0>;
  ARAP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, SOLVE_THREADS, 0));
  if (per_sm < 1) { set_error("solve: kernel cannot be co-resident"); return ARAP_ERR_CUDA; }
  const int items = G->M + G->n_groups;
  int grid = std::min(sms * std::min(per_sm, 4), (items + QUADS_PER_BLOCK - 1) / QUADS_PER_BLOCK);
  grid = std::max(1, std::min(grid, 2048));
  void* args[] = {(void*)&S};
  ARAP_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(SOLVE_THREADS), args, 0, st));
  return ARAP_OK;
}
