// Stage (c), fast path: the Gauss-Newton / Jacobi-PCG solve of solve.cu with the per-node state resident in
// shared memory for the whole solve.
//
// Same mathematics, same phases and the same two grid barriers per PCG iteration as solve.cu (see its header for
// the reference citations).  What changes is where the data lives: nodes are dealt round-robin to the CTAs
// (owner(i) = i % gridDim, one CTA per SM); a CTA keeps x, r, z, p, h, 1/diag, its own row values u, the per-edge
// constants and the graph slices of ITS nodes in shared memory.  Only what other CTAs need crosses L2:
//   z, p (the search-direction pieces neighbours recombine as p = z + beta p_old), x + h for residual evaluations,
//   u_in (row values delivered to the destination node's in-edge slots) and u_con (constraint rows).
// That cuts the per-iteration L2 traffic from ~45 MB to ~12 MB at 16k nodes and removes most dependent global loads
// from both phases, which were latency-bound.  FPS node order is spatially random, so round-robin ownership also
// spreads the control-region nodes (extra constraint work) evenly.
//
// Used when the per-CTA slice fits in shared memory (<= ~27k nodes at k = 10 on 148 SMs); otherwise solve.cu runs.
#include <cooperative_groups.h>
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "common.cuh"
#include "kernels.h"
#include "solve_dev.h"

namespace cg = cooperative_groups;

namespace arapgs {

#ifndef ARAP_SM_THREADS
#define ARAP_SM_THREADS 512
#endif
constexpr int SM_THREADS = ARAP_SM_THREADS;
constexpr int SM_NRED = 3;
constexpr int SM_RB = 4;        // E_reg rows a thread has in flight (one batch covers 119 nodes x k = 10 at 512 threads)
constexpr int SM_GB = 3;        // unknowns a thread has in flight in the gather phase (one batch covers 128 nodes)
constexpr int SM_GMAXG = 160;   // groups per CTA the shared-memory row table can index

struct D4 { double a, b, c, d; };
__device__ __forceinline__ D4 ld4(const double* p) {
  const double2 lo = *reinterpret_cast<const double2*>(p), hi = *reinterpret_cast<const double2*>(p + 2);
  return D4{lo.x, lo.y, hi.x, hi.y};
}
__device__ __forceinline__ D4 ld4cg(const double* p) {  // data published by other CTAs: read through L2
  const double2 lo = __ldcg(reinterpret_cast<const double2*>(p)), hi = __ldcg(reinterpret_cast<const double2*>(p + 2));
  return D4{lo.x, lo.y, hi.x, hi.y};
}
__device__ __forceinline__ void st4(double* p, const D4& v) {
  *reinterpret_cast<double2*>(p) = make_double2(v.a, v.b);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v.c, v.d);
}
__device__ __forceinline__ unsigned long long gtime2() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__device__ __forceinline__ void rot_lin(const D4& A0, const D4& A1, const D4& A2, const D4& P0, const D4& P1, const D4& P2, double w, double (&u)[6]) {
  const double a0p1 = fma(A0.a, P0.b, fma(A1.a, P1.b, A2.a * P2.b)), a1p0 = fma(A0.b, P0.a, fma(A1.b, P1.a, A2.b * P2.a));
  const double a0p2 = fma(A0.a, P0.c, fma(A1.a, P1.c, A2.a * P2.c)), a2p0 = fma(A0.c, P0.a, fma(A1.c, P1.a, A2.c * P2.a));
  const double a1p2 = fma(A0.b, P0.c, fma(A1.b, P1.c, A2.b * P2.c)), a2p1 = fma(A0.c, P0.b, fma(A1.c, P1.b, A2.c * P2.b));
  u[0] = w * (a0p1 + a1p0); u[1] = w * (a0p2 + a2p0); u[2] = w * (a1p2 + a2p1);
  u[3] = 2.0 * w * fma(A0.a, P0.a, fma(A1.a, P1.a, A2.a * P2.a));
  u[4] = 2.0 * w * fma(A0.b, P0.b, fma(A1.b, P1.b, A2.b * P2.b));
  u[5] = 2.0 * w * fma(A0.c, P0.c, fma(A1.c, P1.c, A2.c * P2.c));
}
__device__ __forceinline__ void rot_res(const D4& A0, const D4& A1, const D4& A2, double w, double (&f)[6]) {
  f[0] = w * fma(A0.a, A0.b, fma(A1.a, A1.b, A2.a * A2.b));
  f[1] = w * fma(A0.a, A0.c, fma(A1.a, A1.c, A2.a * A2.c));
  f[2] = w * fma(A0.b, A0.c, fma(A1.b, A1.c, A2.b * A2.c));
  f[3] = w * (fma(A0.a, A0.a, fma(A1.a, A1.a, A2.a * A2.a)) - 1.0);
  f[4] = w * (fma(A0.b, A0.b, fma(A1.b, A1.b, A2.b * A2.b)) - 1.0);
  f[5] = w * (fma(A0.c, A0.c, fma(A1.c, A1.c, A2.c * A2.c)) - 1.0);
}
__device__ __forceinline__ double rot_t(const D4& Aj, double w, const double (&u)[6], int c) {
  if (c == 0) return w * fma(u[0], Aj.b, fma(u[1], Aj.c, 2.0 * u[3] * Aj.a));
  if (c == 1) return w * fma(u[0], Aj.a, fma(u[2], Aj.c, 2.0 * u[4] * Aj.b));
  return w * fma(u[1], Aj.a, fma(u[2], Aj.b, 2.0 * u[5] * Aj.c));
}

// shared-memory slice of one CTA
struct Loc {
  double *xs, *rs, *zs, *ps, *hs, *ds, *us;  // [NL*12] x6, [NL*K*3]
  double* uro;                               // [NL*6] E_rot rows of J p (row phase -> gather phase)
  float4* be;                                // [NL*K]  (g_q - g_i as float, 1)
  int *nbr, *o2i;                            // [NL*K]
  int *inb, *ine, *cb, *ce, *sic, *fr;       // [NL]
  // shared-memory copies of the constraint tables of this CTA (valid when use_g / use_c): every remote gather of a
  // phase can then be issued from shared-memory indices alone, i.e. in one L2 round
  int* lcb;                                  // [NL] local start of a node's constraint entries
  int* goff;                                 // [SM_GMAXG + 1] entry offsets of the CTA's groups (g = lg * B + b)
  int* gq; double* gc;                       // [gcap], [gcap * 4]  row entries: node (-1 = excluded), w_con wei (v_c - g_q, 1)
  double* cpart;                             // [gcap * 3] per-entry products of the constraint rows
  int* cg; double* cc;                       // [ccap], [ccap * 4]  column entries of the CTA's nodes: group, coefficients
  int use_g, use_c, ng, ngent;
  int b, B, nloc;
};

// Grid barrier + deterministic reduction of the first NV of SM_NRED scalars, flag-in-data style (the scheme of NCCL's LL
// protocol): every CTA publishes its partial sums as 8-byte words {stamp : value half}, 64 bytes per CTA, and then
// spins on ALL CTAs' words until they carry this barrier's stamp.  No atomic, and stamp + data arrive in the same 8-byte
// single-copy-atomic access, so the barrier costs one store propagation plus one L2 read round instead of atomic ->
// poll -> read partials.  Two slot sets (barrier parity): a CTA can run at most one barrier ahead of the slowest
// reader.  Slots are zeroed by the launcher (stamps start at 1).
constexpr int LL_WORDS = 8;   // u64 words per CTA slot
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
constexpr int LL_GATHER = 160;   // threads of CTA 0 that poll the slots (5 warps: one slot each at 148 CTAs)
struct BarrierSmem { double part[SM_THREADS / 32][SM_NRED], tot[SM_NRED], slot[LL_GATHER][SM_NRED]; };
__device__ __forceinline__ BarrierSmem& barrier_smem() {   // one instance for all NV instantiations
  __shared__ BarrierSmem bs;
  return bs;
}
template <int NV>
__device__ __forceinline__ void barrier_reduce(const SolveDev& S, unsigned* counter, int& phase, double (&v)[SM_NRED]) {
  BarrierSmem& bs = barrier_smem();
  double (&s_part)[SM_THREADS / 32][SM_NRED] = bs.part;
  double (&s_tot)[SM_NRED] = bs.tot;
  double (&s_slot)[LL_GATHER][SM_NRED] = bs.slot;
#pragma unroll
  for (int q = 0; q < NV; q++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int q = 0; q < NV; q++) s_part[warp][q] = v[q];
  __syncthreads();
  unsigned long long* set = reinterpret_cast<unsigned long long*>(counter) + (size_t)(phase & 1) * gridDim.x * LL_WORDS;
  unsigned long long* res = reinterpret_cast<unsigned long long*>(counter) + (size_t)2 * gridDim.x * LL_WORDS + (size_t)(phase & 1) * LL_WORDS;
  const unsigned long long stamp = (unsigned long long)(unsigned)(phase + 1) << 32;
  if (warp == 0) {
    double a[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) {
      a[q] = lane < SM_THREADS / 32 ? s_part[lane][q] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    }
    if (lane == 0) {
      __threadfence();   // the CTA's published vectors (ordered before by the __syncthreads above) precede the stamp
      unsigned long long* mine = set + (size_t)blockIdx.x * LL_WORDS;
      mine[7] = gtime2();   // arrival time (diagnostics: barrier skew)
#pragma unroll
      for (int q = 0; q < NV; q++) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(a[q]);
        ll_store(mine + 2 * q, stamp | (bits & 0xffffffffull));
        ll_store(mine + 2 * q + 1, stamp | (bits >> 32));
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < LL_GATHER) {   // the gatherer: O(B) polling instead of every CTA reading every slot
    double a[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) a[q] = 0.0;
    for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += LL_GATHER) {
      const unsigned long long* src = set + (size_t)bb * LL_WORDS;
      unsigned long long wv[2 * NV];
      bool ok;
      do {
        ok = true;
#pragma unroll
        for (int t = 0; t < 2 * NV; t++) { wv[t] = ll_load(src + t); ok = ok && ((wv[t] & 0xffffffff00000000ull) == stamp); }
      } while (!ok);
#pragma unroll
      for (int q = 0; q < NV; q++)
        a[q] += __longlong_as_double((long long)((wv[2 * q] & 0xffffffffull) | (wv[2 * q + 1] << 32)));
    }
#pragma unroll
    for (int q = 0; q < NV; q++) s_slot[threadIdx.x][q] = a[q];
    asm volatile("bar.sync 1, %0;" ::"n"(LL_GATHER) : "memory");
    if (warp == 0) {   // fixed order: deterministic sums
#pragma unroll
      for (int q = 0; q < NV; q++) {
        double t = 0.0;
#pragma unroll
        for (int r = 0; r < LL_GATHER / 32; r++) t += s_slot[lane + 32 * r][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        a[q] = t;
      }
      if (lane == 0) {
        __threadfence();
#pragma unroll
        for (int q = 0; q < NV; q++) {
          const unsigned long long bits = (unsigned long long)__double_as_longlong(a[q]);
          ll_store(res + 2 * q, stamp | (bits & 0xffffffffull));
          ll_store(res + 2 * q + 1, stamp | (bits >> 32));
        }
      }
    }
  }
  if (warp == 0) {
    unsigned long long rv = 0;
    if (lane < 2 * NV) {
      do { rv = ll_load(res + lane); } while ((rv & 0xffffffff00000000ull) != stamp);
    }
    __threadfence();
#pragma unroll
    for (int q = 0; q < NV; q++) {
      const unsigned long long lo = __shfl_sync(0xffffffffu, rv, 2 * q), hi = __shfl_sync(0xffffffffu, rv, 2 * q + 1);
      if (lane == 0) s_tot[q] = __longlong_as_double((long long)((lo & 0xffffffffull) | (hi << 32)));
    }
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NV; q++) v[q] = s_tot[q];
  __syncthreads();
  phase++;
}

// Published (global) vectors keep each node as [A row-major (9) | t (3)], 96 bytes: the three t components that the
// E_reg rows of a neighbour need sit in ONE 32-byte sector.  Shared memory keeps rows as (A_j0, A_j1, A_j2, t_j).
__device__ __forceinline__ int pub(int qi) { const int j = qi >> 2, c = qi & 3; return c < 3 ? 3 * j + c : 9 + j; }

// remote vector value va[o] + sc*vb[o] (published arrays, through L2)
__device__ __forceinline__ double rcomb1(const double* va, const double* vb, double sc, size_t o) {
  return vb ? fma(sc, __ldcg(vb + o), __ldcg(va + o)) : __ldcg(va + o);
}
// row j of node q: (A_j0, A_j1, A_j2, t_j)
__device__ __forceinline__ D4 rcomb4(const double* va, const double* vb, double sc, int q, int j) {
  const double* a = va + (size_t)q * 12;
  D4 v{__ldcg(a + 3 * j), __ldcg(a + 3 * j + 1), __ldcg(a + 3 * j + 2), __ldcg(a + 9 + j)};
  if (vb) {
    const double* b = vb + (size_t)q * 12;
    const D4 w{__ldcg(b + 3 * j), __ldcg(b + 3 * j + 1), __ldcg(b + 3 * j + 2), __ldcg(b + 9 + j)};
    v.a = fma(sc, w.a, v.a); v.b = fma(sc, w.b, v.b); v.c = fma(sc, w.c, v.c); v.d = fma(sc, w.d, v.d);
  }
  return v;
}

// Row phase over this CTA's nodes (+ its share of the constraint groups).
//   MODE 1: u = J v; the CTA's own v is in L.ps (already formed); remote v = ga + sc*gb (published z, p_old).
//   MODE 0: f(v) nonlinear residual; own v in `own` (shared, already formed); remote v = ga (published x + h).
template <int K, int MODE>
__device__ __forceinline__ double rows_smem(const SolveDev& S, const Loc& L, const double* own, const double* ga, const double* gb, double sc,
                                            unsigned long long* tmark = nullptr, unsigned long long* wst = nullptr) {
  const int k = K;
  const int k3 = 3 * k;
  double sq = 0.0;
  // E_reg rows, SM_RB per thread at a time: all remote loads of a batch are issued before any is consumed (the phase is
  // bound by L2 round trips, not bandwidth).  The neighbour's free flag rides in the sign bit of L.nbr.
  for (int t0 = threadIdx.x; t0 < L.nloc * k3; t0 += SM_RB * SM_THREADS) {
    double ta[SM_RB], tb[SM_RB]; int qf[SM_RB];
#pragma unroll
    for (int r = 0; r < SM_RB; r++) {
      const int t = t0 + r * SM_THREADS;
      qf[r] = -1; ta[r] = 0.0; tb[r] = 0.0;
      if (t < L.nloc * k3) {
        const int li = t / k3, rem = t - li * k3, s = rem / 3, j = rem - 3 * s;
        if (L.fr[li]) {
          qf[r] = L.nbr[li * k + s];
          if (qf[r] >= 0) {
            const size_t o = (size_t)qf[r] * 12 + 9 + j;
            ta[r] = __ldcg(ga + o);
            if (gb) tb[r] = __ldcg(gb + o);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < SM_RB; r++) {
      const int t = t0 + r * SM_THREADS;
      if (t >= L.nloc * k3) continue;
      const int li = t / k3, rem = t - li * k3, s = rem / 3, j = rem - 3 * s;
      if (!L.fr[li]) continue;
      const int e = li * k + s;
      const D4 v = ld4(own + (size_t)li * 12 + 4 * j);
      const double tq = qf[r] >= 0 ? fma(sc, tb[r], ta[r]) : 0.0;
      double val;
      if (MODE == 1) {
        const float4 b = L.be[e];
        val = S.w_reg * ((fma(v.c, (double)b.z, fma(v.b, (double)b.y, v.a * (double)b.x)) + v.d) - tq);
      } else {
        const int i = li * L.B + L.b, q = qf[r] & 0x7fffffff;
        const double gi0 = S.node_pos[3 * i], gi1 = S.node_pos[3 * i + 1], gi2 = S.node_pos[3 * i + 2];
        const double gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
        const double gij = j == 0 ? gi0 : j == 1 ? gi1 : gi2, gqj = j == 0 ? gq0 : j == 1 ? gq1 : gq2;
        val = S.w_reg * ((((fma(v.c, gq2 - gi2, fma(v.b, gq1 - gi1, v.a * (gq0 - gi0))) + gij) + v.d) - gqj) - tq);
      }
      L.us[(size_t)e * 3 + j] = val;
      const int slot = L.o2i[e];
      if (slot >= 0) S.u_in[(size_t)slot * 3 + j] = val;
      sq = fma(val, val, sq);
      if (s == 0) {
        const double sv = S.w_reg * v.d;
        sq = fma((double)L.sic[li] * sv, sv, sq);
      }
    }
  }
  if (tmark) tmark[0] = gtime2();
  if (wst && (threadIdx.x & 31) == 0) wst[(threadIdx.x >> 5) * 6 + 0] = gtime2();
  // E_rot rows: one thread per local node, dealt from the end of the CTA
  for (int li = SM_THREADS - 1 - (int)threadIdx.x; li < L.nloc; li += SM_THREADS) {
    if (!L.fr[li]) continue;
    const size_t ob = (size_t)li * 12;
    double u[6];
    const D4 V0 = ld4(own + ob), V1 = ld4(own + ob + 4), V2 = ld4(own + ob + 8);
    if (MODE == 1) {
      const D4 A0 = ld4(L.xs + ob), A1 = ld4(L.xs + ob + 4), A2 = ld4(L.xs + ob + 8);
      rot_lin(A0, A1, A2, V0, V1, V2, S.w_rot, u);
#pragma unroll
      for (int t = 0; t < 6; t++) L.uro[li * 6 + t] = u[t];
    } else rot_res(V0, V1, V2, S.w_rot, u);
#pragma unroll
    for (int t = 0; t < 6; t++) sq = fma(u[t], u[t], sq);
  }
  if (tmark) tmark[1] = gtime2();
  if (wst && (threadIdx.x & 31) == 0) wst[(threadIdx.x >> 5) * 6 + 1] = gtime2();
  // constraint rows, shared-memory tables: (entry, j) per thread.  The gathers are issued here and consumed after the
  // E_rot rows.
  const bool csm = MODE == 1 && L.use_g;
  const int ne3 = csm ? L.ngent * 3 : 0;
  double cra[4], crb[4]; int ce0 = -1, cj0 = 0;
  if (csm && (int)threadIdx.x < ne3) {
    const int e = threadIdx.x / 3; cj0 = threadIdx.x - 3 * e;
    const int q = L.gq[e];
    ce0 = e;
    if (q >= 0) {
      const double* a = ga + (size_t)q * 12; const double* bq = gb + (size_t)q * 12;
      cra[0] = __ldcg(a + 3 * cj0); cra[1] = __ldcg(a + 3 * cj0 + 1); cra[2] = __ldcg(a + 3 * cj0 + 2); cra[3] = __ldcg(a + 9 + cj0);
      crb[0] = __ldcg(bq + 3 * cj0); crb[1] = __ldcg(bq + 3 * cj0 + 1); crb[2] = __ldcg(bq + 3 * cj0 + 2); crb[3] = __ldcg(bq + 9 + cj0);
    } else { ce0 = -2 - e; }
  }
  if (csm) {
    if (ce0 != -1) {
      double acc = 0.0;
      const int e = ce0 >= 0 ? ce0 : -2 - ce0;
      if (ce0 >= 0) {
        const D4 c = ld4(L.gc + (size_t)e * 4);
        acc = fma(c.c, fma(sc, crb[2], cra[2]), fma(c.b, fma(sc, crb[1], cra[1]), fma(c.a, fma(sc, crb[0], cra[0]), c.d * fma(sc, crb[3], cra[3]))));
      }
      L.cpart[e * 3 + cj0] = acc;
    }
    for (int t = threadIdx.x + SM_THREADS; t < ne3; t += SM_THREADS) {   // more entries than threads: not overlapped
      const int e = t / 3, j = t - 3 * e;
      const int q = L.gq[e];
      double acc = 0.0;
      if (q >= 0) {
        const D4 c = ld4(L.gc + (size_t)e * 4);
        const D4 v = rcomb4(ga, gb, sc, q, j);
        acc = fma(c.c, v.c, fma(c.b, v.b, fma(c.a, v.a, c.d * v.d)));
      }
      L.cpart[t] = acc;
    }
    if (wst && (threadIdx.x & 31) == 0) wst[(threadIdx.x >> 5) * 6 + 2] = gtime2();
    __syncthreads();
    if (wst && (threadIdx.x & 31) == 0) wst[(threadIdx.x >> 5) * 6 + 3] = gtime2();
    const int l16 = threadIdx.x & 15;
    const unsigned tmask = 0xFFFFu << (threadIdx.x & 16);
    for (int t = (threadIdx.x >> 4); t < 3 * L.ng; t += (SM_THREADS >> 4)) {
      const int lg = t / 3, j = t - 3 * lg;
      double acc = 0.0;
      for (int e = L.goff[lg] + l16; e < L.goff[lg + 1]; e += 16) acc += L.cpart[e * 3 + j];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(tmask, acc, o, 16);
      if (l16 == 0) {
        S.u_con[(size_t)(lg * L.B + L.b) * 3 + j] = acc;
        sq = fma(acc, acc, sq);
      }
    }
  } else
  // constraint rows: 16-lane team per (group, component); groups dealt round-robin to CTAs
  {
    const int l16 = threadIdx.x & 15;
    const unsigned tmask = 0xFFFFu << (threadIdx.x & 16);
    const int nteams = SM_THREADS >> 4;
    const int ng = S.n_groups > L.b ? (S.n_groups - L.b + L.B - 1) / L.B : 0;   // groups g = lg*B + b
    for (int t = (threadIdx.x >> 4); t < 3 * ng; t += nteams) {
      const int lg = t / 3, j = t - 3 * lg;
      const int g = lg * L.B + L.b;
      const int mb = S.grp_off[g], me = S.grp_off[g + 1];
      double acc = 0.0;
      if (MODE == 1) {
        // precomputed entries (node or -1, w_con wei (v_c - g_q, 1)): one independent load level before the gather
        for (int m0 = mb; m0 < me; m0 += 4) {
          int qq[4]; D4 cc[4];
#pragma unroll
          for (int r = 0; r < 4; r++) {
            qq[r] = -1;
            if (m0 + r < me && l16 < k) {
              const size_t id = (size_t)(m0 + r) * k + l16;
              qq[r] = S.gent_q[id];
              cc[r] = ld4(S.gent_c + id * 4);
            }
          }
#pragma unroll
          for (int r = 0; r < 4; r++) {
            if (qq[r] < 0) continue;
            const D4 v = rcomb4(ga, gb, sc, qq[r], j);
            acc = fma(cc[r].c, v.c, fma(cc[r].b, v.b, fma(cc[r].a, v.a, fma(cc[r].d, v.d, acc))));
          }
        }
      } else {
      for (int m = mb; m < me; m++) {
        const int c = S.grp_member[m];
        if (l16 < k) {
          const int q = S.anc_idx[c * k + l16];
          const double wei = S.anc_w[c * k + l16];
          const float vc0 = S.node_pos[3 * c], vc1 = S.node_pos[3 * c + 1], vc2 = S.node_pos[3 * c + 2];
          if (!S.node_free[q]) {
            acc = fma(wei, j == 0 ? (double)vc0 : j == 1 ? (double)vc1 : (double)vc2, acc);
          } else {
            const float gq0 = S.node_pos[3 * q], gq1 = S.node_pos[3 * q + 1], gq2 = S.node_pos[3 * q + 2];
            const D4 v = rcomb4(ga, gb, sc, q, j);
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
        S.u_con[(size_t)g * 3 + j] = val;
        sq = fma(val, val, sq);
      }
    }
  }
  return sq;
}

// (J^T u) for local unknown (li, j, c)
template <int K>
__device__ __forceinline__ double gather_smem(const SolveDev& S, const Loc& L, int li, int j, int c, const D4& Aj, const double (&urot)[6], double vt) {
  const int k = K;
  const double* ur = L.us + (size_t)li * k * 3 + j;
  double y;
  if (c < 3) {
    y = rot_t(Aj, S.w_rot, urot, c);
    const float* be = reinterpret_cast<const float*>(L.be + (size_t)li * k) + c;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < K; s++) acc = fma((double)be[4 * s], ur[3 * s], acc);
    y = fma(S.w_reg, acc, y);
  } else {
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < K; s++) acc += ur[3 * s];
    const double* ui = S.u_in + j;
    double acc2 = 0.0;
    const int ib = L.inb[li], ie = L.ine[li];
    for (int t0 = ib; t0 < ie; t0 += 8) {
      double uu[8];
#pragma unroll
      for (int t = 0; t < 8; t++) uu[t] = (t0 + t < ie) ? __ldcg(ui + (size_t)(t0 + t) * 3) : 0.0;
#pragma unroll
      for (int t = 0; t < 8; t++) acc2 += uu[t];
    }
    y = S.w_reg * (acc - acc2);
    y = fma((double)L.sic[li] * S.w_reg * S.w_reg, vt, y);
  }
  const int cb = L.cb[li], ce = L.ce[li];
  for (int t0 = cb; t0 < ce; t0 += 4) {
    double cc[4], uu[4]; int gg[4];
#pragma unroll
    for (int t = 0; t < 4; t++) { const bool ok = t0 + t < ce; gg[t] = ok ? S.cin_grp[t0 + t] : -1; cc[t] = ok ? S.ccoef[(size_t)(t0 + t) * 4 + c] : 0.0; }
#pragma unroll
    for (int t = 0; t < 4; t++) uu[t] = gg[t] >= 0 ? __ldcg(S.u_con + (size_t)gg[t] * 3 + j) : 0.0;
#pragma unroll
    for (int t = 0; t < 4; t++) y = fma(cc[t], uu[t], y);
  }
  return y;
}

// (J^T u) for local unknown (li, j, c) inside the PCG loop, in two steps so that a thread can put the remote gathers of
// all its unknowns in flight before consuming any (the phase is bound by L2 round trips).  The four lanes (c = 0..3) of
// a (node, j) quad split the gathers (in-edge rows, constraint rows); partial sums / values are exchanged with quad
// shuffles.  `act` is uniform over the quad.
struct GatherLd { double uu[4], u[4]; int g[4]; };
template <int K>
__device__ __forceinline__ void gather_issue(const SolveDev& S, const Loc& L, bool act, int li, int j, int c, GatherLd& G) {
#pragma unroll
  for (int r = 0; r < 4; r++) { G.uu[r] = 0.0; G.u[r] = 0.0; G.g[r] = -1; }
  if (!act) return;
  const double* ui = S.u_in + j;
  const int ib = L.inb[li], ie = L.ine[li];
#pragma unroll
  for (int r = 0; r < 4; r++) { const int t = ib + 4 * r + c; if (t < ie) G.uu[r] = __ldcg(ui + (size_t)t * 3); }
  const int cb = L.cb[li], ce = L.ce[li];
  if (cb < ce) {
    const int* cgp = L.use_c ? L.cg + (L.lcb[li] - cb) : S.cin_grp;
#pragma unroll
    for (int r = 0; r < 4; r++) { const int t = cb + 4 * r + c; if (t < ce) G.g[r] = cgp[t]; }
#pragma unroll
    for (int r = 0; r < 4; r++) if (G.g[r] >= 0) G.u[r] = __ldcg(S.u_con + (size_t)G.g[r] * 3 + j);
  }
}
template <int K>
__device__ __forceinline__ double gather_finish(const SolveDev& S, const Loc& L, bool act, int li, int j, int c, const D4& Aj, double vt,
                                                const GatherLd& G) {
  const unsigned qmask = 0xFu << (threadIdx.x & 28);
  if (!act) return 0.0;
  const double* ur = L.us + (size_t)li * K * 3 + j;
  double y = 0.0, own = 0.0;
  if (c < 3) {
    double u[6];
#pragma unroll
    for (int t = 0; t < 6; t++) u[t] = L.uro[li * 6 + t];
    y = rot_t(Aj, S.w_rot, u, c);
    const float* be = reinterpret_cast<const float*>(L.be + (size_t)li * K) + c;
#pragma unroll
    for (int s = 0; s < K; s++) own = fma((double)be[4 * s], ur[3 * s], own);
  } else {
#pragma unroll
    for (int s = 0; s < K; s++) own += ur[3 * s];
  }
  // in-edge rows, -w each: lane c took slots ib + c, ib + c + 4, ...
  const double* ui = S.u_in + j;
  const int ib = L.inb[li], ie = L.ine[li];
  double a2 = (G.uu[0] + G.uu[1]) + (G.uu[2] + G.uu[3]);
  for (int base = ib + 16; base < ie; base += 16) {   // in-degree > 16: rare
    double uu[4];
#pragma unroll
    for (int r = 0; r < 4; r++) { const int t = base + 4 * r + c; uu[r] = t < ie ? __ldcg(ui + (size_t)t * 3) : 0.0; }
    a2 += (uu[0] + uu[1]) + (uu[2] + uu[3]);
  }
  // constraint rows: lane c fetched u_con of entries cb + c, cb + c + 4, ...; it forms their products with all four
  // coefficients, and a quad reduce-scatter (3 exchanges) leaves every lane with the sum for its own component
  const int cb = L.cb[li], ce = L.ce[li];
  double yc = 0.0;
  if (cb < ce) {
    const int* cgp = L.use_c ? L.cg + (L.lcb[li] - cb) : S.cin_grp;
    const double* ccp = L.use_c ? L.cc + (ptrdiff_t)(L.lcb[li] - cb) * 4 : S.ccoef;
    double part[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int t = cb + 4 * r + c;
      if (t < ce) {
        const D4 cf = ld4(ccp + (ptrdiff_t)t * 4);
        part[0] = fma(cf.a, G.u[r], part[0]); part[1] = fma(cf.b, G.u[r], part[1]);
        part[2] = fma(cf.c, G.u[r], part[2]); part[3] = fma(cf.d, G.u[r], part[3]);
      }
    }
    for (int base = cb + 16; base < ce; base += 16) {   // more than 16 entries: not overlapped
      int g[4]; double u[4];
#pragma unroll
      for (int r = 0; r < 4; r++) { const int t = base + 4 * r + c; g[r] = t < ce ? cgp[t] : -1; }
#pragma unroll
      for (int r = 0; r < 4; r++) u[r] = g[r] >= 0 ? __ldcg(S.u_con + (size_t)g[r] * 3 + j) : 0.0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const int t = base + 4 * r + c;
        if (t < ce) {
          const D4 cf = ld4(ccp + (ptrdiff_t)t * 4);
          part[0] = fma(cf.a, u[r], part[0]); part[1] = fma(cf.b, u[r], part[1]);
          part[2] = fma(cf.c, u[r], part[2]); part[3] = fma(cf.d, u[r], part[3]);
        }
      }
    }
    const bool hi = (c & 2) != 0, odd = (c & 1) != 0;
    const double k0 = (hi ? part[2] : part[0]) + __shfl_xor_sync(qmask, hi ? part[0] : part[2], 2, 4);
    const double k1 = (hi ? part[3] : part[1]) + __shfl_xor_sync(qmask, hi ? part[1] : part[3], 2, 4);
    yc = (odd ? k1 : k0) + __shfl_xor_sync(qmask, odd ? k0 : k1, 1, 4);
  }
  a2 += __shfl_xor_sync(qmask, a2, 1, 4);
  a2 += __shfl_xor_sync(qmask, a2, 2, 4);
  if (c < 3) y = fma(S.w_reg, own, y);
  else {
    y = S.w_reg * (own - a2);
    y = fma((double)L.sic[li] * S.w_reg * S.w_reg, vt, y);
  }
  return y + yc;
}

template <int K>
__device__ __forceinline__ double diag_smem(const SolveDev& S, const Loc& L, int li, int j, int c, const D4& Aj) {
  double d;
  if (c < 3) {
    const double w2 = S.w_rot * S.w_rot;
    const double o1 = c == 0 ? Aj.b : Aj.a, o2 = c == 2 ? Aj.b : Aj.c, own = c == 0 ? Aj.a : c == 1 ? Aj.b : Aj.c;
    d = w2 * (o1 * o1 + o2 * o2 + 4.0 * own * own);
    const float* be = reinterpret_cast<const float*>(L.be + (size_t)li * K) + c;
    double acc = 0.0;
#pragma unroll
    for (int s = 0; s < K; s++) acc = fma((double)be[4 * s], (double)be[4 * s], acc);
    d = fma(S.w_reg * S.w_reg, acc, d);
  } else {
    d = S.w_reg * S.w_reg * (double)(K + (L.ine[li] - L.inb[li]) + L.sic[li]);
  }
  int cur = -1; double sa = 0.0;
  for (int t = L.cb[li]; t < L.ce[li]; t++) {
    const int g = S.cin_grp[t];
    if (g != cur) { d = fma(sa, sa, d); sa = 0.0; cur = g; }
    sa += S.ccoef[(size_t)t * 4 + c];
  }
  return fma(sa, sa, d);
}

// diagnostics: time from the last CTA's arrival at barrier `ph` to block 0's exit
__device__ __noinline__ void barrier_skew(unsigned* counter, int ph, unsigned long long t_exit, double* out) {
  const unsigned long long* set = reinterpret_cast<const unsigned long long*>(counter) + (size_t)(ph & 1) * gridDim.x * LL_WORDS;
  unsigned long long mx = 0;
  for (int bb = 0; bb < (int)gridDim.x; bb++) {
    const unsigned long long t = ll_load(set + (size_t)bb * LL_WORDS + 7);
    mx = t > mx ? t : mx;
  }
  out[0] += (double)(t_exit - mx);   // last arrival -> block 0's exit: the barrier mechanism itself
}

// Slice capacities are compile-time (NL nodes per CTA, GCAP / CCAP constraint-table entries): every shared-memory
// array then sits at a constant address and the address arithmetic that made up ~12% of the instructions disappears.
template <int K, int NL> struct SmCaps {
  // NL = 136: the slice of a solve that leaves ~30 SMs to a concurrent kernel (arap_params.solver_ctas) at 16k nodes
  static constexpr int G = NL <= 112 ? (K <= 10 ? 768 : 640) : NL <= 136 ? (K <= 10 ? 512 : 384) : (K <= 10 ? 192 : 0);
  static constexpr int C = NL <= 112 ? (K <= 10 ? 1280 : 1024) : NL <= 136 ? (K <= 10 ? 768 : 640) : (K <= 10 ? 320 : 0);
};
template <int K, int NL>
__global__ void __launch_bounds__(SM_THREADS, 1) k_solve_smem(SolveDev S, unsigned* counter) {
  constexpr int gcap = SmCaps<K, NL>::G, ccap = SmCaps<K, NL>::C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Loc L;
  {
    double* d = reinterpret_cast<double*>(smem_raw);
    L.xs = d; d += (size_t)NL * 12; L.rs = d; d += (size_t)NL * 12; L.zs = d; d += (size_t)NL * 12;
    L.ps = d; d += (size_t)NL * 12; L.hs = d; d += (size_t)NL * 12; L.ds = d; d += (size_t)NL * 12;
    L.us = d; d += (size_t)NL * K * 3;
    L.uro = d; d += (size_t)NL * 6;
    if ((NL * K * 3) & 1) d += 1;  // keep 16-byte alignment for the float4 array
    L.be = reinterpret_cast<float4*>(d);
    int* ip = reinterpret_cast<int*>(L.be + (size_t)NL * K);
    L.nbr = ip; ip += (size_t)NL * K; L.o2i = ip; ip += (size_t)NL * K;
    L.inb = ip; ip += NL; L.ine = ip; ip += NL; L.cb = ip; ip += NL; L.ce = ip; ip += NL; L.sic = ip; ip += NL; L.fr = ip; ip += NL;
    L.lcb = ip; ip += NL; L.goff = ip; ip += SM_GMAXG + 1; L.gq = ip; ip += gcap; L.cg = ip; ip += ccap;
    if ((ip - reinterpret_cast<int*>(smem_raw)) & 3) ip += 4 - ((ip - reinterpret_cast<int*>(smem_raw)) & 3);   // 16-byte alignment
    double* dd = reinterpret_cast<double*>(ip);
    L.gc = dd; dd += (size_t)gcap * 4; L.cc = dd; dd += (size_t)ccap * 4; L.cpart = dd;
  }
  const int M = S.M, B = gridDim.x, b = blockIdx.x, tid = threadIdx.x;
  L.b = b; L.B = B; L.nloc = M > b ? (M - b + B - 1) / B : 0;
  const int nloc = L.nloc, NU = nloc * 12;
  int phase = 0;
  double red[SM_NRED];

  // ---- slice set-up: graph meta, per-edge constants, x = identity
  for (int li = tid; li < nloc; li += SM_THREADS) {
    const int i = li * B + b;
    L.fr[li] = S.node_free[i]; L.inb[li] = S.in_off[i]; L.ine[li] = S.in_off[i + 1];
    L.cb[li] = S.cin_off[i]; L.ce[li] = S.cin_off[i + 1]; L.sic[li] = S.static_in_cnt[i];
    for (int t = S.cin_off[i]; t < S.cin_off[i + 1]; t++) {   // w_con wei (v_c - g_q, 1) (Deform.cpp:325-328)
      const int m = S.cin_member[t];
      const double wv = S.w_con * S.anc_w[m * K + S.cin_slot[t]];
      double* c = S.ccoef + (size_t)t * 4;
      c[0] = wv * (double)(S.node_pos[3 * m] - S.node_pos[3 * i]);
      c[1] = wv * (double)(S.node_pos[3 * m + 1] - S.node_pos[3 * i + 1]);
      c[2] = wv * (double)(S.node_pos[3 * m + 2] - S.node_pos[3 * i + 2]);
      c[3] = wv;
    }
  }
  for (int e = tid; e < nloc * K; e += SM_THREADS) {
    const int li = e / K, s = e - li * K, i = li * B + b;
    const int q = S.nbr[i * K + s];
    L.nbr[e] = S.node_free[q] ? q : (q | (int)0x80000000); L.o2i[e] = S.out_to_in[i * K + s];
    L.be[e] = make_float4(S.node_pos[3 * q] - S.node_pos[3 * i], S.node_pos[3 * q + 1] - S.node_pos[3 * i + 1],
                          S.node_pos[3 * q + 2] - S.node_pos[3 * i + 2], 1.0f);   // float differences (Deform.cpp:254-256)
  }
  {  // constraint-row entries, shared by all CTAs: (member m, slot s) -> node (or -1 if excluded), w_con wei (v_c - g_q, 1)
    const int n_mem = S.n_groups > 0 ? S.grp_off[S.n_groups] : 0;
    for (int id = b * SM_THREADS + tid; id < n_mem * K; id += B * SM_THREADS) {
      const int m = id / K, sl = id - m * K, c = S.grp_member[m];
      const int q = S.anc_idx[c * K + sl];
      const double wv = S.w_con * S.anc_w[c * K + sl];
      S.gent_q[id] = S.node_free[q] ? q : -1;
      double* gc = S.gent_c + (size_t)id * 4;
      gc[0] = wv * (double)(S.node_pos[3 * c] - S.node_pos[3 * q]);
      gc[1] = wv * (double)(S.node_pos[3 * c + 1] - S.node_pos[3 * q + 1]);
      gc[2] = wv * (double)(S.node_pos[3 * c + 2] - S.node_pos[3 * q + 2]);
      gc[3] = wv;
    }
  }
  for (int t = tid; t < NU; t += SM_THREADS) {
    const int li = t / 12, qi = t - 12 * li, c = qi & 3, jj = qi >> 2;
    const size_t go = (size_t)(li * B + b) * 12 + pub(qi);
    const double xv = (c == jj) ? 1.0 : 0.0;
    L.xs[t] = xv; L.hs[t] = 0.0; L.ps[t] = 0.0; L.rs[t] = 0.0; L.zs[t] = 0.0; L.ds[t] = 0.0;
    S.x[go] = xv; S.z[go] = 0.0; S.p0[go] = 0.0; S.p1[go] = 0.0;   // S.x doubles as the published x + h
  }
  {  // shared-memory copies of this CTA's constraint tables
    __shared__ int s_flag[4];
    __syncthreads();
    if (tid == 0) {
      const int ng = S.n_groups > b ? (S.n_groups - b + B - 1) / B : 0;
      int off = 0;
      const bool fits = ng <= SM_GMAXG;
      for (int lg = 0; lg < ng && fits; lg++) { L.goff[lg] = off; off += (S.grp_off[lg * B + b + 1] - S.grp_off[lg * B + b]) * K; }
      if (fits) L.goff[ng] = off;
      s_flag[0] = fits && off <= gcap; s_flag[1] = ng; s_flag[2] = off;
    }
    if (tid == 32) {
      int off = 0;
      for (int li = 0; li < nloc; li++) { L.lcb[li] = off; off += L.ce[li] - L.cb[li]; }
      s_flag[3] = off <= ccap;
    }
    __syncthreads();
    L.use_g = s_flag[0]; L.ng = s_flag[1]; L.ngent = s_flag[2]; L.use_c = s_flag[3];
    if (L.use_g)
      for (int lg = 0; lg < L.ng; lg++) {
        const int base = L.goff[lg], len = L.goff[lg + 1] - base, mb = S.grp_off[lg * B + b];
        for (int t = tid; t < len; t += SM_THREADS) {
          const int m = mb + t / K, sl = t - (t / K) * K, c = S.grp_member[m];
          const int q = S.anc_idx[c * K + sl];
          const double wv = S.w_con * S.anc_w[c * K + sl];
          L.gq[base + t] = S.node_free[q] ? q : -1;
          double* gc = L.gc + (size_t)(base + t) * 4;
          gc[0] = wv * (double)(S.node_pos[3 * c] - S.node_pos[3 * q]);
          gc[1] = wv * (double)(S.node_pos[3 * c + 1] - S.node_pos[3 * q + 1]);
          gc[2] = wv * (double)(S.node_pos[3 * c + 2] - S.node_pos[3 * q + 2]);
          gc[3] = wv;
        }
      }
    if (L.use_c)
      for (int li = tid; li < nloc; li += SM_THREADS)   // same thread that wrote these ccoef rows above
        for (int t = L.cb[li]; t < L.ce[li]; t++) {
          const int lt = L.lcb[li] + (t - L.cb[li]);
          L.cg[lt] = S.cin_grp[t];
#pragma unroll
          for (int c = 0; c < 4; c++) L.cc[(size_t)lt * 4 + c] = S.ccoef[(size_t)t * 4 + c];
        }
  }
  red[0] = red[1] = red[2] = 0.0;
  barrier_reduce<1>(S, counter, phase, red);

  const int n_warm = S.warm ? min((int)S.warm[0], S.warm_systems) : 0;   // rewritten by block 0 at the very end, i.e. after barriers every CTA passes after this read
  int gn_iters = 0, halvings = 0, total_cg = 0, flag = 0;
  double energy = 0.0, normh = 0.0, last_rel = 0.0, abs_target = -1.0, E0 = 0.0;
  bool have_f = false;
  __shared__ double s_time[8];      // block 0 / thread 0: phase and row-phase timers
  __shared__ int s_cg_gn[8];
  __shared__ double s_skew[6];
  if (tid < 6) s_skew[tid] = 0.0;
  if (tid < 8) { s_time[tid] = 0.0; s_cg_gn[tid] = 0; }

  for (int gn = 0; gn < S.max_gn; gn++) {
    gn_iters = gn + 1;
    if (!have_f) {   // S.x holds the current x of every node (published at init / after every accepted step)
      red[0] = rows_smem<K, 0>(S, L, L.xs, S.x, nullptr, 0.0);
      barrier_reduce<1>(S, counter, phase, red);
      E0 = red[0];
    }
    energy = E0;
    // ---- gradient, Jacobi preconditioner
    // Warm start.  Consecutive drag steps solve nearly the same systems, so system gn starts from the solution h' the
    // previous step found for ITS system gn instead of from 0.  It is folded into the loop as a first iteration with
    // p = h' and the exact line-search step alpha = (g.h') / (h'.H h') — the usual CG formula, r.p / p.Hp — after which
    // plain PCG continues with beta = 0.  The optimal alpha rescales the guess: ~1 while the drag is steady, ~-1 when it
    // reverses, ~0 when it pauses (forcing alpha = 1 cost 855 instead of ~200 iterations on the first step of a pause,
    // because the old solution is then a worse start than zero).  The stopping rule is unchanged (residual relative to
    // |g|), so the answer is the same to the solver tolerance; only the number of iterations drops.
    const bool warm = gn < n_warm;
    double* warm_h = S.warm ? S.warm + 8 + (size_t)gn * S.M * 12 : nullptr;
    double rz_l = 0.0, gg_l = 0.0, xx_l = 0.0;
    for (int t = tid; t < NU; t += SM_THREADS) {
      const int li = t / 12, qi = t - 12 * li, j = qi >> 2, c = qi & 3;
      if (!L.fr[li]) continue;
      const size_t ob = (size_t)li * 12;
      const D4 A0 = ld4(L.xs + ob), A1 = ld4(L.xs + ob + 4), A2 = ld4(L.xs + ob + 8);
      double f[6]; rot_res(A0, A1, A2, S.w_rot, f);
      const D4 Aj = j == 0 ? A0 : j == 1 ? A1 : A2;
      const double xv = c == 0 ? Aj.a : c == 1 ? Aj.b : c == 2 ? Aj.c : Aj.d;
      const double g = -gather_smem<K>(S, L, li, j, c, Aj, f, xv);
      const double di = 1.0 / diag_smem<K>(S, L, li, j, c, Aj);
      const double zv = g * di;
      // warm start: the first search direction is the previous drag step's solution of this system (see the PCG loop)
      const double z0 = warm ? warm_h[(size_t)(li * B + b) * 12 + qi] : zv;
      L.ds[t] = di; L.rs[t] = g; L.zs[t] = z0; L.hs[t] = 0.0;
      S.z[(size_t)(li * B + b) * 12 + pub(qi)] = z0;
      rz_l = fma(g, z0, rz_l); gg_l = fma(g, g, gg_l); xx_l = fma(xv, xv, xx_l);   // r.p of the first direction (= r.z when cold)
    }
    red[0] = rz_l; red[1] = gg_l; red[2] = xx_l;
    barrier_reduce<3>(S, counter, phase, red);
    double rz = red[0]; const double gg = red[1]; const double normv = sqrt(red[2]);
    if (abs_target < 0.0) abs_target = S.cg_tol * S.cg_tol * gg;
    // Every system is solved to the absolute target above or to the relative residual S.eta0 of ITS OWN right-hand
    // side, whichever is looser.  A solve error delta_k in Gauss-Newton step k reaches the final iterate damped by the
    // contraction of the remaining steps, i.e. as ~ eta0 |h_last|, and |h_last| is below the stop threshold: the first
    // system (10 decades under the absolute rule) needs no more than the later ones get.
    const double target = fmax(abs_target, S.eta0 * S.eta0 * gg);
    const int cg_before = total_cg;

    // ---- PCG
    double beta = 0.0; int cur = 0;
    if (gg > 0.0) {
      for (int it = 0; it < S.max_cg; it++) {
        double* pnew = cur ? S.p1 : S.p0;
        const double* pold = cur ? S.p0 : S.p1;
        const unsigned long long t0 = gtime2();
        // own p = z + beta p (shared) and publish it; remote p is recombined from the published z and p_old
        for (int t = tid; t < NU; t += SM_THREADS) {
          const double pv = fma(beta, L.ps[t], L.zs[t]);
          L.ps[t] = pv;
          pnew[(size_t)((t / 12) * B + b) * 12 + pub(t % 12)] = pv;
        }
        __syncthreads();
        unsigned long long tm[2];
        const unsigned long long t0b = gtime2();
        __shared__ unsigned long long s_wst[(SM_THREADS / 32) * 6];
        red[0] = rows_smem<K, 1>(S, L, L.ps, S.z, pold, beta, tm, (b == 0 && it == 50) ? s_wst : nullptr);
        if (b == 0 && it == 50 && (tid & 31) == 0) s_wst[(tid >> 5) * 6 + 4] = gtime2();
        const unsigned long long t1 = gtime2();
        barrier_reduce<1>(S, counter, phase, red);
        const unsigned long long t2 = gtime2();
        if (b == 0 && tid == 0 && it == 50) {   // diagnostics: latest warp at each stage of the row phase, relative to its start
          for (int st = 0; st < 5; st++) {
            unsigned long long mx = 0;
            for (int wv = 0; wv < SM_THREADS / 32; wv++) mx = s_wst[wv * 6 + st] > mx ? s_wst[wv * 6 + st] : mx;
            s_skew[st] += (double)(mx - t0b);
          }
          barrier_skew(counter, phase - 1, t2, s_skew + 5);
        }
        const double pHp = red[0];
        const double alpha = (warm && it == 0 && !(pHp > 0.0)) ? 0.0 : rz / pHp;   // zero warm guess: p = 0
        double rzn_l = 0.0, rr_l = 0.0;
        for (int t0 = 0; t0 < NU; t0 += SM_GB * SM_THREADS) {   // NU is a multiple of 4: quads are all in or all out
          GatherLd G[SM_GB];
#pragma unroll
          for (int r = 0; r < SM_GB; r++) {
            const int t = t0 + r * SM_THREADS + tid;
            const bool in = t < NU;
            const int li = in ? t / 12 : 0, qi = in ? t - 12 * li : 0;
            gather_issue<K>(S, L, in && L.fr[li], li, qi >> 2, qi & 3, G[r]);
          }
#pragma unroll
          for (int r = 0; r < SM_GB; r++) {
            const int t = t0 + r * SM_THREADS + tid;
            const bool in = t < NU;
            const int li = in ? t / 12 : 0, qi = in ? t - 12 * li : 0, j = qi >> 2, c = qi & 3;
            const bool act = in && L.fr[li];
            const D4 Aj = ld4(L.xs + (size_t)li * 12 + 4 * j);
            const double pv = act ? L.ps[t] : 0.0;
            const double y = gather_finish<K>(S, L, act, li, j, c, Aj, pv, G[r]);
            if (!act) continue;
            L.hs[t] = fma(alpha, pv, L.hs[t]);
            const double rv = fma(-alpha, y, L.rs[t]);
            L.rs[t] = rv;
            const double zv = rv * L.ds[t];
            L.zs[t] = zv;
            S.z[(size_t)(li * B + b) * 12 + pub(qi)] = zv;
            rzn_l = fma(rv, zv, rzn_l); rr_l = fma(rv, rv, rr_l);
          }
        }
        red[0] = rzn_l; red[1] = rr_l;
        const unsigned long long t3 = gtime2();
        barrier_reduce<2>(S, counter, phase, red);
        const unsigned long long t4 = gtime2();
        if (tid == 0) {
          s_time[4] += (double)(t0b - t0); s_time[5] += (double)(tm[0] - t0b); s_time[6] += (double)(tm[1] - tm[0]); s_time[7] += (double)(t1 - tm[1]);
          s_time[0] += (double)(t1 - t0); s_time[1] += (double)(t2 - t1); s_time[2] += (double)(t3 - t2); s_time[3] += (double)(t4 - t3);
        }
        total_cg++;
        const double rzn = red[0], rr = red[1];
        last_rel = sqrt(rr / gg);
        cur ^= 1;
        if ((!(pHp > 0.0) && !(warm && it == 0)) || !(rr == rr)) { flag |= 1; break; }   // a zero warm guess is fine (p = 0)
        if (rr <= target || rr <= 1e-30 * gg) break;
        beta = (warm && it == 0) ? 0.0 : rzn / rz; rz = rzn;
        if (it == S.max_cg - 1) flag |= 2;
      }
    }
    if (flag & 1) {   // numeric breakdown (uniform over the grid: pHp / rr come out of the grid reduction): h may hold NaN.  Take a zero
      // step instead — x keeps the last good iterate (identity on the first system), the loop ends through the |h| test, the
      // caller sees flag bit 0, and arap_apply gets finite transforms.
      for (int t = tid; t < NU; t += SM_THREADS) L.hs[t] = 0.0;
      __syncthreads();
    }
    if (S.warm && gn < SOLVE_WARM_MAX)   // this system's solution (before step halving) seeds the next drag step
      for (int t = tid; t < NU; t += SM_THREADS) warm_h[(size_t)((t / 12) * B + b) * 12 + (t % 12)] = L.hs[t];

    // ---- step halving (Deform.cpp:144-156): publish x + h, evaluate, accept or halve
    bool accepted = false;
    for (double alpha_ls = 1.0; alpha_ls > 1e-15; alpha_ls *= 0.5) {
      double hh_l = 0.0;
      for (int t = tid; t < NU; t += SM_THREADS) {
        const double hv = L.hs[t];
        hh_l = fma(hv, hv, hh_l);
        const double xv = L.xs[t] + hv;
        L.zs[t] = xv;                                                       // zs is free between linear solves: holds x + h
        S.x[(size_t)((t / 12) * B + b) * 12 + pub(t % 12)] = xv;
      }
      red[0] = hh_l;
      barrier_reduce<1>(S, counter, phase, red);
      const double hh = red[0];
      red[0] = rows_smem<K, 0>(S, L, L.zs, S.x, nullptr, 0.0);
      barrier_reduce<1>(S, counter, phase, red);
      const double E1 = red[0];
      if (!(E1 <= E0)) {   // also rejects a non-finite energy (the reference's `>` would accept NaN)
        for (int t = tid; t < NU; t += SM_THREADS) L.hs[t] *= 0.5;
        halvings++;
        normh = 0.5 * sqrt(hh);
        __syncthreads();
      } else {
        for (int t = tid; t < NU; t += SM_THREADS) L.xs[t] = L.zs[t];
        normh = sqrt(hh);
        E0 = E1; have_f = true; accepted = true;   // S.x already holds the accepted x; row buffers hold f(x)
        __syncthreads();
        break;
      }
    }
    if (!accepted) {
      // restore the published x; f(x) must be recomputed
      for (int t = tid; t < NU; t += SM_THREADS) S.x[(size_t)((t / 12) * B + b) * 12 + pub(t % 12)] = L.xs[t];
      red[0] = 0.0;
      barrier_reduce<1>(S, counter, phase, red);
      have_f = false;
    }
    if (gn < 8 && tid == 0) s_cg_gn[gn] = total_cg - cg_before;
    if (normh < (normv + 1e-6) * 1e-6) break;
  }

  // putFreeInputs (Deform.hpp:140-151)
  for (int t = tid; t < NU; t += SM_THREADS) {
    const int li = t / 12, qi = t - 12 * li, jj = qi >> 2, c = qi & 3, i = li * B + b;
    const double v = L.xs[t];
    if (c < 3) S.rot_out[(size_t)i * 9 + jj + 3 * c] = v; else S.trans_out[(size_t)i * 3 + jj] = v;
  }
  if (b == 0 && tid == 0) {
    if (S.warm) S.warm[0] = (flag & 1) ? 0.0 : (double)min(gn_iters, SOLVE_WARM_MAX);
    S.stats[0] = gn_iters; S.stats[1] = energy; S.stats[2] = halvings; S.stats[3] = normh;
    S.stats[4] = total_cg; S.stats[5] = last_rel; S.stats[6] = flag;
    S.stats[8] = s_time[0]; S.stats[9] = s_time[1]; S.stats[10] = s_time[2]; S.stats[11] = s_time[3]; S.stats[12] = gridDim.x;
    for (int t = 0; t < 8; t++) S.stats[16 + t] = s_cg_gn[t];
    for (int t = 0; t < 6; t++) S.stats[24 + t] = s_skew[t];
    S.stats[13] = s_time[4]; S.stats[14] = s_time[5]; S.stats[15] = s_time[6]; S.stats[7] = s_time[7];
  }
}

size_t solve_smem_bytes(int NL, int K, int gcap, int ccap) {
  size_t d = (size_t)NL * 12 * 6 + (size_t)NL * K * 3 + (size_t)NL * 6;
  if ((NL * K * 3) & 1) d += 1;
  return d * 8 + (size_t)NL * K * 16 + (size_t)NL * K * 8 + (size_t)NL * 7 * 4 + (size_t)(SM_GMAXG + 1) * 4 +
         (size_t)gcap * (4 + 32 + 24) + (size_t)ccap * (4 + 32) + 96;
}

// returns ARAP_OK if launched, -1 if the slice does not fit (caller falls back to the global-memory kernel)
int launch_solve_smem(const SolveDev& S, unsigned* counter, cudaStream_t st, int max_ctas) {
  if (S.k != 8 && S.k != 10 && S.k != 12) return -1;
  int dev = 0, sms = 0, max_smem = 0;
  ARAP_CUDA_TRY(cudaGetDevice(&dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  ARAP_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  // small graphs: fewer CTAs (>= ~24 nodes each) make the barriers cheaper
  int grid = std::max(1, std::min(sms, (S.M + 23) / 24));
  if (max_ctas > 0) grid = std::min(grid, max_ctas);   // SMs left to concurrent kernels (arap_params.solver_ctas)
  if (const char* ev = getenv("ARAP_SOLVE_GRID")) grid = std::max(1, std::min(sms, atoi(ev)));   // measurement aid
  const int NL = (S.M + grid - 1) / grid;
  void* kern = nullptr; size_t smem = 0;
  auto pick = [&](auto kc) {
    constexpr int KK = decltype(kc)::value;
    if (NL <= 112) { kern = (void*)k_solve_smem<KK, 112>; smem = solve_smem_bytes(112, KK, SmCaps<KK, 112>::G, SmCaps<KK, 112>::C); }
    else if (NL <= 136) { kern = (void*)k_solve_smem<KK, 136>; smem = solve_smem_bytes(136, KK, SmCaps<KK, 136>::G, SmCaps<KK, 136>::C); }
    else if (NL <= 176) { kern = (void*)k_solve_smem<KK, 176>; smem = solve_smem_bytes(176, KK, SmCaps<KK, 176>::G, SmCaps<KK, 176>::C); }
  };
  if (S.k == 8) pick(std::integral_constant<int, 8>{});
  else if (S.k == 10) pick(std::integral_constant<int, 10>{});
  else pick(std::integral_constant<int, 12>{});
  if (!kern) return -1;   // slice does not fit: the global-memory kernel runs
  cudaFuncAttributes fa;
  ARAP_CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)kern));
  if (smem + fa.sharedSizeBytes > (size_t)max_smem) return -1;
  ARAP_CUDA_TRY(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  ARAP_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)kern, SM_THREADS, smem));
  if (per_sm < 1) return -1;
  ARAP_CUDA_TRY(cudaMemsetAsync(counter, 0, (size_t)2 * (grid + 1) * LL_WORDS * sizeof(unsigned long long), st));
  SolveDev Sc = S; unsigned* cnt = counter;
  void* args[] = {(void*)&Sc, (void*)&cnt};
  ARAP_CUDA_TRY(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(SM_THREADS), args, smem, st));
  return ARAP_OK;
}

}  // namespace arapgs
