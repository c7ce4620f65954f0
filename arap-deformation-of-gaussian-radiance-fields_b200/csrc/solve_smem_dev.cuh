// Device-side pieces shared by the two shared-memory-resident solver kernels (solve_smem.cu: two-phase matrix-free PCG;
// solve_pipe.cu: one-barrier pipelined PCG on the explicit normal-equation stencil): constants, the per-CTA slice `Loc`,
// the flag-in-data grid barrier, the row / gather / diagonal evaluators.  Included inside namespace arapgs.
#pragma once

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
  // solve_pipe.cu only
  double *ms, *ws, *ss;                      // [NL*12] m = D^-1 w (the published vector), w, s of the pipelined recurrences
  double* Cm;                                // [NL*16] sum_s c c^T, c = (g_q - g_i, 1): the E_reg self block (rows of 4)
  double* cu;                                // [ccap*3] row value u_gj of the group of each column entry
  int *eslot, *erd, *emeta;                  // [ccap] slot the entry publishes to; first slot of its group / big-group index; li | big << 31
  int nent;
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
static __device__ __noinline__ void barrier_skew(unsigned* counter, int ph, unsigned long long t_exit, double* out) {
  const unsigned long long* set = reinterpret_cast<const unsigned long long*>(counter) + (size_t)(ph & 1) * gridDim.x * LL_WORDS;
  unsigned long long mx = 0;
  for (int bb = 0; bb < (int)gridDim.x; bb++) {
    const unsigned long long t = ll_load(set + (size_t)bb * LL_WORDS + 7);
    mx = t > mx ? t : mx;
  }
  out[0] += (double)(t_exit - mx);   // last arrival -> block 0's exit: the barrier mechanism itself
}
