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

#include "solve_smem_dev.cuh"

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
