// Device-side argument block shared by the two solver kernels (solve.cu: global-memory vectors; solve_smem.cu:
// shared-memory-resident per-node state).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace arapgs {

struct SolveDev {
  int M, k, n_groups;
  const float* node_pos;      // M x 3
  const int* nbr;             // M x k
  const int* in_off;          // M + 1: in-edges from FREE sources
  const int* out_to_in;       // M x k: slot of edge (i,s) in u_in, or -1
  const int* anc_idx;         // M x k
  const double* anc_w;        // M x k
  const uint8_t* node_free;   // M
  const int* static_in_cnt;   // M
  const int* grp_off;         // n_groups + 1 -> members
  const int* grp_member;
  const float* grp_aim;       // n_groups x 3
  const int* cin_off;         // M + 1 -> constraint entries touching the node, sorted by group
  const int* cin_grp;
  const int* cin_member;
  const int* cin_slot;
  double w_rot, w_reg, w_con;  // square-rooted (Deform.hpp:452-454)
  int max_gn, max_cg;
  double cg_tol, eta0;
  // work (double).  Vectors: [M][3][4]
  double *x, *h, *r, *z, *p0, *p1, *dinv, *bedge /* M x k x 4 */, *ccoef /* cin entries x 4 */;
  double *u_reg /* M x k x 3 */, *u_in /* in-edges x 3 */, *u_con /* groups x 3 */;
  double* gent_c;              // constraint-row entries x 4 (shared-memory kernel)
  int* gent_q;                 // node of each entry, -1 if excluded
  double* partial;             // 2 x gridDim x NRED
  double *rot_out, *trans_out, *stats;
  // warm start of the linear solves (shared-memory kernel): warm[0] = number of Gauss-Newton systems whose solution the
  // previous solve saved, warm + 8 + g * M * 12 = that solution for system g < SOLVE_WARM_MAX; null = off
  double* warm;
  int warm_systems;            // how many Gauss-Newton systems of a step may be warm-started (<= SOLVE_WARM_MAX)
};
constexpr int SOLVE_WARM_MAX = 4;


// solve_smem.cu: returns ARAP_OK if launched, a positive error code on CUDA failure, -1 if the per-CTA slice does not
// fit in shared memory (the caller then runs the global-memory kernel).
int launch_solve_smem(const SolveDev& S, unsigned* counter, cudaStream_t st, int max_ctas);
// solve_pipe.cu: one-barrier pipelined PCG; `extra` = solve_pipe_extra_doubles() doubles of workspace, 16-byte aligned.
// Same return convention.
int launch_solve_pipe(const SolveDev& S, unsigned* counter, double* extra, cudaStream_t st, int max_ctas);
size_t solve_pipe_extra_doubles(int M, int k, int n_groups);

}  // namespace arapgs
