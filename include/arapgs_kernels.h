/* Kernel layer of libarapgs: the C ABI one level below include/arapgs.h.
 *
 * Stateless entry points: plain device pointers + sizes + an explicit cudaStream_t, `int` status (codes of arapgs.h,
 * message through arap_last_error()).  The session layer (arap_*, arapgs.h) is built on these; the -m gpu kernel
 * parity tests call them directly.  Every group cites the reference interface it replaces. */
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif


/* ---- stage (d) apply + per-step sample passes (csrc/apply.cu) -----------------
 * Reference interfaces replaced (src/projects/gaussianviewer/renderer/):
 *   arapk_node_xf / arapk_lbs_points / arapk_lbs_tiles   DeformGraph::predict_mesh / predict_samples, Deform.hpp:230-268
 *                                                        (callers UpdatePosition / UpdatePositionforSamples, GaussianView.cpp:3021-3077)
 *   arapk_lbs_build_tiles                                set-up of the staged LBS tables (ours; no reference counterpart)
 *   arapk_end_points                                     GaussianView::GetEndPoints, GaussianView.cpp:4643-4668
 *   arapk_fit_gaussians                                  GaussianView::UpdateAsSixPointsWithdrawBad, GaussianView.cpp:3081-3166
 *   arapk_node_quats / arapk_rotate_sample_shs           FastUpdateSamplesSH + RotateSHs, GaussianView.cpp:3169-3186, cudakdtree.cu:201-222
 *   arapk_static_flags                                   CheckStaticSamples / CheckMovedGaussians, GaussianView.cpp:2024-2109 */
int arapk_node_xf(int M, const double* rot, const double* trans, const float* node_pos, void* node_xf,
                  void* node_xf32 /* M x 64 B, may be null */, cudaStream_t st);
int arapk_lbs_points(const float* in, float* out, long long P, int k, const uint16_t* ridx, const double* rw,
                     const void* node_xf, const uint8_t* skip, int group, cudaStream_t st);
long long arapk_lbs_tile_count(long long rows);
int arapk_lbs_tile_cap(void);
int arapk_lbs_build_tiles(long long rows, int k, const uint16_t* ridx, uint32_t* slots, uint16_t* tile_cnt,
                          uint16_t* tile_nodes, cudaStream_t st);
int arapk_lbs_tiles(const float* in, float* out, long long P, int k, const uint32_t* slots, const double* rw,
                    const uint16_t* ridx, const uint16_t* tile_cnt, const uint16_t* tile_nodes, const void* node_xf,
                    const uint8_t* skip, int group, int magic, cudaStream_t st);
int arapk_end_points(long long N, const float* pos, const float* rot, const float* scale, float* ends, cudaStream_t st);
int arapk_fit_gaussians(long long N, const float* ends, const float* scale_backup, const uint8_t* is_static, float* pos,
                        float* rot, float* scale, float* shs, cudaStream_t st);
/* tolerance mode (arap_params.lbs_mode = 3): neighbour UNIONS with dense float weights and float records, same reference
 * interfaces as above (DeformGraph::predict_mesh / predict_samples, Deform.hpp:230-268; UpdateAsSixPointsWithdrawBad,
 * GaussianView.cpp:3081-3166).  arapk_sunion_build / arapk_gunion_build: two-pass set-up (first call with the table pointers
 * NULL returns the sizes); arapk_lbs_union32: a row family; arapk_apply_union: end-point skinning + six-point fit + SH rotation
 * in one pass.  node_xf32 = the float records arapk_node_xf writes next to the fp64 ones (64 B per node). */
int arapk_sunion_build(long long rows, int k, const uint16_t* ridx, const float* wf, const double* wd, int* boff, uint16_t* blist,
                       float* bw, long long* rows_out, int* scratch, cudaStream_t st);
int arapk_lbs_union32(const float* in, float* out, long long P, const int* boff, const uint16_t* blist, const float* bw,
                      const void* node_xf32, const uint8_t* skip, int group, cudaStream_t st);
long long arapk_gtile_count(long long n_gaussians);
int arapk_gtile_cap(void);
int arapk_gunion_build(long long N, int k, const uint16_t* end_ridx, const double* end_rw, int* uoff, int* woff, uint32_t* usw,
                       uint16_t* unode, float* uw, uint16_t* gtile_cnt, uint16_t* gtile_nodes, long long* rows_out,
                       long long* words_out, int* scratch, cudaStream_t st);
int arapk_apply_union(long long N, const void* node_xf32, const uint16_t* gtile_cnt, const uint16_t* gtile_nodes, const int* uoff,
                      const int* woff, const uint32_t* usw, const uint16_t* unode, const float* uw, float* ends,
                      const float* scale_backup, const uint8_t* is_static, float* pos, float* rot, float* scale, float* shs,
                      cudaStream_t st);
/* the same pass with the fused multi-GPU epilogue: every tile's final pos / rot / scale is also stored into the peers' arrays
 * (pointers already offset to THIS rank's range; peers = NULL or n = 0: no epilogue) */
#define ARAP_MAX_PEERS 7
typedef struct ArapPeerPush { float* pos[ARAP_MAX_PEERS]; float* rot[ARAP_MAX_PEERS]; float* scale[ARAP_MAX_PEERS]; int n;
                               int multicast; /* 1: n = 1 and the pointers are an NVSwitch multicast mapping (multimem.st) */ } ArapPeerPush;
int arapk_apply_union_push(long long N, const void* node_xf32, const uint16_t* gtile_cnt, const uint16_t* gtile_nodes, const int* uoff,
                      const int* woff, const uint32_t* usw, const uint16_t* unode, const float* uw, float* ends,
                      const float* scale_backup, const uint8_t* is_static, float* pos, float* rot, float* scale, float* shs,
                      const ArapPeerPush* peers, cudaStream_t st);
/* multi-GPU receivers: repeat the owner's SH update of k_fit_gaussians from (old rotation, new rotation) on a held copy */
int arapk_replay_shs(long long N, const float* rot_old, const float* rot_new, const uint8_t* is_static, float* shs,
                     cudaStream_t st);
int arapk_node_quats(int M, const double* rot, float* q_xyzw, cudaStream_t st);
int arapk_rotate_sample_shs(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                            const uint8_t* is_static, float* feature, cudaStream_t st);
/* the same pass carrying the multi-GPU exchange: a few CTAs of the launch stream this rank's (final) pose arrays pos / rot / scale [n]
 * into the peers' copies while the others rotate the SH rows, so the transfer is spread under the longest kernel of the step */
typedef struct ArapPosePush { ArapPeerPush peers; const float* pos; const float* rot; const float* scale; long long n;
                               int copy_ctas; /* set by the launcher */ } ArapPosePush;
int arapk_rotate_sample_shs_push(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                            const uint8_t* is_static, float* feature, const ArapPosePush* push, cudaStream_t st);
/* deferred sample SH rotation (arap_params.lazy_sample_sh): compose the step's blended sample quaternion onto an accumulator;
 * the rows are rotated later by arapk_replay_shs with rot_old = NULL (identity) and rot_new = the accumulator */
int arapk_accumulate_sample_quats(long long S, int k, const float* w, const uint16_t* idx, const float* q_xyzw,
                                  const uint8_t* is_static, float* qacc_wxyz, cudaStream_t st);
int arapk_fill_identity_quats(long long S, float* q_wxyz, cudaStream_t st);
int arapk_static_flags(long long G, int group, int k, const uint16_t* idx, const uint8_t* node_static, uint8_t* out,
                       cudaStream_t st);
int arapk_sh_rotate_test(const float* R9, float* shs48_dev, int fast, cudaStream_t st);

/* ---- stage (b) FPS + kNN (csrc/knn.cu) -------------------------------------------
 *   arapk_fps                                            farthest_control_points_sampling, helper.cpp:139-195
 *   arapk_knn_build / arapk_knn_query                    DeformGraph::findNearestNodes + computeWeights, Deform.hpp:153-208
 *                                                        (callers setupWeights*, GaussianView.cpp:2818-2918; setupEdges, Deform.hpp:84-95)
 *   arapk_minmax                                         getOverallAABB, GaussianView.cpp:3601-3631 */
int arapk_minmax(const float* pts, long long N, float* out6_dev, cudaStream_t st);
int arapk_fps(const float* pos, long long N, int node_num, int* out_idx_dev, void* scratch, size_t scratch_bytes,
              int* out_count_host, cudaStream_t st);
/* the same selection over cell-ordered points, pruned by the density grid (bit-identical sequence) */
size_t arapk_fps_grid_scratch_bytes(long long N, int G);
int arapk_fps_grid(const float* pos, long long N, int node_num, const int* cell_prefix, const float* min3_host, float step, int G,
                   int* out_idx_dev, void* scratch, size_t scratch_bytes, int* out_count_host, cudaStream_t st);
size_t arapk_knn_workspace_bytes(int M);
size_t arapk_knn_index_struct_bytes();
int arapk_knn_build(const float* nodes_dev, int M, void* workspace, size_t workspace_bytes, void* index_out, cudaStream_t st);
int arapk_knn_query(const void* index, const float* queries_dev, long long Q, int k, uint32_t* idx_plain, double* w_plain,
                    uint16_t* idx_blk, double* w_blk, float* wf_blk, uint32_t* idx_kq, void* slow_scratch,
                    size_t slow_scratch_bytes, int* n_slow_host, cudaStream_t st);

/* ---- stage (c) solve (csrc/solve.cu, csrc/solve_smem.cu) -------------------------
 *   arapk_solve                                          Deform ctor + Deform::optimize + putFreeInputs,
 *                                                        Deform.hpp:414-455, Deform.cpp:95-169, 378-581, Deform.hpp:140-151 */
typedef struct ArapSolveGraph {
  int M, k, n_groups;
  const float* node_pos;      // M x 3, current node positions (device)
  const int* nbr;             // M x k out-neighbours (Node.Neighbor)
  const int* in_off;          // M + 1: in-edges whose source node is free (row values land in u_in[in_off[q] ..])
  const int* out_to_in;       // M x k: for edge (i, s) its slot in the destination's in-edge range, -1 if i is excluded
  const int* anc_idx;         // M x k: anchor vertex kNN row of each node
  const double* anc_w;        // M x k
  const uint8_t* node_free;   // M: 1 = unknown, 0 = excluded (identity)
  const int* static_in_cnt;   // M: number of (excluded node, slot) pairs pointing at the node
  const int* grp_off;         // n_groups + 1: constraint groups -> members
  const int* grp_member;      // node ids
  const float* grp_aim;       // n_groups x 3 target (node aim or block centre aim)
  const int* cin_off;         // M + 1: constraint entries touching a node, sorted by group
  const int* cin_grp;
  const int* cin_member;
  const int* cin_slot;
  long long n_cin_entries;    // cin_off[M]
} ArapSolveGraph;

typedef struct ArapSolveParams {
  double w_rot, w_reg, w_con;  // un-rooted weights (1, 10, 100 by default)
  int max_gn_iters;            // MAX_ITERS 30
  int max_cg_iters;
  double cg_tol;               // relative residual of the first linear system; later ones reuse its absolute value
  int force_global_kernel;     // 1: skip the shared-memory-resident fast path (tests)
  double newton_eta0;          // > 0: inexact Newton forcing (shared-memory kernel only), see arapgs.h
  double* warm_buf;            // device, arapk_solve_warm_doubles(M) doubles, zeroed by the caller whenever the unknown set changes;
                               // null = every PCG solve starts from 0.  See arapgs.h (arap_params.warm_start).
  int max_ctas;                // 0 = one CTA per SM; n > 0 = at most n CTAs (shared-memory kernel)
  int warm_systems;            // 0 = all (SOLVE_WARM_MAX), n > 0 = only the first n Gauss-Newton systems of a step
  int pipelined;               // 1 = one-barrier pipelined PCG kernel (csrc/solve_pipe.cu) when the slice fits it; the caller must have
                               // checked arapk_solve_pipe_eligible() for this constraint set (the kernel re-checks: stats flag bit 2)
} ArapSolveParams;

/* host-side check for ArapSolveParams.pipelined: host copies of grp_off (n_groups + 1) and cin_off (M + 1) */
int arapk_solve_pipe_eligible(int M, int k, int n_groups, const int* grp_off_host, const int* cin_off_host, int max_ctas);

size_t arapk_solve_warm_doubles(int M);

size_t arapk_solve_workspace_bytes(int M, int k, int n_groups);
int arapk_solve(const ArapSolveGraph* G, const ArapSolveParams* P, void* workspace, size_t workspace_bytes,
                double* rot_out, double* trans_out, double* stats_dev, cudaStream_t st);

/* ---- stage (a) density grid (csrc/grid.cu) ---------------------------------------
 *   arapk_cell_assign / arapk_permute_gaussians          GetGsGrid + host re-order, cudakdtree.cu:403-423, GaussianView.cpp:3922-3953
 *   arapk_gs_aabbs                                       per-Gaussian cutoff boxes, GaussianView.cpp:3961-4019
 *   arapk_footprint_count / arapk_footprint_fill         GetBoxesGsGrid + serial fill, cudakdtree.cu:425-451, GaussianView.cpp:4077-4100
 *   arapk_valid_cells / arapk_emit_samples               GaussianView.cpp:4040-4053, 4111-4133
 *   arapk_ada_lpf                                        GetAdaLpfRatio, GaussianView.cpp:4670-4751
 *   arapk_grid_eval                                      Rasterizer::forward3d_grid call sites, GaussianView.cpp:4159-4186 */
int arapk_cell_assign(const float* pos, long long N, const float* min3_host, float step, int G, int* cell_out,
                      int* prefix_out /* G^3 inclusive */, int* new_idx_out /* N, may be null */, void* scratch,
                      size_t scratch_bytes, cudaStream_t st);
size_t arapk_grid_scratch_bytes(long long N, int G);
int arapk_permute_gaussians(long long N, const int* new_idx, const float* pos, const float* rot, const float* scale,
                            const float* opacity, const float* shs, float* pos_o, float* rot_o, float* scale_o,
                            float* opacity_o, float* shs_o, cudaStream_t st);
int arapk_gs_aabbs(long long N, const float* pos, const float* rot, const float* scale, const float* opacity, float* aabb,
                   float* clip /* may be null */, float* smax /* may be null */, cudaStream_t st);
int arapk_footprint_count(long long N, const float* aabb, const float* min3_host, float step, int G, int padding,
                          int* prefix_out /* G^3 inclusive */, long long* total_host, void* scratch, size_t scratch_bytes,
                          cudaStream_t st);
int arapk_footprint_fill(long long N, const float* aabb, const float* min3_host, float step, int G, int padding,
                         const int* prefix, int* lists_out, void* scratch, size_t scratch_bytes, cudaStream_t st);
/* multi-GPU grid sharding: the same passes restricted to the x-slab [xlo, xhi) of cells; per-x-layer Gaussian counts */
int arapk_footprint_count_slab(long long N, const float* aabb, const float* min3_host, float step, int G, int padding, int xlo, int xhi,
                               int* prefix_out, long long* total_host, void* scratch, size_t scratch_bytes, cudaStream_t st);
int arapk_footprint_fill_slab(long long N, const float* aabb, const float* min3_host, float step, int G, int padding, int xlo, int xhi,
                              const int* prefix, int* lists_out, void* scratch, size_t scratch_bytes, cudaStream_t st);
int arapk_xlayer_hist(const float* pos, long long N, const float* min3_host, float step, int G, int* hist_dev /* G */, cudaStream_t st);
/* host only: balanced x-slab cuts [world + 1] from the per-layer counts (every rank at least one layer) */
int arapk_slab_cuts(const int* layer_count_host, int G, int world, int* cuts_out);
int arapk_valid_cells(const int* prefix, int G, int* valid_out, int* count_host, void* scratch, size_t scratch_bytes,
                      cudaStream_t st);
int arapk_emit_samples(const int* valid, int V, const float* min3_host, float step, int G, float* out, cudaStream_t st);
int arapk_ada_lpf(const float* samples, const int* valid, int V, float lpf_parameter, float* out /* G^3 x 9 */,
                  cudaStream_t st);
/* JudgeEmptyGrid, GaussianView.cpp:4272-4318; scratch >= G^3 bytes */
int arapk_judge_empty_grid(const int* valid, int V, const float* aim_opacity, int G, int* empty_out, void* scratch,
                           size_t scratch_bytes, cudaStream_t st);
int arapk_grid_eval(const int* valid, int V, const int* prefix, const int* lists, const float* samples, const float* pos,
                    const float* rot, const float* scale, const float* opacity, const float* shs, const float* ada_lpf,
                    float* out_feature, float* out_opacity, cudaStream_t st);


#ifdef __cplusplus
}  /* extern "C" */
#endif
