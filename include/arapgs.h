/* arapgs.h — C ABI of libarapgs: B200-native (sm_100a) ARAP deformation of
 * Gaussian radiance fields.
 *
 * Drop-in boundary for the deformation path of the reference viewer
 * (XinhaoT/ARAP-Deformation-of-Gaussian-Radiance-Fields, SIBR gaussianViewer).
 * The reference has no plugin / FFI layer: its deformation API is the set of
 * C++ entry points GaussianView calls.  Each function below names the reference
 * call site(s) it replaces, relative to
 *   src/projects/gaussianviewer/renderer/   (GV = GaussianView.cpp, DH = Deform.hpp,
 *   DC = Deform.cpp, HC = helper.cpp, CK = cudakdtree.cu).
 *
 * Conventions:
 *   - plain pointers and sizes only; all functions return an int status
 *     (ARAP_OK = 0) and never exit(); arap_last_error() gives the message.
 *   - an arap_ctx owns every device buffer; the caller may borrow the SoA
 *     pointers for the rasteriser (arap_device_view).
 *   - one explicit CUDA stream per ctx; calls are not re-entrant per ctx and
 *     only synchronise where the host reads a result.
 *   - there is NO CPU fallback: every entry point needs the CUDA device.
 */
#ifndef ARAPGS_H
#define ARAPGS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARAP_OK 0
#define ARAP_ERR_INVALID 1
#define ARAP_ERR_CUDA 2
#define ARAP_ERR_STATE 3
#define ARAP_ERR_IO 4
#define ARAP_ERR_KNN_TIES 5
#define ARAP_ERR_UNSUPPORTED 6
#define ARAP_ERR_NUMERIC 7

#define ARAP_KNN_MAX 12 /* helper.hpp:48 */

typedef struct arap_ctx arap_ctx;

/* Tunables the reference keeps as GaussianView members / ImGui widgets
 * (GV.hpp:322, 446, 455-457, 482, 598; DC:3-4). */
typedef struct arap_params {
  int grid_num;        /* 64 (paper default) .. 128; <ply>_config.txt field 0 (GV:459-487) */
  int padding;         /* 1  (GV.hpp:446) */
  int knn_k;           /* 10 (GV.hpp:598; README: 8 to reproduce the paper) */
  int node_num;        /* 150 (GV.hpp:322) */
  int high_quality;    /* config field 2 (GV:459-487).  Recorded and reported back only: its sole consumers are inside the external
                          CudaRasterizer fork (forward3d_grid / L1loss3d), whose source is not in the reference tree, and
                          ExpandOpRange of the stage-II optimiser (GV:1700-1702, out of scope).  It changes nothing here. */
  float lpf_parameter; /* 0.2 (GV.hpp:482) */
  double w_rot, w_reg, w_con; /* 1, 10, 100 (DH:56-58) */
  int max_gn_iters;    /* 30 (DC:3) */
  int max_cg_iters;    /* PCG iteration cap per linear system */
  double cg_tol;       /* relative residual of the first linear system of a step */
  int skip_static_endpoints; /* 0 = reference behaviour (all endpoints skinned) */
  int solver_global_memory;  /* 1 = force the global-memory solver kernel (default 0: shared-memory-resident kernel when it fits) */
  int lbs_mode;        /* skinning (LBS) kernels.  0 (default) = bit-faithful: the reference's `float += double` chain (DH:239-246) with node
                          records staged per 128-row tile in shared memory and the float rounding on the FP64 pipe; 1 = the same from
                          global memory; 2 = staged records + conversion instructions — 0, 1, 2 produce identical bits.
                          3 = tolerance mode: end-point skinning fused into the six-point fit and sample skinning evaluated as
                          p' = p + blend(A - I | t').(p - c; 1) in float (csrc/apply.cu) — the correctly rounded exact skinning in all
                          but a few per cent of the coordinates, i.e. within ~1 float ulp of p of the reference's chain (which carries
                          that much rounding noise of its own); end points of excluded-only Gaussians are not touched.
                          Node and mesh-point positions always use the bit-faithful kernel. */
  double newton_eta0;  /* each Gauss-Newton linear system stops at relative residual newton_eta0 of its own right-hand side
                          or at the cg_tol target, whichever is looser (0 = cg_tol target only).  Default 1e-6: the error of
                          system k reaches the result damped by the remaining Gauss-Newton steps, ~ newton_eta0 x (last step
                          length); measured node-transform deviation from the reference's direct solves stays at the 1e-11 level
                          of the cg_tol-only rule with ~25% fewer PCG iterations (1e-4 is where the parity bar is reached). */
  int warm_start;      /* (shared-memory solver only; the global-memory fallback kernel always starts from zero)
                          1 (default): every PCG solve of a drag step starts from the solution the previous step found for the same
                          Gauss-Newton system, scaled by an exact line search (zero after arap_set_blocks / a graph build).  Same stopping rule, same answer to the
                          solver tolerance, fewer iterations while the drag is coherent.  0 = start from zero like the first step;
                          n > 1 = warm-start only the first n - 1 systems of a step. */
  int solver_ctas;     /* 0 (default): the solve uses one CTA per SM.  n > 0: at most n CTAs, leaving the other SMs to kernels that run
                          beside it — the multi-GPU driver reserves SMs for the NCCL all-gather of the previous step's SoA, which
                          otherwise cannot overlap the solve (a 512-thread solver CTA fills an SM's register file). */
  int lazy_sample_sh;  /* 0 (default): the aim features of the samples are rotated every drag step (FastUpdateSamplesSH, GV:1519), as in the
                          reference.  1: a step only composes its blended sample rotation onto a per-sample quaternion (32 B instead of
                          384 B of traffic per sample); the feature rows are rotated once by the accumulated rotation when they are
                          consumed — arap_grid_update_lists (stroke end), arap_get_device_view, the download calls, or
                          arap_sample_features_materialize.  Same result up to float rounding (SH rotation composes exactly). */
  int fps_mode;        /* node sampling (HC:139-195): 0 (default) = selection loop pruned by the density grid, 1 = one pass over all
                          candidate points per node (first version).  Same node sequence, bit for bit. */
  int solver_pipelined; /* 1 (default): the linear systems of the Gauss-Newton solve run as pipelined PCG on the explicit J^T J stencil with
                          ONE grid barrier per iteration (csrc/solve_pipe.cu) whenever the constraint set fits that kernel (checked at
                          arap_set_blocks); 0 = the two-barrier matrix-free PCG (csrc/solve_smem.cu).  Same preconditioner, same stopping
                          rules, iterates equal to rounding. */
} arap_params;

typedef struct arap_solve_stats {
  int gn_iters;        /* Gauss-Newton iterations taken */
  int cg_iters;        /* total PCG iterations (one-barrier kernel: products with J^T J, i.e. iterations + 1-2 set-up products per system) */
  int halvings;        /* step-halving count */
  int flags;           /* bit0: numeric breakdown, bit1: PCG hit the iteration cap, bit2: constraint set does not fit the selected solver kernel (identity transforms returned) */
  double energy;       /* f.f at the last linearisation point (Deform::optimize return) */
  double normh;        /* |h| of the last accepted step */
  double last_rel_residual;
  double phase_ns[4];  /* block 0's time (summed over PCG iterations) in: row phase, barrier 1, gather/update phase, barrier 2 (two-barrier kernel);
                          stencil + recurrences + publication, CTA sync + E_rot rows, barrier, 0 (one-barrier kernel, solver_pipelined) */
  int grid_blocks;     /* cooperative grid size used */
  double row_sub_ns[4]; /* row phase split: form p, E_reg rows, E_rot rows, constraint rows (two-barrier shared-memory kernel); one-barrier kernel:
                           group sums + gathers issued, CTA sync, first row's stencil, recurrences + publication */
  int cg_iters_gn[8];  /* PCG iterations of the first 8 Gauss-Newton iterations */
  double barrier_skew_ns[6]; /* diagnostics sampled at PCG iteration 50 of each Gauss-Newton iteration (summed), block 0: time from the
                                start of the row phase until its LAST warp has finished E_reg rows, E_rot rows, constraint gathers,
                                the CTA sync, the whole phase; [5] = last CTA's arrival at the barrier -> block 0's exit.
                                One-barrier kernel: [0..3] = a CTA's own work (everything but the barrier) summed over the PCG iterations:
                                mean, max, min over the CTAs, block 0's */
} arap_solve_stats;

typedef struct arap_grid_info {
  int grid_num, padding;
  int valid_cells;          /* valid_grid_num */
  long long samples;        /* 64 * valid_cells */
  long long pairs;          /* grid_gs_prefix_sum[G^3-1] */
  float aabb_min[3], aabb_max[3];
  float grid_step;
} arap_grid_info;

/* Borrowed device pointers: exactly the Rasterizer::forward /
 * forward3d_grid argument arrays (GV:1106-1112, 4159-4186).
 * Validity: pos / rot / scale / opacity / shs keep their addresses from arap_set_gaussians until the next
 * arap_set_gaussians or arap_destroy (arap_grid_build re-orders them in place, arap_apply / arap_step update them in place).
 * node_pos / node_rot / node_trans are stable from a graph build (arap_graph_build_* / arap_replay with rebuild) until the
 * next one; every arap_apply writes the new node positions to the same address.  The grid arrays (valid_grid ..
 * ada_lpf_ratio, sample_pos, end_points) are re-allocated by arap_grid_build and arap_grid_update_lists; aim_feature /
 * aim_opacity by arap_grid_build + arap_grid_eval.  Fetch the view again after any of those calls. */
typedef struct arap_device_view {
  long long n_gaussians;
  float* pos;      /* N x 3 */
  float* rot;      /* N x 4 (w,x,y,z) */
  float* scale;    /* N x 3, linear */
  float* opacity;  /* N, post-sigmoid */
  float* shs;      /* N x 48, 16 coefficients x RGB interleaved */
  int n_nodes;
  float* node_pos; /* M x 3 */
  double* node_rot;   /* M x 9 column-major (DeformGraph::rot) of the last solve */
  double* node_trans; /* M x 3 */
  long long n_samples;
  float* sample_pos;     /* S x 3 */
  float* aim_feature;    /* S x 48 */
  float* aim_opacity;    /* S */
  int* valid_grid;       /* V */
  int* grid_gs_prefix_sum; /* G^3, inclusive */
  int* grided_gs_idx;    /* pairs */
  int* gs_init_grid_idx; /* N */
  float* ada_lpf_ratio;  /* G^3 x 9 */
  float* end_points;     /* N x 6 x 3 */
  int* empty_grid;       /* V: JudgeEmptyGrid flags (GV:4272-4318), set by arap_grid_eval(ctx, 0); NULL before */
  float* cur_feature;    /* S x 48: arap_grid_eval(ctx, 1) output (UpdateFeatures, GV:4159-4186); NULL before */
  float* cur_opacity;    /* S */
} arap_device_view;

/* Device time (ms, CUDA events on the ctx stream) of the last run of each set-up / stroke-end stage, read with
 * arap_setup_timing.  T_stroke = SCENE_AABB + FOOTPRINT_LISTS + GRID_EVAL of an arap_grid_update_lists +
 * arap_grid_eval(ctx, 1) pair; T_graph = FPS + NODE_GRAPH + KNN_ENDS + KNN_SAMPLES + TILE_TABLES of a graph build. */
enum {
  ARAP_ST_SCENE_AABB = 0,     /* getOverallAABB */
  ARAP_ST_CELL_ASSIGN = 1,    /* GetGsGrid + scan, both passes of arap_grid_build (a2) */
  ARAP_ST_REORDER = 2,        /* cell-order permutation of the SoA (a2) */
  ARAP_ST_FOOTPRINT_LISTS = 3,/* cutoff boxes + count + scan + fill + per-cell order (a3-a5) */
  ARAP_ST_SAMPLES = 4,        /* valid cells, sample emit, rest-state adaptive LPF, end points (a6, a8, d1) */
  ARAP_ST_GRID_EVAL = 5,      /* forward3d_grid (+ JudgeEmptyGrid for the aim field) (a7) */
  ARAP_ST_FPS = 6,            /* farthest_control_points_sampling (b5) */
  ARAP_ST_NODE_GRAPH = 7,     /* bucket index + node kNN + edges (b3) */
  ARAP_ST_KNN_ENDS = 8,       /* setupWeightsforEnds: 6N queries (b1, b2, b4) */
  ARAP_ST_KNN_SAMPLES = 9,    /* setupWeightsforSamples: S queries */
  ARAP_ST_TILE_TABLES = 10,   /* per-tile node lists / slots of the skinning kernels (ours) */
  ARAP_SETUP_STAGES = 11
};
int arap_setup_timing(arap_ctx* ctx, float* ms, int n /* <= ARAP_SETUP_STAGES */);

/* ---- lifecycle ----------------------------------------------------------- */
const char* arap_last_error(void);
int arap_version(void);
int arap_default_params(arap_params* p);
/* stream: a cudaStream_t (or NULL for a ctx-owned stream). Replaces the cudaMalloc block of GaussianView::init (GV:611-841). */
int arap_create(arap_ctx** out, int device, void* stream, const arap_params* params);
int arap_destroy(arap_ctx* ctx);
int arap_set_params(arap_ctx* ctx, const arap_params* params);
int arap_sync(arap_ctx* ctx);

/* ---- Gaussians ----------------------------------------------------------- */
/* SoA upload (post-activation values, as loadPly_Origin produces: GV:43-153). Host or device source. */
int arap_set_gaussians(arap_ctx* ctx, long long n, const float* pos, const float* rot, const float* scale,
                       const float* opacity, const float* shs, int src_is_device);
int arap_download_gaussians(arap_ctx* ctx, float* pos, float* rot, float* scale, float* opacity, float* shs);
int arap_get_device_view(arap_ctx* ctx, arap_device_view* out);

/* ---- stage (a): density grid ---------------------------------------------- */
/* getOverallAABB + getGridSamples + GetAdaLpfRatio (GV:3601-3631, 3870-4149, 4670-4751).
 * Re-orders the Gaussians into cell order exactly like GV:3938-3953. */
int arap_grid_build(arap_ctx* ctx);
/* UpdateContainingRelationship (GV:3634-3743): rebuild boxes/lists for the current Gaussians, keeping samples. */
int arap_grid_update_lists(arap_ctx* ctx);
/* Rasterizer::forward3d_grid call sites (GV:4159-4186 cur, 4222-4249 aim). which: 0 = aim, 1 = current. */
int arap_grid_eval(arap_ctx* ctx, int which);
/* GetAdaLpfRatio (GV:4670-4751) on the current, deformed sample positions (call sites GV:958, 1704-1706);
 * arap_grid_build computes the rest-state value (GV:725). */
int arap_ada_lpf_update(arap_ctx* ctx);
int arap_download_ada_lpf(arap_ctx* ctx, float* out /* G^3 x 9 */);
/* JudgeEmptyGrid (GV:4272-4318), evaluated on device at the end of arap_grid_eval(ctx, 0). */
int arap_download_empty_grid(arap_ctx* ctx, int* out /* V */);
int arap_grid_info_get(arap_ctx* ctx, arap_grid_info* out);
int arap_download_grid(arap_ctx* ctx, int* valid, int* prefix, int* lists, float* sample_pos, int* gs_init_grid_idx);
int arap_download_features(arap_ctx* ctx, int which, float* feature, float* opacity);
int arap_download_samples(arap_ctx* ctx, float* sample_pos, float* aim_feature);
/* arap_params.lazy_sample_sh = 1: rotate the aim features by the rotations accumulated since the last call (stream-ordered; no-op otherwise) */
int arap_sample_features_materialize(arap_ctx* ctx);

/* ---- stage (b): graph, kNN, weights --------------------------------------- */
/* farthest_control_points_sampling over the Gaussian centres + DeformGraph ctor + setupWeights* (GV:698-754, HC:139-195, DH:54-95). */
int arap_graph_build_fps(arap_ctx* ctx, int node_num, int k);
/* LoadDeformation: nodes given as Gaussian (or mesh point) indices (GV:4961-5010). */
int arap_graph_build_anchors(arap_ctx* ctx, const int* anchor_idx, int m, int k);
/* LoadMeshForGraph + "Build Graph on Mesh" + RebuildGraph (GV:4939-4959, 4825-4856): mesh points become the
 * candidate points; nodes = FPS over them with node_num = all. on_mesh=0 keeps nodes on Gaussians but still skins the mesh points. */
int arap_set_mesh_points(arap_ctx* ctx, const float* pts, int n, int nodes_on_mesh);
/* The reference's other skinned point sets: family 0 = mesh_points of the textured mesh (setupWeightsforMesh GV:2834-2845),
 * family 1 = soup_points of <ply>_soup.obj (setupWeightsforSoup GV:2862-2873); both are moved every step by predict_mesh
 * (UpdatePosition GV:2989-3020) with the bit-faithful kernel.  Call before the graph build.  arap_download_points: family -1 =
 * the simplified points of arap_set_mesh_points. */
#define ARAP_POINT_FAMILIES 2
int arap_set_points(arap_ctx* ctx, int family, const float* pts, long long n);
int arap_download_points(arap_ctx* ctx, int family, float* pts);
/* DeformGraph::computeWeights for arbitrary device queries (DH:187-208): idx Q x k (uint32), w Q x k (double), device outputs. */
int arap_knn_weights(arap_ctx* ctx, const float* queries_dev, long long q, int k, uint32_t* idx_dev, double* w_dev);
int arap_download_graph(arap_ctx* ctx, int* anchor, float* node_pos, int* nbr /* M x k */);
int arap_download_rows(arap_ctx* ctx, int family /*0 ends,1 samples,2 mesh,3 nodes*/, uint32_t* idx, double* w);

/* ---- control regions (GV:1169-1387, 1996-2109) ---------------------------- */
/* blocks as CSR over node ids; types: 1 active, 0 pinned, -1 excluded. Absorbs UpdateIndicies + CheckStaticSamples. */
int arap_set_blocks(arap_ctx* ctx, int n_blocks, const int* block_off, const uint32_t* block_nodes, const int* block_types);
int arap_download_static_flags(arap_ctx* ctx, uint8_t* gaussians, uint8_t* samples);
/* the six end points per Gaussian (GetEndPoints GV:4643-4668, skinned every step GV:3021-3040): N x 18 floats */
int arap_download_end_points(arap_ctx* ctx, float* ends);

/* ---- aims (GV:2920-2983, 5240-5247) ---------------------------------------- */
int arap_aim_translate(arap_ctx* ctx, const float delta[3]);                 /* UpdateAimPosition */
int arap_aim_twist(arap_ctx* ctx, const float axis[4], int y);               /* UpdateAimPositionTwist */
int arap_aim_scale(arap_ctx* ctx, int y);                                    /* UpdateAimPositionScale */
int arap_aim_set(arap_ctx* ctx, const float* aim /* M x 3 host */);          /* script aims (GV:1924-1937) */
int arap_aim_get(arap_ctx* ctx, float* aim);
int arap_aim_reload(arap_ctx* ctx);                                          /* ReloadAimPositions */

/* ---- stage (c) + (d): one drag step (GV:1481-1522) ------------------------- */
/* Deform deform(...); real_time_deform(); putFreeInputs  (DC:77-169, DH:140-151) */
int arap_solve(arap_ctx* ctx, int constraints_on_center);
int arap_solve_stats_get(arap_ctx* ctx, arap_solve_stats* out); /* synchronises */
/* UpdatePositionforSamples; UpdatePosition; UpdateAsSixPointsWithdrawBad; FastUpdateSamplesSH; ReloadAimPositions; resetRT */
int arap_apply(arap_ctx* ctx);
/* solve + apply */
int arap_step(arap_ctx* ctx, int constraints_on_center);
/* Makes `stream` (a cudaStream_t) wait until the deformed SoA (arap_device_view: pos/rot/scale/shs) of the last
 * arap_apply / arap_step is final, i.e. until its six-point fit has run — the per-step sample SH pass that follows
 * (GV:1519) does not touch the SoA.  Lets a consumer (the rasteriser, or the multi-GPU all-gather of the deformed
 * Gaussians) start on its own stream while that pass is still running.  The caller must make the ctx stream wait
 * for the consumer before the next arap_apply overwrites the SoA.  No reference counterpart (GV:1640-1647 copies
 * after the whole step). */
int arap_soa_ready_wait(arap_ctx* ctx, void* stream);
/* The other half: `event` (a cudaEvent_t the caller recorded on its own stream after its last read of the SoA; not
 * owned, must stay alive until the next arap_apply) — the next arap_apply makes the ctx stream wait for it right
 * before the kernel that overwrites the SoA (the six-point fit), so the consumer also overlaps the next solve and
 * end-point pass.  One-shot: consumed by that arap_apply. */
int arap_soa_release_event(arap_ctx* ctx, void* event);
int arap_download_nodes(arap_ctx* ctx, float* node_pos, double* rot, double* trans);
/* per-stage device time of the last arap_step in ms: [0] solve [1] samples lbs [2] endpoints+mesh+nodes lbs [3] fit [4] sample SH [5] total */
int arap_last_step_timing(arap_ctx* ctx, float* ms6);
int arap_enable_timing(arap_ctx* ctx, int on);
/* stage timings (6 floats per step, same order) of the most recent steps since timing was enabled, oldest first (ring of 128) */
int arap_step_timings(arap_ctx* ctx, float* ms, int max_steps, int* n_out);

/* ---- multi-GPU: exchange of the deformed Gaussians between the ranks of one node --------------------
 * One process (and one arap_ctx) per GPU; every rank owns a contiguous index range of the Gaussians (equal counts) and runs
 * the replicated node solve.  No reference counterpart (the reference is single-GPU; its per-frame upload GV:1640-1647 is what
 * the exchange replaces on the non-owning ranks).  Host protocol:
 *   rank 0: arap_comm_unique_id(id) -> hand the 128 bytes to the other ranks out of band (MPI, a socket, torch.distributed ...)
 *   all:    arap_set_gaussians(own shard) ... arap_comm_init(ctx, id, rank, world)   [collective; fetch arap_device_view again:
 *           the SoA now lives inside the gathered arrays]
 *   step:   arap_step(ctx, ...); arap_comm_exchange(ctx);                            [asynchronous: overlaps the sample passes]
 *   frame:  arap_comm_materialize_sh(ctx); arap_comm_sync(ctx); read arap_comm_view  [only when this rank consumes remote SH rows]
 * arap_comm_exchange all-gathers pos / rot / scale (40 B per Gaussian, in place, one grouped NCCL call over NVLink); the SH rows
 * of remote Gaussians are brought up to date lazily by arap_comm_materialize_sh from the gathered rotations. */
typedef struct arap_gathered_view {
  int rank, world;
  long long n_per_rank;
  float* pos;    /* world * n x 3, rank-major */
  float* rot;    /* world * n x 4 */
  float* scale;  /* world * n x 3 */
  float* shs;    /* world * n x 48; remote ranges valid after arap_comm_materialize_sh + arap_comm_sync */
  void* side_stream; /* the cudaStream_t the exchange runs on */
} arap_gathered_view;
int arap_comm_unique_id(char id_out[128]);
int arap_comm_init(arap_ctx* ctx, const char id[128], int rank, int world);
int arap_comm_exchange(arap_ctx* ctx);
int arap_comm_materialize_sh(arap_ctx* ctx);
int arap_comm_sync(arap_ctx* ctx);
int arap_comm_view(arap_ctx* ctx, arap_gathered_view* out);
int arap_comm_destroy(arap_ctx* ctx);
/* ONE scene sharded over the ranks (BASELINE configs[4]).  Rank r holds the Gaussians [r n, (r + 1) n) of a
 * scene that is already in cell order (the order a single-GPU arap_grid_build leaves behind; contiguous index ranges are then
 * x-slabs up to boundary effects).  arap_comm_grid_build replaces arap_grid_build for such a session (after arap_comm_init):
 * one grid over all ranks' Gaussians (scene box GV:3601-3631 over the gathered positions, identical on every rank), of which
 * this rank bins (GV:3961-4100) and evaluates (GV:4159-4186) only its x-slab of cells [x_lo, x_hi) — including the Gaussians
 * of other ranks whose padded footprint reaches into the slab (the halo; they are read from the gathered arrays, no extra
 * exchange).  x_lo < 0: automatic cuts balanced by the number of Gaussians per x-layer (same result on every rank).
 * The samples, their skinning tables, the per-step sample passes and the stroke-end rebuild (arap_grid_update_lists: call
 * arap_comm_exchange after the last step first) then cover that slab only; the union over the ranks is the single-GPU grid. */
int arap_comm_grid_build(arap_ctx* ctx, int x_lo, int x_hi);
/* mode 0 (default): arap_comm_exchange is an in-place NCCL all-gather.  mode 1: the exchange is FUSED into the apply kernel —
 * every tile's final pos / rot / scale is stored straight into the other ranks' gathered arrays over NVLink (peer mappings via
 * cudaIpc: one process per GPU, one node, <= 8 ranks), ordered by per-rank epoch flags instead of a collective; arap_comm_exchange
 * then only waits for the ranks' "done" flags.  mode 2: the same with ONE store per value into an NVSwitch multicast mapping of the
 * gathered pose arrays (multimem.st; CUDA driver VMM API, the multicast object's descriptor is handed from rank 0 to the other
 * processes over a unix socket): the switch replicates the store into every rank's copy, so a rank's NVLink egress is 40 bytes
 * per Gaussian whatever the number of ranks (mode 1: 40 bytes per peer).  ARAP_ERR_UNSUPPORTED (on every rank alike) if the
 * platform has no multicast; the pose pointers of arap_device_view / arap_comm_view change (fetch them again).
 * Collective (call on every rank after arap_comm_init, once, from mode 0); needs lbs_mode = 3.
 * Like the all-gather the fused modes require every rank to run the same sequence of arap_step / arap_apply calls. */
int arap_comm_set_mode(arap_ctx* ctx, int mode);
int arap_comm_slab_get(arap_ctx* ctx, int* x_lo, int* x_hi);

/* ---- deform.txt / graph.obj / config / scripts (host IO, byte-compatible) -- */
typedef struct arap_history arap_history;
int arap_history_load(const char* path, arap_history** out);                 /* LoadDeformation grammar (GV:4961-5095) */
int arap_history_save(const arap_history* h, const char* path);              /* RecordDeformation (GV:4859-4935) */
int arap_history_free(arap_history* h);
int arap_history_new(arap_history** out, int nodes_on_mesh, const int* node_anchor, int n_nodes);
int arap_history_add_block(arap_history* h, const uint32_t* nodes, int n);
int arap_history_add_move(arap_history* h, int op_type, const float* movements /* n x 3 */, int n, const int* block_types,
                          int n_types, int energy_on_center, const float twist_axis[4]);
/* summary: [0] nodes_on_mesh [1] n_nodes [2] total_ops [3] move_ops [4] n_blocks [5] n_moves */
int arap_history_summary(const arap_history* h, int* out6);
int arap_history_nodes(const arap_history* h, int* out);
int arap_history_ops(const arap_history* h, int* out);
int arap_history_block(const arap_history* h, int i, uint32_t* nodes_out, int* n);
int arap_history_move(const arap_history* h, int i, float* movements_out, int* n, int* block_types_out, int* n_types,
                      int* energy_on_center, float* twist_axis4);
/* LoadMeshPoints: `v x y z` lines only (HC:264-282). Two-call pattern: pts==NULL returns the count. */
int arap_graph_obj_load(const char* path, float* pts, int* n);
int arap_graph_obj_save(const char* path, const float* pts, int n);          /* writeVectorToObj (HC:1111-1126) */
/* loadPly_Origin / loadPly (GV:43-271): 3DGS point_cloud.ply, binary little-endian, 62 float properties per vertex (63 with the
 * trailing `index` of the reference's "soup" files).  Applies the reference's activations (normalised quaternion w,x,y,z; exp
 * scale; sigmoid opacity), interleaves f_rest into 16 x RGB and orders the Gaussians by the Morton code of their position in
 * the cloud's bounding box (ties keep file order; the reference's std::sort leaves them unspecified).  Two-call pattern:
 * pos == NULL returns the vertex count in *n; otherwise *n is the capacity of the arrays on entry.  rot/scale/opacity/shs/index/
 * aabb may be NULL.  index = the file's `index` property when present, else the vertex's position in the file. */
int arap_ply_load(const char* path, long long* n, float* pos, float* rot, float* scale, float* opacity, float* shs, int* index,
                  float aabb_min[3], float aabb_max[3]);
/* savePly (GV:273-339): header text identical to the reference's; Gaussians outside [box_min, box_max] (NULL = keep all) or with
 * skip[i] != 0 (the reference's too_small) are dropped; *written (may be NULL) = vertices written. */
int arap_ply_save(const char* path, long long n, const float* pos, const float* rot, const float* scale, const float* opacity,
                  const float* shs, const float box_min[3], const float box_max[3], const uint8_t* skip, long long* written);
/* <ply>_config.txt: grid_num is_synthetic conf2 (GV:459-487) */
int arap_config_load(const char* path, int* grid_num, int* is_synthetic, int* has_soup, int* high_quality);
/* Run a loaded history through the state machine of GV:1757-1916 (blocks added/deleted, per-move block types,
 * op types 1..4).  Rebuilds the graph from the recorded anchors first when rebuild != 0 (LoadDeformation) or keeps the
 * current graph (LoadDeformation_wo_rebuild).  Returns the number of drag steps run. */
int arap_replay(arap_ctx* ctx, const arap_history* h, int rebuild_graph, int* steps_run);
/* LoadDeformScript0/1 (GV:2512-2600) on the current graph + the script loop of GV:1918-1993. */
int arap_run_script(arap_ctx* ctx, int script_id, int* steps_run);
/* PointRotateByAxis (HC:1077-1100), exposed for tests */
int arap_point_rotate_by_axis(const float point[3], const float center[3], const float axis[4], float radian, float out[3]);

#ifdef __cplusplus
}
#endif
#endif /* ARAPGS_H */
