// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17 + OpenMP, no Eigen: Eigen is absent from this image)
// of the reference's ARAP deformation path.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// PARITY STATUS: the reference has no tests and cannot be built here (no Eigen,
// no GL stack, CudaRasterizer fetched from the network).  This restatement is
// pinned by the reference's fixtures only (FPS known-answer on graph.obj,
// deform.txt / config grammar, analytic invariants — see tests/test_oracle_*).
// Numeric outputs of solve / apply are "parity unpinned": they follow the
// reference's arithmetic operation by operation, with the few Eigen-internal
// summation orders (un-vendored dependency, version unpinned) stated inline.
//
// Every function cites the reference file:line it follows.  Abbreviations:
//   GV  = src/projects/gaussianviewer/renderer/GaussianView.cpp
//   DH  = .../renderer/Deform.hpp      DC = .../renderer/Deform.cpp
//   HC  = .../renderer/helper.cpp      CK = .../renderer/cudakdtree.cu
//
// Build: g++ -O3 -fopenmp -shared -fPIC (no -march=native, no -ffast-math,
// -ffp-contract=off) — the reference adds neither (CMakeLists.txt:91).

#include <algorithm>
#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>
#include <omp.h>

#define ORC_API extern "C" __attribute__((visibility("default")))

namespace {

constexpr int KNN_MAX = 12;  // helper.hpp:48

// ---------------------------------------------------------------------------
// small 3x3 helpers (double)
// ---------------------------------------------------------------------------
struct M3d { double m[3][3]; };

// One-sided cyclic Jacobi eigen-decomposition of a symmetric 3x3 (double).
// Used to restate Eigen::JacobiSVD -> polar factors (HC:429-440): the polar
// decomposition M = R S is unique for non-singular M, so any accurate double
// algorithm reproduces U V^T and V Sigma V^T to ~1e-15.
static void sym_eig3(const double A_in[3][3], double V[3][3], double w[3]) {
  double A[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { A[i][j] = A_in[i][j]; V[i][j] = (i == j); }
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * diag || off == 0.0) break;
    for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
      if (A[p][q] == 0.0) continue;
      double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
      double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 3; k++) {  // A <- A J
        double akp = A[k][p], akq = A[k][q];
        A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
      }
      for (int k = 0; k < 3; k++) {  // A <- J^T A
        double apk = A[p][k], aqk = A[q][k];
        A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
      }
      for (int k = 0; k < 3; k++) {
        double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
      }
    }
  }
  for (int i = 0; i < 3; i++) w[i] = A[i][i];
}

// polar decomposition M = R S (S symmetric PSD) in double; HC:429-440
// (getOthogonalMatrixWithK: R = U V^T, S = V Sigma V^T, K = diag(S)).
static void polar3(const double M[3][3], double R[3][3], double S[3][3]) {
  double MtM[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0; for (int k = 0; k < 3; k++) s += M[k][i] * M[k][j];
    MtM[i][j] = s;
  }
  double V[3][3], w[3];
  sym_eig3(MtM, V, w);
  double sg[3], isg[3];
  for (int i = 0; i < 3; i++) { sg[i] = std::sqrt(std::max(w[i], 0.0)); isg[i] = sg[i] > 0 ? 1.0 / sg[i] : 0.0; }
  double Sinv[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0, si = 0;
    for (int k = 0; k < 3; k++) { s += V[i][k] * sg[k] * V[j][k]; si += V[i][k] * isg[k] * V[j][k]; }
    S[i][j] = s; Sinv[i][j] = si;
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    double s = 0; for (int k = 0; k < 3; k++) s += M[i][k] * Sinv[k][j];
    R[i][j] = s;
  }
}

// ---------------------------------------------------------------------------
// quaternions (float), Eigen 3.4 public algorithms.  Storage here: w,x,y,z.
// ---------------------------------------------------------------------------
struct Qf { float w, x, y, z; };

// Eigen::Quaternionf(Matrix3f) — SURVEY Appendix B.4 (Eigen/src/Geometry/Quaternion.h,
// quaternionbase_assign_impl<Other,3,3>); m is row-major m[r][c].
static Qf quat_from_matrix(const float m[3][3]) {
  Qf q; float qv[3];
  float t = m[0][0] + m[1][1] + m[2][2];
  if (t > 0.0f) {
    t = std::sqrt(t + 1.0f);
    q.w = 0.5f * t; t = 0.5f / t;
    q.x = (m[2][1] - m[1][2]) * t; q.y = (m[0][2] - m[2][0]) * t; q.z = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[i][i] - m[j][j] - m[k][k] + 1.0f);
    qv[i] = 0.5f * t; t = 0.5f / t;
    q.w = (m[k][j] - m[j][k]) * t;
    qv[j] = (m[j][i] + m[i][j]) * t;
    qv[k] = (m[k][i] + m[i][k]) * t;
    q.x = qv[0]; q.y = qv[1]; q.z = qv[2];
  }
  return q;
}
// squaredNorm over coeffs (x,y,z,w): SSE2 predux order (a0+a2)+(a1+a3) — Eigen
// internal, parity unpinned at 1 ulp.
static inline float quat_n2(const Qf& q) { return (q.x * q.x + q.z * q.z) + (q.y * q.y + q.w * q.w); }
static Qf quat_normalized(const Qf& q) {
  float n = std::sqrt(quat_n2(q));
  return Qf{q.w / n, q.x / n, q.y / n, q.z / n};
}
static Qf quat_mul(const Qf& a, const Qf& b) {
  Qf r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
static Qf quat_inverse(const Qf& q) {
  float n2 = quat_n2(q);
  if (n2 > 0.0f) return Qf{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return Qf{0, 0, 0, 0};
}
// Eigen QuaternionBase::toRotationMatrix
static void quat_to_matrix(const Qf& q, float m[3][3]) {
  float tx = 2.0f * q.x, ty = 2.0f * q.y, tz = 2.0f * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  m[0][0] = 1.0f - (tyy + tzz); m[0][1] = txy - twz; m[0][2] = txz + twy;
  m[1][0] = txy + twz; m[1][1] = 1.0f - (txx + tzz); m[1][2] = tyz - twx;
  m[2][0] = txz - twy; m[2][1] = tyz + twx; m[2][2] = 1.0f - (txx + tyy);
}

// ---------------------------------------------------------------------------
// SH rotation, degree 1..3 (HC:938-1075; device twin CK:11-199).
// The reference's Construct_SH_Rotation_Matrix is the Ivanic-Ruedenberg
// recurrence written out entry by entry; this is the same recurrence in loop
// form: float products inside P(), double coefficient combination, one
// rounding to float per entry (as the reference's `double * float-expr`
// assignments do).
// ---------------------------------------------------------------------------
struct SHRot { float b1[3][3], b2[5][5], b3[7][7]; };

template <int L, typename Prev>
static inline float shP(int i, int a, int b, const float r1[3][3], const Prev& prev) {
  // prev is band L-1, centred indexing via offset (L-1)
  const int o = L - 1;
  const float ri1 = r1[i + 1][2], rim1 = r1[i + 1][0], ri0 = r1[i + 1][1];
  if (b == L) return ri1 * prev[a + o][L - 1 + o] - rim1 * prev[a + o][-L + 1 + o];
  if (b == -L) return ri1 * prev[a + o][-L + 1 + o] + rim1 * prev[a + o][L - 1 + o];
  return ri0 * prev[a + o][b + o];
}

template <int L, typename Prev, typename Out>
static void sh_band(const float r1[3][3], const Prev& prev, Out& out) {
  for (int m = -L; m <= L; m++) for (int n = -L; n <= L; n++) {
    const int d = (m == 0);
    const int am = std::abs(m);
    const double denom = (std::abs(n) == L) ? double(2 * L * (2 * L - 1)) : double((L + n) * (L - n));
    const double u = std::sqrt(double((L + m) * (L - m)) / denom);
    const double v = 0.5 * std::sqrt(double((1 + d) * (L + am - 1) * (L + am)) / denom) * (1 - 2 * d);
    const double w = -0.5 * std::sqrt(double((L - am - 1) * (L - am)) / denom) * (1 - d);
    double acc = 0.0;
    if (u != 0.0) acc += u * (double)shP<L>(0, m, n, r1, prev);
    if (v != 0.0) {
      double V;
      if (m == 0) V = (double)(shP<L>(1, 1, n, r1, prev) + shP<L>(-1, -1, n, r1, prev));
      else if (m > 0) {
        if (m == 1) V = std::sqrt(2.0) * (double)shP<L>(1, 0, n, r1, prev);
        else V = (double)(shP<L>(1, m - 1, n, r1, prev) - shP<L>(-1, -m + 1, n, r1, prev));
      } else {
        if (m == -1) V = std::sqrt(2.0) * (double)shP<L>(-1, 0, n, r1, prev);
        else V = (double)(shP<L>(1, m + 1, n, r1, prev) + shP<L>(-1, -m - 1, n, r1, prev));
      }
      acc += v * V;
    }
    if (w != 0.0) {
      double W;
      if (m > 0) W = (double)(shP<L>(1, m + 1, n, r1, prev) + shP<L>(-1, -m - 1, n, r1, prev));
      else W = (double)(shP<L>(1, m - 1, n, r1, prev) - shP<L>(-1, -m + 1, n, r1, prev));
      acc += w * W;
    }
    out[m + L][n + L] = (float)acc;
  }
}

static void sh_rotation_matrices(const float R[3][3], SHRot& s) {
  // band 1 in (y,z,x) order: sh1(i,j) = R((i+1)%3,(j+1)%3)   (HC:984-992)
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) s.b1[i][j] = R[(i + 1) % 3][(j + 1) % 3];
  sh_band<2>(s.b1, s.b1, s.b2);
  sh_band<3>(s.b1, s.b2, s.b3);
}

// rotate 16x3 interleaved SH coefficients in place (HC:938-975), without the
// odd-index sign flips (callers add them: GV:3138-3154 / CK:160-196).
static void sh_apply(const SHRot& s, float* shs) {
  for (int c = 0; c < 3; c++) {
    float in[16], out[16];
    for (int i = 0; i < 16; i++) in[i] = shs[i * 3 + c];
    out[0] = in[0];
    for (int i = 0; i < 3; i++) { float a = 0; for (int k = 0; k < 3; k++) a += s.b1[i][k] * in[1 + k]; out[1 + i] = a; }
    for (int i = 0; i < 5; i++) { float a = 0; for (int k = 0; k < 5; k++) a += s.b2[i][k] * in[4 + k]; out[4 + i] = a; }
    for (int i = 0; i < 7; i++) { float a = 0; for (int k = 0; k < 7; k++) a += s.b3[i][k] * in[9 + k]; out[9 + i] = a; }
    for (int i = 0; i < 16; i++) shs[i * 3 + c] = out[i];
  }
}
static void sh_rotate_flipped(const float R[3][3], float* shs) {
  SHRot s; sh_rotation_matrices(R, s);
  for (int k = 1; k < 16; k += 2) { shs[k * 3] = -shs[k * 3]; shs[k * 3 + 1] = -shs[k * 3 + 1]; shs[k * 3 + 2] = -shs[k * 3 + 2]; }
  sh_apply(s, shs);
  for (int k = 1; k < 16; k += 2) { shs[k * 3] = -shs[k * 3]; shs[k * 3 + 1] = -shs[k * 3 + 1]; shs[k * 3 + 2] = -shs[k * 3 + 2]; }
}

// ---------------------------------------------------------------------------
// kNN: DeformGraph::findNearestNodes (DH:153-185) — literal.
// ---------------------------------------------------------------------------
static inline float knn_dist(const float* p, const float* n) {
  float t0 = p[0] - n[0], t1 = p[1] - n[1], t2 = p[2] - n[2];
  return std::sqrt(t0 * t0 + t1 * t1 + t2 * t2);
}
static void find_nearest(const float* nodes, int M, const float* pos, int kq, uint32_t* idx,
                         std::vector<double>& d, std::vector<uint32_t>& index) {
  d.resize(M); index.resize(M);
  for (int j = 0; j < M; j++) { index[j] = j; d[j] = knn_dist(pos, nodes + 3 * j); }
  for (int i = 0; i < kq; i++) {  // selection sort, scan from the end, strict <
    int m = i;
    for (int j = M - 1; j > i; j--) if (d[j] < d[m]) m = j;
    idx[i] = index[m];
    std::swap(d[i], d[m]); std::swap(index[i], index[m]);
  }
}

}  // namespace

// ===========================================================================
// exported C API (ctypes)
// ===========================================================================

ORC_API int orc_num_threads() { return omp_get_max_threads(); }
ORC_API void orc_set_threads(int n) { omp_set_num_threads(n); }

// --- farthest_control_points_sampling + FetchFirstNodeIdx (HC:139-195) -----
// pts_distance (HC:60-63): std::pow(float,int) promotes to double, so the
// squares and their sum are double, sqrt is double, the result is rounded to
// float on return.
ORC_API int orc_fps(const float* pos, int n, int node_num, int* out) {
  int m = std::min(node_num, n);
  if (m <= 0) return 0;
  int first = 0; float best = (float)-100000.0;  // `Infinity` macro, HC:47
  for (int i = 0; i < n; i++) {
    float c = 0.0f; for (int j = 0; j < 3; j++) c += pos[3 * i + j];
    if (c > best) { best = c; first = i; }
  }
  int cnt = 0; out[cnt++] = first;
  std::vector<float> dist(n, FLT_MAX);
  // the `if (first_node_idx == i) continue;` quirk (HC:164) only skips one
  // loop index; the body does not depend on i, so the sequence is unchanged.
  while (cnt < m) {
    const float* l = pos + 3 * out[cnt - 1];
#pragma omp parallel for schedule(static)
    for (int j = 0; j < n; j++) {
      double dx = (double)(l[0] - pos[3 * j]), dy = (double)(l[1] - pos[3 * j + 1]), dz = (double)(l[2] - pos[3 * j + 2]);
      float dd = (float)std::sqrt(dx * dx + dy * dy + dz * dz);
      if (dd < dist[j]) dist[j] = dd;
    }
    int arg = 0; float mx = dist[0];  // std::max_element: first maximum
    for (int j = 1; j < n; j++) if (dist[j] > mx) { mx = dist[j]; arg = j; }
    out[cnt++] = arg;
  }
  return cnt;
}

// --- findNearestNodes + computeWeights (DH:153-208) ------------------------
// idx_out: Q x (k+1) uint32; w_out: Q x k double (may be null -> kNN only).
ORC_API void orc_knn_weights(const float* nodes, int M, const float* queries, long long Q, int k,
                             uint32_t* idx_out, double* w_out) {
#pragma omp parallel
  {
    std::vector<double> d; std::vector<uint32_t> index;
#pragma omp for schedule(dynamic, 64)
    for (long long qi = 0; qi < Q; qi++) {
      const float* p = queries + 3 * qi;
      uint32_t* idx = idx_out + qi * (k + 1);
      find_nearest(nodes, M, p, k + 1, idx, d, index);
      if (!w_out) continue;
      double* w = w_out + qi * k;
      double dmax = knn_dist(p, nodes + 3 * idx[k]);
      double sum = 0.0;
      for (int j = 0; j < k; j++) {
        double dist = knn_dist(p, nodes + 3 * idx[j]);
        double u = 1.0 - dist / dmax;
        w[j] = u * u;  // pow(x, 2.0)
        sum += w[j];
      }
      if (k == 1) w[0] = 1.0; else for (int j = 0; j < k; j++) w[j] /= sum;
    }
  }
}

// --- setupEdges (DH:84-95): neighbours = idx[1..k] of a k+1 query ----------
ORC_API void orc_graph_edges(const float* nodes, int M, int k, uint32_t* nbr_out) {
  std::vector<uint32_t> idx((size_t)M * (k + 1));
  orc_knn_weights(nodes, M, nodes, M, k, idx.data(), nullptr);
  for (int i = 0; i < M; i++) for (int j = 1; j <= k; j++) nbr_out[(size_t)i * k + j - 1] = idx[(size_t)i * (k + 1) + j];
}

// --- LBS: predict_mesh / predict_samples (DH:230-268) ----------------------
// points: P x 3 float, updated in place.  rows: idx P x k (uint32), w P x k
// (double).  node_pos M x 3 float, rot M x 9 double (column-major), trans M x 3.
// skip: optional P flags (static samples are skipped, GV:3064).
// Double products, float accumulator rounded after each neighbour.
ORC_API void orc_lbs_points(float* points, long long P, int k, const uint32_t* idx, const double* w,
                            const float* node_pos, const double* rot, const double* trans,
                            const int* skip) {
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < P; i++) {
    if (skip && skip[i]) continue;
    float cur[3] = {points[3 * i], points[3 * i + 1], points[3 * i + 2]};
    float out[3] = {0.0f, 0.0f, 0.0f};
    for (int j = 0; j < k; j++) {
      uint32_t nd = idx[i * k + j];
      const float* pn = node_pos + 3 * nd;
      const double* r = rot + 9 * nd; const double* t = trans + 3 * nd;
      float t0 = cur[0] - pn[0], t1 = cur[1] - pn[1], t2 = cur[2] - pn[2];
      double wj = w[i * k + j];
      out[0] += wj * (r[0] * t0 + r[3] * t1 + r[6] * t2 + t[0] + pn[0]);
      out[1] += wj * (r[1] * t0 + r[4] * t1 + r[7] * t2 + t[1] + pn[1]);
      out[2] += wj * (r[2] * t0 + r[5] * t1 + r[8] * t2 + t[2] + pn[2]);
    }
    points[3 * i] = out[0]; points[3 * i + 1] = out[1]; points[3 * i + 2] = out[2];
  }
}

// --- GetEndPoints (GV:4643-4668) -------------------------------------------
// ends: N x 6 x 3, order gaussian*6 + axis*2 + {+,-}
ORC_API void orc_end_points(const float* pos, const float* rot, const float* scale, long long N, float* ends) {
#pragma omp parallel for schedule(static)
  for (long long g = 0; g < N; g++) {
    Qf q = quat_normalized(Qf{rot[4 * g], rot[4 * g + 1], rot[4 * g + 2], rot[4 * g + 3]});
    float R[3][3]; quat_to_matrix(q, R);
    for (int i = 0; i < 3; i++) {
      float e = (scale[3 * g + i] + 1e-3f) * 2.0f;  // axis_padding, end_coeff (GV.hpp:116-117)
      for (int c = 0; c < 3; c++) {
        float v = R[c][i] * e;
        ends[(g * 6 + 2 * i) * 3 + c] = pos[3 * g + c] + v;
        ends[(g * 6 + 2 * i + 1) * 3 + c] = pos[3 * g + c] - v;
      }
    }
  }
}

// --- UpdateAsSixPointsWithdrawBad (GV:3081-3166) ---------------------------
// Inputs: ends N x 18, scale_backup (rest scales) N x 3, static flags (may be
// null).  In/out: pos, rot (w,x,y,z), scale, shs (N x 48).
ORC_API void orc_fit_gaussians(const float* ends, const float* scale_backup, const unsigned char* is_static,
                               long long N, float* pos, float* rot, float* scale, float* shs) {
#pragma omp parallel for schedule(static)
  for (long long g = 0; g < N; g++) {
    if (is_static && is_static[g]) continue;
    Qf oq{rot[4 * g], rot[4 * g + 1], rot[4 * g + 2], rot[4 * g + 3]};
    const float* e = ends + g * 18;
    float c[3];
    // P.rowwise().mean(): Eigen's unrolled redux splits in halves:
    // (p0+(p1+p2)) + (p3+(p4+p5)), then / 6 (Eigen internal, parity unpinned at 1 ulp)
    for (int r = 0; r < 3; r++) {
      float p0 = e[0 + r], p1 = e[3 + r], p2 = e[6 + r], p3 = e[9 + r], p4 = e[12 + r], p5 = e[15 + r];
      c[r] = ((p0 + (p1 + p2)) + (p3 + (p4 + p5))) / 6.0f;
    }
    // M = P * pinv(Q) with Q = [+-e_i]: column i = (P_2i - P_2i+1)/2 (closed form)
    double M[3][3];
    for (int i = 0; i < 3; i++) for (int r = 0; r < 3; r++) {
      float a = e[(2 * i) * 3 + r] - c[r], b = e[(2 * i + 1) * 3 + r] - c[r];
      float mf = 0.5f * a + (-0.5f) * b;
      M[r][i] = (double)mf;
    }
    double R[3][3], S[3][3];
    polar3(M, R, S);
    float Rf[3][3]; float K[3];
    for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) Rf[i][j] = (float)R[i][j]; K[i] = (float)S[i][i]; }
    Qf q = quat_normalized(quat_from_matrix(Rf));
    rot[4 * g] = q.w; rot[4 * g + 1] = q.x; rot[4 * g + 2] = q.y; rot[4 * g + 3] = q.z;
    for (int i = 0; i < 3; i++) {
      float s0 = scale_backup[3 * g + i];
      scale[3 * g + i] = K[i] / ((s0 + 1e-3f) * 2.0f) * s0;
    }
    pos[3 * g] = c[0]; pos[3 * g + 1] = c[1]; pos[3 * g + 2] = c[2];
    Qf rq = quat_normalized(quat_mul(q, quat_inverse(oq)));
    float Rs[3][3]; quat_to_matrix(rq, Rs);
    sh_rotate_flipped(Rs, shs + g * 48);
  }
}

// --- SH helpers exported for direct tests -----------------------------------
ORC_API void orc_sh_rotate(const float* R9_rowmajor, float* shs48, int flip_odd) {
  float R[3][3]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = R9_rowmajor[3 * i + j];
  if (flip_odd) sh_rotate_flipped(R, shs48);
  else { SHRot s; sh_rotation_matrices(R, s); sh_apply(s, shs48); }
}
ORC_API void orc_sh_matrices(const float* R9_rowmajor, float* b1, float* b2, float* b3) {
  float R[3][3]; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[i][j] = R9_rowmajor[3 * i + j];
  SHRot s; sh_rotation_matrices(R, s);
  memcpy(b1, s.b1, sizeof(s.b1)); memcpy(b2, s.b2, sizeof(s.b2)); memcpy(b3, s.b3, sizeof(s.b3));
}
ORC_API void orc_polar(const double* M9_rowmajor, double* R9, double* S9) {
  double M[3][3], R[3][3], S[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[i][j] = M9_rowmajor[3 * i + j];
  polar3(M, R, S);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { R9[3 * i + j] = R[i][j]; S9[3 * i + j] = S[i][j]; }
}

// --- FastUpdateSamplesSH host part (GV:3169-3186; HC:506-517, 816-820) -----
// node rot (M x 9 double, column-major) -> float matrix -> Newton polar
// iteration (tol 1e-6 max-abs, returns the iterate BEFORE the converged one)
// -> quaternion, normalised.  out: M x 4 in Eigen coeff order (x,y,z,w).
ORC_API void orc_node_quats(const double* rot, int M, float* q_xyzw) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < M; n++) {
    float A[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = (float)rot[9 * n + c * 3 + r];
    for (int it = 0; it < 100; it++) {
      // next = 0.5*(M + (M^T)^-1); 3x3 inverse via cofactors / determinant
      float T[3][3]; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) T[r][c] = A[c][r];
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
      float det = T[0][0] * cof[0][0] + T[0][1] * cof[0][1] + T[0][2] * cof[0][2];
      float invdet = 1.0f / det;
      float nx[3][3]; float md = 0.0f;
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        float inv = cof[c][r] * invdet;  // inverse = adjugate / det
        nx[r][c] = 0.5f * (A[r][c] + inv);
        md = std::max(md, std::fabs(A[r][c] - nx[r][c]));
      }
      if (md < 1e-6f) break;
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) A[r][c] = nx[r][c];
    }
    Qf q = quat_normalized(quat_from_matrix(A));
    q_xyzw[4 * n] = q.x; q_xyzw[4 * n + 1] = q.y; q_xyzw[4 * n + 2] = q.z; q_xyzw[4 * n + 3] = q.w;
  }
}

// --- RotateSHs + Q_SlerpCUDA (CK:113-148, 201-222) -------------------------
// w: S x k float, idx: S x k int, node quats (x,y,z,w), feature S x 48 in place.
// The reference negates end_q in global memory when the dot is negative (a
// benign race, q == -q); here the flip is applied to a local copy.
ORC_API void orc_rotate_sample_shs(long long S, int k, const float* w, const int* idx, const float* q_xyzw,
                                   const int* is_static, float* feature) {
#pragma omp parallel for schedule(static)
  for (long long s = 0; s < S; s++) {
    if (is_static && is_static[s]) continue;
    Qf wq{1.0f, 0.0f, 0.0f, 0.0f};
    float last = 0.0f;
    for (int j = 0; j < k; j++) {
      float cw = w[s * k + j];
      float t = cw / (cw + last);
      const float* e = q_xyzw + 4 * idx[s * k + j];
      Qf eq{e[3], e[0], e[1], e[2]};
      // float products summed in float, then widened (CK:117-120)
      float cf = wq.x * eq.x + wq.y * eq.y + wq.z * eq.z + wq.w * eq.w;
      double cosa = (double)cf;
      if (cosa < 0) { eq.x = -eq.x; eq.y = -eq.y; eq.z = -eq.z; eq.w = -eq.w; cosa = -cosa; }
      double rA, rB, td = (double)t;
      if (cosa > 0.99995f) { rA = 1.0f - td; rB = td; }
      else {
        double sina = std::sqrt(1.0f - cosa * cosa);
        double ang = std::atan2(sina, cosa);
        rA = std::sin((1.0f - td) * ang) / sina;
        rB = std::sin(td * ang) / sina;
      }
      Qf l;
      l.x = (float)(rA * wq.x + rB * eq.x); l.y = (float)(rA * wq.y + rB * eq.y);
      l.z = (float)(rA * wq.z + rB * eq.z); l.w = (float)(rA * wq.w + rB * eq.w);
      wq = quat_normalized(l);
      last += cw;
    }
    float R[3][3]; quat_to_matrix(quat_normalized(wq), R);
    sh_rotate_flipped(R, feature + s * 48);
  }
}

// --- CheckStaticSamples / CheckMovedGaussians flags (GV:2024-2109) ---------
// node_static: M flags.  rows: P x k.  group: rows per output flag (1 for
// samples, 6 for Gaussians' endpoints).  out[g] = 1 iff every neighbour of
// every row in the group is static.
ORC_API void orc_static_flags(const uint32_t* idx, long long P, int k, int group, const unsigned char* node_static,
                              unsigned char* out) {
  long long G = P / group;
#pragma omp parallel for schedule(static)
  for (long long g = 0; g < G; g++) {
    unsigned char st = 1;
    for (long long r = g * group; r < (g + 1) * group; r++)
      for (int j = 0; j < k; j++) if (!node_static[idx[r * k + j]]) st = 0;
    out[g] = st;
  }
}

// ===========================================================================
// Gauss-Newton embedded-deformation solve (DC:6-581, DH:414-455)
// ===========================================================================
namespace {

struct SolveProblem {
  int M, k;
  const float* node_pos;       // M x 3 current positions
  const uint32_t* nbr;         // M x k  (Node.Neighbor)
  const uint32_t* anchor_idx;  // M x k  (cand_vertices[Vertex_index].Neighbor_Nodes)
  const double* anchor_w;      // M x k
  std::vector<int> free_rank;  // M: rank in free set or -1
  std::vector<int> free_nodes, static_nodes;
  int n_blocks; const int* block_off; const uint32_t* block_nodes; const int* block_type;
  const float* aim;            // M x 3
  bool on_center;
  double w_rot, w_reg, w_con;  // already square-rooted (DH:452-454)
  double weight_factor;        // control_weight_factor (DH:430); 1.0 in practice
  std::vector<std::vector<uint32_t>> center_sel;  // SelectKeyControls result per block
  std::vector<std::array<float, 3>> center_aim;
  int rows;
  std::vector<std::vector<int>> node_blocks;  // for BelongsToSameBlock
};

static bool same_block(const SolveProblem& P, uint32_t a, uint32_t b) {  // DC:20-32
  for (int ba : P.node_blocks[a]) for (int bb : P.node_blocks[b]) if (ba == bb) return true;
  return false;
}

struct Row { int n; int col[4 * KNN_MAX + 8]; double val[4 * KNN_MAX + 8]; };

// CalcEnergyFunc (DC:378-581)
static void eval_f(const SolveProblem& P, const double* x, std::vector<double>& f) {
  f.assign(P.rows, 0.0);
  const int n = (int)P.free_nodes.size(), k = P.k;
  int index = 0;
  for (int i = 0; i < n; i++) {
    const double* a = x + 12 * i;
    auto dot = [&](int c0, int c1) { return a[3 * c0] * a[3 * c1] + a[3 * c0 + 1] * a[3 * c1 + 1] + a[3 * c0 + 2] * a[3 * c1 + 2]; };
    f[index + 0] = P.w_rot * dot(0, 1); f[index + 1] = P.w_rot * dot(0, 2); f[index + 2] = P.w_rot * dot(1, 2);
    f[index + 3] = P.w_rot * (dot(0, 0) - 1.0); f[index + 4] = P.w_rot * (dot(1, 1) - 1.0); f[index + 5] = P.w_rot * (dot(2, 2) - 1.0);
    index += 6;
  }
  for (int r = 0; r < n; r++) {
    int i = P.free_nodes[r];
    const double* a = x + 12 * r;
    double gj[3] = {P.node_pos[3 * i], P.node_pos[3 * i + 1], P.node_pos[3 * i + 2]};
    for (int t = 0; t < k; t++) {
      int j = P.nbr[(size_t)i * k + t];
      double gk[3] = {P.node_pos[3 * j], P.node_pos[3 * j + 1], P.node_pos[3 * j + 2]};
      double tk[3] = {0, 0, 0};
      int rk = P.free_rank[j];
      if (rk >= 0) { tk[0] = x[12 * rk + 9]; tk[1] = x[12 * rk + 10]; tk[2] = x[12 * rk + 11]; }
      double ws = same_block(P, i, j) ? P.weight_factor : 1.0;
      double d[3] = {gk[0] - gj[0], gk[1] - gj[1], gk[2] - gj[2]};
      for (int c = 0; c < 3; c++) {
        double v = (a[c] * d[0] + a[c + 3] * d[1] + a[c + 6] * d[2]) + gj[c] + a[9 + c] - gk[c] - tk[c];
        f[index + c] = P.w_reg * v * ws;
      }
      index += 3;
    }
  }
  for (int i : P.static_nodes) {
    for (int t = 0; t < k; t++) {
      int q = P.nbr[(size_t)i * k + t];
      int rk = P.free_rank[q];
      for (int c = 0; c < 3; c++) f[index + c] = P.w_reg * (rk >= 0 ? -x[12 * rk + 9 + c] : -0.0);
      index += 3;
    }
  }
  auto skin = [&](uint32_t node, double out[3]) {
    double ve[3] = {P.node_pos[3 * node], P.node_pos[3 * node + 1], P.node_pos[3 * node + 2]};
    out[0] = out[1] = out[2] = 0.0;
    for (int j = 0; j < k; j++) {
      uint32_t q = P.anchor_idx[(size_t)node * k + j];
      double wei = P.anchor_w[(size_t)node * k + j];
      int rk = P.free_rank[q];
      if (rk < 0) { for (int c = 0; c < 3; c++) out[c] += wei * ve[c]; }
      else {
        const double* a = x + 12 * rk;
        double gq[3] = {P.node_pos[3 * q], P.node_pos[3 * q + 1], P.node_pos[3 * q + 2]};
        double d[3] = {ve[0] - gq[0], ve[1] - gq[1], ve[2] - gq[2]};
        for (int c = 0; c < 3; c++) out[c] += wei * ((a[c] * d[0] + a[c + 3] * d[1] + a[c + 6] * d[2]) + gq[c] + a[9 + c]);
      }
    }
  };
  for (int b = 0; b < P.n_blocks; b++) {
    if (P.block_type[b] == -1) continue;
    if (P.on_center) {
      for (uint32_t node : P.center_sel[b]) {
        double nv[3]; skin(node, nv);
        for (int c = 0; c < 3; c++) f[index + c] += P.w_con * (nv[c] - (double)P.center_aim[b][c]);
      }
      index += 3;
    } else {
      for (int t = P.block_off[b]; t < P.block_off[b + 1]; t++) {
        uint32_t node = P.block_nodes[t];
        double nv[3]; skin(node, nv);
        for (int c = 0; c < 3; c++) f[index + c] = P.w_con * (nv[c] - (double)P.aim[3 * node + c]);
        index += 3;
      }
    }
  }
}

// FastCalcJacobiMat (DC:180-376): rows emitted in the reference's order.
template <typename F>
static void for_each_jrow(const SolveProblem& P, const double* x, F&& emit) {
  const int n = (int)P.free_nodes.size(), k = P.k;
  Row row;
  for (int i = 0; i < n; i++) {
    int k0 = 12 * i; const double* a = x + k0; double w = P.w_rot;
    const int pr[3][2] = {{0, 1}, {0, 2}, {1, 2}};
    for (int r = 0; r < 3; r++) {
      int c0 = pr[r][0], c1 = pr[r][1]; row.n = 0;
      for (int t = 0; t < 3; t++) { row.col[row.n] = k0 + 3 * c0 + t; row.val[row.n++] = a[3 * c1 + t] * w; }
      for (int t = 0; t < 3; t++) { row.col[row.n] = k0 + 3 * c1 + t; row.val[row.n++] = a[3 * c0 + t] * w; }
      emit(row);
    }
    for (int j = 0; j < 3; j++) {
      row.n = 0;
      for (int t = 0; t < 3; t++) { row.col[row.n] = k0 + 3 * j + t; row.val[row.n++] = 2 * a[3 * j + t] * w; }
      emit(row);
    }
  }
  for (int r = 0; r < n; r++) {
    int i = P.free_nodes[r]; int k1 = 12 * r;
    const float* vi = P.node_pos + 3 * i;
    for (int t = 0; t < k; t++) {
      int q = P.nbr[(size_t)i * k + t];
      const float* vk = P.node_pos + 3 * q;
      int rk = P.free_rank[q];
      double ws = same_block(P, i, q) ? P.weight_factor : 1.0;
      double w = P.w_reg;
      for (int j = 0; j < 3; j++) {
        row.n = 0;
        // float differences (DC:254-256), then double products
        row.col[row.n] = k1 + j;     row.val[row.n++] = (vk[0] - vi[0]) * w * ws;
        row.col[row.n] = k1 + j + 3; row.val[row.n++] = (vk[1] - vi[1]) * w * ws;
        row.col[row.n] = k1 + j + 6; row.val[row.n++] = (vk[2] - vi[2]) * w * ws;
        row.col[row.n] = k1 + j + 9; row.val[row.n++] = w;  // no weight_scale (DC:257)
        if (rk >= 0) { row.col[row.n] = 12 * rk + j + 9; row.val[row.n++] = -w * ws; }
        emit(row);
      }
    }
  }
  for (int i : P.static_nodes) {
    for (int t = 0; t < k; t++) {
      int q = P.nbr[(size_t)i * k + t]; int rk = P.free_rank[q];
      for (int j = 0; j < 3; j++) {
        row.n = 0;
        if (rk >= 0) { row.col[row.n] = 12 * rk + j + 9; row.val[row.n++] = -P.w_reg; }
        emit(row);
      }
    }
  }
  auto con_entries = [&](uint32_t node, int kk, Row& rw) {
    const float* vi = P.node_pos + 3 * node;
    for (int j = 0; j < k; j++) {
      uint32_t q = P.anchor_idx[(size_t)node * k + j];
      double wei = P.anchor_w[(size_t)node * k + j];
      int rk = P.free_rank[q]; if (rk < 0) continue;
      const float* vk = P.node_pos + 3 * q; int k1 = 12 * rk; double w = P.w_con;
      rw.col[rw.n] = k1 + kk;     rw.val[rw.n++] = wei * (vi[0] - vk[0]) * w;
      rw.col[rw.n] = k1 + 3 + kk; rw.val[rw.n++] = wei * (vi[1] - vk[1]) * w;
      rw.col[rw.n] = k1 + 6 + kk; rw.val[rw.n++] = wei * (vi[2] - vk[2]) * w;
      rw.col[rw.n] = k1 + 9 + kk; rw.val[rw.n++] = wei * w;
    }
  };
  for (int b = 0; b < P.n_blocks; b++) {
    if (P.block_type[b] == -1) continue;
    if (P.on_center) {
      // one 3-row group per block; entries of all selected nodes summed
      // (setFromTriplets sums duplicates, DC:374).  Rows can be long: emit per
      // node as separate partial rows is NOT equivalent for J^T J, so gather.
      for (int kk = 0; kk < 3; kk++) {
        std::unordered_map<int, double> acc;
        for (uint32_t node : P.center_sel[b]) {
          Row rw; rw.n = 0; con_entries(node, kk, rw);
          for (int t = 0; t < rw.n; t++) acc[rw.col[t]] += rw.val[t];
        }
        // emit as chunks is wrong; use the long-row path
        std::vector<int> cols; std::vector<double> vals;
        for (auto& kv : acc) { cols.push_back(kv.first); vals.push_back(kv.second); }
        // sort for determinism
        std::vector<int> ord(cols.size()); std::iota(ord.begin(), ord.end(), 0);
        std::sort(ord.begin(), ord.end(), [&](int a, int b2) { return cols[a] < cols[b2]; });
        Row dummy; dummy.n = -1;  // signal long row
        emit(dummy, cols, vals, ord);
      }
    } else {
      for (int t = P.block_off[b]; t < P.block_off[b + 1]; t++) {
        uint32_t node = P.block_nodes[t];
        for (int kk = 0; kk < 3; kk++) { row.n = 0; con_entries(node, kk, row); emit(row); }
      }
    }
  }
}

// ---- block-sparse SPD solver (12x12 blocks), minimum-degree ordering ------
// Restates Eigen::SimplicialCholesky (DC:103-133) as "an exact sparse SPD
// direct solve": the solution of (J^T J) h = g is unique to rounding.
struct BlockSPD {
  int n = 0;
  std::vector<std::unordered_map<int, int>> rowmap;  // rowmap[i][j] -> block id (i>=j stored; also j>i mirrored lazily)
  std::vector<std::array<double, 144>> blocks;       // block (i,j): row-major 12x12
  void init(int n_) { n = n_; rowmap.assign(n, {}); blocks.clear(); }
  double* at(int i, int j) {  // stores lower triangle only (i >= j)
    auto it = rowmap[i].find(j);
    if (it != rowmap[i].end()) return blocks[it->second].data();
    int id = (int)blocks.size(); blocks.emplace_back(); blocks.back().fill(0.0);
    rowmap[i][j] = id; return blocks[id].data();
  }
};

static void chol12(double* A) {  // in place lower Cholesky of row-major 12x12 (full symmetric input)
  for (int j = 0; j < 12; j++) {
    double s = A[j * 12 + j];
    for (int t = 0; t < j; t++) s -= A[j * 12 + t] * A[j * 12 + t];
    double d = std::sqrt(s); A[j * 12 + j] = d;
    for (int i = j + 1; i < 12; i++) {
      double v = A[i * 12 + j];
      for (int t = 0; t < j; t++) v -= A[i * 12 + t] * A[j * 12 + t];
      A[i * 12 + j] = v / d;
    }
    for (int t = j + 1; t < 12; t++) A[j * 12 + t] = 0.0;
  }
}

struct SparseChol {
  int n;
  std::vector<int> perm, pos;             // perm[p] = node at elimination position p
  std::vector<std::vector<int>> colrows;  // positions > p in column p (sorted)
  std::vector<std::vector<double>> L;     // per column: (1 + rows) blocks of 144 (diag first)

  void analyze(const BlockSPD& H) {
    n = H.n;
    std::vector<std::vector<int>> adj(n);
    for (int i = 0; i < n; i++) for (auto& kv : H.rowmap[i]) if (kv.first != i) { adj[i].push_back(kv.first); adj[kv.first].push_back(i); }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<char> done(n, 0); perm.resize(n); pos.assign(n, -1);
    std::vector<std::vector<int>> elim_nbrs(n);
    std::vector<int> mark(n, -1);
    for (int p = 0; p < n; p++) {
      int best = -1; size_t bd = SIZE_MAX;
      for (int v = 0; v < n; v++) if (!done[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
      int v = best; done[v] = 1; perm[p] = v; pos[v] = p;
      elim_nbrs[v] = adj[v];
      const std::vector<int>& N = elim_nbrs[v];
      for (int u : N) {
        // adj[u] = (adj[u] U N) \ {u, v}
        for (int t : adj[u]) mark[t] = u;
        mark[u] = u; mark[v] = u;
        std::vector<int>& au = adj[u];
        au.erase(std::remove(au.begin(), au.end(), v), au.end());
        for (int t : N) if (mark[t] != u) { au.push_back(t); mark[t] = u; }
      }
      adj[v].clear(); adj[v].shrink_to_fit();
    }
    colrows.assign(n, {});
    for (int p = 0; p < n; p++) {
      int v = perm[p];
      for (int u : elim_nbrs[v]) colrows[p].push_back(pos[u]);
      std::sort(colrows[p].begin(), colrows[p].end());
    }
  }

  double* blk(int rowp, int colp) {  // rowp >= colp
    if (rowp == colp) return L[colp].data();
    auto& r = colrows[colp];
    auto it = std::lower_bound(r.begin(), r.end(), rowp);
    return L[colp].data() + 144 * (1 + (it - r.begin()));
  }

  void factorize(const BlockSPD& H) {
    L.assign(n, {});
    for (int p = 0; p < n; p++) L[p].assign(144 * (1 + colrows[p].size()), 0.0);
    for (int i = 0; i < n; i++) for (auto& kv : H.rowmap[i]) {
      int j = kv.first; const double* src = H.blocks[kv.second].data();
      int pi = pos[i], pj = pos[j];
      if (pi >= pj) { double* d = blk(pi, pj); for (int t = 0; t < 144; t++) d[t] += src[t]; }
      else { double* d = blk(pj, pi); for (int a = 0; a < 12; a++) for (int b = 0; b < 12; b++) d[b * 12 + a] += src[a * 12 + b]; }
    }
    for (int p = 0; p < n; p++) {
      double* D = L[p].data();
      chol12(D);
      const int nr = (int)colrows[p].size();
      // L_ip = H_ip * D^-T  (solve X D^T = B, D lower)
#pragma omp parallel for schedule(static) if (nr > 16)
      for (int t = 0; t < nr; t++) {
        double* B = L[p].data() + 144 * (1 + t);
        for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) {
          double v = B[r * 12 + c];
          for (int s = 0; s < c; s++) v -= B[r * 12 + s] * D[c * 12 + s];
          B[r * 12 + c] = v / D[c * 12 + c];
        }
      }
      // trailing update: for a<=b in rows: blk(rb, ra) -= L_b L_a^T
#pragma omp parallel for schedule(dynamic, 1) if (nr > 8)
      for (int a = 0; a < nr; a++) {
        const double* La = L[p].data() + 144 * (1 + a);
        int ra = colrows[p][a];
        for (int b = a; b < nr; b++) {
          const double* Lb = L[p].data() + 144 * (1 + b);
          int rb = colrows[p][b];
          double* T = blk(rb, ra);
          for (int r = 0; r < 12; r++) for (int c = 0; c < 12; c++) {
            double s = 0; for (int t = 0; t < 12; t++) s += Lb[r * 12 + t] * La[c * 12 + t];
            T[r * 12 + c] -= s;
          }
        }
      }
    }
  }

  void solve(const double* g, double* h) const {
    std::vector<double> y((size_t)12 * n);
    for (int p = 0; p < n; p++) for (int c = 0; c < 12; c++) y[12 * p + c] = g[12 * perm[p] + c];
    for (int p = 0; p < n; p++) {  // forward
      const double* D = L[p].data(); double* yp = &y[12 * p];
      for (int r = 0; r < 12; r++) { double v = yp[r]; for (int s = 0; s < r; s++) v -= D[r * 12 + s] * yp[s]; yp[r] = v / D[r * 12 + r]; }
      for (size_t t = 0; t < colrows[p].size(); t++) {
        const double* B = L[p].data() + 144 * (1 + t); double* yr = &y[12 * colrows[p][t]];
        for (int r = 0; r < 12; r++) { double s = 0; for (int c = 0; c < 12; c++) s += B[r * 12 + c] * yp[c]; yr[r] -= s; }
      }
    }
    for (int p = n - 1; p >= 0; p--) {  // backward
      const double* D = L[p].data(); double* yp = &y[12 * p];
      for (size_t t = 0; t < colrows[p].size(); t++) {
        const double* B = L[p].data() + 144 * (1 + t); const double* yr = &y[12 * colrows[p][t]];
        for (int c = 0; c < 12; c++) { double s = 0; for (int r = 0; r < 12; r++) s += B[r * 12 + c] * yr[r]; yp[c] -= s; }
      }
      for (int r = 11; r >= 0; r--) { double v = yp[r]; for (int s = r + 1; s < 12; s++) v -= D[s * 12 + r] * yp[s]; yp[r] = v / D[r * 12 + r]; }
    }
    for (int p = 0; p < n; p++) for (int c = 0; c < 12; c++) h[12 * perm[p] + c] = y[12 * p + c];
  }
};

}  // namespace

// stats: [0]=GN iterations, [1]=final energy f.f, [2]=total line-search halvings,
//        [3]=last |h|, [4]=rows, [5]=unknowns
ORC_API int orc_solve(int M, int k, const float* node_pos, const uint32_t* nbr,
                      const uint32_t* anchor_idx, const double* anchor_w,
                      const unsigned char* node_static,
                      int n_blocks, const int* block_off, const uint32_t* block_nodes, const int* block_type,
                      const float* aim, int on_center, double w_rot, double w_reg, double w_con,
                      double weight_factor, int max_iters,
                      double* rot_out, double* trans_out, double* stats) {
  SolveProblem P;
  P.M = M; P.k = k; P.node_pos = node_pos; P.nbr = nbr; P.anchor_idx = anchor_idx; P.anchor_w = anchor_w;
  P.n_blocks = n_blocks; P.block_off = block_off; P.block_nodes = block_nodes; P.block_type = block_type;
  P.aim = aim; P.on_center = on_center != 0; P.weight_factor = weight_factor;
  P.w_rot = std::sqrt(w_rot); P.w_reg = std::sqrt(w_reg); P.w_con = std::sqrt(w_con);
  P.free_rank.assign(M, -1);
  for (int i = 0; i < M; i++) { if (node_static && node_static[i]) P.static_nodes.push_back(i); else { P.free_rank[i] = (int)P.free_nodes.size(); P.free_nodes.push_back(i); } }
  P.node_blocks.assign(M, {});
  int control_blocks = 0, control_nodes = 0;  // GetControlNums (DC:6-18)
  for (int b = 0; b < n_blocks; b++) {
    if (block_type[b] == -1) continue;
    control_blocks++; control_nodes += block_off[b + 1] - block_off[b];
    for (int t = block_off[b]; t < block_off[b + 1]; t++) P.node_blocks[block_nodes[t]].push_back(b);
  }
  // SelectKeyControls (DC:34-75): FPS picks min(20,|block|) nodes but the code
  // then takes the FIRST that-many nodes of the block (c_indices[i], DC:59-60);
  // std::set de-duplicates and orders them.
  P.center_sel.assign(n_blocks, {}); P.center_aim.assign(n_blocks, {0, 0, 0});
  if (P.on_center) {
    for (int b = 0; b < n_blocks; b++) {
      if (block_type[b] == -1) continue;
      int sz = block_off[b + 1] - block_off[b];
      int nsel = std::min(20, sz);  // CONTROL_NODE_NUM (DC:4)
      float c[3] = {0, 0, 0};
      std::vector<uint32_t> sel;
      for (int t = 0; t < nsel; t++) {
        uint32_t nd = block_nodes[block_off[b] + t];
        sel.push_back(nd);
        for (int d = 0; d < 3; d++) c[d] += aim[3 * nd + d];
      }
      std::sort(sel.begin(), sel.end()); sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
      P.center_sel[b] = sel;
      for (int d = 0; d < 3; d++) P.center_aim[b][d] = c[d] / (float)nsel;  // Pos / size_t -> float divide
    }
  }
  const int n = (int)P.free_nodes.size();
  P.rows = n * (6 + 3 * k) + (P.on_center ? 3 * control_blocks : 3 * control_nodes) + (int)P.static_nodes.size() * 3 * k;
  const int nx = 12 * n;
  std::vector<double> x(nx, 0.0);  // setIdentityRots (DC:83-93)
  for (int i = 0; i < n; i++) { x[12 * i] = 1.0; x[12 * i + 4] = 1.0; x[12 * i + 8] = 1.0; }

  std::vector<double> f, f1, g(nx), h(nx), xt(nx);
  BlockSPD H; SparseChol chol; bool analyzed = false;
  int iters = 0, halvings = 0; double normh = 0.0, energy = 0.0;
  for (int iter = 0; iter < max_iters; iter++) {
    iters = iter + 1;
    eval_f(P, x.data(), f);
    H.init(n); std::fill(g.begin(), g.end(), 0.0);
    int ridx = 0;
    struct Emit {
      BlockSPD& H; std::vector<double>& g; const std::vector<double>& f; int& ridx;
      void operator()(const Row& r) {
        double fr = f[ridx++];
        for (int a = 0; a < r.n; a++) {
          g[r.col[a]] -= r.val[a] * fr;
          for (int b = 0; b < r.n; b++) {
            int ba = r.col[a] / 12, bb = r.col[b] / 12;
            if (ba >= bb) H.at(ba, bb)[(r.col[a] % 12) * 12 + (r.col[b] % 12)] += r.val[a] * r.val[b];
          }
        }
      }
      void operator()(const Row&, const std::vector<int>& cols, const std::vector<double>& vals, const std::vector<int>& ord) {
        double fr = f[ridx++];
        for (int ia : ord) {
          g[cols[ia]] -= vals[ia] * fr;
          for (int ib : ord) {
            int ba = cols[ia] / 12, bb = cols[ib] / 12;
            if (ba >= bb) H.at(ba, bb)[(cols[ia] % 12) * 12 + (cols[ib] % 12)] += vals[ia] * vals[ib];
          }
        }
      }
    } emit{H, g, f, ridx};
    for_each_jrow(P, x.data(), emit);
    if (ridx != P.rows) { fprintf(stderr, "orc_solve: row count mismatch %d vs %d\n", ridx, P.rows); return -1; }
    // make sure every free node has a diagonal block even if isolated
    for (int i = 0; i < n; i++) H.at(i, i);
    if (!analyzed) { chol.analyze(H); analyzed = true; }
    chol.factorize(H);
    chol.solve(g.data(), h.data());

    double normv = 0; for (double v : x) normv += v * v; normv = std::sqrt(normv);
    double old_e = 0; for (double v : f) old_e += v * v;
    for (double alpha = 1.0; alpha > 1e-15; alpha *= 0.5) {  // DC:144-156
      for (int t = 0; t < nx; t++) xt[t] = x[t] + h[t];
      eval_f(P, xt.data(), f1);
      double new_e = 0; for (double v : f1) new_e += v * v;
      if (new_e > old_e) { for (double& v : h) v *= 0.5; halvings++; }
      else { x = xt; break; }
    }
    normh = 0; for (double v : h) normh += v * v; normh = std::sqrt(normh);
    energy = old_e;
    if (normh < (normv + 1e-6) * 1e-6) break;
  }
  // putFreeInputs (DH:140-151)
  for (int i = 0; i < M; i++) {
    for (int t = 0; t < 9; t++) rot_out[9 * i + t] = (t % 4 == 0) ? 1.0 : 0.0;
    for (int t = 0; t < 3; t++) trans_out[3 * i + t] = 0.0;
  }
  for (int r = 0; r < n; r++) {
    int i = P.free_nodes[r];
    for (int t = 0; t < 9; t++) rot_out[9 * i + t] = x[12 * r + t];
    for (int t = 0; t < 3; t++) trans_out[3 * i + t] = x[12 * r + 9 + t];
  }
  if (stats) {
    // `return fx.dot(fx)` is the energy at the last linearisation point (DC:168)
    stats[0] = iters; stats[1] = energy; stats[2] = halvings; stats[3] = normh; stats[4] = P.rows; stats[5] = nx;
  }
  return 0;
}

// Jacobian/residual export at a given x (per node M x 12: rot 9 col-major + trans 3),
// for tests: COO triplets in the reference's row order; columns index 12*rank.
// Returns nnz (or -needed if cap too small).
ORC_API long long orc_jacobian(int M, int k, const float* node_pos, const uint32_t* nbr,
                               const uint32_t* anchor_idx, const double* anchor_w, const unsigned char* node_static,
                               int n_blocks, const int* block_off, const uint32_t* block_nodes, const int* block_type,
                               const float* aim, int on_center, double w_rot, double w_reg, double w_con,
                               const double* rot, const double* trans,
                               long long cap, int* rows_out, int* cols_out, double* vals_out, double* f_out, int* dims_out);

// residual only (for tests comparing f at a given x laid out per node M x 12)
ORC_API double orc_energy(int M, int k, const float* node_pos, const uint32_t* nbr,
                          const uint32_t* anchor_idx, const double* anchor_w, const unsigned char* node_static,
                          int n_blocks, const int* block_off, const uint32_t* block_nodes, const int* block_type,
                          const float* aim, int on_center, double w_rot, double w_reg, double w_con,
                          const double* rot, const double* trans) {
  SolveProblem P;
  P.M = M; P.k = k; P.node_pos = node_pos; P.nbr = nbr; P.anchor_idx = anchor_idx; P.anchor_w = anchor_w;
  P.n_blocks = n_blocks; P.block_off = block_off; P.block_nodes = block_nodes; P.block_type = block_type;
  P.aim = aim; P.on_center = on_center != 0; P.weight_factor = 1.0;
  P.w_rot = std::sqrt(w_rot); P.w_reg = std::sqrt(w_reg); P.w_con = std::sqrt(w_con);
  P.free_rank.assign(M, -1);
  for (int i = 0; i < M; i++) { if (node_static && node_static[i]) P.static_nodes.push_back(i); else { P.free_rank[i] = (int)P.free_nodes.size(); P.free_nodes.push_back(i); } }
  P.node_blocks.assign(M, {});
  int cb = 0, cn = 0;
  for (int b = 0; b < n_blocks; b++) if (block_type[b] != -1) { cb++; cn += block_off[b + 1] - block_off[b]; }
  P.center_sel.assign(n_blocks, {}); P.center_aim.assign(n_blocks, {0, 0, 0});
  if (P.on_center) for (int b = 0; b < n_blocks; b++) {
    if (block_type[b] == -1) continue;
    int sz = block_off[b + 1] - block_off[b], nsel = std::min(20, sz); float c[3] = {0, 0, 0};
    std::vector<uint32_t> sel;
    for (int t = 0; t < nsel; t++) { uint32_t nd = block_nodes[block_off[b] + t]; sel.push_back(nd); for (int d = 0; d < 3; d++) c[d] += aim[3 * nd + d]; }
    std::sort(sel.begin(), sel.end()); sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
    P.center_sel[b] = sel; for (int d = 0; d < 3; d++) P.center_aim[b][d] = c[d] / (float)nsel;
  }
  int n = (int)P.free_nodes.size();
  P.rows = n * (6 + 3 * k) + (P.on_center ? 3 * cb : 3 * cn) + (int)P.static_nodes.size() * 3 * k;
  std::vector<double> x(12 * (size_t)n);
  for (int r = 0; r < n; r++) { int i = P.free_nodes[r]; for (int t = 0; t < 9; t++) x[12 * r + t] = rot[9 * i + t]; for (int t = 0; t < 3; t++) x[12 * r + 9 + t] = trans[3 * i + t]; }
  std::vector<double> f; eval_f(P, x.data(), f);
  double e = 0; for (double v : f) e += v * v; return e;
}

// ===========================================================================
// Density grid (GV:3601-4149, CK:403-451)
// ===========================================================================

// getOverallAABB (GV:3601-3631).  Eigen float vectors times double literals
// evaluate in float (the literal is converted to the expression's scalar).
ORC_API void orc_overall_aabb(const float* pos, long long N, float* out6) {
  float mn[3] = {pos[0], pos[1], pos[2]}, mx[3] = {pos[0], pos[1], pos[2]};
  for (long long i = 0; i < N; i++) for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], pos[3 * i + c]); mx[c] = std::max(mx[c], pos[3 * i + c]); }
  const float f11 = (float)1.1, fgrow = (float)(1.0 + (1.0) / 128);  // NUM_SAMPLES_PER_DIM 128
  for (int c = 0; c < 3; c++) {
    float mean = (mx[c] + mn[c]) / 2.0f;
    float nmin = (mn[c] - mean) * f11 + mean;
    float nmax = ((mx[c] - mean) * f11 + mean - nmin) * fgrow + nmin;
    out6[c] = std::min(nmin, -0.75f);
    out6[3 + c] = std::max(nmax, 0.75f);
  }
}

// grid_step (GV:3873-3877)
ORC_API float orc_grid_step(const float* aabb6, int G) {
  float xs = (aabb6[3] - aabb6[0]) / G, ys = (aabb6[4] - aabb6[1]) / G, zs = (aabb6[5] - aabb6[2]) / G;
  return std::max(std::max(xs, ys), zs);
}

static inline int cell_of(float p, float mn, float step) { return int(std::floor((p - mn) / step)); }

// GetGsGrid + inclusive scan + host re-order indices (CK:403-423, GV:3896-3934)
// cell_out[N], prefix_out[G^3] inclusive, new_idx[N] (destination slot of i).
ORC_API void orc_cell_assign(const float* pos, long long N, const float* min3, float step, int G,
                             int* cell_out, int* prefix_out, int* new_idx) {
  long long GC = (long long)G * G * G;
  std::vector<int> cnt(GC, 0);
  for (long long i = 0; i < N; i++) {
    int c = cell_of(pos[3 * i], min3[0], step) * G * G + cell_of(pos[3 * i + 1], min3[1], step) * G + cell_of(pos[3 * i + 2], min3[2], step);
    cell_out[i] = c; cnt[c]++;
  }
  int run = 0; for (long long c = 0; c < GC; c++) { run += cnt[c]; prefix_out[c] = run; }
  if (new_idx) {
    std::vector<int> used(GC, 0);
    for (long long i = 0; i < N; i++) { int c = cell_out[i]; new_idx[i] = (c == 0 ? 0 : prefix_out[c - 1]) + used[c]++; }
  }
}

// per-Gaussian cutoff AABB (GV:3961-4019): aabb N x 6 (min xyz, max xyz),
// clip N x 3 (scale_3d_clip), smax N (scale_3d_max)
ORC_API void orc_gs_aabbs(const float* pos, const float* rot, const float* scale, const float* opacity, long long N,
                          float* aabb, float* clip, float* smax) {
#pragma omp parallel for schedule(static)
  for (long long g = 0; g < N; g++) {
    Qf q = quat_normalized(Qf{rot[4 * g], rot[4 * g + 1], rot[4 * g + 2], rot[4 * g + 3]});
    float R[3][3]; quat_to_matrix(q, R);
    float s3[3];
    if (opacity[g] <= 1.0f / 255.0f) { s3[0] = s3[1] = s3[2] = 0.0f; }
    else {
      for (int i = 0; i < 3; i++) s3[i] = std::sqrt(-2.0f * std::log(1.0f / 255.0f / opacity[g])) * (scale[3 * g + i] + 0.0f);
    }
    if (clip) for (int i = 0; i < 3; i++) clip[3 * g + i] = s3[i];
    if (smax) { float m = s3[0]; if (s3[1] > m) m = s3[1]; if (s3[2] > m) m = s3[2]; smax[g] = m; }
    float mn[3] = {pos[3 * g], pos[3 * g + 1], pos[3 * g + 2]}, mx[3] = {mn[0], mn[1], mn[2]};
    for (int d = 0; d < 3; d++) for (int c = 0; c < 3; c++) {
      float v = R[c][d] * s3[d];
      float l = pos[3 * g + c] + v, r = pos[3 * g + c] + (-v);
      mn[c] = std::min(mn[c], l); mx[c] = std::max(mx[c], l);
      mn[c] = std::min(mn[c], r); mx[c] = std::max(mx[c], r);
    }
    for (int c = 0; c < 3; c++) { aabb[6 * g + c] = mn[c]; aabb[6 * g + 3 + c] = mx[c]; }
  }
}

static inline void cell_range(const float* a, const float* min3, float step, int G, int padding, int lo[3], int hi[3]) {
  for (int c = 0; c < 3; c++) {
    lo[c] = std::max(int(std::floor((a[c] - min3[c]) / step) - padding), 0);
    hi[c] = std::min(int(std::floor((a[3 + c] - min3[c]) / step) + padding), G - 1);
  }
}

// GetBoxesGsGrid + scan (CK:425-451, 533-553): prefix[G^3] inclusive; returns P
ORC_API long long orc_footprint_count(const float* aabb, long long N, const float* min3, float step, int G, int padding, int* prefix_out) {
  long long GC = (long long)G * G * G;
  std::vector<int> cnt(GC, 0);
  for (long long g = 0; g < N; g++) {
    int lo[3], hi[3]; cell_range(aabb + 6 * g, min3, step, G, padding, lo, hi);
    for (int x = lo[0]; x <= hi[0]; x++) for (int y = lo[1]; y <= hi[1]; y++) for (int z = lo[2]; z <= hi[2]; z++) cnt[(long long)x * G * G + y * G + z]++;
  }
  int run = 0; for (long long c = 0; c < GC; c++) { run += cnt[c]; prefix_out[c] = run; }
  return run;
}

// serial host fill (GV:4077-4100): lists ascending in Gaussian index per cell
ORC_API void orc_footprint_fill(const float* aabb, long long N, const float* min3, float step, int G, int padding,
                                const int* prefix, int* lists) {
  long long GC = (long long)G * G * G;
  std::vector<int> used(GC, 0);
  for (long long g = 0; g < N; g++) {
    int lo[3], hi[3]; cell_range(aabb + 6 * g, min3, step, G, padding, lo, hi);
    for (int x = lo[0]; x <= hi[0]; x++) for (int y = lo[1]; y <= hi[1]; y++) for (int z = lo[2]; z <= hi[2]; z++) {
      long long c = (long long)x * G * G + y * G + z;
      lists[(c == 0 ? 0 : prefix[c - 1]) + used[c]++] = (int)g;
    }
  }
}

// valid cells (GV:4040-4053): cells with a non-empty list; returns V
ORC_API int orc_valid_cells(const int* prefix, int G, int* valid_out) {
  long long GC = (long long)G * G * G; int v = 0;
  for (long long c = 0; c < GC; c++) { int n = prefix[c] - (c ? prefix[c - 1] : 0); if (n != 0) { if (valid_out) valid_out[v] = (int)c; v++; } }
  return v;
}

// sample emit (GV:4111-4133): V x 64 x 3
ORC_API void orc_emit_samples(const int* valid, int V, const float* min3, float step, int G, float* out) {
  const int spd = 4;  // SAMPLES_PER_GRID
  float interval = step / spd;
  for (int i = 0; i < V; i++) {
    int c = valid[i];
    int xi = c / (G * G), yi = (c - xi * G * G) / G, zi = c % G;
    // float + int*float (float) + 0.5*interval (double) -> rounded to float
    float gx = (float)((double)(min3[0] + xi * step) + 0.5 * interval);
    float gy = (float)((double)(min3[1] + yi * step) + 0.5 * interval);
    float gz = (float)((double)(min3[2] + zi * step) + 0.5 * interval);
    for (int s = 0; s < 64; s++) {
      int sx = s / 16, sy = (s - sx * 16) / 4, sz = s % 4;
      float* o = out + ((size_t)i * 64 + s) * 3;
      o[0] = gx + sx * interval; o[1] = gy + sy * interval; o[2] = gz + sz * interval;
    }
  }
}

// GetAdaLpfRatio (GV:4670-4751): per valid cell, least-squares 3x3 M mapping
// the unit-cube corners to the 8 corner samples; lpf = M M^T * 0.2.
// The 24x9 normal equations decouple per output row: M = P Q^T (Q Q^T)^-1 with
// Q Q^T = 2 I + 2 (ones) -> inverse = 0.5 I - 0.125 ones.  out: G^3 x 9
// (row-major; the matrix is symmetric), cells not in `valid` left untouched.
ORC_API void orc_ada_lpf(const float* samples, const int* valid, int V, float lpf_parameter, float* out) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < V; i++) {
    const float* base = samples + (size_t)i * 64 * 3;
    float PQt[3][3] = {{0}};  // sum over corners of P(:,c) q(c)^T
    float Psum[3] = {0, 0, 0};
    for (int col = 0; col < 8; col++) {
      int idx = 0; float q[3] = {0, 0, 0};
      if (col % 2 == 1) { q[0] = 1.0f; idx += 48; }
      if ((col / 2) % 2 == 1) { q[1] = 1.0f; idx += 12; }
      if ((col / 4) % 2 == 1) { q[2] = 1.0f; idx += 3; }
      float p[3];
      for (int r = 0; r < 3; r++) p[r] = (base[idx * 3 + r] - base[r]) / 3.0f;
      for (int r = 0; r < 3; r++) { Psum[r] += p[r]; for (int c = 0; c < 3; c++) PQt[r][c] += p[r] * q[c]; }
    }
    float Mx[3][3];
    for (int r = 0; r < 3; r++) {
      float rs = PQt[r][0] + PQt[r][1] + PQt[r][2];
      for (int c = 0; c < 3; c++) Mx[r][c] = 0.5f * PQt[r][c] - 0.125f * rs;
    }
    (void)Psum;
    float* o = out + (size_t)valid[i] * 9;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
      float s = 0; for (int t = 0; t < 3; t++) s += Mx[r][t] * Mx[c][t];
      o[3 * r + c] = s * lpf_parameter;
    }
  }
}

// forward3d_grid — EXTERNAL (XinhaoT/CudaRasterizer @ 96ea96c, source absent;
// call site GV:4159-4186).  PARITY UNPINNED.  Restated from the parameter list
// and the paper's field definition:
//   w_g(x)    = alpha_g * exp(-1/2 d^T (Sigma_g + LPF_cell)^-1 d),  d = x - mu_g
//   feature(x)= sum_g w_g(x) * SH_g      opacity(x) = sum_g w_g(x)
// over the Gaussians of the sample's cell list, in list order (ascending index),
// float accumulation.  Contributions with exponent power < ln(1/255 / alpha)
// (outside the footprint cutoff that defines the lists, GV:3978) are dropped.
ORC_API void orc_grid_eval(const int* valid, int V, const int* prefix, const int* lists, const float* samples,
                           const float* pos, const float* rot, const float* scale, const float* opacity, const float* shs,
                           const float* ada_lpf, float* out_feature, float* out_opacity) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < V; i++) {
    int c = valid[i];
    int beg = c ? prefix[c - 1] : 0, end = prefix[c];
    const float* lpf = ada_lpf + (size_t)c * 9;
    std::vector<float> feat(64 * 48, 0.0f), opa(64, 0.0f);
    for (int t = beg; t < end; t++) {
      int g = lists[t];
      float a = opacity[g];
      if (a <= 1.0f / 255.0f) continue;
      Qf q = quat_normalized(Qf{rot[4 * g], rot[4 * g + 1], rot[4 * g + 2], rot[4 * g + 3]});
      float R[3][3]; quat_to_matrix(q, R);
      float Sg[3][3];
      for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) {
        float s = 0; for (int d = 0; d < 3; d++) s += R[r][d] * (scale[3 * g + d] * scale[3 * g + d]) * R[cc][d];
        Sg[r][cc] = s + lpf[3 * r + cc];
      }
      // symmetric inverse
      float c00 = Sg[1][1] * Sg[2][2] - Sg[1][2] * Sg[2][1];
      float c01 = Sg[1][2] * Sg[2][0] - Sg[1][0] * Sg[2][2];
      float c02 = Sg[1][0] * Sg[2][1] - Sg[1][1] * Sg[2][0];
      float det = Sg[0][0] * c00 + Sg[0][1] * c01 + Sg[0][2] * c02;
      if (!(det > 0.0f)) continue;
      float id = 1.0f / det;
      float i00 = c00 * id, i01 = c01 * id, i02 = c02 * id;
      float i11 = (Sg[0][0] * Sg[2][2] - Sg[0][2] * Sg[2][0]) * id;
      float i12 = (Sg[0][2] * Sg[1][0] - Sg[0][0] * Sg[1][2]) * id;
      float i22 = (Sg[0][0] * Sg[1][1] - Sg[0][1] * Sg[1][0]) * id;
      float cut = std::log(1.0f / 255.0f / a);  // negative
      for (int s = 0; s < 64; s++) {
        const float* x = samples + ((size_t)i * 64 + s) * 3;
        float d0 = x[0] - pos[3 * g], d1 = x[1] - pos[3 * g + 1], d2 = x[2] - pos[3 * g + 2];
        float pw = -0.5f * (i00 * d0 * d0 + i11 * d1 * d1 + i22 * d2 * d2) - (i01 * d0 * d1 + i02 * d0 * d2 + i12 * d1 * d2);
        if (pw > 0.0f || pw < cut) continue;
        float w = a * std::exp(pw);
        opa[s] += w;
        float* f = &feat[s * 48];
        const float* sh = shs + (size_t)g * 48;
        for (int u = 0; u < 48; u++) f[u] += w * sh[u];
      }
    }
    memcpy(out_feature + (size_t)i * 64 * 48, feat.data(), sizeof(float) * 64 * 48);
    memcpy(out_opacity + (size_t)i * 64, opa.data(), sizeof(float) * 64);
  }
}

// JudgeEmptyGrid (GV:4272-4318)
ORC_API void orc_judge_empty(const int* valid, int V, int G, const float* aim_opacity, int* empty_out) {
  long long GC = (long long)G * G * G;
  std::vector<char> bad(GC, 1);
  for (int i = 0; i < V; i++) {
    float tot = 0.0f; for (int j = 0; j < 64; j++) tot += aim_opacity[(size_t)i * 64 + j];
    if (tot > 1e-6) {
      int idx = valid[i]; int z = idx % G, y = (idx / G) % G, x = idx / (G * G);
      auto at = [&](int xx, int yy, int zz) { bad[(long long)xx * G * G + yy * G + zz] = 0; };
      at(x, y, z); at(std::max(x - 1, 0), y, z); at(std::min(x + 1, G - 1), y, z);
      at(x, std::max(y - 1, 0), z); at(x, std::min(y + 1, G - 1), z);
      at(x, y, std::max(z - 1, 0)); at(x, y, std::min(z + 1, G - 1));
    }
  }
  for (int i = 0; i < V; i++) empty_out[i] = bad[valid[i]] ? 1 : 0;
}

// PointRotateByAxis (HC:1077-1100)
ORC_API void orc_rotate_by_axis(const float* point, const float* center, const float* axis4, float radian, float* out) {
  float cost = std::cos(radian), sint = std::sin(radian);
  float norm = std::sqrt(axis4[0] * axis4[0] + axis4[1] * axis4[1] + axis4[2] * axis4[2]);
  float x = axis4[0] / norm, y = axis4[1] / norm, z = axis4[2] / norm;
  out[0] = (x * x * (1 - cost) + cost) * point[0] + (x * y * (1 - cost) - z * sint) * point[1] + (x * z * (1 - cost) + y * sint) * point[2];
  out[1] = (y * x * (1 - cost) + z * sint) * point[0] + (y * y * (1 - cost) + cost) * point[1] + (y * z * (1 - cost) - x * sint) * point[2];
  out[2] = (z * x * (1 - cost) - y * sint) * point[0] + (z * y * (1 - cost) + x * sint) * point[1] + (z * z * (1 - cost) + cost) * point[2];
  float a = center[0], b = center[1], c = center[2];
  out[0] += (a * (y * y + z * z) - x * (b * y + c * z)) * (1 - cost) + (b * z - c * y) * sint;
  out[1] += (b * (x * x + z * z) - y * (a * x + c * z)) * (1 - cost) + (c * x - a * z) * sint;
  out[2] += (c * (x * x + y * y) - z * (a * x + b * y)) * (1 - cost) + (a * y - b * x) * sint;
}

ORC_API long long orc_jacobian(int M, int k, const float* node_pos, const uint32_t* nbr,
                               const uint32_t* anchor_idx, const double* anchor_w, const unsigned char* node_static,
                               int n_blocks, const int* block_off, const uint32_t* block_nodes, const int* block_type,
                               const float* aim, int on_center, double w_rot, double w_reg, double w_con,
                               const double* rot, const double* trans,
                               long long cap, int* rows_out, int* cols_out, double* vals_out, double* f_out, int* dims_out) {
  SolveProblem P;
  P.M = M; P.k = k; P.node_pos = node_pos; P.nbr = nbr; P.anchor_idx = anchor_idx; P.anchor_w = anchor_w;
  P.n_blocks = n_blocks; P.block_off = block_off; P.block_nodes = block_nodes; P.block_type = block_type;
  P.aim = aim; P.on_center = on_center != 0; P.weight_factor = 1.0;
  P.w_rot = std::sqrt(w_rot); P.w_reg = std::sqrt(w_reg); P.w_con = std::sqrt(w_con);
  P.free_rank.assign(M, -1);
  for (int i = 0; i < M; i++) { if (node_static && node_static[i]) P.static_nodes.push_back(i); else { P.free_rank[i] = (int)P.free_nodes.size(); P.free_nodes.push_back(i); } }
  P.node_blocks.assign(M, {});
  int cb = 0, cn = 0;
  for (int b = 0; b < n_blocks; b++) if (block_type[b] != -1) { cb++; cn += block_off[b + 1] - block_off[b]; }
  P.center_sel.assign(n_blocks, {}); P.center_aim.assign(n_blocks, {0, 0, 0});
  if (P.on_center) for (int b = 0; b < n_blocks; b++) {
    if (block_type[b] == -1) continue;
    int sz = block_off[b + 1] - block_off[b], nsel = std::min(20, sz); float c[3] = {0, 0, 0};
    std::vector<uint32_t> sel;
    for (int t = 0; t < nsel; t++) { uint32_t nd = block_nodes[block_off[b] + t]; sel.push_back(nd); for (int d = 0; d < 3; d++) c[d] += aim[3 * nd + d]; }
    std::sort(sel.begin(), sel.end()); sel.erase(std::unique(sel.begin(), sel.end()), sel.end());
    P.center_sel[b] = sel; for (int d = 0; d < 3; d++) P.center_aim[b][d] = c[d] / (float)nsel;
  }
  int n = (int)P.free_nodes.size();
  P.rows = n * (6 + 3 * k) + (P.on_center ? 3 * cb : 3 * cn) + (int)P.static_nodes.size() * 3 * k;
  std::vector<double> x(12 * (size_t)n);
  for (int r = 0; r < n; r++) { int i = P.free_nodes[r]; for (int t = 0; t < 9; t++) x[12 * r + t] = rot[9 * i + t]; for (int t = 0; t < 3; t++) x[12 * r + 9 + t] = trans[3 * i + t]; }
  std::vector<double> f; eval_f(P, x.data(), f);
  if (f_out) for (int r = 0; r < P.rows; r++) f_out[r] = f[r];
  if (dims_out) { dims_out[0] = P.rows; dims_out[1] = 12 * n; }
  long long nnz = 0; int ridx = 0;
  struct Emit {
    long long& nnz; int& ridx; long long cap; int* R; int* Cc; double* V;
    void operator()(const Row& r) { for (int a = 0; a < r.n; a++) { if (nnz < cap) { R[nnz] = ridx; Cc[nnz] = r.col[a]; V[nnz] = r.val[a]; } nnz++; } ridx++; }
    void operator()(const Row&, const std::vector<int>& cols, const std::vector<double>& vals, const std::vector<int>& ord) {
      for (int ia : ord) { if (nnz < cap) { R[nnz] = ridx; Cc[nnz] = cols[ia]; V[nnz] = vals[ia]; } nnz++; } ridx++; }
  } emit{nnz, ridx, cap, rows_out, cols_out, vals_out};
  for_each_jrow(P, x.data(), emit);
  return nnz <= cap ? nnz : -nnz;
}
