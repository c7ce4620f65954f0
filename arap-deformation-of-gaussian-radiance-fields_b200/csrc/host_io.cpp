// Host-only side of the reference's deformation API: deform.txt record /
// replay, graph.obj, <ply>_config.txt, the built-in deform scripts, and the
// replay state machine.  File formats are byte-compatible with the reference.
//
//   RecordDeformation / LoadDeformation   GaussianView.cpp:4859-4935, 4961-5095
//   replay state machine                  GaussianView.cpp:1757-1916
//   LoadDeformScript0/1, RunDeformScript  GaussianView.cpp:2512-2600, 2790-2816, 1918-1993
//   LoadMeshPoints / writeVectorToObj     helper.cpp:264-282, 1111-1126
//   config                                GaussianView.cpp:459-487, helper.cpp:198-230
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/arapgs.h"
#include "session.h"

namespace arapgs { void set_error(const std::string& msg); }
using arapgs::set_error;

// DeformHistory (helper.hpp:172-198)
struct arap_history {
  int nodes_on_mesh = 0;
  std::vector<int> nodes;            // anchor (Gaussian or mesh point) index per node
  int total_operations = 0, move_operations = 0;
  std::vector<int> operation_types;  // 0 add block, 1 drag (centre), 2 twist, 3 scale, 4 drag (per node), <0 delete block -(i+1)
  std::vector<std::vector<uint32_t>> block_nodes;
  std::vector<std::vector<float>> mouse_movements;  // n x 3 each
  std::vector<std::vector<int>> blocks_types_moves;
  std::vector<int> energy_on_centers;
  std::vector<float> twist_axis;     // 4 per move
};

// Counts come from an untrusted file: every count is bounded by the file size (each element takes at least two bytes of
// text) before anything is resized, and the cross-section consistency the replay relies on is checked here, so that a
// malformed file ends in ARAP_ERR_IO instead of an out-of-bounds read or a std::bad_alloc crossing the C boundary.
static int history_load_impl(const char* path, arap_history** out) {
  std::ifstream in(path);
  if (!in.is_open()) { set_error(std::string("history_load: cannot open ") + path); return ARAP_ERR_IO; }
  in.seekg(0, std::ios::end);
  const long long file_bytes = (long long)in.tellg();
  in.seekg(0, std::ios::beg);
  const long long max_count = file_bytes / 2 + 1;
  std::unique_ptr<arap_history> hp(new arap_history());
  arap_history* h = hp.get();
  std::string name; long long size = 0, inner = 0;
  auto fail = [&](const char* what) { set_error(std::string("history_load: malformed file at ") + what); return ARAP_ERR_IO; };
  auto count_ok = [&](long long c, long long per) { return c >= 0 && c <= max_count / per; };
  if (!(in >> h->nodes_on_mesh)) return fail("node type");
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Nodes:");
  h->nodes.resize((size_t)size);
  for (long long i = 0; i < size; i++) if (!(in >> h->nodes[i])) return fail("node indices");
  if (!(in >> name >> h->total_operations)) return fail("Total_Operations:");
  if (!(in >> name >> h->move_operations)) return fail("Move_Operations:");
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Operation_Types:");
  h->operation_types.resize((size_t)size);
  for (long long i = 0; i < size; i++) if (!(in >> h->operation_types[i])) return fail("operation types");
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Block_Nodes:");
  h->block_nodes.resize((size_t)size);
  for (long long i = 0; i < size; i++) {
    if (!(in >> inner) || !count_ok(inner, 1)) return fail("block size");
    h->block_nodes[i].resize((size_t)inner);
    for (long long j = 0; j < inner; j++) if (!(in >> h->block_nodes[i][j])) return fail("block nodes");
  }
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Mouse_Movements:");
  h->mouse_movements.resize((size_t)size);
  for (long long i = 0; i < size; i++) {
    if (!(in >> inner) || !count_ok(inner, 3)) return fail("movement count");
    h->mouse_movements[i].resize((size_t)inner * 3);
    for (long long j = 0; j < inner * 3; j++) if (!(in >> h->mouse_movements[i][j])) return fail("movements");
  }
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Blocks_Types_Moves:");
  h->blocks_types_moves.resize((size_t)size);
  for (long long i = 0; i < size; i++) {
    if (!(in >> inner) || !count_ok(inner, 1)) return fail("block types count");
    h->blocks_types_moves[i].resize((size_t)inner);
    for (long long j = 0; j < inner; j++) if (!(in >> h->blocks_types_moves[i][j])) return fail("block types");
  }
  if (!(in >> name >> size) || !count_ok(size, 1)) return fail("Energy_on_Center:");
  h->energy_on_centers.resize((size_t)size);
  for (long long i = 0; i < size; i++) if (!(in >> h->energy_on_centers[i])) return fail("energy flags");
  if (!(in >> name >> size) || !count_ok(size, 4)) return fail("Twist_Axis:");
  h->twist_axis.resize((size_t)size * 4);
  for (long long i = 0; i < size * 4; i++) if (!(in >> h->twist_axis[i])) return fail("twist axes");
  // RecordDeformation writes one entry of every per-move section per move operation (GV:4893-4932), and the replay
  // indexes all of them with the same move counter (GV:1815-1899).
  const size_t moves = h->mouse_movements.size();
  if (h->blocks_types_moves.size() != moves || h->energy_on_centers.size() != moves || h->twist_axis.size() != 4 * moves)
    return fail("per-move sections: Mouse_Movements / Blocks_Types_Moves / Energy_on_Center / Twist_Axis counts differ");
  size_t n_add = 0, n_move = 0;
  for (size_t i = 0; i < h->operation_types.size() && (long long)i < (long long)h->total_operations; i++) {
    const int op = h->operation_types[i];
    if (op > 4) return fail("operation types: unknown op");
    n_add += op == 0; n_move += op > 0;
  }
  if (h->total_operations < 0 || (size_t)h->total_operations > h->operation_types.size()) return fail("Total_Operations exceeds the Operation_Types count");
  if (n_add > h->block_nodes.size() || n_move > moves) return fail("operation types: more add / move ops than recorded blocks / movements");
  *out = hp.release();
  return ARAP_OK;
}

extern "C" int arap_history_load(const char* path, arap_history** out) {
  if (!path || !out) { set_error("history_load: bad arguments"); return ARAP_ERR_INVALID; }
  try { return history_load_impl(path, out); }
  catch (const std::exception& e) { set_error(std::string("history_load: ") + e.what()); return ARAP_ERR_IO; }
}

// Same token stream as RecordDeformation: default ostream float formatting, one space after every token.
extern "C" int arap_history_save(const arap_history* h, const char* path) {
  if (!h || !path) { set_error("history_save: bad arguments"); return ARAP_ERR_INVALID; }
  std::ofstream o(path);
  if (!o.is_open()) { set_error(std::string("history_save: cannot open ") + path); return ARAP_ERR_IO; }
  o << (h->nodes_on_mesh ? 1 : 0) << " ";
  o << "Nodes: " << h->nodes.size() << " ";
  for (int v : h->nodes) o << v << " ";
  o << "Total_Operations: " << h->total_operations << " ";
  o << "Move_Operations: " << h->move_operations << " ";
  o << "Operation_Types: " << h->operation_types.size() << " ";
  for (int v : h->operation_types) o << v << " ";
  o << "Block_Nodes: " << h->block_nodes.size() << " ";
  for (auto& b : h->block_nodes) { o << b.size() << " "; for (uint32_t v : b) o << v << " "; }
  o << "Mouse_Movements: " << h->mouse_movements.size() << " ";
  for (auto& m : h->mouse_movements) { o << m.size() / 3 << " "; for (float v : m) o << v << " "; }
  o << "Blocks_Types_Moves: " << h->blocks_types_moves.size() << " ";
  for (auto& b : h->blocks_types_moves) { o << b.size() << " "; for (int v : b) o << v << " "; }
  o << "Energy_on_Center: " << h->energy_on_centers.size() << " ";
  for (int v : h->energy_on_centers) o << v << " ";
  o << "Twist_Axis: " << h->twist_axis.size() / 4 << " ";
  for (float v : h->twist_axis) o << v << " ";
  return ARAP_OK;
}

extern "C" int arap_history_free(arap_history* h) { delete h; return ARAP_OK; }

extern "C" int arap_history_new(arap_history** out, int nodes_on_mesh, const int* node_anchor, int n_nodes) {
  if (!out || n_nodes < 0) return ARAP_ERR_INVALID;
  auto* h = new arap_history();
  h->nodes_on_mesh = nodes_on_mesh ? 1 : 0;
  if (n_nodes) h->nodes.assign(node_anchor, node_anchor + n_nodes);
  *out = h;
  return ARAP_OK;
}
extern "C" int arap_history_add_block(arap_history* h, const uint32_t* nodes, int n) {
  if (!h || n < 0) return ARAP_ERR_INVALID;
  h->block_nodes.emplace_back(nodes, nodes + n);
  h->operation_types.push_back(0); h->total_operations++;
  return ARAP_OK;
}
extern "C" int arap_history_add_move(arap_history* h, int op_type, const float* mv, int n, const int* bt, int n_types,
                                     int energy_on_center, const float twist_axis[4]) {
  if (!h || op_type < 1 || op_type > 4 || n < 0) return ARAP_ERR_INVALID;
  h->mouse_movements.emplace_back(mv, mv + (size_t)n * 3);
  h->blocks_types_moves.emplace_back(bt, bt + n_types);
  h->energy_on_centers.push_back(energy_on_center);
  for (int i = 0; i < 4; i++) h->twist_axis.push_back(twist_axis ? twist_axis[i] : 0.f);
  h->operation_types.push_back(op_type); h->total_operations++; h->move_operations++;
  return ARAP_OK;
}
extern "C" int arap_history_summary(const arap_history* h, int* o) {
  if (!h || !o) return ARAP_ERR_INVALID;
  o[0] = h->nodes_on_mesh; o[1] = (int)h->nodes.size(); o[2] = h->total_operations; o[3] = h->move_operations;
  o[4] = (int)h->block_nodes.size(); o[5] = (int)h->mouse_movements.size();
  return ARAP_OK;
}
extern "C" int arap_history_nodes(const arap_history* h, int* out) { if (!h) return ARAP_ERR_INVALID; memcpy(out, h->nodes.data(), sizeof(int) * h->nodes.size()); return ARAP_OK; }
extern "C" int arap_history_ops(const arap_history* h, int* out) { if (!h) return ARAP_ERR_INVALID; memcpy(out, h->operation_types.data(), sizeof(int) * h->operation_types.size()); return ARAP_OK; }
extern "C" int arap_history_block(const arap_history* h, int i, uint32_t* nodes_out, int* n) {
  if (!h || i < 0 || i >= (int)h->block_nodes.size()) return ARAP_ERR_INVALID;
  if (n) *n = (int)h->block_nodes[i].size();
  if (nodes_out) memcpy(nodes_out, h->block_nodes[i].data(), sizeof(uint32_t) * h->block_nodes[i].size());
  return ARAP_OK;
}
extern "C" int arap_history_move(const arap_history* h, int i, float* mv, int* n, int* bt, int* n_types, int* eoc, float* axis4) {
  if (!h || i < 0 || i >= (int)h->mouse_movements.size()) return ARAP_ERR_INVALID;
  if (n) *n = (int)h->mouse_movements[i].size() / 3;
  if (mv) memcpy(mv, h->mouse_movements[i].data(), sizeof(float) * h->mouse_movements[i].size());
  if (n_types) *n_types = i < (int)h->blocks_types_moves.size() ? (int)h->blocks_types_moves[i].size() : 0;
  if (bt && i < (int)h->blocks_types_moves.size()) memcpy(bt, h->blocks_types_moves[i].data(), sizeof(int) * h->blocks_types_moves[i].size());
  if (eoc) *eoc = i < (int)h->energy_on_centers.size() ? h->energy_on_centers[i] : 0;
  if (axis4) for (int t = 0; t < 4; t++) axis4[t] = (size_t)(4 * i + t) < h->twist_axis.size() ? h->twist_axis[4 * i + t] : 0.f;
  return ARAP_OK;
}

// LoadMeshPoints: only lines starting with "v " contribute (helper.cpp:264-282)
extern "C" int arap_graph_obj_load(const char* path, float* pts, int* n) {
  if (!path || !n) { set_error("graph_obj_load: bad arguments"); return ARAP_ERR_INVALID; }
  std::ifstream f(path);
  if (!f.is_open()) { set_error(std::string("graph_obj_load: cannot open ") + path); return ARAP_ERR_IO; }
  std::string line; int cnt = 0;
  while (std::getline(f, line)) {
    if (line.substr(0, 2) == "v ") {
      std::istringstream ss(line.substr(2));
      float p[3] = {0, 0, 0};
      ss >> p[0] >> p[1] >> p[2];
      if (pts) { pts[3 * cnt] = p[0]; pts[3 * cnt + 1] = p[1]; pts[3 * cnt + 2] = p[2]; }
      cnt++;
    }
  }
  *n = cnt;
  return ARAP_OK;
}
extern "C" int arap_graph_obj_save(const char* path, const float* pts, int n) {
  std::ofstream f(path);
  if (!f.is_open()) { set_error(std::string("graph_obj_save: cannot open ") + path); return ARAP_ERR_IO; }
  for (int i = 0; i < n; i++) f << "v " << pts[3 * i] << " " << pts[3 * i + 1] << " " << pts[3 * i + 2] << "\n";
  return ARAP_OK;
}

extern "C" int arap_config_load(const char* path, int* grid_num, int* is_synthetic, int* has_soup, int* high_quality) {
  std::ifstream f(path);
  if (!f.is_open()) { set_error(std::string("config_load: cannot open ") + path); return ARAP_ERR_IO; }
  std::vector<int> v; std::string line;
  while (std::getline(f, line)) { std::istringstream iss(line); int x; while (iss >> x) v.push_back(x); }
  if (v.size() < 3) { set_error("config_load: expected `grid_num is_synthetic conf2`"); return ARAP_ERR_IO; }
  if (grid_num) *grid_num = v[0];
  if (is_synthetic) *is_synthetic = v[1];
  int soup = 0, hq = 0;
  if (v[2] == 1) { soup = 1; hq = 1; } else if (v[2] == 2) { soup = 0; hq = 1; }
  if (has_soup) *has_soup = soup;
  if (high_quality) *high_quality = hq;
  return ARAP_OK;
}

// ------------------------------------------------------------------ 3DGS PLY (GV:43-339)
// Binary little-endian vertex records of 62 floats (x y z nx ny nz f_dc[3] f_rest[45] opacity scale[3] rot[4]) or 63 with a
// trailing `index` property (the reference's "soup" files, helper.hpp:98-108).  The loader applies the reference's
// activations (normalised quaternion, exp scale, sigmoid opacity), interleaves the channel-major f_rest block into
// 16 x RGB, and orders the Gaussians by the 63-bit Morton code of their position inside the cloud's bounding box.
namespace {
struct PlyHeader { long long count = 0; int props = 0; std::streampos data = 0; };
int ply_header(std::ifstream& f, const char* path, PlyHeader& h) {
  std::string line;
  bool fmt_ok = false, ended = false;
  if (!std::getline(f, line) || line.substr(0, 3) != "ply") { set_error(std::string("ply_load: not a PLY file: ") + path); return ARAP_ERR_IO; }
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ss(line); std::string a, b;
    ss >> a;
    if (a == "format") { ss >> b; fmt_ok = b == "binary_little_endian"; }
    else if (a == "element") { ss >> b; if (b == "vertex") ss >> h.count; }
    else if (a == "property") { ss >> b; if (b != "float") { set_error("ply_load: only float properties are supported"); return ARAP_ERR_IO; } h.props++; }
    else if (a == "end_header") { ended = true; break; }
  }
  if (!ended || !fmt_ok || h.count < 0 || (h.props != 62 && h.props != 63)) {
    set_error("ply_load: expected binary_little_endian with 62 (or 63, with index) float properties per vertex");
    return ARAP_ERR_IO;
  }
  h.data = f.tellg();
  return ARAP_OK;
}
inline float sigmoidf(float x) { return 1.0f / (1.0f + std::exp(-x)); }              // helper.hpp sigmoid
inline float inverse_sigmoidf(float x) { return std::log(x / (1.0f - x)); }          // helper.hpp inverse_sigmoid
}  // namespace

extern "C" int arap_ply_load(const char* path, long long* n, float* pos, float* rot, float* scale, float* opacity, float* shs,
                             int* index, float aabb_min[3], float aabb_max[3]) {
  if (!path || !n) { set_error("ply_load: bad arguments"); return ARAP_ERR_INVALID; }
  std::ifstream f(path, std::ios_base::binary);
  if (!f.good()) { set_error(std::string("ply_load: cannot open ") + path); return ARAP_ERR_IO; }
  PlyHeader h;
  int rc = ply_header(f, path, h); if (rc) return rc;
  if (!pos) { *n = h.count; return ARAP_OK; }   // first call: the count
  if (*n < h.count) { set_error("ply_load: output arrays are smaller than the vertex count"); return ARAP_ERR_INVALID; }
  const size_t P = (size_t)h.props, cnt = (size_t)h.count;
  std::vector<float> rec(cnt * P);
  f.read(reinterpret_cast<char*>(rec.data()), (std::streamsize)(rec.size() * sizeof(float)));
  if ((size_t)f.gcount() != rec.size() * sizeof(float)) { set_error("ply_load: truncated vertex data"); return ARAP_ERR_IO; }
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (size_t i = 0; i < cnt; i++)
    for (int c = 0; c < 3; c++) { mn[c] = std::min(mn[c], rec[i * P + c]); mx[c] = std::max(mx[c], rec[i * P + c]); }
  // Morton order (GV:91-116): 21 bits per axis of floor((2^21 - 1) * (p - min) / (max - min)), x in bit 3i, y in 3i+1, z in 3i+2.
  // The reference's std::sort leaves the order of equal codes unspecified; ties keep file order here.
  std::vector<std::pair<uint64_t, uint32_t>> order(cnt);
  for (size_t i = 0; i < cnt; i++) {
    uint64_t code = 0;
    int q[3];
    for (int c = 0; c < 3; c++) {
      const float rel = (rec[i * P + c] - mn[c]) / (mx[c] - mn[c]);
      q[c] = (int)((float)((1 << 21) - 1) * rel);
    }
    for (int b = 0; b < 21; b++)
      for (int c = 0; c < 3; c++) code |= (uint64_t)((unsigned)q[c] & (1u << b)) << (2 * b + c);
    order[i] = {code, (uint32_t)i};
  }
  std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  for (size_t k = 0; k < cnt; k++) {
    const float* r = rec.data() + (size_t)order[k].second * P;
    for (int c = 0; c < 3; c++) pos[3 * k + c] = r[c];
    if (rot) {
      const float* q = r + 58;
      float l2 = 0.f;
      for (int j = 0; j < 4; j++) l2 += q[j] * q[j];
      const float len = std::sqrt(l2);
      for (int j = 0; j < 4; j++) rot[4 * k + j] = q[j] / len;
    }
    if (scale) for (int j = 0; j < 3; j++) scale[3 * k + j] = std::exp(r[55 + j]);
    if (opacity) opacity[k] = sigmoidf(r[54]);
    if (shs) {
      float* o = shs + 48 * k; const float* sh = r + 6;
      o[0] = sh[0]; o[1] = sh[1]; o[2] = sh[2];
      for (int j = 1; j < 16; j++) { o[3 * j] = sh[(j - 1) + 3]; o[3 * j + 1] = sh[(j - 1) + 18]; o[3 * j + 2] = sh[(j - 1) + 33]; }
    }
    if (index) index[k] = P == 63 ? (int)std::lround(r[62]) : (int)order[k].second;
  }
  if (aabb_min) for (int c = 0; c < 3; c++) aabb_min[c] = mn[c];
  if (aabb_max) for (int c = 0; c < 3; c++) aabb_max[c] = mx[c];
  *n = h.count;
  return ARAP_OK;
}

// savePly (GV:273-339): Gaussians outside [box_min, box_max] or flagged `skip` are dropped; scale -> log, opacity ->
// inverse sigmoid, SH back to channel-major f_rest; header text identical to the reference's.
extern "C" int arap_ply_save(const char* path, long long n, const float* pos, const float* rot, const float* scale, const float* opacity,
                             const float* shs, const float box_min[3], const float box_max[3], const uint8_t* skip, long long* written) {
  if (!path || n < 0 || !pos || !rot || !scale || !opacity || !shs) { set_error("ply_save: bad arguments"); return ARAP_ERR_INVALID; }
  auto keep = [&](long long i) {
    if (box_min && box_max)
      for (int c = 0; c < 3; c++) if (pos[3 * i + c] < box_min[c] || pos[3 * i + c] > box_max[c]) return false;
    return !(skip && skip[i]);
  };
  long long count = 0;
  for (long long i = 0; i < n; i++) count += keep(i);
  std::ofstream f(path, std::ios_base::binary);
  if (!f.is_open()) { set_error(std::string("ply_save: cannot open ") + path); return ARAP_ERR_IO; }
  f << "ply\nformat binary_little_endian 1.0\nelement vertex " << count << "\n";
  for (const char* p : {"x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"}) f << "property float " << p << "\n";
  for (int i = 0; i < 45; i++) f << "property float f_rest_" << i << "\n";
  for (const char* p : {"opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"}) f << "property float " << p << "\n";
  f << "end_header\n";
  std::vector<float> rec((size_t)count * 62, 0.0f);
  size_t k = 0;
  for (long long i = 0; i < n; i++) {
    if (!keep(i)) continue;
    float* r = rec.data() + k * 62; k++;
    for (int c = 0; c < 3; c++) r[c] = pos[3 * i + c];
    const float* sh = shs + 48 * i;
    r[6] = sh[0]; r[7] = sh[1]; r[8] = sh[2];
    for (int j = 1; j < 16; j++) { r[6 + (j - 1) + 3] = sh[3 * j]; r[6 + (j - 1) + 18] = sh[3 * j + 1]; r[6 + (j - 1) + 33] = sh[3 * j + 2]; }
    r[54] = inverse_sigmoidf(opacity[i]);
    for (int j = 0; j < 3; j++) r[55 + j] = std::log(scale[3 * i + j]);
    for (int j = 0; j < 4; j++) r[58 + j] = rot[4 * i + j];
  }
  f.write(reinterpret_cast<const char*>(rec.data()), (std::streamsize)(rec.size() * sizeof(float)));
  if (!f.good()) { set_error("ply_save: write failed"); return ARAP_ERR_IO; }
  if (written) *written = count;
  return ARAP_OK;
}

extern "C" int arap_point_rotate_by_axis(const float point[3], const float center[3], const float axis[4], float radian, float out[3]) {
  float cost = std::cos(radian), sint = std::sin(radian);
  float norm = std::sqrt(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
  float x = axis[0] / norm, y = axis[1] / norm, z = axis[2] / norm;
  out[0] = (x * x * (1 - cost) + cost) * point[0] + (x * y * (1 - cost) - z * sint) * point[1] + (x * z * (1 - cost) + y * sint) * point[2];
  out[1] = (y * x * (1 - cost) + z * sint) * point[0] + (y * y * (1 - cost) + cost) * point[1] + (y * z * (1 - cost) - x * sint) * point[2];
  out[2] = (z * x * (1 - cost) - y * sint) * point[0] + (z * y * (1 - cost) + x * sint) * point[1] + (z * z * (1 - cost) + cost) * point[2];
  float a = center[0], b = center[1], c = center[2];
  out[0] += (a * (y * y + z * z) - x * (b * y + c * z)) * (1 - cost) + (b * z - c * y) * sint;
  out[1] += (b * (x * x + z * z) - y * (a * x + c * z)) * (1 - cost) + (c * x - a * z) * sint;
  out[2] += (c * (x * x + y * y) - z * (a * x + b * y)) * (1 - cost) + (a * y - b * x) * sint;
  return ARAP_OK;
}

namespace {
int push_blocks(arap_ctx* ctx, const std::vector<std::vector<uint32_t>>& blocks, const std::vector<int>& types) {
  std::vector<int> off{0}; std::vector<uint32_t> nodes;
  for (auto& b : blocks) { nodes.insert(nodes.end(), b.begin(), b.end()); off.push_back((int)nodes.size()); }
  if (nodes.empty()) nodes.push_back(0);
  return arap_set_blocks(ctx, (int)blocks.size(), off.data(), nodes.data(), types.empty() ? nullptr : types.data());
}
}  // namespace

// The replay state machine of GaussianView::onUpdate (GV:1757-1916), headless: one drag step per recorded movement.
extern "C" int arap_replay(arap_ctx* ctx, const arap_history* h, int rebuild_graph, int* steps_run) {
  if (!ctx || !h) { set_error("replay: bad arguments"); return ARAP_ERR_INVALID; }
  int rc;
  if (rebuild_graph) {
    int k = arapgs::session_k(ctx);
    if (k <= 0) { set_error("replay: build a graph first (its k is reused; k is not stored in deform.txt)"); return ARAP_ERR_STATE; }
    if ((rc = arap_graph_build_anchors(ctx, h->nodes.data(), (int)h->nodes.size(), k))) return rc;
  }
  std::vector<std::vector<uint32_t>> blocks; std::vector<int> types;
  if ((rc = push_blocks(ctx, blocks, types))) return rc;
  int add_idx = 0, move_idx = 0, steps = 0;
  for (int step = 0; step < h->total_operations && step < (int)h->operation_types.size(); step++) {
    const int op = h->operation_types[step];
    if (op < 0) {
      const int cur = -(op + 1);
      if (!blocks.empty() && cur < (int)blocks.size()) { blocks.erase(blocks.begin() + cur); types.erase(types.begin() + cur); if ((rc = push_blocks(ctx, blocks, types))) return rc; }
    } else if (op == 0) {
      if (add_idx >= (int)h->block_nodes.size()) { set_error("replay: more add-block ops than blocks"); return ARAP_ERR_IO; }
      types.push_back(blocks.empty() ? 1 : 0);  // first block active, later ones pinned (GV:1805-1811)
      blocks.push_back(h->block_nodes[add_idx++]);
      if ((rc = push_blocks(ctx, blocks, types))) return rc;
    } else {
      if (move_idx >= (int)h->mouse_movements.size()) { set_error("replay: more move ops than movements"); return ARAP_ERR_IO; }
      if (move_idx < (int)h->blocks_types_moves.size()) {
        const auto& bt = h->blocks_types_moves[move_idx];
        if (bt.size() != blocks.size()) { set_error("replay: block-type list does not match the block count"); return ARAP_ERR_IO; }
        types = bt;
        if ((rc = push_blocks(ctx, blocks, types))) return rc;
      }
      const auto& mv = h->mouse_movements[move_idx];
      static const float no_axis[4] = {0.f, 0.f, 0.f, 0.f};
      const float* axis = no_axis;
      if (op == 2) {   // the axis of a twist is read only for twists, and only when the file really has one for this move
        if (h->twist_axis.size() < 4 * ((size_t)move_idx + 1)) { set_error("replay: twist move without a Twist_Axis entry"); return ARAP_ERR_IO; }
        axis = &h->twist_axis[4 * (size_t)move_idx];
        if (!(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2] > 0.f)) { set_error("replay: zero or non-finite twist axis"); return ARAP_ERR_IO; }
      }
      for (size_t s = 0; s + 2 < mv.size() + 0 && s < mv.size(); s += 3) {
        int on_center = 0;
        if (op == 1) { if ((rc = arap_aim_translate(ctx, &mv[s]))) return rc; on_center = 1; }
        else if (op == 2) { if ((rc = arap_aim_twist(ctx, axis, (int)mv[s]))) return rc; }
        else if (op == 3) { if ((rc = arap_aim_scale(ctx, (int)mv[s]))) return rc; }
        else if (op == 4) { if ((rc = arap_aim_translate(ctx, &mv[s]))) return rc; }
        if ((rc = arap_step(ctx, on_center))) return rc;
        steps++;
      }
      move_idx++;
    }
  }
  if (steps_run) *steps_run = steps;
  return arap_sync(ctx);
}

// LoadDeformScript0 (bend) / LoadDeformScript1 (twist) + RunDeformScript + the script loop.
extern "C" int arap_run_script(arap_ctx* ctx, int script_id, int* steps_run) {
  if (!ctx) return ARAP_ERR_INVALID;
  if (script_id != 0 && script_id != 1) { set_error("run_script: only scripts 0 (bend) and 1 (twist) are implemented"); return ARAP_ERR_UNSUPPORTED; }
  const int M = arapgs::session_num_nodes(ctx);
  const std::vector<uint32_t> block1 = {0, 20, 53, 59, 63, 64, 67, 68, 145, 167, 189, 190, 192, 196, 197, 199};   // GV:2517
  const std::vector<uint32_t> block2 = {1, 7, 21, 26, 54, 70, 80, 127, 176, 178, 179, 181, 183, 184, 185, 186};   // GV:2518
  if (M < 200) { set_error("run_script: the stripes scripts index nodes up to 199"); return ARAP_ERR_STATE; }
  int rc;
  std::vector<float> node_pos((size_t)M * 3);
  if ((rc = arap_download_nodes(ctx, node_pos.data(), nullptr, nullptr))) return rc;
  const int inter_steps = 50;
  const float center_start[3] = {0.0f, -1.5f, 0.0f};
  const double Pi = 3.1415926535;  // helper.hpp:47
  std::vector<std::vector<float>> aims(inter_steps);
  for (int df = 1; df <= inter_steps; df++) {
    auto& a = aims[df - 1]; a.resize(block2.size() * 3);
    for (size_t i = 0; i < block2.size(); i++) {
      const float* sp = &node_pos[3 * block2[i]];
      float o[3];
      if (script_id == 0) {
        const float kk = 5.4f;
        const float aa = (float)(3.0 * Pi * Pi / (kk * kk));
        const float axis[4] = {0.0f, 0.0f, 1.0f, 0.0f};
        const float x = (float)((float)df * (-kk / Pi) / (float)inter_steps);
        const float y = aa * x * x;
        const float radian = (float)(-(float)df * Pi / (float)inter_steps);
        arap_point_rotate_by_axis(sp, center_start, axis, radian, o);
        o[0] += x; o[1] += y; o[2] += 0.0f;
      } else {
        const float axis[4] = {0.0f, 1.0f, 0.0f, 0.0f};
        const float radian = (float)(-(float)df * 1.5f * Pi / (float)inter_steps);
        arap_point_rotate_by_axis(sp, center_start, axis, radian, o);
      }
      a[3 * i] = o[0]; a[3 * i + 1] = o[1]; a[3 * i + 2] = o[2];
    }
  }
  std::vector<std::vector<uint32_t>> blocks = {block1, block2};
  std::vector<int> types = {0, 1};
  if ((rc = push_blocks(ctx, blocks, types))) return rc;
  const std::vector<float> temp_aim = node_pos;  // temp_aim_nodes = deform_graph.nodes (GV:2812)
  std::vector<float> aim((size_t)M * 3);
  int steps = 0;
  for (int s = 0; s < inter_steps; s++) {
    if ((rc = arap_aim_get(ctx, aim.data()))) return rc;
    for (uint32_t n : block1) for (int c = 0; c < 3; c++) aim[3 * n + c] = temp_aim[3 * n + c];
    for (size_t i = 0; i < block2.size(); i++) for (int c = 0; c < 3; c++) aim[3 * block2[i] + c] = aims[s][3 * i + c];
    if ((rc = arap_aim_set(ctx, aim.data()))) return rc;
    if ((rc = arap_sync(ctx))) return rc;  // `aim` is reused next iteration
    if ((rc = arap_step(ctx, 0))) return rc;
    steps++;
  }
  if (steps_run) *steps_run = steps;
  return arap_sync(ctx);
}
