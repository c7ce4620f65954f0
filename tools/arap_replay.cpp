// Headless replay: point_cloud.ply + deform.txt (+ graph.obj) -> deformed point_cloud.ply, through the C ABI only.
// The same sequence the reference's viewer runs for "Load Deformation" + "Run Historical Deform"
// (GaussianView.cpp:425-882 init, 4961-5095 LoadDeformation, 1757-1916 RunHistoricalDeform, 273-339 savePly),
// without a window.  Needs a CUDA device (libarapgs has no CPU path).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/arapgs.h"

#define CHECK(call)                                                                 \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != ARAP_OK) {                                                           \
      std::fprintf(stderr, "arap_replay: %s failed (%d): %s\n", #call, rc_, arap_last_error()); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

static void usage() {
  std::fprintf(stderr,
               "usage: arap_replay <point_cloud.ply> <deform.txt> <out.ply> [--grid G] [--k K] [--nodes M] [--hq 0|1]\n"
               "                   [--config <ply>_config.txt] [--graph graph.obj] [--device D] [--no-rebuild]\n");
}

int main(int argc, char** argv) {
  if (argc < 4) { usage(); return 2; }
  const char *ply = argv[1], *deform = argv[2], *out = argv[3], *graph = nullptr, *config = nullptr;
  int grid = 64, k = 10, nodes = 0, hq = 0, device = 0, rebuild = 1;
  for (int i = 4; i < argc; i++) {
    auto val = [&](int& dst) { if (i + 1 < argc) dst = std::atoi(argv[++i]); };
    if (!std::strcmp(argv[i], "--grid")) val(grid);
    else if (!std::strcmp(argv[i], "--k")) val(k);
    else if (!std::strcmp(argv[i], "--nodes")) val(nodes);
    else if (!std::strcmp(argv[i], "--hq")) val(hq);
    else if (!std::strcmp(argv[i], "--device")) val(device);
    else if (!std::strcmp(argv[i], "--graph") && i + 1 < argc) graph = argv[++i];
    else if (!std::strcmp(argv[i], "--config") && i + 1 < argc) config = argv[++i];
    else if (!std::strcmp(argv[i], "--no-rebuild")) rebuild = 0;
    else { usage(); return 2; }
  }
  if (config) { int synth = 0, soup = 0; CHECK(arap_config_load(config, &grid, &synth, &soup, &hq)); }

  long long n = 0;
  CHECK(arap_ply_load(ply, &n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
  std::vector<float> pos(3 * n), rot(4 * n), scale(3 * n), opacity(n), shs(48 * n);
  float bmin[3], bmax[3];
  CHECK(arap_ply_load(ply, &n, pos.data(), rot.data(), scale.data(), opacity.data(), shs.data(), nullptr, bmin, bmax));
  std::printf("loaded %lld Gaussians, box [%g %g %g] - [%g %g %g]\n", n, bmin[0], bmin[1], bmin[2], bmax[0], bmax[1], bmax[2]);

  arap_history* h = nullptr;
  CHECK(arap_history_load(deform, &h));
  int sum[6] = {0};
  CHECK(arap_history_summary(h, sum));

  arap_params prm;
  CHECK(arap_default_params(&prm));
  prm.grid_num = grid; prm.knn_k = k; prm.high_quality = hq;
  if (nodes > 0) prm.node_num = nodes;
  arap_ctx* ctx = nullptr;
  CHECK(arap_create(&ctx, device, nullptr, &prm));
  CHECK(arap_set_gaussians(ctx, n, pos.data(), rot.data(), scale.data(), opacity.data(), shs.data(), 0));
  CHECK(arap_grid_build(ctx));
  CHECK(arap_grid_eval(ctx, 0));
  if (graph) {   // nodes on the mesh points of graph.obj (LoadMeshForGraph, GV:2818-2918)
    int m = 0;
    CHECK(arap_graph_obj_load(graph, nullptr, &m));
    std::vector<float> pts(3 * (size_t)m);
    CHECK(arap_graph_obj_load(graph, pts.data(), &m));
    CHECK(arap_set_mesh_points(ctx, pts.data(), m, 1));
    std::printf("graph.obj: %d mesh points\n", m);
  }
  CHECK(arap_graph_build_fps(ctx, nodes > 0 ? nodes : prm.node_num, k));   // GaussianView::init builds the FPS graph; the replay reuses its k
  int steps = 0;
  CHECK(arap_replay(ctx, h, rebuild, &steps));
  CHECK(arap_sync(ctx));
  std::printf("replayed %d drag steps\n", steps);
  CHECK(arap_download_gaussians(ctx, pos.data(), rot.data(), scale.data(), opacity.data(), shs.data()));
  long long written = 0;
  CHECK(arap_ply_save(out, n, pos.data(), rot.data(), scale.data(), opacity.data(), shs.data(), nullptr, nullptr, nullptr, &written));
  std::printf("wrote %lld Gaussians to %s\n", written, out);
  arap_history_free(h);
  arap_destroy(ctx);
  return 0;
}
