#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): grid, graph, per-node and centre constraints, a few drag
steps with the one-barrier solver (cold + warm-started), a stroke-end rebuild, and a three-slab sharded-scene grid.

    compute-sanitizer --tool racecheck python tests/studies/sanitizer_solve.py"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
pkg = ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
sc = scenes.make_scene("sphere1m", n=6000)
for on_center in (False, True):
    s = pkg.Session(device=0, grid_num=16, knn_k=10, node_num=400, lbs_mode=3)
    s.set_gaussians(sc["pos"], sc["rot"], sc["scale"], sc["opacity"], sc["shs"])
    s.grid_build(); s.grid_eval(0)
    g = s.graph_build_fps()
    npz = g["node_pos"]
    blocks = [np.nonzero(npz[:, 2] > 0.3)[0].astype(np.uint32), np.nonzero(npz[:, 2] < -0.3)[0].astype(np.uint32),
              np.nonzero((npz[:, 0] > 0.4) & (np.abs(npz[:, 2]) < 0.2))[0].astype(np.uint32)]
    s.set_blocks(blocks, [1, 0, 0])
    for step in range(3):
        s.aim_translate([0.002, 0.0, 0.01]); s.step(on_center)
        st = s.solve_stats()
        assert st["flags"] == 0, st
    s.grid_update_lists(); s.grid_eval(1)
    print("on_center", on_center, "gn", st["gn_iters"], "products", st["cg_iters"], "kernel phases", st["phase_ns"][3] == 0)
    s.close()
full = pkg.Session(device=0, grid_num=16, knn_k=10, node_num=100)
full.set_gaussians(sc["pos"], sc["rot"], sc["scale"], sc["opacity"], sc["shs"])
full.grid_build(); o = full.download_gaussians(); full.close()
for lo, hi in ((0, 5), (5, 11), (11, 16)):
    s = pkg.Session(device=0, grid_num=16, knn_k=10, node_num=100)
    s.set_gaussians(o["pos"], o["rot"], o["scale"], o["opacity"], o["shs"])
    s.comm_init(pkg.comm_unique_id(), 0, 1)
    print("slab", lo, hi, s.comm_grid_build(lo, hi)["valid_cells"]); s.grid_eval(0); s.close()
print("done")
