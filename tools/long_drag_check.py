#!/usr/bin/env python
"""Robustness check (GPU): a long drag on the bench workload with direction changes, pauses and a block edit; prints the
per-step Gauss-Newton / PCG iteration counts and fails on any solver flag.  Not a test (needs ~30 s on a B200)."""
import importlib, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge

pkg = ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
wl = sys.argv[1] if len(sys.argv) > 1 else "sphere1m"
cfg = scenes.CONFIGS[wl]
sc = scenes.make_scene(wl, n=int(sys.argv[2]) if len(sys.argv) > 2 else cfg["n"])
s = pkg.Session(device=0, grid_num=cfg["grid"], knn_k=cfg["k"], node_num=cfg["nodes"])
s.set_gaussians(sc["pos"], sc["rot"], sc["scale"], sc["opacity"], sc["shs"])
s.grid_build(); s.grid_eval(0)
g = s.graph_build_fps()
blocks, types = scenes.cap_blocks(g["node_pos"])
s.set_blocks(blocks, types)
log = []
def run(n, d):
    for _ in range(n):
        s.aim_translate(d); s.step(False)
        st = s.solve_stats()
        assert st["flags"] == 0, st
        log.append((st["gn_iters"], st["cg_iters"], st["halvings"]))
run(40, [0, 0, 0.002]); run(3, [0, 0, 0]); run(30, [0, 0, -0.002]); run(20, [0.002, 0.001, 0])
s.set_blocks(blocks[::-1], types)           # swap active / pinned caps: warm-start buffer is invalidated
run(20, [0, 0, 0.003])
for y in (10, -10, 5):
    s.aim_twist([0.0, 0.0, 1.0, 0.0], y); s.step(False); st = s.solve_stats(); assert st["flags"] == 0, st
    log.append((st["gn_iters"], st["cg_iters"], st["halvings"]))
out = s.download_gaussians()
assert all(np.isfinite(out[k]).all() for k in out)
a = np.array(log)
print(json.dumps({"steps": len(log), "gn_iters_max": int(a[:, 0].max()), "cg_iters_mean": float(a[:, 1].mean()), "cg_iters_max": int(a[:, 1].max()),
                  "halvings": int(a[:, 2].sum()), "cg_first10": a[:10, 1].tolist(), "cg_after_pause": a[40:46, 1].tolist(),
                  "cg_after_reverse": a[43:49, 1].tolist(), "cg_after_block_edit": a[93:97, 1].tolist(), "cg_twist": a[-3:, 1].tolist()}))
