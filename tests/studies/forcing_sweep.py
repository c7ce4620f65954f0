"""GPU experiment: node-transform / covariance deviation from the oracle vs (newton_eta0, cg_tol)."""
import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as ge
pkg = ge.load_package(); import importlib; scenes = importlib.import_module(ge.PKG + ".scenes")
from oracle.session import OracleSession
from test_gpu_session import _cov
kw = dict(grid_num=32, knn_k=8, node_num=150)
sc = scenes.make_scene("sphere1m", n=30000)
def run_oracle():
    o = OracleSession(sc, **kw); o.grid_build(); o.graph_build_fps()
    return o
o = run_oracle()
blocks, types = None, None
res_o = []
for eta0, tol in [(0, 1e-10), (1e-6, 1e-10), (1e-5, 1e-10), (1e-4, 1e-10), (1e-3, 1e-10), (1e-5, 1e-9), (1e-4, 1e-9)]:
    s = pkg.Session(device=0, **kw)
    s.set_gaussians(sc["pos"], sc["rot"], sc["scale"], sc["opacity"], sc["shs"]); s.grid_build()
    g = s.graph_build_fps()
    s.set_params(newton_eta0=eta0, cg_tol=tol)
    o = run_oracle()
    blocks, types = scenes.cap_blocks(g["node_pos"], lo=-0.3, hi=0.3)
    s.set_blocks(blocks, types); o.set_blocks(blocks, types)
    dmax = 0; its = []
    for step in range(4):
        s.aim_translate([0.002, 0.0, 0.01]); o.aim_translate([0.002, 0.0, 0.01])
        s.solve(False); so = o.solve(False); st = s.solve_stats()
        _, rot, trans = s.download_nodes()
        dmax = max(dmax, np.abs(rot - o.rot).max(), np.abs(trans - o.trans).max())
        its.append((st["gn_iters"], so["iters"], st["cg_iters"]))
        s.apply(); o.apply()
    out = s.download_gaussians()
    Cg, Co = _cov(out["rot"], out["scale"]), _cov(o.g["rot"], o.g["scale"])
    rel = np.linalg.norm(Cg - Co, axis=(1, 2)) / np.linalg.norm(Co, axis=(1, 2))
    print(f"eta0={eta0:g} cg_tol={tol:g}: node dmax {dmax:.2e}  pos {np.abs(out['pos']-o.g['pos']).max():.2e}  cov rel max {rel.max():.2e} p99.9 {np.quantile(rel,0.999):.2e}  its {its}", flush=True)
