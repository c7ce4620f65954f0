#!/usr/bin/env python
"""CPU study (scipy, no GPU): classical Jacobi-PCG against pipelined PCG (Ghysels & Vanroose 2014; one reduction per
iteration, the recurrences of csrc/solve_pipe.cu) on the oracle's Jacobian of the first Gauss-Newton system of a drag step.

    python tests/studies/pipelined_cg.py [nodes=4000]

M=4000: identical iteration counts at 1e-6 / 1e-8 (208 / 309), 411 vs 413 at 1e-10; solutions agree to 3e-9 relative."""
import importlib, sys, time
from pathlib import Path
import numpy as np
import scipy.sparse as sp
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
import oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = 200000; k = 10
sc = scenes.make_scene("sphere1m", n=N)
O.set_threads(8)
anchor = O.fps(sc["pos"], M)
nodes = sc["pos"][anchor].copy()
nbr = O.graph_edges(nodes, k)
idx, w = O.knn_weights(nodes, nodes, k)
blocks, types = scenes.cap_blocks(nodes)
aim = nodes.copy(); aim[blocks[0]] += np.array([0, 0, 0.02], np.float32)
rot = np.tile(np.eye(3).reshape(9), (M, 1)); trans = np.zeros((M, 3))
R, C, V, f, (m, n) = O.jacobian(nodes, nbr, idx[:, :k], w, np.zeros(M, np.uint8), blocks, types, aim, False, rot, trans)
J = sp.csr_matrix((V, (R, C)), shape=(m, n))
H = (J.T @ J).tocsr()
g = -(J.T @ f)
d = H.diagonal(); di = 1.0 / d
g0 = np.linalg.norm(g)
def pcg(tol):
    x = np.zeros(n); r = g.copy(); z = r * di; p = z.copy(); rz = r @ z
    for it in range(1, 5000):
        q = H @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= tol * g0: return it, x
        z = r * di; rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
def pipe(tol):
    x = np.zeros(n); r = g.copy(); u = r * di; w_ = H @ u
    z = np.zeros(n); s = np.zeros(n); p = np.zeros(n)
    gam_old = 1.0; alpha_old = 1.0
    for it in range(0, 5000):
        gam = r @ u; delta = w_ @ u; rr = r @ r
        if np.sqrt(rr) <= tol * g0: return it, x
        mm = w_ * di; nn = H @ mm
        if it > 0:
            beta = gam / gam_old; alpha = gam / (delta - beta * gam / alpha_old)
        else:
            beta = 0.0; alpha = gam / delta
        z = nn + beta * z; s = w_ + beta * s; p = u + beta * p
        x += alpha * p; r -= alpha * s; w_ -= alpha * z; u = r * di
        gam_old = gam; alpha_old = alpha
for tol in (1e-6, 1e-8, 1e-10):
    i1, x1 = pcg(tol); i2, x2 = pipe(tol)
    t1 = np.linalg.norm(g - H @ x1) / g0; t2 = np.linalg.norm(g - H @ x2) / g0
    print(f"tol {tol:g}: PCG {i1} its true rel res {t1:.2e}; pipelined {i2} its true rel res {t2:.2e}; |x1-x2|/|x1| {np.linalg.norm(x1-x2)/np.linalg.norm(x1):.2e}")
