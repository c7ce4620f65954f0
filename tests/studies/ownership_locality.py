#!/usr/bin/env python
"""CPU study (numpy / scipy, no GPU) for the solver's next step: how local the node graph becomes when a CTA owns a spatially compact
set of nodes instead of every 148th one, and what small block preconditioners buy.

    python tests/studies/ownership_locality.py [nodes=4000]

shells scene, k = 10, 148 CTAs.  4 000 nodes: 0.8 % of the kNN edges stay inside a CTA under round-robin ownership, 54 % under a balanced
recursive coordinate bisection; PCG iterations Jacobi 90, 4x4 (node, component) block-Jacobi 88, 12x12 node block-Jacobi 80.
16 000 nodes: 0.7 % / 71 %; 144 / 143 / 122."""
import importlib, sys
from pathlib import Path
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import __graft_entry__ as ge
ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
import oracle as O
M = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = 200000; k = 10
sc = scenes.make_scene("shells6m", n=N)
O.set_threads(8)
anchor = O.fps(sc["pos"], M)
nodes = sc["pos"][anchor].copy()
nbr = O.graph_edges(nodes, k)
idx, w = O.knn_weights(nodes, nodes, k)
# (1) locality of the graph under round-robin and under a balanced recursive coordinate bisection into B parts
B = 148
def rcb(ids, parts):
    if parts == 1: return [ids]
    a = parts // 2
    ax = np.argmax(nodes[ids].max(0) - nodes[ids].min(0))
    order = ids[np.argsort(nodes[ids, ax], kind="stable")]
    cut = len(ids) * a // parts
    return rcb(order[:cut], a) + rcb(order[cut:], parts - a)
parts = rcb(np.arange(M), B)
own_rcb = np.empty(M, int)
for b, p in enumerate(parts): own_rcb[p] = b
own_rr = np.arange(M) % B
for name, own in (("round-robin", own_rr), ("recursive coordinate bisection", own_rcb)):
    loc = (own[nbr] == own[:, None]).mean()
    print(f"{name}: {100*loc:.1f}% of the kNN edges stay inside the CTA; nodes per CTA {np.bincount(own).min()}..{np.bincount(own).max()}")
# (2) PCG iterations with 4x4 (node, component) block-Jacobi vs Jacobi vs 12x12
blocks, types = scenes.cap_blocks(nodes)
aim = nodes.copy(); aim[blocks[0]] += np.array([0, 0, 0.002], np.float32)
rot = np.tile(np.eye(3).reshape(9), (M, 1)); trans = np.zeros((M, 3))
R, C, V, f, (m, n) = O.jacobian(nodes, nbr, idx[:, :k], w, np.zeros(M, np.uint8), blocks, types, aim, False, rot, trans)
J = sp.csr_matrix((V, (R, C)), shape=(m, n)); H = (J.T @ J).tocsr(); g = -(J.T @ f)
def pcg(apply_M, tol=1e-6):
    x = np.zeros(n); r = g.copy(); z = apply_M(r); p = z.copy(); rz = r @ z; g0 = np.linalg.norm(g)
    for it in range(1, 5000):
        q = H @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= tol * g0: return it
        z = apply_M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
d = H.diagonal()
print("Jacobi", pcg(lambda r: r / d))
Hd = H.toarray() if n <= 20000 else None
# 4x4 blocks: unknown index u = 12 i + (j + 3c | 9 + j): rows (j): [j, j+3, j+6, 9+j]
Hc = H.tocoo()
def blk_inv(groups_of):
    key = groups_of(np.arange(n))
    sel = key[Hc.row] == key[Hc.col]
    Bm = sp.csr_matrix((Hc.data[sel], (Hc.row[sel], Hc.col[sel])), shape=(n, n)).tocsc()
    import scipy.sparse.linalg as spla
    lu = spla.splu(Bm + 1e-14 * sp.eye(n, format="csc"))
    return lambda r: lu.solve(r)
def row_group(u):
    i, q = u // 12, u % 12
    j = np.where(q < 9, q % 3, q - 9)
    return i * 3 + j
print("4x4 (node, component) block-Jacobi", pcg(blk_inv(row_group)))
print("12x12 node block-Jacobi", pcg(blk_inv(lambda u: u // 12)))
