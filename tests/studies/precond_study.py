#!/usr/bin/env python
"""CPU study (scipy, no GPU): PCG iteration counts of the first Gauss-Newton system of a drag step under different
preconditioners, on the oracle's Jacobian (reference arithmetic) of a scene with the bench's node density.

    python tests/studies/precond_study.py [nodes=4000] [gaussians=200000]

Unknown order of the oracle's Jacobian = the reference's: 12 per free node (A column-major 9, t 3)."""
import importlib, sys, time
from pathlib import Path
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge
ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
import oracle as O

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200000
k = 10
sc = scenes.make_scene("sphere1m", n=N)
O.set_threads(8)
anchor = O.fps(sc["pos"], M)
nodes = sc["pos"][anchor].copy()
nbr = O.graph_edges(nodes, k)
idx, w = O.knn_weights(nodes, nodes, k)
blocks, types = scenes.cap_blocks(nodes)
aim = nodes.copy(); aim[blocks[0]] += np.array([0, 0, 0.002], np.float32)
rot = np.tile(np.eye(3).reshape(9), (M, 1)); trans = np.zeros((M, 3))
R, C, V, f, (m, n) = O.jacobian(nodes, nbr, idx[:, :k], w, np.zeros(M, np.uint8), blocks, types, aim, False, rot, trans)
J = sp.csr_matrix((V, (R, C)), shape=(m, n))
H = (J.T @ J).tocsr()
g = -(J.T @ f)
print(f"M={M} unknowns={n} rows={m} nnz(H)={H.nnz} active={len(blocks[0])} pinned={len(blocks[1])}")

def pcg(apply_M, tol=1e-6, maxit=5000):
    x = np.zeros(n); r = g.copy(); z = apply_M(r); p = z.copy(); rz = r @ z; g0 = np.linalg.norm(g)
    for it in range(1, maxit + 1):
        q = H @ p; a = rz / (p @ q); x += a * p; r -= a * q
        if np.linalg.norm(r) <= tol * g0: return it
        z = apply_M(r); rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn
    return maxit

d = H.diagonal()
res = {}
res["Jacobi (current)"] = pcg(lambda r: r / d)

def block_jacobi(bs):
    nb = n // bs
    Hb = np.zeros((nb, bs, bs))
    Hc = H.tocoo()
    sel = (Hc.row // bs) == (Hc.col // bs)
    Hb[Hc.row[sel] // bs, Hc.row[sel] % bs, Hc.col[sel] % bs] = Hc.data[sel]
    inv = np.linalg.inv(Hb)
    return lambda r: np.einsum("bij,bj->bi", inv, r.reshape(nb, bs)).reshape(-1)
res["block-Jacobi 12x12 (node)"] = pcg(block_jacobi(12))

# two-level additive: Jacobi + coarse correction on piecewise-constant aggregates (12 dofs per aggregate)
def two_level(n_agg, smoother):
    from scipy.cluster.vq import kmeans2
    _, lab = kmeans2(nodes.astype(np.float64), n_agg, minit="++", seed=1)
    free = np.arange(M)   # all nodes free in this scene
    rows = np.arange(n); cols = lab[rows // 12] * 12 + rows % 12
    P = sp.csr_matrix((np.ones(n), (rows, cols)), shape=(n, n_agg * 12))
    Hc = (P.T @ H @ P).tocsc() + 1e-12 * sp.eye(n_agg * 12)
    lu = spla.splu(Hc)
    return lambda r: smoother(r) + P @ lu.solve(P.T @ r)
for na in (16, 64, 148, 256):
    res[f"two-level additive: Jacobi + {na} aggregates x 12"] = pcg(two_level(na, lambda r: r / d))
res["two-level additive: block-Jacobi 12 + 64 aggregates"] = pcg(two_level(64, block_jacobi(12)))

# aggregates a kernel can get for free: equal chunks of the nodes in Morton order (= spatially clustered CTA ownership)
def morton_chunks(n_agg):
    q = ((nodes - nodes.min(0)) / (nodes.max(0) - nodes.min(0) + 1e-9) * 1023).astype(np.int64)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF; v = (v | (v << 8)) & 0x0300F00F; v = (v | (v << 4)) & 0x030C30C3; return (v | (v << 2)) & 0x09249249
    key = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    lab = np.empty(M, np.int64); lab[np.argsort(key, kind="stable")] = np.arange(M) * n_agg // M
    return lab
def two_level_lab(lab, smoother, n_agg):
    rows = np.arange(n); cols = lab[rows // 12] * 12 + rows % 12
    P = sp.csr_matrix((np.ones(n), (rows, cols)), shape=(n, n_agg * 12))
    lu = spla.splu((P.T @ H @ P).tocsc() + 1e-12 * sp.eye(n_agg * 12))
    return lambda r: smoother(r) + P @ lu.solve(P.T @ r)
def rcb(n_parts):   # recursive coordinate bisection with proportional splits: balanced, compact boxes (host set-up cost: O(M log M))
    lab = np.zeros(M, np.int64)
    def rec(ids, p0, p):
        if p == 1:
            lab[ids] = p0; return
        pl = p // 2
        ax = np.argmax(nodes[ids].max(0) - nodes[ids].min(0))
        order = ids[np.argsort(nodes[ids, ax], kind="stable")]
        cut = len(ids) * pl // p
        rec(order[:cut], p0, pl); rec(order[cut:], p0 + pl, p - pl)
    rec(np.arange(M), 0, n_parts)
    return lab
res["two-level additive: Jacobi + 148 RCB boxes x 12 (balanced)"] = pcg(two_level_lab(rcb(148), lambda r: r / d, 148))
res["two-level additive: block-Jacobi 12 + 148 RCB boxes"] = pcg(two_level_lab(rcb(148), block_jacobi(12), 148))
for na in (148, 296):
    res[f"two-level additive: Jacobi + {na} Morton chunks x 12"] = pcg(two_level_lab(morton_chunks(na), lambda r: r / d, na))
res["two-level additive: block-Jacobi 12 + 148 Morton chunks"] = pcg(two_level_lab(morton_chunks(148), block_jacobi(12), 148))

# Chebyshev-accelerated Jacobi (degree 3) as preconditioner: 3 extra mat-vecs, no reductions
lmax = spla.eigsh(sp.diags(1 / np.sqrt(d)) @ H @ sp.diags(1 / np.sqrt(d)), k=1, which="LA", return_eigenvectors=False)[0]
def cheb(deg, lo_frac=0.06):
    lo, hi = lo_frac * lmax, 1.05 * lmax
    th, de = (hi + lo) / 2, (hi - lo) / 2
    def ap(r):
        x = np.zeros(n); rr = r.copy(); sig = th / de; rho = 1 / sig
        dvec = (rr / d) / th
        for _ in range(deg):
            x += dvec; rr = r - H @ x
            rho_n = 1 / (2 * sig - rho)
            dvec = rho_n * rho * dvec + 2 * rho_n / de * (rr / d); rho = rho_n
        return x
    return ap
for deg in (2, 4):
    it = pcg(cheb(deg)); res[f"Chebyshev({deg}) of Jacobi  [mat-vecs = {deg + 1} x its]"] = f"{it}  ({it * (deg + 1)} mat-vecs)"

try:
    ilu = spla.spilu(H.tocsc(), drop_tol=1e-4, fill_factor=4)
    res["ILU(drop 1e-4) (reference point, not parallel)"] = pcg(ilu.solve)
except Exception as e:  # noqa
    res["ILU"] = f"failed: {e}"
for kname, v in res.items():
    print(f"{kname:60s} {v}")
