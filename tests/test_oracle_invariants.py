"""Analytic invariants that pin the oracle where the reference has no tests. CPU only."""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

from oracle.session import OracleSession


def sh_basis(d):
    """3DGS real SH basis (degree 3) at unit direction d — the basis the rasteriser evaluates."""
    x, y, z = d
    C0, C1 = 0.28209479177387814, 0.4886025119029199
    C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
    C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658,
          1.445305721320277, -0.5900435899266435]
    xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
    return np.array([C0, -C1 * y, C1 * z, -C1 * x, C2[0] * xy, C2[1] * yz, C2[2] * (2 * zz - xx - yy), C2[3] * xz,
                     C2[4] * (xx - yy), C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                     C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy), C3[5] * z * (xx - yy),
                     C3[6] * x * (xx - 3 * yy)])


def test_sh_rotation_rotates_the_radiance_function(orc):
    """f'(R d) == f(d): pins the rotation matrices (and the odd-index sign flips) against the 3DGS basis."""
    rng = np.random.default_rng(1)
    for seed in range(4):
        R = Rotation.random(random_state=seed).as_matrix().astype(np.float32)
        sh = rng.normal(size=(16, 3)).astype(np.float32)
        sh2 = orc.sh_rotate(R, sh.reshape(-1)).reshape(16, 3)
        for _ in range(10):
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            assert np.allclose(sh_basis(d) @ sh, sh_basis(R.astype(np.float64) @ d) @ sh2, atol=3e-6)


def test_sh_rotation_inverse_and_composition(orc):
    rng = np.random.default_rng(2)
    R1 = Rotation.random(random_state=5).as_matrix().astype(np.float32)
    R2 = Rotation.random(random_state=6).as_matrix().astype(np.float32)
    sh = rng.normal(size=48).astype(np.float32)
    assert np.allclose(orc.sh_rotate(R1.T.copy(), orc.sh_rotate(R1, sh)), sh, atol=2e-6)
    assert np.allclose(orc.sh_rotate(R2, orc.sh_rotate(R1, sh)), orc.sh_rotate((R2 @ R1).astype(np.float32), sh), atol=3e-6)
    b1, b2, b3 = orc.sh_matrices(R1)
    for b in (b1, b2, b3):
        assert np.allclose(b @ b.T, np.eye(len(b)), atol=2e-6)


def test_polar_matches_svd(orc):
    rng = np.random.default_rng(3)
    for _ in range(20):
        M = rng.normal(size=(3, 3)) * np.array([1.0, 0.1, 0.01])
        U, s, Vt = np.linalg.svd(M)
        R, S = orc.polar(M)
        assert np.allclose(R, U @ Vt, atol=1e-9) and np.allclose(S, Vt.T @ np.diag(s) @ Vt, atol=1e-12)


def test_knn_selection_sort_semantics(orc):
    """Without ties the partial selection sort is a plain ascending order; weights follow (1-d/dmax)^2 normalised."""
    rng = np.random.default_rng(4)
    nodes = rng.normal(size=(300, 3)).astype(np.float32)
    q = rng.normal(size=(500, 3)).astype(np.float32)
    k = 10
    idx, w = orc.knn_weights(nodes, q, k)
    t = q[:, None, :] - nodes[None]
    d = np.sqrt(((t[..., 0] * t[..., 0] + t[..., 1] * t[..., 1]) + t[..., 2] * t[..., 2]).astype(np.float32)).astype(np.float32)
    order = np.argsort(d, axis=1, kind="stable")[:, :k + 1]
    ties = np.array([len(np.unique(np.sort(d[i])[:k + 2])) < k + 2 for i in range(len(q))])
    assert np.array_equal(idx[~ties], order[~ties].astype(np.uint32))
    dd = np.take_along_axis(d, idx.astype(np.int64), 1).astype(np.float64)
    u = (1.0 - dd[:, :k] / dd[:, k:k + 1]) ** 2
    assert np.allclose(w, u / u.sum(1, keepdims=True), rtol=1e-15, atol=0)
    assert np.allclose(w.sum(1), 1.0)


def test_knn_tie_break_is_position_based(orc):
    """position i keeps a tie against later positions; among later positions the highest wins."""
    nodes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [3, 3, 3]], np.float32)
    idx, _ = orc.knn_weights(nodes, np.zeros((1, 3), np.float32), 3, weights=False)
    # step 0: position 0 ties with everyone -> stays (node 0). step 1: position 1 holds d=1 -> node 1. ...
    assert idx[0].tolist() == [0, 1, 2, 3]
    nodes2 = np.array([[3, 3, 3], [0, 1, 0], [0, 0, 1], [-1, 0, 0], [0, -1, 0], [0, 0, -1], [1, 0, 0]], np.float32)
    idx2, _ = orc.knn_weights(nodes2, np.zeros((1, 3), np.float32), 3, weights=False)
    # step 0: d[0] is large; scanning from the end, the LAST position with the minimum (6) wins; node 0 moves to position 6.
    # step 1: position 1 holds the minimum -> node 1; etc.
    assert idx2[0].tolist() == [6, 1, 2, 3]


def _session(scenes, n=3000, nodes=60, k=6, **kw):
    sc = scenes.make_scene("sphere1m", n=n)
    o = OracleSession(sc, grid_num=16, knn_k=k, node_num=nodes, with_samples=kw.pop("with_samples", False), **kw)
    o.grid_build()
    o.graph_build_fps()
    return o


def test_identity_aims_give_identity_solve_and_noop_apply(scenes):
    o = _session(scenes)
    o.set_blocks([np.arange(o.M, dtype=np.uint32)], [1])
    before = {k: v.copy() for k, v in o.g.items()}
    st = o.step(False)
    assert st["iters"] == 1 and st["energy"] < 1e-20
    assert np.allclose(o.last_rot, np.tile(np.eye(3).reshape(-1), (o.M, 1)), atol=1e-12) and np.allclose(o.last_trans, 0, atol=1e-12)
    assert np.allclose(o.g["pos"], before["pos"], atol=2e-7)
    assert np.allclose(o.g["scale"], before["scale"], rtol=2e-4)   # (s+1e-3)*2 end-point round trip in float
    assert np.allclose(o.g["shs"], before["shs"], atol=2e-6)


def test_rigid_aims_give_rigid_motion(scenes):
    """Rigid aims on all nodes => every Gaussian follows the rigid motion; scales unchanged; SH rotated by R."""
    o = _session(scenes)
    o.set_blocks([np.arange(o.M, dtype=np.uint32)], [1])
    R = Rotation.from_rotvec([0.02, -0.03, 0.05]).as_matrix()
    t = np.array([0.01, -0.02, 0.005])
    before = {k: v.copy() for k, v in o.g.items()}
    o.aim_set((o.node_pos.astype(np.float64) @ R.T + t).astype(np.float32))
    st = o.step(False)
    assert st["energy"] < 1e-10
    assert np.allclose(o.g["pos"], before["pos"].astype(np.float64) @ R.T + t, atol=5e-7)
    assert np.allclose(o.g["scale"], before["scale"], rtol=5e-4)
    q0 = Rotation.from_quat(before["rot"][:, [1, 2, 3, 0]]); q1 = Rotation.from_quat(o.g["rot"][:, [1, 2, 3, 0]])
    assert np.allclose((q1 * q0.inv()).as_matrix(), R[None], atol=2e-4)
    expect = np.stack([__import__("oracle").sh_rotate(R.astype(np.float32), s) for s in before["shs"][:50]])
    assert np.allclose(o.g["shs"][:50], expect, atol=2e-4)


def test_solve_first_step_matches_scipy_direct_solve(orc, scenes):
    """(J^T J) h = -J^T f solved by scipy's sparse LU equals the oracle's block LDL^T step."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    o = _session(scenes, n=4000, nodes=150, k=8)
    blocks, types = scenes.cap_blocks(o.node_pos, lo=-0.3, hi=0.3)
    o.set_blocks(blocks, types)
    o.aim_translate([0.0, 0.0, 0.02])
    I = np.tile(np.eye(3).reshape(-1), (o.M, 1)); Z = np.zeros((o.M, 3))
    R, Cc, V, f, (m, n) = orc.jacobian(o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, False, I, Z)
    J = sp.coo_matrix((V, (R, Cc)), shape=(m, n)).tocsc()
    h = spl.spsolve((J.T @ J).tocsc(), -(J.T @ f))
    rot, trans, st = orc.solve(o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, False, max_iters=1)
    x1 = np.concatenate([rot, trans], 1).reshape(-1)
    x0 = np.concatenate([I, Z], 1).reshape(-1)
    assert np.allclose(x1 - x0, h, atol=1e-9)
    assert np.isclose(f @ f, orc.energy(o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, False, I, Z))


@pytest.mark.parametrize("on_center", [False, True])
def test_jacobian_matches_finite_differences_of_the_residual(orc, scenes, on_center):
    """The analytic Jacobian (Deform.cpp:180-376) is the derivative of the residual (Deform.cpp:378-581) at a generic x:
    central differences of f over every unknown of a few nodes, per-node and centre constraints, with an excluded block."""
    o = _session(scenes, n=3000, nodes=60, k=6)
    rng = np.random.default_rng(3)
    z = o.node_pos[:, 2]
    blocks = [np.nonzero(z > 0.25)[0].astype(np.uint32), np.nonzero(z < -0.25)[0].astype(np.uint32),
              np.nonzero(np.abs(z) < 0.05)[0].astype(np.uint32)[:3], np.nonzero(o.node_pos[:, 0] > 0.4)[0].astype(np.uint32)[:4]]
    o.set_blocks(blocks, [1, 0, 0, -1])
    o.aim_translate([0.01, 0.0, 0.03])
    rot = np.tile(np.eye(3).reshape(-1), (o.M, 1)) + rng.normal(size=(o.M, 9)) * 0.05
    trans = rng.normal(size=(o.M, 3)) * 0.02
    args = (o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, on_center)
    R, Cc, V, f0, (m, n) = orc.jacobian(*args, rot, trans)
    J = np.zeros((m, n)); np.add.at(J, (R, Cc), V)
    free = np.nonzero(o.node_static == 0)[0]
    assert n == 12 * len(free)
    x = np.concatenate([rot, trans], 1)
    eps = 1e-6
    for u in rng.choice(n, 40, replace=False):
        node, comp = free[u // 12], u % 12
        xp, xm = x.copy(), x.copy()
        xp[node, comp] += eps; xm[node, comp] -= eps
        fp = orc.jacobian(*args, xp[:, :9].copy(), xp[:, 9:].copy())[3]
        fm = orc.jacobian(*args, xm[:, :9].copy(), xm[:, 9:].copy())[3]
        assert np.allclose((fp - fm) / (2 * eps), J[:, u], atol=2e-7), (u, np.abs((fp - fm) / (2 * eps) - J[:, u]).max())


@pytest.mark.parametrize("on_center", [False, True])
def test_explicit_normal_equation_stencil_equals_JtJ(orc, scenes, on_center):
    """csrc/solve_pipe.cu applies H = J^T J as an explicit stencil (header of that file): per node i and component j the 4-vector
    y_ij = (A_j0, A_j1, A_j2, t_j) sees  w_reg^2 [C_i y_ij - sum_s c_is t_{q(s) j}] from its own E_reg rows, the mirror terms of
    its in-edges, w_reg^2 (static in-edges) on t, the node-local E_rot block, and sum_g r_g r_g^T from the constraint rows with
    r_g = w_con wei (v_c - g_q, 1) on y_qj.  Built here entry by entry from those formulas (numpy) and compared with J^T J of the
    oracle's Jacobian (Deform.cpp:180-376) at a generic x, per-node and centre constraints, with an excluded block."""
    o = _session(scenes, n=3000, nodes=60, k=6)
    rng = np.random.default_rng(5)
    z = o.node_pos[:, 2]
    blocks = [np.nonzero(z > 0.25)[0].astype(np.uint32), np.nonzero(z < -0.25)[0].astype(np.uint32),
              np.nonzero(np.abs(z) < 0.05)[0].astype(np.uint32)[:3], np.nonzero(o.node_pos[:, 0] > 0.4)[0].astype(np.uint32)[:4]]
    types = [1, 0, 0, -1]
    o.set_blocks(blocks, types)
    o.aim_translate([0.01, 0.0, 0.03])
    M, k = o.M, o.nbr.shape[1]
    rot = np.tile(np.eye(3).reshape(-1), (M, 1)) + rng.normal(size=(M, 9)) * 0.05      # column-major A (Deform.hpp:29-36)
    trans = rng.normal(size=(M, 3)) * 0.02
    R, Cc, V, f0, (m, n) = orc.jacobian(o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, on_center, rot, trans)
    J = np.zeros((m, n)); np.add.at(J, (R, Cc), V)
    want = J.T @ J
    free = o.node_static == 0
    rank = np.cumsum(free) - 1
    assert n == 12 * free.sum()
    w_rot2, w_reg2, w_con = 1.0, 10.0, 10.0
    g = o.node_pos.astype(np.float32)

    def y_idx(i, j):      # unknown indices of (A_j0, A_j1, A_j2, t_j): x[j + 3c] = A[j, c], x[9 + j] = t_j
        b = 12 * rank[i]
        return np.array([b + j, b + j + 3, b + j + 6, b + 9 + j])
    H = np.zeros((n, n))
    for i in range(M):
        for s in range(k):
            q = int(o.nbr[i, s])
            c = np.append((g[q] - g[i]).astype(np.float64), 1.0)          # float position differences (Deform.cpp:254-256)
            for j in range(3):
                if free[i]:
                    yi = y_idx(i, j)
                    H[np.ix_(yi, yi)] += w_reg2 * np.outer(c, c)            # C_i
                    if free[q]:
                        tq = y_idx(q, j)[3]
                        H[yi, tq] -= w_reg2 * c; H[tq, yi] -= w_reg2 * c     # own rows <-> the neighbour's t, and its mirror (in-edge partial)
                        H[tq, tq] += w_reg2                                  # in-degree term
                elif free[q]:
                    tq = y_idx(q, j)[3]
                    H[tq, tq] += w_reg2                                      # static in-edge
    for i in np.nonzero(free)[0]:
        A = rot[i].reshape(3, 3).T                                           # A[j, c]
        Jr = np.zeros((6, 9))                                                # columns: x[j + 3c]
        for r, (a, b) in enumerate(((0, 1), (0, 2), (1, 2))):
            for j in range(3):
                Jr[r, j + 3 * a] = A[j, b]; Jr[r, j + 3 * b] = A[j, a]
        for c in range(3):
            for j in range(3):
                Jr[3 + c, j + 3 * c] = 2.0 * A[j, c]
        b0 = 12 * rank[i]
        H[b0:b0 + 9, b0:b0 + 9] += w_rot2 * Jr.T @ Jr
    groups = []
    for blk, t in zip(blocks, types):
        if t == -1:
            continue
        groups += [sorted(set(int(v) for v in blk[:20]))] if on_center else [[int(v)] for v in blk]
    for members in groups:
        for j in range(3):
            r = np.zeros(n)
            for c_node in members:
                for s in range(k):
                    q = int(o.anc_idx[c_node, s])
                    if free[q]:
                        r[y_idx(q, j)] += w_con * o.anc_w[c_node, s] * np.append((g[c_node] - g[q]).astype(np.float64), 1.0)
            H += np.outer(r, r)
    scale = np.abs(want).max()
    assert np.abs(H - want).max() <= 1e-9 * scale, np.abs(H - want).max() / scale


def test_pipelined_pcg_recurrences_solve_the_gauss_newton_system(orc, scenes):
    """The recurrences of csrc/solve_pipe.cu (pipelined PCG, Ghysels & Vanroose 2014, with the diagonal preconditioner folded in:
    u = D^-1 r, q = D^-1 s, m = D^-1 w; dots taken before the product of the same iteration; warm start x0 = alpha h' with the
    exact line-search alpha) restated in numpy on the oracle's first Gauss-Newton system: same solution as a direct solve, the
    same iteration count as classical Jacobi-PCG (+-2), and the warm start from a nearby solution needs fewer products."""
    o = _session(scenes, n=4000, nodes=150, k=8)
    blocks, types = scenes.cap_blocks(o.node_pos, lo=-0.3, hi=0.3)
    o.set_blocks(blocks, types)
    o.aim_translate([0.0, 0.0, 0.02])
    I = np.tile(np.eye(3).reshape(-1), (o.M, 1)); Z = np.zeros((o.M, 3))
    R, Cc, V, f, (m, n) = orc.jacobian(o.node_pos, o.nbr, o.anc_idx, o.anc_w, o.node_static, o.blocks, o.block_types, o.aim, False, I, Z)
    J = np.zeros((m, n)); np.add.at(J, (R, Cc), V)
    H = J.T @ J; g = -(J.T @ f); di = 1.0 / np.diag(H)
    exact = np.linalg.solve(H, g)

    def classical(tol):
        x = np.zeros(n); r = g.copy(); z = r * di; p = z.copy(); rz = r @ z
        for it in range(1, 5000):
            q = H @ p; a = rz / (p @ q); x += a * p; r -= a * q
            if r @ r <= tol * tol * (g @ g): return it, x
            z = r * di; rzn = r @ z; p = z + (rzn / rz) * p; rz = rzn

    def pipelined(tol, warm=None):
        x = np.zeros(n); r = g.copy(); products = 0
        if warm is not None:                                   # x0 = alpha h', alpha = g.h' / h'.H h'
            q = H @ warm; products += 1
            a = (g @ warm) / (warm @ q) if warm @ q > 0 else 0.0
            x = a * warm; r = g - a * q
        w = H @ (r * di); products += 1
        z = np.zeros(n); s = np.zeros(n); p = np.zeros(n)
        gam_old = alpha_old = 1.0
        for it in range(5000):
            u = r * di
            gam, dlt, rr = r @ u, w @ u, r @ r                 # one reduction, before this iteration's product
            if rr <= tol * tol * (g @ g): return products, x
            nn = H @ (w * di); products += 1
            beta = gam / gam_old if it else 0.0
            alpha = gam / (dlt - beta * gam / alpha_old) if it else gam / dlt
            z = nn + beta * z; s = w + beta * s; p = u + beta * p
            x = x + alpha * p; r = r - alpha * s; w = w - alpha * z
            gam_old, alpha_old = gam, alpha
    it_c, x_c = classical(1e-8)
    it_p, x_p = pipelined(1e-8)
    assert abs(it_p - 1 - it_c) <= 2, (it_p, it_c)
    assert np.abs(x_p - exact).max() <= 1e-6 * np.abs(exact).max() and np.abs(x_p - x_c).max() <= 1e-6 * np.abs(exact).max()
    it_w, x_w = pipelined(1e-8, warm=exact * 1.02 + 1e-4 * np.abs(exact).max() * np.sin(np.arange(n)))
    assert it_w < 0.8 * it_p and np.abs(x_w - exact).max() <= 1e-6 * np.abs(exact).max(), (it_w, it_p)


def test_six_point_fit_recovers_a_rigid_motion_and_axis_stretch(orc, scenes):
    """End points moved by x -> Q x + t: the fit returns rot = Q rot, same scales, pos = Q pos + t; stretched along the
    Gaussian's own axes by (a, b, c): scale_i' = ((s_i + 1e-3) a_i - 1e-3 ... ) as GaussianView.cpp:3124-3132 defines it."""
    sc = scenes.make_scene("sphere1m", n=500)
    N = sc["n"]
    ends = orc.end_points(sc["pos"], sc["rot"], sc["scale"]).reshape(N, 6, 3).astype(np.float64)
    Q = Rotation.from_rotvec([0.3, -0.2, 0.5]).as_matrix(); t = np.array([0.1, -0.05, 0.2])
    moved = (ends @ Q.T + t).astype(np.float32).reshape(N, 18)
    out = {k: sc[k].copy() for k in ("pos", "rot", "scale", "shs")}
    orc.fit_gaussians(moved, sc["scale"], np.zeros(N, np.uint8), out["pos"], out["rot"], out["scale"], out["shs"])
    assert np.abs(out["pos"] - (sc["pos"].astype(np.float64) @ Q.T + t)).max() <= 2e-6
    assert (np.abs(out["scale"] - sc["scale"]) / sc["scale"]).max() <= 2e-3          # float end points: 1e-7 / (2 s) relative
    Rn = Rotation.from_quat(out["rot"][:, [1, 2, 3, 0]]).as_matrix()
    R0 = Rotation.from_quat((sc["rot"] / np.linalg.norm(sc["rot"], axis=1, keepdims=True))[:, [1, 2, 3, 0]]).as_matrix()
    assert np.abs(Rn - Q @ R0).max() <= 2e-4
    # stretch along the Gaussian's first axis by 1.5: K_0 scales by 1.5, the others stay
    c = ends.mean(1, keepdims=True)
    d = ends - c
    d[:, 0:2] *= 1.5
    out2 = {k: sc[k].copy() for k in ("pos", "rot", "scale", "shs")}
    orc.fit_gaussians((c + d).astype(np.float32).reshape(N, 18), sc["scale"], np.zeros(N, np.uint8), out2["pos"], out2["rot"], out2["scale"], out2["shs"])
    s0 = sc["scale"].astype(np.float64)
    want0 = 1.5 * (s0[:, 0] + 1e-3) * 2.0 / ((s0[:, 0] + 1e-3) * 2.0) * s0[:, 0]
    assert (np.abs(out2["scale"][:, 0] - want0) / want0).max() <= 2e-3
    assert (np.abs(out2["scale"][:, 1:] - sc["scale"][:, 1:]) / sc["scale"][:, 1:]).max() <= 2e-3


def test_lbs_and_sample_sh_are_noops_for_identity_transforms(orc):
    rng = np.random.default_rng(1)
    M, P, k = 50, 2000, 6
    nodes = rng.normal(size=(M, 3)).astype(np.float32)
    pts = rng.normal(size=(P, 3)).astype(np.float32)
    idx, w = orc.knn_weights(nodes, pts, k)
    I = np.tile(np.eye(3).reshape(-1), (M, 1)); Z = np.zeros((M, 3))
    out = orc.lbs_points(pts.copy(), idx[:, :k], w, nodes, I, Z)
    assert np.abs(out - pts).max() <= 3e-7                                             # float accumulator, k roundings
    q = orc.node_quats(I)
    assert np.allclose(q, [0, 0, 0, 1], atol=1e-7)
    feat = rng.normal(size=(P, 48)).astype(np.float32)
    rot = feat.copy()
    orc.rotate_sample_shs(w.astype(np.float32), idx[:, :k].astype(np.int32), q, np.zeros(P, np.int32), rot)
    assert np.abs(rot - feat).max() <= 1e-6


def test_excluded_nodes_keep_identity_and_static_flags(scenes):
    o = _session(scenes, with_samples=True)
    lo = np.nonzero(o.node_pos[:, 2] < -0.2)[0].astype(np.uint32)
    hi = np.nonzero(o.node_pos[:, 2] > 0.3)[0].astype(np.uint32)
    o.set_blocks([hi, lo], [1, -1])
    assert o.gs_static.sum() > 0 and o.sample_static.sum() > 0 and o.gs_static.sum() < o.N
    before = o.g["pos"].copy()
    o.aim_translate([0, 0, 0.01])
    o.step(False)
    assert np.allclose(o.last_rot[lo], np.eye(3).reshape(-1)) and np.allclose(o.last_trans[lo], 0)
    st = o.gs_static.astype(bool)
    assert np.array_equal(o.g["pos"][st], before[st]) and not np.array_equal(o.g["pos"][~st], before[~st])


def test_grid_invariants(scenes):
    o = _session(scenes, with_samples=True)
    G = o.G
    assert o.cell_prefix[-1] == o.N and np.all(np.diff(o.gs_init_grid_idx) >= 0)      # Gaussians are in cell order
    cnt = np.diff(np.concatenate([[0], o.fp_prefix]))
    assert np.array_equal(np.nonzero(cnt)[0], o.valid) and len(o.sample_pos) == 64 * len(o.valid)
    for c in o.valid[:50]:
        b, e = (o.fp_prefix[c - 1] if c else 0), o.fp_prefix[c]
        assert np.all(np.diff(o.lists[b:e]) > 0)                                       # ascending Gaussian index per cell
    # every Gaussian sits in its own cell's list (padding >= 0)
    for g in range(0, o.N, 97):
        c = o.gs_init_grid_idx[g]; b, e = (o.fp_prefix[c - 1] if c else 0), o.fp_prefix[c]
        assert g in o.lists[b:e]
    # samples of a cell lie inside the cell
    c = o.valid[0]; x, y, z = c // (G * G), (c // G) % G, c % G
    lo = o.aabb[:3] + np.array([x, y, z]) * o.gstep
    assert np.all(o.sample_pos[:64] > lo) and np.all(o.sample_pos[:64] < lo + o.gstep)
    # undeformed adaptive LPF = (step/4)^2 * 0.2 * I  (GaussianView.cpp:3881)
    assert np.allclose(o.ada_lpf[c].reshape(3, 3), np.eye(3) * (o.gstep / 4) ** 2 * 0.2, rtol=1e-3, atol=1e-12)


def test_grid_eval_is_the_lpf_widened_mixture(scenes):
    """Field evaluation (forward3d_grid is external): the oracle's value at a sample equals the definition
    feature(x) = sum_g alpha_g exp(-1/2 (x - mu_g)^T (Sigma_g + LPF_cell)^-1 (x - mu_g)) SH_g over the cell's list, with the
    1/255 footprint cut-off — recomputed here in float64 numpy for a few cells."""
    o = _session(scenes, n=4000, with_samples=True)
    o.grid_eval(0)
    g = o.g
    for ci in (0, len(o.valid) // 2, len(o.valid) - 1):
        c = o.valid[ci]
        b, e = (o.fp_prefix[c - 1] if c else 0), o.fp_prefix[c]
        lpf = o.ada_lpf[c].reshape(3, 3).astype(np.float64)
        x = o.sample_pos[ci * 64:(ci + 1) * 64].astype(np.float64)
        feat, opa = np.zeros((64, 48)), np.zeros(64)
        for gi in o.lists[b:e]:
            a = float(g["opacity"][gi])
            if a <= 1.0 / 255.0:
                continue
            q = g["rot"][gi].astype(np.float64); q /= np.linalg.norm(q)
            R = Rotation.from_quat(q[[1, 2, 3, 0]]).as_matrix()
            S = R @ np.diag(g["scale"][gi].astype(np.float64) ** 2) @ R.T + lpf
            d = x - g["pos"][gi].astype(np.float64)
            pw = -0.5 * np.einsum("si,ij,sj->s", d, np.linalg.inv(S), d)
            wgt = np.where((pw <= 0) & (pw >= np.log(1.0 / 255.0 / a) + 1e-4), a * np.exp(pw), 0.0)   # margin: float cut-off ties
            near_cut = np.abs(pw - np.log(1.0 / 255.0 / a)) < 1e-4
            wgt = np.where(near_cut, np.nan, wgt)
            opa += wgt; feat += wgt[:, None] * g["shs"][gi].astype(np.float64)
        ok = np.isfinite(opa)
        assert ok.sum() >= 48
        got_o, got_f = o.aim_opacity[ci * 64:(ci + 1) * 64], o.aim_feature[ci * 64:(ci + 1) * 64]
        assert np.allclose(got_o[ok], opa[ok], rtol=2e-4, atol=1e-6)
        assert np.allclose(got_f[ok], feat[ok], rtol=2e-4, atol=2e-5)
