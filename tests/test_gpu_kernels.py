"""Kernel-layer parity (-m gpu): CUDA kernels called through the C ABI vs the CPU oracle on the same seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

from gpu_util import KnnIndex, dev, ptr, stream

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------- stage (b)
@pytest.mark.parametrize("M,Q,k", [(300, 20000, 10), (4000, 60000, 8), (13, 500, 12), (2000, 5000, 1)])
def test_knn_indices_and_weights_bit_exact(pkg, orc, M, Q, k):
    rng = np.random.default_rng(M + k)
    nodes = (rng.normal(size=(M, 3)) * [0.5, 0.3, 0.2]).astype(np.float32)
    q = (rng.normal(size=(Q, 3)) * 0.6).astype(np.float32)
    q[:10] = nodes[:10]                                # zero distances
    q[10:20] = 50.0 + rng.normal(size=(10, 3))         # far outside the node box
    idx_g, w_g, _ = KnnIndex(pkg, nodes).query(q, k)
    idx_o, w_o = orc.knn_weights(nodes, q, k)
    assert np.array_equal(idx_g, idx_o)                # k+1 neighbour indices, bit-exact
    assert np.array_equal(w_g, w_o)                    # double weights, bit-exact


def test_knn_lattice_ties_bit_exact(pkg, orc, golden):
    """Regular lattices produce many exact distance ties: the selection-sort tie-break must be reproduced."""
    g = np.arange(6, dtype=np.float32) * 0.25
    nodes = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    rng = np.random.default_rng(0)
    nodes = nodes[rng.permutation(len(nodes))]
    h = np.arange(11, dtype=np.float32) * 0.125
    q = np.stack(np.meshgrid(h, h, h, indexing="ij"), -1).reshape(-1, 3)
    idx_g, w_g, nslow = KnnIndex(pkg, nodes).query(q, 10)
    idx_o, w_o = orc.knn_weights(nodes, q, 10)
    assert nslow > 100                                 # the exact slow path really ran
    assert np.array_equal(idx_g, idx_o) and np.array_equal(w_g, w_o)
    # the stripes graph (box lattice) against lattice samples
    pts = pkg.graph_obj_load(golden / "stripes_graph.obj")
    q2 = (np.stack(np.meshgrid(h, h * 2 - 1.25, h, indexing="ij"), -1).reshape(-1, 3) - [0.3, 0, 0.3]).astype(np.float32)
    idx_g, w_g, _ = KnnIndex(pkg, pts).query(q2, 10)
    idx_o, w_o = orc.knn_weights(pts, q2, 10)
    assert np.array_equal(idx_g, idx_o) and np.array_equal(w_g, w_o)


def test_fps_bit_exact(pkg, orc, golden):
    lib = pkg.lib()
    rng = np.random.default_rng(7)
    for pts, m in [(rng.normal(size=(50000, 3)).astype(np.float32), 700),
                   (pkg.graph_obj_load(golden / "stripes_graph.obj"), 200),
                   (rng.normal(size=(40, 3)).astype(np.float32), 100)]:
        d = dev(pts)
        out = torch.zeros(max(m, 1), dtype=torch.int32, device="cuda")
        scratch = torch.empty(len(pts) * 4 + 65536 + 512, dtype=torch.uint8, device="cuda")
        cnt = C.c_int()
        pkg.check(lib.arapk_fps(ptr(d), C.c_longlong(len(pts)), m, ptr(out), ptr(scratch), C.c_size_t(scratch.numel()), C.byref(cnt), stream()))
        torch.cuda.synchronize()
        ref = orc.fps(pts, m)
        assert cnt.value == len(ref) and np.array_equal(out.cpu().numpy()[:cnt.value], ref)


# ----------------------------------------------------------------------------- stage (d)
def _blocked(idx, w, k):
    """plain rows -> the library's 32-row blocked tables (uint16 idx, double w)."""
    P = len(idx)
    nb = (P + 31) // 32
    bi = np.zeros((nb, k, 32), np.uint16); bw = np.zeros((nb, k, 32), np.float64)
    ii = np.zeros((nb * 32, k), np.uint16); ww = np.zeros((nb * 32, k), np.float64)
    ii[:P], ww[:P] = idx[:, :k], w
    bi[:] = ii.reshape(nb, 32, k).transpose(0, 2, 1); bw[:] = ww.reshape(nb, 32, k).transpose(0, 2, 1)
    return bi, bw


def _random_transforms(rng, M, mag=0.05):
    R = Rotation.from_rotvec(rng.normal(size=(M, 3)) * mag).as_matrix()
    A = R + rng.normal(size=(M, 3, 3)) * mag * 0.1
    rot = np.ascontiguousarray(A.transpose(0, 2, 1).reshape(M, 9))   # column-major
    trans = rng.normal(size=(M, 3)) * mag * 0.2
    return rot, trans


@pytest.mark.parametrize("k", [8, 10, 12, 5])
def test_lbs_points_bit_exact(pkg, orc, k):
    """LBS emulates the reference's double-product / float-accumulate arithmetic: outputs must be bit-identical."""
    lib = pkg.lib()
    rng = np.random.default_rng(k)
    M, P = 500, 100003
    nodes = (rng.normal(size=(M, 3)) * 0.4).astype(np.float32)
    pts = (rng.normal(size=(P, 3)) * 0.4).astype(np.float32)
    idx, w = orc.knn_weights(nodes, pts, k)
    rot, trans = _random_transforms(rng, M)
    skip = (rng.uniform(size=P) < 0.1).astype(np.uint8)
    ref = orc.lbs_points(pts.copy(), idx[:, :k], w, nodes, rot, trans, skip=skip.astype(np.int32))
    bi, bw = _blocked(idx, w, k)
    xf = torch.empty(M * 112, dtype=torch.uint8, device="cuda")
    d_rot, d_trans, d_nodes, d_pts = dev(rot), dev(trans), dev(nodes), dev(pts)
    pkg.check(lib.arapk_node_xf(M, ptr(d_rot), ptr(d_trans), ptr(d_nodes), ptr(xf), None, stream()))
    d_bi, d_bw, d_skip = dev(bi), dev(bw), dev(skip)
    pkg.check(lib.arapk_lbs_points(ptr(d_pts), ptr(d_pts), C.c_longlong(P), k, ptr(d_bi), ptr(d_bw), ptr(xf), ptr(d_skip), 1, stream()))
    torch.cuda.synchronize()
    out = d_pts.cpu().numpy()
    mism = np.nonzero(out != ref)[0]
    # fused multiply-adds differ from the reference's separate mul/add only below 1e-16 relative; a float rounding
    # flip needs that to straddle a rounding boundary (probability ~1e-8 per op)
    assert len(mism) <= 2, (len(mism), np.abs(out - ref).max())
    assert np.abs(out - ref).max() <= 1.2e-7


@pytest.mark.parametrize("k", [10, 8, 12, 5, 1])
def test_lbs_tiles_bit_exact(pkg, orc, k):
    """The staged-record LBS kernel (per-tile distinct node lists, one-byte slots, optional FP64-pipe float rounding) must
    give the same bits as the global-gather kernel and the oracle, on tiles that fit the staging area and tiles that do not."""
    lib = pkg.lib()
    rng = np.random.default_rng(100 + k)
    M, P = 3000, 150011
    nodes = (rng.normal(size=(M, 3)) * 0.4).astype(np.float32)
    pts = (rng.normal(size=(P, 3)) * 0.4).astype(np.float32)
    half = P // 2                                                     # first half spatially coherent (few nodes per tile), rest random
    ctr = (rng.normal(size=(half // 125 + 1, 3)) * 0.4).astype(np.float32)
    pts[:half] = (np.repeat(ctr, 125, axis=0)[:half] + rng.normal(size=(half, 3)) * 0.01).astype(np.float32)
    pts[7] = 0.0                                                      # exact zeros through the rounding trick
    idx, w = orc.knn_weights(nodes, pts, k)
    rot, trans = _random_transforms(rng, M)
    rot[3] = np.eye(3).reshape(9); trans[3] = 0.0
    skip = (rng.uniform(size=P) < 0.1).astype(np.uint8)
    ref = orc.lbs_points(pts.copy(), idx[:, :k], w, nodes, rot, trans, skip=skip.astype(np.int32))
    bi, bw = _blocked(idx, w, k)
    xf = torch.empty(M * 112, dtype=torch.uint8, device="cuda")
    d_rot, d_trans, d_nodes = dev(rot), dev(trans), dev(nodes)
    pkg.check(lib.arapk_node_xf(M, ptr(d_rot), ptr(d_trans), ptr(d_nodes), ptr(xf), None, stream()))
    d_bi, d_bw, d_skip = dev(bi), dev(bw), dev(skip)
    nt, cap = lib.arapk_lbs_tile_count(P), lib.arapk_lbs_tile_cap()
    d_slots = torch.zeros(((P + 31) // 32) * 32 * 3, dtype=torch.int32, device="cuda")
    d_cnt = torch.zeros(nt, dtype=torch.int16, device="cuda")
    d_tn = torch.zeros(nt * cap, dtype=torch.int16, device="cuda")
    pkg.check(lib.arapk_lbs_build_tiles(C.c_longlong(P), k, ptr(d_bi), ptr(d_slots), ptr(d_cnt), ptr(d_tn), stream()))
    torch.cuda.synchronize()
    cnt = d_cnt.cpu().numpy().astype(np.int64)
    if k > 1:
        assert (cnt > 0).sum() > nt // 4 and (cnt == 0).sum() > nt // 4, ((cnt > 0).sum(), nt)   # both paths exercised
    # distinct lists are what the rows reference
    t0 = int(np.nonzero(cnt > 0)[0][0])
    staged = d_tn.cpu().numpy().view(np.uint16)[t0 * cap:t0 * cap + cnt[t0]]
    assert np.array_equal(np.unique(staged), np.unique(idx[t0 * 128:(t0 + 1) * 128, :k]))   # unused slots repeat a node of the tile
    sl = d_slots.cpu().numpy().view(np.uint8).reshape(-1, 3, 32, 4)                       # [block][word][lane][byte]
    for r in range(t0 * 128, min((t0 + 1) * 128, P), 17):
        got = [staged[sl[r // 32, j // 4, r % 32, j % 4]] for j in range(k)]
        assert np.array_equal(got, idx[r, :k])                                            # slot -> node is the row's neighbour list
    d_a = dev(pts)
    pkg.check(lib.arapk_lbs_points(ptr(d_a), ptr(d_a), C.c_longlong(P), k, ptr(d_bi), ptr(d_bw), ptr(xf), ptr(d_skip), 1, stream()))
    outs = [d_a.cpu().numpy()]
    for magic in (0, 1):
        d_b = dev(pts)
        pkg.check(lib.arapk_lbs_tiles(ptr(d_b), ptr(d_b), C.c_longlong(P), k, ptr(d_slots), ptr(d_bw), ptr(d_bi), ptr(d_cnt), ptr(d_tn),
                                      ptr(xf), ptr(d_skip), 1, magic, stream()))
        torch.cuda.synchronize()
        outs.append(d_b.cpu().numpy())
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])     # three kernels, same bits
    mism = np.nonzero(outs[2] != ref)[0]
    assert len(mism) <= 2, (len(mism), np.abs(outs[2] - ref).max())
    assert np.abs(outs[2] - ref).max() <= 1.2e-7


def test_end_points_and_fit_match_oracle(pkg, orc, scenes):
    lib = pkg.lib()
    sc = scenes.make_scene("sphere1m", n=50001)
    N = sc["n"]
    ends_ref = orc.end_points(sc["pos"], sc["rot"], sc["scale"])
    d = {k: dev(sc[k]) for k in ("pos", "rot", "scale", "opacity", "shs")}
    d_ends = torch.zeros(N * 18, dtype=torch.float32, device="cuda")
    pkg.check(lib.arapk_end_points(C.c_longlong(N), ptr(d["pos"]), ptr(d["rot"]), ptr(d["scale"]), ptr(d_ends), stream()))
    torch.cuda.synchronize()
    assert np.array_equal(d_ends.cpu().numpy().reshape(N, 6, 3), ends_ref)             # float-only arithmetic: bit-exact
    # deform the endpoints with a smooth affine field, then fit
    rng = np.random.default_rng(3)
    A = np.eye(3) + rng.normal(size=(3, 3)) * 0.05
    ends = (ends_ref.reshape(-1, 3).astype(np.float64) @ A.T + [0.01, 0.02, -0.01]).astype(np.float32)
    ends += rng.normal(size=ends.shape).astype(np.float32) * 1e-4
    static = (rng.uniform(size=N) < 0.2).astype(np.uint8)
    ref = {k: sc[k].copy() for k in ("pos", "rot", "scale", "shs")}
    orc.fit_gaussians(ends.reshape(N, 18), sc["scale"], static, ref["pos"], ref["rot"], ref["scale"], ref["shs"])
    d_ends = dev(ends); d_sb = dev(sc["scale"]); d_st = dev(static)
    pkg.check(lib.arapk_fit_gaussians(C.c_longlong(N), ptr(d_ends), ptr(d_sb), ptr(d_st), ptr(d["pos"]), ptr(d["rot"]), ptr(d["scale"]), ptr(d["shs"]), stream()))
    torch.cuda.synchronize()
    out = {k: d[k].cpu().numpy() for k in ref}
    st = static.astype(bool)
    for k in ref:                                                                       # static Gaussians untouched
        assert np.array_equal(out[k][st], sc[k][st])
    assert np.array_equal(out["pos"], ref["pos"])                                       # centre: float-only, bit-exact
    # north_star tolerance: <= 1e-5 relative on deformed means / covariances
    assert (np.abs(out["scale"] - ref["scale"]) / ref["scale"]).max() <= 1e-5
    assert np.abs(out["rot"] - ref["rot"]).max() <= 2e-6
    cov = lambda q, s: (lambda R: R * (s ** 2)[:, None, :] @ R.transpose(0, 2, 1))(Rotation.from_quat(q[:, [1, 2, 3, 0]]).as_matrix())
    Cg, Co = cov(out["rot"][~st], out["scale"][~st]), cov(ref["rot"][~st], ref["scale"][~st])
    rel = np.linalg.norm(Cg - Co, axis=(1, 2)) / np.linalg.norm(Co, axis=(1, 2))
    assert rel.max() <= 1e-5, rel.max()
    assert np.abs(out["shs"] - ref["shs"]).max() <= 2e-6


def test_replay_shs_repeats_the_fit_bit_for_bit(pkg, scenes):
    """Multi-GPU receivers rebuild remote SH rows from (old rotation, new rotation): same bits as the owner's fit."""
    lib = pkg.lib()
    sc = scenes.make_scene("sphere1m", n=40007)
    N = sc["n"]
    rng = np.random.default_rng(9)
    d = {k: dev(sc[k]) for k in ("pos", "rot", "scale", "shs")}
    d_ends = torch.zeros(N * 18, dtype=torch.float32, device="cuda")
    pkg.check(lib.arapk_end_points(C.c_longlong(N), ptr(d["pos"]), ptr(d["rot"]), ptr(d["scale"]), ptr(d_ends), stream()))
    static = dev((rng.uniform(size=N) < 0.2).astype(np.uint8))
    held = {"rot": d["rot"].clone(), "shs": d["shs"].clone()}                      # what a receiver holds
    d_sb = dev(sc["scale"])
    for step in range(3):
        A = np.eye(3) + rng.normal(size=(3, 3)) * 0.05
        ends = (d_ends.cpu().numpy().reshape(-1, 3).astype(np.float64) @ A.T).astype(np.float32)
        d_ends = dev(ends.reshape(-1))
        pkg.check(lib.arapk_fit_gaussians(C.c_longlong(N), ptr(d_ends), ptr(d_sb), ptr(static), ptr(d["pos"]), ptr(d["rot"]), ptr(d["scale"]), ptr(d["shs"]), stream()))
        rot_prev = held["rot"].clone()
        held["rot"].copy_(d["rot"])                                                # the 16 bytes that travel
        pkg.check(lib.arapk_replay_shs(C.c_longlong(N), ptr(rot_prev), ptr(held["rot"]), ptr(static), ptr(held["shs"]), stream()))
        torch.cuda.synchronize()
        assert torch.equal(held["shs"], d["shs"]), step
    assert not torch.equal(d["shs"], dev(sc["shs"]))                               # the rows did change


def test_sh_rotation_device_matches_oracle(pkg, orc):
    lib = pkg.lib()
    rng = np.random.default_rng(11)
    for seed in range(5):
        R = Rotation.random(random_state=seed).as_matrix().astype(np.float32)
        sh = rng.normal(size=48).astype(np.float32)
        d_R, d_sh = dev(R), dev(sh)
        pkg.check(lib.arapk_sh_rotate_test(ptr(d_R), ptr(d_sh), 0, stream()))
        torch.cuda.synchronize()
        assert np.array_equal(d_sh.cpu().numpy(), orc.sh_rotate(R, sh))                 # reference rounding (double coefficients): bit-exact
        d_sh = dev(sh)
        pkg.check(lib.arapk_sh_rotate_test(ptr(d_R), ptr(d_sh), 1, stream()))
        torch.cuda.synchronize()
        assert np.abs(d_sh.cpu().numpy() - orc.sh_rotate(R, sh)).max() <= 1e-6          # float production version: <= 2 ulp per matrix entry


def test_node_quats_and_sample_sh_rotation(pkg, orc):
    lib = pkg.lib()
    rng = np.random.default_rng(5)
    M, S, k = 300, 20000, 10
    rot, _ = _random_transforms(rng, M, 0.1)
    q_ref = orc.node_quats(rot)
    d_rot = dev(rot); d_q = torch.zeros(M * 4, dtype=torch.float32, device="cuda")
    pkg.check(lib.arapk_node_quats(M, ptr(d_rot), ptr(d_q), stream()))
    torch.cuda.synchronize()
    q_g = d_q.cpu().numpy().reshape(M, 4)
    assert np.abs(q_g - q_ref).max() <= 1e-6
    nodes = (rng.normal(size=(M, 3)) * 0.4).astype(np.float32)
    smp = (rng.normal(size=(S, 3)) * 0.4).astype(np.float32)
    idx, w = orc.knn_weights(nodes, smp, k)
    feat = rng.normal(size=(S, 48)).astype(np.float32)
    static = (rng.uniform(size=S) < 0.3).astype(np.uint8)
    ref = feat.copy()
    orc.rotate_sample_shs(w.astype(np.float32), idx[:, :k].astype(np.int32), q_ref, static.astype(np.int32), ref)
    bi, bw = _blocked(idx, w, k)
    d_feat, d_bi, d_wf, d_st, d_qr = dev(feat), dev(bi), dev(bw.astype(np.float32)), dev(static), dev(q_ref)
    pkg.check(lib.arapk_rotate_sample_shs(C.c_longlong(S), k, ptr(d_wf), ptr(d_bi), ptr(d_qr), ptr(d_st), ptr(d_feat), stream()))
    torch.cuda.synchronize()
    out = d_feat.cpu().numpy()
    assert np.array_equal(out[static.astype(bool)], feat[static.astype(bool)])
    assert np.abs(out - ref).max() <= 5e-6                                              # double sin/atan2: libm vs CUDA, ~1e-7


def test_static_flags(pkg, orc):
    lib = pkg.lib()
    rng = np.random.default_rng(9)
    M, P, k = 200, 6 * 5000, 8
    idx = rng.integers(0, M, size=(P, k)).astype(np.uint32)
    ns = (rng.uniform(size=M) < 0.9).astype(np.uint8)
    bi, _ = _blocked(idx, np.zeros((P, k)), k)
    for group in (1, 6):
        out = torch.zeros(P // group, dtype=torch.uint8, device="cuda")
        d_bi, d_ns = dev(bi), dev(ns)
        pkg.check(lib.arapk_static_flags(C.c_longlong(P // group), group, k, ptr(d_bi), ptr(d_ns), ptr(out), stream()))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), orc.static_flags(idx, group, ns))


def test_selectable_kernel_variants_keep_parity():
    """Variants chosen by environment variables (read once per process): the TMA version of the sample SH rotation (tensor-map
    loads / stores, 64-byte swizzle) and the per-query kNN walk of the first round must pass the same oracle comparisons."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for env, sel in ((dict(ARAP_ROT_TMA="1"), "test_node_quats_and_sample_sh_rotation or test_drag_steps_match_oracle"),
                     (dict(ARAP_KNN_TILE="0"), "test_knn_indices_and_weights_bit_exact or test_knn_lattice_ties_bit_exact")):
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_kernels.py"), os.path.join(here, "test_gpu_session.py"),
                            "-m", "gpu", "-x", "-q", "-k", sel], env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, (env, r.stdout[-2000:] + r.stderr[-2000:])


def test_one_barrier_solver_eligibility_check(pkg):
    """arapk_solve_pipe_eligible (host side, csrc/solve_pipe.cu): the one-barrier kernel is selected only when every CTA's slice of
    constraint entries fits its shared-memory table, there are at most 8 multi-member groups and none has more than 20 members;
    the session asks it at the first solve after arap_set_blocks and otherwise runs the two-barrier kernel."""
    lib = pkg.lib()
    M, k = 3000, 10

    def ask(grp_sizes, per_node, max_ctas=0):
        grp_off = np.concatenate([[0], np.cumsum(grp_sizes)]).astype(np.int32)
        cin_off = np.concatenate([[0], np.cumsum(per_node)]).astype(np.int32)
        return lib.arapk_solve_pipe_eligible(M, k, len(grp_sizes), grp_off.ctypes.data_as(C.c_void_p), cin_off.ctypes.data_as(C.c_void_p), max_ctas)
    ones = np.ones(500, np.int64)
    assert ask(ones, np.full(M, 2)) == 1                      # 125 CTAs x 24 nodes x 2 entries
    assert ask(ones, np.full(M, 50)) == 0                     # 1200 entries per CTA > 1024
    assert ask(ones, np.full(M, 6), max_ctas=24) == 0         # 125 nodes per CTA (the NL = 136 slice: 672 entries) x 6 = 750
    assert ask(ones, np.full(M, 5), max_ctas=24) == 1
    assert ask(ones, np.full(M, 2), max_ctas=18) == 0         # 167 nodes per CTA: no slice of this kernel
    assert ask(np.array([20, 20, 20, 1, 1]), np.full(M, 2)) == 1
    assert ask(np.array([21, 1]), np.full(M, 2)) == 0         # more members than a centre-constraint group can have (DC:4)
    assert ask(np.full(9, 2), np.full(M, 2)) == 0             # nine multi-member groups
