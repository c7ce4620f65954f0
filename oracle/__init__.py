"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path).

ctypes bindings over ``oracle/arap_oracle.cpp``, the CPU restatement of the
reference's ARAP deformation path (see that file's header for the parity
status: *parity unpinned* except for the reference's fixtures).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libarap_oracle.so"
_SRC = _HERE / "arap_oracle.cpp"

# The reference adds neither -march=native nor -ffast-math (CMakeLists.txt:91).
CXXFLAGS = ["-O3", "-fopenmp", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC",
            "-Wall", "-Wno-unused-function", "-Wno-array-bounds"]


def build(force: bool = False) -> Path:
    """Compile the oracle in-tree (oracle/_build/, git-ignored via *.so)."""
    if not force and _SO.exists() and _SO.stat().st_mtime >= _SRC.stat().st_mtime:
        return _SO
    _SO.parent.mkdir(exist_ok=True)
    subprocess.check_call(["g++", *CXXFLAGS, "-o", str(_SO), str(_SRC)])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _SO.exists():
            build()
        _lib = C.CDLL(str(_SO))
        _lib.orc_grid_step.restype = C.c_float
        _lib.orc_footprint_count.restype = C.c_longlong
        _lib.orc_energy.restype = C.c_double
        _lib.orc_jacobian.restype = C.c_longlong
    return _lib


def _p(a, dtype):
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, dtype)
    return a.ctypes.data_as(C.c_void_p)


f32, f64, u32, i32, u8 = np.float32, np.float64, np.uint32, np.int32, np.uint8


def num_threads() -> int:
    return lib().orc_num_threads()


def set_threads(n: int) -> None:
    lib().orc_set_threads(int(n))


# --------------------------------------------------------------------------- graph
def fps(pos: np.ndarray, node_num: int) -> np.ndarray:
    pos = np.ascontiguousarray(pos, f32)
    out = np.zeros(min(node_num, len(pos)), i32)
    n = lib().orc_fps(_p(pos, f32), len(pos), int(node_num), _p(out, i32))
    return out[:n]


def knn_weights(nodes: np.ndarray, queries: np.ndarray, k: int, weights: bool = True):
    """-> idx [Q,k+1] uint32, w [Q,k] float64 (None if weights=False)."""
    nodes = np.ascontiguousarray(nodes, f32)
    queries = np.ascontiguousarray(queries, f32).reshape(-1, 3)
    Q = len(queries)
    idx = np.zeros((Q, k + 1), u32)
    w = np.zeros((Q, k), f64) if weights else None
    lib().orc_knn_weights(_p(nodes, f32), len(nodes), _p(queries, f32), C.c_longlong(Q), int(k), _p(idx, u32), _p(w, f64))
    return idx, w


def graph_edges(nodes: np.ndarray, k: int) -> np.ndarray:
    nodes = np.ascontiguousarray(nodes, f32)
    out = np.zeros((len(nodes), k), u32)
    lib().orc_graph_edges(_p(nodes, f32), len(nodes), int(k), _p(out, u32))
    return out


# --------------------------------------------------------------------------- apply
def lbs_points(points, idx, w, node_pos, rot, trans, skip=None):
    """In-place LBS of points [P,3] with rows idx[P,k], w[P,k]."""
    assert points.dtype == f32 and points.flags["C_CONTIGUOUS"]
    P, k = idx.shape
    idx = np.ascontiguousarray(idx, u32)
    w = np.ascontiguousarray(w, f64)
    skip = None if skip is None else np.ascontiguousarray(skip, i32)
    lib().orc_lbs_points(_p(points, f32), C.c_longlong(P), int(k), _p(idx, u32), _p(w, f64),
                         _p(np.ascontiguousarray(node_pos, f32), f32), _p(np.ascontiguousarray(rot, f64), f64),
                         _p(np.ascontiguousarray(trans, f64), f64), _p(skip, i32))
    return points


def end_points(pos, rot, scale):
    N = len(pos)
    ends = np.zeros((N, 6, 3), f32)
    lib().orc_end_points(_p(np.ascontiguousarray(pos, f32), f32), _p(np.ascontiguousarray(rot, f32), f32),
                         _p(np.ascontiguousarray(scale, f32), f32), C.c_longlong(N), _p(ends, f32))
    return ends


def fit_gaussians(ends, scale_backup, is_static, pos, rot, scale, shs):
    """In-place six-point fit; pos/rot/scale/shs are float32 C arrays."""
    N = len(pos)
    st = None if is_static is None else np.ascontiguousarray(is_static, u8)
    for a in (pos, rot, scale, shs):
        assert a.dtype == f32 and a.flags["C_CONTIGUOUS"]
    lib().orc_fit_gaussians(_p(np.ascontiguousarray(ends, f32), f32), _p(np.ascontiguousarray(scale_backup, f32), f32),
                            _p(st, u8), C.c_longlong(N), _p(pos, f32), _p(rot, f32), _p(scale, f32), _p(shs, f32))


def sh_rotate(R, shs48, flip_odd=True):
    R = np.ascontiguousarray(R, f32)
    out = np.ascontiguousarray(shs48, f32).copy()
    lib().orc_sh_rotate(_p(R, f32), _p(out, f32), int(bool(flip_odd)))
    return out


def sh_matrices(R):
    R = np.ascontiguousarray(R, f32)
    b1, b2, b3 = np.zeros((3, 3), f32), np.zeros((5, 5), f32), np.zeros((7, 7), f32)
    lib().orc_sh_matrices(_p(R, f32), _p(b1, f32), _p(b2, f32), _p(b3, f32))
    return b1, b2, b3


def polar(M):
    M = np.ascontiguousarray(M, f64)
    R, S = np.zeros((3, 3), f64), np.zeros((3, 3), f64)
    lib().orc_polar(_p(M, f64), _p(R, f64), _p(S, f64))
    return R, S


def node_quats(rot):
    rot = np.ascontiguousarray(rot, f64)
    q = np.zeros((len(rot), 4), f32)
    lib().orc_node_quats(_p(rot, f64), len(rot), _p(q, f32))
    return q


def rotate_sample_shs(w, idx, q_xyzw, is_static, feature):
    S, k = idx.shape
    assert feature.dtype == f32 and feature.flags["C_CONTIGUOUS"]
    st = None if is_static is None else np.ascontiguousarray(is_static, i32)
    lib().orc_rotate_sample_shs(C.c_longlong(S), int(k), _p(np.ascontiguousarray(w, f32), f32),
                                _p(np.ascontiguousarray(idx, i32), i32), _p(np.ascontiguousarray(q_xyzw, f32), f32),
                                _p(st, i32), _p(feature, f32))


def static_flags(idx, group, node_static):
    idx = np.ascontiguousarray(idx, u32)
    P, k = idx.shape
    out = np.zeros(P // group, u8)
    lib().orc_static_flags(_p(idx, u32), C.c_longlong(P), int(k), int(group), _p(np.ascontiguousarray(node_static, u8), u8), _p(out, u8))
    return out


# --------------------------------------------------------------------------- solve
def _blocks_csr(blocks):
    off = np.zeros(len(blocks) + 1, i32)
    for i, b in enumerate(blocks):
        off[i + 1] = off[i] + len(b)
    nodes = np.concatenate([np.asarray(b, u32) for b in blocks]) if blocks else np.zeros(0, u32)
    return off, np.ascontiguousarray(nodes, u32)


def solve(node_pos, nbr, anchor_idx, anchor_w, node_static, blocks, block_types, aim, on_center,
          w_rot=1.0, w_reg=10.0, w_con=100.0, weight_factor=1.0, max_iters=30):
    """Gauss-Newton solve (Deform::real_time_deform). -> rot[M,9], trans[M,3], stats dict."""
    node_pos = np.ascontiguousarray(node_pos, f32)
    M = len(node_pos)
    k = nbr.shape[1]
    off, bn = _blocks_csr(blocks)
    bt = np.ascontiguousarray(block_types, i32)
    rot, trans, stats = np.zeros((M, 9), f64), np.zeros((M, 3), f64), np.zeros(8, f64)
    rc = lib().orc_solve(M, k, _p(node_pos, f32), _p(np.ascontiguousarray(nbr, u32), u32),
                         _p(np.ascontiguousarray(anchor_idx[:, :k], u32), u32), _p(np.ascontiguousarray(anchor_w, f64), f64),
                         _p(np.ascontiguousarray(node_static, u8), u8), len(blocks), _p(off, i32), _p(bn, u32), _p(bt, i32),
                         _p(np.ascontiguousarray(aim, f32), f32), int(bool(on_center)),
                         C.c_double(w_rot), C.c_double(w_reg), C.c_double(w_con), C.c_double(weight_factor), int(max_iters),
                         _p(rot, f64), _p(trans, f64), _p(stats, f64))
    if rc != 0:
        raise RuntimeError("orc_solve failed")
    return rot, trans, dict(iters=int(stats[0]), energy=stats[1], halvings=int(stats[2]), normh=stats[3],
                            rows=int(stats[4]), unknowns=int(stats[5]))


def energy(node_pos, nbr, anchor_idx, anchor_w, node_static, blocks, block_types, aim, on_center, rot, trans,
           w_rot=1.0, w_reg=10.0, w_con=100.0):
    node_pos = np.ascontiguousarray(node_pos, f32)
    M, k = len(node_pos), nbr.shape[1]
    off, bn = _blocks_csr(blocks)
    return lib().orc_energy(M, k, _p(node_pos, f32), _p(np.ascontiguousarray(nbr, u32), u32),
                            _p(np.ascontiguousarray(anchor_idx[:, :k], u32), u32), _p(np.ascontiguousarray(anchor_w, f64), f64),
                            _p(np.ascontiguousarray(node_static, u8), u8), len(blocks), _p(off, i32), _p(bn, u32),
                            _p(np.ascontiguousarray(block_types, i32), i32), _p(np.ascontiguousarray(aim, f32), f32),
                            int(bool(on_center)), C.c_double(w_rot), C.c_double(w_reg), C.c_double(w_con),
                            _p(np.ascontiguousarray(rot, f64), f64), _p(np.ascontiguousarray(trans, f64), f64))


def jacobian(node_pos, nbr, anchor_idx, anchor_w, node_static, blocks, block_types, aim, on_center, rot, trans,
             w_rot=1.0, w_reg=10.0, w_con=100.0):
    """-> (rows, cols, vals, f, (m, n)) COO Jacobian in the reference's row order, residual f."""
    node_pos = np.ascontiguousarray(node_pos, f32)
    M, k = len(node_pos), nbr.shape[1]
    off, bn = _blocks_csr(blocks)
    cap = M * (27 + 15 * k) + 3 * 4 * k * (int(off[-1]) + 8) * 3 + 1024
    R, Cc, V = np.zeros(cap, i32), np.zeros(cap, i32), np.zeros(cap, f64)
    dims = np.zeros(2, i32)
    fbuf = np.zeros(M * (6 + 6 * k) + 3 * int(off[-1]) + 64, f64)
    nnz = lib().orc_jacobian(M, k, _p(node_pos, f32), _p(np.ascontiguousarray(nbr, u32), u32),
                             _p(np.ascontiguousarray(anchor_idx[:, :k], u32), u32), _p(np.ascontiguousarray(anchor_w, f64), f64),
                             _p(np.ascontiguousarray(node_static, u8), u8), len(blocks), _p(off, i32), _p(bn, u32),
                             _p(np.ascontiguousarray(block_types, i32), i32), _p(np.ascontiguousarray(aim, f32), f32),
                             int(bool(on_center)), C.c_double(w_rot), C.c_double(w_reg), C.c_double(w_con),
                             _p(np.ascontiguousarray(rot, f64), f64), _p(np.ascontiguousarray(trans, f64), f64),
                             C.c_longlong(cap), _p(R, i32), _p(Cc, i32), _p(V, f64), _p(fbuf, f64), _p(dims, i32))
    assert nnz >= 0, "jacobian cap too small"
    return R[:nnz], Cc[:nnz], V[:nnz], fbuf[:dims[0]].copy(), (int(dims[0]), int(dims[1]))


# --------------------------------------------------------------------------- grid
def overall_aabb(pos):
    pos = np.ascontiguousarray(pos, f32)
    out = np.zeros(6, f32)
    lib().orc_overall_aabb(_p(pos, f32), C.c_longlong(len(pos)), _p(out, f32))
    return out


def grid_step(aabb, G):
    return np.float32(lib().orc_grid_step(_p(np.ascontiguousarray(aabb, f32), f32), int(G)))


def cell_assign(pos, min3, step, G):
    pos = np.ascontiguousarray(pos, f32)
    N = len(pos)
    cell, prefix, new_idx = np.zeros(N, i32), np.zeros(G ** 3, i32), np.zeros(N, i32)
    lib().orc_cell_assign(_p(pos, f32), C.c_longlong(N), _p(np.ascontiguousarray(min3, f32), f32), C.c_float(step), int(G),
                          _p(cell, i32), _p(prefix, i32), _p(new_idx, i32))
    return cell, prefix, new_idx


def gs_aabbs(pos, rot, scale, opacity):
    N = len(pos)
    aabb, clip, smax = np.zeros((N, 6), f32), np.zeros((N, 3), f32), np.zeros(N, f32)
    lib().orc_gs_aabbs(_p(np.ascontiguousarray(pos, f32), f32), _p(np.ascontiguousarray(rot, f32), f32),
                       _p(np.ascontiguousarray(scale, f32), f32), _p(np.ascontiguousarray(opacity, f32), f32),
                       C.c_longlong(N), _p(aabb, f32), _p(clip, f32), _p(smax, f32))
    return aabb, clip, smax


def footprint_lists(aabb, min3, step, G, padding=1):
    aabb = np.ascontiguousarray(aabb, f32)
    N = len(aabb)
    m3 = np.ascontiguousarray(min3, f32)
    prefix = np.zeros(G ** 3, i32)
    P = lib().orc_footprint_count(_p(aabb, f32), C.c_longlong(N), _p(m3, f32), C.c_float(step), int(G), int(padding), _p(prefix, i32))
    lists = np.zeros(P, i32)
    lib().orc_footprint_fill(_p(aabb, f32), C.c_longlong(N), _p(m3, f32), C.c_float(step), int(G), int(padding), _p(prefix, i32), _p(lists, i32))
    return prefix, lists


def valid_cells(prefix, G):
    prefix = np.ascontiguousarray(prefix, i32)
    V = lib().orc_valid_cells(_p(prefix, i32), int(G), None)
    out = np.zeros(V, i32)
    lib().orc_valid_cells(_p(prefix, i32), int(G), _p(out, i32))
    return out


def emit_samples(valid, min3, step, G):
    valid = np.ascontiguousarray(valid, i32)
    out = np.zeros((len(valid) * 64, 3), f32)
    lib().orc_emit_samples(_p(valid, i32), len(valid), _p(np.ascontiguousarray(min3, f32), f32), C.c_float(step), int(G), _p(out, f32))
    return out


def ada_lpf(samples, valid, G, lpf_parameter=0.2):
    out = np.zeros((G ** 3, 9), f32)
    lib().orc_ada_lpf(_p(np.ascontiguousarray(samples, f32), f32), _p(np.ascontiguousarray(valid, i32), i32), len(valid),
                      C.c_float(lpf_parameter), _p(out, f32))
    return out


def grid_eval(valid, prefix, lists, samples, pos, rot, scale, opacity, shs, lpf):
    V = len(valid)
    feat, opa = np.zeros((V * 64, 48), f32), np.zeros(V * 64, f32)
    a = lambda x, t: _p(np.ascontiguousarray(x, t), t)
    lib().orc_grid_eval(a(valid, i32), V, a(prefix, i32), a(lists, i32), a(samples, f32), a(pos, f32), a(rot, f32),
                        a(scale, f32), a(opacity, f32), a(shs, f32), a(lpf, f32), _p(feat, f32), _p(opa, f32))
    return feat, opa


def judge_empty(valid, G, aim_opacity):
    out = np.zeros(len(valid), i32)
    lib().orc_judge_empty(_p(np.ascontiguousarray(valid, i32), i32), len(valid), int(G),
                          _p(np.ascontiguousarray(aim_opacity, f32), f32), _p(out, i32))
    return out


def rotate_by_axis(point, center, axis4, radian):
    out = np.zeros(3, f32)
    lib().orc_rotate_by_axis(_p(np.ascontiguousarray(point, f32), f32), _p(np.ascontiguousarray(center, f32), f32),
                             _p(np.ascontiguousarray(axis4, f32), f32), C.c_float(radian), _p(out, f32))
    return out
