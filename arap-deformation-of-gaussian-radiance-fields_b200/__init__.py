"""arapgs-b200: B200-native ARAP deformation of Gaussian radiance fields.

Thin ctypes binding over ``libarapgs.so`` (the C ABI of ``include/arapgs.h``).
This module is plumbing only: every computation happens in the CUDA library,
and importing/using it without the built library or without a CUDA device
raises — there is no CPU fallback.

The directory name contains hyphens; import it with
``importlib.import_module("arap-deformation-of-gaussian-radiance-fields_b200")``
(``__graft_entry__.load_package()`` does that).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build  # noqa: F401  (re-exported: pkg.build.build())

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "libarapgs.so"

f32, f64, u32, i32, u8 = np.float32, np.float64, np.uint32, np.int32, np.uint8


class ArapError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libarapgs error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("grid_num", C.c_int), ("padding", C.c_int), ("knn_k", C.c_int), ("node_num", C.c_int),
                ("high_quality", C.c_int), ("lpf_parameter", C.c_float),
                ("w_rot", C.c_double), ("w_reg", C.c_double), ("w_con", C.c_double),
                ("max_gn_iters", C.c_int), ("max_cg_iters", C.c_int), ("cg_tol", C.c_double),
                ("skip_static_endpoints", C.c_int), ("solver_global_memory", C.c_int), ("lbs_mode", C.c_int),
                ("newton_eta0", C.c_double), ("warm_start", C.c_int), ("solver_ctas", C.c_int), ("lazy_sample_sh", C.c_int), ("fps_mode", C.c_int), ("solver_pipelined", C.c_int)]


class SolveStats(C.Structure):
    _fields_ = [("gn_iters", C.c_int), ("cg_iters", C.c_int), ("halvings", C.c_int), ("flags", C.c_int),
                ("energy", C.c_double), ("normh", C.c_double), ("last_rel_residual", C.c_double),
                ("phase_ns", C.c_double * 4), ("grid_blocks", C.c_int), ("row_sub_ns", C.c_double * 4), ("cg_iters_gn", C.c_int * 8), ("barrier_skew_ns", C.c_double * 6)]


class GridInfo(C.Structure):
    _fields_ = [("grid_num", C.c_int), ("padding", C.c_int), ("valid_cells", C.c_int), ("samples", C.c_longlong),
                ("pairs", C.c_longlong), ("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3), ("grid_step", C.c_float)]


class DeviceView(C.Structure):
    _fields_ = [("n_gaussians", C.c_longlong), ("pos", C.c_void_p), ("rot", C.c_void_p), ("scale", C.c_void_p),
                ("opacity", C.c_void_p), ("shs", C.c_void_p), ("n_nodes", C.c_int), ("node_pos", C.c_void_p),
                ("node_rot", C.c_void_p), ("node_trans", C.c_void_p), ("n_samples", C.c_longlong),
                ("sample_pos", C.c_void_p), ("aim_feature", C.c_void_p), ("aim_opacity", C.c_void_p),
                ("valid_grid", C.c_void_p), ("grid_gs_prefix_sum", C.c_void_p), ("grided_gs_idx", C.c_void_p),
                ("gs_init_grid_idx", C.c_void_p), ("ada_lpf_ratio", C.c_void_p), ("end_points", C.c_void_p),
                ("empty_grid", C.c_void_p), ("cur_feature", C.c_void_p), ("cur_opacity", C.c_void_p)]


class GatheredView(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("n_per_rank", C.c_longlong), ("pos", C.c_void_p), ("rot", C.c_void_p),
                ("scale", C.c_void_p), ("shs", C.c_void_p), ("side_stream", C.c_void_p)]


_lib = None


def lib() -> C.CDLL:
    """Load libarapgs.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not _SO.exists():
            raise ImportError(f"{_SO} is missing: run __graft_entry__.build() (there is no CPU fallback)")
        _lib = C.CDLL(str(_SO))
        _lib.arap_last_error.restype = C.c_char_p
        _lib.arapk_knn_workspace_bytes.restype = C.c_size_t
        _lib.arapk_knn_index_struct_bytes.restype = C.c_size_t
        _lib.arapk_solve_workspace_bytes.restype = C.c_size_t
        _lib.arapk_grid_scratch_bytes.restype = C.c_size_t
        _lib.arapk_lbs_tile_count.restype = C.c_longlong
        _lib.arapk_lbs_tile_count.argtypes = [C.c_longlong]
        _lib.arapk_grid_scratch_bytes.argtypes = [C.c_longlong, C.c_int]
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ArapError(rc, lib().arap_last_error().decode(errors="replace"))


def _np(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype)
    return a


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C ABI (rank 0; hand the 128 bytes to the other ranks out of band)."""
    buf = C.create_string_buffer(128)
    check(lib().arap_comm_unique_id(buf))
    return buf.raw


def default_params(**kw) -> Params:
    p = Params()
    check(lib().arap_default_params(C.byref(p)))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class History:
    """deform.txt record (reference DeformHistory, helper.hpp:172-198)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load(cls, path):
        h = C.c_void_p()
        check(lib().arap_history_load(str(path).encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def new(cls, nodes_on_mesh, anchors):
        h = C.c_void_p()
        a = _np(anchors, i32)
        check(lib().arap_history_new(C.byref(h), int(nodes_on_mesh), _ptr(a), len(a)))
        return cls(h)

    def save(self, path):
        check(lib().arap_history_save(self._h, str(path).encode()))

    def add_block(self, nodes):
        a = _np(nodes, u32)
        check(lib().arap_history_add_block(self._h, _ptr(a), len(a)))

    def add_move(self, op_type, movements, block_types, energy_on_center=0, twist_axis=(0, 0, 0, 0)):
        mv = _np(movements, f32).reshape(-1, 3)
        bt = _np(block_types, i32)
        ax = _np(twist_axis, f32)
        check(lib().arap_history_add_move(self._h, int(op_type), _ptr(mv), len(mv), _ptr(bt), len(bt), int(energy_on_center), _ptr(ax)))

    def summary(self):
        o = np.zeros(6, i32)
        check(lib().arap_history_summary(self._h, _ptr(o)))
        return dict(nodes_on_mesh=int(o[0]), n_nodes=int(o[1]), total_ops=int(o[2]), move_ops=int(o[3]), n_blocks=int(o[4]), n_moves=int(o[5]))

    def nodes(self):
        s = self.summary()
        o = np.zeros(s["n_nodes"], i32)
        check(lib().arap_history_nodes(self._h, _ptr(o)))
        return o

    def ops(self):
        o = np.zeros(self.summary()["total_ops"], i32)
        check(lib().arap_history_ops(self._h, _ptr(o)))
        return o

    def block(self, i):
        n = C.c_int()
        check(lib().arap_history_block(self._h, i, None, C.byref(n)))
        o = np.zeros(n.value, u32)
        check(lib().arap_history_block(self._h, i, _ptr(o), C.byref(n)))
        return o

    def move(self, i):
        n, nt, eoc = C.c_int(), C.c_int(), C.c_int()
        check(lib().arap_history_move(self._h, i, None, C.byref(n), None, C.byref(nt), C.byref(eoc), None))
        mv, bt, ax = np.zeros((n.value, 3), f32), np.zeros(nt.value, i32), np.zeros(4, f32)
        check(lib().arap_history_move(self._h, i, _ptr(mv), C.byref(n), _ptr(bt), C.byref(nt), C.byref(eoc), _ptr(ax)))
        return dict(movements=mv, block_types=bt, energy_on_center=eoc.value, twist_axis=ax)

    def __del__(self):
        try:
            if self._h:
                lib().arap_history_free(self._h)
                self._h = None
        except Exception:
            pass


def graph_obj_load(path) -> np.ndarray:
    n = C.c_int()
    check(lib().arap_graph_obj_load(str(path).encode(), None, C.byref(n)))
    pts = np.zeros((n.value, 3), f32)
    check(lib().arap_graph_obj_load(str(path).encode(), _ptr(pts), C.byref(n)))
    return pts


def ply_load(path) -> dict:
    """3DGS PLY -> activated, Morton-ordered SoA (arap_ply_load).  Keys: pos, rot (w,x,y,z), scale, opacity, shs, index, aabb_min/max."""
    n = C.c_longlong()
    check(lib().arap_ply_load(str(path).encode(), C.byref(n), None, None, None, None, None, None, None, None))
    N = n.value
    out = dict(pos=np.zeros((N, 3), f32), rot=np.zeros((N, 4), f32), scale=np.zeros((N, 3), f32), opacity=np.zeros(N, f32),
               shs=np.zeros((N, 48), f32), index=np.zeros(N, np.int32), aabb_min=np.zeros(3, f32), aabb_max=np.zeros(3, f32))
    check(lib().arap_ply_load(str(path).encode(), C.byref(n), _ptr(out["pos"]), _ptr(out["rot"]), _ptr(out["scale"]), _ptr(out["opacity"]),
                              _ptr(out["shs"]), _ptr(out["index"]), _ptr(out["aabb_min"]), _ptr(out["aabb_max"])))
    return out


def ply_save(path, g: dict, box_min=None, box_max=None, skip=None) -> int:
    """Write a 3DGS PLY in the reference's format (arap_ply_save); returns the number of vertices written."""
    pos, rot, scale = _np(g["pos"], f32), _np(g["rot"], f32), _np(g["scale"], f32)
    op, shs = _np(g["opacity"], f32), _np(g["shs"], f32)
    bmin = _np(box_min, f32) if box_min is not None else None
    bmax = _np(box_max, f32) if box_max is not None else None
    sk = _np(skip, np.uint8) if skip is not None else None
    w = C.c_longlong()
    check(lib().arap_ply_save(str(path).encode(), C.c_longlong(len(pos)), _ptr(pos), _ptr(rot), _ptr(scale), _ptr(op), _ptr(shs),
                              _ptr(bmin) if bmin is not None else None, _ptr(bmax) if bmax is not None else None,
                              _ptr(sk) if sk is not None else None, C.byref(w)))
    return w.value


def graph_obj_save(path, pts) -> None:
    p = _np(pts, f32)
    check(lib().arap_graph_obj_save(str(path).encode(), _ptr(p), len(p)))


def config_load(path):
    g, s, soup, hq = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib().arap_config_load(str(path).encode(), C.byref(g), C.byref(s), C.byref(soup), C.byref(hq)))
    return dict(grid_num=g.value, is_synthetic=s.value, has_soup=soup.value, high_quality=hq.value)


class Session:
    """One deformation context (mirrors the deformation API of GaussianView)."""

    def __init__(self, device: int = 0, stream: int | None = None, **params):
        self._ctx = C.c_void_p()
        self.params = default_params(**params)
        check(lib().arap_create(C.byref(self._ctx), int(device), C.c_void_p(stream) if stream else None, C.byref(self.params)))

    def close(self):
        if self._ctx:
            lib().arap_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.params, k, v)
        check(lib().arap_set_params(self._ctx, C.byref(self.params)))

    def sync(self):
        check(lib().arap_sync(self._ctx))

    # -- Gaussians
    def set_gaussians(self, pos, rot, scale, opacity, shs):
        pos, rot, scale, opacity, shs = _np(pos, f32), _np(rot, f32), _np(scale, f32), _np(opacity, f32), _np(shs, f32)
        n = len(pos)
        assert rot.shape == (n, 4) and scale.shape == (n, 3) and opacity.shape[0] == n and shs.reshape(n, -1).shape[1] == 48
        check(lib().arap_set_gaussians(self._ctx, C.c_longlong(n), _ptr(pos), _ptr(rot), _ptr(scale), _ptr(opacity), _ptr(shs), 0))
        self.N = n

    def set_gaussians_device(self, n, pos, rot, scale, opacity, shs):
        """Device pointers (ints), e.g. torch tensors' data_ptr()."""
        check(lib().arap_set_gaussians(self._ctx, C.c_longlong(n), *(C.c_void_p(int(p)) for p in (pos, rot, scale, opacity, shs)), 1))
        self.N = n

    def download_gaussians(self):
        n = self.N
        out = dict(pos=np.zeros((n, 3), f32), rot=np.zeros((n, 4), f32), scale=np.zeros((n, 3), f32),
                   opacity=np.zeros(n, f32), shs=np.zeros((n, 48), f32))
        check(lib().arap_download_gaussians(self._ctx, *(_ptr(out[k]) for k in ("pos", "rot", "scale", "opacity", "shs"))))
        return out

    def device_view(self) -> DeviceView:
        v = DeviceView()
        check(lib().arap_get_device_view(self._ctx, C.byref(v)))
        return v

    # -- grid
    def grid_build(self):
        check(lib().arap_grid_build(self._ctx))
        return self.grid_info()

    def grid_update_lists(self):
        check(lib().arap_grid_update_lists(self._ctx))

    def grid_eval(self, which=0):
        check(lib().arap_grid_eval(self._ctx, int(which)))

    def grid_info(self):
        g = GridInfo()
        check(lib().arap_grid_info_get(self._ctx, C.byref(g)))
        return dict(grid_num=g.grid_num, padding=g.padding, valid_cells=g.valid_cells, samples=g.samples, pairs=g.pairs,
                    aabb_min=np.array(g.aabb_min[:], f32), aabb_max=np.array(g.aabb_max[:], f32), grid_step=np.float32(g.grid_step))

    def download_grid(self):
        gi = self.grid_info()
        G = gi["grid_num"]
        out = dict(valid=np.zeros(gi["valid_cells"], i32), prefix=np.zeros(G ** 3, i32), lists=np.zeros(gi["pairs"], i32),
                   sample_pos=np.zeros((gi["samples"], 3), f32), gs_init_grid_idx=np.zeros(self.N, i32))
        check(lib().arap_download_grid(self._ctx, *(_ptr(out[k]) for k in ("valid", "prefix", "lists", "sample_pos", "gs_init_grid_idx"))))
        return out

    def ada_lpf_update(self):
        check(lib().arap_ada_lpf_update(self._ctx))

    def download_ada_lpf(self):
        G = self.grid_info()["grid_num"]
        out = np.zeros((G ** 3, 9), f32)
        check(lib().arap_download_ada_lpf(self._ctx, _ptr(out)))
        return out

    def download_empty_grid(self):
        out = np.zeros(self.grid_info()["valid_cells"], i32)
        check(lib().arap_download_empty_grid(self._ctx, _ptr(out)))
        return out

    def download_features(self, which=0):
        S = self.grid_info()["samples"]
        f, o = np.zeros((S, 48), f32), np.zeros(S, f32)
        check(lib().arap_download_features(self._ctx, int(which), _ptr(f), _ptr(o)))
        return f, o

    def sample_features_materialize(self):
        check(lib().arap_sample_features_materialize(self._ctx))

    def download_samples(self, features=True):
        S = self.grid_info()["samples"]
        p = np.zeros((S, 3), f32)
        f = np.zeros((S, 48), f32) if features else None
        check(lib().arap_download_samples(self._ctx, _ptr(p), _ptr(f)))
        return p, f

    # -- graph
    def set_mesh_points(self, pts, nodes_on_mesh=True):
        p = _np(pts, f32).reshape(-1, 3)
        check(lib().arap_set_mesh_points(self._ctx, _ptr(p), len(p), int(bool(nodes_on_mesh))))

    def set_points(self, family, pts):
        """family 0 = textured-mesh points, 1 = soup points (arap_set_points); skinned every step."""
        p = _np(pts, f32).reshape(-1, 3)
        self._npts = getattr(self, "_npts", {})
        self._npts[int(family)] = len(p)
        check(lib().arap_set_points(self._ctx, int(family), _ptr(p), C.c_longlong(len(p))))

    def download_points(self, family, n=None):
        n = self._npts[int(family)] if n is None else n
        out = np.zeros((n, 3), f32)
        check(lib().arap_download_points(self._ctx, int(family), _ptr(out)))
        return out

    def graph_build_fps(self, node_num=None, k=None):
        check(lib().arap_graph_build_fps(self._ctx, int(node_num or self.params.node_num), int(k or self.params.knn_k)))
        return self.download_graph()

    def graph_build_anchors(self, anchors, k=None):
        a = _np(anchors, i32)
        check(lib().arap_graph_build_anchors(self._ctx, _ptr(a), len(a), int(k or self.params.knn_k)))
        return self.download_graph()

    def download_graph(self):
        v = self.device_view()
        M = v.n_nodes
        # k is recovered from the neighbour table size via a probe download
        anchor, pos = np.zeros(M, i32), np.zeros((M, 3), f32)
        nbr = np.zeros((M, ARAP_KNN_MAX), i32)
        check(lib().arap_download_graph(self._ctx, _ptr(anchor), _ptr(pos), None))
        self.M = M
        return dict(anchor=anchor, node_pos=pos)

    def download_edges(self, k):
        nbr = np.zeros((self.M, k), i32)
        check(lib().arap_download_graph(self._ctx, None, None, _ptr(nbr)))
        return nbr

    def download_rows(self, family, rows, k):
        fam = dict(ends=0, samples=1, mesh=2, nodes=3)[family]
        idx, w = np.zeros((rows, k), u32), np.zeros((rows, k), f64)
        check(lib().arap_download_rows(self._ctx, fam, _ptr(idx), _ptr(w)))
        return idx, w

    # -- blocks / aims
    def set_blocks(self, blocks, types):
        off = np.zeros(len(blocks) + 1, i32)
        for i, b in enumerate(blocks):
            off[i + 1] = off[i] + len(b)
        nodes = np.concatenate([np.asarray(b, u32) for b in blocks]) if len(blocks) else np.zeros(1, u32)
        nodes = _np(nodes, u32)
        t = _np(types, i32) if len(blocks) else np.zeros(1, i32)
        check(lib().arap_set_blocks(self._ctx, len(blocks), _ptr(off), _ptr(nodes), _ptr(t)))

    def download_end_points(self):
        e = np.zeros((self.N, 18), f32)
        check(lib().arap_download_end_points(self._ctx, _ptr(e)))
        return e

    def static_flags(self):
        g, s = np.zeros(self.N, u8), np.zeros(self.grid_info()["samples"], u8)
        check(lib().arap_download_static_flags(self._ctx, _ptr(g), _ptr(s)))
        return g, s

    def aim_translate(self, delta):
        d = _np(delta, f32)
        check(lib().arap_aim_translate(self._ctx, _ptr(d)))

    def aim_twist(self, axis4, y):
        a = _np(axis4, f32)
        check(lib().arap_aim_twist(self._ctx, _ptr(a), int(y)))

    def aim_scale(self, y):
        check(lib().arap_aim_scale(self._ctx, int(y)))

    def aim_set(self, aim):
        a = _np(aim, f32)
        check(lib().arap_aim_set(self._ctx, _ptr(a)))
        self.sync()

    def aim_get(self):
        a = np.zeros((self.M, 3), f32)
        check(lib().arap_aim_get(self._ctx, _ptr(a)))
        return a

    # -- step
    def solve(self, on_center=False):
        check(lib().arap_solve(self._ctx, int(bool(on_center))))

    def solve_stats(self):
        s = SolveStats()
        check(lib().arap_solve_stats_get(self._ctx, C.byref(s)))
        return dict(gn_iters=s.gn_iters, cg_iters=s.cg_iters, halvings=s.halvings, flags=s.flags, energy=s.energy,
                    normh=s.normh, last_rel_residual=s.last_rel_residual, phase_ns=list(s.phase_ns), grid_blocks=s.grid_blocks, row_sub_ns=list(s.row_sub_ns),
                    cg_iters_gn=list(s.cg_iters_gn), barrier_skew_ns=list(s.barrier_skew_ns))

    def apply(self):
        check(lib().arap_apply(self._ctx))

    def step(self, on_center=False):
        check(lib().arap_step(self._ctx, int(bool(on_center))))

    def soa_release_event(self, event):
        """Hand over a CUDA event (integer handle) the next apply waits for before it overwrites the SoA."""
        check(lib().arap_soa_release_event(self._ctx, C.c_void_p(int(event))))

    def soa_ready_wait(self, stream):
        """Make CUDA stream `stream` (integer handle) wait for the last step's deformed SoA (see arap_soa_ready_wait)."""
        check(lib().arap_soa_ready_wait(self._ctx, C.c_void_p(int(stream))))

    def download_nodes(self):
        pos, rot, trans = np.zeros((self.M, 3), f32), np.zeros((self.M, 9), f64), np.zeros((self.M, 3), f64)
        check(lib().arap_download_nodes(self._ctx, _ptr(pos), _ptr(rot), _ptr(trans)))
        return pos, rot, trans

    SETUP_STAGES = ("scene_aabb", "cell_assign", "reorder", "footprint_lists", "samples", "grid_eval", "fps", "node_graph",
                    "knn_ends", "knn_samples", "tile_tables")

    def setup_timing(self):
        """Device ms of the last run of every set-up / stroke-end stage (arap_setup_timing)."""
        ms = np.zeros(len(self.SETUP_STAGES), f32)
        check(lib().arap_setup_timing(self._ctx, _ptr(ms), len(ms)))
        return {k: float(v) for k, v in zip(self.SETUP_STAGES, ms)}

    # -- multi-GPU exchange (arap_comm_*)
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        check(lib().arap_comm_init(self._ctx, C.c_char_p(unique_id), int(rank), int(world)))

    def comm_exchange(self):
        check(lib().arap_comm_exchange(self._ctx))

    def comm_set_mode(self, mode):
        """0 = NCCL all-gather per exchange, 1 = exchange fused into the apply kernel (peer stores over NVLink)."""
        check(lib().arap_comm_set_mode(self._ctx, int(mode)))

    def comm_grid_build(self, x_lo=-1, x_hi=-1):
        """One scene sharded over the ranks: grid over everybody's Gaussians, this rank's x-slab of cells (arap_comm_grid_build)."""
        check(lib().arap_comm_grid_build(self._ctx, int(x_lo), int(x_hi)))
        return self.grid_info()

    def comm_slab(self):
        lo, hi = C.c_int(), C.c_int()
        check(lib().arap_comm_slab_get(self._ctx, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def comm_materialize_sh(self):
        check(lib().arap_comm_materialize_sh(self._ctx))

    def comm_sync(self):
        check(lib().arap_comm_sync(self._ctx))

    def comm_view(self) -> GatheredView:
        v = GatheredView()
        check(lib().arap_comm_view(self._ctx, C.byref(v)))
        return v

    def comm_destroy(self):
        check(lib().arap_comm_destroy(self._ctx))

    def enable_timing(self, on=True):
        check(lib().arap_enable_timing(self._ctx, int(bool(on))))

    def last_step_timing(self):
        ms = np.zeros(6, f32)
        check(lib().arap_last_step_timing(self._ctx, _ptr(ms)))
        return dict(solve=float(ms[0]), samples_lbs=float(ms[1]), points_lbs=float(ms[2]), fit=float(ms[3]), sample_sh=float(ms[4]), total=float(ms[5]))

    def step_timings(self, max_steps=128):
        ms = np.zeros((max_steps, 6), f32)
        n = C.c_int()
        check(lib().arap_step_timings(self._ctx, _ptr(ms), int(max_steps), C.byref(n)))
        return ms[:n.value]   # columns: solve, samples_lbs, points_lbs, fit, sample_sh, total

    # -- replay
    def replay(self, history: History, rebuild_graph=True) -> int:
        n = C.c_int()
        check(lib().arap_replay(self._ctx, history._h, int(bool(rebuild_graph)), C.byref(n)))
        v = self.device_view()
        self.M = v.n_nodes
        return n.value

    def run_script(self, script_id: int) -> int:
        n = C.c_int()
        check(lib().arap_run_script(self._ctx, int(script_id), C.byref(n)))
        return n.value


ARAP_KNN_MAX = 12
