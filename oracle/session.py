"""ORACLE — TEST INFRASTRUCTURE ONLY.

Python mirror of the deformation state machine of the reference's GaussianView
(init -> grid -> graph -> blocks -> aims -> per-step driver -> replay), built on
the CPU restatement in ``arap_oracle.cpp``.  File:line citations refer to
``src/projects/gaussianviewer/renderer/GaussianView.cpp`` (GV) of the reference.
"""
from __future__ import annotations

import time

import numpy as np

import oracle as O

f32 = np.float32


class OracleSession:
    def __init__(self, gaussians: dict, grid_num=64, padding=1, knn_k=10, node_num=150, lpf_parameter=0.2,
                 w_rot=1.0, w_reg=10.0, w_con=100.0, max_gn_iters=30, with_samples=True):
        self.g = {k: np.ascontiguousarray(gaussians[k], f32).copy() for k in ("pos", "rot", "scale", "opacity", "shs")}
        self.g["shs"] = self.g["shs"].reshape(len(self.g["pos"]), 48)
        self.N = len(self.g["pos"])
        self.G, self.padding, self.k, self.node_num = grid_num, padding, knn_k, node_num
        self.lpf_parameter = lpf_parameter
        self.w = (w_rot, w_reg, w_con)
        self.max_gn_iters = max_gn_iters
        self.with_samples = with_samples
        self.mesh_pts = np.zeros((0, 3), f32)
        self.nodes_on_mesh = False
        self.blocks, self.block_types = [], []
        self.aim_feature = None
        self.timing = {}

    # ---- grid (GV:3601-3631, 3870-4149, 4670-4751)
    def grid_build(self):
        g = self.g
        self.aabb = O.overall_aabb(g["pos"])
        self.gstep = O.grid_step(self.aabb, self.G)
        cell, prefix, new_idx = O.cell_assign(g["pos"], self.aabb[:3], self.gstep, self.G)
        for k in g:  # re-order (GV:3938-3953): out[new_idx[i]] = in[i]
            out = np.empty_like(g[k]); out[new_idx] = g[k]; g[k] = out
        self.new_idx = new_idx
        self.scale_backup = g["scale"].copy()
        self.gs_init_grid_idx, self.cell_prefix, _ = O.cell_assign(g["pos"], self.aabb[:3], self.gstep, self.G)
        self._build_lists()
        self.valid = O.valid_cells(self.fp_prefix, self.G)
        self.sample_pos = O.emit_samples(self.valid, self.aabb[:3], self.gstep, self.G)
        self.ada_lpf = O.ada_lpf(self.sample_pos, self.valid, self.G, self.lpf_parameter)
        self.sample_static = np.zeros(len(self.sample_pos), np.uint8)
        self.gs_static = np.zeros(self.N, np.uint8)
        self.ends = O.end_points(g["pos"], g["rot"], g["scale"]).reshape(self.N * 6, 3)
        return dict(valid_cells=len(self.valid), samples=len(self.sample_pos), pairs=len(self.lists), grid_step=self.gstep,
                    aabb_min=self.aabb[:3].copy(), aabb_max=self.aabb[3:].copy())

    def _build_lists(self):
        g = self.g
        self.gs_aabb, _, _ = O.gs_aabbs(g["pos"], g["rot"], g["scale"], g["opacity"])
        self.fp_prefix, self.lists = O.footprint_lists(self.gs_aabb, self.aabb[:3], self.gstep, self.G, self.padding)

    def grid_update_lists(self):  # UpdateContainingRelationship (GV:3634-3743)
        self.aabb = O.overall_aabb(self.g["pos"])
        self.gstep = O.grid_step(self.aabb, self.G)
        self._build_lists()

    def ada_lpf_update(self):  # GetAdaLpfRatio on the deformed samples (GV:958, 1704-1706)
        self.ada_lpf = O.ada_lpf(self.sample_pos, self.valid, self.G, self.lpf_parameter)

    def grid_eval(self, which=0):
        g = self.g
        f, o = O.grid_eval(self.valid, self.fp_prefix, self.lists, self.sample_pos, g["pos"], g["rot"], g["scale"], g["opacity"], g["shs"], self.ada_lpf)
        if which == 0:
            self.aim_feature, self.aim_opacity = f, o
            self.empty_grid = O.judge_empty(self.valid, self.G, o)   # GPUSetupSamplesFeatures -> JudgeEmptyGrid (GV:4268)
        else:
            self.cur_feature, self.cur_opacity = f, o
        return f, o

    # ---- graph (GV:698-754, 4825-4856, 4939-5010)
    def set_mesh_points(self, pts, nodes_on_mesh=True):
        self.mesh_pts = np.ascontiguousarray(pts, f32).reshape(-1, 3).copy()
        self.nodes_on_mesh = bool(nodes_on_mesh) and len(self.mesh_pts) > 0

    def _cand(self):
        return self.mesh_pts if self.nodes_on_mesh else self.g["pos"]

    def graph_build_fps(self, node_num=None, k=None):
        node_num = node_num or self.node_num
        if self.nodes_on_mesh:
            node_num = len(self.mesh_pts)
        return self.graph_build_anchors(O.fps(self._cand(), node_num), k)

    def graph_build_anchors(self, anchors, k=None):
        self.k = k or self.k
        k = self.k
        self.anchor = np.asarray(anchors, np.int32).copy()
        self.M = len(self.anchor)
        self.node_pos = self._cand()[self.anchor].copy()
        self.node_rest = self.node_pos.copy()
        self.aim = self.node_pos.copy()
        self.nbr = O.graph_edges(self.node_rest, k)
        idx, w = O.knn_weights(self.node_rest, self.node_rest, k)       # cand_vertices[Vertex_index] rows
        self.anc_idx, self.anc_w = idx[:, :k].copy(), w
        self.end_idx, self.end_w = self._rows(self.ends)
        if self.with_samples:
            self.smp_idx, self.smp_w = self._rows(self.sample_pos)
        self.mesh_idx, self.mesh_w = self._rows(self.mesh_pts)
        self.rot = np.tile(np.eye(3).reshape(-1), (self.M, 1))
        self.trans = np.zeros((self.M, 3))
        self.set_blocks([], [])
        return dict(anchor=self.anchor, node_pos=self.node_pos.copy())

    def _rows(self, pts):
        if len(pts) == 0:
            return np.zeros((0, self.k), np.uint32), np.zeros((0, self.k))
        idx, w = O.knn_weights(self.node_rest, pts, self.k)
        return np.ascontiguousarray(idx[:, :self.k]), w

    # ---- blocks (GV:1996-2087)
    def set_blocks(self, blocks, types):
        self.blocks = [np.asarray(b, np.uint32) for b in blocks]
        self.block_types = list(types)
        st = np.zeros(self.M, np.uint8)
        for b, t in zip(self.blocks, self.block_types):
            if t < 0:
                st[b] = 1
        self.node_static = st
        if self.with_samples and len(self.sample_pos):
            self.sample_static = O.static_flags(self.smp_idx, 1, st)
        self.gs_static = O.static_flags(self.end_idx, 6, st)

    # ---- aims (GV:2920-2983)
    def _active_entries(self):
        e = [b for b, t in zip(self.blocks, self.block_types) if t == 1]
        return np.concatenate(e) if e else np.zeros(0, np.uint32)

    def aim_translate(self, delta):
        d = np.asarray(delta, f32)
        for i in self._active_entries():
            self.aim[i] = self.aim[i] + d

    def _active_center(self):
        c = np.zeros(3, f32)
        ent = self._active_entries()
        for i in ent:
            c = (c + self.node_pos[i]).astype(f32)
        return (c / f32(len(ent))).astype(f32)

    def aim_twist(self, axis4, y):
        ent = self._active_entries()
        if len(ent) == 0:
            return
        radian = f32(0.005) * f32(int(y))
        c = self._active_center()
        for i in ent:
            self.aim[i] = O.rotate_by_axis(self.aim[i], c, np.asarray(axis4, f32), radian)

    def aim_scale(self, y):
        ent = self._active_entries()
        if len(ent) == 0:
            return
        s = f32(f32(0.002) * f32(int(y)) + f32(1.0))
        c = self._active_center()
        for i in ent:
            self.aim[i] = (c + s * (self.aim[i] - c)).astype(f32)

    def aim_set(self, aim):
        self.aim = np.ascontiguousarray(aim, f32).copy()

    # ---- one drag step (GV:1481-1522)
    def solve(self, on_center=False):
        t0 = time.perf_counter()
        self.rot, self.trans, self.stats = O.solve(self.node_pos, self.nbr, self.anc_idx, self.anc_w, self.node_static,
                                                   self.blocks, self.block_types, self.aim, on_center,
                                                   *self.w, max_iters=self.max_gn_iters)
        self.timing["solve"] = time.perf_counter() - t0
        return self.stats

    def apply(self):
        g = self.g
        t0 = time.perf_counter()
        if self.with_samples and len(self.sample_pos):
            O.lbs_points(self.sample_pos, self.smp_idx, self.smp_w, self.node_pos, self.rot, self.trans, skip=self.sample_static.astype(np.int32))
        t1 = time.perf_counter()
        if len(self.mesh_pts):
            O.lbs_points(self.mesh_pts, self.mesh_idx, self.mesh_w, self.node_pos, self.rot, self.trans)
        O.lbs_points(self.ends, self.end_idx, self.end_w, self.node_pos, self.rot, self.trans)
        nxt = self.node_pos.copy()
        O.lbs_points(nxt, self.anc_idx, self.anc_w, self.node_pos, self.rot, self.trans)
        t2 = time.perf_counter()
        O.fit_gaussians(self.ends.reshape(self.N, 18), self.scale_backup, self.gs_static, g["pos"], g["rot"], g["scale"], g["shs"])
        t3 = time.perf_counter()
        if self.with_samples and self.aim_feature is not None:
            q = O.node_quats(self.rot)
            O.rotate_sample_shs(self.smp_w.astype(f32), self.smp_idx.astype(np.int32), q, self.sample_static.astype(np.int32), self.aim_feature)
        t4 = time.perf_counter()
        self.node_pos = nxt
        self.aim = self.node_pos.copy()                      # ReloadAimPositions
        self.rot = np.tile(np.eye(3).reshape(-1), (self.M, 1))  # resetRT
        self.trans = np.zeros((self.M, 3))
        self.timing.update(samples_lbs=t1 - t0, points_lbs=t2 - t1, fit=t3 - t2, sample_sh=t4 - t3)

    def step(self, on_center=False):
        st = self.solve(on_center)
        rot, trans = self.rot.copy(), self.trans.copy()
        self.apply()
        self.last_rot, self.last_trans = rot, trans
        return st

    # ---- replay (GV:1757-1916)
    def replay(self, hist: dict, rebuild_graph=True, max_steps=None):
        if rebuild_graph:
            self.graph_build_anchors(hist["nodes"], self.k)
        blocks, types = [], []
        self.set_blocks(blocks, types)
        add_idx = move_idx = steps = 0
        for op in hist["operation_types"][:hist["total_operations"]]:
            if op < 0:
                cur = -(op + 1)
                if blocks and cur < len(blocks):
                    del blocks[cur]; del types[cur]
                    self.set_blocks(blocks, types)
            elif op == 0:
                types.append(1 if not blocks else 0)
                blocks.append(hist["block_nodes"][add_idx]); add_idx += 1
                self.set_blocks(blocks, types)
            else:
                types = list(hist["blocks_types_moves"][move_idx])
                self.set_blocks(blocks, types)
                axis = hist["twist_axis"][move_idx]
                for mv in hist["mouse_movements"][move_idx]:
                    if max_steps is not None and steps >= max_steps:
                        return steps
                    on_center = False
                    if op == 1:
                        self.aim_translate(mv); on_center = True
                    elif op == 2:
                        self.aim_twist(axis, int(mv[0]))
                    elif op == 3:
                        self.aim_scale(int(mv[0]))
                    elif op == 4:
                        self.aim_translate(mv)
                    self.step(on_center)
                    steps += 1
                move_idx += 1
        return steps

    # ---- scripts (GV:2512-2600, 2790-2816, 1918-1993)
    def run_script(self, script_id, max_steps=None):
        block1 = [0, 20, 53, 59, 63, 64, 67, 68, 145, 167, 189, 190, 192, 196, 197, 199]
        block2 = [1, 7, 21, 26, 54, 70, 80, 127, 176, 178, 179, 181, 183, 184, 185, 186]
        inter = 50
        Pi = 3.1415926535
        start = self.node_pos.copy()
        center_start = np.array([0.0, -1.5, 0.0], f32)
        aims = []
        for df in range(1, inter + 1):
            cur = []
            for n in block2:
                if script_id == 0:
                    kk = f32(5.4)
                    aa = f32(3.0 * Pi * Pi / float(kk * kk))
                    x = f32(float(f32(df)) * (-float(kk) / Pi) / float(f32(inter)))
                    y = f32(f32(aa * x) * x)
                    radian = f32(-float(f32(df)) * Pi / float(f32(inter)))
                    o = O.rotate_by_axis(start[n], center_start, np.array([0, 0, 1, 0], f32), radian)
                    o = (o + np.array([x, y, f32(0.0)], f32)).astype(f32)
                else:
                    radian = f32(float(f32(-float(f32(df)) * float(f32(1.5)))) * Pi / float(f32(inter)))
                    o = O.rotate_by_axis(start[n], center_start, np.array([0, 1, 0, 0], f32), radian)
                cur.append(o)
            aims.append(np.array(cur, f32))
        self.set_blocks([block1, block2], [0, 1])
        temp_aim = self.node_pos.copy()
        steps = 0
        for s in range(inter):
            if max_steps is not None and steps >= max_steps:
                break
            self.aim[block1] = temp_aim[block1]
            self.aim[block2] = aims[s]
            self.step(False)
            steps += 1
        return steps


def parse_deform_txt(path) -> dict:
    """Whitespace token stream of RecordDeformation (GV:4859-4935)."""
    tok = open(path).read().split()
    p = [0]

    def nxt():
        v = tok[p[0]]; p[0] += 1; return v

    h = {}
    h["nodes_on_mesh"] = int(nxt()); nxt(); n = int(nxt())
    h["nodes"] = np.array([int(nxt()) for _ in range(n)], np.int32)
    nxt(); h["total_operations"] = int(nxt()); nxt(); h["move_operations"] = int(nxt())
    nxt(); n = int(nxt()); h["operation_types"] = [int(nxt()) for _ in range(n)]
    nxt(); n = int(nxt()); h["block_nodes"] = []
    for _ in range(n):
        m = int(nxt()); h["block_nodes"].append(np.array([int(nxt()) for _ in range(m)], np.uint32))
    nxt(); n = int(nxt()); h["mouse_movements"] = []
    for _ in range(n):
        m = int(nxt()); h["mouse_movements"].append(np.array([f32(nxt()) for _ in range(3 * m)], f32).reshape(m, 3))
    nxt(); n = int(nxt()); h["blocks_types_moves"] = []
    for _ in range(n):
        m = int(nxt()); h["blocks_types_moves"].append([int(nxt()) for _ in range(m)])
    nxt(); n = int(nxt()); h["energy_on_centers"] = [int(nxt()) for _ in range(n)]
    nxt(); n = int(nxt()); h["twist_axis"] = [np.array([f32(nxt()) for _ in range(4)], f32) for _ in range(n)]
    h["trailing_tokens"] = len(tok) - p[0]
    return h
