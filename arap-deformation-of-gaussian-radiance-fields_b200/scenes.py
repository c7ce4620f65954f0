"""Seeded synthetic Gaussian clouds for the BASELINE.json configs.

Both reference `point_cloud.ply` files are missing from the reference tree, so
every config is a seeded stand-in: numpy PCG64, seed = 20250401 + config index.
Attribute law (all configs): log-scale ~ N(ln s_med, 0.4^2) per axis, opacity =
sigmoid(N(2,1)), quaternion = normalised N(0, I4), SH DC ~ U(-1,1), rest ~
N(0, 0.05^2).  Values are post-activation, i.e. what the reference's PLY loader
hands to the viewer (GaussianView.cpp:120-152).
"""
from __future__ import annotations

import numpy as np

BASE_SEED = 20250401

CONFIGS = {
    # name: (config index, default N, scale median, grid, nodes, k)
    "stripes": dict(index=0, n=200_000, s_med=0.008, grid=64, nodes=200, k=10),
    "pinocchio": dict(index=1, n=30_000, s_med=0.008, grid=64, nodes=501, k=8),
    "sphere1m": dict(index=2, n=1_000_000, s_med=0.004, grid=64, nodes=4000, k=10),
    "shells6m": dict(index=3, n=6_000_000, s_med=0.002, grid=128, nodes=16000, k=10,
                     samples_at_n=58_776_512),   # 64 x valid cells of the full scene (grid_build on the GPU arm)
    # ONE scene sharded over the ranks (BASELINE configs[4]): make_scene_shard, arap_comm_grid_build
    "shells50m": dict(index=4, n=50_000_000, s_med=0.001, grid=128, nodes=16000, k=10, sharded_scene=True),
}


def _attributes(rng, n, s_med):
    scale = np.exp(rng.normal(np.log(s_med), 0.4, size=(n, 3))).astype(np.float32)
    opacity = (1.0 / (1.0 + np.exp(-rng.normal(2.0, 1.0, size=n)))).astype(np.float32)
    q = rng.normal(size=(n, 4)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    shs = np.empty((n, 48), np.float32)
    shs[:, :3] = rng.uniform(-1.0, 1.0, size=(n, 3))
    shs[:, 3:] = rng.normal(0.0, 0.05, size=(n, 45))
    return q.astype(np.float32), scale, opacity, shs


def _unit(rng, n):
    d = rng.normal(size=(n, 3))
    return d / np.linalg.norm(d, axis=1, keepdims=True)


def _positions(name, rng, n):
    if name == "stripes":
        # surface shell (thickness 0.02) of the box [-0.3,0.3] x [-1.5,1.5] x [-0.3,0.3]
        half = np.array([0.3, 1.5, 0.3])
        area = np.array([half[1] * half[2], half[0] * half[2], half[0] * half[1]])
        face = rng.choice(3, size=n, p=area / area.sum())
        p = rng.uniform(-1.0, 1.0, size=(n, 3)) * half
        sign = rng.choice([-1.0, 1.0], size=n)
        depth = rng.uniform(0.0, 0.02, size=n)
        p[np.arange(n), face] = sign * (half[face] - depth)
        return p
    if name == "pinocchio":
        # capsule-like body: a ball of radius 0.2 swept along y over [-0.4, 0.4] (bounding length 1.2)
        p = _unit(rng, n) * 0.2 * np.cbrt(rng.uniform(size=(n, 1)))
        p[:, 1] += rng.uniform(-0.4, 0.4, size=n)
        return p
    if name == "sphere1m":
        return _unit(rng, n) * (0.5 + rng.normal(0.0, 0.01, size=(n, 1)))
    if name in ("shells6m", "shells50m"):
        nv = n // 10
        ns = n - nv
        radii = rng.choice([0.3, 0.5, 0.7], size=ns, p=np.array([0.09, 0.25, 0.49]) / 0.83)
        shell = _unit(rng, ns) * (radii[:, None] + rng.normal(0.0, 0.005, size=(ns, 1)))
        vol = _unit(rng, nv) * 0.7 * np.cbrt(rng.uniform(size=(nv, 1)))
        p = np.concatenate([shell, vol])
        return p[rng.permutation(n)]
    raise KeyError(name)


def make_scene(name: str, n: int | None = None, seed_offset: int = 0):
    """-> dict(pos, rot, scale, opacity, shs) float32 arrays + config metadata."""
    cfg = CONFIGS[name]
    n = int(n or cfg["n"])
    rng = np.random.Generator(np.random.PCG64(BASE_SEED + cfg["index"] + 1000 * seed_offset))
    pos = _positions(name, rng, n).astype(np.float32)
    rot, scale, opacity, shs = _attributes(rng, n, cfg["s_med"])
    return dict(pos=np.ascontiguousarray(pos), rot=rot, scale=scale, opacity=opacity, shs=shs, name=name,
                grid=cfg["grid"], nodes=cfg["nodes"], k=cfg["k"], n=n)


def _attributes32(rng, n, s_med):
    """Same law as _attributes, float32 draws (the 50M-Gaussian scene: 2.4 G normal variates)."""
    scale = np.exp(np.float32(np.log(s_med)) + np.float32(0.4) * rng.standard_normal((n, 3), dtype=np.float32))
    opacity = (1.0 / (1.0 + np.exp(-(np.float32(2.0) + rng.standard_normal(n, dtype=np.float32))))).astype(np.float32)
    q = rng.standard_normal((n, 4), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    shs = np.empty((n, 48), np.float32)
    shs[:, :3] = rng.random((n, 3), dtype=np.float32) * np.float32(2.0) - np.float32(1.0)
    shs[:, 3:] = np.float32(0.05) * rng.standard_normal((n, 45), dtype=np.float32)
    return q, scale.astype(np.float32), opacity, shs


SHARD_BLOCKS = 64      # attribute streams of a sharded scene: one per 1/64 of the globally ordered Gaussians
SHARD_CHUNK = 2_000_000


def make_scene_shard(name: str, n_total: int, rank: int, world: int):
    """Rank `rank`'s contiguous part [rank n, (rank + 1) n) of ONE scene of n_total Gaussians in global cell order (x-major cells
    of a fixed 128^3 grid over [-0.8, 0.8]^3: what a single-GPU grid build would leave behind, up to its data-dependent box).
    The scene does not depend on `world`: positions come from per-chunk streams (every rank draws all of them and sorts the
    cell keys), the i.i.d. attributes from one stream per 1/64 of the ordered Gaussians (a rank draws only its own)."""
    cfg = CONFIGS[name]
    n_total = int(n_total)
    assert SHARD_BLOCKS % world == 0 and n_total % SHARD_BLOCKS == 0, "n_total must divide into 64 blocks, world into 64"
    n = n_total // world
    pos = np.empty((n_total, 3), np.float32)
    for c, lo in enumerate(range(0, n_total, SHARD_CHUNK)):
        m = min(SHARD_CHUNK, n_total - lo)
        rng = np.random.Generator(np.random.PCG64(BASE_SEED + cfg["index"] + 1000 * (100 + c)))
        pos[lo:lo + m] = _positions(name, rng, m).astype(np.float32)
    G = 128
    cell = np.clip(np.floor((pos + np.float32(0.8)) * np.float32(G / 1.6)), 0, G - 1).astype(np.int32)
    key = (cell[:, 0] * G + cell[:, 1]) * G + cell[:, 2]
    del cell
    order = np.argsort(key, kind="stable")
    del key
    own = order[rank * n:(rank + 1) * n]
    del order
    out = dict(pos=np.ascontiguousarray(pos[own]), name=name, grid=cfg["grid"], nodes=cfg["nodes"], k=cfg["k"], n=n, n_total=n_total)
    del pos
    blk = n_total // SHARD_BLOCKS
    parts = []
    for b in range(rank * n // blk, (rank + 1) * n // blk):
        rng = np.random.Generator(np.random.PCG64(BASE_SEED + cfg["index"] + 1000 * (5000 + b)))
        parts.append(_attributes32(rng, blk, cfg["s_med"]))
    for i, kk in enumerate(("rot", "scale", "opacity", "shs")):
        out[kk] = np.ascontiguousarray(np.concatenate([p[i] for p in parts]))
    return out


def cap_blocks(node_pos: np.ndarray, axis: int = 2, lo: float = -0.4, hi: float = 0.4):
    """Two box-selected caps: nodes above `hi` (active, type 1) and below `lo` (pinned, type 0)."""
    active = np.nonzero(node_pos[:, axis] > hi)[0].astype(np.uint32)
    pinned = np.nonzero(node_pos[:, axis] < lo)[0].astype(np.uint32)
    return [active, pinned], [1, 0]
