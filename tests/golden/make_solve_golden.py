"""Generates tests/golden/solve_<name>.npz: node transforms of the ORACLE's Gauss-Newton solve (block-sparse Cholesky,
oracle/arap_oracle.cpp, restating Deform.cpp:95-169) at the node counts of BASELINE configs[2] (4 000) and configs[3]
(16 000) — sizes at which the oracle takes minutes, so the -m gpu test compares against these committed vectors instead
of running it.  The solve depends on the node set only: the scene is a seeded subsample of the config's cloud
(same law), nodes = FPS over it (bit-exact on both sides, checked through the anchor hash).

    python tests/golden/make_solve_golden.py sphere1m 200000 4000 2
    python tests/golden/make_solve_golden.py shells6m 400000 16000 1

Stored: anchor hash, block sizes, per step: GN iterations, halvings, energy, and rot/trans of 512 sampled nodes + the
max-norm of all transforms; node positions after the step at the sampled nodes.
"""
import hashlib
import importlib
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge  # noqa: E402

ge.load_package()
scenes = importlib.import_module(ge.PKG + ".scenes")
import oracle as O  # noqa: E402
from oracle.session import OracleSession  # noqa: E402

DRAG = np.array([0.0, 0.0, 0.002], np.float32)


def main(name, n, M, steps, grid):
    k = 10
    sc = scenes.make_scene(name, n=n)
    o = OracleSession(sc, grid_num=grid, knn_k=k, node_num=M, with_samples=False)
    o.grid_build()                                   # the cell re-order decides which Gaussian index a node anchors to
    anchors = O.fps(o.g["pos"], M)
    o.anchor, o.M = anchors, M
    o.node_pos = o.g["pos"][anchors].copy(); o.node_rest = o.node_pos.copy(); o.aim = o.node_pos.copy()
    o.nbr = O.graph_edges(o.node_rest, k)
    idx, w = O.knn_weights(o.node_rest, o.node_rest, k)
    o.anc_idx, o.anc_w = idx[:, :k].copy(), w
    o.node_static = np.zeros(M, np.uint8)
    blocks, types = scenes.cap_blocks(o.node_pos)
    o.blocks = [np.asarray(b, np.uint32) for b in blocks]; o.block_types = types
    rng = np.random.Generator(np.random.PCG64(7))
    sample = np.sort(rng.choice(M, size=min(512, M), replace=False)).astype(np.int32)
    out = dict(name=name, n=n, M=M, k=k, grid=grid, steps=0, drag=DRAG, sample=sample,
               anchor_sha1=hashlib.sha1(np.ascontiguousarray(anchors, np.int32).tobytes()).hexdigest(),
               block_sizes=np.array([len(b) for b in blocks]))
    dst = Path(__file__).resolve().parent / f"solve_{name}_{M}.npz"
    for s in range(steps):
        o.aim_translate(DRAG)
        t0 = time.time()
        st = o.solve(False)
        print(f"step {s}: {time.time() - t0:.1f} s {st}", flush=True)
        out[f"gn_{s}"] = int(st["iters"]); out[f"halvings_{s}"] = int(st["halvings"]); out[f"energy_{s}"] = float(st["energy"])
        out[f"rot_{s}"] = o.rot[sample].copy(); out[f"trans_{s}"] = o.trans[sample].copy()
        out[f"rot_absmax_{s}"] = float(np.abs(o.rot - np.eye(3).reshape(-1)).max()); out[f"trans_absmax_{s}"] = float(np.abs(o.trans).max())
        nxt = o.node_pos.copy()
        O.lbs_points(nxt, o.anc_idx, o.anc_w, o.node_pos, o.rot, o.trans)        # node positions (GV:3041-3046)
        o.node_pos = nxt; o.aim = nxt.copy()
        out[f"node_pos_{s}"] = nxt[sample].copy()
        out["steps"] = s + 1
        np.savez_compressed(dst, **out)
    print("wrote", dst)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 64)
