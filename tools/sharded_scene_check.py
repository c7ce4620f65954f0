#!/usr/bin/env python
"""Multi-GPU check of the sharded-scene path (arap_comm_grid_build) against the single-GPU grid.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_scene_check.py

Every rank builds the single-GPU reference of the whole (small) scene itself, then holds part `rank` of the cell-ordered
Gaussians, joins the communicator and builds its x-slab of the one grid.  Checked per rank, bit for bit: valid cells, per-cell
lists (global Gaussian indices, halo Gaussians of the other ranks included), sample positions, the evaluated field — at rest and,
after a drag with arap_comm_exchange every step, at the stroke end (lists rebuilt over the gathered, deformed Gaussians; remote SH
rows brought up to date from the gathered rotations: compared with a tolerance).  The union of the slabs is the whole grid."""
import importlib, os, sys
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as ge


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = ge.load_package()
    scenes = importlib.import_module(ge.PKG + ".scenes")
    n = 40000 - 40000 % world
    sc = scenes.make_scene("sphere1m", n=n)
    kw = dict(grid_num=32, knn_k=10, node_num=200)
    full = pkg.Session(device=local, **kw)
    full.set_gaussians(sc["pos"], sc["rot"], sc["scale"], sc["opacity"], sc["shs"])
    gi = full.grid_build(); full.grid_eval(0)
    ordered = full.download_gaussians()
    g = full.graph_build_fps()
    blocks, types = scenes.cap_blocks(g["node_pos"], lo=-0.3, hi=0.3)
    ref = full.download_grid(); rf, ro = full.download_features(0)
    full.set_blocks(blocks, types)
    for _ in range(3):
        full.aim_translate([0.0, 0.01, 0.02]); full.step(False)
    full.grid_update_lists(); full.grid_eval(1)
    ref2 = full.download_grid(); rf2, ro2 = full.download_features(1)
    deformed = full.download_gaussians()

    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    m = n // world
    own = {kk: np.ascontiguousarray(ordered[kk][rank * m:(rank + 1) * m]) for kk in ("pos", "rot", "scale", "opacity", "shs")}
    s = pkg.Session(device=local, **kw)
    s.set_gaussians(own["pos"], own["rot"], own["scale"], own["opacity"], own["shs"])
    s.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    si = s.comm_grid_build()
    lo, hi = s.comm_slab()
    s.grid_eval(0)
    d = s.download_grid(); f, o = s.download_features(0)
    G = gi["grid_num"]
    sel = np.nonzero((ref["valid"] // (G * G) >= lo) & (ref["valid"] // (G * G) < hi))[0]
    rows = (sel[:, None] * 64 + np.arange(64)[None, :]).reshape(-1)
    assert np.array_equal(d["valid"], ref["valid"][sel]), "valid cells"
    assert np.array_equal(d["sample_pos"], ref["sample_pos"][rows]), "sample positions"
    rp = np.concatenate([[0], ref["prefix"]]); dp = np.concatenate([[0], d["prefix"]])
    for c in d["valid"]:
        assert np.array_equal(d["lists"][dp[c]:dp[c + 1]], ref["lists"][rp[c]:rp[c + 1]]), ("list of cell", c)
    remote = int(((d["lists"] // m) != rank).sum())
    assert np.array_equal(f, rf[rows]) and np.array_equal(o, ro[rows]), "field at rest"
    # drag: nodes are the FPS nodes of the whole scene (positions handed over: nodes on "mesh"), replicated solve
    s.set_mesh_points(g["node_pos"], True)
    s.graph_build_fps()
    s.set_blocks(blocks, types)
    for _ in range(3):
        s.aim_translate([0.0, 0.01, 0.02]); s.step(False); s.comm_exchange()
    s.grid_update_lists(); s.grid_eval(1)
    d2 = s.download_grid(); f2, o2 = s.download_features(1)
    out = s.download_gaussians()
    for kk in ("pos", "rot", "scale"):
        assert np.array_equal(out[kk], deformed[kk][rank * m:(rank + 1) * m]), ("deformed", kk)
    assert np.array_equal(d2["sample_pos"], ref2["sample_pos"][rows]), "advected samples"
    for c in d2["valid"]:
        assert np.array_equal(d2["lists"][np.concatenate([[0], d2["prefix"]])[c]:d2["prefix"][c]], ref2["lists"][np.concatenate([[0], ref2["prefix"]])[c]:ref2["prefix"][c]]), ("stroke-end list of cell", c)
    err = float(np.abs(f2 - rf2[rows]).max()); erro = float(np.abs(o2 - ro2[rows]).max())
    assert err <= 2e-5 and erro == 0.0, (err, erro)      # remote SH rows: one rotation by the accumulated quaternion instead of three
    tot = torch.tensor([si["valid_cells"], si["pairs"]], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    assert int(tot[0]) == gi["valid_cells"], (int(tot[0]), gi["valid_cells"])
    print(f"rank {rank}/{world}: slab [{lo}, {hi}) {si['valid_cells']} valid cells, {si['pairs']} list pairs ({remote} of them Gaussians of other ranks); "
          f"rest state bit-identical; stroke end: lists identical, field max abs diff {err:.2e}; union {int(tot[0])} = {gi['valid_cells']} cells, "
          f"pairs {int(tot[1])} vs {gi['pairs']} single-GPU", flush=True)
    s.close(); full.close()

    # ---- transport check: the exchange fused into the apply kernel (arap_comm_set_mode(1): peer stores over NVLink, epoch flags)
    # against the NCCL all-gather, both on the fused tolerance-mode apply (lbs_mode 3): gathered arrays, stroke-end lists and
    # field must be identical bit for bit — only the transport differs.
    par = importlib.import_module(ge.PKG + ".parallel")

    def run(mode):
        t = pkg.Session(device=local, lbs_mode=3, **kw)
        t.set_gaussians(own["pos"], own["rot"], own["scale"], own["opacity"], own["shs"])
        ib = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ib.copy_(torch.frombuffer(bytearray(pkg.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ib, 0)
        t.comm_init(bytes(ib.cpu().numpy().tobytes()), rank, world)
        t.comm_set_mode(mode)
        t.comm_grid_build(); t.grid_eval(0)
        t.set_mesh_points(g["node_pos"], True); t.graph_build_fps(); t.set_blocks(blocks, types)
        for step in range(5):
            t.aim_translate([0.0, 0.01, 0.02]); t.step(False); t.comm_exchange()
            if step == 2:                       # a consumer of remote rows in mid-stroke: the next pushes must wait for it
                t.comm_materialize_sh()
        t.comm_sync()
        gv = t.comm_view()
        got = {name: torch.as_tensor(par.DevArray(getattr(gv, name), (m * world, w)), device="cuda").clone().cpu().numpy() for name, w in (("pos", 3), ("rot", 4), ("scale", 3))}
        t.grid_update_lists(); t.grid_eval(1)
        dg = t.download_grid(); ff, oo = t.download_features(1)
        t.sync(); dist.barrier()
        t.close()
        return got, dg, ff, oo
    a_got, a_dg, a_f, a_o = run(0)
    for mode, label in ((1, "fused peer-store exchange"), (2, "fused NVSwitch-multicast exchange")):
      try:
        b_got, b_dg, b_f, b_o = run(mode)
      except pkg.ArapError as e:
        print(f"rank {rank}/{world}: mode {mode} not available here: {e}", flush=True)
        continue
      for kk in a_got:
        assert np.array_equal(a_got[kk], b_got[kk]), ("gathered", mode, kk)
        assert not np.array_equal(a_got[kk][(1 - rank) * m:(2 - rank) * m] if world == 2 else a_got[kk], ordered[kk][(1 - rank) * m:(2 - rank) * m] if world == 2 else ordered[kk]), "remote range did not move"
      for kk in ("valid", "prefix", "lists", "sample_pos"):
        assert np.array_equal(a_dg[kk], b_dg[kk]), ("stroke-end grid", mode, kk)
      assert np.array_equal(a_f, b_f) and np.array_equal(a_o, b_o), ("stroke-end field", mode)
      print(f"rank {rank}/{world}: {label} == NCCL all-gather (gathered pos / rot / scale after 5 steps, stroke-end lists and field: bit-identical)", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
