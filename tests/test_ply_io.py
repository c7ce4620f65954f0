"""3DGS PLY loader / saver (reference GaussianView.cpp:43-339) — host C++, no GPU needed."""
import numpy as np
import pytest


def _morton_order(pos):
    """numpy restatement of GV:91-116 (float32 arithmetic, 21 bits per axis, x in bit 3i)."""
    pos = pos.astype(np.float32)
    mn, mx = pos.min(0), pos.max(0)
    rel = (pos - mn) / (mx - mn)
    q = (np.float32((1 << 21) - 1) * rel).astype(np.int64)
    code = np.zeros(len(pos), np.uint64)
    for b in range(21):
        for c in range(3):
            code |= ((q[:, c].astype(np.uint64) >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + c)
    return np.argsort(code, kind="stable"), mn, mx


def _cloud(n, seed=7):
    r = np.random.default_rng(seed)
    q = r.normal(size=(n, 4)).astype(np.float32)
    return dict(pos=r.uniform(-1, 1, (n, 3)).astype(np.float32), rot=(q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32),
                scale=np.exp(r.uniform(-6, -3, (n, 3))).astype(np.float32), opacity=r.uniform(0.05, 0.95, n).astype(np.float32),
                shs=r.normal(size=(n, 48)).astype(np.float32))


def test_ply_header_and_record_layout(pkg, tmp_path):
    g = _cloud(5)
    p = tmp_path / "pc.ply"
    assert pkg.ply_save(p, g) == 5
    raw = p.read_bytes()
    props = ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)] + \
            ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    header = "ply\nformat binary_little_endian 1.0\nelement vertex 5\n" + "".join(f"property float {s}\n" for s in props) + "end_header\n"
    assert raw.startswith(header.encode())                                   # byte-identical to savePly's header (GV:296-309)
    rec = np.frombuffer(raw[len(header):], np.float32).reshape(5, 62)
    assert np.array_equal(rec[:, :3], g["pos"]) and np.all(rec[:, 3:6] == 0)
    assert np.array_equal(rec[:, 6:9], g["shs"][:, :3])                      # f_dc
    assert np.array_equal(rec[:, 9:24], g["shs"][:, 3::3])                   # f_rest: channel-major, 15 per channel
    assert np.array_equal(rec[:, 24:39], g["shs"][:, 4::3]) and np.array_equal(rec[:, 39:54], g["shs"][:, 5::3])
    assert np.allclose(rec[:, 54], np.log(g["opacity"] / (1 - g["opacity"])), rtol=1e-6)
    assert np.allclose(rec[:, 55:58], np.log(g["scale"]), rtol=1e-6) and np.array_equal(rec[:, 58:62], g["rot"])


def test_ply_round_trip_is_morton_ordered(pkg, tmp_path):
    g = _cloud(4000)
    p = tmp_path / "pc.ply"
    pkg.ply_save(p, g)
    out = pkg.ply_load(p)
    order, mn, mx = _morton_order(g["pos"])
    assert np.array_equal(out["aabb_min"], mn) and np.array_equal(out["aabb_max"], mx)
    assert np.array_equal(out["index"], order.astype(np.int32))              # bit-exact Morton order (integer work)
    assert np.array_equal(out["pos"], g["pos"][order]) and np.array_equal(out["shs"], g["shs"][order])
    assert np.allclose(out["scale"], g["scale"][order], rtol=2e-6) and np.allclose(out["opacity"], g["opacity"][order], rtol=2e-6, atol=1e-7)
    assert np.allclose(out["rot"], g["rot"][order], atol=2e-7)
    assert np.allclose(np.linalg.norm(out["rot"], axis=1), 1.0, atol=1e-6)


def test_ply_save_filters_and_soup_index(pkg, tmp_path):
    g = _cloud(300)
    skip = np.zeros(300, np.uint8); skip[::7] = 1
    inside = np.all((g["pos"] >= -0.5) & (g["pos"] <= 0.5), axis=1)
    p = tmp_path / "crop.ply"
    n = pkg.ply_save(p, g, box_min=[-0.5] * 3, box_max=[0.5] * 3, skip=skip)
    assert n == int((inside & (skip == 0)).sum()) and len(pkg.ply_load(p)["pos"]) == n
    # "soup" variant: a trailing index property is carried through the Morton re-order
    raw = p.read_bytes()
    head, body = raw.split(b"end_header\n", 1)
    rec = np.frombuffer(body, np.float32).reshape(n, 62)
    soup = np.concatenate([rec, (np.arange(n, dtype=np.float32) * 3 + 1)[:, None]], axis=1)
    q = tmp_path / "soup.ply"
    q.write_bytes(head + b"property float index\nend_header\n" + soup.astype(np.float32).tobytes())
    out = pkg.ply_load(q)
    order, _, _ = _morton_order(rec[:, :3])
    assert np.array_equal(out["index"], (order * 3 + 1).astype(np.int32))


def test_ply_errors(pkg, tmp_path):
    bad = tmp_path / "bad.ply"
    bad.write_bytes(b"ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    with pytest.raises(pkg.ArapError):
        pkg.ply_load(bad)
    with pytest.raises(pkg.ArapError):
        pkg.ply_load(tmp_path / "missing.ply")
