"""Oracle + host logic pinned against the reference's own fixtures (SURVEY 8(c) known-answer tests). CPU only."""
import numpy as np

from oracle.session import parse_deform_txt

BLOCK1 = [0, 20, 53, 59, 63, 64, 67, 68, 145, 167, 189, 190, 192, 196, 197, 199]   # GaussianView.cpp:2517
BLOCK2 = [1, 7, 21, 26, 54, 70, 80, 127, 176, 178, 179, 181, 183, 184, 185, 186]   # GaussianView.cpp:2518


def _obj_points(path):
    pts = [[np.float32(t) for t in l[2:].split()[:3]] for l in open(path) if l[:2] == "v "]
    return np.array(pts, np.float32)


def test_fps_known_answer_on_graph_obj(orc, golden):
    """LoadDeformScript0's hard-coded node blocks are FPS-order indices: block1 must be the y=+1.5 cap, block2 the y=-1.5 cap."""
    pts = _obj_points(golden / "stripes_graph.obj")
    assert pts.shape == (200, 3)
    idx = orc.fps(pts, 200)
    assert sorted(idx.tolist()) == list(range(200))
    assert np.all(pts[idx[BLOCK1], 1] > 1.49)
    assert np.all(pts[idx[BLOCK2], 1] < -1.49)


def test_fps_clamps_and_first_node(orc):
    rng = np.random.default_rng(0)
    p = rng.normal(size=(50, 3)).astype(np.float32)
    idx = orc.fps(p, 80)                       # node_num > N is clamped (helper.cpp:143-146)
    assert len(idx) == 50 and len(set(idx.tolist())) == 50
    s = (p[:, 0] + p[:, 1]) + p[:, 2]          # FetchFirstNodeIdx: float (x+y)+z, first maximum
    assert idx[0] == int(np.argmax(s))


def test_deform_txt_known_answer(golden):
    h = parse_deform_txt(golden / "pinocchio_deform.txt")
    assert h["nodes_on_mesh"] == 0 and len(h["nodes"]) == 501 and h["nodes"].max() == 29548
    assert h["total_operations"] == 7 and h["move_operations"] == 3
    assert h["operation_types"] == [0, 0, 0, 0, 4, 4, 4]
    assert [len(b) for b in h["block_nodes"]] == [211, 281, 7, 2]
    assert [len(m) for m in h["mouse_movements"]] == [244, 51, 45]
    assert h["blocks_types_moves"] == [[-1, -1, -1, 1]] * 3
    assert h["energy_on_centers"] == [0, 0, 0] and len(h["twist_axis"]) == 3
    assert h["trailing_tokens"] == 0


def test_host_history_parser_matches_and_roundtrips(pkg, golden, tmp_path):
    """The product's C++ deform.txt reader/writer (host-only, no GPU) against the oracle parser; byte-compatible rewrite."""
    src = golden / "pinocchio_deform.txt"
    h = pkg.History.load(src)
    ref = parse_deform_txt(src)
    s = h.summary()
    assert s == dict(nodes_on_mesh=0, n_nodes=501, total_ops=7, move_ops=3, n_blocks=4, n_moves=3)
    assert np.array_equal(h.nodes(), ref["nodes"]) and h.ops().tolist() == ref["operation_types"]
    for i in range(4):
        assert np.array_equal(h.block(i), ref["block_nodes"][i])
    for i in range(3):
        m = h.move(i)
        assert np.array_equal(m["movements"], ref["mouse_movements"][i])
        assert m["block_types"].tolist() == ref["blocks_types_moves"][i]
        assert np.array_equal(m["twist_axis"], ref["twist_axis"][i])
    out = tmp_path / "deform_out.txt"
    h.save(out)
    assert open(out).read().split() == open(src).read().split()   # same token stream (RecordDeformation formatting)
    assert open(out).read() == open(src).read()


def test_history_builder(pkg, tmp_path):
    h = pkg.History.new(0, [5, 3, 9])
    h.add_block([0, 1])
    h.add_block([2])
    h.add_move(2, [[10, 0, 0], [12, 0, 0]], [1, 0], 0, (0, 1, 0, 0))
    p = tmp_path / "h.txt"
    h.save(p)
    r = parse_deform_txt(p)
    assert r["operation_types"] == [0, 0, 2] and r["total_operations"] == 3 and r["move_operations"] == 1
    assert r["mouse_movements"][0].tolist() == [[10, 0, 0], [12, 0, 0]] and r["twist_axis"][0].tolist() == [0, 1, 0, 0]


def test_config_and_graph_obj_loaders(pkg, golden, tmp_path):
    assert pkg.config_load(golden / "point_cloud_config.txt") == dict(grid_num=64, is_synthetic=1, has_soup=0, high_quality=0)
    (tmp_path / "c.txt").write_text("128 1 2 ")
    assert pkg.config_load(tmp_path / "c.txt") == dict(grid_num=128, is_synthetic=1, has_soup=0, high_quality=1)
    pts = pkg.graph_obj_load(golden / "stripes_graph.obj")
    assert np.array_equal(pts, _obj_points(golden / "stripes_graph.obj"))
    pkg.graph_obj_save(tmp_path / "g.obj", pts[:5])
    assert pkg.graph_obj_load(tmp_path / "g.obj").shape == (5, 3)


def test_errors_are_reported_not_fatal(pkg, tmp_path):
    import pytest
    with pytest.raises(pkg.ArapError) as e:
        pkg.History.load(tmp_path / "missing.txt")
    assert e.value.code == 4
    (tmp_path / "bad.txt").write_text("0 Nodes: 3 1 2")
    with pytest.raises(pkg.ArapError):
        pkg.History.load(tmp_path / "bad.txt")


def test_untrusted_deform_txt_is_rejected_with_a_status(pkg, golden, tmp_path):
    """deform.txt is untrusted input: inconsistent per-move sections (e.g. fewer Twist_Axis rows than moves, which the replay
    would index out of bounds) and absurd counts must end in ARAP_ERR_IO, never in a crash across the C boundary."""
    import pytest
    tok = open(golden / "pinocchio_deform.txt").read().split()
    i = tok.index("Twist_Axis:")
    short = tok[:i] + ["Twist_Axis:", "1"] + tok[i + 2:i + 6]                 # 3 moves, 1 axis
    (tmp_path / "short_axis.txt").write_text(" ".join(short) + " ")
    with pytest.raises(pkg.ArapError) as e:
        pkg.History.load(tmp_path / "short_axis.txt")
    assert e.value.code == 4 and "per-move sections" in str(e.value)
    huge = list(tok); huge[tok.index("Block_Nodes:") + 2] = "2000000000"      # block size far beyond the file size
    (tmp_path / "huge.txt").write_text(" ".join(huge) + " ")
    with pytest.raises(pkg.ArapError) as e:
        pkg.History.load(tmp_path / "huge.txt")
    assert e.value.code == 4
    huge2 = list(tok); huge2[tok.index("Nodes:") + 1] = "-5"
    (tmp_path / "neg.txt").write_text(" ".join(huge2) + " ")
    with pytest.raises(pkg.ArapError):
        pkg.History.load(tmp_path / "neg.txt")
    ops = list(tok); j = tok.index("Operation_Types:"); ops[j + 2] = "9"      # unknown op code
    (tmp_path / "ops.txt").write_text(" ".join(ops) + " ")
    with pytest.raises(pkg.ArapError):
        pkg.History.load(tmp_path / "ops.txt")
