"""Host side (C++ mirror of Obj.hs / BIH.hs / Lib.hs `render` plumbing) against the oracle.  CPU only."""
import json
import os
import subprocess

import numpy as np
import pytest

import pysqt
from oracle import oracle as O
from pysqt import scenes
from golden.make_golden import sha
from common import build_pair

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "scene_obj_golden.json")))


def flat_nodes_as_u32(hs):
    return np.stack([hs.nodes["lmax"].view(np.uint32), hs.nodes["rmin"].view(np.uint32), hs.nodes["a"], hs.nodes["b"]], 1)


def test_parser_and_bih_match_oracle_and_golden(host_scene, oracle_scene):
    v9, mi = host_scene.parsed_tris()
    ov9, omi = oracle_scene.tris()
    assert np.array_equal(v9.view(np.uint32), ov9.view(np.uint32)) and np.array_equal(mi, omi)
    assert sha(v9) == GOLD["sha_tris"]
    assert host_scene.stats() == GOLD["bih"]
    root, nodes, leaf = oracle_scene.export_bih()
    assert np.array_equal(host_scene.root, root)
    assert np.array_equal(flat_nodes_as_u32(host_scene), nodes) and sha(flat_nodes_as_u32(host_scene)) == GOLD["sha_nodes"]
    assert np.array_equal(host_scene.tris["orig_index"], leaf.astype(np.uint32))


def test_flattened_triangles_carry_reference_edges(host_scene):
    """e1 = v1 - v0, e2 = v2 - v0 in binary32 (Geometry.hs:130-131), leaf (`flatten`) order, original index kept."""
    v9, mi = host_scene.parsed_tris()
    t = host_scene.tris
    src = v9[t["orig_index"]].reshape(-1, 3, 3)
    assert np.array_equal(t["v0"], src[:, 0])
    assert np.array_equal(t["e1"], src[:, 1] - src[:, 0]) and np.array_equal(t["e2"], src[:, 2] - src[:, 0])
    assert np.array_equal(t["material"], mi[t["orig_index"]].astype(np.uint32))
    assert sorted(t["orig_index"].tolist()) == list(range(host_scene.n_tris))
    # every leaf owns a contiguous range and the ranges tile [0, n_tris)
    leaves = host_scene.nodes[(host_scene.nodes["b"] & 0x80000000) != 0]
    first, cnt = leaves["a"].astype(np.int64), (leaves["b"] & 0x7fffffff).astype(np.int64)
    order = np.argsort(first, kind="stable")
    assert np.array_equal(np.cumsum(cnt[order]) - cnt[order], first[order]) and cnt.sum() == host_scene.n_tris


@pytest.mark.parametrize("gen,n", [("cornell", 3000), ("soup", 20000), ("mesh", 30000)])
def test_bih_builder_matches_oracle_on_synthetic_scenes(gen, n):
    v9, mi, mats = {"cornell": scenes.cornell_box, "soup": scenes.triangle_soup, "mesh": scenes.subdivided_mesh}[gen](n)
    osc, hs = build_pair(v9, mi, mats)
    root, nodes, leaf = osc.export_bih()
    assert np.array_equal(hs.root, root) and np.array_equal(flat_nodes_as_u32(hs), nodes)
    assert np.array_equal(hs.tris["orig_index"], leaf.astype(np.uint32))


def test_parallel_bih_build_is_bit_identical(monkeypatch):
    """Above 2^17 triangles the host builder splits the top of the tree on one thread and builds the subtrees
    concurrently; the spliced tree must equal the sequential build and the oracle's literal build bit for bit."""
    v9, mi, mats = scenes.triangle_soup(150_000, seed=11)
    v9 = v9 * np.float32(8.0)
    monkeypatch.setenv("SQT_BIH_THREADS", "6")
    par = pysqt.HostScene.from_arrays(v9, mi, mats)
    monkeypatch.setenv("SQT_BIH_THREADS", "1")
    seq = pysqt.HostScene.from_arrays(v9, mi, mats)
    osc = O.Scene.from_arrays(v9, mi, mats)
    osc.make_bih()
    root, nodes, leaf = osc.export_bih()
    for hs in (par, seq):
        assert np.array_equal(hs.root, root) and np.array_equal(flat_nodes_as_u32(hs), nodes)
        assert np.array_equal(hs.tris["orig_index"], leaf.astype(np.uint32))


def test_degenerate_builds():
    mats = np.array([[0, .5, .5, .5, 0, 0, 0, 0]], np.float32)
    # fewer than leafLimit triangles: a single Leaf (BIH.hs:69)
    v9 = np.random.default_rng(0).normal(size=(14, 9)).astype(np.float32)
    osc, hs = build_pair(v9, np.zeros(14, np.int32), mats)
    assert hs.n_nodes == 1 and hs.nodes["b"][0] == (14 | 0x80000000)
    # identical centroids: nothing is < the split plane -> Branch (Leaf empty) (Leaf everything)  (BIH.hs:70-72)
    v9 = np.tile(np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32), (20, 1))
    osc, hs = build_pair(v9, np.zeros(20, np.int32), mats)
    # (the sequential binary32 mean may round above or below the common centroid, so either side can be the empty one)
    assert hs.n_nodes == 3 and sorted(int(x & 0x7fffffff) for x in hs.nodes["b"][1:]) == [0, 20]
    root, nodes, leaf = osc.export_bih()
    assert np.array_equal(flat_nodes_as_u32(hs), nodes)
    # no triangles at all
    osc, hs = build_pair(np.zeros((0, 9), np.float32), np.zeros(0, np.int32), mats)
    assert hs.n_nodes == 1 and hs.n_tris == 0


def test_camera_and_params(camera):
    assert np.array_equal(camera, O.load_camera(os.path.join(pysqt.ROOT, "data", "camera")))
    for lit in (True, False):
        p, q = pysqt.make_params(1920, 1080, 7, max_depth=5, seed=3, literal=lit), O.make_params(1920, 1080, 7, max_depth=5, seed=3, literal=lit)
        assert (p.rows, p.cols, p.xdiv, p.ydiv, p.seed_stride, p.spp, p.max_depth, p.seed) == (
            q.rows, q.cols, q.xdiv, q.ydiv, q.seed_stride, q.spp, q.max_depth, q.seed)


def test_parser_error_behaviour(tmp_path):
    bad = tmp_path / "bad.obj"
    bad.write_text("o Cube\nv 0 0 0\n")                 # no mtllib: the reference dies on an irrefutable pattern (Obj.hs:51)
    with pytest.raises(pysqt.SqtError, match="Irrefutable pattern"):
        pysqt.HostScene.load(str(bad), str(tmp_path))
    cam = tmp_path / "camera"
    cam.write_text("0 7\n")
    with pytest.raises(pysqt.SqtError, match="Failed to parse /data/camera"):        # Obj.hs:65
        pysqt.load_camera(str(cam))
    with pytest.raises(pysqt.SqtError, match="does not exist"):
        pysqt.HostScene.load(str(tmp_path / "missing.obj"), str(tmp_path))
    obj = tmp_path / "ok.obj"
    obj.write_text("mtllib m.sq\no A\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl M\ns off\nf 1 2 3\nf 1 2 9\n")
    (tmp_path / "m.sq").write_text("newmtl M\nreflective 0.5 1 1 1\nemissive 0 0 0 0\n")
    with pytest.raises(pysqt.SqtError, match="index too large"):                    # vs !! (a-1)
        pysqt.HostScene.load(str(obj), str(tmp_path))


def test_objects_with_unknown_material_are_dropped_and_yz_swapped(tmp_path):
    obj = tmp_path / "s.obj"
    obj.write_text("mtllib m.sq\no A\nv 0 1 2\nv 1 0 0\nv 0 0 1\nusemtl M\ns off\nf 1 2 3\no B\nv 5 5 5\nusemtl Nope\nf 1 2 4\n")
    (tmp_path / "m.sq").write_text("newmtl M\nreflective 0.25 0.1 0.2 0.3\nemissive 2 1 1 1\n\nnewmtl Other\nreflective 0 0 0 0\nemissive 0 0 0 0\n")
    hs = pysqt.HostScene.load(str(obj), str(tmp_path))
    osc = O.Scene.load(str(obj), str(tmp_path))
    v9, mi = hs.parsed_tris()
    assert hs.n_tris == 1 == osc.n_tris and np.array_equal(v9, osc.tris()[0])
    assert list(v9[0][:3]) == [0, 2, 1]                  # swapYZ (Obj.hs:112-113)
    assert hs.mats["reflective"][0] == np.float32(0.25) and hs.n_mats == 2


def test_png_writer_roundtrip(tmp_path):
    from PIL import Image
    img = np.random.default_rng(1).integers(0, 256, (37, 53, 3), dtype=np.uint8)
    path = str(tmp_path / "out.png")
    pysqt.write_png(path, img)
    assert np.array_equal(np.asarray(Image.open(path)), img)
    with pytest.raises(pysqt.SqtError, match="unsupported image format"):
        pysqt.write_png(str(tmp_path / "out.bmp"), img)


def test_cli_fails_loudly_without_gpu_or_runs(tmp_path):
    """The CLI keeps the reference's flags (Main.hs:13-33).  Without a B200 it must fail with the no-device error,
    never fall back to a CPU render."""
    exe = os.path.join(pysqt.ROOT, "squigly-trace_b200", "squigly-trace")
    out = str(tmp_path / "r.png")
    r = subprocess.run([exe, "-s", "2", "-d", "32,32", "-p", out, "--objpath", os.path.join(pysqt.ROOT, "data", "scene.obj")],
                       cwd=pysqt.ROOT, capture_output=True, text=True)
    if r.returncode == 0:
        assert os.path.exists(out) and "Took" in r.stdout
    else:
        assert "no CUDA device" in r.stderr or "compute capability" in r.stderr
        assert not os.path.exists(out)
    h = subprocess.run([exe, "--help"], capture_output=True, text=True)
    for flag in ("--samples", "--dimensions", "--savepath", "--objpath", "--camerapath", "--debug", "--debugpath", "--cast"):
        assert flag in h.stdout


def test_obj_parser_is_as_strict_as_parsec_and_read(tmp_path):
    """Obj.hs:96-147 through parsec + `read`: text starting with `s` that is neither `s on` nor `s off` is a parse error
    (parsec's `string` consumes before it fails), `1.` / `.5` are `Prelude.read: no parse`; trailing input after the last
    object is ignored (no `eof` in the reference) -- here with a warning."""
    import pysqt
    (tmp_path / "m.sq").write_text("newmtl A\nreflective 0.5 1 1 1\nemissive 0 0 0 0\n")
    base = "mtllib m.sq\no X\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl A\n%sf 1 2 3\n"

    def load(text):
        (tmp_path / "s.obj").write_text(text)
        return pysqt.HostScene.load(str(tmp_path / "s.obj"), str(tmp_path))
    assert load(base % "s off\n").n_tris == 1 and load(base % "s on\n").n_tris == 1 and load(base % "").n_tris == 1
    for bad in (base % "s 1\n", (base % "").replace("v 1 0 0", "v 1. 0 0"), (base % "").replace("v 1 0 0", "v .5 0 0"),
                (base % "").replace("v 1 0 0", "v -.5 0 0")):
        with pytest.raises(pysqt.SqtError, match="Irrefutable pattern"):
            load(bad)
    assert load(base % "" + "garbage\n").n_tris == 1
