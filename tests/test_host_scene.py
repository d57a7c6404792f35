"""Host data path: OBJ/MTL loader, SAH BVH, camera record -- byte-exact against what the
REFERENCE's own host sources produce (golden fixtures made by tests/golden/make_golden.py; live comparison when
oracle/_ref and /root/reference are present)."""
import math
import os

import numpy as np
import pytest

import yune_b200 as yb
from tests.helpers import (REF_GEOMETRY, load_golden_scene, masked_nodes_equal, random_soup, tris_equal, write_obj)


@pytest.mark.parametrize("name", ["cornellbox", "teapot"])
def test_loader_and_bvh_reproduce_golden_buffers(tmp_path, name):
    """Export the golden triangles to OBJ/MTL text, load them with the product's Scene and rebuild the BVH: every buffer
    must come back bit-identical to the reference-built golden (the builder is deterministic in the triangle list)."""
    tris, mats, nodes = load_golden_scene(name)
    obj = str(tmp_path / "scene.obj")
    write_obj(obj, tris, mats)
    s = yb.Scene().loadModel(obj)
    assert s.num_triangles == tris.size
    assert tris_equal(s.vert_data, tris)
    assert s.mat_data.tobytes() == mats.tobytes()
    assert masked_nodes_equal(nodes, s.bvh), "BVH differs from the reference's"


def test_golden_shapes_match_survey():
    """Known-good shapes from SURVEY.md 8c pin 1 (probed on the reference): 15 nodes (7 inner / 8 leaves), 2199 nodes
    (1099 / 1012 / 88 empty), max leaf 10."""
    for name, n_tris, n_nodes, kinds in (("cornellbox", 60, 15, (7, 8, 0)), ("teapot", 6380, 2199, (1099, 1012, 88))):
        tris, mats, nodes = load_golden_scene(name)
        assert tris.size == n_tris and nodes.size == n_nodes and mats.size == 7
        inner = int((nodes["child_idx"] > 0).sum()); leaf = int((nodes["child_idx"] == -1).sum()); empty = int((nodes["child_idx"] == -2).sum())
        assert (inner, leaf, empty) == kinds
        assert nodes["vert_len"].max() == 10
        np.testing.assert_allclose(nodes["p_min"][0][:3], [-1.02, -1.00383, -4.049096], rtol=0, atol=1e-6)
        np.testing.assert_allclose(nodes["p_max"][0][:3], [1.2, 1.18617, -2.019096], rtol=0, atol=1e-6)


@pytest.mark.skipif(not os.path.isdir(REF_GEOMETRY), reason="needs the reference tree")
@pytest.mark.parametrize("fn", ["cornellbox.obj", "cornellbox-teapot.obj"])
@pytest.mark.parametrize("bins", [20, 2, 8, 64])
def test_live_against_reference_host_code(ref_host, fn, bins):
    """Same .obj through the reference's Scene/BVH sources (oracle/_ref) and through ours, for several bin counts."""
    path = os.path.join(REF_GEOMETRY, fn)
    rt, rm, rn, rroot = ref_host.load(path, bins)
    s = yb.Scene().loadModel(path, bvh_bins=bins)
    assert tris_equal(s.vert_data, rt)
    assert s.mat_data.tobytes() == rm.tobytes()
    assert masked_nodes_equal(rn, s.bvh)
    assert s.root.tobytes() == rroot.tobytes()


def test_material_dialect_quirks(tmp_path):
    """Lowercase custom keys; px/py crossed; unknown usemtl -> material 0; defaults of newMaterial (appendix B#3)."""
    (tmp_path / "q.mtl").write_text("# comment\nnewmtl a\nkd 0.1 0.2 0.3\npx 7\npy 9\nis_specular 1\n\nnewmtl b\nks 0.5 0.5 0.5\nn 1.5\nalpha_x 0.25\n")
    (tmp_path / "q.obj").write_text("mtllib q.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nusemtl b\nf 1//1 2//1 3//1\nusemtl nosuch\nf 1//1 2//1 4//1\no thing\nf 2/5/1 3/5/1 4/5/1\n")
    s = yb.Scene().loadModel(str(tmp_path / "q.obj"))
    m = s.mat_data
    assert m.size == 2
    assert (m["px"][0], m["py"][0]) == (9.0, 7.0)                 # crossed
    assert m["is_specular"][0] == 1 and m["is_transmissive"][0] == 0
    np.testing.assert_array_equal(m["kd"][1], np.float32([0.3, 0.3, 0.3, 1]))     # default kd
    assert m["alpha_y"][1] == 100 and m["alpha_x"][1] == 0.25 and m["n"][1] == 1.5
    assert list(s.vert_data["matID"]) == [1, 0, 0]                 # unknown name -> 0, and 'o' does not reset it
    assert s.vert_data["v1"][0][3] == 1 and s.vert_data["vn1"][0][3] == 0
    # flat triangles get the 0.2 slab on zero-extent axes (appendix B#4): triangle 0 lies in z = 0
    assert s.bvh.size == 1 and s.bvh["vert_len"][0] == 3
    assert s.bvh["p_max"][0][2] == np.float32(1.0)                 # union with the third triangle reaches z = 1
    assert s.root[6] == np.float32(1.0)


def test_missing_mtllib_uses_default_material_without_writing(tmp_path):
    (tmp_path / "n.obj").write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    s = yb.Scene().loadModel(str(tmp_path / "n.obj"))
    assert s.mat_data.size == 1 and s.vert_data["matID"][0] == 0
    assert not (tmp_path / "n.mtl").exists()      # the reference would create it (src/Scene.cpp:139-168); we must not


def test_error_behaviour(tmp_path):
    with pytest.raises(yb.YuneError, match="Error opening"):
        yb.Scene().loadModel(str(tmp_path / "nope.obj"))
    (tmp_path / "m.obj").write_text("mtllib gone.mtl\nv 0 0 0\n")
    with pytest.raises(yb.YuneError, match="Error opening material file"):
        yb.Scene().loadModel(str(tmp_path / "m.obj"))
    (tmp_path / "e.mtl").write_text("# nothing\n")
    (tmp_path / "e.obj").write_text("mtllib e.mtl\n")
    with pytest.raises(yb.YuneError, match="Bad Material file"):
        yb.Scene().loadModel(str(tmp_path / "e.obj"))
    (tmp_path / "f.mtl").write_text("newmtl a\n")
    (tmp_path / "f.obj").write_text("mtllib f.mtl\nv 0 0 0\nv 1 0 0\nf 1 2\n")
    with pytest.raises(yb.YuneError, match="fewer than 3"):
        yb.Scene().loadModel(str(tmp_path / "f.obj"))
    # RendererCore::loadScene reports instead of raising (src/RendererCore.cpp:124-137)
    r = yb.RendererCore(yb.CUDAManager(), 8, 8)
    assert r.loadScene(str(tmp_path / "nope.obj")) is False and "Error opening" in r.cl_manager.last_message


def test_duplicate_geometry_terminates(tmp_path):
    """The reference builder never terminates on > 10 coincident centroids (appendix B#20); ours must report it."""
    (tmp_path / "d.mtl").write_text("newmtl a\n")
    lines = ["mtllib d.mtl", "v 0 0 0", "v 1 0 0", "v 0 1 0", "vn 0 0 1"] + ["f 1//1 2//1 3//1"] * 11
    (tmp_path / "d.obj").write_text("\n".join(lines) + "\n")
    with pytest.raises(yb.YuneError, match="does not terminate"):
        yb.Scene().loadModel(str(tmp_path / "d.obj"))


def test_empty_scene_and_tiny_scene(tmp_path):
    (tmp_path / "z.mtl").write_text("newmtl a\n")
    (tmp_path / "z.obj").write_text("mtllib z.mtl\n")
    s = yb.Scene().loadModel(str(tmp_path / "z.obj"))
    assert s.num_triangles == 0 and s.bvh.size == 1 and s.bvh["child_idx"][0] == -2     # root stays "empty"
    (tmp_path / "t.obj").write_text("mtllib z.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n")
    s = yb.Scene().loadModel(str(tmp_path / "t.obj"))
    assert s.bvh.size == 1 and s.bvh["child_idx"][0] == -1 and s.bvh["vert_len"][0] == 1


def test_camera_record():
    """Cam = rows of view2world + view_plane_dist = 1/tan(fov*3.14/360) evaluated in double (src/Camera.cpp:60-82, 93-103)."""
    cam = yb.default_camera()
    for i, r in enumerate(("r1", "r2", "r3", "r4")):
        want = np.zeros(4, np.float32); want[i] = 1
        np.testing.assert_array_equal(np.abs(cam[r][0]), want)
    assert cam["view_plane_dist"][0] == np.float32(1 / math.tan(60.0 * 3.14 / 360))
    assert yb.default_camera(90.0)["view_plane_dist"][0] == np.float32(1 / math.tan(90.0 * 3.14 / 360))


def _glm_rotate(angle, axis):
    """glm::rotate(mat4(1), angle, axis) (GLM's matrix_transform.inl), column-major -> returned as a numpy matrix M with M @ v."""
    a = np.asarray(axis, np.float64); a = a / np.linalg.norm(a)
    c, s = np.cos(angle), np.sin(angle)
    t = (1 - c) * a
    R = np.eye(4)
    R[0, 0] = c + t[0] * a[0]; R[1, 0] = t[0] * a[1] + s * a[2]; R[2, 0] = t[0] * a[2] - s * a[1]
    R[0, 1] = t[1] * a[0] - s * a[2]; R[1, 1] = c + t[1] * a[1]; R[2, 1] = t[1] * a[2] + s * a[0]
    R[0, 2] = t[2] * a[0] + s * a[1]; R[1, 2] = t[2] * a[1] - s * a[0]; R[2, 2] = c + t[2] * a[2]
    return R


def test_camera_set_orientation_follows_reference():
    """Camera::setOrientation (src/Camera.cpp:119-166) restated with numpy in double: key moves (one axis per call, priority
    z, x, y, step 0.1), yaw about world +Y, pitch about the local side axis and refused once up.y would turn negative."""
    import yune_b200 as yb
    cam = yb.Camera(60.0)
    side = np.array([1.0, 0, 0, 0]); up = np.array([0, 1.0, 0, 0]); look = np.array([0, 0, -1.0, 0]); eye = np.array([0, 0, 0, 1.0])
    M = np.stack([side, up, -look, eye], axis=1)               # columns
    assert not cam.is_changed or cam.setBuffer() is not None
    rng = np.random.default_rng(5)
    moves = [((0, 0, 1, 0), 0, 0), ((1, 0, 0, 0), 0, 0), ((0, -1, 0, 0), 0, 0), ((1, 1, -1, 0), 0, 0),
             ((0, 0, 0, 0), 0.0, 2.0), ((0, 0, 0, 0), 1.5, 0.0), ((0, 0, 1, 0), -0.7, 0.9), ((0, 0, 0, 0), 40.0, 0.0)]
    moves += [(tuple(rng.integers(-1, 2, 3)) + (0,), float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2))) for _ in range(12)]
    for d, pitch, yaw in moves:
        cam.setOrientation(d, pitch, yaw)
        assert cam.is_changed
        # --- restatement ---
        if d[2] > 0: eye = eye + look * 0.1
        elif d[2] < 0: eye = eye - look * 0.1
        elif d[0] > 0: eye = eye + side * 0.1
        elif d[0] < 0: eye = eye - side * 0.1
        elif d[1] > 0: eye = eye + up * 0.1
        elif d[1] < 0: eye = eye - up * 0.1
        if pitch == 0 and yaw == 0:
            M[:, 3] = eye
        else:
            rotx = _glm_rotate(pitch * 0.25, side[:3]); roty = _glm_rotate(yaw * 0.25, (0, 1, 0))
            M[:, 3] = (0, 0, 0, 1)
            if (rotx @ up)[1] >= 0: M = rotx @ M
            M = roty @ M
            M[:, 3] = eye
            side = M[:, 0] / np.linalg.norm(M[:, 0]); up = M[:, 1] / np.linalg.norm(M[:, 1]); look = -M[:, 2] / np.linalg.norm(M[:, 2])
        rec = cam.setBuffer()
        assert not cam.is_changed
        ours = np.stack([rec["r1"][0], rec["r2"][0], rec["r3"][0], rec["r4"][0]])      # rows of view2world (Cam.r1..r4)
        np.testing.assert_allclose(ours, M, atol=2e-5)
    assert up[1] >= -1e-6                                        # the pitch limit held through the 40-unit pitch
    cam.resetCamera()
    np.testing.assert_array_equal(cam.setBuffer().view(np.uint8), yb.default_camera(60.0).view(np.uint8))


def test_synthetic_c4_scene_and_in_memory_geometry(tmp_path):
    """configs[3] generator at a small subdivision: 40 wall triangles + 2 icospheres; setGeometry (no OBJ text) must build
    exactly what loadModel builds from the equivalent OBJ, and -- when the reference is present -- what the reference builds."""
    from yune_b200.scenes import synthetic_c4, icosphere
    from tests.refbind import have_ref, RefHost
    v, f = icosphere(3)
    assert v.shape == (642, 3) and f.shape == (1280, 3)
    np.testing.assert_allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-12)
    tris, mats, nodes = load_golden_scene("cornellbox")
    T = synthetic_c4(tris, 3)
    assert T.size == 40 + 2 * 1280 and (T["matID"][40:] == 3).all()
    a = yb.Scene().setGeometry(T, mats)
    obj = str(tmp_path / "c4.obj")
    write_obj(obj, T, mats)
    b = yb.Scene().loadModel(obj)
    assert tris_equal(a.vert_data, b.vert_data) and a.bvh.tobytes() == b.bvh.tobytes() and a.root.tobytes() == b.root.tobytes()
    if have_ref():
        rt, rm, rn, rroot = RefHost().load(obj)
        assert tris_equal(rt, a.vert_data) and masked_nodes_equal(rn, a.bvh)
    # spheres rest inside the box and do not touch each other
    p = np.stack([T["v1"], T["v2"], T["v3"]], 1)[40:, :, :3]
    assert p[..., 1].min() > -1.0039 and np.abs(p[: 1280].mean((0, 1)) - [-0.45, -0.55, -3.2]).max() < 1e-3


def test_threaded_bvh_build_equals_sequential_and_reference(tmp_path):
    """The builder decides the nodes of one level in parallel and splits the passes over nodes of >= 131072 primitives across
    threads; 655 400 triangles (two icospheres of 327 680) exercise both.  Output must be byte-identical to the single-thread
    build and -- masked for the reference's uninitialised bytes -- to the reference's own builder."""
    from yune_b200.scenes import synthetic_c4
    from tests.refbind import have_ref, RefHost
    tris, mats, nodes = load_golden_scene("cornellbox")
    T = synthetic_c4(tris, 7)
    assert T.size == 40 + 2 * 327680
    a = yb.Scene().setGeometry(T, mats)
    os.environ["YUNE_BVH_THREADS"] = "1"
    try:
        b = yb.Scene().setGeometry(T, mats)
    finally:
        del os.environ["YUNE_BVH_THREADS"]
    assert a.bvh.tobytes() == b.bvh.tobytes() and a.bvh.size > 90000
    if have_ref():
        obj = str(tmp_path / "c4_7.obj")
        write_obj(obj, T, mats)
        rt, rm, rn, rroot = RefHost().load(obj)
        assert tris_equal(rt, a.vert_data) and masked_nodes_equal(rn, a.bvh)


@pytest.mark.skipif(not os.path.isdir(REF_GEOMETRY), reason="needs the reference tree")
@pytest.mark.parametrize("kind,n,bins,seed", [("uniform", 900, 20, 1), ("uniform", 21, 20, 2), ("uniform", 11, 20, 3), ("uniform", 10, 20, 4),
                                              ("clustered", 1500, 20, 5), ("clustered", 700, 3, 6), ("flats", 1200, 20, 7),
                                              ("flats", 400, 2, 8), ("mixed", 2000, 20, 9), ("mixed", 800, 64, 10)])
def test_random_soups_against_reference_host_code(ref_host, tmp_path, kind, n, bins, seed):
    """Randomised geometry through the reference's compiled Scene/BVH sources and through ours: triangles, root box and every
    node (masked for the bytes the reference leaves uninitialised) must be identical -- including the leaf threshold (10 / 11
    primitives), the SAH-vs-median decision (> 20 primitives and bins > 2, src/BVH.cpp:89-128) and empty children."""
    T = random_soup(np.random.default_rng(seed), n, kind)
    mats = load_golden_scene("cornellbox")[1]
    obj = str(tmp_path / "soup.obj")
    write_obj(obj, T, mats)
    rt, rm, rn, rroot = ref_host.load(obj, bins)
    s = yb.Scene().loadModel(obj, bvh_bins=bins)
    assert tris_equal(s.vert_data, rt) and tris_equal(s.vert_data, T)
    assert s.mat_data.tobytes() == rm.tobytes()
    assert masked_nodes_equal(rn, s.bvh), "%d vs %d nodes" % (rn.size, s.bvh.size)
    assert s.root.tobytes() == rroot.tobytes()
    # every triangle sits in exactly one leaf
    leaves = s.bvh[s.bvh["child_idx"] == -1]
    ids = np.concatenate([l["vert_list"][: l["vert_len"]] for l in leaves]) if leaves.size else np.zeros(0, int)
    assert np.array_equal(np.sort(ids), np.arange(n))


QUIRK_MTL = """# materials in the reference's dialect
newmtl   first
kd 0.1 0.2 0.3
ks .5 .25 0.125
ke 1 2 3
n 1.33
k 2.5
px 7
py 9
alpha_x 0.5
alpha_y 0.75
is_specular 1
is_transmissive 1
Kd 0.9 0.9 0.9
unknown_key 4 5 6

newmtl second
kd 0.7 0.7 0.7

newmtl third
ks 1 1 1
px 40
"""

QUIRK_OBJ = """mtllib q.mtl
# comment line
o first_object
v 0 0 0
v 1 0 0
v 0 1 0
v  0.25   0.25   1.5
v -1e-3 2.5E+0 -.5
vt 0.5 0.5
vt 0 1
vn 0 0 1
vn 0 1 0
vn 0.6 0 0.8

usemtl second
f 1//1 2//1 3//1
f 1/1/2 2/2/2 4/1/3
s off
g group_name
usemtl third
f 2//3 3//3 5//3
usemtl nosuch
f 5//1 4//2 1//3
o second_object
usemtl first
f 3//2 4//2 5//2
f 1//1 3//1 4//1
"""


@pytest.mark.skipif(not os.path.isdir(REF_GEOMETRY), reason="needs the reference tree")
@pytest.mark.parametrize("newline", ["\n", "\r\n"])
def test_text_quirks_live_against_reference_loader(ref_host, tmp_path, newline):
    """OBJ / MTL text the shipped scenes do not contain -- vt lines and a/b/c corners, exponents, repeated blanks, s / g / o
    statements, an unknown usemtl, unknown and wrongly-cased MTL keys, CRLF line ends -- through the reference's compiled
    Scene.cpp and through ours: triangles, materials, BVH and root box identical."""
    (tmp_path / "q.mtl").write_bytes(QUIRK_MTL.replace("\n", newline).encode())
    (tmp_path / "q.obj").write_bytes(QUIRK_OBJ.replace("\n", newline).encode())
    path = str(tmp_path / "q.obj")
    rt, rm, rn, rroot = ref_host.load(path)
    s = yb.Scene().loadModel(path)
    assert rt.size == 6 and rm.size == 3
    assert tris_equal(s.vert_data, rt)
    assert s.mat_data.tobytes() == rm.tobytes()
    assert masked_nodes_equal(rn, s.bvh)
    assert s.root.tobytes() == rroot.tobytes()


_QBASE = "mtllib q.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 1 1 0\nvn 0 0 1\n"
SMALL_QUIRKS = {
    "quad_face_keeps_first_three_corners": _QBASE + "usemtl a\nf 1//1 2//1 4//1 3//1\n",
    "tabs": _QBASE.replace(" ", "\t") + "usemtl\ta\nf\t1//1\t2//1\t3//1\n",
    "trailing_spaces": _QBASE + "usemtl a   \nf 1//1 2//1 3//1   \n",
    "no_usemtl": _QBASE + "f 1//1 2//1 3//1\n",
    "v_with_w": "mtllib q.mtl\nv 0 0 0 1\nv 1 0 0 1\nv 0 1 0 1\nvn 0 0 1\nf 1//1 2//1 3//1\n",
    "leading_spaces": "mtllib q.mtl\n  v 0 0 0\n  v 1 0 0\n v 0 1 0\n vn 0 0 1\n  f 1//1 2//1 3//1\n",
    "comment_before_mtllib": "# c\n\nmtllib q.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//1 3//1\n",
    "usemtl_twice": _QBASE + "usemtl a\nusemtl zzz\nf 1//1 2//1 3//1\nusemtl a\nf 2//1 3//1 4//1\n",
    "vn_before_v": "mtllib q.mtl\nvn 0 0 1\nv 0 0 0\nv 1 0 0\nv 0 1 0\nf 1//1 2//1 3//1\n",
    "no_final_newline": _QBASE + "f 1//1 2//1 3//1",
    "vt_corners": _QBASE + "vt 0 0\nf 1/1/1 2/1/1 3/1/1\n",
}


@pytest.mark.skipif(not os.path.isdir(REF_GEOMETRY), reason="needs the reference tree")
@pytest.mark.parametrize("case", sorted(SMALL_QUIRKS))
def test_small_text_quirks_live_against_reference_loader(ref_host, tmp_path, case):
    (tmp_path / "q.mtl").write_text("newmtl a\nkd 0.5 0.5 0.5\n")
    (tmp_path / "q.obj").write_text(SMALL_QUIRKS[case])
    path = str(tmp_path / "q.obj")
    rt, rm, rn, rroot = ref_host.load(path)
    s = yb.Scene().loadModel(path)
    assert rt.size >= 1 and tris_equal(s.vert_data, rt) and s.mat_data.tobytes() == rm.tobytes()
    assert masked_nodes_equal(rn, s.bvh) and s.root.tobytes() == rroot.tobytes()
