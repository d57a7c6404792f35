"""Shared helpers of the test-suite (scene fixtures, OBJ export, image statistics)."""
import os

import numpy as np

from tests.refbind import load_golden_scene, TRI_DTYPE, MAT_DTYPE, NODE_DTYPE  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_GEOMETRY = "/root/reference/geometry"


def transmissive(tris):
    """Config C2: teapot triangles (material 3 'teapot') use material 4 'mirror-transmissive'."""
    t = tris.copy()
    t["matID"] = np.where(t["matID"] == 3, 4, t["matID"])
    return t


def golden_scene_object(name, transmissive_teapot=False):
    """A yune_b200.Scene filled from the golden buffers (what the reference's host code produces for the shipped scenes)."""
    import yune_b200 as yb
    s = yb.Scene()
    tris, mats, nodes = load_golden_scene(name)
    if transmissive_teapot:
        tris = transmissive(tris)
    s.vert_data, s.mat_data, s.bvh = tris, mats, nodes
    s.num_triangles = int(tris.size)
    return s


def masked_nodes_equal(a, b):
    """Byte comparison of BVHNodeGPU arrays ignoring the slots the reference leaves uninitialised (appendix B#2)."""
    if a.shape != b.shape:
        return False
    ok = (a["p_min"] == b["p_min"]).all() and (a["p_max"] == b["p_max"]).all() and (a["child_idx"] == b["child_idx"]).all() \
        and (a["vert_len"] == b["vert_len"]).all()
    n = np.maximum(a["vert_len"], 0)
    for j in range(10):
        m = j < n
        ok = ok and (a["vert_list"][m, j] == b["vert_list"][m, j]).all()
    return bool(ok)


def tris_equal(a, b):
    """TriangleGPU comparison ignoring the never-written pad (appendix B#1)."""
    return all((a[f] == b[f]).all() for f in ("v1", "v2", "v3", "vn1", "vn2", "vn3", "matID")) and a.shape == b.shape


def write_obj(path, tris, mats, mtl_name="scene.mtl", mat_names=None):
    """Serialise device-layout triangles back to OBJ + the reference's MTL dialect with round-trip-exact floats (%.9g)."""
    d = os.path.dirname(path)
    names = mat_names or ["m%d" % i for i in range(mats.size)]
    with open(os.path.join(d, mtl_name), "w") as f:
        for i, m in enumerate(mats):
            f.write("newmtl %s\n" % names[i])
            for key in ("ke", "kd", "ks"):
                f.write("%s %.9g %.9g %.9g\n" % ((key,) + tuple(float(x) for x in m[key][:3])))
            # the reference's loader crosses px and py (src/Scene.cpp:212-215): 'px' fills py
            f.write("n %.9g\nk %.9g\npx %.9g\npy %.9g\nalpha_x %.9g\nalpha_y %.9g\nis_specular %d\nis_transmissive %d\n\n" % (
                m["n"], m["k"], m["py"], m["px"], m["alpha_x"], m["alpha_y"], m["is_specular"], m["is_transmissive"]))
    with open(path, "w") as f:
        f.write("mtllib %s\n" % mtl_name)
        for t in tris:
            for k in ("v1", "v2", "v3"):
                f.write("v %.9g %.9g %.9g\n" % tuple(float(x) for x in t[k][:3]))
            for k in ("vn1", "vn2", "vn3"):
                f.write("vn %.9g %.9g %.9g\n" % tuple(float(x) for x in t[k][:3]))
        cur = None
        for i, t in enumerate(tris):
            if t["matID"] != cur:
                cur = t["matID"]
                f.write("usemtl %s\n" % names[int(cur)])
            b = 3 * i + 1
            f.write("f %d//%d %d//%d %d//%d\n" % (b, b, b + 1, b + 1, b + 2, b + 2))


def luminance(img):
    return 0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]


def rel_rmse(a, b, mask=None):
    """Per-pixel relative RMSE with +1e-2 in the denominator (SURVEY.md 8c pin 3)."""
    a = a[..., :3].astype(np.float64); b = b[..., :3].astype(np.float64)
    if mask is None:
        mask = np.isfinite(a).all(-1) & np.isfinite(b).all(-1)
    e = (a[mask] - b[mask]) / (b[mask] + 1e-2)
    return float(np.sqrt((e ** 2).mean()))
