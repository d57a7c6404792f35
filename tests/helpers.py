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


def random_soup(rng, n, kind):
    """Random triangle soups that stress different branches of BVH::createBVH: uniform (SAH wins), clustered (SAH loses, spatial
    median with empty children), axis-aligned flats (the +0.2 padding of src/TriangleCPU.cpp:63-67), slivers of mixed size."""
    T = np.zeros(n, dtype=TRI_DTYPE)
    if kind == "uniform":
        c = rng.uniform(-1, 1, (n, 1, 3)); p = c + rng.uniform(-0.05, 0.05, (n, 3, 3))
    elif kind == "clustered":
        centres = rng.uniform(-4, 4, (6, 3))
        c = centres[rng.integers(0, 6, n)][:, None, :] + rng.normal(0, 0.02, (n, 1, 3)); p = c + rng.normal(0, 0.01, (n, 3, 3))
    elif kind == "flats":
        c = rng.uniform(-1, 1, (n, 1, 3)); p = c + rng.uniform(-0.1, 0.1, (n, 3, 3))
        ax = rng.integers(0, 3, n)
        p[np.arange(n), :, ax] = c[np.arange(n), 0, ax][:, None]            # zero extent on one axis
    else:  # mixed sizes over 4 decades, long slivers
        c = rng.uniform(-10, 10, (n, 1, 3)); s = 10.0 ** rng.uniform(-3, 1, (n, 1, 1))
        p = c + s * rng.uniform(-1, 1, (n, 3, 3)) * np.array([1.0, 0.02, 1.0])
    p = p.astype(np.float32)
    for k, name in enumerate(("v1", "v2", "v3")):
        T[name][:, :3] = p[:, k]; T[name][:, 3] = 1.0
    nrm = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0]).astype(np.float32)
    for name in ("vn1", "vn2", "vn3"):
        T[name][:, :3] = nrm
    T["matID"] = rng.integers(0, 7, n)
    return T


def soup_rays(rng, T, m):
    """m rays for a random_soup scene: origins in (and a little around) the scene box, half aimed at points inside triangles,
    half near them; some axis-degenerate directions (NaN-guarded slabs).  Returns (origin+direction [m, 6], segment lengths)."""
    p = np.stack([T["v1"], T["v2"], T["v3"]], 1)[..., :3].astype(np.float64)
    lo, hi = p.min((0, 1)), p.max((0, 1))
    o = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), (m, 3))
    w = rng.dirichlet([1, 1, 1], m)
    target = (p[rng.integers(0, T.size, m)] * w[:, :, None]).sum(1)
    target[m // 2:] += rng.normal(0, 0.02, (m - m // 2, 3)) * (hi - lo)
    d = target - o
    k = max(m // 100, 1)
    d[:k, 0] = 0; d[k:2 * k, 1] = 0; d[2 * k:3 * k, 2] = 0; d[3 * k:3 * k + k // 3, :2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = (np.linalg.norm(target - o, axis=1) * rng.uniform(0.2, 1.5, m)).astype(np.float32)
    return od, tm


def luminance(img):
    return 0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]


def rel_rmse(a, b, mask=None):
    """Per-pixel relative RMSE with +1e-2 in the denominator (SURVEY.md 8c pin 3)."""
    a = a[..., :3].astype(np.float64); b = b[..., :3].astype(np.float64)
    if mask is None:
        mask = np.isfinite(a).all(-1) & np.isfinite(b).all(-1)
    e = (a[mask] - b[mask]) / (b[mask] + 1e-2)
    return float(np.sqrt((e ** 2).mean()))
