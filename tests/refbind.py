"""tests/refbind.py -- ctypes bindings of the CHECKER libraries (test infrastructure only):

  oracle/_ref/libyune_ref_host.so     the reference's own Scene/BVH sources (oracle/ref_host_api.cpp)
  oracle/_ref/libyune_ref_kernels.so  the reference's own OpenCL kernels compiled as C++ (oracle/gen_ref_kernels.py)
  oracle/libyune_oracle.so            the hand restatement (oracle/yune_oracle.cpp)

Nothing under yune_b200/ imports this module.
"""
import ctypes as C
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_HOST = os.path.join(ROOT, "oracle", "_ref", "libyune_ref_host.so")
REF_KERNELS = os.path.join(ROOT, "oracle", "_ref", "libyune_ref_kernels.so")
ORACLE = os.path.join(ROOT, "oracle", "libyune_oracle.so")

TRI_DTYPE = np.dtype([("v1", "<f4", 4), ("v2", "<f4", 4), ("v3", "<f4", 4), ("vn1", "<f4", 4), ("vn2", "<f4", 4), ("vn3", "<f4", 4),
                      ("matID", "<i4"), ("pad", "<f4", 3)])
NODE_DTYPE = np.dtype([("p_min", "<f4", 4), ("p_max", "<f4", 4), ("vert_list", "<i4", 10), ("child_idx", "<i4"), ("vert_len", "<i4")])
MAT_DTYPE = np.dtype([("ke", "<f4", 4), ("kd", "<f4", 4), ("ks", "<f4", 4), ("n", "<f4"), ("k", "<f4"), ("px", "<f4"), ("py", "<f4"),
                      ("alpha_x", "<f4"), ("alpha_y", "<f4"), ("is_specular", "<i4"), ("is_transmissive", "<i4")])
P = C.c_void_p


def ptr(a):
    return a.ctypes.data_as(P) if a is not None else None


def ncores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def default_cam_array(y_fov=60.0):
    """Default Cam record restated from src/Camera.cpp:36-103 (identity rows, vpd = 1/tan(fov*3.14/360) in double)."""
    cam = np.zeros(20, np.float32)
    cam[0] = cam[5] = cam[10] = cam[15] = 1.0
    cam[16] = np.float32(1 / math.tan(np.float32(y_fov) * 3.14 / 360))
    return cam


def frame_rands(seed, n):
    """Per-frame `rand` list: mt19937(seed) raw outputs, the stand-in for RendererCore's clock-seeded engine
    (src/RendererCore.cpp:61-63, 268).  std::uniform_int_distribution over the full 32-bit range returns the raw word."""
    rs = np.random.RandomState(seed)           # MT19937
    return [int(x) for x in rs.randint(0, 2 ** 32, size=n, dtype=np.uint64)]


def have_ref():
    return os.path.exists(REF_HOST) and os.path.exists(REF_KERNELS)


class RefHost:
    def __init__(self):
        self.lib = C.CDLL(REF_HOST)
        self.lib.yref_scene_load.restype = C.c_void_p
        self.lib.yref_scene_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        self.lib.yref_last_error.restype = C.c_char_p

    def load(self, path, bins=-1):
        h = self.lib.yref_scene_load(path.encode(), os.path.basename(path).encode(), bins)
        if not h:
            raise RuntimeError(self.lib.yref_last_error().decode())
        h = C.c_void_p(h)
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.yref_scene_counts(h, C.byref(a), C.byref(b), C.byref(c))
        tris = np.zeros(a.value, TRI_DTYPE); mats = np.zeros(b.value, MAT_DTYPE); nodes = np.zeros(c.value, NODE_DTYPE)
        root = np.zeros(8, np.float32)
        self.lib.yref_scene_copy(h, ptr(tris), ptr(mats), ptr(nodes), ptr(root))
        self.lib.yref_scene_free(h)
        return tris, mats, nodes, root


class RefKernels:
    """The reference kernels on host threads. variant in {"udpt", "udpt_mis", "bdpt"}."""

    def __init__(self):
        self.lib = C.CDLL(REF_KERNELS)

    def frame(self, variant, out, inp, cam, tris, mats, nodes, gi, reset, rand, W, H, threads=None, blocks=(2, 2)):
        f = getattr(self.lib, "yref_%s_frame" % variant)
        f(ptr(out), ptr(inp), ptr(cam), int(tris.size), ptr(tris), ptr(mats), int(nodes.size), ptr(nodes), int(gi), int(reset),
          C.c_uint(rand & 0xffffffff), W, H, blocks[0], blocks[1], threads or ncores())

    def render(self, variant, cam, tris, mats, nodes, W, H, rands, gi=1, threads=None):
        """Progressive render exactly like RendererCore drives the kernel: ping-pong images, reset on frame 0."""
        a = np.zeros((H, W, 4), np.float32); b = np.zeros_like(a)
        for i, r in enumerate(rands):
            self.frame(variant, b, a, cam, tris, mats, nodes, gi, 1 if i == 0 else 0, r, W, H, threads)
            a, b = b, a
        return a

    def primary(self, variant, cam, tris, nodes, rand, jitter_mode, W, H, threads=None):
        n = W * H
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32); od = np.zeros((n, 6), np.float32)
        f = getattr(self.lib, "yref_%s_primary" % variant)
        f(ptr(tri), ptr(light), ptr(t), ptr(od), ptr(cam), int(tris.size), ptr(tris), int(nodes.size), ptr(nodes),
          C.c_uint(rand & 0xffffffff), int(jitter_mode), W, H, threads or ncores())
        return tri, light, t, od

    def trace(self, variant, od6, tmax, shadow, tris, nodes, threads=None):
        od6 = np.ascontiguousarray(od6, np.float32); n = od6.shape[0]
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        f = getattr(self.lib, "yref_%s_trace" % variant)
        f(n, ptr(od6), ptr(tmax), int(shadow), ptr(tri), ptr(light), ptr(t), int(tris.size), ptr(tris), int(nodes.size), ptr(nodes), threads or ncores())
        return tri, light, t

    def tonemap(self, img):
        img = np.ascontiguousarray(img, np.float32); out = np.zeros_like(img)
        self.lib.yref_tonemap_frame(ptr(img), ptr(img), ptr(out), 0, img.shape[1], img.shape[0])
        return out


def load_golden_scene(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", "scene_%s.npz" % name))
    return z["vert_data"].view(TRI_DTYPE).reshape(-1), z["mat_data"].view(MAT_DTYPE).reshape(-1), z["bvh"].view(NODE_DTYPE).reshape(-1)
