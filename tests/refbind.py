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
    cam[16] = np.float32(1 / math.tan(float(np.float32(y_fov)) * 3.14 / 360))   # double arithmetic, like the C++
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
        self._padded = None

    def _guarded(self, tris):
        """bdpt.cl reads scene_data[-1] when the camera ray hits the light (createEyePath continues with triangle_ID = -1,
        bdpt.cl:518-528 -> sampleGlossyPdf :1055).  On a GPU that is a silent out-of-bounds read whose value is discarded
        (the pixel returns the constant light colour); on the host it can fault.  The harness therefore hands the reference
        a buffer with one readable zero record in front of element 0.  Reference code is unchanged."""
        buf = np.zeros(tris.size + 1, TRI_DTYPE)
        buf[1:] = tris
        self._padded = buf                      # keep alive for the duration of the call
        return C.c_void_p(buf.ctypes.data + TRI_DTYPE.itemsize)

    def frame(self, variant, out, inp, cam, tris, mats, nodes, gi, reset, rand, W, H, threads=None, blocks=(2, 2)):
        f = getattr(self.lib, "yref_%s_frame" % variant)
        f(ptr(out), ptr(inp), ptr(cam), int(tris.size), self._guarded(tris), ptr(mats), int(nodes.size), ptr(nodes), int(gi), int(reset),
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


class YorConfig(C.Structure):
    _fields_ = [("integrator", C.c_int), ("mis", C.c_int), ("rng_mode", C.c_int), ("n_lights", C.c_int), ("rr_threshold", C.c_int),
                ("oren_nayar", C.c_int), ("heap_size", C.c_int), ("bdpt_bounces", C.c_int), ("seed", C.c_uint32), ("threads", C.c_int)]


class Oracle:
    """oracle/libyune_oracle.so -- the hand restatement.  variant in {"udpt", "udpt_mis", "bdpt"}."""

    def __init__(self):
        if not os.path.exists(ORACLE):
            raise RuntimeError("oracle/libyune_oracle.so missing: run `python -m yune_b200.build --oracle`")
        self.lib = C.CDLL(ORACLE)

    @staticmethod
    def config(variant="udpt", rng_mode=0, seed=0, lights=None, oren_nayar=0, rr_threshold=-1, heap_size=-1, threads=None, bdpt_bounces=0):
        c = YorConfig()
        c.integrator = 1 if variant == "bdpt" else 0
        c.mis = 1 if variant == "udpt_mis" else 0
        c.rng_mode = rng_mode
        c.n_lights = 0 if lights is None else int(lights.size)
        c.rr_threshold = rr_threshold; c.oren_nayar = oren_nayar; c.heap_size = heap_size; c.bdpt_bounces = bdpt_bounces
        c.seed = seed & 0xffffffff
        c.threads = threads or ncores()
        return c

    def frame(self, cfg, out, inp, cam, tris, mats, nodes, gi, reset, frame_arg, W, H, lights=None):
        self.lib.yor_render_frame(C.byref(cfg), ptr(lights), ptr(out), ptr(inp), ptr(cam), ptr(tris), int(tris.size), ptr(mats),
                                  ptr(nodes), int(nodes.size), int(gi), int(reset), C.c_uint32(frame_arg & 0xffffffff), W, H)

    def render(self, cfg, cam, tris, mats, nodes, W, H, frame_args, gi=1, lights=None):
        a = np.zeros((H, W, 4), np.float32); b = np.zeros_like(a)
        for i, r in enumerate(frame_args):
            self.frame(cfg, b, a, cam, tris, mats, nodes, gi, 1 if i == 0 else 0, r, W, H, lights)
            a, b = b, a
        return a

    def samples(self, cfg, cam, tris, mats, nodes, W, H, frame_arg, gi=1, lights=None):
        out = np.zeros((H, W, 4), np.float32)
        self.frame(cfg, out, out, cam, tris, mats, nodes, gi, 1, frame_arg, W, H, lights)
        return out

    def primary(self, cfg, cam, tris, nodes, rand, jitter_mode, W, H, lights=None):
        n = W * H
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32); od = np.zeros((n, 6), np.float32)
        work = np.zeros(4, np.uint64)
        self.lib.yor_primary(C.byref(cfg), ptr(lights), ptr(cam), ptr(tris), int(tris.size), ptr(nodes), int(nodes.size),
                             C.c_uint32(rand & 0xffffffff), int(jitter_mode), W, H, ptr(tri), ptr(light), ptr(t), ptr(od), ptr(work))
        return tri, light, t, od, work

    def trace(self, cfg, od6, tmax, shadow, tris, nodes, lights=None):
        od6 = np.ascontiguousarray(od6, np.float32); n = od6.shape[0]
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        self.lib.yor_trace(C.byref(cfg), ptr(lights), n, ptr(od6), ptr(tmax), int(shadow), ptr(tris), int(tris.size), ptr(nodes), int(nodes.size),
                           ptr(tri), ptr(light), ptr(t))
        return tri, light, t

    def count_work(self, od6, tmax, shadow, tris, nodes):
        od6 = np.ascontiguousarray(od6, np.float32); work = np.zeros(2, np.uint64)
        self.lib.yor_count_work(int(od6.shape[0]), ptr(od6), ptr(tmax), int(shadow), ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(work))
        return int(work[0]), int(work[1])

    def tonemap(self, img):
        img = np.ascontiguousarray(img, np.float32); out = np.zeros_like(img)
        self.lib.yor_tonemap(ptr(img), ptr(out), int(img.size // 4))
        return out
