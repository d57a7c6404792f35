"""Host-side mirror of the reference's operator surface, over the C ABI.

  Scene        <- yune::Scene        (include/Scene.h:40-64)        loadModel / loadBVH / reloadMatFile, vert_data, mat_data, bvh
  CUDAManager  <- yune::CLManager    (include/CLManager.h:50-88)    setup, createRenderProgram, setup*Buffer
  RendererCore <- yune::RendererCore (include/RendererCore.h:46-90) setup, enqueueKernels (headless: blocks until done), stats

Error behaviour follows the reference: Scene methods raise (std::runtime_error there, YuneError here);
CUDAManager methods that return bool in the reference return bool here and keep the message in
`last_message` (the reference pushes it to the GUI callback, src/CLManager.cpp:261-265); setup() raises.
"""
import ctypes as C
import math
import os

import numpy as np

from . import _native

TRI_DTYPE = np.dtype([("v1", "<f4", 4), ("v2", "<f4", 4), ("v3", "<f4", 4), ("vn1", "<f4", 4), ("vn2", "<f4", 4), ("vn3", "<f4", 4),
                      ("matID", "<i4"), ("pad", "<f4", 3)])
NODE_DTYPE = np.dtype([("p_min", "<f4", 4), ("p_max", "<f4", 4), ("vert_list", "<i4", 10), ("child_idx", "<i4"), ("vert_len", "<i4")])
MAT_DTYPE = np.dtype([("ke", "<f4", 4), ("kd", "<f4", 4), ("ks", "<f4", 4), ("n", "<f4"), ("k", "<f4"), ("px", "<f4"), ("py", "<f4"),
                      ("alpha_x", "<f4"), ("alpha_y", "<f4"), ("is_specular", "<i4"), ("is_transmissive", "<i4")])
CAM_DTYPE = np.dtype([("r1", "<f4", 4), ("r2", "<f4", 4), ("r3", "<f4", 4), ("r4", "<f4", 4), ("view_plane_dist", "<f4"), ("pad", "<f4", 3)])
QUAD_DTYPE = np.dtype([("pos", "<f4", 4), ("normal", "<f4", 4), ("ke", "<f4", 4), ("kd", "<f4", 4), ("ks", "<f4", 4),
                       ("edge_l", "<f4", 4), ("edge_w", "<f4", 4), ("phong_exponent", "<f4"), ("pad", "<f4", 3)])
assert TRI_DTYPE.itemsize == 112 and NODE_DTYPE.itemsize == 80 and MAT_DTYPE.itemsize == 80 and CAM_DTYPE.itemsize == 80 and QUAD_DTYPE.itemsize == 128


class YuneError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def write_image(path, rgba, save_ext=None):
    """Encode a bottom-up RGBA float image (H, W, 4) into `path`; format = `save_ext` or else the extension of `path`
    (.hdr .pfm .png .jpg .ppm): the product's own encoders behind yune_write_image (include/yune_host.h).  True on success."""
    a = np.ascontiguousarray(rgba, np.float32)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("write_image expects an (H, W, 4) float image")
    return _native.load().yune_write_image(path.encode(), save_ext.encode() if save_ext else None, _ptr(a), int(a.shape[1]), int(a.shape[0])) == 0


def default_camera(y_fov=60.0):
    """The reference's default camera (src/Camera.cpp:93-103) as the 80-byte Cam record."""
    cam = np.zeros(1, CAM_DTYPE)
    _native.load().yune_camera_default(C.c_float(y_fov), _ptr(cam))
    return cam


class Camera:
    """yune::Camera (src/Camera.cpp): pose state + setOrientation + setBuffer, computed by the native host library."""

    def __init__(self, y_fov=60.0):
        self._lib = _native.load()
        self._h = self._lib.yune_camera_create(C.c_float(y_fov))
        if not self._h:
            raise MemoryError("yune_camera_create")

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.yune_camera_destroy(self._h); self._h = None

    def setOrientation(self, direction=(0, 0, 0, 0), pitch=0.0, yaw=0.0):
        d = np.asarray(direction, np.float32)
        self._lib.yune_camera_set_orientation(self._h, _ptr(d), C.c_float(pitch), C.c_float(yaw))
        return self

    def resetCamera(self):
        self._lib.yune_camera_reset(self._h); return self

    @property
    def is_changed(self):
        return bool(self._lib.yune_camera_is_changed(self._h))

    def setBuffer(self):
        cam = np.zeros(1, CAM_DTYPE)
        self._lib.yune_camera_set_buffer(self._h, _ptr(cam))
        return cam


def quad_light(pos, normal, ke, edge_l, edge_w):
    q = np.zeros(1, QUAD_DTYPE)
    q["pos"][0] = list(pos) + [1.0]
    q["normal"][0] = list(normal) + [0.0]
    q["ke"][0] = list(ke) + [0.0]
    q["edge_l"][0] = list(edge_l) + [0.0]
    q["edge_w"][0] = list(edge_w) + [0.0]
    return q


LIGHT_UDPT = quad_light((-0.1979, 0.92, -3.1972), (0, -1, 0), (16, 16, 16), (0.4, 0, 0), (0, 0, 0.4))        # udpt.cl:97-106
LIGHT_BDPT = quad_light((-0.1979, 0.703, -3.1972), (0, 1, 0), (18.3, 16.2, 14.5), (0.4, 0, 0), (0, 0, 0.4))  # bdpt.cl:106-115


class Scene:
    """OBJ/MTL -> device-layout arrays + SAH BVH, computed by the native host library (csrc/host)."""

    def __init__(self):
        self._lib = _native.load()
        self._h = C.c_void_p(self._lib.yune_scene_create())
        if not self._h:
            raise YuneError(self._lib.yune_scene_last_error(None).decode())
        self.vert_data = np.zeros(0, TRI_DTYPE)
        self.mat_data = np.zeros(0, MAT_DTYPE)
        self.bvh = np.zeros(0, NODE_DTYPE)
        self.main_camera = default_camera()
        self.root = np.zeros(8, np.float32)
        self.num_triangles = 0

    def __del__(self):
        try:
            if self._h:
                self._lib.yune_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def _refresh(self):
        L = self._lib
        nt, nm, nn = L.yune_scene_num_triangles(self._h), L.yune_scene_num_materials(self._h), L.yune_scene_num_bvh_nodes(self._h)

        def view(p, n, dt):
            if n == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * (n * dt.itemsize)).from_address(p)
            return np.frombuffer(buf, dtype=dt, count=n).copy()
        self.vert_data = view(L.yune_scene_vert_data(self._h), nt, TRI_DTYPE)
        self.mat_data = view(L.yune_scene_mat_data(self._h), nm, MAT_DTYPE)
        self.bvh = view(L.yune_scene_bvh_data(self._h), nn, NODE_DTYPE)
        L.yune_scene_root_aabb(self._h, _ptr(self.root))
        self.num_triangles = nt

    def loadModel(self, filepath, filename=None, bvh_bins=20):
        """Scene::loadModel (src/Scene.cpp:133-383). `filename` is accepted for signature parity and ignored."""
        if self._lib.yune_scene_load_model(self._h, os.fspath(filepath).encode(), int(bvh_bins)) != 0:
            raise YuneError(self._lib.yune_scene_last_error(self._h).decode())
        self._refresh()
        return self

    def setGeometry(self, vert_data, mat_data, bvh_bins=20):
        """loadModel for geometry already in memory (synthetic scenes): same centroid/AABB/BVH code path."""
        t = np.ascontiguousarray(vert_data, TRI_DTYPE); m = np.ascontiguousarray(mat_data, MAT_DTYPE)
        if self._lib.yune_scene_set_geometry(self._h, _ptr(t), int(t.size), _ptr(m), int(m.size), int(bvh_bins)) != 0:
            raise YuneError(self._lib.yune_scene_last_error(self._h).decode())
        self._refresh()
        return self

    def loadBVH(self, bvh_bins):
        if self._lib.yune_scene_load_bvh(self._h, int(bvh_bins)) != 0:
            raise YuneError(self._lib.yune_scene_last_error(self._h).decode())
        self._refresh()

    def reloadMatFile(self):
        if self._lib.yune_scene_reload_mat_file(self._h) != 0:
            raise YuneError(self._lib.yune_scene_last_error(self._h).decode())
        self._refresh()


class CUDAManager:
    """Replacement of yune::CLManager: owns the device context and every device buffer."""

    def __init__(self):
        self._lib = _native.load()
        self._ctx = C.c_void_p()
        self.last_message = ""
        self.rk_file = ""
        self.rk_compiler_opts = ""
        self.ppk_file = ""

    # -- CLManager::setup(): raises like the reference does (exceptions propagate to main, main.cpp:15-21)
    def setup(self, device=0):
        rc = self._lib.yune_setup(int(device), C.byref(self._ctx))
        if rc != 0:
            raise YuneError(self._lib.yune_last_error(None).decode())
        return self

    def close(self):
        if self._ctx:
            self._lib.yune_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ok(self, rc):
        if rc != 0:
            self.last_message = self._lib.yune_last_error(self._ctx).decode()
            return False
        return True

    def check(self, rc):
        if not self._ok(rc):
            raise YuneError(self.last_message)

    def createRenderProgram(self, fn, path="", reload=False, compiler_opts=None):
        """`fn` is the reference kernel file name ("udpt.cl", "bdpt.cl").  If `path` names a real file, its
        '#yune-preproc compiler-opts' line is honoured like src/CLManager.cpp:182-204 does."""
        opts = compiler_opts
        if opts is None and path and os.path.isfile(path):
            with open(path, "r", errors="replace") as f:
                for line in f:
                    t = line.split()
                    if len(t) >= 3 and t[0] == "#yune-preproc" and t[1] == "compiler-opts":
                        opts = t[2]
        ok = self._ok(self._lib.yune_create_render_program(self._ctx, fn.encode(), (opts or "").encode()))
        if ok:
            self.rk_file, self.rk_compiler_opts = fn, opts or ""
        return ok

    def createPostProcProgram(self, fn, path="", reload=False):
        ok = self._ok(self._lib.yune_create_postproc_program(self._ctx, fn.encode(), b""))
        if ok:
            self.ppk_file = fn
        return ok

    def setupCameraBuffer(self, cam):
        cam = np.ascontiguousarray(cam)
        self.check(self._lib.yune_setup_camera_buffer(self._ctx, _ptr(cam)))

    def setupImageBuffers(self, width, height):
        return self._ok(self._lib.yune_setup_image_buffers(self._ctx, int(width), int(height)))

    def setupBVHBuffer(self, bvh_data, bvh_size=0.0, scene_size=0.0):
        a = np.ascontiguousarray(bvh_data)
        return self._ok(self._lib.yune_setup_bvh_buffer(self._ctx, _ptr(a), int(a.size)))

    def setupVertexBuffer(self, vert_data, scene_size=0.0):
        a = np.ascontiguousarray(vert_data)
        return self._ok(self._lib.yune_setup_vertex_buffer(self._ctx, _ptr(a), int(a.size)))

    def buildBVHOnDevice(self, leaf_max=2):
        """Build the BVH on the GPU from the uploaded vertex buffer (yune_build_bvh_on_device) instead of uploading one."""
        return self._ok(self._lib.yune_build_bvh_on_device(self._ctx, int(leaf_max)))

    def bvhInfo(self):
        n, ni, d, ms = C.c_int(), C.c_int(), C.c_int(), C.c_float()
        self.check(self._lib.yune_bvh_info(self._ctx, C.byref(n), C.byref(ni), C.byref(d), C.byref(ms)))
        return dict(n_nodes=n.value, n_inner=ni.value, depth=d.value, device_build_ms=ms.value)

    def readBVHBuffer(self):
        """The BVHNodeGPU array in use (uploaded or device-built), reference record layout."""
        a = np.zeros(self.bvhInfo()["n_nodes"], NODE_DTYPE)
        self.check(self._lib.yune_read_bvh_buffer(self._ctx, _ptr(a), int(a.size)))
        return a

    def setupMatBuffer(self, mat_data):
        a = np.ascontiguousarray(mat_data)
        return self._ok(self._lib.yune_setup_mat_buffer(self._ctx, _ptr(a), int(a.size)))

    def setLightSources(self, quads):
        if quads is None:
            return self._ok(self._lib.yune_set_light_sources(self._ctx, None, 0))
        a = np.ascontiguousarray(quads)
        return self._ok(self._lib.yune_set_light_sources(self._ctx, _ptr(a), int(a.size)))

    def setOption(self, key, value):
        self.check(self._lib.yune_set_option(self._ctx, key.encode(), float(value)))

    def getOption(self, key):
        v = C.c_double()
        self.check(self._lib.yune_get_option(self._ctx, key.encode(), C.byref(v)))
        return v.value


class CUDAGroup:
    """Several devices of one node behind one object (include/yune_cuda.h: yune_group_*): the scene is replicated, a render is
    sharded by SAMPLE INDEX, one NCCL sum-reduce merges the accumulation buffers on the root."""

    def __init__(self, n_devices=0, devices=None):
        self._lib = _native.load()
        self._g = C.c_void_p()
        arr = None
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices); n_devices = len(devices)
        rc = self._lib.yune_group_create(int(n_devices), arr, C.byref(self._g))
        if rc != 0:
            raise YuneError(self._lib.yune_group_last_error(None).decode())
        self.size = self._lib.yune_group_size(self._g)

    def close(self):
        if self._g:
            self._lib.yune_group_destroy(self._g); self._g = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise YuneError(self._lib.yune_group_last_error(self._g).decode())

    def manager(self, rank):
        """A CUDAManager view of one rank's context (owned by the group: do not close it)."""
        m = CUDAManager.__new__(CUDAManager)
        m._lib = self._lib; m._ctx = C.c_void_p(self._lib.yune_group_ctx(self._g, int(rank)))
        m.last_message = ""; m.rk_file = m.rk_compiler_opts = m.ppk_file = ""
        m.close = lambda: None
        return m

    def setup(self, scene, width, height, kernel="udpt.cl", compiler_opts="", cam=None):
        L, g = self._lib, self._g
        self.width, self.height = int(width), int(height)
        v, mt, b = np.ascontiguousarray(scene.vert_data), np.ascontiguousarray(scene.mat_data), np.ascontiguousarray(scene.bvh)
        cam = np.ascontiguousarray(cam if cam is not None else scene.main_camera)
        self.check(L.yune_group_create_render_program(g, kernel.encode(), compiler_opts.encode()))
        self.check(L.yune_group_create_postproc_program(g, b"tonemap.cl", b""))
        self.check(L.yune_group_setup_vertex_buffer(g, _ptr(v), int(v.size)))
        self.check(L.yune_group_setup_mat_buffer(g, _ptr(mt), int(mt.size)))
        self.check(L.yune_group_setup_bvh_buffer(g, _ptr(b), int(b.size)))
        self.check(L.yune_group_setup_image_buffers(g, self.width, self.height))
        self.check(L.yune_group_setup_camera_buffer(g, _ptr(cam)))
        return self

    def setOption(self, key, value):
        self.check(self._lib.yune_group_set_option(self._g, key.encode(), float(value)))

    def render(self, spp_begin, spp_count, seed=12345, gi_check=True, reset=True):
        self.check(self._lib.yune_group_render(self._g, int(spp_begin), int(spp_count), int(bool(gi_check)), int(seed) & 0xffffffff, int(bool(reset))))
        return self.stats()

    def reduce(self, root=0):
        self.check(self._lib.yune_group_reduce(self._g, int(root)))
        return self.stats()

    def stats(self):
        st = _native.GroupStats()
        self._lib.yune_group_get_stats(self._g, C.byref(st))
        return st

    def readSum(self, rank=0):
        img = np.zeros((self.height, self.width, 4), np.float32)
        m = self.manager(rank)
        m.check(self._lib.yune_read_sum(m._ctx, _ptr(img)))
        return img


def _group_read_sum_fixed(self, rank=0):
    img = np.zeros((self.height, self.width, 4), np.int64)
    m = self.manager(rank)
    m.check(self._lib.yune_read_sum_fixed(m._ctx, _ptr(img)))
    return img


CUDAGroup.readSumFixed = _group_read_sum_fixed


def shard_samples(spp_begin, spp_count, rank, n_ranks):
    """yune_shard_samples: (begin, count) of `rank`'s contiguous, balanced slice."""
    b, c = C.c_int(), C.c_int()
    _native.load().yune_shard_samples(int(spp_begin), int(spp_count), int(rank), int(n_ranks), C.byref(b), C.byref(c))
    return b.value, c.value


class RendererCore:
    """Headless counterpart of yune::RendererCore: uploads the scene (setup) and renders frames (enqueueKernels)."""

    def __init__(self, cuda_manager, width, height):
        self.cl_manager = cuda_manager
        self._lib = cuda_manager._lib
        self.width, self.height = int(width), int(height)
        self.render_scene = None
        self.samples_taken = 0
        self.seed = 12345
        self.gi_check = True
        self.stats = _native.Stats()

    @property
    def _ctx(self):
        return self.cl_manager._ctx

    def loadScene(self, path, fn=None, bvh_bins=20):
        """RendererCore::loadScene (src/RendererCore.cpp:124-137): returns False and keeps the message on failure."""
        try:
            self.render_scene = Scene().loadModel(path, fn, bvh_bins)
            return True
        except YuneError as e:
            self.cl_manager.last_message = str(e)
            return False

    def setup(self, scene=None, gi_check=True):
        """RendererCore::setup (src/RendererCore.cpp:155-246): upload vertex/material/BVH buffers, image buffers, camera."""
        if scene is not None:
            self.render_scene = scene
        s = self.render_scene
        m = self.cl_manager
        ok = m.setupVertexBuffer(s.vert_data) and m.setupMatBuffer(s.mat_data) and m.setupBVHBuffer(s.bvh) \
            and m.setupImageBuffers(self.width, self.height)
        if not ok:
            return False
        m.setupCameraBuffer(s.main_camera)
        self.gi_check = bool(gi_check)
        self.samples_taken = 0
        return True

    def enqueueKernels(self, frames=1, gi_check=None, reset=None):
        """Render `frames` more samples per pixel (one reference frame = 1 spp, src/RendererCore.cpp:483-486)."""
        if gi_check is not None and bool(gi_check) != self.gi_check:
            self.gi_check = bool(gi_check)
            reset = True                                   # GI toggle forces a reset (src/RendererCore.cpp:556-565)
        if reset is None:
            reset = self.samples_taken == 0
        if reset:
            self.samples_taken = 0
        self.cl_manager.check(self._lib.yune_render(self._ctx, self.samples_taken, int(frames), int(self.gi_check), self.seed, int(bool(reset))))
        self.samples_taken += int(frames)
        self._lib.yune_get_stats(self._ctx, C.byref(self.stats))
        return self.stats

    def finish(self):
        """Option "pipeline": complete the paths the last enqueueKernels left in flight (yune_finish)."""
        self.cl_manager.check(self._lib.yune_finish(self._ctx))
        self._lib.yune_get_stats(self._ctx, C.byref(self.stats))
        return self.stats

    def postProcess(self):
        self.cl_manager.check(self._lib.yune_tonemap(self._ctx))
        self._lib.yune_get_stats(self._ctx, C.byref(self.stats))

    def _read(self, fn):
        img = np.zeros((self.height, self.width, 4), np.float32)
        self.cl_manager.check(fn(self._ctx, _ptr(img)))
        return img

    def readHDR(self):
        return self._read(self._lib.yune_read_hdr)

    def readSum(self):
        return self._read(self._lib.yune_read_sum)

    def readSumFixed(self):
        """Option "deterministic": the fixed-point accumulation buffer (H, W, 4) int64, unit 2^-24 (r, g, b) and sample count."""
        img = np.zeros((self.height, self.width, 4), np.int64)
        self.cl_manager.check(self._lib.yune_read_sum_fixed(self._ctx, _ptr(img)))
        return img

    def readLDR(self):
        return self._read(self._lib.yune_read_ldr)

    def saveImage(self, save_fn, save_ext=None):
        """RendererCore::saveImage (src/RendererCore.cpp:608-646): ".hdr" (and ".pfm") from the float image, ".png" / ".jpg"
        (and ".ppm") from the 8-bit view of the tonemapped one.  Returns False with cl_manager.last_message set on failure."""
        ext = (save_ext or os.path.splitext(save_fn)[1]).lower()
        if ext in (".png", ".jpg", ".jpeg", ".ppm"):
            self.postProcess()
            img = self.readLDR()
        elif ext in (".hdr", ".pfm"):
            img = self.readHDR()
        else:
            self.cl_manager.last_message = "unsupported image extension (use .hdr, .png, .jpg, .pfm or .ppm)"
            return False
        return write_image(save_fn, img, ext)          # the file is save_fn as given, like the reference

    def writeSum(self, img):
        a = np.ascontiguousarray(img, np.float32)
        self.cl_manager.check(self._lib.yune_write_sum(self._ctx, _ptr(a)))

    def captureRays(self, iteration, max_rays):
        self.cl_manager.check(self._lib.yune_debug_capture_rays(self._ctx, int(iteration), int(max_rays)))

    def readCaptured(self, which):
        n, nq = C.c_int(), C.c_int()
        self.cl_manager.check(self._lib.yune_debug_read_captured(self._ctx, int(which), None, None, C.byref(n), C.byref(nq)))
        od = np.zeros((n.value, 6), np.float32); tm = np.zeros(n.value, np.float32)
        if n.value:
            self.cl_manager.check(self._lib.yune_debug_read_captured(self._ctx, int(which), _ptr(od), _ptr(tm), C.byref(n), C.byref(nq)))
        return od, tm, nq.value

    def tracePrimary(self, jitter_mode=0, rand=0):
        n = self.width * self.height
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        self.cl_manager.check(self._lib.yune_trace_primary(self._ctx, int(jitter_mode), int(rand) & 0xffffffff, _ptr(tri), _ptr(light), _ptr(t)))
        return tri, light, t

    def traceRays(self, od6, tmax=None, any_hit=False):
        od6 = np.ascontiguousarray(od6, np.float32)
        n = od6.shape[0]
        tm = None if tmax is None else np.ascontiguousarray(tmax, np.float32)
        tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        self.cl_manager.check(self._lib.yune_trace_rays(self._ctx, n, _ptr(od6), _ptr(tm), int(bool(any_hit)), _ptr(tri), _ptr(light), _ptr(t)))
        return tri, light, t
