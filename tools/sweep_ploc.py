"""PLOC search radius x leaf size on the C4 scene with the reference builder's tree uploaded and the own tree built on the device
(option device_layout = 1): device build time and render speed (measurement aid; run under gpurun)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from yune_b200.scenes import synthetic_c4
from tests.refbind import load_golden_scene
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W, H = 3840, 2160
tris, mats, _ = load_golden_scene("cornellbox")
T = synthetic_c4(tris, 9)
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, W, H)
assert m.createRenderProgram("udpt.cl")
m.setOption("pool_slots", 1 << 24); m.setOption("device_layout", 1)
assert m.setupVertexBuffer(T) and m.setupMatBuffer(mats) and m.setupImageBuffers(W, H)
m.setupCameraBuffer(yb.default_camera())
for radius, leaf in ((8, 2), (16, 2), (32, 2), (64, 2), (16, 1), (16, 3), (32, 3), (16, 4)):
    m.setOption("ploc_radius", radius)
    assert m.buildBVHOnDevice(leaf), m.last_message
    info = m.bvhInfo()
    r.enqueueKernels(1)
    m.setOption("time_stages", 4)
    st = r.enqueueKernels(spp, reset=True)
    m.setOption("time_stages", 0)
    n = max(st.timed_iterations, 1)
    print(json.dumps(dict(radius=radius, leaf_max=leaf, build_ms=round(info["device_build_ms"], 1), depth=info["depth"], n_inner=info["n_inner"],
                          msamples_s=round(st.samples / st.render_ms / 1e3, 1), trace_ms=round(st.trace_ms / n, 4))), flush=True)
