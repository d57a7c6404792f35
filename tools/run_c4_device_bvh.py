"""configs[3] ("C4", ~10.5 M triangles) with the BVH built ON THE DEVICE (yune_build_bvh_on_device) against the host path:
scene-to-first-sample time and render speed of both trees.  Measurement aid (run under gpurun).
    python tools/run_c4_device_bvh.py [subdiv 9] [spp 16] [leaf_max 2] [host: 0|1]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from yune_b200.scenes import synthetic_c4
from tests.refbind import load_golden_scene
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 9
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
leaf_max = int(sys.argv[3]) if len(sys.argv) > 3 else 2
with_host = int(sys.argv[4]) if len(sys.argv) > 4 else 1
W, H = 3840, 2160
tris, mats, _ = load_golden_scene("cornellbox")
T = synthetic_c4(tris, subdiv)
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, W, H)
assert m.createRenderProgram("udpt.cl")
m.setOption("pool_slots", 1 << 24)
res = dict(triangles=int(T.size), leaf_max=leaf_max, spp=spp)

def render(tag):
    st = r.enqueueKernels(1)
    m.setOption("time_stages", 4)
    st = r.enqueueKernels(spp, reset=True)
    m.setOption("time_stages", 0)
    n = max(st.timed_iterations, 1)
    res[tag] = dict(msamples_s=st.samples / st.render_ms / 1e3, avg_trace_ms=st.trace_ms / n, avg_shade_ms=st.shade_ms / n, iterations=int(st.iterations))
    img = r.readHDR()
    res[tag]["mean_luminance"] = float((0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]).mean())

# device path: vertex + material upload, build on the GPU, first sample (builder 1 = PLOC, 0 = linear BVH)
for builder, tag in ((1, "device_ploc"), (0, "device_lbvh")):
    m.setOption("device_builder", builder)
    t0 = time.time()
    assert m.setupVertexBuffer(T) and m.setupMatBuffer(mats) and m.setupImageBuffers(W, H)
    m.setupCameraBuffer(yb.default_camera())
    assert m.buildBVHOnDevice(leaf_max), m.last_message
    t1 = time.time()
    m.check(r._lib.yune_render(r._ctx, 0, 1, 1, 1, 1))
    t2 = time.time()
    render(tag)
    res[tag].update(m.bvhInfo()); res[tag]["build_call_s"] = t1 - t0; res[tag]["scene_to_first_sample_s"] = t2 - t0
    print(json.dumps(res), flush=True)
m.setOption("device_builder", 1)
if with_host:
    t0 = time.time()
    sc = yb.Scene().setGeometry(T, mats)
    t1 = time.time()
    for tag, dl in (("host_tree_device_layout", 1), ("host", 0)):      # the reference builder's tree uploaded; own tree built on the device / on the host
        m.setOption("device_layout", dl)
        t1b = time.time()
        assert r.setup(sc), m.last_message
        m.check(r._lib.yune_render(r._ctx, 0, 1, 1, 1, 1))
        t2 = time.time()
        render(tag)
        res[tag].update(m.bvhInfo()); res[tag]["bvh_build_s"] = t1 - t0; res[tag]["upload_to_first_sample_s"] = t2 - t1b; res[tag]["scene_to_first_sample_s"] = (t1 - t0) + (t2 - t1b)
        print(json.dumps(res), flush=True)
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_c4_device_bvh.json"), "w"), indent=1)
