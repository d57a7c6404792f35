"""A small tour of every device path for compute-sanitizer (memcheck / synccheck): both device builders, device layout, host
layout, all trace-kernel variants (reference tree, own tree staged / unstaged, 4-wide, watertight), sorted queues, pipelined calls,
BDPT, tonemap.  Run under gpurun:   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yune_b200 as yb
from tests.refbind import load_golden_scene
from tests.helpers import random_soup, soup_rays, transmissive
rng = np.random.default_rng(1)
tris, mats, nodes = load_golden_scene("teapot")
m = yb.CUDAManager().setup(0)
m.setOption("max_iterations", 20000)
W = H = 48
r = yb.RendererCore(m, W, H)
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = transmissive(tris), mats, nodes
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS") and m.createPostProcProgram("tonemap.cl")
assert r.setup(sc), m.last_message
od, tm = soup_rays(rng, tris, 4000)
def tour(tag):
    st = r.enqueueKernels(2, reset=True)
    r.traceRays(od); r.traceRays(od, tm, any_hit=True); r.tracePrimary(1, 7)
    print(tag, "ok", st.iterations, flush=True)
tour("host layout + reinsertion")
for key, val in (("smem_nodes", 100), ("accel", 0), ("accel", 2), ("accel", 1), ("isect", 1), ("isect", 0), ("device_layout", 1), ("device_builder", 0), ("device_builder", 1),
                 ("sort_rays", 1), ("sort_rays", 0), ("deterministic", 0), ("deterministic", 1), ("smem_nodes", -1)):
    m.setOption(key, val)
    tour("%s=%s" % (key, val))
for builder in (1, 0):
    m.setOption("device_builder", builder)
    for leaf in (1, 2, 10):
        assert m.buildBVHOnDevice(leaf), m.last_message
        m.readBVHBuffer()
        tour("device BVH builder %d leaf_max %d" % (builder, leaf))
m.setOption("device_builder", 1)
assert r.setup(sc)
m.setOption("pipeline", 1)
for f in range(4):
    r.enqueueKernels(1, reset=(f == 0))
r.finish(); m.setOption("pipeline", 0)
print("pipelined ok", flush=True)
r.postProcess(); r.readLDR(); r.readHDR(); r.readSumFixed()
T = random_soup(rng, 3000, "mixed")
soup = yb.Scene().setGeometry(T, mats)
assert r.setup(soup); tour("soup")
assert m.buildBVHOnDevice(2); tour("soup device BVH")
assert m.createRenderProgram("bdpt.cl") and r.setup(sc)
st = r.enqueueKernels(1, reset=True); print("bdpt ok", st.iterations, flush=True)
m.close()
print("DONE")
