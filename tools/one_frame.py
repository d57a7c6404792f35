"""Two 1-spp calls of the C2 scene at 1024^2 for an ncu launch list of the second one (which kernels a reference "frame" costs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from bench import load_scene
tris, mats, nodes = load_scene()
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, 1024, 1024)
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS")
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc)
for k, v in (kv.split("=") for kv in sys.argv[1:]):
    m.setOption(k, float(v))
st = r.enqueueKernels(1)
st = r.enqueueKernels(1)
print("ms", st.render_ms, "iterations", st.iterations)
