"""Short C2 render for ncu (a number printed under a profiler is never a bench value)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from bench import load_scene
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tris, mats, nodes = load_scene()
m = yb.CUDAManager().setup(0)
for kv in sys.argv[3:]:
    k, v = kv.split("="); m.setOption(k, float(v))
r = yb.RendererCore(m, size, size)
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS")
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc)
st = r.enqueueKernels(spp)
n=max(st.timed_iterations,1)
print("timed iters", st.timed_iterations, "avg shade ms", st.shade_ms/n, "avg trace ms", st.trace_ms/n)
print("spp", spp, "ms", st.render_ms, "Msamples/s", st.samples / st.render_ms / 1e3, "iters", st.iterations, "ext", st.extend_rays, "shadow", st.shadow_rays)
