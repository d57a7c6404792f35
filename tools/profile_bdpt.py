"""Short C3 (bdpt.cl, -DMIS, two lights) render for ncu (a number printed under a profiler is never a bench value)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from tests.refbind import load_golden_scene
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
size = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
tris, mats, nodes = load_golden_scene("teapot")
mats = mats.copy(); mats["alpha_x"] = 0.25                  # C3: Oren-Nayar walls
m = yb.CUDAManager().setup(0)
for kv in sys.argv[3:]:
    k, v = kv.split("="); m.setOption(k, float(v))
m.setOption("oren_nayar", 1)
r = yb.RendererCore(m, size, size)
assert m.createRenderProgram("bdpt.cl", compiler_opts="-DMIS")
second = yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))
m.setLightSources(np.concatenate([yb.LIGHT_BDPT, second]))
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc)
m.setOption("time_stages", 4)
st = r.enqueueKernels(spp)
n = max(st.timed_iterations, 1)
print("timed iters", st.timed_iterations, "avg shade ms", st.shade_ms / n, "avg trace ms", st.trace_ms / n)
print("spp", spp, "ms", st.render_ms, "Msamples/s", st.samples / st.render_ms / 1e3, "iters", st.iterations, "ext", st.extend_rays, "shadow", st.shadow_rays)
