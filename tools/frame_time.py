"""ms per reference "frame" (1 spp over the image, src/RendererCore.cpp:483-486) and per small batches, C2 scene (measurement aid)."""
import os, sys, time
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
r.enqueueKernels(4)
for spp in (1, 2, 4, 8, 16, 64):
    ms = []
    for rep in range(5):
        st = r.enqueueKernels(spp, reset=(rep == 0))
        ms.append(st.render_ms)
    ms.sort()
    print("spp/call %3d  median %.2f ms  = %.2f ms/frame  %.0f Msamples/s  iterations %d  pool %d" %
          (spp, ms[2], ms[2] / spp, 1024 * 1024 * spp / ms[2] / 1e3, st.iterations, int(m.getOption("pool_slots_in_use"))), flush=True)
