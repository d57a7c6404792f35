"""ms per reference "frame" (1 spp over the image, src/RendererCore.cpp:483-486) and per small batches, C2 scene (measurement aid).
Second table: the same with option "pipeline" (a call returns when its samples are handed out; paths in flight are carried)."""
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
m.setOption("pipeline", 1)
for pool in (0, 1 << 20, 1 << 21, 1 << 22):
    m.setOption("pool_slots", pool)
    for spp in (1, 4):
        n = 96 // spp
        r.enqueueKernels(spp, reset=True)
        t0 = time.perf_counter(); dev = 0.0; its = 0
        for f in range(n):
            st = r.enqueueKernels(spp); dev += st.render_ms; its += st.iterations
        t1 = time.perf_counter()
        carried = st.carried_paths
        fin = r.finish().finish_ms
        print("pipelined  pool %8d  spp/call %d  %.3f ms/call wall  %.3f ms/call device  = %.3f ms/frame  %.0f Msamples/s  %.1f iterations/call  carried %d  finish %.2f ms" %
              (int(m.getOption("pool_slots_in_use")), spp, (t1 - t0) * 1e3 / n, dev / n, (t1 - t0) * 1e3 / n / spp, 1024 * 1024 * spp * n / (t1 - t0) / 1e6, its / n, carried, fin), flush=True)
