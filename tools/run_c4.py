"""configs[3] ("C4") at full size: ~10.5 M triangles, 3840x2160.  Development / measurement aid (run under gpurun)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from yune_b200.scenes import synthetic_c4
from tests.refbind import load_golden_scene
subdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 9
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
W, H = (3840, 2160) if len(sys.argv) <= 3 else (int(sys.argv[3]), int(sys.argv[4]))
tris, mats, _ = load_golden_scene("cornellbox")
t0 = time.time(); T = synthetic_c4(tris, subdiv); t1 = time.time()
sc = yb.Scene().setGeometry(T, mats); t2 = time.time()
print("triangles", T.size, "nodes", sc.bvh.size, "gen %.1fs  BVH build %.1fs" % (t1 - t0, t2 - t1), flush=True)
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, W, H)
assert m.createRenderProgram("udpt.cl")
t3 = time.time(); assert r.setup(sc), m.last_message
st = r.enqueueKernels(1); t4 = time.time()
print("upload + relayout + first frame %.1fs" % (t4 - t3), flush=True)
m.setOption("time_stages", 4)
st = r.enqueueKernels(spp, reset=True)
n = max(st.timed_iterations, 1)
res = dict(triangles=int(T.size), nodes=int(sc.bvh.size), width=W, height=H, spp=spp, ms=st.render_ms, msamples_s=st.samples / st.render_ms / 1e3,
           mrays_s=(st.extend_rays + st.shadow_rays) / st.render_ms / 1e3, iterations=st.iterations, extend_rays=st.extend_rays, shadow_rays=st.shadow_rays,
           avg_trace_ms=st.trace_ms / n, avg_shade_ms=st.shade_ms / n, bvh_build_s=t2 - t1)
img = r.readHDR()
res["mean_luminance"] = float((0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]).mean())
res["nonfinite_pixels"] = int((~np.isfinite(img[..., :3]).all(-1)).sum())
print(json.dumps(res))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "c4_subdiv%d.json" % subdiv), "w"), indent=1)
np.save(os.path.join(ROOT, "gpurun_out", "c4_preview.npy"), img[::8, ::8].copy())
