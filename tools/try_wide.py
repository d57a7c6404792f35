"""First device run of the EXPERIMENTAL 4-wide trace kernel (option accel = 2).  Run it under a timeout --
a scheduling bug in a persistent kernel does not return:

    gpurun --timeout 120 -- 'timeout 90 python tools/try_wide.py'

Step 1 checks hit records against the oracle (primary rays at 64x64, 100 000 random rays, closest and any hit); step 2, only if
step 1 is bit-exact, renders the bench workload (C2, 1024x1024) at 64 spp with accel 1 and accel 2 and prints both timings."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from tests.helpers import golden_scene_object
from tests.refbind import Oracle, default_cam_array

oracle = Oracle(); cfg = Oracle.config("udpt"); CAM = default_cam_array()
m = yb.CUDAManager().setup(0)
sc = golden_scene_object("teapot", transmissive_teapot=True)
bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.uint32)
m.setOption("accel", 2)
r = yb.RendererCore(m, 64, 64)
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS") and r.setup(sc), m.last_message
tri, light, t = r.tracePrimary(1, 12345)
otri, olight, ot, _, _ = oracle.primary(cfg, CAM, sc.vert_data, sc.bvh, 12345, 1, 64, 64)
ok = bool((tri == otri).all() and (light == olight).all() and (bits(t) == bits(ot)).all())
print("primary rays bit-exact:", ok, flush=True)
rng = np.random.RandomState(3); n = 100000
o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
d = rng.normal(size=(n, 3)); d[:500, 0] = 0; d[500:1000, 1] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
od = np.concatenate([o, d], 1).astype(np.float32); tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
a = r.traceRays(od); b = oracle.trace(cfg, od, None, 0, sc.vert_data, sc.bvh)
ok2 = bool((a[0] == b[0]).all() and (a[1] == b[1]).all() and (bits(a[2]) == bits(b[2])).all())
sa = r.traceRays(od, tm, any_hit=True); sb = oracle.trace(cfg, od, tm, 1, sc.vert_data, sc.bvh)
ok3 = bool((((sb[0] >= 0) | (sb[1] >= 0)) == (sa[0] >= 0)).all())
print("random rays bit-exact:", ok2, " occlusion answers equal:", ok3, flush=True)
if not (ok and ok2 and ok3):
    sys.exit(1)
res = {}
for accel in (1, 2):
    m.setOption("accel", accel); m.setOption("time_stages", 4)
    r = yb.RendererCore(m, 1024, 1024)
    assert r.setup(sc), m.last_message
    r.enqueueKernels(8)                                   # warm-up: layout upload, pool allocation
    st = r.enqueueKernels(64, reset=True)
    k = max(st.timed_iterations, 1)
    res[accel] = dict(ms=st.render_ms, msamples_s=st.samples / st.render_ms / 1e3, avg_trace_ms=st.trace_ms / k, avg_shade_ms=st.shade_ms / k,
                      iterations=int(st.iterations), mean=float(r.readHDR()[..., :3].mean()))
print(json.dumps(res))
