"""Short render of a BASELINE config for ncu (a number printed under a profiler is never a bench value).

    python tools/profile_run.py <spp> <size or WxH> [option=value ...] [config=c2|c3|c4]

Prints the steady-state rays per launch (pool full) so that per-ray figures can be derived from the captured launch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from bench import build_scene, CONFIGS
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 4
size = sys.argv[2] if len(sys.argv) > 2 else "1024"
W, H = (int(size.split("x")[0]), int(size.split("x")[1])) if "x" in size else (int(size), int(size))
opts = dict(kv.split("=") for kv in sys.argv[3:])
cfg = dict(CONFIGS[opts.pop("config", "c2")])
tris, mats, nodes, lights = build_scene(cfg)
m = yb.CUDAManager().setup(0)
for k, v in opts.items():
    m.setOption(k, float(v))
r = yb.RendererCore(m, W, H)
assert m.createRenderProgram(cfg["kernel"], compiler_opts=cfg["opts"])
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc), m.last_message
if lights is not None:
    assert m.setLightSources(lights)
m.setOption("oren_nayar", 1 if cfg.get("oren_nayar") else 0)
st = r.enqueueKernels(spp)
n = max(st.timed_iterations, 1)
print("timed iters", st.timed_iterations, "avg shade ms", st.shade_ms / n, "avg trace ms", st.trace_ms / n)
print("spp", spp, "ms", st.render_ms, "Msamples/s", st.samples / st.render_ms / 1e3, "iters", st.iterations, "ext", st.extend_rays, "shadow", st.shadow_rays)
si = max(st.steady_iterations, 1)
print("STEADY iterations %d ext_per_launch %.1f shadow_per_launch %.1f pool_slots %d" % (st.steady_iterations, st.steady_extend_rays / si, st.steady_shadow_rays / si, int(m.getOption("pool_slots_in_use"))))
