"""BASELINE.json configs[0] (C1) and configs[2] (C3) at full size, for the record (run under gpurun; C2 is bench.py, C4 tools/run_c4.py)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from tests.refbind import load_golden_scene

def run(name, scene, prog, opts, W, H, spp, lights=None, oren_nayar=False):
    tris, mats, nodes = load_golden_scene(scene)
    m = yb.CUDAManager().setup(0)
    if oren_nayar:
        mats = mats.copy(); mats["alpha_x"] = 0.25            # sigma^2 of the diffuse walls
        m.setOption("oren_nayar", 1)
    r = yb.RendererCore(m, W, H)
    assert m.createRenderProgram(prog, compiler_opts=opts), m.last_message
    if lights is not None: m.setLightSources(lights)
    sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
    assert r.setup(sc), m.last_message
    r.enqueueKernels(4)
    m.setOption("time_stages", 8)
    st = r.enqueueKernels(spp, reset=True)
    n = max(st.timed_iterations, 1)
    img = r.readHDR()
    res = dict(config=name, scene=scene, program=prog, opts=opts, width=W, height=H, spp=spp, ms=st.render_ms,
               msamples_s=st.samples / st.render_ms / 1e3, mrays_s=(st.extend_rays + st.shadow_rays) / st.render_ms / 1e3,
               iterations=st.iterations, avg_trace_ms=st.trace_ms / n, avg_shade_ms=st.shade_ms / n,
               mean_luminance=float((0.212671 * img[..., 0] + 0.715160 * img[..., 1] + 0.072169 * img[..., 2]).mean()),
               nonfinite_pixels=int((~np.isfinite(img[..., :3]).all(-1)).sum()))
    print(json.dumps(res), flush=True)
    m.destroy() if hasattr(m, "destroy") else None
    return res

out = []
out.append(run("C1", "cornellbox", "udpt.cl", "", 512, 512, 64))
out.append(run("C1 x16 spp (steady state)", "cornellbox", "udpt.cl", "", 512, 512, 1024))
second = yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))       # same second light as tests/test_gpu_parity.py
out.append(run("C3", "teapot", "bdpt.cl", "-DMIS", 1024, 1024, 512, lights=np.concatenate([yb.LIGHT_BDPT, second]), oren_nayar=True))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_c1_c3.json"), "w"), indent=1)
