"""Sweep engine knobs on the C2 workload (development aid)."""
import itertools, os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from bench import load_scene
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grid = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {"trace_block": [256, 512, 1024], "phase_min": [8, 12, 16], "smem_nodes": [1099, 512]}
tris, mats, nodes = load_scene()
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, 1024, 1024)
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS")
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = tris, mats, nodes
assert r.setup(sc)
r.enqueueKernels(8)
keys = list(grid)
res = []
for combo in itertools.product(*[grid[k] for k in keys]):
    for k, v in zip(keys, combo):
        m.setOption(k, v)
    try:
        st = r.enqueueKernels(spp, reset=True)
        v = st.samples / st.render_ms / 1e3
    except yb.YuneError as e:
        v = 0.0; print("fail", combo, e)
    res.append((v, combo))
    print(dict(zip(keys, combo)), "%.1f Msamples/s  iters %d" % (v, st.iterations), flush=True)
print("best", max(res))
