import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from tests.helpers import golden_scene_object
from tests.refbind import Oracle, default_cam_array
W = 80
m = yb.CUDAManager().setup(0)
sc = golden_scene_object("cornellbox")
r = yb.RendererCore(m, W, W)
assert m.createRenderProgram("bdpt.cl"); assert r.setup(sc)
o = Oracle(); cam = default_cam_array()
out = {}
for bounces in (2, 3, 20):
    m.setOption("bdpt_bounces", bounces)
    m.check(r._lib.yune_render(r._ctx, 0, 1, 1, 555, 1))
    out["ours_b%d" % bounces] = r.readSum()
    out["ref_b%d" % bounces] = o.samples(Oracle.config("bdpt", rng_mode=1, seed=555, bdpt_bounces=bounces), cam, sc.vert_data, sc.mat_data, sc.bvh, W, W, 0)
np.savez(os.path.join(ROOT, "gpurun_out", "dbg_bdpt.npz"), **out)
