"""Design aid, CPU only: steps per ray if the own traversal tree were collapsed into nodes of 2 (shipped) / 3 / 4 / 6 / 8 children
(tests/hostcheck hc_wide_stats).  python tools/wide_tree_model.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C, os
from tests.helpers import load_golden_scene, ROOT
from tests.refbind import Oracle, default_cam_array, ptr
hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))
oracle = Oracle(); cfg = Oracle.config("udpt")
for scene in ("teapot", "cornellbox"):
    tris, mats, nodes = load_golden_scene(scene)
    W = 256
    _, _, _, od_p, _ = oracle.primary(cfg, default_cam_array(), tris, nodes, 12345, 1, W, W)
    rng = np.random.RandomState(11); n = 100000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od_r = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    for width in (2, 3, 4, 6, 8):
        line = "%s width %d |" % (scene, width)
        for name, od, t, anyq in (("primary", od_p, None, 0), ("random", od_r, None, 0), ("shadow", od_r, tm, 1)):
            out = np.zeros(4, np.uint64)
            assert hc.hc_wide_stats(od.shape[0], ptr(od), ptr(t) if t is not None else None, anyq, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), width, ptr(out)) == 0
            m = od.shape[0]
            line += " %s visits %.2f boxes %.2f tri %.2f hit %.3f |" % (name, out[0] / m, out[1] / m, out[2] / m, out[3] / m)
        print(line)
