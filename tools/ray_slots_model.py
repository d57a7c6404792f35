"""Design aid, CPU only: the trace scheduling with R ray slots per lane (tests/hostcheck hc_trace_warp_multi), binary and 4-wide
tree, shipped knobs: warp-level steps per ray, lanes per step, modelled issue slots per ray.   python tools/ray_slots_model.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import load_golden_scene, ROOT
from tests.refbind import Oracle, default_cam_array, ptr
hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))
oracle = Oracle(); cfg = Oracle.config("udpt")
KNOBS = np.array([12, 24, 16, 8], np.int32)
tris, mats, nodes = load_golden_scene("teapot")
rng = np.random.RandomState(11); n = 40000
o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
od = np.concatenate([o, d], 1).astype(np.float32)
tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
base = {}
for accel in (1, 2):
    for R in (1, 2, 3):
        for kind, t, anyq in (("closest", None, 0), ("shadow", tm, 1)):
            util = np.zeros(4, np.uint64); tri = np.zeros(n, np.int32)
            assert hc.hc_trace_warp_multi(n, ptr(od), ptr(t) if t is not None else None, anyq, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(tri), None, accel, ptr(KNOBS), R, ptr(util)) == 0
            u = [float(v) for v in util]
            slots = (u[0] * (88 if accel == 1 else 134) + u[2] * 110) * 32 / n
            base.setdefault((accel, kind), tri)
            assert (tri == base[(accel, kind)]).all()
            print("accel %d R %d %-8s node %.2f (%.1f lanes) tri %.2f (%.1f lanes) slots/ray %.0f" % (accel, R, kind, u[0]*32/n, u[1]/u[0], u[2]*32/n, u[3]/u[2], slots))
