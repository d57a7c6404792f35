"""Design aid, CPU only: WARP-level steps of the trace scheduling (tests/hostcheck hc_trace_warp = the host restatement of
k_trace's postponed leaves / votes / refill) over the shipped binary own tree (accel 1) and over the 4-wide collapse of it
(accel 2), same rays, same knobs.  Prints node steps and triangle steps per ray as a warp executes them (a step costs its
issue slots whether 5 or 32 lanes take part) and the lanes active in them.   python tools/warp_step_model.py"""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.helpers import load_golden_scene, ROOT
from tests.refbind import Oracle, default_cam_array, ptr

hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))
oracle = Oracle(); cfg = Oracle.config("udpt")
KNOBS = np.array([12, 24, 16, 8], np.int32)          # refill_idle, phase_min, inner_min, inner_chain: the shipped defaults
NODE_SLOTS = {1: 88.0, 2: 134.0}                     # issue slots of one node step: measured (binary) / estimated (4-wide), DESIGN.md 10
TRI_SLOTS = 110.0
for scene in ("teapot", "cornellbox"):
    tris, mats, nodes = load_golden_scene(scene)
    W = 192
    _, _, _, od_p, _ = oracle.primary(cfg, default_cam_array(), tris, nodes, 12345, 1, W, W)
    rng = np.random.RandomState(11); n = 40000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od_r = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    for name, od, t, anyq in (("primary (coherent)", od_p, None, 0), ("random interior", od_r, None, 0), ("shadow segments", od_r, tm, 1)):
        m = od.shape[0]; res = {}
        for accel in (1, 2):
            util = np.zeros(4, np.uint64); tri = np.zeros(m, np.int32)
            assert hc.hc_trace_warp(m, ptr(od), ptr(t) if t is not None else None, anyq, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size),
                                    ptr(tri), None, accel, ptr(KNOBS), ptr(util)) == 0
            u = [float(x) for x in util]
            slots = (u[0] * NODE_SLOTS[accel] + u[2] * TRI_SLOTS) * 32 / m          # issue slots per ray at 32 rays per warp
            res[accel] = (u[0] * 32 / m, u[1] / max(u[0], 1), u[2] * 32 / m, u[3] / max(u[2], 1), slots, tri)
        assert (res[1][5] == res[2][5]).all()
        for accel in (1, 2):
            r = res[accel]
            print("%-10s %-20s %s: node steps/ray %.2f (%.1f lanes)  tri steps/ray %.2f (%.1f lanes)  modelled slots/ray %.0f" %
                  (scene, name, "binary" if accel == 1 else "4-wide", r[0], r[1], r[2], r[3], r[4]))
        print("%-10s %-20s 4-wide / binary = %.3f" % (scene, name, res[2][4] / res[1][4]))
