"""Traversal LOGIC of the product (trace_core.h, lights.h, relayout.cpp compiled for the host by tests/hostcheck --
a test-only build, never loaded by the package) against the oracle: depth-first ordered walk with pruning must return
the reference's breadth-first answer bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from tests.helpers import ROOT, GOLDEN, load_golden_scene
from tests.refbind import NODE_DTYPE, TRI_DTYPE as TRI_DTYPE_, Oracle, default_cam_array, ptr

LIGHT_UDPT = np.zeros(32, np.float32)
LIGHT_UDPT[0:4] = [-0.1979, 0.92, -3.1972, 1]; LIGHT_UDPT[4:8] = [0, -1, 0, 0]; LIGHT_UDPT[8:12] = [16, 16, 16, 0]
LIGHT_UDPT[20:24] = [0.4, 0, 0, 0]; LIGHT_UDPT[24:28] = [0, 0, 0.4, 0]


@pytest.fixture(scope="module")
def hc():
    return C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))


def _trace(hc, od, tmax, any_hit, tris, nodes, leaf_split=0, accel=0):
    n = od.shape[0]
    tri = np.zeros(n, np.int32); light = np.zeros(n, np.int32); t = np.zeros(n, np.float32); work = np.zeros(2, np.uint64)
    rc = hc.hc_trace(n, ptr(od), ptr(tmax), int(any_hit), ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(LIGHT_UDPT), 1,
                     ptr(tri), ptr(light), ptr(t), ptr(work), int(leaf_split), int(accel))
    assert rc == 0
    return tri, light, t, work


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_primary_rays(hc, oracle, scene):
    tris, mats, nodes = load_golden_scene(scene)
    g = np.load(os.path.join(GOLDEN, "primary_%s.npz" % scene))
    W = int(g["width"])
    for jm in (0, 1):
        otri, olight, ot, od, _ = oracle.primary(Oracle.config("udpt"), default_cam_array(), tris, nodes, int(g["rand"]), jm, W, W)
        tri, light, t, work = _trace(hc, od, None, 0, tris, nodes)
        assert (tri == g["tri_j%d" % jm]).all() and (light == g["light_j%d" % jm]).all()
        assert (t.view(np.uint32) == ot.view(np.uint32)).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_random_rays_closest_and_any(hc, oracle, scene):
    tris, mats, nodes = load_golden_scene(scene)
    rng = np.random.RandomState(11); n = 60000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:500, 0] = 0; d[500:1000, 1] = 0; d[1000:1500, 2] = 0; d[1500:1600, :2] = 0      # exercise the NaN-guarded slabs
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    tri, light, t, _ = _trace(hc, od, None, 0, tris, nodes)
    assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all()
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, nodes)
    atri, alight, _, _ = _trace(hc, od, tm, 1, tris, nodes)
    assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
@pytest.mark.parametrize("leaf_split,accel", [(2, 0), (1, 0), (0, 1), (0, 2)])
def test_refined_leaves_and_own_tree_return_the_same_hits(hc, oracle, scene, leaf_split, accel):
    """Both accelerations of the reference walk -- padded subtrees inside big leaves (accel 0, leaf_split) and our own SAH
    tree with the exact leaf-box filter (accel 1) -- must give the reference's hit record bit for bit while testing far
    fewer triangles."""
    tris, mats, nodes = load_golden_scene(scene)
    rng = np.random.RandomState(23); n = 40000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d[:300, 0] = 0; d[300:600, 1] = 0; d[600:900, 2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    tri, light, t, work = _trace(hc, od, None, 0, tris, nodes, leaf_split, accel)
    assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all()
    _, _, _, work0 = _trace(hc, od, None, 0, tris, nodes, 0, 0)
    assert work[1] < 0.6 * work0[1]                       # triangle tests
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, nodes)
    atri, alight, _, _ = _trace(hc, od, tm, 1, tris, nodes, leaf_split, accel)
    assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()


def test_layout_rejects_malformed_input(hc):
    tris, mats, nodes = load_golden_scene("cornellbox")
    od = np.zeros((1, 6), np.float32); od[0, 5] = -1
    bad = nodes.copy(); bad["child_idx"][0] = 10 ** 6
    tri = np.zeros(1, np.int32); light = np.zeros(1, np.int32)
    assert hc.hc_trace(1, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(bad), int(bad.size), ptr(LIGHT_UDPT), 1, ptr(tri), ptr(light), None, None, 0, 0) == -1
    bad = nodes.copy(); leaf = int(np.nonzero(bad["vert_len"] > 0)[0][0]); bad["vert_list"][leaf, 0] = 9999
    assert hc.hc_trace(1, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(bad), int(bad.size), ptr(LIGHT_UDPT), 1, ptr(tri), ptr(light), None, None, 0, 0) == -1


@pytest.mark.parametrize("accel", [0, 1, 2])
def test_device_scheduling_restatement_is_order_free(hc, oracle, accel):
    """The DEVICE walk postpones leaves, votes per step and refills idle lanes (kernels.cu); tests/hostcheck restates that
    scheduling for simulated 32-lane warps on top of trace_core.h's arithmetic.  Whatever the knobs, and whichever rays share a
    warp, the hit records must equal the oracle's bit for bit (closest hit, exact ties by reference rank) and the occlusion
    answers must be the same."""
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(31); n = 20000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d[:200, 0] = 0; d[200:400, 1] = 0; d[400:600, 2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    cfg = Oracle.config("udpt")
    # the oracle's traceRay runs the analytic light first; compare on rays the light does not bound
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, nodes)
    free = olight < 0
    lanes = {}
    for knobs in ((12, 24, 16, 8), (1, 1, 1, 0), (32, 32, 33, 0), (6, 8, 1, 64), (20, 2, 30, 3)):
        k = np.array(knobs, np.int32); util = np.zeros(4, np.uint64)
        tri = np.zeros(n, np.int32); t = np.zeros(n, np.float32)
        assert hc.hc_trace_warp(n, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(tri), ptr(t), accel, ptr(k), ptr(util)) == 0
        assert (tri[free] == otri[free]).all(), knobs
        assert (t[free].view(np.uint32) == ot[free].view(np.uint32)).all(), knobs
        lanes[knobs] = (util[1] / max(util[0], 1), util[3] / max(util[2], 1))
        occ = np.zeros(n, np.int32)
        assert hc.hc_trace_warp(n, ptr(od), ptr(tm), 1, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(occ), None, accel, ptr(k), None) == 0
        sfree = slight < 0
        assert ((occ[sfree] >= 0) == (stri[sfree] >= 0)).all(), knobs
    # the default knobs keep most lanes of a node step busy; one-lane-at-a-time settings do not (sanity of the statistics)
    assert lanes[(12, 24, 16, 8)][0] > 16 and lanes[(12, 24, 16, 8)][0] > lanes[(32, 32, 33, 0)][0]


def test_threaded_layout_build_equals_sequential(hc):
    """relayout.cpp builds the two subtrees of big nodes of its own SAH tree on two threads; the traversal layout (pair records
    in breadth-first order, triangles in leaf order, shading records, leaf boxes) must not depend on it."""
    import yune_b200 as yb
    from yune_b200.scenes import synthetic_c4
    tris, mats, nodes = load_golden_scene("cornellbox")
    sc = yb.Scene().setGeometry(synthetic_c4(tris, 6), mats)          # 163 880 triangles: the parallel path starts at 65 536
    h = []
    for threads in (None, "1"):
        if threads: os.environ["YUNE_BVH_THREADS"] = threads
        try:
            out = np.zeros(4, np.uint64)
            assert hc.hc_layout_hash(ptr(sc.vert_data), int(sc.vert_data.size), ptr(sc.bvh), int(sc.bvh.size), 0, 1, ptr(out)) == 0
        finally:
            os.environ.pop("YUNE_BVH_THREADS", None)
        h.append(out.tolist())
    assert h[0] == h[1] and all(x != 0 for x in h[0])


@pytest.mark.parametrize("kind,n,seed,offset", [("uniform", 1500, 31, 0.0), ("clustered", 1200, 32, 0.0), ("flats", 1000, 33, 0.0),
                                                ("mixed", 1800, 34, 0.0), ("uniform", 1500, 35, 1000.0), ("flats", 1000, 36, 10000.0)])
def test_random_soups_closest_and_any(hc, oracle, kind, n, seed, offset):
    """Geometry the shipped scenes do not have -- overlapping boxes everywhere, axis-aligned flats whose boxes carry the
    reference's +0.2 padding, slivers over four decades of size, empty children -- walked by the reference's breadth-first
    queue (oracle) and by the product's three walks (reference tree, refined leaves, own tree + leaf-box filter): same hit
    record bit for bit, same occlusion answer.  `offset` moves the scene away from the origin (coordinates of 1e3 / 1e4 with
    0.05-sized triangles: the own-tree slab test o * (1/d) cancels heavily and still must never lose a reference hit)."""
    import yune_b200 as yb
    from tests.helpers import random_soup, soup_rays
    rng = np.random.default_rng(seed)
    T = random_soup(rng, n, kind)
    for k in ("v1", "v2", "v3"):
        T[k][:, :3] += np.float32([offset, -offset / 2, 2 * offset])
    sc = yb.Scene().setGeometry(T, load_golden_scene("cornellbox")[1])
    tris, nodes = sc.vert_data, sc.bvh
    od, tm = soup_rays(rng, T, 30000)
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    assert (otri >= 0).mean() > 0.3
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, nodes)
    for leaf_split, accel in ((0, 0), (2, 0), (0, 1), (0, 2)):
        tri, light, t, _ = _trace(hc, od, None, 0, tris, nodes, leaf_split, accel)
        assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all(), (leaf_split, accel)
        atri, _, _, _ = _trace(hc, od, tm, 1, tris, nodes, leaf_split, accel)
        assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all(), (leaf_split, accel)


@pytest.mark.parametrize("scene", ["cornellbox", "teapot", "soup"])
def test_brute_force_mode_bvh_size_zero(hc, oracle, scene):
    """Kernel arg 6 bvh_size == 0: the reference intersects every triangle in index order with no box test at all
    (udpt.cl:280-284).  The product answers that mode with its own tree and a filter that always passes; hit records must be
    those of the reference's loop bit for bit -- they differ from the BVH mode's where a box test rejected a grazing hit or
    an exact tie went to the other triangle -- and the occlusion answers must agree."""
    from tests.helpers import random_soup, soup_rays
    from tests.refbind import have_ref, RefKernels
    rng = np.random.default_rng(77)
    if scene == "soup":
        tris = random_soup(rng, 1200, "flats")
        od, tm = soup_rays(rng, tris, 20000)
    else:
        tris, mats, nodes = load_golden_scene(scene)
        n = 6000 if scene == "teapot" else 60000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:300, 0] = 0; d[300:600, 1] = 0; d[600:900, 2] = 0
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        od = np.concatenate([o, d], 1).astype(np.float32)
        tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    none = np.zeros(0, NODE_DTYPE)
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, none)
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, none)
    if have_ref():      # the oracle's loop is the reference's: same answers from the compiled kernel text
        rtri, rlight, rt = RefKernels().trace("udpt", od, None, 0, tris, none)
        assert (rtri == otri).all() and (rlight == olight).all() and (rt.view(np.uint32) == ot.view(np.uint32)).all()
    for accel in (1, 0, 2):                               # 0 is answered like 1 in this mode (there is no reference tree to walk)
        tri, light, t, work = _trace(hc, od, None, 0, tris, none, 0, accel)
        assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all()
        assert work[1] < 0.2 * od.shape[0] * tris.size        # and nowhere near n_rays x n_triangles tests
        atri, _, _, _ = _trace(hc, od, tm, 1, tris, none, 0, accel)
        assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()
    # the device scheduling restatement on the same layout
    tri_w = np.zeros(od.shape[0], np.int32); t_w = np.zeros(od.shape[0], np.float32)
    knobs = np.array([12, 24, 16, 8], np.int32)
    assert hc.hc_trace_warp(od.shape[0], ptr(od), None, 0, ptr(tris), int(tris.size), None, 0, ptr(tri_w), ptr(t_w), 1, ptr(knobs), None) == 0
    free = olight < 0
    assert (tri_w[free] == otri[free]).all()


def test_wide_tree_model_finds_the_same_hits(hc):
    """tests/hostcheck hc_wide_stats (design aid behind DESIGN.md's wide-node figures): collapsing the own tree into nodes of up
    to 4 or 8 children must not change which rays hit, and must cut the node visits roughly in half at equal box tests."""
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(5); n = 20000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    res = {}
    for width in (2, 4, 8):
        out = np.zeros(4, np.uint64)
        assert hc.hc_wide_stats(n, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), width, ptr(out)) == 0
        res[width] = [int(x) for x in out]
    assert res[2][3] == res[4][3] == res[8][3]
    assert res[4][0] < 0.6 * res[2][0] and res[4][1] < 1.1 * res[2][1] and res[8][0] < res[4][0]
    assert hc.hc_wide_stats(n, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 9, ptr(out)) == -1


@pytest.mark.parametrize("accel", [1, 2])
def test_ray_slots_model_is_order_free(hc, oracle, accel):
    """hc_trace_warp_multi (design aid behind DESIGN.md's ray-slot figures): two or three ray slots per lane change when a ray
    advances, never its hit record; lanes per triangle step go up."""
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(41); n = 12000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    otri, olight, ot = oracle.trace(Oracle.config("udpt"), od, None, 0, tris, nodes)
    free = olight < 0
    k = np.array([12, 24, 16, 8], np.int32); lanes = {}
    for R in (1, 2, 3):
        tri = np.zeros(n, np.int32); t = np.zeros(n, np.float32); util = np.zeros(4, np.uint64)
        assert hc.hc_trace_warp_multi(n, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(tri), ptr(t), accel, ptr(k), R, ptr(util)) == 0
        assert (tri[free] == otri[free]).all() and (t[free].view(np.uint32) == ot[free].view(np.uint32)).all(), R
        lanes[R] = float(util[3]) / float(util[2])
    assert lanes[2] > lanes[1] + 4 and lanes[3] > lanes[2]
    assert hc.hc_trace_warp_multi(n, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(tri), ptr(t), accel, ptr(k), 5, None) == -1


def test_wide_layout_shape(hc):
    """accel 2 layout of the teapot scene: about half as many records as the binary tree, two thirds of its depth (stack bound 3 per level), and small
    enough for the shared-memory staging area (3900 x 56 B today)."""
    tris, mats, nodes = load_golden_scene("teapot")
    b = np.zeros(4, np.int32); w = np.zeros(4, np.int32)
    assert hc.hc_layout_info(ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 0, 1, ptr(b)) == 0
    assert hc.hc_layout_info(ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 0, 2, ptr(w)) == 0
    assert b[0] == w[0] and b[1] == w[1] == tris.size           # same binary tree underneath, every triangle in a leaf
    assert 0.4 * b[0] < w[3] < 0.55 * b[0] and w[2] <= 0.7 * b[2] and 3 * w[2] + 2 <= 64
    assert w[3] * 112 <= 3900 * 56


def _soup_with_duplicates(seed, n):
    """A soup in which every triangle exists twice (indices i and i + n): every hit is an exact tie in t, so the hit index
    says which copy the walk kept -- the reference keeps the one its breadth-first queue reaches first (udpt.cl:373)."""
    import yune_b200 as yb
    from tests.helpers import random_soup, soup_rays
    rng = np.random.default_rng(seed)
    T = random_soup(rng, n, "uniform")
    T2 = np.concatenate([T, T])
    sc = yb.Scene().setGeometry(T2, load_golden_scene("cornellbox")[1])
    od, tm = soup_rays(rng, T, 20000)
    return sc.vert_data, sc.bvh.copy(), od, tm


def test_hand_made_tree_whose_boxes_do_not_nest_falls_back_to_the_reference_walk(hc, oracle):
    """accel 1 decides 'the reference would have reached this triangle' from the leaf box alone, which needs every box
    to contain its children's boxes.  A tree that does not nest (here: inner boxes shrunk, so the reference prunes subtrees whose
    leaves a ray would pass) must be walked as uploaded: the layout reports accel 0 and the hits are the oracle's."""
    tris, nodes, od, tm = _soup_with_duplicates(5, 1500)
    inner = np.nonzero(nodes["child_idx"] > 0)[0]
    victims = inner[(inner > 0)][::3]
    mid = 0.5 * (nodes["p_min"][victims] + nodes["p_max"][victims])
    nodes["p_min"][victims] = mid - 0.25 * (mid - nodes["p_min"][victims])
    nodes["p_max"][victims] = mid + 0.25 * (nodes["p_max"][victims] - mid)
    for accel in (1, 2):
        assert hc.hc_layout_accel(ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 0, accel) == 0
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    for accel in (0, 1, 2):
        tri, light, t, _ = _trace(hc, od, None, 0, tris, nodes, 0, accel)
        assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all(), accel


def test_tie_rank_follows_the_breadth_first_queue_not_the_node_index(hc, oracle):
    """the visiting rank that breaks exact ties must be the position in the reference's breadth-first queue.  Here
    two DIFFERENT reference-built trees over the same triangles (copy i in tree A, copy i + n in tree B) hang under one root, A's
    nodes stored before B's: child indices are after their parents but the array is not in breadth-first order.  Every hit is an
    exact tie between the two copies; the queue reaches the shallower leaf first, the node index always says A."""
    import yune_b200 as yb
    from tests.helpers import random_soup, soup_rays
    rng = np.random.default_rng(6)
    n = 1500
    T = random_soup(rng, n, "uniform")
    mats = load_golden_scene("cornellbox")[1]
    A = yb.Scene().setGeometry(T, mats, bvh_bins=20).bvh.copy()
    B = yb.Scene().setGeometry(T, mats, bvh_bins=3).bvh.copy()
    na, nb = A.size, B.size
    nodes = np.zeros(1 + na + nb, NODE_DTYPE)
    a_inner, b_inner = A["child_idx"] > 0, B["child_idx"] > 0
    A["child_idx"][a_inner] += 2
    B["child_idx"][b_inner] += na + 1
    b_leaf = (B["child_idx"] == -1) & (B["vert_len"] > 0)
    for j in range(10):
        m = b_leaf & (B["vert_len"] > j)
        B["vert_list"][m, j] += n
    nodes[1], nodes[2], nodes[3:2 + na], nodes[2 + na:] = A[0], B[0], A[1:], B[1:]
    nodes["p_min"][0] = np.minimum(A["p_min"][0], B["p_min"][0]); nodes["p_max"][0] = np.maximum(A["p_max"][0], B["p_max"][0])
    nodes["child_idx"][0] = 1; nodes["vert_len"][0] = 0
    tris = np.concatenate([T, T])
    od, tm = soup_rays(rng, T, 20000)
    assert hc.hc_layout_accel(ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 0, 1) == 1       # boxes nest
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    assert ((otri >= n).sum() > 100) and ((otri >= 0) & (otri < n)).sum() > 100                         # both copies win somewhere
    for leaf_split, accel in ((0, 0), (2, 0), (0, 1), (0, 2)):
        tri, light, t, _ = _trace(hc, od, None, 0, tris, nodes, leaf_split, accel)
        assert (tri == otri).all() and (t.view(np.uint32) == ot.view(np.uint32)).all(), (leaf_split, accel)


# ---- option "isect" = 1: the watertight perf-mode intersection (north star; SURVEY 7 "ship both") ----
WT = 1 | 256        # hc_trace's accel argument: own tree, bit 8 = watertight


def _icosphere(level, radius=1.0, centre=(0.0, 0.0, 0.0)):
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(level):
        cache, nf = {}, []
        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    V = (np.array(v) * radius + np.array(centre)).astype(np.float32)       # ONE float32 position per vertex, shared by its triangles
    return V, np.array(f, np.int32)


def test_watertight_mode_matches_parity_mode_except_at_edges(hc):
    """isect 1 against isect 0 on the shipped scenes: the hit triangle may differ only for rays within rounding distance of an
    edge / vertex (or of a reference box face, where the reference itself drops the hit), at most 1e-5 of the rays;
    t agrees to 4e-6 absolute + 2e-6 relative (both tests are accurate to a few ulps of the vertex coordinates)."""
    for scene in ("cornellbox", "teapot"):
        tris, mats, nodes = load_golden_scene(scene)
        rng = np.random.RandomState(11); n = 300000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        od = np.concatenate([o, d], 1).astype(np.float32)
        t0, l0, tt0, _ = _trace(hc, od, None, 0, tris, nodes, 0, 1)
        t1, l1, tt1, _ = _trace(hc, od, None, 0, tris, nodes, 0, WT)
        assert (t0 != t1).mean() <= 1e-5, scene
        same = (t0 == t1) & (t0 >= 0)
        assert (np.abs(tt0[same] - tt1[same]) <= 4e-6 + 2e-6 * np.abs(tt0[same])).all()
        tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
        a0, _, _, _ = _trace(hc, od, tm, 1, tris, nodes, 0, 1)
        a1, _, _, _ = _trace(hc, od, tm, 1, tris, nodes, 0, WT)
        assert ((a0 >= 0) != (a1 >= 0)).mean() <= 1e-4        # a segment that ends within rounding distance of a surface may flip


def test_watertight_mode_leaks_no_ray_through_shared_edges_and_vertices(hc):
    """The property that names the mode: rays from inside a closed mesh aimed EXACTLY at its vertices and at points on its edges
    (float32 positions, float32 directions) all hit something in isect 1.  (The reference's Moller-Trumbore decides u, v >= 0 and
    u + v <= 1 per triangle with differently rounded edge vectors and lets some of these rays through; that count is only
    reported, it is the reference's behaviour.)"""
    import yune_b200 as yb
    V, F = _icosphere(4)                                         # 5120 triangles, 2562 shared vertices
    T = np.zeros(F.shape[0], TRI_DTYPE_)
    for k, name in enumerate(("v1", "v2", "v3")):
        T[name][:, :3] = V[F[:, k]]; T[name][:, 3] = 1.0
        T["vn" + name[1]][:, :3] = V[F[:, k]]
    sc = yb.Scene().setGeometry(T, load_golden_scene("cornellbox")[1])
    tris, nodes = sc.vert_data, sc.bvh
    rng = np.random.default_rng(3)
    origins = rng.uniform(-0.3, 0.3, (8, 3)).astype(np.float32)
    e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
    w = rng.uniform(0, 1, (e.shape[0], 1)).astype(np.float32)
    targets = np.concatenate([V, (V[e[:, 0]] * w + V[e[:, 1]] * (1 - w)).astype(np.float32), 0.5 * (V[e[:, 0]] + V[e[:, 1]])]).astype(np.float32)
    leaks_wt = leaks_mt = total = 0
    for o in origins:
        d = targets - o
        d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        od = np.concatenate([np.broadcast_to(o, d.shape), d], 1).astype(np.float32)
        t1, _, _, _ = _trace(hc, od, None, 0, tris, nodes, 0, WT)
        t0, _, _, _ = _trace(hc, od, None, 0, tris, nodes, 0, 1)
        leaks_wt += int((t1 < 0).sum()); leaks_mt += int((t0 < 0).sum()); total += od.shape[0]
    print("rays at vertices / edges: %d, leaked in parity mode (reference behaviour): %d, in watertight mode: %d" % (total, leaks_mt, leaks_wt))
    assert total > 100000 and leaks_wt == 0


def test_watertight_mode_needs_the_own_tree(hc):
    tris, mats, nodes = load_golden_scene("cornellbox")
    od = np.zeros((1, 6), np.float32); od[0, 5] = -1
    tri = np.zeros(1, np.int32); light = np.zeros(1, np.int32); t = np.zeros(1, np.float32); work = np.zeros(2, np.uint64)
    for accel in (0, 2):
        assert hc.hc_trace(1, ptr(od), None, 0, ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), ptr(LIGHT_UDPT), 1,
                           ptr(tri), ptr(light), ptr(t), ptr(work), 0, accel | 256) == -1


def test_reference_leaves_for_the_device_layout(hc):
    """relayout.cpp: referenceLeavesForDevice -- what the device layout path (option device_layout, bvh_build.cu) takes from an
    uploaded tree: every triangle's reference leaf and visiting rank (leaves in the order of the reference's breadth-first queue,
    slots in vert_list order) and the leaves' uploaded boxes; trees whose boxes do not nest, or that list a triangle twice or
    not at all, are left to the host path."""
    tris, mats, nodes = load_golden_scene("teapot")
    n = int(tris.size)
    leaf = np.zeros(n, np.int32); rank = np.zeros(n, np.int32); nl = C.c_int(0)
    boxes = np.zeros((int((nodes["vert_len"] > 0).sum()), 2, 4), np.float32)
    assert hc.hc_reference_leaves(ptr(tris), n, ptr(nodes), int(nodes.size), ptr(leaf), ptr(rank), ptr(boxes), C.byref(nl)) == 1
    is_leaf = (nodes["child_idx"] == -1) & (nodes["vert_len"] > 0)
    ids = np.nonzero(is_leaf)[0]                                   # reference-built: node-index order == queue order
    assert nl.value == ids.size
    assert np.array_equal(np.sort(rank), np.arange(n))
    k = 0
    for li, i in enumerate(ids):
        for j in range(int(nodes["vert_len"][i])):
            t = int(nodes["vert_list"][i, j])
            assert leaf[t] == li and rank[t] == k
            k += 1
        assert (boxes[li, 0, :3] == nodes["p_min"][i, :3]).all() and (boxes[li, 1, :3] == nodes["p_max"][i, :3]).all()
    # a triangle listed by two leaves / boxes that do not nest: not usable (the host path handles them)
    twice = nodes.copy(); a, b = ids[0], ids[1]
    twice["vert_list"][b, 0] = twice["vert_list"][a, 0]
    assert hc.hc_reference_leaves(ptr(tris), n, ptr(twice), int(twice.size), ptr(leaf), ptr(rank), None, C.byref(nl)) == 0
    shrunk = nodes.copy(); inner = np.nonzero(nodes["child_idx"] > 0)[0][1]
    shrunk["p_max"][inner, :3] = shrunk["p_min"][inner, :3] + 1e-3
    assert hc.hc_reference_leaves(ptr(tris), n, ptr(shrunk), int(shrunk.size), ptr(leaf), ptr(rank), None, C.byref(nl)) == 0
    bad = nodes.copy(); bad["child_idx"][0] = nodes.size + 5
    assert hc.hc_reference_leaves(ptr(tris), n, ptr(bad), int(bad.size), ptr(leaf), ptr(rank), None, C.byref(nl)) == -1


def test_rays_whose_origin_over_direction_overflows(hc, oracle):
    """1/d finite but o * (1/d) = inf (|d_k| ~ 1e-38 next to |o_k| ~ 1e1): the fused slab test of the own tree would
    see -inf on both planes of every box.  Such rays take the guarded form; hits must stay the oracle's."""
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(3); n = 6000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1) * 8.0 + np.array([0, 0, 21.0])
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    tiny = np.float32([1e-38, 3e-38, -2e-38, 1.2e-38])
    for k in range(3):
        d[k * 2000:(k + 1) * 2000, k] = tiny[rng.randint(0, 4, 2000)]
    od = np.concatenate([o, d], 1).astype(np.float32)
    inv = 1.0 / od[:, 3:].astype(np.float64)
    assert np.isfinite((1.0 / od[:, 3:]).astype(np.float32)).all() and (np.abs(od[:, :3] * inv.astype(np.float32)) == np.inf).any()
    cfg = Oracle.config("udpt")
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    for leaf_split, accel in ((0, 0), (0, 1), (0, 2)):
        tri, light, t, _ = _trace(hc, od, None, 0, tris, nodes, leaf_split, accel)
        assert (tri == otri).all() and (light == olight).all() and (t.view(np.uint32) == ot.view(np.uint32)).all(), (leaf_split, accel)


def test_reinsertion_passes_make_the_own_tree_cheaper_and_keep_every_hit(hc, monkeypatch):
    """relayout.cpp: OptTree -- reinsertion passes over the host-built own tree (default up to 2^18 triangles).  The own tree's
    topology is free (the exact leaf-box filter decides the hits), so the passes may only change the WORK: on the teapot scene the
    box tests per ray drop by more than 5 % and every hit record stays bit-identical."""
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(19); n = 60000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    res = {}
    for passes in (0, 2):
        monkeypatch.setenv("YUNE_OWN_OPT", str(passes))
        res[passes] = _trace(hc, od, None, 0, tris, nodes, 0, 1)
    monkeypatch.delenv("YUNE_OWN_OPT")
    dflt = _trace(hc, od, None, 0, tris, nodes, 0, 1)
    for a in (res[2], dflt):
        assert (a[0] == res[0][0]).all() and (a[1] == res[0][1]).all() and (a[2].view(np.uint32) == res[0][2].view(np.uint32)).all()
    assert res[2][3][0] < 0.95 * res[0][3][0], (res[0][3], res[2][3])
    assert dflt[3][0] == res[2][3][0]                      # two passes are the default at this size
