"""The oracle must be right before anything is compared with it.  The reference has no tests of its own,
so the restatement (oracle/yune_oracle.cpp) is pinned against the reference ITSELF:
  * committed fixtures rendered by the reference's own kernel text (tests/golden/hdr_*.npz, primary_*.npz, kat.npz);
  * live, when oracle/_ref is present: bit-identity frame by frame for every kernel variant."""
import os

import numpy as np
import pytest

from tests.helpers import GOLDEN, load_golden_scene, transmissive
from tests.refbind import Oracle, default_cam_array, frame_rands

CAM = default_cam_array()


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("cfg,scene,variant,tr", [
    ("c1_udpt_128", "cornellbox", "udpt", False), ("c1_udptmis_128", "cornellbox", "udpt_mis", False),
    ("c2_udptmis_96", "teapot", "udpt_mis", True), ("c3_bdpt_64", "teapot", "bdpt", False)])
def test_restatement_reproduces_reference_frames_bit_exactly(oracle, cfg, scene, variant, tr):
    g = np.load(os.path.join(GOLDEN, "hdr_%s.npz" % cfg))
    img, spp, seed = g["image"], int(g["spp"]), int(g["seed"])
    tris, mats, nodes = load_golden_scene(scene)
    if tr:
        tris = transmissive(tris)
    W = img.shape[1]
    # every frame of the fixture through the same progressive protocol (a 64-frame fixture costs ~2 s on 8 cores)
    full = oracle.render(Oracle.config(variant), CAM, tris, mats, nodes, W, W, frame_rands(seed, spp))
    assert (full[..., 3] == spp).all()
    assert (_bits(full) == _bits(img)).all(), "restatement differs from the reference's kernel output"


def test_c1_golden_full_length(oracle):
    """One full-length fixture end to end (64 progressive frames of udpt.cl on the Cornell box)."""
    g = np.load(os.path.join(GOLDEN, "hdr_c1_udpt_128.npz"))
    tris, mats, nodes = load_golden_scene("cornellbox")
    mine = oracle.render(Oracle.config("udpt"), CAM, tris, mats, nodes, 128, 128, frame_rands(int(g["seed"]), int(g["spp"])))
    assert (_bits(mine) == _bits(g["image"])).all()
    assert (mine[..., 3] == 64).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_primary_hits_match_reference(oracle, scene):
    g = np.load(os.path.join(GOLDEN, "primary_%s.npz" % scene))
    tris, mats, nodes = load_golden_scene(scene)
    W = int(g["width"])
    for jm in (0, 1):
        tri, light, t, od, work = oracle.primary(Oracle.config("udpt"), CAM, tris, nodes, int(g["rand"]), jm, W, W)
        assert (tri == g["tri_j%d" % jm]).all() and (light == g["light_j%d" % jm]).all()
        assert (_bits(t) == _bits(g["t_j%d" % jm])).all()
        assert work[2] == 0                          # no BFS-queue overflow on the shipped scenes (HEAP_SIZE 1500)


def test_known_answers(oracle):
    k = np.load(os.path.join(GOLDEN, "kat.npz"))
    assert [oracle.lib.yor_wang_hash(int(s)) & 0xffffffff for s in k["seeds"]] == list(k["wang_hash"])
    assert [oracle.lib.yor_xor_shift(int(s)) & 0xffffffff for s in k["seeds"]] == list(k["xor_shift"])
    out = oracle.tonemap(k["tonemap_in"])
    assert (_bits(out) == _bits(k["tonemap_out"])).all()
    # hand-computed: wang_hash(0) -- (0^61)^(0>>16)=61; *9=549; ^(549>>4)=549^34=519; *0x27d4eb2d; ^>>15
    s = 61 * 9; s ^= s >> 4; s = (s * 0x27d4eb2d) & 0xffffffff; s ^= s >> 15
    assert k["wang_hash"][0] == s
    x = 1; x ^= (x << 13) & 0xffffffff; x ^= x >> 17; x ^= (x << 5) & 0xffffffff
    assert k["xor_shift"][1] == x


@pytest.mark.parametrize("variant", ["udpt", "udpt_mis", "bdpt"])
def test_live_bit_identity_with_compiled_reference(oracle, ref_kernels, variant):
    """Every frame, every variant, both scenes, incl. the transmissive teapot: restatement == reference kernel text."""
    for scene, W, tr in (("cornellbox", 48, False), ("teapot", 40, False), ("teapot", 40, True)):
        tris, mats, nodes = load_golden_scene(scene)
        if tr:
            tris = transmissive(tris)
        rands = frame_rands(4242, 4)
        a = ref_kernels.render(variant, CAM, tris, mats, nodes, W, W, rands)
        b = oracle.render(Oracle.config(variant), CAM, tris, mats, nodes, W, W, rands)
        assert (_bits(a) == _bits(b)).all()
    # direct-light-only mode (GI_CHECK = 0, udpt.cl:458)
    tris, mats, nodes = load_golden_scene("cornellbox")
    a = ref_kernels.render(variant, CAM, tris, mats, nodes, 32, 32, [7, 8], gi=0)
    b = oracle.render(Oracle.config(variant), CAM, tris, mats, nodes, 32, 32, [7, 8], gi=0)
    assert (_bits(a) == _bits(b)).all()


def test_live_trace_identity(oracle, ref_kernels):
    tris, mats, nodes = load_golden_scene("teapot")
    rng = np.random.RandomState(3); n = 20000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:100, 0] = 0; d[100:200, 1] = 0; d[200:300, 2] = 0
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.05, 2.5, n).astype(np.float32)
    for shadow, t in ((0, None), (1, tm)):
        a = ref_kernels.trace("udpt", od, t, shadow, tris, nodes)
        b = oracle.trace(Oracle.config("udpt"), od, t, shadow, tris, nodes)
        if shadow:
            assert (((a[0] >= 0) | (a[1] >= 0)) == ((b[0] >= 0) | (b[1] >= 0))).all()
        else:
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all()


def test_counter_stream_is_order_free(oracle):
    """rng_mode 1: the value of a sample depends only on (seed, pixel, sample index), not on what was rendered before."""
    tris, mats, nodes = load_golden_scene("cornellbox")
    cfg = Oracle.config("udpt_mis", rng_mode=1, seed=77)
    a = oracle.samples(cfg, CAM, tris, mats, nodes, 24, 24, 5)
    oracle.samples(cfg, CAM, tris, mats, nodes, 24, 24, 3)
    b = oracle.samples(cfg, CAM, tris, mats, nodes, 24, 24, 5)
    assert (_bits(a) == _bits(b)).all()
    c = oracle.samples(Oracle.config("udpt_mis", rng_mode=1, seed=78), CAM, tris, mats, nodes, 24, 24, 5)
    assert not (_bits(a) == _bits(c)).all()
