"""GPU parity tests (run on a B200, through the C ABI).  Bars:
  * integer / index work (hit triangle ids, light ids, occlusion flags): BIT-EXACT against the oracle;
  * hit distances t: bit-exact (same IEEE +,-,*,/ chain, no FMA);
  * radiance: the product uses CUDA's sinf/cosf/powf where the oracle uses glibc's, so per-sample values agree to a few
    ulp unless a comparison flips; tolerance: >= 99.5 % of pixels within 1e-3 relative of the oracle's value for the same
    (seed, pixel, sample), and mean luminance within 0.2 %;
  * against the reference's own RNG stream (different random numbers): statistical agreement, mean luminance within 1 %
    (64 spp fixtures) and relative RMSE within the noise expectation (SURVEY.md 8c pin 3)."""
import os

import numpy as np
import pytest

import yune_b200 as yb
from tests.helpers import GOLDEN, golden_scene_object, load_golden_scene, luminance, rel_rmse, transmissive
from tests.refbind import Oracle, default_cam_array, frame_rands

pytestmark = pytest.mark.gpu
CAM = default_cam_array()


def _renderer(m, scene, W, H, kernel="udpt.cl", opts="", transmissive_teapot=False):
    sc = golden_scene_object(scene, transmissive_teapot)
    r = yb.RendererCore(m, W, H)
    assert m.createRenderProgram(kernel, compiler_opts=opts), m.last_message
    assert r.setup(sc), m.last_message
    return r, sc


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("scene,W", [("cornellbox", 512), ("teapot", 1024)])
def test_primary_hit_ids_bit_exact_at_config_resolution(gpu_manager, oracle, scene, W):
    """SURVEY.md 8c pin 2 at the C1 / C2 resolutions: pixel centres and the reference's jitter for rand = 12345."""
    r, sc = _renderer(gpu_manager, scene, W, W)
    for jm in (0, 1):
        tri, light, t = r.tracePrimary(jm, 12345)
        otri, olight, ot, od, work = oracle.primary(Oracle.config("udpt"), CAM, sc.vert_data, sc.bvh, 12345, jm, W, W)
        assert (tri == otri).all(), "%d primary hit ids differ" % int((tri != otri).sum())
        assert (light == olight).all()
        assert (_bits(t) == _bits(ot)).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_primary_hits_against_committed_golden(gpu_manager, scene):
    g = np.load(os.path.join(GOLDEN, "primary_%s.npz" % scene))
    W = int(g["width"])
    r, sc = _renderer(gpu_manager, scene, W, W)
    for jm in (0, 1):
        tri, light, t = r.tracePrimary(jm, int(g["rand"]))
        assert (tri == g["tri_j%d" % jm]).all() and (light == g["light_j%d" % jm]).all()
        assert (_bits(t) == _bits(g["t_j%d" % jm])).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_random_rays_bit_exact(gpu_manager, oracle, scene):
    """Bounce-like and shadow-like rays, incl. axis-degenerate directions (NaN-guarded slabs) and short segments."""
    r, sc = _renderer(gpu_manager, scene, 64, 64)
    rng = np.random.RandomState(5); n = 300000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d[:2000, 0] = 0; d[2000:4000, 1] = 0; d[4000:6000, 2] = 0; d[6000:6500, :2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
    cfg = Oracle.config("udpt")
    tri, light, t = r.traceRays(od)
    otri, olight, ot = oracle.trace(cfg, od, None, 0, sc.vert_data, sc.bvh)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    atri, alight, _ = r.traceRays(od, tm, any_hit=True)
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, sc.vert_data, sc.bvh)
    assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()
    # empty input is a no-op
    e = r.traceRays(np.zeros((0, 6), np.float32))
    assert e[0].size == 0


SAMPLE_CASES = [("cornellbox", "udpt", "", False, 96), ("cornellbox", "udpt_mis", "-DMIS", False, 96),
                ("teapot", "udpt", "", False, 96), ("teapot", "udpt_mis", "-DMIS", True, 96), ("teapot", "udpt_mis", "-DMIS", False, 64)]


@pytest.mark.parametrize("scene,variant,opts,tr,W", SAMPLE_CASES)
def test_per_sample_radiance_matches_oracle(gpu_manager, oracle, scene, variant, opts, tr, W):
    """Same counter-based stream on both sides -> the SAME sample, pixel by pixel (1 spp images, three sample indices)."""
    r, sc = _renderer(gpu_manager, scene, W, W, opts=opts, transmissive_teapot=tr)
    r.seed = 2024
    cfg = Oracle.config(variant, rng_mode=1, seed=2024)
    close_frac, lum_ours, lum_ref = [], 0.0, 0.0
    for s in (0, 1, 7):
        gpu_manager.check(r._lib.yune_render(r._ctx, s, 1, 1, r.seed, 1))
        ours = r.readSum()
        ref = oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, s)
        assert (ours[..., 3] == 1).all()
        a, b = ours[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
        close = (np.abs(a - b) <= 1e-3 * np.abs(b) + 1e-6).all(-1)
        close_frac.append(close.mean())
        lum_ours += luminance(a).mean(); lum_ref += luminance(b).mean()
    assert min(close_frac) >= 0.995, "per-sample agreement %s" % close_frac
    assert abs(lum_ours - lum_ref) / lum_ref < 2e-3


def test_extensions_two_lights_and_oren_nayar_match_oracle(gpu_manager, oracle):
    """Config C3's ingredients on the unidirectional integrator: a second quad light (sampleLights' N-light branch,
    udpt.cl:667-697, with the light-pick draw) and Oren-Nayar on pure-diffuse lobes (udpt-primitives.cl:681-725,
    sigma^2 = alpha_x), both sides fed the same light list / materials."""
    lights = np.concatenate([yb.LIGHT_UDPT, yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))])
    m = gpu_manager
    for variant, opts in (("udpt", ""), ("udpt_mis", "-DMIS")):
        r, sc = _renderer(m, "teapot", 80, 80, opts=opts)
        mats = sc.mat_data.copy(); mats["alpha_x"] = 0.25            # sigma^2 for the diffuse walls
        assert m.setupMatBuffer(mats) and m.setLightSources(lights)
        m.setOption("oren_nayar", 1)
        try:
            r.seed = 404
            cfg = Oracle.config(variant, rng_mode=1, seed=404, lights=lights, oren_nayar=1)
            fr = []
            for s in (0, 3):
                m.check(r._lib.yune_render(r._ctx, s, 1, 1, r.seed, 1))
                ours = r.readSum()
                ref = oracle.samples(cfg, CAM, sc.vert_data, mats, sc.bvh, 80, 80, s, lights=lights)
                close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
                fr.append(close.mean())
            assert min(fr) >= 0.995, fr
            # the second light is really used: the image differs from the single-light render
            m.setLightSources(None)
            m.check(r._lib.yune_render(r._ctx, 3, 1, 1, r.seed, 1))
            assert np.abs(r.readSum()[..., :3] - ours[..., :3]).mean() > 1e-3
        finally:
            m.setOption("oren_nayar", 0); m.setLightSources(None)


@pytest.mark.parametrize("scene,two_lights,W", [("cornellbox", False, 80), ("teapot", False, 64), ("teapot", True, 64)])
def test_bdpt_per_sample_radiance_matches_oracle(gpu_manager, oracle, scene, two_lights, W):
    """bdpt.cl semantics (createLightPath / createEyePath / all-pairs connections, bdpt.cl:432-640), same counter stream on
    both sides.  Config C3 adds a second quad light: the emitter of the light path is then drawn by |ke| * area.

    Tolerance.  The reference's connection ray ends EXACTLY on the light-path vertex (ray length = distance to the point, no
    epsilon, bdpt.cl:604-610), so the triangle that vertex lies on is hit at t == length up to rounding and the strict
    't < ray->length' (udpt.cl:373) turns each connection into a coin flip decided by the last bit.  Re-compiling the CPU
    oracle itself with FMA contraction changes 13-31 % of the per-sample values of bdpt.cl (0 % for udpt.cl) -- measured,
    see DESIGN.md section 2.  The product's shading arithmetic differs from the oracle's in the last bit (CUDA sinf/cosf/powf,
    reciprocal-multiply normalisation), so per-sample agreement sits in that same band (measured 78-91 %); required: >= 65 % of the
    pixels, and the means must agree to 0.5 %."""
    m = gpu_manager
    lights = None
    if two_lights:
        lights = np.concatenate([yb.LIGHT_BDPT, yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))])
    r, sc = _renderer(m, scene, W, W, kernel="bdpt.cl")
    try:
        if lights is not None:
            assert m.setLightSources(lights)
        r.seed = 555
        cfg = Oracle.config("bdpt", rng_mode=1, seed=555, lights=lights)
        fr, lo, lr = [], 0.0, 0.0
        for s in (0, 2):
            m.check(r._lib.yune_render(r._ctx, s, 1, 1, r.seed, 1))
            ours = r.readSum()
            ref = oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, s, lights=lights)
            assert (ours[..., 3] == 1).all()
            a, b = ours[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
            fr.append((np.abs(a - b) <= 1e-3 * np.abs(b) + 1e-6).all(-1).mean())
            lo += luminance(a).mean(); lr += luminance(b).mean()
        assert min(fr) >= 0.65, fr
        assert abs(lo - lr) / lr < 5e-3
    finally:
        m.setLightSources(None)
        assert m.createRenderProgram("udpt.cl")


def test_bdpt_image_agrees_with_reference_kernel_statistically(gpu_manager):
    g = np.load(os.path.join(GOLDEN, "hdr_c3_bdpt_64.npz"))
    ref, spp = g["image"], int(g["spp"])
    r, sc = _renderer(gpu_manager, "teapot", 64, 64, kernel="bdpt.cl")
    try:
        r.enqueueKernels(spp * 4)                      # 4x the reference's samples: our noise is then the smaller term
        ours = r.readHDR()
        la, lb = luminance(ours).mean(), luminance(ref).mean()
        assert abs(la - lb) / lb < 0.02, (la, lb)
        # noise floor (SURVEY.md 8c pin 3): two independent renders of ours at the REFERENCE's spp differ by sqrt(2) x the
        # single-image noise; ours-vs-reference at equal spp must sit within 1.5 x that (+ a small absolute term)
        r.seed = 101; r.enqueueKernels(spp, reset=True); a = r.readHDR()
        r.seed = 202; r.enqueueKernels(spp, reset=True); b = r.readHDR()
        floor = rel_rmse(a, b)
        assert rel_rmse(a, ref) < 1.5 * floor + 1e-3, (rel_rmse(a, ref), floor)
    finally:
        assert gpu_manager.createRenderProgram("udpt.cl")


def test_synthetic_c4_scene_hits_and_samples(gpu_manager, oracle):
    """configs[3] at subdivision 6 (163,880 triangles): deep tree, leaves refined, nodes beyond the shared-memory prefix.
    Hits bit-exact against the oracle's breadth-first walk (queue unbounded: the reference's 1500-entry cap would silently
    drop subtrees on big scenes), per-sample radiance within tolerance."""
    from yune_b200.scenes import synthetic_c4
    m = gpu_manager
    tris, mats, _ = load_golden_scene("cornellbox")
    sc = yb.Scene().setGeometry(synthetic_c4(tris, 6), mats)
    W = 160
    r = yb.RendererCore(m, W, W)
    assert m.createRenderProgram("udpt.cl") and r.setup(sc), m.last_message
    cfg = Oracle.config("udpt", heap_size=0)
    tri, light, t = r.tracePrimary(1, 999)
    otri, olight, ot, od, work = oracle.primary(cfg, CAM, sc.vert_data, sc.bvh, 999, 1, W, W)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    rng = np.random.RandomState(2); n = 100000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    od6 = np.concatenate([o, d], 1).astype(np.float32)
    a = r.traceRays(od6); b = oracle.trace(cfg, od6, None, 0, sc.vert_data, sc.bvh)
    assert (a[0] == b[0]).all() and (_bits(a[2]) == _bits(b[2])).all()
    r.seed = 3
    m.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1))
    ours = r.readSum()
    ref = oracle.samples(Oracle.config("udpt", rng_mode=1, seed=3, heap_size=0), CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, 0)
    close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
    assert close.mean() >= 0.995, close.mean()


def test_direct_light_only_mode(gpu_manager, oracle):
    """GI_CHECK = 0 (kernel arg 8, udpt.cl:458): direct lighting at the first hit only."""
    r, sc = _renderer(gpu_manager, "cornellbox", 64, 64)
    r.seed = 9
    gpu_manager.check(r._lib.yune_render(r._ctx, 0, 1, 0, r.seed, 1))
    ours = r.readSum()
    ref = oracle.samples(Oracle.config("udpt", rng_mode=1, seed=9), CAM, sc.vert_data, sc.mat_data, sc.bvh, 64, 64, 0, gi=0)
    close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
    assert close.mean() >= 0.995


@pytest.mark.parametrize("cfg,scene,opts,tr", [("c1_udpt_128", "cornellbox", "", False), ("c1_udptmis_128", "cornellbox", "-DMIS", False),
                                               ("c2_udptmis_96", "teapot", "-DMIS", True)])
def test_image_agrees_with_reference_kernel_statistically(gpu_manager, cfg, scene, opts, tr):
    """Against frames rendered by the reference's OWN kernel text with its own RNG (tests/golden/hdr_*.npz)."""
    g = np.load(os.path.join(GOLDEN, "hdr_%s.npz" % cfg))
    ref, spp = g["image"], int(g["spp"])
    W = ref.shape[1]
    r, sc = _renderer(gpu_manager, scene, W, W, opts=opts, transmissive_teapot=tr)
    r.enqueueKernels(spp)
    ours = r.readHDR()
    assert (ours[..., 3] == spp).all()
    assert np.isfinite(ours).all() and np.isfinite(ref).all()
    la, lb = luminance(ours).mean(), luminance(ref).mean()
    assert abs(la - lb) / lb < 0.01, (la, lb)
    # noise floor: two independent renders of ours at the same spp differ by sqrt(2) x the single-image noise
    r.seed = 777
    r.enqueueKernels(spp, reset=True)
    other = r.readHDR()
    floor = rel_rmse(ours, other)
    assert rel_rmse(ours, ref) < 1.5 * floor + 1e-3, (rel_rmse(ours, ref), floor)


def test_sample_ranges_add_up_and_pool_size_is_invisible(gpu_manager):
    """render [0,4) + [4,8) == render [0,8); shards over sample index sum to the unsharded image (the multi-GPU rule,
    SURVEY.md 8e); and the wavefront pool size does not change the image.  fp32 atomics: tolerance = summation order."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 64, 64, opts="-DMIS", transmissive_teapot=True)
    r.seed = 31
    m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); full = r.readSum()
    m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1)); a = r.readSum()
    m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 1)); b = r.readSum()
    m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 0)); ab = r.readSum()      # accumulate on top: a-range then b-range twice
    np.testing.assert_allclose(a + b, full, rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose(b + b, ab, rtol=2e-5, atol=1e-5)
    # default accumulation is fixed point (option "deterministic"): in the integer domain the shards add up EXACTLY
    m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); full_fix = r.readSumFixed()
    m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1)); a_fix = r.readSumFixed()
    m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 1)); b_fix = r.readSumFixed()
    np.testing.assert_array_equal(a_fix + b_fix, full_fix)
    m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 0))
    np.testing.assert_array_equal(r.readSumFixed(), b_fix + b_fix)
    m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1))
    np.testing.assert_array_equal(r.readSum(), full)                             # and a repeated render is bit-identical
    old = m.getOption("pool_slots")
    try:
        m.setOption("pool_slots", 2048)
        m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); small = r.readSum()
    finally:
        m.setOption("pool_slots", old)
    np.testing.assert_array_equal(small, full)                                   # the pool size is invisible, bit for bit
    assert (full[..., 3] == 8).all()
    # zero samples is a no-op that still succeeds
    m.check(r._lib.yune_render(r._ctx, 0, 0, 1, r.seed, 0))


@pytest.mark.parametrize("accel,leaf_split,device_layout", [(0, 0, -1), (0, 2, -1), (1, 0, 0), (1, 0, 1), (2, 0, -1), (2, 3, -1)])
def test_every_acceleration_mode_is_bit_exact(gpu_manager, oracle, accel, leaf_split, device_layout):
    """The walks (reference tree as is / with refined leaves / own tree + exact leaf-box filter, the own tree built by the host's
    binned-SAH builder or by the device's PLOC builder / the same over 4-wide records) against the oracle."""
    m = gpu_manager
    old = (m.getOption("accel"), m.getOption("leaf_split"))
    try:
        m.setOption("accel", accel); m.setOption("leaf_split", leaf_split); m.setOption("device_layout", device_layout)
        r, sc = _renderer(m, "teapot", 256, 256)
        tri, light, t = r.tracePrimary(1, 4711)
        otri, olight, ot, od, _ = oracle.primary(Oracle.config("udpt"), CAM, sc.vert_data, sc.bvh, 4711, 1, 256, 256)
        assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
        rng = np.random.RandomState(17); n = 200000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:1000, 0] = 0; d[1000:2000, 1] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
        od6 = np.concatenate([o, d], 1).astype(np.float32)
        a = r.traceRays(od6); b = oracle.trace(Oracle.config("udpt"), od6, None, 0, sc.vert_data, sc.bvh)
        assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all()
        tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
        sa = r.traceRays(od6, tm, any_hit=True); sb = oracle.trace(Oracle.config("udpt"), od6, tm, 1, sc.vert_data, sc.bvh)
        assert (((sb[0] >= 0) | (sb[1] >= 0)) == (sa[0] >= 0)).all()
        assert m.getOption("layout_built_on_device") == (1 if accel == 1 and device_layout != 0 else 0)
    finally:
        m.setOption("accel", old[0]); m.setOption("leaf_split", old[1]); m.setOption("device_layout", -1)


def test_trace_result_independent_of_warp_scheduling(gpu_manager, oracle):
    """The trace kernel postpones leaves (a lane parks the leaf it reaches and keeps walking) and votes per step between a
    node step and a triangle step; when a triangle is tested relative to the rest of the walk depends on the knobs below and
    on which rays share a warp.  The hit records must not: closest hit with exact ties by reference rank is order-free."""
    m = gpu_manager
    keys = ("refill_idle", "phase_min", "inner_min", "inner_chain", "trace_block", "smem_nodes")
    old = [m.getOption(k) for k in keys]
    try:
        r, sc = _renderer(m, "teapot", 128, 128)
        rng = np.random.RandomState(23); n = 100000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:500, 2] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
        od6 = np.concatenate([o, d], 1).astype(np.float32)
        tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
        b = oracle.trace(Oracle.config("udpt"), od6, None, 0, sc.vert_data, sc.bvh)
        sb = oracle.trace(Oracle.config("udpt"), od6, tm, 1, sc.vert_data, sc.bvh)
        for combo in ((12, 24, 16, 8, 1024, -1), (1, 1, 1, 0, 32, 0), (32, 32, 33, 0, 256, 100), (6, 8, 1, 64, 512, 2340), (20, 2, 30, 3, 1024, 3900)):
            for k, v in zip(keys, combo):
                m.setOption(k, v)
            a = r.traceRays(od6)
            assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all(), combo
            sa = r.traceRays(od6, tm, any_hit=True)
            assert (((sb[0] >= 0) | (sb[1] >= 0)) == (sa[0] >= 0)).all(), combo
    finally:
        for k, v in zip(keys, old):
            m.setOption(k, v)


def test_shade_result_independent_of_scheduling(gpu_manager):
    """The persistent shade kernel sorts slots into per-block lists and runs a round when a list is full; which slots meet
    in a round depends on pool size and grid size, the samples must not.  Repeated renders also exercise the flush pass and
    the list bookkeeping (a missing barrier there once showed up only as a rare illegal address)."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 96, 96, opts="-DMIS", transmissive_teapot=True)
    r.seed = 11
    old_pool = m.getOption("pool_slots")
    ref = None
    try:
        for pool, per_sm in ((0, 0), (1 << 22, 0), (4096, 1), (5000, 3), (1025, 2), (1 << 16, 0), (4096, 1)):
            m.setOption("pool_slots", pool); m.setOption("shade_blocks_per_sm", per_sm)
            m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1))
            a = r.readSum()
            assert (a[..., 3] == 4).all()
            if ref is None: ref = a
            else: np.testing.assert_allclose(a, ref, rtol=2e-5, atol=1e-5)
    finally:
        m.setOption("pool_slots", old_pool); m.setOption("shade_blocks_per_sm", 0)


def test_bdpt_result_independent_of_scheduling(gpu_manager):
    """Same property for the bidirectional kernel: its rounds (light vertex / eye vertex / regenerate / camera) and the
    flattened connection task list regroup with the pool size; every sample must come out the same."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 64, 64, kernel="bdpt.cl", opts="-DMIS")
    r.seed = 23
    old_pool = m.getOption("pool_slots")
    ref = None
    try:
        for pool in (0, 1 << 22, 4096, 1025, 1 << 15, 4096):
            m.setOption("pool_slots", pool)
            m.check(r._lib.yune_render(r._ctx, 0, 3, 1, r.seed, 1))
            a = r.readSum()
            assert (a[..., 3] == 3).all()
            if ref is None: ref = a
            else: np.testing.assert_allclose(a, ref, rtol=2e-5, atol=1e-5)
    finally:
        m.setOption("pool_slots", old_pool)


def test_headless_cli(gpu_manager, tmp_path):
    """C++ front end (csrc/app/yune_headless.cpp over csrc/host/RendererCore.cpp): OBJ in, .pfm out, same image as the
    Python harness renders through the same C ABI."""
    import subprocess
    from tests.helpers import write_obj, ROOT
    tris, mats, nodes = load_golden_scene("cornellbox")
    obj = str(tmp_path / "cb.obj"); out = str(tmp_path / "cb.pfm")
    write_obj(obj, tris, mats)
    exe = os.path.join(ROOT, "yune_b200", "yune_headless")
    p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "8", "--seed", "5", "--out", out], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Total Triangles Loaded: 60" in p.stdout and "BVH Size: 15 Nodes" in p.stdout
    with open(out, "rb") as f:
        assert f.readline() == b"PF\n" and f.readline() == b"64 48\n" and f.readline() == b"-1.0\n"
        img = np.frombuffer(f.read(), "<f4").reshape(48, 64, 3)
    r, sc = _renderer(gpu_manager, "cornellbox", 64, 48)
    r.seed = 5
    r.enqueueKernels(8)
    np.testing.assert_allclose(img, r.readHDR()[..., :3], rtol=1e-4, atol=1e-5)
    bad = subprocess.run([exe, "--obj", str(tmp_path / "missing.obj")], capture_output=True, text=True)
    assert bad.returncode == 1 and "Error opening" in bad.stderr
    # round 2: the viewer's call pattern (frame by frame, pipelined) gives the same image; the BVH can be built on the device
    out2 = str(tmp_path / "cb2.pfm")
    p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "8", "--seed", "5", "--frame-by-frame", "--pipeline", "--out", out2], capture_output=True, text=True)
    assert p.returncode == 0 and "pipelined" in p.stdout, p.stderr
    with open(out2, "rb") as f:
        f.readline(); f.readline(); f.readline()
        img2 = np.frombuffer(f.read(), "<f4").reshape(48, 64, 3)
    np.testing.assert_array_equal(img2, img)
    out3 = str(tmp_path / "cb3.pfm")
    p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "8", "--seed", "5", "--device-bvh", "--option", "sort_rays=1", "--out", out3], capture_output=True, text=True)
    assert p.returncode == 0 and "BVH built on the device" in p.stdout, p.stderr
    with open(out3, "rb") as f:
        f.readline(); f.readline(); f.readline()
        img3 = np.frombuffer(f.read(), "<f4").reshape(48, 64, 3)
    assert np.isfinite(img3).all() and abs(img3.mean() / img.mean() - 1) < 0.05      # another tree: same estimator, grazing hits may differ
    bad = subprocess.run([exe, "--obj", obj, "--option", "nosuch=1"], capture_output=True, text=True)
    assert bad.returncode == 1 and "unknown option" in bad.stderr


def test_tonemap_matches_oracle(gpu_manager, oracle):
    r, sc = _renderer(gpu_manager, "cornellbox", 64, 64)
    r.enqueueKernels(4)
    r.postProcess()
    hdr, ldr = r.readHDR(), r.readLDR()
    ref = oracle.tonemap(hdr)
    np.testing.assert_allclose(ldr, ref, rtol=2e-6, atol=1e-7)
    # Reinhard with L_white = 1 is the identity up to rounding (appendix B#14): ldr = hdr^(1/2.2)
    np.testing.assert_allclose(ldr[..., :3], hdr[..., :3] ** (1 / 2.2), rtol=1e-4, atol=1e-5)


def test_error_behaviour_on_device(gpu_manager):
    m = yb.CUDAManager().setup(0)
    try:
        r = yb.RendererCore(m, 16, 16)
        assert r._lib.yune_render(m._ctx, 0, 1, 1, 0, 1) == -3            # nothing set up
        assert m.createRenderProgram("nosuch.cl") is False and "unknown render kernel" in m.last_message
        assert m.createRenderProgram("udpt.cl", compiler_opts="-DFOO") is False
        assert m.createPostProcProgram("tonemap.cl") is True
        tris, mats, nodes = load_golden_scene("cornellbox")
        bad = nodes.copy(); bad["child_idx"][0] = 10 ** 6
        assert m.setupVertexBuffer(tris) and m.setupMatBuffer(mats) and m.setupBVHBuffer(bad) and m.setupImageBuffers(16, 16)
        m.setupCameraBuffer(yb.default_camera())
        assert r._lib.yune_render(m._ctx, 0, 1, 1, 0, 1) == -4            # malformed BVH rejected at upload, not traversed
        # round-2 entry points: state errors, bad arguments
        m2 = yb.CUDAManager().setup(0)
        try:
            assert m2.buildBVHOnDevice(2) is False and "vertex buffer" in m2.last_message        # nothing to build from
            assert not m2._ok(m2._lib.yune_bvh_info(m2._ctx, None, None, None, None))
            assert m2._ok(m2._lib.yune_finish(m2._ctx))                                           # nothing in flight: a no-op
            assert m2.setupVertexBuffer(tris) and m2.setupMatBuffer(mats)
            m2.setOption("accel", 0)
            assert m2.buildBVHOnDevice(2) is False and "accel 1" in m2.last_message
            m2.setOption("accel", 1)
            assert m2.buildBVHOnDevice(2), m2.last_message
            small = np.zeros(3, nodes.dtype)
            assert not m2._ok(m2._lib.yune_read_bvh_buffer(m2._ctx, small.ctypes.data, 3)) and "capacity" in m2.last_message
            for key, bad_value in (("isect", 2), ("device_builder", 5), ("ploc_radius", 0), ("sort_bits", 31), ("own_tree_passes", 99)):
                assert not m2._ok(m2._lib.yune_set_option(m2._ctx, key.encode(), float(bad_value))), key
            assert not m2._ok(m2._lib.yune_set_option(m2._ctx, b"layout_built_on_device", 1.0)) and "read-only" in m2.last_message
        finally:
            m2.close()
    finally:
        m.close()


@pytest.mark.parametrize("W,H", [(1, 1), (7, 3), (33, 65), (250, 2)])
def test_ragged_image_sizes_match_oracle(gpu_manager, oracle, W, H):
    """Edge cases of the pixel / sample bookkeeping: images that are not square, not a multiple of the 256-slot chunk, smaller
    than a warp.  Primary hits bit-exact, one sample per pixel equal to the oracle's for the same (seed, pixel, sample), every
    pixel counted exactly once per sample, progressive accumulation over three calls."""
    r, sc = _renderer(gpu_manager, "teapot", W, H, opts="-DMIS", transmissive_teapot=True)
    tri, light, t = r.tracePrimary(1, 99)
    otri, olight, ot, _, _ = oracle.primary(Oracle.config("udpt"), CAM, sc.vert_data, sc.bvh, 99, 1, W, H)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    r.seed = 5
    cfg = Oracle.config("udpt_mis", rng_mode=1, seed=5)
    gpu_manager.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1))
    ours = r.readSum()
    assert ours.shape[:2] == (H, W) and (ours[..., 3] == 1).all()
    ref = oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, W, H, 0)
    close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
    assert close.mean() >= (0.99 if W * H >= 100 else 1.0 - 1.0 / (W * H) - 1e-9), close.mean()
    for begin, count in ((1, 2), (3, 1)):
        gpu_manager.check(r._lib.yune_render(r._ctx, begin, count, 1, r.seed, 0))
    assert (r.readSum()[..., 3] == 4).all()


@pytest.mark.parametrize("scene", ["cornellbox", "teapot"])
def test_brute_force_mode_bvh_size_zero(gpu_manager, oracle, scene):
    """Kernel arg 6 bvh_size == 0 (udpt.cl:280-284: every triangle in index order, no box tests): hit records equal to the
    oracle's loop bit for bit, same occlusion answers, one sample per pixel equal to the oracle's for the same (seed, pixel,
    sample).  tests/test_traversal_hostcheck.py pins the same mode on the CPU, also against the compiled reference kernel."""
    m = gpu_manager
    sc = golden_scene_object(scene)
    sc.bvh = sc.bvh[:0]
    W = 48
    r = yb.RendererCore(m, W, W)
    assert m.createRenderProgram("udpt.cl") and r.setup(sc), m.last_message
    cfg = Oracle.config("udpt")
    tri, light, t = r.tracePrimary(1, 77)
    otri, olight, ot, _, _ = oracle.primary(cfg, CAM, sc.vert_data, sc.bvh, 77, 1, W, W)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    rng = np.random.default_rng(78); n = 20000 if scene == "teapot" else 200000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d[:300, 0] = 0; d[300:600, 1] = 0; d[600:900, 2] = 0
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.01, 2.5, n).astype(np.float32)
    a = r.traceRays(od); b = oracle.trace(cfg, od, None, 0, sc.vert_data, sc.bvh)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all()
    sa = r.traceRays(od, tm, any_hit=True); sb = oracle.trace(cfg, od, tm, 1, sc.vert_data, sc.bvh)
    assert (((sb[0] >= 0) | (sb[1] >= 0)) == (sa[0] >= 0)).all()
    r.seed = 9
    m.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1))
    ours = r.readSum()
    ref = oracle.samples(Oracle.config("udpt", rng_mode=1, seed=9), CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, 0)
    close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
    assert close.mean() >= 0.99, close.mean()


def test_headless_cli_writes_png_and_jpg(gpu_manager, tmp_path):
    """saveImage's 8-bit formats through the C++ front end (RendererCore::saveImage -> postProcess -> ImageIO.cpp): the decoded
    .png is the 8-bit view of the tonemapped image the Python harness reads back for the same render, the .jpg is close."""
    import subprocess
    Image = pytest.importorskip("PIL.Image")
    from tests.helpers import write_obj, ROOT
    tris, mats, nodes = load_golden_scene("cornellbox")
    obj = str(tmp_path / "cb.obj")
    write_obj(obj, tris, mats)
    exe = os.path.join(ROOT, "yune_b200", "yune_headless")
    r, sc = _renderer(gpu_manager, "cornellbox", 64, 48)
    assert gpu_manager.createPostProcProgram("tonemap.cl")
    r.seed = 5
    r.enqueueKernels(8)
    r.postProcess()
    want = (np.clip(np.nan_to_num(r.readLDR()[::-1, :, :3]), 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    for ext, tol in ((".png", 1), (".jpg", 6)):      # a second render: fp32 accumulation order may move a value across a rounding step
        out = str(tmp_path / ("cb" + ext))
        p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "8", "--seed", "5", "--out", out], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        got = np.asarray(Image.open(out).convert("RGB"))
        assert got.shape == (48, 64, 3)
        assert np.abs(got.astype(int) - want.astype(int)).max() <= tol, ext
    assert r.saveImage(str(tmp_path / "py.png")) and np.array_equal(np.asarray(Image.open(str(tmp_path / "py.png")).convert("RGB")), want)
    # "Save At Samples" (src/RendererCore.cpp:447-449): the image after 3 of the 8 samples, written to the name as given
    shot = str(tmp_path / "after3")
    p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "8", "--seed", "5", "--save-at", "3", "--save-at-out", shot,
                        "--save-at-ext", ".png"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r.samples_taken = 0
    r.enqueueKernels(3)
    r.postProcess()
    want3 = (np.clip(np.nan_to_num(r.readLDR()[::-1, :, :3]), 0, 1) * np.float32(255) + np.float32(0.5)).astype(np.uint8)
    got3 = np.asarray(Image.open(shot).convert("RGB"))
    assert np.abs(got3.astype(int) - want3.astype(int)).max() <= 1 and np.abs(got3.astype(int) - want.astype(int)).max() > 1


@pytest.mark.parametrize("kind,n,seed", [("uniform", 1500, 31), ("clustered", 1200, 32), ("flats", 1000, 33), ("mixed", 1800, 34)])
def test_random_soups_bit_exact(gpu_manager, oracle, kind, n, seed):
    """Random triangle soups (tests/helpers.random_soup: overlapping boxes, padded flats, slivers over four decades, empty
    children) -- the same scenes and rays tests/test_traversal_hostcheck.py walks on the CPU -- through the device kernels:
    closest hits bit-exact against the oracle's breadth-first walk, same occlusion answers, for the own-tree walk (default)
    and the walk over the reference tree."""
    from tests.helpers import random_soup, soup_rays
    m = gpu_manager
    rng = np.random.default_rng(seed)
    T = random_soup(rng, n, kind)
    sc = yb.Scene().setGeometry(T, load_golden_scene("cornellbox")[1])
    od, tm = soup_rays(rng, T, 200000)
    cfg = Oracle.config("udpt", heap_size=0)
    otri, olight, ot = oracle.trace(cfg, od, None, 0, sc.vert_data, sc.bvh)
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, sc.vert_data, sc.bvh)
    assert (otri >= 0).mean() > 0.3
    try:
        for accel in (1, 0, 2):
            m.setOption("accel", accel)
            r = yb.RendererCore(m, 32, 32)
            assert m.createRenderProgram("udpt.cl") and r.setup(sc), m.last_message
            tri, light, t = r.traceRays(od)
            assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all(), accel
            atri, _, _ = r.traceRays(od, tm, any_hit=True)
            assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all(), accel
    finally:
        m.setOption("accel", 1)


# ------------------------------------------------------------------------------------------------------------------
# round 2: the parity gaps the round-1 review named
# ------------------------------------------------------------------------------------------------------------------
C3_LIGHTS = lambda: np.concatenate([yb.LIGHT_BDPT, yb.quad_light((0.6, 0.0, -3.6), (-1, 0, 0), (8, 8, 8), (0, 0.3, 0), (0, 0, 0.3))])


def test_config_c3_as_configured_matches_oracle(gpu_manager, oracle):
    """configs[2] exactly as BASELINE.json names it: naive BDPT (bdpt.cl:432-640) + TWO quad lights + Oren-Nayar walls
    (sigma^2 = alpha_x = 0.25) + the Phong teapot, one sample per pixel against the oracle with the same counter stream.
    Bar = the BDPT bar (connection rays end exactly on the vertex: >= 65 % of pixels within 1e-3, means within 0.5 %) plus a
    tight statistical pin over 16 samples per pixel: mean luminance within 0.5 % and relRMSE within the noise floor."""
    m = gpu_manager
    lights = C3_LIGHTS()
    W = 64
    r, sc = _renderer(m, "teapot", W, W, kernel="bdpt.cl")
    mats = sc.mat_data.copy(); mats["alpha_x"] = 0.25
    try:
        assert m.setupMatBuffer(mats) and m.setLightSources(lights)
        m.setOption("oren_nayar", 1)
        r.seed = 3003
        cfg = Oracle.config("bdpt", rng_mode=1, seed=3003, lights=lights, oren_nayar=1)
        fr, acc_o, acc_r = [], 0.0, 0.0
        ours16 = np.zeros((W, W, 3)); ref16 = np.zeros((W, W, 3))
        for s_ in range(16):
            m.check(r._lib.yune_render(r._ctx, s_, 1, 1, r.seed, 1))
            ours = r.readSum()
            ref = oracle.samples(cfg, CAM, sc.vert_data, mats, sc.bvh, W, W, s_, lights=lights)
            a, b = ours[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
            fr.append((np.abs(a - b) <= 1e-3 * np.abs(b) + 1e-6).all(-1).mean())
            ours16 += a; ref16 += b
        assert min(fr) >= 0.65, fr
        lo, lr = luminance(ours16).mean(), luminance(ref16).mean()
        assert abs(lo - lr) / lr < 5e-3, (lo, lr)
        # the same 16 samples on both sides: what is left after the coin-flip connections is far below Monte-Carlo noise;
        # bound it by the noise floor of two disjoint 16-sample renders of ours
        other = np.zeros((W, W, 3))
        for s_ in range(16, 32):
            m.check(r._lib.yune_render(r._ctx, s_, 1, 1, r.seed, 1)); other += r.readSum()[..., :3]
        assert rel_rmse(ours16 / 16, ref16 / 16) < 0.5 * rel_rmse(ours16 / 16, other / 16), (rel_rmse(ours16 / 16, ref16 / 16), rel_rmse(ours16 / 16, other / 16))
        # Oren-Nayar and the second light are really in the estimate
        m.setOption("oren_nayar", 0)
        m.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1)); plain = r.readSum()[..., :3]
        m.setOption("oren_nayar", 1); m.setLightSources(None)
        m.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1)); one_light = r.readSum()[..., :3]
        m.setLightSources(lights)
        m.check(r._lib.yune_render(r._ctx, 0, 1, 1, r.seed, 1)); full = r.readSum()[..., :3]
        assert np.abs(full - plain).mean() > 1e-4 and np.abs(full - one_light).mean() > 1e-3
    finally:
        m.setOption("oren_nayar", 0); m.setLightSources(None)
        assert m.createRenderProgram("udpt.cl")


# sequences of (direction key, pitch, yaw) as the GUI would send them (one call per frame; rotation_speed 0.25 rad per unit,
# move_speed 0.1 per call): forward + look up / left, backward + look down / right + up, strafe left + look up / right
MOVED_CAMERAS = [[((0, 0, 1, 0), 0.3, -0.5)] * 2 + [((1, 0, 0, 0), 0.0, 0.2)] * 2,
                 [((0, 0, -1, 0), -0.25, 0.45)] * 3 + [((0, 1, 0, 0), 0.1, 0.0)] * 2,
                 [((-1, 0, 0, 0), 0.4, 0.6)] * 2 + [((0, 0, 1, 0), 0.0, 0.0)] * 5]


@pytest.mark.parametrize("moves", MOVED_CAMERAS)
def test_moved_camera_primary_hits_and_samples(gpu_manager, oracle, moves):
    """Camera::setOrientation (src/Camera.cpp:119-166) -> setBuffer -> createRay (udpt.cl:213-238) with a view matrix that is
    NOT axis aligned: the only case in which matrix rounding reaches the kernel.  The Cam record comes
    from the product's host library (pinned against a restatement of the reference's GLM arithmetic in tests/test_host_scene.py),
    the same 80 bytes go to the oracle: primary hits bit-exact, per-sample radiance within tolerance, and the image really moved."""
    m = gpu_manager
    W, H = 96, 72
    r, sc = _renderer(m, "teapot", W, H, opts="-DMIS", transmissive_teapot=True)
    camera = yb.Camera()
    for direction, pitch, yaw in moves:
        camera.setOrientation(direction, pitch, yaw)
    cam = camera.setBuffer()
    rows = np.stack([cam["r1"][0][:3], cam["r2"][0][:3], cam["r3"][0][:3]])
    assert np.abs(rows).max() < 1.0 - 5e-3                                  # genuinely rotated: no axis-aligned row
    assert np.abs(np.array([cam["r1"][0][3], cam["r2"][0][3], cam["r3"][0][3]])).max() > 0.05     # and translated
    try:
        m.setupCameraBuffer(cam)
        for jm in (0, 1):
            tri, light, t = r.tracePrimary(jm, 2468)
            otri, olight, ot, _, _ = oracle.primary(Oracle.config("udpt"), cam, sc.vert_data, sc.bvh, 2468, jm, W, H)
            assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
        dtri, _, _ = oracle.primary(Oracle.config("udpt"), CAM, sc.vert_data, sc.bvh, 2468, 0, W, H)[:3]
        assert (dtri != otri).mean() > 0.05 and (otri >= 0).mean() > 0.3       # not the default view, and it sees the scene
        r.seed = 77
        cfg = Oracle.config("udpt_mis", rng_mode=1, seed=77)
        for s_ in (0, 5):
            m.check(r._lib.yune_render(r._ctx, s_, 1, 1, r.seed, 1))
            ours = r.readSum()
            ref = oracle.samples(cfg, cam, sc.vert_data, sc.mat_data, sc.bvh, W, H, s_)
            close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
            assert close.mean() >= 0.995, close.mean()
    finally:
        m.setupCameraBuffer(yb.default_camera())


@pytest.mark.parametrize("rr", [0, 2, 12])
def test_rr_threshold_option_matches_oracle(gpu_manager, oracle, rr):
    """RR_THRESHOLD (udpt.cl:6, 514-523) as a run-time option, non-default values on both sides."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 64, 64, opts="-DMIS", transmissive_teapot=True)
    try:
        m.setOption("rr_threshold", rr)
        r.seed = 31
        cfg = Oracle.config("udpt_mis", rng_mode=1, seed=31, rr_threshold=rr)
        dflt = Oracle.config("udpt_mis", rng_mode=1, seed=31)
        for s_ in (0, 1):
            m.check(r._lib.yune_render(r._ctx, s_, 1, 1, r.seed, 1))
            ours = r.readSum()
            ref = oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, 64, 64, s_)
            close = (np.abs(ours[..., :3] - ref[..., :3]) <= 1e-3 * np.abs(ref[..., :3]) + 1e-6).all(-1)
            assert close.mean() >= 0.995, close.mean()
        other = oracle.samples(dflt, CAM, sc.vert_data, sc.mat_data, sc.bvh, 64, 64, 1)
        assert np.abs(other[..., :3] - ref[..., :3]).mean() > 1e-4              # the option changes the estimate
    finally:
        m.setOption("rr_threshold", -1)


@pytest.mark.parametrize("bounces,rr", [(3, -1), (6, 1), (32, 8)])
def test_bdpt_bounces_and_rr_options_match_oracle(gpu_manager, oracle, bounces, rr):
    """BDPT_BOUNCES (bdpt.cl:7: path-vertex budget of both sub-paths) and bdpt.cl's RR_THRESHOLD, non-default on both sides."""
    m = gpu_manager
    W = 64
    r, sc = _renderer(m, "teapot", W, W, kernel="bdpt.cl")
    try:
        m.setOption("bdpt_bounces", bounces); m.setOption("rr_threshold", rr)
        r.seed = 909
        cfg = Oracle.config("bdpt", rng_mode=1, seed=909, bdpt_bounces=bounces, rr_threshold=rr)
        fr, lo, lr = [], 0.0, 0.0
        for s_ in (0, 1, 2):
            m.check(r._lib.yune_render(r._ctx, s_, 1, 1, r.seed, 1))
            ours = r.readSum()
            ref = oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, s_)
            a, b = ours[..., :3].astype(np.float64), ref[..., :3].astype(np.float64)
            fr.append((np.abs(a - b) <= 1e-3 * np.abs(b) + 1e-6).all(-1).mean())
            lo += luminance(a).mean(); lr += luminance(b).mean()
        assert min(fr) >= 0.65, fr
        assert abs(lo - lr) / lr < 1e-2, (lo, lr)
        if bounces == 3:        # a three-vertex budget is a visibly different (darker) estimator than the default 20
            full = oracle.samples(Oracle.config("bdpt", rng_mode=1, seed=909), CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, 0)
            assert luminance(full[..., :3]).mean() > 1.02 * luminance(oracle.samples(cfg, CAM, sc.vert_data, sc.mat_data, sc.bvh, W, W, 0)[..., :3]).mean()
    finally:
        m.setOption("bdpt_bounces", 20); m.setOption("rr_threshold", -1)
        assert m.createRenderProgram("udpt.cl")


def test_synthetic_c4_at_full_size_hits_bit_exact(gpu_manager, oracle):
    """configs[3] at the size that is TIMED: subdivision 9 = 10,485,800 triangles, 3.1 M reference nodes, own tree deeper than
    the staged prefix (two-path trace kernel, > 2340 staged records).  Primary hits and 200,000 random closest / any-hit rays
    bit-exact against the oracle's walk of the reference BVH (queue unbounded, appendix B#13)."""
    from yune_b200.scenes import synthetic_c4
    m = gpu_manager
    tris, mats, _ = load_golden_scene("cornellbox")
    sc = yb.Scene().setGeometry(synthetic_c4(tris, 9), mats)
    assert sc.vert_data.size == 10485800
    W, H = 192, 108
    r = yb.RendererCore(m, W, H)
    assert m.createRenderProgram("udpt.cl") and r.setup(sc), m.last_message
    cfg = Oracle.config("udpt", heap_size=0)
    tri, light, t = r.tracePrimary(1, 4242)
    otri, olight, ot, _, _ = oracle.primary(cfg, CAM, sc.vert_data, sc.bvh, 4242, 1, W, H)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    assert (tri >= 10).mean() > 0.04                          # the spheres (triangles after the ten wall triangles) are in view
    rng = np.random.RandomState(12); n = 200000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d[:500, 0] = 0; d[500:1000, 2] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
    od6 = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
    a = r.traceRays(od6); b = oracle.trace(cfg, od6, None, 0, sc.vert_data, sc.bvh)
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and (_bits(a[2]) == _bits(b[2])).all()
    sa = r.traceRays(od6, tm, any_hit=True); sb = oracle.trace(cfg, od6, tm, 1, sc.vert_data, sc.bvh)
    assert (((sb[0] >= 0) | (sb[1] >= 0)) == (sa[0] >= 0)).all()


def _device_count():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("n_dev", [2, 4])
def test_group_reduced_image_equals_single_gpu(oracle, n_dev):
    """The multi-GPU rule on hardware, through the C ABI (yune_group_*): n devices render disjoint sample ranges, ncclReduce(sum)
    to the root -- the reduced buffer must equal what ONE device renders for the whole range (fp32 summation order only), and
    every pixel must have been counted exactly spp times.  Needs >= n_dev GPUs (gpurun --gpus N); skipped on a single-GPU box."""
    if _device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    sc = golden_scene_object("teapot", transmissive_teapot=True)
    W, H, spp, seed = 160, 96, 37, 4321                       # 37 samples: uneven shards
    one = yb.CUDAGroup(1).setup(sc, W, H, compiler_opts="-DMIS")
    try:
        one.setOption("deterministic", 0)                      # this test pins the FLOAT reduce; the integer one has its own below
        one.render(0, spp, seed=seed); one.reduce(0)
        ref = one.readSum(0)
    finally:
        one.close()
    g = yb.CUDAGroup(n_dev).setup(sc, W, H, compiler_opts="-DMIS")
    try:
        g.setOption("deterministic", 0)
        st = g.render(0, spp, seed=seed)
        assert st.n_devices == n_dev and st.samples == W * H * spp
        parts = [g.readSum(r) for r in range(n_dev)]
        for r in range(n_dev):
            assert (parts[r][..., 3] == yb.shard_samples(0, spp, r, n_dev)[1]).all()
        st = g.reduce(0)
        assert st.reduce_ms > 0
        total = g.readSum(0)
        assert (total[..., 3] == spp).all()
        np.testing.assert_allclose(total, np.sum(parts, axis=0, dtype=np.float32), rtol=1e-6, atol=1e-6)      # NCCL summed exactly these buffers
        np.testing.assert_allclose(total, ref, rtol=2e-5, atol=1e-5)                                         # == the single-device image
        for r in range(1, n_dev):
            np.testing.assert_array_equal(g.readSum(r), parts[r])                                            # non-root buffers untouched
        # a second, accumulating pass continues the sample range on every rank
        g.render(spp, 3, seed=seed, reset=False)
        assert sum(float(g.readSum(r)[0, 0, 3]) for r in range(1, n_dev)) + float(g.readSum(0)[0, 0, 3]) == spp + sum(yb.shard_samples(0, spp, r, n_dev)[1] for r in range(1, n_dev)) + 3
    finally:
        g.close()


def test_group_of_one_and_headless_cli_multi_gpu(tmp_path):
    """A group of one needs no NCCL and equals the plain context; `yune_headless --gpus N` (C++ host over yune_group_*) writes
    the same picture as --gpus 1."""
    import subprocess
    from tests.helpers import write_obj, ROOT
    sc = golden_scene_object("cornellbox")
    g = yb.CUDAGroup(1).setup(sc, 64, 48)
    try:
        g.render(0, 8, seed=5); g.reduce(0)
        img = g.readSum(0)
        assert (img[..., 3] == 8).all() and g.stats().reduce_ms == 0.0
    finally:
        g.close()
    if _device_count() < 2:
        return
    tris, mats, nodes = load_golden_scene("cornellbox")
    obj = str(tmp_path / "cb.obj"); write_obj(obj, tris, mats)
    exe = os.path.join(ROOT, "yune_b200", "yune_headless")
    outs = []
    for n in (1, 2):
        out = str(tmp_path / ("cb%d.pfm" % n))
        p = subprocess.run([exe, "--obj", obj, "--width", "64", "--height", "48", "--spp", "9", "--seed", "5", "--gpus", str(n), "--out", out], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        with open(out, "rb") as f:
            f.readline(); f.readline(); f.readline()
            outs.append(np.frombuffer(f.read(), "<f4").reshape(48, 64, 3))
    np.testing.assert_allclose(outs[0], outs[1], rtol=1e-4, atol=1e-5)


def test_deterministic_accumulation_is_bit_exact(gpu_manager):
    """Option "deterministic": fixed-point (2^-24) integer accumulation.  Same bits from run to run, for any pool size, and for
    any split of the sample range into calls: [0,4) then [4,8) on top == [0,8), and the integer buffers of two separate shards
    ADD UP exactly to the unsharded one (the multi-GPU rule; test_group_reduced_image... does it with real devices).  Against
    the default float accumulation the image agrees to summation-order tolerance."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 96, 64, opts="-DMIS", transmissive_teapot=True)
    r.seed = 31
    assert m.getOption("deterministic") == 1                                    # the default
    m.setOption("deterministic", 0)
    m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); float_mode = r.readSum()
    old_pool = m.getOption("pool_slots")
    try:
        m.setOption("deterministic", 1)
        m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); full, full_fix = r.readSum(), r.readSumFixed()
        assert (full_fix[..., 3] == 8).all() and (full[..., 3] == 8).all()
        np.testing.assert_allclose(full, float_mode, rtol=2e-5, atol=1e-5)
        np.testing.assert_array_equal(full[..., :3], (full_fix[..., :3].astype(np.float32) * np.float32(2.0 ** -24)))     # the float view is derived
        for rep in range(2):                                                     # run to run
            m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1))
            np.testing.assert_array_equal(r.readSumFixed(), full_fix); np.testing.assert_array_equal(r.readSum(), full)
        m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1)); a_fix = r.readSumFixed()
        m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 0))                  # accumulate the second half on top
        np.testing.assert_array_equal(r.readSumFixed(), full_fix); np.testing.assert_array_equal(r.readSum(), full)
        m.check(r._lib.yune_render(r._ctx, 4, 4, 1, r.seed, 1)); b_fix = r.readSumFixed()
        np.testing.assert_array_equal(a_fix + b_fix, full_fix)                   # shards over sample index sum exactly
        for pool in (2048, 5000, 1 << 20):                                       # scheduling
            m.setOption("pool_slots", pool)
            m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1))
            np.testing.assert_array_equal(r.readSumFixed(), full_fix)
        # tonemap reads the derived buffer
        r.postProcess()
        assert np.isfinite(r.readLDR()).all()
        # switching the mode off keeps the image; switching it on again rebuilds the integers from it
        m.setOption("deterministic", 0)
        np.testing.assert_array_equal(r.readSum(), full)
        with pytest.raises(yb.YuneError):
            r.readSumFixed()
    finally:
        m.setOption("deterministic", 1); m.setOption("pool_slots", old_pool)


@pytest.mark.parametrize("n_dev", [2, 4])
def test_group_deterministic_image_is_identical_for_any_gpu_count(n_dev):
    """north_star: "runs are reproducible".  With "deterministic" = 1 the group reduces the INTEGER buffers (ncclInt64 sum): the
    image of n devices equals the image of one device bit for bit."""
    if _device_count() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    sc = golden_scene_object("teapot", transmissive_teapot=True)
    W, H, spp, seed = 128, 80, 21, 99
    imgs = []
    for n in (1, n_dev):
        g = yb.CUDAGroup(n).setup(sc, W, H, compiler_opts="-DMIS")
        try:
            g.render(0, spp, seed=seed); g.reduce(0)            # "deterministic" = 1 is the default
            imgs.append((g.readSum(0), g.readSumFixed(0)))
        finally:
            g.close()
    np.testing.assert_array_equal(imgs[0][1], imgs[1][1])
    np.testing.assert_array_equal(imgs[0][0], imgs[1][0])
    assert (imgs[1][1][..., 3] == spp).all()


def test_watertight_perf_mode_on_device(gpu_manager):
    """Option "isect" = 1 (north star: a watertight ray-triangle test; SURVEY 7: ship a parity mode and a perf mode).
    (a) the device walk returns what the host build of the same headers returns, bit for bit (tests/test_traversal_hostcheck.py
    pins that one: <= 1e-5 of the rays differ from parity mode, no ray leaks through shared edges / vertices of a closed mesh);
    (b) against parity mode ON THE DEVICE: hit ids differ for <= 1e-5 of the rays; (c) no leak through the icosphere's edges and
    vertices; (d) a render in perf mode is the render in parity mode up to the last bits of (t, u, v), pixel by pixel."""
    import ctypes as C
    from tests.helpers import ROOT
    from tests.refbind import ptr
    from tests.test_traversal_hostcheck import _trace, _icosphere, WT
    from tests.refbind import TRI_DTYPE
    hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))
    m = gpu_manager
    assert m.getOption("accel") == 1
    try:
        r, sc = _renderer(m, "teapot", 64, 64)
        rng = np.random.RandomState(41); n = 300000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:1000, 0] = 0; d[1000:2000, 1] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
        od = np.concatenate([o, d], 1).astype(np.float32)
        tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
        p_tri, p_light, p_t = r.traceRays(od)
        m.setOption("isect", 1)
        w_tri, w_light, w_t = r.traceRays(od)
        h_tri, h_light, h_t, _ = _trace(hc, od, None, 0, sc.vert_data, sc.bvh, 0, WT)
        assert (w_tri == h_tri).all() and (w_light == h_light).all() and (_bits(w_t) == _bits(h_t)).all()          # (a)
        assert (w_tri != p_tri).mean() <= 1e-5                                                                      # (b)
        wa = r.traceRays(od, tm, any_hit=True); ha = _trace(hc, od, tm, 1, sc.vert_data, sc.bvh, 0, WT)
        assert ((wa[0] >= 0) == (ha[0] >= 0)).all()
        # (c) closed mesh, rays aimed exactly at vertices and edge points
        V, F = _icosphere(4)
        T = np.zeros(F.shape[0], TRI_DTYPE)
        for k, name in enumerate(("v1", "v2", "v3")):
            T[name][:, :3] = V[F[:, k]]; T[name][:, 3] = 1.0; T["vn" + name[1]][:, :3] = V[F[:, k]]
        ball = yb.Scene().setGeometry(T, load_golden_scene("cornellbox")[1])
        r2 = yb.RendererCore(m, 32, 32)
        assert r2.setup(ball), m.last_message
        e = np.concatenate([F[:, [0, 1]], F[:, [1, 2]], F[:, [2, 0]]])
        targets = np.concatenate([V, 0.5 * (V[e[:, 0]] + V[e[:, 1]])]).astype(np.float32)
        o2 = np.float32([0.1, -0.2, 0.05])
        d2 = targets - o2; d2 = (d2 / np.linalg.norm(d2, axis=1, keepdims=True)).astype(np.float32)
        od2 = np.concatenate([np.broadcast_to(o2, d2.shape), d2], 1).astype(np.float32)
        tri2, light2, _ = r2.traceRays(od2)
        assert (tri2 >= 0).all()
        # (d) images
        r3, _ = _renderer(m, "teapot", 96, 96, opts="-DMIS")           # (the glossy teapot: a glass one amplifies last-bit differences into other paths)
        r3.seed = 5
        m.check(r3._lib.yune_render(r3._ctx, 0, 16, 1, r3.seed, 1)); img_w = r3.readSum()
        m.setOption("isect", 0)
        m.check(r3._lib.yune_render(r3._ctx, 0, 16, 1, r3.seed, 1)); img_p = r3.readSum()
        # (t, u, v) of the two tests differ in the last bits, so the samples do too; a pixel is off by more only where a path
        # took another branch because of them or grazed an edge (measured on B200 with the GLASS teapot, which amplifies them:
        # 95.5 % of the pixels within 1e-3, means 1.3 % apart at 16 spp -- two noise realisations)
        close = (np.abs(img_w - img_p)[..., :3] <= 1e-3 * np.abs(img_p[..., :3]) + 1e-4).all(-1).mean()
        assert close >= 0.9, close
        assert abs(luminance(img_w).mean() / luminance(img_p).mean() - 1) < 1e-2
        # the perf mode needs the own tree
        m.setOption("accel", 0); m.setOption("isect", 1)
        assert not m._ok(r3._lib.yune_render(r3._ctx, 0, 1, 1, 1, 1)) and "accel 1" in m.last_message
    finally:
        m.setOption("accel", 1); m.setOption("isect", 0)


def test_pipelined_frames_add_up_to_one_big_call(gpu_manager):
    """Option "pipeline": the reference's call pattern (ONE sample per pixel per call, src/RendererCore.cpp:483-486) without a full
    drain per call.  A call returns when its samples are handed out and leaves paths in flight; after yune_finish the image is,
    bit for bit (fixed-point accumulation), the image of one call over the whole sample range.  A change of seed finishes the
    carried paths under the old seed first; reset discards them."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 96, 96, opts="-DMIS", transmissive_teapot=True)
    r.seed = 9
    lib, ctx = r._lib, r._ctx
    N = 12
    m.check(lib.yune_render(ctx, 0, N, 1, r.seed, 1)); whole = r.readSumFixed()
    m.check(lib.yune_render(ctx, 0, N, 1, 77, 1)); whole77 = r.readSumFixed()
    try:
        m.setOption("pipeline", 1)
        carried = 0
        for f in range(N):
            m.check(lib.yune_render(ctx, f, 1, 1, r.seed, 1 if f == 0 else 0))
            lib.yune_get_stats(ctx, __import__("ctypes").byref(r.stats)); carried = max(carried, r.stats.carried_paths)
        assert carried > 0                                     # something really was left in flight
        partial = r.readSumFixed()
        assert (partial[..., 3] <= N).all() and (partial[..., 3] < N).any()
        st = r.finish()
        assert st.carried_paths == 0
        np.testing.assert_array_equal(r.readSumFixed(), whole)
        r.finish()                                             # nothing in flight: a no-op
        np.testing.assert_array_equal(r.readSumFixed(), whole)
        # seed change between pipelined calls: the carried paths finish under the seed they were started with
        half = N // 2
        for f in range(half):
            m.check(lib.yune_render(ctx, f, 1, 1, r.seed, 1 if f == 0 else 0))
        for f in range(half, N):
            m.check(lib.yune_render(ctx, f, 1, 1, 77, 0))
        r.finish()
        m.check(lib.yune_render(ctx, 0, half, 1, r.seed, 1)); a = r.readSumFixed()
        m.check(lib.yune_render(ctx, half, N - half, 1, 77, 0)); r.finish()
        mixed = r.readSumFixed()
        m.setOption("pipeline", 0)
        m.check(lib.yune_render(ctx, 0, half, 1, r.seed, 1)); m.check(lib.yune_render(ctx, half, N - half, 1, 77, 0))
        np.testing.assert_array_equal(mixed, r.readSumFixed())
        # reset discards what is in flight; a non-pipelined call after pipelined ones completes everything
        m.setOption("pipeline", 1)
        m.check(lib.yune_render(ctx, 0, 1, 1, r.seed, 1))
        m.check(lib.yune_render(ctx, 0, 1, 1, r.seed, 1))
        m.setOption("pipeline", 0)
        m.check(lib.yune_render(ctx, 1, N - 1, 1, r.seed, 0))
        np.testing.assert_array_equal(r.readSumFixed(), whole)
    finally:
        m.setOption("pipeline", 0)


@pytest.mark.parametrize("builder", [1, 0])
@pytest.mark.parametrize("scene,leaf_max", [("teapot", 2), ("cornellbox", 1), ("uniform", 4), ("flats", 2), ("mixed", 10), ("tiny", 2)])
def test_bvh_built_on_device_keeps_the_node_contract(gpu_manager, oracle, scene, leaf_max, builder):
    """the BVH built ON THE GPU (yune_build_bvh_on_device, bvh_build.cu; builder 1 = PLOC, 0 = linear BVH) is handed
    out in the reference's BVHNodeGPU format and walked by the device from the layout emitted next to it.  Pins: (a) the downloaded
    array is a well-formed reference tree whose boxes nest (the host layout code accepts it for accel 1) and holds every triangle
    exactly once; (b) device hits == the ORACLE's reference-style walk (udpt.cl:288-431) of that downloaded array, bit for bit,
    closest and any-hit; (c) uploading the downloaded array like any host-built one gives the same hits again; (d) a render with it
    agrees with the render over the reference builder's tree (same estimator; hits differ only where a ray grazes a box face)."""
    import ctypes as C
    from tests.helpers import ROOT, random_soup, soup_rays
    from tests.refbind import ptr
    hc = C.CDLL(os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so"))
    m = gpu_manager
    rng = np.random.default_rng(61)
    mats = load_golden_scene("cornellbox")[1]
    if scene in ("teapot", "cornellbox"):
        sc = golden_scene_object(scene)
        n = 200000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:1000, 0] = 0; d[1000:2000, 1] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
        od = np.concatenate([o, d], 1).astype(np.float32); tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
    else:
        T = random_soup(rng, 3 if scene == "tiny" else 3000, "uniform" if scene == "tiny" else scene)
        sc = yb.Scene().setGeometry(T, mats)
        od, tm = soup_rays(rng, T, 60000)
    r = yb.RendererCore(m, 64, 64)
    assert m.createRenderProgram("udpt.cl", compiler_opts=""), m.last_message
    assert r.setup(sc), m.last_message
    host_hits = r.traceRays(od)
    m.setOption("device_builder", builder)
    assert m.buildBVHOnDevice(leaf_max), m.last_message
    m.setOption("device_builder", 1)
    info = m.bvhInfo()
    nodes = m.readBVHBuffer()
    tris = sc.vert_data
    # (a)
    assert nodes.size == info["n_nodes"] and info["device_build_ms"] > 0
    leaf = (nodes["child_idx"] == -1) & (nodes["vert_len"] > 0)
    inner = nodes["child_idx"] > 0
    assert (leaf | inner).all() and (nodes["vert_len"][leaf] <= leaf_max).all()
    listed = np.concatenate([nodes["vert_list"][i, :nodes["vert_len"][i]] for i in np.nonzero(leaf)[0]])
    assert np.array_equal(np.sort(listed), np.arange(tris.size))
    assert hc.hc_layout_accel(ptr(tris), int(tris.size), ptr(nodes), int(nodes.size), 0, 1) == 1
    # (b)
    cfg = Oracle.config("udpt")
    tri, light, t = r.traceRays(od)
    otri, olight, ot = oracle.trace(cfg, od, None, 0, tris, nodes)
    assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
    atri, alight, _ = r.traceRays(od, tm, any_hit=True)
    stri, slight, _ = oracle.trace(cfg, od, tm, 1, tris, nodes)
    assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()
    assert (tri != host_hits[0]).mean() < 1e-3                     # the two trees only disagree where a ray grazes a box face
    # (d) before (c) replaces the device-built layout
    if scene == "teapot":
        r.seed = 3
        m.check(r._lib.yune_render(r._ctx, 0, 16, 1, r.seed, 1)); img_dev = r.readSum()
    # (c)
    assert m.setupBVHBuffer(nodes), m.last_message
    tri2, light2, t2 = r.traceRays(od)
    assert (tri2 == tri).all() and (_bits(t2) == _bits(t)).all()
    if scene == "teapot":
        m.check(r._lib.yune_render(r._ctx, 0, 16, 1, r.seed, 1))
        np.testing.assert_array_equal(r.readSum(), img_dev)       # same tree through both layout paths: the same image, bit for bit
        assert r.setup(sc)
        m.check(r._lib.yune_render(r._ctx, 0, 16, 1, r.seed, 1)); img_host = r.readSum()
        assert abs(luminance(img_dev).mean() / luminance(img_host).mean() - 1) < 2e-3
        assert (img_dev == img_host).all(-1).mean() > 0.99


@pytest.mark.parametrize("builder", [1, 0])
@pytest.mark.parametrize("scene", ["teapot", "cornellbox", "uniform", "flats", "mixed"])
def test_own_tree_built_on_the_device_keeps_the_uploaded_trees_hits(gpu_manager, oracle, scene, builder):
    """Option "device_layout" = 1 (default from 2^20 triangles: C4): the walk's own tree for an UPLOADED BVH is built by the device
    builder instead of the host's binned-SAH builder.  The uploaded tree still decides every hit -- triangle records carry ITS
    leaves and visiting ranks, the filter tests ITS boxes: hit records bit-identical to the oracle's walk of the uploaded tree,
    and a render is bit for bit the render with the host-built own tree."""
    from tests.helpers import random_soup, soup_rays
    m = gpu_manager
    rng = np.random.default_rng(71)
    if scene in ("teapot", "cornellbox"):
        sc = golden_scene_object(scene, transmissive_teapot=(scene == "teapot"))
        n = 200000
        o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
        d = rng.normal(size=(n, 3)); d[:1000, 0] = 0; d[1000:2000, 1] = 0; d /= np.linalg.norm(d, axis=1, keepdims=True)
        od = np.concatenate([o, d], 1).astype(np.float32); tm = rng.uniform(0.001, 2.5, n).astype(np.float32)
    else:
        T = random_soup(rng, 2500, scene)
        sc = yb.Scene().setGeometry(T, load_golden_scene("cornellbox")[1])
        od, tm = soup_rays(rng, T, 60000)
    r = yb.RendererCore(m, 64, 64)
    assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS"), m.last_message
    try:
        assert r.setup(sc), m.last_message
        r.seed = 17
        m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1)); img_host = r.readSum()
        m.setOption("device_builder", builder); m.setOption("device_layout", 1)
        cfg = Oracle.config("udpt")
        tri, light, t = r.traceRays(od)
        otri, olight, ot = oracle.trace(cfg, od, None, 0, sc.vert_data, sc.bvh)
        assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all()
        atri, alight, _ = r.traceRays(od, tm, any_hit=True)
        stri, slight, _ = oracle.trace(cfg, od, tm, 1, sc.vert_data, sc.bvh)
        assert (((stri >= 0) | (slight >= 0)) == (atri >= 0)).all()
        m.check(r._lib.yune_render(r._ctx, 0, 4, 1, r.seed, 1))
        np.testing.assert_array_equal(r.readSum(), img_host)
    finally:
        m.setOption("device_layout", -1); m.setOption("device_builder", 1)


def test_rays_whose_origin_over_direction_overflows_on_device(gpu_manager, oracle):
    """ADVICE r1 on the device: 1/d finite, o * (1/d) infinite (subnormal direction components): lane_init's guard sends the ray
    through the guarded slab form; hit records are the oracle's, with the own tree built by either builder."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 64, 64)
    rng = np.random.RandomState(3); n = 6000
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1) * 8.0 + np.array([0, 0, 21.0])
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    tiny = np.float32([1e-38, 3e-38, -2e-38, 1.2e-38])
    for k in range(3):
        d[k * 2000:(k + 1) * 2000, k] = tiny[rng.randint(0, 4, 2000)]
    od = np.concatenate([o, d], 1).astype(np.float32)
    otri, olight, ot = oracle.trace(Oracle.config("udpt"), od, None, 0, sc.vert_data, sc.bvh)
    try:
        for dl in (0, 1):
            m.setOption("device_layout", dl)
            tri, light, t = r.traceRays(od)
            assert (tri == otri).all() and (light == olight).all() and (_bits(t) == _bits(ot)).all(), dl
    finally:
        m.setOption("device_layout", -1)


def test_sorted_ray_queues_do_not_change_the_image(gpu_manager):
    """Option "sort_rays" (both ray queues radix-sorted by the Morton cell of the ray origin before the trace kernel; off by
    default): the order in which rays are traced is invisible -- with fixed-point
    accumulation the image is bit for bit the unsorted one, for every number of key bits."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 96, 96, opts="-DMIS", transmissive_teapot=True)
    r.seed = 23
    try:
        m.setOption("sort_rays", 0)
        m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1)); plain = r.readSumFixed()
        for bits in (18, 6, 30):
            m.setOption("sort_rays", 1); m.setOption("sort_bits", bits)
            m.check(r._lib.yune_render(r._ctx, 0, 8, 1, r.seed, 1))
            np.testing.assert_array_equal(r.readSumFixed(), plain)
    finally:
        m.setOption("sort_rays", 0); m.setOption("sort_bits", 18)


def test_pipelined_frames_with_the_bidirectional_integrator(gpu_manager):
    """Option "pipeline" with bdpt.cl: a BDPT slot carries its light path and pending connections across calls like a udpt slot
    carries its ray; frames + yune_finish == one call, bit for bit."""
    m = gpu_manager
    r, sc = _renderer(m, "teapot", 64, 64, kernel="bdpt.cl")
    r.seed = 13
    lib, ctx = r._lib, r._ctx
    N = 6
    m.check(lib.yune_render(ctx, 0, N, 1, r.seed, 1)); whole = r.readSumFixed()
    try:
        m.setOption("pipeline", 1)
        for f in range(N):
            m.check(lib.yune_render(ctx, f, 1, 1, r.seed, 1 if f == 0 else 0))
        r.finish()
        np.testing.assert_array_equal(r.readSumFixed(), whole)
    finally:
        m.setOption("pipeline", 0)
        m.createRenderProgram("udpt.cl")
