"""First GPU bring-up check (development aid, not a test): parity of the trace kernels vs the reference's traceRay,
a statistical image check and a first timing.  Run under gpurun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import yune_b200 as yb
from tests.refbind import RefKernels, load_golden_scene, default_cam_array, have_ref

out = {}
def scene_from_golden(name, transmissive=False):
    s = yb.Scene.__new__(yb.Scene)
    tris, mats, nodes = load_golden_scene(name)
    if transmissive:
        tris = tris.copy(); tris["matID"] = np.where(tris["matID"] == 3, 4, tris["matID"])
    s.vert_data, s.mat_data, s.bvh = tris, mats, nodes
    s.main_camera = yb.default_camera(); s._h = None
    return s

kern = RefKernels() if have_ref() else None
cam = default_cam_array()
m = yb.CUDAManager().setup(0)
m.setOption("max_iterations", 20000)
for name, W in (("cornellbox", 512), ("teapot", 1024)):
    sc = scene_from_golden(name)
    r = yb.RendererCore(m, W, W)
    assert m.createRenderProgram("udpt.cl"), m.last_message
    assert r.setup(sc), m.last_message
    for jm in (0, 1):
        t0 = time.time(); tri, light, t = r.tracePrimary(jm, 12345); dt = time.time() - t0
        if kern:
            rtri, rlight, rt, od = kern.primary("udpt", cam, sc.vert_data, sc.bvh, 12345, jm, W, W)
            out["primary_%s_j%d" % (name, jm)] = dict(tri_mismatch=int((tri != rtri).sum()), light_mismatch=int((light != rlight).sum()),
                                                      t_bits=int((t.view(np.uint32) != rt.view(np.uint32)).sum()), n=int(tri.size), sec=dt)
        print(name, jm, out.get("primary_%s_j%d" % (name, jm)), flush=True)
    rng = np.random.RandomState(7); n = 1 << 20
    o = np.stack([rng.uniform(-1, 1, n), rng.uniform(-1, 0.98, n), rng.uniform(-4, -2, n)], 1)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:1000, 0] = 0; d[1000:2000, 1] = 0; d[2000:2500, 2] = 0
    od = np.concatenate([o, d], 1).astype(np.float32)
    tm = rng.uniform(0.05, 2.5, n).astype(np.float32)
    tri, light, t = r.traceRays(od)
    atri, alight, _ = r.traceRays(od, tm, any_hit=True)
    if kern:
        rtri, rlight, rt = kern.trace("udpt", od, None, 0, sc.vert_data, sc.bvh)
        stri, slight, _ = kern.trace("udpt", od, tm, 1, sc.vert_data, sc.bvh)
        out["random_%s" % name] = dict(tri_mismatch=int((tri != rtri).sum()), light_mismatch=int((light != rlight).sum()),
                                       t_bits=int((t.view(np.uint32) != rt.view(np.uint32)).sum()),
                                       any_mismatch=int((((stri >= 0) | (slight >= 0)) != (atri >= 0)).sum()), n=n)
    print(name, out.get("random_%s" % name), flush=True)

# image sanity: C1 128^2 64 spp vs the reference's golden image
for variant, opts, gold in (("udpt", "", "hdr_c1_udpt_128.npz"), ("udpt_mis", "-DMIS", "hdr_c1_udptmis_128.npz")):
    sc = scene_from_golden("cornellbox")
    r = yb.RendererCore(m, 128, 128)
    assert m.createRenderProgram("udpt.cl", compiler_opts=opts), m.last_message
    assert r.setup(sc), m.last_message
    st = r.enqueueKernels(64)
    img = r.readHDR()
    g = np.load(os.path.join(ROOT, "tests", "golden", gold))["image"]
    fin = np.isfinite(img[..., :3]).all(-1) & np.isfinite(g[..., :3]).all(-1)
    out["image_c1_" + variant] = dict(mean_ours=float(img[..., :3][fin].mean()), mean_ref=float(g[..., :3][fin].mean()),
                                      nonfinite_ours=int((~np.isfinite(img[..., :3]).all(-1)).sum()), nonfinite_ref=int((~np.isfinite(g[..., :3]).all(-1)).sum()),
                                      alpha_min=float(img[..., 3].min()), alpha_max=float(img[..., 3].max()), ms=st.render_ms, iters=st.iterations,
                                      ext=st.extend_rays, shad=st.shadow_rays)
    print(variant, out["image_c1_" + variant], flush=True)
    np.save(os.path.join(ROOT, "gpurun_out", "c1_%s.npy" % variant), img)

# first timing: C2 at 1024^2, 16 spp
sc = scene_from_golden("teapot", transmissive=True)
r = yb.RendererCore(m, 1024, 1024)
assert m.createRenderProgram("udpt.cl", compiler_opts="-DMIS")
assert r.setup(sc)
for pool in (1 << 20, 1 << 21, 1 << 22):
    m.setOption("pool_slots", pool)
    r.enqueueKernels(4, reset=True)
    st = r.enqueueKernels(16, reset=True)
    rays = st.extend_rays + st.shadow_rays
    out["c2_pool_%d" % pool] = dict(ms=st.render_ms, msamples_s=st.samples / st.render_ms / 1e3, mrays_s=rays / st.render_ms / 1e3,
                                    iters=st.iterations, ext=st.extend_rays, shad=st.shadow_rays, launches=st.kernel_launches)
    print(pool, out["c2_pool_%d" % pool], flush=True)
img = r.readHDR(); np.save(os.path.join(ROOT, "gpurun_out", "c2.npy"), img[::4, ::4].copy())
m.setOption("time_stages", 1); m.setOption("pool_slots", 1 << 20)
st = r.enqueueKernels(8, reset=True)
out["c2_stages"] = dict(ms=st.render_ms, shade_ms=st.shade_ms, trace_ms=st.trace_ms, iters=st.iterations)
print(out["c2_stages"])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "first.json"), "w"), indent=1)
