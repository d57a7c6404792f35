"""The reference's udpt.cl on the box's GPU through NVIDIA's OpenCL (oracle/_ref/yune_ref_ocl) next to the product, same scene,
same image size (measurement aid; run under gpurun).   python tools/run_ref_ocl.py [c1|c2] [frames]
Prints the driver's line, the product's Msamples/s for the same number of samples, and how the two images compare."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import yune_b200 as yb
from bench import CONFIGS, build_scene, reference_opencl_run, SEED
from tests.helpers import luminance, rel_rmse
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = dict(CONFIGS[name], name=name)
scene = build_scene(cfg)
W, H = cfg["W"], cfg["H"]
ref = reference_opencl_run(cfg, scene, W, H, frames, want_image=True)
img = ref.pop("image", None)
print(json.dumps(ref))
if img is None:
    sys.exit(0)
m = yb.CUDAManager().setup(0)
r = yb.RendererCore(m, W, H)
assert m.createRenderProgram(cfg["kernel"], compiler_opts=cfg["opts"])
sc = yb.Scene(); sc.vert_data, sc.mat_data, sc.bvh = scene[0], scene[1], scene[2]
assert r.setup(sc)
r.enqueueKernels(frames)
st = r.enqueueKernels(frames, reset=True)
ours = r.readHDR()
res = {"ours_msamples_s": st.samples / st.render_ms / 1e3, "ours_ms_per_frame": st.render_ms / frames, "speedup_vs_reference_on_same_gpu": st.samples / st.render_ms / 1e3 / ref["msamples_s"],
       "count_ok": bool((img[..., 3] == frames).all()), "mean_lum_ref": float(luminance(img).mean()), "mean_lum_ours": float(luminance(ours).mean()),
       "mean_lum_ratio": float(luminance(ours).mean() / luminance(img).mean()), "rel_rmse": rel_rmse(ours, img)}
print(json.dumps(res))
