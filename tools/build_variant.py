"""Development aid: compile the library with extra -D flags into gpurun-visible build/variants/<name>.so
(select it with YUNE_B200_LIB=<path>).   python tools/build_variant.py mb2 -DYUNE_SHADE_MIN_BLOCKS=2"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from yune_b200 import build as B
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "build", "variants"); os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, name + ".so")
srcs = [os.path.join(B.CSRC, s) for s in B.CUDA_SOURCES + B.HOST_SOURCES]
cmd = ["nvcc"] + B.NVCC_FLAGS + flags + ["-shared", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(B.CSRC, "cuda"),
                                          "-I", os.path.join(B.CSRC, "host"), "-o", out] + srcs
r = subprocess.run(cmd, capture_output=True, text=True)
open(out + ".ptxas.log", "w").write(r.stdout + r.stderr)
if r.returncode: sys.exit(r.stderr[-3000:])
for l in (r.stdout + r.stderr).splitlines():
    if "k_shade_denseILb1" in l or "k_traceILb0ELi1" in l: show = 2
    elif "show" in dir() and show > 0:
        print(l.strip()); show -= 1
print(out)
