"""Build recipe for the native library (yune_b200/libyune_b200.so) and the test-only checkers.

    python -m yune_b200.build            # product library (nvcc, sm_100a only)
    python -m yune_b200.build --oracle   # + oracle/ C++ restatement and, when /root/reference exists, oracle/_ref

The product is ONE shared library exporting the C ABI of include/yune_cuda.h and include/yune_host.h.
It is built in-tree so it travels with the repository snapshot; there is no JIT and no CPU variant.
"""
import os, subprocess, sys, shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "yune_b200", "csrc")
LIB = os.path.join(ROOT, "yune_b200", "libyune_b200.so")

CUDA_SOURCES = ["cuda/kernels.cu", "cuda/bdpt.cu", "cuda/context.cu", "cuda/group.cu", "cuda/bvh_build.cu"]
HOST_SOURCES = ["cuda/relayout.cpp", "host/BVH.cpp", "host/Scene.cpp", "host/Camera.cpp", "host/RendererCore.cpp", "host/ImageIO.cpp", "host/host_capi.cpp"]
APP = os.path.join(ROOT, "yune_b200", "yune_headless")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-fmad=false",                    # parity: products and sums are never fused (see strict_math.h)
              "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_deps():
    deps = []
    for d in ("cuda", "host", "app"):
        for f in os.listdir(os.path.join(CSRC, d)):
            deps.append(os.path.join(CSRC, d, f))
    for f in os.listdir(os.path.join(ROOT, "include")):
        deps.append(os.path.join(ROOT, "include", f))
    return deps


def build_library(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not force and not _newer(LIB, _all_deps()):
        return LIB
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build yune_b200/libyune_b200.so (and there is no CPU fallback)")
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES]
    cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CSRC, "cuda"),
                                 "-I", os.path.join(CSRC, "host"), "-o", LIB] + srcs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed")
    with open(os.path.join(ROOT, "yune_b200", "build_ptxas.log"), "w") as f:
        f.write(r.stdout + r.stderr)
    # headless command-line front end (C++ host, links the library)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CSRC, "host"),
                           os.path.join(CSRC, "app", "yune_headless.cpp"), "-o", APP, "-L", os.path.join(ROOT, "yune_b200"),
                           "-lyune_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB


def build_hostcheck(force=False):
    """tests/hostcheck: TEST-ONLY host compile of the traversal headers (logic check without a GPU)."""
    out = os.path.join(ROOT, "tests", "hostcheck", "libyune_hostcheck.so")
    srcs = [os.path.join(ROOT, "tests", "hostcheck", "hostcheck.cpp"), os.path.join(CSRC, "cuda", "relayout.cpp")]
    if not force and not _newer(out, srcs + _all_deps()):
        return out
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(CSRC, "cuda"), "-o", out] + srcs
    subprocess.check_call(cmd)
    return out


def build_oracle(force=False):
    """oracle/: the C++ restatement (always) and oracle/_ref from the reference sources (when present)."""
    odir = os.path.join(ROOT, "oracle")
    outs = []
    port = os.path.join(odir, "yune_oracle.cpp")
    if os.path.exists(port):
        out = os.path.join(odir, "libyune_oracle.so")
        if force or _newer(out, [port]):
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                                   "-I", os.path.join(ROOT, "include"), "-o", out, port])
        outs.append(out)
    ref = os.environ.get("YUNE_REFERENCE", "/root/reference")
    if os.path.isdir(ref):
        refdir = os.path.join(odir, "_ref")
        os.makedirs(refdir, exist_ok=True)
        host_so = os.path.join(refdir, "libyune_ref_host.so")
        host_srcs = [os.path.join(odir, "ref_host_api.cpp")] + [os.path.join(ref, "src", f) for f in
                                                                ("Scene.cpp", "BVH.cpp", "BVHNodeCPU.cpp", "TriangleCPU.cpp")]
        if force or _newer(host_so, host_srcs):
            subprocess.check_call(["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-w",
                                   "-I", os.path.join(odir, "shim"), "-I", os.path.join(ref, "include"), "-o", host_so] + host_srcs)
        kern_so = os.path.join(refdir, "libyune_ref_kernels.so")
        kdeps = [os.path.join(odir, f) for f in ("gen_ref_kernels.py", "clc_shim.inc", "ref_kernel_driver.inc", "ref_tonemap_driver.inc")]
        if force or _newer(kern_so, kdeps):
            subprocess.check_call([sys.executable, os.path.join(odir, "gen_ref_kernels.py")])
        ocl = os.path.join(refdir, "yune_ref_ocl")      # the reference's udpt.cl behind an OpenCL host (runs it on the box's GPU)
        odeps = [os.path.join(odir, f) for f in ("gen_ref_ocl.py", "ref_ocl_driver.c")]
        if force or _newer(ocl, odeps):
            subprocess.check_call([sys.executable, os.path.join(odir, "gen_ref_ocl.py")])
        outs += [host_so, kern_so, ocl]
    return outs


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_library(force=force, verbose=True))
    if "--oracle" in sys.argv:
        print(build_hostcheck(force=force))
        print(build_oracle(force=force))
