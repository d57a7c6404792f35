#!/usr/bin/env python3
"""oracle/gen_ref_kernels.py -- TEST INFRASTRUCTURE ONLY.

Builds oracle/_ref/libyune_ref_kernels.so from the REFERENCE's own OpenCL kernel sources, read where
they lie under /root/reference (never copied into the repo): kernels/legacy/udpt.cl (with and without
-DMIS), kernels/legacy/bdpt.cl and kernels/post-proc/tonemap.cl.  Each file is adapted mechanically
(the same steps CLManager::createRenderProgram performs, src/CLManager.cpp:182-204, plus C++ spelling of
OpenCL vector literals) and compiled as C++ against oracle/clc_shim.inc.  The adapted text only ever
exists in a temporary build directory.

Text adaptations (regular expressions, no semantic edits):
  * strip CRs and '#yune-preproc' lines;
  * '(floatN)(' / '(int2)(' / '(float4) 2.2f'  ->  'floatN(' ... constructor calls;
  * '.xyz' -> '.xyz()';
  * file-scope '__constant' -> 'static const'; other address-space / access qualifiers are defined away.
"""
import os, re, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("YUNE_REFERENCE", "/root/reference")

def adapt(text):
    text = text.replace("\r", "")
    text = "\n".join(l for l in text.split("\n") if not l.startswith("#yune-preproc"))
    text = re.sub(r"\((float4|float3|int2)\)\s*\(", r"\1(", text)
    text = re.sub(r"\((float4|float3)\)\s*([0-9.]+f?)", r"\1(\2)", text)
    text = re.sub(r"\.xyz\b", ".xyz()", text)
    text = re.sub(r"(?m)^__constant\b", "static const", text)
    return text

PRELUDE = """// GENERATED in a temporary directory by oracle/gen_ref_kernels.py -- do not commit.
#include <cmath>
#include <climits>
#include <cstdint>
#include <cstddef>
#define __kernel
#define __global
#define __constant const
#define constant const
#define __write_only
#define __read_only
%(defines)s
namespace %(ns)s {
#include "clc_shim.inc"
#define YREF_NAME(x) %(ns)s_##x
"""

VARIANTS = [
    # (namespace / symbol prefix, source, extra defines, driver include)
    ("yref_udpt",     "kernels/legacy/udpt.cl",      "",            "ref_kernel_driver.inc"),
    ("yref_udpt_mis", "kernels/legacy/udpt.cl",      "#define MIS", "ref_kernel_driver.inc"),
    ("yref_bdpt",     "kernels/legacy/bdpt.cl",      "",            "ref_kernel_driver.inc"),
    ("yref_tonemap",  "kernels/post-proc/tonemap.cl", "",           "ref_tonemap_driver.inc"),
]

def main():
    if not os.path.isdir(REF):
        print("reference tree not present (%s): keeping any prebuilt oracle/_ref" % REF)
        return 0
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="yref_build_") as tmp:
        objs = []
        for ns, src, defines, driver in VARIANTS:
            with open(os.path.join(REF, src), "r", newline="") as f:
                body = adapt(f.read())
            cpp = os.path.join(tmp, ns + ".cpp")
            with open(cpp, "w") as f:
                f.write(PRELUDE % {"ns": ns, "defines": defines})
                f.write(body)
                f.write('\n#include "%s"\n}\n' % driver)
            obj = os.path.join(tmp, ns + ".o")
            cmd = ["g++", "-std=c++14", "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
                   "-w", "-I", HERE, "-c", cpp, "-o", obj]
            subprocess.check_call(cmd)
            objs.append(obj)
        so = os.path.join(out_dir, "libyune_ref_kernels.so")
        subprocess.check_call(["g++", "-shared", "-fopenmp", "-o", so] + objs)
        print("built", so)
    return 0

if __name__ == "__main__":
    sys.exit(main())
