"""Build oracle/_ref/yune_ref_ocl: the reference's udpt.cl, as it lies in the reference tree, embedded in oracle/ref_ocl_driver.c
(an OpenCL host that runs it on the box's GPU through NVIDIA's OpenCL driver).  TEST / MEASUREMENT INFRASTRUCTURE.

Only the '#yune-preproc ...' header lines are removed from the kernel text -- the reference's own loader removes them before
clCreateProgramWithSource (src/CLManager.cpp:182-204).  The text only ever exists as a C string literal in a temporary
directory and inside the (git-ignored) binary."""
import os, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YUNE_REFERENCE", "/root/reference")


def main():
    src = os.path.join(REF, "kernels", "legacy", "udpt.cl")
    text = open(src, "r", errors="replace").read().splitlines(keepends=True)
    while text and text[0].split()[:1] == ["#yune-preproc"]:
        text.pop(0)
    out_dir = os.path.join(ROOT, "oracle", "_ref")
    os.makedirs(out_dir, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        lit = os.path.join(tmp, "kernel_text.c")
        with open(lit, "w") as f:
            f.write("const char* yune_ref_kernel_text =\n")
            for line in text:
                f.write('"' + line.rstrip("\n").replace("\\", "\\\\").replace('"', '\\"') + '\\n"\n')
            f.write(";\n")
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-o", os.path.join(out_dir, "yune_ref_ocl"),
                               os.path.join(ROOT, "oracle", "ref_ocl_driver.c"), lit, "-ldl"])
    print(os.path.join(out_dir, "yune_ref_ocl"))


if __name__ == "__main__":
    sys.exit(main())
