"""Compile libsvgp_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python -m svgp_vae_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsvgp_b200.so")
SOURCES = ["kernel_matrix.cu", "simt_gemm.cu", "linalg_f64.cu", "rowterms.cu", "tc_engine.cu", "i8_planes.cu", "tc_i8_engine.cu"]
# NOT -arch=sm_100a: that also emits generic compute_100 PTX, on which tcgen05.* does not assemble
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo", "-shared",
              "-Xcompiler", "-fPIC"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "svgp_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    tmp = OUT + ".tmp%d" % os.getpid()             # a failed compile must not take the working library with it
    cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd))
    try:
        subprocess.run(cmd, check=True)
        os.replace(tmp, OUT)
    finally:
        if os.path.exists(tmp):
            os.remove(tmp)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
