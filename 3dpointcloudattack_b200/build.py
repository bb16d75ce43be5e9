"""Build libpcdist.so (the sm_100a CUDA library behind the C ABI in include/pcdist.h).

    python 3dpointcloudattack_b200/build.py [--force]

nvcc cross-compiles without a GPU.  The .so is written IN-TREE next to this file so that it
travels to the GPU box with the repo snapshot (it is git-ignored, not gpurun-ignored).
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpcdist.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=false",            # never contract: every FMA in the kernels is an explicit intrinsic
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "pcdist.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra_flags=(), out=OUT):
    """extra_flags / out: development variants only (tools/), e.g. -DPCD_SWEEP_TRACE."""
    if out == OUT and not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
