"""Build the CPU oracle shared library (TEST INFRASTRUCTURE ONLY).

    python oracle/build.py

Produces oracle/libpcd_oracle.so from oracle/pcd_oracle.c.  -ffp-contract=off keeps
gcc from fusing the explicit mul/add pairs; -mfma makes fmaf() a single vfmadd.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "pcd_oracle.c")
OUT = os.path.join(HERE, "libpcd_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-mfma", "-mavx2",
           "-fopenmp", "-shared", "-fPIC", "-o", OUT, SRC, "-lm"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
