"""Builds tests/emul/libemul_math.so (test infrastructure: the product's HD arithmetic compiled for the host)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def build():
    out = os.path.join(HERE, "libemul_math.so")
    srcs = [os.path.join(HERE, "emul_math.cpp"), os.path.join(HERE, "..", "..", "lgm_b200", "csrc", "splat_math.cuh")]
    if (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([cxx, "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-mfma", "-x", "c++", srcs[0], "-o", out])
    return out


if __name__ == "__main__":
    print(build())
