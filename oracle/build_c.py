"""Builds the C helpers of the oracle into oracle/_build/ (git-ignored).  TEST INFRASTRUCTURE ONLY."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")


def build_dct(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, "libdct_ducc.so")
    srcs = [os.path.join(HERE, "dct_ducc_c.c"), os.path.join(HERE, "dct_ducc_generated.h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, srcs[0]], check=True)
    return out


if __name__ == "__main__":
    print(build_dct(True))
