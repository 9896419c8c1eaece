"""Build the C restatement (oracle/sdp_oracle.c) into oracle/_build/liboracle.so.

TEST INFRASTRUCTURE ONLY.  Flags: -O2 -ffp-contract=off (no FMA contraction,
like the reference's object), no -ffast-math, no -march.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sdp_oracle.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")


def build(force=False, verbose=False):
    if (not force and os.path.exists(LIB)
            and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-strict-overflow", "-fPIC", "-shared",
           "-std=c99", "-Wall", SRC, "-o", LIB, "-lm"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 or verbose:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("gcc failed building the oracle")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
