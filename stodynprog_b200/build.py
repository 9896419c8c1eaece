"""Build recipe of the CUDA extension (in-tree shared library, sm_100a only).

`python -m stodynprog_b200.build` compiles stodynprog_b200/csrc/sdp_b200.cu into
stodynprog_b200/_lib/libsdp_b200.so with nvcc.  The library has a plain C ABI
(include/sdp_b200.h) and links the CUDA runtime statically: it has no
dependency on torch or Python.  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
SRC = os.path.join(PKG_DIR, "csrc", "sdp_b200.cu")
INCLUDE = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(PKG_DIR, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsdp_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "--fmad=false",          # never contract a*b+c: parity contract, SURVEY.md App. A
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",   # (host pass: sdp_interp_host)
    "-cudart", "static",
]


def _nvcc():
    cand = os.environ.get("NVCC") or "nvcc"
    if os.path.sep not in cand:
        for d in os.environ.get("PATH", "").split(os.pathsep) + ["/usr/local/cuda/bin"]:
            p = os.path.join(d, cand)
            if os.path.isfile(p) and os.access(p, os.X_OK):
                return p
    return cand


def is_stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [SRC, os.path.join(INCLUDE, "sdp_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the shared library if missing or older than its sources.
    Returns the library path.  `out`: alternative output path (tuning variants,
    always rebuilt)."""
    if out is None and not force and not is_stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    out = out or LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra_flags) + ["-I", INCLUDE, SRC, "-o", out]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libsdp_b200.so (exit %d)" % res.returncode)
    return out


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "--ptxas-v" in sys.argv else []
    print(build(force=True, verbose=True, extra_flags=flags))
