"""Build recipe for oracle/_ref: the reference's only native component.

TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Compiles the reference's Cython
interpolation routine *from the source where it lies* under /root/reference
(stodynprog/dolointerpolation/multilinear_cython.pyx) into a shared object under
oracle/_ref/.  No reference source is copied into the repository: Cython's
generated C file is written to a temporary directory and deleted, only the
compiled .so stays (git-ignored, but it travels to the GPU box).

Flags follow the reference's own setup.py:18-22 (`-O3`, no OpenMP, no -march,
no -ffast-math) so the object has x86 `cvttsd2si` casts, true divisions and
no FMA contraction - the arithmetic contract of SURVEY.md App. A.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("STODYNPROG_REFERENCE", "/root/reference")
PYX = os.path.join(REF_ROOT, "stodynprog", "dolointerpolation", "multilinear_cython.pyx")
OUT_DIR = os.path.join(HERE, "_ref")
SO_NAME = "multilinear_cython" + sysconfig.get_config_var("EXT_SUFFIX")


def ref_so_path():
    return os.path.join(OUT_DIR, SO_NAME)


def build(force=False, verbose=False):
    """Returns the path of the built .so, or None when the reference is absent
    (GPU box) and no prebuilt file exists."""
    so = ref_so_path()
    if os.path.exists(so) and not force:
        return so
    if not os.path.exists(PYX):
        return None
    import numpy as np
    os.makedirs(OUT_DIR, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="sdp_ref_build_")
    try:
        c_file = os.path.join(tmp, "multilinear_cython.c")
        cmd = [sys.executable, "-m", "cython", "-3", PYX, "-o", c_file]
        subprocess.run(cmd, check=True, capture_output=not verbose)
        inc = sysconfig.get_paths()["include"]
        cmd = ["gcc", "-shared", "-fPIC", "-O3", "-DNDEBUG", "-fno-strict-overflow",
               "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
               "-I", inc, "-I", np.get_include(), c_file, "-o", so]
        subprocess.run(cmd, check=True, capture_output=not verbose)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return so


PY_STAGE = os.path.join(OUT_DIR, "py")      # oracle/_ref/py/stodynprog/*.py (git-ignored)


def stage_reference_python(force=False):
    """Build artefact for the GPU box: the reference's own Python package, copied VERBATIM from
    where it lies (/root/reference/stodynprog/**/*.py) into the git-ignored oracle/_ref/py/, next
    to its compiled routine.  It never enters the repository's history; it travels with the
    snapshot so that `bench.py --impl reference` can time the unmodified
    DPSolver._value_at_state_vect (stodynprog.py:639-691) on the GPU box's host cores, where
    /root/reference does not exist.  Returns the directory to put on sys.path, or None."""
    src = os.path.join(REF_ROOT, "stodynprog")
    dst = os.path.join(PY_STAGE, "stodynprog")
    if os.path.exists(os.path.join(dst, "stodynprog.py")) and not force:
        return PY_STAGE
    if not os.path.exists(os.path.join(src, "stodynprog.py")):
        return None
    for base, dirs, files in os.walk(src):
        rel = os.path.relpath(base, src)
        for f in files:
            if f.endswith(".py"):
                os.makedirs(os.path.join(dst, rel), exist_ok=True)
                shutil.copyfile(os.path.join(base, f), os.path.join(dst, rel, f))
    return PY_STAGE


if __name__ == "__main__":
    print("oracle/_ref/py:", stage_reference_python(force="--force" in sys.argv))
    p = build(force="--force" in sys.argv, verbose=True)
    print("oracle/_ref:", p)
