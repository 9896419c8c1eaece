"""Import the UNMODIFIED reference (pierre-haessig/stodynprog) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to generate the
committed golden fixtures, and by the (container-only) cross-check tests.  The
reference tree is never copied or modified: its Python package is imported from
where it lies (/root/reference) and its Cython routine is the object compiled by
oracle/build_ref.py.  On the GPU box /root/reference does not exist: there the
package is imported from the verbatim, git-ignored copy that `__graft_entry__.build()` stages
under oracle/_ref/py (build_ref.stage_reference_python) - for `bench.py --impl reference` only;
without it `load_reference()` returns None.

Three shims are needed on py3.12 / numpy 2.x; they are
applied to the *environment*, not to the reference's files:
  1. stub `matplotlib` / `matplotlib.pyplot` modules (stodynprog.py:13 imports
     pyplot at module import; matplotlib is not installed),
  2. `inspect.getargspec` rebuilt on `getfullargspec` (stodynprog.py:32-33,120,174),
  3. `np.int = int` (stodynprog.py:265; dolointerpolation/multilinear.py:70).
"""
import collections
import importlib.machinery
import importlib.util
import inspect
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("STODYNPROG_REFERENCE", "/root/reference")
_cache = {}


def _package_root():
    """directory holding the reference's `stodynprog` package: the read-only reference tree in
    the build container, else the verbatim copy staged by build_ref.stage_reference_python()
    (git-ignored oracle/_ref/py, the only form in which it reaches the GPU box)"""
    if os.path.exists(os.path.join(REF_ROOT, "stodynprog", "stodynprog.py")):
        return REF_ROOT
    from . import build_ref
    if os.path.exists(os.path.join(build_ref.PY_STAGE, "stodynprog", "stodynprog.py")):
        return build_ref.PY_STAGE
    return None


def reference_available():
    return _package_root() is not None


def _apply_shims():
    import numpy as np
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
            import matplotlib.pyplot  # noqa: F401
        except Exception:
            mpl = types.ModuleType("matplotlib")
            plt = types.ModuleType("matplotlib.pyplot")
            pylab = types.ModuleType("matplotlib.pylab")
            mpl.pyplot = plt
            mpl.pylab = pylab
            mpl.rcParams = {}
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
            sys.modules["matplotlib.pylab"] = pylab
    if not hasattr(inspect, "getargspec"):
        ArgSpec = collections.namedtuple("ArgSpec", "args varargs keywords defaults")

        def getargspec(func):
            f = inspect.getfullargspec(func)
            return ArgSpec(f.args, f.varargs, f.varkw, f.defaults)
        inspect.getargspec = getargspec
    if not hasattr(np, "int"):
        np.int = int


def load_reference_cython():
    """The reference's compiled interpolation routine (oracle/_ref), or None."""
    if "cy" in _cache:
        return _cache["cy"]
    from . import build_ref
    so = build_ref.build()
    mod = None
    if so is not None:
        name = "stodynprog.dolointerpolation.multilinear_cython"
        loader = importlib.machinery.ExtensionFileLoader(name, so)
        spec = importlib.util.spec_from_file_location(name, so, loader=loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
    _cache["cy"] = mod
    return mod


def load_reference():
    """Returns the reference `stodynprog` package (with SysDescription, DPSolver),
    or None when /root/reference is absent."""
    if "pkg" in _cache:
        return _cache["pkg"]
    if not reference_available():
        _cache["pkg"] = None
        return None
    _apply_shims()
    cy = load_reference_cython()
    # the compiled routine lives outside the (read-only) reference tree:
    sys.modules["stodynprog.dolointerpolation.multilinear_cython"] = cy
    root = _package_root()
    if root not in sys.path:
        sys.path.insert(0, root)
    # the reference package's __init__ does `from stodynprog import tests`
    # which needs nose: give it an inert stub.
    if "nose" not in sys.modules:
        try:
            import nose  # noqa: F401
        except Exception:
            nose = types.ModuleType("nose")
            tools = types.ModuleType("nose.tools")
            import unittest
            tc = unittest.TestCase()
            tools.assert_true = tc.assertTrue
            tools.assert_equal = tc.assertEqual
            tools.assert_raises = tc.assertRaises
            nose.tools = tools
            nose.run = lambda *a, **k: None
            sys.modules["nose"] = nose
            sys.modules["nose.tools"] = tools
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import stodynprog as ref
    _cache["pkg"] = ref
    return ref
