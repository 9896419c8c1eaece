"""Interpolation front-ends over the K2 kernel (`sdp_interp`).

* `MlinInterpolator` - the callable `DPSolver.interp_on_state` returns
  (reference stodynprog/stodynprog.py:255-290): same constructor, attributes
  (`ndim`, `_xmin`, `_xmax`, `_xshape`, `values`), broadcasting call semantics,
  and it stays picklable (only numpy attributes), like the reference's
  `P_sto_law.dat`.
* `multilinear_interpolation`, `MultilinearInterpolator`, `mlinspace` - the dolo
  API vendored by the reference (dolointerpolation/multilinear_cython.pyx:17-49,
  dolointerpolation/multilinear.py:15-91).

Batches are evaluated on the GPU (K2).  Calls with at most `HOST_MAX_POINTS` points - the
reference's simulation loops evaluate the policy interpolant one scalar point per time step
(examples/20 Searev storage control/storage_control.py:217,246) - go to the library's host
routine `sdp_interp_host` instead: same operations in the same order, bit-identical results,
no launch and no PCIe crossing.  It is a latency path of the same native library, not a
fallback: without the CUDA extension or without a GPU the package raises.
"""
import ctypes
import os

import numpy as np

from . import _cabi

# a GPU call costs ~100 us of copies, launch and synchronisation: what the host routine needs
# for a few thousand points
HOST_MAX_POINTS = int(os.environ.get("SDP_INTERP_HOST_MAX_POINTS", "2048"))

__all__ = ["MlinInterpolator", "multilinear_interpolation", "MultilinearInterpolator", "mlinspace"]

_engine = None


def _default_engine():
    global _engine
    if _engine is None:
        from .engine import Engine
        _engine = Engine()
    return _engine


def multilinear_interpolation(smin, smax, orders, values, s, engine=None):
    """Multilinear interpolation of `values` (n_v, prod(orders)), C-order rows, at
    the points `s` (d, n_s) on the even grid [smin, smax] with `orders` points
    per axis.  Returns a fresh (n_v, n_s) array of the dtype of `values` (fp64 or
    fp32).  Points outside the grid are linearly extrapolated from the border
    cell, exactly like the reference (multilinear_cython.pyx:17-49).
    """
    values = np.asarray(values)
    s = np.asarray(s)
    if values.ndim != 2 or s.ndim != 2:
        raise ValueError("values and s must be 2-D arrays (n_v, n_grid) and (d, n_s)")
    d = s.shape[0]
    if not 1 <= d <= 4:
        # same type and text as the reference's dispatcher (pyx:46-47)
        raise Exception("Can't interpolate in dimension strictly greater than 5")
    dtype = np.float32 if values.dtype == np.float32 else np.float64
    if values.dtype != dtype or s.dtype != dtype:
        # the reference's typed memoryviews reject mixed / other dtypes
        raise ValueError("Buffer dtype mismatch: values and s must both be float64 (or both float32)")
    orders = [int(o) for o in np.asarray(orders).ravel()]
    if len(orders) != d or len(smin) != d or len(smax) != d:
        raise ValueError("smin, smax and orders must have one entry per row of s")
    if values.shape[1] != int(np.prod(orders)):
        raise ValueError("values has %d columns, the grid has %d points"
                         % (values.shape[1], int(np.prod(orders))))
    g = _cabi.SdpGrid()
    g.d = d
    for k in range(d):
        g.order[k] = orders[k]
        g.smin[k] = float(smin[k])
        g.smax[k] = float(smax[k])
    eng = engine or _default_engine()          # (raises without the extension or without a GPU)
    values = np.ascontiguousarray(values)
    s = np.ascontiguousarray(s)
    n_v, n_s = values.shape[0], s.shape[1]
    if n_v * n_s <= HOST_MAX_POINTS and getattr(eng, "_cuda", False):
        out = np.empty((n_v, n_s), dtype=dtype)
        fn = eng.lib.sdp_interp_host_f32 if dtype == np.float32 else eng.lib.sdp_interp_host
        rc = fn(ctypes.byref(g), n_v, values.ctypes.data_as(ctypes.c_void_p), n_s,
                s.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        _cabi.check(rc, "sdp_interp_host")
        return out
    return eng.interp(g, values, s)


class MlinInterpolator(object):
    """Multilinear interpolant on a rectangular grid given by 1-D axes
    (reference stodynprog.py:255-290)."""

    def __init__(self, *x_grid):
        self.ndim = len(x_grid)
        self._xmin = np.array([x[0] for x in x_grid])
        self._xmax = np.array([x[-1] for x in x_grid])
        self._xshape = np.array([len(x) for x in x_grid], dtype=int)
        self.values = None

    def set_values(self, values):
        assert values.ndim == self.ndim
        assert values.shape == tuple(self._xshape)
        self.values = np.ascontiguousarray(np.atleast_2d(values.ravel()))

    def __call__(self, *x_interp):
        """evaluate at coordinates `x_interp` (one broadcastable array per axis);
        the output has the broadcast shape of the inputs."""
        assert len(x_interp) == self.ndim
        x_mesh = np.broadcast_arrays(*x_interp)
        shape = x_mesh[0].shape
        x_stack = np.vstack([np.asarray(x).astype(float).ravel() for x in x_mesh])
        a = multilinear_interpolation(self._xmin.astype(float), self._xmax.astype(float),
                                      self._xshape, self.values.astype(float, copy=False), x_stack)
        return a.reshape(shape)


def mlinspace(smin, smax, orders):
    """(d, prod(orders)) array enumerating the grid points, last index fastest
    (reference dolointerpolation/multilinear.py:15-21)."""
    if len(orders) == 1:
        res = np.atleast_2d(np.linspace(np.array(smin), np.array(smax), int(np.asarray(orders)[0])))
        return res.reshape(1, -1).copy()
    meshes = np.meshgrid(*[np.linspace(smin[i], smax[i], int(orders[i])) for i in range(len(orders))],
                         indexing='ij')
    return np.vstack([m.flatten() for m in meshes])


class MultilinearInterpolator(object):
    """dolo-style interpolator object: `smin, smax, orders`, `.grid`,
    `.set_values(values)` with one row per interpolated function, call with a
    (d, n_s) array (reference dolointerpolation/multilinear.py:23-91)."""

    __grid__ = None

    def __init__(self, smin, smax, orders, values=None, dtype=np.float64):
        self.smin = np.array(smin, dtype=dtype)
        self.smax = np.array(smax, dtype=dtype)
        self.orders = np.array(orders, dtype=int)
        self.d = len(orders)
        self.dtype = dtype
        if values is not None:
            self.set_values(values)

    @property
    def grid(self):
        if self.__grid__ is None:
            self.__grid__ = mlinspace(self.smin, self.smax, self.orders)
        return self.__grid__

    def set_values(self, values):
        self.values = np.ascontiguousarray(values, dtype=self.dtype)

    def interpolate(self, s):
        s = np.ascontiguousarray(s, dtype=self.dtype)
        return multilinear_interpolation(self.smin, self.smax, self.orders, self.values, s)

    def __call__(self, s):
        return self.interpolate(s)
