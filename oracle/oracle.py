"""ctypes front-end of the C restatement (oracle/sdp_oracle.c).

TEST INFRASTRUCTURE ONLY - see oracle/README.md.  "Parity pinned": these
functions are checked against the reference's known-answer tests, against the
reference's own compiled routine (oracle/_ref) and against golden fixtures
generated from the unmodified reference (tests/test_oracle.py).
"""
import ctypes

import numpy as np

from . import build as _build

_lib = None
_dp = ctypes.POINTER(ctypes.c_double)
_fp = ctypes.POINTER(ctypes.c_float)
_lp = ctypes.POINTER(ctypes.c_int64)
_ip = ctypes.POINTER(ctypes.c_int32)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(_build.build())
        L.oracle_interp.restype = ctypes.c_int
        L.oracle_interp.argtypes = [ctypes.c_int, _dp, _dp, _lp, ctypes.c_int64, _dp,
                                    ctypes.c_int64, _dp, _dp]
        L.oracle_interp_f32.restype = ctypes.c_int
        L.oracle_interp_f32.argtypes = [ctypes.c_int, _fp, _fp, _lp, ctypes.c_int64, _fp,
                                        ctypes.c_int64, _fp, _fp]
        L.oracle_cell_search.restype = ctypes.c_int
        L.oracle_cell_search.argtypes = [ctypes.c_int, _dp, _dp, _lp, ctypes.c_int64, _dp, _ip, _dp]
        L.oracle_backup.restype = ctypes.c_int
        L.oracle_backup.argtypes = [ctypes.c_int, _dp, _dp, _lp, _dp, ctypes.c_int64, ctypes.c_int64,
                                    ctypes.POINTER(_dp), _lp, _lp, _dp, ctypes.c_int64,
                                    ctypes.c_int64, _dp, _dp, _lp, _dp]
        L.oracle_policy_backup.restype = ctypes.c_int
        L.oracle_policy_backup.argtypes = [ctypes.c_int, _dp, _dp, _lp, _dp, ctypes.c_int64,
                                           ctypes.c_int64, _dp, _dp, ctypes.c_int64, _dp, _dp]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _grid_args(smin, smax, orders):
    smin = np.ascontiguousarray(smin, dtype=np.float64)
    smax = np.ascontiguousarray(smax, dtype=np.float64)
    orders = np.ascontiguousarray(orders, dtype=np.int64)
    return smin, smax, orders


def interp(smin, smax, orders, values, s):
    """same contract as the reference's multilinear_interpolation (pyx:17-49)"""
    values = np.ascontiguousarray(values)
    s = np.ascontiguousarray(s)
    d, n_s = s.shape
    n_v = values.shape[0]
    if not 1 <= d <= 4:
        raise Exception("Can't interpolate in dimension strictly greater than 5")
    orders64 = np.ascontiguousarray(orders, dtype=np.int64)
    if values.dtype == np.float32:
        smin = np.ascontiguousarray(smin, dtype=np.float32)
        smax = np.ascontiguousarray(smax, dtype=np.float32)
        s = s.astype(np.float32, copy=False)
        out = np.zeros((n_v, n_s), dtype=np.float32)
        rc = lib().oracle_interp_f32(d, smin.ctypes.data_as(_fp), smax.ctypes.data_as(_fp),
                                     orders64.ctypes.data_as(_lp), n_v, values.ctypes.data_as(_fp),
                                     n_s, s.ctypes.data_as(_fp), out.ctypes.data_as(_fp))
    else:
        smin, smax, orders64 = _grid_args(smin, smax, orders)
        values = values.astype(np.float64, copy=False)
        s = s.astype(np.float64, copy=False)
        out = np.zeros((n_v, n_s))
        rc = lib().oracle_interp(d, _d(smin), _d(smax), orders64.ctypes.data_as(_lp), n_v, _d(values),
                                 n_s, _d(s), _d(out))
    assert rc == 0
    return out


def cell_search(smin, smax, orders, s):
    """-> (cell int32 [n], lam fp64 [d][n])"""
    smin, smax, orders = _grid_args(smin, smax, orders)
    s = np.ascontiguousarray(s, dtype=np.float64)
    d, n = s.shape
    cell = np.zeros(n, dtype=np.int32)
    lam = np.zeros((d, n))
    rc = lib().oracle_cell_search(d, _d(smin), _d(smax), orders.ctypes.data_as(_lp), n, _d(s),
                                  cell.ctypes.data_as(_ip), _d(lam))
    assert rc == 0
    return cell, lam


def backup(smin, smax, orders, J_next, U, W, coords, g, p, want_all=False):
    """one state's backup from un-broadcast (Ueff, Weff) coordinate / cost arrays.
    Returns (J_opt, ind_opt[, J_all])."""
    smin, smax, orders = _grid_args(smin, smax, orders)
    d = len(orders)
    J_next = np.ascontiguousarray(J_next, dtype=np.float64).ravel()
    arrs, us, ws = [], [], []
    for a in list(coords) + [g]:
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.ndim == 2
        arrs.append(a)
        us.append(a.shape[1] if a.shape[0] > 1 else 0)
        ws.append(1 if a.shape[1] > 1 else 0)
    ptrs = (_dp * d)(*[_d(a) for a in arrs[:d]])
    us_a = np.array(us[:d], dtype=np.int64)
    ws_a = np.array(ws[:d], dtype=np.int64)
    J_opt = ctypes.c_double()
    ind = ctypes.c_int64()
    J_all = np.zeros(U) if want_all else None
    pp = np.ascontiguousarray(p, dtype=np.float64) if p is not None else None
    rc = lib().oracle_backup(d, _d(smin), _d(smax), orders.ctypes.data_as(_lp), _d(J_next), U, W,
                             ptrs, us_a.ctypes.data_as(_lp), ws_a.ctypes.data_as(_lp),
                             _d(arrs[d]), us[d], ws[d], _d(pp) if pp is not None else None,
                             ctypes.byref(J_opt), ctypes.byref(ind),
                             _d(J_all) if want_all else None)
    assert rc == 0
    if want_all:
        return J_opt.value, ind.value, J_all
    return J_opt.value, ind.value


def policy_backup(smin, smax, orders, J_in, s, g, p):
    """fixed-policy backup. s: (d, n, W) dense coordinates; g: (n, W) or (n,)"""
    smin, smax, orders = _grid_args(smin, smax, orders)
    s = np.ascontiguousarray(s, dtype=np.float64)
    d, n, W = s.shape
    g = np.ascontiguousarray(g, dtype=np.float64)
    g_ws = 1 if g.ndim == 2 and g.shape[1] == W and W > 1 else 0
    J_in = np.ascontiguousarray(J_in, dtype=np.float64).ravel()
    p = np.ascontiguousarray(p, dtype=np.float64)
    out = np.zeros(n)
    rc = lib().oracle_policy_backup(d, _d(smin), _d(smax), orders.ctypes.data_as(_lp), _d(J_in), n, W,
                                    _d(s.reshape(d, n * W)), _d(g), g_ws, _d(p), _d(out))
    assert rc == 0
    return out
