"""numpy model of the C ABI (include/sdp_b200.h) for host-logic tests.

TEST INFRASTRUCTURE ONLY.  It lets the tests drive the product's *host* code
(tabulation, descriptors, work items, slabs, collectives, argmin -> control
values) on a machine without a GPU, by standing in for libsdp_b200.so behind
`Engine(_test_lib=...)`.  It reads and writes the caller's (CPU torch) buffers
through raw pointers with exactly the table layout the header documents, and
does its arithmetic with the C oracle.  It is deliberately slow and simple.
"""
import ctypes

import numpy as np

from oracle import oracle as oc
from stodynprog_b200 import _cabi


def _arr(ptr, n, ctype):
    p = ptr.value if hasattr(ptr, "value") else ptr
    if n == 0 or not p:
        return np.zeros(0, dtype=ctype)
    return np.ctypeslib.as_array((ctype * n).from_address(p))


def _grid(gref):
    g = gref._obj
    d = g.d
    return d, np.array(g.smin[:d]), np.array(g.smax[:d]), np.array(g.order[:d], dtype=np.int64)


def _better(av, ai, bv, bi):
    an, bn = av != av, bv != bv
    if an or bn:
        return (ai < bi) if (an and bn) else an
    if av < bv:
        return True
    if av > bv:
        return False
    return ai < bi


class FakeLib(object):
    def __init__(self):
        self.launches = 0
        self.err = b""

    def sdp_version(self):
        return _cabi.SDP_ABI_VERSION

    def sdp_last_error(self):
        return self.err

    def sdp_launch_count(self):
        return self.launches

    def sdp_cell_setup(self, gref, n, s, cell, lam, stream):
        d, smin, smax, orders = _grid(gref)
        S = _arr(s, d * n, ctypes.c_double).reshape(d, n)
        c, l = oc.cell_search(smin, smax, orders, S)
        _arr(cell, n, ctypes.c_int32)[:] = c
        _arr(lam, d * n, ctypes.c_double)[:] = l.reshape(-1)
        self.launches += 1
        return 0

    def sdp_build_tables(self, gref, W, g_per_w, n_states, desc, staging, cell, lam, lam_plane, g,
                         max_Upad, stream):
        d, smin, smax, orders = _grid(gref)
        D = np.frombuffer((ctypes.c_uint8 * (n_states * _cabi.STATE_DESC_DTYPE.itemsize))
                          .from_address(desc.value), dtype=_cabi.STATE_DESC_DTYPE)
        # staging / table extents are not passed through the ABI: map generously
        top = 0
        for r in D:
            for k in range(d + 1):
                top = max(top, int(r["src"][k]) + (int(r["U"]) - 1) * int(r["us"][k])
                          + (W - 1) * int(r["ws"][k]) + 1)
        stag = _arr(staging, top, ctypes.c_double)
        for r in D:
            U, Upad, eo, go = int(r["U"]), int(r["Upad"]), int(r["entry_off"]), int(r["g_off"])
            assert Upad % 4 == 0 and Upad >= U and Upad <= max_Upad
            u = np.arange(U)[None, :]
            w = np.arange(W)[:, None]
            coords = np.stack([stag[int(r["src"][k]) + u * int(r["us"][k]) + w * int(r["ws"][k])]
                               .reshape(-1) for k in range(d)])
            c, l = oc.cell_search(smin, smax, orders, coords)
            cell_blk = _arr(cell.value + 4 * eo, W * Upad, ctypes.c_int32).reshape(W, Upad)
            cell_blk[:] = 0
            cell_blk[:, :U] = c.reshape(W, U)
            for k in range(d):
                lam_blk = _arr(lam.value + 8 * (k * lam_plane + eo), W * Upad,
                               ctypes.c_double).reshape(W, Upad)
                lam_blk[:] = 0
                lam_blk[:, :U] = l[k].reshape(W, U)
            gsrc = stag[int(r["src"][d]) + u * int(r["us"][d]) + w * int(r["ws"][d])]
            if g_per_w:
                g_blk = _arr(g.value + 8 * go, W * Upad, ctypes.c_double).reshape(W, Upad)
                g_blk[:] = 0
                g_blk[:, :U] = gsrc
            else:
                g_blk = _arr(g.value + 8 * go, Upad, ctypes.c_double)
                g_blk[:] = 0
                g_blk[:U] = gsrc[0]
        self.launches += 1
        return 0

    def sdp_sweep(self, gref, tref, J_prev, part_val, part_idx, J_out, argmin_out, stream):
        d, smin, smax, orders = _grid(gref)
        T = tref._obj
        n_grid = int(np.prod(orders))
        J = _arr(J_prev, n_grid, ctypes.c_double)
        strides = np.concatenate([np.cumprod(orders[::-1])[::-1][1:], [1]]).astype(np.int64)
        items = np.frombuffer((ctypes.c_uint8 * (T.n_items * _cabi.ITEM_DTYPE.itemsize))
                              .from_address(T.items), dtype=_cabi.ITEM_DTYPE)
        p = _arr(T.p, T.W, ctypes.c_double) if T.expect else np.ones(T.W)
        pv = _arr(part_val, T.n_items, ctypes.c_double)
        pi = _arr(part_idx, T.n_items, ctypes.c_int32)
        W = T.W
        for n_it, it in enumerate(items):
            Upad, cnt = int(it["Upad"]), int(it["u_count"])
            acc = np.zeros(cnt)
            for w in range(W):
                off = int(it["entry_base"]) + w * Upad
                c = _arr(T.cell + 4 * off, cnt, ctypes.c_int32).astype(np.int64)
                lam = [_arr(T.lam + 8 * (k * T.lam_plane + off), cnt, ctypes.c_double) for k in range(d)]
                vals = None
                # nested lerp, last axis innermost
                def rec(base, k):
                    if k == d:
                        return J[base]
                    a = rec(base, k + 1)
                    b = rec(base + strides[k], k + 1)
                    return (1 - lam[k]) * a + lam[k] * b
                v = rec(c, 0)
                if T.g_per_w:
                    gv = _arr(T.g + 8 * (int(it["g_base"]) + w * Upad), cnt, ctypes.c_double)
                else:
                    gv = _arr(T.g + 8 * int(it["g_base"]), cnt, ctypes.c_double)
                jg = gv + v
                acc = acc + jg * p[w] if T.expect else jg
            nan = np.isnan(acc)
            j = int(np.argmax(nan)) if nan.any() else int(np.argmin(acc))   # first NaN, else first min
            pv[n_it], pi[n_it] = acc[j], int(it["u_begin"]) + j
        ib = _arr(T.item_begin, T.n_states + 1, ctypes.c_int64)
        Jo = _arr(J_out, T.n_states, ctypes.c_double)
        ao = _arr(argmin_out, T.n_states, ctypes.c_int32)
        for i in range(T.n_states):
            bv, bi = np.inf, 2 ** 31 - 1
            for k in range(ib[i], ib[i + 1]):
                if _better(pv[k], int(pi[k]), bv, bi):
                    bv, bi = pv[k], int(pi[k])
            Jo[i], ao[i] = bv, bi
        self.launches += 2
        return 0

    def sdp_policy_eval(self, gref, W, g_per_w, p, cell, lam, lam_plane, g, n_states, state_begin,
                        n_grid, J_a, J_b, n_iter, rel_dp, ref_index, hist, stream):
        d, smin, smax, orders = _grid(gref)
        strides = np.concatenate([np.cumprod(orders[::-1])[::-1][1:], [1]]).astype(np.int64)
        P = _arr(p, W, ctypes.c_double)
        C = _arr(cell, W * n_states, ctypes.c_int32).astype(np.int64).reshape(W, n_states)
        L = [_arr(lam.value + 8 * k * lam_plane, W * n_states, ctypes.c_double).reshape(W, n_states)
             for k in range(d)]
        Gv = _arr(g, (W if g_per_w else 1) * n_states, ctypes.c_double).reshape(-1, n_states)
        A, B = _arr(J_a, n_grid, ctypes.c_double), _arr(J_b, n_grid, ctypes.c_double)
        H = _arr(hist, n_iter, ctypes.c_double) if rel_dp else None
        src, dst = A, B
        for it in range(n_iter):
            acc = np.zeros(n_states)
            for w in range(W):
                lam_w = [L[k][w] for k in range(d)]

                def rec(base, k):
                    if k == d:
                        return src[base]
                    a = rec(base, k + 1)
                    b = rec(base + strides[k], k + 1)
                    return (1 - lam_w[k]) * a + lam_w[k] * b
                acc = acc + (Gv[w if g_per_w else 0] + rec(C[w], 0)) * P[w]
            dst[state_begin:state_begin + n_states] = acc
            if rel_dp:
                H[it] = dst[ref_index]
                dst -= H[it]
            src, dst = dst, src
        self.launches += n_iter
        return 0

    def sdp_rel_shift(self, J, n, ref_index, ref_out, stream):
        A = _arr(J, n, ctypes.c_double)
        r = _arr(ref_out, 1, ctypes.c_double)
        r[0] = A[ref_index]
        A -= r[0]
        return 0

    def sdp_supnorm_diff(self, a, b, n, out, stream):
        d = np.abs(_arr(a, n, ctypes.c_double) - _arr(b, n, ctypes.c_double))
        d = d[~np.isnan(d)]
        _arr(out, 1, ctypes.c_double)[0] = d.max() if d.size else 0.0
        return 0

    def sdp_interp(self, gref, n_v, values, n_s, s, out, stream):
        d, smin, smax, orders = _grid(gref)
        n_grid = int(np.prod(orders))
        V = _arr(values, n_v * n_grid, ctypes.c_double).reshape(n_v, n_grid)
        S = _arr(s, d * n_s, ctypes.c_double).reshape(d, n_s)
        _arr(out, n_v * n_s, ctypes.c_double)[:] = oc.interp(smin, smax, orders, V, S).reshape(-1)
        return 0

    def sdp_interp_f32(self, gref, n_v, values, n_s, s, out, stream):
        d, smin, smax, orders = _grid(gref)
        n_grid = int(np.prod(orders))
        V = _arr(values, n_v * n_grid, ctypes.c_float).reshape(n_v, n_grid)
        S = _arr(s, d * n_s, ctypes.c_float).reshape(d, n_s)
        _arr(out, n_v * n_s, ctypes.c_float)[:] = oc.interp(smin, smax, orders, V, S).reshape(-1)
        return 0
