"""CPU port of the reference solver loops (numpy, per-state Python loop).

TEST INFRASTRUCTURE ONLY - the checker of the parity tests and the timed
`cpu_baseline` ("port") of bench.py.  It follows the reference's structure
statement by statement so that it both (a) produces the reference's results
and (b) costs what the reference costs on a CPU:

  value_iteration      stodynprog/stodynprog.py:466-534
  _value_at_state_vect :639-691   (np.inner for the expectation, ndarray.argmin)
  control_grids        :432-463
  eval_policy          :693-775
  policy_iteration     :777-812
  bellman_recursion    :536-591
  MlinInterpolator     :255-290

The interpolation call goes to the C restatement (oracle/sdp_oracle.c), or to
the reference's own compiled Cython routine (oracle/_ref) with interp="ref".
Problem descriptions are the product's `SysDescription` objects (pure holders
of callables - no arithmetic lives there).

Validated in this container against the unmodified reference: bit-identical J
and policies on configs #1-#4 (tests/test_oracle.py, tests/golden/).
"""
import itertools

import numpy as np

from . import oracle as _oc

__all__ = ["PortSolver", "PortInterpolator", "port_api"]


class PortInterpolator(object):
    """stodynprog.py:255-290"""

    def __init__(self, x_grid, values, interp="c"):
        self.ndim = len(x_grid)
        self.smin = np.array([x[0] for x in x_grid], dtype=float)
        self.smax = np.array([x[-1] for x in x_grid], dtype=float)
        self.orders = np.array([len(x) for x in x_grid], dtype=np.int64)
        self.values = np.ascontiguousarray(np.atleast_2d(np.asarray(values, dtype=float).ravel()))
        if interp == "ref":
            from .ref_loader import load_reference_cython
            cy = load_reference_cython()
            if cy is None:
                raise RuntimeError("oracle/_ref is not built")
            self._fn = lambda s: cy.multilinear_interpolation(self.smin, self.smax, self.orders,
                                                              self.values, s)
        else:
            self._fn = lambda s: _oc.interp(self.smin, self.smax, self.orders, self.values, s)

    def __call__(self, *x_interp):
        mesh = np.broadcast_arrays(*x_interp)                       # :281
        shape = mesh[0].shape
        stack = np.vstack([np.asarray(x).astype(float).ravel() for x in mesh])   # :283
        return np.asarray(self._fn(stack)).reshape(shape)           # :285-287


class PortSolver(object):
    """numpy restatement of DPSolver; same public surface as far as the hot
    path goes."""

    def __init__(self, sys, interp="c"):
        self.sys = sys
        self.interp_backend = interp
        self.state_grid = [[0.] for s in sys.state]
        self.perturb_grid = [[0.] for p in sys.perturb]
        self.perturb_proba = [[1.] for p in sys.perturb]
        self.control_steps = (1.,) * len(sys.control)

    # grids: stodynprog.py:335-389
    def discretize_perturb(self, *a):
        assert len(a) == len(self.sys.perturb) * 3
        self.perturb_grid, self.perturb_proba = [], []
        for i in range(len(self.sys.perturb)):
            grid = np.linspace(*a[3 * i:3 * i + 3])
            if self.sys.perturb_types[i] == 'continuous':
                proba = self.sys.perturb_laws[i].pdf(grid)
                proba /= proba.sum()
            else:
                proba = self.sys.perturb_laws[i].pmf(grid)
                assert np.allclose(proba.sum(), 1.)
            self.perturb_grid.append(grid)
            self.perturb_proba.append(proba)
        return self.perturb_grid, self.perturb_proba

    def discretize_state(self, *a):
        assert len(a) == len(self.sys.state) * 3
        self.state_grid = [np.linspace(*a[3 * i:3 * i + 3]) for i in range(len(self.sys.state))]
        self._state_grid_shape = tuple(len(g) for g in self.state_grid)
        self._state_ref_ind = tuple(n // 2 for n in self._state_grid_shape)
        self._state_ref = tuple(g[i] for g, i in zip(self.state_grid, self._state_ref_ind))
        return self.state_grid

    @property
    def state_grid_full(self):
        nd = len(self.state_grid)
        return np.broadcast_arrays(*[np.reshape(g, (1,) * i + (-1,) + (1,) * (nd - i - 1))
                                     for i, g in enumerate(self.state_grid)])

    def interp_on_state(self, A):
        if A.shape != self._state_grid_shape:
            raise ValueError('array `A` should be of shape {:s}, not {:s}'.format(
                str(self._state_grid_shape), str(A.shape)))
        return PortInterpolator(self.state_grid, A, self.interp_backend)

    # stodynprog.py:432-463
    def control_grids(self, state_k, t_k=None):
        if t_k is not None:
            state_k = (t_k,) + state_k
        boxes = self.sys.control_box(*state_k, **self.sys.params)
        grids, dims = [], []
        for (u_min, u_max), step in zip(boxes, self.control_steps):
            n_interv = (u_max - u_min) / step
            if n_interv < 0.1:
                npts, grid = 1, np.array([(u_min + u_max) / 2])
            else:
                npts = int(np.ceil(n_interv) + 1)
                grid = np.linspace(u_min, u_max, npts)
            grids.append(grid)
            dims.append(npts)
        return grids, tuple(dims)

    def _perturb_product(self):
        """Several perturbations - a TODO of the reference (stodynprog.py:614,666,679-683,728), so
        there is no reference behaviour to restate; the convention defined here (and documented
        in DESIGN.md) is the natural extension of its one-perturbation code: grid j enters dyn /
        cost on its own axis behind the controls, shape (1,)*j + (W_j,) + (1,)*(m-1-j); the
        joint law is the product of the (independent) marginals; the expectation is np.inner
        over the C-order flattened product grid.  Returns (w_args, w_shape, p_flat)."""
        m = len(self.perturb_grid)
        w_shape = tuple(len(g) for g in self.perturb_grid)
        w_args = tuple(np.asarray(g).reshape((1,) * j + (-1,) + (1,) * (m - 1 - j))
                       for j, g in enumerate(self.perturb_grid))
        p = np.ones(1)
        for q in self.perturb_proba:
            p = np.multiply.outer(p, np.asarray(q, dtype=float)).reshape(-1)
        return w_args, w_shape, p

    # stodynprog.py:639-691
    def value_at_state(self, x_k, J_interp, t_k=None, want_index=False):
        u_grids, control_dims = self.control_grids(x_k, t_k)
        nc = len(u_grids)
        nb_perturb = len(self.perturb_grid)
        n_w_axes = max(nb_perturb, 1)
        for i in range(nc):
            u_grids[i].shape = (1,) * i + (-1,) + (1,) * (nc - 1 - i + n_w_axes)
        if nb_perturb >= 2:
            w_args, w_shape, p_flat = self._perturb_product()
        else:
            w_args = tuple(self.perturb_grid)
        args = x_k + tuple(u_grids) + w_args
        if t_k is not None:
            args = (t_k,) + args
        x_next = self.sys.dyn(*args, **self.sys.params)
        g = self.sys.cost(*args, **self.sys.params)
        J_grid = g + J_interp(*x_next)
        if nb_perturb == 0:
            J = J_grid
        elif nb_perturb == 1:
            J = np.inner(J_grid, self.perturb_proba[0])
            assert J.shape == control_dims
        else:
            J_grid = np.broadcast_to(J_grid, control_dims + w_shape).reshape(control_dims + (-1,))
            J = np.inner(J_grid, p_flat)
        flat = J.argmin()
        ind = np.unravel_index(flat, control_dims)
        u_opt = [u_grids[i].flatten()[ind[i]] for i in range(nc)]
        if want_index:
            return J[ind], u_opt, int(flat), J
        return J[ind], u_opt

    # stodynprog.py:466-534
    def value_iteration(self, J_next, rel_dp=False, report_time=False, want_index=False,
                        state_slice=None):
        dims = tuple(len(g) for g in self.state_grid)
        ref_ind = getattr(self, '_state_ref_ind', None)
        if rel_dp:
            J_next, J_ref = J_next
            assert J_next[ref_ind] == 0.
        nc = len(self.sys.control)
        J_k = np.zeros(dims)
        pol_k = np.zeros(dims + (nc,))
        idx_k = np.zeros(dims, dtype=np.int64)
        J_interp = self.interp_on_state(J_next)
        pairs = zip(itertools.product(*[range(n) for n in dims]), itertools.product(*self.state_grid))
        if state_slice is not None:
            pairs = itertools.islice(pairs, *state_slice)
        for ind_x, x_k in pairs:
            if want_index:
                J_k[ind_x], pol_k[ind_x], idx_k[ind_x], _ = self.value_at_state(x_k, J_interp, None, True)
            else:
                J_k[ind_x], pol_k[ind_x] = self.value_at_state(x_k, J_interp)
        if rel_dp:
            J_ref = J_k[ref_ind]
            J_k -= J_ref
            J_k = J_k, J_ref
        if want_index:
            return J_k, pol_k, idx_k
        return J_k, pol_k

    # stodynprog.py:536-591
    def bellman_recursion(self, t_fin, J_fin, t_ini=0, report_time=False):
        dims = tuple(len(g) for g in self.state_grid)
        nc = len(self.sys.control)
        assert t_ini == 0
        J = np.zeros((t_fin - t_ini,) + dims)
        pol = np.zeros((t_fin - t_ini,) + dims + (nc,))
        for t_k in range(t_ini, t_fin)[::-1]:
            k = t_k - t_ini
            J_interp = self.interp_on_state(J_fin if t_k == t_fin - 1 else J[k + 1])
            for ind_x, x_k in zip(itertools.product(*[range(n) for n in dims]),
                                  itertools.product(*self.state_grid)):
                J[k][ind_x], pol[k][ind_x] = self.value_at_state(x_k, J_interp, t_k)
        return J, pol

    # stodynprog.py:693-775
    def eval_policy(self, pol, n_iter, rel_dp=False, J_zero=None, report_time=False,
                    J_ref_full=False):
        dims = self._state_grid_shape
        ns = len(self.sys.state)
        J_pol = np.zeros(dims) if J_zero is None else J_zero
        assert J_pol.shape == dims
        J_ref = np.zeros(n_iter)
        ref_ind = self._state_ref_ind
        nc = len(self.sys.control)
        assert pol.shape == dims + (nc,)
        w_k, w_proba = self.perturb_grid[0], self.perturb_proba[0]
        m = len(self.perturb_grid)
        w_args, w_shape = (w_k,), (len(w_k),)
        if m >= 2:
            w_args, w_shape, w_proba = self._perturb_product()
        sg = tuple(np.reshape(self.state_grid[i], (1,) * i + (-1,) + (1,) * (ns - 1 - i + m))
                   for i in range(ns))
        for k in range(n_iter):
            J_interp = self.interp_on_state(J_pol)
            u_k = [pol[..., i].reshape(dims + (1,) * m) for i in range(nc)]
            args = sg + tuple(u_k) + w_args
            x_next = self.sys.dyn(*args, **self.sys.params)
            g = self.sys.cost(*args, **self.sys.params)
            J_grid = g + J_interp(*x_next)
            if m >= 2:
                J_grid = np.broadcast_to(J_grid, dims + w_shape).reshape(dims + (-1,))
            J_pol = np.inner(J_grid, w_proba)
            if rel_dp:
                J_ref[k] = J_pol[ref_ind]
                J_pol -= J_ref[k]
        if rel_dp:
            return J_pol, (J_ref if J_ref_full else J_ref[-1])
        return J_pol

    # stodynprog.py:777-812
    def policy_iteration(self, pol_init, n_val, n_pol=1, rel_dp=False):
        pol = pol_init
        J_pol = self.eval_policy(pol, n_val, rel_dp)
        self.ref_costs = [J_pol[1]] if rel_dp else []
        for k in range(n_pol):
            _, pol = self.value_iteration(J_pol, rel_dp=rel_dp)
            J_pol = self.eval_policy(pol, n_val, rel_dp)
            if rel_dp:
                self.ref_costs.append(J_pol[1])
        return J_pol, pol


class _PortAPI(object):
    """`api` object for the tests/workloads.py factories"""

    def __init__(self, interp="c"):
        from stodynprog_b200.sysdesc import SysDescription
        self.SysDescription = SysDescription
        self._interp = interp

    def DPSolver(self, sys, **kw):
        return PortSolver(sys, interp=self._interp)


def port_api(interp="c"):
    return _PortAPI(interp)
