"""Host tabulation stage: evaluate the user's callables, exactly as the reference
does, and hand the un-broadcast results to the device table-build kernel.

What the reference does per state inside its hot loop (stodynprog.py:639-677):
    u_grids = control_grids(x_k, t_k)            # control_box + np.linspace
    x_next  = sys.dyn(x_k..., u_grids..., w)     # broadcast over (U1[,U2..],W)
    g       = sys.cost(...)
    J_next_interp(*x_next)                       # np.broadcast_arrays + ravel + cell search
Here the same calls are made once per (system, grids) - the tables are invariant
across sweeps - and the broadcast expansion + cell search run on the GPU
(`sdp_build_tables`), producing the dense (cell, lam, g) tables the sweep kernel
streams.  Nothing in this module does interpolation or minimisation arithmetic.
"""
import itertools

import numpy as np

from . import _cabi

__all__ = ["control_grid_counts", "control_axis_values", "HostStateTable", "tabulate_states"]


def _npts_for(width, step):
    """number of points of one control axis, reference stodynprog.py:446-457.
    Returns 1 for the 'single point at the centre' branch (n_interv < 0.1)."""
    n_interv = width / step
    if n_interv < 0.1:
        return 1
    return int(np.ceil(n_interv) + 1)


def control_grid_counts(intervals, control_steps):
    """(lo, hi, npts) per control for one state's admissible box
    (reference DPSolver.control_grids, stodynprog.py:445-460)."""
    lo, hi, npts = [], [], []
    for (u_min, u_max), step in zip(intervals, control_steps):
        lo.append(u_min)
        hi.append(u_max)
        npts.append(_npts_for(u_max - u_min, step))
    return lo, hi, npts


def make_control_grid(u_min, u_max, npts):
    """the 1-D control grid itself: centre point when npts == 1, else
    np.linspace(u_min, u_max, npts) (reference stodynprog.py:449-458)."""
    if npts == 1:
        return np.array([(u_min + u_max) / 2])
    return np.linspace(u_min, u_max, npts)


def control_axis_values(lo, hi, npts, idx):
    """Vectorised value of control-grid point `idx` for many states at once:
    element-wise the same arithmetic as np.linspace (numpy/_core/function_base.py:
    y = arange(num)*step + start with step = (stop-start)/(num-1), y[-1] = stop;
    (i/div)*delta + start when step == 0) and the centre point for npts == 1.
    Used to map argmin indices back to control VALUES (stodynprog.py:689)."""
    lo = np.asarray(lo, dtype=float)
    hi = np.asarray(hi, dtype=float)
    npts = np.asarray(npts)
    idx = np.asarray(idx)
    div = np.maximum(npts - 1, 1).astype(float)
    delta = hi - lo
    with np.errstate(invalid="ignore", over="ignore"):
        step = delta / div
        y = np.where(step == 0, (idx / div) * delta, idx * step) + lo
        y = np.where(idx == npts - 1, hi, y)
        y = np.where(npts == 1, (lo + hi) / 2, y)
    return y


class HostStateTable(object):
    """Host-side record of one shard's control discretisation."""

    def __init__(self, n_states, nb_control):
        self.lo = np.zeros((n_states, nb_control))
        self.hi = np.zeros((n_states, nb_control))
        self.npts = np.ones((n_states, nb_control), dtype=np.int64)

    @property
    def U(self):
        return self.npts.prod(axis=1)


def _compact(a, control_dims, U, W, what):
    """Reduce one dyn/cost output to a C-contiguous (Ueff, Weff) fp64 array with
    Ueff in {1,U}, Weff in {1,W}: the un-broadcast form of what
    MlinInterpolator.__call__ would expand (stodynprog.py:281-283:
    broadcast_arrays -> astype(float) -> ravel)."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(float)
    nb = len(control_dims) + 1
    if a.ndim > nb:
        raise ValueError("%s returned an array of rank %d; expected something broadcastable "
                         "to controls + perturbation shape %s" % (what, a.ndim, control_dims + (W,)))
    shape = (1,) * (nb - a.ndim) + a.shape
    w_eff = shape[-1]
    if w_eff != 1 and w_eff != W:
        raise ValueError("%s output has %d entries along the perturbation axis, expected 1 or %d"
                         % (what, w_eff, W))
    ctrl = shape[:-1]
    if all(s == 1 for s in ctrl):
        return np.ascontiguousarray(a.reshape(1, w_eff)), 1, w_eff
    for s, c in zip(ctrl, control_dims):
        if s != 1 and s != c:
            raise ValueError("%s output shape %s does not broadcast to the control grid %s"
                             % (what, a.shape, control_dims))
    a = a.reshape(shape)
    if ctrl != tuple(control_dims):
        a = np.broadcast_to(a, tuple(control_dims) + (w_eff,))
    return np.ascontiguousarray(a.reshape(U, w_eff)), U, w_eff


class _ChunkWriter(object):
    """Accumulates staged arrays + descriptors and flushes them to the device."""

    def __init__(self, d, flush_fn, max_doubles):
        self.d = d
        self.flush_fn = flush_fn
        self.max_doubles = max_doubles
        self.reset()

    def reset(self):
        self.arrays = []
        self.n_doubles = 0
        self.desc = []
        self.max_Upad = 0

    def add_state(self, entry_off, g_off, U, Upad, compact):
        """compact: list of d+1 tuples (array(Ueff,Weff), Ueff, Weff)"""
        rec = np.zeros((), dtype=_cabi.STATE_DESC_DTYPE)
        rec["entry_off"] = entry_off
        rec["g_off"] = g_off
        rec["U"] = U
        rec["Upad"] = Upad
        slots = list(range(self.d)) + [self.d]   # slot d holds the stage cost
        for k, (arr, u_eff, w_eff) in zip(slots, compact):
            rec["src"][k] = self.n_doubles
            rec["us"][k] = w_eff if u_eff > 1 else 0
            rec["ws"][k] = 1 if w_eff > 1 else 0
            self.arrays.append(arr.ravel())
            self.n_doubles += arr.size
        self.desc.append(rec)
        self.max_Upad = max(self.max_Upad, Upad)
        if self.n_doubles >= self.max_doubles:
            self.flush()

    def flush(self):
        if not self.desc:
            return
        staging = np.concatenate(self.arrays) if self.arrays else np.zeros(1)
        desc = np.array(self.desc, dtype=_cabi.STATE_DESC_DTYPE)
        self.flush_fn(desc, staging, self.max_Upad)
        self.reset()


def scan_control_boxes(sys, control_steps, states, t_k=None):
    """First pass: admissible box and control counts of every state of the shard
    (one `control_box` call per state, like the reference's control_grids)."""
    nb_control = len(sys.control)
    n = len(states)
    tab = HostStateTable(n, nb_control)
    params = sys.params
    for i, x_k in enumerate(states):
        args = x_k if t_k is None else (t_k,) + x_k
        intervals = sys.control_box(*args, **params)
        lo, hi, npts = control_grid_counts(intervals, control_steps)
        if len(lo) != nb_control:
            raise ValueError("control_box returned %d intervals, control_steps has %d steps and "
                             "the system has %d controls" % (len(intervals), len(control_steps), nb_control))
        tab.lo[i] = lo
        tab.hi[i] = hi
        tab.npts[i] = npts
    return tab


def tabulate_states(sys, states, host_tab, perturb_grid, t_k, entry_off, g_off, Upad, g_per_w,
                    flush_fn, max_doubles=8 << 20):
    """Second pass: call dyn/cost per state with the reference's argument
    shapes (stodynprog.py:655-676) and stage the un-broadcast outputs.

    Returns True on success, or False if a state's cost turned out to depend on
    the perturbation while `g_per_w` is 0 (the caller restarts in dense-g mode).
    """
    d = len(sys.state)
    nb_control = len(sys.control)
    W = len(perturb_grid[0]) if len(perturb_grid) > 0 else 1
    params = sys.params
    writer = _ChunkWriter(d, flush_fn, max_doubles)
    w_args = tuple(perturb_grid)
    for i, x_k in enumerate(states):
        npts = host_tab.npts[i]
        control_dims = tuple(int(n) for n in npts)
        U = int(np.prod(control_dims)) if nb_control else 1
        u_grids = []
        for c in range(nb_control):
            ug = make_control_grid(host_tab.lo[i, c], host_tab.hi[i, c], control_dims[c])
            # control c varies along axis c, the perturbation along the last axis
            ug.shape = (1,) * c + (-1,) + (1,) * (nb_control - c)
            u_grids.append(ug)
        args = x_k + tuple(u_grids) + w_args
        if t_k is not None:
            args = (t_k,) + args
        x_next = sys.dyn(*args, **params)
        g_k = sys.cost(*args, **params)
        if len(x_next) != d:
            raise ValueError("dyn returned %d next-state components, expected %d" % (len(x_next), d))
        compact = [_compact(c, control_dims, U, W, "dyn") for c in x_next]
        compact.append(_compact(g_k, control_dims, U, W, "cost"))
        # the joint broadcast of (g, x_next...) must cover the whole control grid
        # (stodynprog.py:683 asserts J.shape == control_dims)
        if U > 1 and not any(u_eff == U for _, u_eff, _ in compact):
            raise AssertionError("dyn/cost outputs do not span the control grid %s" % (control_dims,))
        if compact[-1][2] > 1 and not g_per_w:
            return False
        writer.add_state(int(entry_off[i]), int(g_off[i]), U, int(Upad[i]), compact)
    writer.flush()
    return True


def state_tuples(state_grid, begin, end):
    """states [begin, end) of the C-order flattened grid as tuples of numpy
    scalars, the same objects `itertools.product(*self.state_grid)` yields in the
    reference (stodynprog.py:478)."""
    it = itertools.product(*state_grid)
    return list(itertools.islice(it, begin, end))
