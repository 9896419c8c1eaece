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

__all__ = ["scan_control_boxes", "scan_control_boxes_batched", "scan_control_boxes_by_axes", "scan_control_boxes_parallel", "state_tuples_at", "control_grid_counts", "control_axis_values", "HostStateTable", "tabulate_states",
           "tabulate_states_batched", "GDependsOnW", "BatchedMismatch", "NotFactorable",
           "probe_factor_mask", "check_factorable", "perturb_layout", "joint_proba"]


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
    Used to map argmin indices back to control VALUES (stodynprog.py:689) and to
    lay out the control grids of a chunk of states for the batched tabulation."""
    lo = np.asarray(lo, dtype=float)
    hi = np.asarray(hi, dtype=float)
    npts = np.asarray(npts)
    idx = np.asarray(idx)
    div = np.maximum(npts - 1, 1).astype(float)
    delta = hi - lo
    with np.errstate(invalid="ignore", over="ignore"):
        step = delta / div
        y = idx * step
        y += lo
        flat = step == 0
        if np.any(flat):                       # rare: degenerate or denormal-step boxes
            y = np.where(flat, (idx / div) * delta + lo, y)
        y = np.where(idx == npts - 1, hi, y)
        single = npts == 1
        if np.any(single):
            y = np.where(single, (lo + hi) / 2, y)
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


def perturb_layout(perturb_grid):
    """(w_args, w_shape, W) of a solver's perturbation grids.

    One perturbation: its grid goes to dyn/cost as the reference passes it, a (W,) vector on the
    last axis (stodynprog.py:667-672).  Several perturbations - which the reference leaves as a
    TODO (stodynprog.py:614,666,679-683,728) - each get an axis of their own behind the control
    axes: grid j is passed with shape (1,)*j + (W_j,) + (1,)*(m-1-j), the outputs broadcast to
    controls + (W_1, .., W_m), and the product grid is flattened in C order into ONE axis of
    W = W_1*..*W_m nodes whose probabilities are the products of the marginals
    (`joint_proba`; the laws of a SysDescription are independent).  The expectation is then the
    reference's np.inner over that axis, nodes summed in C order of the product."""
    m = len(perturb_grid)
    if m == 0:
        return (), (), 1
    grids = [np.asarray(g) for g in perturb_grid]
    w_shape = tuple(len(g) for g in grids)
    if m == 1:
        return (grids[0],), w_shape, w_shape[0]
    w_args = tuple(g.reshape((1,) * j + (-1,) + (1,) * (m - 1 - j)) for j, g in enumerate(grids))
    return w_args, w_shape, int(np.prod(w_shape))


def joint_proba(perturb_proba):
    """probabilities of the flattened product grid of `perturb_layout` (C order)"""
    p = np.ones(1)
    for q in perturb_proba:
        p = np.multiply.outer(p, np.asarray(q, dtype=float)).reshape(-1)
    return np.ascontiguousarray(p)


def _fold_w(a, lead, w_shape, what):
    """output of dyn/cost with `lead` leading axes (controls, or states + controls) followed by
    one axis per perturbation -> the same with ONE trailing axis: the C-order flattened product of
    the perturbation axes, or size 1 when the output spans none of them"""
    m = len(w_shape)
    if m <= 1:
        return a
    if a.ndim > lead + m:
        raise ValueError("%s returned an array of rank %d; expected something broadcastable to "
                         "%d leading axes + the perturbation grid %s" % (what, a.ndim, lead, w_shape))
    a = a.reshape((1,) * (lead + m - a.ndim) + a.shape)
    tail = a.shape[lead:]
    for n_a, n_w in zip(tail, w_shape):
        if n_a != 1 and n_a != n_w:
            raise ValueError("%s output shape %s does not broadcast to the perturbation grid %s"
                             % (what, a.shape, w_shape))
    if all(n_a == 1 for n_a in tail):
        return a.reshape(a.shape[:lead] + (1,))
    a = np.broadcast_to(a, a.shape[:lead] + tuple(w_shape))
    return a.reshape(a.shape[:lead] + (-1,))


def _compact(a, control_dims, U, W, what, w_shape=()):
    """Reduce one dyn/cost output to a C-contiguous (Ueff, Weff) fp64 array with
    Ueff in {1,U}, Weff in {1,W}: the un-broadcast form of what
    MlinInterpolator.__call__ would expand (stodynprog.py:281-283:
    broadcast_arrays -> astype(float) -> ravel)."""
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(float)
    a = _fold_w(a, len(control_dims), w_shape, what)
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
    """Accumulates staged arrays + per-state descriptors (SdpStateDesc) and
    flushes them to the device at chunk boundaries.  `align` = number of states
    a flush must be a multiple of (32 for the state-minor layout)."""

    def __init__(self, d, flush_fn, max_doubles, align=1):
        self.d = d
        self.flush_fn = flush_fn
        self.max_doubles = max_doubles
        self.align = align
        self.reset()

    def reset(self):
        self.arrays = []
        self.n_doubles = 0
        self.descs = []          # list of structured arrays
        self.n_states = 0

    def add_array(self, arr):
        """stage one array, return its offset (in doubles)"""
        off = self.n_doubles
        self.arrays.append(arr.reshape(-1))
        self.n_doubles += arr.size
        return off

    def add_descs(self, recs):
        self.descs.append(recs)
        self.n_states += len(recs)
        if self.n_doubles >= self.max_doubles and self.n_states % self.align == 0:
            self.flush()

    def flush(self):
        if not self.n_states:
            return
        # (the staged arrays travel as a list: the caller copies each straight into its
        # page-locked upload buffer instead of concatenating first)
        staging = self.arrays if self.arrays else [np.zeros(1)]
        desc = np.concatenate(self.descs)
        self.flush_fn(desc, staging)
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


# ---------------------------------------------------------------------------
# the same scan on several host cores.  The reference's own examples write their box
# functions with np.max((a, b)) / np.min((a, b)) on scalars, which neither vectorise nor run
# fast (10 us per state: 11 s for the 10^6 states of config #5, by far the longest stage of
# a table build).  The calls are independent, so the range of states is cut into chunks that
# forked workers scan one state at a time, exactly as above; the user's callables reach the
# workers through the fork (closures cannot be pickled), only float arrays come back.
# ---------------------------------------------------------------------------
_SCAN_JOB = None      # (sys, control_steps, state_grid, t_k), set just before the workers fork


def _scan_chunk(bounds):
    sys, control_steps, state_grid, t_k = _SCAN_JOB
    tab = scan_control_boxes(sys, control_steps, state_tuples_at(state_grid, bounds[0], bounds[1]), t_k)
    return tab.lo, tab.hi, tab.npts


def scan_control_boxes_parallel(sys, control_steps, state_grid, begin, end, t_k=None, procs=2,
                                timeout=None):
    """`scan_control_boxes` for the states [begin, end) of the C-order grid on `procs` forked
    workers.  Returns a HostStateTable, or None when forking is not available, a worker
    failed or the workers did not answer within `timeout` seconds (the caller then scans
    serially, which also reproduces any exception of the user's control_box in place)."""
    global _SCAN_JOB
    import multiprocessing as mp
    import os
    import time
    import warnings
    n = end - begin
    try:
        procs = min(procs, len(os.sched_getaffinity(0)))     # cores this process may use
    except AttributeError:
        pass
    if procs < 2 or n < 2 * procs or "fork" not in mp.get_all_start_methods():
        return None
    n_chunks = procs * 4
    cuts = [begin + n * k // n_chunks for k in range(n_chunks + 1)]
    chunks = [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    if timeout is None:
        # ten times what a timed serial sample predicts for the workers' share
        k = min(n, 256)
        t0 = time.perf_counter()
        try:
            scan_control_boxes(sys, control_steps, state_tuples_at(state_grid, begin, begin + k), t_k)
        except Exception:
            return None      # (the serial scan reproduces the user's exception in place)
        timeout = 10.0 + 10.0 * (time.perf_counter() - t0) / k * n / procs
    _SCAN_JOB = (sys, control_steps, state_grid, t_k)
    pool = None
    try:
        pool = mp.get_context("fork").Pool(procs)
        parts = pool.map_async(_scan_chunk, chunks).get(timeout=timeout)
    except Exception as e:
        warnings.warn("parallel control_box scan failed (%s: %s); scanning serially"
                      % (type(e).__name__, e))
        return None
    finally:
        _SCAN_JOB = None
        if pool is not None:
            pool.terminate()
    tab = HostStateTable(n, len(sys.control))
    tab.lo = np.concatenate([p[0] for p in parts], axis=0)
    tab.hi = np.concatenate([p[1] for p in parts], axis=0)
    tab.npts = np.concatenate([p[2] for p in parts], axis=0)
    return tab


def scan_control_boxes_by_axes(sys, control_steps, state_grid, begin, end, t_k=None,
                               verify_min=2048, verify_frac=1.0 / 128, seed=0):
    """First pass for box functions that read only SOME of the state variables - the rule in the
    reference's examples: the admissible storage power depends on the stored energy alone
    (examples/howto storage-AR1.ipynb:214, storage_control.py:67-79), so the 10^6 states of
    config #5 hold 2 000 distinct boxes - without vectorising the user's function (the examples'
    `np.max((a, b))` on scalars cannot take arrays).

    1. Hypothesis: from the middle state of the grid, each state axis is moved alone through up
       to 9 of its values; an axis whose moves never change the box (bit for bit) is taken as
       ignored.
    2. `control_box` is called the reference's way (stodynprog.py:440), once per point of the
       product of the axes that are NOT ignored, the ignored ones held at their middle value.
    3. The table is broadcast to the states [begin, end) of the C-order grid and CHECKED: at
       least `verify_min` (and `verify_frac` of the) states, seeded random plus the first and
       last, are re-evaluated one by one with their own coordinates and must give the same
       (lo, hi, npts) bit for bit - the trust level of the batched dyn/cost evaluation.
    Returns a HostStateTable, or None when no axis can be dropped, the saving is below 4x, or
    the check fails (the caller then scans every state)."""
    nb_control = len(sys.control)
    dims = [len(g) for g in state_grid]
    d = len(dims)
    n = end - begin
    if n < 4096 or nb_control == 0 or d < 2:
        return None
    grids = [np.asarray(g) for g in state_grid]
    mid = [m // 2 for m in dims]

    def box_at(idx):
        x_k = tuple(grids[k][i] for k, i in enumerate(idx))
        return scan_control_boxes(sys, control_steps, [x_k], t_k)

    def same(a, b):
        return (np.array_equal(a.lo.view(np.int64), b.lo.view(np.int64))
                and np.array_equal(a.hi.view(np.int64), b.hi.view(np.int64))
                and np.array_equal(a.npts, b.npts))

    try:
        base = box_at(mid)
        used = []
        for k in range(d):
            moves = np.unique(np.linspace(0, dims[k] - 1, min(9, dims[k])).astype(int))
            if any(not same(base, box_at(mid[:k] + [int(i)] + mid[k + 1:])) for i in moves):
                used.append(k)
        n_sub = int(np.prod([dims[k] for k in used])) if used else 1
        if len(used) == d or n_sub * 4 > n:
            return None
        # the boxes over the product of the axes the function reads
        sub_dims = [dims[k] for k in used]
        sub_states = []
        for sub in itertools.product(*[range(m) for m in sub_dims]):
            idx = list(mid)
            for k, i in zip(used, sub):
                idx[k] = i
            sub_states.append(tuple(grids[k][i] for k, i in enumerate(idx)))
        sub_tab = scan_control_boxes(sys, control_steps, sub_states, t_k)
        full_idx = np.unravel_index(np.arange(begin, end), dims)
        sub_flat = np.ravel_multi_index([full_idx[k] for k in used], sub_dims) if used \
            else np.zeros(n, dtype=np.int64)
        tab = HostStateTable(n, nb_control)
        tab.lo, tab.hi, tab.npts = sub_tab.lo[sub_flat], sub_tab.hi[sub_flat], sub_tab.npts[sub_flat]
        # check on states evaluated with their own coordinates
        n_check = min(n, max(verify_min, int(n * verify_frac)))
        picks = np.random.default_rng(seed).choice(n, size=n_check, replace=False)
        picks = np.unique(np.concatenate([picks, [0, n - 1]]))
        states = [tuple(grids[k][full_idx[k][i]] for k in range(d)) for i in picks]
        ref = scan_control_boxes(sys, control_steps, states, t_k)
    except Exception:
        return None          # (the per-state scan reproduces the user's exception in place)
    ok = (np.array_equal(tab.lo[picks].view(np.int64), ref.lo.view(np.int64))
          and np.array_equal(tab.hi[picks].view(np.int64), ref.hi.view(np.int64))
          and np.array_equal(tab.npts[picks], ref.npts))
    return tab if ok else None


def scan_control_boxes_batched(sys, control_steps, state_grid, begin, end, t_k=None, verify=16):
    """First pass in ONE `control_box` call for states [begin, end) of the C-order
    grid: every state variable enters as an (S,) array.  Works for box functions
    written with element-wise numpy (np.maximum/np.where/...); the reference's own
    examples use `np.max((a, b))`, which does not vectorise - those raise or fail
    the check below and the caller falls back to the per-state scan.

    `verify` sample states are re-evaluated one by one, the reference's way
    (stodynprog.py:440-460), and must agree bit-for-bit on (lo, hi, npts).
    Returns a HostStateTable, or None when the batched call cannot be trusted."""
    nb_control = len(sys.control)
    S = end - begin
    if S <= 0 or nb_control == 0:
        return None
    dims = [len(g) for g in state_grid]
    idx = np.unravel_index(np.arange(begin, end), dims)
    cols = tuple(np.asarray(state_grid[k])[idx[k]] for k in range(len(dims)))
    args = cols if t_k is None else (t_k,) + cols
    try:
        with np.errstate(all="ignore"):
            intervals = sys.control_box(*args, **sys.params)
            if len(intervals) != nb_control or len(control_steps) != nb_control:
                return None
            tab = HostStateTable(S, nb_control)
            for c, ((u_min, u_max), step) in enumerate(zip(intervals, control_steps)):
                lo = np.asarray(u_min, dtype=float)
                hi = np.asarray(u_max, dtype=float)
                if lo.ndim > 1 or hi.ndim > 1 or lo.size not in (1, S) or hi.size not in (1, S):
                    return None
                tab.lo[:, c] = lo.reshape(-1)
                tab.hi[:, c] = hi.reshape(-1)
                n_interv = (tab.hi[:, c] - tab.lo[:, c]) / step
                if not np.all(np.isfinite(n_interv)) or np.any(n_interv >= 2 ** 31):
                    return None
                tab.npts[:, c] = np.where(n_interv < 0.1, 1, (np.ceil(n_interv) + 1).astype(np.int64))
    except Exception:
        return None
    picks = np.unique(np.linspace(0, S - 1, min(verify, S)).astype(int))
    states = [tuple(col[i] for col in cols) for i in picks]
    try:
        ref = scan_control_boxes(sys, control_steps, states, t_k)
    except Exception:
        return None
    same = (np.array_equal(tab.lo[picks].view(np.int64), ref.lo.view(np.int64))
            and np.array_equal(tab.hi[picks].view(np.int64), ref.hi.view(np.int64))
            and np.array_equal(tab.npts[picks], ref.npts))
    return tab if same else None


class NotFactorable(Exception):
    """a chunk of staged dyn/cost outputs does not have the (x,u) + (x,w)
    structure the factored tables were being built for: restart dense"""


def probe_factor_mask(sys, x_k, host_tab, i, perturb_grid, t_k):
    """Look at what dyn/cost return for ONE state (the reference's own call,
    stodynprog.py:674-676) and classify every next-state coordinate by the
    axes its un-broadcast output spans: controls only, perturbation only, or
    neither.  Returns the bit mask of the coordinates to put in the (x,u) part
    of factored tables, or None when some coordinate spans
    both axes, the cost depends on w, or the split would be one-sided."""
    d = len(sys.state)
    w_args, w_shape, W = perturb_layout(perturb_grid)
    compact, control_dims, U = _eval_one_state(sys, x_k, host_tab, i, w_args, t_k, W, w_shape)
    if compact[-1][2] > 1:
        return None                       # g depends on w
    mask = 0
    free = []                             # coordinates that depend on neither u nor w
    for k in range(d):
        _, u_eff, w_eff = compact[k]
        if u_eff > 1 and w_eff > 1:
            return None
        if u_eff > 1:
            mask |= 1 << k
        elif w_eff == 1:
            free.append(k)
    full = (1 << d) - 1
    if mask == 0 and free:
        mask |= 1 << free.pop(0)          # need at least one (x,u) coordinate
    if mask == full or mask == 0:
        return None
    return mask


def check_factorable(desc, d, u_mask):
    """every state record of a chunk must agree with the split: (x,u)
    coordinates and g have no perturbation stride, (x,w) coordinates no control
    stride.  Raises NotFactorable otherwise."""
    for k in range(d):
        if (u_mask >> k) & 1:
            ok = not np.any(desc["ws"][:, k])
        else:
            ok = not np.any(desc["cs"][:, k, :])
        if not ok:
            raise NotFactorable()
    if np.any(desc["ws"][:, d]):
        raise NotFactorable()


class GDependsOnW(Exception):
    """the stage cost depends on the perturbation but the tables were being
    built with one g per (state, control): restart with a dense g table"""


def _eval_one_state(sys, x_k, host_tab, i, w_args, t_k, W, w_shape=(), only=None):
    """dyn/cost of one state with the reference's argument shapes
    (stodynprog.py:655-676).  Returns [(array(Ueff,Weff), Ueff, Weff)] * (d+1).
    `w_args`, `W`, `w_shape`: see perturb_layout."""
    d = len(sys.state)
    nb_control = len(sys.control)
    control_dims = tuple(int(n) for n in host_tab.npts[i])
    U = int(np.prod(control_dims)) if nb_control else 1
    n_w_axes = max(len(w_shape), 1)
    u_grids = []
    for c in range(nb_control):
        ug = make_control_grid(host_tab.lo[i, c], host_tab.hi[i, c], control_dims[c])
        # control c varies along axis c, the perturbation(s) along the last axis (axes)
        ug.shape = (1,) * c + (-1,) + (1,) * (nb_control - 1 - c + n_w_axes)
        u_grids.append(ug)
    args = x_k + tuple(u_grids) + w_args
    if t_k is not None:
        args = (t_k,) + args
    if only == "cost":          # (see _eval_state_chunk)
        return [_compact(sys.cost(*args, **sys.params), control_dims, U, W, "cost", w_shape)], control_dims, U
    x_next = sys.dyn(*args, **sys.params)
    if len(x_next) != d:
        raise ValueError("dyn returned %d next-state components, expected %d" % (len(x_next), d))
    compact = [_compact(c, control_dims, U, W, "dyn", w_shape) for c in x_next]
    if only == "dyn":
        return compact, control_dims, U
    g_k = sys.cost(*args, **sys.params)
    compact.append(_compact(g_k, control_dims, U, W, "cost", w_shape))
    # the joint broadcast of (g, x_next...) must cover the whole control grid
    # (stodynprog.py:683 asserts J.shape == control_dims)
    if U > 1 and not any(u_eff == U for _, u_eff, _ in compact):
        raise AssertionError("dyn/cost outputs do not span the control grid %s" % (control_dims,))
    return compact, control_dims, U


def tabulate_states(sys, states, host_tab, perturb_grid, t_k, entry_off, g_off, Upad, g_per_w,
                    flush_fn, max_doubles=8 << 20, align=1, valid=None):
    """Second pass, per-state mode: call dyn/cost once per state exactly like
    the reference's hot loop and stage the un-broadcast outputs.
    Raises GDependsOnW if a cost depends on w while `g_per_w` is 0.
    `valid` (bool per entry of `states`, layout CF): False marks a padding position
    that repeats a real state; its record gets U = 0 (no admissible control)."""
    d = len(sys.state)
    nb_control = len(sys.control)
    if nb_control > _cabi.SDP_MAX_C:
        raise NotImplementedError("more than %d control variables" % _cabi.SDP_MAX_C)
    w_args, w_shape, W = perturb_layout(perturb_grid)
    writer = _ChunkWriter(d, flush_fn, max_doubles, align)
    block = 1024
    for b0 in range(0, len(states), block):
        b1 = min(b0 + block, len(states))
        recs = np.zeros(b1 - b0, dtype=_cabi.STATE_DESC_DTYPE)
        recs["npts"] = 1
        for i in range(b0, b1):
            compact, control_dims, U = _eval_one_state(sys, states[i], host_tab, i, w_args, t_k, W, w_shape)
            if compact[-1][2] > 1 and not g_per_w:
                raise GDependsOnW()
            r = recs[i - b0]
            r["U"] = U
            r["npts"][:nb_control] = control_dims
            # C-order strides of the flattened control index, per axis
            tail = np.ones(nb_control + 1, dtype=np.int64)
            for c in range(nb_control - 1, -1, -1):
                tail[c] = tail[c + 1] * control_dims[c]
            for k, (arr, u_eff, w_eff) in enumerate(compact):
                r["src"][k] = writer.add_array(arr)
                r["ws"][k] = 1 if w_eff > 1 else 0
                if u_eff > 1:
                    r["cs"][k][:nb_control] = w_eff * tail[1:nb_control + 1]
        recs["entry_off"] = entry_off[b0:b1]
        recs["g_off"] = g_off[b0:b1]
        recs["Upad"] = Upad[b0:b1]
        if valid is not None:
            recs["U"] = np.where(valid[b0:b1], recs["U"], 0)
        writer.add_descs(recs)
    writer.flush()


def _strides_or_zero(shape):
    """element strides of a C-contiguous array of `shape`, 0 on size-1 axes"""
    st = np.ones(len(shape), dtype=np.int64)
    for k in range(len(shape) - 2, -1, -1):
        st[k] = st[k + 1] * shape[k + 1]
    return np.where(np.asarray(shape) == 1, 0, st)


def _eval_state_chunk(sys, state_cols, lo, hi, npts, w_args, t_k, W, grid_cache=None, w_shape=(),
                      only=None):
    """dyn/cost for a chunk of S states in ONE call: state variables enter as
    (S,1,..,1) arrays, control c as an (S,..,n_c_max,..,1) array whose row s is
    that state's own control grid (padded by repeating its last point), the
    perturbation as the reference's (W,) vector.  Element-wise numpy arithmetic
    gives the same values as the reference's per-state calls; this is verified
    on sample states by the caller before the mode is trusted."""
    S, nb_control = npts.shape
    d = len(state_cols)
    nmax = [int(npts[:, c].max()) for c in range(nb_control)]
    full_shape = (S,) + tuple(nmax) + (W,)
    rank = len(full_shape)
    n_w_axes = max(len(w_shape), 1)
    rank_in = 1 + nb_control + n_w_axes         # rank of the arguments (one axis per perturbation)
    xs = tuple(col.reshape((S,) + (1,) * (rank_in - 1)) for col in state_cols)
    # the padded control grids of the chunk; a time-dependent recursion whose boxes do
    # not change from one instant to the next (the common case) reuses them
    key = None
    us = None
    if grid_cache is not None:
        key = (lo.tobytes(), hi.tobytes(), npts.tobytes(), n_w_axes)
        us = grid_cache.get(key)
    if us is None:
        us = []
        for c in range(nb_control):
            j = np.arange(nmax[c])[None, :]
            n_c = npts[:, c][:, None]
            vals = control_axis_values(lo[:, c][:, None], hi[:, c][:, None], n_c, np.minimum(j, n_c - 1))
            vals = vals.reshape((S,) + (1,) * c + (nmax[c],) + (1,) * (nb_control - 1 - c + n_w_axes))
            vals.flags.writeable = False      # shared between calls: user code must not mutate it
            us.append(vals)
        if grid_cache is not None:
            if len(grid_cache) >= 64:
                grid_cache.clear()
            grid_cache[key] = us
    args = xs + tuple(us) + w_args
    if t_k is not None:
        args = (t_k,) + args
    # `only`: "cost" / "dyn" evaluate one of the two callables (time-dependent recursions whose
    # dynamics do not depend on the instant re-evaluate the cost alone)
    pairs = []
    if only != "cost":
        x_next = sys.dyn(*args, **sys.params)
        if len(x_next) != d:
            raise ValueError("dyn returned %d next-state components, expected %d" % (len(x_next), d))
        pairs += [("dyn", c) for c in x_next]
    if only != "dyn":
        pairs.append(("cost", sys.cost(*args, **sys.params)))
    outs = []
    for what, a in pairs:
        a = np.asarray(a)
        if a.dtype != np.float64:
            a = a.astype(float)
        a = _fold_w(a, 1 + nb_control, w_shape, what)
        if a.ndim > rank:
            raise ValueError("%s output of rank %d does not broadcast to %s" % (what, a.ndim, full_shape))
        a = np.ascontiguousarray(a.reshape((1,) * (rank - a.ndim) + a.shape))
        for n_a, n_f in zip(a.shape, full_shape):
            if n_a != 1 and n_a != n_f:
                raise ValueError("%s output shape %s does not broadcast to %s" % (what, a.shape, full_shape))
        outs.append(a)
    return outs, nmax


def tabulate_states_batched(sys, state_grid, begin, end, host_tab, perturb_grid, t_k, entry_off,
                            g_off, Upad, g_per_w, flush_fn, chunk_states=4096,
                            max_doubles=16 << 20, align=1, verify=8, grid_cache=None,
                            flat_index=None, valid=None, record=None, threads=1):
    """Second pass, batched mode: one dyn/cost call per chunk of states.
    `verify` sample states of EVERY chunk are re-evaluated per state, the
    reference's way, and compared bit-for-bit; a mismatch raises
    BatchedMismatch (the caller falls back to the per-state mode).
    By default the states are [begin, end) of the C-order grid; `flat_index` gives
    them explicitly instead (layout CF walks the grid column by column), with
    `host_tab`, `entry_off`, ... indexed by position in that list, and `valid`
    marking the real states (False: a padding position that repeats a real state and
    gets U = 0).
    `threads` > 1 (DPSolver.host_threads, opt-in): the chunks are evaluated - and checked - on that
    many host threads (numpy releases the interpreter lock inside its loops) while the calling
    thread stages and uploads them in order.  Same calls, same arguments, same tables; the
    reference calls dyn/cost on the calling thread only, so callables that are not thread-safe
    must keep the default of 1."""
    d = len(sys.state)
    nb_control = len(sys.control)
    if nb_control > _cabi.SDP_MAX_C:
        raise NotImplementedError("more than %d control variables" % _cabi.SDP_MAX_C)
    w_args, w_shape, W = perturb_layout(perturb_grid)
    dims = [len(g) for g in state_grid]
    writer = _ChunkWriter(d, flush_fn, max_doubles, align)
    n = end - begin if flat_index is None else len(flat_index)
    chunk_states = max(align, chunk_states // align * align)

    def evaluate(b0):
        """the user's callables on one chunk (and its check): independent of every other chunk"""
        b1 = min(b0 + chunk_states, n)
        S = b1 - b0
        flat = np.arange(begin + b0, begin + b1) if flat_index is None else flat_index[b0:b1]
        idx = np.unravel_index(flat, dims)
        cols = [np.asarray(state_grid[k])[idx[k]] for k in range(d)]
        npts = host_tab.npts[b0:b1]
        outs, nmax = _eval_state_chunk(sys, cols, host_tab.lo[b0:b1], host_tab.hi[b0:b1], npts,
                                       w_args, t_k, W, grid_cache, w_shape)
        if outs[-1].shape[-1] > 1 and not g_per_w:
            raise GDependsOnW()
        if verify:
            # every chunk is another region of the state space, where the branches of a user's
            # np.where / clipping may differ: each is checked on its own sample states
            _verify_chunk(sys, cols, host_tab, b0, outs, w_args, t_k, W, min(verify, S),
                          None if valid is None else valid[b0:b1], w_shape)
        return b1, cols, npts, outs

    for b0, (b1, cols, npts, outs) in _in_order(evaluate, range(0, n, chunk_states), threads):
        S = b1 - b0
        recs = np.zeros(S, dtype=_cabi.STATE_DESC_DTYPE)
        recs["npts"] = 1
        recs["npts"][:, :nb_control] = npts
        recs["U"] = npts.prod(axis=1) if nb_control else 1
        if valid is not None:
            recs["U"] = np.where(valid[b0:b1], recs["U"], 0)
        recs["entry_off"] = entry_off[b0:b1]
        recs["g_off"] = g_off[b0:b1]
        recs["Upad"] = Upad[b0:b1]
        for k, a in enumerate(outs):
            st = _strides_or_zero(a.shape)
            base = writer.add_array(a)
            recs["src"][:, k] = base + np.arange(S) * st[0]
            recs["cs"][:, k, :nb_control] = st[1:1 + nb_control]
            recs["ws"][:, k] = st[-1]
        if record is not None:
            # what a recursion whose dynamics ignore the instant needs to re-tabulate the cost
            # alone (Engine.recursion_fast): the chunk's state columns and un-broadcast outputs,
            # and where the cost sits in the staging buffer
            record.append(dict(cols=cols, outs=outs, g_base=int(base), b0=b0, b1=b1))
        writer.add_descs(recs)
    writer.flush()


def _in_order(fn, keys, threads=1, ahead=2):
    """yield (key, fn(key)) in the order of `keys`; with threads > 1 up to ahead*threads calls
    run on worker threads ahead of the consumer.  An exception of a call surfaces, at its turn, in
    the consumer; calls not yet started are then dropped."""
    keys = list(keys)
    if threads <= 1 or len(keys) <= 1:
        for k in keys:
            yield k, fn(k)
        return
    from concurrent.futures import ThreadPoolExecutor
    window = max(2, ahead * threads)
    pool = ThreadPoolExecutor(max_workers=threads, thread_name_prefix="sdp-tabulate")
    pending = []
    try:
        nxt = 0
        for k in keys:
            while nxt < len(keys) and len(pending) < window:
                pending.append(pool.submit(fn, keys[nxt]))
                nxt += 1
            yield k, pending.pop(0).result()
    finally:
        for f in pending:
            f.cancel()
        pool.shutdown(wait=True)


class BatchedMismatch(Exception):
    """the user's callables do not give bit-identical results when evaluated on
    a chunk of states at once"""


def _verify_chunk(sys, cols, host_tab, b0, outs, w_args, t_k, W, n_check, valid=None, w_shape=()):
    """bit-for-bit comparison of the chunk evaluation against the reference's
    per-state evaluation on a few sample states, after full broadcast"""
    S = len(cols[0])
    nb_control = len(sys.control)
    picks = np.unique(np.linspace(0, S - 1, n_check).astype(int))
    if valid is not None:
        picks = picks[np.asarray(valid)[picks]]      # (padding positions repeat a real state)
    for s in picks:
        x_k = tuple(col[s] for col in cols)
        compact, control_dims, U = _eval_one_state(sys, x_k, host_tab, b0 + s, w_args, t_k, W, w_shape)
        full = tuple(control_dims) + (W,)
        for a, (ref, u_eff, w_eff) in zip(outs, compact):
            row = a[s if a.shape[0] > 1 else 0]                     # (n1.., Weff)
            sl = tuple(slice(0, control_dims[c] if row.shape[c] > 1 else 1)
                       for c in range(nb_control))
            got = np.ascontiguousarray(np.broadcast_to(row[sl], full))
            shape = (tuple(control_dims) if u_eff > 1 else (1,) * nb_control) + (w_eff,)
            want = np.ascontiguousarray(np.broadcast_to(ref.reshape(shape), full))
            if not np.array_equal(got.view(np.int64), want.view(np.int64)):
                raise BatchedMismatch()


def state_tuples_at(state_grid, begin, end):
    """the same tuples as `state_tuples`, built from the unravelled indices (no walk from
    the start of the grid)"""
    dims = [len(g) for g in state_grid]
    idx = np.unravel_index(np.arange(begin, end), dims)
    cols = [np.asarray(g)[i] for g, i in zip(state_grid, idx)]
    return list(zip(*cols))


def state_tuples(state_grid, begin, end):
    """states [begin, end) of the C-order flattened grid as tuples of numpy
    scalars, the same objects `itertools.product(*self.state_grid)` yields in the
    reference (stodynprog.py:478)."""
    it = itertools.product(*state_grid)
    return list(itertools.islice(it, begin, end))
