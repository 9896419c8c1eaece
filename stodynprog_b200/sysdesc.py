"""Problem description: the host-side holder of the user's callables.

Mirror of the reference's `SysDescription` (stodynprog/stodynprog.py:56-247) and
its helpers `_zero_cost` (:19-21) and `_enforce_sig_len` (:24-54).  It is pure
boundary: no device work happens here.  Same constructor, same attribute names
(including the reference's spelling `stationnary`), same error types and error
texts (one of them is pinned by the reference's test_stodynprog.py:58).

Differences, both forced by the interpreter and invisible to callers:
  * signatures are read with `inspect.getfullargspec` (`getargspec`, used by the
    reference at :32-33,120,174, was removed in Python 3.11);
  * nothing imports matplotlib.
"""
import inspect

__all__ = ["SysDescription"]


def _zero_cost(*x):
    """g(x) = 0 whatever the arguments; default terminal cost
    (reference stodynprog.py:19-21)."""
    return 0.


def _signature(fun):
    """(positional argument names, name of the **kwargs catcher or None)."""
    spec = inspect.getfullargspec(fun)
    return list(spec.args), spec.varkw


def _enforce_sig_len(fun, args, with_params, shortname=None):
    """Check that `fun` takes exactly len(args) positional arguments and takes
    `**kwargs` iff the system has parameters.  Raises ValueError otherwise,
    returns True on success (reference stodynprog.py:24-54; message formats
    :43-51)."""
    fun_args, kw_name = _signature(fun)
    prefix = (shortname if shortname is not None else '') + "'{:s}' ".format(fun.__name__)

    if len(fun_args) != len(args):
        raise ValueError(prefix + 'should accept {:d} args ({:s}), not {:d}'.format(
            len(args), ', '.join(args), len(fun_args)))
    if with_params and kw_name is None:
        raise ValueError(prefix + 'should accept extra keyword arguments')
    if not with_params and kw_name is not None:
        raise ValueError(prefix + 'should not accept extra keyword arguments')
    return True


class SysDescription(object):
    """Dynamical system seen by dynamic programming:

      * dynamics   x_{k+1} = f_k(x_k, u_k, w_k)        -> `dyn`
      * stage cost g_k(x_k, u_k, w_k)                   -> `cost`
      * admissible controls, a box U_k(x_k)             -> `control_box`
      * perturbation laws (scipy.stats frozen laws)     -> `perturb_laws`

    `dims` is (n_state, n_control[, n_perturb]).  For a time-dependent system
    (`stationnary=False`) every callable takes the instant `k` first.
    `params` is a dict splatted as keyword arguments into every callable.
    (reference stodynprog.py:56-99)
    """

    def __init__(self, dims, stationnary=True, name='', params=None):
        self.name = name
        self.stationnary = bool(stationnary)
        self.params = params if params is not None else {}

        if len(dims) == 3:
            n_state, n_control, n_perturb = dims
        elif len(dims) == 2:
            n_state, n_control = dims
            n_perturb = 0
        else:
            raise ValueError('dims tuple should be of len 2 or 3')

        # placeholder names, replaced by the names read from `dyn`'s signature
        self.state = ['x{:d}'.format(i + 1) for i in range(n_state)]
        self.control = ['u{:d}'.format(i + 1) for i in range(n_control)]
        self.perturb = ['w{:d}'.format(i + 1) for i in range(n_perturb)]

        self._dyn_args = self.state + self.control + self.perturb
        if not self.stationnary:
            self._dyn_args.insert(0, 'time_k')

        self._dyn = None
        self._cost = None
        self._control_box = None
        self._terminal_cost = _zero_cost
        self._perturb_laws = None

    # -- properties -------------------------------------------------------
    @property
    def stochastic(self):
        """True when the system has at least one perturbation."""
        return len(self.perturb) > 0

    @property
    def dyn(self):
        """dynamics function x_{k+1} = f_k(x_k, u_k, w_k)"""
        return self._dyn

    @dyn.setter
    def dyn(self, dyn):
        # arity check, then adopt the variable names of the signature
        # (reference stodynprog.py:111-131)
        if _enforce_sig_len(dyn, self._dyn_args, bool(self.params), 'dynamics function'):
            self._dyn = dyn
        names, _ = _signature(dyn)
        self._dyn_args = names
        if not self.stationnary:
            names = names[1:]
        ns, nc, nw = len(self.state), len(self.control), len(self.perturb)
        self.state = names[0:ns]
        self.control = names[ns:ns + nc]
        self.perturb = names[ns + nc:ns + nc + nw]

    @property
    def control_box(self):
        """admissible controls U_k(x_k) as a box [u1_min,u1_max] x [u2_min,u2_max] x ...
        (reference stodynprog.py:133-150)"""
        return self._control_box

    @control_box.setter
    def control_box(self, control_box):
        args = list(self.state)
        if not self.stationnary:
            args.insert(0, 'time_k')
        if _enforce_sig_len(control_box, args, bool(self.params),
                            'control description function'):
            self._control_box = control_box

    @property
    def cost(self):
        """stage cost g_k(x_k, u_k, w_k)"""
        return self._cost

    @cost.setter
    def cost(self, cost):
        if _enforce_sig_len(cost, self._dyn_args, bool(self.params), 'cost function'):
            self._cost = cost

    @property
    def terminal_cost(self):
        """terminal cost g(x_K) (stored only; like the reference, no solver
        entry point consumes it - stodynprog.py:165-179)"""
        return self._terminal_cost

    @terminal_cost.setter
    def terminal_cost(self, cost):
        cost_args, _ = _signature(cost)
        if len(cost_args) != len(self.state):
            raise ValueError('cost function should accept '
                             '{:d} args instead of {:d}'.format(
                                 len(self.state), len(cost_args)))
        self._terminal_cost = cost

    @property
    def perturb_laws(self):
        """distribution laws of the perturbations `w_k`"""
        return self._perturb_laws

    @perturb_laws.setter
    def perturb_laws(self, laws):
        # one law per perturbation; a law with a `pdf` is continuous, with a
        # `pmf` discrete (reference stodynprog.py:186-207)
        if len(laws) != len(self.perturb):
            raise ValueError('{:d} perturbation laws should be provided'
                             .format(len(self.perturb)))
        self._perturb_laws = laws
        self.perturb_types = []
        for law in laws:
            kind = None
            try:
                law.pdf(0)
                kind = 'continuous'
            except AttributeError:
                try:
                    law.pmf(0)
                    kind = 'discrete'
                except AttributeError:
                    raise ValueError('perturbation law {:s} should either have '
                                     'a pdf or a pmf method'.format(repr(law)))
            self.perturb_types.append(kind)

    # -- reporting --------------------------------------------------------
    def print_summary(self):
        """text summary, same layout as the reference (stodynprog.py:209-242)"""
        print('Dynamical system "{}" description'.format(self.name))
        print('* behavioral properties: {}, {}'.format(
            'stationnary' if self.stationnary else 'time dependent',
            'stochastic' if self.stochastic else 'deterministic'))

        print('* functions:')
        functions = [('dynamics', self.dyn), ('cost', self.cost),
                     ('control box', self.control_box)]
        width = max(len(label) for label, _ in functions) + 1
        for label, fun in functions:
            where = ('{0.__module__}.{0.__name__}'.format(fun) if fun is not None
                     else 'None (to be defined)')
            print('  - {0:{width}}: {1}'.format(label, where, width=width))

        print('* variables')
        vectors = [('state', self.state), ('control', self.control)]
        if self.stochastic:
            vectors.append(('perturbation', self.perturb))
        width = max(len(label) for label, _ in vectors) + 1
        for label, names in vectors:
            print('  - {0:{width}}: {1} (dim {2:d})'.format(
                label, ', '.join(names), len(names), width=width))

    def __repr__(self):
        return '<SysDescription "{:s}" at 0x{:x}>'.format(self.name, id(self))
