"""The reference's example problems, restated as factories.

Each factory takes `api` - any object exposing `SysDescription` and `DPSolver`
(this package, the unmodified reference, or the oracle port) - and returns a
configured solver plus what a driver needs (initial policy, horizon, ...).
Using one definition for all three keeps the parity tests honest: the same
callables, grids and steps go to every implementation.

Config numbers follow BASELINE.json / SURVEY.md §8:
  #1 inventory control          doc/example_inventory.py:28-91
  #2 deterministic PV storage   examples/01 .../pv_storage_control.py:33-106
  #3 storage + AR(1)            examples/howto storage-AR1.ipynb (cells 2-8)
  #4 SEAREV + storage           examples/20 .../storage_control.py:36-134, searev_data.py:16-22,70-81
  #5 storage + AR(1), 2000x500 states x <=256 controls x 9 nodes (synthetic)
"""
import numpy as np
import scipy.stats as stats

__all__ = ["inventory", "pv_storage", "pv_production", "storage_ar1", "storage_ar1_large",
           "searev", "backups_per_sweep"]


class Problem(object):
    """what a factory returns"""

    def __init__(self, name, sys, solver, **extra):
        self.name = name
        self.sys = sys
        self.solver = solver
        self.__dict__.update(extra)


# ---------------------------------------------------------------------------
# #1 shop inventory (1 state, 1 control, 1 discrete perturbation)
# ---------------------------------------------------------------------------
def inventory(api, **solver_kw):
    h, p, c = 0.5, 3, 1          # holding, shortage, ordering unit costs
    demand = stats.rv_discrete(values=([0, 1, 2, 3], [0.2, 0.4, 0.3, 0.1])).freeze()

    def dyn_inv(x, u, w):
        return (x + u - w,)

    def admissible_orders(x):
        return ((0, 10),)

    def op_cost(x, u, w):
        return np.where(x > 0, x * h, -x * p) + u * c

    sys = api.SysDescription((1, 1, 1), name='Shop Inventory')
    sys.dyn = dyn_inv
    sys.perturb_laws = [demand]
    sys.control_box = admissible_orders
    sys.cost = op_cost
    solver = api.DPSolver(sys, **solver_kw)
    solver.discretize_state(-3, 6, 10)
    solver.discretize_perturb(0, 3, 4)
    solver.control_steps = (1,)
    return Problem('inventory', sys, solver, J0=np.zeros(10))


# ---------------------------------------------------------------------------
# #2 deterministic storage smoothing a PV production (time-dependent)
# ---------------------------------------------------------------------------
def pv_production(n_days=10, phi=0.8, seed=0):
    """hourly PV production: clear-sky half-sine times an AR(1) 'cloud' factor
    mapped to [0,1], rounded to 4 decimals - the recipe of
    examples/01 .../pv_prod_generator.py:19-52 (which wrote pv_prod.csv)."""
    from scipy.signal import lfilter
    n = 24 * n_days
    t = np.arange(n) * 1.
    sine = -np.cos(2 * np.pi * t / 24)
    clear = np.where(sine < 0, 0, sine)
    rs = np.random.RandomState(seed)
    ar = lfilter([1], [1, -phi], rs.normal(size=n))
    cloud = stats.norm.cdf(ar, scale=1 / np.sqrt(1 - phi ** 2))
    return t, np.round(clear * cloud, 4)


def pv_storage(api, n_E=50, horizon=None, **solver_kw):
    t, P_prod_data = pv_production()
    dt = t[1] - t[0]
    E_rated, P_rated, a = 2, 1, 0.0
    if horizon is None:
        horizon = len(P_prod_data)

    def dyn_sto(k, E_sto, P_sto):
        return (E_sto + (P_sto - a * abs(P_sto)) * dt,)

    def admissible_controls(k, E_sto):
        P_neg = np.max((-E_sto / (1 + a) / dt, -P_rated))
        P_pos = np.min(((E_rated - E_sto) / (1 - a) / dt, P_rated))
        return ((P_neg, P_pos),)

    def cost_model(k, E_sto, P_sto):
        P_grid = P_prod_data[k] - P_sto
        over = np.where(P_grid > 0.4, P_grid - 0.4, 0)
        neg = np.where(P_grid < 0, P_grid, 0)
        return over ** 2 + neg ** 2 + 0 * P_sto ** 2

    sys = api.SysDescription((1, 1, 0), name='Deterministic Storage for PV', stationnary=False)
    sys.dyn = dyn_sto
    sys.control_box = admissible_controls
    sys.cost = cost_model
    solver = api.DPSolver(sys, **solver_kw)
    solver.discretize_state(0, E_rated, n_E)
    solver.control_steps = (.001,)
    return Problem('pv_storage', sys, solver, horizon=horizon, J_fin=np.zeros(n_E),
                   P_prod_data=P_prod_data)


# ---------------------------------------------------------------------------
# #3 / #5 energy storage absorbing an AR(1) mismatch (2 states, 2 controls, 1 perturbation)
# ---------------------------------------------------------------------------
def storage_ar1(api, n_E=41, n_P=61, n_w=9, steps=(0.001, 0.1), P_rated=4., **solver_kw):
    dt = 1.
    p_scale, p_corr = 1., 0.8
    E_rated = 10.
    innov_scale = p_scale * np.sqrt(1 - p_corr)      # sic: the notebook's formula
    innov_law = stats.norm(loc=0, scale=innov_scale)
    P_tol_reduced = 0.9 * 0.5

    def dyn_sto(E, P_mis, P_sto, P_cur, innov):
        return (E + P_sto * dt, p_corr * P_mis + innov)

    def admissible_controls(E, P_mis):
        P_neg = np.max((-E / dt, -P_rated))
        P_pos = np.min(((E_rated - E) / dt, +P_rated))
        return ((P_neg, P_pos), (0, 0))           # curtailment disabled

    def cost_thres_quad(E, P_mis, P_sto, P_cur, innov):
        P_dev = P_mis - P_cur - P_sto
        above = (P_dev - P_tol_reduced) ** 2
        under = (P_dev + P_tol_reduced) ** 2
        cost = np.where(P_dev > P_tol_reduced, above, 0. * P_dev)
        return np.where(P_dev < -P_tol_reduced, under, cost)

    def P_sto_empirical(E, P_mis):
        P_neg = np.max((-E / dt, -P_rated))
        P_pos = np.min(((E_rated - E) / dt, +P_rated))
        return P_neg if P_mis < P_neg else (P_pos if P_mis > P_pos else P_mis)

    sys = api.SysDescription((2, 2, 1), name='Storage + AR(1)')
    sys.dyn = dyn_sto
    sys.control_box = admissible_controls
    sys.cost = cost_thres_quad
    sys.perturb_laws = [innov_law]
    solver = api.DPSolver(sys, **solver_kw)
    solver.discretize_state(0, E_rated, n_E, -4 * p_scale, 4 * p_scale, n_P)
    solver.discretize_perturb(-4 * innov_scale, 4 * innov_scale, n_w)
    solver.control_steps = tuple(steps)

    def initial_policy():
        E_g, P_g = solver.state_grid_full
        pol = np.zeros(solver._state_grid_shape + (2,))
        pol[..., 0] = np.vectorize(P_sto_empirical)(E_g, P_g)
        return pol

    return Problem('storage_ar1_%dx%d' % (n_E, n_P), sys, solver, initial_policy=initial_policy,
                   J0=np.zeros((n_E, n_P)))


def storage_ar1_large(api, n_E=2000, n_P=500, **solver_kw):
    """config #5: the storage-AR1 callables on a 2000 x 500 grid, control step
    8/255 (129..256 controls per state), 9 perturbation nodes."""
    prob = storage_ar1(api, n_E=n_E, n_P=n_P, n_w=9, steps=(8. / 255, 0.1), P_rated=4., **solver_kw)
    prob.name = 'storage_ar1_large_%dx%d' % (n_E, n_P)
    return prob


# ---------------------------------------------------------------------------
# #4 SEAREV wave-energy converter + storage (3 states, 1 control, 1 perturbation)
# ---------------------------------------------------------------------------
def searev(api, n_E=31, n_S=61, n_A=61, **solver_kw):
    damp, torque_max, power_max = 4.e6, 2e6, 1.1     # PTO: N/(rad/s), N.m, MW
    dt = 0.1
    c1, c2, innov_std = 1.9799, -0.9879, 0.00347     # AR(2) speed model at 0.1 s
    E_rated, P_rated, a = 10, 1.1, 0.00

    def searev_power(speed):
        tor = speed * damp
        tor = np.where(tor > torque_max, torque_max, tor)
        tor = np.where(tor < -torque_max, -torque_max, tor)
        P_prod = tor * speed / 1e6
        return np.where(P_prod > power_max, power_max, P_prod)

    def dyn_searev_sto(E_sto, Speed, Accel, P_sto, innov):
        E_sto_n = E_sto + (P_sto - a * abs(P_sto)) * dt
        Speed_n = (c1 + c2) * Speed - dt * c2 * Accel + innov
        Accel_n = (c1 + c2 - 1) / dt * Speed - c2 * Accel + innov / dt
        return (E_sto_n, Speed_n, Accel_n)

    def admissible_controls(E_sto, Speed, Accel):
        P_neg = np.max((-E_sto / (1 + a) / dt, -P_rated))
        P_pos = np.min(((E_rated - E_sto) / (1 - a) / dt, P_rated))
        return ((P_neg, P_pos),)

    def cost_model(E_sto, Speed, Accel, P_sto, innov):
        P_grid = searev_power(Speed) - P_sto
        return (P_grid / power_max) ** 2

    sys = api.SysDescription((3, 1, 1), name='Searev + Storage')
    sys.dyn = dyn_searev_sto
    sys.control_box = admissible_controls
    sys.cost = cost_model
    sys.perturb_laws = [stats.norm(loc=0, scale=innov_std)]
    solver = api.DPSolver(sys, **solver_kw)
    solver.discretize_state(0, E_rated, n_E, -4 * .254, 4 * 0.254, n_S, -4 * .227, 4 * .227, n_A)
    solver.discretize_perturb(-3 * innov_std, 3 * innov_std, 9)
    solver.control_steps = (.001,)

    def initial_policy():
        E_g, S_g, A_g = solver.state_grid_full
        pol_lin = searev_power(S_g) - P_rated * E_g / E_rated
        return pol_lin[..., np.newaxis]

    return Problem('searev_%dx%dx%d' % (n_E, n_S, n_A), sys, solver, initial_policy=initial_policy,
                   J0=np.zeros((n_E, n_S, n_A)))


def backups_per_sweep(solver, t_k=None):
    """admissible (x,u,w) triples of one sweep: sum_x U(x) * W"""
    import itertools
    W = len(solver.perturb_grid[0]) if len(solver.perturb_grid) else 1
    total = 0
    for x_k in itertools.product(*solver.state_grid):
        _, dims = solver.control_grids(x_k, t_k)
        total += int(np.prod(dims))
    return total * W
