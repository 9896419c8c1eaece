"""developer timing: finite-horizon recursion of config #2 (PV storage, T=240, 50 states,
1001..2001 controls, deterministic) - ours vs the oracle port on the host CPU; and the
SEAREV policy iteration (config #4) end to end."""
import contextlib
import io
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from oracle.ref_port import port_api  # noqa: E402


def timed(fn):
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        out = fn()
    return time.perf_counter() - t0, out


prob = wl.pv_storage(sdp)
T = prob.horizon
timed(lambda: prob.solver.bellman_recursion(8, prob.J_fin))          # warm-up (library load, allocator)
t_ours, (J, pol) = timed(lambda: prob.solver.bellman_recursion(T, prob.J_fin))
ora = wl.pv_storage(port_api("c"))
t_port, (Jo, polo) = timed(lambda: ora.solver.bellman_recursion(T, ora.J_fin))
print("pv_storage bellman_recursion T=%d: ours %.3f s (%.2f ms/instant), port on CPU %.3f s (%.2f ms/instant); "
      "policies equal: %s, max |dJ| %.2e" % (T, t_ours, 1e3 * t_ours / T, t_port, 1e3 * t_port / T,
                                            np.array_equal(pol, polo), np.max(np.abs(J - Jo))))

if "--searev" in sys.argv:
    prob = wl.searev(sdp)
    sv = prob.solver
    t_setup, Tb = timed(lambda: sv.sweep_tables())
    t_pi, (Jp, polp) = timed(lambda: sv.policy_iteration(prob.initial_policy(), 1000, 5, rel_dp=True))
    print("searev 31x61x61: table setup %.2f s, policy_iteration(n_val=1000, n_pol=5) %.2f s "
          "(reference: 785 s, SURVEY.md 8a)" % (t_setup, t_pi))
