"""developer timing: where the end-to-end time of DPSolver.value_iteration(J_host)
goes on config #5 (host phases timed with a device sync after each)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402

prob = wl.storage_ar1_large(sdp)
sv = prob.solver
eng = sv.engine
T = sv.sweep_tables()
dims = sv._state_grid_shape
n_grid = int(np.prod(dims))
J_h = np.random.default_rng(0).standard_normal(dims)
for _ in range(3):
    J_h, pol_h = sv.value_iteration(J_h, report_time=False)


def sync():
    torch.cuda.synchronize()


acc = {}


def tick(name, t0):
    sync()
    t1 = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t1 - t0)
    return t1


R = 20
for _ in range(R):
    sync()
    t = time.perf_counter()
    key = sv._cache_key(None)
    t = tick("cache_key", t)
    J_prev, J_new = eng.J_pair(n_grid)
    eng.begin_call(n_grid)
    t = tick("J_pair", t)
    eng.upload_J(J_h, J_prev)
    t = tick("upload_J (stage + H2D 8 MB)", t)
    eng.sweep(T, J_prev, J_new)
    t = tick("sweep", t)
    pol_dev = eng.policy_values(T, eng.gather_argmin(T))
    t = tick("policy_values (K3)", t)
    outs = eng.to_host(J_new, pol_dev)
    t = tick("to_host (D2H 24 MB)", t)
    J_k = outs[0].reshape(dims)
    pol_k = outs[1].reshape(dims + (2,))
    t = tick("reshape", t)
tot = 0.0
for k, v in acc.items():
    print("%-34s %8.3f ms" % (k, 1e3 * v / R))
    tot += v
print("%-34s %8.3f ms" % ("sum of phases", 1e3 * tot / R))
sync()
t0 = time.perf_counter()
for _ in range(R):
    J_h, pol_h = sv.value_iteration(J_h, report_time=False)
sync()
print("%-34s %8.3f ms" % ("value_iteration (whole call)", 1e3 * (time.perf_counter() - t0) / R))
