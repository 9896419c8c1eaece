"""developer timing: the timeline of DPSolver.value_iteration(J_host) on config #5 (one rank):
where the runs' sweeps, combines and copies sit relative to the start of the call
(Engine._trace: sweep_to_host records an event after every step)."""
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
sv.host_threads = "auto"
eng = sv.engine
T = sv.sweep_tables()
dims = sv._state_grid_shape
J_h = np.random.default_rng(0).standard_normal(dims)
for _ in range(3):
    J_h, pol_h = sv.value_iteration(J_h, report_time=False)
torch.cuda.synchronize()
R = 20
t0 = time.perf_counter()
for _ in range(R):
    J_h, pol_h = sv.value_iteration(J_h, report_time=False)
print("value_iteration: %.3f ms per call; pieces %s; bands %s" % (1e3 * (time.perf_counter() - t0) / R, os.environ.get("SDP_COLUMN_PIECES", "default"),
                                                        T.bands["rows"] if T.column else None))
acc, host = {}, 0.0
main = torch.cuda.current_stream(eng.device)
for rep in range(R):
    torch.cuda.synchronize()
    eng._trace = []
    e0 = torch.cuda.Event(enable_timing=True)
    h0 = time.perf_counter()
    e0.record(main)
    J_h, pol_h = sv.value_iteration(J_h, report_time=False)
    host += time.perf_counter() - h0
    for name, e in eng._trace:
        acc[name] = acc.get(name, 0.0) + e0.elapsed_time(e)
eng._trace = None
print("ms after the start of the call (device clock), mean of %d calls:" % R)
for k, v in acc.items():
    print("  %-30s %8.3f" % (k, v / R))
print("  %-30s %8.3f" % ("host: call returned", 1e3 * host / R))
# the device-resident sweep of the same tables (what bench.py's `value` times)
n_grid = int(np.prod(dims))
J_a, J_b = eng.J_pair(n_grid)
eng.begin_call(n_grid)
eng.upload_J(J_h, J_a)
for _ in range(3):
    eng.sweep(T, J_a, J_b)
torch.cuda.synchronize()
ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ea.record(torch.cuda.current_stream(eng.device))
for _ in range(20):
    eng.sweep(T, J_a, J_b)
    J_a, J_b = J_b, J_a
eb.record(torch.cuda.current_stream(eng.device))
eb.synchronize()
print("device-resident sweep: %.4f ms" % (ea.elapsed_time(eb) / 20))
