"""layout CF on config #5: work-item length x row bands x item hand-out (one-GPU tuning run)"""
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
from stodynprog_b200 import _cabi  # noqa: E402
from stodynprog_b200.engine import Engine  # noqa: E402

n_E, n_P = 2000, 500
prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=n_P)
sv = prob.solver
sv.column_hoist = "on"
eng = sv.engine
lib = eng.lib
J0_host = np.random.default_rng(0).standard_normal((n_E, n_P))
J0 = eng.to_device(J0_host.reshape(-1))
_cabi.check(lib.sdp_set_option(b"col_threads", 640), "opt")


def timed(T, n=8, warm=2):
    a, b = J0.clone(), torch.empty_like(J0)
    for _ in range(warm):
        eng.sweep(T, a, b)
        a, b = b, a
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for k in range(n):
        eng.sweep(T, a, b)
        a, b = b, a
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n, float(a.sum().item())


def e2e(n=8):
    J_h = J0_host
    for _ in range(2):
        J_h, pol_h = sv.value_iteration(J_h, report_time=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        J_h, pol_h = sv.value_iteration(J_h, report_time=False)
    return 1e3 * (time.perf_counter() - t0) / n, float(J_h.sum())


for bands in ("1", "3", "auto"):
    for chunk in (64, 32, 16):
        Engine.COLUMN_BANDS = bands
        eng.item_chunk_auto, eng.item_chunk = False, chunk
        sv._table_cache = {}
        T = sv.sweep_tables()
        for dyn in (0, 1):
            _cabi.check(lib.sdp_set_option(b"col_dynamic", dyn), "opt")
            ms, chk = timed(T)
            ems, echk = e2e()
            print("bands=%-4s chunk=%-3d items=%-7d dynamic=%d: %.3f ms/sweep = %4.0f G/s | e2e %.3f ms = %4.0f G/s"
                  " | checksums %.10e %.10e" % (bands, chunk, T.n_items, dyn, ms, T.n_backups_local / ms / 1e6, ems,
                                               T.n_backups_total / ems / 1e6, chk, echk), flush=True)
        del T
