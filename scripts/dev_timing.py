"""developer timing sweep (not part of the bench contract): time K1 on a
workload for several item chunk sizes / controls-per-lane settings."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stodynprog_b200 as sdp  # noqa: E402
from stodynprog_b200 import workloads as wl  # noqa: E402


def time_sweeps(prob, n=10, warm=3):
    sv = prob.solver
    t0 = time.perf_counter()
    T = sv.sweep_tables()
    t_setup = time.perf_counter() - t0
    eng = sv.engine
    rng = np.random.default_rng(0)
    J_prev = eng.to_device(rng.standard_normal(int(np.prod(sv._state_grid_shape))))
    J_new = torch.empty_like(J_prev)
    for _ in range(warm):
        eng.sweep(T, J_prev, J_new)
        J_prev, J_new = J_new, J_prev
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(n):
        eng.sweep(T, J_prev, J_new)
        J_prev, J_new = J_new, J_prev
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(n)])
    gb = T.n_backups_local * T.algorithmic_bytes_per_backup / 1e9
    best, med = ms.min(), np.median(ms)
    return dict(setup_s=round(t_setup, 2), items=T.n_items, backups=T.n_backups_local,
                ms_best=round(float(best), 4), ms_med=round(float(med), 4),
                gbackups_s=round(T.n_backups_local / med / 1e6, 2),
                alg_GBs=round(gb / (med / 1e3), 1), table_GB=round(T.device_bytes / 1e9, 2))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ar1"
    chunks = [int(c) for c in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["512"])]
    print("SDP_UPL =", os.environ.get("SDP_UPL", "4"))
    for chunk in chunks:
        if which == "ar1":
            prob = wl.storage_ar1(sdp, item_chunk=chunk)
        elif which == "large":
            n_E = int(os.environ.get("N_E", "500"))
            prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=500, item_chunk=chunk)
        elif which == "searev":
            prob = wl.searev(sdp, n_E=int(os.environ.get("N_E", "5")), item_chunk=chunk)
        print(which, "chunk", chunk, time_sweeps(prob), flush=True)
