"""developer timing sweep (not part of the bench contract): time K1 on a
workload for several launch-tuning settings, tables built once."""
import itertools
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


def time_sweeps(sv, T, n=10, warm=3):
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
    med = float(np.median(ms))
    return dict(ms_med=round(med, 4), gbackups_s=round(T.n_backups_local / med / 1e6, 1),
                alg_GBs=round(gb / (med / 1e3), 0), frac=round(gb / (med / 1e3) / 6548.8, 3),
                checksum=float(J_prev.sum().item()))


def opt(lib, **kw):
    for k, v in kw.items():
        _cabi.check(lib.sdp_set_option(k.encode(), int(v)), "sdp_set_option")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ar1"
    n_E = int(os.environ.get("N_E", "500"))
    layout = os.environ.get("LAYOUT", "auto")
    chunk = int(os.environ.get("CHUNK", "512"))
    if which == "ar1":
        prob = wl.storage_ar1(sdp, item_chunk=chunk)
    elif which == "large":
        prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=500, item_chunk=chunk)
    elif which == "searev":
        prob = wl.searev(sdp, n_E=int(os.environ.get("N_E", "5")), item_chunk=chunk)
    sv = prob.solver
    sv.table_layout = layout
    t0 = time.perf_counter()
    T = sv.sweep_tables()
    print(which, "layout", "B" if T.tiled else "A", "setup %.1fs" % (time.perf_counter() - t0),
          "items", T.n_items, "table GB %.2f" % (T.device_bytes / 1e9), flush=True)
    lib = sv.engine.lib
    if T.tiled:
        opt(lib, tma=0)
        for wb in (1, 3):
            opt(lib, wb=wb)
            print("ldg wb=%d" % wb, time_sweeps(sv, T), flush=True)
        opt(lib, tma=2)
        for rb in (2, 4, 8):
            opt(lib, rb=rb)
            print("pipe rb=%d" % rb, time_sweeps(sv, T), flush=True)
        if os.environ.get("TMA_SWEEP"):
            opt(lib, tma=1)
            for R, S, NW in itertools.product((4, 8), (2, 3), (1, 2, 4, 8)):
                opt(lib, tma_rows=R, tma_stages=S, tma_warps=NW)
                print("tma R=%d S=%d NW=%d" % (R, S, NW), time_sweeps(sv, T), flush=True)
    else:
        for upl in (4, 2):
            opt(lib, upl=upl)
            print("upl=%d" % upl, time_sweeps(sv, T), flush=True)
