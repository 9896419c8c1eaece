"""developer experiment: streaming-kernel time (pre-pass + column sweep, layout CF) of every shard
of an N-way cut of config #5 on ONE GPU, the grid cut into whole rows of axis 0 or into whole
columns - what a rank of an N-GPU run executes per sweep, without the exchange.
    python scripts/dev_shard_emulation.py [N ...]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from stodynprog_b200 import _cabi  # noqa: E402
from stodynprog_b200.engine import Engine, partition_by_weight  # noqa: E402

Ns = [int(a) for a in sys.argv[1:] if int(a) > 1] if len(sys.argv) > 1 else [8]
# OPTS="col_dynamic=1,col_threads=768": library options; CHUNK=16: controls per work item
for kv in filter(None, os.environ.get("OPTS", "").split(",")):
    k, v = kv.split("=")
    _cabi.check(_cabi.load_library().sdp_set_option(k.encode(), int(v)), "sdp_set_option")
AXES = os.environ.get("AXES", "rows,columns").split(",")
Engine.COLUMN_BANDS = os.environ.get("BANDS", "1")    # (a rank of a multi-GPU run sweeps one band)

prob = wl.storage_ar1_large(sdp)
sv = prob.solver
if os.environ.get("CHUNK"):
    sv._item_chunk = int(os.environ["CHUNK"])
eng = sv.engine
print("OPTS=%s CHUNK=%s" % (os.environ.get("OPTS", ""), os.environ.get("CHUNK", "auto")))
T = eng.build_sweep_tables(sv)                 # the scan happens here, once
U_all = T.host_full.U.astype(np.int64)
n_rows, n_cols = sv._state_grid_shape
n_grid = len(U_all)
J = torch.from_numpy(np.random.default_rng(0).standard_normal(n_grid)).to(eng.device)


def time_partials(T, reps=12):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record()
        rc = eng.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables), eng._ptr(J),
                                        eng._ptr(T.part_val), eng._ptr(T.part_idx), eng.stream)
        _cabi.check(rc, "sdp_sweep_partials")
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs[3:]]))


t1 = time_partials(T)
print("N=1 (bands %s): %.4f ms, layout %s, items %d, chunk %d" % (T.bands["rows"], t1, T.layout_name, T.n_items, T.item_chunk), flush=True)
del T
for N in Ns:
    row_w = (U_all + 1).reshape(n_rows, n_cols).sum(axis=1)
    rb = [int(b) * n_cols for b in partition_by_weight(row_w, N)]
    col_w = (U_all + 1).reshape(n_rows, n_cols).sum(axis=0)
    cb = [int(b) for b in partition_by_weight(col_w, N)]
    for axis, bounds in (("rows", rb), ("columns", cb)):
        if axis not in AXES:
            continue
        ts = []
        for r in range(N):
            sv._slab_override = sv._col_override = None
            if axis == "rows":
                sv._slab_override = (bounds[r], bounds[r + 1])
            else:
                sv._col_override = (bounds[r], bounds[r + 1])
            T = eng.build_sweep_tables(sv)
            assert T.column
            ts.append(time_partials(T))
            info = (T.n_items, T.item_chunk, T.n_segs)
            del T
        print("N=%d %-7s ms per shard: %s  max %.4f mean %.4f  (items %d chunk %d segs %d)  ideal %.4f  "
              "kernel scaling %.2fx" % (N, axis, " ".join("%.4f" % x for x in ts), max(ts), np.mean(ts),
                                        info[0], info[1], info[2], t1 / N, t1 / max(ts)), flush=True)
