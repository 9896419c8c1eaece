"""target for ncu captures: build one workload and run a few sweeps.
    python scripts/ncu_target.py {ar1|large|searev} {on|off} [n_sweeps]
env: N_E (large), SEAREV_N_E, LAYOUT, COLUMN (on|off|auto: layout CF), BANDS (row bands of
layout CF on one rank: auto|1|3..), COL_SHARD="i/N" (only the columns of shard i of an N-way
cut by columns: what one rank of an N-GPU run sweeps)"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from dev_factored import make  # noqa: E402
from stodynprog_b200 import _cabi  # noqa: E402
from stodynprog_b200.engine import Engine, partition_by_weight  # noqa: E402

which, compress = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
if os.environ.get("BANDS"):
    Engine.COLUMN_BANDS = os.environ["BANDS"]
sv = make(which, compress, os.environ.get("LAYOUT", "state_minor" if which == "large" else "auto"))
sv.column_hoist = os.environ.get("COLUMN", "auto")
eng = sv.engine
shard = os.environ.get("COL_SHARD")
if shard:
    i, N = (int(x) for x in shard.split("/"))
    Engine.COLUMN_BANDS = "1"
    T = eng.build_sweep_tables(sv)                      # (the scan; gives the control counts)
    U = T.host_full.U.astype(np.int64)
    rows, cols = sv._state_grid_shape
    cb = partition_by_weight((U + 1).reshape(rows, cols).sum(axis=0), N)
    del T
    sv._col_override = (int(cb[i]), int(cb[i + 1]))
    T = eng.build_sweep_tables(sv)
    J = eng.to_device(np.random.default_rng(0).standard_normal(rows * cols))
    for _ in range(n):                                 # the streaming pass of the shard (pre-pass + sweep)
        rc = eng.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables), eng._ptr(J),
                                        eng._ptr(T.part_val), eng._ptr(T.part_idx), eng.stream)
        _cabi.check(rc, "sdp_sweep_partials")
    torch.cuda.synchronize()
    print(which, "column shard", shard, T.layout_name, T.n_states, "states;", _cabi.last_kernel())
    sys.exit(0)
T = sv.sweep_tables()
J_prev = eng.to_device(np.random.default_rng(0).standard_normal(int(np.prod(sv._state_grid_shape))))
J_new = torch.empty_like(J_prev)
for _ in range(n):
    eng.sweep(T, J_prev, J_new)
    J_prev, J_new = J_new, J_prev
torch.cuda.synchronize()
print(which, compress, T.layout_name, T.n_states, "states;", _cabi.last_kernel(), "done")
