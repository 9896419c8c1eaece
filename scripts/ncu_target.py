"""target for ncu captures: build one workload and run a few sweeps.
    python scripts/ncu_target.py {ar1|large|searev} {on|off} [n_sweeps]
env: N_E (large), SEAREV_N_E, LAYOUT, COLUMN (on|off|auto: layout CF)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from dev_factored import make  # noqa: E402

which, compress = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
sv = make(which, compress, os.environ.get("LAYOUT", "state_minor" if which == "large" else "auto"))
sv.column_hoist = os.environ.get("COLUMN", "auto")
T = sv.sweep_tables()
eng = sv.engine
J_prev = eng.to_device(np.random.default_rng(0).standard_normal(int(np.prod(sv._state_grid_shape))))
J_new = torch.empty_like(J_prev)
for _ in range(n):
    eng.sweep(T, J_prev, J_new)
    J_prev, J_new = J_new, J_prev
torch.cuda.synchronize()
print(which, compress, T.layout_name, "done")
