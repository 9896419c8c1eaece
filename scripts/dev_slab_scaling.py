"""developer timing: how the streaming kernel scales when the state slab shrinks
(what one rank of an N-GPU run sees): config #5 with n_E = 2000/N states rows.
Prints kernel ms per sweep and backups/s for the default (factored) and dense tables."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from dev_timing import time_sweeps  # noqa: E402

for compress in ("auto", "off"):
    for n_E in (2000, 1000, 500, 250):
        for chunk in ((None, 512) if n_E <= 500 else (None,)):
            prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=500, item_chunk=chunk)
            sv = prob.solver
            sv.table_compress = compress
            T = sv.sweep_tables()
            r = time_sweeps(sv, T, n=20, warm=5)
            print("compress=%s n_E=%d layout=%s chunk=%d items=%d  %s" % (
                compress, n_E, T.layout_name, T.item_chunk, T.n_items, r), flush=True)
            del sv, T, prob
            torch.cuda.empty_cache()
