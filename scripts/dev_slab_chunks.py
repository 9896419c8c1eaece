"""developer experiment: streaming-kernel time of each slab of an N-way cut of config #5 on ONE
GPU, for several work-item sizes (controls per warp).  Shows the wave quantisation of short
launches: a slab of an 8-GPU run is ~4 000 tiles = 2-3 waves of resident warps.
    python scripts/dev_slab_chunks.py [N] [compress]"""
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
from stodynprog_b200.engine import partition_by_weight  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
compress = sys.argv[2] if len(sys.argv) > 2 else "auto"
chunks = [int(c) for c in os.environ.get("CHUNKS", "512,256,128,64,32").split(",")]

prob = wl.storage_ar1_large(sdp)
sv = prob.solver
sv.table_compress = compress
eng = sv.engine
sv._slab_override = (0, 32)
T = eng.build_sweep_tables(sv)                 # the scan happens here, once
U_all = T.host_full.U.astype(np.int64)
bounds = partition_by_weight(U_all + 1, N)
n_grid = len(U_all)
J = torch.from_numpy(np.random.default_rng(0).standard_normal(n_grid)).to(eng.device)
print("slab bounds", bounds)
res = np.zeros((len(chunks), N))
for ci, chunk in enumerate(chunks):
    eng.item_chunk, eng.item_chunk_auto = chunk, False
    for r in range(N):
        sv._slab_override = (bounds[r], bounds[r + 1])
        T = eng.build_sweep_tables(sv)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(9)]
        for a, b in evs:
            a.record()
            rc = eng.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables), eng._ptr(J),
                                            eng._ptr(T.part_val), eng._ptr(T.part_idx), eng.stream)
            _cabi.check(rc, "sdp_sweep_partials")
            b.record()
        torch.cuda.synchronize()
        res[ci, r] = np.median([a.elapsed_time(b) for a, b in evs[3:]])
        n_items = T.n_items
        del T
    print("chunk %4d items(last slab) %6d  ms per slab: %s  max %.4f mean %.4f" % (
        chunk, n_items, " ".join("%.4f" % x for x in res[ci]), res[ci].max(), res[ci].mean()), flush=True)
