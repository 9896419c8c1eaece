"""one-GPU check of layout CF (column-shared hoist) on config #5: bit-identity against
layout BF on the full grid, then timings over CTA size / controls per iteration / CTA
segments per SM.      python scripts/dev_column.py [n_E] [n_P]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stodynprog_b200 as sdp  # noqa: E402
from stodynprog_b200 import workloads as wl, _cabi  # noqa: E402

n_E = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_P = int(sys.argv[2]) if len(sys.argv) > 2 else 500
prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=n_P)
sv = prob.solver
eng = sv.engine
lib = eng.lib
t0 = time.perf_counter()
sv.column_hoist = "on"
Tc = sv.sweep_tables()
t1 = time.perf_counter()
sv.column_hoist = "off"
Tb = sv.sweep_tables()
t2 = time.perf_counter()
print("tables: CF %.1f s (%s, chunk %d, %d items, %d segs), BF %.1f s (%s, chunk %d, %d items)"
      % (t1 - t0, Tc.layout_name, Tc.item_chunk, Tc.n_items, Tc.n_segs, t2 - t1, Tb.layout_name,
         Tb.item_chunk, Tb.n_items), flush=True)
n_grid = n_E * n_P
J0 = eng.to_device(np.random.default_rng(0).standard_normal(n_grid))


def sweep_pair(T, J, n):
    a, b = J.clone(), torch.empty_like(J)
    for _ in range(n):
        eng.sweep(T, a, b)
        a, b = b, a
    torch.cuda.synchronize()
    return a, T.argmin[:T.n_states].clone()


for n in (1, 3):
    Jb, ab = sweep_pair(Tb, J0, n)
    Jc, ac = sweep_pair(Tc, J0, n)
    same_J = bool(torch.equal(Jb.view(torch.int64), Jc.view(torch.int64)))
    same_a = bool(torch.equal(ab, ac))
    print("after %d sweeps: J bit-identical %s, argmin identical %s (%d states differ)"
          % (n, same_J, same_a, int((ab != ac).sum().item())), flush=True)


def timed(T, n=10, warm=3):
    a, b = J0.clone(), torch.empty_like(J0)
    for _ in range(warm):
        eng.sweep(T, a, b)
        a, b = b, a
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s.record()
    for k in range(n):
        eng.sweep(T, a, b, events=ev[k])
        a, b = b, a
    e.record()
    torch.cuda.synchronize()
    k1 = float(np.median([x.elapsed_time(y) for x, y in ev]))
    tot = s.elapsed_time(e) / n
    return "%.3f ms/sweep (kernel %.3f ms) = %.0f G backups/s" % (tot, k1, T.n_backups_local / tot / 1e6)


print("BF            :", timed(Tb), flush=True)
sm = torch.cuda.get_device_properties(0).multi_processor_count
for pre in (1, 0):
    _cabi.check(lib.sdp_set_option(b"col_prepass", pre), "opt")
    print("CF default, column tables %s:" % ("from the pre-pass" if pre else "gathered by every CTA"),
          timed(Tc), flush=True)
lib.sdp_set_option(b"col_prepass", 1)
for per_sm in (1, 2, 4):
    eng.set_column_segments(Tc, sm * per_sm)
    for threads in (512, 384, 256):
        for ub, pf in ((2, 2), (2, 1), (1, 2), (1, 1)):
            _cabi.check(lib.sdp_set_option(b"col_threads", threads), "opt")
            _cabi.check(lib.sdp_set_option(b"col_ub", ub), "opt")
            _cabi.check(lib.sdp_set_option(b"col_pf", pf), "opt")
            print("CF segs/SM=%d threads=%d ub=%d pf=%d:" % (per_sm, threads, ub, pf), timed(Tc), flush=True)
lib.sdp_set_option(b"col_threads", 512)
lib.sdp_set_option(b"col_ub", 2)
lib.sdp_set_option(b"col_pf", 2)
eng.set_column_segments(Tc, sm)
# work-item length (tables rebuilt: host tabulation is cached only for the control boxes)
if os.environ.get("COLUMN_CHUNKS"):
    for chunk in (16, 64, 128):
        sv2 = wl.storage_ar1_large(sdp, n_E=n_E, n_P=n_P, item_chunk=chunk).solver
        sv2.column_hoist = "on"
        T2 = sv2.sweep_tables()
        print("CF item_chunk=%d (%d items):" % (chunk, T2.n_items), timed(T2), flush=True)
        del T2, sv2
