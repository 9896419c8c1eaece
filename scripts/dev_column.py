"""one-GPU check of layout CF (column-shared hoist) on config #5: bit-identity against
layout BF on the full grid, device-resident and end-to-end timings, then a sweep over CTA
size / controls per iteration / groups in flight / CTA segments per SM.
      python scripts/dev_column.py [n_E] [n_P]"""
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

n_E = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_P = int(sys.argv[2]) if len(sys.argv) > 2 else 500
prob = wl.storage_ar1_large(sdp, n_E=n_E, n_P=n_P)
sv = prob.solver
eng = sv.engine
lib = eng.lib
n_grid = n_E * n_P
dims = (n_E, n_P)
J0_host = np.random.default_rng(0).standard_normal(dims)
J0 = eng.to_device(J0_host.reshape(-1))


def build(colmode, bands):
    Engine.COLUMN_BANDS = bands
    sv.column_hoist = colmode
    sv._table_cache = {}
    t0 = time.perf_counter()
    T = sv.sweep_tables()
    print("tables %-22s bands=%-4s %.1f s: chunk %d, %d items, %d segs, %.2f GB"
          % (T.layout_name, bands, time.perf_counter() - t0, T.item_chunk, T.n_items, T.n_segs,
             T.device_bytes / 1e9), flush=True)
    return T


def sweeps(T, n):
    a, b = J0.clone(), torch.empty_like(J0)
    for _ in range(n):
        eng.sweep(T, a, b)
        a, b = b, a
    torch.cuda.synchronize()
    return a, T.argmin[:T.n_states].clone()


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
    return "%.3f ms/sweep (streaming pass %.3f ms) = %.0f G backups/s" % (tot, k1, T.n_backups_local / tot / 1e6)


def e2e(n=8):
    J_h = J0_host
    for _ in range(2):
        J_h, pol_h = sv.value_iteration(J_h, report_time=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        J_h, pol_h = sv.value_iteration(J_h, report_time=False)
    dt = (time.perf_counter() - t0) / n
    T = sv.last_tables
    return "e2e %.3f ms/call = %.0f G backups/s (overlapped result copy: %s)" % (
        1e3 * dt, T.n_backups_total / dt / 1e9, eng.can_overlap_results(T)), J_h, pol_h


quick = bool(os.environ.get("COLUMN_QUICK"))
Tc = build("on", "1")
ref = {}
for n in (1, 3):
    ref[n] = sweeps(Tc, n)
print("CF 1 band     :", timed(Tc), flush=True)
msg, Jc_h, polc_h = e2e()
print("CF 1 band     :", msg, flush=True)
for pre in (0, 1, 2):
    _cabi.check(lib.sdp_set_option(b"col_prepass", pre), "opt")
    print("CF 1 band, column tables %s:" % ("gathered by every CTA", "pre-pass + vector-load copy",
                                            "pre-pass + TMA bulk copy")[pre], timed(Tc), flush=True)
sm = torch.cuda.get_device_properties(0).multi_processor_count
for per_sm in ((1,) if quick else (1, 2)):
    eng.set_column_segments(Tc, sm * per_sm)
    for threads in (512, 576, 640, 704, 768):
        for ub, pf in ((2, 2), (2, 1), (1, 2)):
            _cabi.check(lib.sdp_set_option(b"col_threads", threads), "opt")
            _cabi.check(lib.sdp_set_option(b"col_ub", ub), "opt")
            _cabi.check(lib.sdp_set_option(b"col_pf", pf), "opt")
            print("CF segs/SM=%d threads=%d ub=%d pf=%d:" % (per_sm, threads, ub, pf), timed(Tc, n=6, warm=2),
                  flush=True)
lib.sdp_set_option(b"col_threads", 640)
lib.sdp_set_option(b"col_ub", 2)
lib.sdp_set_option(b"col_pf", 2)
del Tc

Tb = build("off", "1")
for n in (1, 3):
    Jb, ab = sweeps(Tb, n)
    Jc, ac = ref[n]
    print("after %d sweeps, CF vs BF: J bit-identical %s, argmin identical %s (%d states differ)"
          % (n, bool(torch.equal(Jb.view(torch.int64), Jc.view(torch.int64))), bool(torch.equal(ab, ac)),
             int((ab != ac).sum().item())), flush=True)
print("BF            :", timed(Tb), flush=True)
msg, Jb_h, polb_h = e2e()
print("BF            :", msg, flush=True)
print("e2e results CF == BF:", np.array_equal(Jc_h.view(np.int64), Jb_h.view(np.int64)),
      np.array_equal(polc_h, polb_h), flush=True)
del Tb

for bands in ("auto", "3", "2"):
    T5 = build("on", bands)
    for threads in (512, 640):
        _cabi.check(lib.sdp_set_option(b"col_threads", threads), "opt")
        print("CF bands %s threads=%d:" % (T5.bands["rows"], threads), timed(T5), flush=True)
        J5, a5 = sweeps(T5, 3)
        print("   vs CF 1 band after 3 sweeps: J %s argmin %s"
              % (bool(torch.equal(J5.view(torch.int64), ref[3][0].view(torch.int64))),
                 bool(torch.equal(a5, ref[3][1]))), flush=True)
        msg, J5_h, pol5_h = e2e()
        print("   ", msg, "; results == BF:", np.array_equal(J5_h.view(np.int64), Jb_h.view(np.int64)),
              np.array_equal(pol5_h, polb_h), flush=True)
    lib.sdp_set_option(b"col_threads", 640)
    del T5
