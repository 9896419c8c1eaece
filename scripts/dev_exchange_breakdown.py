"""developer experiment (torchrun, one process per GPU): where a sharded sweep of config #5 spends
its time on every rank - streaming kernel (pre-pass + column sweep), fused combine + peer stores +
epoch, flag wait - from CUDA events between the launches of Engine.sweep's peer-memory path.
    torchrun --nproc-per-node N scripts/dev_exchange_breakdown.py [K]"""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

K = int(sys.argv[1]) if len(sys.argv) > 1 else 30
os.environ.setdefault("SDP_P2P_TIMEOUT_S", "60")
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from stodynprog_b200 import _cabi  # noqa: E402

sv = wl.storage_ar1_large(sdp).solver
eng = sv.engine
T = sv.sweep_tables()
n_grid = int(np.prod(sv._state_grid_shape))
J_prev, J_new = eng.J_pair(n_grid)
eng.begin_call(n_grid)
eng.upload_J(np.random.default_rng(0).standard_normal(n_grid), J_prev)
px = eng._peer[n_grid]
assert px is not None, "needs the peer-memory exchange"
st = eng.torch_stream
lib = eng.lib


FOLD = os.environ.get("FOLD", "0") == "1"     # the wait rides in the next sweep's pre-pass
pending = [None]


def one(J_prev, J_new, ev):
    k_new = px.index_of(J_new)
    ev[0].record(st)
    if pending[0] is not None:
        _cabi.check(lib.sdp_sweep_partials_after(ctypes.byref(T.grid), ctypes.byref(T.c_tables), eng._ptr(J_prev),
                                                 eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                                 ctypes.byref(pending[0]), eng.stream), "partials_after")
        pending[0] = None
    else:
        _cabi.check(lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables), eng._ptr(J_prev),
                                           eng._ptr(T.part_val), eng._ptr(T.part_idx), eng.stream), "partials")
    ev[1].record(st)
    if T.col_bounds is not None:
        rc = lib.sdp_sweep_finalize_p2p_cols(ctypes.byref(T.c_tables), eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                             eng._ptr(T.argmin), ctypes.byref(px.peers[k_new]),
                                             T.col_bounds[-1], T.col_bounds[rank], eng.stream)
    else:
        rc = lib.sdp_sweep_finalize_p2p(ctypes.byref(T.c_tables), eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                        eng._ptr(T.argmin), ctypes.byref(px.peers[k_new]), T.state_begin,
                                        eng.stream)
    _cabi.check(rc, "finalize_p2p")
    ev[2].record(st)
    if FOLD:
        pending[0] = px.peers[k_new]
    else:
        _cabi.check(lib.sdp_p2p_wait(ctypes.byref(px.peers[k_new]), eng.stream), "wait")
    ev[3].record(st)


evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
for _ in range(5):
    one(J_prev, J_new, evs[0])
    J_prev, J_new = J_new, J_prev
dist.barrier()
torch.cuda.synchronize()
for k in range(K):
    one(J_prev, J_new, evs[k])
    J_prev, J_new = J_new, J_prev
if pending[0] is not None:
    _cabi.check(lib.sdp_p2p_wait(ctypes.byref(pending[0]), eng.stream), "wait")
torch.cuda.synchronize()
seg = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(3)] for e in evs[3:]])
gap = np.array([evs[k][3].elapsed_time(evs[k + 1][0]) for k in range(3, K - 1)])
tot = evs[3][0].elapsed_time(evs[K - 1][3]) / (K - 3)
mine = torch.tensor(list(seg.mean(axis=0)) + [gap.mean(), tot], dtype=torch.float64, device="cuda")
allr = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
dist.all_gather(allr, mine)
if rank == 0:
    print("shards of %s, %d ranks, %s; wait folded into the pre-pass: %s; ms per sweep" % ("columns" if T.col_bounds is not None else "rows",
                                                      dist.get_world_size(), _cabi.last_kernel(), FOLD))
    print("rank  kernel  combine+stores+epoch  wait  gap   step")
    for r, x in enumerate(allr):
        print("%4d  %.4f  %.4f                %.4f %.4f %.4f" % ((r,) + tuple(float(v) for v in x)))
# where the combine + exchange launch spends its time: the same combine with local stores only
# (sdp_sweep_finalize), then the fused launch with parts of the exchange switched off (timing
# only: `dbg_exchange` gives wrong results)
def time_fn(fn, reps=20):
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in e:
        a.record(st)
        fn()
        b.record(st)
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in e[3:]]))


def fused():
    k_new = px.index_of(J_new)
    if T.col_bounds is not None:
        rc = lib.sdp_sweep_finalize_p2p_cols(ctypes.byref(T.c_tables), eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                             eng._ptr(T.argmin), ctypes.byref(px.peers[k_new]),
                                             T.col_bounds[-1], T.col_bounds[rank], eng.stream)
    else:
        rc = lib.sdp_sweep_finalize_p2p(ctypes.byref(T.c_tables), eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                        eng._ptr(T.argmin), ctypes.byref(px.peers[k_new]), T.state_begin, eng.stream)
    _cabi.check(rc, "finalize_p2p")
    _cabi.check(lib.sdp_p2p_wait(ctypes.byref(px.peers[k_new]), eng.stream), "wait")


def local():
    _cabi.check(lib.sdp_sweep_finalize(ctypes.byref(T.c_tables), eng._ptr(T.part_val), eng._ptr(T.part_idx),
                                       eng._ptr(T.J_out), eng._ptr(T.argmin), eng.stream), "finalize")


res = {"local combine": time_fn(local)}
for name, dbg in (("fused + wait", 0), ("  no remote stores", 1), ("  no system fence", 2), ("  relaxed flag stores", 4),
                  ("  none of the three", 7)):
    dist.barrier()
    lib.sdp_set_option(b"dbg_exchange", dbg)
    res[name] = time_fn(fused)
lib.sdp_set_option(b"dbg_exchange", 0)
if rank == 0:
    for k, v in res.items():
        print("%-24s %.4f ms" % (k, v))
dist.barrier()
dist.destroy_process_group()
