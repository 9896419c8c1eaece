"""target of the compute-sanitizer passes (scripts/gpu.sh sanitize): every sweep kernel family on
small grids - dense A / B (TMA ring: mbarriers + bulk copies), factored AF (hoisted) / BF, layout
CF (column-table pre-pass, bulk-copied table hand-over, several row bands, transposing combine),
fixed-policy backup, K0 builds, K2 / K3 - each checked against the oracle port so that a run that
passes the sanitizer is also a correct one.
    compute-sanitizer --tool memcheck|racecheck python scripts/sanitizer_target.py
With torchrun (2 ranks) it runs the sharded path instead: fused combine + peer stores + epoch
flags, flag wait folded into the pre-pass, rows and columns."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import stodynprog_b200 as sdp  # noqa: E402
import workloads as wl  # noqa: E402
from stodynprog_b200.engine import Engine  # noqa: E402
from oracle.ref_port import port_api  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("SDP_P2P_TIMEOUT_S", "300")
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))


def check(name, sv, so, J0, sweeps=2):
    J = J0
    for _ in range(sweeps):
        Jg, polg = sv.value_iteration(J, report_time=False)
        Jo, polo = so.value_iteration(J)
        assert np.array_equal(polg, polo), name
        assert np.max(np.abs(Jg - Jo) / np.maximum(np.abs(Jo), 1e-300)) <= 1e-10, name
        J = Jo
    if rank == 0:
        print("ok  %-34s %s" % (name, sv.last_tables.layout_name), flush=True)


port = port_api()
if world == 1:
    cases = [("dense A", dict(table_layout="control_minor", table_compress="off")),
             ("dense B (TMA ring)", dict(table_layout="state_minor", table_compress="off")),
             ("factored AF (hoisted)", dict(table_layout="control_minor")),
             ("factored BF", dict(table_layout="state_minor", column_hoist="off")),
             ("column CF, one band", dict(table_layout="state_minor", column_hoist="on"))]
    for name, knobs in cases:
        kw = dict(n_E=70, n_P=6, n_w=9, steps=(0.3, 0.1))
        sv, so = wl.storage_ar1(sdp, **kw).solver, wl.storage_ar1(port, **kw).solver
        for k, v in knobs.items():
            setattr(sv, k, v)
        check(name, sv, so, np.random.default_rng(1).standard_normal((70, 6)))
    Engine.COLUMN_BANDS = "3"
    kw = dict(n_E=330, n_P=3, n_w=3, steps=(2.0, 0.1))
    sv, so = wl.storage_ar1(sdp, **kw).solver, wl.storage_ar1(port, **kw).solver
    sv.table_layout, sv.column_hoist = "state_minor", "on"
    check("column CF, three bands, W = 3", sv, so, np.random.default_rng(2).standard_normal((330, 3)), 1)
    Engine.COLUMN_BANDS = "auto"
    # the host path of large sweeps: pieces of columns on alternating streams, combine + control
    # values in one launch (the small variant beside the next sweep, the wide one at the end),
    # 2-D copies into the result arrays
    saved = Engine.OVERLAP_MIN_BACKUPS, Engine.OVERLAP_MIN_ITEMS
    Engine.OVERLAP_MIN_BACKUPS = Engine.OVERLAP_MIN_ITEMS = 0
    kw = dict(n_E=100, n_P=37, n_w=9, steps=(0.5, 0.1))
    sv, so = wl.storage_ar1(sdp, **kw).solver, wl.storage_ar1(port, **kw).solver
    sv.table_layout, sv.column_hoist = "state_minor", "on"
    check("column CF, pieces of columns", sv, so, np.random.default_rng(6).standard_normal((100, 37)))
    assert sv.engine.can_overlap_results(sv.last_tables) and len(sv.engine._chunk_plan(sv.last_tables)) == 4
    Engine.OVERLAP_MIN_BACKUPS, Engine.OVERLAP_MIN_ITEMS = saved
    prob, ora = wl.searev(sdp, n_E=7, n_S=9, n_A=9), wl.searev(port, n_E=7, n_S=9, n_A=9)
    prob.solver.control_steps = ora.solver.control_steps = (.05,)
    check("SEAREV 3-D, AF", prob.solver, ora.solver, np.zeros((7, 9, 9)), 1)
    (Jd, Jr), pol = prob.solver.policy_iteration(prob.initial_policy(), 5, 1, rel_dp=True)
    (Jdo, Jro), polo = ora.solver.policy_iteration(ora.initial_policy(), 5, 1, rel_dp=True)
    assert np.array_equal(pol, polo) and abs(Jr - Jro) <= 1e-10 * abs(Jro)
    print("ok  policy iteration (fixed-policy backup, relative DP)", flush=True)
    pv, pvo = wl.pv_storage(sdp, horizon=6), wl.pv_storage(port, horizon=6)
    pv.solver.control_steps = pvo.solver.control_steps = (.02,)
    J, pol = pv.solver.bellman_recursion(6, pv.J_fin, report_time=False)
    Jo, polo = pvo.solver.bellman_recursion(6, pvo.J_fin)
    assert np.array_equal(pol, polo) and pv.solver.last_recursion is not None
    print("ok  bellman_recursion, fast path", flush=True)
    f = prob.solver.interp_on_state(np.random.default_rng(3).standard_normal((7, 9, 9)))
    x = np.random.default_rng(4).uniform(-1, 1, (3, 5000))
    f(x[0] * 10, x[1], x[2])
    print("ok  K2 interpolation (5000 points)", flush=True)
else:
    for axis in ("rows", "columns"):
        kw = dict(n_E=32 * world + 8, n_P=4 * world + 3, n_w=9, steps=(0.5, 0.1))
        sv, so = wl.storage_ar1(sdp, **kw).solver, wl.storage_ar1(port, **kw).solver
        sv.table_layout, sv.column_hoist, sv.slab_axis = "state_minor", "on", axis
        J0 = np.random.default_rng(5).standard_normal(sv._state_grid_shape)
        check("column CF, %d ranks, shards of %s" % (world, axis), sv, so, J0)
        Js, pols, info = sv.solve_value_iteration(J_zero=J0, max_iter=3)     # flag wait folded into the pre-pass
        J = J0
        for _ in range(3):
            J, polo = so.value_iteration(J)
        assert np.array_equal(pols, polo)
        if rank == 0:
            print("ok  device-resident loop, %s" % axis, flush=True)
    sv = wl.storage_ar1(sdp, steps=(0.05, 0.1)).solver
    pol0 = wl.storage_ar1(sdp, steps=(0.05, 0.1)).initial_policy()
    (Jd, Jr), pol = sv.policy_iteration(pol0, 5, 1, rel_dp=True)
    if rank == 0:
        print("ok  sharded policy iteration (fused fixed-policy backup + exchange), J_ref %.6g" % Jr, flush=True)
    dist.barrier()
    dist.destroy_process_group()
print("SANITIZER_TARGET_DONE rank %d" % rank)
