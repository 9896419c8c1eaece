"""world_size-2 tests on CPU (gloo): the sharded sweep pipeline - slab
partition, per-rank tabulation, per-sweep all-gather of J, all-reduce-max of the
residual, gathered argmin -> policy - driven through the numpy model of the C
ABI, against the single-process result."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(sdp, wl, lib, layout):
    """the sharded test problem: storage + AR(1) on a 9 x 11 grid, or - layout CF, which
    needs slabs of whole rows of axis 0 and at least 32 rows per rank - on a 70 x 5 grid"""
    if layout.startswith("column"):
        prob = wl.storage_ar1(sdp, n_E=70, n_P=5, n_w=3, steps=(0.5, 0.1), _test_lib=lib)
        prob.solver.table_layout = "state_minor"
        prob.solver.column_hoist = "on"
        if layout == "column_by_columns":       # the grid cut into whole columns instead of rows
            prob.solver.slab_axis = "columns"
    else:
        prob = wl.storage_ar1(sdp, n_E=9, n_P=11, steps=(0.01, 0.1), _test_lib=lib)
        prob.solver.table_layout = layout
    J0 = np.random.default_rng(0).standard_normal(prob.solver._state_grid_shape)
    return prob, prob.solver, J0


def _worker(rank, world, port, out_dir, layout):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import stodynprog_b200 as sdp
        import workloads as wl
        from fake_lib import FakeLib
        prob, sv, J0 = _problem(sdp, wl, FakeLib(), layout)
        J1, pol1 = sv.value_iteration(J0, report_time=False)
        T = sv.last_tables
        assert T.column == layout.startswith("column")
        assert (T.col_bounds is not None) == (layout == "column_by_columns")
        (Jd, Jr), pol2 = sv.value_iteration((J1 - J1[sv._state_ref_ind], 0.), rel_dp=True, report_time=False)
        Je, ref = sv.eval_policy(pol1, 6, rel_dp=True, report_time=False)
        Js, pols, info = sv.solve_value_iteration(J_zero=J0, max_iter=3, tol=0.0)
        # host_results = "root": only rank 0's input is read, only rank 0 gets arrays
        sv.host_results = "root"
        Jq, polq = sv.value_iteration(J0 if rank == 0 else None, report_time=False)
        (Jqd, Jqr), polq2 = sv.value_iteration((J1 - J1[sv._state_ref_ind], 0.) if rank == 0 else None,
                                                rel_dp=True, report_time=False)
        Jqs, polqs, _ = sv.solve_value_iteration(J_zero=J0 if rank == 0 else None, max_iter=3, tol=0.0)
        if rank == 0:
            root_ok = (np.array_equal(Jq, J1) and np.array_equal(polq, pol1) and np.array_equal(Jqd, Jd)
                       and Jqr == Jr and np.array_equal(polq2, pol2) and np.array_equal(Jqs, Js)
                       and np.array_equal(polqs, pols))
        else:
            root_ok = all(x is None for x in (Jq, polq, Jqd, Jqr, polq2, Jqs, polqs))
        sv.host_results = "all"
        assert root_ok, "host_results='root' mismatch on rank %d" % rank
        np.savez(os.path.join(out_dir, "rank%d.npz" % rank), J1=J1, pol1=pol1, Jd=Jd, Jr=Jr, pol2=pol2,
                 Je=Je, ref=ref, Js=Js, pols=pols, resid=np.array(info["residuals"]),
                 bounds=np.array(T.bounds if T.bounds is not None else T.col_bounds), n_local=T.n_states,
                 backups=T.n_backups_local,
                 total=T.n_backups_total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("layout", ["control_minor", "state_minor", "column", "column_by_columns"])
def test_sharded_sweep_world2_matches_single_process(tmp_path, layout):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), layout), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % k)) for k in range(world)]

    # single-process run of the same thing
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stodynprog_b200 as sdp
    import workloads as wl
    from fake_lib import FakeLib
    prob, sv, J0 = _problem(sdp, wl, FakeLib(), layout)
    J1, pol1 = sv.value_iteration(J0, report_time=False)
    (Jd, Jr), pol2 = sv.value_iteration((J1 - J1[sv._state_ref_ind], 0.), rel_dp=True, report_time=False)
    Je, ref = sv.eval_policy(pol1, 6, rel_dp=True, report_time=False)
    Js, pols, info = sv.solve_value_iteration(J_zero=J0, max_iter=3, tol=0.0)

    for k in range(world):
        # every rank holds the full, identical result
        assert np.array_equal(r[k]["J1"], J1) and np.array_equal(r[k]["pol1"], pol1)
        assert np.array_equal(r[k]["Jd"], Jd) and r[k]["Jr"] == Jr and np.array_equal(r[k]["pol2"], pol2)
        assert np.array_equal(r[k]["Je"], Je) and r[k]["ref"] == ref
        assert np.array_equal(r[k]["Js"], Js) and np.array_equal(r[k]["pols"], pols)
        assert np.array_equal(r[k]["resid"], np.array(info["residuals"]))
    # the slabs are contiguous, disjoint, cover the grid and are balanced by controls
    b = r[0]["bounds"]
    n_grid = J0.size
    if layout == "column_by_columns":
        # the ranks hold whole columns: 2 + 3 of the 5, every row of them
        assert list(b) == list(r[1]["bounds"]) and b[0] == 0 and b[-1] == J0.shape[1] and 0 < b[1] < 5
        assert int(r[0]["n_local"]) == b[1] * J0.shape[0] and int(r[1]["n_local"]) == (5 - b[1]) * J0.shape[0]
        assert int(r[0]["backups"]) + int(r[1]["backups"]) == int(r[0]["total"])
        return
    assert list(b) == list(r[1]["bounds"]) and b[0] == 0 and b[-1] == n_grid
    assert int(r[0]["n_local"]) + int(r[1]["n_local"]) == n_grid
    if layout == "column":
        assert b[1] % J0.shape[1] == 0          # whole rows of axis 0 per rank
    assert int(r[0]["backups"]) + int(r[1]["backups"]) == int(r[0]["total"])
    assert abs(int(r[0]["backups"]) - int(r[1]["backups"])) / int(r[0]["total"]) < 0.1


def _coll_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stodynprog_b200.engine import Collective
        c = Collective()
        bounds = [0, 3, 10]
        full = torch.arange(10, dtype=torch.float64) * 1.5
        local = full[bounds[rank]:bounds[rank + 1]].clone()
        got = c.all_gather_slabs(local, bounds)
        out = torch.zeros(10, dtype=torch.float64)
        c.all_gather_slabs(local, bounds, out=out)
        m = c.all_reduce_max(torch.tensor([float(rank + 1)], dtype=torch.float64))
        objs = c.all_gather_object({"rank": rank})
        ok = bool(torch.equal(got, full) and torch.equal(out, full) and m.item() == world
                  and [o["rank"] for o in objs] == list(range(world)))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write(str(ok))
    finally:
        dist.destroy_process_group()


def test_collective_uneven_slabs_world2(tmp_path):
    mp.spawn(_coll_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for k in range(2):
        assert open(os.path.join(str(tmp_path), "ok%d" % k)).read() == "True"


def _share_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from stodynprog_b200.engine import Collective
        from stodynprog_b200.hostshare import HostShare
        hs = HostShare(Collective(), n_grid=12, nc=2, cuda=False)      # (no cudaHostRegister on CPU)
        ok = True
        # the segment is one piece of memory: what a rank writes into its part of a slot, the
        # others read after the flag barrier
        a, b = 12 * rank // world, 12 * (rank + 1) // world
        hs.slot_J(1).numpy()[a:b] = 100 * rank + np.arange(a, b)
        hs.slot_pol(1).numpy()[2 * a:2 * b] = -rank
        hs.publish(0, 7 + rank)
        hs.barrier()
        want = np.concatenate([100 * r + np.arange(12 * r // world, 12 * (r + 1) // world) for r in range(world)])
        ok &= np.array_equal(hs.slot_J(1).numpy(), want)
        ok &= [hs.read(r, 0) for r in range(world)] == [7 + r for r in range(world)]
        # arrays handed out keep their slot busy until they die; views of them count
        J, pol = hs.hand_out(1, (3, 4), (3, 4, 2), writable=rank == 0)
        ok &= hs.slot_of(J) == 1 and hs.slot_of(np.zeros(12)) is None and hs.live_mask() == 2
        ok &= J.flags.writeable == (rank == 0) and np.array_equal(J.reshape(-1), want)
        row = J[1]
        del J, pol
        ok &= hs.live_mask() == 2
        del row
        ok &= hs.live_mask() == 0
        for _ in range(50):
            hs.barrier()
        open(os.path.join(out_dir, "share%d" % rank), "w").write(str(bool(ok)))
    finally:
        dist.destroy_process_group()


def test_host_share_world2(tmp_path):
    """the shared page-locked result segment (hostshare.py) without a GPU: mapping, flag barrier,
    control words, slot liveness"""
    mp.spawn(_share_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for k in range(2):
        assert open(os.path.join(str(tmp_path), "share%d" % k)).read() == "True"
    assert not [f for f in os.listdir("/dev/shm") if f.startswith("sdp_b200_")]
