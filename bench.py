#!/usr/bin/env python
"""bench.py - Bellman backups/s of the sweep hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one full Bellman sweep (value_iteration's hot path) over the
configured state grid.  Workload: BASELINE.json configs[4], the storage-AR1
problem scaled to 2000 x 500 states x <=256 controls x 9 perturbation nodes
(1 848 240 000 admissible backups per sweep, 38.6 GB of tables), sharded over
the N GPUs (strong scaling: total work fixed).  One JSON line on stdout.

  value      whole-job backups/s, tables and J resident in HBM, CUDA events on
             the launching stream, barrier + synchronize on both sides, max over ranks
  e2e        the same metric through the public API with HOST arrays:
             DPSolver.value_iteration(J_host) -> (J_host, pol_host), i.e. H2D of J,
             sweep, D2H of J and argmin, host mapping argmin -> control values
  roofline   HBM: algorithmic bytes (4 + 8d + 8/W per backup) of the
             streaming kernel / its own CUDA-event duration, vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's numpy/Cython loop, 1 core (the
             reference is single-threaded), on a bounded random sample of states

--impl reference times the reference's CPU path (oracle port; its interpolation
runs through the reference's own compiled Cython routine when oracle/_ref is
present) on bounded samples of the same workload and prints the same line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "bellman_backups_per_sec"
UNIT = "backups/s"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc = None
        self.lines = []
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_problem(api, args, **solver_kw):
    from stodynprog_b200 import workloads as wl
    if args.workload == "ar1":
        return wl.storage_ar1(api, **solver_kw), "howto storage-AR1 41x61 (BASELINE configs[2])"
    prob = wl.storage_ar1_large(api, n_E=args.n_E, n_P=args.n_P, **solver_kw)
    name = "synthetic storage-AR1 %dx%d states x <=256 controls x 9 nodes (BASELINE configs[4])" \
        % (args.n_E, args.n_P)
    return prob, name


def cpu_port_sample(args, n_states, seed=0, interp="c", J=None):
    """time the oracle port's per-state backup (what the reference's value_iteration
    does for each state, stodynprog.py:511-515) on `n_states` random states of the
    workload.  Returns (backups, seconds)."""
    from oracle.ref_port import port_api
    prob, _ = make_problem(port_api(interp), args)
    sv = prob.solver
    dims = sv._state_grid_shape
    rng = np.random.default_rng(seed)
    if J is None:
        J = np.random.default_rng(0).standard_normal(dims)
    J_interp = sv.interp_on_state(J)
    n_grid = int(np.prod(dims))
    picks = rng.choice(n_grid, size=min(n_states, n_grid), replace=False)
    W = len(sv.perturb_grid[0])
    backups = 0
    t0 = time.perf_counter()
    for flat in picks:
        idx = np.unravel_index(flat, dims)
        x_k = tuple(g[i] for g, i in zip(sv.state_grid, idx))
        _, _, _, Jall = sv.value_at_state(x_k, J_interp, None, True)
        backups += Jall.size * W
    return backups, time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference's CPU path, bounded samples, rank 0 only"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build as ob
    ob.build()
    interp = "c"
    kind = "port"
    try:
        from oracle import build_ref
        if build_ref.build() is not None:
            from oracle.ref_loader import load_reference_cython
            if load_reference_cython() is not None:
                interp = "ref"
    except Exception:
        pass
    _, name = make_problem(__import__("oracle.ref_port", fromlist=["port_api"]).port_api(interp), args)
    n_sample = args.cpu_sample
    for _ in range(args.warmup):
        cpu_port_sample(args, max(n_sample // 10, 10), seed=99, interp=interp)
    tot_b, tot_t = 0, 0.0
    for k in range(args.steps):
        b, t = cpu_port_sample(args, n_sample, seed=k, interp=interp)
        tot_b += b
        tot_t += t
    value = tot_b / tot_t
    sample = ("%d random states per step (seeded) of the %s grid, per-state numpy loop of the "
              "reference restated in oracle/ref_port.py; interpolation through %s"
              % (n_sample, "x".join(str(n) for n in (args.n_E, args.n_P)),
                 "the reference's own compiled Cython routine (oracle/_ref)" if interp == "ref"
                 else "the C restatement (oracle/sdp_oracle.c)"))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def kernel_name(T):
    """the streaming kernel the library launches for this table layout (defaults of
    csrc/sdp_b200.cu: TMA ring R=8 for layout B, hoisted inner interpolation for AF
    with u_mask == 1)"""
    d = T.d
    if T.factored:
        if T.column:
            return "k_column_table<%d> + k_sweep_fact_column<%d,%d,2,2,%s,640>" % (
                d, d, 3 if T.W <= 3 else (5 if T.W <= 5 else 9), "true" if T.W in (3, 5, 9) else "false")
        if T.tiled:
            return "k_sweep_fact_tiled<%d,%d,%d>" % (d, T.u_mask, 3 if T.W <= 3 else (5 if T.W <= 5 else 9))
        if T.u_mask == 1:
            wm = 3 if T.W <= 3 else (5 if T.W <= 5 else 9)
            return "k_sweep_fact_hoist_c<%d,%d>" % (d, wm) if T.W <= 9 else "k_sweep_fact_hoist<%d,2>" % d
        return "k_sweep_fact<%d,%d,4>" % (d, T.u_mask)
    return "k_sweep_tiled_tma<%d,8>" % d if T.tiled else "k_sweep<%d,4>" % d


def ncu_counters(workload, T, world):
    """what the committed `ncu --set full` capture of this workload's streaming kernel says
    (profiles/r1_traffic.json; N=1 only): {"traffic": DRAM bytes per launch, pipe
    utilisations, source file}, or {}"""
    if world != 1:
        return {}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            e = json.load(f).get("%s/%s" % (workload, T.layout_name))
    except Exception:
        return {}
    if e is None:
        return {}
    return e if isinstance(e, dict) else {"traffic": e}


def ncu_traffic(workload, T, world):
    return ncu_counters(workload, T, world)


def time_sweeps(eng, T, J_prev, J_new, K, barrier):
    """K timed sweeps, CUDA events on the launching stream around the whole loop and
    around each streaming-kernel launch.  Returns (ms_total, k1_ms array, J_prev, J_new)."""
    import torch
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for k in range(K):
        eng.sweep(T, J_prev, J_new, events=kev[k])
        J_prev, J_new = J_new, J_prev
    end.record()
    barrier()
    return start.elapsed_time(end), np.array([a.elapsed_time(b) for a, b in kev]), J_prev, J_new


def roofline_of(T, k1_ms, ms_total, K, peak, peak_src, traffic):
    b_alg = T.algorithmic_bytes_per_backup
    k1 = float(np.mean(k1_ms))
    achieved = T.n_backups_local * b_alg / (k1 * 1e-3) / 1e9
    ncu = traffic if isinstance(traffic, dict) else {}
    traffic = ncu.get("traffic")
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
         "kernel": kernel_name(T), "kernel_ms": k1, "kernel_share_of_step": k1 * K / ms_total,
         "algorithmic_bytes_per_backup": b_alg,
         "algorithmic_bytes_per_launch": T.n_backups_local * b_alg,
         "table_layout": T.layout_name,
         "table_bytes_resident": T.device_bytes,
         "streamed_bytes_per_backup": T.streamed_bytes_per_backup,
         "streamed_GBs": T.device_bytes / (k1 * 1e-3) / 1e9}
    if len(ncu) > 1:
        r["ncu"] = {k: v for k, v in ncu.items() if k != "traffic"}
        # the pipe the committed ncu capture shows closest to its peak: what actually bounds the
        # kernel when `frac` (dense algorithmic bytes over the HBM peak) is not the binding ratio
        pipes = {"l1_data_pipe_pct": "L1 / shared-memory data pipe (wavefronts)", "fp64_pipe_pct": "fp64 pipe",
                 "dram_pct": "HBM", "l1_pct": "L1", "l2_pct": "L2"}
        seen = [(float(ncu[k]), name) for k, name in pipes.items() if k in ncu]
        if seen:
            top = max(seen)
            r["limiter"] = {"unit": top[1], "frac_of_peak": round(top[0] / 100.0, 3),
                            "source": ncu.get("source")}
    if T.column:
        r["note"] = ("column-shared hoist over factored (x,u)+(x,w) tables: the kernel streams %.2f B "
                     "per backup instead of the dense layout's %.2f B and reads the inner "
                     "interpolation from a per-column table in shared memory, so `achieved` (dense "
                     "algorithmic bytes / time, SURVEY.md 8d) exceeds the HBM peak; bound by shared-"
                     "memory wavefronts / the fp64 pipe, see `dense_layout` for the HBM-bound kernel "
                     "on the same workload" % (T.streamed_bytes_per_backup, b_alg))
    elif T.factored:
        r["note"] = ("factored (x,u)+(x,w) tables: the kernel streams %.2f B per backup instead of "
                     "the dense layout's %.2f B, so `achieved` (dense algorithmic bytes / time, "
                     "SURVEY.md 8d) exceeds the HBM peak; the kernel is bound by the L1 wavefronts "
                     "of the corner gathers / the fp64 pipe, see `dense_layout` for the "
                     "HBM-bound kernel on the same workload" % (T.streamed_bytes_per_backup, b_alg))
    else:
        r["padding_fill"] = T.n_backups_local / max(T.n_entries, 1)
    return r


def run_ours(args):
    # a rank that dies must fail the run quickly, not leave its peers spinning on flags
    os.environ.setdefault("SDP_P2P_TIMEOUT_S", "120")
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import stodynprog_b200 as sdp
    from stodynprog_b200 import _cabi
    from stodynprog_b200 import build as product_build
    if rank == 0:
        product_build.build()
    if world > 1:
        dist.barrier()

    prob, name = make_problem(sdp, args)
    sv = prob.solver
    if args.layout != "auto":
        sv.table_layout = args.layout
    sv.table_compress = args.compress
    sv.column_hoist = args.column_hoist
    if args.item_chunk:
        sv._item_chunk = args.item_chunk
    eng = sv.engine
    t0 = time.perf_counter()
    T = sv.sweep_tables()
    setup_s = time.perf_counter() - t0
    dims = sv._state_grid_shape
    n_grid = int(np.prod(dims))

    J_host = np.random.default_rng(0).standard_normal(dims)
    J_prev, J_new = eng.J_pair(n_grid)
    eng.begin_call(n_grid)
    eng.upload_J(J_host, J_prev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks / throttle reasons are sampled from the warm-up to the end of the e2e loop
    # (the device-timed region alone is only K x 3 ms: too short for nvidia-smi's 100 ms period)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.2)
    for _ in range(args.warmup):
        eng.sweep(T, J_prev, J_new)
        J_prev, J_new = J_new, J_prev

    K = args.steps
    launches0 = _cabi.launch_count()
    ms_total, k1_ms, J_prev, J_new = time_sweeps(eng, T, J_prev, J_new, K, barrier)
    launches = _cabi.launch_count() - launches0
    t = torch.tensor([ms_total, float(np.mean(k1_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total_max, k1_ms_max = float(t[0]), float(t[1])
    k1_per_rank = None
    if world > 1:
        # slab balance: every rank's mean streaming-kernel time and slab size
        mine = torch.tensor([float(np.mean(k1_ms)), float(T.n_backups_local), float(T.n_states)],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        k1_per_rank = {"balance": "re-cut by measured slab times" if T.slab_recut else "equal admissible controls",
                       "kernel_ms_before_recut": T.slab_times_ms,
                       "kernel_ms": [round(float(x[0]), 4) for x in allr],
                       "backups": [int(x[1]) for x in allr], "states": [int(x[2]) for x in allr]}
    total_backups = T.n_backups_total
    value = total_backups * K / (ms_total_max * 1e-3)

    # roofline of the streaming kernel on this rank's slab
    peak, peak_src = read_peaks()
    roofline = roofline_of(T, k1_ms, ms_total, K, peak, peak_src, ncu_traffic(args.workload, T, world))
    J_keep = J_prev.clone()

    # end-to-end through the public API, host arrays in and out
    n_e2e = max(3, min(K, 10))

    def time_e2e(mode):
        """n_e2e calls of value_iteration with host arrays; `mode` = solver.host_results"""
        sv.host_results = mode
        J_h = J_keep.cpu().numpy().reshape(dims)
        for _ in range(2):
            J_h, pol_h = sv.value_iteration(J_h, report_time=False)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            J_h, pol_h = sv.value_iteration(J_h, report_time=False)
        torch.cuda.synchronize()
        e2e_local = time.perf_counter() - t0
        te = torch.tensor([e2e_local], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te[0])
        nb_J = 8 * n_grid
        nb_pol = 8 * n_grid * len(sv.sys.control)
        copying = world if mode == "all" else 1       # ranks that move host data
        return {"value": total_backups * n_e2e / e2e_s, "unit": UNIT,
                "h2d_bytes_per_step": nb_J * copying, "d2h_bytes_per_step": (nb_J + nb_pol) * copying,
                "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e,
                "api": "DPSolver.value_iteration(J_host) -> (J_host, pol_host)" + (
                    "" if world == 1 else
                    ", solver.host_results = 'all': every rank uploads J and downloads (J, pol)" if mode == "all"
                    else ", solver.host_results = 'root': rank 0 uploads J (handed to the other ranks "
                         "over NVLink) and rank 0 alone downloads (J, pol)")}

    e2e = time_e2e("all")
    clocks = sampler.stop() if rank == 0 else None
    e2e_all = None
    if world > 1:
        # 8 ranks pulling 24 MB each through shared PCIe roots take ~2 ms, one rank 0.45 ms
        # (scripts/dev_pcie_contention.py): the headline uses the root-only result mode
        e2e_all = e2e
        e2e = time_e2e("root")
        sv.host_results = "all"

    # the same workload through the dense (x,u,w) tables: the HBM-bound kernel the
    # roofline target is stated for
    dense = None
    if T.factored and not args.no_dense:
        sv.table_compress = "off"
        sv.column_hoist = "off"
        Td = sv.sweep_tables()
        Ja, Jb = eng.J_pair(n_grid)
        eng.begin_call(n_grid)
        Ja.copy_(J_keep)
        for _ in range(3):
            eng.sweep(Td, Ja, Jb)
            Ja, Jb = Jb, Ja
        Kd = max(5, min(K, 10))
        ms_d, k1_d, Ja, Jb = time_sweeps(eng, Td, Ja, Jb, Kd, barrier)
        td = torch.tensor([ms_d], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        dense = {"value": total_backups * Kd / (float(td[0]) * 1e-3), "unit": UNIT, "steps": Kd,
                 "ms_per_step": float(td[0]) / Kd,
                 "roofline": roofline_of(Td, k1_d, ms_d, Kd, peak, peak_src,
                                         ncu_traffic(args.workload, Td, world))}
        sv.table_compress = args.compress
        sv.column_hoist = args.column_hoist
        del Td
        sv.clear_tables()
        torch.cuda.empty_cache()

    setup = torch.tensor([setup_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(setup, op=dist.ReduceOp.MAX)

    extra = None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import build as ob
        ob.build()
        b, tcpu = cpu_port_sample(args, args.cpu_sample, seed=0)
        cpu = {"value": b / tcpu, "unit": UNIT, "cores": 1, "kind": "port",
               "host_cpus": os.cpu_count(), "seconds": tcpu,
               "sample": "%d random states (seeded) of the same grid, %d backups, per-state "
                         "numpy loop of the reference restated in oracle/ref_port.py (the "
                         "reference is single-threaded: prange compiled without OpenMP)"
                         % (args.cpu_sample, b)}
    if rank == 0 and world == 1 and args.workload == "large" and not args.no_extra:
        extra = measure_config3(sdp, peak, peak_src)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": args.warmup, "ms_per_step": ms_total_max / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "states": n_grid, "backups_per_sweep": total_backups,
                       "state_dims": list(dims), "perturbation_nodes": T.W,
                       "table_layout": T.layout_name,
                       "row_bands": (T.bands["rows"] if T.column else None),
                       "tabulate_mode": T.tabulate_mode, "item_chunk": T.item_chunk,
                       "parallelism": "state slabs x%d, %s" % (
                           world, "one rank" if world == 1 else
                           ("J slab stored into every rank's buffer by the combine kernel over NVLink "
                            "peer memory + flag wait" if eng.peer_exchange(n_grid) is not None
                            else "NCCL all-gather of J per sweep")),
                       "l2": "tables streamed once per sweep (%.1f GB per GPU) >> 126 MB L2; "
                             "no flush needed" % (T.device_bytes / 1e9),
                       "J_init": "default_rng(0).standard_normal, then fed back sweep to sweep"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "clocks": clocks, "setup_seconds": float(setup[0]),
        }
        if e2e_all is not None:
            line["e2e_all_ranks"] = e2e_all
        if k1_per_rank is not None:
            line["slabs"] = k1_per_rank
        if dense is not None:
            line["dense_layout"] = dense
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if extra is not None:
            line["storage_ar1_41x61"] = extra
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_config3(sdp, peak, peak_src, steps=50, warmup=5):
    """BASELINE configs[2] (the 41x61 storage-AR1 grid of the notebook, 142 762 509
    backups per sweep): the grid the north star's 60 % roofline target is quoted on.
    Reported for the default (factored) tables and for the dense tables."""
    import torch
    from stodynprog_b200 import workloads as wl
    out = {"workload": "howto storage-AR1 41x61 x 4001..8001 controls x 9 nodes"}
    for compress in ("auto", "off"):
        prob = wl.storage_ar1(sdp)
        sv = prob.solver
        sv.table_compress = compress
        eng = sv.engine
        T = sv.sweep_tables()
        J_prev, J_new = eng.J_pair(41 * 61)
        eng.upload_J(np.random.default_rng(0).standard_normal(41 * 61), J_prev)
        for _ in range(warmup):
            eng.sweep(T, J_prev, J_new)
            J_prev, J_new = J_new, J_prev
        ms_total, k1_ms, J_prev, J_new = time_sweeps(eng, T, J_prev, J_new, steps,
                                                     torch.cuda.synchronize)
        out["backups_per_sweep"] = T.n_backups_total
        out["default" if compress == "auto" else "dense_layout"] = {
            "ms_per_step": ms_total / steps, "value": T.n_backups_total * steps / (ms_total * 1e-3),
            "unit": UNIT,
            "roofline": roofline_of(T, k1_ms, ms_total, steps, peak, peak_src, ncu_traffic("ar1", T, 1))}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="large", choices=["large", "ar1"])
    ap.add_argument("--n-E", dest="n_E", type=int, default=2000)
    ap.add_argument("--n-P", dest="n_P", type=int, default=500)
    ap.add_argument("--layout", default="auto", choices=["auto", "control_minor", "state_minor"])
    ap.add_argument("--compress", default="auto", choices=["auto", "off", "on"],
                    help="factored (x,u)+(x,w) tables (auto: whenever the system allows)")
    ap.add_argument("--column-hoist", dest="column_hoist", default="auto", choices=["auto", "on", "off"],
                    help="layout CF, one inner-interpolation table per grid column (auto: SDP_COLUMN_HOIST)")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-layout sub-measurement")
    ap.add_argument("--item-chunk", dest="item_chunk", type=int, default=0)
    ap.add_argument("--cpu-sample", dest="cpu_sample", type=int, default=None,
                    help="states per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.cpu_sample is None:
        # ~200 us per state for the 2000x500 grid (<=256 controls); ~1.4 ms for the 41x61 grid
        args.cpu_sample = (60000 if args.impl == "ours" else 8000) if args.workload == "large" else 2501
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
