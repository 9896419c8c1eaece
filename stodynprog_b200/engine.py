"""Device side of the solver: table residency, sweep launches, collectives.

torch is plumbing here: it owns the device buffers, the stream and the
process group.  Every compute step goes through the C ABI (`_cabi`), i.e. the
hand-written sm_100a kernels in csrc/sdp_b200.cu.  There is no CPU path.

Multi-GPU model: one process per GPU; the C-order flattened state
grid is cut into contiguous slabs balanced by the number of admissible
controls; each rank holds the tables of its slab only; the value function J
(8 bytes per state) is replicated.  One exchange per sweep: all-gather of the
new J slab (+ an all-reduce-max of the sup-norm residual when asked).
"""
import ctypes
import os

import numpy as np

from . import _cabi
from . import tabulate as tb

__all__ = ["Engine", "SweepTables", "PolicyTables", "partition_by_weight", "rebalance_bounds",
           "column_order", "column_segments", "row_aligned"]

# solver.column_hoist = "auto" uses layout CF (column-shared hoist) whenever it applies unless
# SDP_COLUMN_HOIST=0; "on" / "off" on the solver override it.  Measured on config #5, one B200
# (profiles/r1_column_tuning.txt): 1.21 ms per sweep against 2.83 ms for layout BF.
COLUMN_HOIST_DEFAULT = os.environ.get("SDP_COLUMN_HOIST", "1") != "0"
# solver.slab_axis = "auto": how a grid in layout CF is cut over several ranks ("auto" | "rows" |
# "columns").  By columns every rank tabulates and loads the tables of its own columns only;
# measured on config #5 (profiles/r2_shard_emulation.txt, one rank's streaming kernel): 0.179 ms
# against 0.233 ms per sweep for 1/8 of the grid, 0.61 against 0.65 ms for 1/2.  "auto" cuts
# by columns whenever layout CF applies and every rank gets at least 4 columns.
SLAB_AXIS_DEFAULT = os.environ.get("SDP_SLAB_AXIS", "auto")
# solver.column_pairs = "auto": layout CF with two rows per lane (3 shared-memory reads for 2
# backups instead of 4).  OFF by default: measured on config #5 (profiles/r2_emu_variants_pairs.txt)
# 1.58-1.67 ms per sweep against 1.15 ms with one row per lane.  The reads saved come back as bank
# conflicts: a half-warp's 16 lanes then span up to 32 table rows, and no placement of the rows in
# the 16 eight-byte bank pairs serves both the stride-2 pattern of the interior of the grid and the
# stride <= 1 patterns of the clipped control boxes without collisions.
COLUMN_PAIRS_DEFAULT = os.environ.get("SDP_COLUMN_PAIRS", "0") != "0"


def _torch():
    import torch
    return torch


def partition_by_weight(weights, world):
    """Cut range(len(weights)) into `world` contiguous slabs of nearly equal
    total weight.  Returns the world+1 boundaries (monotone, first 0, last n)."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    bounds = [0]
    if n == 0:
        return [0] * (world + 1)
    csum = np.cumsum(w)
    total = csum[-1]
    for r in range(1, world):
        target = total * r / world
        b = int(np.searchsorted(csum, target, side="left")) + 1
        # pick the nearer of the two candidate cuts
        if b - 1 > bounds[-1] and abs(csum[b - 2] - target) <= abs(csum[b - 1] - target):
            b -= 1
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return bounds


def rebalance_bounds(U_all, bounds, times, tolerance=1.03):
    """Re-cut contiguous slabs from measured slab times: the weight U(x)+1 of every
    state is scaled by its slab's time per unit weight (a piecewise-constant cost
    density), then the grid is cut into slabs of equal estimated time.  Returns the
    new boundaries, or None if the slabs are balanced within `tolerance` (slowest /
    mean), a time is missing, or nothing would move."""
    t = np.asarray(times, dtype=float)
    world = len(t)
    old = [int(b) for b in bounds]
    if world < 2 or not np.all(t > 0) or t.max() <= tolerance * t.mean():
        return None
    w = np.asarray(U_all, dtype=np.float64) + 1.0
    for r in range(world):
        sl = slice(old[r], old[r + 1])
        tot = w[sl].sum()
        if tot > 0:
            w[sl] *= t[r] / tot
    new = [int(b) for b in partition_by_weight(w, world)]
    return None if new == old else new


def make_items(unit_U, chunk, unit_off, per_entry, g_unit_off, g_per_entry, Upad):
    """Work-item table (SdpItem records) + first item of every unit.

    A unit (a state in layout A, a tile of 32 states in layout B) with unit_U controls is
    cut into ceil(unit_U / chunk) runs of EQUAL length (a multiple of 4, at most `chunk`):
    140 controls with chunk 128 become 72 + 68, not 128 + 12 - a short run costs a warp
    the same prologue (item, w-part, first row) as a long one.  Control u of a unit sits
    `u * per_entry` table entries after `unit_off[unit]` (g: `u * g_per_entry` after
    `g_unit_off[unit]`); `Upad` is layout A's row pitch per unit (None for layout B)."""
    unit_U = np.asarray(unit_U, dtype=np.int64)
    units = len(unit_U)
    n_it = (unit_U + chunk - 1) // chunk
    per_unit = (((unit_U + np.maximum(n_it, 1) - 1) // np.maximum(n_it, 1)) + 3) // 4 * 4
    item_begin = np.zeros(units + 1, dtype=np.int64)
    np.cumsum(n_it, out=item_begin[1:])
    n_items = int(item_begin[-1])
    st = np.repeat(np.arange(units, dtype=np.int64), n_it)
    kk = np.arange(n_items, dtype=np.int64) - item_begin[st]
    per = per_unit[st]
    items = np.zeros(n_items, dtype=_cabi.ITEM_DTYPE)
    items["u_begin"] = kk * per
    items["u_count"] = np.minimum(per, unit_U[st] - kk * per)
    items["state"] = st
    assert n_items == 0 or int(items["u_count"].min()) >= 1
    items["entry_base"] = np.asarray(unit_off)[st] + kk * per * per_entry
    items["g_base"] = np.asarray(g_unit_off)[st] + kk * per * g_per_entry
    items["Upad"] = 0 if Upad is None else np.asarray(Upad)[st]
    return items, item_begin


def fill_c_tables(T):
    """the SdpTables record (include/sdp_b200.h) of a SweepTables"""
    c = _cabi.SdpTables()
    c.cell = T.cell.data_ptr()
    c.lam = T.lam.data_ptr()
    c.lam_plane = T.lam_plane
    c.g = T.g.data_ptr()
    c.g_per_w = T.g_per_w
    c.W = T.W
    c.expect = T.expect
    if T.u_mask:
        c.layout = _cabi.LAYOUT_STATE_MINOR_FACTORED if T.tiled else _cabi.LAYOUT_CONTROL_MINOR_FACTORED
        if T.column:
            c.layout = _cabi.LAYOUT_COLUMN_FACTORED
            c.n_cols, c.tiles_per_col = T.n_cols, T.tiles_per_col
            c.seg_begin, c.n_segs = T.seg_begin.data_ptr(), T.n_segs
            c.col_table = T.col_table.data_ptr()
            c.col_table_ready = 0
            if T.item_order is not None:
                # several bands: the whole-list launch walks the bands of a column back to back
                c.item_order = T.item_order.data_ptr()
                c.run_end = T.run_end_ord.data_ptr()
            else:
                c.run_end = T.run_end.data_ptr()
            if getattr(T, "pairs", False):
                c.col_pairs = 1
                c.pos_row = T.pos_row.data_ptr()
        c.u_mask = T.u_mask
        c.cell_w = T.cell_w.data_ptr()
        c.lam_w = T.lam_w.data_ptr()
        c.lam_w_plane = T.lam_w_plane
    else:
        c.layout = _cabi.LAYOUT_STATE_MINOR if T.tiled else _cabi.LAYOUT_CONTROL_MINOR
    c.p = T.p.data_ptr()
    c.p_host = T.p_host.ctypes.data
    c.items = T.items.data_ptr()
    c.n_items = T.n_items
    c.item_begin = T.item_begin.data_ptr()
    c.n_states = T.n_states
    c.U = T.U_dev.data_ptr()
    return c


# Work items (warps) wanted per sweep launch.  A B200 keeps 148 SMs x 12..24 warps of these
# kernels resident, so a slab of a multi-GPU run (config #5 cut in 8: ~4 000 tiles of ~200
# controls) is only 2-3 waves of equally long items and its time is quantised by whole waves:
# measured per slab (profiles/r1_slab_chunks.txt) 0.373..0.449 ms with one item per tile,
# 0.384..0.390 ms with runs of <= 64 controls (about 9 waves) - 13 % on the slowest slab,
# which is the one the whole sweep waits for.  Runs are cut evenly (see the item table), so
# shorter runs cost nothing measurable on a grid that is already long (2.843 vs 2.853 ms).
ITEMS_TARGET = 148 * 96


def pick_item_chunk(unit_U, min_chunk, max_chunk=512, target=ITEMS_TARGET):
    """largest power-of-two-scaled chunk in [min_chunk, max_chunk] giving at least
    `target` work items for units with `unit_U` controls each"""
    unit_U = np.asarray(unit_U, dtype=np.int64)
    chunk = max_chunk
    while chunk > min_chunk and int(((unit_U + chunk - 1) // chunk).sum()) < target:
        chunk //= 2
    return max(chunk, min_chunk)


class _ColumnsNotApplicable(Exception):
    """the cut by columns was chosen by "auto" but layout CF turned out not to apply to the
    built tables: build_sweep_tables starts again with slabs of rows"""


class ColumnHoistRefused(Exception):
    """layout CF was demanded (column_hoist = 'on') for tables whose (x,w) part varies
    along a column of the grid"""


def pair_positions(r0, r1, pair_ok):
    """Positions of the rows [r0, r1) of a band for the two-rows-per-lane sweep: rows r, r+1 with
    pair_ok[r] share a lane (positions 2j, 2j+1), any other row gets a lane of its own with a
    padding position (-1) beside it; padded with -1 to a multiple of 64 (whole pairs of tiles).
    Returns the int64 array position -> row."""
    out = []
    r = r0
    while r < r1:
        if r + 1 < r1 and pair_ok[r]:
            out += [r, r + 1]
            r += 2
        else:
            out += [r, -1]
            r += 1
    out += [-1] * (-len(out) % 64)
    return np.asarray(out, dtype=np.int64)


def column_order(n_states, n_cols, band_rows=None, pair_ok=None):
    """Position order of layout CF for a slab of whole rows of state axis 0.

    The slab's local states are i = row*n_cols + col (C-order).  Layout CF walks them
    band by band (`band_rows`: row boundaries [0, ..., n_rows]; default one band), inside a
    band column by column, every column of a band padded to whole tiles of 32 rows.
    Returns (order, valid, band_tiles, band_tile_begin, tile_col, pos_row):
      order[p]  local state at position p (a padding position repeats the last row of
                its column in the band), valid[p] False on padding positions;
      band_tiles[b] tiles per column in band b; band_tile_begin[b] its first tile;
      tile_col[t] the column of tile t;
      pos_row   None, or - `pair_ok` given: two rows per lane, see pair_positions - per band the
                int64 array position (of every column) -> row of the band (row - r0), -1 = padding."""
    n_rows = n_states // n_cols
    assert n_rows * n_cols == n_states and n_rows >= 1
    if band_rows is None:
        band_rows = [0, n_rows]
    band_rows = [int(r) for r in band_rows]
    assert band_rows[0] == 0 and band_rows[-1] == n_rows and all(a < b for a, b in zip(band_rows, band_rows[1:]))
    cols = np.arange(n_cols, dtype=np.int64)
    orders, valids, band_tiles, band_tile_begin, tile_col, pos_row = [], [], [], [0], [], []
    for r0, r1 in zip(band_rows[:-1], band_rows[1:]):
        if pair_ok is None:
            tpc = (r1 - r0 + 31) // 32
            row = r0 + np.arange(32 * tpc, dtype=np.int64)
            valid_row = row < r1
            row = np.minimum(row, r1 - 1)
        else:
            pr = pair_positions(r0, r1, pair_ok)
            tpc = len(pr) // 32
            valid_row = pr >= 0
            # (a padding position repeats the nearest real row before it: any real state does)
            row = np.maximum.accumulate(np.where(valid_row, pr, r0))
            pos_row.append(np.where(valid_row, pr - r0, -1))
        orders.append((row[None, :] * n_cols + cols[:, None]).reshape(-1))
        valids.append(np.broadcast_to(valid_row[None, :], (n_cols, len(row))).reshape(-1))
        band_tiles.append(tpc)
        band_tile_begin.append(band_tile_begin[-1] + n_cols * tpc)
        tile_col.append(np.repeat(cols, tpc))
    return (np.concatenate(orders), np.concatenate(valids), band_tiles, band_tile_begin,
            np.concatenate(tile_col), pos_row if pair_ok is not None else None)


def item_run_ends(run_key):
    """run_end[i] = index one past the last item of the run of equal consecutive keys
    containing item i (layout CF: the items of one band and column)"""
    key = np.asarray(run_key)
    n = len(key)
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.concatenate([np.flatnonzero(key[1:] != key[:-1]) + 1, [n]]).astype(np.int64)
    return ends[np.searchsorted(ends, np.arange(n), side="right")]


def column_segments(item_u_count, n_ctas):
    """Layout CF: cut the item list (ordered by tile, hence by column) into at most
    `n_ctas` contiguous runs of equal weight, one per CTA - one CTA per SM, since the
    column table fills its shared memory - so that a CTA meets few column changes.
    Weight of an item: its controls plus a fixed cost.  Returns int64 [n_segs + 1]."""
    n_items = len(item_u_count)
    n_segs = max(1, min(int(n_ctas), n_items))
    seg = partition_by_weight(np.asarray(item_u_count, dtype=np.float64) + 2.0, n_segs)
    return np.asarray(seg, dtype=np.int64)


def row_aligned(bounds, n_cols):
    """slab boundaries moved to the nearest multiple of n_cols (whole rows of axis 0),
    kept monotone"""
    out = [int(bounds[0])]
    for b in bounds[1:-1]:
        out.append(max(out[-1], int(round(float(b) / n_cols)) * n_cols))
    out.append(int(bounds[-1]))
    return [min(b, out[-1]) for b in out]


class _DeviceBoundLib(object):
    """The C ABI launches on the calling thread's CURRENT device (it takes raw pointers and a
    stream, and never calls cudaSetDevice).  An Engine made for `device` must therefore make that
    device current around every entry point, or DPSolver(sys, device='cuda:1') would launch on
    GPU 0 against GPU 1 pointers.  `torch.cuda.device` is a no-op when the device is already
    current (one C++ call)."""

    def __init__(self, lib, device):
        self._lib = lib
        self._guard = _torch().cuda.device(device)
        self._wrapped = {}

    def __getattr__(self, name):
        fn = self._wrapped.get(name)
        if fn is None:
            raw = getattr(self._lib, name)
            if not name.startswith("sdp_") or name in ("sdp_version", "sdp_last_error", "sdp_launch_count",
                                                       "sdp_last_kernel", "sdp_set_option"):
                return raw
            guard = self._guard

            def fn(*args):
                with guard:
                    return raw(*args)
            self._wrapped[name] = fn
        return fn


class Collective(object):
    """Thin wrapper over torch.distributed for the per-sweep exchange.
    Works with NCCL (CUDA tensors, on the current stream) and gloo (CPU tensors,
    used by the world_size-2 host tests)."""

    def __init__(self, group=None):
        torch = _torch()
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.world = dist.get_world_size(group)
            self.rank = dist.get_rank(group)
        else:
            self.world = 1
            self.rank = 0
        self._plans = {}

    def all_gather_object(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def plan(self, bounds, device):
        """index plan for gathering uneven contiguous slabs through one padded
        all_gather_into_tensor"""
        torch = _torch()
        key = (tuple(bounds), str(device))
        if key not in self._plans:
            counts = np.diff(np.asarray(bounds))
            maxc = int(counts.max()) if len(counts) else 0
            idx = np.concatenate([r * maxc + np.arange(c) for r, c in enumerate(counts)]) \
                if maxc > 0 else np.zeros(0, dtype=np.int64)
            self._plans[key] = (maxc, torch.from_numpy(idx.astype(np.int64)).to(device))
        return self._plans[key]

    def all_gather_slabs(self, local, bounds, out=None):
        """local: 1-D tensor holding this rank's slab [bounds[rank], bounds[rank+1]).
        Returns the concatenation over ranks (length bounds[-1])."""
        torch = _torch()
        if self.world == 1:
            if out is not None:
                out.copy_(local)
                return out
            return local
        maxc, index = self.plan(bounds, local.device)
        pad_local = torch.zeros(maxc, dtype=local.dtype, device=local.device)
        pad_local[:local.numel()] = local
        pad_full = torch.empty(maxc * self.world, dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(pad_full, pad_local, group=self.group)
        if out is not None:
            torch.index_select(pad_full, 0, index, out=out)
            return out
        return pad_full.index_select(0, index)

    def all_gather_indexed(self, local, maxc, index, out=None):
        """all-gather of unequal 1-D pieces followed by one index_select: element g of the
        result is element index[g] of the rank-major padded concatenation (rank r's piece at
        [r*maxc, r*maxc + len)).  Used when the pieces are not contiguous slabs of the
        result (grid sharded by columns)."""
        torch = _torch()
        if self.world == 1:
            res = local.index_select(0, index)
            if out is not None:
                out.copy_(res)
                return out
            return res
        pad_local = torch.zeros(maxc, dtype=local.dtype, device=local.device)
        pad_local[:local.numel()] = local
        pad_full = torch.empty(maxc * self.world, dtype=local.dtype, device=local.device)
        self.dist.all_gather_into_tensor(pad_full, pad_local, group=self.group)
        if out is not None:
            torch.index_select(pad_full, 0, index, out=out)
            return out
        return pad_full.index_select(0, index)

    def all_reduce_max(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.group)


class PeerExchange(object):
    """Symmetric (peer-mapped) J double buffer + flag words for the fused
    combine + all-gather of the sweep (`sdp_sweep_finalize_p2p`): every rank's
    combine kernel stores its slab of the new J straight into every rank's
    buffer over NVLink and publishes an epoch; consumers wait on local flags.

    Buffer protocol (ping-pong): a sweep reads J[k] and writes J[1-k] on all
    ranks; the per-sweep wait guarantees that a rank can only be one sweep ahead
    of its peers, so the buffer it writes is never one a peer still reads.
    `barrier()` must separate public calls (a peer may still be copying the last
    result out of the buffer the next call will write)."""

    def __init__(self, engine, n_grid):
        import torch
        import torch.distributed._symmetric_memory as symm
        coll = engine.coll
        self.engine = engine
        self.n_grid = n_grid
        world, rank = coll.world, coll.rank
        if world > _cabi.SDP_MAX_PEERS:
            raise ValueError("peer exchange supports at most %d ranks" % _cabi.SDP_MAX_PEERS)
        group = coll.group if coll.group is not None else coll.dist.group.WORLD
        n_pad = (n_grid + 31) // 32 * 32
        # [J0 | J1 | flags (world u64, padded to 32 words)]
        self.buf = symm.empty(2 * n_pad + 32, dtype=torch.float64, device=engine.device)
        self.buf.zero_()
        self.hdl = symm.rendezvous(self.buf, group)
        torch.cuda.synchronize(engine.device)
        self.hdl.barrier()                      # zeros visible everywhere before first use
        self.n_pad = n_pad
        self._views = {}
        self.J = [self.buf[:n_grid], self.buf[n_pad:n_pad + n_grid]]
        self.local = torch.zeros(4, dtype=torch.int64, device=engine.device)   # [epoch, done]
        # every rank's full-grid argmin, filled by the peers' combine kernels like J (SdpPeers.A)
        self.abuf = symm.empty(n_pad, dtype=torch.int32, device=engine.device)
        self.abuf.zero_()
        self.ahdl = symm.rendezvous(self.abuf, group)
        torch.cuda.synchronize(engine.device)
        self.ahdl.barrier()
        self.argmin = self.abuf[:n_grid]
        aptrs = [int(p) for p in self.ahdl.buffer_ptrs]
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.peers = []
        self.peers_J_only = []      # the same exchanges without the argmin (sweeps of a device-resident loop)
        for with_argmin in (True, False):
            for k in range(2):
                P = _cabi.SdpPeers()
                P.world, P.rank = world, rank
                for r in range(world):
                    P.J[r] = ptrs[r] + 8 * k * n_pad
                    P.flags[r] = ptrs[r] + 8 * 2 * n_pad
                    P.A[r] = aptrs[r] if with_argmin else None
                P.epoch = self.local.data_ptr()
                P.done = self.local.data_ptr() + 8
                (self.peers if with_argmin else self.peers_J_only).append(P)

    def peer_view(self, r, k):
        """rank r's copy of J buffer k as a tensor on this device (peer-mapped)"""
        import torch
        key = (r, k)
        if key not in self._views:
            self._views[key] = self.hdl.get_buffer(r, (self.n_grid,), torch.float64, k * self.n_pad)
        return self._views[key]

    def index_of(self, t):
        """0/1 when `t` is one of the two J buffers, else None"""
        for k in range(2):
            if t.data_ptr() == self.J[k].data_ptr() and t.numel() == self.n_grid:
                return k
        return None

    def barrier(self):
        eng = self.engine
        rc = eng.lib.sdp_p2p_barrier(ctypes.byref(self.peers[0]), eng.stream)
        _cabi.check(rc, "sdp_p2p_barrier")


class SweepTables(object):
    """Dense (cell, lam, g) tables of one slab of states, resident in HBM,
    plus the host-side control discretisation needed to turn argmin indices
    back into control values."""

    def __init__(self):
        self.grid = None           # _cabi.SdpGrid
        self.d = 0
        self.W = 1
        self.expect = 1
        self.g_per_w = 0
        self.bounds = None         # slab boundaries over ranks (world+1)
        self.state_begin = 0
        self.n_states = 0
        self.host_full = None      # HostStateTable of ALL states (replicated)
        self.cell = self.lam = self.g = self.p = None
        self.p_host = None
        self.items = self.item_begin = None
        self.part_val = self.part_idx = None
        self.J_out = self.argmin = None
        self.lam_plane = 0
        self.n_items = 0
        self.n_entries = 0
        self.n_backups_local = 0   # admissible (x,u,w) triples in this slab
        self.n_backups_total = 0
        self.c_tables = None       # _cabi.SdpTables
        self.tiled = False         # layout B (state-minor) when True
        self.u_mask = 0            # factored layouts: coordinates of the (x,u) part; 0 = dense
        self.cell_w = self.lam_w = None
        self.lam_w_plane = 0
        self.U_dev = None
        self.tabulate_mode = None
        self.setup_seconds = 0.0
        self.item_chunk = 0        # controls per work item used for these tables
        self.item_begin_host = self.unit_U_host = None
        self.chunk_plan = None     # see Engine._chunk_plan
        # layout CF (column-shared hoist): BF tables over column-major tiles, see column_order()
        self.column = False
        self.n_cols = self.tiles_per_col = 0
        self.seg_begin = None      # device int64 [n_segs+1]: item range of every CTA
        self.n_segs = 0
        self.item_u_count_host = None
        self.col_table = None      # device fp64 scratch: the column tables of the current sweep
        self.run_end = None        # device int64 [n_items]: end of every item's (band, column) run
        self.pairs = False         # layout CF with two rows per lane (SdpTables.col_pairs)
        self.pos_row = self.work_dev = self.work_host = None
        self.item_order = None     # several bands: device int64 [n_items], items column by column (all bands)
        self.run_end_ord = None    # ... and the end of every position's column run in that order
        self.bands = None          # dict(rows, tiles, tile_begin, tile_col), see column_order
        self.band_views = None     # per band: (SdpTables view for the combine pass, first state, states)
        self.sm_count = 148
        # grid sharded by COLUMNS (layout CF, solver.slab_axis = "columns"): this rank holds the
        # columns [col_bounds[rank], col_bounds[rank+1]) of every row; local state row*n_cols + lc
        self.col_bounds = None
        self.gather_index = None   # device int64 [n_grid]: see Collective.all_gather_indexed
        self.gather_maxc = 0
        self.slab_times_ms = None  # measured per-rank sweep times (several ranks, see _measured_bounds)
        self.slab_recut = False    # True when those times moved the slab boundaries

    @property
    def algorithmic_bytes_per_backup(self):
        """4 + 8 d + 8 kappa  """
        kappa = 1.0 if self.g_per_w else 1.0 / self.W
        return 4.0 + 8.0 * self.d + 8.0 * kappa

    @property
    def factored(self):
        return self.u_mask != 0

    @property
    def layout_name(self):
        if self.column:
            return "column_factored"
        return ("state_minor" if self.tiled else "control_minor") + ("_factored" if self.u_mask else "")

    @property
    def device_bytes(self):
        n = 0
        for t in (self.cell, self.lam, self.g, self.items, self.item_begin, self.cell_w, self.lam_w):
            if t is not None:
                n += t.numel() * t.element_size()
        return n

    @property
    def streamed_bytes_per_backup(self):
        """bytes of table the sweep kernel actually reads per admissible (x,u,w)"""
        return self.device_bytes / max(self.n_backups_local, 1)


class PolicyTables(object):
    """[w][n_states] planes for the fixed-policy backup (eval_policy)."""

    def __init__(self):
        self.grid = None
        self.W = 1
        self.g_per_w = 0
        self.cell = self.lam = self.g = self.p = None
        self.ref_cell = self.ref_lam = self.ref_g = None   # reference state's entries (several ranks)
        self.lam_plane = 0
        self.state_begin = 0
        self.n_states = 0
        self.bounds = None


class Engine(object):
    """Owns the device, the stream, the process group and the launches."""

    def __init__(self, device=None, group=None, item_chunk=None, _test_lib=None):
        torch = _torch()
        self.coll = Collective(group)
        # controls per work item (one warp each).  None = adaptive: 512, halved while the
        # slab yields fewer warps than a few waves of the machine (small grids, or
        # one slab of a grid cut over 8 GPUs), see build_sweep_tables
        self.item_chunk_auto = not item_chunk
        self.item_chunk = int(item_chunk) if item_chunk else 512
        if self.item_chunk % 4:
            raise ValueError("item_chunk must be a multiple of 4")
        if _test_lib is not None:
            # TEST SEAM ONLY (tests/fake_lib.py): a numpy model of the C ABI used to
            # exercise the host logic (descriptors, items, slabs, collectives) on a
            # machine without a GPU.  Never set by the package itself.
            self.lib = _test_lib
            self.device = torch.device("cpu")
            self._cuda = False
            self._peer = {}
            self._grid_cache = {}
            self._scan_cache = None
            return
        self._peer = {}                          # n_grid -> PeerExchange | None
        self._grid_cache = {}
        self._scan_cache = None
        self.lib = _cabi.load_library()          # raises if the extension is missing
        if not torch.cuda.is_available():
            raise _cabi.SdpLibraryError(
                "no CUDA device visible: stodynprog_b200 has no CPU path "
                "(the CPU oracle lives under oracle/ and is test infrastructure only)")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._cuda = True
        self.lib = _DeviceBoundLib(self.lib, self.device)

    # -- helpers ----------------------------------------------------------
    @property
    def stream(self):
        torch = _torch()
        if not self._cuda:
            return ctypes.c_void_p(0)
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def torch_stream(self):
        """the launching stream (torch's current stream of THIS engine's device)"""
        return _torch().cuda.current_stream(self.device)

    def _ptr(self, t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)

    def to_device(self, a, dtype=None):
        torch = _torch()
        a = np.ascontiguousarray(a)
        if not a.flags.writeable:
            a = a.copy()
        t = torch.from_numpy(a)
        if dtype is not None:
            t = t.to(dtype)
        if self._cuda and t.numel() * t.element_size() >= (1 << 20):
            # large inputs go through a pinned staging buffer (async DMA)
            pin = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            pin.copy_(t)
            return pin.to(self.device, non_blocking=True)
        return t.to(self.device, non_blocking=False)

    _TORCH_DTYPES = {"float64": "float64", "int32": "int32", "int64": "int64", "uint8": "uint8"}

    def to_device_packed(self, arrays):
        """Several host arrays -> ONE page-locked staging buffer -> one asynchronous
        H2D copy; returns the device views (same shapes and dtypes).  Table builds
        upload a dozen small arrays; one by one each is a synchronous pageable copy
        (a stream synchronisation apiece), which dominates small or time-dependent
        problems that rebuild their tables at every instant."""
        torch = _torch()
        arrs = []
        for a in arrays:
            a = np.ascontiguousarray(a)
            if a.dtype.name not in self._TORCH_DTYPES:
                a = a.view(np.uint8).reshape(-1)          # structured records travel as bytes
            arrs.append(a)
        if not self._cuda:
            return [torch.from_numpy(a.copy()) for a in arrs]
        offs, total = [], 0
        for a in arrs:
            total = (total + 15) // 16 * 16
            offs.append(total)
            total += a.nbytes
        pin = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=True)
        pn = pin.numpy()
        for a, o in zip(arrs, offs):
            if a.nbytes:
                pn[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        buf = pin.to(self.device, non_blocking=True)
        outs = []
        for a, o in zip(arrs, offs):
            t = buf[o:o + a.nbytes].view(getattr(torch, self._TORCH_DTYPES[a.dtype.name]))
            outs.append(t.reshape(a.shape))
        return outs

    def to_device_concat(self, records, parts):
        """a structured record array + a LIST of fp64 arrays -> one page-locked buffer -> one
        asynchronous H2D copy; returns (records as device bytes, the parts concatenated without gaps
        as one device fp64 tensor).  Each part is copied once, straight into the upload buffer."""
        torch = _torch()
        rec = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
        n_parts = int(sum(a.size for a in parts))
        off = (rec.nbytes + 15) // 16 * 16
        total = off + 8 * n_parts
        host = torch.empty(max(total, 16), dtype=torch.uint8, pin_memory=self._cuda)
        hn = host.numpy()
        hn[:rec.nbytes] = rec
        dst = hn[off:off + 8 * n_parts].view(np.float64)
        pos = 0
        for a in parts:
            dst[pos:pos + a.size] = a.reshape(-1)
            pos += a.size
        buf = host.to(self.device, non_blocking=True) if self._cuda else host
        return buf[:rec.nbytes], buf[off:off + 8 * n_parts].view(torch.float64)

    def sync(self):
        if self._cuda:
            _torch().cuda.synchronize(self.device)

    # -- J buffers / peer exchange ------------------------------------------
    def peer_exchange(self, n_grid):
        """PeerExchange for grids of n_grid points, or None (one rank, CPU test
        seam, SDP_P2P=0, or symmetric memory not available: the NCCL all-gather
        path is used instead)."""
        if n_grid not in self._peer:
            px = None
            if self._cuda and self.coll.world > 1 and os.environ.get("SDP_P2P", "1") != "0":
                try:
                    px = PeerExchange(self, n_grid)
                except Exception as e:            # no P2P / symmetric memory on this box
                    import warnings
                    warnings.warn("peer-memory exchange unavailable (%s: %s); using NCCL all-gather"
                                  % (type(e).__name__, e))
                    px = None
                # every rank must take the same path
                ok = self.coll.all_gather_object(px is not None)
                if not all(ok):
                    px = None
            self._peer[n_grid] = px
        return self._peer[n_grid]

    def host_share(self, n_grid, nc):
        """HostShare for results of [n_grid] states x nc controls, or None: one rank, no peer
        exchange (the argmin must travel with J), SDP_HOST_SHARE=0, or shared memory /
        cudaHostRegister not available - same answer on every rank"""
        key = ("host_share", n_grid, nc)
        if key not in self._peer:
            hs = None
            if (self.coll.world > 1 and self.peer_exchange(n_grid) is not None
                    and os.environ.get("SDP_HOST_SHARE", "1") != "0"):
                from .hostshare import HostShare
                err = None
                try:
                    hs = HostShare(self.coll, n_grid, nc, self._cuda)
                except Exception as e:          # every rank must still take part in the vote below
                    err = e
                ok = self.coll.all_gather_object(hs is not None)
                if not all(ok):
                    if err is not None:
                        import warnings
                        warnings.warn("shared host results unavailable (%s: %s)" % (type(err).__name__, err))
                    hs = None
            self._peer[key] = hs
        return self._peer[key]

    def sweep_shared(self, hs, T, J_next, rel_ref_index=None, want_results=True, while_waiting=None):
        """One sweep with host arrays in and out through the shared page-locked segment `hs`:
        every rank uploads 1/N of J_next and hands it to its peers over NVLink, sweeps its shard
        (the combine kernel stores J AND the argmin into every rank's buffers), maps 1/N of the
        policy to control values and copies 1/N of (J, pol) into the shared result slot over its
        own PCIe link.  Returns (J, pol, J_ref) as views of the slot (None, None, J_ref when
        `want_results` is False), or None when no slot is free (the caller takes the private-
        copy path; same decision on every rank)."""
        import time
        torch = _torch()
        world, rank = self.coll.world, self.coll.rank
        n_grid, nc = hs.n_grid, hs.nc
        px = self.peer_exchange(n_grid)
        tm = [time.perf_counter()]
        # 1. where is the input, which slot takes the output (rank 0's array decides)
        src = -2
        if rank == 0:
            src = hs.slot_of(J_next)
            if src is None:
                hs.stage_J().numpy()[:] = np.asarray(J_next, dtype=np.float64).reshape(-1)
                src = -1
        hs.publish(0, src if rank == 0 else 0)
        hs.publish(1, hs.live_mask())
        hs.barrier()
        src = hs.read(0, 0)
        live = 0
        for r in range(world):
            live |= hs.read(r, 1)
        out = next((s_ for s_ in range(hs.n_slots) if s_ != src and not (live >> s_) & 1), None)
        if out is None:
            return None
        J_src = hs.stage_J() if src < 0 else hs.slot_J(src)
        tm.append(time.perf_counter())
        # 2. upload 1/N, hand it to the peers
        # (no begin_call barrier: every rank passed the host barrier above after synchronising its
        # previous call, so nobody still touches the J buffers)
        J_prev, J_new = self.J_pair(n_grid)
        self.flush_exchange()
        a, b = n_grid * rank // world, n_grid * (rank + 1) // world
        k = px.index_of(J_prev)
        J_prev[a:b].copy_(J_src[a:b], non_blocking=True)
        rc = self.lib.sdp_p2p_broadcast(self._ptr(J_prev), a, b - a, ctypes.byref(px.peers_J_only[k]), self.stream)
        _cabi.check(rc, "sdp_p2p_broadcast")
        rc = self.lib.sdp_p2p_wait(ctypes.byref(px.peers_J_only[k]), self.stream)
        _cabi.check(rc, "sdp_p2p_wait")
        tm.append(time.perf_counter())
        # 3. the sweep; J_new and the argmin are complete on every rank after the flag wait
        ref_out = torch.zeros(1, dtype=torch.float64, device=self.device) if rel_ref_index is not None else None
        self.sweep(T, J_prev, J_new, rel_ref_index=rel_ref_index, ref_out=ref_out)
        tm.append(time.perf_counter())
        # 4. 1/N of the results -> the shared slot
        pol = torch.empty((b - a, max(nc, 1)), dtype=torch.float64, device=self.device)
        if nc and b > a:
            rc = self.lib.sdp_policy_values(
                b - a, nc, ctypes.c_void_p(T.lo_dev.data_ptr() + 8 * nc * a),
                ctypes.c_void_p(T.hi_dev.data_ptr() + 8 * nc * a),
                ctypes.c_void_p(T.npts_dev.data_ptr() + 4 * nc * a),
                ctypes.c_void_p(px.argmin.data_ptr() + 4 * a), self._ptr(pol), self.stream)
            _cabi.check(rc, "sdp_policy_values")
        hs.slot_J(out)[a:b].copy_(J_new[a:b], non_blocking=True)
        if nc:
            hs.slot_pol(out)[a * nc:b * nc].copy_(pol.view(-1)[:(b - a) * nc], non_blocking=True)
        ref_host = None
        if ref_out is not None:
            ref_host = self.host_result_buffer((1,), torch.float64)
            ref_host.copy_(ref_out, non_blocking=True)
        tm.append(time.perf_counter())
        if while_waiting is not None:
            while_waiting()
        tm.append(time.perf_counter())
        self.torch_stream.synchronize() if self._cuda else None
        tm.append(time.perf_counter())
        hs.barrier()
        tm.append(time.perf_counter())
        # host-side seconds: [agree on slots, enqueue upload + hand-over, enqueue sweep, enqueue result
        # copies, cache check, wait for the GPU, host barrier] (diagnostics, bench --shared-timing)
        self.last_shared_timing = [b_ - a_ for a_, b_ in zip(tm[:-1], tm[1:])]
        J_ref = float(ref_host[0]) if ref_host is not None else None
        if not want_results:
            return None, None, J_ref
        return hs, out, J_ref

    def J_pair(self, n_grid):
        """two device fp64 [n_grid] buffers to ping-pong sweeps between; with several
        ranks they live in symmetric memory so that the sweep's combine kernel can
        store new values directly into every rank's copy"""
        torch = _torch()
        px = self.peer_exchange(n_grid)
        if px is not None:
            return px.J[0], px.J[1]
        return (torch.empty(n_grid, dtype=torch.float64, device=self.device),
                torch.empty(n_grid, dtype=torch.float64, device=self.device))

    def begin_call(self, n_grid):
        """separate two public calls that reuse the symmetric J buffers"""
        self.flush_exchange()
        px = self._peer.get(n_grid)
        if px is not None:
            px.barrier()

    def share_J(self, J, root=0):
        """make rank `root`'s device copy of J (one of the J_pair buffers) the copy of every
        rank: peer-memory stores over NVLink + flag barrier, or an NCCL broadcast.
        Collective: every rank calls it."""
        if self.coll.world == 1:
            return
        self.flush_exchange()
        px = self._peer.get(J.numel())
        k = px.index_of(J) if px is not None else None
        if k is None:
            self.coll.dist.broadcast(J, src=root, group=self.coll.group)
            return
        if self.coll.rank == root:
            for r in range(self.coll.world):
                if r != root:
                    px.peer_view(r, k).copy_(J, non_blocking=True)
        px.barrier()

    def upload_J(self, J_host, dst):
        """host fp64 array -> device buffer `dst`, asynchronously on the current stream.

        The arrays this package returns live in page-locked memory (`to_host`), so
        the usual loop `J, pol = solver.value_iteration(J)` hands back a buffer the
        DMA engine can read directly: no staging copy.  Any other array is staged
        through a pinned buffer in a few chunks, so that the H2D of one chunk
        overlaps the host copy of the next.  Every public call synchronises the
        stream before returning, so the caller's array is never read after that."""
        torch = _torch()
        a = np.ascontiguousarray(np.asarray(J_host, dtype=np.float64).reshape(-1))
        if not self._cuda:
            dst.copy_(torch.from_numpy(a))
            return dst
        src = None
        if a.flags.writeable:
            t = torch.from_numpy(a)
            if t.is_pinned():
                src = t
        if src is not None:
            dst.copy_(src, non_blocking=True)
            return dst
        n = a.size
        pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
        pin_np = pin.numpy()
        step = max((n + 3) // 4, 1 << 16)
        for b0 in range(0, n, step):
            b1 = min(b0 + step, n)
            pin_np[b0:b1] = a[b0:b1]
            dst[b0:b1].copy_(pin[b0:b1], non_blocking=True)
        return dst

    # -- slab balance -----------------------------------------------------
    REBALANCE_MIN_BACKUPS = 500 * 1000 * 1000    # "auto": only sweeps long enough to matter
    REBALANCE_TOLERANCE = 1.03                   # slowest / mean slab time that is left alone
    REBALANCE_SKEW = None                        # test hook (tests/multi_gpu_check.py)

    def _measured_bounds(self, T, U_all):
        """Slab boundaries equalising the MEASURED sweep time of the ranks.

        Cutting the grid by the number of admissible controls balances the backups,
        not the time: the cost of a backup depends on where its corners fall (L1 / L2
        hit rates differ between regions of the state space) and on the GPU (measured
        on 8 B200s, config #5: 0.375 .. 0.454 ms for slabs of equal backup count).
        Every rank times the streaming kernel on its slab; the per-state weights of a
        slab are scaled by its measured time per weight and the grid is cut again.
        Returns the new boundaries, or None when the slabs are already balanced.
        Same data on every rank (all-gathered), so every rank takes the same decision;
        results do not depend on the partition."""
        torch = _torch()
        coll = self.coll
        n_grid = len(U_all)
        J = torch.zeros(n_grid, dtype=torch.float64, device=self.device)
        t_mine = 0.0
        if T.n_states > 0 and T.n_items > 0:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(7)]
            for a, b in evs:
                a.record(self.torch_stream)
                rc = self.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables),
                                                 self._ptr(J), self._ptr(T.part_val),
                                                 self._ptr(T.part_idx), self.stream)
                _cabi.check(rc, "sdp_sweep_partials")
                b.record(self.torch_stream)
            self.sync()
            t_mine = float(np.median([a.elapsed_time(b) for a, b in evs[2:]]))
        t = np.asarray(coll.all_gather_object(t_mine), dtype=float)
        if self.REBALANCE_SKEW is not None:      # test hook: pretend some slabs are slower
            t = t * np.asarray(self.REBALANCE_SKEW, dtype=float)[:len(t)]
        T.slab_times_ms = [float(x) for x in t]
        return rebalance_bounds(U_all, T.bounds, t, self.REBALANCE_TOLERANCE)

    # -- sweep tables -----------------------------------------------------
    # host cores for the per-state control_box scan of large grids (one rank; forked workers).
    # Opt-in (SDP_SCAN_PROCS=8): forking a process that holds a CUDA context and BLAS threads is
    # only safe for plain-numpy box functions; the scans above cover the common cases without it
    SCAN_PROCS = int(os.environ.get("SDP_SCAN_PROCS", "1"))
    SCAN_PARALLEL_MIN_STATES = 200 * 1000

    def _scan_boxes(self, solver, t_k, state_grid, n_grid, mode):
        """Pass 1 of a table build: the admissible box and control-grid sizes of every state of the
        grid.  Every rank scans an equal share (one vectorised call; else one call per distinct
        box; else per state, optionally on forked workers), the host table is replicated.
        Returns (HostStateTable of the whole grid, this rank's state tuples if they were made)."""
        sys = solver.sys
        coll = self.coll
        world, rank = coll.world, coll.rank
        nb_control = len(sys.control)
        eq = [n_grid * r // world for r in range(world + 1)]
        mine = None
        # the scan used to be by far the longest host stage: the last one is kept, so that a second
        # layout of the same problem (dense after factored, a re-cut of the slabs) does not repeat
        # it; DPSolver.clear_tables() drops it
        scan_key = (t_k, id(sys.control_box), repr(sorted(sys.params.items())) if sys.params else "",
                    tuple(float(c) for c in solver.control_steps),
                    tuple(g.tobytes() for g in state_grid), world)
        host_full = None
        if self._scan_cache is not None and self._scan_cache[0] == scan_key:
            host_full = self._scan_cache[1]
            # the box function may read globals / closure cells that changed since the scan
            # (the reference calls it afresh in every sweep): re-check three states
            probe = sorted({0, n_grid // 2, n_grid - 1})
            now = tb.scan_control_boxes(sys, solver.control_steps,
                                        [tb.state_tuples_at(state_grid, i, i + 1)[0] for i in probe], t_k)
            if not (np.array_equal(now.lo.view(np.int64), host_full.lo[probe].view(np.int64))
                    and np.array_equal(now.hi.view(np.int64), host_full.hi[probe].view(np.int64))
                    and np.array_equal(now.npts, host_full.npts[probe])):
                host_full = None
        if host_full is None:
            part = None
            if mode != "per_state":
                # one vectorised control_box call, trusted only if sample states agree
                # bit-for-bit with the reference's per-state calls
                part = tb.scan_control_boxes_batched(sys, solver.control_steps, state_grid,
                                                     eq[rank], eq[rank + 1], t_k)
            if part is None and mode != "per_state":
                # box functions that read only some of the state variables (the storage
                # examples): one call per distinct box, checked on sample states
                part = tb.scan_control_boxes_by_axes(sys, solver.control_steps, state_grid,
                                                     eq[rank], eq[rank + 1], t_k)
            if part is None and world == 1 and self.SCAN_PROCS > 1 and n_grid >= self.SCAN_PARALLEL_MIN_STATES:
                # box functions that do not vectorise (np.max((a, b)) on scalars, as in the
                # reference's examples): one call per state, on several host cores
                part = tb.scan_control_boxes_parallel(sys, solver.control_steps, state_grid,
                                                      eq[rank], eq[rank + 1], t_k, self.SCAN_PROCS)
            if part is None:
                mine = tb.state_tuples(state_grid, eq[rank], eq[rank + 1])
                part = tb.scan_control_boxes(sys, solver.control_steps, mine, t_k)
            parts = coll.all_gather_object((part.lo, part.hi, part.npts))
            host_full = tb.HostStateTable(n_grid, nb_control)
            host_full.lo = np.concatenate([p[0] for p in parts], axis=0)
            host_full.hi = np.concatenate([p[1] for p in parts], axis=0)
            host_full.npts = np.concatenate([p[2] for p in parts], axis=0)
            self._scan_cache = (scan_key, host_full)
        return host_full, mine

    def build_sweep_tables(self, solver, t_k=None, reuse=None):
        """Tabulate the user's callables over this rank's shard and build the tables on the
        device (see _build_sweep_tables).  With several ranks and slab_axis "auto" the grid is
        cut by columns when layout CF is expected to apply; if the built tables refuse it (same
        decision on every rank), the build starts again with slabs of rows."""
        try:
            return self._build_sweep_tables(solver, t_k, reuse, None)
        except _ColumnsNotApplicable:
            return self._build_sweep_tables(solver, t_k, None, "rows")

    def _build_sweep_tables(self, solver, t_k, reuse, forced_axis):
        """Tabulate the user's callables over this rank's slab and build the dense
        tables on the device.  `reuse`: a SweepTables whose device buffers are
        recycled when the sizes match (time-dependent recursion).

        Solver knobs read here:
          solver.table_layout   : "auto" | "control_minor" (A) | "state_minor" (B)
          solver.table_compress : "auto" | "off" | "on"  (factored (x,u) + (x,w) tables)
          solver.tabulate       : "auto" | "per_state" | "batched"
        """
        import time
        torch = _torch()
        t0 = time.perf_counter()
        sys = solver.sys
        state_grid = [np.asarray(g, dtype=float) for g in solver.state_grid]
        d = len(state_grid)
        grid = _cabi.make_grid(state_grid)
        for ax in state_grid:
            if len(ax) < 2:
                raise ValueError("every state variable needs at least 2 grid points "
                                 "(the reference's interpolation reads out of bounds and "
                                 "divides 0/0 on a 1-point axis, SURVEY.md App. A.2)")
        n_grid = int(np.prod([len(ax) for ax in state_grid]))
        # several perturbations (a TODO of the reference, stodynprog.py:666,679-683): their product
        # grid is flattened in C order into one axis of W nodes (tabulate.perturb_layout)
        nb_perturb = len(solver.perturb_grid)
        W = tb.perturb_layout(solver.perturb_grid)[2]
        if W > 4096:
            raise ValueError("the product perturbation grid has %d nodes; at most 4096 are supported" % W)
        coll = self.coll
        world, rank = coll.world, coll.rank
        dev = self.device

        # pass 1: control boxes (replicated host table, needed to map argmin -> control values)
        mode = getattr(solver, "tabulate", "auto")
        nb_control = len(sys.control)
        eq = [n_grid * r // world for r in range(world + 1)]
        host_full, mine = self._scan_boxes(solver, t_k, state_grid, n_grid, mode)
        U_all = host_full.U.astype(np.int64)
        if U_all.max(initial=0) >= 2 ** 31 - 4:
            raise ValueError("more than 2^31 control combinations for one state")

        # layout CF (column-shared hoist, include/sdp_b200.h): wanted by solver.column_hoist,
        # possible when the column table fits shared memory; needs slabs of whole rows of
        # state axis 0, a factored split with u_mask == 1 and a w-part that is the same for
        # all the states of a column (checked on the built tables, see build_for)
        n_rows0 = len(state_grid[0])
        n_cols = n_grid // n_rows0
        col_mode = getattr(solver, "column_hoist", "auto")
        if col_mode not in ("auto", "on", "off"):
            raise ValueError("column_hoist must be 'auto', 'on' or 'off'")
        col_wanted = col_mode == "on" or (col_mode == "auto" and COLUMN_HOIST_DEFAULT)
        col_candidate = bool(
            col_wanted and d in (2, 3) and nb_perturb >= 1 and 1 < W <= _cabi.FACTORED_MAX_W_REG
            and 8 * _cabi.column_pitch(n_rows0, W) <= _cabi.COLUMN_MAX_SMEM_BYTES
            and getattr(solver, "table_layout", "auto") in ("auto", "state_minor")
            and getattr(solver, "table_compress", "auto") != "off"
            and n_rows0 >= 32 * world and nb_control <= _cabi.SDP_MAX_C)
        col_refused = [False]       # set when the built w-part turns out to vary along a column
        # layout CF with two rows per lane (SdpTables.col_pairs): solver.column_pairs / SDP_COLUMN_PAIRS
        pair_mode = getattr(solver, "column_pairs", "auto")
        if pair_mode not in ("auto", "on", "off"):
            raise ValueError("column_pairs must be 'auto', 'on' or 'off'")
        pairs = bool(col_candidate and (pair_mode == "on" or (pair_mode == "auto" and COLUMN_PAIRS_DEFAULT))
                     and 8 * _cabi.column_pitch(n_rows0, W, pairs=True) <= _cabi.COLUMN_MAX_SMEM_BYTES)
        # several ranks: slabs of whole rows of axis 0 ("rows"), or - layout CF only - whole
        # columns ("columns": every rank then tabulates and loads the tables of its own columns
        # only, so the per-column costs divide by the number of ranks)
        slab_axis = getattr(solver, "slab_axis", "auto")
        if slab_axis not in ("auto", "rows", "columns"):
            raise ValueError("slab_axis must be 'auto', 'rows' or 'columns'")
        if slab_axis == "auto":
            slab_axis = SLAB_AXIS_DEFAULT
        axis_auto = slab_axis == "auto" or forced_axis is not None
        if forced_axis is not None:
            slab_axis = forced_axis
        elif slab_axis == "auto":
            slab_axis = "columns" if (col_candidate and n_cols >= 4 * world) else "rows"
        by_columns = world > 1 and slab_axis == "columns"
        # developer experiments (scripts/dev_shard_emulation.py): the tables of ONE column shard
        # [c0, c1) of the grid on a single rank, as a rank of a multi-GPU run would hold them
        col_override = getattr(solver, "_col_override", None) if world == 1 else None
        if col_override is not None:
            by_columns = True
        if by_columns and not (col_candidate and n_cols >= world):
            raise ValueError("slab_axis='columns' needs layout CF (column_hoist) and at least one "
                             "grid column per rank")

        def build_for(bounds, reuse):
            """tables of this rank's slab: states [bounds[rank], bounds[rank+1]) of the C-order
            grid, or (by_columns) the columns [bounds[rank], bounds[rank+1]) of every row"""
            if by_columns:
                c0, c1 = bounds[rank], bounds[rank + 1]
                # local state row*(c1-c0) + lc  <->  grid state row*n_cols + c0 + lc
                glob = (np.arange(n_rows0, dtype=np.int64)[:, None] * n_cols
                        + np.arange(c0, c1, dtype=np.int64)[None, :]).reshape(-1)
                sb, se, n = 0, n_grid, len(glob)
                n_cols_loc = c1 - c0
            else:
                sb, se = bounds[rank], bounds[rank + 1]
                n = se - sb
                glob = slice(sb, se)
                n_cols_loc = n_cols
            host = tb.HostStateTable(n, nb_control)
            host.lo, host.hi, host.npts = host_full.lo[glob], host_full.hi[glob], host_full.npts[glob]
            U = U_all[glob]

            # table layout (see include/sdp_b200.h): lane <-> control (A) when states
            # have many controls, lane <-> state (B) when there are many states with
            # few controls each
            layout = getattr(solver, "table_layout", "auto")
            if layout == "auto":
                mean_U = float(U.mean()) if n else 0.0
                layout = "state_minor" if (n >= 32 * 1024 and mean_U <= 1024) else "control_minor"
            tiled = layout == "state_minor"
            w_grid = [np.asarray(g) for g in solver.perturb_grid]

            # factored ("broadcast-compressed") tables when every next-state coordinate
            # depends on (x,u) only or on (x,w) only and g does not depend on w; probed on
            # the slab's state with most controls, then checked on every staged chunk
            compress = getattr(solver, "table_compress", "auto")
            u_mask = 0
            w_cap = _cabi.FACTORED_MAX_W_REG if tiled else _cabi.FACTORED_MAX_W_SMEM
            if (compress != "off" and n > 0 and nb_perturb >= 1 and 1 < W <= w_cap and d in (2, 3)
                    and nb_control <= _cabi.SDP_MAX_C):
                i_probe = int(np.argmax(U))
                # (host row i_probe is grid state glob[i_probe] when the shard is whole columns)
                g_probe = int(glob[i_probe]) if by_columns else sb + i_probe
                x_probe = tb.state_tuples_at(state_grid, g_probe, g_probe + 1)[0]
                u_mask = tb.probe_factor_mask(sys, x_probe, host, i_probe, w_grid, t_k) or 0
            if compress == "on" and not u_mask:
                raise ValueError("table_compress='on' but the system's dyn/cost do not have the "
                                 "(x,u) + (x,w) structure (or d, W are outside the supported range)")
            if world > 1:
                # all ranks must agree (they run the same kernels on the same layout)
                u_mask = min(coll.all_gather_object(u_mask))
            column = bool(col_candidate and not col_refused[0] and tiled and u_mask == 1 and n > 0
                          and (by_columns or (sb % n_cols == 0 and se % n_cols == 0)))
            if world > 1:
                column = bool(min(coll.all_gather_object(column)))
            if by_columns and not column:
                if axis_auto:
                    raise _ColumnsNotApplicable()
                raise ValueError("slab_axis='columns' but layout CF does not apply to these tables")
            if col_mode == "on" and not column:
                raise ValueError("column_hoist='on' but layout CF does not apply: it needs the "
                                 "state-minor layout, a factored (x,u)+(x,w) split with state axis 0 "
                                 "alone following the control, at most 9 perturbation nodes, a w-part "
                                 "that does not depend on axis 0, and order[0]*(W|1)*8 bytes of "
                                 "shared memory")
            pos = {}

            def positions(col):
                """(n_eff, U_eff, host_eff, flat_eff, valid, bands): the slab's states in table
                order - C-order, or for layout CF band by band, column by column, with padding
                (column_order)"""
                if col not in pos:
                    if not col:
                        pos[col] = (n, U, host, None, None, None)
                    else:
                        bands = self._column_bands(U.reshape(n // n_cols_loc, n_cols_loc).sum(axis=1), W)
                        pair_ok = None
                        if pairs:
                            # rows that may share a lane: neighbours on axis 0 with the same control
                            # grid sizes in every column (their backups then read overlapping table
                            # rows); a speed hint only - the kernel handles any pair
                            npts_rc = host.npts.reshape(n // n_cols_loc, n_cols_loc * max(nb_control, 1))
                            pair_ok = np.all(npts_rc[1:] == npts_rc[:-1], axis=1)
                        order, valid, band_tiles, band_tile_begin, tile_col, pos_row = \
                            column_order(n, n_cols_loc, bands, pair_ok)
                        h = tb.HostStateTable(len(order), nb_control)
                        h.lo, h.hi, h.npts = host.lo[order], host.hi[order], host.npts[order]
                        flat = glob[order] if by_columns else sb + order
                        pos[col] = (len(order), np.where(valid, U[order], 0), h, flat, valid,
                                    dict(rows=bands, tiles=band_tiles, tile_begin=band_tile_begin,
                                         tile_col=tile_col, pos_row=pos_row))
                return pos[col]

            def sizes(u_mask, col):
                """entry offsets of the layout: dense tables have W entries per control,
                factored tables one"""
                Wf = 1 if u_mask else W
                n, U = positions(col)[:2]
                if tiled:
                    n_tiles = (n + 31) // 32
                    Upad_t = np.zeros(n_tiles * 32, dtype=np.int64)
                    Upad_t[:n] = U
                    tile_U = Upad_t.reshape(n_tiles, 32).max(axis=1)
                    if col and pairs:
                        # the two tiles of a pair are swept by one warp: same number of controls
                        tile_U = np.repeat(tile_U.reshape(-1, 2).max(axis=1), 2)
                    tile_off = np.zeros(n_tiles + 1, dtype=np.int64)
                    np.cumsum(tile_U * Wf * 32, out=tile_off[1:])
                    return (n_tiles, tile_U, tile_off, int(tile_off[-1]), np.zeros(n + 1, dtype=np.int64),
                            np.zeros(n, dtype=np.int64))
                Upad = (U + 3) // 4 * 4
                entry_off = np.zeros(n + 1, dtype=np.int64)
                np.cumsum(Wf * Upad, out=entry_off[1:])
                return 0, None, None, int(entry_off[-1]), entry_off, Upad

            prev_mode = reuse.tabulate_mode if reuse is not None else None
            T = reuse if (reuse is not None and reuse.W == W and reuse.d == d and reuse.tiled == tiled) \
                else SweepTables()
            T.grid, T.d, T.W = grid, d, W
            T.tiled = tiled
            T.expect = 1 if nb_perturb >= 1 else 0
            T.bounds, T.state_begin, T.n_states = bounds, sb, n
            T.col_bounds = None
            if by_columns and col_override is not None:
                T.bounds, T.state_begin, T.col_bounds = None, 0, [int(b) for b in bounds]
            elif by_columns:
                # the exchange goes by grid position, not by slab: T.bounds stays None
                T.bounds, T.state_begin, T.col_bounds = None, 0, [int(b) for b in bounds]
                widths = np.diff(np.asarray(bounds, dtype=np.int64))
                T.gather_maxc = int(widths.max()) * n_rows0
                cols = np.arange(n_cols, dtype=np.int64)
                owner = np.searchsorted(np.asarray(bounds[1:], dtype=np.int64), cols, side="right")
                lc = cols - np.asarray(bounds, dtype=np.int64)[owner]
                rows = np.arange(n_rows0, dtype=np.int64)[:, None]
                idx = owner[None, :] * T.gather_maxc + rows * widths[owner][None, :] + lc[None, :]
                T.gather_index = self.to_device(idx.reshape(-1))
            T.host_full = host_full
            T.n_backups_local = int(U.sum()) * W
            T.n_backups_total = int(U_all.sum()) * W
            # host copy of the probabilities, kept alive with the tables (SdpTables.p_host)
            T.p_host = tb.joint_proba(solver.perturb_proba) if nb_perturb >= 1 else np.ones(1)
            # (one packed upload for the small per-table arrays)
            small = [T.p_host]
            if nb_control:
                small += [host_full.lo.reshape(-1), host_full.hi.reshape(-1),
                          host_full.npts.astype(np.int32).reshape(-1)]
            small = self.to_device_packed(small)
            T.p = small[0]
            # replicated control discretisation, for the argmin -> control value kernel
            T.lo_dev, T.hi_dev, T.npts_dev = (small[1], small[2], small[3]) if nb_control else (None, None, None)
            T.nb_control = nb_control
            T.tabulate_mode = None

            def ensure(name, numel, dtype):
                """(re)allocate T.<name> only when the size changes (time-dependent
                recursions rebuild same-sized tables at every instant)"""
                t = getattr(T, name)
                if t is None or t.numel() != numel or t.dtype != dtype:
                    setattr(T, name, None)        # release before allocating the new size
                    setattr(T, name, torch.empty(numel, dtype=dtype, device=dev))

            keep_staging = bool(getattr(solver, "_keep_staging", False))
            flushes, chunk_record = [], []

            def build(g_per_w, batched, u_mask, col):
                L = {"u_mask": u_mask}
                n, U, host, flat_eff, valid, _ = positions(col)
                n_tiles, tile_U, tile_off, n_entries, entry_off, Upad = sizes(u_mask, col)
                lam_plane = (n_entries + 3) // 4 * 4
                n_u = bin(u_mask).count("1")
                L.update(n_tiles=n_tiles, tile_U=tile_U, tile_off=tile_off, n_entries=n_entries,
                         entry_off=entry_off, Upad=Upad, lam_plane=lam_plane)
                ensure("cell", max(lam_plane, 4), torch.int32)
                ensure("lam", max(lam_plane, 4) * (n_u if u_mask else d), torch.float64)
                if u_mask:
                    n_wp = (n_tiles * 32 if tiled else n) * W
                    lam_w_plane = (n_wp + 3) // 4 * 4
                    ensure("cell_w", max(lam_w_plane, 4), torch.int32)
                    ensure("lam_w", max(lam_w_plane, 4) * (d - n_u), torch.float64)
                else:
                    T.cell_w = T.lam_w = None
                    lam_w_plane = 0
                L["lam_w_plane"] = lam_w_plane
                tile_g_off = None
                if u_mask:
                    # one g per (x,u) entry, indexed like the u-part
                    g_off, g_len = entry_off, n_entries
                    tile_g_off = tile_off
                elif tiled:
                    if g_per_w:
                        tile_g_off = tile_off
                    else:
                        tile_g_off = np.zeros(n_tiles + 1, dtype=np.int64)
                        np.cumsum(tile_U * 32, out=tile_g_off[1:])
                    g_off = np.zeros(n + 1, dtype=np.int64)
                    g_len = int(tile_g_off[-1])
                else:
                    if g_per_w:
                        g_off = entry_off
                    else:
                        g_off = np.zeros(n + 1, dtype=np.int64)
                        np.cumsum(Upad, out=g_off[1:])
                    g_len = int(g_off[-1])
                if tiled:
                    tile_off_dev, tile_g_off_dev, tile_U_dev = self.to_device_packed(
                        [tile_off, tile_g_off, tile_U.astype(np.int32)])
                L.update(g_off=g_off, tile_g_off=tile_g_off)
                ensure("g", max(g_len, 4), torch.float64)
                done = [0]      # states flushed so far (chunks arrive in order)
                del flushes[:], chunk_record[:]
                chunk_grids = {}

                def flush(desc, staging):
                    if u_mask:
                        tb.check_factorable(desc, d, u_mask)
                    desc_dev, stag_dev = self.to_device_concat(desc, staging)
                    n_staging = int(stag_dev.numel())
                    ns = len(desc)
                    first = done[0]
                    if tiled:
                        assert first % 32 == 0
                        t_first = first // 32
                        nt = (ns + 31) // 32
                        t_off = ctypes.c_void_p(tile_off_dev.data_ptr() + 8 * t_first)
                        t_U = ctypes.c_void_p(tile_U_dev.data_ptr() + 4 * t_first)
                        t_Umax = int(tile_U[t_first:t_first + nt].max())
                    max_Upad = int(desc["Upad"].max()) if not tiled else 0

                    def launch(desc_ptr, stag_ptr, g_ptr):
                        """K0 on this chunk: `desc_ptr` / `stag_ptr` the chunk's descriptors and
                        staged outputs, `g_ptr` the stage-cost table to fill (a recursion whose
                        dynamics ignore the instant calls it again per instant with the
                        descriptors pointing at that instant's cost)"""
                        if tiled and u_mask:
                            w0 = t_first * W * 32
                            rc = self.lib.sdp_build_tables_factored_tiled(
                                ctypes.byref(grid), W, u_mask, ns, desc_ptr, stag_ptr,
                                nt, t_off, t_U, t_Umax, self._ptr(T.cell), self._ptr(T.lam), lam_plane,
                                g_ptr, ctypes.c_void_p(T.cell_w.data_ptr() + 4 * w0),
                                ctypes.c_void_p(T.lam_w.data_ptr() + 8 * w0), lam_w_plane, self.stream)
                            _cabi.check(rc, "sdp_build_tables_factored_tiled")
                        elif tiled:
                            rc = self.lib.sdp_build_tables_tiled(
                                ctypes.byref(grid), W, g_per_w, ns, desc_ptr, stag_ptr,
                                nt, t_off, ctypes.c_void_p(tile_g_off_dev.data_ptr() + 8 * t_first), t_U,
                                t_Umax, self._ptr(T.cell), self._ptr(T.lam), lam_plane, g_ptr,
                                self.stream)
                            _cabi.check(rc, "sdp_build_tables_tiled")
                        elif u_mask:
                            w0 = first * W
                            rc = self.lib.sdp_build_tables_factored(
                                ctypes.byref(grid), W, u_mask, ns, desc_ptr, stag_ptr,
                                self._ptr(T.cell), self._ptr(T.lam), lam_plane, g_ptr,
                                max_Upad, ctypes.c_void_p(T.cell_w.data_ptr() + 4 * w0),
                                ctypes.c_void_p(T.lam_w.data_ptr() + 8 * w0), lam_w_plane, self.stream)
                            _cabi.check(rc, "sdp_build_tables_factored")
                        else:
                            rc = self.lib.sdp_build_tables(ctypes.byref(grid), W, g_per_w, ns,
                                                           desc_ptr, stag_ptr,
                                                           self._ptr(T.cell), self._ptr(T.lam), lam_plane,
                                                           g_ptr, max_Upad, self.stream)
                            _cabi.check(rc, "sdp_build_tables")

                    launch(self._ptr(desc_dev), self._ptr(stag_dev), self._ptr(T.g))
                    # (`keep`: the device arrays the launch closure points into)
                    flushes.append(dict(desc=desc, n_staging=n_staging, stag_dev=stag_dev, launch=launch,
                                        first=first, keep=(desc_dev, tile_off_dev, tile_g_off_dev, tile_U_dev)
                                        if tiled else (desc_dev,)) if keep_staging else None)
                    done[0] += ns
                    # the staging tensors are freed by torch's caching allocator in
                    # stream order, so no synchronisation is needed here

                align = 32 if tiled else 1
                if batched:
                    # time-dependent recursion: the callables are the same at every instant, so
                    # the full bit-for-bit check of the batched evaluation is made at the first
                    # instant and a one-state check afterwards; unchanged control boxes reuse
                    # the chunk's control grids
                    tb.tabulate_states_batched(sys, state_grid, sb, se, host, w_grid, t_k, entry_off,
                                               g_off, Upad, g_per_w, flush, align=align,
                                               verify=1 if prev_mode == "batched" else 8,
                                               # (chunks of a column-ordered grid repeat the same rows of
                                               # control grids: kept for the duration of this build)
                                               grid_cache=self._grid_cache if t_k is not None else chunk_grids,
                                               flat_index=flat_eff, valid=valid,
                                               record=chunk_record if keep_staging else None)
                else:
                    states = mine if (mine is not None and (sb, se) == (eq[rank], eq[rank + 1])) else \
                        tb.state_tuples(state_grid, sb, se)
                    if col:
                        states = [states[i] for i in flat_eff - sb]      # (by_columns: sb = 0, all states)
                    tb.tabulate_states(sys, states, host, w_grid, t_k, entry_off, g_off, Upad,
                                       g_per_w, flush, align=align, valid=valid)
                return L

            # mode / layout resolution: batched evaluation is tried first in "auto"
            # mode and abandoned if it fails or is not bit-identical to the
            # reference's per-state calls on the sample states; factored tables are
            # abandoned for dense ones as soon as one chunk does not fit the split
            g_per_w = T.g_per_w if (reuse is T and not u_mask) else 0
            batched = mode in ("auto", "batched")
            L = None
            while L is None:
                try:
                    col = column and u_mask == 1
                    built = build(g_per_w, batched, u_mask, col)
                    T.tabulate_mode = "batched" if batched else "per_state"
                    if col:
                        # the hoisted table is shared by a column only if the (x,w) part of its
                        # states is the same; checked bit for bit on the built tables
                        ok = self._column_w_part_ok(T, W, n_cols_loc, positions(True)[5], positions(True)[4],
                                                    built["lam_w_plane"])
                        if world > 1:
                            ok = bool(min(coll.all_gather_object(ok)))
                        if not ok:
                            if by_columns and axis_auto and col_mode != "on":
                                raise _ColumnsNotApplicable()
                            if col_mode == "on" or by_columns:
                                raise ColumnHoistRefused("column_hoist='on' but the (x,w) part of the "
                                                         "next state depends on state axis 0")
                            col_refused[0] = True
                            column = False
                            built = None         # rebuild in C-order (layout BF)
                    L = built
                except tb.NotFactorable:
                    if compress == "on" or world > 1:
                        # (with several ranks a silent per-rank fallback would desynchronise the layouts)
                        raise ValueError("dyn/cost outputs do not keep the (x,u) + (x,w) structure "
                                         "seen on the probe state; use solver.table_compress = 'off'")
                    u_mask = 0
                except tb.GDependsOnW:
                    if g_per_w:
                        raise
                    if u_mask:
                        u_mask = 0
                    else:
                        g_per_w = 1      # the cost depends on w: dense g table
                except tb.BatchedMismatch:
                    if mode != "auto":
                        raise
                    batched = False      # not bit-identical to per-state calls
                except ColumnHoistRefused as e:
                    raise ValueError(str(e))
                except _ColumnsNotApplicable:
                    raise
                except Exception:
                    if not (batched and mode == "auto"):
                        raise
                    batched = False      # callables not vectorisable over states
            T.u_mask = u_mask
            T.g_per_w = g_per_w
            col = column and u_mask == 1
            T.column = col
            T.bands = positions(True)[5] if col else None
            T.n_cols, T.tiles_per_col = (n_cols_loc, T.bands["tiles"][0]) if col else (0, 0)
            U_eff = positions(col)[1]
            n_tiles, tile_U, tile_off = L["n_tiles"], L["tile_U"], L["tile_off"]
            entry_off, Upad, g_off, tile_g_off = L["entry_off"], L["Upad"], L["g_off"], L["tile_g_off"]
            lam_plane = L["lam_plane"]
            T.n_entries, T.lam_plane, T.lam_w_plane = L["n_entries"], lam_plane, L["lam_w_plane"]
            Wf = 1 if u_mask else W     # table entries per control

            # work items: one warp per run of at most `item_chunk` controls
            unit_U = tile_U if tiled else U
            chunk = self.item_chunk
            sm_count = torch.cuda.get_device_properties(dev).multi_processor_count if self._cuda else 148
            if self.item_chunk_auto:
                # layout A walks 128 controls per warp iteration, layout B one; layout CF has one
                # CTA per SM whose 16 warps share the items of a few column pieces: several
                # rounds of items per piece keep the warps of a CTA level
                chunk = pick_item_chunk(unit_U, 128 if not tiled else 32,
                                        **({"target": sm_count * 16 * 48} if col else {}))
            T.item_chunk = chunk
            if col and not pairs and self.COLUMN_TAIL_FRACTION > 0 and chunk >= 32:
                # layout CF: the warps of a CTA meet at a barrier at the end of every (band, column)
                # run of tiles and wait for the one that took the last item; the last tiles of every
                # run are cut into runs of half the length, handed out last, so that the wait is
                # half an item shorter (it weighs on the short sweeps of a multi-GPU shard)
                bt, tb0 = T.bands["tiles"], T.bands["tile_begin"]
                chunk_arr = np.full(len(unit_U), chunk, dtype=np.int64)
                for b, tpc in enumerate(bt):
                    tail = int(round(tpc * self.COLUMN_TAIL_FRACTION))
                    if tail > 0:
                        t_in = np.arange(tb0[b + 1] - tb0[b]) % tpc
                        chunk_arr[tb0[b]:tb0[b + 1]][t_in >= tpc - tail] = chunk // 2
                chunk = chunk_arr
            if tiled:
                per_entry = Wf * 32
                g_unit_off = tile_off if (T.g_per_w or u_mask) else tile_g_off
                # (layout CF: the Upad field of an item carries the column of its tile)
                items, item_begin = make_items(unit_U, chunk, tile_off, per_entry, g_unit_off,
                                               per_entry if (T.g_per_w or u_mask) else 32,
                                               T.bands["tile_col"] if col else None)
            else:
                items, item_begin = make_items(unit_U, chunk, entry_off, 1, g_off, 1, Upad)
            n_items = len(items)
            T.n_items = n_items
            T.item_begin_host = item_begin
            T.unit_U_host = np.asarray(unit_U, dtype=np.int64)
            T.chunk_plan = None
            up = [items if n_items else np.zeros(1, dtype=_cabi.ITEM_DTYPE), item_begin,
                  U_eff.astype(np.int32) if len(U_eff) else np.zeros(1, dtype=np.int32)]
            T.item_u_count_host = items["u_count"].copy() if col else None
            T.item_order = T.run_end_ord = T.pos_row = None
            T.pairs = bool(col and pairs)
            T.work_host = None
            if col:
                T.sm_count = sm_count
                # items of one band and column are consecutive (tiles are ordered that way)
                n_bands = len(T.bands["tiles"])
                tile_band = np.repeat(np.arange(n_bands), np.diff(T.bands["tile_begin"]))
                st_of_item = items["state"].astype(np.int64)
                item_col = T.bands["tile_col"][st_of_item]
                # the WORK list: what a warp takes - every item, or (two rows per lane) the items of
                # the first tile of every pair, each carrying the index of the same run of controls
                # in the second tile (item.g_base; the two tiles are cut alike)
                work = np.arange(n_items, dtype=np.int64)
                if T.pairs:
                    n_it = np.diff(item_begin)[st_of_item]
                    first = (st_of_item % 2) == 0        # (tiles per column and band are even)
                    items["g_base"] = np.where(first, work + n_it, -1)
                    work = work[first]
                T.work_host = work
                w_band, w_col, w_cnt = tile_band[st_of_item][work], item_col[work], T.item_u_count_host[work]
                if n_bands > 1 or T.pairs:
                    # the launch over the whole list (device-resident sweeps) walks the work column
                    # by column, the bands of a column back to back: one table load per column
                    by_col = np.argsort(w_col * n_bands + w_band, kind="stable")
                    order = work[by_col]
                    up.append(column_segments(w_cnt[by_col], sm_count * self.COLUMN_SEGS_PER_SM))
                    up.append(item_run_ends(w_band * n_cols_loc + w_col))      # (natural order, per band)
                    up += [order, item_run_ends(w_col[by_col])]
                else:
                    up.append(column_segments(w_cnt, sm_count * self.COLUMN_SEGS_PER_SM))
                    up.append(item_run_ends(w_band * n_cols_loc + w_col))
                T.n_segs = len(up[3]) - 1
                if T.pairs:
                    up += [work, np.concatenate(T.bands["pos_row"]).astype(np.int32)]
                up[0] = items            # (g_base now holds the partners)
            else:
                T.n_segs, T.seg_begin, T.run_end = 0, None, None
            up = self.to_device_packed(up)
            T.items, T.item_begin, T.U_dev = up[0], up[1], up[2]
            if col:
                T.seg_begin, T.run_end = up[3], up[4]
                if n_bands > 1 or T.pairs:
                    T.item_order, T.run_end_ord = up[5], up[6]
                T.work_dev = None
                if T.pairs:
                    T.work_dev, T.pos_row = up[-2], up[-1]
                ensure("col_table", n_cols_loc * _cabi.column_pitch(n_rows0, W, T.pairs), torch.float64)
            else:
                T.col_table = None
            T.band_views = None
            n_part = max(n_items, 1) * (32 if tiled else 1)
            T.part_val = torch.empty(n_part, dtype=torch.float64, device=dev)
            T.part_idx = torch.empty(n_part, dtype=torch.int32, device=dev)
            T.J_out = torch.empty(max(n, 1), dtype=torch.float64, device=dev)
            T.argmin = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
            T.c_tables = fill_c_tables(T)
            # (kept for Engine.recursion_fast: one chunk, one flush, batched evaluation)
            T.build_record = None
            if keep_staging and T.tabulate_mode == "batched" and len(flushes) == 1 and len(chunk_record) == 1 \
                    and not col:
                T.build_record = dict(flushes[0], host=host, w_grid=w_grid, **chunk_record[0])
            return T

        # slabs balanced by admissible controls ...
        if col_override is not None:
            bounds = [int(col_override[0]), int(col_override[1])]
        elif by_columns:
            # whole columns per rank, cut by the admissible controls of the columns
            col_w = (U_all + 1).reshape(n_rows0, n_cols).sum(axis=0)
            bounds = [int(b) for b in partition_by_weight(col_w, world)]
        elif world > 1 and col_candidate:
            # whole rows of axis 0 per rank (layout CF); a row is 1/n_rows0 of the grid
            row_w = (U_all + 1).reshape(n_rows0, n_cols).sum(axis=1)
            bounds = [int(b) * n_cols for b in partition_by_weight(row_w, world)]
        else:
            bounds = partition_by_weight(U_all + 1, world) if world > 1 else [0, n_grid]
        override = getattr(solver, "_slab_override", None)
        if override is not None and world == 1:
            # developer experiments (scripts/dev_slab_chunks.py): the tables of ONE slab of
            # the grid, as a rank of a multi-GPU run would hold them; only the streaming
            # kernel can be run on such tables (J_out covers the slab, not the grid)
            bounds = [int(override[0]), int(override[1])]
        T = build_for(bounds, reuse)
        # ... then, with several ranks, by the measured cost of a backup in each slab
        balance = getattr(solver, "slab_balance", "auto")
        # (layout CF cuts whole rows of axis 0 and keeps the cut by admissible controls unless
        # the measured re-cut is asked for explicitly)
        if world > 1 and self._cuda and balance != "controls" and not by_columns and \
                (balance == "measured" or (T.n_backups_total >= self.REBALANCE_MIN_BACKUPS and not T.column)):
            new_bounds = self._measured_bounds(T, U_all)
            if new_bounds is not None and T.column:
                new_bounds = row_aligned(new_bounds, n_cols)
                if new_bounds == [int(b) for b in T.bounds]:
                    new_bounds = None
            if new_bounds is not None:
                T = build_for(new_bounds, T)
                T.slab_recut = True
        self.sync()
        T.setup_seconds = time.perf_counter() - t0
        return T

    COLUMN_SEGS_PER_SM = int(os.environ.get("SDP_COLUMN_SEGS_PER_SM", "1"))
    # share of the tiles at the end of every (band, column) run whose work items are half as long
    COLUMN_TAIL_FRACTION = float(os.environ.get("SDP_COLUMN_TAIL_FRACTION", "0"))
    # launch shape of a band's sweep (SdpTables.col_launch_hint): 640 threads, round-robin - measured
    # on config #5 with ~10 column pieces per CTA (profiles/r2_emu_variants_bands.txt): 1.256 ms
    # against 1.346 ms for the 768-thread first-come-first-served shape that is best (1.146 against
    # 1.180 ms) when a CTA meets 3-4 long pieces
    BAND_LAUNCH_HINT = 640 | (1 << 16)

    def set_column_segments(self, T, n_ctas):
        """layout CF: re-cut the item list of built tables into `n_ctas` CTA segments
        (developer tuning, scripts/dev_column.py)"""
        assert T.column
        w = T.item_u_count_host[T.work_host]
        if T.item_order is not None:
            w = T.item_u_count_host[T.item_order.cpu().numpy()]
        seg = column_segments(w, n_ctas)
        T.seg_begin = self.to_device_packed([seg])[0]
        T.n_segs = len(seg) - 1
        T.c_tables = fill_c_tables(T)
        T.band_views = T.chunk_plan = None

    def _column_w_part_ok(self, T, W, n_cols, bands, valid, lam_w_plane):
        """layout CF: True when, in every column, the (x,w) part of all real states -
        partial cell index and weights of every perturbation node - is bit-identical to
        that of the column's first row (lane 0 of its first tile in the first band, the
        one the sweep kernels read)"""
        torch = _torch()
        live_all = torch.from_numpy(np.ascontiguousarray(valid)).to(self.device)
        planes = [T.cell_w] + [T.lam_w[j * lam_w_plane:(j + 1) * lam_w_plane].view(torch.int64)
                               for j in range(T.d - 1)]
        for t in planes:
            ref = None
            for tpc, t0, t1 in zip(bands["tiles"], bands["tile_begin"][:-1], bands["tile_begin"][1:]):
                v = t[t0 * W * 32:t1 * W * 32].view(n_cols, tpc, W, 32)
                if ref is None:
                    ref = v[:, :1, :, :1]
                live = live_all[t0 * 32:t1 * 32].view(n_cols, tpc, 1, 32)
                if not bool(((v == ref) | ~live).all().item()):
                    return False
        return True

    # layout CF on one rank: the rows are cut into bands of decreasing size so that the
    # results of a band can travel to the host while the next bands are swept (sweep_to_host)
    # Measured on config #5 (profiles/r1_column_tuning.txt), ms per device-resident sweep /
    # ms per value_iteration call with host arrays: 1 band 1.21 / 1.93, 3 bands (45 / 25 /
    # 30 % of the controls) 1.29 / 1.75, 5 bands (45 / 25 / 15 / 10 / 5 %) 1.75 / 1.72 - a
    # small band gives a CTA only a round or two of items per column table.
    COLUMN_BANDS = os.environ.get("SDP_COLUMN_BANDS", "auto")    # "auto" | number of bands
    COLUMN_BAND_FRACTIONS = (0.45, 0.25, 0.30)

    def _column_bands(self, row_weight, W):
        """row boundaries of the bands of layout CF for a slab whose rows weigh `row_weight`
        (admissible controls per row)"""
        n_rows = len(row_weight)
        mode = self.COLUMN_BANDS
        one = [0, n_rows]
        if self.coll.world > 1:
            return one               # the fused combine + exchange publishes one epoch per sweep
        if mode == "auto":
            big = (self._cuda and self.coll.world == 1
                   and float(np.sum(row_weight)) * W >= self.OVERLAP_MIN_BACKUPS
                   and os.environ.get("SDP_OVERLAP", "1") != "0")
            fractions = self.COLUMN_BAND_FRACTIONS if big else (1.0,)
        else:
            k = max(1, int(mode))
            fractions = self.OVERLAP_FRACTIONS[:k - 1] + (1.0,) if k <= len(self.OVERLAP_FRACTIONS) \
                else tuple([1.0 / k] * k)
        if len(fractions) == 1 or n_rows < 64 * len(fractions):
            return one
        csum = np.cumsum(np.asarray(row_weight, dtype=np.float64))
        cuts, acc = [0], 0.0
        for f in fractions[:-1]:
            acc += f
            r = int(np.searchsorted(csum, acc * csum[-1], side="left")) + 1
            r = (r + 31) // 32 * 32                  # whole tiles: no padding lanes inside the slab
            if cuts[-1] < r < n_rows:
                cuts.append(r)
        return cuts + [n_rows]

    def _band_views(self, T):
        """layout CF: per band (SdpTables view for the combine pass, first state, states)"""
        if T.band_views is None:
            views = []
            B = T.bands
            for b, tpc in enumerate(B["tiles"]):
                v = _cabi.SdpTables.from_buffer_copy(T.c_tables)
                v.item_begin = T.c_tables.item_begin + 8 * int(B["tile_begin"][b])
                s0, s1 = int(B["rows"][b]) * T.n_cols, int(B["rows"][b + 1]) * T.n_cols
                v.n_states = s1 - s0
                v.tiles_per_col = int(tpc)
                if getattr(T, "pairs", False):
                    v.pos_row = T.pos_row.data_ptr() + 4 * 32 * int(sum(B["tiles"][:b]))
                views.append((v, s0, s1 - s0))
            T.band_views = views
        return T.band_views

    def _finalize(self, T, J_out, argmin, stream=None):
        """the combine pass: one launch, or for layout CF one per band"""
        stream = self.stream if stream is None else stream
        if not T.column:
            rc = self.lib.sdp_sweep_finalize(ctypes.byref(T.c_tables), self._ptr(T.part_val),
                                             self._ptr(T.part_idx), self._ptr(J_out), self._ptr(argmin), stream)
            _cabi.check(rc, "sdp_sweep_finalize")
            return
        for v, s0, ns in self._band_views(T):
            rc = self.lib.sdp_sweep_finalize(ctypes.byref(v), self._ptr(T.part_val), self._ptr(T.part_idx),
                                             ctypes.c_void_p(J_out.data_ptr() + 8 * s0),
                                             ctypes.c_void_p(argmin.data_ptr() + 4 * s0), stream)
            _cabi.check(rc, "sdp_sweep_finalize")

    def sweep_local(self, T, J_prev, events=None, J_out=None):
        """Enqueue K1 on this rank's slab.  J_prev: device fp64 [n_grid].
        Results land in `J_out` (default T.J_out) / T.argmin (slab-local).  `events`: optional
        (start, end) torch.cuda.Event pair recorded around the streaming kernel
        alone (bench roofline)."""
        J_out = T.J_out if J_out is None else J_out
        if events is None and not T.column:
            rc = self.lib.sdp_sweep(ctypes.byref(T.grid), ctypes.byref(T.c_tables), self._ptr(J_prev),
                                    self._ptr(T.part_val), self._ptr(T.part_idx),
                                    self._ptr(J_out), self._ptr(T.argmin), self.stream)
            _cabi.check(rc, "sdp_sweep")
            return
        if events is not None:
            events[0].record(self.torch_stream)
        rc = self.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables),
                                         self._ptr(J_prev), self._ptr(T.part_val),
                                         self._ptr(T.part_idx), self.stream)
        _cabi.check(rc, "sdp_sweep_partials")
        if events is not None:
            events[1].record(self.torch_stream)
        self._finalize(T, J_out, T.argmin)

    # Device-resident iterations (solve_value_iteration, the bench loop) may leave the flag wait
    # of a sweep's exchange to the NEXT sweep, whose first kernel then starts with it
    # (sdp_sweep_partials_after): one launch less per sweep.  Every other consumer of J calls
    # flush_exchange() first.  SDP_FOLD_WAIT=0 keeps the separate sdp_p2p_wait launch.
    FOLD_WAIT = os.environ.get("SDP_FOLD_WAIT", "1") != "0"
    _pending_wait = None

    def flush_exchange(self):
        """enqueue the flag wait a sweep(..., defer_wait=True) left pending"""
        pend, self._pending_wait = self._pending_wait, None
        if pend is not None:
            rc = self.lib.sdp_p2p_wait(ctypes.byref(pend), self.stream)
            _cabi.check(rc, "sdp_p2p_wait")

    def sweep(self, T, J_prev, J_new, rel_ref_index=None, ref_out=None, resid_out=None,
              events=None, defer_wait=False, want_argmin=True):
        """One full Bellman sweep: K1 on the slab, all-gather of the J slab into
        J_new (device fp64 [n_grid]), optional relative-DP shift and optional
        sup-norm residual max|J_new - J_prev| (all-reduced).
        `defer_wait`: the caller's next use of J_new is another sweep (or it calls
        flush_exchange() itself): the arrival of the peers' slabs is then awaited by that
        sweep's first kernel.
        `want_argmin` False: an intermediate sweep of a device-resident loop, whose policy
        nobody reads - the combine kernel then sends J alone to the peers (the shard's own
        argmin is still written; gather_argmin falls back to collecting those)."""
        n = T.n_states
        sb = T.state_begin
        px = self._peer.get(J_new.numel()) if self.coll.world > 1 else None
        k_new = px.index_of(J_new) if px is not None else None
        pend, self._pending_wait = self._pending_wait, None
        if k_new is None and pend is not None:
            self._pending_wait = pend
            self.flush_exchange()
            pend = None
        if k_new is not None:
            # fused combine + all-gather: K1, then the combine kernel stores the slab
            # into every rank's J_new over NVLink and publishes the epoch
            if events is not None:
                events[0].record(self.torch_stream)
            if pend is not None:
                rc = self.lib.sdp_sweep_partials_after(ctypes.byref(T.grid), ctypes.byref(T.c_tables),
                                                       self._ptr(J_prev), self._ptr(T.part_val),
                                                       self._ptr(T.part_idx), ctypes.byref(pend), self.stream)
                _cabi.check(rc, "sdp_sweep_partials_after")
            else:
                rc = self.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(T.c_tables),
                                                 self._ptr(J_prev), self._ptr(T.part_val),
                                                 self._ptr(T.part_idx), self.stream)
                _cabi.check(rc, "sdp_sweep_partials")
            if events is not None:
                events[1].record(self.torch_stream)
            P_out = px.peers[k_new] if want_argmin else px.peers_J_only[k_new]
            if T.col_bounds is not None:
                rc = self.lib.sdp_sweep_finalize_p2p_cols(
                    ctypes.byref(T.c_tables), self._ptr(T.part_val), self._ptr(T.part_idx), self._ptr(T.argmin),
                    ctypes.byref(P_out), T.col_bounds[-1], T.col_bounds[self.coll.rank], self.stream)
                _cabi.check(rc, "sdp_sweep_finalize_p2p_cols")
            else:
                rc = self.lib.sdp_sweep_finalize_p2p(ctypes.byref(T.c_tables), self._ptr(T.part_val),
                                                     self._ptr(T.part_idx), self._ptr(T.argmin),
                                                     ctypes.byref(P_out), sb, self.stream)
                _cabi.check(rc, "sdp_sweep_finalize_p2p")
            self._argmin_in_px = bool(want_argmin)
            if defer_wait and self.FOLD_WAIT and rel_ref_index is None and resid_out is None:
                self._pending_wait = px.peers[k_new]
            else:
                rc = self.lib.sdp_p2p_wait(ctypes.byref(px.peers[k_new]), self.stream)
                _cabi.check(rc, "sdp_p2p_wait")
        else:
            self._argmin_in_px = False
            if self.coll.world == 1:
                self.sweep_local(T, J_prev, events, J_out=J_new)     # (one rank: straight into J_new)
            else:
                self.sweep_local(T, J_prev, events)
            if self.coll.world == 1:
                pass
            elif T.col_bounds is not None:
                self.coll.all_gather_indexed(T.J_out[:n], T.gather_maxc, T.gather_index, out=J_new)
            else:
                self.coll.all_gather_slabs(T.J_out[:n], T.bounds, out=J_new)
        if rel_ref_index is not None:
            rc = self.lib.sdp_rel_shift(self._ptr(J_new), J_new.numel(), int(rel_ref_index),
                                        self._ptr(ref_out), self.stream)
            _cabi.check(rc, "sdp_rel_shift")
        if resid_out is not None:
            if T.col_bounds is not None:
                # every rank holds all of J: any partition of the grid does for the residual
                world, rank, n_grid = self.coll.world, self.coll.rank, J_new.numel()
                sb = n_grid * rank // world
                n = n_grid * (rank + 1) // world - sb
            a = J_new[sb:sb + n]
            b = J_prev[sb:sb + n]
            rc = self.lib.sdp_supnorm_diff(self._ptr(a), self._ptr(b), n, self._ptr(resid_out),
                                           self.stream)
            _cabi.check(rc, "sdp_supnorm_diff")
            self.coll.all_reduce_max(resid_out)

    # -- one sweep with host results, D2H overlapped with compute -------------
    OVERLAP_MIN_BACKUPS = 100 * 1000 * 1000     # below this the result copy is not worth hiding
    OVERLAP_MIN_ITEMS = 4096
    OVERLAP_FRACTIONS = (0.45, 0.25, 0.15, 0.10, 0.05)

    def can_overlap_results(self, T):
        if T.column and len(T.bands["tiles"]) < 2:
            return False             # layout CF streams its results band by band
        return (self._cuda and self.coll.world == 1 and T.n_items >= self.OVERLAP_MIN_ITEMS
                and T.n_backups_local >= self.OVERLAP_MIN_BACKUPS
                and os.environ.get("SDP_OVERLAP", "1") != "0")

    def _chunk_plan(self, T):
        """cut the slab's work units (states, or tiles of 32 states) into a few runs of
        decreasing size; each run is swept by its own launches (sub-range views of the
        same tables: shifted item / item_begin pointers), so that its results can
        travel to the host while the next run computes"""
        if T.chunk_plan is not None:
            return T.chunk_plan
        if T.column:
            # one run per band: its own CTA segments over the band's items (absolute item
            # indices, the column tables are tabulated once before the first run)
            plan = []
            tb0 = T.bands["tile_begin"]
            for b, (view, s0, ns) in enumerate(self._band_views(T)):
                # (positions of the work list - every item, or with two rows per lane the items
                # of the first tile of every pair - that belong to the band)
                i0, i1 = (int(np.searchsorted(T.work_host, T.item_begin_host[tb0[k]])) for k in (b, b + 1))
                seg = i0 + column_segments(T.item_u_count_host[T.work_host[i0:i1]],
                                           T.sm_count * self.COLUMN_SEGS_PER_SM)
                seg_dev = self.to_device_packed([seg])[0]
                cp = _cabi.SdpTables.from_buffer_copy(T.c_tables)
                cp.seg_begin, cp.n_segs, cp.col_table_ready = seg_dev.data_ptr(), len(seg) - 1, 1
                # a band is walked in the list's own order: many short column pieces per CTA
                cp.item_order = T.work_dev.data_ptr() if T.work_dev is not None else 0
                cp.run_end = T.run_end.data_ptr()
                cp.col_launch_hint = self.BAND_LAUNCH_HINT
                plan.append(dict(tab_p=cp, tab_f=view, s0=s0, s1=s0 + ns, keep=seg_dev,
                                 pv=ctypes.c_void_p(T.part_val.data_ptr()),
                                 pi=ctypes.c_void_p(T.part_idx.data_ptr())))
            T.chunk_plan = plan
            return plan
        us = 32 if T.tiled else 1
        n_units = len(T.unit_U_host)
        csum = np.cumsum(T.unit_U_host)
        total = float(csum[-1])
        cuts = [0]
        acc = 0.0
        for f in self.OVERLAP_FRACTIONS[:-1]:
            acc += f
            b = int(np.searchsorted(csum, acc * total, side="left")) + 1
            cuts.append(min(max(b, cuts[-1]), n_units))
        cuts.append(n_units)
        plan = []
        width = 32 if T.tiled else 1
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b <= a:
                continue
            i0, i1 = int(T.item_begin_host[a]), int(T.item_begin_host[b])
            s0, s1 = a * us, min(b * us, T.n_states)
            cp = _cabi.SdpTables.from_buffer_copy(T.c_tables)
            cp.items = T.c_tables.items + _cabi.ITEM_DTYPE.itemsize * i0
            cp.n_items = i1 - i0
            cf = _cabi.SdpTables.from_buffer_copy(T.c_tables)
            cf.item_begin = T.c_tables.item_begin + 8 * a
            cf.n_states = s1 - s0
            plan.append(dict(tab_p=cp, tab_f=cf, s0=s0, s1=s1,
                             pv=ctypes.c_void_p(T.part_val.data_ptr() + 8 * width * i0),
                             pi=ctypes.c_void_p(T.part_idx.data_ptr() + 4 * width * i0)))
        T.chunk_plan = plan
        return plan

    def sweep_to_host(self, T, J_prev, J_new, while_waiting=None):
        """One sweep of a single-rank slab, returning (J, pol) as host arrays in
        page-locked memory.  The slab is swept in a few runs on two alternating
        streams (so that the tail of one run is filled by the next); as soon as a run
        is combined and its argmin mapped to control values (K3), a copy stream sends
        that run's J and policy to the host while the following runs still compute.
        Same kernels on sub-ranges of the same tables: bit-identical to `sweep`."""
        torch = _torch()
        dev = self.device
        n, nc = T.n_states, T.nb_control
        if not hasattr(self, "_side"):
            self._side = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
            self._copy = torch.cuda.Stream(dev)
        main = torch.cuda.current_stream(dev)
        pol = torch.empty((n, nc), dtype=torch.float64, device=dev)
        J_pin = self.host_result_buffer((n,), torch.float64)
        pol_pin = self.host_result_buffer((n, nc), torch.float64)
        if T.column:
            rc = self.lib.sdp_column_table(ctypes.byref(T.grid), ctypes.byref(T.c_tables), self._ptr(J_prev),
                                           self.stream)
            _cabi.check(rc, "sdp_column_table")
        ev0 = torch.cuda.Event()
        ev0.record(main)
        for k, ch in enumerate(self._chunk_plan(T)):
            st = self._side[k % 2]
            sp = ctypes.c_void_p(st.cuda_stream)
            st.wait_event(ev0)
            s0, s1 = ch["s0"], ch["s1"]
            rc = self.lib.sdp_sweep_partials(ctypes.byref(T.grid), ctypes.byref(ch["tab_p"]),
                                             self._ptr(J_prev), ch["pv"], ch["pi"], sp)
            _cabi.check(rc, "sdp_sweep_partials")
            rc = self.lib.sdp_sweep_finalize(ctypes.byref(ch["tab_f"]), self._ptr(T.part_val),
                                             self._ptr(T.part_idx),
                                             ctypes.c_void_p(J_new.data_ptr() + 8 * s0),
                                             ctypes.c_void_p(T.argmin.data_ptr() + 4 * s0), sp)
            _cabi.check(rc, "sdp_sweep_finalize")
            if nc:
                rc = self.lib.sdp_policy_values(
                    s1 - s0, nc, ctypes.c_void_p(T.lo_dev.data_ptr() + 8 * nc * s0),
                    ctypes.c_void_p(T.hi_dev.data_ptr() + 8 * nc * s0),
                    ctypes.c_void_p(T.npts_dev.data_ptr() + 4 * nc * s0),
                    ctypes.c_void_p(T.argmin.data_ptr() + 4 * s0),
                    ctypes.c_void_p(pol.data_ptr() + 8 * nc * s0), sp)
                _cabi.check(rc, "sdp_policy_values")
            ev = torch.cuda.Event()
            ev.record(st)
            self._copy.wait_event(ev)
            with torch.cuda.stream(self._copy):
                J_pin[s0:s1].copy_(J_new[s0:s1], non_blocking=True)
                if nc:
                    pol_pin[s0:s1].copy_(pol[s0:s1], non_blocking=True)
        done = torch.cuda.Event()
        done.record(self._copy)
        main.wait_event(done)
        if while_waiting is not None:
            while_waiting()          # host work hidden behind the sweep (the GPU is busy)
        done.synchronize()
        return self.result_array(J_pin), self.result_array(pol_pin)

    def gather_argmin(self, T):
        """full-grid int32 argmin (device) of the last sweep.  With the peer-memory exchange every
        rank already holds it (the combine kernel stores the argmin next to J, SdpPeers.A): valid
        until the next sweep; otherwise gathered over the ranks"""
        if self.coll.world > 1 and T.host_full is not None:
            px = self._peer.get(int(T.host_full.lo.shape[0]))
            if px is not None and getattr(self, "_argmin_in_px", False):
                self.flush_exchange()
                return px.argmin
        if T.col_bounds is not None:
            return self.coll.all_gather_indexed(T.argmin[:T.n_states], T.gather_maxc, T.gather_index)
        return self.coll.all_gather_slabs(T.argmin[:T.n_states], T.bounds)

    def policy_values(self, T, argmin_full):
        """K3: full-grid argmin (device int32 [N]) -> control values (device fp64 [N][nc])"""
        torch = _torch()
        n = argmin_full.numel()
        nc = T.nb_control
        pol = torch.empty((n, nc), dtype=torch.float64, device=self.device)
        if nc:
            rc = self.lib.sdp_policy_values(n, nc, self._ptr(T.lo_dev), self._ptr(T.hi_dev),
                                            self._ptr(T.npts_dev), self._ptr(argmin_full),
                                            self._ptr(pol), self.stream)
            _cabi.check(rc, "sdp_policy_values")
        return pol

    # Result arrays are handed to the caller as numpy views of page-locked buffers (no
    # extra host copy, and the next call can DMA them back without staging).  Page-locked
    # memory is a limited resource and a caller may keep every J_k of a long iteration:
    # above this budget of LIVE result bytes new results go to ordinary pageable memory.
    PINNED_RESULT_BUDGET = int(os.environ.get("SDP_PINNED_RESULT_BUDGET", str(4 << 30)))
    _pinned_live = [0]

    def host_result_buffer(self, shape, dtype):
        """uninitialised host tensor for a result: page-locked while the budget lasts"""
        torch = _torch()
        nbytes = int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size()
        if not self._cuda or Engine._pinned_live[0] + nbytes > self.PINNED_RESULT_BUDGET:
            return torch.empty(shape, dtype=dtype)
        return torch.empty(shape, dtype=dtype, pin_memory=True)

    def result_array(self, t):
        """numpy view of a host result tensor; a page-locked one is counted against the
        budget until the caller drops the array (views made from it keep it alive)"""
        import weakref
        a = t.numpy()
        if self._cuda and t.is_pinned():
            nbytes = a.nbytes
            live = Engine._pinned_live
            live[0] += nbytes

            def release(n=nbytes, live=live):
                live[0] -= n
            weakref.finalize(a, release)
        return a

    def to_host(self, *tensors, while_waiting=None):
        """device tensors -> fresh numpy arrays (through pinned buffers, one sync);
        `while_waiting()` runs on the host between the enqueue and the synchronisation"""
        torch = _torch()
        self.flush_exchange()
        if not self._cuda:
            if while_waiting is not None:
                while_waiting()
            return [t.numpy().copy() for t in tensors]
        outs = [self.host_result_buffer(tuple(t.shape), t.dtype) for t in tensors]
        for o, t in zip(outs, tensors):
            o.copy_(t, non_blocking=True)
        if while_waiting is not None:
            while_waiting()
        torch.cuda.current_stream(self.device).synchronize()
        return [self.result_array(o) for o in outs]

    def _pinned_scratch(self, n_doubles):
        """page-locked fp64 scratch of at least `n_doubles`, kept on the engine (allocating and
        pinning 100+ MB costs as much as copying it); None without CUDA"""
        if not self._cuda:
            return None
        torch = _torch()
        buf = getattr(self, "_pin_scratch", None)
        if buf is None or buf.numel() < n_doubles:
            self._pin_scratch = None
            buf = self._pin_scratch = torch.empty(n_doubles, dtype=torch.float64, pin_memory=True)
        return buf

    # -- time-dependent recursion, fast path ----------------------------------
    RECURSION_MAX_G_BYTES = 8 << 30      # per-instant stage-cost tables kept resident

    def recursion_fast(self, solver, t_ini, t_fin, J_fin, J_out, pol_out):
        """bellman_recursion (stodynprog.py:536-591) for systems whose dynamics and admissible
        controls do not depend on the instant, only the stage cost does - the reference's
        time-dependent examples (examples/01 .../pv_storage_control.py:43-65: the PV production
        enters the cost alone).  The cell / weight tables are then the same at every instant:
        they are built once, the cost is tabulated for all instants up front, the T sweeps
        are enqueued back to back with J and the policy staying on the device, and one copy
        brings (J, pol) of all instants to the host.

        Nothing is assumed: the box scan and the batched dyn evaluation of the first, middle
        and last instant are compared bit for bit, and at every instant dyn / control_box of
        one sample state (another one each time) are compared with the tables and the batched
        cost is checked on sample states against the reference's per-state call.  Any
        difference -> returns None and the caller takes the per-instant path.
        Returns a dict of timings on success."""
        import time
        torch = _torch()
        t0 = time.perf_counter()
        sys = solver.sys
        if self.coll.world > 1 or t_fin - t_ini < 3:
            return None
        t_base = t_fin - 1
        solver._keep_staging = True
        try:
            T = self.build_sweep_tables(solver, t_base)
        finally:
            solver._keep_staging = False
        rec = getattr(T, "build_record", None)
        if rec is None or T.n_states != len(rec["desc"]):
            return None
        d, nc, W = T.d, T.nb_control, T.W
        host, cols, outs0 = rec["host"], rec["cols"], rec["outs"]
        S = T.n_states
        g0 = outs0[-1]
        n_T = t_fin - t_ini
        g_len = T.g.numel()
        if n_T * g_len * 8 > self.RECURSION_MAX_G_BYTES:
            return None
        state_grid = [np.asarray(g, dtype=float) for g in solver.state_grid]
        w_args, w_shape, _ = tb.perturb_layout(rec["w_grid"])
        state_of = lambda i: tuple(c[i] for c in cols)            # noqa: E731

        def same_bits(a, b):
            return a.shape == b.shape and np.array_equal(a.view(np.int64), b.view(np.int64))

        # 1. dyn / control_box at the first and the middle instant against the last one
        for t_chk in sorted({t_ini, (t_ini + t_fin) // 2}):
            if t_chk == t_base:
                continue
            box = tb.scan_control_boxes(sys, solver.control_steps, [state_of(i) for i in range(S)], t_chk)
            if not (same_bits(box.lo, host.lo) and same_bits(box.hi, host.hi) and np.array_equal(box.npts, host.npts)):
                return None
            dyn_t, _ = tb._eval_state_chunk(sys, cols, host.lo, host.hi, host.npts, w_args, t_chk, W,
                                            self._grid_cache, w_shape, only="dyn")
            if not all(same_bits(a, b) for a, b in zip(dyn_t, outs0[:d])):
                return None

        # 2. the stage cost of every instant (one batched call each), checked on sample states
        # (filled in place in page-locked memory kept between recursions: one asynchronous upload)
        g_pin = self._pinned_scratch(n_T * g0.size)
        g_src = (g_pin.numpy() if g_pin is not None else np.empty(n_T * g0.size)).reshape((n_T,) + g0.shape)
        k_base = t_base - t_ini
        g_src[k_base] = g0
        for k in range(n_T):
            t_k = t_ini + k
            if t_k == t_base:
                continue
            i_chk = (7 * k) % S
            x_chk = state_of(i_chk)
            box = tb.scan_control_boxes(sys, solver.control_steps, [x_chk], t_k)
            if not (same_bits(box.lo[0], host.lo[i_chk]) and same_bits(box.hi[0], host.hi[i_chk])
                    and np.array_equal(box.npts[0], host.npts[i_chk])):
                return None
            (g_t,), nmax = tb._eval_state_chunk(sys, cols, host.lo, host.hi, host.npts, w_args, t_k, W,
                                                self._grid_cache, w_shape, only="cost")
            if g_t.shape != g0.shape:
                return None
            full = tuple(int(n) for n in host.npts[i_chk]) + (W,)
            for i_s, what in ((i_chk, None), ((i_chk + S // 2) % S, "cost")):
                compact, cdims, U = tb._eval_one_state(sys, state_of(i_s), host, i_s, w_args, t_k, W, w_shape,
                                                       only=what)
                cdims = tuple(cdims)
                pairs = [(g_t, compact[-1])] if what == "cost" else \
                    list(zip(list(outs0[:d]) + [g_t], compact))
                for a, (ref, u_eff, w_eff) in pairs:
                    row = a[i_s if a.shape[0] > 1 else 0]
                    sl = tuple(slice(0, cdims[c] if row.shape[c] > 1 else 1) for c in range(nc))
                    shape = (cdims if u_eff > 1 else (1,) * nc) + (w_eff,)
                    fullb = cdims + (W,)
                    got = np.ascontiguousarray(np.broadcast_to(row[sl], fullb))
                    want = np.ascontiguousarray(np.broadcast_to(ref.reshape(shape), fullb))
                    if not same_bits(got, want):
                        return None
            g_src[k] = g_t
        t_tab = time.perf_counter()

        # 3. device: staging of instant t_base + the cost of all instants behind it; per-instant
        # descriptors differ only in where the cost is read
        n_stag = rec["n_staging"]
        g_size = g0.size
        stag_all = torch.empty(n_stag + n_T * g_size, dtype=torch.float64, device=self.device)
        stag_all[:n_stag].copy_(rec["stag_dev"])
        if g_pin is not None:
            stag_all[n_stag:].copy_(g_pin[:n_T * g_size], non_blocking=True)
        else:
            stag_all[n_stag:].copy_(self.to_device(g_src.reshape(-1)))
        desc_all = np.repeat(rec["desc"][None, :], n_T, axis=0)
        shift = n_stag + np.arange(n_T, dtype=np.int64) * g_size - rec["g_base"]
        desc_all["src"][:, :, d] += shift[:, None]
        desc_dev = self.to_device_packed([desc_all.reshape(-1)])[0]
        desc_bytes = desc_all.dtype.itemsize * S
        g_all = torch.empty((n_T, g_len), dtype=torch.float64, device=self.device)
        for k in range(n_T):
            rec["launch"](ctypes.c_void_p(desc_dev.data_ptr() + k * desc_bytes), self._ptr(stag_all),
                          ctypes.c_void_p(g_all.data_ptr() + 8 * k * g_len))

        # 4. the T sweeps back to back, J and the policy on the device
        n_grid = S
        J_all = torch.empty((n_T + 1, n_grid), dtype=torch.float64, device=self.device)
        J_all[n_T].copy_(self.to_device(np.ascontiguousarray(np.asarray(J_fin, dtype=np.float64).reshape(-1))))
        arg_all = torch.empty((n_T, n_grid), dtype=torch.int32, device=self.device)
        pol_all = torch.empty((n_T, n_grid, max(nc, 1)), dtype=torch.float64, device=self.device)
        self.sync()
        t_up = time.perf_counter()
        for k in range(n_T - 1, -1, -1):
            v = _cabi.SdpTables.from_buffer_copy(T.c_tables)
            v.g = g_all.data_ptr() + 8 * k * g_len
            rc = self.lib.sdp_sweep(ctypes.byref(T.grid), ctypes.byref(v), self._ptr(J_all[k + 1]),
                                    self._ptr(T.part_val), self._ptr(T.part_idx), self._ptr(J_all[k]),
                                    self._ptr(arg_all[k]), self.stream)
            _cabi.check(rc, "sdp_sweep")
            if nc:
                rc = self.lib.sdp_policy_values(n_grid, nc, self._ptr(T.lo_dev), self._ptr(T.hi_dev),
                                                self._ptr(T.npts_dev), self._ptr(arg_all[k]),
                                                self._ptr(pol_all[k]), self.stream)
                _cabi.check(rc, "sdp_policy_values")
        J_h, pol_h = self.to_host(J_all[:n_T], pol_all)
        t_end = time.perf_counter()
        J_out[...] = J_h.reshape(J_out.shape)
        if nc:
            pol_out[...] = pol_h.reshape(pol_out.shape)
        return {"tabulate_s": t_tab - t0, "upload_build_s": t_up - t_tab, "sweeps_s": t_end - t_up,
                "instants": n_T, "tables": T}

    # -- policy tables ----------------------------------------------------
    def build_policy_tables(self, solver, pol):
        """Evaluate dyn/cost for the fixed policy on the whole grid exactly as
        eval_policy does (one broadcast call, stodynprog.py:731-755), expand on the
        host to dense [d][W][N] coordinates and run the cell search (K0a) for this
        rank's equal-count slab."""
        torch = _torch()
        sys = solver.sys
        state_grid_1d = [np.asarray(g, dtype=float) for g in solver.state_grid]
        state_dims = tuple(len(g) for g in state_grid_1d)
        nb_state = len(state_dims)
        nb_control = len(sys.control)
        grid = _cabi.make_grid(state_grid_1d)
        n_grid = int(np.prod(state_dims))
        solver.perturb_grid[0]                        # IndexError if deterministic, like :726
        w_args, w_shape, W = tb.perturb_layout(solver.perturb_grid)
        w_proba = tb.joint_proba(solver.perturb_proba)
        n_w_axes = len(w_shape)
        state_grid = tuple(np.reshape(g, (1,) * i + (-1,) + (1,) * (nb_state - 1 - i + n_w_axes))
                           for i, g in enumerate(solver.state_grid))
        u_k = [pol[..., i].reshape(state_dims + (1,) * n_w_axes) for i in range(nb_control)]
        args = state_grid + tuple(u_k) + w_args
        x_next = sys.dyn(*args, **sys.params)
        g_k = sys.cost(*args, **sys.params)
        if n_w_axes > 1:
            # one trailing axis: the C-order flattened product of the perturbation axes
            x_next = tuple(tb._fold_w(self._as_full(c, state_dims + w_shape), nb_state, w_shape, "dyn")
                           for c in x_next)
            g_k = tb._fold_w(self._as_full(g_k, state_dims + w_shape), nb_state, w_shape, "cost")
        full = state_dims + (W,)
        world, rank = self.coll.world, self.coll.rank
        bounds = [n_grid * r // world for r in range(world + 1)]
        sb, se = bounds[rank], bounds[rank + 1]
        n = se - sb

        def dense_wn(a, allow_compact_w):
            a = np.asarray(a)
            if a.dtype != np.float64:
                a = a.astype(float)
            if a.ndim > len(full):
                raise ValueError("dyn/cost output of rank %d does not broadcast to %s" % (a.ndim, full))
            a = a.reshape((1,) * (len(full) - a.ndim) + a.shape)
            if allow_compact_w and a.shape[-1] == 1:
                flat = np.broadcast_to(a, state_dims + (1,)).reshape(n_grid)
                return np.ascontiguousarray(flat[sb:se]), 0
            flat = np.broadcast_to(a, full).reshape(n_grid, W)
            return np.ascontiguousarray(flat[sb:se].T), 1     # [W][n]

        P = PolicyTables()
        P.grid, P.W, P.bounds, P.state_begin, P.n_states = grid, W, bounds, sb, n
        coords = np.stack([dense_wn(c, False)[0].reshape(-1) for c in x_next]) if n else \
            np.zeros((nb_state, 0))
        g_arr, g_per_w = dense_wn(g_k, True)
        P.g_per_w = g_per_w
        P.lam_plane = W * n
        s_dev = self.to_device(coords)
        P.cell = torch.empty(max(W * n, 1), dtype=torch.int32, device=self.device)
        P.lam = torch.empty(max(W * n, 1) * nb_state, dtype=torch.float64, device=self.device)
        rc = self.lib.sdp_cell_setup(ctypes.byref(grid), W * n, self._ptr(s_dev), self._ptr(P.cell),
                                     self._ptr(P.lam), self.stream)
        _cabi.check(rc, "sdp_cell_setup")
        P.g = self.to_device(g_arr.reshape(-1)) if n else torch.zeros(1, dtype=torch.float64, device=self.device)
        P.p = self.to_device(w_proba)
        if world > 1:
            # every rank keeps the W table entries of relative DP's reference state (the
            # fused backup + shift kernel recomputes that state's value in every block)
            ref_flat = int(np.ravel_multi_index(solver._state_ref_ind, state_dims))
            rc_coords = np.stack([np.broadcast_to(self._as_full(c, full), full).reshape(n_grid, W)[ref_flat]
                                  for c in x_next])                       # [d][W]
            s_ref = self.to_device(np.ascontiguousarray(rc_coords))
            P.ref_cell = torch.empty(W, dtype=torch.int32, device=self.device)
            P.ref_lam = torch.empty(W * nb_state, dtype=torch.float64, device=self.device)
            rc = self.lib.sdp_cell_setup(ctypes.byref(grid), W, self._ptr(s_ref), self._ptr(P.ref_cell),
                                         self._ptr(P.ref_lam), self.stream)
            _cabi.check(rc, "sdp_cell_setup")
            g_full = self._as_full(g_k, full)
            if g_per_w:
                P.ref_g = self.to_device(np.ascontiguousarray(
                    np.broadcast_to(g_full, full).reshape(n_grid, W)[ref_flat]))
            else:
                P.ref_g = self.to_device(np.ascontiguousarray(
                    np.broadcast_to(g_full, state_dims + (1,)).reshape(n_grid)[ref_flat:ref_flat + 1]))
        return P

    @staticmethod
    def _as_full(a, full):
        """dyn/cost output as an fp64 array of the rank of `full` (leading axes of size 1 added)"""
        a = np.asarray(a)
        if a.dtype != np.float64:
            a = a.astype(float)
        return a.reshape((1,) * (len(full) - a.ndim) + a.shape)

    def policy_eval(self, P, J_a, J_b, n_iter, rel_dp, ref_index, J_ref_hist):
        """n_iter fixed-policy backups, ping-pong between J_a and J_b (device fp64
        [n_grid]).  Returns the tensor holding the final value function."""
        n_grid = J_a.numel()
        self.flush_exchange()
        if self.coll.world == 1:
            rc = self.lib.sdp_policy_eval(ctypes.byref(P.grid), P.W, P.g_per_w, self._ptr(P.p),
                                          self._ptr(P.cell), self._ptr(P.lam), P.lam_plane,
                                          self._ptr(P.g), P.n_states, 0, n_grid,
                                          self._ptr(J_a), self._ptr(J_b), int(n_iter),
                                          1 if rel_dp else 0, int(ref_index),
                                          self._ptr(J_ref_hist) if rel_dp else ctypes.c_void_p(0),
                                          self.stream)
            _cabi.check(rc, "sdp_policy_eval")
            return J_a if n_iter % 2 == 0 else J_b
        cur, nxt = J_a, J_b
        sb, n = P.state_begin, P.n_states
        px = self._peer.get(n_grid)
        if px is not None and px.index_of(J_a) is not None and px.index_of(J_b) is not None:
            # fused backup + all-gather over peer memory: the kernel stores the slab's new
            # values into every rank's buffer and publishes an epoch (no NCCL call)
            for k in range(n_iter):
                k_new = px.index_of(nxt)
                null = ctypes.c_void_p(0)
                rc = self.lib.sdp_policy_eval_p2p(
                    ctypes.byref(P.grid), P.W, P.g_per_w, self._ptr(P.p), self._ptr(P.cell),
                    self._ptr(P.lam), P.lam_plane, self._ptr(P.g), n, sb, n_grid, self._ptr(cur),
                    ctypes.byref(px.peers[k_new]),
                    self._ptr(P.ref_cell) if rel_dp else null, self._ptr(P.ref_lam) if rel_dp else null,
                    self._ptr(P.ref_g) if rel_dp else null,
                    ctypes.c_void_p(J_ref_hist.data_ptr() + 8 * k) if rel_dp else null, self.stream)
                _cabi.check(rc, "sdp_policy_eval_p2p")
                rc = self.lib.sdp_p2p_wait(ctypes.byref(px.peers[k_new]), self.stream)
                _cabi.check(rc, "sdp_p2p_wait")
                cur, nxt = nxt, cur
            return cur
        for k in range(n_iter):
            rc = self.lib.sdp_policy_eval(ctypes.byref(P.grid), P.W, P.g_per_w, self._ptr(P.p),
                                          self._ptr(P.cell), self._ptr(P.lam), P.lam_plane,
                                          self._ptr(P.g), n, sb, n_grid,
                                          self._ptr(cur), self._ptr(nxt), 1, 0, 0,
                                          ctypes.c_void_p(0), self.stream)
            _cabi.check(rc, "sdp_policy_eval")
            self.coll.all_gather_slabs(nxt[sb:sb + n].clone(), P.bounds, out=nxt)
            if rel_dp:
                ref_ptr = ctypes.c_void_p(J_ref_hist.data_ptr() + 8 * k)
                rc = self.lib.sdp_rel_shift(self._ptr(nxt), n_grid, int(ref_index), ref_ptr, self.stream)
                _cabi.check(rc, "sdp_rel_shift")
            cur, nxt = nxt, cur
        return cur

    # -- interpolation ----------------------------------------------------
    def interp(self, grid, values, s):
        """values: host (n_v, n_grid); s: host (d, n_s) -> host (n_v, n_s).
        fp64 or fp32 according to values.dtype."""
        torch = _torch()
        f32 = values.dtype == np.float32
        n_v, n_s = values.shape[0], s.shape[1]
        v_dev = self.to_device(values)
        s_dev = self.to_device(s)
        out = torch.empty((n_v, n_s), dtype=torch.float32 if f32 else torch.float64, device=self.device)
        fn = self.lib.sdp_interp_f32 if f32 else self.lib.sdp_interp
        rc = fn(ctypes.byref(grid), n_v, self._ptr(v_dev), n_s, self._ptr(s_dev), self._ptr(out), self.stream)
        _cabi.check(rc, "sdp_interp")
        return out.cpu().numpy()
