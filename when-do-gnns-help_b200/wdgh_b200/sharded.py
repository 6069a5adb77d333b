"""Partitioned execution of the hot path over N GPUs (one process per GPU).

Nodes are owned 1-D: rank r holds a contiguous block of rows of the adjacency (CSR with GLOBAL column ids), the
matching feature rows and labels.  Two pipelines share that ownership:

* `ShardedStats` / `CudaShardedStats` -- 1-D row partition.  One step =
      all-gather(labels, degree scales)                      -- NCCL over NVLink (gloo in CPU tests)
      features of the other ranks                            -- copy-engine pulls of peer-mapped shards with one
                                                                aggregation phase per arriving shard (or NCCL all-gather)
      local  A_hat[rows_r, :] X  and local label statistics  -- the same CUDA kernels as on one GPU
      all-reduce(class histograms + counters)                -- SUM (MAX for the bincount length)
* `Grid2D` / `Cuda2DShardedStats` -- 2-D (row group x column group) partition for even N: a rank aggregates the block
  A[rows of its row group, columns of its column group], needs only its column group's feature shards, stores the
  foreign row slices straight into their owners' memory from inside the aggregation kernel (NVLink peer stores) and
  finishes its own slice -- reduction over the received slices, self loop, scale -- in its last aggregation phase.

The aggregation output stays row-sharded in both.  Only `torch.distributed` plumbing and index arithmetic live here;
in `ShardedStats` the local compute is injected through two methods so that the gloo tests can drive the same
plumbing on CPU.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


class RowPartition:
    """Equal blocks of ceil(n / world) rows; the last block may be short (buffers are padded)."""

    def __init__(self, n: int, world: int):
        if n < 0 or world < 1:
            raise ValueError("bad partition")
        self.n, self.world = int(n), int(world)
        self.block = (self.n + self.world - 1) // self.world if self.n else 0

    def bounds(self, rank: int):
        r0 = min(rank * self.block, self.n)
        return r0, min(r0 + self.block, self.n)

    def rows(self, rank: int) -> int:
        r0, r1 = self.bounds(rank)
        return r1 - r0

    def owner(self, node):
        return np.asarray(node) // max(self.block, 1)

    @property
    def padded(self) -> int:
        return self.block * self.world


def shard_csr(rowptr, col, val, r0, r1):
    """Rows [r0, r1) of a CSR as (local rowptr rebased to 0, col slice with global ids, val slice)."""
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    local = rowptr[r0:r1 + 1] - rowptr[r0]
    return local, col[e0:e1], (None if val is None else val[e0:e1])


def _all_gather_rows(local_padded, group):
    """[block, ...] per rank -> [world*block, ...] on every rank."""
    world = dist.get_world_size(group)
    out = local_padded.new_empty((world * local_padded.shape[0],) + tuple(local_padded.shape[1:]))
    try:
        dist.all_gather_into_tensor(out, local_padded.contiguous(), group=group)
    except (RuntimeError, NotImplementedError):  # backends without the flat variant
        parts = list(out.chunk(world, dim=0))
        dist.all_gather(parts, local_padded.contiguous(), group=group)
    return out


def _pad_rows(t, rows):
    if t.shape[0] == rows:
        return t
    pad = t.new_zeros((rows - t.shape[0],) + tuple(t.shape[1:]))
    return torch.cat([t, pad], 0)


class ShardedStats:
    """Row-sharded A_hat X + label statistics.  Subclasses provide the local compute."""

    def __init__(self, part: RowPartition, rank: int, x_local, labels_local, num_classes: int, group=None):
        self.part, self.rank, self.group = part, rank, group
        self.c = int(num_classes)
        self.x_local = _pad_rows(x_local, part.block)
        self.labels_local = _pad_rows(labels_local, part.block)

    # -- local compute, overridden -------------------------------------------------------------
    def local_degree_scale(self, norm, add_self_loop):  # -> float32 [rows_r] (+ optional uint8 degree codes)
        raise NotImplementedError

    def local_compute(self, x_full, labels_full, dinv_full, norm, add_self_loop):
        """-> (y_local [rows_r, d], counters int64 [H + 2C + C*C], node_sum float64 [1])"""
        raise NotImplementedError

    # -- the step ------------------------------------------------------------------------------
    def gather_inputs(self, norm, add_self_loop):
        x_full = _all_gather_rows(self.x_local, self.group)
        labels_full = _all_gather_rows(self.labels_local, self.group)
        dinv_full = None
        self.code_full = None
        if norm != _lib.NORM_NONE:
            loc = self.local_degree_scale(norm, add_self_loop)
            dinv, code = loc if isinstance(loc, tuple) else (loc, None)
            dinv_full = _all_gather_rows(_pad_rows(dinv, self.part.block), self.group)
            if code is not None:
                self.code_full = _all_gather_rows(_pad_rows(code, self.part.block), self.group)
        return x_full, labels_full, dinv_full

    def reduce_counters(self, counters, node_sum):
        nb = counters[_lib.SC_NBINS:_lib.SC_NBINS + 1].clone()
        counters[_lib.SC_NBINS] = 0
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(nb, op=dist.ReduceOp.MAX, group=self.group)
        counters[_lib.SC_NBINS] = nb[0]
        dist.all_reduce(node_sum, op=dist.ReduceOp.SUM, group=self.group)
        return counters, node_sum

    def step(self, norm=_lib.NORM_SYM, add_self_loop=True):
        x_full, labels_full, dinv_full = self.gather_inputs(norm, add_self_loop)
        y_local, counters, node_sum = self.local_compute(x_full, labels_full, dinv_full, norm, add_self_loop)
        counters, node_sum = self.reduce_counters(counters, node_sum)
        return y_local, counters, node_sum


def _symmetric_features(x_local, rank, world, block, group):
    """Peer-mapped feature buffer indexed by GLOBAL node id: this rank's shard lives in place at rows
    [rank*block, (rank+1)*block), every other block is NaN until a pull fills it (an aggregation phase that reads a
    block nobody filled shows up as NaN instead of as a plausible number).

    Returns (x_full, handle, why): handle is None when the ranks could not ALL map each other (then x_full is plain
    memory for an NCCL all-gather); the decision is taken collectively so that every rank runs the same path."""
    d = int(x_local.shape[1])
    dev = x_local.device
    ok, why, x_full, hdl = 1, "", None, None
    try:
        import torch.distributed._symmetric_memory as symm_mem
        x_full = symm_mem.empty((world * block, d), dtype=torch.float32, device=dev)
        hdl = symm_mem.rendezvous(x_full, group=group if group is not None else dist.group.WORLD)
        hdl.get_buffer((rank + 1) % world, (world * block, d), torch.float32)
    except Exception as e:  # no peer access on this box
        ok, why = 0, repr(e)
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        hdl = None
        x_full = torch.empty((world * block, d), dtype=torch.float32, device=dev)
        why = why or "a peer rank could not map the symmetric buffer"
    x_full.fill_(float("nan"))
    x_full[rank * block:rank * block + x_local.shape[0]].copy_(x_local)
    if x_local.shape[0] < block:
        x_full[rank * block + x_local.shape[0]:(rank + 1) * block].zero_()
    torch.cuda.synchronize()
    dist.barrier(group=group)
    return x_full, hdl, why


class CudaShardedStats(ShardedStats):
    """The product, 1-D row partition: local compute = the CUDA kernels of libwdgh_b200.so on this rank's GPU.

    phased (default for binary graphs with a row-group width): every rank maps its peers' feature shards (symmetric
    memory over NVLink) and pulls them with the copy engines; the aggregation runs as one `wdgh_spmm_csr_ranged`
    phase per source rank -- local columns straight away, then one `y +=` phase per arriving shard, self loop + scale
    in the last one.  Otherwise: NCCL all-gather of the features, then one aggregation launch."""

    def __init__(self, graph_local, part, rank, x_local, labels32_local, num_classes, group=None, phased=None):
        from . import graph as G
        self._G = G
        self.g = graph_local
        d = int(x_local.shape[1])
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        ok_width = (d % 4 == 0) and (d >= 128 or d in (32, 64)) and graph_local.val is None
        self.phased = bool(ok_width and world > 1) if phased is None else bool(phased)
        self._seg = None
        self._skip = None
        if graph_local.row_offset != part.bounds(rank)[0] or graph_local.n != part.rows(rank):
            raise ValueError("graph shard does not match the partition")
        super().__init__(part, rank, x_local, labels32_local, num_classes, group)
        self._scratch = None
        self._y = None
        self._x_full = None
        self._hdl = None
        self._peers = None
        self._why_no_peers = ""
        if self.phased:
            blk = part.block
            self._x_full, self._hdl, self._why_no_peers = _symmetric_features(self.x_local, rank, world, blk, group)
            self.x_local = self._x_full[rank * blk:(rank + 1) * blk]   # the shard lives in place from now on
            if self._hdl is not None:
                self._peers = [self._hdl.get_buffer(b, (world * blk, d), torch.float32)[b * blk:(b + 1) * blk]
                               for b in range(world)]
                self._copy_stream = torch.cuda.Stream()
                self._events = [torch.cuda.Event() for _ in range(world)]

    def local_degree_scale(self, norm, add_self_loop):
        self.g._dinv.clear()  # recomputed every step: it is part of the timed path
        dinv, _, code = self.g.degree_scale(norm, add_self_loop)
        return dinv, code

    def local_compute(self, x_full, labels_full, dinv_full, norm, add_self_loop):
        G = self._G
        if self._y is None or self._y.shape[1] != x_full.shape[1]:
            self._y = torch.empty((self.g.n, x_full.shape[1]), dtype=torch.float32, device=x_full.device)
        y, self._scratch = G.spmm_structure_fused(self.g, x_full, labels_full, self.c, norm, add_self_loop,
                                                  out=self._y, dinv=dinv_full, deg_code=self.code_full,
                                                  scratch=self._scratch)
        return y, self._scratch[0], self._scratch[1]

    def step(self, norm=_lib.NORM_SYM, add_self_loop=True):
        """Same result as ShardedStats.step, with the feature transfer (the only large one: (N-1)/N of the feature
        matrix per rank over NVLink) issued asynchronously so that the label pass, which needs only the gathered
        labels, runs underneath it."""
        G = self._G
        labels_full = _all_gather_rows(self.labels_local, self.group)
        dinv_full = None
        self.code_full = None
        if norm != _lib.NORM_NONE:
            dinv, code = self.local_degree_scale(norm, add_self_loop)
            dinv_full = _all_gather_rows(_pad_rows(dinv, self.part.block), self.group)
            if code is not None:
                self.code_full = _all_gather_rows(_pad_rows(code, self.part.block), self.group)
        world = dist.get_world_size(self.group)
        d = int(self.x_local.shape[1])
        if self.phased:
            return self._step_phased(labels_full, dinv_full, norm, add_self_loop, world, d)
        if self._x_full is None:
            self._x_full = self.x_local.new_empty((world * self.x_local.shape[0], d))
        work = dist.all_gather_into_tensor(self._x_full, self.x_local, group=self.group, async_op=True)
        self._scratch = G.structure_counts_raw(self.g, labels_full, self.c, self._scratch)   # under the all-gather
        if self._y is None or self._y.shape[1] != d:
            self._y = torch.empty((self.g.n, d), dtype=torch.float32, device=self.x_local.device)
        work.wait()
        G.spmm(self.g, self._x_full, norm, add_self_loop, out=self._y, dinv=dinv_full, deg_code=self.code_full)
        counters, node_sum = self.reduce_counters(self._scratch[0], self._scratch[1])
        return self._y, counters, node_sum

    # -- phased step: peer-mapped feature shards, copy-engine pulls, one aggregation phase per arriving shard -----
    def reset_graph(self):
        """The CSR arrays were overwritten in place (same shapes): drop everything derived from them."""
        self.g._plan = None
        self.g._dinv.clear()
        self._seg = None
        self._skip = None

    def _step_phased(self, labels_full, dinv_full, norm, add_self_loop, world, d):
        G = self._G
        g, r = self.g, self.rank
        if self._seg is None:  # per graph, like the plan
            bounds = [b * self.part.block for b in range(world)] + [max(world * self.part.block, g.n_global)]
            self._seg = G.column_segments(g, bounds)
            self._skip = G.heavy_flags(g) if g.n_chunks else None
        x_full, seg = self._x_full, self._seg
        block = self.part.block
        if self._y is None or self._y.shape[1] != d:
            self._y = torch.empty((g.n, d), dtype=torch.float32, device=self.x_local.device)
        cur = torch.cuda.current_stream()
        if self._peers is not None:
            # the pulls start once everything queued so far has finished: the previous step still reads x_full, and the
            # small all-gathers above order this step after the peers' writes to their shards
            self._copy_stream.wait_stream(cur)
            order = [(r - k) % world for k in range(1, world)]
            with torch.cuda.stream(self._copy_stream):
                for src in order:
                    x_full[src * block:(src + 1) * block].copy_(self._peers[src], non_blocking=True)
                    self._events[src].record(self._copy_stream)
        else:
            work = dist.all_gather_into_tensor(x_full, self.x_local, group=self.group, async_op=True)
        # under the transfers: label pass, then the entries whose source node is local (the shard is in place)
        self._scratch = G.structure_counts_raw(g, labels_full, self.c, self._scratch)
        if self._peers is not None:
            G.spmm_ranged(g, seg[r], seg[r + 1], x_full, self._y, norm, add_self_loop, dinv_full, self.code_full,
                          self._skip, accumulate=False, finalize=False, run_split_rows=False)
            for i, src in enumerate(order):  # one phase per shard, in arrival order
                cur.wait_event(self._events[src])
                last = i == len(order) - 1
                G.spmm_ranged(g, seg[src], seg[src + 1], x_full, self._y, norm, add_self_loop, dinv_full,
                              self.code_full, self._skip, accumulate=True, finalize=last, run_split_rows=last)
        else:
            # NCCL writes every block of x_full, the own one included: local columns from the shard meanwhile
            G.spmm_ranged(g, seg[r], seg[r + 1], self.x_local, self._y, norm, add_self_loop, dinv_full,
                          self.code_full, self._skip, accumulate=False, finalize=False, run_split_rows=False,
                          x_row0=g.row_offset)
            work.wait()
            first, last = r == 0, r == world - 1
            if not first:  # columns owned by ranks 0 .. r-1
                G.spmm_ranged(g, seg[0], seg[r], x_full, self._y, norm, add_self_loop, dinv_full, self.code_full,
                              self._skip, accumulate=True, finalize=last, run_split_rows=last)
            if not last:   # columns owned by ranks r+1 .. world-1
                G.spmm_ranged(g, seg[r + 1], seg[world], x_full, self._y, norm, add_self_loop, dinv_full,
                              self.code_full, self._skip, accumulate=True, finalize=True, run_split_rows=True)
        counters, node_sum = self.reduce_counters(self._scratch[0], self._scratch[1])
        return self._y, counters, node_sum


# ===============================================================================================
# 2-D (row group x column group) partition: less NVLink traffic than the 1-D row partition
# ===============================================================================================
class Grid2D:
    """world = pr x pc ranks; rank r = i * pc + j sits in row group i and column group j.

    Nodes stay owned 1-D (rank r owns [r*block, (r+1)*block)).  Row group i = the rows of ranks i*pc .. i*pc+pc-1;
    column group j = the nodes owned by the ranks with (rank % pc) == j.  Rank (i, j) aggregates the block
    A[rows of group i, columns of group j]: it needs only the feature shards of its column group (pr shards
    instead of all pr*pc) and produces partial sums for the rows of its whole row group, which are then reduced
    onto the owner of each row slice."""

    def __init__(self, n: int, world: int, pr: int):
        if world % pr:
            raise ValueError("world must be a multiple of pr")
        self.n, self.world, self.pr, self.pc = int(n), int(world), int(pr), world // pr
        self.part = RowPartition(n, world)

    def coords(self, rank: int):
        return rank // self.pc, rank % self.pc

    def row_group_ranks(self, i: int):
        return [i * self.pc + s for s in range(self.pc)]

    def col_group_ranks(self, j: int):
        return [ii * self.pc + j for ii in range(self.pr)]

    def col_in_group(self, col, j: int):
        """Boolean mask: which column ids belong to column group j."""
        return (col // max(self.part.block, 1)) % self.pc == j

    def filter_slice(self, rowptr, col, j: int):
        """CSR (torch tensors, any device) of one rank's rows -> the same rows restricted to column group j."""
        rows = int(rowptr.shape[0]) - 1
        keep = self.col_in_group(col.to(torch.int64), j)
        rid = torch.repeat_interleave(torch.arange(rows, device=col.device), rowptr[1:] - rowptr[:-1])
        out = torch.zeros(rows + 1, dtype=torch.int64, device=col.device)
        out[1:] = torch.cumsum(torch.bincount(rid[keep], minlength=rows), 0)
        return out, col[keep].contiguous()

    def schedule(self, rank: int):
        """Order in which `rank` aggregates the row slices of its block: [(k, s, owner)], k = 1 .. pc.

        Foreign slices first (slice s = (j + k) % pc is sent to its owner (i, s) and lands in the owner's
        receive slot k), the rank's own slice (k = pc, nothing to send) last."""
        i, j = self.coords(rank)
        return [(k, (j + k) % self.pc, i * self.pc + (j + k) % self.pc) for k in range(1, self.pc + 1)]

    def slot_source(self, rank: int, k: int) -> int:
        """Rank whose partial sums arrive in receive slot k (1 <= k < pc) of `rank`."""
        i, j = self.coords(rank)
        return i * self.pc + (j - k) % self.pc


class Cuda2DShardedStats:
    """A_hat X on the 2-D partition (2 row groups x N/2 column groups) + label statistics on the 1-D row shard.

    Rank (i, j) aggregates the block A[rows of row group i, columns of column group j]: the pc row slices of its row
    group, restricted to the two feature shards of its column group (its own, in place, and the partner's, pulled).
    One step:
      1. all-gather(labels) -- also the step barrier; then the partner shard is PULLED by the copy engines;
      2. under the pull: degree scales, label pass, and the own row slice over the columns of the OWN shard
         (raw partial sums into receive slot 0);
      3. the pc - 1 foreign row slices, full column range: the aggregation kernel stores every finished row straight
         into its owner's receive slot over NVLink (peer-mapped memory) -- compute and transfer are one kernel.  These
         launches are NVLink-bound from pc = 4 on; overlap=True runs them on half of the CTA slots while the own
         slice's partner-shard columns are aggregated next to them on a second stream (`y +=` into slot 0).  Measured
         at N = 8: 26.5 ms against 25.05 ms without (the second pass over slot 0 costs more HBM time than the
         overlap saves), so it is opt-in;
      4. a tiny all-reduce orders the peers' stores before
      5. the last launch on the own slice: slot 0 and the pc - 1 slices received from the peers enter every row's
         entry stream as virtual trailing entries next to the self loop, then the scale is applied -- with overlap the
         column range is empty and the launch is a pure streaming reduction, without it the launch also aggregates the
         partner-shard columns.  No separate reduce kernel, no local copy of the foreign partials;
      6. all-reduce of the class histograms + counters.
    Per rank 1 + (pc - 1) shards cross NVLink in each direction instead of N - 1."""

    def __init__(self, grid: Grid2D, rank: int, slice_graphs, graph_1d, x_local, labels32_local, num_classes, group=None,
                 overlap=None, foreign_ctas_per_sm=16, own_ctas_per_sm=16):
        from . import graph as G
        import torch.distributed._symmetric_memory as symm_mem
        if grid.pr != 2:
            raise ValueError("the 2-D pipeline is written for 2 row groups (one partner shard per rank)")
        self._G, self.grid, self.rank, self.group = G, grid, rank, group
        self.i, self.j = grid.coords(rank)
        self.slices = slice_graphs          # pc CSRGraphs: rows of rank i*pc+s, columns of group j (global ids)
        self.g1d = graph_1d                 # rows of this rank, all columns (label pass, degree scale)
        self.c = int(num_classes)
        part, pc = grid.part, grid.pc
        blk, d = part.block, int(x_local.shape[1])
        dev = x_local.device
        self.labels_local = _pad_rows(labels32_local, blk)
        wgroup = group if group is not None else dist.group.WORLD
        # symmetric (peer-mapped) buffers.  x_full is indexed by global node id; my own shard lives in place (write
        # features through `x_shard`), the partner shard is pulled into its global position, the other blocks are
        # never touched (they stay NaN: reading one would be a bug and shows).
        self.x_full, self._hx, why = _symmetric_features(x_local, rank, grid.world, blk, group)
        if self._hx is None:
            raise RuntimeError(f"the 2-D pipeline needs peer-mapped memory between all ranks: {why}")
        self.x_shard = self.x_full[rank * blk:(rank + 1) * blk]
        # receive slots: [0] = my own slice over my own shard's columns (local), [k] = partial of my rows from the
        # rank k places to the left in my row group (stored by that rank's kernel over NVLink)
        self.recv = symm_mem.empty((pc, blk, d), dtype=torch.float32, device=dev)
        self.recv.fill_(float("nan"))
        self._hr = symm_mem.rendezvous(self.recv, group=wgroup)
        self.partner = (1 - self.i) * pc + self.j
        self._peer_x = self._hx.get_buffer(self.partner, (grid.world * blk, d), torch.float32)[
            self.partner * blk:(self.partner + 1) * blk]
        self._peer_recv = {r: self._hr.get_buffer(r, (pc, blk, d), torch.float32)
                           for r in grid.row_group_ranks(self.i) if r != rank}
        self._copy, self._side = torch.cuda.Stream(), torch.cuda.Stream()
        self._ev_ready, self._ev_x, self._ev_side = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.overlap = bool(overlap) and pc > 1      # None / False: off (see the class docstring for the measurement)
        self._ctas = (int(foreign_ctas_per_sm), int(own_ctas_per_sm))
        self._tiny = torch.zeros(1, dtype=torch.int32, device=dev)
        self._scratch = None
        self._y = torch.empty((self.g1d.n, d), dtype=torch.float32, device=dev)
        self._skip = [G.heavy_flags(g) if g.n_chunks else None for g in self.slices]
        g_own = self.slices[self.j]
        self._seg_own = G.column_segments(g_own, [rank * blk, (rank + 1) * blk])   # [lo, hi) = my shard's columns
        self.stage_ms = None            # WDGH_STAGE_TIMES=1: per-stage device times of the last step
        self._trace = os.environ.get("WDGH_STAGE_TIMES") == "1"
        torch.cuda.synchronize()
        dist.barrier(group=group)

    def _mark(self, marks, name):
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    def step(self, norm=_lib.NORM_SYM, add_self_loop=True):
        G, grid = self._G, self.grid
        part, pc, j, r = grid.part, grid.pc, self.j, self.rank
        blk = part.block
        cur = torch.cuda.current_stream()
        marks = [] if self._trace else None
        self._mark(marks, "start")
        # 1. labels; every peer has now finished its previous step (its reads of the receive slots and of x_full) and
        #    the writes of its shard, so the pull and, later, the stores into the peers' slots may start
        labels_full = _all_gather_rows(self.labels_local, self.group)
        self._ev_ready.record(cur)
        self._copy.wait_event(self._ev_ready)
        with torch.cuda.stream(self._copy):
            self.x_full[self.partner * blk:(self.partner + 1) * blk].copy_(self._peer_x, non_blocking=True)
            self._ev_x.record(self._copy)
        # 2. under the pull
        dinv_full = code_full = None
        if norm != _lib.NORM_NONE:
            self.g1d._dinv.clear()
            dinv, _, code = self.g1d.degree_scale(norm, add_self_loop)
            dinv_full = _all_gather_rows(_pad_rows(dinv, blk), self.group)
            code_full = _all_gather_rows(_pad_rows(code, blk), self.group) if code is not None else None
        self._mark(marks, "small all-gathers")
        self._scratch = G.structure_counts_raw(self.g1d, labels_full, self.c, self._scratch)
        self._mark(marks, "label pass")
        g_own = self.slices[j]
        lo, hi = self._seg_own[0], self._seg_own[1]
        G.spmm_ranged(g_own, lo, hi, self.x_full, self.recv[0], norm, add_self_loop, dinv_full, code_full,
                      self._skip[j], accumulate=False, finalize=False, run_split_rows=False)
        self._mark(marks, "own slice, own columns")
        cur.wait_event(self._ev_x)
        self._mark(marks, "wait pull")
        rb, re = (g_own.rowptr[:-1], lo) if self.i == 1 else (hi, g_own.rowptr[1:])   # the partner shard's columns
        if self.overlap:
            # own slice, partner columns (HBM-bound) next to the NVLink-bound foreign slices, each on part of the SMs
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                G.spmm_ranged(g_own, rb, re, self.x_full, self.recv[0], norm, add_self_loop, dinv_full, code_full,
                              self._skip[j], accumulate=True, finalize=False, run_split_rows=False,
                              ctas_per_sm=self._ctas[1])
                self._ev_side.record(self._side)
        # 3. foreign slices: rows leave for their owner as they are finished
        for k, s, owner in grid.schedule(r)[:-1]:
            g = self.slices[s]
            G.spmm_ranged(g, g.rowptr[:-1], g.rowptr[1:], self.x_full, self._peer_recv[owner][k], norm, add_self_loop,
                          dinv_full, code_full, self._skip[s], accumulate=False, finalize=False, run_split_rows=True,
                          ctas_per_sm=self._ctas[0] if self.overlap else 0)
            self._mark(marks, f"slice {s} -> rank {owner}")
        if self.overlap:
            cur.wait_event(self._ev_side)
            self._mark(marks, "own slice, partner columns (second stream)")
        # 4. every rank's stores have landed once this all-reduce completes (stream-ordered after the kernels)
        if pc > 1:
            dist.all_reduce(self._tiny, group=self.group)
            self._mark(marks, "barrier")
        # 5. slot 0 + received slices + self loop + scale (+ the partner columns when they were not overlapped)
        if self.overlap:
            rb = re = g_own.rowptr[:-1]
        G.spmm_ranged(g_own, rb, re, self.x_full, self._y, norm, add_self_loop, dinv_full, code_full, self._skip[j],
                      accumulate=False, finalize=True, run_split_rows=True, extra=self.recv, extra_split=pc - 1)
        self._mark(marks, "own slice: " + ("" if self.overlap else "partner columns + ") + "reduce + finalize")
        counters, node_sum = ShardedStats.reduce_counters(self, self._scratch[0], self._scratch[1])
        if marks is not None:
            self._mark(marks, "counter all-reduce")
            torch.cuda.synchronize()
            self.stage_ms = [(b[0], a[1].elapsed_time(b[1])) for a, b in zip(marks[:-1], marks[1:])]
        return self._y, counters, node_sum
