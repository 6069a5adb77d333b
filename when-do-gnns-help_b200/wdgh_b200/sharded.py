"""1-D row-partitioned execution of the hot path over N GPUs (one process per GPU).

Rank r owns a contiguous block of rows of the adjacency (CSR with GLOBAL column ids), the
matching feature rows and labels.  One step =
    all-gather(features, labels, degree scales)            -- NCCL over NVLink (gloo in CPU tests)
    local  A_hat[rows_r, :] X  and local label statistics  -- the same CUDA kernels as on one GPU
    all-reduce(class histograms + counters)                -- SUM (MAX for the bincount length)
There is no other data-path collective: the aggregation output stays row-sharded.

Only `torch.distributed` plumbing and index arithmetic live here; the local compute is injected
through two methods so that the world_size-2 gloo tests can drive the same plumbing on CPU.
"""
from __future__ import annotations

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


class CudaShardedStats(ShardedStats):
    """The product: local compute = the CUDA kernels of libwdgh_b200.so on this rank's GPU."""

    def __init__(self, graph_local, part, rank, x_local, labels32_local, num_classes, group=None, slabs=None):
        from . import graph as G
        self._G = G
        self.g = graph_local
        d = int(x_local.shape[1])
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if slabs is None:
            # Column slabs of the feature matrix (gather of slab s+1 overlapping the aggregation of slab s) are
            # supported but OFF: measured on 2 x B200, 1B-entry graph, d = 128: 1 slab 76.2 ms, 2 slabs 91.2 ms,
            # 4 slabs 130.6 ms per step -- the gather kernel is bound by the number of random row requests, not
            # by their size, so every extra slab costs almost a full aggregation pass.
            slabs = 1
        if d % slabs or (d // slabs) % 32:
            raise ValueError("feature width must split into slabs that are multiples of 32 columns")
        self.slabs = int(slabs)
        if graph_local.row_offset != part.bounds(rank)[0] or graph_local.n != part.rows(rank):
            raise ValueError("graph shard does not match the partition")
        super().__init__(part, rank, x_local, labels32_local, num_classes, group)
        self._scratch = None
        self._y = None
        self._x_full = None
        # slab-major copy of the local feature rows: each slab is one contiguous all-gather payload
        ds = d // self.slabs
        self._x_slabs = ([self.x_local] if self.slabs == 1 else
                         [self.x_local[:, k * ds:(k + 1) * ds].contiguous() for k in range(self.slabs)])

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
        """Same result as ShardedStats.step, with the feature all-gather (the only large transfer:
        (N-1)/N of the feature matrix per rank over NVLink) issued asynchronously so that the label pass,
        which needs only the gathered labels, runs underneath it."""
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
        ds = d // self.slabs
        if self._x_full is None:
            self._x_full = [xs.new_empty((world * xs.shape[0], ds)) for xs in self._x_slabs]
        # all slab gathers are queued back to back on the collective stream ...
        works = [dist.all_gather_into_tensor(buf, xs, group=self.group, async_op=True)
                 for buf, xs in zip(self._x_full, self._x_slabs)]
        # ... the label pass and then the aggregation of slab s run while slab s+1 is still in flight
        self._scratch = G.structure_counts_raw(self.g, labels_full, self.c, self._scratch)
        if self._y is None or self._y.shape[1] != d:
            self._y = torch.empty((self.g.n, d), dtype=torch.float32, device=self.x_local.device)
        for k, (work, buf) in enumerate(zip(works, self._x_full)):
            work.wait()
            G.spmm(self.g, buf, norm, add_self_loop, out=self._y[:, k * ds:(k + 1) * ds], dinv=dinv_full,
                   deg_code=self.code_full)
        counters, node_sum = self.reduce_counters(self._scratch[0], self._scratch[1])
        return self._y, counters, node_sum
