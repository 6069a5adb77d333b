"""N>1 host logic on CPU: world_size-2 gloo run of the row-partitioned step.

The partition arithmetic, the all-gather of features / labels / degree scales and the all-reduce
of the counters are the product code (wdgh_b200.sharded.ShardedStats); the local compute is
injected from the oracle (tests may use it), and the result must equal the single-process oracle.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ref_port as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _graph(n=203, seed=5):
    rng = np.random.default_rng(seed)
    deg = rng.integers(0, 9, n)
    deg[7] = 90
    src = np.repeat(np.arange(n), deg)
    dst = rng.integers(0, n, src.shape[0])
    row, col, _ = O.coalesce(src, dst, None, n)
    keep = row != col
    row, col = row[keep], col[keep]
    labels = rng.integers(0, 4, n).astype(np.int64)
    labels[rng.random(n) < 0.1] = -1
    x = rng.standard_normal((n, 12)).astype(np.float32)
    return row, col, labels, x


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wdgh_b200 import _lib
        from wdgh_b200.sharded import RowPartition, ShardedStats, shard_csr

        row, col, labels, x = _graph()
        n, C = labels.shape[0], 4
        part = RowPartition(n, world)
        r0, r1 = part.bounds(rank)
        rowptr = O.csr_from_coo(row, n)
        lrp, lcol, _ = shard_csr(rowptr, col, None, r0, r1)

        class OracleShard(ShardedStats):
            def local_degree_scale(self, norm, add_self_loop):
                deg = np.diff(lrp).astype(np.float64) + (1.0 if add_self_loop else 0.0)
                return torch.from_numpy((1.0 / np.sqrt(deg)).astype(np.float32))

            def local_compute(self, x_full, labels_full, dinv_full, norm, add_self_loop):
                rows = r1 - r0
                lrow = np.repeat(np.arange(rows), np.diff(lrp)) + r0
                xf, dv, lab = x_full.numpy()[:n], dinv_full.numpy()[:n], labels_full.numpy()[:n].astype(np.int64)
                # local rows of D^-1/2 (A+I) D^-1/2 X
                y = np.zeros((rows, xf.shape[1]), np.float32)
                np.add.at(y, lrow - r0, (dv[lcol][:, None] * xf[lcol]))
                y = (y + dv[r0:r1, None] * xf[r0:r1]) * dv[r0:r1, None]
                s = O.structure_counts(lrow, lcol, lab, n, num_classes=C)
                cnt = np.zeros(_lib.sc_words(C), np.int64)
                cnt[_lib.SC_MATCH_ALL], cnt[_lib.SC_MATCH_LAB] = s["match_all"], s["match_lab"]
                cnt[_lib.SC_N_LAB], cnt[_lib.SC_N_SELF] = s["n_lab"], s["n_selfloop"]
                d_loc, m_loc = s["deg_nsl"][r0:r1], s["match_nsl"][r0:r1]
                cnt[_lib.SC_N_EMPTY] = int((np.diff(lrp) == 0).sum())
                nz = np.nonzero(d_loc)[0]
                cnt[_lib.SC_NBINS] = (nz.max() + r0 + 1) if nz.size else 0
                cnt[_lib.SC_N_NODES_NSL] = nz.size
                ll = lab[r0:r1]
                cnt[_lib.SC_HEADER:_lib.SC_HEADER + C] = np.bincount(ll[ll >= 0], minlength=C)
                cnt[_lib.SC_HEADER + C:_lib.SC_HEADER + 2 * C] = np.bincount(ll[ll >= 0], weights=np.diff(lrp)[ll >= 0],
                                                                            minlength=C)
                cnt[_lib.SC_HEADER + 2 * C:_lib.SC_HEADER + 2 * C + C * C] = s["hist"].ravel()
                node_sum = (m_loc[nz].astype(np.float32) / d_loc[nz].astype(np.float32)).astype(np.float64).sum()
                return torch.from_numpy(y), torch.from_numpy(cnt), torch.tensor([node_sum, 0.0], dtype=torch.float64)

        pipe = OracleShard(part, rank, torch.from_numpy(x[r0:r1]), torch.from_numpy(labels[r0:r1].astype(np.int32)), C)
        y, cnt, node_sum = pipe.step(_lib.NORM_SYM, True)
        out.put((rank, r0, r1, y.numpy(), cnt.numpy(), float(node_sum[0])))
    finally:
        dist.destroy_process_group()


def test_row_partition_bounds():
    from wdgh_b200.sharded import RowPartition
    for n, w in ((10, 3), (203, 2), (8, 8), (5, 8), (0, 2), (1000, 7)):
        p = RowPartition(n, w)
        spans = [p.bounds(r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert sum(p.rows(r) for r in range(w)) == n and p.padded >= n
        if n:
            assert all(int(p.owner(v)) == r for r in range(w) for v in range(*p.bounds(r)))


def test_shard_csr_roundtrip():
    from wdgh_b200.sharded import RowPartition, shard_csr
    row, col, _, _ = _graph()
    n = 203
    rowptr = O.csr_from_coo(row, n)
    part = RowPartition(n, 3)
    cols = []
    for r in range(3):
        lrp, lcol, _ = shard_csr(rowptr, col, None, *part.bounds(r))
        assert lrp[0] == 0 and lrp[-1] == lcol.shape[0] and len(lrp) == part.rows(r) + 1
        cols.append(lcol)
    assert np.array_equal(np.concatenate(cols), col)


@pytest.mark.timeout(120)
def test_world2_step_matches_single_process_oracle():
    from wdgh_b200 import _lib
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([out.get(timeout=100) for _ in range(world)])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    row, col, labels, x = _graph()
    n, C = labels.shape[0], 4
    ones = np.ones(row.shape[0], np.float32)
    r2, c2, v2 = O.sys_normalized_adjacency(row, col, ones, n)
    ref_y = O.spmm(r2, c2, v2, n, x)
    y = np.concatenate([r[3] for r in res])
    np.testing.assert_allclose(y, ref_y, rtol=1e-5, atol=1e-6)
    s = O.structure_counts(row, col, labels, n, num_classes=C)
    for r in res:  # identical all-reduced counters on both ranks
        cnt = r[4]
        assert cnt[_lib.SC_MATCH_ALL] == s["match_all"] and cnt[_lib.SC_N_LAB] == s["n_lab"]
        assert np.array_equal(cnt[_lib.SC_HEADER + 2 * C:_lib.SC_HEADER + 2 * C + C * C].reshape(C, C), s["hist"])
        assert np.array_equal(cnt[_lib.SC_HEADER:_lib.SC_HEADER + C], s["class_count"])
        assert cnt[_lib.SC_NBINS] == np.nonzero(s["deg_nsl"])[0].max() + 1
        assert cnt[_lib.SC_N_NODES_NSL] == int((s["deg_nsl"] > 0).sum())
        keep = s["deg_nsl"] > 0
        ref_sum = (s["match_nsl"][keep].astype(np.float32) / s["deg_nsl"][keep].astype(np.float32)).astype(np.float64).sum()
        assert abs(r[5] - ref_sum) < 1e-9


def _worker_2d(rank, world, port, out):
    """One rank of the 2-D partition on CPU: Grid2D's slicing + schedule (product code) drive a gloo exchange."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from wdgh_b200.sharded import Grid2D, shard_csr

        row, col, labels, x = _graph()
        n, d = labels.shape[0], x.shape[1]
        grid = Grid2D(n, world, 2)
        blk, pc = grid.part.block, grid.pc
        i, j = grid.coords(rank)
        rowptr = O.csr_from_coo(row, n)
        deg = np.diff(rowptr).astype(np.float64) + 1.0
        dinv = (1.0 / np.sqrt(deg)).astype(np.float32)
        # features of my column group only (what the pulls deliver); everything else stays NaN on purpose
        x_cols = np.full((world * blk, d), np.nan, np.float32)
        for src in grid.col_group_ranks(j):
            a, b = grid.part.bounds(src)
            x_cols[a:b] = x[a:b]
        partial, sends = {}, []
        for k, s, owner in grid.schedule(rank):
            r0, r1 = grid.part.bounds(owner)
            lrp, lcol, _ = shard_csr(rowptr, col, None, r0, r1)
            frp, fcol = grid.filter_slice(torch.from_numpy(lrp), torch.from_numpy(lcol), j)
            frp, fcol = frp.numpy(), fcol.numpy()
            p = np.zeros((blk, d), np.float32)
            np.add.at(p, np.repeat(np.arange(r1 - r0), np.diff(frp)), dinv[fcol][:, None] * x_cols[fcol])
            assert np.isfinite(p).all()
            partial[s] = p
            if owner != rank:
                sends.append(dist.isend(torch.from_numpy(p), dst=owner, tag=k))
        recv = [torch.empty((blk, d)) for _ in range(pc)]
        reqs = [dist.irecv(recv[k], src=grid.slot_source(rank, k), tag=k) for k in range(1, pc)]
        [q.wait() for q in reqs + sends]
        r0, r1 = grid.part.bounds(rank)
        acc = partial[j].copy()
        for k in range(1, pc):
            acc += recv[k].numpy()
        y = (acc[:r1 - r0] + dinv[r0:r1, None] * x[r0:r1]) * dinv[r0:r1, None]
        out.put((rank, y))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world4_2d_partition_matches_single_process_oracle():
    world = 4
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_2d, args=(r, world, port, out)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([out.get(timeout=150) for _ in range(world)], key=lambda t: t[0])
    [p.join(30) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    row, col, labels, x = _graph()
    n = labels.shape[0]
    r2, c2, v2 = O.sys_normalized_adjacency(row, col, np.ones(row.shape[0], np.float32), n)
    ref_y = O.spmm(r2, c2, v2, n, x)
    np.testing.assert_allclose(np.concatenate([r[1] for r in res]), ref_y, rtol=1e-5, atol=1e-6)
