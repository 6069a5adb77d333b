"""Host-side logic of the mirror that needs no GPU: RNG-compatible splits, accuracy, masks."""
import numpy as np
import pytest
import torch

from oracle import ref_port as O


def test_splits_consume_rng_like_the_reference():
    from wdgh_b200 import util_funcs as uf
    labels = torch.from_numpy(np.random.default_rng(0).integers(0, 5, 1000))
    for pct in (0.6, 0.3):
        torch.manual_seed(123)
        a = uf.random_disassortative_splits(labels, torch.tensor(5), pct)
        after_a = torch.rand(1)
        torch.manual_seed(123)
        b = O.random_disassortative_splits(labels, 5, pct)
        after_b = torch.rand(1)
        for m1, m2 in zip(a, b):
            assert m1.dtype == torch.bool and torch.equal(m1, m2)
        assert torch.equal(after_a, after_b)          # same number of RNG draws
        tr, va, te = a
        assert not (tr & va).any() and not (tr & te).any() and not (va & te).any()
        assert int(tr.sum() + va.sum() + te.sum()) == 1000
        assert int(va.sum()) == 200


def test_accuracy_and_mask():
    from wdgh_b200 import util_funcs as uf
    out = torch.tensor([[0.1, 0.9], [0.8, 0.2], [0.3, 0.7], [0.6, 0.4]])
    labels = torch.tensor([1, 0, 0, 0])
    assert float(uf.accuracy(labels, out)) == float(O.accuracy(labels, out)) == 0.75
    m = uf.index_to_mask(torch.tensor([0, 3]), 5)
    assert m.tolist() == [True, False, False, True, False]


def test_rand_train_test_idx_partitions_labelled_nodes():
    from wdgh_b200 import util_funcs as uf
    label = torch.tensor([0, 1, -1, 1, 0, -1, 1, 0, 0, 1])
    np.random.seed(4)
    tr, va, te = uf.rand_train_test_idx(label)
    allidx = torch.cat([tr, va, te]).sort().values
    assert allidx.tolist() == [0, 1, 3, 4, 6, 7, 8, 9]
    assert len(tr) == 4 and len(va) == 1
    np.random.seed(4)
    perm = np.random.permutation(8)
    assert tr.tolist() == torch.where(label != -1)[0][perm[:4]].tolist()


def test_row_partition_is_used_by_bench_defaults():
    from wdgh_b200.sharded import RowPartition
    p = RowPartition(50_000_000, 8)
    assert p.block == 6_250_000 and p.bounds(7) == (43_750_000, 50_000_000)


def test_grid2d_blocks_tile_the_matrix():
    """Every (row, column) pair falls into exactly one rank's block; groups have the documented members."""
    from wdgh_b200.sharded import Grid2D
    g = Grid2D(1000, 8, 2)
    assert (g.pr, g.pc) == (2, 4) and g.coords(6) == (1, 2)
    assert g.row_group_ranks(1) == [4, 5, 6, 7] and g.col_group_ranks(2) == [2, 6]
    col = torch.arange(1000)
    owner = col // g.part.block
    hits = torch.zeros(1000, dtype=torch.int64)
    for j in range(g.pc):
        m = g.col_in_group(col, j)
        assert set(owner[m].tolist()) <= set(g.col_group_ranks(j))
        hits += m
    assert bool((hits == 1).all())
    # ranks of a row group cover the rows of that group once each
    rows = sorted(r for s in g.row_group_ranks(0) for r in range(*g.part.bounds(s)))
    assert rows == list(range(0, 4 * g.part.block))


@pytest.mark.parametrize("world,pr", [(4, 2), (8, 2), (8, 4), (6, 2)])
def test_grid2d_schedule_and_receive_slots_agree(world, pr):
    """What rank r pushes in step k lands in slot k of the owner, and the owner expects exactly r there; every
    foreign slice of a row group is produced once per column group, the own slice comes last."""
    from wdgh_b200.sharded import Grid2D
    g = Grid2D(997, world, pr)
    for r in range(world):
        i, j = g.coords(r)
        sched = g.schedule(r)
        assert [k for k, _, _ in sched] == list(range(1, g.pc + 1))
        assert sorted(s for _, s, _ in sched) == list(range(g.pc))           # every slice of the row group once
        assert sched[-1][1:] == (j, r)                                         # own slice last, nothing to send
        for k, s, owner in sched[:-1]:
            assert g.coords(owner) == (i, s) and owner != r
            assert g.slot_source(owner, k) == r
        # the pc - 1 receive slots of r are fed by pc - 1 distinct ranks of its row group
        sources = [g.slot_source(r, k) for k in range(1, g.pc)]
        assert sorted(sources + [r]) == g.row_group_ranks(i)


def test_bench_clock_sampler_keeps_only_samples_inside_the_window():
    import datetime
    import bench

    def line(t, sm, cap):
        ts = datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return f"{ts}, {sm}, 1965, 512.3, Not Active, Not Active, Not Active, {cap}, 97"

    t = 1_792_000_000.0
    text = "\n".join([line(t - 5.0, 210, "Not Active"), line(t + 0.1, 1600, "Active"), line(t + 0.2, 1700, "Not Active"),
                      line(t + 0.3, 1650, "Not Active"), line(t + 9.0, 300, "Not Active"), "garbage"])
    out = bench.ClockSampler.parse(text, t, t + 0.35)
    assert out == {"sm_mhz": 1650.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    assert bench.ClockSampler.parse("", t, t + 1)["samples"] == 0


def test_bench_generator_is_partition_and_flag_independent(monkeypatch):
    """The synthetic graph must be the same whoever generates a row range: any sharding, with or without features
    (the 2-D partition regenerates its peers' rows with want_x=False)."""
    import bench
    monkeypatch.setattr(bench, "TILE_ROWS", 1000)
    dev = torch.device("cpu")
    n = 4700
    rp, col, x, lab = bench.gen_rows(0, n, n, 6.0, 5, 0.3, 8, dev)
    assert rp[0] == 0 and rp[-1] == col.shape[0] and x.shape == (n, 8) and lab.shape == (n,)
    for r0, r1 in ((0, 1175), (1175, 2350), (2350, 4700), (3300, 3301)):
        for want_x in (True, False):
            rp_s, col_s, x_s, lab_s = bench.gen_rows(r0, r1, n, 6.0, 5, 0.3, 8, dev, want_x=want_x)
            assert torch.equal(rp_s, rp[r0:r1 + 1] - rp[r0])
            assert torch.equal(col_s, col[rp[r0]:rp[r1]])
            assert torch.equal(lab_s, lab[r0:r1])
            if want_x:
                assert torch.equal(x_s, x[r0:r1])
            else:
                assert x_s is None


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract's
    keys; under torchrun only rank 0 works and prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--nodes", "20000",
           "--cpu-sample-nodes", "20000", "--steps", "1", "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=root, env={**os.environ, "RANK": "0"})
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert out.returncode == 0 and len(lines) == 1, out.stderr[-1500:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "GEdges/s" and line["higher_is_better"] is True
    assert line["metric"] == "aggregation_plus_homophily_throughput" and line["value"] > 0
    # the unmodified reference when a reference tree is reachable (/root/reference here, oracle/_ref on the GPU box)
    from oracle import ref_shim
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_shim.default_root() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and "SAMPLE" in line["cpu_baseline"]["sample"]
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("SAMPLE") and line["config"]["stored_entries"] > 0
    assert line["config"]["nodes"] == 20000
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=120, cwd=root, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""


def test_similarity_near_ties_counts_exact_ties():
    """oracle.similarity_near_ties (the sweep's tolerance for soft LAS): a node whose aggregated label distribution is
    uniform has ratio = 1 in exact arithmetic whatever the class sizes are -- those nodes, and only those, are counted."""
    from oracle import ref_port as O
    c, per = 4, 6
    n = c * per
    labels = np.repeat(np.arange(c), per)
    # ring inside every class (strongly homophilous: ratio > 1 for everybody) ...
    src = np.arange(n)
    dst = (src // per) * per + (src % per + 1) % per
    row, col = np.concatenate([src, dst]), np.concatenate([dst, src])
    oh = np.eye(c, dtype=np.float32)[labels]

    def rownorm(row, col):
        a = np.zeros((n, n))
        a[row, col] = 1.0
        a = a / a.sum(1, keepdims=True)
        r, c_ = np.nonzero(a)
        return r, c_, a[r, c_]
    r, c_, v = rownorm(row, col)
    assert O.similarity_near_ties(oh, r, c_, v, n, oh) == 0
    # ... then node 0 is rewired to exactly one neighbour of every class: its row of A X is uniform -> an exact tie
    keep = (row != 0) & (col != 0)
    extra_dst = np.array([1, per, 2 * per, 3 * per])
    row2 = np.concatenate([row[keep], np.zeros(4, np.int64)])
    col2 = np.concatenate([col[keep], extra_dst])
    r, c_, v = rownorm(row2, col2)
    assert O.similarity_near_ties(oh, r, c_, v, n, oh) == 1
