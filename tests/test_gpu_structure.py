"""Bit-exact checks of the integer structure kernels against the oracle (stable argsort / bincount)."""
import pytest
import torch

from oracle import phc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(ei, n):
    from phc_gnn_b200.graph import EdgeStructure
    dev = torch.device("cuda:0")
    s = EdgeStructure(ei.to(dev), n)
    s.validate()
    rowptr, col, perm = O.csr_by_target(ei, n)
    assert torch.equal(s.rowptr.cpu().long(), rowptr)
    assert torch.equal(s.perm.cpu().long(), perm)
    assert torch.equal(s.col.cpu().long(), col)
    rowptr_t, col_t, perm_t = O.csr_by_source(ei, n)
    assert torch.equal(s.rowptr_t.cpu().long(), rowptr_t)
    assert torch.equal(s.perm_t.cpu().long(), perm_t)
    assert torch.equal(s.col_t.cpu().long(), col_t)


@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (5, 0, 0), (7, 20, 1), (100, 1000, 2), (3000, 7000, 3), (16000, 300000, 4),
                                      (1025, 5000, 5), (50, 5000, 6)])
def test_csr_random(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, e), generator=g, dtype=torch.int64)
    _check(ei, n)


def test_csr_hub_and_isolated():
    g = torch.Generator().manual_seed(0)
    n = 300
    hub = torch.stack([torch.randint(0, n, (700,), generator=g), torch.full((700,), 17)])       # in-degree 700
    fan = torch.stack([torch.full((90,), 200), torch.randint(0, 100, (90,), generator=g)])       # out-degree 90
    ei = torch.cat([hub, fan], 1)[:, torch.randperm(790, generator=g)]
    _check(ei.contiguous(), n)


def test_csr_workloads():
    from phc_gnn_b200.synthetic import workloads, make_batch
    for name, b in (("hiv", 16), ("mnist", 4), ("ppa", 2)):
        d = make_batch(workloads(4)[name], seed=1, batch_graphs=b)
        _check(d.edge_index, d.x.size(0))


def test_out_of_range_is_flagged():
    from phc_gnn_b200.graph import EdgeStructure
    s = EdgeStructure(torch.tensor([[0, 1, 9], [1, 0, 0]]).cuda(), 3)
    with pytest.raises(IndexError):
        s.validate()


@pytest.mark.parametrize("sizes", [[3, 4, 5], [1], [0, 2, 0, 0, 3, 0], [5, 0], [0, 0, 7]])
def test_graph_ptr(sizes):
    from phc_gnn_b200.graph import SegmentStructure
    batch = torch.cat([torch.full((s,), i, dtype=torch.int64) for i, s in enumerate(sizes)]) if sum(sizes) else torch.zeros(0, dtype=torch.int64)
    s = SegmentStructure(batch.cuda(), len(sizes))
    s.validate()
    assert torch.equal(s.graph_ptr.cpu().long(), O.graph_ptr(batch, len(sizes)))


def test_graph_ptr_unsorted_is_flagged():
    from phc_gnn_b200.graph import SegmentStructure
    s = SegmentStructure(torch.tensor([0, 2, 1]).cuda(), 3)
    with pytest.raises(ValueError):
        s.validate()


def _random_graph_with_isolated(seed, N, E, loops, dup_loops=False):
    g = torch.Generator().manual_seed(seed)
    active = torch.randperm(N, generator=g)[: max(2, int(0.7 * N))]
    src = active[torch.randint(0, active.numel(), (E,), generator=g)]
    dst = active[torch.randint(0, active.numel(), (E,), generator=g)]
    lp = torch.randint(0, N, (loops,), generator=g)                      # self loops, some on otherwise isolated nodes
    if dup_loops and loops > 1:
        lp[1] = lp[0]
    ei = torch.cat([torch.stack([src, dst]), torch.stack([lp, lp])], dim=1)
    perm = torch.randperm(ei.size(1), generator=g)
    ei = ei[:, perm]
    attr = torch.randint(0, 5, (ei.size(1), 3), generator=g)
    return ei, attr


@pytest.mark.parametrize("seed,N,E,loops,dup", [(0, 50, 120, 6, False), (1, 3000, 7000, 40, True), (2, 15616, 290000, 0, False),
                                                (3, 10, 0, 3, False), (4, 1, 0, 0, False), (5, 257, 1024, 1024, True)])
def test_remove_isolated_nodes_bit_exact(seed, N, E, loops, dup):
    """csrc/prep.cu against the CPU restatement of torch_geometric 1.6.1 remove_isolated_nodes: mask, relabelled
    edge list, edge order and gathered edge features are identical."""
    from oracle import phc_oracle as O
    from phc_gnn_b200.prep import remove_isolated_nodes
    ei, attr = _random_graph_with_isolated(seed, N, E, loops, dup)
    want_ei, want_attr, want_mask = O.remove_isolated_nodes(ei, attr, N)
    got_ei, got_attr, got_mask = remove_isolated_nodes(ei.to(DEV), attr.to(DEV), N)
    assert torch.equal(got_mask.cpu(), want_mask)
    assert torch.equal(got_ei.cpu(), want_ei)
    assert torch.equal(got_attr.cpu(), want_attr)
    # properties: every kept node touches a non-loop edge; applying the transform again changes nothing
    if want_ei.numel():
        nl = want_ei[:, want_ei[0] != want_ei[1]]
        deg = torch.bincount(nl.reshape(-1), minlength=int(want_mask.sum()))
        assert bool((deg > 0).all())
    again_ei, again_attr, again_mask = remove_isolated_nodes(got_ei, got_attr, int(want_mask.sum()))
    assert bool(again_mask.all()) and torch.equal(again_ei, got_ei) and torch.equal(again_attr, got_attr)


def test_remove_isolated_nodes_transform_on_batch():
    """The transform object filters node-level tensors (x, batch) and leaves graph-level ones (y) alone; the model runs on
    the result."""
    from phc_gnn_b200.prep import RemoveIsolatedNodes
    from phc_gnn_b200.synthetic import make_batch, tiny, workloads
    wl = tiny(workloads(4)["hiv"], 32, 2, 6, 8, 14)
    b = make_batch(wl, seed=3)
    # append three isolated nodes to the last graph
    extra = 3
    b.x = torch.cat([b.x, b.x[:extra]], 0)
    b.batch = torch.cat([b.batch, b.batch[-1:].repeat(extra)], 0)
    d = RemoveIsolatedNodes()(b.to(DEV))
    assert d.x.size(0) == b.x.size(0) - extra and d.batch.size(0) == d.x.size(0)
    assert torch.equal(d.edge_index.cpu(), b.edge_index) and torch.equal(d.y.cpu(), b.y)
