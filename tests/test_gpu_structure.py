"""Bit-exact checks of the integer structure kernels against the oracle (stable argsort / bincount)."""
import pytest
import torch

from oracle import phc_oracle as O

pytestmark = pytest.mark.gpu


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
