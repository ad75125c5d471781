"""CPU restatement of the PyG mini-batch collate (oracle/phc_oracle.py::collate): a hand-worked known answer and the
round trip against the synthetic batches (which are built graph by graph in the same way)."""
import types

import pytest
import torch

from oracle import phc_oracle as O


def test_collate_known_answer():
    g0 = types.SimpleNamespace(x=torch.tensor([[1, 2], [3, 4], [5, 6]]), edge_index=torch.tensor([[0, 1, 2], [1, 2, 0]]),
                               edge_attr=torch.tensor([[10], [11], [12]]), y=torch.tensor([[0.5]]))
    g1 = types.SimpleNamespace(x=torch.tensor([[7, 8], [9, 10]]), edge_index=torch.tensor([[0, 1], [1, 0]]),
                               edge_attr=torch.tensor([[20], [21]]), y=torch.tensor([[1.5]]))
    x, ei, ea, batch, y = O.collate([g1, g0, g1])
    assert x.tolist() == [[7, 8], [9, 10], [1, 2], [3, 4], [5, 6], [7, 8], [9, 10]]
    assert ei.tolist() == [[0, 1, 2, 3, 4, 5, 6], [1, 0, 3, 4, 2, 6, 5]]
    assert ea.view(-1).tolist() == [20, 21, 10, 11, 12, 20, 21]
    assert batch.tolist() == [0, 0, 1, 1, 1, 2, 2]
    assert y.view(-1).tolist() == [1.5, 0.5, 1.5]


@pytest.mark.parametrize("wl_name", ["hiv", "zinc", "mnist", "ppa"])
def test_split_then_collate_is_identity(wl_name):
    from phc_gnn_b200.synthetic import make_batch, workloads
    data = make_batch(workloads(4)[wl_name], seed=2, batch_graphs=5)
    graphs = O.split_batch(data)
    assert len(graphs) == 5 and all(int(g.edge_index.min()) >= 0 and int(g.edge_index.max()) < g.x.size(0) for g in graphs)
    x, ei, ea, batch, y = O.collate(graphs)
    assert torch.equal(x, data.x) and torch.equal(ei, data.edge_index) and torch.equal(ea, data.edge_attr)
    assert torch.equal(batch, data.batch) and torch.equal(y, data.y)


def test_device_store_refuses_cpu():
    from phc_gnn_b200.prep import DeviceGraphStore
    from phc_gnn_b200.synthetic import make_batch, workloads
    graphs = O.split_batch(make_batch(workloads(4)["zinc"], seed=0, batch_graphs=2))
    with pytest.raises(RuntimeError):
        DeviceGraphStore(graphs, "cpu")


@pytest.mark.parametrize("G,B,W", [(100, 8, 1), (100, 8, 4), (37, 5, 2), (64, 8, 8), (10, 4, 4)])
def test_epoch_sampler_shards_by_graph(G, B, W):
    """Every graph is visited by exactly one rank per epoch, all ranks run the same number of steps, epochs reshuffle,
    and the same (seed, epoch) gives the same order on every rank."""
    import numpy as np
    from phc_gnn_b200.prep import EpochSampler
    samplers = [EpochSampler(G, B, rank=r, world=W, seed=3) for r in range(W)]
    per_rank = [list(s) for s in samplers]
    steps = {len(b) for b in per_rank}
    assert len(steps) == 1 and steps.pop() == len(samplers[0])
    seen = np.concatenate([ids for b in per_rank for ids in b])
    assert len(seen) == len(set(seen.tolist()))                          # disjoint
    dropped = G - len(seen)
    assert 0 <= dropped < W                                              # only a tail smaller than the world is dropped
    for step in range(len(per_rank[0])):
        sizes = [len(per_rank[r][step]) for r in range(W)]
        assert max(sizes) - min(sizes) <= 1 and max(sizes) <= B
    again = [list(EpochSampler(G, B, rank=r, world=W, seed=3)) for r in range(W)]
    assert all(np.array_equal(a, b) for r in range(W) for a, b in zip(per_rank[r], again[r]))
    samplers[0].set_epoch(1)
    if G > B * W:
        assert not all(np.array_equal(a, b) for a, b in zip(per_rank[0], list(samplers[0])))
    full = [ids for ids in EpochSampler(G, B, rank=0, world=W, seed=3, drop_last=True)]
    assert all(len(ids) == B for ids in full)
    ordered = np.concatenate(list(EpochSampler(G, B, shuffle=False)))
    assert ordered.tolist() == list(range(G))


def test_plan_collate_offsets_match_the_oracle_batch():
    """Host half of the device collate: the prefix sums place every selected graph where the oracle's collate puts it."""
    import numpy as np
    from phc_gnn_b200.prep import plan_collate
    from phc_gnn_b200.synthetic import make_batch, split_graphs, workloads
    graphs = split_graphs(make_batch(workloads(4)["hiv"], seed=4, batch_graphs=9))
    nptr = np.concatenate([[0], np.cumsum([g.x.size(0) for g in graphs])]).astype(np.int64)
    eptr = np.concatenate([[0], np.cumsum([g.edge_index.size(1) for g in graphs])]).astype(np.int64)
    ids = [7, 0, 7, 3, 8]
    ids_np, onp, oep = plan_collate(torch.tensor(ids), nptr, eptr)
    x, ei, ea, batch, y = O.collate([graphs[i] for i in ids])
    assert ids_np.tolist() == ids and onp[-1] == x.size(0) and oep[-1] == ei.size(1)
    assert onp.tolist() == O.graph_ptr(batch, len(ids)).tolist()
    for b, g in enumerate(ids):
        assert torch.equal(x[onp[b]:onp[b + 1]], graphs[g].x)
        assert torch.equal(ei[:, oep[b]:oep[b + 1]] - int(onp[b]), graphs[g].edge_index)
    e_ids, e_onp, e_oep = plan_collate([], nptr, eptr)
    assert e_ids.shape == (0,) and e_onp.tolist() == [0] and e_oep.tolist() == [0]
    for bad in ([9], [-1], [0, 12]):
        with pytest.raises(IndexError):
            plan_collate(bad, nptr, eptr)


def test_balanced_partition_equalises_rank_work():
    import numpy as np
    from phc_gnn_b200.prep import EpochSampler, balanced_partition
    rng = np.random.default_rng(0)
    sizes = rng.integers(187, 301, 512)
    parts = balanced_partition(sizes, 8)
    assert sorted(np.concatenate(parts).tolist()) == list(range(512)) and all(len(p) == 64 for p in parts)
    loads = np.array([sizes[p].sum() for p in parts])
    naive = np.array([sizes[r * 64:(r + 1) * 64].sum() for r in range(8)])
    assert loads.max() - loads.min() <= 20 < naive.max() - naive.min()            # a handful of nodes (0.1 %) instead of hundreds
    # the sampler deals every global batch the same way: ranks are disjoint, cover the batch, equal counts
    costs = rng.integers(10, 40, 100)
    seen = []
    for r in range(4):
        batches = list(EpochSampler(100, 5, rank=r, world=4, seed=3, costs=costs))
        assert len(batches) == 5 and all(len(b) == 5 for b in batches)
        seen.append(batches)
    for step in range(5):
        ids = np.concatenate([seen[r][step] for r in range(4)])
        assert len(set(ids.tolist())) == 20
        loads = [costs[seen[r][step]].sum() for r in range(4)]
        assert max(loads) - min(loads) <= costs.max()


def test_make_batch_with_given_sizes():
    import numpy as np
    from phc_gnn_b200.synthetic import graph_sizes, make_batch, workloads
    wl = workloads(4)["hiv"]
    gs = graph_sizes(wl, 5, 16)
    b = make_batch(wl, seed=1, sizes=gs[:6])
    assert b.num_graphs == 6 and torch.bincount(b.batch).tolist() == gs[:6].tolist()
    assert torch.equal(make_batch(wl, seed=9, batch_graphs=4).x, make_batch(wl, seed=9, batch_graphs=4).x)
