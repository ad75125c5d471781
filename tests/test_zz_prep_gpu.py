"""Device-side mini-batch collate (phc_collate_batch through prep.DeviceGraphStore) against the CPU restatement of the PyG
collate — integer / byte work, so bit-exact."""
import numpy as np
import pytest
import torch

from oracle import phc_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _same(got, want, what):
    assert got.dtype == want.dtype and tuple(got.shape) == tuple(want.shape), f"{what}: {got.dtype}{tuple(got.shape)} vs {want.dtype}{tuple(want.shape)}"
    g, w = got.cpu().contiguous(), want.contiguous()
    if g.numel():                                  # byte-wise: the pcba targets hold NaNs
        assert torch.equal(g.reshape(-1).view(torch.uint8), w.reshape(-1).view(torch.uint8)), what


@pytest.mark.parametrize("wl_name,graphs", [("hiv", 24), ("zinc", 24), ("mnist", 6), ("ppa", 4), ("pcba", 40)])
def test_collate_matches_oracle(wl_name, graphs):
    from phc_gnn_b200.prep import DeviceGraphStore
    from phc_gnn_b200.synthetic import make_batch, workloads
    data = make_batch(workloads(4)[wl_name], seed=5, batch_graphs=graphs)
    dataset = O.split_batch(data)
    store = DeviceGraphStore(dataset, DEV)
    # the whole dataset in order reproduces the batch it was cut from
    full = store.collate(list(range(graphs)))
    store.check_status()
    for k in ("x", "edge_index", "edge_attr", "batch", "y"):
        _same(getattr(full, k), getattr(data, k), f"{wl_name} in-order {k}")
    assert full.num_graphs == graphs
    # shuffled selection with repeats, twice in a row (staging buffer reuse)
    rng = np.random.default_rng(1)
    for _ in range(2):
        ids = rng.integers(0, graphs, size=graphs + 3)
        got = store.collate(ids)
        store.check_status()
        x, ei, ea, batch, y = O.collate([dataset[i] for i in ids.tolist()])
        _same(got.x, x, f"{wl_name} x")
        _same(got.edge_index, ei, f"{wl_name} edge_index")
        _same(got.edge_attr, ea, f"{wl_name} edge_attr")
        _same(got.batch, batch, f"{wl_name} batch")
        _same(got.y, y, f"{wl_name} y")
    # a single graph and the empty selection
    one = store.collate([graphs - 1])
    store.check_status()
    _same(one.edge_index, dataset[-1].edge_index, "single graph edge_index")
    _same(one.x, dataset[-1].x, "single graph x")
    none = store.collate([])
    assert none.num_graphs == 0 and none.x.size(0) == 0 and none.edge_index.shape == (2, 0)
    with pytest.raises(IndexError):
        store.collate([graphs])


def test_collated_batch_trains(monkeypatch):
    """A batch assembled on the device goes through the model like a host-built one (same logits)."""
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc_gnn_b200.prep import DeviceGraphStore
    from phc_gnn_b200.synthetic import make_batch, tiny, workloads
    wl = tiny(workloads(4)["hiv"], 16, 2, 12, 5, 9, head=[12, 8])
    data = make_batch(wl, seed=3)
    torch.manual_seed(0)
    np.random.seed(0)
    model = PHMSkipConnectAdd(**wl.model).to(DEV).eval()
    store = DeviceGraphStore(O.split_batch(data), DEV)
    with torch.no_grad():
        a = model(store.collate(list(range(data.num_graphs))))
        b = model(data.to(DEV))
    assert torch.equal(a, b)


def test_device_loader_epoch():
    """DeviceLoader = sampler + collate (+ transform): one epoch over a small store on two 'ranks' reproduces the oracle's
    collate of the same ids, and the RemoveIsolatedNodes transform is applied per batch."""
    from phc_gnn_b200.prep import DeviceGraphStore, DeviceLoader, EpochSampler, RemoveIsolatedNodes
    from phc_gnn_b200.synthetic import make_batch, split_graphs, workloads
    data = make_batch(workloads(4)["hiv"], seed=9, batch_graphs=21)
    dataset = split_graphs(data)
    store = DeviceGraphStore(dataset, DEV)
    seen = []
    for rank in range(2):
        sampler = EpochSampler(len(dataset), 4, rank=rank, world=2, seed=5)
        loader = DeviceLoader(store, sampler)
        assert len(loader) == 3
        for ids, batch in zip(sampler, loader):
            store.check_status()
            x, ei, ea, bvec, y = O.collate([dataset[i] for i in ids.tolist()])
            _same(batch.x, x, "x")
            _same(batch.edge_index, ei, "edge_index")
            _same(batch.edge_attr, ea, "edge_attr")
            _same(batch.batch, bvec, "batch")
            _same(batch.y, y, "y")
            seen += ids.tolist()
    assert len(seen) == len(set(seen)) == 21              # global batches of 8, 8 and 5 graphs; the short one is split 3 + 2
    loader = DeviceLoader(store, EpochSampler(len(dataset), 8, seed=1), transform=RemoveIsolatedNodes())
    for ids, batch in zip(loader.sampler, loader):
        x, ei, ea, bvec, y = O.collate([dataset[i] for i in ids.tolist()])
        ei2, ea2, mask = O.remove_isolated_nodes(ei, ea, x.size(0))
        _same(batch.edge_index, ei2, "transformed edge_index")
        _same(batch.x, x[mask], "transformed x")
