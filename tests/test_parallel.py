"""World-size-2 gloo test (CPU) of the data-parallel host logic: parameter broadcast, flat gradient
bucket, averaged all-reduce, grads left as views of the flat buffer.  (The GPU kernels cannot run on
CPU, so a small dense model stands in; the bucket is model-agnostic.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from phc_gnn_b200.parallel import DataParallelPHC
        torch.manual_seed(100 + rank)                      # different init per rank: broadcast must fix it
        model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 3))
        dp = DataParallelPHC(model)
        w0 = [p.detach().clone() for p in model.parameters()]
        torch.manual_seed(7 + rank)
        x = torch.randn(4, 6)
        dp(x).square().sum().backward()
        local = [p.grad.detach().clone() for p in model.parameters()]
        dp.reduce_gradients()
        flat = dp.bucket.flat
        views_ok = all(p.grad.data_ptr() >= flat.data_ptr() and
                       p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4 for p in model.parameters())
        # second step: grads produced by autograd are fresh tensors again and must be re-packed
        for p in model.parameters():
            p.grad = None
        dp(x).square().sum().backward()
        dp.reduce_gradients()
        out[rank] = dict(w0=w0, local=local, reduced=[p.grad.detach().clone() for p in model.parameters()], views_ok=views_ok,
                         has_module=hasattr(dp, "module") and dp.module is model)
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    for x, y in zip(a["w0"], b["w0"]):
        assert torch.equal(x, y)                           # replicas start identical (broadcast from rank 0)
    for ga, gb, ra, rb in zip(a["local"], b["local"], a["reduced"], b["reduced"]):
        torch.testing.assert_close(ra, (ga + gb) / 2)
        assert torch.equal(ra, rb)                         # replicas see bit-identical reduced gradients
    assert a["views_ok"] and b["views_ok"] and a["has_module"]


def test_get_model_blocks_sees_through_wrapper():
    from phc_gnn_b200.nn import get_model_blocks
    from phc_gnn_b200.parallel import DataParallelPHC

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.convs = torch.nn.Linear(2, 2)
            self.pooling = torch.nn.Identity()
    m = M()
    dp = DataParallelPHC(m)
    assert len(get_model_blocks(dp, "convs", lr=0.1)[0]["params"]) == 2
    assert get_model_blocks(dp, "pooling") == [] and get_model_blocks(dp, "nope") == []
    assert get_model_blocks(m, "convs", lr=0.1)[0]["lr"] == 0.1


def _sharded_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from phc_gnn_b200.parallel import DataParallelPHC
        from phc_gnn_b200.prep import EpochSampler
        torch.manual_seed(0)
        feats, target = torch.randn(40, 6), torch.randn(40, 1)          # one row per "graph"
        torch.manual_seed(5 + rank)
        model = torch.nn.Linear(6, 1)
        dp = DataParallelPHC(model)                                     # broadcast from rank 0
        sampler = EpochSampler(40, 4, rank=rank, world=world, seed=9, drop_last=True)
        steps = []
        for ids in sampler:
            for p in model.parameters():
                p.grad = None
            idx = torch.from_numpy(ids)
            ((dp(feats[idx]) - target[idx]) ** 2).mean().backward()
            dp.reduce_gradients()
            steps.append((ids.tolist(), [p.grad.detach().clone() for p in model.parameters()]))
        out[rank] = dict(steps=steps, w=[p.detach().clone() for p in model.parameters()])
    finally:
        dist.destroy_process_group()


def test_sharded_epoch_equals_global_batches_gloo():
    """Data parallelism shards by graph: with EpochSampler's per-rank slices of every global batch, the all-reduced
    (averaged) gradient of each step equals the single-process gradient on the union of the ranks' graphs."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_sharded_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert len(a["steps"]) == len(b["steps"]) == 5
    torch.manual_seed(0)
    feats, target = torch.randn(40, 6), torch.randn(40, 1)
    model = torch.nn.Linear(6, 1)
    with torch.no_grad():
        for p, w in zip(model.parameters(), a["w"]):
            p.copy_(w)
    seen = []
    for (ia, ga), (ib, gb) in zip(a["steps"], b["steps"]):
        assert not set(ia) & set(ib) and len(ia) == len(ib) == 4
        seen += ia + ib
        idx = torch.tensor(ia + ib)
        for p in model.parameters():
            p.grad = None
        ((model(feats[idx]) - target[idx]) ** 2).mean().backward()
        for p, x, y in zip(model.parameters(), ga, gb):
            assert torch.equal(x, y)
            torch.testing.assert_close(x, p.grad, rtol=1e-5, atol=1e-6)
    assert len(set(seen)) == 40


class _StagedToy(torch.nn.Module):
    """Stand-in with the layout optim.staged_parameters understands (convs / norms / pooling / downstream) and the
    ``_report_stage`` protocol of the PHC models: a tensor hook on every layer output reports completed stages."""

    def __init__(self, L=3, F=5):
        super().__init__()
        self.enc = torch.nn.Linear(4, F)
        self.convs = torch.nn.ModuleList([torch.nn.Linear(F, F) for _ in range(L)])
        self.norms = torch.nn.ModuleList([torch.nn.LayerNorm(F) for _ in range(L)])
        self.pooling = torch.nn.Linear(F, F)
        self.downstream = torch.nn.Linear(F, 2)

    def forward(self, x):
        h0 = self.enc(x)
        h = h0
        L = len(self.convs)
        for i in range(L):
            h = torch.relu(self.norms[i](self.convs[i](h))) + h0          # "first" skip: h0 feeds every layer
            cb = self.__dict__.get("_stage_hook")
            if cb is not None and h.requires_grad:
                def _hook(_g, k=L - 1 - i, cb=cb):
                    cb(k)
                h.register_hook(_hook)
        return self.downstream(self.pooling(h).mean(0, keepdim=True))


def _staged_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from phc_gnn_b200.optim import staged_parameters
        from phc_gnn_b200.parallel import GradientBucket
        torch.manual_seed(3)
        model = _StagedToy()
        params, ends = staged_parameters(model)
        assert len(ends) == 4 and ends[-1] == len(params) == len(list(model.parameters()))
        # stage 0 = head, then layers 2, 1, and (layer 0 + encoder) last
        assert {id(p) for p in params[:ends[0]]} == {id(p) for m in (model.pooling, model.downstream) for p in m.parameters()}
        assert {id(p) for p in params[ends[2]:]} == {id(p) for m in (model.convs[0], model.norms[0], model.enc) for p in m.parameters()}
        bucket = GradientBucket(params, ends, min_chunk_bytes=0)
        assert bucket.enable_overlap(model, None)
        torch.manual_seed(10 + rank)
        x = torch.randn(6, 4)
        calls = []
        orig = bucket.stage_ready
        object.__setattr__(model, "_stage_hook", lambda k: (calls.append((k, bucket._done)), orig(k)))
        bucket.overlap = False                              # hooks fire but send nothing: the plain local gradients
        model(x).square().sum().backward()
        local = [p.grad.detach().clone() for p in params]
        calls.clear()
        bucket.overlap = True                               # the same backward with the overlapped reduction live
        for p in params:
            p.grad = None
        model(x).square().sum().backward()
        in_flight = bucket._done
        bucket.reduce()
        out[rank] = dict(local=local, reduced=[p.grad.detach().clone() for p in params], calls=calls, in_flight=in_flight,
                         launched=bucket.launched_last, ends=ends)
    finally:
        dist.destroy_process_group()


def test_staged_overlapped_allreduce_gloo():
    """GradientBucket with stages: slices go out as backward reports them (head first), the remainder with reduce(); the
    result is the plain average and identical on both ranks."""
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_staged_worker, args=(world, port, out), nprocs=world, join=True)
    a, b = out[0], out[1]
    assert [k for k, _ in a["calls"]][:3] == [0, 1, 2] or [k for k, _ in a["calls"]][-3:] == [0, 1, 2]
    assert a["in_flight"] == a["ends"][2]                  # three stages were sent during backward, the last one by reduce()
    assert a["launched"] == 4
    for ga, gb, ra, rb in zip(a["local"], b["local"], a["reduced"], b["reduced"]):
        torch.testing.assert_close(ra, (ga + gb) / 2)
        assert torch.equal(ra, rb)


def _uneven_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from phc_gnn_b200.parallel import DataParallelPHC
        torch.manual_seed(0)
        feats, target = torch.randn(5, 6), torch.randn(5, 1)
        model = torch.nn.Linear(6, 1)
        dp = DataParallelPHC(model)
        idx = torch.tensor([0, 1, 2]) if rank == 0 else torch.tensor([3, 4])          # a short last batch: 3 + 2 graphs
        ((dp(feats[idx]) - target[idx]) ** 2).mean().backward()
        dp.bucket.set_batch_share(len(idx), 5, world)
        dp.reduce_gradients()
        out[rank] = [p.grad.detach().clone() for p in model.parameters()]
    finally:
        dist.destroy_process_group()


def test_uneven_shards_reduce_to_the_global_batch_gradient_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_uneven_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    feats, target = torch.randn(5, 6), torch.randn(5, 1)
    model = torch.nn.Linear(6, 1)
    ((model(feats) - target) ** 2).mean().backward()
    for p, a, b in zip(model.parameters(), out[0], out[1]):
        assert torch.equal(a, b)
        torch.testing.assert_close(a, p.grad, rtol=1e-5, atol=1e-6)
