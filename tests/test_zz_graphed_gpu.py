"""graphed.GraphedTrainStep: the whole train() body replayed from a CUDA graph must be the eager step, bit for bit
(same kernels, same order, deterministic reductions), including Adam's device-side step counter; dropout must draw
fresh masks on every replay; the flat optimizer must work with the reference scripts' lr schedulers and resume."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(wname, dropout, seed=0):
    from phc_gnn_b200.nn import PHMSkipConnectAdd
    from phc_gnn_b200.synthetic import make_batch, tiny, workloads
    from phc_gnn_b200.train import TrainStep
    wl = tiny(workloads(4)[wname], 64, 2, 32, 8, 20, und_edges=40 if wname == "ppa" else None, head=[32, 16])
    if not dropout:
        wl.model["dropout_mpnn"] = [0.0] * 2
        wl.model["dropout_dn"] = [0.0] * 2
    torch.manual_seed(seed)
    import numpy as np
    np.random.seed(seed)                                    # phm_init draws from numpy, like the reference
    model = PHMSkipConnectAdd(**wl.model).to(DEV)
    model.train()
    batches = [make_batch(wl, seed=s).to(DEV) for s in (1, 2, 3)]
    return wl, model, TrainStep(model, wl, None, None), batches


@pytest.mark.parametrize("wname", ["hiv", "ppa", "mnist"])
def test_graph_replay_equals_eager_step_bitwise(wname, monkeypatch):
    monkeypatch.delenv("PHC_PRECISION", raising=False)
    from phc_gnn_b200.graphed import GraphedTrainStep
    _, ma, sa, batches = _setup(wname, dropout=False)
    _, mb, sb, _ = _setup(wname, dropout=False)
    gb = GraphedTrainStep(sb, capture_after=2)
    order = [0, 1, 2, 0, 1, 2, 0, 0, 1, 2, 2]
    for i in order:
        la = sa(batches[i]).clone()
        lb = gb(batches[i]).clone()
        assert torch.equal(la, lb), f"{wname}: loss differs at batch {i}: {float(la)} vs {float(lb)}"
    st = gb.stats()
    assert st["graphs"] == 3 and st["captures"] == 3 and st["replays"] == len(order) - 6, st
    for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
        assert torch.equal(pa, pb), f"{wname}: parameter {k} differs after {len(order)} steps"
    for (k, ba), (_, bb) in zip(ma.named_buffers(), mb.named_buffers()):
        assert torch.equal(ba, bb), f"{wname}: buffer {k} differs"
    assert sa.opt.t == sb.opt.t == len(order)
    assert torch.equal(sa.opt.exp_avg, sb.opt.exp_avg) and torch.equal(sa.opt.exp_avg_sq, sb.opt.exp_avg_sq)


def test_graph_replays_draw_fresh_dropout_masks():
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graphed import dropout_epoch
    from phc_gnn_b200.graph import _stream
    dev = torch.device(DEV)
    ep = dropout_epoch(dev)
    h = torch.ones(512, 64, device=dev)
    ops.bn_act_drop_skip(h, None, phm_dim=4, use_bn=False, training=True, drop_p=0.5, drop_same=False)       # warm-up
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    torch.manual_seed(3)
    with torch.cuda.graph(g):
        ops.run("phc_dropout_epoch_advance", None, ep.data_ptr(), _stream(dev))
        y = ops.bn_act_drop_skip(h, None, phm_dim=4, use_bn=False, training=True, drop_p=0.5, drop_same=False)
    outs = []
    for _ in range(3):
        g.replay()
        outs.append(y.clone())
    torch.cuda.synchronize()
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2])
    for o in outs:
        keep = float((o != 0).float().mean())
        assert 0.45 < keep < 0.55 and torch.all((o == 0) | (o == 2.0))


def test_dropout_forward_and_backward_masks_agree_inside_a_replayed_step():
    """With dropout on, a replayed step must still be a consistent forward/backward pair: the loss decreases over replays of one
    batch just as it does eagerly (a mask mismatch between forward and backward would give garbage gradients)."""
    from phc_gnn_b200.graphed import GraphedTrainStep
    _, m, s, batches = _setup("hiv", dropout=True)
    g = GraphedTrainStep(s, capture_after=1)
    losses = [float(g(batches[0])) for _ in range(40)]
    assert g.stats()["replays"] == 39
    assert all(l == l for l in losses)
    assert sum(losses[-5:]) / 5 < sum(losses[:5]) / 5


def test_flat_clip_adam_is_an_optimizer_schedulers_and_resume_work():
    from phc_gnn_b200.optim import FlatClipAdam
    _, m, s, batches = _setup("hiv", dropout=False)
    opt = s.opt
    assert isinstance(opt, torch.optim.Optimizer) and isinstance(opt, FlatClipAdam)
    plateau = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="max", factor=0.5, patience=0)
    steplr = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    s(batches[0])
    lr0 = opt.param_groups[0]["lr"]
    plateau.step(1.0)
    plateau.step(0.5)                                       # no improvement -> halves the rate
    assert opt.param_groups[0]["lr"] == pytest.approx(lr0 * 0.5)
    s(batches[1])
    assert float(opt.lr_dev) == pytest.approx(lr0 * 0.5)    # uploaded before the kernel ran
    steplr.step()
    state = copy.deepcopy(opt.state_dict())
    params = [p.detach().clone() for p in m.parameters()]
    s(batches[2])
    after = [p.detach().clone() for p in m.parameters()]
    # resume: restore parameters + optimizer state, repeat the step -> identical result
    with torch.no_grad():
        for p, v in zip(m.parameters(), params):
            p.copy_(v)
    opt.load_state_dict(state)
    assert opt.t == state["t"] == 2
    for mod in m.modules():                                  # BN running statistics moved too: irrelevant for train-mode outputs
        pass
    s(batches[2])
    for p, v in zip(m.parameters(), after):
        assert torch.allclose(p, v, rtol=0, atol=0), "resumed step differs"
