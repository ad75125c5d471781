"""Quaternion family and legacy-layout parameters on the CUDA path against the vectors recorded from
the UNMODIFIED reference (tests/golden/family, oracle/make_golden_family.py).  The quaternion models run the same kernels
as the PHM models with n = 4 and a frozen Hamilton rule — this file is also the coverage of ``learn_phm=False``."""
import pytest
import torch

from gpu_util import assert_close
from oracle import phc_oracle as O
from test_family_host import load_family, product_class, quaternion_cases

pytestmark = pytest.mark.gpu
RTOL = 1e-4      # north_star fp32 tolerance
DEV = "cuda:0"


def _tol(ref, rtol=RTOL):
    return max(rtol * float(ref.abs().max()), 2e-6)


def _run(fx, fuse):
    from phc.quaternion.regularization import quaternion_weight_regularization
    m = product_class(fx)(**fx["cfg"])
    m.load_quaternion_state_dict(fx["state"])
    m = m.to(DEV)
    m.fuse_edge_encoder = fuse != "none"
    m.fuse_layer = fuse == "direct"
    data = fx["batch"].to(DEV)
    m.train()
    logits = m(data)
    reg = quaternion_weight_regularization(m, device=DEV, p=2)
    loss = O.task_loss(logits, data.y, fx["loss_kind"]) + fx["reg_scale"] * reg
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: v.detach().cpu() for k, v in m.quaternion_named_gradients().items()}
    running = {k: v.detach().cpu() for k, v in m.quaternion_state_dict().items() if "running" in k or "tracked" in k}
    m.eval()
    with torch.no_grad():
        ev = m(data).cpu()
    return m, logits.detach().cpu(), reg.detach().cpu(), loss.detach().cpu(), grads, running, ev


@pytest.mark.parametrize("fuse", ["direct", "none"], ids=["one-call-per-layer", "separate-ops"])
@pytest.mark.parametrize("name", quaternion_cases())
def test_quaternion_model_matches_reference_golden(name, fuse, monkeypatch):
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    fx = load_family(name)
    m, logits, reg, loss, grads, running, ev = _run(fx, fuse)
    assert_close(logits, fx["logits_train"], RTOL, _tol(fx["logits_train"]), f"{name}: train logits")
    assert_close(reg, fx["reg"], RTOL, 1e-6, f"{name}: regulariser")
    assert_close(loss, fx["loss"], RTOL, _tol(fx["loss"]), f"{name}: loss")
    assert_close(ev, fx["logits_eval"], RTOL, _tol(fx["logits_eval"]), f"{name}: eval logits")
    for k, g in fx["grads"].items():
        assert k in grads, f"{name}: no gradient for {k}"
        assert_close(grads[k], g, 5 * RTOL, 10 * _tol(g), f"{name}: grad {k}")
    for k, v in fx["running_after"].items():
        assert_close(running[k].float(), v.float(), RTOL, max(1e-5, _tol(v.float())), f"{name}: {k}")
    for mod in m.modules():                                      # the rule stayed frozen and untouched
        r = getattr(mod, "phm_rule", None)
        if isinstance(r, torch.nn.Parameter):
            assert r.grad is None and not r.requires_grad


def test_quaternion_regulariser_p1_matches_oracle():
    from phc.quaternion.regularization import quaternion_weight_regularization
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd
    fx = load_family("quaternion_hiv_softmax_mlp")
    m = QuaternionSkipConnectAdd(**fx["cfg"])
    m.load_quaternion_state_dict(fx["state"])
    m = m.to(DEV)
    got = quaternion_weight_regularization(m, device=DEV, p=1)
    want = O.quaternion_weight_regularization({k: v for k, v in fx["state"].items()}, fx["cfg"], 1)
    assert_close(got.detach().cpu(), want, RTOL, 1e-6, "p=1 regulariser")


def test_legacy_phm_linear_on_device(monkeypatch):
    """PHMLinear_Old vectors from the reference through the CUDA PHMLinear on converted parameters."""
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    from phc.hypercomplex.layers import PHMLinear
    from phc_gnn_b200 import legacy
    for key, fx in load_family("legacy_phmlinear").items():
        n = fx["n"]
        lin = PHMLinear(n * fx["in_per"], n * fx["out_per"], n, c_init="standard")
        lin.load_state_dict(legacy.convert_legacy_phm_state_dict(fx["state"]), strict=True)
        lin = lin.to(DEV)
        x = fx["x"].to(DEV).requires_grad_(True)
        y = lin(x)
        assert_close(y.detach().cpu(), fx["y"], RTOL, _tol(fx["y"]), f"{key}: y")
        y.backward(fx["gy"].to(DEV))
        assert_close(x.grad.cpu(), fx["gx"], RTOL, _tol(fx["gx"]), f"{key}: dx")
        g = legacy.to_legacy_phm_state_dict({k: p.grad.detach().cpu() for k, p in lin.named_parameters()})
        for k, want in fx["grads"].items():
            assert_close(g[k], want, 5 * RTOL, 10 * _tol(want), f"{key}: grad {k}")


def test_quaternion_train_steps(monkeypatch):
    """The timed unit (train.TrainStep: forward, loss + quaternion regulariser, backward, clip, flat Adam) on a quaternion
    model against the same steps taken by the CPU oracle on the reference's quaternion parameters (dropout off, one
    repeated batch): per-step losses agree, weights move, the Hamilton rule stays frozen and bit-identical."""
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    import numpy as np
    from test_family_host import leaves
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd
    from phc_gnn_b200 import legacy
    from phc_gnn_b200.quaternion import quaternion_weight_regularization
    from phc_gnn_b200.synthetic import make_batch, tiny, workloads
    from phc_gnn_b200.train import TrainStep
    wl = tiny(workloads(4)["zinc"], 16, 2, 24, 5, 12, head=[16, 8])
    wl.lr, wl.weight_decay = 5e-3, 0.1
    kw = {k: v for k, v in wl.model.items() if k not in ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type")}
    kw["dropout_mpnn"] = [0.0] * len(kw["mp_layers"])
    kw["dropout_dn"] = [0.0] * len(kw["downstream_layers"])
    torch.manual_seed(0)
    np.random.seed(0)
    model = QuaternionSkipConnectAdd(init="quaternion", **kw)
    pq = leaves(model.quaternion_state_dict(), torch.float64)
    model = model.to(DEV)
    model.train()
    step = TrainStep(model, wl)
    assert step.regulariser is quaternion_weight_regularization
    host = make_batch(wl, seed=1)
    data = host.to(DEV)
    w0 = model.downstream.affine[0].W.detach().clone()
    got = [float(step(data)) for _ in range(4)]

    train = [v for v in pq.values() if v.requires_grad]
    opt = torch.optim.Adam(train, lr=wl.lr)
    want = []
    for _ in range(4):
        opt.zero_grad()
        logits = O.quaternion_model_forward(pq, kw, host, training=True)
        loss = O.task_loss(logits, host.y, wl.loss) + wl.lr * wl.weight_decay * O.quaternion_weight_regularization(pq, kw, 2)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(train, max_norm=wl.grad_clip, norm_type=2)
        opt.step()
        want.append(float(loss.detach()))
    assert abs(got[0] - want[0]) <= 1e-4 * abs(want[0]), (got, want)
    for g, w in zip(got[1:], want[1:]):
        # later steps: Adam normalises gradients, so entries whose gradient is zero in exact arithmetic (a bias in front
        # of a batch norm) take +-lr noise steps on both sides; they do not reach the loss
        assert abs(g - w) <= 2e-3 * abs(w), (got, want)
    assert not torch.equal(model.downstream.affine[0].W.detach(), w0)
    for mod in model.modules():
        r = getattr(mod, "phm_rule", None)
        if isinstance(r, torch.nn.Parameter):
            assert torch.equal(r.detach().cpu(), legacy.hamilton_rule())


@pytest.mark.parametrize("fuse", ["direct", "none"], ids=["one-call-per-layer", "separate-ops"])
@pytest.mark.parametrize("name", __import__("test_family_host").phm_option_cases())
def test_phm_constructor_options_match_reference_golden(name, fuse, monkeypatch):
    """Constructor options outside the 14 main fixtures — naive encoders (embedding / linear), add_self_loops=False,
    bias=False, learn_phm=False (whose reference semantics keep the GINE MLP rules trainable), fixed softmax beta — and
    PHMSkipConnectConcat at phm_dim = 1 (the only n the reference's concat model runs for)."""
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    from phc.hypercomplex.regularization import phm_weight_regularization
    from test_family_host import phm_product_class
    fx = load_family(name)
    m = phm_product_class(fx)(**fx["cfg"])
    m.load_state_dict(fx["state"], strict=True)
    m = m.to(DEV)
    m.fuse_edge_encoder = fuse != "none"
    m.fuse_layer = fuse == "direct"
    data = fx["batch"].to(DEV)
    m.train()
    logits = m(data)
    reg = phm_weight_regularization(m, p=2)
    loss = O.task_loss(logits, data.y, fx["loss_kind"]) + fx["reg_scale"] * reg
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().cpu() for k, p in m.named_parameters() if p.grad is not None}
    running = {k: v.detach().cpu() for k, v in m.state_dict().items() if "running" in k or "tracked" in k}
    m.eval()
    with torch.no_grad():
        ev = m(data).cpu()
    assert_close(logits.detach().cpu(), fx["logits_train"], RTOL, _tol(fx["logits_train"]), f"{name}: train logits")
    assert_close(reg.detach().cpu(), fx["reg"], RTOL, 1e-6, f"{name}: regulariser")
    assert_close(loss.detach().cpu(), fx["loss"], RTOL, _tol(fx["loss"]), f"{name}: loss")
    assert_close(ev, fx["logits_eval"], RTOL, _tol(fx["logits_eval"]), f"{name}: eval logits")
    for k, g in fx["grads"].items():
        assert k in grads, f"{name}: no gradient for {k}"
        assert_close(grads[k], g, 5 * RTOL, 10 * _tol(g), f"{name}: grad {k}")
    trainable = {k for k, p in m.named_parameters() if p.requires_grad}
    assert trainable == set(fx["grads"]), f"{name}: trainable parameter sets differ: {sorted(trainable ^ set(fx['grads']))}"
    for k, v in fx["running_after"].items():
        assert_close(running[k].float(), v.float(), RTOL, max(1e-5, _tol(v.float())), f"{name}: {k}")
