"""Operator-level parity of the CUDA kernels (through the C ABI) against the CPU oracle and the
reference-recorded known answers.  fp32 tolerance: rtol 1e-4 (north_star), atol scaled per case."""
import itertools

import pytest
import torch

from gpu_util import assert_close, grad_tol
from oracle import phc_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4
DEV = "cuda:0"


def _graph(n, e, seed, dup_free=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.randint(0, n, (2, 3 * e), generator=g, dtype=torch.int64)
    if dup_free:
        key = torch.unique(ei[0] * n + ei[1])
        key = key[torch.randperm(key.numel(), generator=g)][:e]
        ei = torch.stack([key // n, key % n])
    return ei.contiguous()


@pytest.mark.parametrize("reduce,act,F,self_loop", [
    ("add", "identity", 16, True), ("add", "relu", 20, False), ("mean", "identity", 12, True), ("mean", "swish", 8, False),
    ("max", "identity", 16, True), ("max", "relu", 16, False), ("min", "elu", 12, True), ("softmax", "identity", 16, True),
    ("softmax", "relu", 8, False), ("softmax", "swish", 20, True), ("add", "identity", 7, True), ("max", "lrelu", 9, False),
    ("softmax", "selu", 5, True), ("add", "identity", 500, True), ("softmax", "identity", 200, True)])
def test_aggregate_fwd_bwd(reduce, act, F, self_loop):
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import EdgeStructure
    N, E = 57, 300
    ei = _graph(N, E, 3)
    E = ei.size(1)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, F, generator=g)
    ea = torch.randn(E, F, generator=g)
    beta = torch.tensor(0.8)
    gout = torch.randn(N, F, generator=g)
    # oracle in fp64
    xo, eo, bo = (t.double().requires_grad_(True) for t in (x, ea, beta))
    ref = O.propagate(xo, ei, eo, reduce, act, bo)
    if self_loop:
        ref = ref + xo
    ref.backward(gout.double())
    xs, es, bs = (t.to(DEV).requires_grad_(True) for t in (x, ea, beta))
    s = EdgeStructure(ei.to(DEV), N)
    out = ops.aggregate(xs, es, s, reduce, act, bs, self_loop)
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 1e-5, "out")
    assert_close(xs.grad.cpu(), xo.grad.float(), RTOL, grad_tol(xo.grad, RTOL), "dx")
    assert_close(es.grad.cpu(), eo.grad.float(), RTOL, grad_tol(eo.grad, RTOL), "dea")
    if reduce == "softmax":
        assert_close(bs.grad.cpu(), bo.grad.float(), 1e-3, grad_tol(bo.grad, 1e-3), "dbeta")


@pytest.mark.parametrize("aggregators,scalers,act,F,n", [
    (["mean", "min", "max", "std"], ["identity", "amplification", "attenuation"], "relu", 16, 4),      # the reference default
    (["sum", "var"], ["linear", "inverse_linear"], "identity", 12, 2),
    (["std", "max", "mean", "min", "sum", "var"], ["attenuation", "identity"], "swish", 20, 5),
    (["mean", "min", "max", "std"], ["identity", "amplification", "attenuation"], "elu", 500, 4),      # Fc = 125: scalar stores
    (["max"], ["identity"], "relu", 7, 1)])
def test_pna_aggregate_fwd_bwd(aggregators, scalers, act, F, n):
    """csrc/pna.cu against the oracle restatement of PHMPNAConvSimple.aggregate (fp64); isolated nodes included."""
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import EdgeStructure
    N, E = 61, 260
    ei = _graph(N - 4, E, 5)                     # the last 4 nodes have no edges at all
    E = ei.size(1)
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, F, generator=g)
    ea = torch.randn(E, F, generator=g)
    avg = O.pna_avg_deg(torch.tensor([0, 4, 9, 3, 1]))
    S, T = len(scalers), len(aggregators)
    gout = torch.randn(N, S * T * F, generator=g)
    xo, eo = (t.double().requires_grad_(True) for t in (x, ea))
    msg = O.activation(xo[ei[0]] + eo, act)
    ref = O.pna_aggregate(msg, ei[1], N, n, aggregators, scalers, avg)
    ref.backward(gout.double())
    xs, es = (t.to(DEV).requires_grad_(True) for t in (x, ea))
    s = EdgeStructure(ei.to(DEV), N)
    out = ops.pna_aggregate(xs, es, s, n, aggregators, scalers, avg["log"], avg["lin"], msg_act=act)
    assert out.shape == (N, S * T * F)
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 2e-5, "out")
    # var = E[m^2] - E[m]^2 cancels in fp32: its gradient inherits that relative error through 1/(2 std)
    loose = 20 if any(a in ("std", "var") for a in aggregators) else 1
    assert_close(xs.grad.cpu(), xo.grad.float(), loose * RTOL, loose * grad_tol(xo.grad, RTOL), "dx")
    assert_close(es.grad.cpu(), eo.grad.float(), loose * RTOL, loose * grad_tol(eo.grad, RTOL), "dea")
    # run-to-run bitwise determinism
    out2 = ops.pna_aggregate(xs.detach(), es.detach(), s, n, aggregators, scalers, avg["log"], avg["lin"], msg_act=act)
    assert torch.equal(out2, out.detach())


def test_aggregate_empty_rows_and_isolated_nodes():
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import EdgeStructure
    N, F = 9, 8
    ei = torch.tensor([[0, 1, 2, 2], [1, 1, 1, 3]])
    x, ea = torch.randn(N, F), torch.randn(4, F)
    s = EdgeStructure(ei.to(DEV), N)
    for red in ("add", "mean", "max", "min", "softmax"):
        ref = O.propagate(x, ei, ea, red, "identity", torch.tensor(1.0))
        out = ops.aggregate(x.to(DEV), ea.to(DEV), s, red, "identity", torch.tensor(1.0, device=DEV), False)
        assert_close(out.cpu(), ref, RTOL, 1e-6, red)
        assert float(out[5:].abs().max()) == 0.0          # untouched rows are exactly zero


def test_aggregate_is_bitwise_deterministic():
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import EdgeStructure
    N, F = 2000, 64
    ei = _graph(N, 30000, 5)
    x = torch.randn(N, F, device=DEV, requires_grad=True)
    ea = torch.randn(ei.size(1), F, device=DEV, requires_grad=True)
    beta = torch.tensor(1.0, device=DEV, requires_grad=True)
    outs = []
    for _ in range(2):
        s = EdgeStructure(ei.to(DEV), N)
        out = ops.aggregate(x, ea, s, "softmax", "relu", beta, True)
        gx, ge, gb = torch.autograd.grad(out.square().sum(), (x, ea, beta))
        outs.append((out.detach().clone(), gx, ge, gb))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


@pytest.mark.parametrize("gated,F,n", [(False, 16, 4), (True, 16, 4), (True, 20, 5), (False, 9, 3), (True, 9, 3), (True, 500, 4)])
def test_pool_fwd_bwd(gated, F, n):
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import SegmentStructure
    sizes = [5, 1, 0, 17, 40, 3]
    batch = torch.cat([torch.full((s,), i, dtype=torch.int64) for i, s in enumerate(sizes)])
    N, B = batch.numel(), len(sizes)
    x = torch.randn(N, F)
    z = torch.randn(N, F // n) if gated else None
    gout = torch.randn(B, F)
    xo = x.double().requires_grad_(True)
    zo = z.double().requires_grad_(True) if gated else None
    xin = (xo.reshape(N, n, -1) * torch.sigmoid(zo)[:, None, :]).reshape(N, F) if gated else xo
    ref = O.seg_sum(xin, batch, B)
    ref.backward(gout.double())
    xs = x.to(DEV).requires_grad_(True)
    zs = z.to(DEV).requires_grad_(True) if gated else None
    seg = SegmentStructure(batch.to(DEV), B)
    out = ops.segment_pool(xs, batch.to(DEV), seg, n, gate_logits=zs)
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 1e-5, "pool")
    assert_close(xs.grad.cpu(), xo.grad.float(), RTOL, 1e-5, "dx")
    if gated:
        assert_close(zs.grad.cpu(), zo.grad.float(), RTOL, 1e-5, "dz")


@pytest.mark.parametrize("M,F,n,act,affine,skip,training", [
    (64, 16, 4, "relu", True, True, True), (300, 20, 5, "identity", True, False, True), (257, 12, 3, "swish", True, True, True),
    (129, 9, 3, "elu", False, True, True), (64, 16, 4, "relu", True, True, False), (1000, 200, 4, "lrelu", True, True, True),
    (6, 8, 2, "selu", True, False, True)])
def test_norm_act_skip(M, F, n, act, affine, skip, training):
    from phc.hypercomplex.norm import PHMNorm
    g = torch.Generator().manual_seed(1)
    h = torch.randn(M, F, generator=g) * 2.0 + 3.0          # non-zero mean: exercises the variance algorithm
    sk = torch.randn(M, F, generator=g) if skip else None
    gout = torch.randn(M, F, generator=g)
    norm = PHMNorm(F, n, affine=affine)
    p = {}
    with torch.no_grad():
        for c, bn in enumerate(norm.bn.bn):
            if affine:
                bn.weight.copy_(1 + 0.3 * torch.randn(F // n, generator=g)); bn.bias.copy_(0.2 * torch.randn(F // n, generator=g))
            bn.running_mean.copy_(0.1 * torch.randn(F // n, generator=g)); bn.running_var.copy_(1 + torch.rand(F // n, generator=g))
            for k in ("weight", "bias", "running_mean", "running_var"):
                v = getattr(bn, k)
                p[f"n.bn.bn.{c}.{k}"] = None if v is None else v.detach().clone().double()
    for k, v in p.items():
        if v is not None and "running" not in k:
            v.requires_grad_(True)
    ho = h.double().requires_grad_(True)
    ref = O.activation(O.phm_norm(ho, p, "n", n, training), act)
    if skip:
        ref = ref + sk.double()
    ref.backward(gout.double())
    norm = norm.to(DEV)
    norm.train(training)
    hs = h.to(DEV).requires_grad_(True)
    out = norm.fused(hs, skip=sk.to(DEV) if skip else None, act=act)
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 2e-5, "y")
    assert_close(hs.grad.cpu(), ho.grad.float(), 2e-4, grad_tol(ho.grad, 2e-4), "dh")
    for c, bn in enumerate(norm.bn.bn):
        if affine:
            assert_close(bn.weight.grad.cpu(), p[f"n.bn.bn.{c}.weight"].grad.float(), 2e-4, grad_tol(p[f"n.bn.bn.{c}.weight"].grad, 2e-4), "dgamma")
            assert_close(bn.bias.grad.cpu(), p[f"n.bn.bn.{c}.bias"].grad.float(), 2e-4, grad_tol(p[f"n.bn.bn.{c}.bias"].grad, 2e-4), "dbeta")
        assert_close(bn.running_mean.cpu(), p[f"n.bn.bn.{c}.running_mean"].float(), RTOL, 1e-5, "running_mean")
        assert_close(bn.running_var.cpu(), p[f"n.bn.bn.{c}.running_var"].float(), RTOL, 1e-5, "running_var")
        assert int(bn.num_batches_tracked) == (1 if training else 0)


@pytest.mark.parametrize("M,F,Fr,n,act,use_norm", [(257, 48, 24, 4, "relu", True), (1000, 500, 500, 4, "relu", True), (64, 10, 7, 1, "swish", True),
                                                  (300, 36, 12, 3, "elu", False), (33, 8, 9, 2, "identity", True)])
def test_norm_act_into_concat_buffer(M, F, Fr, n, act, use_norm):
    """concat_right: [ act(norm(h)) | right ] as ONE buffer (strided store / strided gradient read) == torch.cat of the two,
    values and every gradient bit for bit (same kernels, only the row stride differs)."""
    from phc.hypercomplex.norm import PHMNorm
    from phc_gnn_b200.nn import norm_act_drop_skip
    g = torch.Generator().manual_seed(2)
    h = (torch.randn(M, F, generator=g) * 1.5 + 0.5).to(DEV)
    right = torch.randn(M, Fr, generator=g).to(DEV)
    gout = torch.randn(M, F + Fr, generator=g).to(DEV)
    res = []
    for fused in (True, False):
        torch.manual_seed(5)
        norm = PHMNorm(F, n).to(DEV) if use_norm else None
        hs, rs = h.clone().requires_grad_(True), right.clone().requires_grad_(True)
        if fused:
            out = norm_act_drop_skip(norm, hs, None, act, n, True, concat_right=rs)
        else:
            out = torch.cat([norm_act_drop_skip(norm, hs, None, act, n, True), rs], dim=-1)
        assert out.shape == (M, F + Fr) and out.is_contiguous()
        out.backward(gout)
        res.append((out.detach(), hs.grad, rs.grad) + (tuple(p.grad for p in norm.parameters()) if use_norm else ()))
    for a, b in zip(*res):
        assert torch.equal(a, b)


def test_dropout_invariants():
    # mirrors reference phc/hypercomplex/tests/test_ops_equal_quaternion.py:62-103
    from phc.hypercomplex.layers import phm_dropout
    n, M, Fc, p = 4, 512, 32, 0.3
    x = torch.randn(M, n * Fc, device=DEV) + 5.0
    torch.manual_seed(0)
    y = phm_dropout(x, n, p=p, training=True, same=False)
    kept = y != 0
    assert_close(y[kept], (x / (1 - p))[kept], 1e-6, 1e-6, "scale")
    frac = kept.float().mean().item()
    assert abs(frac - (1 - p)) < 0.02
    y2 = phm_dropout(x, n, p=p, training=True, same=True)
    z = (y2 == 0).view(M, n, Fc)
    assert bool((z == z[:, :1]).all())                     # identical zero pattern across components
    assert abs((~z).float().mean().item() - (1 - p)) < 0.03
    assert phm_dropout(x, n, p=p, training=False) is x
    assert phm_dropout(x, n, p=0.0, training=True) is x
    # backward uses the same mask as forward
    xr = x.clone().requires_grad_(True)
    torch.manual_seed(5)
    yr = phm_dropout(xr, n, p=p, training=True)
    yr.sum().backward()
    assert torch.equal(xr.grad != 0, yr != 0)
    torch.manual_seed(5)
    assert torch.equal(phm_dropout(x, n, p=p, training=True), yr.detach())     # seeded => reproducible
    with pytest.raises(AssertionError):
        phm_dropout(x, n, p=1.5)


@pytest.mark.parametrize("n,dims,Fc,R", [(4, [119, 4, 12, 12, 10, 6, 6, 2, 2], 8, 333), (2, [28], 6, 100), (4, [5, 6, 2], 50, 2000),
                                         (3, [1], 5, 40), (5, [4], 4, 1)])
def test_embedding_encoder(n, dims, Fc, R):
    from phc.hypercomplex.encoder import PHMEncoder
    g = torch.Generator().manual_seed(2)
    enc = PHMEncoder(Fc, dims, n)
    idx = torch.stack([torch.randint(0, d, (R,), generator=g) for d in dims], 1)
    if len(dims) == 1:
        idx = idx[:, 0]
    p = {f"e.{k}": v.detach().clone().double().requires_grad_(True) for k, v in enc.state_dict().items()}
    ref = O.encoder(idx, p, "e", n, dims, torch.float64)
    gout = torch.randn(R, n * Fc, generator=g)
    ref.backward(gout.double())
    enc = enc.to(DEV)
    out = enc(idx.to(DEV))
    assert out.shape == (R, n, Fc)
    out.reshape(R, -1).backward(gout.to(DEV))
    assert_close(out.detach().cpu().reshape(R, -1), ref.detach().float(), 1e-6, 1e-6, "embed")
    for k, v in enc.named_parameters():
        assert_close(v.grad.cpu(), p["e." + k].grad.float(), RTOL, grad_tol(p["e." + k].grad, RTOL), k)


@pytest.mark.parametrize("n,D,Fc,R", [(4, 7, 125, 3000), (4, 3, 56, 500), (2, 1, 6, 77), (4, 5, 5, 10)])
def test_linear_encoder(n, D, Fc, R):
    from phc.hypercomplex.encoder import PHMEncoder
    g = torch.Generator().manual_seed(2)
    enc = PHMEncoder(Fc, D, n)
    feat = torch.rand(R, D, generator=g)
    p = {f"e.{k}": v.detach().clone().double().requires_grad_(True) for k, v in enc.state_dict().items()}
    ref = O.encoder(feat, p, "e", n, D, torch.float64)
    gout = torch.randn(R, n * Fc, generator=g)
    ref.backward(gout.double())
    enc = enc.to(DEV)
    out = enc.flat(feat.to(DEV))
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 1e-5, "linear encoder")
    for k, v in enc.named_parameters():
        assert_close(v.grad.cpu(), p["e." + k].grad.float(), RTOL, grad_tol(p["e." + k].grad, RTOL), k)


def _phm_linear_case(n, fin, fout, M, precision, rtol, seed=0):
    from phc_gnn_b200 import ops
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(n, n, n, generator=g) * 0.5
    W = torch.randn(n, fin // n, fout // n, generator=g) * 0.2
    b = torch.randn(fout, generator=g)
    x = torch.randn(M, fin, generator=g)
    res = torch.randn(M, fout, generator=g)
    gy = torch.randn(M, fout, generator=g)
    po = {"l.phm_rule": A.double().requires_grad_(True), "l.W": W.double().requires_grad_(True), "l.b": b.double().requires_grad_(True)}
    xo = x.double().requires_grad_(True)
    ro = res.double().requires_grad_(True)
    ref = O.phm_linear(xo, po, "l") + ro
    ref.backward(gy.double())
    t = [v.to(DEV).requires_grad_(True) for v in (x, A, W, b, res)]
    y = ops.phm_linear(t[0], t[1], t[2], t[3], t[4], precision=precision)
    y.backward(gy.to(DEV))
    scale = float(ref.detach().abs().max())
    assert_close(y.detach().cpu(), ref.detach().float(), rtol, rtol * scale, "y")
    for got, want, nm in ((t[0].grad, xo.grad, "dx"), (t[1].grad, po["l.phm_rule"].grad, "dA"), (t[2].grad, po["l.W"].grad, "dW"),
                          (t[3].grad, po["l.b"].grad, "db"), (t[4].grad, ro.grad, "dres")):
        assert_close(got.cpu(), want.float(), rtol, grad_tol(want, rtol), nm)


@pytest.mark.parametrize("n,fin,fout,M", [(4, 16, 24, 9), (2, 10, 6, 7), (3, 9, 12, 5), (5, 20, 10, 6), (1, 7, 5, 4), (8, 64, 32, 130),
                                          (4, 200, 200, 333), (4, 500, 500, 300), (2, 180, 180, 257), (4, 224, 224, 1000),
                                          (4, 512, 768, 64), (16, 32, 48, 50)])
def test_phm_linear_fp32(n, fin, fout, M):
    _phm_linear_case(n, fin, fout, M, precision=0, rtol=RTOL)


@pytest.mark.parametrize("n,fin,fout,M", [(4, 128, 128, 512), (4, 500, 500, 1000), (2, 180, 180, 3000), (4, 200, 200, 3333),
                                          (1, 224, 56, 9000), (5, 200, 200, 777), (8, 512, 512, 640), (4, 512, 768, 600),
                                          (3, 33, 300, 515), (16, 512, 64, 520), (4, 224, 224, 8936), (4, 500, 512, 64), (4, 512, 256, 33),
                                          (2, 180, 80, 128), (4, 256, 148, 64), (1, 500, 125, 200)])
def test_phm_linear_tf32x3_tensor_core(n, fin, fout, M):
    """tcgen05 path (3-term tf32 split): fp32-class accuracy, same rtol as the FFMA path."""
    _phm_linear_case(n, fin, fout, M, precision=1, rtol=RTOL)


BF16_RTOL = 2e-2    # stated tolerance of the bf16-operand tensor-core mode (8-bit mantissa operands, fp32 accumulate)


@pytest.mark.parametrize("n,fin,fout,M", [(4, 500, 500, 1000), (2, 180, 180, 3000), (4, 200, 200, 3333), (1, 224, 56, 2000),
                                          (5, 200, 200, 777), (4, 512, 256, 64)])
def test_phm_linear_bf16_tensor_core(n, fin, fout, M):
    """"bf16" precision mode: operands rounded to bf16, single tensor-core pass, fp32 accumulation."""
    _phm_linear_case(n, fin, fout, M, precision=2, rtol=BF16_RTOL)


def test_phm_linear_tensor_core_is_deterministic():
    from phc_gnn_b200 import ops
    g = torch.Generator().manual_seed(3)
    n, F, M = 4, 200, 3000
    t = [v.to(DEV).requires_grad_(True) for v in (torch.randn(M, F, generator=g), torch.randn(n, n, n, generator=g),
                                                 torch.randn(n, F // n, F // n, generator=g), torch.randn(F, generator=g))]
    outs = []
    for _ in range(2):
        y = ops.phm_linear(*t, precision=1)
        outs.append((y.detach().clone(),) + torch.autograd.grad(y.square().sum(), t))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_phm_linear_known_answers(ops_golden):
    from phc_gnn_b200 import ops
    for key, fx in ops_golden.items():
        if not key.startswith("phmlinear"):
            continue
        t = {k: fx[k].to(DEV).requires_grad_(k in ("x", "A", "W", "b")) for k in ("x", "A", "W", "b", "gy")}
        y = ops.phm_linear(t["x"], t["A"], t["W"], t["b"], precision=0)
        assert_close(y.detach().cpu(), fx["y"], RTOL, 1e-5, key)
        y.backward(t["gy"])
        for got, want in ((t["x"].grad, fx["gx"]), (t["A"].grad, fx["gA"]), (t["W"].grad, fx["gW"]), (t["b"].grad, fx["gb"])):
            assert_close(got.cpu(), want, RTOL, grad_tol(want, RTOL), key)


def test_weight_regulariser():
    from phc.hypercomplex.layers import PHMLinear
    from phc.hypercomplex.regularization import phm_weight_regularization
    torch.manual_seed(0)
    m = torch.nn.Sequential(PHMLinear(16, 24, 4), PHMLinear(24, 8, 4), PHMLinear(10, 6, 2), PHMLinear(500, 500, 4)).to(DEV)
    with torch.no_grad():
        m[0].W[:, 0, 0] = 0.0                                  # zero-norm column: gradient must be 0, not NaN
    reg = phm_weight_regularization(m, p=2)
    (3.0 * reg).backward()
    ref = 0.0
    ws = [l.W.detach().double().cpu().requires_grad_(True) for l in m]
    for w in ws:
        ref = ref + w.norm(p=2, dim=0).mean()
    (3.0 * ref).backward()
    assert_close(reg.detach().cpu(), ref.detach().float(), RTOL, 1e-6, "reg")
    for l, w in zip(m, ws):
        g = torch.nan_to_num(w.grad.float())
        assert_close(l.W.grad.cpu(), g, RTOL, grad_tol(g, RTOL), "dW")
    r1 = phm_weight_regularization(m, p=1)                      # other norms take the generic path
    assert torch.isfinite(r1)


@pytest.mark.parametrize("reduce,act,linear,F,n", [
    ("add", "identity", True, 500, 4), ("mean", "relu", True, 224, 4), ("max", "identity", True, 500, 4), ("min", "swish", True, 20, 5),
    ("softmax", "identity", False, 200, 4), ("softmax", "relu", False, 16, 4), ("add", "identity", False, 512, 4),
    ("max", "elu", False, 180, 2), ("mean", "identity", False, 12, 3), ("softmax", "swish", True, 64, 1)])
def test_conv_aggregate_with_fused_encoder(reduce, act, linear, F, n):
    """Aggregation with the edge encoder fused in == encoder followed by aggregation (oracle, fp64)."""
    from phc.hypercomplex.encoder import PHMEncoder
    from phc_gnn_b200 import ops
    from phc_gnn_b200.graph import EdgeStructure
    N, E = 301, 2500
    ei = _graph(N, E, 9)
    E = ei.size(1)
    g = torch.Generator().manual_seed(4)
    dims = 7 if linear else [5, 6, 2]
    enc = PHMEncoder(F // n, dims, n)
    attr = torch.rand(E, 7, generator=g) if linear else torch.stack([torch.randint(0, d, (E,), generator=g) for d in dims], 1)
    x = torch.randn(N, F, generator=g)
    beta = torch.tensor(0.9)
    gout = torch.randn(N, F, generator=g)
    po = {f"e.{k}": v.detach().clone().double().requires_grad_(True) for k, v in enc.state_dict().items()}
    xo, bo = x.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = O.propagate(xo, ei, O.encoder(attr, po, "e", n, dims, torch.float64), reduce, act, bo) + xo
    ref.backward(gout.double())
    enc = enc.to(DEV)
    assert enc.can_fuse(F)
    linear_, params, vocab = enc.fusable_params()
    xs, bs = x.to(DEV).requires_grad_(True), beta.to(DEV).requires_grad_(True)
    s = EdgeStructure(ei.to(DEV), N)
    out = ops.conv_aggregate_fused(xs, attr.to(DEV), s, linear=linear_, params=params, phm_dim=n, vocab=vocab, reduce=reduce,
                                   msg_act=act, beta=bs, self_loop=True)
    out.backward(gout.to(DEV))
    assert_close(out.detach().cpu(), ref.detach().float(), RTOL, 2e-5, "out")
    assert_close(xs.grad.cpu(), xo.grad.float(), RTOL, grad_tol(xo.grad, RTOL), "dx")
    for k, v in enc.named_parameters():
        assert_close(v.grad.cpu(), po["e." + k].grad.float(), 2e-4, grad_tol(po["e." + k].grad, 2e-4), k)
    if reduce == "softmax":
        assert_close(bs.grad.cpu(), bo.grad.float(), 1e-3, grad_tol(bo.grad, 1e-3), "dbeta")


@pytest.mark.parametrize("fin,fout,M,residual", [(500, 500, 1000, False), (256, 256, 4099, True), (200, 200, 333, False)])
def test_phm_linear_epilogue_batch_norm_statistics(fin, fout, M, residual):
    """phc_phm_linear_fwd_bnstats: the tensor-core epilogue's per-32-row chunk moments, merged by
    phc_bn_act_drop_skip_fwd_partials, give the batch-norm of the separate stats kernel."""
    import ctypes
    from phc_gnn_b200 import _lib
    from phc_gnn_b200.graph import _stream
    from phc_gnn_b200.ops import run
    lib = _lib.load()
    n = 4
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, fin, generator=g).to(DEV)
    A = torch.randn(n, n, n, generator=g).to(DEV)
    W = (torch.randn(n, fin // n, fout // n, generator=g) * 0.1).to(DEV)
    b = torch.randn(fout, generator=g).to(DEV)
    res = torch.randn(M, fout, generator=g).to(DEV) if residual else None
    gamma = torch.rand(fout, generator=g).to(DEV) + 0.5
    beta = torch.randn(fout, generator=g).to(DEV)
    st = _stream(torch.device(DEV))
    y = torch.empty(M, fout, device=DEV)
    nb = lib.phc_phm_linear_fwd_workspace_bytes(M, fin, fout, n, 1)
    ws = torch.empty(nb, dtype=torch.uint8, device=DEV)
    part = torch.full(((M + 31) // 32, 2, fout), float("nan"), device=DEV)
    produced = ctypes.c_int(0)
    run("phc_phm_linear_fwd_bnstats", None, x.data_ptr(), A.data_ptr(), W.data_ptr(), b.data_ptr(), 0 if res is None else res.data_ptr(),
        y.data_ptr(), M, fin, fout, n, 0, 1, ws.data_ptr(), ws.numel(), part.data_ptr(), ctypes.addressof(produced), st)
    torch.cuda.synchronize()
    assert produced.value == 1, "the n = 4 tensor-core path did not report batch-norm partials"
    assert torch.isfinite(part).all()
    outs = []
    for use_partials in (True, False):
        rm, rv = torch.zeros(fout, device=DEV), torch.ones(fout, device=DEV)
        tracked = torch.zeros(n, dtype=torch.int64, device=DEV)
        o = torch.empty_like(y)
        mean, rstd = torch.empty(fout, device=DEV), torch.empty(fout, device=DEV)
        if use_partials:
            run("phc_bn_act_drop_skip_fwd_partials", None, y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                tracked.data_ptr(), n, 0, M, fout, n, 1, 0.1, 1e-5, 1, 0.0, 0, 0, o.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                part.data_ptr(), 32, st)
        else:
            w2 = torch.empty(lib.phc_bn_workspace_bytes(M, fout), dtype=torch.uint8, device=DEV)
            run("phc_bn_act_drop_skip_fwd", None, y.data_ptr(), gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                tracked.data_ptr(), n, 0, M, fout, n, 1, 1, 0.1, 1e-5, 1, 0.0, 0, 0, o.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                w2.data_ptr(), w2.numel(), st)
        torch.cuda.synchronize()
        outs.append((o.cpu(), mean.cpu(), rstd.cpu(), rm.cpu(), rv.cpu(), tracked.cpu()))
    for a, c, what in zip(outs[0], outs[1], ("y", "mean", "rstd", "running_mean", "running_var", "tracked")):
        assert_close(a.float(), c.float(), 1e-5, 1e-5, what)
    y64 = y.double().cpu()
    assert_close(outs[0][1].double(), y64.mean(0), 1e-5, 1e-6, "mean vs fp64")
    assert_close(outs[0][2].double(), 1.0 / torch.sqrt(y64.var(0, unbiased=False) + 1e-5), 1e-4, 1e-6, "rstd vs fp64")


@pytest.mark.parametrize("reduce,linear,F,n,self_loop", [("add", True, 500, 4, True), ("mean", True, 64, 4, False),
                                                         ("add", False, 200, 4, True), ("mean", False, 180, 2, True)])
def test_conv_forward_from_node_sums(reduce, linear, F, n, self_loop):
    """phc_conv_fused_fwd_sums (encoder applied to per-node feature sums, pure row gather) == the per-edge fused kernel
    == encoder followed by aggregation (fp64 oracle); isolated nodes included."""
    import ctypes
    from phc.hypercomplex.encoder import PHMEncoder
    from phc_gnn_b200 import _lib, ops
    from phc_gnn_b200.graph import EdgeStructure, _stream
    from phc_gnn_b200.ops import _ptr_array, run
    N, E = 333, 2900
    ei = _graph(N - 5, E, 11)                                      # the last five nodes have no edges
    E = ei.size(1)
    g = torch.Generator().manual_seed(6)
    dims = 7 if linear else [5, 6, 2]
    enc = PHMEncoder(F // n, dims, n)
    attr = torch.rand(E, 7, generator=g) if linear else torch.stack([torch.randint(0, d, (E,), generator=g) for d in dims], 1)
    x = torch.randn(N, F, generator=g)
    po = {f"e.{k}": v.detach().clone().double() for k, v in enc.state_dict().items()}
    ref = O.propagate(x.double(), ei, O.encoder(attr, po, "e", n, dims, torch.float64), reduce, "identity", torch.tensor(1.0).double())
    if self_loop:
        ref = ref + x.double()
    enc = enc.to(DEV)
    linear_, params, vocab = enc.fusable_params()
    s = EdgeStructure(ei.to(DEV), N)
    xs, at = x.to(DEV), attr.to(DEV)
    at = at.float().contiguous() if linear_ else at.long().contiguous()
    edge = ops.conv_aggregate_fused(xs, at, s, linear=linear_, params=params, phm_dim=n, vocab=vocab, reduce=reduce, msg_act="identity",
                                    beta=None, self_loop=self_loop)
    lib = _lib.load()
    st = _stream(torch.device(DEV))
    red = ops.REDUCE_IDS[reduce]
    enc_dim = at.size(1)
    rows = enc_dim + 1 if linear_ else int(sum(vocab))
    vc = (ctypes.c_int * max(len(vocab), 1))(*vocab) if vocab else None
    sums = torch.empty((N, rows), device=DEV)
    run("phc_edge_feature_sums", None, at.data_ptr(), 0 if linear_ else 1, enc_dim, vc, s.rowptr.data_ptr(), s.perm.data_ptr(), N,
        int(red == 1), sums.data_ptr(), st)
    out = torch.empty_like(xs)
    ws = torch.empty(lib.phc_conv_fused_fwd_sums_workspace_bytes(F, rows), dtype=torch.uint8, device=DEV)
    run("phc_conv_fused_fwd_sums", None, xs.data_ptr(), sums.data_ptr(), 0 if linear_ else 1, enc_dim, vc, _ptr_array(params),
        s.rowptr.data_ptr(), s.col.data_ptr(), N, F, n, red, int(self_loop), out.data_ptr(), ws.data_ptr(), ws.numel(), st)
    torch.cuda.synchronize()
    assert_close(out.cpu(), ref.float(), RTOL, 5e-5, "node-sums kernel vs fp64 oracle")
    assert_close(out.cpu(), edge.detach().cpu(), RTOL, 5e-5, "node-sums kernel vs per-edge kernel")


@pytest.mark.parametrize("kind,B,C", [("ce", 64, 37), ("ce", 128, 10), ("bce", 128, 1), ("bce_masked", 512, 128), ("l1", 128, 1)])
def test_fused_task_loss_matches_the_eager_expression(kind, B, C, monkeypatch):
    """phc_task_loss (loss + gradient in one launch, csrc/loss.cu) against the eager chain of the reference's train() bodies
    (train_hiv.py:174-178, train_zinc.py:192, train_ppa.py:200) on the same logits: value and d(loss)/d(logits), also under a
    non-unit upstream gradient."""
    from phc_gnn_b200 import train
    g = torch.Generator().manual_seed(3)
    logits = (torch.randn(B, C, generator=g) * 3).to(DEV)
    if kind == "ce":
        y = torch.randint(0, C, (B,), generator=g).to(DEV)
    elif kind == "l1":
        y = torch.randn(B, generator=g).to(DEV)
    else:
        y = torch.randint(0, 2, (B, C), generator=g).float()
        if kind == "bce_masked":
            y[torch.rand(B, C, generator=g) < 0.4] = float("nan")
        y = y.to(DEV)
    outs = []
    for fused in (False, True):
        monkeypatch.setattr(train, "FUSED_LOSS", fused)
        l = logits.clone().requires_grad_(True)
        loss = train.task_loss(l, y, kind)
        (loss * 0.37).backward()
        outs.append((loss.detach().cpu(), l.grad.cpu()))
    assert_close(outs[1][0], outs[0][0], 1e-5, 1e-6, f"{kind}: loss")
    assert_close(outs[1][1], outs[0][1], 1e-5, 1e-7, f"{kind}: d loss / d logits")


def test_embedding_index_validation_reports_what_the_kernels_clamp():
    """nn.Embedding raises on an index outside its table; the embedding kernels clamp it (a device assertion would poison the context) and
    ops.validate_indices reports it."""
    from phc_gnn_b200 import ops
    vocab = [5, 6, 2]
    g = torch.Generator().manual_seed(1)
    idx = torch.stack([torch.randint(0, d, (300,), generator=g) for d in vocab], 1).to(DEV)
    ops.validate_indices(idx, vocab)                                  # in range: silent
    bad = idx.clone()
    bad[123, 1] = 6
    with pytest.raises(IndexError):
        ops.validate_indices(bad, vocab)
    bad = idx.clone()
    bad[7, 0] = -1
    with pytest.raises(IndexError):
        ops.validate_indices(bad, vocab)
    tables = [torch.randn(d, 8, generator=g).to(DEV) for d in vocab]
    out = ops.embed_sum(bad, tables, 1, vocab)                        # clamped, finite, no fault
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
