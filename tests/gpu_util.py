"""Helpers shared by the -m gpu parity tests: run the product model (CUDA kernels through the
C ABI) and the CPU oracle on identical inputs/weights."""
import torch

from oracle import phc_oracle as O


def cuda():
    return torch.device("cuda:0")


def product_model(cfg, state, device):
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    m = PHMSkipConnectAdd(**cfg)
    m.load_state_dict(state, strict=True)
    return m.to(device)


def oracle_params(state, dtype=torch.float32):
    p = {}
    for k, v in state.items():
        v = v.detach().cpu().clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if "running" not in k:
                v.requires_grad_(True)
        p[k] = v
    return p


def product_train_eval(cfg, state, batch, loss_kind, reg_scale, device, fuse_edge_encoder=True, fuse_layer=True, bucket=True):
    """-> dict(logits, loss, reg, grads, running, logits_eval) from the CUDA path.  bucket: register a flat gradient buffer for the
    parameters, as FlatClipAdam / DataParallelPHC do — the precondition of the layers' in-place parameter-gradient path."""
    from phc.hypercomplex.regularization import phm_weight_regularization
    m = product_model(cfg, state, device)
    if bucket:
        from phc_gnn_b200.optim import ordered_parameters
        from phc_gnn_b200.parallel import GradientBucket
        m._test_bucket = GradientBucket(ordered_parameters(m))
        m._test_bucket._ensure(torch.device(device))
    m.fuse_edge_encoder = fuse_edge_encoder
    m.fuse_layer = fuse_layer
    data = batch.to(device)
    m.train()
    logits = m(data)
    reg = phm_weight_regularization(m, p=2)
    loss = O.task_loss(logits, data.y, loss_kind) + reg_scale * reg
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if p.grad is not None}
    running = {n: v.detach().cpu().clone() for n, v in m.state_dict().items() if "running" in n or "tracked" in n}
    m.eval()
    with torch.no_grad():
        ev = m(data)
    return dict(logits=logits.detach().cpu(), loss=loss.detach().cpu(), reg=reg.detach().cpu(), grads=grads, running=running,
                logits_eval=ev.cpu(), model=m)


def oracle_train_eval(cfg, state, batch, loss_kind, reg_scale, dtype=torch.float32):
    p = oracle_params(state, dtype)
    logits = O.model_forward(p, cfg, batch, training=True)
    reg = O.weight_regularization(p, 2)
    loss = O.task_loss(logits, batch.y, loss_kind) + reg_scale * reg
    loss.backward()
    grads = {k: v.grad.detach().clone().float() for k, v in p.items() if v.requires_grad and v.grad is not None}
    running = {k: v.detach().clone().float() for k, v in p.items() if "running" in k}
    with torch.no_grad():
        ev = O.model_forward(p, cfg, batch, training=False)
    return dict(logits=logits.detach().float(), loss=loss.detach().float(), reg=reg.detach().float(), grads=grads,
                running=running, logits_eval=ev.float())


def assert_close(a, b, rtol, atol, what=""):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol, msg=lambda m: f"{what}: {m}")


def grad_tol(ref: torch.Tensor, rtol: float):
    """absolute tolerance scaled to the tensor's magnitude (gradients span many orders)."""
    # floor: gradients that are exactly zero in exact arithmetic (e.g. a bias feeding a batch-norm)
    # come out as +-1e-7 rounding noise on both sides
    return max(rtol * float(ref.abs().max()), 2e-6)
