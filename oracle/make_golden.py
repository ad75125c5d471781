"""Generate tests/golden/*.pt by running the UNMODIFIED reference (dev container only).

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

TEST INFRASTRUCTURE.  The reference package is imported from /root/reference with
oracle/refshim standing in for torch_scatter / torch_geometric / ogb (not installable
offline).  For every case we build the reference's ``PHMSkipConnectAdd`` with a tiny
configuration, overwrite every parameter / buffer with seeded values (so the reference's
uninitialised bias element, SURVEY.md D8, is defined), run one train-mode forward+backward
(dropout 0 so no RNG is involved, BN on batch statistics) and one eval-mode forward, and
store inputs, weights, outputs and gradients.  The GPU box never sees /root/reference —
only these small fixtures travel.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

# The reference's ``phc`` has no __init__.py (a namespace package), so the product's drop-in ``phc`` package at the
# repo root would shadow it on import: pin the name to the reference tree explicitly and verify below.
import types  # noqa: E402
_ref_phc = types.ModuleType("phc")
_ref_phc.__path__ = ["/root/reference/phc"]
sys.modules["phc"] = _ref_phc

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from phc_gnn_b200.synthetic import workloads, tiny, make_batch  # noqa: E402


def cases():
    w4 = workloads(4)
    out = {}
    c = tiny(w4["hiv"], 16, 2, 6, 5, 9, head=[12, 8]); out["hiv_n4_softmax_mlp"] = c
    c = tiny(w4["hiv"], 16, 2, 5, 4, 8, head=[8]); c.model.update(msg_encoder="relu", activation="swish"); out["hiv_n4_softmax_relu_msg"] = c
    c = tiny(workloads(2)["zinc"], 12, 3, 6, 4, 9, head=[12, 6]); out["zinc_n2_sum_mlp_last"] = c
    c = tiny(w4["zinc"], 12, 2, 6, 4, 9, head=[12, 8]); out["zinc_n4_sum_mlp_last"] = c
    c = tiny(workloads(5)["zinc"], 20, 2, 5, 4, 9, head=[10]); c.model.update(atom_encoded_dim=20); out["zinc_n5_sum_mlp"] = c
    c = tiny(w4["pcba"], 16, 3, 7, 4, 9, head=[24, 8]); c.model.update(target_dim=5); out["pcba_n4_sum_lin"] = c
    c = tiny(w4["mnist"], 16, 2, 4, 7, 10, head=[16, 8]); c.extra["k"] = 3; out["mnist_n4_mean_lin"] = c
    c = tiny(w4["ppa"], 20, 2, 3, 10, 14, und_edges=30, head=[16, 8]); c.model.update(target_dim=7); out["ppa_n4_sum_mlp"] = c
    c = tiny(w4["ppa"], 20, 2, 3, 10, 14, und_edges=30, head=[16]); c.model.update(target_dim=7, msg_aggr="max"); out["ppa_n4_max_mlp"] = c
    c = tiny(workloads(3)["zinc"], 12, 2, 5, 4, 9, head=[9])
    c.model.update(msg_aggr="min", pooling="globalsum", activation="lrelu", msg_encoder="elu", mlp=False, sc_type="first")
    out["zinc_n3_min_lin_globalsum"] = c
    c = tiny(w4["cifar"], 8, 2, 3, 6, 8, head=[8]); c.extra["k"] = 2
    c.model.update(msg_aggr="softmax", initial_beta=0.7, learn_beta=True, activation="elu", msg_encoder="swish")
    out["cifar_n4_softmax_lin_swish"] = c
    c = tiny(w4["hiv"], 8, 1, 4, 3, 6, head=[8]); c.model.update(norm_mp=None, norm_dn=None, msg_aggr="mean", mlp=True, activation="selu")
    out["hiv_n4_mean_nonorm"] = c
    pna = dict(msg_aggr="pna", aggregators=["mean", "min", "max", "std"], scalers=["identity", "amplification", "attenuation"],
               deg=torch.tensor([0, 9, 31, 22, 7, 2]), post_layers=1)
    c = tiny(w4["hiv"], 16, 2, 6, 5, 9, head=[12, 8]); c.model.update(pna); out["hiv_n4_pna"] = c
    c = tiny(workloads(2)["zinc"], 12, 2, 5, 4, 9, head=[10]); c.model.update(pna)
    c.model.update(aggregators=["sum", "var", "max"], scalers=["linear", "inverse_linear", "identity"], post_layers=2, activation="elu")
    out["zinc_n2_pna_post2"] = c
    return out


def seeded_fill(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            if name.endswith(".b") or name.endswith("bias"):
                prm.copy_(0.1 * torch.randn(prm.shape, generator=g))
            elif name.endswith("phm_rule"):
                prm.add_(0.15 * torch.randn(prm.shape, generator=g))
            elif name.endswith("beta"):
                pass
            elif ".bn." in name and name.endswith("weight"):
                prm.copy_(1.0 + 0.2 * torch.randn(prm.shape, generator=g))
            elif name.endswith(".W"):
                prm.copy_(0.35 * torch.randn(prm.shape, generator=g))
        for name, buf in model.named_buffers():
            if name.endswith("running_mean"):
                buf.copy_(0.1 * torch.randn(buf.shape, generator=g))
            elif name.endswith("running_var"):
                buf.copy_(1.0 + 0.3 * torch.rand(buf.shape, generator=g))


def ref_loss(logits, y, kind):
    if kind in ("bce", "bce_masked"):
        mask = ~torch.isnan(y)
        return F.binary_cross_entropy_with_logits(input=logits[mask], target=y[mask])
    if kind == "l1":
        return (logits.squeeze() - y).abs().mean()
    return F.cross_entropy(logits, y.view(-1))


def main():
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc.hypercomplex.regularization import phm_weight_regularization
    from phc.hypercomplex.layers import PHMLinear
    from phc.hypercomplex.utils import get_multiplication_matrices
    import phc.hypercomplex.undirectional.models as _ref_models
    assert _ref_models.__file__.startswith("/root/reference/"), f"not the reference: {_ref_models.__file__}"
    outdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(outdir, exist_ok=True)
    # fixture k (the seed) follows the alphabetical order of the first twelve cases; later additions are appended
    # so that regenerating never changes an existing fixture
    ordered = sorted(cases().items(), key=lambda kv: ("pna" in kv[0], kv[0]))
    for k, (name, wl) in enumerate(ordered):
        torch.manual_seed(100 + k)
        import numpy as np
        np.random.seed(100 + k)
        kw = dict(wl.model)
        kw["dropout_mpnn"] = [0.0] * len(kw["mp_layers"])
        kw["dropout_dn"] = [0.0] * len(kw["downstream_layers"])
        model = PHMSkipConnectAdd(**kw)
        seeded_fill(model, 7 + k)
        data = make_batch(wl, seed=k)
        state0 = {n: v.clone() for n, v in model.state_dict().items()}
        model.train()
        logits = model(data)
        reg = phm_weight_regularization(model, p=2)
        loss = ref_loss(logits, data.y, wl.loss) + 0.01 * reg
        loss.backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        state1 = {n: v.clone() for n, v in model.state_dict().items() if "running" in n or "tracked" in n}
        model.eval()
        with torch.no_grad():
            logits_eval = model(data)
        fx = dict(name=name, cfg=kw, loss_kind=wl.loss, reg_scale=0.01,
                  data=dict(x=data.x, edge_index=data.edge_index, edge_attr=data.edge_attr, batch=data.batch,
                            y=data.y, num_graphs=data.num_graphs),
                  state=state0, logits_train=logits.detach(), loss=loss.detach(), reg=reg.detach(), grads=grads,
                  running_after=state1, logits_eval=logits_eval, n_params=model.get_number_of_params_())
        path = os.path.join(outdir, name + ".pt")
        torch.save(fx, path)
        print(f"{name:32s} N={data.x.size(0):4d} E={data.edge_index.size(1):4d} params={fx['n_params']:6d} "
              f"loss={float(loss):.6f} {os.path.getsize(path) / 1024:.0f} KiB")

    # op-level known answers: rule matrices and PHMLinear forward/backward for several n
    ops = {}
    for n in (1, 2, 3, 4, 5, 8):
        ops[f"rule_standard_{n}"] = torch.stack(get_multiplication_matrices(n, type="standard"), 0)
    g = torch.Generator().manual_seed(5)
    for n, fin, fout, m in ((4, 16, 24, 9), (2, 10, 6, 7), (3, 9, 12, 5), (5, 20, 10, 6), (1, 7, 5, 4)):
        lin = PHMLinear(fin, fout, n, c_init="standard")
        with torch.no_grad():
            lin.phm_rule.add_(0.2 * torch.randn(lin.phm_rule.shape, generator=g))
            lin.W.copy_(torch.randn(lin.W.shape, generator=g))
            lin.b.copy_(torch.randn(lin.b.shape, generator=g))
        x = torch.randn(m, fin, generator=g, requires_grad=True)
        y = lin(x)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        ops[f"phmlinear_n{n}"] = dict(x=x.detach(), A=lin.phm_rule.detach().clone(), W=lin.W.detach().clone(),
                                      b=lin.b.detach().clone(), y=y.detach(), gy=gy, gx=x.grad.clone(),
                                      gA=lin.phm_rule.grad.clone(), gW=lin.W.grad.clone(), gb=lin.b.grad.clone())
    torch.save(ops, os.path.join(outdir, "ops.pt"))
    print("ops.pt", os.path.getsize(os.path.join(outdir, "ops.pt")) // 1024, "KiB")


if __name__ == "__main__":
    main()
