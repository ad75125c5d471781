"""Generate tests/golden/family/*.pt: the reference's OTHER parameter layouts run by the UNMODIFIED reference
(dev container only — needs /root/reference).

    python oracle/make_golden_family.py

TEST INFRASTRUCTURE (see oracle/make_golden.py for how the reference is imported).  Two kinds of fixture:

* ``legacy_phmlinear.pt`` — ``PHMLinear_Old`` (reference phc/hypercomplex/layers.py:114-192, the layout of the shipped
  checkpoints) with seeded parameters: legacy state dict, input, output, and the gradients of input and parameters;
* ``quaternion_*.pt`` — the reference's ``QuaternionSkipConnectAdd`` (phc/quaternion/undirectional/models.py:25-230) and,
  for the names containing "concat", ``QuaternionSkipConnectConcat`` (:233-448) on tiny seeded configurations: quaternion-layout state dict, batch, train/eval logits, loss, regulariser
  (``quaternion_weight_regularization``), gradients under the quaternion parameter names, running statistics.

Before anything is written the script checks, with the reference alone, the two layout relations that
phc_gnn_b200/legacy.py implements: PHMLinear_Old == PHMLinear on converted parameters, and the quaternion model ==
the reference's PHMSkipConnectAdd(phm_dim=4, learn_phm=False) on converted parameters.
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

import types  # noqa: E402
_ref_phc = types.ModuleType("phc")
_ref_phc.__path__ = ["/root/reference/phc"]
sys.modules["phc"] = _ref_phc

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from phc_gnn_b200.synthetic import workloads, tiny, make_batch  # noqa: E402
from phc_gnn_b200 import legacy  # noqa: E402

PHM_ONLY = ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type")


def quaternion_cases():
    w4 = workloads(4)
    out = {}
    c = tiny(w4["hiv"], 16, 2, 6, 5, 9, head=[12, 8]); out["quaternion_hiv_softmax_mlp"] = c
    c = tiny(w4["zinc"], 12, 2, 6, 4, 9, head=[12, 8]); out["quaternion_zinc_sum_mlp"] = c
    c = tiny(w4["pcba"], 16, 2, 7, 4, 9, head=[24, 8]); c.model.update(target_dim=5, activation="swish"); out["quaternion_pcba_sum_lin"] = c
    c = tiny(w4["mnist"], 16, 2, 4, 7, 10, head=[16, 8]); c.extra["k"] = 3
    c.model.update(msg_aggr="softmax", initial_beta=0.8, learn_beta=True, msg_encoder="relu", pooling="globalsum")
    out["quaternion_mnist_softmax_lin"] = c
    # QuaternionSkipConnectConcat (names contain "concat"): layer widths may differ, skip = component-wise concat
    c = tiny(w4["hiv"], 16, 2, 6, 5, 9, head=[12, 8]); c.model.update(mp_layers=[16, 24], msg_aggr="sum", mlp=False)
    out["quaternion_concat_hiv_sum_lin"] = c
    c = tiny(w4["zinc"], 12, 2, 6, 4, 9, head=[12]); c.model.update(mp_layers=[8, 12], msg_aggr="softmax", initial_beta=1.0,
                                                                     learn_beta=True, mlp=True, pooling="globalsum")
    out["quaternion_concat_zinc_softmax_mlp"] = c
    return out


def quaternion_kwargs(phm_kwargs):
    kw = {k: v for k, v in phm_kwargs.items() if k not in PHM_ONLY}
    kw["init"] = "glorot-uniform"
    kw["dropout_mpnn"] = [0.0] * len(kw["mp_layers"])
    kw["dropout_dn"] = [0.0] * len(kw["downstream_layers"])
    return kw


def seeded_fill(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in model.named_parameters():
            leaf = name.rsplit(".", 1)[-1]
            if leaf.startswith("b_") or leaf == "bias":
                prm.copy_(0.1 * torch.randn(prm.shape, generator=g))
            elif leaf == "beta":
                pass
            elif ".bn." in name and leaf == "weight":
                prm.copy_(1.0 + 0.2 * torch.randn(prm.shape, generator=g))
            elif leaf.startswith("W_"):
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


def legacy_fixture(outdir):
    from phc.hypercomplex.layers import PHMLinear, PHMLinear_Old
    g = torch.Generator().manual_seed(11)
    fx = {}
    for n, fin, fout, m in ((4, 5, 7, 9), (2, 6, 3, 5), (5, 4, 4, 6), (3, 2, 5, 4)):      # PER-COMPONENT widths (legacy convention)
        old = PHMLinear_Old(fin, fout, n, c_init="standard")
        with torch.no_grad():
            for a in old.phm_rule:
                # rebind, never write in place: PHMLinear_Old's rule parameters share storage with the reference's module-level
                # rule tables (utils.py), and an in-place update would corrupt them for everything generated afterwards
                a.data = a.data + 0.2 * torch.randn(a.shape, generator=g)
            for w in old.W:
                w.copy_(torch.randn(w.shape, generator=g))
            for b in old.b:
                b.copy_(torch.randn(b.shape, generator=g))
        x = torch.randn(m, n * fin, generator=g, requires_grad=True)
        y = old(x)
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        state = {k: v.detach().clone() for k, v in old.state_dict().items()}
        grads = {k: p.grad.clone() for k, p in old.named_parameters()}
        # relation check inside the reference: the current PHMLinear on converted parameters
        new = PHMLinear(n * fin, n * fout, n, c_init="standard")
        new.load_state_dict(legacy.convert_legacy_phm_state_dict(state), strict=True)
        err = float((new(x.detach()) - y.detach()).abs().max())
        assert err <= 1e-5, f"legacy relation broken for n={n}: {err}"
        back = legacy.to_legacy_phm_state_dict(legacy.convert_legacy_phm_state_dict(state))
        assert all(torch.equal(back[k], state[k]) for k in state) and list(back) == list(state)
        fx[f"n{n}"] = dict(n=n, in_per=fin, out_per=fout, state=state, x=x.detach(), y=y.detach(), gy=gy, gx=x.grad.clone(),
                           grads=grads)
        print(f"legacy PHMLinear_Old n={n} {fin}->{fout}: current-layout max err {err:.2e}")
    torch.save(fx, os.path.join(outdir, "legacy_phmlinear.pt"))


def quaternion_fixtures(outdir):
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    from phc.quaternion.regularization import quaternion_weight_regularization
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    import phc.quaternion.undirectional.models as _ref_models
    assert _ref_models.__file__.startswith("/root/reference/"), f"not the reference: {_ref_models.__file__}"
    # seeds follow the alphabetical order of the Add cases; the Concat cases were appended later (existing fixtures never change)
    for k, (name, wl) in enumerate(sorted(quaternion_cases().items(), key=lambda kv: ("concat" in kv[0], kv[0]))):
        torch.manual_seed(300 + k)
        np.random.seed(300 + k)
        kw = quaternion_kwargs(wl.model)
        concat = "concat" in name
        model = (QuaternionSkipConnectConcat if concat else QuaternionSkipConnectAdd)(**kw)
        seeded_fill(model, 40 + k)
        data = make_batch(wl, seed=50 + k)
        state0 = {n: v.clone() for n, v in model.state_dict().items()}
        model.train()
        logits = model(data)
        reg = quaternion_weight_regularization(model, device="cpu", p=2)
        loss = ref_loss(logits, data.y, wl.loss) + 0.01 * reg
        loss.backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        state1 = {n: v.clone() for n, v in model.state_dict().items() if "running" in n or "tracked" in n}
        model.eval()
        with torch.no_grad():
            logits_eval = model(data)

        err_t = err_e = float('nan')
        if not concat:          # (the reference's PHM concat model raises for n > 1, SURVEY.md D2: nothing to compare with)
            # relation check inside the reference: PHM(n=4, Hamilton rule) on converted parameters
            pkw = dict(wl.model)
            pkw.update(dropout_mpnn=kw["dropout_mpnn"], dropout_dn=kw["dropout_dn"], phm_dim=4, learn_phm=False, sc_type="first")
            phm = PHMSkipConnectAdd(**pkw)
            phm.load_state_dict(legacy.quaternion_to_phm_state_dict(state0), strict=True)
            phm.train()
            err_t = float((phm(data) - logits.detach()).abs().max())
            phm.eval()
            with torch.no_grad():
                err_e = float((phm(data) - logits_eval).abs().max())
            scale = float(logits.detach().abs().max())
            assert err_t <= 2e-5 * max(1.0, scale) and err_e <= 2e-5 * max(1.0, scale), (name, err_t, err_e, scale)
        back = legacy.phm_to_quaternion_state_dict(legacy.quaternion_to_phm_state_dict(state0))
        assert set(back) == set(state0) and all(torch.equal(back[n], state0[n]) for n in state0), name

        fx = dict(name=name, model="concat" if concat else "add", cfg=kw, loss_kind=wl.loss, reg_scale=0.01,
                  data=dict(x=data.x, edge_index=data.edge_index, edge_attr=data.edge_attr, batch=data.batch,
                            y=data.y, num_graphs=data.num_graphs),
                  state=state0, logits_train=logits.detach(), loss=loss.detach(), reg=reg.detach(), grads=grads,
                  running_after=state1, logits_eval=logits_eval, n_params=model.get_number_of_params_())
        path = os.path.join(outdir, name + ".pt")
        torch.save(fx, path)
        print(f"{name:32s} N={data.x.size(0):4d} E={data.edge_index.size(1):4d} params={fx['n_params']:6d} "
              f"loss={float(loss):.6f} phm-vs-quaternion err train {err_t:.1e} eval {err_e:.1e}  {os.path.getsize(path) / 1024:.0f} KiB")


def phm_option_cases():
    """PHM configurations whose constructor options the 14 fixtures of make_golden.py do not exercise: the naive encoders
    (embedding and linear), add_self_loops=False (both convolution types), bias=False, a frozen rule on a PHM model."""
    w4 = workloads(4)
    out = {}
    c = tiny(w4["hiv"], 16, 2, 6, 5, 9, head=[12, 8])
    c.model.update(naive_encoder=True, add_self_loops=False, mlp=False, msg_aggr="sum")
    out["phm_hiv_naive_enc_noloops_lin"] = c
    c = tiny(workloads(2)["zinc"], 12, 2, 6, 4, 9, head=[12, 6])
    c.model.update(bias=False, add_self_loops=False, mlp=True, msg_aggr="mean", learn_phm=False)
    out["phm_zinc_n2_nobias_noloops_mlp_frozen"] = c
    c = tiny(w4["mnist"], 16, 2, 4, 7, 10, head=[16, 8]); c.extra["k"] = 3
    c.model.update(naive_encoder=True, msg_aggr="softmax", initial_beta=1.3, learn_beta=False, mlp=True)
    out["phm_mnist_naive_linear_enc_softmax_fixed_beta"] = c
    # PHMSkipConnectConcat runs in the reference only for phm_dim = 1: pins the concat model's structure
    c = tiny(workloads(1)["hiv"], 8, 2, 6, 4, 8, head=[6, 4]); c.model.update(mp_layers=[8, 12], msg_aggr="sum", mlp=True)
    out["phm_concat_n1_hiv_sum_mlp"] = c
    c = tiny(workloads(1)["mnist"], 6, 2, 4, 7, 10, head=[5]); c.extra["k"] = 3
    c.model.update(mp_layers=[6, 4], msg_aggr="max", mlp=False, pooling="globalsum")
    out["phm_concat_n1_mnist_max_lin"] = c
    return out


def phm_option_fixtures(outdir):
    sys.path.insert(0, HERE)
    from make_golden import seeded_fill as phm_fill        # same seeding rules as the main PHM fixtures
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd, PHMSkipConnectConcat
    from phc.hypercomplex.regularization import phm_weight_regularization
    # seeds follow the alphabetical order of the Add cases; the Concat cases were appended later
    for k, (name, wl) in enumerate(sorted(phm_option_cases().items(), key=lambda kv: ("concat" in kv[0], kv[0]))):
        torch.manual_seed(500 + k)
        np.random.seed(500 + k)
        kw = dict(wl.model)
        kw["dropout_mpnn"] = [0.0] * len(kw["mp_layers"])
        kw["dropout_dn"] = [0.0] * len(kw["downstream_layers"])
        concat = "concat" in name
        model = (PHMSkipConnectConcat if concat else PHMSkipConnectAdd)(**kw)
        phm_fill(model, 60 + k)
        data = make_batch(wl, seed=70 + k)
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
        fx = dict(name=name, model="phm_concat" if concat else "phm", cfg=kw, loss_kind=wl.loss, reg_scale=0.01,
                  data=dict(x=data.x, edge_index=data.edge_index, edge_attr=data.edge_attr, batch=data.batch,
                            y=data.y, num_graphs=data.num_graphs),
                  state=state0, logits_train=logits.detach(), loss=loss.detach(), reg=reg.detach(), grads=grads,
                  running_after=state1, logits_eval=logits_eval, n_params=model.get_number_of_params_())
        path = os.path.join(outdir, name + ".pt")
        torch.save(fx, path)
        print(f"{name:48s} N={data.x.size(0):4d} E={data.edge_index.size(1):4d} params={fx['n_params']:6d} "
              f"loss={float(loss):.6f}  {os.path.getsize(path) / 1024:.0f} KiB")


def init_fixture(outdir):
    """Known answers of the reference's quaternion initialisers under fixed seeds (phc/quaternion/inits.py:40-113)."""
    from phc.quaternion.inits import orthogonal_init, quaternion_init
    fx = {}
    for fin, fout in ((6, 5), (4, 9)):
        np.random.seed(21)
        torch.manual_seed(21)
        fx[f"quaternion_{fin}_{fout}"] = torch.stack(quaternion_init(fin, fout), dim=0)          # [4, out, in]
        torch.manual_seed(22)
        w = torch.stack(orthogonal_init(fin, fout), dim=0).double()
        fx[f"orthogonal_{fin}_{fout}_gram"] = torch.einsum("cok,coj->kj", w, w) if fout >= fin else torch.einsum("cok,cpk->op", w, w)
        fx[f"orthogonal_{fin}_{fout}_std"] = w.std()
    torch.save(fx, os.path.join(outdir, "inits_quaternion.pt"))
    print("inits_quaternion.pt:", {k: tuple(v.shape) for k, v in fx.items()})


INIT_MODEL_KW = dict(phm_dim=4, atom_encoded_dim=16, mp_layers=[16, 16], dropout_mpnn=[0.0, 0.0], downstream_layers=[12, 8], mlp=True,
                     msg_aggr="softmax", initial_beta=1.0, learn_beta=True)


def init_state_fixture(outdir):
    """Initial state dict of a freshly constructed reference model under fixed seeds (numpy + torch): what
    ``reset_parameters`` draws — phm_init (scipy chi, torch uniform, numpy phase), embedding / linear initialisers, rules."""
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    np.random.seed(3)
    torch.manual_seed(3)
    m = PHMSkipConnectAdd(**INIT_MODEL_KW)
    torch.save(dict(kw=INIT_MODEL_KW, seed=3, state={k: v.clone() for k, v in m.state_dict().items()}),
               os.path.join(outdir, "inits_phm_model.pt"))
    print("inits_phm_model.pt:", len(m.state_dict()), "tensors")


def main():
    outdir = os.path.join(ROOT, "tests", "golden", "family")
    os.makedirs(outdir, exist_ok=True)
    legacy_fixture(outdir)
    init_fixture(outdir)
    init_state_fixture(outdir)
    quaternion_fixtures(outdir)
    phm_option_fixtures(outdir)


if __name__ == "__main__":
    main()
