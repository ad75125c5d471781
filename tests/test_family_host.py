"""The reference's other parameter layouts on the CPU: the oracle's quaternion / legacy restatements
against vectors recorded from the UNMODIFIED reference (oracle/make_golden_family.py -> tests/golden/family), and the
product's relabelling functions (phc_gnn_b200/legacy.py — tensor re-labelling only, no kernel) against both."""
import os

import pytest
import torch

from conftest import golden_dir
from oracle import phc_oracle as O

RTOL, ATOL = 1e-4, 2e-5
FAMILY = os.path.join(golden_dir(), "family")
SHIPPED = {"hiv": ("/root/reference/benchmarks/hiv/experiment1/run_1/model.pt", 110909),
           "zinc": ("/root/reference/benchmarks/zinc/experiment1/run_1/model.pt", 106291)}


def quaternion_cases():
    return sorted(f[:-3] for f in os.listdir(FAMILY) if f.startswith("quaternion_") and f.endswith(".pt"))


def load_family(name):
    from phc_gnn_b200.synthetic import GraphBatch
    fx = torch.load(os.path.join(FAMILY, name + ".pt"), weights_only=False)
    if "data" in fx:
        d = fx["data"]
        fx["batch"] = GraphBatch(d["x"], d["edge_index"], d["edge_attr"], d["batch"], d["y"], d["num_graphs"])
    return fx


def oracle_forward(fx):
    return O.quaternion_concat_model_forward if fx.get("model") == "concat" else O.quaternion_model_forward


def product_class(fx):
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    return QuaternionSkipConnectConcat if fx.get("model") == "concat" else QuaternionSkipConnectAdd


def leaves(state, dtype=torch.float32):
    p = {}
    for k, v in state.items():
        v = v.clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if "running" not in k:
                v.requires_grad_(True)
        p[k] = v
    return p


@pytest.mark.parametrize("name", quaternion_cases())
def test_quaternion_oracle_matches_reference(name):
    fx = load_family(name)
    pq = leaves(fx["state"])
    data, cfg = fx["batch"], fx["cfg"]
    forward = oracle_forward(fx)
    logits = forward(pq, cfg, data, training=True)
    torch.testing.assert_close(logits, fx["logits_train"], rtol=RTOL, atol=ATOL)
    reg = O.quaternion_weight_regularization(pq, cfg, 2)
    torch.testing.assert_close(reg, fx["reg"], rtol=RTOL, atol=ATOL)
    loss = O.task_loss(logits, data.y, fx["loss_kind"]) + fx["reg_scale"] * reg
    torch.testing.assert_close(loss, fx["loss"], rtol=RTOL, atol=ATOL)
    loss.backward()
    for k, g in fx["grads"].items():
        assert pq[k].grad is not None, k
        torch.testing.assert_close(pq[k].grad, g, rtol=5e-4, atol=5e-5, msg=lambda m: f"{k}: {m}")
    with torch.no_grad():
        ev = forward(pq, cfg, data, training=False)
    torch.testing.assert_close(ev, fx["logits_eval"], rtol=RTOL, atol=ATOL)
    assert sum(v.numel() for v in pq.values() if v.requires_grad) == fx["n_params"]


def test_quaternion_linear_is_phm_with_hamilton_rule():
    """The written-out Hamilton product == the PHM contraction on the relabelled parameters (oracle, both directions
    restated independently), and the product's relabelling produces exactly those PHM tensors."""
    from phc_gnn_b200 import legacy
    g = torch.Generator().manual_seed(3)
    pq = {f"l.W_{c}": torch.randn(5, 3, generator=g) for c in "rijk"}
    pq.update({f"l.b_{c}": torch.randn(5, generator=g) for c in "rijk"})
    x = torch.randn(7, 12, generator=g)
    want = O.quaternion_linear(x, pq, "l")
    got = O.phm_linear(x, O.quaternion_as_phm(pq), "l")
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    prod = legacy.quaternion_to_phm_state_dict(pq)
    ora = O.quaternion_as_phm(pq)
    assert set(prod) == set(ora)
    for k in prod:
        assert torch.equal(prod[k], ora[k]), k
    back = legacy.phm_to_quaternion_state_dict(prod)
    assert set(back) == set(pq) and all(torch.equal(back[k], pq[k]) for k in pq)


@pytest.mark.parametrize("name", quaternion_cases())
def test_quaternion_state_dict_relabelling(name):
    """Reference quaternion state dict -> product model (strict load) -> back, bit-exact; the relabelled parameters run
    through the PHM oracle reproduce the reference's quaternion outputs."""
    from phc_gnn_b200 import legacy
    fx = load_family(name)
    model = product_class(fx)(**fx["cfg"])
    model.load_quaternion_state_dict(fx["state"])
    assert model.get_number_of_params_() == fx["n_params"]
    back = model.quaternion_state_dict()
    assert set(back) == set(fx["state"])
    for k, v in fx["state"].items():
        assert torch.equal(back[k], v), k
    for m in model.modules():
        if hasattr(m, "phm_rule") and isinstance(m.phm_rule, torch.nn.Parameter):
            assert not m.phm_rule.requires_grad
            assert torch.equal(m.phm_rule.detach(), legacy.hamilton_rule())
    if fx.get("model") != "concat":
        p = leaves(model.state_dict())
        cfg = dict(O.quaternion_cfg(fx["cfg"]))
        logits = O.model_forward(p, cfg, fx["batch"], training=True)
        torch.testing.assert_close(logits, fx["logits_train"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("init", ["orthogonal", "quaternion", "glorot-uniform", "glorot-normal"])
def test_quaternion_reset_parameters(init):
    """reset_parameters keeps the Hamilton rule (the PHM layer would re-draw it from c_init, reference layers.py:281) and
    gives the QLinear bias pattern; 'orthogonal' weights have orthonormal quaternion columns (unit norm, as the reference's)."""
    import numpy as np
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd
    from phc_gnn_b200 import legacy
    from phc_gnn_b200.quaternion import _qconj, _qmul
    torch.manual_seed(0)
    np.random.seed(0)
    m = QuaternionSkipConnectAdd(atom_encoded_dim=16, mp_layers=[16, 16], dropout_mpnn=[0.0, 0.0], downstream_layers=[24, 8],
                                 init=init, mlp=True)
    m.reset_parameters()
    lin = m.downstream.affine[0]                         # 4 x [4 -> 6]
    assert torch.equal(lin.phm_rule.detach(), legacy.hamilton_rule())
    assert torch.equal(lin.b.detach(), torch.cat([torch.zeros(6), torch.full((18,), 0.2)]))
    assert torch.isfinite(lin.W).all() and float(lin.W.abs().max()) > 0
    if init == "orthogonal":
        w = lin.W.detach().double().permute(0, 2, 1)                # tall orientation [6 rows, 4 cols]
        gram = torch.stack([torch.stack([_qmul(_qconj(w[:, :, a]), w[:, :, b]).sum(dim=1) for b in range(4)], dim=1)
                            for a in range(4)], dim=1)               # [4, cols, cols]
        eye = torch.zeros_like(gram)
        eye[0] = torch.eye(4, dtype=torch.float64)
        assert float((gram - eye).abs().max()) < 1e-5


def test_scripts_family_switch():
    """benchmarks/train_hiv.py:182 picks the regulariser by ``hasattr(model, "phm_dim")``."""
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    kw = dict(atom_encoded_dim=16, mp_layers=[16, 16], dropout_mpnn=[0.0, 0.0], downstream_layers=[16, 8])
    assert hasattr(PHMSkipConnectAdd(**kw), "phm_dim")
    for cls in (QuaternionSkipConnectAdd, QuaternionSkipConnectConcat):
        m = cls(init="glorot-uniform", **kw)
        assert not hasattr(m, "phm_dim") and m._n == 4
        import pickle
        m2 = pickle.loads(pickle.dumps(m))                      # scripts save whole modules (train_hiv.py:344)
        assert not hasattr(m2, "phm_dim") and m2._n == 4


def test_load_state_dict_accepts_the_other_layouts():
    """model.load_state_dict takes quaternion-named and legacy-named state dicts directly."""
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc_gnn_b200 import legacy
    fx = load_family("quaternion_zinc_sum_mlp")
    q = product_class(fx)(**fx["cfg"])
    q.load_state_dict(fx["state"], strict=True)
    want = legacy.quaternion_to_phm_state_dict(fx["state"])
    assert all(torch.equal(v, want[k]) for k, v in q.state_dict().items())
    kw = dict(phm_dim=2, atom_encoded_dim=8, mp_layers=[8, 8], dropout_mpnn=[0.0, 0.0], downstream_layers=[8, 4], mlp=True)
    a, b = PHMSkipConnectAdd(**kw), PHMSkipConnectAdd(**kw)
    old = legacy.to_legacy_phm_state_dict(a.state_dict())
    assert legacy.is_legacy_phm_state_dict(old) and not legacy.is_legacy_phm_state_dict(a.state_dict())
    b.load_state_dict(old, strict=True)
    assert all(torch.equal(v, a.state_dict()[k]) for k, v in b.state_dict().items())


def test_q_batch_norm_is_refused():
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd
    with pytest.raises(NotImplementedError):
        QuaternionSkipConnectAdd(atom_encoded_dim=16, mp_layers=[16], dropout_mpnn=[0.0], norm_mp="q-batch-norm")


def test_legacy_phm_linear_relation():
    """PHMLinear_Old vectors recorded from the reference: the oracle's restatement of the legacy layer reproduces them,
    and the product's converted parameters reproduce them through the CURRENT layer's oracle (values and gradients)."""
    from phc_gnn_b200 import legacy
    fxs = load_family("legacy_phmlinear")
    for key, fx in fxs.items():
        n = fx["n"]
        old = leaves(fx["state"])
        x = fx["x"].clone().requires_grad_(True)
        y = O.legacy_phm_linear(x, old, "", n)
        torch.testing.assert_close(y, fx["y"], rtol=RTOL, atol=ATOL)
        y.backward(fx["gy"])
        torch.testing.assert_close(x.grad, fx["gx"], rtol=RTOL, atol=ATOL)
        for k, g in fx["grads"].items():
            torch.testing.assert_close(old[k].grad, g, rtol=RTOL, atol=1e-4, msg=lambda m: f"{key} {k}: {m}")
        new_state = legacy.convert_legacy_phm_state_dict(fx["state"])
        assert list(new_state) == ["phm_rule", "W", "b"]
        assert new_state["W"].shape == (n, fx["in_per"], fx["out_per"]) and new_state["b"].shape == (n * fx["out_per"],)
        new = {"l." + k: v.clone().requires_grad_(True) for k, v in new_state.items()}
        x2 = fx["x"].clone().requires_grad_(True)
        y2 = O.phm_linear(x2, new, "l")
        torch.testing.assert_close(y2, fx["y"], rtol=RTOL, atol=ATOL)
        y2.backward(fx["gy"])
        torch.testing.assert_close(x2.grad, fx["gx"], rtol=RTOL, atol=ATOL)
        grads_old_layout = legacy.to_legacy_phm_state_dict({k[2:]: v.grad for k, v in new.items()})
        for k, g in fx["grads"].items():
            torch.testing.assert_close(grads_old_layout[k], g, rtol=RTOL, atol=1e-4, msg=lambda m: f"{key} {k}: {m}")
        back = legacy.to_legacy_phm_state_dict(new_state)
        assert list(back) == list(fx["state"]) and all(torch.equal(back[k], fx["state"][k]) for k in back)


@pytest.mark.parametrize("which", sorted(SHIPPED))
def test_shipped_checkpoint_import(which):
    """The reference's shipped ``model.pt`` (legacy layout, pre-rename module paths): config recovered, strict load into the
    current model class, known parameter count, and a finite eval forward through the oracle.  Needs /root/reference
    (dev container); the GPU box runs the layout tests above on the committed fixtures instead."""
    path, n_params = SHIPPED[which]
    if not os.path.exists(path):
        pytest.skip("reference checkpoints are only present in the dev container")
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc_gnn_b200 import legacy
    from phc_gnn_b200.synthetic import make_batch, workloads
    kind, cfg, sd = legacy.read_legacy_checkpoint(path)
    assert kind == "PHMSkipConnectAdd"
    assert not legacy.is_legacy_phm_state_dict(sd)
    model = PHMSkipConnectAdd(**cfg)
    model.load_state_dict(sd, strict=True)
    assert model.get_number_of_params_() == n_params
    wl = workloads(cfg["phm_dim"])[which]
    assert cfg["atom_input_dims"] == wl.model["atom_input_dims"] and cfg["bond_input_dims"] == wl.model["bond_input_dims"]
    batch = make_batch(wl, seed=1, batch_graphs=8)
    p = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        out = O.model_forward(p, cfg, batch, training=False)
    assert out.shape == (8, cfg["target_dim"]) and torch.isfinite(out).all()


def phm_option_cases():
    return sorted(f[:-3] for f in os.listdir(FAMILY) if f.startswith("phm_") and f.endswith(".pt"))


@pytest.mark.parametrize("name", phm_option_cases())
def test_oracle_matches_reference_on_constructor_options(name):
    """naive encoders, add_self_loops=False, bias=False, frozen rule, fixed softmax beta — reference-recorded."""
    fx = load_family(name)
    p = leaves(fx["state"])
    for k, v in p.items():                                  # frozen parameters (learn_phm / learn_beta False) have no gradient
        if v.requires_grad and k not in fx["grads"]:
            v.requires_grad_(False)
    data, cfg = fx["batch"], fx["cfg"]
    forward = O.concat_model_forward if fx["model"] == "phm_concat" else O.model_forward     # concat: phm_dim = 1 only (D2)
    logits = forward(p, cfg, data, training=True)
    torch.testing.assert_close(logits, fx["logits_train"], rtol=RTOL, atol=ATOL)
    reg = O.weight_regularization(p, 2)
    torch.testing.assert_close(reg, fx["reg"], rtol=RTOL, atol=ATOL)
    loss = O.task_loss(logits, data.y, fx["loss_kind"]) + fx["reg_scale"] * reg
    torch.testing.assert_close(loss, fx["loss"], rtol=RTOL, atol=ATOL)
    loss.backward()
    for k, g in fx["grads"].items():
        assert p[k].grad is not None, k
        torch.testing.assert_close(p[k].grad, g, rtol=5e-4, atol=5e-5, msg=lambda m: f"{k}: {m}")
    with torch.no_grad():
        ev = forward(p, cfg, data, training=False)
    torch.testing.assert_close(ev, fx["logits_eval"], rtol=RTOL, atol=ATOL)
    assert sum(v.numel() for v in p.values() if v.requires_grad) == fx["n_params"]


def phm_product_class(fx):
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd, PHMSkipConnectConcat
    return PHMSkipConnectConcat if fx["model"] == "phm_concat" else PHMSkipConnectAdd


@pytest.mark.parametrize("name", phm_option_cases())
def test_product_model_takes_option_state_dicts(name):
    fx = load_family(name)
    m = phm_product_class(fx)(**fx["cfg"])
    m.load_state_dict(fx["state"], strict=True)
    assert m.get_number_of_params_() == fx["n_params"]


def test_quaternion_initialisers_against_reference_known_answers():
    """Same seeds -> the reference's 'quaternion' initialisation to rounding (same RNG streams in the same order);
    'orthogonal': same invariants (real part of the column / row Gram matrix = identity, same element spread)."""
    import numpy as np
    from phc_gnn_b200.quaternion import quaternion_init, quaternion_orthogonal_init
    fx = torch.load(os.path.join(FAMILY, "inits_quaternion.pt"), weights_only=False)
    for fin, fout in ((6, 5), (4, 9)):
        np.random.seed(21)
        torch.manual_seed(21)
        got = quaternion_init(fin, fout)                                       # [4, in, out]
        torch.testing.assert_close(got.permute(0, 2, 1), fx[f"quaternion_{fin}_{fout}"], rtol=1e-6, atol=1e-7)
        torch.manual_seed(22)
        w = quaternion_orthogonal_init(fin, fout).double().permute(0, 2, 1)   # reference orientation [4, out, in]
        gram = torch.einsum("cok,coj->kj", w, w) if fout >= fin else torch.einsum("cok,cpk->op", w, w)
        want = fx[f"orthogonal_{fin}_{fout}_gram"]
        torch.testing.assert_close(want, torch.eye(want.size(0), dtype=torch.float64), rtol=0, atol=1e-5)   # the reference's property
        torch.testing.assert_close(gram, torch.eye(gram.size(0), dtype=torch.float64), rtol=0, atol=1e-5)   # ... and ours
        assert abs(float(w.std()) - float(fx[f"orthogonal_{fin}_{fout}_std"])) < 0.02


def test_model_initialisation_reproduces_the_reference_under_the_same_seeds():
    """A freshly constructed PHM model draws the reference's initial weights (same RNG streams in the same order) — except
    element ``b[out/n]`` of every PHMLinear bias, which the reference leaves uninitialised (0.2 here)."""
    import numpy as np
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    fx = torch.load(os.path.join(FAMILY, "inits_phm_model.pt"), weights_only=False)
    np.random.seed(fx["seed"])
    torch.manual_seed(fx["seed"])
    m = PHMSkipConnectAdd(**fx["kw"])
    got = m.state_dict()
    assert set(got) == set(fx["state"])                      # registration order differs, names do not
    for k, want in fx["state"].items():
        g = got[k]
        if k.endswith(".b"):
            p = g.numel() // fx["kw"]["phm_dim"]
            keep = torch.ones_like(g, dtype=torch.bool)
            keep[p] = False
            assert torch.equal(g[keep], want[keep]), k
            assert float(g[p]) == pytest.approx(0.2)
        else:
            assert torch.equal(g, want), k
