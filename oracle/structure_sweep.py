"""Constructor-option sweep shared by oracle/make_structure_fixture.py (runs it on the UNMODIFIED reference, dev container
only) and tests/test_structure_sweep.py (runs it on the product): for every configuration the model is only CONSTRUCTED and
summarised — state-dict keys and shapes, the set of trainable parameters, the parameter count.  TEST INFRASTRUCTURE."""
import hashlib

import numpy as np
import torch


def configurations():
    """[(tag, family, kwargs)] — family in {"add", "cat", "qadd", "qcat"}."""
    base = dict(atom_encoded_dim=24, mp_layers=[24, 24], dropout_mpnn=[0.0, 0.0], downstream_layers=[24, 12], target_dim=3)
    out = []
    for n in (1, 2, 3, 4, 6):
        for mlp in (False, True):
            for aggr in ("add", "mean", "max", "softmax", "pna"):
                for variant in range(6):
                    kw = dict(base, phm_dim=n, mlp=mlp, msg_aggr=aggr)
                    if aggr == "softmax":
                        kw.update(initial_beta=1.0, learn_beta=bool(variant % 2))
                    if aggr == "pna":
                        kw.update(aggregators=["mean", "max"], scalers=["identity", "amplification"], deg=torch.tensor([0, 3, 5, 2]),
                                  post_layers=1 + variant % 2)
                    if variant == 1:
                        kw.update(naive_encoder=True)
                    if variant == 2:
                        kw.update(bias=False, norm_mp=None, norm_dn=None)
                    if variant == 3:
                        kw.update(learn_phm=False, pooling="globalsum")
                    if variant == 4:
                        kw.update(atom_input_dims=5, bond_input_dims=2, sc_type="last")
                    if variant == 5:
                        kw.update(real_trafo="sum", atom_input_dims=[7], bond_input_dims=[3])
                    out.append((f"add|n{n}|mlp{int(mlp)}|{aggr}|v{variant}", "add", kw))
    for mlp in (False, True):          # the concat model is only constructible AND runnable at phm_dim = 1 in the reference
        for variant in range(3):
            kw = dict(base, phm_dim=1, mlp=mlp, msg_aggr="add", mp_layers=[24, 12])
            if variant == 1:
                kw.update(sc_type="last", pooling="globalsum")
            if variant == 2:
                kw.update(naive_encoder=True, atom_input_dims=4, bond_input_dims=1)
            out.append((f"cat|n1|mlp{int(mlp)}|v{variant}", "cat", kw))
    for fam in ("qadd", "qcat"):
        for mlp in (False, True):
            for aggr in ("add", "softmax", "max"):
                for variant in (0, 1, 3):
                    kw = dict(base, mlp=mlp, msg_aggr=aggr, init="glorot-uniform")
                    if aggr == "softmax":
                        kw.update(initial_beta=1.0, learn_beta=True)
                    if variant == 1:
                        kw.update(naive_encoder=True)
                    if variant == 3:
                        kw.update(atom_input_dims=5, bond_input_dims=2)
                    if fam == "qcat":
                        kw.update(mp_layers=[24, 12])
                    out.append((f"{fam}|mlp{int(mlp)}|{aggr}|v{variant}", fam, kw))
    return out


def summarise(state_shapes: dict, trainable, n_params: int) -> dict:
    """state_shapes: {key: shape list} under the REFERENCE's parameter names."""
    h = hashlib.sha1()
    for k in sorted(state_shapes):
        h.update(f"{k}:{tuple(state_shapes[k])};".encode())
    t = hashlib.sha1()
    for k in sorted(trainable):
        t.update((k + ";").encode())
    return dict(keys=len(state_shapes), state=h.hexdigest(), trainable=t.hexdigest(), n_params=int(n_params))


def build(classes: dict, family: str, kw: dict):
    np.random.seed(0)
    torch.manual_seed(0)
    return classes[family](**kw)
