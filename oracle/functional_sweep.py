"""Random-configuration sweep shared by oracle/make_functional_fixture.py (runs the UNMODIFIED reference, dev container only)
and tests/test_functional_sweep.py (runs the CPU oracle): model options, inputs and weights are all derived from seeds, so
the fixture only stores the reference's outputs (logits, loss, per-parameter gradient summaries).  TEST INFRASTRUCTURE."""
import random
import zlib

import torch

from phc_gnn_b200.synthetic import make_batch, tiny, workloads

ACTS = ["relu", "lrelu", "elu", "selu", "swish"]


def configurations(seed: int = 0, count: int = 120):
    """[(tag, workload, kwargs, batch_seed)], deterministic in (seed, count)."""
    rng = random.Random(seed)
    out = []
    for it in range(count):
        n = rng.choice([1, 2, 3, 4, 5])
        wname = rng.choice(["hiv", "zinc", "mnist", "pcba", "ppa"])
        width = n * rng.choice([2, 3, 4])
        layers = rng.choice([1, 2, 3])
        head = [n * rng.choice([2, 3]) for _ in range(rng.choice([1, 2]))]
        wl = tiny(workloads(n)[wname], width, layers, rng.choice([3, 5]), 3, 7, und_edges=8 if wname == "ppa" else None, head=head)
        if wname == "mnist":
            wl.extra["k"] = 2
        kw = dict(wl.model)
        aggr = rng.choice(["sum", "mean", "max", "min", "softmax", "pna"])
        kw.update(msg_aggr=aggr, mlp=rng.random() < 0.5, activation=rng.choice(ACTS), msg_encoder=rng.choice(ACTS + ["identity"]),
                  pooling=rng.choice(["globalsum", "softattention"]), sc_type=rng.choice(["first", "last"]),
                  naive_encoder=rng.random() < 0.25, bias=rng.random() < 0.8, add_self_loops=rng.random() < 0.75,
                  learn_phm=rng.random() < 0.8, real_trafo=rng.choice(["linear", "linear", "linear", "sum", "norm"]),
                  same_dropout=rng.random() < 0.3)
        if rng.random() < 0.25:
            kw.update(norm_mp=None)
        if rng.random() < 0.25:
            kw.update(norm_dn=None)
        if aggr == "softmax":
            kw.update(initial_beta=rng.choice([0.5, 1.0, 2.0]), learn_beta=rng.random() < 0.7)
        if aggr == "pna":
            kw.update(aggregators=rng.sample(["mean", "min", "max", "std", "sum", "var"], 3),
                      scalers=rng.sample(["identity", "amplification", "attenuation", "linear", "inverse_linear"], 2),
                      deg=[0, 4, 9, 6, 2, 1], post_layers=rng.choice([1, 2]))
        kw["dropout_mpnn"] = [0.0] * layers
        kw["dropout_dn"] = [0.0] * len(head)
        if wname == "pcba":
            kw["target_dim"] = 4
        if wname == "ppa":
            kw["target_dim"] = 5
        if kw["real_trafo"] != "linear" and kw["pooling"] == "softattention" and n > 1:
            kw["pooling"] = "globalsum"            # the reference raises for that combination
        tag = (f"{it}:{wname} n{n} w{width} L{layers} {aggr} mlp{int(kw['mlp'])} {kw['activation']}/{kw['msg_encoder']} {kw['pooling']} "
               f"sc={kw['sc_type']} naive{int(kw['naive_encoder'])} bias{int(kw['bias'])} loops{int(kw['add_self_loops'])} "
               f"lphm{int(kw['learn_phm'])} rt={kw['real_trafo']} nm={kw.get('norm_mp')} nd={kw.get('norm_dn')}")
        out.append((tag, wl, kw, 10_000 + it))
    return out


def model_kwargs(kw: dict) -> dict:
    k = dict(kw)
    if "deg" in k:
        k["deg"] = torch.tensor(k["deg"])
    return k


def batch_for(wl, kw, batch_seed):
    data = make_batch(wl, seed=batch_seed)
    if wl.name == "ppa":
        data.y = data.y % 5
    if wl.name == "pcba":
        data.y = data.y[:, :4].contiguous()
    return data


def fill_by_name(named_tensors, seed: int) -> None:
    """Overwrite parameters / buffers in place with values that depend only on (seed, tensor name, shape) — independent of
    the order in which a module tree registers them."""
    with torch.no_grad():
        for name, t in named_tensors:
            if not t.is_floating_point():
                continue
            g = torch.Generator().manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 31))
            leaf = name.rsplit(".", 1)[-1]
            r = torch.randn(t.shape, generator=g)
            if leaf == "running_var":
                t.copy_(1.0 + 0.3 * r.abs())
            elif leaf == "running_mean":
                t.copy_(0.1 * r)
            elif leaf == "beta":
                pass
            elif leaf == "phm_rule":
                t.add_(0.15 * r)
            elif ".bn." in name and leaf == "weight":
                t.copy_(1.0 + 0.2 * r)
            elif leaf in ("b", "bias"):
                t.copy_(0.1 * r)
            else:
                t.copy_(0.35 * r)


def loss_fn(logits, y, kind, target_dim, task_loss):
    if logits.size(-1) != target_dim:              # a non-linear real_trafo leaves n * target_dim outputs (D6)
        return logits.square().mean()
    return task_loss(logits, y, kind)


def grad_summary(named_grads) -> list:
    """[[L2 norm, sum], ...] of the gradients in the order of the sorted parameter names (names are not stored)."""
    return [[float(g.double().norm()), float(g.double().sum())] for _, g in sorted(named_grads, key=lambda kv: kv[0]) if g is not None]


def quaternion_configurations(seed: int = 1, count: int = 48):
    """[(tag, concat, workload, kwargs, batch_seed)] for QuaternionSkipConnectAdd / Concat (norm_mp / norm_dn stay on: the
    reference's quaternion models raise without them)."""
    rng = random.Random(seed)
    out = []
    for it in range(count):
        wname = rng.choice(["hiv", "zinc", "mnist", "pcba", "ppa"])
        concat = rng.random() < 0.4
        width = 4 * rng.choice([1, 2, 3])
        layers = rng.choice([1, 2, 3])
        head = [4 * rng.choice([1, 2]) for _ in range(rng.choice([1, 2]))]
        wl = tiny(workloads(4)[wname], width, layers, rng.choice([3, 5]), 3, 7, und_edges=8 if wname == "ppa" else None, head=head)
        if wname == "mnist":
            wl.extra["k"] = 2
        kw = {k: v for k, v in wl.model.items() if k not in ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type")}
        aggr = rng.choice(["sum", "mean", "max", "min", "softmax"])
        kw.update(init="glorot-uniform", msg_aggr=aggr, mlp=rng.random() < 0.5, activation=rng.choice(ACTS),
                  msg_encoder=rng.choice(["identity", "relu", "swish"]), pooling=rng.choice(["globalsum", "softattention"]),
                  naive_encoder=rng.random() < 0.3, bias=rng.random() < 0.8)
        if concat:
            kw["mp_layers"] = [4 * rng.choice([1, 2, 3]) for _ in range(layers)]
        if aggr == "softmax":
            kw.update(initial_beta=rng.choice([0.5, 1.5]), learn_beta=rng.random() < 0.7)
        kw["dropout_mpnn"] = [0.0] * layers
        kw["dropout_dn"] = [0.0] * len(head)
        if wname == "pcba":
            kw["target_dim"] = 4
        if wname == "ppa":
            kw["target_dim"] = 5
        tag = (f"q{it}:{'cat' if concat else 'add'} {wname} w{width} L{layers} {aggr} mlp{int(kw['mlp'])} {kw['activation']}/"
               f"{kw['msg_encoder']} {kw['pooling']} naive{int(kw['naive_encoder'])} bias{int(kw['bias'])}")
        out.append((tag, concat, wl, kw, 20_000 + it))
    return out
