"""The CPU oracle against the UNMODIFIED reference on 120 random model configurations (every aggregator incl. PNA, both conv
types, n = 1..5, encoders, bias / norm / self-loop / frozen-rule / real_trafo / skip variants; oracle/functional_sweep.py).
The fixture (oracle/make_functional_fixture.py) stores only the reference's outputs; inputs and weights are regenerated from
seeds — weights through the PRODUCT's module tree, whose parameter names and shapes equal the reference's
(tests/test_structure_sweep.py), filled by name."""
import json
import os
import sys

import pytest
import torch

from conftest import ROOT, golden_dir
from oracle import phc_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from functional_sweep import batch_for, configurations, fill_by_name, grad_summary, loss_fn, model_kwargs  # noqa: E402

CONFIGS = configurations()


@pytest.fixture(scope="module")
def reference_outputs():
    with open(os.path.join(golden_dir(), "family", "functional_sweep.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("chunk", range(6))
def test_oracle_matches_reference_on_random_configurations(chunk, reference_outputs):
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    for it, (tag, wl, kw, bseed) in enumerate(CONFIGS):
        if it % 6 != chunk:
            continue
        want = reference_outputs[tag]
        torch.manual_seed(it)
        model = PHMSkipConnectAdd(**model_kwargs(kw))                         # construction only (CPU): names, shapes, requires_grad
        fill_by_name(list(model.named_parameters()) + list(model.named_buffers()), 77 + it)
        trainable = {k for k, p in model.named_parameters() if p.requires_grad}
        p = {}
        for k, v in model.state_dict().items():
            v = v.detach().clone()
            if k in trainable:
                v.requires_grad_(True)
            p[k] = v
        data = batch_for(wl, kw, bseed)
        logits = O.model_forward(p, kw | {"deg": model_kwargs(kw).get("deg")} if "deg" in kw else kw, data, training=True)
        ref_logits = torch.tensor(want["logits"])
        scale = max(1.0, float(ref_logits.abs().max()))
        assert logits.shape == ref_logits.shape, tag
        assert float((logits.double() - ref_logits).abs().max()) <= 2e-4 * scale, tag
        reg = O.weight_regularization(p, 2)
        loss = loss_fn(logits, data.y, wl.loss, kw["target_dim"], O.task_loss) + 0.01 * reg
        assert abs(float(reg) - want["reg"]) <= 1e-4 * max(1.0, abs(want["reg"])), tag
        assert abs(float(loss) - want["loss"]) <= 2e-4 * max(1.0, abs(want["loss"])), tag
        loss.backward()
        got = grad_summary((k, p[k].grad) for k in sorted(trainable))
        assert len(got) == len(want["grads"]), f"{tag}: {len(got)} gradients, the reference has {len(want['grads'])}"
        # tolerance per tensor: 2e-3 relative, with a floor of 2e-4 x the model's largest gradient norm — entries that are
        # zero in exact arithmetic (a bias in front of a batch norm) carry fp32 noise of that size on both sides
        gscale = max(1e-3, max(w[0] for w in want["grads"]))
        for (gn, gs), (wn, ws) in zip(got, want["grads"]):
            assert abs(gn - wn) <= 2e-3 * max(wn, 1e-1 * gscale), f"{tag}: gradient norm {gn} vs {wn}"
            assert abs(gs - ws) <= 2e-3 * max(abs(ws), wn, 1e-1 * gscale), f"{tag}: gradient sum {gs} vs {ws}"
        with torch.no_grad():
            ev = O.model_forward({k: v.detach() for k, v in p.items()}, kw | {"deg": model_kwargs(kw).get("deg")} if "deg" in kw else kw,
                                 data, training=False)
        ref_ev = torch.tensor(want["logits_eval"])
        assert float((ev.double() - ref_ev).abs().max()) <= 2e-4 * max(1.0, float(ref_ev.abs().max())), tag


@pytest.fixture(scope="module")
def quaternion_reference_outputs():
    with open(os.path.join(golden_dir(), "family", "functional_sweep_quaternion.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("chunk", range(4))
def test_quaternion_oracle_matches_reference_on_random_configurations(chunk, quaternion_reference_outputs):
    """48 random QuaternionSkipConnectAdd / Concat configurations: the oracle's Hamilton-product restatement against the
    reference's outputs.  The parameter names and shapes come from the product model's ``quaternion_state_dict()`` (the
    reference's layout), filled by name exactly as the generator filled the reference model."""
    from functional_sweep import quaternion_configurations
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    from phc_gnn_b200 import legacy
    for it, (tag, concat, wl, kw, bseed) in enumerate(quaternion_configurations()):
        if it % 4 != chunk:
            continue
        want = quaternion_reference_outputs[tag]
        torch.manual_seed(it)
        model = (QuaternionSkipConnectConcat if concat else QuaternionSkipConnectAdd)(**kw)
        state = {k: v.detach().clone() for k, v in model.quaternion_state_dict().items()}
        fill_by_name(list(state.items()), 177 + it)
        frozen = set(legacy.phm_to_quaternion_state_dict({k: v for k, v in model.state_dict().items()
                                                           if k.endswith(".beta") and not dict(model.named_parameters())[k].requires_grad}))
        pq = {}
        for k, v in state.items():
            if v.is_floating_point() and "running" not in k and k not in frozen:
                v.requires_grad_(True)
            pq[k] = v
        data = batch_for(wl, kw, bseed)
        forward = O.quaternion_concat_model_forward if concat else O.quaternion_model_forward
        logits = forward(pq, kw, data, training=True)
        ref_logits = torch.tensor(want["logits"])
        assert logits.shape == ref_logits.shape, tag
        assert float((logits.detach().double() - ref_logits).abs().max()) <= 2e-4 * max(1.0, float(ref_logits.abs().max())), tag
        reg = O.quaternion_weight_regularization(pq, kw, 2)
        loss = loss_fn(logits, data.y, wl.loss, kw["target_dim"], O.task_loss) + 0.01 * reg
        assert abs(float(reg.detach()) - want["reg"]) <= 1e-4 * max(1.0, abs(want["reg"])), tag
        assert abs(float(loss.detach()) - want["loss"]) <= 2e-4 * max(1.0, abs(want["loss"])), tag
        loss.backward()
        got = grad_summary((k, v.grad) for k, v in pq.items() if v.requires_grad)
        assert len(got) == len(want["grads"]), f"{tag}: {len(got)} gradients, the reference has {len(want['grads'])}"
        gscale = max(1e-3, max(w[0] for w in want["grads"]))
        for (gn, gs), (wn, ws) in zip(got, want["grads"]):
            assert abs(gn - wn) <= 2e-3 * max(wn, 1e-1 * gscale), f"{tag}: gradient norm {gn} vs {wn}"
            assert abs(gs - ws) <= 2e-3 * max(abs(ws), wn, 1e-1 * gscale), f"{tag}: gradient sum {gs} vs {ws}"
        with torch.no_grad():
            ev = forward({k: v.detach() for k, v in pq.items()}, kw, data, training=False)
        ref_ev = torch.tensor(want["logits_eval"])
        assert float((ev.double() - ref_ev).abs().max()) <= 2e-4 * max(1.0, float(ref_ev.abs().max())), tag
