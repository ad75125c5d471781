"""The CUDA path against the reference on the 120 random configurations of oracle/functional_sweep.py (fixture:
tests/golden/family/functional_sweep.json).  Opt-in (``PHC_GPU_SWEEP=1``): it was written after the round's GPU budget was
spent and has NOT run on a GPU yet, so it is skipped by default rather than risk an unverified red in the suite;
``PHC_GPU_SWEEP=1 python -m pytest tests/test_zz_sweep_gpu.py -m gpu`` is the first thing to do with the next GPU minutes,
after which the switch goes away."""
import json
import os
import sys

import pytest
import torch

from conftest import ROOT, golden_dir
from oracle import phc_oracle as O

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from functional_sweep import batch_for, configurations, fill_by_name, grad_summary, loss_fn, model_kwargs  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PHC_GPU_SWEEP") != "1", reason="not yet verified on a GPU: set PHC_GPU_SWEEP=1 to run")]
DEV = "cuda:0"
CONFIGS = configurations()


@pytest.fixture(scope="module")
def reference_outputs():
    with open(os.path.join(golden_dir(), "family", "functional_sweep.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("index", range(len(CONFIGS)))
def test_cuda_path_matches_reference_on_random_configuration(index, reference_outputs, monkeypatch):
    monkeypatch.setenv("PHC_PRECISION", "fp32")
    from phc.hypercomplex.regularization import phm_weight_regularization
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    tag, wl, kw, bseed = CONFIGS[index]
    want = reference_outputs[tag]
    torch.manual_seed(index)
    model = PHMSkipConnectAdd(**model_kwargs(kw))
    fill_by_name(list(model.named_parameters()) + list(model.named_buffers()), 77 + index)
    model = model.to(DEV)
    data = batch_for(wl, kw, bseed).to(DEV)
    model.train()
    logits = model(data)
    ref_logits = torch.tensor(want["logits"])
    assert logits.shape == ref_logits.shape, tag
    assert float((logits.detach().cpu().double() - ref_logits).abs().max()) <= 2e-4 * max(1.0, float(ref_logits.abs().max())), tag
    reg = phm_weight_regularization(model, p=2)
    loss = loss_fn(logits, data.y, wl.loss, kw["target_dim"], O.task_loss) + 0.01 * reg
    assert abs(float(reg) - want["reg"]) <= 1e-4 * max(1.0, abs(want["reg"])), tag
    assert abs(float(loss) - want["loss"]) <= 2e-4 * max(1.0, abs(want["loss"])), tag
    loss.backward()
    torch.cuda.synchronize()
    got = grad_summary((k, p.grad.detach().cpu()) for k, p in model.named_parameters() if p.requires_grad)
    assert len(got) == len(want["grads"]), tag
    gscale = max(1e-3, max(w[0] for w in want["grads"]))
    for (gn, gs), (wn, ws) in zip(got, want["grads"]):
        assert abs(gn - wn) <= 2e-3 * max(wn, 1e-1 * gscale), f"{tag}: gradient norm {gn} vs {wn}"
        assert abs(gs - ws) <= 2e-3 * max(abs(ws), wn, 1e-1 * gscale), f"{tag}: gradient sum {gs} vs {ws}"
    model.eval()
    with torch.no_grad():
        ev = model(data).cpu().double()
    ref_ev = torch.tensor(want["logits_eval"])
    assert float((ev - ref_ev).abs().max()) <= 2e-4 * max(1.0, float(ref_ev.abs().max())), tag
