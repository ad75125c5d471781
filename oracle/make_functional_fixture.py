"""tests/golden/family/functional_sweep.json: 120 random model configurations run by the UNMODIFIED reference (dev container
only — needs /root/reference): train-mode logits, loss, per-parameter gradient summaries, eval-mode logits.  Inputs and
weights are regenerated from seeds by the test (oracle/functional_sweep.py), so only outputs are stored.

    python oracle/make_functional_fixture.py
"""
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "refshim"))
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import types  # noqa: E402
_ref_phc = types.ModuleType("phc")
_ref_phc.__path__ = ["/root/reference/phc"]
sys.modules["phc"] = _ref_phc

import numpy as np  # noqa: E402
import torch  # noqa: E402

from functional_sweep import (batch_for, configurations, fill_by_name, grad_summary, loss_fn, model_kwargs,  # noqa: E402
                              quaternion_configurations)
from make_golden import ref_loss  # noqa: E402


def _round(obj):
    """9 significant digits: fp32 outputs carry no more."""
    if isinstance(obj, float):
        return float(f"{obj:.9g}")
    if isinstance(obj, list):
        return [_round(v) for v in obj]
    if isinstance(obj, dict):
        return {k: _round(v) for k, v in obj.items()}
    return obj


def main():
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc.hypercomplex.regularization import phm_weight_regularization
    import phc.hypercomplex.undirectional.models as _ref_models
    assert _ref_models.__file__.startswith("/root/reference/"), f"not the reference: {_ref_models.__file__}"
    out = {}
    for it, (tag, wl, kw, bseed) in enumerate(configurations()):
        torch.manual_seed(it)
        np.random.seed(it)
        model = PHMSkipConnectAdd(**model_kwargs(kw))
        fill_by_name(list(model.named_parameters()) + list(model.named_buffers()), 77 + it)
        data = batch_for(wl, kw, bseed)
        model.train()
        logits = model(data)
        reg = phm_weight_regularization(model, p=2)
        loss = loss_fn(logits, data.y, wl.loss, kw["target_dim"], ref_loss) + 0.01 * reg
        loss.backward()
        grads = grad_summary((k, p.grad) for k, p in model.named_parameters())
        model.eval()
        with torch.no_grad():
            ev = model(data)
        out[tag] = _round(dict(logits=logits.detach().double().tolist(), loss=float(loss), reg=float(reg), grads=grads,
                               logits_eval=ev.double().tolist()))
    path = os.path.join(ROOT, "tests", "golden", "family", "functional_sweep.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(len(out), "configurations ->", path, os.path.getsize(path) // 1024, "KiB")
    quaternion_main()


def quaternion_main():
    """Same for the quaternion family (48 configurations of QuaternionSkipConnectAdd / Concat) -> functional_sweep_quaternion.json;
    weights are filled under the reference's own parameter names."""
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    from phc.quaternion.regularization import quaternion_weight_regularization
    out = {}
    for it, (tag, concat, wl, kw, bseed) in enumerate(quaternion_configurations()):
        torch.manual_seed(it)
        np.random.seed(it)
        model = (QuaternionSkipConnectConcat if concat else QuaternionSkipConnectAdd)(**kw)
        fill_by_name(list(model.named_parameters()) + list(model.named_buffers()), 177 + it)
        data = batch_for(wl, kw, bseed)
        model.train()
        logits = model(data)
        reg = quaternion_weight_regularization(model, device="cpu", p=2)
        loss = loss_fn(logits, data.y, wl.loss, kw["target_dim"], ref_loss) + 0.01 * reg
        loss.backward()
        grads = grad_summary((k, p.grad) for k, p in model.named_parameters())
        model.eval()
        with torch.no_grad():
            ev = model(data)
        out[tag] = _round(dict(logits=logits.detach().double().tolist(), loss=float(loss), reg=float(reg), grads=grads,
                               logits_eval=ev.double().tolist()))
    path = os.path.join(ROOT, "tests", "golden", "family", "functional_sweep_quaternion.json")
    with open(path, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print(len(out), "quaternion configurations ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
