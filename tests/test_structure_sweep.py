"""Drop-in structure of the module API over a sweep of 342 constructor configurations: for each, the product model is
constructed (CPU, no kernel involved) and its state-dict keys and shapes, trainable-parameter set and parameter count are
compared with what the UNMODIFIED reference builds (tests/golden/family/structure_sweep.json, oracle/make_structure_fixture.py).
Quaternion models are compared under the reference's parameter names (legacy.phm_to_quaternion_state_dict)."""
import json
import os
import sys

import pytest
import torch

from conftest import ROOT, golden_dir

sys.path.insert(0, os.path.join(ROOT, "oracle"))
from structure_sweep import build, configurations, summarise  # noqa: E402


@pytest.fixture(scope="module")
def reference_structure():
    with open(os.path.join(golden_dir(), "family", "structure_sweep.json")) as fh:
        return json.load(fh)


def _classes():
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd, PHMSkipConnectConcat
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    return dict(add=PHMSkipConnectAdd, cat=PHMSkipConnectConcat, qadd=QuaternionSkipConnectAdd, qcat=QuaternionSkipConnectConcat)


@pytest.mark.parametrize("family", ["add", "cat", "qadd", "qcat"])
def test_structure_matches_reference(family, reference_structure):
    from phc_gnn_b200 import legacy
    classes = _classes()
    checked = 0
    for tag, fam, kw in configurations():
        if fam != family:
            continue
        m = build(classes, fam, kw)
        shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
        trainable = [k for k, p in m.named_parameters() if p.requires_grad]
        if fam.startswith("q"):                      # compare under the reference's W_r.. / b_r.. / qlinear / bn.bn.r names
            dummy = {k: torch.empty(s) for k, s in shapes.items()}
            shapes = {k: list(v.shape) for k, v in legacy.phm_to_quaternion_state_dict(dummy).items()}
            trainable = list(legacy.phm_to_quaternion_state_dict({k: dummy[k] for k in trainable} |
                                                                 {k: dummy[k] for k in dummy if k.endswith(".W")}).keys())
            trainable = [k for k in trainable if k in shapes and "running" not in k and "tracked" not in k]
        got = summarise(shapes, trainable, m.get_number_of_params_())
        assert got == reference_structure[tag], f"{tag}: {got} != {reference_structure[tag]}"
        checked += 1
    assert checked == sum(1 for _, fam, _ in configurations() if fam == family) and checked > 0
