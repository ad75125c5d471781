"""tests/golden/family/structure_sweep.json: the constructor-option sweep of oracle/structure_sweep.py run on the UNMODIFIED
reference (dev container only — needs /root/reference).  TEST INFRASTRUCTURE.

    python oracle/make_structure_fixture.py
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

from structure_sweep import build, configurations, summarise  # noqa: E402


def main():
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd, PHMSkipConnectConcat
    from phc.quaternion.undirectional.models import QuaternionSkipConnectAdd, QuaternionSkipConnectConcat
    import phc.quaternion.undirectional.models as _ref_models
    assert _ref_models.__file__.startswith("/root/reference/"), f"not the reference: {_ref_models.__file__}"
    classes = dict(add=PHMSkipConnectAdd, cat=PHMSkipConnectConcat, qadd=QuaternionSkipConnectAdd, qcat=QuaternionSkipConnectConcat)
    out = {}
    for tag, fam, kw in configurations():
        m = build(classes, fam, kw)
        out[tag] = summarise({k: list(v.shape) for k, v in m.state_dict().items()},
                             [k for k, p in m.named_parameters() if p.requires_grad], m.get_number_of_params_())
    path = os.path.join(ROOT, "tests", "golden", "family", "structure_sweep.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print(len(out), "configurations ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
