#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun).
#   bash baseline/install_ref.sh
# 1. the contract's offline pip install (the reference's setup.py lists packages 'phc', 'phc.hypercomplex',
#    'phc.quaternion', 'benchmarks' only);
# 2. setup.py omits the sub-packages phc/hypercomplex/undirectional and phc/quaternion/undirectional (the model and
#    message-passing modules): they are copied verbatim from the same source tree, nothing is edited.
# torch_scatter / torch_geometric / ogb cannot be installed offline: bench.py --impl reference puts
# oracle/refshim (pure-torch restatements of the pinned versions' documented semantics) on the path for them.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$REF"
for sub in hypercomplex/undirectional quaternion/undirectional; do
  mkdir -p "$HERE/_ref/phc/$sub"
  cp "$REF/phc/$sub/"*.py "$HERE/_ref/phc/$sub/"
done
rm -rf "$HERE/_ref/benchmarks"      # scripts, datasets and checkpoints: not on the timed path
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
python - <<PY
import hashlib, os, sys
ref, dst = "$REF/phc", "$HERE/_ref/phc"
bad = []
for root, _, files in os.walk(dst):
    for f in files:
        if f.endswith(".py"):
            a = os.path.join(root, f); b = os.path.join(ref, os.path.relpath(a, dst))
            if hashlib.sha256(open(a, "rb").read()).digest() != hashlib.sha256(open(b, "rb").read()).digest():
                bad.append(a)
assert not bad, f"installed files differ from the reference: {bad}"
print("baseline/_ref: every installed module is byte-identical to", ref)
PY
