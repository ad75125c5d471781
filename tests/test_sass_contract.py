"""Static check of the built library's SASS (no GPU needed): the PHMLinear kernels issue 5th-generation tensor-core instructions
fed by TMA, nothing falls back to legacy mma.sync, no kernel uses a floating-point atomic (determinism), and the tensor-memory-operand
kernels carry no lane-serialisation loop around their MMA / TMA issue (`elect.sync`: DESIGN.md section 4, late round 2)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    from phc_gnn_b200 import _lib
    lib = _lib.build()
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, timeout=600).stdout
    per, cur = {}, None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur[m.group(1) + m.group(2)] += 1
    assert per, "no SASS found in the library (was it built for sm_100a?)"
    return per


def _kernels(sass, needle):
    return {k: v for k, v in sass.items() if needle in k}


def test_tensor_core_kernels_use_tcgen05_and_tma(sass):
    mix, dh = _kernels(sass, "phm_tc_mix_v3_kernel"), _kernels(sass, "phm_tc_dh_v2_kernel")
    assert mix and dh
    for name, c in list(mix.items()) + list(dh.items()):
        assert c["UTCHMMA"] > 0, f"{name}: no tcgen05.mma"
        assert c["STTM"] > 0 and c["LDTM"] > 0, f"{name}: operands / accumulators do not go through tensor memory"
        assert c["UTMALDG"] > 0, f"{name}: no TMA tensor load"
        assert c["BRA.U.ANY"] == 0, f"{name}: lane-serialisation loop around the issue (use elect.sync, not lane == 0)"
    assert all(c["FFMA2"] > 0 for c in mix.values()), "rule mixing lost its packed fp32x2 arithmetic"
    assert sum(c["HMMA"] for c in sass.values()) == 0, "legacy mma.sync in the library"


def test_no_floating_point_atomics_anywhere(sass):
    bad = [k for k, c in sass.items() for op, n in c.items()
           if n and re.match(r"(ATOM|ATOMG|ATOMS|RED|REDG)\.", op) and re.search(r"\.F(16|32|64)", op)]
    assert not bad, f"floating-point atomics (non-deterministic summation order) in: {sorted(set(bad))[:5]}"
