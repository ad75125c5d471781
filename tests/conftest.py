import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_collection_modifyitems(config, items):
    """-m gpu tests skip (rather than fail with "no NVIDIA driver") on a box without a CUDA device."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def golden_cases():
    return sorted(f[:-3] for f in os.listdir(golden_dir()) if f.endswith(".pt") and f != "ops.pt")


@pytest.fixture(scope="session")
def ops_golden():
    import torch
    return torch.load(os.path.join(golden_dir(), "ops.pt"), weights_only=False)


def load_golden(name):
    import torch
    from phc_gnn_b200.synthetic import GraphBatch
    fx = torch.load(os.path.join(golden_dir(), name + ".pt"), weights_only=False)
    d = fx["data"]
    fx["batch"] = GraphBatch(d["x"], d["edge_index"], d["edge_attr"], d["batch"], d["y"], d["num_graphs"])
    return fx
