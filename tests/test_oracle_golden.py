"""The CPU oracle (oracle/phc_oracle.py) against vectors recorded from the UNMODIFIED
reference (oracle/make_golden.py).  This is what pins the oracle (prompt section 3)."""
import pytest
import torch

from conftest import golden_cases, load_golden
from oracle import phc_oracle as O

RTOL, ATOL = 1e-4, 2e-5   # fp32 vs fp32, different summation order only


def _params(fx, dtype=torch.float32):
    p = {}
    for k, v in fx["state"].items():
        v = v.clone()
        if v.is_floating_point():
            v = v.to(dtype)
            if "running" not in k:
                v.requires_grad_(True)
        p[k] = v
    return p


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_reference(name):
    fx = load_golden(name)
    p = _params(fx)
    data, cfg = fx["batch"], fx["cfg"]
    logits = O.model_forward(p, cfg, data, training=True)
    torch.testing.assert_close(logits, fx["logits_train"], rtol=RTOL, atol=ATOL)
    reg = O.weight_regularization(p, 2)
    torch.testing.assert_close(reg, fx["reg"], rtol=RTOL, atol=ATOL)
    loss = O.task_loss(logits, data.y, fx["loss_kind"]) + fx["reg_scale"] * reg
    torch.testing.assert_close(loss, fx["loss"], rtol=RTOL, atol=ATOL)
    loss.backward()
    for k, g in fx["grads"].items():
        assert p[k].grad is not None, k
        torch.testing.assert_close(p[k].grad, g, rtol=5e-4, atol=5e-5, msg=lambda m: f"{k}: {m}")
    for k, v in fx["running_after"].items():
        torch.testing.assert_close(p[k], v, rtol=RTOL, atol=ATOL, msg=lambda m: f"{k}: {m}")
    with torch.no_grad():
        ev = O.model_forward(p, cfg, data, training=False)
    torch.testing.assert_close(ev, fx["logits_eval"], rtol=RTOL, atol=ATOL)
    n_params = sum(v.numel() for v in p.values() if v.requires_grad)
    assert n_params == fx["n_params"]


def test_oracle_fp64_agrees(  ):
    fx = load_golden("hiv_n4_softmax_mlp")
    p = _params(fx, torch.float64)
    logits = O.model_forward(p, fx["cfg"], fx["batch"], training=True)
    torch.testing.assert_close(logits.float(), fx["logits_train"], rtol=RTOL, atol=ATOL)


def test_phm_linear_known_answers(ops_golden):
    for key, fx in ops_golden.items():
        if not key.startswith("phmlinear"):
            continue
        x = fx["x"].clone().requires_grad_(True)
        p = {"l.phm_rule": fx["A"].clone().requires_grad_(True), "l.W": fx["W"].clone().requires_grad_(True),
             "l.b": fx["b"].clone().requires_grad_(True)}
        y = O.phm_linear(x, p, "l")
        torch.testing.assert_close(y, fx["y"], rtol=RTOL, atol=ATOL)
        y.backward(fx["gy"])
        torch.testing.assert_close(x.grad, fx["gx"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(p["l.phm_rule"].grad, fx["gA"], rtol=RTOL, atol=1e-4)
        torch.testing.assert_close(p["l.W"].grad, fx["gW"], rtol=RTOL, atol=ATOL)
        torch.testing.assert_close(p["l.b"].grad, fx["gb"], rtol=RTOL, atol=ATOL)


def test_structure_oracle_small():
    ei = torch.tensor([[0, 1, 2, 2, 3, 0], [1, 0, 1, 3, 2, 1]])
    rowptr, col, perm = O.csr_by_target(ei, 5)
    assert rowptr.tolist() == [0, 1, 4, 5, 6, 6]
    assert perm.tolist() == [1, 0, 2, 5, 4, 3]
    assert col.tolist() == [1, 0, 2, 0, 3, 2]
    assert O.graph_ptr(torch.tensor([0, 0, 2, 2, 2]), 4).tolist() == [0, 2, 2, 5, 5]


def test_remove_isolated_nodes_known_answer():
    """Hand-worked example of torch_geometric 1.6.1 remove_isolated_nodes (documented behaviour: nodes without a non-loop
    edge are dropped together with their self loops; surviving self loops move behind the other edges)."""
    import torch
    from oracle import phc_oracle as O
    #           0->1  2->2  1->0  1->1  3->3  4->1
    ei = torch.tensor([[0, 2, 1, 1, 3, 4], [1, 2, 0, 1, 3, 1]])
    attr = torch.arange(6).view(6, 1)
    out, ea, mask = O.remove_isolated_nodes(ei, attr, 6)
    assert mask.tolist() == [True, True, False, False, True, False]
    assert out.tolist() == [[0, 1, 2, 1], [1, 0, 1, 1]]          # node 4 -> id 2; self loop of node 1 last
    assert ea.view(-1).tolist() == [0, 2, 5, 3]
