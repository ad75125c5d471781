"""CPU-side checks (no GPU): the C-ABI library builds, loads and exports every symbol declared in
include/phc_b200.h; the module API mirrors the reference's names / state-dict keys / parameter
counts; host-side helpers behave; and the product path refuses to run on CPU tensors."""
import ctypes
import os

import pytest
import torch

from conftest import ROOT, golden_cases, load_golden


def test_library_builds_loads_and_exports_header_symbols():
    from phc_gnn_b200 import _lib
    path = _lib.build()
    assert os.path.exists(path)
    decls = _lib.parse_header()
    assert len(decls) >= 20
    lib = ctypes.CDLL(path)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in include/phc_b200.h but not exported"
    loaded = _lib.load()
    assert loaded.phc_version() >= 100
    # pure host-side queries (no device work)
    assert loaded.phc_csr_workspace_bytes(10, 20) >= 4 * (40 + 40)
    assert loaded.phc_bn_workspace_bytes(1000, 64) > 0
    assert loaded.phc_phm_linear_bwd_workspace_bytes(128, 16, 16, 4, 0) > 0


def test_invalid_arguments_are_reported_not_thrown():
    from phc_gnn_b200 import _lib
    lib = _lib.load()
    rc = lib.phc_phm_linear_fwd(0, 0, 0, 0, 0, 0, 4, 10, 16, 4, 0, 0, 0, 0, 0)   # 10 % 4 != 0
    assert rc == 1
    assert b"not divisible" in lib.phc_last_error()
    with pytest.raises(AssertionError):
        _lib.check(rc, "phc_phm_linear_fwd")


@pytest.mark.parametrize("name", golden_cases())
def test_state_dict_and_param_count_match_reference(name):
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    fx = load_golden(name)
    model = PHMSkipConnectAdd(**fx["cfg"])
    model.load_state_dict(fx["state"], strict=True)
    assert model.get_number_of_params_() == fx["n_params"]
    assert set(k for k, _ in model.named_parameters()) >= set(fx["grads"].keys())


def test_known_parameter_counts():
    # published by the reference: benchmarks/inference.ipynb cells 18/25/32, benchmarks/README.md:93
    from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
    from phc_gnn_b200.synthetic import workloads
    assert PHMSkipConnectAdd(**workloads(4)["hiv"].model).get_number_of_params_() == 110909
    pcba2 = dict(workloads(2)["pcba"].model)
    assert PHMSkipConnectAdd(**pcba2).get_number_of_params_() == 1690328
    pcba8 = dict(workloads(8)["pcba"].model)
    assert PHMSkipConnectAdd(**pcba8).get_number_of_params_() == 688832
    zinc5 = dict(workloads(5)["zinc"].model)
    zinc5.update(atom_encoded_dim=200, mp_layers=[200] * 4, downstream_layers=[180, 80])   # benchmarks/zinc/experiment1/params.json
    assert PHMSkipConnectAdd(**zinc5).get_number_of_params_() == 106291


def test_rules_match_reference(ops_golden):
    from phc.hypercomplex.utils import get_multiplication_matrices
    for n in (1, 2, 3, 4, 5, 8):
        got = torch.stack(get_multiplication_matrices(n, type="standard"), 0)
        assert torch.equal(got, ops_golden[f"rule_standard_{n}"])
    r = get_multiplication_matrices(3, type="random")
    assert len(r) == 3 and r[0].shape == (3, 3) and float(torch.stack(r).abs().max()) <= 1.0


def test_kronecker_helpers_agree():
    # mirrors reference phc/hypercomplex/tests/test_kronecker_product.py
    from phc.hypercomplex.kronecker import kronecker_product, kronecker_product_einsum_batched, kronecker_product_single
    A, B = torch.randn(4, 4, 4), torch.randn(4, 16, 8)
    k1 = torch.stack([kronecker_product_single(a, b) for a, b in zip(A, B)], 0)
    torch.testing.assert_close(kronecker_product(A, B), k1)
    torch.testing.assert_close(kronecker_product_einsum_batched(A, B), k1)


def test_sum_kronecker_equals_quaternion_real_representation():
    # mirrors reference phc/hypercomplex/tests/test_realrepr_sumkronecker.py with RealP written out
    from phc.hypercomplex.kronecker import kronecker_product_einsum_batched
    from phc.hypercomplex.utils import get_multiplication_matrices
    r, i, j, k = (torch.randn(5, 3) for _ in range(4))
    realp = torch.cat([torch.cat([r, -i, -j, -k], 1), torch.cat([i, r, -k, j], 1),
                       torch.cat([j, k, r, -i], 1), torch.cat([k, -j, i, r], 1)], 0)
    A = torch.stack(get_multiplication_matrices(4), 0)
    H = kronecker_product_einsum_batched(A, torch.stack([r, i, j, k], 0)).sum(0)
    torch.testing.assert_close(H, realp)


def test_flat_alias_repacks_after_data_swap():
    from phc_gnn_b200.flat import alias_flat, is_packed
    ps = [torch.nn.Parameter(torch.randn(3)), torch.nn.Parameter(torch.randn(3))]
    want = torch.cat([p.detach() for p in ps])
    flat = alias_flat(None, ps)
    assert is_packed(flat, ps) and torch.equal(flat, want)
    flat[0] = 42.0
    assert float(ps[0][0]) == 42.0                       # same memory
    ps[1].data = torch.zeros(3)                           # what the reference's reset_parameters does
    assert not is_packed(flat, ps)
    flat2 = alias_flat(flat, ps)
    assert is_packed(flat2, ps) and torch.equal(flat2[3:], torch.zeros(3))


def test_flat_alias_reuses_enclosing_buffer():
    from phc_gnn_b200.flat import alias_flat
    big = torch.arange(10.0)
    ps = [torch.nn.Parameter(torch.zeros(2)), torch.nn.Parameter(torch.zeros(3))]
    ps[0].data = big[4:6]
    ps[1].data = big[6:9]                                 # back to back inside a larger buffer (the optimizer's)
    v = alias_flat(None, ps)
    assert v.data_ptr() == big[4:].data_ptr() and v.numel() == 5 and torch.equal(v, big[4:9])


def test_cpu_tensors_are_refused():
    from phc.hypercomplex.layers import PHMLinear
    lin = PHMLinear(8, 8, 4, c_init="standard")
    with pytest.raises(RuntimeError, match="no CPU path"):
        lin(torch.randn(3, 8))


def test_phmlinear_asserts_like_reference():
    from phc.hypercomplex.layers import PHMLinear
    with pytest.raises(AssertionError):
        PHMLinear(10, 8, 4)
    with pytest.raises(AssertionError):
        PHMLinear(8, 8, 4, w_init="glorot_uniform")     # the reference wants hyphens here (layers.py:228)


def test_import_changes_nothing_in_torch_and_compat_is_opt_in(monkeypatch):
    """``import phc`` must not touch global torch behaviour (no TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD, no patched schedulers); the shims
    the reference's unchanged scripts need are installed by phc.compat.enable() only."""
    import subprocess
    import sys
    code = ("import os, torch, inspect; import phc; "
            "assert 'TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD' not in os.environ; "
            "assert not getattr(torch.optim.lr_scheduler.StepLR.__init__, '_phc_compat', False); print('clean')")
    env = {k: v for k, v in os.environ.items() if k not in ("PHC_COMPAT", "TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD")}
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, env=env)
    assert out.returncode == 0 and "clean" in out.stdout, out.stderr[-1500:]
    import phc.compat
    saved = {c: c.__init__ for c in (torch.optim.lr_scheduler.ReduceLROnPlateau, torch.optim.lr_scheduler.StepLR)}
    try:
        phc.compat.enable()
        opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))])
        torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="max", factor=0.5, patience=3, verbose=True)
        torch.optim.lr_scheduler.StepLR(opt, step_size=10, gamma=0.5, verbose=True)
        assert "TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD" not in os.environ or os.environ.get("PHC_COMPAT") == "trust-checkpoints"
    finally:
        for c, init in saved.items():
            c.__init__ = init
    original = torch.load
    with phc.compat.trusted_load():
        assert torch.load is not original
    assert torch.load is original


def test_legacy_unpickler_refuses_foreign_globals(tmp_path):
    import pickle
    from phc_gnn_b200 import legacy
    path = tmp_path / "evil.pkl"
    with open(path, "wb") as fh:
        pickle.dump(os.system, fh)                      # a global outside torch / collections / numpy
    with open(path, "rb") as fh, pytest.raises(pickle.UnpicklingError):
        legacy._LegacyUnpickler(fh).load()


def test_synthetic_batches_are_deterministic_and_shaped():
    from phc_gnn_b200.synthetic import workloads, make_batch
    wl = workloads(4)["hiv"]
    a, b = make_batch(wl, seed=3, batch_graphs=8), make_batch(wl, seed=3, batch_graphs=8)
    assert torch.equal(a.edge_index, b.edge_index) and torch.equal(a.x, b.x)
    assert a.x.shape[1] == 9 and a.edge_attr.shape[1] == 3 and a.edge_index.dtype == torch.int64
    assert bool((a.batch[1:] >= a.batch[:-1]).all())
    ei = a.edge_index
    assert not bool((ei[0] == ei[1]).any())
    assert torch.unique(ei[0] * 100000 + ei[1]).numel() == ei.size(1)     # no duplicate edges
    k = make_batch(workloads(4)["mnist"], seed=1, batch_graphs=4)
    deg = torch.bincount(k.edge_index[1], minlength=k.x.size(0))
    assert bool((deg == 8).all())
