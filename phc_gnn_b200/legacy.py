"""Importers for the reference's other parameter layouts.

Two layouts exist next to the current ``phc.hypercomplex`` one, and both describe maps the PHM kernels already run:

* the **legacy PHM layout** of ``PHMLinear_Old`` (reference phc/hypercomplex/layers.py:114-192) that the shipped
  checkpoints ``benchmarks/{hiv,zinc}/experiment1/run_1/model.pt`` use: per-component ``ParameterList``s
  ``W.i`` ``[out/n, in/n]``, ``b.i`` ``[out/n]``, ``phm_rule.i`` ``[n, n]``, evaluated as ``H = sum_i kron(A_i, W_i)``,
  ``y = H x`` (reference layers.py:58-78).  The current layer computes ``y = x H'`` with
  ``H' = sum_i kron(A'_i, W'_i)`` (layers.py:198-219), hence ``W'_i = W_i^T``, ``A'_i = A_i^T``, ``b' = cat(b_i)``;
* the **quaternion layout** of ``QLinear`` (reference phc/quaternion/layers.py:50-126): ``W_r, W_i, W_j, W_k``
  ``[out, in]`` and ``b_r..b_k``, evaluated as the Hamilton product ``W (x) q``.  That is the PHM layer with n = 4,
  ``W'_c = W_c^T`` and the fixed rule ``A'_c = (left-multiplication matrix of the c-th unit)^T``
  (``get_multiplication_matrices(4, "standard")`` holds the untransposed matrices, reference
  phc/hypercomplex/utils.py:5-22 and phc/quaternion/tests/test_realrepr_sumkronecker.py:14-33).

Everything here is tensor re-labelling on whatever device the tensors live on; no kernel is involved.  The shipped
checkpoints are whole-module pickles of classes that no longer exist under their pickled paths
(``hypercomplex.layers.PHMLinear`` ...), so ``read_legacy_checkpoint`` unpickles them onto placeholder
``nn.Module`` classes and recovers (model class name, constructor arguments, converted state dict).
"""
from __future__ import annotations

import pickle
import re
import types
from collections import OrderedDict
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .functional import get_multiplication_matrices

_LEGACY_KEY = re.compile(r"^(?:(.*)\.)?(W|b|phm_rule)\.(\d+)$")
_QUAT_W = re.compile(r"^(?:(.*)\.)?W_([rijk])$")
_QUAT_B = re.compile(r"^(?:(.*)\.)?b_([rijk])$")
_COMP = {"r": 0, "i": 1, "j": 2, "k": 3}
_COMP_NAME = "rijk"


def _join(prefix, leaf: str) -> str:
    return f"{prefix}.{leaf}" if prefix else leaf


def hamilton_rule() -> torch.Tensor:
    """[4,4,4] rule A' with  y = x (sum_c A'_c (x) W_c^T)  ==  Hamilton product  W (x) q."""
    return torch.stack([a.t() for a in get_multiplication_matrices(4, type="standard")], dim=0).contiguous()


# ---------------------------------------------------------------------------------------------- legacy PHM layout
def is_legacy_phm_state_dict(sd: Dict[str, torch.Tensor]) -> bool:
    return any(_LEGACY_KEY.match(k) for k in sd)


def convert_legacy_phm_state_dict(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """``PHMLinear_Old`` keys -> current ``PHMLinear`` keys; every other entry is passed through unchanged."""
    groups: Dict[Tuple[str, str], Dict[int, torch.Tensor]] = {}
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in sd.items():
        m = _LEGACY_KEY.match(k)
        if m is None:
            out[k] = v
            continue
        prefix, what, idx = m.group(1) or "", m.group(2), int(m.group(3))
        key = (prefix, what)
        if key not in groups:
            groups[key] = {}
            out[_join(prefix, what)] = None            # keeps the position of the first component
        groups[key][idx] = v
    for (prefix, what), parts in groups.items():
        n = len(parts)
        assert sorted(parts) == list(range(n)), f"{prefix}.{what}: components {sorted(parts)} are not 0..{n - 1}"
        seq = [parts[i] for i in range(n)]
        if what == "b":
            out[_join(prefix, "b")] = torch.cat([t.reshape(-1) for t in seq], dim=0)
        else:                                             # W_i^T / A_i^T
            out[_join(prefix, what)] = torch.stack([t.t() for t in seq], dim=0).contiguous()
    return out


def to_legacy_phm_state_dict(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Inverse of :func:`convert_legacy_phm_state_dict` (current layout -> ``PHMLinear_Old`` keys)."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in sd.items():
        prefix, _, leaf = k.rpartition(".")
        if leaf in ("W", "phm_rule") and v.dim() == 3:
            for i in range(v.size(0)):
                out[f"{k}.{i}"] = v[i].t().contiguous()
        elif leaf == "b" and v.dim() == 1 and _join(prefix, "W") in sd:
            n = sd[_join(prefix, "W")].size(0)
            for i, part in enumerate(v.chunk(n)):
                out[f"{k}.{i}"] = part.clone()
        else:
            out[k] = v
    return out


# ---------------------------------------------------------------------------------------------- quaternion layout
def _quat_module_key(k: str) -> str:
    """Module-path differences between the quaternion and the PHM trees: QMLP's ``qlinear1/2`` are PHMMLP's
    ``linear1/2`` (reference quaternion/layers.py:139-142 vs hypercomplex/layers.py:326-331); the encoders' and the
    naive batch-norm's children named r,i,j,k are list entries 0..3 (quaternion/encoder.py:63-79, quaternion/norm.py:279-300
    vs hypercomplex/encoder.py:18-25, hypercomplex/norm.py)."""
    k = k.replace(".qlinear1.", ".linear1.").replace(".qlinear2.", ".linear2.")
    k = re.sub(r"\.bn\.bn\.([rijk])\.", lambda m: f".bn.bn.{_COMP[m.group(1)]}.", k)
    k = re.sub(r"^(atomencoder|bondencoders\.\d+)\.([rijk])\.", lambda m: f"{m.group(1)}.encoders.{_COMP[m.group(2)]}.", k)
    return k


def _phm_module_key_to_quat(k: str) -> str:
    k = k.replace(".linear1.", ".qlinear1.").replace(".linear2.", ".qlinear2.")
    k = re.sub(r"\.bn\.bn\.([0-3])\.", lambda m: f".bn.bn.{_COMP_NAME[int(m.group(1))]}.", k)
    k = re.sub(r"^(atomencoder|bondencoders\.\d+)\.encoders\.([0-3])\.", lambda m: f"{m.group(1)}.{_COMP_NAME[int(m.group(2))]}.", k)
    return k


def is_quaternion_state_dict(sd: Dict[str, torch.Tensor]) -> bool:
    return any(_QUAT_W.match(k) for k in sd)


def quaternion_to_phm_state_dict(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """State dict of a reference ``QuaternionSkipConnectAdd/Concat`` -> the PHM(n=4, Hamilton rule) layout."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    ws: Dict[str, Dict[int, torch.Tensor]] = {}
    bs: Dict[str, Dict[int, torch.Tensor]] = {}
    for k, v in sd.items():
        mw, mb = _QUAT_W.match(k), _QUAT_B.match(k)
        if mw:
            prefix = _quat_module_key((mw.group(1) or "") + ".")[:-1]
            if prefix not in ws:
                ws[prefix] = {}
                out[_join(prefix, "phm_rule")] = None
                out[_join(prefix, "W")] = None
            ws[prefix][_COMP[mw.group(2)]] = v
        elif mb:
            prefix = _quat_module_key((mb.group(1) or "") + ".")[:-1]
            if prefix not in bs:
                bs[prefix] = {}
                out[_join(prefix, "b")] = None
            bs[prefix][_COMP[mb.group(2)]] = v
        else:
            out[_quat_module_key(k)] = v
    for prefix, parts in ws.items():
        assert sorted(parts) == [0, 1, 2, 3], f"{prefix}: quaternion weight needs W_r, W_i, W_j, W_k"
        w = torch.stack([parts[c].t() for c in range(4)], dim=0).contiguous()
        out[_join(prefix, "W")] = w
        out[_join(prefix, "phm_rule")] = hamilton_rule().to(device=w.device, dtype=w.dtype)
    for prefix, parts in bs.items():
        assert sorted(parts) == [0, 1, 2, 3], f"{prefix}: quaternion bias needs b_r, b_i, b_j, b_k"
        out[_join(prefix, "b")] = torch.cat([parts[c].reshape(-1) for c in range(4)], dim=0)
    return out


def phm_to_quaternion_state_dict(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Inverse relabelling (also maps a gradient dict of the PHM model onto the quaternion parameter names);
    ``phm_rule`` entries are dropped — the quaternion layout has no rule parameter."""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, v in sd.items():
        prefix, _, leaf = k.rpartition(".")
        qprefix = _phm_module_key_to_quat(prefix + ".")[:-1]
        if leaf == "phm_rule":
            continue
        if leaf == "W" and v.dim() == 3 and v.size(0) == 4:
            for c in range(4):
                out[_join(qprefix, f"W_{_COMP_NAME[c]}")] = v[c].t().contiguous()
        elif leaf == "b" and v.dim() == 1 and _join(prefix, "W") in sd:
            for c, part in enumerate(v.chunk(4)):
                out[_join(qprefix, f"b_{_COMP_NAME[c]}")] = part.clone()
        else:
            out[_phm_module_key_to_quat(k)] = v
    return out


# ---------------------------------------------------------------------------------------------- shipped checkpoints
class _Placeholder(nn.Module):
    """Stands in for a pickled reference class: ``nn.Module`` pickles restore ``__dict__`` only, so parameters, buffers,
    children and the constructor arguments the reference stored as attributes all survive."""


def _placeholder_function(*args, **kwargs):
    raise RuntimeError("placeholder for a pickled reference function; legacy modules are not executable")


_FOREIGN_ROOTS = ("hypercomplex", "quaternion", "phc", "torch_geometric", "torch_scatter", "ogb")
# everything else a module pickle may legitimately name: tensor rebuilders, containers, dtypes — nothing executable beyond them
_ALLOWED_ROOTS = ("torch", "collections", "numpy")
# inert stdlib data classes found in the shipped checkpoints (PyG's MessagePassing keeps inspect.signature() results)
_ALLOWED_GLOBALS = {("copyreg", "_reconstructor"), ("inspect", "Parameter"), ("inspect", "_ParameterKind"), ("inspect", "_empty"),
                    ("inspect", "Signature"), ("typing", "Any"), ("typing", "Optional"), ("typing", "Union"),
                    ("_operator", "getitem"), ("typing", "Tuple"), ("typing", "List"), ("typing", "Dict"), ("typing", "Callable")}
_ALLOWED_BUILTINS = {"set", "frozenset", "list", "dict", "tuple", "int", "float", "bool", "str", "bytes", "complex", "slice", "range",
                     "object", "long", "unicode", "type"}


class _LegacyUnpickler(pickle.Unpickler):
    _classes: Dict[Tuple[str, str], object] = {}

    def find_class(self, module, name):
        if module.split(".")[0] in _FOREIGN_ROOTS:
            key = (module, name)
            if key not in self._classes:
                self._classes[key] = (type(name, (_Placeholder,), {"__module__": "phc_legacy." + module})
                                      if name[:1].isupper() else _placeholder_function)
            return self._classes[key]
        root = module.split(".")[0]
        if root in _ALLOWED_ROOTS or (root in ("builtins", "__builtin__") and name in _ALLOWED_BUILTINS) or \
                (module, name) in _ALLOWED_GLOBALS:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"legacy checkpoint names {module}.{name}: only torch / collections / numpy globals and the "
                                     f"reference's own (placeholder) classes are accepted")


def _pickle_module():
    mod = types.ModuleType("pickle")
    mod.Unpickler = _LegacyUnpickler
    mod.load = lambda f, **kw: _LegacyUnpickler(f, **kw).load()
    return mod


def _legacy_constructor_args(root: nn.Module) -> dict:
    """Constructor arguments of the current ``PHMSkipConnectAdd/Concat`` from the attributes the legacy model stored.
    The legacy model kept widths PER COMPONENT (``atom_encoded_dim``, ``mp_layers``, ``downstream_layers`` were divided
    by ``phm_dim`` in its constructor, as the quaternion models still do, reference quaternion/undirectional/models.py:56-58);
    the current constructor takes total widths."""
    a = vars(root)
    n = int(a["phm_dim"])
    cfg = dict(phm_dim=n, learn_phm=bool(a.get("learn_phm", True)), phm_rule=None,
               atom_input_dims=a["atom_input_dims"], bond_input_dims=a["bond_input_dims"],
               atom_encoded_dim=int(a["atom_encoded_dim"]) * n, naive_encoder=bool(a.get("naive_encoder", False)),
               w_init=a.get("init", "phm"), c_init="standard", same_dropout=bool(a.get("same_dropout", False)),
               mp_layers=[int(d) * n for d in a["mp_layers"]], bias=bool(a.get("bias", True)),
               dropout_mpnn=list(a["dropout_mpnn"]), norm_mp=a.get("norm_mp"), add_self_loops=bool(a.get("add_self_loops", True)),
               msg_aggr=a.get("msg_aggr_type", "add"), node_aggr=a.get("node_aggr_type", "sum"), mlp=bool(a.get("mlp_mp", False)),
               pooling=a.get("pooling_type", "softattention"), activation=a.get("activation_str", "relu"),
               real_trafo=a.get("real_trafo_type", "linear"), downstream_layers=[int(d) * n for d in a["downstream_layers"]],
               target_dim=int(a.get("target_dim", 1)), dropout_dn=a.get("dropout_dn", [0.2, 0.1]), norm_dn=a.get("norm_dn_type"),
               msg_encoder=a.get("msg_encoder_str", "identity"), sc_type=a.get("sc_type", "first"))
    if cfg["w_init"] not in ("phm", "glorot_uniform", "glorot_normal"):
        cfg["w_init"] = "phm"
    convs = root._modules.get("convs")
    if convs is not None and len(convs._modules):
        inner = next(iter(convs._modules.values()))._modules.get("transform")
        if inner is not None and "initial_beta" in vars(inner):          # softmax aggregation
            cfg["initial_beta"] = float(vars(inner)["initial_beta"])
            cfg["learn_beta"] = bool(vars(inner)["learn_beta"])
    return cfg


def read_legacy_checkpoint(path: str) -> Tuple[str, dict, "OrderedDict[str, torch.Tensor]"]:
    """-> (model class name, constructor kwargs for the current class, state dict in the current layout), on the CPU."""
    root = torch.load(path, pickle_module=_pickle_module(), weights_only=False, map_location="cpu")
    if isinstance(root, dict):                                   # a plain state dict
        sd = root
        kind, cfg = "state_dict", {}
    else:
        assert isinstance(root, nn.Module), f"{path}: expected a pickled module or a state dict, got {type(root)}"
        kind, sd = type(root).__name__, root.state_dict()
        cfg = _legacy_constructor_args(root) if kind.startswith("PHMSkipConnect") else {}
    if is_legacy_phm_state_dict(sd):
        sd = convert_legacy_phm_state_dict(sd)
    return kind, cfg, sd


def load_legacy_checkpoint(path: str, device=None) -> nn.Module:
    """The shipped ``model.pt`` as a current-layout model running on the B200 kernels (eval mode, as saved)."""
    from . import nn as _nn
    kind, cfg, sd = read_legacy_checkpoint(path)
    if kind not in ("PHMSkipConnectAdd", "PHMSkipConnectConcat"):
        raise ValueError(f"{path}: no importer for a pickled {kind}")
    model = getattr(_nn, kind)(**cfg)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model.to(device) if device is not None else model
