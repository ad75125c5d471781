"""The reference's quaternion model family served by the PHM kernels.

A quaternion linear map ``W (x) q`` (Hamilton product, reference phc/quaternion/layers.py:50-126) is the PHM layer with
n = 4, a FIXED multiplication rule and ``W'_c = W_c^T`` (see legacy.py; the relation is checked inside the reference by
oracle/make_golden_family.py).  Every other block of ``QuaternionSkipConnectAdd/Concat``
(phc/quaternion/undirectional/models.py:25-448) has the same arithmetic as its PHM counterpart at n = 4: the component-wise
batch norm (quaternion/norm.py:279-300), split activations and dropout, the GINE-style convolutions
(quaternion/undirectional/messagepassing.py), soft-attention / sum pooling (quaternion/pooling.py) and the downstream
network with its real transform (quaternion/downstream.py).  So the family is a thin subclass of the PHM models: n = 4,
``learn_phm=False``, the Hamilton rule, ``sc_type="first"``, the reference's constructor signature and initialisers —
and the tcgen05 PHMLinear, fused aggregation and batch-norm kernels run it unchanged.  Parameters live in the PHM layout;
``load_quaternion_state_dict`` / ``quaternion_state_dict`` translate from / to the reference's ``W_r..W_k`` names.

Model construction draws its weights from the reference's distributions but not with its RNG sequence (the reference
re-initialises every module several times while building; only the stand-alone initialisers are seed-compatible).

Not covered: ``norm="q-batch-norm"`` (4x4 whitening with a Cholesky factor per feature, quaternion/norm.py:104-276) has no PHM
counterpart and no kernel here; it raises at construction.
"""
from __future__ import annotations

import math
from typing import Optional, Union

import numpy as np
import torch

from . import legacy, ops
from .functional import glorot_normal, glorot_uniform, phm_cat
from .nn import ATOM_FEAT_DIMS, BOND_FEAT_DIMS, PHMLinear, PHMMLP, PHMSkipConnectAdd, PHMSkipConnectConcat, PHMSoftAttentionPooling

_INITS = ("glorot-normal", "glorot-uniform", "quaternion", "orthogonal")


# ------------------------------------------------------------------------------------------------- initialisers
def quaternion_init(in_features: int, out_features: int, criterion: str = "glorot", low: float = 0, high: float = 1) -> torch.Tensor:
    """[4, in, out] polar initialisation — reference phc/quaternion/inits.py:40-76: chi(4)-distributed modulus, a random
    unit imaginary axis, uniform phase, and the reference's extra cos^2 weighting of the three imaginary parts.  RNG
    streams are drawn in the reference's order (scipy chi, torch uniform x3, numpy uniform x4)."""
    from scipy.stats import chi
    if criterion == "glorot":
        s = 1.0 / math.sqrt(2 * (in_features + out_features))
    elif criterion == "he":
        s = 1.0 / math.sqrt(2 * in_features)
    else:
        raise ValueError("Invalid criterion: " + criterion)
    shape = (in_features, out_features)
    modulus = torch.from_numpy(chi.rvs(df=4, loc=0, scale=s, size=shape)).to(torch.float64)
    axis = torch.zeros(4, *shape, dtype=torch.float64)
    for c in range(1, 4):
        axis[c] = torch.empty(shape, dtype=torch.float64).uniform_(low, high)      # fp64 draws, as the reference's
    axis = axis / axis.norm(p=2, dim=0).clamp_min(1e-10)
    theta = torch.from_numpy(np.random.uniform(low=-np.pi, high=np.pi, size=shape)).to(torch.float64)
    share = torch.stack([torch.cos(torch.from_numpy(np.random.uniform(low=-s, high=s, size=shape)).to(torch.float64)) ** 2
                         for _ in range(3)], dim=0)
    share = share / share.sum(dim=0, keepdim=True)
    w = torch.empty(4, *shape, dtype=torch.float64)
    w[0] = modulus * torch.cos(theta)
    w[1:] = modulus * axis[1:] * torch.sin(theta) * share
    return w.to(torch.float32)


def _qmul(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Hamilton product of quaternion arrays stored as [4, ...]."""
    ar, ai, aj, ak = a
    br, bi, bj, bk = b
    return torch.stack([ar * br - ai * bi - aj * bj - ak * bk,
                        ar * bi + ai * br + aj * bk - ak * bj,
                        ar * bj - ai * bk + aj * br + ak * bi,
                        ar * bk + ai * bj - aj * bi + ak * br], dim=0)


def _qconj(a: torch.Tensor) -> torch.Tensor:
    return torch.cat([a[:1], -a[1:]], dim=0)


def quaternion_orthogonal_init(in_features: int, out_features: int, scale: float = 1.0) -> torch.Tensor:
    """[4, in, out] weight whose quaternion matrix has orthonormal columns (rows when it is wide) — the property the
    reference gets from a quaternion Householder QR of a Gaussian matrix (phc/quaternion/inits.py:79-113, qr.py:65-108; its
    ``Q /= 2`` undoes the factor 2 its QR carries, cf. phc/quaternion/tests/test_quat_qr.py:16-25, so the result has UNIT
    columns — checked against the reference: same column / row norms and element std).  Here: modified Gram-Schmidt in
    quaternion arithmetic (fp64) on the tall orientation; the two constructions agree up to a unit-quaternion phase per
    column, which the Gaussian draw makes immaterial."""
    rows, cols = max(in_features, out_features), min(in_features, out_features)
    a = torch.zeros(4, rows, cols, dtype=torch.float64).normal_(std=scale)
    for j in range(cols):
        v = a[:, :, j]
        for _ in range(2 if j else 1):                          # one re-orthogonalisation pass keeps fp64 orthogonality
            if j:
                # coefficients <q_l, v> = sum_r conj(q_l[r]) * v[r] for all previous columns l, then v -= q_l * coeff_l
                coeff = _qmul(_qconj(a[:, :, :j]), v.unsqueeze(-1).expand(4, rows, j)).sum(dim=1)         # [4, j]
                v = v - _qmul(a[:, :, :j], coeff.unsqueeze(1).expand(4, rows, j)).sum(dim=2)
        a[:, :, j] = v / v.pow(2).sum().sqrt()
    if in_features < out_features:                              # built as [out, in]: transpose the quaternion matrix
        a = a.permute(0, 2, 1)
    return a.contiguous().to(torch.float32)


@torch.no_grad()
def init_quaternion_linear(lin: PHMLinear, init: str) -> None:
    """``QLinear.reset_parameters`` (reference quaternion/layers.py:76-107) on a PHM-layout layer: weight by ``init``,
    bias 0 for the real part and 0.2 for the imaginary parts, rule = Hamilton."""
    assert lin.phm_dim == 4
    k, p = lin._in_feats_per_axis, lin._out_feats_per_axis
    if init == "quaternion":
        lin.W.copy_(quaternion_init(k, p))
    elif init == "orthogonal":
        lin.W.copy_(quaternion_orthogonal_init(k, p))
    elif init == "glorot-normal":
        for c in range(4):
            glorot_normal(lin.W[c])
    elif init == "glorot-uniform":
        for c in range(4):
            glorot_uniform(lin.W[c])
    else:
        raise ValueError(init)
    if lin.b is not None:
        lin.b[:p] = 0.0
        lin.b[p:] = 0.2
    lin.phm_rule.copy_(legacy.hamilton_rule().to(lin.phm_rule.device))


# ------------------------------------------------------------------------------------------------- models
class _QuaternionMixin(object):
    def _check(self, init, norm_mp, norm_dn, atom_encoded_dim, mp_layers, downstream_layers):
        assert init in _INITS, f"init variable '{init}' wrong."
        for nm in (norm_mp, norm_dn):
            if nm == "q-batch-norm":
                raise NotImplementedError("norm 'q-batch-norm' (4x4 whitening, reference quaternion/norm.py:104-276) is not "
                                          "implemented on the B200 path; use 'naive-batch-norm'")
            assert nm in ["None", None, "naive-batch-norm"]
        for d in [atom_encoded_dim] + list(mp_layers) + list(downstream_layers):
            assert d % 4 == 0, f"width {d} is not divisible by 4 (the reference would silently floor it)"

    def _hide_phm_attributes(self):
        """The training scripts tell the families apart by ``hasattr(model, "phm_dim")`` (reference
        benchmarks/train_hiv.py:182 picks quaternion_weight_regularization when it is absent), so a quaternion model must
        not expose it; the forward reads the private ``_n``."""
        self.__dict__.pop("phm_dim", None)
        for m in self.modules():          # every rule is the fixed Hamilton rule (the PHM GINE convs keep theirs trainable, nn.py)
            if isinstance(m, PHMLinear):
                m.phm_rule.requires_grad_(False)
                m.learn_phm = False

    def reset_parameters(self):
        super().reset_parameters()
        for m in self.modules():
            if isinstance(m, PHMLinear):
                init_quaternion_linear(m, self.init)

    def load_quaternion_state_dict(self, state_dict, strict: bool = True):
        """Load a state dict of the reference's quaternion model (``W_r``.. names, reference quaternion/layers.py:60-69)."""
        return self.load_state_dict(legacy.quaternion_to_phm_state_dict(state_dict), strict=strict)

    def quaternion_state_dict(self):
        return legacy.phm_to_quaternion_state_dict(self.state_dict())

    def quaternion_named_gradients(self):
        """{reference parameter name: gradient} — the PHM-layout gradients relabelled (dW_c = dW'_c^T)."""
        g = {n: p.grad for n, p in self.named_parameters() if p.grad is not None}
        return legacy.phm_to_quaternion_state_dict(g)


class QuaternionSkipConnectAdd(_QuaternionMixin, PHMSkipConnectAdd):
    """reference phc/quaternion/undirectional/models.py:25-230, on the PHM kernels (n = 4, Hamilton rule)."""

    def __init__(self, atom_input_dims: Union[int, list] = ATOM_FEAT_DIMS, atom_encoded_dim: int = 196,
                 bond_input_dims: Union[int, list] = BOND_FEAT_DIMS, naive_encoder: bool = False, init: str = "orthogonal",
                 same_dropout: bool = False, mp_layers: list = [196, 196, 196], bias: bool = True,
                 dropout_mpnn: list = [0.0, 0.0, 0.0], norm_mp: Optional[str] = "naive-batch-norm", add_self_loops: bool = True,
                 msg_aggr: str = "add", node_aggr: str = "sum", mlp: bool = False, pooling: str = "softattention",
                 activation: str = "relu", real_trafo: str = "linear", downstream_layers: list = [256, 128], target_dim: int = 1,
                 dropout_dn: Union[list, float] = [0.2, 0.1], norm_dn: Optional[str] = "naive-batch-norm",
                 msg_encoder: str = "identity", **kwargs) -> None:
        self._check(init, norm_mp, norm_dn, atom_encoded_dim, mp_layers, downstream_layers)
        self.init = init
        for k in ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type"):
            kwargs.pop(k, None)
        PHMSkipConnectAdd.__init__(
            self, phm_dim=4, learn_phm=False, phm_rule=None, atom_input_dims=atom_input_dims, atom_encoded_dim=atom_encoded_dim,
            bond_input_dims=bond_input_dims, naive_encoder=naive_encoder, w_init="phm", c_init="standard", same_dropout=same_dropout,
            mp_layers=mp_layers, bias=bias, dropout_mpnn=dropout_mpnn, norm_mp=norm_mp, add_self_loops=add_self_loops,
            msg_aggr=msg_aggr, node_aggr=node_aggr, mlp=mlp, pooling=pooling, activation=activation, real_trafo=real_trafo,
            downstream_layers=downstream_layers, target_dim=target_dim, dropout_dn=dropout_dn, norm_dn=norm_dn,
            msg_encoder=msg_encoder, sc_type="first", **kwargs)
        self._hide_phm_attributes()


class QuaternionSkipConnectConcat(_QuaternionMixin, PHMSkipConnectConcat):
    """reference phc/quaternion/undirectional/models.py:234-448, on the PHM kernels (n = 4, Hamilton rule)."""

    def _skip_concat(self, z: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        """``qcat([q, atom_encoded], dim=-1)`` (reference phc/quaternion/undirectional/models.py:407, algebra.py cat): every component's block becomes
        [z_c | skip_c] — in the flat layout that is the component-aware ``phm_cat``, not a flat concat."""
        return phm_cat([z, skip], 4)

    def __init__(self, atom_input_dims: Union[int, list] = ATOM_FEAT_DIMS, atom_encoded_dim: int = 128,
                 bond_input_dims: Union[int, list] = BOND_FEAT_DIMS, naive_encoder: bool = False, init: str = "orthogonal",
                 same_dropout: bool = False, mp_layers: list = [128, 196, 256], bias: bool = True,
                 dropout_mpnn: list = [0.0, 0.0, 0.0], norm_mp: Optional[str] = "naive-batch-norm", add_self_loops: bool = True,
                 msg_aggr: str = "add", node_aggr: str = "sum", mlp: bool = False, pooling: str = "softattention",
                 activation: str = "relu", real_trafo: str = "linear", downstream_layers: list = [256, 128], target_dim: int = 1,
                 dropout_dn: Union[list, float] = [0.2, 0.1], norm_dn: Optional[str] = "naive-batch-norm",
                 msg_encoder: str = "identity", **kwargs) -> None:
        self._check(init, norm_mp, norm_dn, atom_encoded_dim, mp_layers, downstream_layers)
        self.init = init
        for k in ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type"):
            kwargs.pop(k, None)
        PHMSkipConnectConcat.__init__(
            self, phm_dim=4, learn_phm=False, phm_rule=None, atom_input_dims=atom_input_dims, atom_encoded_dim=atom_encoded_dim,
            bond_input_dims=bond_input_dims, naive_encoder=naive_encoder, w_init="phm", c_init="standard", same_dropout=same_dropout,
            mp_layers=mp_layers, bias=bias, dropout_mpnn=dropout_mpnn, norm_mp=norm_mp, add_self_loops=add_self_loops,
            msg_aggr=msg_aggr, node_aggr=node_aggr, mlp=mlp, pooling=pooling, activation=activation, real_trafo=real_trafo,
            downstream_layers=downstream_layers, target_dim=target_dim, dropout_dn=dropout_dn, norm_dn=norm_dn,
            msg_encoder=msg_encoder, sc_type="first", **kwargs)
        self._hide_phm_attributes()


# ------------------------------------------------------------------------------------------------- regulariser
def quaternion_weights(model) -> list:
    """The [4, in, out] weight stacks ``quaternion_weight_regularization`` sums over, in the reference's order
    (phc/quaternion/regularization.py:27-97): message-passing transforms, the soft-attention pooling layer — for which
    the reference stacks ``W_r, W_i, W_k, W_k`` (its line 81 repeats W_k and drops W_j; reproduced) — and the downstream
    affine layers."""
    root = getattr(model, "module", model)
    ws = []
    for mp in root.convs:
        t = mp.transform.transform
        if isinstance(t, PHMMLP):
            ws += [t.linear1.W, t.linear2.W]
        elif isinstance(t, PHMLinear):
            ws.append(t.W)
    if isinstance(root.pooling, PHMSoftAttentionPooling):
        w = root.pooling.linear.W
        ws.append(w.index_select(0, torch.tensor([0, 1, 3, 3], device=w.device)))
    for lin in root.downstream.affine:
        ws.append(lin.W)
    return ws


def quaternion_weight_regularization(model, device=None, p: int = 1):
    """sum over the model's quaternion weights of ``stack(W_r..W_k).norm(p, dim=0).mean()`` — reference
    phc/quaternion/regularization.py:27-97 (undirectional models).  The component norm does not depend on the [out,in] /
    [in,out] orientation, so the PHM-layout stacks are used as they are; p = 2 runs the fused regulariser kernel."""
    assert p in [1, 2]
    ws = quaternion_weights(model)
    if p == 2 and ws and all(w.is_cuda for w in ws) and len(ws) <= 256:
        return ops.weight_regularization_l2([w.contiguous() for w in ws])
    reg = 0.0
    for w in ws:
        reg = reg + w.norm(p=p, dim=0).mean()
    return reg
