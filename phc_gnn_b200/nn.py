"""torch.nn.Module mirror of the reference's ``phc.hypercomplex`` operator API, executing on the
hand-written sm_100a kernels (ops.py -> libphc_b200.so).

Class names, constructor signatures, forward signatures, parameter names/shapes and state-dict
keys follow the reference so that ``benchmarks/train_*.py`` and saved
state-dicts keep working; the implementation underneath is new.  Each class cites the reference
location it stands in for.
"""
from __future__ import annotations

from typing import List, Optional, Union

import torch
import torch.nn as nn

from . import ops
from .flat import alias_flat
from .functional import get_multiplication_matrices, glorot_normal, glorot_uniform, phm_init
from .graph import edge_structure, segment_structure

ATOM_FEAT_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]   # ogb.utils.features.get_atom_feature_dims() (ogb 1.2.4)
BOND_FEAT_DIMS = [5, 6, 2]                           # ogb.utils.features.get_bond_feature_dims()

_ACTIVATIONS = ("relu", "lrelu", "elu", "selu", "swish", "identity")


def get_module_activation(activation: str) -> nn.Module:
    """reference phc/quaternion/activations.py:134-147 (stand-alone torch modules; the fused kernels
    take the activation by name instead)."""
    a = activation.lower()
    table = {"relu": nn.ReLU, "lrelu": nn.LeakyReLU, "elu": nn.ELU, "selu": nn.SELU, "swish": nn.SiLU, "identity": nn.Identity}
    return table[a]() if a in table else None


def _norm_on(norm) -> bool:
    return norm not in (None, "None")


# ============================================================================ PHMLinear & friends
class PHMLinear(nn.Module):
    """y = x (sum_i A_i (x) W_i) + b   — reference phc/hypercomplex/layers.py:222-299.

    Parameters: ``phm_rule`` [n,n,n] (trainable iff learn_phm), ``W`` [n, in/n, out/n], ``b`` [out].
    The Kronecker weight is never materialised (kernel: csrc/phm_linear_*.cu).
    """

    def __init__(self, in_features: int, out_features: int, phm_dim: int,
                 phm_rule: Union[None, torch.Tensor] = None, bias: bool = True, w_init: str = "phm",
                 c_init: str = "random", learn_phm: bool = True) -> None:
        super().__init__()
        assert w_init in ["phm", "glorot-normal", "glorot-uniform"]
        assert c_init in ["standard", "random"]
        assert in_features % phm_dim == 0, f"Argument `in_features`={in_features} is not divisble be `phm_dim`{phm_dim}"
        assert out_features % phm_dim == 0, f"Argument `out_features`={out_features} is not divisble be `phm_dim`{phm_dim}"
        self.in_features, self.out_features = in_features, out_features
        self.phm_dim, self.learn_phm = phm_dim, learn_phm
        self._in_feats_per_axis = in_features // phm_dim
        self._out_feats_per_axis = out_features // phm_dim
        self.bias_flag, self.w_init, self.c_init = bias, w_init, c_init
        rule = phm_rule if phm_rule is not None else get_multiplication_matrices(phm_dim, type=c_init)
        self.phm_rule = nn.Parameter(torch.stack([r.detach().clone().float() for r in rule], dim=0), requires_grad=learn_phm)
        self.W = nn.Parameter(torch.empty(phm_dim, self._in_feats_per_axis, self._out_feats_per_axis))
        if bias:
            self.b = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("b", None)
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self):
        n, k, p = self.phm_dim, self._in_feats_per_axis, self._out_feats_per_axis
        if self.w_init == "phm":
            self.W.copy_(phm_init(n, k, p, transpose=False))
        elif self.w_init == "glorot-normal":
            for i in range(n):
                glorot_normal(self.W[i])
        elif self.w_init == "glorot-uniform":
            for i in range(n):
                glorot_uniform(self.W[i])
        else:
            raise ValueError(self.w_init)
        if self.bias_flag:
            # first component 0, the others 0.2.  (The reference leaves element b[out/n] uninitialised —
            # an off-by-one at layers.py:277-278; we give it the evident 0.2.)
            self.b[:p] = 0.0
            self.b[p:] = 0.2
        # the rule is always re-initialised from c_init, as the reference does (layers.py:281)
        self.phm_rule.copy_(torch.stack(get_multiplication_matrices(n, type=self.c_init), dim=0))

    def forward(self, x: torch.Tensor, phm_rule=None, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        return ops.phm_linear(x, self.phm_rule, self.W, self.b, residual)

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, "
                f"phm_dim={self.phm_dim}, bias={self.bias_flag}, w_init={self.w_init}, c_init={self.c_init}, "
                f"learn_phm={self.learn_phm})")


def phm_dropout(x: torch.Tensor, phm_dim: int, p: float = 0.2, training: bool = True, same: bool = False) -> torch.Tensor:
    """reference phc/hypercomplex/layers.py:31-55.  ``same=True`` shares one Bernoulli mask between the
    n components of a feature."""
    assert 0.0 <= p <= 1.0, f"dropout rate must be in [0.0 ; 1.0]. {p} was inserted!"
    if not (training and p > 0.0):
        return x
    return ops.bn_act_drop_skip(x, None, phm_dim=phm_dim, use_bn=False, training=True, drop_p=p, drop_same=same)


class NaivePHMNorm(nn.Module):
    """n independent BatchNorm1d's, one per component column block — reference phc/hypercomplex/norm.py:5-39.
    Executed as one per-column batch-norm over the flat [M,F] matrix (identical arithmetic)."""

    def __init__(self, num_features: int, phm_dim: int, momentum: float = 0.1, eps: float = 1e-5,
                 affine: bool = True, track_running_stats: bool = True) -> None:
        super().__init__()
        assert num_features % phm_dim == 0
        assert momentum is not None, "cumulative moving average (momentum=None) is not supported"
        self.phm_dim = phm_dim
        self.num_features = num_features // phm_dim
        self.momentum, self.eps, self.affine, self.track_running_stats = momentum, eps, affine, track_running_stats
        self.bn = nn.ModuleList([nn.BatchNorm1d(self.num_features, eps, momentum, affine, track_running_stats)
                                 for _ in range(phm_dim)])
        self._flat = [None] * 5
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.bn:
            m.reset_parameters()

    def __getstate__(self):          # flat caches are re-derived after unpickling
        d = dict(self.__dict__)
        d["_flat"] = [None] * 5
        d.pop("_flat_cache", None)
        return d

    def flat_views(self):
        """(gamma, beta, running_mean, running_var, num_batches_tracked) flat vectors aliasing the
        n BatchNorm1d modules' tensors."""
        # hot path: the modules' tensors still sit where the cached flat vectors alias them (checked on the first tensor
        # of each kind; a module is moved / re-initialised as a whole)
        b0 = self.bn[0]
        key = (b0.weight.data_ptr() if self.affine else 0, b0.bias.data_ptr() if self.affine else 0,
               b0.running_mean.data_ptr() if self.track_running_stats else 0)
        cached = self.__dict__.get("_flat_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        groups = []
        if self.affine:
            groups += [[m.weight for m in self.bn], [m.bias for m in self.bn]]
        else:
            groups += [None, None]
        if self.track_running_stats:
            groups += [[m.running_mean for m in self.bn], [m.running_var for m in self.bn],
                       [m.num_batches_tracked for m in self.bn]]
        else:
            groups += [None, None, None]
        for i, g in enumerate(groups):
            self._flat[i] = alias_flat(self._flat[i], g) if g is not None else None
        out = tuple(self._flat)
        b0 = self.bn[0]                  # alias_flat may have re-pointed the tensors: key on the final addresses
        key = (b0.weight.data_ptr() if self.affine else 0, b0.bias.data_ptr() if self.affine else 0,
               b0.running_mean.data_ptr() if self.track_running_stats else 0)
        self.__dict__["_flat_cache"] = (key, out)
        return out

    def autograd_params(self):
        return ([m.weight for m in self.bn] + [m.bias for m in self.bn]) if self.affine else []

    def fused(self, x, skip=None, act: str = "identity", drop_p: float = 0.0, drop_same: bool = False, dropout_training=None,
              concat_right=None):
        training = self.training or not self.track_running_stats
        return ops.bn_act_drop_skip(x, skip, phm_dim=self.phm_dim, flat=self.flat_views(), params=self.autograd_params(),
                                    use_bn=True, training=training, momentum=self.momentum, eps=self.eps, act=act,
                                    drop_p=drop_p if self.training else 0.0, drop_same=drop_same, concat_right=concat_right)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fused(x)

    def __repr__(self):
        return (f"{self.num_features}, phm_dim={self.phm_dim}, eps={self.eps}, momentum={self.momentum}, "
                f"affine={self.affine}, track_running_stats={self.track_running_stats})")


class PHMNorm(nn.Module):
    """reference phc/hypercomplex/norm.py:45-74.  Only "naive-batch-norm" is usable in the reference
    ("naive-naive-batch-norm" is mis-sized there); the same holds here."""

    def __init__(self, num_features: int, phm_dim: int, type: str = "naive-batch-norm", **kwargs):
        super().__init__()
        assert type in ["naive-batch-norm", "naive-naive-batch-norm"]
        if type != "naive-batch-norm":
            raise ValueError("only 'naive-batch-norm' is supported ('naive-naive-batch-norm' is mis-sized in the reference)")
        self.type, self.num_features, self.phm_dim, self.kwargs = type, num_features, phm_dim, kwargs
        self.bn = NaivePHMNorm(num_features=num_features, phm_dim=phm_dim, **kwargs)

    def reset_parameters(self):
        self.bn.reset_parameters()

    def fused(self, x, **kw):
        return self.bn.fused(x, **kw)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.bn(x)

    def __repr__(self):
        return f"{self.__class__.__name__}:(num_features={self.num_features}, phm_dim={self.phm_dim} type={self.type})"


def norm_act_drop_skip(norm: Optional[PHMNorm], x, skip, act: str, phm_dim: int, training: bool, drop_p: float = 0.0,
                       drop_same: bool = False, concat_right: Optional[torch.Tensor] = None) -> torch.Tensor:
    """skip + dropout(act(norm(x))) as ONE kernel pair — reference models.py:206-215, layers.py:349-355,
    downstream.py:98-113 spell this as 4-5 separate modules.  ``concat_right``: [ dropout(act(norm(x))) | concat_right ] as one
    buffer (the concat skip connection of models.py:467 without the concatenation pass)."""
    if norm is not None:
        return norm.fused(x, skip=skip, act=act, drop_p=drop_p, drop_same=drop_same, concat_right=concat_right)
    if act == "identity" and skip is None and concat_right is None and not (training and drop_p > 0.0):
        return x
    return ops.bn_act_drop_skip(x, skip, phm_dim=phm_dim, use_bn=False, training=training, act=act, drop_p=drop_p,
                                drop_same=drop_same, concat_right=concat_right)


class PHMMLP(nn.Module):
    """linear2(act(norm(linear1(x)))) — reference phc/hypercomplex/layers.py:304-369."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int, phm_rule, bias: bool = True, learn_phm: bool = True,
                 activation: str = "relu", norm: Union[None, str] = None, w_init: str = "phm", c_init: str = "standard",
                 factor: float = 1, **kwargs) -> None:
        super().__init__()
        assert norm in ["None", None, "naive-batch-norm", "naive-naive-batch-norm"]
        assert activation.lower() in _ACTIVATIONS
        self.in_features, self.out_features, self.phm_dim = in_features, out_features, phm_dim
        self.bias_flag, self.learn_phm, self.phm_rule = bias, learn_phm, phm_rule
        self.activation_str, self.norm_type, self.factor = activation, norm, factor
        self.w_init, self.c_init = w_init, c_init
        hidden = int(factor * out_features)
        self.linear1 = PHMLinear(in_features, hidden, phm_dim, phm_rule, bias, w_init, c_init, learn_phm)
        self.linear2 = PHMLinear(hidden, out_features, phm_dim, phm_rule, bias, w_init, c_init, learn_phm)
        self.norm_flag = norm in ["naive-batch-norm", "naive-naive-batch-norm"]
        if self.norm_flag:
            self.norm = PHMNorm(num_features=hidden, phm_dim=phm_dim, type=norm, **kwargs)
        self.reset_parameters()

    def reset_parameters(self):
        self.linear1.reset_parameters()
        self.linear2.reset_parameters()
        if self.norm_flag:
            self.norm.reset_parameters()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.linear1(x)
        h = norm_act_drop_skip(self.norm if self.norm_flag else None, h, None, self.activation_str.lower(), self.phm_dim,
                               self.training)
        return self.linear2(h)


class RealTransformer(nn.Module):
    """H^n -> R map — reference phc/hypercomplex/layers.py:372-420.  "linear" is a dense
    nn.Linear(F -> F/n) (run on the PHM kernel with n=1).  The other modes reproduce the reference's
    observable behaviour: its split never splits, so they return the input unreduced."""

    def __init__(self, type: str, in_features: int, phm_dim: int, bias: bool = True) -> None:
        super().__init__()
        assert type in ["linear", "sum", "mean", "norm"]
        self.type, self.in_features, self.phm_dim, self.bias_flag = type, in_features, phm_dim, bias
        self.affine = nn.Linear(in_features, in_features // phm_dim, bias=bias) if type == "linear" else None
        self._unit_rule = None          # [1,1,1] ones on the weight's device, created on first use (NOT a buffer: the
        self.reset_parameters()         # reference module has none, and named_buffers() must list the same tensors)

    def reset_parameters(self):
        if self.type == "linear":
            nn.init.xavier_uniform_(self.affine.weight)
            if self.bias_flag:
                self.affine.bias.data.fill_(0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.type == "linear":
            w = self.affine.weight.t().contiguous().unsqueeze(0)       # [1, in, out]
            if self._unit_rule is None or self._unit_rule.device != w.device:
                self._unit_rule = torch.ones(1, 1, 1, device=w.device)
            return ops.phm_linear(x, self._unit_rule, w, self.affine.bias)
        if self.type == "norm":
            return x.abs()
        return x

    def __repr__(self):
        return f'{self.__class__.__name__}(type="{self.type}", in_features={self.in_features}, phm_dim={self.phm_dim}, bias={self.bias_flag})'


# ============================================================================ encoders
class IntegerEncoder(nn.Module):
    """Sum (or concat) of per-column embeddings — reference phc/quaternion/encoder.py:9-60."""

    def __init__(self, out_dim: int, input_dims: list, combine: str = "sum") -> None:
        super().__init__()
        assert combine in ["sum", "concat"]
        self.combine, self.out_dim, self.input_dims = combine, out_dim, input_dims
        self.embeddings = nn.ModuleList([nn.Embedding(d, out_dim) for d in input_dims])
        self.reset_parameters()

    def reset_parameters(self):
        for e in self.embeddings:
            glorot_uniform(e.weight.data)

    def get_number_of_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 1:
            x = x.unsqueeze(1)
        tables = [e.weight for e in self.embeddings]
        if self.combine == "sum":
            return ops.embed_sum(x, tables, 1, list(self.input_dims))
        return torch.cat([ops.embed_sum(x[:, i], [tables[i]], 1, [self.input_dims[i]]) for i in range(x.size(1))], dim=1)


class PHMEncoder(nn.Module):
    """n independent encoders, stacked on a component axis -> [rows, n, out_dim]
    — reference phc/hypercomplex/encoder.py:7-41.  One fused kernel for all n x #cols tables."""

    def __init__(self, out_dim: int, input_dims: Union[list, int], phm_dim: int, combine: str = "sum"):
        super().__init__()
        self.input_dims, self.out_dim, self.phm_dim, self.combine = input_dims, out_dim, phm_dim, combine
        if isinstance(input_dims, list):
            self.encoders = nn.ModuleList([IntegerEncoder(out_dim, input_dims, combine) for _ in range(phm_dim)])
        elif isinstance(input_dims, int):
            self.encoders = nn.ModuleList([nn.Linear(input_dims, out_dim, bias=True) for _ in range(phm_dim)])
        else:
            print(f"Must insert datatype int or list. Data type {type(input_dims)} was inserted as `input_dims`.")
            raise ValueError
        self.reset_parameters()

    def reset_parameters(self):
        for e in self.encoders:
            e.reset_parameters()

    def fusable_params(self):
        """(is_linear, parameter list in kernel order, vocabulary sizes) for the fused encoder+aggregation kernel."""
        if isinstance(self.input_dims, list):
            return False, [emb.weight for enc in self.encoders for emb in enc.embeddings], list(self.input_dims)
        return True, [e.weight for e in self.encoders] + [e.bias for e in self.encoders], []

    def can_fuse(self, width: int) -> bool:
        if isinstance(self.input_dims, list):
            if self.combine != "sum":
                return False
            return ops.conv_fused_supported(width, self.phm_dim, False, len(self.input_dims), self.input_dims)
        return ops.conv_fused_supported(width, self.phm_dim, True, int(self.input_dims), [])

    def flat(self, x: torch.Tensor) -> torch.Tensor:
        """[rows, n*out_dim] with component c in column block c."""
        if isinstance(self.input_dims, list):
            if self.combine != "sum":
                return torch.stack([enc(x) for enc in self.encoders], dim=1).flatten(1)
            if x.dim() == 1:
                x = x.unsqueeze(1)
            tables = [emb.weight for enc in self.encoders for emb in enc.embeddings]
            return ops.embed_sum(x, tables, self.phm_dim, list(self.input_dims))
        return ops.linear_encoder(x, [e.weight for e in self.encoders], [e.bias for e in self.encoders])

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.flat(x)
        return out.view(out.size(0), self.phm_dim, -1)


class NaivePHMEncoder(nn.Module):
    """One encoder replicated over the n components — reference phc/hypercomplex/encoder.py:45-72."""

    def __init__(self, out_dim: int, input_dims: Union[list, int], phm_dim: int, combine: str = "sum"):
        super().__init__()
        self.input_dims, self.out_dim, self.phm_dim, self.combine = input_dims, out_dim, phm_dim, combine
        if isinstance(input_dims, list):
            self.encoder = IntegerEncoder(out_dim, input_dims, combine)
        elif isinstance(input_dims, int):
            self.encoder = nn.Linear(input_dims, out_dim, bias=True)
        else:
            raise ValueError
        self.reset_parameters()

    def reset_parameters(self):
        self.encoder.reset_parameters()

    def flat(self, x):
        return self.forward(x).flatten(1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if isinstance(self.input_dims, list):
            e = self.encoder(x)
        else:
            e = ops.linear_encoder(x, [self.encoder.weight], [self.encoder.bias])
        return e.unsqueeze(1).expand(-1, self.phm_dim, -1).contiguous()


# ============================================================================ pooling
class PHMGlobalSumPooling(nn.Module):
    """global_add_pool over the batch vector — reference phc/hypercomplex/pooling.py:10-25."""

    def __init__(self, phm_dim: int):
        super().__init__()
        self.phm_dim = phm_dim

    def forward(self, x: torch.Tensor, batch: torch.Tensor, num_graphs: Optional[int] = None) -> torch.Tensor:
        seg = segment_structure(batch, num_graphs)
        return ops.segment_pool(x, batch, seg, self.phm_dim)

    def reset_parameters(self):
        pass

    def __repr__(self):
        return f"{self.__class__.__name__}(phm_dim={self.phm_dim})"


class PHMSoftAttentionPooling(nn.Module):
    """sigmoid-gated sum pooling — reference phc/hypercomplex/pooling.py:29-77:
    g = sigmoid(real_trafo(linear(x))), out[b] = sum_{i in b} g[i] * x[i,c,:] for every component c."""

    def __init__(self, embed_dim: int, phm_dim: int, phm_rule, learn_phm: bool = True, bias: bool = True,
                 w_init: str = "phm", c_init: str = "standard", real_trafo: str = "linear"):
        super().__init__()
        self.embed_dim, self.phm_dim, self.w_init, self.c_init = embed_dim, phm_dim, w_init, c_init
        self.phm_rule, self.learn_phm, self.real_trafo_type, self.bias = phm_rule, learn_phm, real_trafo, bias
        self.linear = PHMLinear(embed_dim, embed_dim, phm_dim, phm_rule, bias, w_init, c_init, learn_phm)
        self.real_trafo = RealTransformer(type=real_trafo, phm_dim=phm_dim, in_features=embed_dim, bias=True)
        self.sum_pooling = PHMGlobalSumPooling(phm_dim=phm_dim)
        self.reset_parameters()

    def reset_parameters(self):
        self.real_trafo.reset_parameters()
        self.linear.reset_parameters()

    def forward(self, x: torch.Tensor, batch: torch.Tensor, num_graphs: Optional[int] = None) -> torch.Tensor:
        logits = self.real_trafo(self.linear(x))
        assert logits.size(-1) == self.embed_dim // self.phm_dim, \
            "gate width must be embed_dim/phm_dim (only real_trafo='linear' produces it, as in the reference)"
        seg = segment_structure(batch, num_graphs)
        return ops.segment_pool(x, batch, seg, self.phm_dim, gate_logits=logits)

    def __repr__(self):
        return (f"{self.__class__.__name__}(embed_dim={self.embed_dim}, phm_dim={self.phm_dim}, learn_phm={self.learn_phm}, "
                f"bias={self.bias}, w_init='{self.w_init}', real_trafo='{self.real_trafo_type}')")


# ============================================================================ downstream head
class PHMDownstreamNet(nn.Module):
    """[PHMLinear -> PHMNorm -> act -> dropout] x len(hidden), PHMLinear, RealTransformer
    — reference phc/hypercomplex/downstream.py:19-130."""

    def __init__(self, in_features: int, phm_dim: int, phm_rule, hidden_layers: list, out_features: int, activation: str,
                 bias: bool, norm: str, w_init: str, c_init: str, dropout: Union[float, list], learn_phm: bool = True,
                 same_dropout: bool = False, real_trafo: str = "linear") -> None:
        super().__init__()
        self.in_features, self.out_features, self.learn_phm = in_features, out_features, learn_phm
        self.phm_rule, self.phm_dim, self.hidden_layers = phm_rule, phm_dim, hidden_layers
        self.activation_str, self.w_init, self.c_init, self.bias = activation, w_init, c_init, bias
        self.dropout = [dropout] * len(hidden_layers) if isinstance(dropout, float) else dropout
        assert len(self.dropout) == len(self.hidden_layers), "dropout list must be of the same size as number of hidden layer"
        self.norm_type, self.same_dropout, self.real_trafo_type = norm, same_dropout, real_trafo
        dims = [in_features] + list(hidden_layers) + [phm_dim * out_features]
        self.affine = nn.ModuleList([PHMLinear(dims[i], dims[i + 1], phm_dim, phm_rule, bias, w_init, c_init, learn_phm)
                                     for i in range(len(dims) - 1)])
        self.real_trafo = RealTransformer(type=real_trafo, in_features=phm_dim * out_features, phm_dim=phm_dim, bias=True)
        self.norm_flag = bool(_norm_on(norm))
        if self.norm_flag:
            self.norm = nn.ModuleList([PHMNorm(num_features=d, phm_dim=phm_dim, type=norm) for d in hidden_layers])
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.affine:
            m.reset_parameters()
        if self.norm_flag:
            for m in self.norm:
                m.reset_parameters()
        self.real_trafo.reset_parameters()

    def forward(self, x: torch.Tensor, verbose=False, **kwargs) -> torch.Tensor:
        act = self.activation_str.lower()
        last = len(self.affine) - 1
        for i, lin in enumerate(self.affine):
            x = lin(x)
            if i < last:
                x = norm_act_drop_skip(self.norm[i] if self.norm_flag else None, x, None, act, self.phm_dim, self.training,
                                       drop_p=self.dropout[i], drop_same=self.same_dropout)
        return self.real_trafo(x)


# ============================================================================ message passing
class _PHMConvBase(nn.Module):
    """Shared machinery of the four conv operators: fused gather + edge add + message activation +
    reduction (+ GIN self term) in one kernel (csrc/aggregate.cu)."""

    def _setup(self, in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, w_init, c_init, aggr,
               msg_encoder):
        self.in_features, self.out_features, self.phm_dim = in_features, out_features, phm_dim
        self.phm_rule, self.learn_phm, self.bias = phm_rule, learn_phm, bias
        self.add_self_loops, self.w_init, self.c_init, self.aggr = add_self_loops, w_init, c_init, aggr
        self.msg_encoder_str = msg_encoder
        assert msg_encoder.lower() in _ACTIVATIONS
        assert aggr in ops.REDUCE_IDS, f"unknown aggregation '{aggr}'"

    def propagate(self, x, edge_index, edge_attr, fuse_self: bool):
        assert x.size(-1) == edge_attr.size(-1)                     # reference messagepassing.py:73/145
        struct = edge_structure(edge_index, x.size(0))
        beta = getattr(self, "beta", None)
        return ops.aggregate(x, edge_attr, struct, self.aggr, self.msg_encoder_str.lower(), beta, self_loop=fuse_self)

    def propagate_fused(self, x, edge_index, raw_edge_attr, encoder, fuse_self: bool):
        """Same as ``propagate`` but takes the RAW edge features and the layer's bond encoder: the [E,F] edge
        embedding is rebuilt on the fly inside the aggregation kernel (csrc/conv_fused.cu)."""
        struct = edge_structure(edge_index, x.size(0))
        beta = getattr(self, "beta", None)
        linear, params, vocab = encoder.fusable_params()
        return ops.conv_aggregate_fused(x, raw_edge_attr, struct, linear=linear, params=params, phm_dim=self.phm_dim, vocab=vocab,
                                        reduce=self.aggr, msg_act=self.msg_encoder_str.lower(), beta=beta, self_loop=fuse_self)

    def layer_forward(self, x, skip, edge_index, raw_edge_attr, encoder, outer_norm, act: str, training: bool, drop_p: float,
                      drop_same: bool):
        """conv + outer norm/act/dropout/skip of models.py:200-217 as ONE autograd node (layer.py)."""
        from .layer import conv_layer
        struct = edge_structure(edge_index, x.size(0))
        linear, params, vocab = encoder.fusable_params()
        mlp = isinstance(self.transform, PHMMLP)
        if mlp:
            lin1, lin2 = self.transform.linear1, self.transform.linear2
            norm1 = self.transform.norm.bn if self.transform.norm_flag else None
            act1 = self.transform.activation_str.lower()
        else:
            lin1, lin2, norm1, act1 = self.transform, None, None, "identity"
        return conv_layer(x, skip, raw_edge_attr, struct, phm_dim=self.phm_dim, enc_linear=linear, enc_params=params, enc_vocab=vocab,
                          reduce=self.aggr, msg_act=self.msg_encoder_str.lower(), beta=getattr(self, "beta", None),
                          add_self_loops=self.add_self_loops, mlp=mlp, lin1=lin1, lin2=lin2, norm1=norm1,
                          norm2=outer_norm.bn if outer_norm is not None else None, act1=act1, act2=act, training=training,
                          drop_p=drop_p, drop_same=drop_same)

    def can_fuse_layer(self, outer_norm) -> bool:
        if not isinstance(self.transform, PHMMLP) and not getattr(self, "same_dim", True):
            return False
        norms = [outer_norm.bn] if outer_norm is not None else []
        if isinstance(self.transform, PHMMLP) and self.transform.norm_flag:
            norms.append(self.transform.norm.bn)
        # one training flag drives the whole fused layer: a norm switched to another mode by hand takes the unfused path
        return all(nm.track_running_stats and nm.training == self.training for nm in norms)

    def _reset_beta(self):
        if getattr(self, "beta", None) is not None:
            self.beta.data.fill_(self.initial_beta)


class PHMConv(_PHMConvBase):
    """aggregate -> PHMLinear, residual self term — reference messagepassing.py:19-88.
    same_dim=True: PHMLinear(agg) + x ; same_dim=False: PHMLinear(agg + x)."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int, phm_rule, learn_phm: bool = True, bias: bool = True,
                 add_self_loops: bool = True, w_init: str = "phm", c_init: str = "standard", aggr: str = "add",
                 same_dim: bool = True, msg_encoder: str = "identity", **kwargs) -> None:
        super().__init__()
        self._setup(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, w_init, c_init, aggr, msg_encoder)
        self.same_dim = same_dim
        self.transform = PHMLinear(in_features, out_features, phm_dim, phm_rule, bias, w_init, c_init, learn_phm)
        if aggr == "softmax":
            self.initial_beta, self.learn_beta = kwargs.get("initial_beta"), kwargs.get("learn_beta")
            self.beta = nn.Parameter(torch.tensor(float(self.initial_beta)), requires_grad=bool(self.learn_beta))
        self.reset_parameters()

    def reset_parameters(self):
        self.transform.reset_parameters()
        self._reset_beta()

    def forward(self, x, edge_index, edge_attr, size=None, encoder=None):
        """``encoder`` given: ``edge_attr`` holds the RAW edge features and the encoder is fused into the aggregation."""
        prop = self.propagate if encoder is None else (lambda a, b, c, fuse_self: self.propagate_fused(a, b, c, encoder, fuse_self))
        if self.same_dim:
            agg = prop(x, edge_index, edge_attr, fuse_self=False)
            return self.transform(agg, residual=x if self.add_self_loops else None)
        agg = prop(x, edge_index, edge_attr, fuse_self=self.add_self_loops)
        return self.transform(agg)


class PHMGINEConv(_PHMConvBase):
    """aggregate + x -> 2-layer PHM MLP — reference messagepassing.py:91-161."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int, phm_rule, learn_phm: bool = True, bias: bool = True,
                 add_self_loops: bool = True, norm=None, activation: str = "relu", w_init: str = "phm", c_init: str = "standard",
                 aggr: str = "add", msg_encoder: str = "identity", **kwargs) -> None:
        super().__init__()
        self._setup(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, w_init, c_init, aggr, msg_encoder)
        self.norm, self.activation_str = norm, activation
        # The reference does not hand ``learn_phm`` to the MLP (messagepassing.py:119-122, :276-278), so the rules of the two
        # MLP layers stay trainable even under learn_phm=False (its parameter count and gradients say so: fixture
        # phm_zinc_n2_nobias_noloops_mlp_frozen).  Reproduced; the quaternion subclasses freeze every rule themselves.
        self.transform = PHMMLP(in_features, out_features, phm_dim, phm_rule, bias=bias, learn_phm=True, activation=activation,
                                norm=norm, w_init=w_init, c_init=c_init, factor=1)
        if aggr == "softmax":
            self.initial_beta, self.learn_beta = kwargs.get("initial_beta"), kwargs.get("learn_beta")
            self.beta = nn.Parameter(torch.tensor(float(self.initial_beta)), requires_grad=bool(self.learn_beta))
        self.reset_parameters()

    def reset_parameters(self):
        self.transform.reset_parameters()
        self._reset_beta()

    def forward(self, x, edge_index, edge_attr, size=None, encoder=None):
        if encoder is not None:
            return self.transform(self.propagate_fused(x, edge_index, edge_attr, encoder, fuse_self=self.add_self_loops))
        return self.transform(self.propagate(x, edge_index, edge_attr, fuse_self=self.add_self_loops))


class PHMConvSoftmax(PHMConv):
    """reference messagepassing.py:164-245 (softmax aggregation with learnable inverse temperature)."""

    def __init__(self, in_features, out_features, phm_dim, phm_rule, learn_phm=True, bias=True, add_self_loops=True,
                 w_init="phm", c_init="standard", aggr="softmax", same_dim=True, msg_encoder="identity", **kwargs):
        super().__init__(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, w_init, c_init,
                         "softmax", same_dim, msg_encoder, **kwargs)


class PHMGINEConvSoftmax(PHMGINEConv):
    """reference messagepassing.py:248-327."""

    def __init__(self, in_features, out_features, phm_dim, phm_rule, learn_phm=True, bias=True, add_self_loops=True,
                 norm=None, activation="relu", w_init="phm", c_init="standard", aggr="softmax", msg_encoder="identity", **kwargs):
        super().__init__(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, norm, activation,
                         w_init, c_init, "softmax", msg_encoder, **kwargs)


class PHMPNAConvSimple(nn.Module):
    """Principal-neighbourhood aggregation conv — reference messagepassing.py:339-453: relu(x_j + e) messages,
    aggregators x degree scalers concatenated component-wise (one kernel, csrc/pna.cu), then
    ``transform`` = PHMLinear(S*T*F -> F) [-> PHMNorm -> act -> PHMLinear(F -> F)] * (post_layers - 1).  No self term.
    As in the reference the inner PHMLinears always learn their rule (``learn_phm`` is not forwarded, :386-399)."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int, phm_rule, learn_phm: bool, bias: bool, activation: str,
                 norm: Optional[str], w_init: str, c_init: str, deg: torch.Tensor,
                 aggregators: List[str] = ["mean", "min", "max", "std"],
                 scalers: List[str] = ["identity", "amplification", "attenuation"], post_layers: int = 1,
                 msg_encoder: str = "relu", **kwargs):
        super().__init__()
        self.in_features, self.out_features, self.bias_flag, self.activation_str = in_features, out_features, bias, activation
        self.norm, self.phm_dim, self.phm_rule, self.w_init, self.c_init, self.learn_phm = norm, phm_dim, phm_rule, w_init, c_init, learn_phm
        self.aggregators_l, self.scalers_l = list(aggregators), list(scalers)
        for a in self.aggregators_l:
            ops.PNA_AGGREGATOR_IDS[a]                      # KeyError on unknown names, like AGGREGATORS[aggr]
        for sc in self.scalers_l:
            ops.PNA_SCALER_IDS[sc]
        self.F_in, self.F_out = in_features, out_features
        self.deg = deg.to(torch.float)
        self.avg_deg = {"lin": self.deg.mean().item(), "log": (self.deg + 1).log().mean().item(),
                        "exp": self.deg.exp().mean().item()}
        wide = len(self.aggregators_l) * len(self.scalers_l) * in_features
        modules = [PHMLinear(in_features=wide, out_features=out_features, bias=bias, phm_dim=phm_dim, phm_rule=phm_rule,
                             w_init=w_init, c_init=c_init)]
        self.post_layers = post_layers
        for _ in range(post_layers - 1):
            if self.norm:
                modules += [PHMNorm(num_features=out_features, phm_dim=phm_dim, type="naive-batch-norm")]
            modules += [get_module_activation(activation)]
            modules += [PHMLinear(in_features=out_features, out_features=out_features, bias=bias, phm_dim=phm_dim,
                                  phm_rule=phm_rule, w_init=w_init, c_init=c_init)]
        self.transform = nn.Sequential(*modules)
        self.msg_encoder_str = msg_encoder
        assert msg_encoder.lower() in _ACTIVATIONS
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.transform:
            if hasattr(m, "reset_parameters"):
                m.reset_parameters()

    def propagate(self, x, edge_index, edge_attr):
        struct = edge_structure(edge_index, x.size(0))
        if edge_attr is None:                                # message() returns x_j unchanged (messagepassing.py:422)
            edge_attr, act = torch.zeros((edge_index.size(1), x.size(1)), dtype=x.dtype, device=x.device), "identity"
        else:
            act = self.msg_encoder_str.lower()
        return ops.pna_aggregate(x, edge_attr, struct, self.phm_dim, self.aggregators_l, self.scalers_l,
                                 self.avg_deg["log"], self.avg_deg["lin"], msg_act=act)

    def forward(self, x, edge_index, edge_attr=None, size=None):
        return self.transform(self.propagate(x, edge_index, edge_attr))

    def can_fuse_layer(self, outer_norm) -> bool:
        return False

    def __repr__(self):
        return (f"{self.__class__.__name__}(in_features={self.in_features}, out_features={self.out_features}, phm_dim={self.phm_dim}, "
                f"aggregators={self.aggregators_l}, scalers={self.scalers_l}, post_layers={self.post_layers})")


class PHMMessagePassing(nn.Module):
    """Dispatcher over the conv operators — reference messagepassing.py:456-518."""

    def __init__(self, in_features: int, out_features: int, phm_dim: int, phm_rule, learn_phm: bool = True, bias: bool = True,
                 add_self_loops: bool = True, norm=None, activation: str = "relu", w_init: str = "phm", c_init: str = "standard",
                 aggr: str = "add", mlp: bool = True, same_dim: bool = True, msg_encoder: str = "identity", **kwargs):
        super().__init__()
        self.in_features, self.out_features, self.phm_dim, self.bias = in_features, out_features, phm_dim, bias
        self.add_self_loops, self.norm, self.learn_phm, self.phm_rule = add_self_loops, norm, learn_phm, phm_rule
        self.activation_str, self.w_init, self.c_init, self.aggr = activation, w_init, c_init, aggr
        self.mlp, self.same_dim, self.msg_encoder_str = mlp, same_dim, msg_encoder
        if aggr == "pna":
            self.transform = PHMPNAConvSimple(in_features=in_features, out_features=out_features, phm_dim=phm_dim, phm_rule=phm_rule,
                                              learn_phm=learn_phm, bias=bias, activation=activation, norm=norm, w_init=w_init,
                                              c_init=c_init, deg=kwargs.get("deg"), aggregators=kwargs.get("aggregators"),
                                              scalers=kwargs.get("scalers"), post_layers=kwargs.get("post_layers"),
                                              msg_encoder="relu")
        elif aggr == "softmax":
            if mlp:
                self.transform = PHMGINEConvSoftmax(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops,
                                                    norm, activation, w_init, c_init, aggr, msg_encoder, **kwargs)
            else:
                self.transform = PHMConvSoftmax(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops,
                                                w_init, c_init, aggr, same_dim, msg_encoder, **kwargs)
        elif mlp:
            self.transform = PHMGINEConv(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, norm,
                                         activation, w_init, c_init, aggr, msg_encoder)
        else:
            self.transform = PHMConv(in_features, out_features, phm_dim, phm_rule, learn_phm, bias, add_self_loops, w_init, c_init,
                                     aggr, same_dim, msg_encoder)
        self.reset_parameters()

    def reset_parameters(self):
        self.transform.reset_parameters()

    def get_num_params(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def forward(self, x, edge_index, edge_attr, size=None, encoder=None):
        if encoder is not None:
            return self.transform(x, edge_index, edge_attr, size, encoder=encoder)
        return self.transform(x, edge_index, edge_attr, size)


# ============================================================================ models
class _PHMSkipConnectBase(nn.Module):
    def _build(self, concat: bool, phm_dim, learn_phm, phm_rule, atom_input_dims, atom_encoded_dim, bond_input_dims,
               naive_encoder, w_init, c_init, same_dropout, mp_layers, bias, dropout_mpnn, norm_mp, add_self_loops, msg_aggr,
               node_aggr, mlp, pooling, activation, real_trafo, downstream_layers, target_dim, dropout_dn, norm_dn, msg_encoder,
               sc_type, kwargs):
        if not concat:
            assert all(x == atom_encoded_dim == mp_layers[0] for x in mp_layers), "dimensionalities need to match for model"
        assert activation.lower() in ["relu", "lrelu", "elu", "selu", "swish"]
        assert len(dropout_mpnn) == len(mp_layers)
        assert pooling in ["globalsum", "softattention"], f"pooling variable '{pooling}' wrong."
        assert norm_mp in [None, "naive-batch-norm", "None", "naive-naive-batch-norm"]
        assert norm_dn in [None, "naive-batch-norm", "None", "naive-naive-batch-norm"]
        assert w_init in ["phm", "glorot_uniform", "glorot_normal"], f"w_init variable '{w_init}' wrong."
        assert c_init in ["standard", "random"], f"c_init variable '{c_init}' wrong."
        if msg_aggr == "sum":
            msg_aggr = "add"
        self.msg_encoder_str, self.phm_rule = msg_encoder, phm_rule
        self.variable_phm = phm_rule is None
        self.phm_dim, self.learn_phm = phm_dim, learn_phm
        self._n = phm_dim                      # the quaternion subclasses drop the public ``phm_dim`` attribute (quaternion.py)
        self.atom_input_dims, self.bond_input_dims = atom_input_dims, bond_input_dims
        self.atom_encoded_dim = atom_encoded_dim // phm_dim
        self.naive_encoder, self.w_init, self.c_init, self.same_dropout = naive_encoder, w_init, c_init, same_dropout
        self.mp_layers, self.bias, self.dropout_mpnn = mp_layers, bias, dropout_mpnn
        self.norm_mp = None if norm_mp == "None" else norm_mp
        self.add_self_loops, self.msg_aggr_type, self.node_aggr_type, self.mlp_mp = add_self_loops, msg_aggr, node_aggr, mlp
        self.pooling_type, self.activation_str, self.real_trafo_type = pooling, activation, real_trafo
        self.downstream_layers, self.target_dim, self.dropout_dn = downstream_layers, target_dim, dropout_dn
        self.norm_dn_type = None if norm_dn == "None" else norm_dn
        self.input_dim = atom_encoded_dim
        self.fuse_edge_encoder = True        # B200 path: rebuild edge embeddings inside the aggregation kernel
        self.fuse_layer = True               # ... and run each message-passing layer as one autograd node
        self.f_act = get_module_activation(activation)
        self.sc_type = sc_type
        Enc = NaivePHMEncoder if naive_encoder else PHMEncoder
        self.atomencoder = Enc(out_dim=self.atom_encoded_dim, input_dims=atom_input_dims, phm_dim=phm_dim, combine="sum")
        convs, norms, bond = [], [], []
        width = self.input_dim
        for i, out_dim in enumerate(mp_layers):
            if concat:
                in_dim = width
            else:
                in_dim = self.input_dim if i == 0 else mp_layers[i - 1]
            bond.append(Enc(out_dim=in_dim // phm_dim, input_dims=bond_input_dims, phm_dim=phm_dim, combine="sum"))
            convs.append(PHMMessagePassing(in_features=in_dim, out_features=out_dim, bias=bias, phm_dim=phm_dim, learn_phm=learn_phm,
                                           phm_rule=phm_rule, norm=self.norm_mp, activation=activation, w_init=w_init, c_init=c_init,
                                           aggr=msg_aggr, mlp=mlp, add_self_loops=add_self_loops, same_dim=not concat,
                                           msg_encoder=msg_encoder, **kwargs))
            norms.append(PHMNorm(num_features=out_dim, phm_dim=phm_dim, type=norm_mp) if self.norm_mp else None)
            width = out_dim + (self.input_dim if concat else 0)      # concat: always with the atom embedding (reference :467,:479-481)
        self.convs, self.norms, self.bondencoders = nn.ModuleList(convs), nn.ModuleList(norms), nn.ModuleList(bond)
        final = width if concat else mp_layers[-1]
        if pooling == "globalsum":
            self.pooling = PHMGlobalSumPooling(phm_dim=phm_dim)
        else:
            self.pooling = PHMSoftAttentionPooling(embed_dim=final, phm_dim=phm_dim, learn_phm=learn_phm, phm_rule=phm_rule,
                                                   w_init=w_init, c_init=c_init, bias=bias, real_trafo=real_trafo)
        self.downstream = PHMDownstreamNet(in_features=final, hidden_layers=downstream_layers, out_features=target_dim,
                                           phm_rule=phm_rule, phm_dim=phm_dim, learn_phm=learn_phm, activation=activation,
                                           bias=bias, norm=self.norm_dn_type, w_init=w_init, c_init=c_init, dropout=dropout_dn,
                                           same_dropout=same_dropout, real_trafo=real_trafo)
        self.reset_parameters()

    def reset_parameters(self):
        self.atomencoder.reset_parameters()
        for enc in self.bondencoders:
            enc.reset_parameters()
        for conv, norm in zip(self.convs, self.norms):
            conv.reset_parameters()
            if norm is not None:
                norm.reset_parameters()
        self.pooling.reset_parameters()
        self.downstream.reset_parameters()

    def get_number_of_params_(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        """Also accepts the reference's other layouts: ``PHMLinear_Old`` keys (``W.0``.., the shipped checkpoints) and, for
        n = 4, the quaternion models' ``W_r``.. keys — relabelled by legacy.py before the regular load."""
        from . import legacy
        if legacy.is_legacy_phm_state_dict(state_dict):
            state_dict = legacy.convert_legacy_phm_state_dict(state_dict)
        elif legacy.is_quaternion_state_dict(state_dict):
            state_dict = legacy.quaternion_to_phm_state_dict(state_dict)
        return super().load_state_dict(state_dict, strict=strict, **kwargs)

    def _encode_edges(self, i: int, edge_attr: torch.Tensor) -> torch.Tensor:
        return self.bondencoders[i].flat(edge_attr)

    def _report_stage(self, h: torch.Tensor, k: int) -> None:
        """Data-parallel overlap (parallel.GradientBucket.enable_overlap): when the gradient of this layer output is complete,
        every parameter of the layers above it and of the head — stages <= k of optim.staged_parameters — has its final gradient."""
        cb = self.__dict__.get("_stage_hook")
        if cb is not None and h.requires_grad:
            def _hook(_grad, k=k, cb=cb):
                cb(k)                                        # returns None: the gradient itself is left untouched
            h.register_hook(_hook)


class PHMSkipConnectAdd(_PHMSkipConnectBase):
    """Message-passing network with additive skip connections — reference
    phc/hypercomplex/undirectional/models.py:24-267."""

    def __init__(self, phm_dim: int = 4, learn_phm: bool = True, phm_rule=None, atom_input_dims: Union[int, list] = ATOM_FEAT_DIMS,
                 atom_encoded_dim: int = 196, bond_input_dims: Union[int, list] = BOND_FEAT_DIMS, naive_encoder: bool = False,
                 w_init: str = "phm", c_init: str = "standard", same_dropout: bool = False, mp_layers: list = [196, 196, 196],
                 bias: bool = True, dropout_mpnn: list = [0.0, 0.0, 0.0], norm_mp: Optional[str] = "naive-batch-norm",
                 add_self_loops: bool = True, msg_aggr: str = "add", node_aggr: str = "sum", mlp: bool = False,
                 pooling: str = "softattention", activation: str = "relu", real_trafo: str = "linear",
                 downstream_layers: list = [256, 128], target_dim: int = 1, dropout_dn: Union[list, float] = [0.2, 0.1],
                 norm_dn: Optional[str] = "naive-batch-norm", msg_encoder: str = "identity", sc_type: str = "first", **kwargs) -> None:
        super().__init__()
        self._build(False, phm_dim, learn_phm, phm_rule, atom_input_dims, atom_encoded_dim, bond_input_dims, naive_encoder, w_init,
                    c_init, same_dropout, mp_layers, bias, dropout_mpnn, norm_mp, add_self_loops, msg_aggr, node_aggr, mlp, pooling,
                    activation, real_trafo, downstream_layers, target_dim, dropout_dn, norm_dn, msg_encoder, sc_type, kwargs)

    def compute_hidden_layer_embedding(self, conv, norm, x, edge_index, edge_attr, dropout_mpnn: float, size=None) -> torch.Tensor:
        """conv -> norm -> act -> dropout -> + skip   (reference models.py:200-217); the last four are one kernel."""
        h = conv(x=x[0], edge_index=edge_index, edge_attr=edge_attr, size=size)
        return norm_act_drop_skip(norm, h, x[1], self.activation_str.lower(), self._n, self.training,
                                  drop_p=dropout_mpnn, drop_same=self.same_dropout)

    def forward(self, data, size=None) -> torch.Tensor:
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if isinstance(self.bond_input_dims, list):
            edge_attr = edge_attr.to(torch.long)
        h0 = self.atomencoder.flat(x)
        h = h0
        L = len(self.mp_layers)
        # sc_type "first": every layer adds h0 — hand each its own alias so that the L skip gradients are summed in one pass; one more
        # alias for h0 as the INPUT of layer 0, so that its gradient joins the same pass instead of an accumulation kernel of its own
        first = ops.fan_out(h0, L + 1) if (self.sc_type == "first" and L > 1) else None
        if first is not None:
            h = first[L]
        for i in range(L):
            if first is not None:
                skip = first[i]
            elif i == 0 or self.sc_type == "first":
                skip = h0
            elif self.sc_type == "last":
                skip = h
            else:
                raise ValueError
            enc = self.bondencoders[i]
            pna = isinstance(self.convs[i].transform, PHMPNAConvSimple)       # PNA reads the materialised edge embedding
            if self.fuse_edge_encoder and not pna and isinstance(enc, PHMEncoder) and enc.can_fuse(h.size(1)):
                conv = self.convs[i].transform
                if self.fuse_layer and conv.can_fuse_layer(self.norms[i]):
                    # the whole layer (conv + norm + act + dropout + skip) as one autograd node
                    h = conv.layer_forward(h, skip, edge_index, edge_attr, enc, self.norms[i], self.activation_str.lower(),
                                           self.training, self.dropout_mpnn[i], self.same_dropout)
                    self._report_stage(h, L - 1 - i)
                    continue
                # bond encoder fused into the aggregation: the [E,F] edge embedding of models.py:238-240 is never formed
                z = self.convs[i](h, edge_index, edge_attr, size, encoder=enc)
                h = norm_act_drop_skip(self.norms[i], z, skip, self.activation_str.lower(), self._n, self.training,
                                       drop_p=self.dropout_mpnn[i], drop_same=self.same_dropout)
                self._report_stage(h, L - 1 - i)
                continue
            e = self._encode_edges(i, edge_attr)
            h = self.compute_hidden_layer_embedding(self.convs[i], self.norms[i], [h, skip], edge_index, e, self.dropout_mpnn[i], size)
            self._report_stage(h, L - 1 - i)
        num_graphs = getattr(data, "num_graphs", None)
        out = self.pooling(h, batch, num_graphs)
        return self.downstream(out)


class PHMSkipConnectConcat(_PHMSkipConnectBase):
    """Skip connections through concatenation — reference models.py:271-517.  The reference's forward
    raises for every phm_dim > 1 (models.py:486 reshapes the layer-0 bond embedding n times too wide,
    SURVEY.md D2): parity is pinned at phm_dim = 1 (fixtures ``phm_concat_n1_*``: layer widths, skip = atom embedding for
    every layer, pooling / downstream widths) and, for the quaternion subclass, by the reference's working
    QuaternionSkipConnectConcat; for phm_dim > 1 this implements the evident intent (flat concat as written at :467)."""

    def __init__(self, phm_dim: int = 4, learn_phm: bool = True, phm_rule=None, atom_input_dims: Union[int, list] = ATOM_FEAT_DIMS,
                 atom_encoded_dim: int = 196, bond_input_dims: Union[int, list] = BOND_FEAT_DIMS, naive_encoder: bool = False,
                 w_init: str = "phm", c_init: str = "standard", same_dropout: bool = False, mp_layers: list = [196, 196, 196],
                 bias: bool = True, dropout_mpnn: list = [0.0, 0.0, 0.0], norm_mp: Optional[str] = "naive-batch-norm",
                 add_self_loops: bool = True, msg_aggr: str = "add", node_aggr: str = "sum", mlp: bool = False,
                 pooling: str = "softattention", activation: str = "relu", real_trafo: str = "linear",
                 downstream_layers: list = [256, 128], target_dim: int = 1, dropout_dn: Union[list, float] = [0.2, 0.1],
                 norm_dn: Optional[str] = "naive-batch-norm", msg_encoder: str = "identity", sc_type: str = "first", **kwargs) -> None:
        super().__init__()
        self._build(True, phm_dim, learn_phm, phm_rule, atom_input_dims, atom_encoded_dim, bond_input_dims, naive_encoder, w_init,
                    c_init, same_dropout, mp_layers, bias, dropout_mpnn, norm_mp, add_self_loops, msg_aggr, node_aggr, mlp, pooling,
                    activation, real_trafo, downstream_layers, target_dim, dropout_dn, norm_dn, msg_encoder, sc_type, kwargs)

    def _skip_concat(self, z: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
        """flat concat, as the reference writes it (models.py:467); the quaternion subclass concatenates per component"""
        return torch.cat([z, skip], dim=-1)

    def forward(self, data, size=None) -> torch.Tensor:
        x, edge_index, edge_attr, batch = data.x, data.edge_index, data.edge_attr, data.batch
        if isinstance(self.bond_input_dims, list):
            edge_attr = edge_attr.to(torch.long)
        h0 = self.atomencoder.flat(x)
        h = h0
        act = self.activation_str.lower()
        for i in range(len(self.mp_layers)):
            skip = h0                                 # the reference concatenates the atom embedding whatever sc_type says (:479-481)
            enc = self.bondencoders[i]
            pna = isinstance(self.convs[i].transform, PHMPNAConvSimple)       # PNA reads the materialised edge embedding
            if self.fuse_edge_encoder and not pna and isinstance(enc, PHMEncoder) and enc.can_fuse(h.size(1)):
                # bond encoder fused into the aggregation (any layer width, F_i + F_0 included): the [E, F] edge embedding of
                # models.py:486-488 is never formed
                z = self.convs[i](h, edge_index, edge_attr, size, encoder=enc)
            else:
                z = self.convs[i](x=h, edge_index=edge_index, edge_attr=self._encode_edges(i, edge_attr), size=size)
            if type(self)._skip_concat is PHMSkipConnectConcat._skip_concat and self.fuse_edge_encoder:
                # flat concat: norm / act / dropout write the left column block of the next layer's input buffer, the atom
                # embedding is copied into the right one — no torch.cat pass (csrc/norm.cu, *_strided)
                h = norm_act_drop_skip(self.norms[i], z, None, act, self._n, self.training, drop_p=self.dropout_mpnn[i],
                                       drop_same=self.same_dropout, concat_right=skip)
                continue
            z = norm_act_drop_skip(self.norms[i], z, None, act, self._n, self.training, drop_p=self.dropout_mpnn[i],
                                   drop_same=self.same_dropout)
            h = self._skip_concat(z, skip)
        out = self.pooling(h, batch, getattr(data, "num_graphs", None))
        return self.downstream(out)


# ============================================================================ regularisers
def phm_weight_regularization(model, p: int = 2, device=None):
    """sum over modules with a ``W`` of W.norm(p, dim=0).mean() — reference regularization.py:15-23."""
    ws = getattr(model, "_phc_W_cache", None)       # the module tree is static: walk it once
    if ws is None:
        ws = [w for _, module in model.named_modules() for w in (getattr(module, "W", None),) if w is not None]
        try:
            object.__setattr__(model, "_phc_W_cache", ws)
        except Exception:
            pass
    if p == 2 and ws and all(w.is_cuda and w.dim() == 3 for w in ws) and len(ws) <= 256:
        return ops.weight_regularization_l2(ws)              # one fused kernel pair (csrc/regularizer.cu)
    reg = 0.0
    for w in ws:
        reg = reg + w.norm(p=p, dim=0).mean()
    return reg


def multiplication_rule_regularization(model, p: int = 1, device=None):
    """reference regularization.py:4-12."""
    reg = 0.0
    for _, module in model.named_modules():
        r = getattr(module, "phm_rule", None)
        if r is not None:
            reg = reg + r.norm(p=p).mean()
    return reg


def get_model_blocks(model, attr: str, **kwargs) -> list:
    """Optimizer parameter group for ``model.<attr>`` (or ``model.module.<attr>`` under a DP wrapper)
    — reference phc/quaternion/regularization.py:4-24."""
    root = getattr(model, "module", model)
    if not isinstance(root, nn.Module) or not hasattr(root, attr):
        return []
    params = list(getattr(root, attr).parameters())
    return [dict(params=params, **kwargs)] if params else []
