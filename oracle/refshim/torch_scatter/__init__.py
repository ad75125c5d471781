"""Pure-torch restatement of the torch-scatter 2.0.5 entry points the reference calls
(reference call sites: phc/hypercomplex/aggregator.py:29,51-53,70-83,
phc/hypercomplex/undirectional/messagepassing.py:4-5). Test infrastructure only."""
import torch


def _expand_index(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src), dim


def _out_shape(src, index, dim, dim_size):
    shape = list(src.shape)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape[dim] = dim_size
    return shape


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    idx, dim = _expand_index(index, src, dim)
    if out is None:
        out = src.new_zeros(_out_shape(src, index, dim, dim_size))
    return out.scatter_add(dim, idx, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    idx, dim = _expand_index(index, src, dim)
    total = scatter_sum(src, index, dim, out, dim_size)
    count = scatter_sum(torch.ones_like(src), index, dim, None, total.size(dim))
    return total / count.clamp(min=1)


def _scatter_ext(src, index, dim, dim_size, mode):
    idx, dim = _expand_index(index, src, dim)
    out = src.new_zeros(_out_shape(src, index, dim, dim_size))
    # untouched (empty) segments keep 0, as torch-scatter fills them
    return out.scatter_reduce(dim, idx, src, mode, include_self=False).clone()


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_ext(src, index, dim, dim_size, "amax"), None


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_ext(src, index, dim, dim_size, "amin"), None


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError(reduce)


from . import composite  # noqa: E402,F401
