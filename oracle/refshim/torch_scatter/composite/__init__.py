"""scatter_softmax as documented for torch-scatter 2.0.5 (eps=1e-12). Test infrastructure only."""
import torch


def scatter_softmax(src, index, dim=-1, eps=1e-12):
    from torch_scatter import scatter_max, scatter_sum, _expand_index
    idx, dim = _expand_index(index, src, dim)
    seg_max = scatter_max(src, index, dim)[0]
    centered = src - seg_max.gather(dim, idx)
    ex = centered.exp()
    denom = scatter_sum(ex, index, dim) + eps
    return ex / denom.gather(dim, idx)
