import torch


def degree(index, num_nodes=None, dtype=None):
    n = int(index.max()) + 1 if num_nodes is None else num_nodes
    out = torch.zeros(n, dtype=dtype, device=index.device)
    return out.scatter_add_(0, index, out.new_ones(index.size(0)))


def remove_isolated_nodes(edge_index, edge_attr=None, num_nodes=None):
    n = int(edge_index.max()) + 1 if num_nodes is None else num_nodes
    mask = torch.zeros(n, dtype=torch.bool)
    mask[edge_index.view(-1)] = True
    assoc = torch.full((n,), -1, dtype=torch.long)
    assoc[mask] = torch.arange(int(mask.sum()))
    return assoc[edge_index], edge_attr, mask
