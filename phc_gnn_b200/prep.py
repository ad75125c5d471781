"""Batch preparation on the device: ``RemoveIsolatedNodes``.

The reference's training loops apply ``torch_geometric.transforms.RemoveIsolatedNodes`` to every batch after moving it
to the GPU (benchmarks/train_hiv.py:171-173, :233-234, :457; benchmarks/utils.py:39-49 has the same transform spelled
out).  PyG runs it as a dozen index kernels plus boolean-mask gathers; here the index work is one C-ABI call
(``phc_remove_isolated_nodes``, csrc/prep.cu) and the only host synchronisation is the read-back of the three output
sizes, which PyG needs as well (``mask.sum()``).
"""
from __future__ import annotations

import copy
from typing import Optional, Tuple

import torch

from . import _lib
from .graph import _stream, require_cuda
from .ops import run


def remove_isolated_nodes(edge_index: torch.Tensor, edge_attr: Optional[torch.Tensor] = None, num_nodes: Optional[int] = None
                          ) -> Tuple[torch.Tensor, Optional[torch.Tensor], torch.Tensor]:
    """Same contract as ``torch_geometric.utils.remove_isolated_nodes`` (1.6.1): -> (edge_index, edge_attr, mask)."""
    require_cuda(edge_index, "edge_index")
    assert edge_index.dim() == 2 and edge_index.size(0) == 2 and edge_index.dtype == torch.int64, \
        "edge_index must be an int64 tensor of shape [2, E]"
    ei = edge_index.contiguous()
    dev = ei.device
    E = ei.size(1)
    N = int(num_nodes) if num_nodes is not None else (int(ei.max()) + 1 if E > 0 else 0)
    lib = _lib.load()
    mask = torch.empty(N, dtype=torch.bool, device=dev)
    assoc = torch.empty(N, dtype=torch.int64, device=dev)
    out = torch.empty((2, E), dtype=torch.int64, device=dev)
    order = torch.empty(E, dtype=torch.int64, device=dev)
    counts = torch.empty(4, dtype=torch.int32, device=dev)
    nb = lib.phc_isolated_workspace_bytes(N, E)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    run("phc_remove_isolated_nodes", dev, ei.data_ptr(), E, N, mask.data_ptr(), assoc.data_ptr(), out.data_ptr(), order.data_ptr(),
        counts.data_ptr(), ws.data_ptr(), nb, _stream(dev), launches=8)
    kept_nodes, kept_edges, kept_loops, status = (int(v) for v in counts.tolist())       # the one host sync
    if status:
        raise IndexError("edge_index refers to nodes outside [0, num_nodes)")
    m = kept_edges + kept_loops
    new_ei = out[:, :m]
    if m != E:
        new_ei = new_ei.contiguous()
    new_attr = None
    if edge_attr is not None:
        new_attr = edge_attr if (m == E and kept_loops == 0) else edge_attr.index_select(0, order[:m])
    return new_ei, new_attr, mask


class RemoveIsolatedNodes(object):
    """Drop-in for ``torch_geometric.transforms.RemoveIsolatedNodes`` / ``benchmarks.utils.CustomRemoveIsolatedNodes`` on a
    CUDA batch: node-level tensors (first dimension == num_nodes, key without "edge") are filtered by the keep mask."""

    def __call__(self, data):
        num_nodes = data.num_nodes
        ei, ea, mask = remove_isolated_nodes(data.edge_index, getattr(data, "edge_attr", None), num_nodes)
        out = copy.copy(data)
        out.edge_index, out.edge_attr = ei, ea
        if bool(mask.all()) if mask.numel() else True:
            return out
        for key, item in list(vars(data).items()):
            if torch.is_tensor(item) and item.dim() > 0 and item.size(0) == num_nodes and "edge" not in key:
                setattr(out, key, item[mask])
        return out

    def __repr__(self):
        return f"{self.__class__.__name__}()"
