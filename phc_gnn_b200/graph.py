"""Per-batch graph structure (CSR by target, CSC by source, graph segment offsets) on the device.

Built once per mini-batch by ``phc_csr_build`` / ``phc_segment_ptr_build`` and shared by all
message-passing layers and by forward and backward.  The cache is keyed on the identity of the
``edge_index`` / ``batch`` tensor objects (never on raw pointers, which the allocator recycles).
"""
from __future__ import annotations

import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib


def _stream(device) -> int:
    """Raw cudaStream_t of torch's current stream (the fast private getter; ~30x cheaper than
    torch.cuda.current_stream().cuda_stream, which matters for launch-bound molecule-sized batches)."""
    idx = device.index if device is not None and device.index is not None else torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(idx)


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f"phc_gnn_b200: {what} must live on a CUDA device (got {t.device}); "
                           "this package has no CPU path — the CPU restatement lives in oracle/ for tests only")


class EdgeStructure(object):
    __slots__ = ("num_nodes", "num_edges", "rowptr", "col", "perm", "rowptr_t", "col_t", "perm_t", "status", "extras", "__weakref__")

    def __init__(self, edge_index: torch.Tensor, num_nodes: int):
        require_cuda(edge_index, "edge_index")
        assert edge_index.dim() == 2 and edge_index.size(0) == 2 and edge_index.dtype == torch.int64, \
            "edge_index must be an int64 tensor of shape [2, E]"
        ei = edge_index.contiguous()
        dev = ei.device
        E, N = ei.size(1), int(num_nodes)
        lib = _lib.load()
        i32 = dict(dtype=torch.int32, device=dev)
        buf = torch.empty(2 * (N + 1) + 4 * E + 1, **i32)
        o = 0
        self.rowptr = buf[o:o + N + 1]; o += N + 1
        self.rowptr_t = buf[o:o + N + 1]; o += N + 1
        self.col = buf[o:o + E]; o += E
        self.perm = buf[o:o + E]; o += E
        self.col_t = buf[o:o + E]; o += E
        self.perm_t = buf[o:o + E]; o += E
        self.status = buf[o:o + 1]
        ws_bytes = lib.phc_csr_workspace_bytes(N, E)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        from .ops import run
        run("phc_csr_build", dev, ei.data_ptr(), E, N, self.rowptr.data_ptr(), self.col.data_ptr(), self.perm.data_ptr(),
            self.rowptr_t.data_ptr(), self.col_t.data_ptr(), self.perm_t.data_ptr(), ws.data_ptr(), ws_bytes,
            self.status.data_ptr(), _stream(dev))
        self.num_nodes, self.num_edges = N, E
        self.extras = {}            # per-batch derived data shared by all layers (e.g. per-node edge-feature sums)

    def validate(self):
        """Synchronising check of the device status word (debug / tests)."""
        s = int(self.status.item())
        if s & 1:
            raise IndexError("edge_index contains node ids outside [0, num_nodes)")


class SegmentStructure(object):
    __slots__ = ("num_nodes", "num_graphs", "graph_ptr", "status", "__weakref__")

    def __init__(self, batch: torch.Tensor, num_graphs: Optional[int] = None):
        require_cuda(batch, "batch")
        assert batch.dim() == 1 and batch.dtype == torch.int64, "batch must be an int64 vector"
        b = batch.contiguous()
        if num_graphs is None:
            # same host sync torch_geometric's global_add_pool performs (size = batch.max()+1)
            num_graphs = int(b.max().item()) + 1 if b.numel() > 0 else 0
        dev = b.device
        lib = _lib.load()
        buf = torch.empty(num_graphs + 2, dtype=torch.int32, device=dev)
        self.graph_ptr = buf[:num_graphs + 1]
        self.status = buf[num_graphs + 1:]
        from .ops import run
        run("phc_segment_ptr_build", dev, b.data_ptr(), b.numel(), num_graphs, self.graph_ptr.data_ptr(),
            self.status.data_ptr(), _stream(dev))
        self.num_nodes, self.num_graphs = b.numel(), int(num_graphs)

    def validate(self):
        s = int(self.status.item())
        if s & 2:
            raise ValueError("batch vector is not ascending (PyG collate order is required)")
        if s & 1:
            raise IndexError("batch vector contains graph ids outside [0, num_graphs)")


# id(tensor) -> (weakref(tensor), version, key, structure)
_CACHE: Dict[int, Tuple[weakref.ref, int, tuple, object]] = {}


def _lookup(t: torch.Tensor, key: tuple):
    hit = _CACHE.get(id(t))
    if hit is not None and hit[0]() is t and hit[1] == t._version and hit[2] == key:
        return hit[3]
    return None


def _store(t: torch.Tensor, key: tuple, s):
    tid = id(t)

    def _evict(_ref, tid=tid):
        _CACHE.pop(tid, None)

    _CACHE[tid] = (weakref.ref(t, _evict), t._version, key, s)
    return s


def edge_structure(edge_index: torch.Tensor, num_nodes: int) -> EdgeStructure:
    key = ("edge", int(num_nodes))
    s = _lookup(edge_index, key)
    return s if s is not None else _store(edge_index, key, EdgeStructure(edge_index, num_nodes))


def segment_structure(batch: torch.Tensor, num_graphs: Optional[int] = None) -> SegmentStructure:
    key = ("seg", num_graphs)
    s = _lookup(batch, key)
    return s if s is not None else _store(batch, key, SegmentStructure(batch, num_graphs))


def clear_cache():
    _CACHE.clear()
