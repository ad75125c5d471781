"""Batch preparation on the device: ``RemoveIsolatedNodes`` and the mini-batch collate (``DeviceGraphStore``).

The reference's training loops apply ``torch_geometric.transforms.RemoveIsolatedNodes`` to every batch after moving it
to the GPU (benchmarks/train_hiv.py:171-173, :233-234, :457; benchmarks/utils.py:39-49 has the same transform spelled
out).  PyG runs it as a dozen index kernels plus boolean-mask gathers; here the index work is one C-ABI call
(``phc_remove_isolated_nodes``, csrc/prep.cu) and the only host synchronisation is the read-back of the three output
sizes, which PyG needs as well (``mask.sum()``).

The reference's ``DataLoader`` collates every mini-batch on the CPU (``Batch.from_data_list``: concatenate, shift each graph's
``edge_index``, build ``batch``; benchmarks/train_hiv.py:481-493) and the loop then copies it to the GPU (:173).  With 180 GB of
HBM the whole dataset fits on the device (ogbg-ppa, the largest, is ~10 GB as int64/fp32 tensors): ``DeviceGraphStore`` keeps it
packed in HBM and ``collate(ids)`` assembles a mini-batch with ONE kernel (``phc_collate_batch``, csrc/prep.cu) — the host only
sends the [B] graph ids and their size prefix sums (a few KB, one pinned copy), and there is no host synchronisation.
"""
from __future__ import annotations

import copy
from typing import Optional, Sequence, Tuple

import numpy as np
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


def plan_collate(ids, node_ptr: np.ndarray, edge_ptr: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Host half of a collate: (ids as int64, node prefix sums [B+1], edge prefix sums [B+1]) of the selected graphs — where
    every graph's nodes / edges start in the assembled batch.  Raises IndexError for an id outside the store."""
    ids_np = np.asarray(ids.cpu().numpy() if torch.is_tensor(ids) else ids, dtype=np.int64).reshape(-1)
    B = int(ids_np.shape[0])
    if B and (ids_np.min() < 0 or ids_np.max() >= len(node_ptr) - 1):
        raise IndexError("graph id outside the store")
    onp = np.zeros(B + 1, dtype=np.int64)
    oep = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(node_ptr[ids_np + 1] - node_ptr[ids_np], out=onp[1:])
    np.cumsum(edge_ptr[ids_np + 1] - edge_ptr[ids_np], out=oep[1:])
    return ids_np, onp, oep


class DeviceGraphStore(object):
    """A graph dataset packed in device memory + the PyG-compatible collate of a mini-batch from it.

    ``graphs``: sequence of objects with ``x`` [n_g, ...], ``edge_index`` int64 [2, e_g] (node ids local to the graph),
    ``edge_attr`` [e_g, ...] or None, ``y`` (one row per graph, any trailing shape) or None — e.g. PyG ``Data`` objects or
    ``synthetic.GraphBatch``es of one graph.  Row payloads must be 4- or 8-byte dtypes (int64 / float32 in the reference).
    """

    def __init__(self, graphs: Sequence, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DeviceGraphStore lives in GPU memory; there is no CPU fallback")
        assert len(graphs) > 0, "empty dataset"
        nn = np.array([int(g.x.size(0)) for g in graphs], dtype=np.int64)
        ne = np.array([int(g.edge_index.size(1)) for g in graphs], dtype=np.int64)
        self.num_graphs = len(graphs)
        self.node_ptr_host = np.concatenate([[0], np.cumsum(nn)]).astype(np.int64)
        self.edge_ptr_host = np.concatenate([[0], np.cumsum(ne)]).astype(np.int64)
        self.node_ptr = torch.from_numpy(self.node_ptr_host).to(device)
        self.edge_ptr = torch.from_numpy(self.edge_ptr_host).to(device)
        self.x = torch.cat([g.x for g in graphs], dim=0).contiguous().to(device)
        ei = torch.cat([g.edge_index for g in graphs], dim=1)
        assert ei.dtype == torch.int64 and ei.dim() == 2 and ei.size(0) == 2, "edge_index must be int64 [2, E]"
        self.edge_index = ei.contiguous().to(device)
        has_attr = getattr(graphs[0], "edge_attr", None) is not None
        self.edge_attr = torch.cat([g.edge_attr for g in graphs], dim=0).contiguous().to(device) if has_attr else None
        has_y = getattr(graphs[0], "y", None) is not None
        if has_y:
            ys = [g.y if g.y.dim() > 0 else g.y.view(1) for g in graphs]
            assert all(y.size(0) == 1 for y in ys), "y must hold one row per graph"
            self.y = torch.cat(ys, dim=0).contiguous().to(device)
        else:
            self.y = None
        for name in ("x", "edge_attr", "y"):
            t = getattr(self, name)
            if t is not None:
                assert t.element_size() in (4, 8), f"{name}: 4- or 8-byte element types only, got {t.dtype}"
                assert t.data_ptr() % 16 == 0
        self.device = device
        self._staging = None                 # pinned [ids | out_node_ptr | out_edge_ptr], reused between calls
        self._copy_done = None
        self._last_status = None

    @staticmethod
    def _row_bytes(t: Optional[torch.Tensor]) -> int:
        if t is None:
            return 0
        return int(t[0].numel()) * t.element_size() if t.size(0) > 0 else int(np.prod(t.shape[1:], dtype=np.int64)) * t.element_size()

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.edge_index, self.edge_attr, self.y, self.node_ptr, self.edge_ptr)
                   if t is not None)

    def collate(self, ids) -> "object":
        """``Batch.from_data_list([dataset[i] for i in ids])`` -> ``synthetic.GraphBatch`` on the device.  ``ids``: host
        sequence / numpy array / CPU tensor of graph indices (any order, repeats allowed)."""
        from .synthetic import GraphBatch
        ids_np, onp, oep = plan_collate(ids, self.node_ptr_host, self.edge_ptr_host)
        B = int(ids_np.shape[0])
        N, E = int(onp[-1]), int(oep[-1])
        words = 3 * B + 2
        if self._staging is None or self._staging.numel() < words:
            self._staging = torch.empty(max(words, 1024), dtype=torch.int64).pin_memory()
        if self._copy_done is not None:
            self._copy_done.synchronize()         # previous batch's ids have left the pinned buffer (normally long done)
        host = self._staging[:words]
        host[:B] = torch.from_numpy(ids_np)
        host[B:2 * B + 1] = torch.from_numpy(onp)
        host[2 * B + 1:] = torch.from_numpy(oep)
        dev = self.device
        meta = host.to(dev, non_blocking=True)
        self._copy_done = torch.cuda.Event()      # the staging buffer is reused by the next call: it waits for this copy
        self._copy_done.record(torch.cuda.current_stream(dev))
        d_ids, d_onp, d_oep = meta[:B], meta[B:2 * B + 1], meta[2 * B + 1:]
        out_ei = torch.empty((2, E), dtype=torch.int64, device=dev)
        out_batch = torch.empty(N, dtype=torch.int64, device=dev)
        out_x = torch.empty((N,) + tuple(self.x.shape[1:]), dtype=self.x.dtype, device=dev)
        out_ea = (torch.empty((E,) + tuple(self.edge_attr.shape[1:]), dtype=self.edge_attr.dtype, device=dev)
                  if self.edge_attr is not None else None)
        out_y = torch.empty((B,) + tuple(self.y.shape[1:]), dtype=self.y.dtype, device=dev) if self.y is not None else None
        status = torch.empty(1, dtype=torch.int32, device=dev)
        p = lambda t: 0 if t is None else t.data_ptr()      # noqa: E731
        run("phc_collate_batch", dev, d_ids.data_ptr(), B, self.num_graphs, self.node_ptr.data_ptr(), self.edge_ptr.data_ptr(),
            d_onp.data_ptr(), d_oep.data_ptr(), self.edge_index.data_ptr(), int(self.edge_index.size(1)), out_ei.data_ptr(), E, N,
            out_batch.data_ptr(), p(self.x), p(out_x), self._row_bytes(self.x), p(self.edge_attr), p(out_ea),
            self._row_bytes(self.edge_attr), p(self.y), p(out_y), self._row_bytes(self.y), status.data_ptr(), _stream(dev), launches=1)
        self._last_status = status               # checked lazily (check_status) so that collate itself never synchronises
        return GraphBatch(out_x, out_ei, out_ea, out_batch, out_y, B)

    def check_status(self) -> None:
        """Raise if the last collate reported an inconsistency (synchronises; for tests / debugging)."""
        s = int(self._last_status.item())
        if s & 1:
            raise IndexError("phc_collate_batch: graph id outside the store")
        if s & 2:
            raise RuntimeError("phc_collate_batch: size prefix sums do not match the selected graphs")


def balanced_partition(costs, world: int) -> list:
    """Split the items of one global batch between ``world`` ranks so that every rank gets the same NUMBER of items (+-1) and
    nearly the same total cost (node or edge count): items sorted by cost, dealt out in serpentine order.  A synchronous
    data-parallel step ends when the slowest rank does, so equalising the per-rank node counts removes the straggler wait
    without changing which graphs form the global batch.  Returns ``world`` index arrays (positions into ``costs``), each in
    ascending position order.  Pure host logic."""
    costs = np.asarray(costs)
    order = np.argsort(-costs, kind="stable")
    parts = [[] for _ in range(world)]
    for j, idx in enumerate(order):
        r = j % (2 * world)
        parts[r if r < world else 2 * world - 1 - r].append(int(idx))
    return [np.sort(np.asarray(p, dtype=np.int64)) for p in parts]


class EpochSampler(object):
    """Graph ids of one rank's mini-batches for one epoch: a seeded permutation of the dataset (``shuffle=True``, the
    scripts' training loaders, benchmarks/train_hiv.py:488-489) cut into global batches of ``world * batch_graphs`` graphs, of
    which rank r takes the r-th slice — every graph is visited once per epoch by exactly one rank, all ranks run the same
    number of steps (a last global batch that cannot give every rank a graph is dropped; a short one is split evenly).
    Pure host logic (numpy): data parallelism shards by graph, no collective is involved."""

    def __init__(self, num_graphs: int, batch_graphs: int, rank: int = 0, world: int = 1, seed: int = 0, shuffle: bool = True,
                 drop_last: bool = False, costs=None):
        """costs (optional, one number per graph, e.g. its node count): the graphs of every global batch are dealt to the ranks by
        ``balanced_partition`` instead of in contiguous slices, so that all ranks get nearly equal work."""
        assert num_graphs > 0 and batch_graphs > 0 and 0 <= rank < world
        self.num_graphs, self.batch_graphs, self.rank, self.world = int(num_graphs), int(batch_graphs), int(rank), int(world)
        self.seed, self.shuffle, self.drop_last = int(seed), bool(shuffle), bool(drop_last)
        self.costs = None if costs is None else np.asarray(costs)
        assert self.costs is None or len(self.costs) == self.num_graphs
        self.epoch = 0

    def set_epoch(self, epoch: int) -> None:
        self.epoch = int(epoch)

    def _global_batches(self):
        order = (np.random.default_rng([self.seed, self.epoch]).permutation(self.num_graphs) if self.shuffle
                 else np.arange(self.num_graphs))
        step = self.world * self.batch_graphs
        for lo in range(0, self.num_graphs, step):
            chunk = order[lo:lo + step]
            if len(chunk) < step and (self.drop_last or len(chunk) < self.world):
                return
            yield chunk

    def __len__(self) -> int:
        return sum(1 for _ in self._global_batches())

    def __iter__(self):
        for chunk in self._global_batches():
            if self.costs is not None and self.world > 1:
                yield np.ascontiguousarray(chunk[balanced_partition(self.costs[chunk], self.world)[self.rank]]).astype(np.int64)
                continue
            per = len(chunk) // self.world                   # == batch_graphs except for a short last batch
            extra = len(chunk) - per * self.world            # the first ``extra`` ranks take one more graph
            lo = self.rank * per + min(self.rank, extra)
            yield np.ascontiguousarray(chunk[lo:lo + per + (1 if self.rank < extra else 0)]).astype(np.int64)


class DeviceLoader(object):
    """Stands in for ``DataLoader(dataset, batch_size, shuffle)`` + ``data.to(device)`` + the per-batch transform of the
    training loops (benchmarks/train_hiv.py:171-173): iterates device-resident mini-batches assembled by
    ``DeviceGraphStore.collate``; ``transform`` is e.g. ``prep.RemoveIsolatedNodes()``."""

    def __init__(self, store: DeviceGraphStore, sampler: EpochSampler, transform=None):
        assert sampler.num_graphs == store.num_graphs
        self.store, self.sampler, self.transform = store, sampler, transform

    def __len__(self) -> int:
        return len(self.sampler)

    def __iter__(self):
        for ids in self.sampler:
            batch = self.store.collate(ids)
            yield self.transform(batch) if self.transform is not None else batch
