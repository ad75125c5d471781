"""Synthetic graph mini-batches shaped like the reference's five benchmark datasets.

There is no network, so the OGB / Benchmarking-GNNs datasets the reference trains on
(reference `benchmarks/train_{hiv,zinc,pcba,mnist,cifar10,ppa}.py`) are replaced by seeded
random graphs with the same node/edge counts, feature columns and dtypes.
The same generator feeds the CUDA path, the CPU oracle and `bench.py`, so both sides of a
parity test always see identical inputs.

A batch is a plain object with the fields the reference reads from a PyG ``Batch``
(`phc/hypercomplex/undirectional/models.py:220`): ``x``, ``edge_index`` (int64 ``[2,E]``,
row 0 = source, row 1 = target), ``edge_attr``, ``batch`` (int64 ``[N]``, ascending), ``y``,
``num_graphs``.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

ATOM_FEAT_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]   # ogb.utils.features.get_atom_feature_dims()
BOND_FEAT_DIMS = [5, 6, 2]                           # ogb.utils.features.get_bond_feature_dims()


class GraphBatch(object):
    """Duck-typed stand-in for ``torch_geometric.data.Batch``."""

    def __init__(self, x, edge_index, edge_attr, batch, y, num_graphs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.batch = batch
        self.y = y
        self.num_graphs = int(num_graphs)

    def to(self, device, non_blocking: bool = False):
        out = copy.copy(self)
        for k in ("x", "edge_index", "edge_attr", "batch", "y"):
            setattr(out, k, getattr(self, k).to(device, non_blocking=non_blocking))
        return out

    def pin_memory(self):
        out = copy.copy(self)
        for k in ("x", "edge_index", "edge_attr", "batch", "y"):
            setattr(out, k, getattr(self, k).pin_memory())
        return out

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_edges(self):
        return self.edge_index.size(1)

    def nbytes(self):
        return sum(getattr(self, k).numel() * getattr(self, k).element_size()
                   for k in ("x", "edge_index", "edge_attr", "batch", "y"))


@dataclass
class Workload:
    """One benchmark configuration: model kwargs (reference run_script_*_phm4.sh) + data shape."""
    name: str
    model: Dict
    loss: str                      # "bce" | "bce_masked" | "l1" | "ce"
    batch_graphs: int
    nodes_lo: int
    nodes_hi: int
    kind: str                      # "mol" | "knn" | "ppa"
    lr: float = 1e-3
    weight_decay: float = 0.0
    grad_clip: float = 2.0
    extra: Dict = field(default_factory=dict)


def _mol_edges(rng: np.random.Generator, nn: int, extra_frac: float = 0.08) -> np.ndarray:
    """Spanning chain + a few ring-closing edges, no self loops, no duplicates; returns the
    undirected edge list ``[m,2]`` with u<v."""
    chain = np.stack([np.arange(nn - 1), np.arange(1, nn)], 1)
    n_extra = int(round(extra_frac * nn)) + 1
    seen = set((int(a), int(b)) for a, b in chain)
    extra = []
    tries = 0
    while len(extra) < n_extra and tries < 20 * n_extra and nn > 2:
        u, v = rng.integers(0, nn, 2)
        tries += 1
        if u == v:
            continue
        a, b = (int(u), int(v)) if u < v else (int(v), int(u))
        if (a, b) in seen:
            continue
        seen.add((a, b))
        extra.append((a, b))
    if extra:
        return np.concatenate([chain, np.asarray(extra, dtype=np.int64)], 0)
    return chain


def _rand_undirected(rng: np.random.Generator, nn: int, m: int) -> np.ndarray:
    """m distinct undirected edges (u<v) on nn nodes, drawn uniformly."""
    m = min(m, nn * (nn - 1) // 2)
    got = np.zeros((0,), dtype=np.int64)
    while got.size < m:
        u = rng.integers(0, nn, 2 * (m - got.size) + 16)
        v = rng.integers(0, nn, u.size)
        keep = u != v
        lo, hi = np.minimum(u, v)[keep], np.maximum(u, v)[keep]
        got = np.unique(np.concatenate([got, lo * nn + hi]))
    got = rng.permutation(got)[:m]
    return np.stack([got // nn, got % nn], 1)


def _symmetrise(und: np.ndarray) -> np.ndarray:
    """[m,2] undirected -> [2,2m] directed with both directions adjacent (OGB convention)."""
    m = und.shape[0]
    ei = np.empty((2, 2 * m), dtype=np.int64)
    ei[0, 0::2], ei[1, 0::2] = und[:, 0], und[:, 1]
    ei[0, 1::2], ei[1, 1::2] = und[:, 1], und[:, 0]
    return ei


def graph_sizes(wl: Workload, seed: int, count: int) -> np.ndarray:
    """Node counts of ``count`` synthetic graphs (the first draw make_batch makes from the same seed)."""
    return np.random.default_rng(seed).integers(wl.nodes_lo, wl.nodes_hi + 1, count)


def make_batch(wl: Workload, seed: int = 0, batch_graphs: Optional[int] = None, sizes=None) -> GraphBatch:
    """sizes (optional): node count of every graph of the batch (e.g. one rank's share of a size-balanced global batch)."""
    rng = np.random.default_rng(seed)
    B = (wl.batch_graphs if batch_graphs is None else batch_graphs) if sizes is None else len(sizes)
    drawn = rng.integers(wl.nodes_lo, wl.nodes_hi + 1, B)
    sizes = drawn if sizes is None else np.asarray(sizes, dtype=drawn.dtype)
    xs, eis, eas, bs = [], [], [], []
    off = 0
    ex = wl.extra
    for g in range(B):
        nn = int(sizes[g])
        if wl.kind == "mol":
            und = _mol_edges(rng, nn)
            ei = _symmetrise(und)
            m = und.shape[0]
            if ex["atom_dims"] is None:
                raise ValueError
            x = np.stack([rng.integers(0, d, nn) for d in ex["atom_dims"]], 1)
            ea_u = np.stack([rng.integers(0, d, m) for d in ex["bond_dims"]], 1)
            ea = np.repeat(ea_u, 2, axis=0)
            if ex.get("squeeze", False):
                x, ea = x[:, 0], ea[:, 0]
        elif wl.kind == "knn":
            k = ex["k"]
            pos = rng.random((nn, 2))
            d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
            np.fill_diagonal(d2, np.inf)
            nbr = np.argsort(d2, axis=1, kind="stable")[:, :k]          # [nn,k]
            dst = np.repeat(np.arange(nn), k)
            src = nbr.reshape(-1)
            ei = np.stack([src, dst], 0).astype(np.int64)
            x = rng.random((nn, ex["node_dim"])).astype(np.float32)
            ea = rng.random((ei.shape[1], ex["edge_dim"])).astype(np.float32)
        elif wl.kind == "ppa":
            und = _rand_undirected(rng, nn, ex["und_edges"])
            ei = _symmetrise(und)
            x = np.zeros((nn,), dtype=np.int64)
            ea_u = rng.random((und.shape[0], ex["edge_dim"])).astype(np.float32)
            ea = np.repeat(ea_u, 2, axis=0)
        else:
            raise ValueError(wl.kind)
        xs.append(x)
        eis.append(ei + off)
        eas.append(ea)
        bs.append(np.full((nn,), g, dtype=np.int64))
        off += nn
    x = torch.from_numpy(np.concatenate(xs, 0))
    ei = torch.from_numpy(np.concatenate(eis, 1))
    ea = torch.from_numpy(np.concatenate(eas, 0))
    b = torch.from_numpy(np.concatenate(bs, 0))
    if wl.loss == "bce":
        y = torch.from_numpy(rng.integers(0, 2, (B, 1)).astype(np.float32))
    elif wl.loss == "bce_masked":
        t = wl.model["target_dim"]
        y = rng.integers(0, 2, (B, t)).astype(np.float32)
        y[rng.random((B, t)) < 0.5] = np.nan
        y = torch.from_numpy(y)
    elif wl.loss == "l1":
        y = torch.from_numpy(rng.standard_normal(B).astype(np.float32))
    elif wl.loss == "ce":
        y = torch.from_numpy(rng.integers(0, wl.model["target_dim"], B).astype(np.int64))
    else:
        raise ValueError(wl.loss)
    return GraphBatch(x, ei, ea, b, y, B)


def _base_model(**kw) -> Dict:
    d = dict(phm_dim=4, learn_phm=True, phm_rule=None, naive_encoder=False, w_init="phm", c_init="standard",
             same_dropout=False, bias=True, norm_mp="naive-batch-norm", add_self_loops=True, node_aggr="sum",
             pooling="softattention", activation="relu", real_trafo="linear", norm_dn="naive-batch-norm",
             msg_encoder="identity", sc_type="first")
    d.update(kw)
    return d


def workloads(phm_dim: int = 4) -> Dict[str, Workload]:
    """The five BASELINE.json configs (C1..C5) with the reference's default hyper-parameters
    (`benchmarks/run_script_{hiv,zinc,pcba,mnist,cifar10,ppa}_phm4.sh`)."""
    n = phm_dim
    w = {}
    w["hiv"] = Workload(
        "hiv", _base_model(phm_dim=n, atom_input_dims=ATOM_FEAT_DIMS, bond_input_dims=BOND_FEAT_DIMS,
                           atom_encoded_dim=200, mp_layers=[200, 200], dropout_mpnn=[0.2, 0.2], mlp=True,
                           msg_aggr="softmax", downstream_layers=[128, 32], dropout_dn=[0.3, 0.1], target_dim=1,
                           initial_beta=1.0, learn_beta=True),
        "bce", 128, 11, 40, "mol", lr=1e-3, weight_decay=0.1,
        extra=dict(atom_dims=ATOM_FEAT_DIMS, bond_dims=BOND_FEAT_DIMS))
    w["zinc"] = Workload(
        "zinc", _base_model(phm_dim=n, atom_input_dims=[28], bond_input_dims=[4], atom_encoded_dim=180,
                            mp_layers=[180] * 4, dropout_mpnn=[0.0] * 4, mlp=True, msg_aggr="sum",
                            downstream_layers=[180, 80], dropout_dn=[0.2, 0.1], target_dim=1, sc_type="last"),
        "l1", 128, 9, 37, "mol", lr=1e-3, weight_decay=0.01,
        extra=dict(atom_dims=[28], bond_dims=[4], squeeze=True))
    w["pcba"] = Workload(
        "pcba", _base_model(phm_dim=n, atom_input_dims=ATOM_FEAT_DIMS, bond_input_dims=BOND_FEAT_DIMS,
                            atom_encoded_dim=512, mp_layers=[512] * 7, dropout_mpnn=[0.1] * 7, mlp=False,
                            msg_aggr="sum", downstream_layers=[768, 256], dropout_dn=[0.3, 0.2], target_dim=128),
        "bce_masked", 512, 12, 40, "mol", lr=5e-4, weight_decay=1e-4,
        extra=dict(atom_dims=ATOM_FEAT_DIMS, bond_dims=BOND_FEAT_DIMS))
    for nm, lo, hi, nd in (("mnist", 66, 75, 3), ("cifar", 110, 125, 5)):
        w[nm] = Workload(
            nm, _base_model(phm_dim=n, atom_input_dims=nd, bond_input_dims=1, atom_encoded_dim=224,
                            mp_layers=[224] * 4, dropout_mpnn=[0.1] * 4, mlp=False, msg_aggr="mean",
                            downstream_layers=[256, 128], dropout_dn=[0.2, 0.1], target_dim=10, sc_type="last"),
            "ce", 128, lo, hi, "knn", lr=1e-3, weight_decay=0.01, extra=dict(k=8, node_dim=nd, edge_dim=1))
    w["ppa"] = Workload(
        "ppa", _base_model(phm_dim=n, atom_input_dims=[1], bond_input_dims=7, atom_encoded_dim=500,
                           mp_layers=[500] * 7, dropout_mpnn=[0.2] * 7, mlp=True, msg_aggr="sum",
                           downstream_layers=[512, 256], dropout_dn=[0.3, 0.2], target_dim=37),
        "ce", 64, 187, 300, "ppa", lr=5e-4, weight_decay=1e-4, extra=dict(und_edges=2266, edge_dim=7))
    return w


def tiny(wl: Workload, width: int, layers: int, graphs: int, nodes_lo: int, nodes_hi: int,
         und_edges: Optional[int] = None, head: Optional[List[int]] = None) -> Workload:
    """Shrink a workload for parity tests / golden fixtures (same structure, small tensors)."""
    out = copy.deepcopy(wl)
    m = out.model
    m["atom_encoded_dim"] = width
    m["mp_layers"] = [width] * layers
    m["dropout_mpnn"] = m["dropout_mpnn"][:1] * layers
    if head is not None:
        m["downstream_layers"] = head
    out.batch_graphs, out.nodes_lo, out.nodes_hi = graphs, nodes_lo, nodes_hi
    if und_edges is not None:
        out.extra["und_edges"] = und_edges
    return out


def split_graphs(data: GraphBatch) -> List[GraphBatch]:
    """Cut a synthetic mini-batch (edges grouped by graph, as ``make_batch`` builds them) into single-graph pieces with
    local node ids — a synthetic DATASET for ``prep.DeviceGraphStore``."""
    B = int(data.num_graphs)
    counts = torch.bincount(data.batch, minlength=B)
    nptr = [0] + torch.cumsum(counts, 0).tolist()
    ecounts = torch.bincount(data.batch[data.edge_index[0]], minlength=B)
    eptr = [0] + torch.cumsum(ecounts, 0).tolist()
    out = []
    for b in range(B):
        n0, n1, e0, e1 = nptr[b], nptr[b + 1], eptr[b], eptr[b + 1]
        out.append(GraphBatch(data.x[n0:n1].clone(), (data.edge_index[:, e0:e1] - n0).clone(), data.edge_attr[e0:e1].clone(),
                              torch.zeros(n1 - n0, dtype=torch.int64), data.y[b:b + 1].clone(), 1))
    return out
