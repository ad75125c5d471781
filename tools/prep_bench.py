#!/usr/bin/env python
"""Micro-benchmark of the batch-preparation kernels against the HBM roofline:

    python tools/prep_bench.py [workload=ppa] [dataset_graphs=2048] [batch_graphs=<workload default>] [reps=50]

* collate: ``DeviceGraphStore.collate`` (phc_collate_batch) — algorithmic bytes = 2 x bytes of the assembled batch (every
  byte read once from the packed store and written once), + 8 B read per directed edge endpoint for the index shift;
* RemoveIsolatedNodes (phc_remove_isolated_nodes) — algorithmic bytes = 16 B per edge read + 16 B per edge written
  (edge_index in, relabelled edge_index out) + the node mask / assoc vectors.

The dataset is built from ``dataset_graphs`` synthetic graphs of the workload's shape (larger than L2 for ppa), batches are
random selections, timing is CUDA events on the launching stream after a warm-up.  Peak = MEASURED_PEAKS.json if present.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from phc_gnn_b200.prep import DeviceGraphStore, remove_isolated_nodes  # noqa: E402
from phc_gnn_b200.synthetic import make_batch, split_graphs, workloads  # noqa: E402


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            d = json.load(fh)
        for k in ("hbm_gbs_burst", "hbm_copy_gbs", "hbm_gbs"):
            if k in d:
                return float(d[k]), "measured"
        for v in d.values():
            if isinstance(v, dict):
                for k in ("burst", "gbs", "GBps"):
                    if k in v:
                        return float(v[k]), "measured"
    except Exception:
        pass
    return 7700.0, "nominal"


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
    wl = workloads(4)[name]
    G = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    B = int(sys.argv[3]) if len(sys.argv) > 3 else wl.batch_graphs
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 50
    dev = torch.device("cuda:0")
    graphs = []
    seed = 0
    while len(graphs) < G:
        graphs += split_graphs(make_batch(wl, seed=seed, batch_graphs=min(256, G - len(graphs))))
        seed += 1
    store = DeviceGraphStore(graphs, dev)
    rng = np.random.default_rng(0)
    sel = [rng.integers(0, G, size=B) for _ in range(reps + 5)]
    for ids in sel[:5]:
        batch = store.collate(ids)
    store.check_status()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    moved = 0
    e0.record()
    for ids in sel[5:]:
        batch = store.collate(ids)
        moved += 2 * batch.nbytes()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    peak, src = hbm_peak()
    gbs = moved / reps / ms / 1e6
    print(f"collate {name}: store {store.nbytes() / 1e6:.1f} MB ({G} graphs), batch {B} graphs = {batch.nbytes() / 1e6:.2f} MB; "
          f"{ms * 1e3:.1f} us per batch incl. host (ids + prefix sums + allocation), {gbs:.0f} GB/s = {gbs / peak:.1%} of {peak:.0f} GB/s ({src})")

    ei, ea = batch.edge_index, batch.edge_attr
    N = batch.num_nodes
    for _ in range(3):
        remove_isolated_nodes(ei, ea, N)
    e0.record()
    for _ in range(reps):
        remove_isolated_nodes(ei, ea, N)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    byt = 32 * ei.size(1) + 9 * N
    print(f"remove_isolated_nodes {name}: N={N} E={ei.size(1)}: {ms * 1e3:.1f} us per call incl. the size read-back, "
          f"{byt / ms / 1e6:.0f} GB/s algorithmic = {byt / ms / 1e6 / peak:.1%} of peak")


    # kernel-only durations (CUPTI via torch.profiler: warm caches, real clocks) — the numbers above include the host side
    import collections
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for ids in sel[5:5 + 10]:
            store.collate(ids)
        for _ in range(10):
            remove_isolated_nodes(ei, ea, N)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            nm = ev.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0]
            agg[nm][0] += 1
            agg[nm][1] += dur
    cb = 2 * batch.nbytes()
    for nm, (cnt, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        line = f"  kernel {nm:40s} n={cnt:3d} avg {tot / cnt:8.2f} us"
        if "collate" in nm:
            line += f"  -> {cb / (tot / cnt) / 1e3:.0f} GB/s = {cb / (tot / cnt) / 1e3 / peak:.1%} of peak (algorithmic 2 x batch bytes)"
        print(line)


if __name__ == "__main__":
    main()
