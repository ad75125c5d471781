#!/usr/bin/env python
"""Unprofiled ms/step of the training step (A/B switches through the environment: PHC_NO_PDL, PHC_B200_LIB, ...)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phc_gnn_b200 import graph
from phc_gnn_b200.nn import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import make_batch, workloads
from phc_gnn_b200.train import TrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
wl = workloads(4)[name]
if len(sys.argv) > 3:               # fewer graphs per batch: the GPU work shrinks, the step time approaches the pure host cost
    wl.batch_graphs = int(sys.argv[3])
if os.environ.get("STEP_AGGR"):   # e.g. STEP_AGGR=max: the workload with another aggregator (run_script_ppa_phm4.sh alternates sum / max)
    wl.model["msg_aggr"] = os.environ["STEP_AGGR"]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = PHMSkipConnectAdd(**wl.model).to(dev)
step = TrainStep(model, wl)
model.train()
batches = [make_batch(wl, seed=i).to(dev) for i in range(8)]
for i in range(8):
    graph.clear_cache(); step(batches[i % 8])
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(steps):
        graph.clear_cache(); step(batches[i % 8])
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / steps)
    print(f"{name}: {e0.elapsed_time(e1) / steps:.3f} ms/step (host submit {1e3 * (t1 - t0) / steps:.3f} ms/step)", flush=True)
print(f"{name} best {best:.3f} ms/step  env: " + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PHC_")))
