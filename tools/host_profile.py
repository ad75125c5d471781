#!/usr/bin/env python
"""cProfile of the host side of one training step (small, launch-bound workloads)."""
import cProfile, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phc_gnn_b200 import graph
from phc_gnn_b200.nn import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import make_batch, workloads
from phc_gnn_b200.train import TrainStep
name = sys.argv[1] if len(sys.argv) > 1 else "zinc"
wl = workloads(4)[name]
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = PHMSkipConnectAdd(**wl.model).to(dev)
step = TrainStep(model, wl)        # flat clip+Adam, as bench.py runs it
model.train()
batches = [make_batch(wl, seed=i).to(dev) for i in range(4)]
for i in range(5):
    graph.clear_cache(); step(batches[i % 4])
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(20):
    graph.clear_cache(); step(batches[i % 4])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host-side submit {1e3*(t1-t0)/20:.2f} ms/step, with sync {1e3*(t2-t0)/20:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    graph.clear_cache(); step(batches[i % 4])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(45)
