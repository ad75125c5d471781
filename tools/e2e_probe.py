#!/usr/bin/env python
"""Where does the end-to-end loop lose time?  (diagnostic)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from phc_gnn_b200 import graph
from phc_gnn_b200.nn import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import make_batch, workloads
from phc_gnn_b200.train import TrainStep, make_optimizer
name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
wl = workloads(4)[name]
dev = torch.device("cuda:0")
for flat in (True, False):
    torch.manual_seed(0)
    model = PHMSkipConnectAdd(**wl.model).to(dev)
    step = TrainStep(model, wl, None if flat else make_optimizer(model, wl.lr))
    model.train()
    host = [make_batch(wl, seed=i).pin_memory() for i in range(4)]
    devb = [b.to(dev) for b in host]
    for i in range(4):
        graph.clear_cache(); step(devb[i % 4])
    torch.cuda.synchronize()
    K = 10
    def timed(fn):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0) / K
    def plain():
        for i in range(K):
            graph.clear_cache(); step(devb[i % 4])
    def lagged_loss():
        lh = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        ev = [torch.cuda.Event(), torch.cuda.Event()]
        for i in range(K):
            graph.clear_cache(); l = step(devb[i % 4])
            lh[i & 1].copy_(l, non_blocking=True); ev[i & 1].record()
            if i > 0:
                ev[(i - 1) & 1].synchronize(); float(lh[(i - 1) & 1])
    def h2d_main():
        for i in range(K):
            d = host[i % 4].to(dev, non_blocking=True)
            graph.clear_cache(); step(d)
    def sync_each():
        for i in range(K):
            graph.clear_cache(); float(step(devb[i % 4]).item())
    print(f"flat_opt={flat}: plain {timed(plain):.2f}  lagged_loss {timed(lagged_loss):.2f}  h2d_on_main {timed(h2d_main):.2f}  "
          f"sync_each {timed(sync_each):.2f} ms/step", flush=True)
