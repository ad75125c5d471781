#!/usr/bin/env python
"""Per-parameter gradient error of a full-size benchmark configuration against the fp64 oracle, for several precision modes.
usage: python tools/diag_fullsize.py ppa [modes...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from gpu_util import oracle_train_eval, product_train_eval
from phc.hypercomplex.undirectional.models import PHMSkipConnectAdd
from phc_gnn_b200.synthetic import workloads, make_batch
wname = sys.argv[1] if len(sys.argv) > 1 else "ppa"
modes = sys.argv[2:] or ["fp32", "tf32x3"]
wl = workloads(4)[wname]
cfg = dict(wl.model)
cfg["dropout_mpnn"] = [0.0] * len(cfg["mp_layers"]); cfg["dropout_dn"] = [0.0] * len(cfg["downstream_layers"])
torch.manual_seed(0); np.random.seed(0)
state = PHMSkipConnectAdd(**cfg).state_dict()
batch = make_batch(wl, seed=11)
want = oracle_train_eval(cfg, state, batch, wl.loss, 0.01, torch.float64)
noise = oracle_train_eval(cfg, state, batch, wl.loss, 0.01, torch.float32)
res = {}
for mode in modes:
    os.environ["PHC_PRECISION"] = mode
    res[mode] = product_train_eval(cfg, state, batch, wl.loss, 0.01, "cuda:0")
def rel(a, b): return float((a.double() - b.double()).abs().max()) / max(float(b.abs().max()), 1e-30)
print(f"{'quantity':58s} {'fp32-oracle':>11s} " + " ".join(f"{m:>11s}" for m in modes))
print(f"{'logits':58s} {rel(noise['logits'], want['logits']):11.2e} " + " ".join(f"{rel(res[m]['logits'], want['logits']):11.2e}" for m in modes))
print(f"{'loss':58s} {rel(noise['loss'], want['loss']):11.2e} " + " ".join(f"{rel(res[m]['loss'], want['loss']):11.2e}" for m in modes))
for k in want["grads"]:
    row = [rel(noise["grads"][k], want["grads"][k])] + [rel(res[m]["grads"][k], want["grads"][k]) for m in modes]
    flag = " <<<" if max(row[1:]) > 1e-3 else ""
    print(f"{k:58s} " + " ".join(f"{v:11.2e}" for v in row) + flag)
