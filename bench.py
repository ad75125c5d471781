#!/usr/bin/env python
"""Throughput benchmark for the PHC-GNN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ppa|cifar|mnist|pcba|zinc|hiv] [--impl reference]

Metric (BASELINE.json): train graphs/sec (fwd+bwd+step); roofline = achieved HBM GB/s of the fused
neighbour-aggregation kernel vs the measured peak.  A step is one iteration of the reference's
train() body (benchmarks/train_hiv.py:170-202) on one synthetic mini-batch per GPU (weak scaling:
per-GPU batch fixed).  N>1 is launched by torchrun, one rank per GPU, gradients all-reduced by NCCL.
Prints ONE JSON line on rank 0.  `--impl reference` times the CPU oracle port of the reference
(oracle/phc_oracle.py; the reference is Python + PyG/torch_scatter which are not installable here).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DEFAULT_PRECISION = "tf32x3"      # what phc_gnn_b200.ops.default_precision() resolves to when PHC_PRECISION is unset

JSON_OUT = sys.stdout
METRIC = "train_graphs_per_sec"
UNIT = "graphs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="ppa")
    ap.add_argument("--phm-dim", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--family", default="phm", choices=["phm", "quaternion"],
                    help="quaternion: the reference's QuaternionSkipConnectAdd on the same kernels (n = 4, frozen Hamilton rule)")
    ap.add_argument("--batches", type=int, default=8, help="distinct synthetic batches cycled per rank")
    ap.add_argument("--precision", default=None, help="fp32 | tf32x3 | bf16 (default: PHC_PRECISION or tf32x3)")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="on: the step is replayed from one CUDA graph per batch shape (phc_gnn_b200/graphed.py; data parallel: the "
                         "all-reduce and the optimizer kernels stay eager between graph launches); off: eager host path (one all-reduce after "
                         "backward; PHC_OVERLAP_ALLREDUCE=1 slices and overlaps it); auto: graph where replay measured a gain (nodes x width per batch < 7.5e6: "
                         "hiv, zinc, mnist, cifar 1.7-2.8x, pcba 1.13x), eager for ppa (kernel-bound: replay 5.13 vs eager 5.09 ms)")
    ap.add_argument("--sharding", default="balanced", choices=["balanced", "random"],
                    help="N > 1: how the graphs of a global batch are dealt to the ranks (balanced: by node count, prep.balanced_partition)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--kernel-timers", action="store_true", help="also print the per-op CUDA-event breakdown to stderr")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak():
    """Dense bf16 TFLOP/s the driver measured on this pool (sustained figure: the kernels are timed inside a long step)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        v = d.get("bf16_tflops_sustained") or d.get("bf16_tflops")
        if v:
            return float(v), "measured (MEASURED_PEAKS.json, sustained)"
    return 2250.0, "nominal dense bf16 (B200_PROFILING.md)"


# tensor-core passes per fp32 product and the MMA rate relative to dense bf16, per precision mode
PASSES = {"tf32x3": (3, 0.5), "bf16x3": (3, 1.0), "bf16": (1, 1.0)}


def traffic_record(kernel):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of ``kernel`` from the committed ``ncu --set full``
    capture named in profiles/traffic.json (one capture per kernel change; bench.py cannot run under ncu itself)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as fh:
        d = json.load(fh).get(kernel)
    return (d["bytes_per_launch"], d["source"]) if d else (None, None)


def phm_linear_roofline(prof, steps, wl, N, precision, step_ms=None):
    """The dominant kernels of the step: the node-level PHMLinear calls (tcgen05 mix kernel forward / dX, dH kernel).  achieved =
    algorithmic FLOPs (2*M*in*out per forward launch, 4*M*in*out per backward call) / CUDA-event time of those calls in
    the instrumented repeat; peak = the measured dense bf16 rate.  Head-level calls (M = graphs per batch) are timed under their own
    key and excluded from both.  Never raises: returns None when the entries are missing."""
    try:
        cf, tf = prof.get("phc_phm_linear_fwd:node", (0, 0.0))
        cb, tb = prof.get("phc_phm_linear_bwd:node", (0, 0.0))
        if not cf or not cb or tf <= 0 or tb <= 0:
            return None
        m = wl.model
        F = m["mp_layers"][0]
        n_lin = cf / steps
        unit = 2.0 * N * F * F
        fwd_us = 1e3 * tf / cf
        bwd_us = 1e3 * tb / cb
        fwd_tf = unit / (fwd_us * 1e-6) / 1e12
        bwd_tf = 2.0 * unit / (bwd_us * 1e-6) / 1e12
        ach = 3.0 * unit / ((fwd_us + bwd_us) * 1e-6) / 1e12
        peak, src = tensor_peak()
        ceiling = None
        if precision in PASSES:
            passes, rate = PASSES[precision]
            ceiling = peak * rate / passes       # fp32-equivalent ceiling: every product costs `passes` MMAs at `rate` x the bf16 rate
        traffic, tsrc = traffic_record("phm_tc_mix")
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": tsrc, "peak_source": src,
                "kernel": "node-level PHMLinear: phm_tc_mix (forward, dX) + phm_tc_dh (dH) + small contraction kernels; fp32-equivalent "
                          "algorithmic FLOPs 6*M*in*out per linear and step",
                "forward_launch": {"avg_us": fwd_us, "tflops": fwd_tf, "flops": unit},
                "backward_call": {"avg_us": bwd_us, "tflops": bwd_tf, "flops": 2.0 * unit},
                "node_level_linears": n_lin, "rows": N, "in_out": F, "algorithmic_flops_per_step": 3.0 * n_lin * unit,
                "share_of_step": ((tf + tb) / steps / step_ms) if step_ms else None,
                "precision": precision, "precision_ceiling_tflops": ceiling,
                "frac_of_precision_ceiling": (ach / ceiling) if ceiling else None,
                "timed": "CUDA events around each C-ABI call on the launching stream, instrumented repeat of the K steps"}
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def workload(args):
    from phc_gnn_b200.synthetic import workloads
    return workloads(args.phm_dim)[args.workload]


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "phc", "hypercomplex", "undirectional", "models.py"))


def _import_reference():
    """The UNMODIFIED reference package installed by baseline/install_ref.sh (byte-identical to /root/reference/phc),
    imported under its own name ``phc`` with oracle/refshim standing in for torch_scatter / torch_geometric / ogb
    (not installable offline).  The product's drop-in ``phc`` package at the repo root must not shadow it: the name is
    pinned to baseline/_ref/phc before anything is imported, and the origin of the model module is asserted."""
    import types
    import warnings
    warnings.filterwarnings("ignore")
    for k in [k for k in sys.modules if k == "phc" or k.startswith("phc.")]:
        del sys.modules[k]
    sys.path.insert(0, os.path.join(ROOT, "oracle", "refshim"))
    pkg = types.ModuleType("phc")
    pkg.__path__ = [os.path.join(REF_DIR, "phc")]
    sys.modules["phc"] = pkg
    import phc.hypercomplex.undirectional.models as ref_models
    from phc.hypercomplex.regularization import phm_weight_regularization as ref_reg
    assert os.path.realpath(ref_models.__file__).startswith(os.path.realpath(REF_DIR)), ref_models.__file__
    return ref_models.PHMSkipConnectAdd, ref_reg


class _ReferenceStepper(object):
    """One iteration of the reference's train() body (benchmarks/train_hiv.py:170-202) on its own model class,
    torch.optim.Adam and clip_grad_norm_, on the host cores.  kind = "reference" when baseline/_ref is present,
    else the oracle port ("port")."""

    def __init__(self, wl):
        import torch.nn.functional as F
        self.wl, self.F = wl, F
        torch.manual_seed(0)
        if reference_available():
            Model, self.reg = _import_reference()
            self.kind = "reference"
            self.model = Model(**wl.model)
            with torch.no_grad():      # the reference leaves one bias element uninitialised: define it
                for n_, p_ in self.model.named_parameters():
                    if not torch.isfinite(p_).all():
                        p_.copy_(torch.nan_to_num(p_, nan=0.2, posinf=0.2, neginf=0.2))
            self.model.train()
            self.opt = torch.optim.Adam(self.model.parameters(), lr=wl.lr)
            self.params = list(self.model.parameters())
        else:
            from oracle import phc_oracle as O
            from phc_gnn_b200.nn import PHMSkipConnectAdd
            self.kind, self.O = "port", O
            state = PHMSkipConnectAdd(**wl.model).state_dict()
            self.p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone())
                      for k, v in state.items()}
            self.opt = torch.optim.Adam(O.trainable(self.p), lr=wl.lr)
            self.gen = torch.Generator().manual_seed(0)

    def step(self, data):
        wl, F = self.wl, self.F
        if self.kind == "port":
            return float(self.O.train_step(self.p, wl.model, data, wl.loss, self.opt, wl.lr, wl.weight_decay, wl.grad_clip, self.gen))
        self.opt.zero_grad()
        logits = self.model(data)
        if wl.loss in ("bce", "bce_masked"):
            mask = ~torch.isnan(data.y)
            loss = F.binary_cross_entropy_with_logits(input=logits[mask], target=data.y[mask].to(torch.float))
        elif wl.loss == "l1":
            loss = (logits.squeeze() - data.y).abs().mean()
        else:
            loss = F.cross_entropy(logits, data.y.view(-1))
        if wl.weight_decay > 0.0:
            loss = loss + wl.lr * wl.weight_decay * self.reg(self.model, p=2).squeeze()
        loss.backward()
        if wl.grad_clip > 0.0:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=wl.grad_clip, norm_type=2)
        self.opt.step()
        return float(loss.item())

    def describe(self):
        return ("the unmodified reference (baseline/_ref, PHMSkipConnectAdd + torch.optim.Adam + clip_grad_norm_) over oracle/refshim"
                if self.kind == "reference" else "oracle/phc_oracle.py (port of the reference; baseline/_ref not installed)")


def run_reference(args):
    """CPU arm: the reference's own implementation of the path on the host cores, at the SAME per-GPU batch as the b200
    arm (one step = one full mini-batch), all host threads; the reference scripts' own setting (6 threads,
    benchmarks/train_hiv.py:632) is timed on a few extra steps and reported next to it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from phc_gnn_b200.synthetic import make_batch
    wl = workload(args)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ref = _ReferenceStepper(wl)
    batches = [make_batch(wl, seed=i) for i in range(min(args.batches, 4))]
    for i in range(args.warmup):
        ref.step(batches[i % len(batches)])
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss = ref.step(batches[(args.warmup + i) % len(batches)])
    dt = time.perf_counter() - t0
    val = wl.batch_graphs * args.steps / dt
    six = None
    if cores > 6:
        torch.set_num_threads(6)
        k6 = max(2, min(args.steps, 5))
        ref.step(batches[0])
        t6 = time.perf_counter()
        for i in range(k6):
            ref.step(batches[i % len(batches)])
        d6 = time.perf_counter() - t6
        six = {"value": wl.batch_graphs * k6 / d6, "unit": UNIT, "cores": 6, "steps": k6,
               "why": "torch.set_num_threads(6) is what the reference scripts run with (benchmarks/train_hiv.py:632)"}
    sample = (f"{args.steps} steps x {wl.batch_graphs} graphs (the full per-GPU batch) of the {wl.name}-shaped workload, "
              f"{ref.describe()}, torch CPU, {cores} threads, {dt:.1f} s")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, wl, wl.batch_graphs, "cpu"),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample, "six_threads": six},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "final_loss": loss}), file=JSON_OUT, flush=True)


def config_dict(args, wl, graphs_per_gpu, l2):
    m = wl.model
    return {"workload": f"{wl.name}-shaped synthetic graphs (BASELINE.json configs: PHC-GNN n={m['phm_dim']}, "
                        f"{len(m['mp_layers'])}x{m['mp_layers'][0]}, aggr={m['msg_aggr']}, mlp={m['mlp']})",
            "graphs_per_gpu_batch": graphs_per_gpu, "phm_dim": m["phm_dim"], "width": m["mp_layers"][0],
            "layers": len(m["mp_layers"]), "aggr": m["msg_aggr"], "parallelism": f"dp{args.gpus}", "l2": l2,
            **({"sharding": getattr(args, "sharding", "balanced")} if args.gpus > 1 else {}),
            **({"family": "quaternion"} if getattr(args, "family", "phm") == "quaternion" else {})}


def cpu_baseline(wl, budget_s: float = 20.0):
    """The reference's CPU implementation (baseline/_ref; else the oracle port) timed on the host cores for a bounded
    sample of the same workload at the same per-GPU batch (rank 0, N=1 only)."""
    from phc_gnn_b200.synthetic import make_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    saved = {k: v for k, v in sys.modules.items() if k == "phc" or k.startswith("phc.")}
    try:
        ref = _ReferenceStepper(wl)
        batch = make_batch(wl, seed=0)
        ref.step(batch)      # warm-up
        t0 = time.perf_counter()
        steps = 0
        while steps < 2 or (time.perf_counter() - t0 < budget_s and steps < 50):
            ref.step(batch)
            steps += 1
        dt = time.perf_counter() - t0
    finally:             # give the name ``phc`` back to the product's drop-in package
        for k in [k for k in sys.modules if k == "phc" or k.startswith("phc.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return {"value": wl.batch_graphs * steps / dt, "unit": UNIT, "cores": cores, "kind": ref.kind,
            "sample": f"{steps} steps x {wl.batch_graphs} graphs (the full per-GPU batch) of the same {wl.name}-shaped workload, "
                      f"{ref.describe()}, torch CPU, {cores} threads, {dt:.1f} s"}


def aggregation_bytes(N, E, F, softmax):
    """Algorithmic HBM bytes of one fused aggregation forward."""
    return 4 * F * (2 * N + E) + 8 * E + 4 * (N + 1) + (8 * N * F if softmax else 0)


def run_b200(args):
    import torch.distributed as dist
    from phc_gnn_b200 import ops, graph
    from phc_gnn_b200.nn import PHMSkipConnectAdd
    from phc_gnn_b200.parallel import DataParallelPHC
    from phc_gnn_b200.synthetic import make_batch
    from phc_gnn_b200.train import TrainStep, make_optimizer

    if args.precision:
        os.environ["PHC_PRECISION"] = args.precision
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep NCCL's own messages off stdout: ONE JSON line there
        dist.init_process_group("nccl", device_id=dev)
    wl = workload(args)
    torch.manual_seed(0)
    import numpy as np
    np.random.seed(0)
    if args.family == "quaternion":
        from phc_gnn_b200.quaternion import QuaternionSkipConnectAdd
        assert args.phm_dim == 4, "--family quaternion is the n = 4 configuration"
        qkw = {k: v for k, v in wl.model.items() if k not in ("phm_dim", "learn_phm", "phm_rule", "w_init", "c_init", "sc_type")}
        model = QuaternionSkipConnectAdd(init="orthogonal", **qkw).to(dev)
    else:
        model = PHMSkipConnectAdd(**wl.model).to(dev)
    dp = DataParallelPHC(model) if world > 1 else None
    step = TrainStep(model, wl, None, dp)        # flat clip+Adam (optim.FlatClipAdam): same update rule, 2 launches
    use_graph = args.graph == "on"
    if args.graph == "auto":
        probe = make_batch(wl, seed=0)
        use_graph = probe.num_nodes * wl.model["mp_layers"][0] < 7.5e6
    if use_graph:
        from phc_gnn_b200.graphed import GraphedTrainStep
        step = GraphedTrainStep(step, max_graphs=2 * args.batches + 4)
    model.train()

    if world > 1 and args.sharding == "balanced":
        # every global batch (world x B graphs, sizes drawn from one seed all ranks share) is dealt to the ranks by node count
        # (prep.balanced_partition): same graphs per rank, nearly equal nodes per rank -> no straggler wait in the synchronous step
        from phc_gnn_b200.prep import balanced_partition
        from phc_gnn_b200.synthetic import graph_sizes
        host = []
        for i in range(args.batches):
            gs = graph_sizes(wl, 7000 + i, world * wl.batch_graphs)
            mine = gs[balanced_partition(gs, world)[rank]]
            host.append(make_batch(wl, seed=rank * 1000 + i, sizes=mine).pin_memory())
    else:
        host = [make_batch(wl, seed=rank * 1000 + i).pin_memory() for i in range(args.batches)]
    devb = [b.to(dev) for b in host]
    ws_bytes = sum(b.num_edges for b in host) / len(host) * wl.model["mp_layers"][0] * 4
    flush = ws_bytes < 256e6          # per-layer edge tensor smaller than 2x L2 -> flush L2 between steps
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one(i, timing_events=None):
        graph.clear_cache()                      # the CSR/segment build is part of every step
        if flush:
            flush_buf.fill_(i & 0xFF)
        if timing_events is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        loss = step(devb[i % len(devb)])
        if timing_events is not None:
            b.record()
            timing_events.append((a, b))
        return loss

    # untimed: one step on every distinct batch first (each batch shape is new to the caching allocator: without this the
    # first timed visit of a batch pays cudaMalloc inside the timed region), then the W warm-up steps
    preroll = 0
    for i in range(len(devb)):
        one(i)
        preroll += 1
    # ... and keep stepping for about a second: clocks, NCCL channels and the allocator reach their steady state (a 2-GPU run
    # measured 5.59 ms/step in a timed region that started 0.3 s after the first kernel, 5.22 ms/step once warm)
    torch.cuda.synchronize()
    t_pre = time.perf_counter()
    extra = 0
    while extra < 200:
        one(extra)
        extra += 1
        if extra % 10 == 0:
            torch.cuda.synchronize()
            flag = torch.tensor([1.0 if time.perf_counter() - t_pre < 1.0 else 0.0], device=dev)
            if world > 1:
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)      # all ranks leave the pre-roll together
            if float(flag.item()) == 0.0:
                break
    preroll = {"steps": preroll + extra, "why": "untimed, BEFORE the W warm-up steps: one step per distinct batch shape (allocator) plus "
               "about one second of stepping until clocks / NCCL channels are in steady state; the reference arm does W only"}
    for i in range(args.warmup):
        one(i)
    # ---- timed region: K steps, device-resident batches, no per-op instrumentation --------------------------
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ops.PROFILE.reset(timing=False)
    evs = []
    for i in range(args.steps):
        loss = one(args.warmup + i, evs)
    barrier()
    clocks = sampler.result()
    launches = ops.PROFILE.launches
    ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    graphs = wl.batch_graphs * args.steps * world
    value = graphs / (ms / 1e3)

    # ---- the same K steps again with a CUDA-event pair around every kernel call (per-op breakdown / roofline);
    #      kept out of the headline region because the event records themselves cost host time ---------------
    barrier()
    ops.PROFILE.reset(timing=True)
    evs2 = []
    for i in range(args.steps):
        one(args.warmup + i, evs2)
    barrier()
    prof = ops.PROFILE.summary()
    ops.PROFILE.reset(timing=False)
    ms_instr = sum(a.elapsed_time(b) for a, b in evs2)

    # the unfused aggregation kernel at the SURVEY 8(d) boundary (edge embedding [E,F] given), same batch shapes:
    # reported next to the fused kernel for reference; it is NOT part of the timed step any more
    unfused = None
    if rank == 0:
        bt = devb[0]
        Fw = wl.model["mp_layers"][0]
        xs = torch.randn(bt.num_nodes, Fw, device=dev)
        es = torch.randn(bt.num_edges, Fw, device=dev)
        st = graph.EdgeStructure(bt.edge_index, bt.num_nodes)
        red = "add" if wl.model["msg_aggr"] in ("sum", "add") else wl.model["msg_aggr"]
        beta_t = torch.ones((), device=dev)
        for _ in range(3):
            ops.aggregate(xs, es, st, red, wl.model["msg_encoder"], beta_t, True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.aggregate(xs, es, st, red, wl.model["msg_encoder"], beta_t, True)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / 10
        byt = aggregation_bytes(bt.num_nodes, bt.num_edges, Fw, red == "softmax")
        unfused = {"kernel": "aggregate_fwd_kernel (edge embedding [E,F] read from HBM)", "avg_launch_us": us,
                   "achieved": byt / us / 1e3, "unit": "GB/s", "in_timed_step": False}
        del xs, es

    # ---- end-to-end: host batches, H2D inside the timed region, loss read back every step -------------
    e2e = None
    if not args.no_e2e:
        # Every step: its batch is copied from pinned host memory (side stream, one step ahead of the compute
        # stream) and its loss is copied back to pinned host memory; the host reads loss i while step i+1 is
        # already queued, so the host never stalls the device (the reference blocks on loss.item() each step).
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
        total_loss = 0.0

        fields = ("x", "edge_index", "edge_attr", "batch", "y")

        def fetch(i):
            # destination tensors come from the compute stream's allocator pool (no cross-stream frees); the copies
            # run on the side stream once the compute stream has reached this point, i.e. one step ahead
            import copy as _copy
            src = host[i % len(host)]
            d = _copy.copy(src)
            for name in fields:
                setattr(d, name, torch.empty_like(getattr(src, name), device=dev))
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ready)
                for name in fields:
                    getattr(d, name).copy_(getattr(src, name), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return d, ev

        def e2e_loop(n_steps, first):
            total = 0.0
            nxt = fetch(first)
            for i in range(n_steps):
                d, ev = nxt
                main.wait_event(ev)
                if i + 1 < n_steps:
                    nxt = fetch(first + i + 1)
                graph.clear_cache()
                l = step(d)
                loss_host[i & 1].copy_(l, non_blocking=True)
                loss_ev[i & 1].record(main)
                if i > 0:
                    loss_ev[(i - 1) & 1].synchronize()
                    total += float(loss_host[(i - 1) & 1]) * wl.batch_graphs
            loss_ev[(n_steps - 1) & 1].synchronize()
            return total + float(loss_host[(n_steps - 1) & 1]) * wl.batch_graphs

        e2e_loop(len(host) + 1, 0)        # untimed: the staging tensors of every batch shape exist in the allocator afterwards
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        total_loss = e2e_loop(args.steps, args.warmup)
        t1.record()
        barrier()
        tt = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": graphs / (float(tt.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(sum(b.nbytes() for b in host) / len(host)), "d2h_bytes_per_step": 4,
               "mean_loss": total_loss / (args.steps * wl.batch_graphs)}

    # ---- end-to-end with the DATASET resident in HBM: per step the host sends 3B+2 int64 (graph ids and
    #      their size prefix sums), the device assembles the mini-batch (prep.DeviceGraphStore.collate = Batch.from_data_list
    #      as one kernel), applies the per-batch transform of the reference's train() where it has one (RemoveIsolatedNodes for
    #      hiv / pcba, train_hiv.py:171-173; one host sync to read the kept sizes), steps, and the loss is read back ------------
    e2e_store = None
    if not args.no_e2e:
        from phc_gnn_b200.prep import DeviceGraphStore, RemoveIsolatedNodes
        from phc_gnn_b200.synthetic import split_graphs
        import numpy as np
        store = DeviceGraphStore([g for b in host for g in split_graphs(b)], dev)
        Bg = wl.batch_graphs
        ids_list = [np.arange(i * Bg, (i + 1) * Bg, dtype=np.int64) for i in range(len(host))]
        transform = RemoveIsolatedNodes() if wl.name in ("hiv", "pcba") else None
        loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_evs = [torch.cuda.Event(), torch.cuda.Event()]
        main_s = torch.cuda.current_stream(dev)

        def store_loop(n_steps, first):
            total = 0.0
            for i in range(n_steps):
                d = store.collate(ids_list[(first + i) % len(ids_list)])
                if transform is not None:
                    d = transform(d)
                graph.clear_cache()
                l = step(d)
                loss_pin[i & 1].copy_(l, non_blocking=True)
                loss_evs[i & 1].record(main_s)
                if i > 0:
                    loss_evs[(i - 1) & 1].synchronize()
                    total += float(loss_pin[(i - 1) & 1]) * Bg
            loss_evs[(n_steps - 1) & 1].synchronize()
            return total + float(loss_pin[(n_steps - 1) & 1]) * Bg

        store_loop(len(host) + 1, 0)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        store_total = store_loop(args.steps, args.warmup)
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_store = {"value": graphs / (float(ts.item()) / 1e3), "unit": UNIT, "collate": "device",
                     "h2d_bytes_per_step": 8 * (3 * Bg + 2), "d2h_bytes_per_step": 4 + (16 if transform is not None else 0),
                     "dataset_bytes_in_hbm": store.nbytes(), "transform": repr(transform) if transform is not None else None,
                     "mean_loss": store_total / (args.steps * Bg)}
        del store

    if rank == 0:
        peak, peak_src = peaks()
        N = sum(b.num_nodes for b in host) / len(host)
        E = sum(b.num_edges for b in host) / len(host)
        F = wl.model["mp_layers"][0]
        sums_path = "phc_conv_fused_fwd_sums" in prof
        fused = sums_path or "phc_conv_fused_fwd" in prof
        key = "phc_conv_fused_fwd_sums" if sums_path else ("phc_conv_fused_fwd" if fused else "phc_aggregate_fwd")
        calls, agg_ms = prof.get(key, (0, 0.0))
        agg = None
        if calls:
            softmax = wl.model["msg_aggr"] == "softmax"
            boundary = aggregation_bytes(N, E, F, softmax)            # SURVEY 8(d) unit: one layer's propagate at the reference's boundary
            us = 1e3 * agg_ms / calls
            dims = wl.model["bond_input_dims"]
            attr_b = (4 * dims if isinstance(dims, int) else 8 * len(dims)) * E
            rows = (dims + 1) if isinstance(dims, int) else sum(dims)
            if sums_path:      # x gathered + out + per-node feature sums + indices + the encoder table: what the kernel must move
                model_b, kname = 4 * F * 2 * N + 4 * rows * N + 4 * E + 4 * (N + 1) + 4 * rows * F, "conv_fwd_sums"
                kdesc = "conv_fwd_sums_kernel (row gather + per-node encoder term; + enc_table_kernel)"
            elif fused:
                model_b, kname = 4 * F * 2 * N + attr_b + 8 * E + 4 * (N + 1) + (8 * N * F if softmax else 0), "conv_fwd"
                kdesc = "conv_fwd_kernel (gather + edge ENCODER + edge add + reduce fused)"
            else:
                model_b, kname, kdesc = boundary, "aggregate_fwd", "aggregate_fwd_kernel (fused gather + edge add + reduce)"
            traffic, tsrc = traffic_record(kname + ":" + wl.name)
            agg = {"bound": "hbm", "kernel": kdesc, "achieved": model_b / us / 1e3, "peak": peak, "unit": "GB/s",
                   "frac": model_b / us / 1e3 / peak, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
                   "algorithmic_bytes_per_launch": model_b, "avg_launch_us": us, "share_of_step": agg_ms / ms,
                   "note": "achieved = the bytes this kernel must move (its own HBM model: x, out, indices and the raw edge features or "
                           "their per-node sums) / time.  The fused kernels never form the [E,F] edge embedding the reference's operator "
                           "boundary reads, so against the SURVEY 8(d) boundary bytes the same time is an EFFECTIVE rate "
                           "(effective_boundary_gbs, may exceed the HBM peak); they are bound by the L2 row gather (l2_gather_gbs = "
                           "4*E*F bytes / time).  The kernel that streams [E,F] from HBM at that boundary is unfused_boundary_kernel.",
                   "boundary_bytes_per_launch": boundary, "effective_boundary_gbs": boundary / us / 1e3,
                   "l2_gather_gbs": 4 * E * F / us / 1e3, "unfused_boundary_kernel": unfused,
                   "timed": "CUDA events around each launch, instrumented repeat of the K steps"}
            if unfused:
                unfused["frac"] = unfused["achieved"] / peak
        precision_name = os.environ.get("PHC_PRECISION", DEFAULT_PRECISION)
        roof = phm_linear_roofline(prof, args.steps, wl, N, precision_name, ms / args.steps)   # share of the headline (uninstrumented) step
        if roof is None:        # a workload without node-level tensor-core linears: the aggregation is the dominant kernel
            roof, agg = agg, None
        breakdown = {k: {"calls_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in sorted(prof.items())}
        if args.kernel_timers:
            print(json.dumps(breakdown, indent=1), file=sys.stderr)
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": {"fp32": "f32", "tf32x3": "f32 (tf32x3 tensor-core split, fp32 accumulate)",
                         "bf16x3": "f32 (bf16x3 tensor-core split: 16-bit-mantissa operands in three bf16 passes, fp32 accumulate)",
                         "bf16": "bf16"}[precision_name],
               "data": "synthetic",
               "config": config_dict(args, wl, wl.batch_graphs, "flushed between steps" if flush else "inputs larger than L2"),
               "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "e2e_device_store": e2e_store, "roofline": roof, "roofline_aggregation": agg,
               "preroll_steps": preroll, "cuda_graph": step.stats() if use_graph else None,
               "ms_per_step_instrumented": ms_instr / args.steps,
               "op_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in breakdown.items()},
               "final_loss": float(loss.item())}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(wl)
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Only the JSON line may appear on stdout: native libraries (NCCL prints its version banner there) are pointed at
    stderr by swapping the descriptors; the JSON is written to the saved original."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    global JSON_OUT
    JSON_OUT = _claim_stdout()
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
