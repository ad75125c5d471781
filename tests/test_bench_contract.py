"""bench.py's output contract on the CPU arm (``--impl reference`` = the unmodified reference from baseline/_ref on the host cores,
or the oracle port when it is not installed): exactly one JSON line on stdout carrying the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "zinc", "--steps", "1",
                           "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [ln for ln in proc.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, proc.stdout
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["metric"] == "train_graphs_per_sec" and rec["unit"] == "graphs/s"
    assert rec["higher_is_better"] is True and rec["steps"] == 1 and rec["warmup"] == 0 and rec["n_gpus"] == 1
    assert rec["value"] > 0 and rec["ms_per_step"] > 0 and rec["vs_baseline"] is None and rec["data"] == "synthetic"
    want_kind = "reference" if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "phc", "hypercomplex", "undirectional", "models.py")) else "port"
    assert rec["cpu_baseline"]["kind"] == want_kind and rec["cpu_baseline"]["cores"] >= 1 and rec["cpu_baseline"]["value"] == rec["value"]
    assert rec["config"]["graphs_per_gpu_batch"] == 128          # the full per-GPU batch of the zinc workload, same as the b200 arm
    assert rec["e2e"] == {"value": rec["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in rec["config"] and "model" not in rec["config"]


def test_other_ranks_of_the_reference_arm_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "zinc",
                           "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert proc.returncode == 0, proc.stderr[-2000:]
    assert proc.stdout.strip() == ""


def test_tensor_roofline_arithmetic():
    """roofline (PHMLinear): algorithmic FLOPs of the node-level linears over their measured time (pure host arithmetic)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    from phc_gnn_b200.synthetic import workloads
    wl = workloads(4)["ppa"]                                   # 7 layers x 2 MLP linears + the pooling linear = 15, F = 500
    steps, N = 4, 10000
    unit = 2.0 * N * 500 * 500
    prof = {"phc_phm_linear_fwd:node": (15 * steps, 1.0 * steps), "phc_phm_linear_bwd:node": (15 * steps, 2.0 * steps),   # 1 ms fwd, 2 ms bwd per step
            "phc_phm_linear_fwd:head": (3 * steps, 0.1 * steps)}
    r = bench.phm_linear_roofline(prof, steps, wl, N, "tf32x3", step_ms=6.0)
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["node_level_linears"] == 15
    assert abs(r["forward_launch"]["tflops"] - 15 * unit / 1e-3 / 1e12) < 1e-6
    assert abs(r["backward_call"]["tflops"] - 15 * 2 * unit / 2e-3 / 1e12) < 1e-6
    assert abs(r["achieved"] - 15 * 3 * unit / 3e-3 / 1e12) < 1e-6 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["precision_ceiling_tflops"] - r["peak"] / 6) < 1e-9 and abs(r["share_of_step"] - 0.5) < 1e-9
    assert abs(bench.phm_linear_roofline(prof, steps, wl, N, "bf16x3")["precision_ceiling_tflops"] - r["peak"] / 3) < 1e-9
    assert bench.phm_linear_roofline({}, steps, wl, N, "tf32x3") is None
    assert bench.phm_linear_roofline(prof, steps, wl, N, "fp32")["frac_of_precision_ceiling"] is None
