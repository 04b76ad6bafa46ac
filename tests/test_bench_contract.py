"""CPU: the reference arm of bench.py (`--impl reference`: the oracle port of the reference's CPU path on the host cores)
must print exactly ONE JSON line on stdout with the keys the driver reads, and the same metric / unit / config as the GPU
arm.  (The GPU arm needs a B200; its line is checked by scripts/gpu_round.sh runs kept under profiles/.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--keyframes", "200"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0
    import bench
    assert d["metric"] == bench.METRIC and d["config"]["workload"] == bench.WORKLOAD


def test_latest_gpu_profile_line_carries_the_contract_keys():
    """the newest committed GPU bench line (profiles/bench_r02m.json) has what the contract asks of the GPU arm"""
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_r02m.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and 0.5 < r["frac"] <= 1.0      # north star: >= 50 % of the INT roofline
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"] * 1.01
    assert d["gpu_launches"] > 0 and d["clocks"]["reasons"] == []
    assert r["bound"] == "tensor" and r["traffic"] and d["sanity"]["cpu_records_equal"] is True and d["sanity"]["cpu_records_compared"] > 40000
    assert d["cpu_baseline"]["reference_faithful_1thread"]["value"] > 0 and d["adapter_queue"]["same_map_as_this_bench"] is True
    assert d["roofline"]["int_pipe_kernel"]["records_equal_tensor_core_path"] is True
    # the multi-GPU line of the same tree: gathered records equal one GPU's, also through the one-process group
    d8 = json.load(open(os.path.join(ROOT, "profiles", "bench_r02m_n8.json")))
    assert d8["n_gpus"] == 8 and d8["sanity"]["gathered_equals_single_gpu"] is True
    assert d8["group"]["records_equal_single_gpu"] is True and d8["group"]["peer_write"]["records_equal"] is True
    assert d8["value"] > 7.5 * d["value"]
