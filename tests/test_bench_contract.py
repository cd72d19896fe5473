"""CPU checks of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys, the clock summary parses
nvidia-smi rows (and survives having none)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] and d["unit"] == "episodes/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_clock_summary_parsing():
    sys.path.insert(0, ROOT)
    import bench
    clk = bench.ClockSampler.__new__(bench.ClockSampler)
    clk.rows = [["1665", "1965", "Not Active", "Not Active", "Not Active", "Active"],
                ["1695", "1965", "Not Active", "Not Active", "Not Active", "Active"],
                ["1965", "1965", "Not Active", "Not Active", "Not Active", "Not Active"]]
    s = clk.summary()
    assert s == {"sm_mhz": 1695, "sm_max_mhz": 1965, "reasons": ["sw_power_cap"], "samples": 3}
    clk.rows = []
    assert clk.summary() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    assert bench.ClockSampler._stamp("2026/10/17 12:34:56.789") is not None and bench.ClockSampler._stamp("garbage") is None
