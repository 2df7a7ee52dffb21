"""The reference arm of bench.py (`--impl reference`) runs on host cores only: its JSON line carries the contract keys
(impl, the same metric / unit / config as the GPU arm, cpu_baseline, e2e with zero copies) and needs no GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", *extra],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, p.stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run()
    assert d["impl"] == "reference" and d["metric"] == "vap_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["cpu_baseline"]["value"] - d["value"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["baseline_config"] == 2 and d["config"]["global_streams"] == 64 and d["config"]["ctx_frames"] == 50
    assert "workload" in d["config"] and "batch=64" in d["config"]["workload"]


def test_reference_arm_other_config():
    d = _run("--config", "5")
    assert d["impl"] == "reference" and d["value"] > 0
    assert d["config"]["baseline_config"] == 5 and d["config"]["head"] == "bc" and d["config"]["ctx_frames"] == 100
