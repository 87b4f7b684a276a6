"""bench.py's reference arm (CPU only) prints exactly one JSON line with the keys the driver reads."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(*args):
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", *args], capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


def test_reference_arm_headline_workload():
    d = _run("--steps", "1", "--warmup", "0", "--cpu-rows", "512")
    assert KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "obs_x_view_x_K_updates_per_s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_config1_uses_the_compiled_reference_when_present():
    d = _run("--workload", "c1", "--steps", "60")
    assert KEYS <= set(d) and d["unit"] == "sweeps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] == 1
