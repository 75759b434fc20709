"""bench.py's reference arm (`--impl reference`: the reference's own CPU code for the path, oracle/_ref) runs without a GPU: one
JSON line on stdout carrying the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "frames/s" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] >= 1 and line["value"] > 0 and abs(line["value"] - 1000.0 / line["ms_per_step"]) < 1e-6 * line["value"] + 1e-9
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and isinstance(cb["sample"], str)
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference" and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libshs_lightcull_ref.so")):
        assert "cull_lights_tiled" in cb["sample"]
