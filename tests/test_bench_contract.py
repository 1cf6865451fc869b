"""CPU: the driver contract of bench.py's reference arm (`--impl reference`): one JSON line on stdout with the same
metric / unit / config as the GPU arm, `impl`, a `cpu_baseline` describing the run and an `e2e` that repeats it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "grid-points/s"
    assert d["metric"].startswith("grid-points/sec fwd+bwd, Darcy 141")
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["scaling"] == "weak"
    assert d["config"]["workload"] == "darcy141_galerkin10_sc2d_b8" and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 8 * 141 * 141 / (d["ms_per_step"] * 1e-3)) < 1e-3 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == dict(value=d["value"], unit=d["unit"], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
