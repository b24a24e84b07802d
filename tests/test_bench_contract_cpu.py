"""CPU: the reference arm of bench.py runs here (it is the CPU port of the reference path) and
prints the JSON line the driver parses."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--ladder-max", "10000"], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "hybrid_top10_qps" and d["unit"] == "queries/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["rows"] == 10_000_000
    cb = d["cpu_baseline"]
    assert [r["rows"] for r in cb["ladder"]] == [10_000] and cb["fit"]["extrapolated"] is True
    assert cb["blas_threads"] is None or cb["blas_threads"] >= 1
