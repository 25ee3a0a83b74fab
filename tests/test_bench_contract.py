"""bench.py's output contract on the reference arm (CPU only): stdout is exactly one JSON line carrying the keys the
driver reads; everything else (library banners, progress) goes to stderr.  The engine arm shares main() and the
descriptor routing, and is exercised on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

from oracle import skirtref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not skirtref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-packages", "1000", "--nlambda", "5"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "packets/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert "workload" in line["config"] and "model" not in line["config"]
