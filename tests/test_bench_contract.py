"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm runs on the host
cores alone, prints exactly ONE JSON line on stdout and carries every key the driver reads; our arm fails loudly
when there is no CUDA device (no CPU fallback)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _util import has_cuda  # noqa: E402

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3", "--ref-batch", "64"],
                         capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("MPC solves/sec") and d["value"] > 0 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["config"]["workload"].startswith("cfg2_thing_demo") and d["data"] == "synthetic" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["reference_per_step_batch"] == 64 and "64 instances" in cb["sample"]   # the label says what was timed


@pytest.mark.skipif(has_cuda(), reason="checks the no-device behaviour")
def test_our_arm_fails_loudly_without_a_gpu():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=600, cwd=str(ROOT))
    assert out.returncode != 0 and out.stdout.strip() == ""
