"""The bench.py contract for the reference arm (runs on the host cores, no GPU): one JSON line with the keys the
driver reads, the same metric / unit / config as the GPU arm, zero copy bytes, and a cpu_baseline that describes
the run itself."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liba52_ref.so"))


@pytest.mark.parametrize("workload", ["decode", "encode"])
def test_reference_arm_json_line(workload):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-seconds", "1", "--workload", workload],
                       capture_output=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.decode().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert ("decoded" if workload == "decode" else "encoded") in d["metric"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["sample"]
    assert cb["kind"] == ("reference" if HAVE_REF else "port")
