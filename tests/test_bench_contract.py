"""The driver-facing contract of bench.py's CPU arm (`--impl reference`): runs here without a GPU.  The arm times the
UNMODIFIED reference (oracle/_ref, a verbatim copy made by oracle/build_ref.py where /root/reference is mounted) or, when
that copy is absent, the plain-C port -- and must print ONE JSON line with the agreed keys either way."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-frames-per-proc", "12"], capture_output=True, text=True, timeout=600, env=env, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ion-frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("ion-frame assignments/sec") and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["vs_baseline"] is None
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if (ROOT / "oracle" / "_ref" / "site_analysis").is_dir():
        assert cb["kind"] == "reference", cb                      # the real thing whenever the copy travelled with the tree
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
