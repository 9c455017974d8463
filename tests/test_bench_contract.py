"""bench.py contract checks that run without a GPU: the reference arm (`--impl reference` = the oracle port of ME's
CPU algorithm on the host cores) prints exactly one JSON line with the keys the driver reads, and ranks other than 0
print nothing."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--cpu-voxels", "8000", "--steps", "1",
                        "--warmup", "0", "--gpus", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return [l for l in p.stdout.splitlines() if l.strip()]


def test_reference_arm_prints_one_json_line():
    lines = _run()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["higher_is_better"] is True
    assert d["metric"] == "MinkUNet34C fwd+bwd voxels/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "not installable" in d["cpu_baseline"]["sample"]  # never labelled "MinkowskiEngine CPU"


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
