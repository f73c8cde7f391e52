"""bench.py's output contract, as far as it can run without a GPU: the reference arm (`--impl reference`) executes the
reference's own CPU estimateGround and must print ONE JSON line with the keys the driver reads; the product arm must
refuse to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--frames-per-gpu", "4",
                          "--ref-frames-per-step", "4"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("frames/sec") and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["steps"] == 1 and j["n_gpus"] == 1 and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and abs(j["cpu_baseline"]["value"] - j["value"]) < 1e-9
    assert j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0 and "workload" in j["config"]


@pytest.mark.skipif(not _no_gpu(), reason="a CUDA device is present")
def test_product_arm_refuses_to_run_without_a_gpu():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1", "--warmup", "1", "--frames-per-gpu", "2", "--no-e2e", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stderr + out.stdout) and not [l for l in out.stdout.splitlines() if l.startswith("{")]
