"""The reference arm of ``bench.py`` (``--impl reference``: the CPU oracle on the host cores) runs without a
GPU and prints one JSON line with the keys the driver reads."""

import json
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--molecules", "3000", "--steps", "2", "--warmup", "1"],
        capture_output=True, text=True, timeout=600, cwd=str(REPO),
    )
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "updates/s"
    assert line["metric"].startswith("molecule*origin updates/sec")
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["steps"] == 2
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("configs[3]") and line["vs_baseline"] is None


def test_gpu_arm_fails_loudly_without_a_device():
    """No CPU fallback: without CUDA the GPU arm exits with an error instead of measuring anything."""
    import torch

    if torch.cuda.is_available():
        return
    out = subprocess.run(
        [sys.executable, str(REPO / "bench.py"), "--steps", "1", "--warmup", "3", "--molecules", "1000", "--origins", "2"],
        capture_output=True, text=True, timeout=600, cwd=str(REPO),
    )
    assert out.returncode != 0
    assert "CUDA" in (out.stderr + out.stdout)
