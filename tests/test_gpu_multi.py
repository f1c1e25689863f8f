"""Origin-sharded run on two real GPUs (torchrun, NCCL + symmetric-memory multicast) against the
single-GPU engine and the CPU oracle.  Skipped on boxes with fewer than two GPUs."""

import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [4099, 32768])  # 4099: not a multiple of 4 (padded slice of the last rank)
def test_two_gpu_sharded_run_matches_single_gpu_and_oracle(tmp_path, n):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "report.json"
    cmd = [
        sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
        "--master-port", str(_free_port()), str(REPO / "tests" / "mgpu_worker.py"), str(out), str(n),
    ]
    proc = subprocess.run(cmd, cwd=REPO, capture_output=True, text=True, timeout=900)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-6000:]
    report = json.loads(out.read_text())
    for mode in ("reader", "sliced"):
        r = report[mode]
        assert r["rows_identical_to_single_gpu"] and r["state_identical_to_single_gpu"], r
        assert r["rows_match_oracle"] and r["passage_mismatch_vs_oracle"] == 0, r
    print(json.dumps(report))
