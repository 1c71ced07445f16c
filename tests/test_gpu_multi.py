"""Multi-GPU (j-band) parity, run on the box when >= 2 GPUs are visible: the bands of a 2-rank NCCL
run are bit-identical to the one-tile run and within 1e-10 of the oracle (tests/dev/mgpu_parity.py)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("cfg,comm", [("mid1", "p2p"), ("mid2", "p2p"), ("mid2", "nccl")])
def test_two_band_parity(cfg, comm, tmp_path):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MGPU_TMP=str(tmp_path), MGPU_COMM=comm)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests/dev/mgpu_parity.py"),
                        cfg, "3"], capture_output=True, text=True, env=env, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"], out
