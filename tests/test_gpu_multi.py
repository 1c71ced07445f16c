"""Multi-GPU (j-band) parity, run on the box when enough GPUs are visible: the bands of an N-rank run are
bit-identical to the one-tile run (fields, strip-ordered xcsum, CRC) and within 1e-10 of the oracle
(tests/dev/mgpu_parity.py).  2 ranks on the small tripolar / periodic grids with both transports, 3 and 4 ranks
(ranks with a neighbour on both sides, the in-kernel barotropic exchange in both directions), and 2, 4 and 8
bands of the full tnx1v4 grid (360 x 385 x 53) under the reference's option set."""
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


CASES = [("mid1", "p2p", 2, 3), ("mid2", "p2p", 2, 3), ("mid2", "nccl", 2, 3), ("mid2", "p2p", 3, 3),
         ("mid2", "p2p", 4, 3), ("mid1", "nccl", 4, 2), ("tnx1v4", "p2p", 2, 1), ("tnx1v4", "p2p", 4, 1),
         ("tnx1v4", "p2p", 8, 1)]


@pytest.mark.parametrize("cfg,comm,nranks,steps", CASES)
def test_band_parity(cfg, comm, nranks, steps, tmp_path):
    if _ngpu() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, MGPU_TMP=str(tmp_path), MGPU_COMM=comm)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests/dev/mgpu_parity.py"),
                        cfg, str(steps)], capture_output=True, text=True, env=env, timeout=1200)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["ok"], out
