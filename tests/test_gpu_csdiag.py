"""Bit-exact parity of the whole hot-path chain in the reference's own diagnostic format: the
' chksum: <name>: 0x%08X' lines that BLOM prints with CSDIAG = .true. (phy/mod_checksum.F90:41-74) after
every routine, i.e. the CRC-32 of every masked field the routine wrote.  tools/csdiag_log.py produces them for
the analytic fuk95 case; tests/golden/fuk95_csdiag.txt is the oracle's log (3 steps, 293 checksums).  The CUDA
path (parity build, through the C ABI) must print the identical text: one flipped bit anywhere in a field
changes its CRC.  (The same log is what a maintainer diffs against the Fortran reference, oracle/_ref/README.md.)"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))
sys.path.insert(0, str(ROOT / "tests" / "dev"))
GOLDEN = ROOT / "tests" / "golden" / "fuk95_csdiag.txt"


def test_oracle_reproduces_committed_log():
    import csdiag_cpu
    import io
    lines = csdiag_cpu.run(steps=3, out=io.StringIO())
    assert lines == GOLDEN.read_text().splitlines()


@pytest.mark.gpu
def test_gpu_log_is_bit_identical_to_golden():
    import csdiag_log
    import io
    lines = csdiag_log.run(steps=3, out=io.StringIO())
    gold = GOLDEN.read_text().splitlines()
    diff = [(i, a, b) for i, (a, b) in enumerate(zip(lines, gold)) if a != b]
    assert len(lines) == len(gold) and not diff, diff[:8]
