"""The csdiag log of tools/csdiag_log.py produced by the ORACLE (CPU restatement, test infrastructure):
the generator of tests/golden/fuk95_csdiag.txt.

    python tests/dev/csdiag_cpu.py --steps 3 > tests/golden/fuk95_csdiag.txt
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))


def run(steps=3, config="fuk95_analytic", ntr=1, options=None, out=sys.stdout):
    from csdiag_log import BLOCKS, csdiag_lines
    from util import Case, prepare_step
    from blom_b200.driver import run_step
    from blom_b200.lib import time_levels
    c = Case(config, ntr=ntr, nstep=1)
    o = c.new_oracle()
    routines, _ = prepare_step(c, (o,), {"ltedtp": "layer", **(options or {})})
    kk = c.dims[2]
    lines = []
    for nstep in range(1, steps + 1):
        lv = time_levels(nstep, kk)
        o.set_scalar("nstep", nstep)
        lines.append(f" step {nstep:6d}")
        for r in routines:
            run_step(o, [r], lv)
            if r in BLOCKS:
                lines += csdiag_lines(o, r, lv, kk, ntr)
    for ln in lines:
        print(ln, file=out)
    return lines


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    run(a.steps)
