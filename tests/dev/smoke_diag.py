"""Development aid: per-routine error growth of the performance (FMA) build against the oracle on a
small case; prints, for each routine, the worst field, where it is and how thick the layer is there.
usage: python tests/dev/smoke_diag.py [config] [parity 0|1] [each|end]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util import Case, interior, prepare_step
from blom_b200.driver import run_step

cfg = sys.argv[1] if len(sys.argv) > 1 else "fuk95"
parity = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
sync_each = (sys.argv[3] != "end") if len(sys.argv) > 3 else True   # "end": download only after the last routine
c = Case(cfg, ntr=1, nstep=1)
o = c.new_oracle(); g = c.new_gpu(parity=parity)
routines, _ = prepare_step(c, (o, g))
kk = c.dims[2]
for r in routines:
    for b in (o, g):
        run_step(b, [r], c.levels)
    if not sync_each and r != "tmsmt2":
        continue
    g.download_all()
    rows = []
    for nm, a in g.arrays.items():
        if a.dtype != np.float64 or nm == "depths":
            continue
        A, B = interior(a), interior(o.arrays[nm])
        d = np.abs(A - B)
        s = np.abs(B).max()
        if s == 0 or not np.isfinite(d).all():
            continue
        e = d.max() / s
        if e > 1e-12:
            idx = np.unravel_index(np.argmax(d), d.shape)
            extra = ""
            if nm in ("temp", "saln", "trc") and idx[0] < 2 * kk:
                extra = f" dp_there={interior(o.arrays['dp'])[idx]:.3e} val={B[idx]:.6g} got={A[idx]:.6g}"
            rows.append((e, nm, idx, extra))
    rows.sort(reverse=True)
    print(f"{r:12s}", "; ".join(f"{nm} {e:.2e} @{idx}{x}" for e, nm, idx, x in rows[:4]) or "all <= 1e-12", flush=True)
g.finalize()
