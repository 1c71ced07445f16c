"""Development timing of individual routines on the GPU (not the contract bench)."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
from blom_b200.driver import HotPath

cfg = sys.argv[1] if len(sys.argv) > 1 else "tnx1v4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
t0 = time.time()
hp = HotPath(cfg, ntr=0, nstep=1)
print("setup %.1fs" % (time.time() - t0), "routines", hp.routines, flush=True)
g = hp.gpu
g.timers_enable(True)
for it in range(reps + 2):
    if it == 2:
        g.timers_reset()
    hp.set_step(1 + it)
    hp.step()
g.sync()
cells = hp.itdm * hp.jtdm * hp.kdm
for k, v in g.timers().items():
    ms = v["ms"] / v["calls"]
    print(f"{k:10s} {ms:8.3f} ms/call  launches/call {v['launches']/v['calls']:.0f}  -> {cells*8/ms/1e6:.1f} GB/s per word/cell")
g.download_all()
print("dp finite:", np.isfinite(hp.arrays["dp"]).all(), "min", hp.arrays["dp"].min())
