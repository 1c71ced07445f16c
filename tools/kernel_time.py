"""Development timing: per-kernel device time (event pair per launch) of the hot path, ms per step.
usage: python tools/kernel_time.py [config] [steps]
  BLOM_OPTIONS=key=value,...        option set of the main measurement
  BLOM_ROUTINES=r1,r2,...           restrict the step to these routines (e.g. ndiff)
  BLOM_AB="k=v,k=v;k=v"             further option sets measured afterwards on the same resident state
                                    (each set is applied on top of the previous ones)"""
import os, sys, time, json
sys.path.insert(0, ".")
from blom_b200.driver import HotPath

cfg = sys.argv[1] if len(sys.argv) > 1 else "tnx0.25v4"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
t0 = time.time()
rts = [r for r in os.environ.get("BLOM_ROUTINES", "").split(",") if r] or None
hp = HotPath(cfg, ntr=0, nstep=1, routines=rts)
g = hp.gpu


def measure(label):
    for _ in range(2):
        hp.advance()
    g.sync()
    g.timers_enable(True); g.timers_reset()
    for _ in range(steps):
        hp.advance()
    g.sync()
    rt = g.timers(); g.timers_enable(False)
    g.ktimers_enable(True)
    for _ in range(steps):
        hp.advance()
    kt = g.ktimers(); g.ktimers_enable(False)
    out = {"config": cfg, "options": label, "setup_s": round(time.time() - t0, 1),
           "routines_ms": {k: round(v["ms"] / v["calls"], 3) for k, v in rt.items() if v["calls"]},
           "kernels_ms_per_step": {k: round(v["ms"] / steps, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["ms"])}}
    out["step_ms"] = round(sum(out["routines_ms"].values()), 3)
    print(json.dumps(out), flush=True)


measure(os.environ.get("BLOM_OPTIONS", ""))
for optset in [s for s in os.environ.get("BLOM_AB", "").split(";") if s]:
    for kv in optset.split(","):
        k, v = kv.split("=", 1)
        g.set_option(k, v)
    measure(optset)
