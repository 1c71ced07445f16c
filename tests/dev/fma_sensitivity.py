"""How much of the FMA (performance) build's deviation from the oracle is the algorithm's own
sensitivity to FP contraction?  Builds a second copy of the ORACLE with -ffp-contract=fast -mfma
(into oracle/_build/fma/, test infrastructure only) and runs one chained step on both: the two CPU
builds differ from each other in the same few nearly massless cells, by the same values, as
libblomgpu.so differs from the oracle on the GPU (profiles/r01_fma_sensitivity.txt).
usage: python tests/dev/fma_sensitivity.py [config]"""
import subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import oracle.oracle as om
from util import Case, interior
from blom_b200.driver import STEP_SEQUENCE

out = ROOT / "oracle" / "_build" / "fma"
out.mkdir(parents=True, exist_ok=True)
objs = []
for src in sorted((ROOT / "oracle").glob("*.cpp")):
    obj = out / (src.stem + ".o")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=fast", "-mfma", "-mavx2",
                    "-c", str(src), "-o", str(obj)], check=True)
    objs.append(str(obj))
subprocess.run(["g++", "-shared", "-fopenmp", "-o", str(out / "liboracle_fma.so"), *objs], check=True)

cfg = sys.argv[1] if len(sys.argv) > 1 else "fuk95"
c = Case(cfg, ntr=1, nstep=1)
o = c.new_oracle()
om.LIB = out / "liboracle_fma.so"
f = c.new_oracle()
m, n, mm, nn, k1m, k1n = c.levels
for b in (o, f):
    b.inieos(); b.numerical_bounds(); b.init_cppm()
    for r in STEP_SEQUENCE:
        if r == "tmsmt1":
            b.tmsmt1(nn)
        elif r == "tmsmt2":
            b.tmsmt2(m, mm, nn, k1m)
        else:
            getattr(b, r)(m, n, mm, nn, k1m, k1n)
print(f"{cfg}: oracle(-ffp-contract=fast) vs oracle(-ffp-contract=off) after one chained step")
for nm, a in f.arrays.items():
    if a.dtype != np.float64 or nm == "depths":
        continue
    A, B = interior(a), interior(o.arrays[nm])
    s = np.abs(B).max()
    if s == 0:
        continue
    d = np.abs(A - B) / s
    if d.max() > 1e-12:
        idx = tuple(int(t) for t in np.unravel_index(np.argmax(d), d.shape))
        print(f"  {nm:10s} max {d.max():.2e} at {idx}   points > 1e-9: {int((d > 1e-9).sum())} of {d.size}")
