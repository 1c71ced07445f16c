"""Development check (uses the test scaffolding): cppm_flux j-pass tile shapes 32x16 and 64x8 give
bit-identical advect results.  usage: python tests/dev/jtile_check.py"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
from util import Case, interior
for cfg in ("tiny2", "tiny4", "fuk95", "mid2"):
    for nstep in (1, 2):
        c = Case(cfg, ntr=1, nstep=nstep)
        res = []
        for tile in ("32x16", "64x8", "32x8"):
            g = c.new_gpu(parity=True)
            g.set_option("cppm_j_tile", tile)
            g.inieos(); g.numerical_bounds(); g.init_cppm()
            g.advect(*c.levels)
            g.download_all()
            res.append({k: v.copy() for k, v in g.arrays.items() if v.dtype == np.float64})
            g.finalize()
        bad = [k for k in res[0] if not (np.array_equal(res[0][k], res[1][k]) and np.array_equal(res[0][k], res[2][k]))]
        print(cfg, nstep, "identical" if not bad else ("DIFF", bad))
