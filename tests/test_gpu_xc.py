"""GPU parity (through the C ABI) of the mod_xc kernels: bit-exact halo/fold,
strip-ordered xcsum, CRC checksums, exact max/min, bigrid masks."""
import json
from pathlib import Path

import numpy as np
import pytest

from blom_b200.lib import BlomGpu
from oracle.oracle import Oracle
from util import Case, interior

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "xctilr_maps.json").read_text())


def coded(ii, jj, nb, nlev):
    k, j, i = np.meshgrid(np.arange(nlev), np.arange(jj + 2 * nb), np.arange(ii + 2 * nb), indexing="ij")
    return (1e6 * (k + 1) + 1000.0 * (j + 7) + (i + 7) + 0.5).astype(np.float64)


@pytest.mark.parametrize("nreg", [0, 1, 2, 3, 4])
def test_xctilr_golden_maps_and_oracle(nreg):
    ii, jj, nb = GOLD["ii"], GOLD["jj"], GOLD["nbdy"]
    g = BlomGpu(ii, jj, 3, nreg)
    o = Oracle(ii, jj, 3, nreg)
    try:
        for itype in (1, 2, 3, 4, 11, 12, 13, 14):
            for (mh, nh) in [(4, 4), (1, 1), (4, 0), (0, 4), (2, 3), (3, 2), (0, 0), (9, 9)]:
                for (l1, ld, koff) in [(1, 3, 1), (2, 3, 1), (1, 2, 2)]:
                    a = coded(ii, jj, nb, 3); b = a.copy()
                    g.register("a", a); o.register("a", b)
                    g.xctilr("a", l1, ld, mh, nh, itype, koff=koff)
                    o.xctilr("a", l1, ld, mh, nh, itype, koff=koff)
                    g.download("a")
                    assert np.array_equal(a.view(np.int64), b.view(np.int64)), (nreg, itype, mh, nh, l1, ld, koff)
            # golden map, independent of the oracle
            a = coded(ii, jj, nb, 3); src = a.copy()
            g.register("a", a)
            g.xctilr("a", 1, 3, nb, nb, itype)
            g.download("a")
            cells = np.array(GOLD["maps"][f"{nreg}_{itype}"]).reshape(jj + 2 * nb, ii + 2 * nb, 3)
            si, sj, sg = cells[..., 0], cells[..., 1], cells[..., 2]
            want = np.where(sg == 0, 0.0, sg * src[:, np.clip(sj + nb - 1, 0, None), np.clip(si + nb - 1, 0, None)])
            assert np.array_equal(a, want), (nreg, itype)
    finally:
        g.finalize()


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4", "fuk95"])
def test_bigrid_sum_crc_minmax(cfg):
    c = Case(cfg)
    g = c.new_gpu(parity=False)
    o = c.new_oracle()
    try:
        assert g.nreg == o.nreg == c.dims[3]
        for nm in ("ip", "iu", "iv", "iq"):
            got = g.fetch(nm, 1, np.int32)[0]
            assert np.array_equal(got, c.masks[nm]), nm
        kk = c.dims[2]
        for nm, msk, it in (("dp", "ip", 1), ("u", "iu", 13), ("v", "iv", 14), ("corioq", "iq", 2)):
            nlev = o.arrays[nm].shape[0]
            for lev in (1, nlev):
                s_g, s_o = g.xcsum(nm, msk, lev), o.xcsum(nm, msk, lev)
                assert np.float64(s_g).view(np.int64) == np.float64(s_o).view(np.int64), (nm, lev, s_g, s_o)
                m = interior(c.masks[msk]) == 1
                vals = interior(o.arrays[nm][lev - 1])[m]
                assert g.xcmax(nm, msk, lev) == vals.max()
                assert g.xcmin(nm, msk, lev) == vals.min()
            assert g.chksum(nm, nlev, it) == o.chksum(nm, nlev, it), nm
        # a one-bit change flips the checksum
        a = g.arrays["dp"]; jm, im = np.argwhere(interior(c.masks["ip"]) == 1)[3]
        a[0, jm + 4, im + 4] = np.nextafter(a[0, jm + 4, im + 4], np.inf)
        g.upload("dp")
        assert g.chksum("dp", 2 * kk, 1) != o.chksum("dp", 2 * kk, 1)
    finally:
        g.finalize()
