"""Pins the oracle's mod_xc restatement (CPU): CRC-32 check value, fold/halo index
maps against the committed golden fixture, strip-ordered xcsum, bigrid."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle.oracle import Oracle
from util import Case, interior

GOLD = json.loads((Path(__file__).parent / "golden" / "xctilr_maps.json").read_text())


def test_crc32_check_value():
    o = Oracle(12, 10, 2, 0)
    assert o.crc32(b"123456789") == 0xCBF43926
    assert o.crc32(b"") == 0
    # chaining == concatenation
    assert o.crc32(b"6789", o.crc32(b"12345")) == 0xCBF43926


def coded(ii, jj, nb, nlev=1):
    a = np.zeros((nlev, jj + 2 * nb, ii + 2 * nb))
    for k in range(nlev):
        for j in range(1 - nb, jj + nb + 1):
            for i in range(1 - nb, ii + nb + 1):
                a[k, j + nb - 1, i + nb - 1] = 1e6 * (k + 1) + 1000.0 * (j + 10) + (i + 10) + 0.5
    return a


@pytest.mark.parametrize("nreg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("itype", [1, 2, 3, 4, 11, 12, 13, 14])
def test_xctilr_matches_golden_maps(nreg, itype):
    ii, jj, nb = GOLD["ii"], GOLD["jj"], GOLD["nbdy"]
    o = Oracle(ii, jj, 2, nreg)
    a = coded(ii, jj, nb, 2)
    src = a.copy()
    o.register("a", a)
    o.xctilr("a", 1, 2, nb, nb, itype)
    cells = GOLD["maps"][f"{nreg}_{itype}"]
    n = 0
    for j in range(1 - nb, jj + nb + 1):
        for i in range(1 - nb, ii + nb + 1):
            si, sj, sg = cells[n]; n += 1
            for k in range(2):
                got = a[k, j + nb - 1, i + nb - 1]
                want = 0.0 if sg == 0 else sg * src[k, sj + nb - 1, si + nb - 1]
                assert got == want, (nreg, itype, i, j, k, got, want)


@pytest.mark.parametrize("mh,nh", [(0, 0), (1, 1), (4, 0), (0, 4), (2, 3), (9, 9)])
def test_xctilr_partial_widths_touch_only_requested_halo(mh, nh):
    ii, jj, nb = 12, 10, 4
    for nreg in (1, 3):
        o = Oracle(ii, jj, 1, nreg)
        a = coded(ii, jj, nb)
        before = a.copy()
        o.register("a", a)
        o.xctilr("a", 1, 1, mh, nh, 1)
        mhl, nhl = min(mh, nb), min(nh, nb)
        changed = a != before
        allowed = np.zeros_like(changed)
        J = slice(nb - nhl, nb + jj + nhl)
        allowed[0, J, nb - mhl:nb] = True
        allowed[0, J, nb + ii:nb + ii + mhl] = True
        allowed[0, nb - nhl:nb, nb:nb + ii] = True
        allowed[0, nb + jj:nb + jj + nhl, nb:nb + ii] = True
        assert not (changed & ~allowed).any()


def test_xctilr_arctic_l1_quirk():
    """serial arctic code: N/S phase honours l1, E/W phase always starts at level 1
    (phy/mod_xc.F90:4265 vs :4363)."""
    ii, jj, nb = 12, 10, 4
    o = Oracle(ii, jj, 2, 2)
    a = coded(ii, jj, nb, 2)
    before = a.copy()
    o.register("a", a)
    o.xctilr("a", 2, 2, 2, 2, 1)
    assert (a[0, nb:nb + jj, :nb - 2] == before[0, nb:nb + jj, :nb - 2]).all()
    assert (a[0, nb:nb + jj, nb - 2:nb] == before[0, nb:nb + jj, nb + ii - 2:nb + ii]).all()  # E/W done on level 1
    assert (a[0, nb - 2:nb, nb:nb + ii] == before[0, nb - 2:nb, nb:nb + ii]).all()  # N/S not done on level 1
    assert (a[1, nb - 2:nb, nb:nb + ii] == 0).all()


def test_xcsum_strip_order():
    rng = np.random.default_rng(5)
    ii, jj, nb = 40, 11, 4
    o = Oracle(ii, jj, 1, 1)
    a = np.zeros((1, jj + 2 * nb, ii + 2 * nb))
    a[0, nb:-nb, nb:-nb] = rng.standard_normal((jj, ii)) * 10.0 ** rng.integers(-8, 8, (jj, ii))
    mask = np.zeros((1, jj + 2 * nb, ii + 2 * nb), dtype=np.int32)
    mask[0, nb:-nb, nb:-nb] = rng.integers(0, 2, (jj, ii))
    o.register("a", a); o.register("msk", mask)
    got = o.xcsum("a", "msk")
    tot = None
    for j in range(jj):
        row = 0.0
        for i1 in range(0, ii, 9):
            s = 0.0
            for i in range(i1, min(i1 + 9, ii)):
                if mask[0, nb + j, nb + i] == 1:
                    s = s + a[0, nb + j, nb + i]
            row = row + s
        tot = row if tot is None else tot + row
    assert got == tot
    assert abs(got - a[0][mask[0] == 1].sum()) <= 1e-6 * np.abs(a).max()


@pytest.mark.parametrize("cfg", ["tiny0", "tiny1", "tiny2", "tiny3", "tiny4"])
def test_bigrid_masks(cfg):
    c = Case(cfg)
    ip, iu, iv, iq = (c.masks[k] for k in ("ip", "iu", "iv", "iq"))
    assert c.nreg == c.dims[3]
    depth = c.grid["depths"][0]
    # bigrid refreshes the depth halo itself: compare on the interior + resolved halo via ip
    assert ((interior(depth) > 0) == (interior(ip) == 1)).all()
    I = (slice(4, -4), slice(4, -4))
    assert (iu[I] == ip[I] * ip[4:-4, 3:-5]).all()
    assert (iv[I] == ip[I] * ip[3:-5, 4:-4]).all()
    q4 = ip[I] * ip[4:-4, 3:-5] * ip[3:-5, 4:-4] * ip[3:-5, 3:-5]
    diag = (ip[I] * ip[3:-5, 3:-5]) | (ip[4:-4, 3:-5] * ip[3:-5, 4:-4])
    assert (iq[I] == (q4 | diag)).all()
    assert ip.sum() > 0.3 * ip.size * 0.5
    # span tables agree with the mask
    o = c.new_oracle()
    ifp, ilp, isp = o.get_int("ifp").reshape(-1, 100), o.get_int("ilp").reshape(-1, 100), o.get_int("isp")
    nb = 4
    for jrow in range(ip.shape[0]):
        m = np.zeros(ip.shape[1], dtype=np.int32)
        for l in range(isp[jrow]):
            m[ifp[jrow, l] + nb - 1: ilp[jrow, l] + nb] = 1
        assert (m == ip[jrow]).all()
