"""Generates tests/golden/xctilr_maps.json: for every region type and grid/field
type, the SOURCE cell (and sign) each halo cell must receive from xctilr with
(mh,nh)=(4,4), derived independently of oracle/xc.cpp from the index rules
documented in SURVEY.md §8(a2) (serial semantics, phy/mod_xc.F90:4262-4419):

  closed edges -> vland (0);  periodic edges -> wrapped image;
  arctic fold (nreg=2), with m(i)=ii+1-i and u(i)=ii+2-i (u(1)=1):
    p: (i,jj+j) <- ( m(i), jj-1-j), j=0..nh      u: same rows with u(i)
    v: (i,jj+j) <- ( m(i), jj-j  ), j=1..nh, and (i,jj) <- (m(i),jj) for i>ii/2
    q: as v with u(i);    vector types (11..14) change sign.
  E/W halos are filled after N/S ones (so corners see N/S results).

The expectation is produced by *pulling* every cell through these rules with a
recursive resolver (a cell whose source is itself a rewritten cell resolves
through it), not by replaying the reference's loops.
Run:  python tests/golden/make_xctilr_golden.py
"""
import json
from pathlib import Path

II, JJ, NB = 12, 10, 4


def resolve(nreg, itype, i, j):
    """-> (si, sj, sign) of the interior source of cell (i,j) after xctilr(4,4), or None for land."""
    it, vec = itype % 10, itype >= 10
    sign = 1
    # E/W first (it is applied last, so it is the outermost rule)
    if i < 1 or i > II:
        if nreg in (0, 4):
            return None
        i = i + II if i < 1 else i - II
    # N/S
    if nreg == 2:
        if j < 1:
            return None
        if it in (1, 3):
            if j >= JJ:
                jh = j - JJ
                i = (II + 1 - i) if it == 1 else (1 if i == 1 else II + 2 - i)
                j = JJ - 1 - jh
                sign = -1 if vec else 1
        else:
            mi = (II + 1 - i) if it == 4 else (1 if i == 1 else II + 2 - i)
            if j > JJ:
                i, j = mi, JJ - (j - JJ)
                sign = -1 if vec else 1
            elif j == JJ and i >= II // 2 + 1:
                i = mi
                sign = -1 if vec else 1
    elif nreg in (0, 1):
        if j < 1 or j > JJ:
            return None
    else:
        if j < 1:
            j += JJ
        elif j > JJ:
            j -= JJ
    return i, j, sign


def main():
    out = {"ii": II, "jj": JJ, "nbdy": NB, "maps": {}}
    for nreg in range(5):
        for itype in (1, 2, 3, 4, 11, 12, 13, 14):
            cells = []
            for j in range(1 - NB, JJ + NB + 1):
                for i in range(1 - NB, II + NB + 1):
                    r = resolve(nreg, itype, i, j)
                    cells.append([0, 0, 0] if r is None else list(r))
            out["maps"][f"{nreg}_{itype}"] = cells
    Path(__file__).with_name("xctilr_maps.json").write_text(json.dumps(out))


if __name__ == "__main__":
    main()
