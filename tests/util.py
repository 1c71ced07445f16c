"""Shared test scaffolding: builds one synthetic case and hands identical copies
to the oracle (CPU restatement, test infrastructure) and to the CUDA library."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from blom_b200 import synth  # noqa: E402
from blom_b200.lib import BlomGpu, time_levels  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


class Case:
    """Synthetic case prepared with the oracle's xctilr/bigrid (CPU)."""

    def __init__(self, config="tiny2", ntr=0, nstep=1, land=True, metric="tripolar", seed=20240611,
                 isopycnic=False, **synth_kw):
        itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
        self.config = config
        self.dims = (itdm, jtdm, kdm, nreg)
        self.ntr, self.nstep = ntr, nstep
        self.syn = synth.make_synth(config, ntr=ntr, land=land, metric=metric, seed=seed, **synth_kw)
        self.grid = self.syn.grid()
        self.state = self.syn.state(self.grid)
        self.scalars = self.syn.scalars(nstep)
        self.levels = time_levels(nstep, kdm)
        prep = self.new_oracle(setup=False)
        synth.fill_halos(prep, {**self.grid, **self.state})
        if isopycnic:  # vcoord='isopyc_bulkml': empty layers 3..kfpla-1, consistent kfpla (needs valid dp halos)
            synth.make_isopycnic(self.state)
        prep.bigrid("depths")
        self.nreg = prep.nreg
        self.masks = {k: prep.get_int(k).reshape(self.syn.ldj, self.syn.ldi).copy() for k in ("ip", "iu", "iv", "iq")}
        synth.derive(self.grid, self.state, self.masks, self.levels, self.scalars, prep)

    def arrays(self):
        return {k: v.copy() for k, v in {**self.grid, **self.state}.items()}

    def new_oracle(self, setup=True):
        itdm, jtdm, kdm, nreg = self.dims
        o = Oracle(itdm, jtdm, kdm, nreg, ntr=self.ntr)
        arrs = self.arrays() if setup else {**self.grid, **self.state}
        o.register_all(arrs)
        o.set_scalars(**self.scalars)
        if setup:
            o.bigrid("depths")
        return o

    def new_gpu(self, parity=True, **kw):
        itdm, jtdm, kdm, nreg = self.dims
        g = BlomGpu(itdm, jtdm, kdm, nreg, ntr=self.ntr, parity=parity, **kw)
        g.register_all(self.arrays())
        g.set_scalars(**self.scalars)
        g.bigrid("depths")
        return g


def prepare_step(c, backends, options=None):
    """Set the namelist options (default: the reference's defaults for the hybrid coordinate, as
    driver.reference_options) on every backend, register the ALE products neutral diffusion consumes
    when ndiff is on the path, and run the setup routines.  Returns (routines, options)."""
    from blom_b200.driver import reference_options, step_routines
    opts = {**reference_options(c.config), **(options or {})}
    routines = step_routines(opts)
    for b in backends:
        for k, v in opts.items():
            b.set_option(k, v)
        if "ndiff" in routines:
            nd = synth.ndiff_inputs(c.syn, c.state, c.levels, ntr=c.ntr)
            b.register_all({k: v.copy() for k, v in nd.items()})
        b.inieos(); b.numerical_bounds(); b.init_cppm()
    return routines, opts


def compare_all(g, o, what, tol=1e-10, skip=("depths",)):
    """Download every registered array of the CUDA backend and compare its interior with the oracle's
    (max-norm relative error <= tol)."""
    g.download_all()
    bad = []
    for nm, a in g.arrays.items():
        if nm in skip or a.dtype != np.float64:
            continue
        err = max_rel_err(interior(a), interior(o.arrays[nm]))
        if not err <= tol:
            bad.append((nm, err))
    assert not bad, (what, sorted(bad, key=lambda t: -t[1])[:6])


def interior(a, nb=4, halo=0):
    """View of the interior (+`halo` rings) of a (nlev, ldj, ldi) array."""
    s = nb - halo
    return a[..., s:a.shape[-2] - s, s:a.shape[-1] - s]


def max_rel_err(a, b, floor=1e-300):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    scale = max(np.max(np.abs(b)), floor)
    return float(np.max(np.abs(a - b)) / scale)


def assert_fma_close(a, b, what, tol=1e-11, max_flip_frac=5e-3, max_flip=2e-2):
    """Comparison of the FMA-contracted performance build with the oracle (-ffp-contract=off).
    CPPM's limiters branch on sign tests of near-cancelling expressions; under FMA contraction a few
    of them take the other branch in nearly massless cells (the oracle itself, compiled with
    -ffp-contract=fast, differs from the oracle in the same cells by the same values:
    tests/dev/fma_sensitivity.py, profiles/r01_fma_sensitivity.txt).  So: every point within `tol` of
    the oracle (relative to the field's max-norm) except at most max(8, max_flip_frac*size) points,
    and none further away than `max_flip`.  Returns the number of points beyond `tol`."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b) / max(np.max(np.abs(b)), 1e-300)
    assert np.isfinite(d).all(), what
    far = d > tol
    assert far.sum() <= max(8, max_flip_frac * far.size) and d.max() <= max_flip, \
        (what, int(far.sum()), far.size, float(d.max()))
    return int(far.sum())


def ulp_diff(a, b):
    """max distance in units of last place between two float64 arrays."""
    ai = np.asarray(a, dtype=np.float64).view(np.int64).astype(np.int64)
    bi = np.asarray(b, dtype=np.float64).view(np.int64).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2 ** 63) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2 ** 63) - bi, bi)
    return int(np.max(np.abs(ai - bi))) if ai.size else 0
