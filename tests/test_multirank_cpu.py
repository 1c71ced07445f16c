"""Host-side logic of the N>1 path on CPU (gloo, world_size 2): band decomposition, band-wise
synthetic state == rows of the one-tile state, NCCL-id style broadcast plumbing, and a host
emulation of the N/S band-edge exchange checked against the oracle's one-tile xctilr."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from blom_b200 import synth  # noqa: E402
from blom_b200.driver import band  # noqa: E402


def test_band_partition_covers_grid():
    for jtdm in (385, 1153, 2165, 46):
        for n in (1, 2, 4, 8):
            rows = [band(jtdm, r, n) for r in range(n)]
            assert rows[0][0] == 0 and sum(jj for _, jj in rows) == jtdm
            for (a, ja), (b, _) in zip(rows, rows[1:]):
                assert a + ja == b
            assert max(jj for _, jj in rows) - min(jj for _, jj in rows) <= 1
            assert min(jj for _, jj in rows) >= 4  # nbdy-wide halos come from the direct neighbour only


def test_weighted_bands_cover_grid_and_balance_wet_columns():
    """driver.balanced_band: contiguous bands that cover the grid, at least 8 rows each, and a smaller spread of
    wet columns per band than equal-height bands on the 0.25 degree grid at 8 ranks."""
    from blom_b200.driver import balanced_band, weighted_bands
    for cfg in ("tnx0.25v4", "tnx0.125v4", "tnx1v4", "mid2"):
        jtdm = synth.CONFIGS[cfg][1]
        for n in (1, 2, 4, 8):
            if jtdm < 8 * n:
                continue
            rows = [balanced_band(cfg, r, n) for r in range(n)]
            assert rows[0][0] == 0 and sum(jj for _, jj in rows) == jtdm
            for (a, ja), (b, _) in zip(rows, rows[1:]):
                assert a + ja == b
            assert min(jj for _, jj in rows) >= 4
    s = synth.Synth.from_config("tnx0.25v4")
    wet = (s.depth_global > 0).sum(axis=1)

    def spread(bands):
        w = np.array([wet[j0:j0 + jj].sum() for j0, jj in bands], dtype=float)
        return w.max() / w.mean()
    equal = [band(1153, r, 8) for r in range(8)]
    weighted = [balanced_band("tnx0.25v4", r, 8) for r in range(8)]
    assert spread(weighted) < spread(equal) and spread(weighted) < 1.08
    # degenerate cost vectors still give legal partitions
    for cost in (np.ones(100), np.r_[np.zeros(90), np.ones(10)], np.r_[np.ones(10), np.zeros(90)]):
        bs = weighted_bands(cost, 4)
        assert bs[0][0] == 0 and sum(jj for _, jj in bs) == 100 and min(jj for _, jj in bs) >= 8


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS["mid2"]
        # the 128-byte communicator id travels from rank 0 like in bench.py
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid[:] = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(uid, 0)
        assert uid[77].item() == 77
        j0, jj = band(jtdm, rank, world)
        syn = synth.Synth(itdm, jtdm, kdm, nreg, ntr=1, j0=j0, jj=jj, baclin=baclin, batrop=batrop)
        gr = syn.grid(); st = syn.state(gr)
        nb = 4
        # N/S exchange of nh rows with the neighbours (host emulation of comm.cu exchange_ns)
        a = st["temp"]
        nh = 3
        reqs = []
        if rank + 1 < world:
            send = torch.from_numpy(np.ascontiguousarray(a[:, nb + jj - nh:nb + jj, nb:nb + itdm]))
            recv_n = torch.empty_like(send)
            reqs += [dist.isend(send, rank + 1), dist.irecv(recv_n, rank + 1)]
        if rank > 0:
            send_s = torch.from_numpy(np.ascontiguousarray(a[:, nb:nb + nh, nb:nb + itdm]))
            recv_s = torch.empty_like(send_s)
            reqs += [dist.isend(send_s, rank - 1), dist.irecv(recv_s, rank - 1)]
        for r in reqs:
            r.wait()
        if rank + 1 < world:
            a[:, nb + jj:nb + jj + nh, nb:nb + itdm] = recv_n.numpy()
        if rank > 0:
            a[:, nb - nh:nb, nb:nb + itdm] = recv_s.numpy()
        # gather the bands (interior + the exchanged inner halo rows) on rank 0
        payload = {"j0": j0, "jj": jj, **gr, **st}
        gathered = [None] * world
        dist.all_gather_object(gathered, payload)
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


def test_gloo_two_ranks_band_state_and_exchange():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS["mid2"]
    one = synth.Synth(itdm, jtdm, kdm, nreg, ntr=1, baclin=baclin, batrop=batrop)
    g1 = one.grid(); s1 = one.state(g1)
    nb = 4
    for b in gathered:
        j0, jj = b["j0"], b["jj"]
        for nm, ref in {**g1, **s1}.items():   # EVERY synthetic array: band rows == one-tile rows
            np.testing.assert_array_equal(b[nm][..., nb:nb + jj, nb:nb + itdm],
                                          ref[..., nb + j0:nb + j0 + jj, nb:nb + itdm], err_msg=nm)
    # exchanged rows equal the neighbour's interior rows == one-tile rows (what xctilr gives inside the domain)
    lo, hi = gathered
    nh = 3
    np.testing.assert_array_equal(lo["temp"][:, nb + lo["jj"]:nb + lo["jj"] + nh, nb:nb + itdm],
                                  s1["temp"][:, nb + lo["jj"]:nb + lo["jj"] + nh, nb:nb + itdm])
    np.testing.assert_array_equal(hi["temp"][:, nb - nh:nb, nb:nb + itdm],
                                  s1["temp"][:, nb + hi["j0"] - nh:nb + hi["j0"], nb:nb + itdm])


def test_fuk95_bands_equal_one_tile():
    """The analytic fuk95 generator is a pure function of global indices: the j-bands of a 2- and a
    4-rank decomposition hold exactly the rows of the one-tile state (what lets a multi-GPU run be
    compared band by band with the one-tile run and the oracle)."""
    from blom_b200.fuk95 import Fuk95
    one = Fuk95(ntr=1)
    g1 = one.grid(); s1 = one.state(g1)
    nb = 4
    for world in (2, 4):
        for rank in range(world):
            j0, jj = band(32, rank, world)
            b = Fuk95(ntr=1, j0=j0, jj=jj)
            gb = b.grid(); sb = b.state(gb)
            for nm in ("depths", "scpx", "corioq"):
                assert np.array_equal(gb[nm][:, nb:nb + jj, nb:-nb], g1[nm][:, nb + j0:nb + j0 + jj, nb:-nb]), nm
            for nm in ("dp", "temp", "saln", "sigma", "trc", "phi", "pb", "difiso", "dpuold", "dpvold"):
                assert np.array_equal(sb[nm][:, nb:nb + jj, nb:-nb], s1[nm][:, nb + j0:nb + j0 + jj, nb:-nb]), (world, rank, nm)
