#!/usr/bin/env python
"""Multi-GPU parity of the j-band path (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dev/mgpu_parity.py tnx1v4 1

Every rank steps its band of the synthetic state through the full hot path; rank 0 then runs
the SAME case on one GPU (one tile) and on the CPU oracle and compares the assembled bands:
  bands vs one tile : bit-identical (same per-cell operations; halos, fold and xcsum order preserved)
  bands vs oracle   : <= 1e-10 of the field max-norm (chained-steps tolerance of DESIGN.md §4)
Prints one JSON line; exit code 1 on mismatch."""
import json
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

FIELDS = ["dp", "temp", "saln", "u", "v", "uflx", "vflx", "utflx", "usflx", "pgfx", "pgfy", "pb", "ubflxs_p",
          "p", "sigma", "umfltd", "vmflsm", "ub", "vb", "nslpx", "nslpy", "nd_trc_rm", "utflld"]


def main():
    import torch
    import torch.distributed as dist
    from blom_b200.driver import HotPath
    from blom_b200.lib import load_library
    import ctypes

    cfg = sys.argv[1] if len(sys.argv) > 1 else "mid2"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", lr))
    buf = ctypes.create_string_buffer(128)
    if rank == 0:
        load_library(True).blomgpu_comm_unique_id(buf)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().numpy().tobytes())
    tmp = os.environ.get("MGPU_TMP") or tempfile.gettempdir()

    comm = os.environ.get("MGPU_COMM", "p2p")   # p2p: CUDA-IPC mailboxes over NVLink; nccl: send/recv
    hp = HotPath(cfg, ntr=1, nstep=1, rank=rank, nranks=world, device=lr, parity=True, comm_uid=uid,
                 options={"comm": comm})
    sums = []
    for _ in range(nsteps):
        hp.advance()
    sums.append(hp.gpu.xcsum("dp", "ip", lev=1))
    crc = hp.gpu.chksum("temp", 2 * hp.kdm, 1)
    hp.gpu.download_all()
    nb = 4
    np.savez(os.path.join(tmp, f"mgpu_band_{rank}.npz"), j0=hp.j0, jj=hp.jj,
             **{f: hp.arrays[f][..., nb:nb + hp.jj, nb:nb + hp.itdm] for f in FIELDS})
    hp.finalize()
    dist.barrier()
    ok = True
    out = {"config": cfg, "n_gpus": world, "steps": nsteps, "comm": comm}
    if rank == 0:
        one = HotPath(cfg, ntr=1, nstep=1, device=lr, parity=True)
        for _ in range(nsteps):
            one.advance()
        s1 = one.gpu.xcsum("dp", "ip", lev=1)
        crc1 = one.gpu.chksum("temp", 2 * one.kdm, 1)
        one.gpu.download_all()
        # CPU oracle on the same case (test infrastructure), same option set and call order
        from util import Case, prepare_step
        from blom_b200.driver import run_step
        from blom_b200.lib import time_levels
        c = Case(cfg, ntr=1, nstep=1)
        o = c.new_oracle()
        routines, _ = prepare_step(c, (o,))
        assert routines == one.routines
        kk = c.dims[2]
        for ns in range(1, nsteps + 1):
            o.set_scalar("nstep", ns)
            run_step(o, routines, time_levels(ns, kk))
        worst_bit, worst_orc = 0.0, 0.0
        bands = [np.load(os.path.join(tmp, f"mgpu_band_{r}.npz")) for r in range(world)]
        for f in FIELDS:
            full = np.concatenate([b[f] for b in bands], axis=-2)
            ref1 = one.arrays[f][..., nb:-nb, nb:-nb]
            refo = o.arrays[f][..., nb:-nb, nb:-nb]
            same = np.array_equal(full, ref1)
            scale = max(np.abs(refo).max(), 1e-300)
            eo = float(np.abs(full - refo).max() / scale)
            e1 = float(np.abs(full - ref1).max() / max(np.abs(ref1).max(), 1e-300))
            worst_bit = max(worst_bit, e1); worst_orc = max(worst_orc, eo)
            if not same or not eo <= 1e-10:
                ok = False
                out.setdefault("bad", []).append([f, e1, eo])
        out.update({"bands_vs_one_tile_max_rel": worst_bit, "bands_vs_oracle_max_rel": worst_orc,
                    "xcsum_equal": sums[0] == s1, "crc_equal": crc == crc1, "crc": f"0x{crc:08X}"})
        ok = ok and sums[0] == s1 and crc == crc1
        out["ok"] = ok
        print(json.dumps(out), flush=True)
        one.finalize()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
