#!/usr/bin/env python
"""bench.py — simulated years per day of the BLOM hot path on B200 (+ roofline, CPU baseline).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

A "step" is one pass of the hot path (driver.step_routines, reference call order of
phy/mod_blom_step.F90:96-227, namelist defaults of the named grid) over one synthetic state.
  value  : SYPD = 86400 / (steps_per_year * t_step), steps_per_year = 365*86400/baclin,
           state resident in HBM, CUDA-event time on the library stream, max over ranks
  e2e    : same metric through the public host API with pinned HOST buffers (wall clock): per step the host
           hands over the time level its own routines wrote (new level of dp,T,S,u,v) and gets both levels
           back; the copies run on their own streams and overlap the kernels (driver.HotPath.step_pipelined)
  roofline: dominant kernel, algorithmic bytes (SURVEY.md §8d word counts) / live CUDA-event
           duration, against the measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline: the oracle (C++ restatement of the reference, kind "port") timed on a
           bounded row-band sample of the same grid; the Fortran reference cannot be built in
           this image (no gfortran/meson/netCDF), see DESIGN.md.
The library flavour that is timed is the parity-green one (-fmad=false, bit-identical to the oracle,
like the reference's -ffp-contract=off release build); `flavours_ms_per_step` also reports the
FMA-contracted flavour measured in a child process.
Multi-GPU (torchrun, one rank per GPU): the global grid is split into j-bands, so the
total work is fixed -> "scaling": "strong".

The GPU measurement can not be lost to the CPU legs: they run after the GPU numbers are complete,
inside try/except, and a failure is reported as {"error": ...} in the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from blom_b200 import synth  # noqa: E402
from blom_b200.driver import HotPath, balanced_band, reference_options, run_step, step_routines  # noqa: E402

METRIC = "simulated_years_per_day_hot_path"

# compulsory 8-byte words moved per interior cell per call of a routine (SURVEY.md §8a/§8d, T=2 scalars):
# distinct 3-D arrays read + written once.  ndiff: R p_src,p_dst (2), tpc_src 5T, t_srcdi 2T, trc_rm T, difiso, pu, pv,
# W trc_rm T + 4 face fluxes x2 (flld, flx) + nslp 2  (phy/mod_ndiff.F90:959-1175)
WORDS = {"advect": 10 + 29, "diffus": 19, "pgforc": 15, "momtum": 29, "tmsmt1": 6, "tmsmt2": 13,
         "eddtra": 21, "init_fluxes": 6, "pbcor1": 18, "pbcor2": 19, "ndiff": 2 + 10 + 4 + 2 + 3 + 2 + 8 + 2}
# with ltedtp='neutral' diffus only refreshes halos (phy/mod_diffus.F90:58-81)
WORDS_NEUTRAL = {"diffus": 0}
# per-kernel ALGORITHMIC 8-byte words per unit and launch (distinct arrays read + written once;
# DESIGN.md §3).  unit: "3d" = interior (i,j,k) cells of the tile, "2d" = (i,j) points,
# "bt" = (i,j) points x barotropic substeps covered by one launch.  The staged momtum kernels exchange
# layer-sized scratch arrays through HBM; those words are NOT algorithmic (29 words per cell
# for the whole routine) and are listed separately in KERNEL_SCRATCH_WORDS: `frac` of such a kernel counts
# only the model arrays it touches, and the routine-level figure is in `routines_roofline`.
KERNELS = {
    "zero_fluxes": (6, "3d"),
    "tmsmt1_kernel": (6, "3d"), "tmsmt2_kernel": (13, "3d"),
    "eddtra_column<u>": (7 + 6, "3d"), "eddtra_column<v>": (7 + 6, "3d"),   # R dp,dpu,p,difint,nslp,T,S  W 6 fluxes
    "advect_flux_area": (10, "3d"),
    "cppm_hedges<i>": (4, "3d"), "cppm_hedges<j>": (4, "3d"),               # R dp,(cross ca)  W hel,her
    "cppm_flux<i>": (12 + 6, "3d"), "cppm_flux<j>": (12 + 6, "3d"),         # R dp,T,S,hel,her,ca,cross ca(2),p,flx(3)  W dp,T,S,flx(3)
    "pbcor_update<1>": (9 + 9, "3d"), "pbcor_update<2>": (9 + 10, "3d"),     # R dp,T,S,flx(6)  W dp,T,S,flx(6)(,sigma)
    "pbcor_finish<1>": (3 + 4, "3d"), "pbcor_finish<2>": (3 + 4, "3d"),
    "pbcor_prep<1>": (3 + 1, "3d"), "pbcor_prep<2>": (3 + 2, "3d"),
    "diffus_flux": (4 + 4 + 4 + 4, "3d"), "diffus_update": (7 + 3, "3d"),
    "pg_p_from_dp": (2, "3d"), "pg_dpuv": (1 + 4, "3d"), "pg_dynh_march": (5 + 8, "3d"), "pg_finalize": (4, "3d"),
    "mt_pressures": (3 + 3, "3d"), "mt_drag": (3, "3d"),
    "mt_aux": (4, "3d"), "mt_vort": (3 + 2, "3d"), "mt_visc": (0, "3d"), "mt_flux1": (2, "3d"),
    "mt_update": (9 + 2, "3d"), "mt_update_v": (9 + 2, "3d"), "mt_column": (8 + 4, "3d"),
    "mt_fused": (21 + 6, "3d"),
    # neutral diffusion: per face column and layer R p_src,p_dst,snapped p_dst (3), interface records 2x4, t_srcdi 2T,
    # tpc_src 5T, difiso, layer means T, face pressure (1)  W 4 face fluxes (read-modify-write), nslp, 2T face
    # convergences; T = 2.  The kernel is latency-bound (data-dependent search per column), its HBM
    # fraction is reported for completeness only.  ndiff_prep: R the ALE products of a cell (p_src, p_dst, t_srcdi 2T,
    # tpc_src 5T), difiso, layer means T  W the column records (24 words per layer), {p_dst, snapped p_dst}, 4 face sums
    "ndiff_face<u>": (3 + 8 + 4 + 10 + 1 + 2 + 1 + 8 + 1 + 4, "3d"), "ndiff_face<v>": (3 + 8 + 4 + 10 + 1 + 2 + 1 + 8 + 1 + 4, "3d"),
    "ndiff_prep": (2 + 4 + 10 + 1 + 2 + 24 + 2 + 4, "3d"), "ndiff_update": (2 + 8 + 2, "3d"),
    "bt_subcycle": (43.2, "bt"), "bt_ueq": (23, "2d"), "bt_veq": (23, "2d"), "bt_continuity": (7, "2d"),
}
KERNEL_SCRATCH_WORDS = {"mt_aux": 7, "mt_vort": 4 + 2, "mt_visc": 2 + 4, "mt_flux1": 8 + 2, "mt_update": 12,
                        "mt_update_v": 12}
# barotropic substep: the reference touches 46R + 7W = 53 distinct 2-D arrays per point.  In every
# block of substeps at least one of the three time weights is exactly zero, and the arrays that only enter through a
# zero weight need not be read: 7 words in blocks 1-3, 14 in blocks 4-5 -> (3*46 + 2*39)/5 = 43.2 words that MUST move
BT_WORDS_PER_SUBSTEP = 43.2


def measured_traffic():
    """dram__bytes_read+write per launch from the committed ncu --set full captures (same config)."""
    p = ROOT / "profiles" / "traffic.json"
    return json.loads(p.read_text()) if p.exists() else {}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons (profiling recipe).  Started before the warm-up so the tool is
    already polling when the timed region begins; only samples stamped inside [mark_start, mark_end]
    are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(device)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark_start(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        import datetime
        self.t1 = datetime.datetime.now()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [[c.strip() for c in r.split(",")] for r in Path(self.f.name).read_text().strip().splitlines()
                if r.count(",") >= 8]
        os.unlink(self.f.name)
        sel = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0], "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                continue
            if self.t0 is None or (self.t0 <= ts <= (self.t1 or ts)):
                sel.append(r)
        if not sel:
            sel = rows[-3:]   # region shorter than the polling interval: nearest samples
            out["note"] = "no sample fell inside the timed region; nearest samples used"
        if not sel:
            return out
        try:
            sm = [float(r[1]) for r in sel]
            out["sm_mhz"] = float(np.median(sm))
            out["sm_max_mhz"] = float(sel[0][2])
            out["power_w_max"] = max(float(r[3]) for r in sel)
        except ValueError:
            return out
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for k, nm in enumerate(names):
            if any("Active" in r[5 + k] and "Not" not in r[5 + k] for r in sel):
                out["reasons"].append(nm)
        out["samples"] = len(sel)
        return out


def steps_per_year(baclin):
    return 365.0 * 86400.0 / baclin


def sypd(t_step_s, baclin):
    return 86400.0 / (steps_per_year(baclin) * t_step_s)


def workload_config(config, world):
    """The `config` object of the JSON line; identical for our arm and the reference arm."""
    itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
    opts = reference_options(config)
    lstep = 2 * int(np.ceil(0.5 * baclin / batrop))
    return {"workload": config, "grid": [itdm, jtdm, kdm], "nreg": nreg, "routines": step_routines(opts),
            "options": opts, "barotropic_substeps": 5 * lstep // 2, "baclin_s": baclin, "n_gpus": world,
            "cache": "3-D state (>= 5 GB per GPU) exceeds the 126 MB L2, every step streams it anew; no flush needed"}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle on a bounded row-band sample of the same grid
# ------------------------------------------------------------------------------------------
BAND_ROWS = 64   # rows per host process; fixed so that the sample does not depend on the core count


def _oracle_worker(config, steps, warmup, rows, widx, barrier, q):
    """One host process stepping a closed band of `rows` rows of the grid with the oracle."""
    try:
        from oracle.oracle import Oracle
        from blom_b200.lib import time_levels
        itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
        opts = reference_options(config)
        routines = step_routines(opts)
        sreg = nreg if rows == jtdm else (1 if nreg in (1, 2, 3) else 0)
        # same generator as the GPU arm; band `widx` of the grid, closed at its own edges
        syn = synth.Synth(itdm, rows, kdm, sreg, baclin=baclin, batrop=batrop, seed=20240611 + widx)
        grid = syn.grid(); state = syn.state(grid)
        o = Oracle(itdm, rows, kdm, sreg)
        arrs = {**grid, **state}
        o.register_all(arrs)
        scal = syn.scalars(1)
        o.set_scalars(**scal)
        for k, v in opts.items():
            o.set_option(k, v)
        synth.fill_halos(o, arrs)
        o.bigrid("depths")
        masks = {k: o.get_int(k).reshape(syn.ldj, syn.ldi) for k in ("ip", "iu", "iv", "iq")}
        levels = time_levels(1, kdm)
        synth.derive(grid, state, masks, levels, scal, o)
        if "ndiff" in routines:
            o.register_all(synth.ndiff_inputs(syn, state, levels))
        o.inieos(); o.numerical_bounds()
        if "advect" in routines:
            o.init_cppm()

        def one(nstep):
            o.set_scalar("nstep", nstep)
            run_step(o, routines, time_levels(nstep, kdm))
        ns = 1
        for _ in range(warmup):
            one(ns); ns += 1
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(steps):
            one(ns); ns += 1
        barrier.wait()
        q.put((widx, (time.perf_counter() - t0) / max(steps, 1)))
    except Exception as e:  # noqa: BLE001
        q.put((widx, repr(e)))
        try:
            barrier.abort()
        except Exception:  # noqa: BLE001
            pass


def oracle_sample(config, steps, warmup, cores=None):
    """The reference algorithm (oracle, C++ restatement) on the host cores: one process per core (at most
    jtdm // BAND_ROWS of them), each stepping its own closed band of BAND_ROWS rows of the grid (bands are
    independent -> no halo exchange cost, which flatters the CPU).  Returns (SYPD scaled to the full
    grid, seconds per sample step, wall seconds of the timed steps, description, processes used)."""
    import multiprocessing as mp
    from oracle.oracle import build
    build()
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    itdm, jtdm, kdm, nreg, baclin, batrop = synth.CONFIGS[config]
    avail = max(1, len(os.sched_getaffinity(0)))
    rows = min(BAND_ROWS, jtdm)
    cores = max(1, min(cores or avail, jtdm // rows))
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(cores)
    q = ctx.Queue()
    procs = [ctx.Process(target=_oracle_worker, args=(config, steps, warmup, rows, w, barrier, q))
             for w in range(cores)]
    for p in procs:
        p.start()
    res = [q.get() for _ in procs]
    for p in procs:
        p.join()
    bad = [r for r in res if not isinstance(r[1], float)]
    if bad:
        raise RuntimeError(f"oracle worker failed: {bad[0][1]}")
    t = max(r[1] for r in res)
    # scale the sample (cores bands of `rows` rows) to the full grid by cell count
    scale = (itdm * jtdm * kdm) / float(itdm * rows * cores * kdm)
    t_full = t * scale
    sample = (f"{cores} host processes (of {avail} usable cores) x {rows} rows x {itdm} x {kdm} layers: independent "
              f"closed bands of the same synthetic generator, {rows * cores} of {jtdm} rows; {steps} timed steps after "
              f"{warmup} warm-up steps; step time scaled by {scale:.3f} (cell count) to the full grid")
    return sypd(t_full, baclin), t, t * steps, sample, cores


def run_reference(args, rank):
    """--impl reference: the reference algorithm on the host cores for exactly --steps timed steps after
    --warmup warm-up steps; every step is the bounded band sample described in cpu_baseline.sample and
    ms_per_step is the measured sample step (steps x ms_per_step = timed wall time), `value` is scaled to
    the full grid."""
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = workload_config(args.config, max(world, args.gpus))
    try:
        v, t, wall, sample, cores = oracle_sample(args.config, args.steps, args.warmup)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "metric": METRIC, "unit": "SYPD", "config": cfg,
                          "unavailable": f"oracle sample failed: {e!r}"[:300]}), flush=True)
        return
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "SYPD",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
            "timed_wall_s": wall,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": "SYPD", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "SYPD", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "C++ restatement of the reference algorithm (oracle/); the Fortran reference cannot be "
                    "built in this image (no Fortran compiler, meson or netCDF).  ms_per_step is one step of the "
                    "band sample; value is that time scaled to the full grid by cell count"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def other_flavour_ms(args, parity):
    """ms per step of the OTHER library flavour, device-resident, measured in a child process after this
    process has released the GPU memory (N=1 only).  Never raises."""
    try:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--config", args.config, "--steps", str(max(3, min(args.steps, 10))),
               "--warmup", str(max(3, min(args.warmup, 3))), "--flavour", "fma" if parity else "parity", "--brief"]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return float(json.loads(ln)["ms_per_step"])
        return {"error": (r.stderr or "no output")[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # tnx0.25v4 is the grid BASELINE.json's north_star targets; it fits one B200 (~76 GiB resident), so the
    # same workload is used at every N and the 1->8 GPU numbers are a strong-scaling series.
    ap.add_argument("--config", default=os.environ.get("BLOM_BENCH_CONFIG", "tnx0.25v4"))
    ap.add_argument("--flavour", default="parity", choices=["parity", "fma"],
                    help="library flavour that is timed: parity (-fmad=false, bit-identical to the oracle) or fma")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-flavour", action="store_true")
    ap.add_argument("--brief", action="store_true", help="device-resident timing only (child runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    parity = args.flavour == "parity"
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()

    if world > 1:
        from blom_b200.lib import load_library
        import ctypes
        buf = ctypes.create_string_buffer(128)
        if rank == 0:
            lib = load_library(parity)
            if lib.blomgpu_comm_unique_id(buf) != 0:
                raise SystemExit("bench.py: " + lib.blomgpu_last_error().decode())
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).cuda()
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())

    t_setup = time.perf_counter()
    hp = HotPath(args.config, nstep=1, rank=rank, nranks=world, device=local_rank, comm_uid=uid,
                 pinned_alloc=None if args.brief else pinned, parity=parity)
    g = hp.gpu
    t_setup = time.perf_counter() - t_setup
    stream = torch.cuda.ExternalStream(g.stream())

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    # ---- device-resident timing ---------------------------------------------------------
    clocks = ClockSampler(local_rank)
    for _ in range(args.warmup):
        hp.advance()
    barrier()
    g.launch_count_reset()
    clocks.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        hp.advance()
    e1.record(stream)
    barrier()
    clocks.mark_end()
    launches = g.launch_count()
    t_ms = e0.elapsed_time(e1)
    tt = torch.tensor([t_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_step = float(tt.item()) / 1e3 / args.steps
    clk = clocks.stop()
    if args.brief:
        if rank == 0:
            print(json.dumps({"flavour": args.flavour, "ms_per_step": 1e3 * t_step, "steps": args.steps,
                              "warmup": args.warmup, "config": args.config}), flush=True)
        hp.finalize()
        return

    # ---- end to end through the host API (pinned host buffers, copies inside) ------------
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        hp.step_pipelined()               # H2D of the new level, the step, D2H of both levels; ends with a sync
        hp.set_step(hp.nstep + 1)
    g.sync()
    w1 = time.perf_counter()
    te = torch.tensor([(w1 - w0) / args.steps], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_e2e = float(te.item())
    h2d, d2h = hp.io_bytes_pipelined()
    if dist is not None:   # whole-job bytes
        tb = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        h2d, d2h = int(tb[0].item()), int(tb[1].item())

    # ---- per-routine and per-kernel device times (2 extra steps, event pair per launch) ---
    g.timers_enable(True); g.timers_reset()
    hp.advance(); hp.advance()
    g.sync()
    rt = g.timers()
    g.timers_enable(False)
    g.ktimers_enable(True)
    hp.advance(); hp.advance()
    kt = g.ktimers()
    g.ktimers_enable(False)

    # per-rank routine times (N > 1): the slowest rank sets the pace, rank 0's own times hide the imbalance
    ranks_rt = None
    if dist is not None:
        ranks_rt = [None] * world
        dist.all_gather_object(ranks_rt, {k: v["ms"] / v["calls"] for k, v in rt.items() if v["calls"]})
    if rank != 0:
        hp.finalize()
        if dist is not None:
            dist.destroy_process_group()
        return

    hbm_peak, peak_src = peaks()
    cells_local = hp.itdm * hp.jj * hp.kdm
    cells2d_local = hp.itdm * hp.jj
    routines_ms = {k: v["ms"] / v["calls"] for k, v in rt.items() if v["calls"]}
    ksum = sum(v["ms"] for v in kt.values()) or 1.0
    # roofline per kernel: algorithmic bytes / live per-launch CUDA-event time; the dominant kernel
    # (largest share of the step) is the headline `roofline`, the rest goes to `roofline_top`
    lstep = hp.scalars["lstep"]
    traffic = measured_traffic().get(args.config, {})
    roofs = []
    for name, v in kt.items():
        if name not in KERNELS or not v["launches"]:
            continue
        w, unit = KERNELS[name]
        dur = v["ms"] / v["launches"] / 1e3
        if unit == "3d":
            nbytes = 8.0 * w * cells_local
        elif unit == "2d":
            nbytes = 8.0 * w * cells2d_local
        else:  # substeps covered by one launch: 5*lstep/2 substeps over the launches of one step
            nbytes = 8.0 * w * cells2d_local * (5 * lstep // 2) / (v["launches"] / 2.0)
        ach = nbytes / dur / 1e9
        r = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
             "frac": ach / hbm_peak, "traffic": traffic.get(name), "peak_source": peak_src,
             "avg_launch_us": dur * 1e6, "share_of_step": v["ms"] / ksum, "algorithmic_bytes_per_launch": nbytes}
        if name in KERNEL_SCRATCH_WORDS:
            r["scratch_words_not_counted"] = KERNEL_SCRATCH_WORDS[name]
        if name.startswith("ndiff_face"):
            r["note"] = ("not a bandwidth kernel: one thread per face column runs the reference's two sequential, "
                         "data-dependent searches (phy/mod_ndiff.F90:160-953); bound by the latency of dependent "
                         "loads and instructions at 16 warps per SM (ncu at tnx1v4, profiles/r02_ncu_full_ndiff_face.txt: "
                         "IPC 1.2 of 4, 16 of 32 lanes active, 9 % of the DRAM peak, 47 % of the stall samples on the "
                         "long scoreboard).  Round 2: per-column source records staged in shared memory with an L2 "
                         "prefetch of the next layer (152.8 -> 115.4 ms for both directions at tnx0.25v4).  The HBM "
                         "fraction is reported because the contract asks for one; `traffic` is measured at tnx1v4 only "
                         "(3.9 GB per launch there for 2.5 GB of algorithmic bytes)")
        if unit == "bt":
            ws_mb = 8.0 * w * cells2d_local / 1e6
            r["note"] = (f"streamed model ({w} words per 2-D point and substep, see BT_WORDS_PER_SUBSTEP); 2-D working set {ws_mb:.0f} MB "
                         + ("fits the 126 MB L2, so DRAM traffic is far below the algorithmic bytes"
                            if ws_mb < 120 else "exceeds the 126 MB L2, so every substep streams from HBM"))
        roofs.append(r)
    roofs.sort(key=lambda r: -r["share_of_step"])
    roof = roofs[0] if roofs else None
    # routine-level roofline: SURVEY §8 words of the whole routine / routine device time
    words = {**WORDS, **(WORDS_NEUTRAL if hp.options.get("ltedtp") == "neutral" else {})}
    rroof = {}
    for r_, ms in routines_ms.items():
        if r_ == "barotp":
            nb_ = 8.0 * BT_WORDS_PER_SUBSTEP * cells2d_local * (5 * lstep // 2)
        elif words.get(r_):
            nb_ = 8.0 * words[r_] * cells_local
        else:
            continue
        rroof[r_] = {"ms": ms, "algorithmic_GB": nb_ / 1e9, "GBps": nb_ / ms / 1e6, "frac": nb_ / ms / 1e6 / hbm_peak}
    # whole-step algorithmic bandwidth
    alg_bytes = 8.0 * hp.cells * sum(words.get(r, 0) for r in hp.routines)
    if "barotp" in hp.routines:
        alg_bytes += 8.0 * BT_WORDS_PER_SUBSTEP * hp.itdm * hp.jtdm * (5 * lstep // 2)
    cfg = workload_config(args.config, world)
    cfg["routines"] = hp.routines
    cfg["options"] = {k: hp.options[k] for k in cfg["options"]}
    line = {
        "metric": METRIC, "value": sypd(t_step, hp.baclin), "unit": "SYPD",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg,
        "flavour": args.flavour + (" (-fmad=false, bit-identical to the oracle)" if parity else " (FMA contraction)"),
        "parallelism": f"j-bands x{world}", "setup_s": round(t_setup, 1),
        "e2e": {"value": sypd(t_e2e, hp.baclin), "unit": "SYPD", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * t_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roof,
        "roofline_top": [{k: r[k] for k in ("kernel", "achieved", "frac", "share_of_step", "avg_launch_us", "traffic")}
                         for r in roofs[1:8]],
        "routines_roofline": rroof,
        "step_algorithmic_GBps": alg_bytes / t_step / 1e9,
        "step_frac_of_hbm_peak": alg_bytes / t_step / 1e9 / hbm_peak,
        "routines_ms": routines_ms,
        **({"routines_ms_max_over_ranks": {k: max(r.get(k, 0.0) for r in ranks_rt) for k in routines_ms},
            "rows_per_rank": [balanced_band(args.config, r, world)[1] for r in range(world)]} if ranks_rt else {}),
        "kernels_ms_per_step": {k: v["ms"] / 2.0 for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["ms"])},
    }
    # the GPU measurement is complete: release the device, then run the legs that may fail
    try:
        hp.finalize()
    except Exception as e:  # noqa: BLE001
        line["finalize_error"] = repr(e)[:200]
    if dist is not None:
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    if world == 1 and not args.no_other_flavour:
        line["flavours_ms_per_step"] = {args.flavour: 1e3 * t_step,
                                        ("fma" if parity else "parity"): other_flavour_ms(args, parity)}
    if world == 1 and not args.no_cpu_baseline:
        try:
            v, t, wall, sample, cores = oracle_sample(args.config, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": "SYPD", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)[:300], "kind": "port"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
