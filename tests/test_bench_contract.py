"""CPU-side checks of bench.py's contract pieces that do not need a GPU: the reference arm prints one
JSON line with the agreed keys, the per-kernel word table covers every kernel the profiles list, and
the measured-traffic table holds the dominant kernel of the default workload."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--config", "tiny2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-1000:] + r.stderr[-1000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "SYPD" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["workload"] == "tiny2"


def test_reference_arm_other_ranks_stay_silent():
    import os
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_kernel_tables_cover_the_profiled_kernels():
    import bench
    last = json.loads((ROOT / "profiles" / "r01_bench_tnx025_n1_s10.json").read_text())
    timed = {k for k, ms in last["kernels_ms_per_step"].items() if ms >= 0.3 and k != "halo_kernel"}
    assert timed <= set(bench.KERNELS), sorted(timed - set(bench.KERNELS))
    traffic = bench.measured_traffic()
    assert traffic["tnx0.25v4"][last["roofline"]["kernel"]] > 0
    assert last["roofline"]["traffic"] is not None and 0 < last["roofline"]["frac"] <= 1
