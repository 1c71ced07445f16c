#!/usr/bin/env python
"""Collect per-kernel DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum, per launch) into
profiles/traffic.json, keyed like bench.py's kernel names.  Inputs: `ncu --set full` reports
(.ncu-rep) or the CSV log of an `ncu --metrics ...,dram__bytes_read.sum,dram__bytes_write.sum --csv`
launch list (the whole-step pass of the profiling recipe).

    python tools/make_traffic.py tnx1v4 gpurun_out/prof_a.ncu-rep [more.ncu-rep ...]
    python tools/make_traffic.py tnx0.25v4 gpurun_out/launches.csv
"""
import io
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def bench_name(ncu_name):
    n = ncu_name.replace("void ", "").replace("unnamed>::", "").split("(")[0].strip()
    m = re.match(r"cppm_flux<(\d)", n)
    if m:
        return "cppm_flux<i>" if m.group(1) == "0" else "cppm_flux<j>"
    m = re.match(r"cppm_hedges(?:_tile)?<(\d)", n)
    if m:
        return "cppm_hedges<i>" if m.group(1) == "0" else "cppm_hedges<j>"
    m = re.match(r"eddtra_column<(\d)", n)
    if m:
        return "eddtra_column<u>" if m.group(1) == "0" else "eddtra_column<v>"
    m = re.match(r"(pbcor_\w+)<(\d)", n)
    if m:
        return f"{m.group(1)}<{m.group(2)}>"
    m = re.match(r"ndiff_face<(\d)", n)
    if m:
        return "ndiff_face<u>" if m.group(1) == "0" else "ndiff_face<v>"
    # occupancy / shape template arguments are not part of bench.py's kernel names
    m = re.match(r"(bt_subcycle|mt_\w+|advect_flux_area|pg_\w+)<", n)
    if m:
        return m.group(1)
    return n


def main():
    cfg, reps = sys.argv[1], sys.argv[2:]
    acc = collections.defaultdict(list)
    for rep in reps:
        if rep.endswith(".csv"):
            lines = [l for l in open(rep) if not l.startswith("==")]
            per = collections.defaultdict(float)
            for r in csv.DictReader(io.StringIO("".join(lines))):
                if r["Metric Name"] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    per[(r["ID"], r["Kernel Name"])] += float(r["Metric Value"].replace(",", "")) * UNIT[r["Metric Unit"]]
            for (_, kn), tot in per.items():
                acc[bench_name(kn)].append(tot)
            continue
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(r[ix[key]].replace(",", "")) * UNIT[units[ix[key]]]
            acc[bench_name(r[ix["Kernel Name"]])].append(tot)
    path = ROOT / "profiles" / "traffic.json"
    data = json.loads(path.read_text()) if path.exists() else {}
    data.setdefault(cfg, {}).update({k: sum(v) / len(v) for k, v in acc.items()})
    path.write_text(json.dumps(data, indent=1, sort_keys=True) + "\n")
    for k, v in sorted(acc.items()):
        print(f"{k:24s} {sum(v) / len(v) / 1e6:10.1f} MB/launch over {len(v)} launches")


if __name__ == "__main__":
    main()
