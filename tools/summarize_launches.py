"""Summarise an ncu launch list (``--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv``)
into per-kernel totals: launches, device time, share of the step, DRAM bytes.  Writes a CSV summary and the JSON that
bench.py reads for ``roofline.traffic``.

    python tools/summarize_launches.py gpurun_out/launches.csv profiles/r01_launches_summary.csv profiles/r01_traffic.json [N K]

With ``N K`` the launch list is cut into N equal consecutive parts (the N identical steps of one bench run) and only
part K (0-based) is summarised - K = 3 is the timed step of ``bench.py --steps 1 --warmup 3``.
"""
import csv
import json
import re
import sys
from collections import defaultdict

src, out_csv, out_json = sys.argv[1:4]
rows = []
with open(src, newline="") as f:
    lines = [ln for ln in f if not ln.startswith("==")]
rd = csv.DictReader(lines)
per = defaultdict(lambda: {"launches": set(), "ns": 0.0, "rd": 0.0, "wr": 0.0})


def to_bytes(value, unit):
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return float(value.replace(",", "")) * mult


def to_ns(value, unit):
    mult = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
    return float(value.replace(",", "")) * mult


records = list(rd)
if len(sys.argv) >= 6:
    n_parts, pick = int(sys.argv[4]), int(sys.argv[5])
    ids = sorted({int(r["ID"]) for r in records})
    per_part = len(ids) // n_parts
    assert per_part * n_parts == len(ids), (len(ids), n_parts)
    keep = set(ids[pick * per_part:(pick + 1) * per_part])
    records = [r for r in records if int(r["ID"]) in keep]
    print(f"{len(ids)} launches = {n_parts} steps x {per_part}; summarising step {pick}")
for r in records:
    name = re.sub(r"<.*", "", r["Kernel Name"]).split("(")[0].strip()
    name = name.split("::")[-1]
    e = per[name]
    e["launches"].add(r["ID"])
    m, v, u = r["Metric Name"], r["Metric Value"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        e["ns"] += to_ns(v, u)
    elif m == "dram__bytes_read.sum":
        e["rd"] += to_bytes(v, u)
    elif m == "dram__bytes_write.sum":
        e["wr"] += to_bytes(v, u)
tot = sum(e["ns"] for e in per.values()) or 1.0
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_us", "share_of_step", "avg_us", "dram_read_MB", "dram_write_MB", "dram_GBps"])
    for name, e in sorted(per.items(), key=lambda kv: -kv[1]["ns"]):
        n = len(e["launches"])
        w.writerow([name, n, f"{e['ns'] / 1e3:.1f}", f"{e['ns'] / tot:.4f}", f"{e['ns'] / 1e3 / n:.2f}", f"{e['rd'] / 1e6:.1f}",
                    f"{e['wr'] / 1e6:.1f}", f"{(e['rd'] + e['wr']) / max(e['ns'], 1.0):.1f}"])
    w.writerow(["TOTAL", sum(len(e["launches"]) for e in per.values()), f"{tot / 1e3:.1f}", "1.0", "", "", "", ""])
js = {name: {"launches": len(e["launches"]), "total_us": e["ns"] / 1e3, "share": e["ns"] / tot,
             "dram_bytes_per_launch": (e["rd"] + e["wr"]) / len(e["launches"])} for name, e in per.items()}
json.dump(js, open(out_json, "w"), indent=1, sort_keys=True)
print(open(out_csv).read())
