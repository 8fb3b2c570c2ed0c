#!/usr/bin/env python
"""Summarise an ncu launch list (--csv --log-file, one row per launch and metric) of `bench.py --ncu-step`:
per kernel: launches, time and share of the step, DRAM bytes and GB/s, tensor-pipe %.
usage: ncu_summarise.py launches.csv out.txt [traffic.json]"""
import csv
import json
import re
import sys
from collections import OrderedDict, defaultdict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"mdk::", "", name)
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    name = re.sub(r"\(mdk::\w+\)$|\(const mdk::\w+\)$|\(\w+Params\)$", "", name)
    return name[:70]


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rd:
        if len(r) != len(hdr):
            continue
        key = r[ix["ID"]]
        d = launches.setdefault(key, dict(name=r[ix["Kernel Name"]]))
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[m] = v
    agg = defaultdict(lambda: dict(n=0, ms=0.0, bytes=0.0, tens=0.0))
    for d in launches.values():
        a = agg[short(d["name"])]
        t = d.get("gpu__time_duration.sum", 0.0)
        a["n"] += 1
        a["ms"] += t
        a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["tens"] += t * d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    total = sum(a["ms"] for a in agg.values())
    with open(out, "w") as f:
        f.write("ncu launch list of ONE eager denoising step (bench.py --ncu-step, config B): per-launch times are\n"
                "cold-cache and serialised -> compare SHARES with bench.py's CUDA-event profile, not absolutes.\n")
        f.write(f"{len(launches)} launches, {total:.1f} ms under ncu\n")
        f.write(f"{'kernel':72s} {'n':>5s} {'ms':>9s} {'share':>7s} {'DRAM GB':>9s} {'GB/s':>8s} {'tensor%':>8s}\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
            gbs = a["bytes"] / (a["ms"] * 1e-3) / 1e9 if a["ms"] > 0 else 0.0
            tp = a["tens"] / a["ms"] if a["ms"] > 0 else 0.0
            f.write(f"{k:72s} {a['n']:5d} {a['ms']:9.3f} {100 * a['ms'] / total:6.2f}% {a['bytes'] / 1e9:9.3f} "
                    f"{gbs:8.0f} {tp:8.1f}\n")
    if len(sys.argv) > 3:
        tr = {}
        for k, a in agg.items():
            base = k.split("<")[0]
            t = tr.setdefault(base, dict(dram_bytes=0.0, launches=0, ms_under_ncu=0.0,
                                         source="ncu launch list of bench.py --ncu-step (one step)"))
            t["dram_bytes"] += a["bytes"]
            t["launches"] += a["n"]
            t["ms_under_ncu"] += a["ms"]
        with open(sys.argv[3], "w") as f:
            json.dump(tr, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
