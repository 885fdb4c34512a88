#!/usr/bin/env python
"""Turn the CSV logs written by tools/gpu_round_check.sh into the text summaries kept under profiles/.

  python tools/ncu_summaries.py launches gpurun_out/launches_TAG.csv > profiles/TAG_launch_list_summary.txt
  python tools/ncu_summaries.py metrics  gpurun_out/mom_TAG.csv      > profiles/TAG_mom_ncu_summary.txt
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def rows(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def launches(path):
    rs = rows(path)
    tot = defaultdict(float)
    cnt = defaultdict(int)
    for r in rs:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ms = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[r["Metric Unit"]]
        k = r["Kernel Name"][:90]
        tot[k] += ms
        cnt[k] += 1
    all_ms = sum(tot.values())
    print(f"ncu launch list (gpu__time_duration.sum, --clock-control none), `python bench.py --steps 2 --warmup 1 --no-e2e "
          f"--no-cpu-baseline`, all {sum(cnt.values())} launches of the process (set-up, 3 steps, the instrumented step, "
          f"diagnostics; cold-cache serialised times: compare shares)")
    print("kernel | launches | total ms | share")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k} | {cnt[k]} | {tot[k]:.3f} | {100 * tot[k] / all_ms:.1f}%")


def metrics(path):
    per = OrderedDict()
    for r in rows(path):
        per.setdefault((r["ID"], r["Kernel Name"][:60]), []).append((r["Metric Name"], r["Metric Unit"], r["Metric Value"]))
    for (_, name), ms in per.items():
        print(name)
        for m, u, v in ms:
            v = float(v.replace(",", ""))
            if u == "ns":
                v, u = v * 1e-6, "ms"
            if u == "byte":
                v, u = v * 1e-9, "GB"
            if u == "Kbyte":
                v, u = v * 1e-6, "GB"
            if u == "Mbyte":
                v, u = v * 1e-3, "GB"
            if u == "Gbyte":
                u = "GB"
            print(f"   {m} {v:.6f} {u}")


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2])
