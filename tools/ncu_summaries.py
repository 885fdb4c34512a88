#!/usr/bin/env python
"""Turn the CSV logs written by tools/gpu_round_check.sh into the text summaries kept under profiles/.

  python tools/ncu_summaries.py launches gpurun_out/launches_TAG.csv > profiles/TAG_launch_list_summary.txt
  python tools/ncu_summaries.py metrics  gpurun_out/mom_TAG.csv      > profiles/TAG_mom_ncu_summary.txt
  ncu -i X.ncu-rep --page raw --csv > X.csv; python tools/ncu_summaries.py full X.csv profiles/ncu_traffic.json 512 "source text"
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


if __name__ == "__main__" and sys.argv[1] in ("launches", "metrics"):
    {"launches": launches, "metrics": metrics}[sys.argv[1]](sys.argv[2])


def full(path, out_json=None, n=None, source=""):
    """`ncu -i X.ncu-rep --page raw --csv > X.csv` of a `--set full` capture -> text summary (stdout) and, with out_json,
    the per-kernel DRAM traffic record bench.py reads (profiles/ncu_traffic.json)"""
    import json
    with open(path, newline="") as f:
        rs = list(csv.reader(f))
    hdr, units, data = rs[0], rs[1], rs[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kernels = {}
    for r in data:
        name = r[ix["Kernel Name"]]
        print(name[:100])
        rec = {}
        for w in want:
            if w in ix:
                print(f"   {w} {r[ix[w]]} {units[ix[w]]}")
                rec[w] = (r[ix[w]], units[ix[w]])
        b = 0.0
        for w in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = rec[w]
            b += float(v.replace(",", "")) * scale.get(u, 1.0)
        base = name.split("<")[0].split("(")[0].replace("void ", "").split("::")[-1]
        kernels.setdefault(base, []).append(b)
    if out_json:
        rec = {"n": n, "source": source, "kernels": {k: {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)} for k, v in kernels.items()}}
        with open(out_json, "w") as f:
            json.dump(rec, f, indent=1)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "full":
    full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None, int(sys.argv[4]) if len(sys.argv) > 4 else None,
         sys.argv[5] if len(sys.argv) > 5 else "")
