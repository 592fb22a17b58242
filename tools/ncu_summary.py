#!/usr/bin/env python
"""Turn ncu output brought back from the GPU box into the small summaries kept under profiles/.

    python tools/ncu_summary.py full  gpurun_out/prof.ncu-rep   profiles/rNN_ncu_full_summary.json
    python tools/ncu_summary.py list  gpurun_out/launches.csv   (prints per-kernel averages and shares)

`full` reads an `ncu --set full` report (needs the local ncu to decode it); `list` reads the CSV of
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv`.
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
        d = {k: (r[hdr.index(k)], units[hdr.index(k)]) for k in KEYS if k in hdr}
        rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * UNIT[d["dram__bytes_read.sum"][1]]
        wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * UNIT[d["dram__bytes_write.sum"][1]]
        stalls = {}
        for i, h in enumerate(hdr):
            if "average_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.25:
                    stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = v
        res[name] = {"traffic_bytes": rd + wr, "dram_read": rd, "dram_write": wr, "duration_us": float(d["gpu__time_duration.sum"][0]),
                     "metrics": {k: v[0] + " " + v[1] for k, v in d.items()}, "stalls_per_issue": stalls}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: {"us": v["duration_us"], "traffic_MB": v["traffic_bytes"] / 1e6} for k, v in res.items()}))


def launches(path):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(lines):
        agg[row["Kernel Name"].split("(")[0]][row["Metric Name"]].append(float(row["Metric Value"].replace(",", "")))
    tot = sum(sum(v["gpu__time_duration.sum"]) for v in agg.values())
    for k, v in agg.items():
        t = v["gpu__time_duration.sum"]
        print(f"{k:58s} n={len(t):3d} avg={sum(t) / len(t) / 1000:8.1f} us share={sum(t) / tot * 100:5.1f}% "
              f"dram_rd={sum(v['dram__bytes_read.sum']) / len(t) / 1e6:7.1f} MB wr={sum(v['dram__bytes_write.sum']) / len(t) / 1e6:6.1f} MB")


if __name__ == "__main__":
    if sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    else:
        launches(sys.argv[2])
