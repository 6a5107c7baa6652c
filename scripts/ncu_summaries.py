#!/usr/bin/env python
"""Summaries of ncu output for profiles/ (run here, no GPU needed: ncu only reads the report).

    python scripts/ncu_summaries.py full   gpurun_out/r01f_full.ncu-rep   profiles/r01f
    python scripts/ncu_summaries.py launch gpurun_out/r01f_launches.csv   profiles/r01f

full   -> <prefix>_ncu_full_fused_step.csv (one row per profiled launch, the metrics DESIGN.md quotes) and
          <prefix>_ncu_dram_traffic.json   (dram__bytes_read + dram__bytes_write per launch, averaged per kernel
          instance and grid size; bench.py reads `roofline.traffic` from it)
launch -> <prefix>_ncu_launch_list_summary.csv (gpu__time_duration.sum of the launch list grouped by kernel)
"""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
LOPS = {0: "none", 1: "drift", 2: "kick", 3: "invx+kick", 4: "c2r"}
SOPS = {0: "none", 1: "scale", 2: "drift", 3: "drift+alias", 4: "psi+rho", 5: "rho", 6: "poisson", 7: "max",
        8: "poisson+inv", 9: "psi+rho+fwdx", 10: "rho+fwdx", 11: "drift+alias+inv", 12: "r2c"}


def pretty(name):
    m = re.search(r"fft_pass_kernel<\(int\)(\d+), \(bool\)(\d), \(int\)(\d+), \(int\)(\d+), \(bool\)(\d)>", name)
    if not m:
        m = re.search(r"fft_pass_kernel<(\d+), (\d), (\d+), (\d+), (\d)>", name)
    if not m:
        return name.split("(")[0]
    n, inv, lop, sop, xl = (int(x) for x in m.groups())
    return f"fft_pass<{n},{'inv' if inv else 'fwd'},{LOPS[lop]},{SOPS[sop]},{'x' if xl else 'yz'}>"


def full(rep, prefix):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    with open(prefix + "_ncu_full_fused_step.csv", "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow(["Kernel Name"] + METRICS)
        w.writerow([""] + [units[col[m]] for m in METRICS])
        for r in data:
            w.writerow([pretty(r[col["Kernel Name"]])] + [r[col[m]] for m in METRICS])
    traffic = {}
    for r in data:
        to_b = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd = float(r[col["dram__bytes_read.sum"]]) * to_b[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * to_b[units[col["dram__bytes_write.sum"]]]
        ms = float(r[col["gpu__time_duration.sum"]]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[units[col["gpu__time_duration.sum"]]]
        key = f"{pretty(r[col['Kernel Name']])} grid={r[col['launch__grid_size']]}"
        t = traffic.setdefault(key, {"launches": 0, "avg_ms": 0.0, "dram_bytes_per_launch": 0.0})
        t["launches"] += 1
        t["avg_ms"] += ms
        t["dram_bytes_per_launch"] += rd + wr
    for t in traffic.values():
        t["avg_ms"] /= t["launches"]
        t["dram_bytes_per_launch"] /= t["launches"]
    json.dump(traffic, open(prefix + "_ncu_dram_traffic.json", "w"), indent=1)
    print("wrote", prefix + "_ncu_full_fused_step.csv", prefix + "_ncu_dram_traffic.json")


def launch(path, prefix):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    head = rows[0]
    col = {h: i for i, h in enumerate(head)}
    agg = {}
    for r in rows[1:]:
        if len(r) <= col["Metric Value"] or r[col["Metric Name"]] != "gpu__time_duration.sum":
            continue
        unit = r[col["Metric Unit"]]
        ms = float(r[col["Metric Value"]].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        a = agg.setdefault(pretty(r[col["Kernel Name"]]), [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(prefix + "_ncu_launch_list_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1]:.3f}", f"{100 * v[1] / tot:.2f}"])
    print("wrote", prefix + "_ncu_launch_list_summary.csv")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
