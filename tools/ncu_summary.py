#!/usr/bin/env python
"""Summarise ncu artefacts into small text files for profiles/ (run here, no GPU needed).

    python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep   profiles/r01_alr_full.txt
    python tools/ncu_summary.py launch gpurun_out/launches.csv   profiles/r01_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
        f.write("# (per-launch values; times under ncu are cold-cache and serialised)\n")
        for r in rows[2:]:
            f.write(f"\n== {r[idx['Kernel Name']]}  (id {r[idx['ID']]})\n")
            for k in KEEP:
                if k in idx:
                    f.write(f"{k:72s} {r[idx[k]]:>16s} {units[idx[k]]}\n")
            try:
                rd = float(r[idx["dram__bytes_read.sum"]].replace(",", ""))
                wr = float(r[idx["dram__bytes_write.sum"]].replace(",", ""))
                ur, uw = units[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_write.sum"]]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = rd * scale[ur] + wr * scale[uw]
                t = float(r[idx["gpu__time_duration.sum"]].replace(",", ""))
                tu = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[units[idx["gpu__time_duration.sum"]]]
                f.write(f"{'traffic = dram read+write per launch':72s} {tot / 1e6:16.3f} MB\n")
                f.write(f"{'dram GB/s under ncu':72s} {tot / (t * tu) / 1e9:16.1f} GB/s\n")
            except Exception as e:  # noqa: BLE001
                f.write(f"# traffic: {e}\n")
            stalls = [(hdr[i], r[i]) for i in range(len(hdr))
                      if "smsp__average_warps_issue_stalled" in hdr[i] and hdr[i].endswith("_per_issue_active.ratio")]
            stalls.sort(key=lambda x: -float(x[1].replace(",", "") or 0))
            f.write("top stall reasons (warps stalled per issue-active cycle):\n")
            for h, v in stalls[:6]:
                f.write(f"   {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):28s} {v}\n")


def launch(csvfile, out):
    lines = [l for l in open(csvfile) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list of {csvfile}\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':70s} {'launches':>9s} {'total us':>12s} {'avg us':>10s} {'share':>7s}\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:70]:70s} {n:9d} {t:12.1f} {t / n:10.1f} {100 * t / tot:6.1f}%\n")


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
