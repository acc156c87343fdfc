#!/usr/bin/env python
"""Read a `.ncu-rep` (from `ncu --set full --clock-control none`, brought back in gpurun_out/) on the CPU box and
(1) print the per-launch summary that gets committed under profiles/, (2) with --traffic KERNEL:KEY record the DRAM
bytes of the first launch whose name contains KERNEL in profiles/ncu_traffic.json - the file bench.py's
`roofline.traffic` reads (so the bench line cites a committed capture instead of a literal).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--traffic attn6_kernel:B2_H48_S47056] [--source profiles/r02_x.txt]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]


def main():
    rep = sys.argv[1]
    traffic = source = None
    a = sys.argv[2:]
    while a:
        if a[0] == "--traffic":
            traffic, a = a[1], a[2:]
        elif a[0] == "--source":
            source, a = a[1], a[2:]
        else:
            a = a[1:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {os.path.basename(rep)}: ncu --page raw, one block per profiled launch")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        print(f"\n## {name[:140]}")
        for m in WANT:
            if m in idx:
                print(f"{m} [{units[idx[m]]}] = {r[idx[m]]}")
        if traffic and traffic.split(":")[0] in name:
            kern, key = traffic.split(":")
            rd = float(r[idx["dram__bytes_read.sum"]]) * SCALE[units[idx["dram__bytes_read.sum"]]]
            wr = float(r[idx["dram__bytes_write.sum"]]) * SCALE[units[idx["dram__bytes_write.sum"]]]
            path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            t = json.load(open(path)) if os.path.exists(path) else {}
            t.setdefault(kern, {})[key] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                                           "ms": float(r[idx["gpu__time_duration.sum"]]) * SCALE[units[idx["gpu__time_duration.sum"]]],
                                           "source": source or os.path.basename(rep)}
            json.dump(t, open(path, "w"), indent=1, sort_keys=True)
            traffic = None


if __name__ == "__main__":
    main()
