#!/usr/bin/env python
"""Per-launch summary of an `ncu --set full ... ; ncu -i rep --page raw --csv` dump (one row per kernel launch):
duration, DRAM bytes read / written, tensor-pipe activity (sm__pipe_tensor_cycles_active — the metric that reads
56-83 % for kernels running at 1.0-1.5 PFLOP/s; the TPC.TriageCompute *_realtime variant does not), DRAM throughput,
L2 hit rate, issue activity, registers, grid.

    python scripts/ncu_summary.py gpurun_out/prof_igemm_raw.csv [label ...] > profiles/r02_igemm_ncu_full.csv
"""
import csv
import sys

COLS = [("time_us", "gpu__time_duration.sum"),
        ("dram_read_MB", "dram__bytes_read.sum"),
        ("dram_write_MB", "dram__bytes_write.sum"),
        ("tensor_pipe_active_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("dram_throughput_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"),
        ("grid", "launch__grid_size"),
        ("cluster", "launch__cluster_dim_x")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    labels = sys.argv[2:]
    hdr, units = rows[0], rows[1]
    kname = hdr.index("Kernel Name")
    idx = [(n, hdr.index(m) if m in hdr else None) for n, m in COLS]
    w = csv.writer(sys.stdout)
    w.writerow(["launch", "layer", "kernel"] + [n for n, _ in idx])
    for i, r in enumerate(rows[2:]):
        if len(r) <= kname:
            continue
        name = r[kname].split("(")[0].replace("void ", "").replace("rovr::", "")
        vals = []
        for n, j in idx:
            v = r[j] if j is not None else ""
            if n == "time_us" and j is not None and units[j] == "ns":
                v = str(float(v) / 1e3)
            if n.endswith("_MB") and j is not None:
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[j], 1.0)
                v = str(round(float(v) * scale, 3))
            vals.append(v)
        w.writerow([i, labels[i] if i < len(labels) else "", name] + vals)


if __name__ == "__main__":
    main()
