#!/usr/bin/env python3
"""Turn `ncu -i X.ncu-rep --page raw --csv` into the small JSON bench.py reads for roofline.traffic.
usage: python tools/ncu_metrics_json.py raw.csv out.json "<how the capture was made>" [algorithmic bytes of the captured launch]
"""
import csv
import json
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches, dram = [], []
    kernel = None
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        kernel = r[col["Kernel Name"]]
        d = {}
        for k in KEEP:
            if k in col:
                d[k] = f"{r[col[k]]} {units[col[k]]}".strip()
        launches.append(d)
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
        dram.append(tot)
    out = {"source": sys.argv[3] if len(sys.argv) > 3 else "", "kernel": kernel, "launches": launches,
           "dram_bytes_per_launch": sum(dram) / max(1, len(dram))}
    if len(sys.argv) > 4:  # 4 B per sample x samples of the captured launch: bench.py pairs `traffic` with THIS, not with its own average launch
        out["algorithmic_bytes_per_launch"] = float(sys.argv[4])
        out["traffic_over_algorithmic"] = out["dram_bytes_per_launch"] / out["algorithmic_bytes_per_launch"]
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print(json.dumps({k: out[k] for k in ("kernel", "dram_bytes_per_launch")}))


if __name__ == "__main__":
    main()
