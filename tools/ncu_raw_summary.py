"""Markdown summary of an `ncu -i X.ncu-rep --page raw --csv` export: one table per profiled launch with the metrics the
roofline discussion uses.   ncu -i rep --page raw --csv > raw.csv ; python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
names, units = rows[hdr], rows[hdr + 1]
ix = {n: i for i, n in enumerate(names)}
for r in rows[hdr + 2:]:
    if len(r) < len(names):
        continue
    print(f"### `{r[ix['Kernel Name']][:110]}`\n\n| metric | value |\n|---|---|")
    for m in WANT:
        if m in ix and r[ix[m]] != "":
            print(f"| {m} | {r[ix[m]]} {units[ix[m]]} |")
    print()
