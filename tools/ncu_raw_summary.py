"""Pick the headline metrics out of `ncu -i X.ncu-rep --page raw --csv` and merge them into a JSON summary.
usage: python tools/ncu_raw_summary.py raw.csv profiles/summary.json section_name"""
import csv, json, sys, os
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]
rows = list(csv.reader(open(sys.argv[1])))
h, units, v = rows[0], rows[1], rows[2]
out = {}
for w in WANT:
    if w in h:
        i = h.index(w)
        out[w] = ("%s %s" % (v[i], units[i])).strip()
path = sys.argv[2]
d = json.load(open(path)) if os.path.exists(path) else {}
d[sys.argv[3]] = out
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
