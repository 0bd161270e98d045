"""DRAM traffic of the dominant launch from an `ncu --set full` capture, in the form bench.py reads (roofline.traffic).
usage: python tools/ncu_traffic.py raw.csv profiles/rNN_traffic.json grid nets width "launch description" visits pops pushes
raw.csv = `ncu -i X.ncu-rep --page raw --csv` of ONE launch; visits / pops / pushes = the counters of that launch (tools/prof_run.py)."""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, units, v = rows[0], rows[1], rows[2]


def val(name):
    i = h.index(name)
    x, u = float(v[i].replace(",", "")), units[i].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
    key = u if u in scale else u.replace("second", "s")
    return x * scale.get(key, 1.0)


rd, wr, dur = val("dram__bytes_read.sum"), val("dram__bytes_write.sum"), val("gpu__time_duration.sum")
visits, pops, pushes = (float(x) for x in sys.argv[7:10])
alg = 36.0 * visits + 28.0 * pops + 20.0 * pushes
out = {"grid": int(sys.argv[3]), "nets": int(sys.argv[4]), "width": int(sys.argv[5]), "launch": sys.argv[6],
       "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr, "duration_s_under_ncu": dur,
       "algorithmic_bytes_same_launch": alg, "amplification": (rd + wr) / alg if alg else None,
       "counters": {"edge_visits": visits, "pops": pops, "label_writes": pushes}}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out))
