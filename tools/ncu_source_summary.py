"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line:
executed warp instructions, stall samples and the dominant stall reason.
usage: python tools/ncu_source_summary.py src.csv [top]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
col = {n: i for i, n in enumerate(h) if n not in ("Source",)}
c_inst = h.index("Instructions Executed"); c_samp = h.index("# Samples")
stalls = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
lines = []
tot_inst = tot_samp = 0
reason_tot = collections.Counter()
for r in rows[hi + 1:]:
    if not r or not r[0] or len(r) < len(h):      # SASS rows, section headers
        continue
    try:
        inst = int(r[c_inst]); samp = int(r[c_samp])
    except ValueError:
        continue
    st = {h[i]: int(r[i]) for i in stalls if r[i] not in ("", "-")}
    for k, v in st.items(): reason_tot[k] += v
    lines.append((int(r[0]), r[1].strip(), inst, samp, max(st, key=st.get) if st else "-"))
    tot_inst += inst; tot_samp += samp
print("total warp instructions %d, stall samples %d" % (tot_inst, tot_samp))
print("by reason (%%): %s" % {k: round(100.0 * v / max(tot_samp, 1), 1) for k, v in reason_tot.most_common(9)})
print("\n   inst%%  samp%%  top stall              line  source")
for ln, src, inst, samp, why in sorted(lines, key=lambda x: -x[2])[:top]:
    print("  %5.1f  %5.1f  %-22s L%-4d %s" % (100.0 * inst / tot_inst, 100.0 * samp / max(tot_samp, 1), why, ln, src[:100]))
