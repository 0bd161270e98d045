#!/bin/sh
# Regenerates tests/golden/*.xz by running the UNMODIFIED reference (oracle/_ref/vpr_ref, built from
# /root/reference by `make -C oracle ref`).  Only works where /root/reference exists.
#   toy : tests/fixtures/gen_blif.py --luts 300  --pis 16 --window 60  --seed 1, W=64
#   mid : tests/fixtures/gen_blif.py --luts 4000 --pis 64 --window 400 --seed 2, W=200
#   hub : tests/fixtures/gen_blif.py --luts 900 --pis 24 --window 120 --seed 5 --hub 700, W=90 (one 84-sink net)
#   duo : tests/fixtures/gen_blif.py --luts 500 --pis 20 --window 80 --seed 7 --clocks 2, W=80 (two clock domains)
#   het : tests/fixtures/gen_blif.py --luts 400 --pis 20 --window 80 --seed 9 --mults 6 on k6_N10_het.xml (height-2 hard blocks), W=70;
#         W=60 is one track short: the reference fails after 50 iterations (golden of the failure)
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd)
REF=$ROOT/oracle/_ref/vpr_ref; W=$(mktemp -d); cd "$W"
cp "$ROOT/tests/fixtures/k6_N10_like.xml" .
python "$ROOT/tests/fixtures/gen_blif.py" toy.blif --luts 300 --pis 16 --window 60 --seed 1 --name toy
python "$ROOT/tests/fixtures/gen_blif.py" mid.blif --luts 4000 --pis 64 --window 400 --seed 2 --name mid
python "$ROOT/tests/fixtures/gen_blif.py" hub.blif --luts 900 --pis 24 --window 120 --seed 5 --name hub --hub 700
python "$ROOT/tests/fixtures/gen_blif.py" duo.blif --luts 500 --pis 20 --window 80 --seed 7 --name duo --clocks 2
for c in toy:64 mid:200 hub:90 duo:80; do
  n=${c%%:*}; w=${c##*:}
  "$REF" flow k6_N10_like.xml $n --nodisp --pack --place > /dev/null
  PF_DUMP_PROBLEM=${n}_w$w.pfp PF_DUMP_RESULT=${n}_w$w.pfr PF_DUMP_NAMES=${n}_w$w.pfn PF_DUMP_TGRAPH=${n}_w$w.pftg PF_DUMP_STA=${n}_w$w.pfsta "$REF" flow k6_N10_like.xml $n --nodisp --route --route_chan_width $w > /dev/null
  if [ $n = toy ] || [ $n = hub ]; then   # text-format goldens (include/pf_text.h): the reference's own print_route / print_place output
    xz -9 -c $n.route > "$HERE/${n}_w$w.route.xz"; xz -9 -c ${n}_w$w.pfn > "$HERE/${n}_w$w.pfn.xz"; cp $n.place "$HERE/"
  fi
  "$REF" inject ${n}_w$w.pfp --result ${n}_w${w}_nt.pfr > /dev/null
  if [ $n != mid ]; then   # breadth-first (Dijkstra) takes the reference 70 s on mid
    PF_DUMP_PROBLEM=${n}_w${w}_bf.pfp PF_DUMP_RESULT=${n}_w${w}_bf.pfr "$REF" flow k6_N10_like.xml $n --nodisp --route --route_chan_width $w --router_algorithm breadth_first > /dev/null
    for f in ${n}_w${w}_bf.pfp ${n}_w${w}_bf.pfr; do xz -9 -c $f > "$HERE/$f.xz"; done
  fi
  for f in ${n}_w$w.pfp ${n}_w$w.pfr ${n}_w${w}_nt.pfr ${n}_w$w.pftg ${n}_w$w.pfsta; do xz -9 -c $f > "$HERE/$f.xz"; done
done
cp toy.blif toy.place "$HERE/"; xz -9 -c toy.net > "$HERE/toy.net.xz"
# heterogeneous fabric
cp "$ROOT/tests/fixtures/k6_N10_het.xml" .
python "$ROOT/tests/fixtures/gen_blif.py" het.blif --luts 400 --pis 20 --window 80 --seed 9 --name het --mults 6
"$REF" flow k6_N10_het.xml het --nodisp --pack --place > /dev/null
PF_DUMP_PROBLEM=het_w60.pfp PF_DUMP_RESULT=het_w60.pfr "$REF" flow k6_N10_het.xml het --nodisp --route --route_chan_width 60 > /dev/null || true
PF_DUMP_PROBLEM=het_w70_bf.pfp PF_DUMP_RESULT=het_w70_bf.pfr "$REF" flow k6_N10_het.xml het --nodisp --route --route_chan_width 70 --router_algorithm breadth_first > /dev/null
PF_DUMP_PROBLEM=het_w70.pfp PF_DUMP_RESULT=het_w70.pfr PF_DUMP_NAMES=het_w70.pfn PF_DUMP_TGRAPH=het_w70.pftg PF_DUMP_STA=het_w70.pfsta "$REF" flow k6_N10_het.xml het --nodisp --route --route_chan_width 70 > /dev/null
"$REF" inject het_w70.pfp --result het_w70_nt.pfr > /dev/null
for f in het_w70.pfp het_w70.pfr het_w70_nt.pfr het_w70.pftg het_w70.pfsta het_w70.pfn het_w70_bf.pfp het_w70_bf.pfr het_w60.pfp het_w60.pfr; do xz -9 -c $f > "$HERE/$f.xz"; done
xz -9 -c het.route > "$HERE/het_w70.route.xz"; xz -9 -c het.net > "$HERE/het.net.xz"; cp het.place het.blif "$HERE/"
# heq: the het circuit with the multiplier's input ports declared logically equivalent and one signal wired to two pins of each
# port, i.e. nets that connect TWICE to one SINK rr node (capacity 4).  The reference routes them, then its own DEBUG cross-check
# timing_driven_check_net_delays (route_timing.c:246) aborts the run on exactly these nets, so the result is dumped the moment
# the routing is legal (PF_DUMP_AT_SUCCESS=1).
python - <<'PY'
import re
s = open("k6_N10_het.xml").read()
top = s.index('<pb_type name="mult" height="2"')          # the cluster-level ports only
s = s[:top] + s[top:].replace('<input name="a" num_pins="4"/>', '<input name="a" num_pins="4" equivalent="true"/>', 1) \
                      .replace('<input name="b" num_pins="4"/>', '<input name="b" num_pins="4" equivalent="true"/>', 1)
open("heq.xml", "w").write(s)
out = []
for line in open("het.blif").read().replace(".model het", ".model heq").split("\n"):
    if line.startswith(".subckt mult4"):
        line = re.sub(r"a\[1\]=\S+", "a[1]=" + re.search(r"a\[0\]=(\S+)", line).group(1), line)
        line = re.sub(r"b\[3\]=\S+", "b[3]=" + re.search(r"b\[2\]=(\S+)", line).group(1), line)
    out.append(line)
open("heq.blif", "w").write("\n".join(out))
PY
"$REF" flow heq.xml heq --nodisp --pack --place > /dev/null
PF_DUMP_AT_SUCCESS=1 PF_DUMP_PROBLEM=heq_w70.pfp PF_DUMP_RESULT=heq_w70.pfr PF_DUMP_TGRAPH=heq_w70.pftg "$REF" flow heq.xml heq --nodisp --route --route_chan_width 70 > /dev/null || true
for f in heq_w70.pfp heq_w70.pfr heq_w70.pftg; do xz -9 -c $f > "$HERE/$f.xz"; done
# two wire types
cp "$ROOT/tests/fixtures/k6_N10_mix.xml" .
python "$ROOT/tests/fixtures/gen_blif.py" mix.blif --luts 350 --pis 18 --window 70 --seed 21 --name mix
"$REF" flow k6_N10_mix.xml mix --nodisp --pack --place > /dev/null
for w in 60 70; do PF_DUMP_PROBLEM=mix_w$w.pfp PF_DUMP_RESULT=mix_w$w.pfr PF_DUMP_TGRAPH=mix_w$w.pftg "$REF" flow k6_N10_mix.xml mix --nodisp --route --route_chan_width $w > /dev/null; done
for f in mix_w70.pfp mix_w70.pfr mix_w70.pftg mix_w60.pfp mix_w60.pfr; do xz -9 -c $f > "$HERE/$f.xz"; done
# pass-transistor wire switches: the switch table of toy_w64 edited (tests/test_oracle_golden.py::unbuffered_toy), routed by the
# reference router in inject mode, timing off and timing-driven with the criticalities of toy_w64.pfr replayed
(cd "$ROOT" && python - "$W" <<'PY'
import sys
sys.path.insert(0, "tests")
from parallel_eda_b200 import pfio
from test_oracle_golden import unbuffered_toy
for timing, tag in ((False, "nt"), (True, "td")):
    pfio.write_problem("%s/toy_w64_unbuf_%s.pfp" % (sys.argv[1], tag), unbuffered_toy("tests/golden/toy_w64.pfp.xz", timing))
PY
)
"$REF" inject toy_w64_unbuf_nt.pfp --result toy_w64_unbuf_nt.pfr > /dev/null
"$REF" inject toy_w64_unbuf_td.pfp --crit toy_w64.pfr --result toy_w64_unbuf_td.pfr > /dev/null
for f in toy_w64_unbuf_nt.pfr toy_w64_unbuf_td.pfr; do xz -9 -c $f > "$HERE/$f.xz"; done
# packed-netlist goldens (pf_net_read, tests/test_net_reader.py): block[] / clb_net[] as the reference's read_netlist held them in a routing run
cp "$ROOT/tests/fixtures/k6_N10_like.xml" "$ROOT/tests/fixtures/k6_N10_het.xml" .
for c in toy:k6_N10_like.xml:64 het:k6_N10_het.xml:70 duo:k6_N10_like.xml:80 mid:k6_N10_like.xml:200; do
  n=${c%%:*}; r=${c#*:}; a=${r%%:*}; w=${r##*:}
  PF_DUMP_NETLIST=$n.netlist "$REF" flow $a $n --nodisp --route --route_chan_width $w > /dev/null
  xz -9e -c $n.netlist > "$HERE/$n.netlist.xz"
done
xz -9e -c duo.net > "$HERE/duo.net.xz"
# the analysis of the finished routing (routing_stats -> do_timing_analysis(.., is_final_analysis = TRUE)): pf_sta_analyze_final
for c in toy:k6_N10_like.xml:64 het:k6_N10_het.xml:70 duo:k6_N10_like.xml:80; do
  n=${c%%:*}; r=${c#*:}; a=${r%%:*}; w=${r##*:}
  PF_DUMP_STA_FINAL=${n}_w${w}_final.pfsta "$REF" flow $a $n --nodisp --route --route_chan_width $w > /dev/null
  for f in ${n}_w${w}_final.pfsta ${n}_w${w}_final.pfsta.slack; do xz -9e -c $f > "$HERE/$f.xz"; done
done
# clock-to-flipflop override constraints: the duo circuit routed with tests/golden/duo_ovr.sdc (same routing problem as duo_w80.pfp)
cp "$HERE/duo_ovr.sdc" .
PF_DUMP_TGRAPH=duo_ovr_w80.pftg PF_DUMP_STA=duo_ovr_w80.pfsta PF_DUMP_STA_FINAL=duo_ovr_w80_final.pfsta PF_DUMP_RESULT=duo_ovr_w80.pfr "$REF" flow k6_N10_like.xml duo --nodisp --route --route_chan_width 80 --sdc_file duo_ovr.sdc > /dev/null
for f in duo_ovr_w80.pftg duo_ovr_w80.pfsta duo_ovr_w80.pfr duo_ovr_w80_final.pfsta duo_ovr_w80_final.pfsta.slack; do xz -9e -c $f > "$HERE/$f.xz"; done
echo "goldens written to $HERE"
