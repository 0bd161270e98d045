#!/bin/sh
# Size-matched stand-ins for BASELINE.json configs[1] / configs[2] (stereovision0 ~11 k LUTs, bgm ~32 k LUTs; the VTR files are
# neither in the reference nor on the box): generated netlists taken through the UNMODIFIED reference's own
# pack -> place -> timing-driven route on the heterogeneous fixture architecture (columns of height-2 hard multipliers).
# Only works where /root/reference exists (oracle/_ref/vpr_ref built by `make -C oracle/ref_build`).  ~6 min of CPU.
set -e
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../../.." && pwd); REF=$ROOT/oracle/_ref/vpr_ref
W=$(mktemp -d); cd "$W"; cp "$ROOT/tests/fixtures/k6_N10_het.xml" .
python "$ROOT/tests/fixtures/gen_blif.py" sv0.blif --luts 11000 --pis 128 --window 500 --seed 31 --name sv0 --mults 40
python "$ROOT/tests/fixtures/gen_blif.py" bgm.blif --luts 32000 --pis 256 --window 800 --seed 37 --name bgm --mults 100
for c in sv0:220 bgm:260; do
  n=${c%%:*}; w=${c##*:}
  "$REF" flow k6_N10_het.xml $n --nodisp --pack --place > ${n}_pp.log
  PF_DUMP_PROBLEM=${n}_w$w.pfp PF_DUMP_RESULT=${n}_w$w.pfr PF_DUMP_TGRAPH=${n}_w$w.pftg "$REF" flow k6_N10_het.xml $n --nodisp --route --route_chan_width $w > ${n}_route$w.log
  for f in ${n}_w$w.pfp ${n}_w$w.pftg $n.net $n.blif; do xz -9e -T0 -c $f > "$HERE/$f.xz"; done
  cp $n.place "$HERE/"
done
echo "fixtures in $HERE; summaries (*.json) are written from the .pfr / route logs by the snippet in DESIGN.md §5"
