#!/bin/sh
# A/B of experiment builds of the device code on a B200 (one gpurun call answers "did it help?").
#   1. here (no GPU; nvcc cross-compiles):   sh tools/ab_bench.sh build deunroll "-DPF_DEUNROLL_COLD=1"
#   2. on the box:  gpurun --timeout 900 -- 'sh tools/ab_bench.sh run deunroll > gpurun_out/ab_deunroll.txt 2>&1'
# `run` alternates product / variant three times (bench.py --steps 10 --warmup 3 each) and prints ms_per_step, the
# kernel time per step and the iteration counts of every run, so drift between runs is visible next to the difference.
set -e
cd "$(dirname "$0")/.."
case "$1" in
build)
  PF_LIB_VARIANT="$2" PF_EXTRA_NVCC_FLAGS="$3" python -m parallel_eda_b200.build --force | tail -1
  ;;
run)
  lib="$PWD/parallel_eda_b200/libpf_router_$2.so"
  [ -f "$lib" ] || { echo "$lib not built"; exit 2; }
  for rep in 1 2 3; do
    for which in product "$2"; do
      if [ "$which" = product ]; then unset PF_ROUTER_LIB; else export PF_ROUTER_LIB="$lib"; fi
      python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%-10s rep $rep: %.2f ms/step, kernel %.2f ms, iterations %s, e2e %.1f ms, clocks %s' % ('$which', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'], 1e3*d['e2e']['s_per_step'], d['clocks']['sm_mhz']))
"
    done
  done
  ;;
*) echo "usage: $0 build <variant> <nvcc flags> | run <variant>"; exit 2;;
esac
