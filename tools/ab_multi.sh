#!/bin/sh
# A/B/C... of experiment builds on one B200: alternates the product library and each named variant `reps` times
#   gpurun -- 'sh tools/ab_multi.sh 2 deunroll fr64 > gpurun_out/ab.txt 2>&1'
cd "$(dirname "$0")/.."
reps="$1"; shift
for rep in $(seq 1 "$reps"); do
  for which in product "$@"; do
    if [ "$which" = product ]; then unset PF_ROUTER_LIB; else export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_$which.so"; fi
    python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%-10s rep $rep: %.2f ms/step, kernel %.2f ms, iterations %s, frac %.4f clocks %s' % ('$which', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'], d['roofline']['frac'], d['clocks']['sm_mhz']))
"
  done
done
