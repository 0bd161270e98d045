#!/bin/sh
# bench.py (no e2e / cpu legs) under several pf_config variants, one line each:  sh tools/gpu_variants.sh "" "validate_commits=-1" ...
for v in "$@"; do
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --cfg "$v" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('cfg[%s]: %.2f ms/step, kernel %.2f ms, iterations %s, wl %s, frac %.4f, over %s' % ('$v', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'], d['route']['wirelength'][-1], d['roofline']['frac'], d['route']['overused_per_iteration']))
"
done
