run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%.2f ms/step, kernel %.2f ms, iterations %s wl %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'][:4], d['route']['wirelength'][:2]))
"; }
for d in 0 4 2 1; do echo "inflight_div $d:"; run --inflight-div $d; done
PF_PHASES=1 python tools/mgpu_phases.py 800 800000 2>&1 >/dev/null | grep -a "PF_PHASES" | cut -c1-250
