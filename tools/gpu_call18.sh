run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%.2f ms/step, kernel %.2f ms, iterations %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'][:3]))
"; }
echo "product:"; run
echo "product far_cap=8192:"; run --cfg "far_cap=8192"
export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_nobk.so"; echo "nobk:"; run; echo "nobk far_cap=8192:"; run --cfg "far_cap=8192"
