run() { python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.1f nets/s %.0f iters %s kernel_ms %.1f frac %.4f'%(d['ms_per_step'],d['value'],d['route']['iterations'],d['roofline']['kernel_ms_per_step'],d['roofline']['frac']))"; }
echo "== 400 default"; run
echo "== 400 div16"; run --inflight-div 16
echo "== 400 div8"; run --inflight-div 8
echo "== 400 div4"; run --inflight-div 4
echo "== 400 batch2"; run --max-batch 2 --pop-slack 0
