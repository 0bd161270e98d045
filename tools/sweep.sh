run() { python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -n 1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ms/step %.1f nets/s %.0f iters %s kernel_ms %.1f frac %.4f'%(d['ms_per_step'],d['value'],d['route']['iterations'],d['roofline']['kernel_ms_per_step'],d['roofline']['frac']))"; }
echo "== 400 auto"; run
echo "== 400 batch1 slack0"; run --max-batch 1 --pop-slack 0
echo "== 400 batch1 slack0 div64"; run --max-batch 1 --pop-slack 0 --inflight-div 64
echo "== 400 batch1 slack0 slots 1184"; run --max-batch 1 --pop-slack 0 --slots 1184
echo "== 100 auto"; run --grid 100 --nets 12500
echo "== 100 b1s0"; run --grid 100 --nets 12500 --max-batch 1 --pop-slack 0
echo "== 100 b32 s.25"; run --grid 100 --nets 12500 --max-batch 32 --pop-slack 0.25
echo "== 200 auto"; run --grid 200 --nets 50000
echo "== 200 b1s0"; run --grid 200 --nets 50000 --max-batch 1 --pop-slack 0
