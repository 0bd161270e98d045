td() { python tools/td_iter_profile.py "$@" 2>&1 >/dev/null | grep -E "^(bgm|sv0)|Error|error" | cut -c1-250; }
td bgm_w260
td bgm_w260 big_slots=592
td bgm_w260 label2_log2=15
td bgm_w260 label2_log2=15 far_cap=32768
td sv0_w220
td sv0_w220 label2_log2=15
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%.2f ms/step, kernel %.2f ms, e2e %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d.get('e2e')))
"
