run() { python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('%.2f ms/step, kernel %.2f ms, iterations %s' % (d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['route']['iterations'][:3]))
"; }
echo "product:"; run
( export PF_ROUTER_LIB="$PWD/parallel_eda_b200/libpf_router_head.so"; echo "head:"; run )
python tools/mgpu_phases.py 2>&1 >/dev/null | grep -E "PF_PHASES rank 0 (it|summary)" | cut -c1-230 > gpurun_out/r02o_phases_n1.txt
python tools/td_iter_profile.py bgm_w260 > gpurun_out/r02o_td_bgm.out 2> gpurun_out/r02o_td_bgm.txt
python tools/td_iter_profile.py sv0_w220 2>&1 >/dev/null | grep -E "^sv0" > gpurun_out/r02o_td_sv0.txt
cat gpurun_out/r02o_phases_n1.txt; grep -E "^bgm|iteration (1|9|13|16|20):|Error|error" gpurun_out/r02o_td_bgm.txt | cut -c1-220; tail -n 3 gpurun_out/r02o_td_bgm.out; grep -c moved gpurun_out/r02o_td_bgm.txt; cat gpurun_out/r02o_td_sv0.txt
echo "product again:"; run
