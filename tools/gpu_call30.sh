# generator with multiply-shift divisions + pinned result arrays: graph hash parity on the GPU, phases of the end-to-end call, bench line
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -n 3
python tools/e2e_phases.py 400 200000 5 gen > gpurun_out/r02B_e2e_phases.txt 2>&1; grep -a "call\|create\|result" gpurun_out/r02B_e2e_phases.txt | tail -n 12
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02B_bench.json 2> gpurun_out/r02B_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02B_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['route']['iterations'], d['roofline']['frac'], d['e2e']['s_per_step'], d['e2e']['phases_s'], d['e2e']['result_check'])"
