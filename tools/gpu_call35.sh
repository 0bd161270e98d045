# generator fill pass on a side stream under the rest of create: graph identity on the GPU, phases of the end-to-end call, bench line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "generated or configs4 or device_built" 2>&1 | tail -n 3
python tools/e2e_phases.py 400 200000 4 gen > gpurun_out/r02H_e2e_phases.txt 2>&1; grep -a "^call\|create " gpurun_out/r02H_e2e_phases.txt | tail -n 6
python bench.py --no-cpu-baseline > gpurun_out/r02H_bench.json 2> gpurun_out/r02H_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02H_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['route']['iterations'][:3], d['roofline']['frac'], d['e2e']['value'], d['e2e']['s_per_step'], d['e2e']['phases_s'], d['e2e']['result_check']['device_check_route']['ok'])"; tail -n 3 gpurun_out/r02H_bench.err
