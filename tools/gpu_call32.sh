# e2e with the previous result released before the next call: phases of the end-to-end call (verbose), bench line
mkdir -p gpurun_out
python tools/e2e_phases.py 400 200000 4 gen > gpurun_out/r02D_e2e_phases.txt 2>&1; grep -a "^call\|create \|result " gpurun_out/r02D_e2e_phases.txt | tail -n 14
python bench.py --no-cpu-baseline > gpurun_out/r02D_bench.json 2> gpurun_out/r02D_bench.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02D_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['route']['iterations'][:3], d['roofline']['frac'], d['e2e']['value'], d['e2e']['s_per_step'], d['e2e']['phases_s'], d['e2e']['result_check']['device_check_route']['ok'])"; tail -n 3 gpurun_out/r02D_bench.err
