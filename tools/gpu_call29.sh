# median-of-3 parity tests, cfg 4 with the loose-fabric divisor 2, phases of the end-to-end call (verbose library output)
python -m pytest tests -m gpu -q > gpurun_out/r02A_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02A_pytest.txt
tail -n 4 gpurun_out/r02A_pytest.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02A_bench.json 2> gpurun_out/r02A_bench.err; cut -c1-200 gpurun_out/r02A_bench.json; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r02A_bench.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['route']['iterations'], d['route']['wirelength'][:3], d['roofline']['frac'], d['e2e']['s_per_step'], d['e2e']['phases_s'])"
python tools/e2e_phases.py 400 200000 4 gen > gpurun_out/r02A_e2e_phases.txt 2>&1; tail -n 12 gpurun_out/r02A_e2e_phases.txt
