#!/bin/sh
# One GPU call that refreshes the round's single-GPU evidence: parity tests, the bench line, the ncu launch list of the bench
# command and one `--set full` capture of the iteration-1 launch of the route kernel (read back here with tools/ncu_traffic.py).
#   gpurun --timeout 1500 -- 'sh tools/gpu_round.sh r02a'
tag="${1:-r02}"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pf_route_kernel -c 1 -o gpurun_out/${tag}_route -f python tools/prof_run.py 400 200000 1 > gpurun_out/${tag}_prof_run.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest.txt; cat gpurun_out/${tag}_bench.json; tail -n 3 gpurun_out/${tag}_prof_run.log
