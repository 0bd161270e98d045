python -m pytest tests/test_gpu_parity.py -q -k "device_built or native_library" > gpurun_out/r02d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02d_pytest.txt
python tools/mgpu_phases.py > gpurun_out/r02d_phases_n1.out 2> gpurun_out/r02d_phases_n1.txt
sh tools/ab_multi.sh 2 noval > gpurun_out/r02d_ab_noval.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02d_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02d_bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/r02d_pytest.txt; grep PF_PHASES gpurun_out/r02d_phases_n1.txt; cat gpurun_out/r02d_ab_noval.txt; cut -c1-2500 gpurun_out/r02d_bench.json; tail -n 3 gpurun_out/r02d_bench.err
