# re-entry check of HEAD on a fresh box: smoke, every GPU test (with durations), the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02C_smoke.txt 2>&1; tail -n 2 gpurun_out/r02C_smoke.txt
python -m pytest tests -m gpu -q --durations=15 > gpurun_out/r02C_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02C_pytest.txt
tail -n 25 gpurun_out/r02C_pytest.txt
python bench.py > gpurun_out/r02C_bench.json 2> gpurun_out/r02C_bench.err; cut -c1-1500 gpurun_out/r02C_bench.json; tail -n 3 gpurun_out/r02C_bench.err
