# last call of the round: every GPU test on HEAD (the device STA tests first: final analysis and override constraints are new)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sta.py tests -m gpu -q -p no:cacheprovider > gpurun_out/r02J_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/r02J_pytest.txt
tail -n 6 gpurun_out/r02J_pytest.txt
